"""CPU suite: oracle (our restatement) vs golden dumps of the compiled reference; host logic; C-ABI symbol check."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import helpers as H
import pipeline_common as PC
from ma_b200 import index, synth

PRESETS = ["illumina", "default", "pacbio", "nanopore", "illuminapaired"]


@pytest.mark.parametrize("preset", PRESETS)
def test_oracle_pipeline_matches_reference_golden(preset, tmp_path):
    gold = PC.load_gold(preset)
    o = H.oracle_align_dump(PC.GOLD_PREFIX, PC.gold_reads(preset), preset, str(tmp_path / "o.dump"), PC.SRAND, 5)
    for k in gold:
        assert np.array_equal(o[k], gold[k]), k


@pytest.mark.parametrize("preset", ["illumina", "illuminapaired", "pacbio"])
def test_oracle_pipeline_matches_reference_golden_with_large_genome_heuristics(preset, tmp_path):
    """ "Minimum Genome Size for Heuristics" = 0: seeding drop-off and SoC minimal length active (gold_<preset>_heur.npz)."""
    gold = PC.load_gold(preset + "_heur")
    o = H.oracle_align_dump(PC.GOLD_PREFIX, PC.gold_reads(preset), preset, str(tmp_path / "o.dump"), PC.SRAND, 5,
                            min_genome_size=0)
    for k in gold:
        assert np.array_equal(o[k], gold[k]), k


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("preset,params", PC.PARAM_VARIATIONS, ids=lambda v: v if isinstance(v, str) else "-".join(v))
def test_oracle_matches_live_reference_parameter_variations(preset, params, tmp_path):
    """Parameters of the presetting set on the reference the way its CLI does (byName()->setByText(), MA_REF_SET in
    oracle/ref_dump.cpp) and on the oracle: every dumped stage identical."""
    env = dict(os.environ)
    env.update(H.ref_param_env(params))
    out = str(tmp_path / "r.dump")
    subprocess.check_call([H.REF_DUMP, "align", PC.GOLD_PREFIX, PC.gold_reads(preset), preset, out, str(PC.SRAND)], env=env)
    ref = H.load_dump(out)
    o = H.oracle_align_dump(PC.GOLD_PREFIX, PC.gold_reads(preset), preset, str(tmp_path / "o.dump"), PC.SRAND, 5,
                            params=params)
    base = PC.load_gold(preset)
    assert any(not np.array_equal(ref[k], base[k]) for k in ("seg", "seed", "soc", "harmseed", "aln", "mq", "pr")), \
        "the variation does not change anything"
    for k in ref:
        assert np.array_equal(o[k], ref[k]), k


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_oracle_matches_live_reference_multi_contig(tmp_path):
    g = synth.random_genome([120_000, 40_000], 11)
    synth.write_genome_txt(str(tmp_path / "g.txt"), g)
    H.run_ref("index", tmp_path / "g.txt", tmp_path / "g")
    reads, _, _, _ = synth.simulate_reads(g, 600, 150, 12, sub_rate=0.015, ins_rate=0.003, del_rate=0.003)
    synth.write_reads_txt(str(tmp_path / "r.txt"), reads)
    H.run_ref("align", tmp_path / "g", tmp_path / "r.txt", "illumina", tmp_path / "r.dump", 5)
    r = H.load_dump(str(tmp_path / "r.dump"))
    o = H.oracle_align_dump(str(tmp_path / "g"), str(tmp_path / "r.txt"), "illumina", str(tmp_path / "o.dump"), 5, 5)
    for k in r:
        assert np.array_equal(o[k], r[k]), k


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("preset,long_reads", PC.REPEAT_RUNS)
def test_oracle_matches_live_reference_repeat_rich(preset, long_reads, tmp_path):
    """Reads with hundreds of ambiguous seeds (tandem / interspersed / low-complexity repeats): tie order of the
    reference's std::sort and heap operations, overlapping SoC windows, ambiguity limits."""
    g = PC.repeat_rich_genome()
    synth.write_genome_txt(str(tmp_path / "g.txt"), g)
    H.run_ref("index", tmp_path / "g.txt", tmp_path / "g")
    synth.write_reads_txt(str(tmp_path / "r.txt"), PC.repeat_rich_reads(g, long_reads))
    H.run_ref("align", tmp_path / "g", tmp_path / "r.txt", preset, tmp_path / "r.dump", PC.SRAND)
    r = H.load_dump(str(tmp_path / "r.dump"))
    assert int(np.diff(r["seed_off"]).max()) > 100
    o = H.oracle_align_dump(str(tmp_path / "g"), str(tmp_path / "r.txt"), preset, str(tmp_path / "o.dump"), PC.SRAND, 5)
    for k in r:
        assert np.array_equal(o[k], r[k]), k


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("preset", ["illumina", "default", "nanopore"])
def test_oracle_matches_live_reference_ragged_read_lengths(preset, tmp_path):
    """Read lengths from 17 to 2000 in one batch (PC.RAGGED_LENGTHS)."""
    reads = PC.ragged_reads(index.load_index(PC.GOLD_PREFIX).forward_codes())
    PC.write_ragged_txt(str(tmp_path / "r.txt"), reads)
    H.run_ref("align", PC.GOLD_PREFIX, tmp_path / "r.txt", preset, tmp_path / "r.dump", PC.SRAND)
    r = H.load_dump(str(tmp_path / "r.dump"))
    o = H.oracle_align_dump(PC.GOLD_PREFIX, str(tmp_path / "r.txt"), preset, str(tmp_path / "o.dump"), PC.SRAND, 5)
    for k in r:
        assert np.array_equal(o[k], r[k]), k


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_oracle_matches_live_reference_above_heuristics_threshold(tmp_path):
    """A 12 Mbp genome (24 M-base text > "Minimum Genome Size for Heuristics"): the large-genome heuristics run
    without any parameter override, as on BASELINE's configurations; paired Illumina reads, every stage."""
    g = synth.random_genome([6_000_000] * 2, 7)
    synth.write_genome_txt(str(tmp_path / "g.txt"), g)
    H.run_ref("index", tmp_path / "g.txt", tmp_path / "g")
    m1, m2, *_ = synth.simulate_pairs(g, 1500, 150, 2017)
    reads = np.empty((3000, 150), dtype=np.uint8)
    reads[0::2], reads[1::2] = m1, m2
    synth.write_reads_txt(str(tmp_path / "r.txt"), reads)
    H.run_ref("align", tmp_path / "g", tmp_path / "r.txt", "illuminapaired", tmp_path / "r.dump", PC.SRAND)
    r = H.load_dump(str(tmp_path / "r.dump"))
    o = H.oracle_align_dump(str(tmp_path / "g"), str(tmp_path / "r.txt"), "illuminapaired", str(tmp_path / "o.dump"),
                            PC.SRAND, 5)
    for k in r:
        assert np.array_equal(o[k], r[k]), k


def test_hostsim_device_routines_match_oracle():
    """The MA_HD routines that the kernels wrap (seeding, SoC/harmonization, NW glue, exact std::sort/heap), compiled
    for the host, against the oracle."""
    d = os.path.join(H.ROOT, "tests", "hostsim")
    subprocess.check_call(["make", "-s", "-C", d])
    assert subprocess.call([os.path.join(d, "test_stl_exact")]) == 0
    for preset in PRESETS:
        rc = subprocess.call([os.path.join(d, "hostsim"), PC.GOLD_PREFIX, PC.gold_reads(preset), preset,
                              str(PC.SRAND)])
        assert rc == 0, preset


def test_index_file_roundtrip(tmp_path):
    ix = index.load_index(PC.GOLD_PREFIX)
    assert ix.ref_len == 120_000 and ix.fwd_len == 60_000 and len(ix.contig_start) == 3
    assert np.array_equal(index.pack_forward(ix.forward_codes()), ix.pac)
    index.store_index(ix, str(tmp_path / "x"))
    for ext in (".bwt", ".sa", ".pac"):
        assert open(PC.GOLD_PREFIX + ext, "rb").read() == open(str(tmp_path / "x") + ext, "rb").read(), ext
    ix2 = index.load_index(str(tmp_path / "x"))
    assert ix2.contig_names == ix.contig_names and np.array_equal(ix2.contig_len, ix.contig_len)


def test_c_abi_exports_every_declared_symbol():
    lib_path = os.path.join(H.ROOT, "ma_b200", "libma_b200.so")
    if not os.path.exists(lib_path):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(lib_path)
    hdr = open(os.path.join(H.ROOT, "include", "ma_b200.h")).read()
    names = set(re.findall(r"\b(ma_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    for n in sorted(names):
        assert hasattr(lib, n), n


def test_no_cpu_fallback_without_device():
    from ma_b200 import api
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.MaB200Error):
        api.Context(0)


def test_synth_generators_are_deterministic():
    g1 = synth.random_genome([5000], 3)
    g2 = synth.random_genome([5000], 3)
    assert np.array_equal(g1[0], g2[0])
    r1 = synth.simulate_reads(g1, 50, 150, 4)[0]
    r2 = synth.simulate_reads(g2, 50, 150, 4)[0]
    assert np.array_equal(r1, r2) and r1.shape == (50, 150)
    a, b, *_ = synth.simulate_pairs(g1, 20, 150, 5)
    assert a.shape == b.shape == (20, 150)


def test_warp_emulated_dp_kernels_match_oracle(tmp_path):
    """The source of the warp-synchronous DP kernels (ksw_bx.cuh: packed banded exact mode, ksw_bn.cuh: its register-resident
    form for narrow bands, ksw_qs.cuh: query-stationary register mode) runs unchanged on a lock-step warp emulator (tests/hostsim/warp_emu.h) against the oracle: every
    kswcpp_extz_t field and the CIGAR for random problems (bands 0..900, N bases, z-drops, score sets whose int8
    arithmetic wraps in the reference) and for the reference's own golden DP calls of two such score sets (the GPU
    suite runs all of them, tests/test_ksw_gpu.py)."""
    d = os.path.join(H.ROOT, "tests", "hostsim")
    subprocess.check_call(["make", "-s", "-C", d])
    assert subprocess.call([os.path.join(d, "bx_sim"), "500", "21", "160"]) == 0
    assert subprocess.call([os.path.join(d, "bx_sim"), "30", "22", "1000"]) == 0
    assert subprocess.call([os.path.join(d, "bx_sim"), "4", "71", "20000"]) == 0  # int32 H rows (lengths beyond int16)
    narrow = dict(os.environ, BX_NARROW="1")  # ksw_bn.cuh (register-resident narrow bands) where it applies
    assert subprocess.call([os.path.join(d, "bx_sim"), "500", "23", "200"], env=narrow) == 0
    assert subprocess.call([os.path.join(d, "bx_sim"), "40", "24", "1500"], env=narrow) == 0
    assert subprocess.call([os.path.join(d, "qs_sim"), "300", "5"]) == 0
    for name in ("_swapped", "_large"):  # (a sample of each set: the emulator runs ~10 problems a second)
        g = np.load(os.path.join(H.GOLDEN, "ksw_golden%s.npz" % name))
        dd = {"ksw_calls": g["calls"].astype(np.int64), "ksw_seq": g["seq"], "ksw_cigar": g["cigar"]}
        sc = [int(x) for x in g["score"]] if "score" in g else [2, 4, 4, 2, 24, 1]
        path = str(tmp_path / ("calls%s.txt" % name))
        n = 0
        with open(path, "w") as f:
            for fl, q, t, c in H.split_ksw_dump(dd):
                if len(q) and len(t) and n < 90:
                    f.write("%d %d %d %s %s\n" % (fl["w"], fl["zdrop"], fl["flag"], "".join(map(str, q)), "".join(map(str, t))))
                    n += 1
        assert n == 90
        assert subprocess.call([os.path.join(d, "bx_sim"), "file", path] + [str(x) for x in sc]) == 0, name
