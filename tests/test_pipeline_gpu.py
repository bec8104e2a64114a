"""GPU parity of the whole path through the C ABI: segments, located seeds, harmonized seed sets and alignment
records must equal the reference's (golden dumps of the compiled reference, and the oracle at larger sizes).

Everything is compared bit-exactly, including the records that depend on Harmonization's double-precision libm calls
(atan/tan/sin/log): a last-ulp difference between CUDA's libm and glibc that changed an emitted alignment would fail
these tests (none has been observed on the golden sets, 10 000 + 100 000 random reads)."""
import os

import numpy as np
import pytest

import helpers as H
import pipeline_common as PC
from ma_b200 import api, index, synth

pytestmark = pytest.mark.gpu
PRESETS = ["illumina", "default", "pacbio", "nanopore"]


def make_ctx(preset, srand=PC.SRAND):
    ctx = api.Context(0, preset)
    p = api.preset(preset)
    p.srand_base = srand
    ctx.set_params(p)
    return ctx


@pytest.fixture(scope="module")
def gold_index():
    return index.load_index(PC.GOLD_PREFIX)


@pytest.mark.parametrize("preset", PRESETS)
def test_all_stages_match_reference_golden(preset, gold_index):
    ctx = make_ctx(preset)
    ctx.index_upload(gold_index)
    got = PC.gpu_stage_dump(ctx, PC.read_reads_txt(PC.gold_reads(preset)))
    PC.assert_same_stages(got, PC.load_gold(preset), what=preset)
    ctx.close()


@pytest.mark.parametrize("preset", PRESETS + ["illuminapaired"])
def test_mapping_quality_and_pairing_match_reference_golden(preset, gold_index):
    """MappingQuality per read and PairedReads per pair of consecutive reads, against the compiled
    reference (flags, double mapping quality bit for bit, order of the returned vectors)."""
    ctx = api.Context(0, preset)
    p = api.preset(preset)
    p.srand_base = PC.SRAND
    ctx.set_params(p)
    ctx.index_upload(gold_index)
    assert p.use_paired_reads == (1 if preset == "illuminapaired" else 0)
    got = PC.gpu_mapq_dump(ctx, PC.read_reads_txt(PC.gold_reads(preset)), p)
    gold = PC.load_gold(preset)
    for k in ("mq_off", "mq", "pr_off", "pr"):
        assert np.array_equal(got[k], gold[k]), (preset, k)
    ctx.close()


@pytest.mark.parametrize("preset", ["illumina", "illuminapaired", "pacbio"])
def test_large_genome_heuristics_match_reference_golden(preset, gold_index):
    """The heuristics that only run for genomes above "Minimum Genome Size for Heuristics" (10 M; seeding drop-off,
    SoC minimal length) — i.e. on BASELINE's 100 Mbp configuration — switched on for the golden genome by setting that
    parameter to 0 (tests/golden/gold_<preset>_heur.npz, written by the compiled reference): every stage, mapping
    quality and pairing."""
    ctx = api.Context(0, preset)
    p = api.preset(preset)
    p.srand_base = PC.SRAND
    p.genome_size_disable = 0
    ctx.set_params(p)
    ctx.index_upload(gold_index)
    gold = PC.load_gold(preset + "_heur")
    base = PC.load_gold(preset)
    assert len(gold["seg"]) < len(base["seg"]) or len(gold["soc"]) < len(base["soc"])  # the heuristics do bite
    reads = PC.read_reads_txt(PC.gold_reads(preset))
    got = PC.gpu_stage_dump(ctx, reads)
    PC.assert_same_stages(got, gold, what=preset + " heuristics")
    mq = PC.gpu_mapq_dump(ctx, reads, p)
    for k in ("mq_off", "mq", "pr_off", "pr"):
        assert np.array_equal(mq[k], gold[k]), (preset, k)
    ctx.close()


@pytest.mark.parametrize("preset,params", PC.PARAM_VARIATIONS, ids=lambda v: v if isinstance(v, str) else "-".join(v))
def test_parameter_variations_against_oracle(preset, params, gold_index, tmp_path):
    """Non-preset values of the presetting's parameters (seeding, SoC, harmonization heuristics, DP, reporting, pairing):
    every stage, mapping quality and pairing against the oracle, which test_pipeline_cpu.py pins to the live reference
    for the same variations."""
    ctx = api.Context(0, preset)
    p = api.preset(preset)
    p.srand_base = PC.SRAND
    for k, v in params.items():
        assert hasattr(p, k), k
        setattr(p, k, v)
    ctx.set_params(p)
    ctx.index_upload(gold_index)
    exp = H.oracle_align_dump(PC.GOLD_PREFIX, PC.gold_reads(preset), preset, str(tmp_path / "o.dump"), PC.SRAND, 5,
                              params=params)
    reads = PC.read_reads_txt(PC.gold_reads(preset))
    got = PC.gpu_stage_dump(ctx, reads)
    PC.assert_same_stages(got, exp, what="%s %s" % (preset, params))
    mq = PC.gpu_mapq_dump(ctx, reads, p)
    for k in ("mq_off", "mq", "pr_off", "pr"):
        assert np.array_equal(mq[k], exp[k]), (preset, params, k)
    ctx.close()


def test_gpu_index_build_is_bit_identical(gold_index):
    ctx = make_ctx("illumina")
    ctx.index_build(gold_index.forward_codes(), gold_index.contig_start, gold_index.contig_len)
    ix = ctx.index_download(gold_index.contig_names)
    assert ix.primary == gold_index.primary
    assert np.array_equal(ix.L2, gold_index.L2)
    assert np.array_equal(ix.bwt, gold_index.bwt)
    assert np.array_equal(ix.sa, gold_index.sa)
    assert np.array_equal(ix.pac[:len(gold_index.pac)], gold_index.pac)
    # and the pipeline on the GPU-built index gives the golden alignments
    got = PC.gpu_stage_dump(ctx, PC.read_reads_txt(PC.gold_reads("illumina")))
    PC.assert_same_stages(got, PC.load_gold("illumina"), what="gpu-built index")
    ctx.close()


def _index_arrays_equal(a, b):
    return (a.primary == b.primary and np.array_equal(a.L2, b.L2) and np.array_equal(a.bwt, b.bwt) and
            np.array_equal(a.sa, b.sa) and np.array_equal(a.pac, b.pac))


@pytest.mark.parametrize("chunk", [1 << 29, 40_000, 1_000])
def test_bucketed_index_builder_is_bit_identical(gold_index, chunk, monkeypatch):
    """The builder for human-sized genomes (index_build.cuh build_index_gpu_large: suffixes sorted bucket by bucket,
    ties refined 29 bases at a time) forced onto small genomes: one chunk, a few chunks, one chunk per 10-mer bin.
    Must equal the reference's index files (golden genome) and the prefix-doubling builder (repeat-rich genome:
    tandem arrays, homopolymer runs, reverse-complemented copies -> many refinement rounds)."""
    monkeypatch.setenv("MA_B200_IB_LARGE", "1")
    monkeypatch.setenv("MA_B200_IB_CHUNK", str(chunk))
    ctx = make_ctx("illumina")
    ctx.index_build(gold_index.forward_codes(), gold_index.contig_start, gold_index.contig_len)
    ix = ctx.index_download(gold_index.contig_names)
    assert ix.primary == gold_index.primary
    assert np.array_equal(ix.L2, gold_index.L2)
    assert np.array_equal(ix.bwt, gold_index.bwt)
    assert np.array_equal(ix.sa, gold_index.sa)
    assert np.array_equal(ix.pac[:len(gold_index.pac)], gold_index.pac)
    g = PC.repeat_rich_genome()
    lens = np.array([len(c) for c in g], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    ctx.index_build(np.concatenate(g), starts, lens)
    big = ctx.index_download(["chr1", "chr2"])
    monkeypatch.setenv("MA_B200_IB_LARGE", "0")
    ctx.index_build(np.concatenate(g), starts, lens)
    assert _index_arrays_equal(big, ctx.index_download(["chr1", "chr2"]))
    # degenerate texts: a homopolymer (every suffix ties until the end of the text), a two-base text, period 2
    for fwd in (np.zeros(3000, dtype=np.uint8), np.array([1, 2], dtype=np.uint8), np.tile([0, 3], 777).astype(np.uint8)):
        monkeypatch.setenv("MA_B200_IB_LARGE", "1")
        ctx.index_build(fwd, np.array([0]), np.array([len(fwd)]))
        a = ctx.index_download(["c"])
        monkeypatch.setenv("MA_B200_IB_LARGE", "0")
        ctx.index_build(fwd, np.array([0]), np.array([len(fwd)]))
        assert _index_arrays_equal(a, ctx.index_download(["c"])), len(fwd)
    ctx.close()


def test_batch_call_equals_staged_calls_and_sharding(gold_index):
    """ma_b200_align_batch == upload/run/download; splitting the batch (read sharding) gives the
    same records when the shard's srand_base is offset by its first read index."""
    reads = PC.read_reads_txt(PC.gold_reads("illumina"))
    data, off = api.pack_reads(reads)
    ctx = make_ctx("illumina")
    ctx.index_upload(gold_index)
    info, alns, runs, st = ctx.align_batch(data, off)
    gold = PC.load_gold("illumina")
    assert st["n_sets"] == len(gold["aln"]) // 8

    def records(info, alns, runs, lo, hi):
        out = []
        for i in range(lo, hi):
            a = alns[info["set_off"][i]:info["set_off"][i] + info["n_sets"][i]]
            for k in np.argsort(a["rank"], kind="stable"):
                x = a[k]
                out.append((i, int(x["begin_q"]), int(x["end_q"]), int(x["begin_ref"]), int(x["end_ref"]),
                            int(x["score"]), tuple(runs[x["run_off"]:x["run_off"] + x["n_runs"]].tolist())))
        return out

    full = records(info, alns, runs, 0, len(reads))
    half = len(reads) // 2
    shards = []
    for lo, hi in [(0, half), (half, len(reads))]:
        c2 = make_ctx("illumina", PC.SRAND + lo)
        c2.index_upload(gold_index)
        d2, o2 = api.pack_reads(reads[lo:hi])
        i2, a2, r2, _ = c2.align_batch(d2, o2)
        shards += [(r[0] + lo,) + r[1:] for r in records(i2, a2, r2, 0, hi - lo)]
        c2.close()
    assert shards == full
    ctx.close()


def _records(info, alns, runs, n):
    """Per read: its alignment records in the reference's result order (slab order is allocation order, not fixed)."""
    out = []
    for i in range(n):
        a = alns[info["set_off"][i]:info["set_off"][i] + info["n_sets"][i]]
        assert (a["read"] == i).all()
        for k in np.argsort(a["rank"], kind="stable"):
            x = a[k]
            out.append((i, int(x["begin_q"]), int(x["end_q"]), int(x["begin_ref"]), int(x["end_ref"]), int(x["score"]),
                        int(x["soc_index"]), tuple(runs[x["run_off"]:x["run_off"] + x["n_runs"]].tolist())))
    return out


def test_pipelined_batch_equals_one_shot(gold_index):
    """The pipelined form of ma_b200_align_batch (sub-batches alternating between two sets of device slabs) returns
    the same records and statistics, for split sizes that do and do not divide the batch."""
    reads = PC.read_reads_txt(PC.gold_reads("illumina"))
    data, off = api.pack_reads(reads)
    ctx = make_ctx("illumina")
    ctx.index_upload(gold_index)
    info, alns, runs, st = ctx.align_batch(data, off)
    full = _records(info, alns, runs, len(reads))
    for split in (7, 64, len(reads) // 2):
        ctx.set_batch_split(split)
        i2, a2, r2, s2 = ctx.align_batch(data, off)
        assert np.array_equal(i2["n_sets"], info["n_sets"])
        assert _records(i2, a2, r2, len(reads)) == full
        for k in ("n_reads", "n_seeds", "n_sets", "n_tasks", "n_runs", "n_ext", "n_invpsi", "dp_cells"):
            assert s2[k] == st[k], k
    # staged calls still work on the same context afterwards
    ctx.align_upload(data, off)
    ctx.align_run()
    i3, a3, r3 = ctx.download_alignments()
    assert _records(i3, a3, r3, len(reads)) == full
    ctx.close()


def test_edge_batches(gold_index):
    ctx = make_ctx("illumina")
    ctx.index_upload(gold_index)
    # empty batch
    info, alns, runs, st = ctx.align_batch(np.zeros(0, np.uint8), np.zeros(1, np.int64))
    assert len(alns) == 0 and st["n_reads"] == 0
    # empty read, 1-base read, all-N read, read shorter than the minimal seed
    fwd = gold_index.forward_codes()
    reads = [np.zeros(0, np.uint8), np.array([2], np.uint8), np.full(40, 4, np.uint8), fwd[100:110].copy(),
             fwd[2000:2150].copy()]
    data, off = api.pack_reads(reads)
    info, alns, runs, st = ctx.align_batch(data, off)
    assert list(info["n_sets"][:4]) == [0, 0, 0, 0] and info["n_sets"][4] >= 1
    a = alns[info["set_off"][4]:info["set_off"][4] + info["n_sets"][4]]
    best = a[np.argmin(a["rank"])]
    assert (best["begin_ref"], best["end_ref"], best["begin_q"], best["end_q"]) == (2000, 2150, 0, 150)
    assert best["score"] == 300
    ctx.close()


def test_no_index_is_an_error():
    ctx = api.Context(0, "illumina")
    ctx.align_upload(np.zeros(4, np.uint8), np.array([0, 4], np.int64))
    with pytest.raises(api.MaB200Error):
        ctx.align_run()
    ctx.close()


def test_config1_scale_against_oracle(tmp_path):
    """BASELINE.json config 1 (1 Mbp genome, 10 k x 150 bp reads, Illumina preset): GPU-built index, every stage of
    every read against the oracle."""
    g = synth.random_genome([1_000_000], 1)
    reads, _, pos, rev = synth.simulate_reads(g, 10_000, 150, 1)
    ctx = make_ctx("illumina")
    ctx.index_build(g[0], np.array([0]), np.array([1_000_000]))
    ix = ctx.index_download(["chr1"])
    index.store_index(ix, str(tmp_path / "g1"))
    synth.write_reads_txt(str(tmp_path / "r.txt"), reads)
    exp = H.oracle_align_dump(str(tmp_path / "g1"), str(tmp_path / "r.txt"), "illumina", str(tmp_path / "o.dump"),
                              PC.SRAND, 5)
    got = PC.gpu_stage_dump(ctx, reads)
    PC.assert_same_stages(got, exp, keys=["seg_off", "seg", "seed_off", "seed"], what="config1 seeding")
    # Harmonization's double-precision libm calls (atan / tan / sin / log): north_star allows only a tolerance that does
    # not change the emitted alignments, so NO read may differ (CUDA's libm has agreed with glibc on every input so far)
    bad = PC.mismatching_reads(got, exp)
    assert bad == 0, "%d of 10000 reads differ" % bad
    PC.assert_same_stages(got, exp, what="config1")
    # accuracy sanity: primary alignment within 5 bp of the simulated origin for > 99 % of the reads
    ok = 0
    for i in range(len(reads)):
        lo = got["aln_off"][i]
        if got["aln_off"][i + 1] > lo:
            b = got["aln"][lo * 8 + 2]
            truth = pos[i] if not rev[i] else 2_000_000 - (pos[i] + 168)  # reverse read = start of revcomp(168-base template)
            ok += abs(int(b) - int(truth)) <= 12
    assert ok > 9900, ok
    st1, st3 = got["_stats"]
    assert st3["n_ext"] == int(exp["work"].sum())
    ctx.close()


@pytest.mark.parametrize("preset,long_reads", PC.REPEAT_RUNS)
def test_repeat_rich_genome_against_oracle(preset, long_reads, tmp_path):
    """Repeat-rich genome (tests/pipeline_common.py repeat_rich_genome; the oracle is pinned to the live reference on it
    in test_pipeline_cpu.py): GPU-built index, reads with up to several hundred seeds, every stage."""
    g = PC.repeat_rich_genome()
    lens = np.array([len(c) for c in g], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    ctx = make_ctx(preset)
    ctx.index_build(np.concatenate(g), starts, lens)
    index.store_index(ctx.index_download(["chr1", "chr2"]), str(tmp_path / "g"))
    reads = PC.repeat_rich_reads(g, long_reads)
    synth.write_reads_txt(str(tmp_path / "r.txt"), reads)
    exp = H.oracle_align_dump(str(tmp_path / "g"), str(tmp_path / "r.txt"), preset, str(tmp_path / "o.dump"), PC.SRAND, 5)
    assert int(np.diff(exp["seed_off"]).max()) > 100
    got = PC.gpu_stage_dump(ctx, reads, keep_segments=8192 if long_reads else 4096)
    PC.assert_same_stages(got, exp, what="repeat-rich " + preset)
    ctx.close()


@pytest.mark.parametrize("preset", ["illumina", "default", "nanopore"])
def test_ragged_read_lengths_against_oracle(preset, gold_index, tmp_path):
    """Read lengths from 17 to 2000 in one batch: minimal seed length, the 800-base switch of the harmonization
    heuristics, reads longer than the DP padding (oracle pinned to the live reference in test_pipeline_cpu.py)."""
    reads = PC.ragged_reads(gold_index.forward_codes())
    PC.write_ragged_txt(str(tmp_path / "r.txt"), reads)
    exp = H.oracle_align_dump(PC.GOLD_PREFIX, str(tmp_path / "r.txt"), preset, str(tmp_path / "o.dump"), PC.SRAND, 5)
    ctx = make_ctx(preset)
    ctx.index_upload(gold_index)
    got = PC.gpu_stage_dump(ctx, reads, keep_segments=8192)
    PC.assert_same_stages(got, exp, what="ragged " + preset)
    mq = PC.gpu_mapq_dump(ctx, reads, api.preset(preset))
    for k in ("mq_off", "mq"):
        assert np.array_equal(mq[k], exp[k]), (preset, k)
    ctx.close()


def test_long_reads_against_oracle(tmp_path):
    """PacBio preset (maxSpan seeding, long banded DP incl. the 1024-column window): 30 x 4 kbp reads, 12 % error."""
    g = synth.random_genome([400_000, 200_000], 21)
    reads, _, _, _ = synth.simulate_long_reads(g, 30, 4000, 22)
    ctx = make_ctx("pacbio")
    ctx.index_build(np.concatenate(g), np.array([0, 400_000]), np.array([400_000, 200_000]))
    ix = ctx.index_download(["c1", "c2"])
    index.store_index(ix, str(tmp_path / "g"))
    synth.write_reads_txt(str(tmp_path / "r.txt"), reads)
    exp = H.oracle_align_dump(str(tmp_path / "g"), str(tmp_path / "r.txt"), "pacbio", str(tmp_path / "o.dump"),
                              PC.SRAND, 5)
    got = PC.gpu_stage_dump(ctx, reads, keep_segments=8192)
    PC.assert_same_stages(got, exp, what="pacbio")
    ctx.close()


@pytest.mark.parametrize("overrides", [{"bandwidth_ext": 24}, {"bandwidth_ext": 70, "zdrop": 30},
                                       {"bandwidth_ext": 40, "padding": 120, "max_gap_area": 5}])
def test_non_default_dp_parameters_against_oracle(gold_index, overrides, tmp_path):
    """Narrow extension bands make the in-band fast mode of the DP kernel hit the band limit (r > w) and fall back
    to the exact 16-aligned mode; small z-drop / padding / gap-area values exercise the dual extension."""
    ctx = api.Context(0, "illumina")
    p = api.preset("illumina")
    p.srand_base = PC.SRAND
    for k, v in overrides.items():
        setattr(p, k, v)
    ctx.set_params(p)
    ctx.index_upload(gold_index)
    got = PC.gpu_stage_dump(ctx, PC.read_reads_txt(PC.gold_reads("illumina")))
    exp = H.oracle_align_dump(PC.GOLD_PREFIX, PC.gold_reads("illumina"), "illumina", str(tmp_path / "o.dump"),
                              PC.SRAND, 5, overrides)
    PC.assert_same_stages(got, exp, what=str(overrides))
    ctx.close()


def test_set_params_rejects_what_is_not_implemented():
    """Values the device path cannot honour are refused (MA_B200_EINVAL) instead of giving non-reference results:
    the non-rectangular SoC of the SV presets, more SoCs than the per-read capacity, nonsense scores."""
    ctx = api.Context(0, "illumina")
    for field, value in (("rectangular_soc", 0), ("max_num_soc", 129), ("max_num_soc", 0), ("seeding_technique", 2),
                         ("match", 0), ("extend2", -1), ("bandwidth_ext", 0)):
        p = api.preset("illumina")
        setattr(p, field, value)
        with pytest.raises(api.MaB200Error, match="error -2"):
            ctx.set_params(p)
    ctx.set_params(api.preset("illumina"))  # the context keeps working
    ctx.close()


def test_read_beyond_a_capacity_is_reported_per_read(gold_index):
    """A DP problem wider than the largest band window (bandwidth_ext 2500, padding 3000, a 2 300-base unalignable read tail) flags
    ONE read (ma_b200_read_info.status, stats.n_failed) — the records of all other reads of the batch are the same as
    without it (the reference has no such capacity; a batch must not die of one read)."""
    reads = [r for r in PC.read_reads_txt(PC.gold_reads("pacbio"))]
    fwd = gold_index.forward_codes()
    rng = np.random.Generator(np.random.PCG64(5))
    bad = np.concatenate([fwd[2000:3500], rng.integers(0, 4, 2300, dtype=np.uint8)])
    out = []
    for batch in (reads, reads + [bad]):
        ctx = api.Context(0, "pacbio")
        p = api.preset("pacbio")
        p.srand_base, p.bandwidth_ext, p.padding = PC.SRAND, 2500, 3000
        ctx.set_params(p)
        ctx.index_upload(gold_index)
        data, off = api.pack_reads(batch)
        info, alns, runs, st = ctx.align_batch(data, off)
        out.append((info, _records(info, alns, runs, len(reads)), st))
        ctx.close()
    (i0, r0, s0), (i1, r1, s1) = out
    n = len(reads)
    # (with these parameters a few of the golden reads run into the capacity themselves: equally in both batches)
    assert np.array_equal(i0["status"], i1["status"][:n]) and (i0["status"] == 0).sum() >= 3
    assert i1["status"][n] == api.READ_EBAND, (i1[n], s1)
    assert s0["n_failed"] == int((i0["status"] != 0).sum()) and s1["n_failed"] == s0["n_failed"] + 1
    ok = set(np.nonzero(i0["status"] == 0)[0].tolist())
    assert [r for r in r0 if r[0] in ok] == [r for r in r1 if r[0] in ok] and len(ok) > 0


def test_two_gpus_sharded_batch_equals_one_gpu(gold_index):
    """SURVEY.md §8(e) on real devices: the index replicated on cuda:0 and cuda:1, the paired golden batch split into the
    contiguous pair shards of ma_b200.dist.shard_pairs (what bench.py --config 2 does per rank), RANSAC streams offset by
    the shard's first read — the concatenated records equal the one-GPU batch. Skips only on a box with one GPU."""
    import torch
    from ma_b200 import dist as madist
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    reads = PC.read_reads_txt(PC.gold_reads("illuminapaired"))
    n = len(reads)

    def run(device, lo, hi):
        ctx = api.Context(device, "illuminapaired")
        p = api.preset("illuminapaired")
        p.srand_base = madist.shard_srand_base(PC.SRAND, lo)
        ctx.set_params(p)
        ctx.index_upload(gold_index)
        data, off = api.pack_reads(reads[lo:hi])
        info, alns, runs, st = ctx.align_batch(data, off)
        rec = [(r[0] + lo,) + r[1:] for r in _records(info, alns, runs, hi - lo)]
        flags = [(int(a["read"]) + lo, int(a["rank"]), int(a["flags"]), float(a["mapq"]) if a["mapq"] == a["mapq"] else -1.0,
                  int(a["rank_mq"]), int(a["pair_rank"])) for a in alns]
        ctx.close()
        return rec, sorted(flags)

    full, full_flags = run(0, 0, n)
    parts, part_flags = [], []
    for rank in range(2):
        lo, hi = madist.shard_pairs(n // 2, rank, 2)
        r, f = run(rank, lo, hi)
        parts += r
        part_flags += f
    assert parts == full and sorted(part_flags) == full_flags


@pytest.mark.parametrize("preset", ["illumina", "illuminapaired", "pacbio"])
def test_reported_only_output_is_the_mapping_quality_vector(preset, gold_index):
    """ma_b200_set_reported_only: the batch call hands back exactly the records with rank_mq >= 0 of the full output
    (same fields, same run words), packed per read, and stats.n_reported counts them."""
    reads = PC.read_reads_txt(PC.gold_reads(preset))
    data, off = api.pack_reads(reads)

    def run(reported_only):
        ctx = make_ctx(preset)
        ctx.index_upload(gold_index)
        ctx.set_reported_only(reported_only)
        info, alns, runs, st = ctx.align_batch(data, off, cap_alns=20000, cap_runs=4_000_000)
        ctx.close()
        rows = []
        for i in range(len(reads)):
            a = alns[info["set_off"][i]:info["set_off"][i] + info["n_sets"][i]]
            assert (a["read"] == i).all()
            for x in a[np.argsort(a["rank"], kind="stable")]:
                rows.append((i, int(x["rank"]), int(x["rank_mq"]), int(x["pair_rank"]), int(x["flags"]), int(x["score"]),
                             int(x["begin_ref"]), int(x["end_ref"]), int(x["begin_q"]), int(x["end_q"]),
                             x["mapq"].tobytes(), tuple(runs[x["run_off"]:x["run_off"] + x["n_runs"]].tolist())))
        return rows, st

    full, st_full = run(False)
    rep, st_rep = run(True)
    assert rep == [r for r in full if r[2] >= 0]
    assert st_rep["n_reported"] == len(rep) and st_rep["n_sets"] == st_full["n_sets"]
    assert len(rep) < len(full) if preset == "illumina" else len(rep) <= len(full)  # (illumina drops records, n-best)


def test_sibling_context_shares_the_index_two_batches_in_flight(gold_index):
    """ma_b200_create_sibling: a second context on the device that views the first one's index. Two host threads, one
    batch each at the same time (what bench.py's e2e figure and maCMD_b200 do): both get the records of a single call;
    destroying the sibling leaves the parent intact."""
    import threading
    reads = PC.read_reads_txt(PC.gold_reads("illumina"))
    data, off = api.pack_reads(reads)
    ctx = make_ctx("illumina")
    ctx.index_upload(gold_index)
    info, alns, runs, _ = ctx.align_batch(data, off)
    expect = _records(info, alns, runs, len(reads))
    sib = ctx.sibling()
    p = api.preset("illumina")
    p.srand_base = PC.SRAND
    sib.set_params(p)
    got = {}

    def work(name, c):
        for _ in range(3):
            i, a, r, _st = c.align_batch(data, off)
            got[name] = _records(i, a, r, len(reads))

    ts = [threading.Thread(target=work, args=("parent", ctx)), threading.Thread(target=work, args=("sibling", sib))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert got["parent"] == expect and got["sibling"] == expect
    sib.close()
    info, alns, runs, _ = ctx.align_batch(data, off)
    assert _records(info, alns, runs, len(reads)) == expect
    ctx.close()
