"""BASELINE.json configs[1] at FULL size (100 Mbp genome, 1 M simulated 2x150 pairs, Illumina_Paired preset) through the
C ABI: the oracle cannot be run at this size, so the checks are size-independent properties of the path —
  * accuracy against the simulation's truth (the primary alignment of a mate starts where the mate was drawn),
  * determinism (two runs give the same records),
  * sharding invariance: a contiguous shard of pairs aligned on its own, with the RANSAC stream offset
    by its first read index, gives the records it has inside the full batch,
  * conservation: every read is reported, mates of a pair stay together, counters add up.
Bit-exact parity with the reference is established at oracle sizes in test_pipeline_gpu.py; this file guards the scale."""
import numpy as np
import pytest

from ma_b200 import api, synth

N_PAIRS = 1_000_000
GENOME_MBP = 100
SRAND = 77


def record_table(info, alns, runs):
    """One row per alignment, ordered by (read, rank): everything the writer consumes, with a checksum of the runs."""
    order = np.lexsort((alns["rank"], alns["read"]))
    a = alns[order]
    n_runs = a["n_runs"].astype(np.int64)
    start = np.repeat(a["run_off"], n_runs)
    within = np.arange(n_runs.sum(), dtype=np.int64) - np.repeat(np.cumsum(n_runs) - n_runs, n_runs)
    words = runs[start + within].astype(np.uint64)
    mixed = words * (within.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1))
    csum = np.zeros(len(a), dtype=np.uint64)
    nz = n_runs > 0
    csum[nz] = np.add.reduceat(mixed, (np.cumsum(n_runs) - n_runs)[nz])
    cols = [a["read"].astype(np.int64), a["rank"].astype(np.int64), a["begin_ref"], a["end_ref"], a["score"],
            a["begin_q"].astype(np.int64), a["end_q"].astype(np.int64), a["flags"].astype(np.int64),
            a["mapq"].view(np.int64), a["rank_mq"].astype(np.int64), a["pair_rank"].astype(np.int64),
            csum.view(np.int64)]
    return np.stack(cols, axis=1)


@pytest.fixture(scope="module")
def workload():
    n_contigs = 10
    genome = synth.random_genome([GENOME_MBP * 1_000_000 // n_contigs] * n_contigs, 2)
    m1, m2, cid, pos, flen, rev = synth.simulate_pairs(genome, N_PAIRS, 150, 2017)
    reads = np.empty((2 * N_PAIRS, 150), dtype=np.uint8)
    reads[0::2], reads[1::2] = m1, m2
    lens = np.array([len(c) for c in genome], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    return genome, reads, (cid, pos, flen, rev), lens, starts


@pytest.mark.gpu
def test_full_size_configs1_properties(workload):
    genome, reads, (cid, pos, flen, rev), lens, starts = workload
    fwd = int(lens.sum())

    def make(srand):
        ctx = api.Context(0, "illumina_paired")
        p = api.preset("illumina_paired")
        p.srand_base = srand
        ctx.set_params(p)
        return ctx

    ctx = make(SRAND)
    ctx.index_build(np.concatenate(genome), starts, lens.tolist())
    data, off = api.pack_reads(reads)
    ctx.align_upload(data, off)
    st = ctx.align_run(api.STAGE_MAPQ)
    info, alns, runs = ctx.download_alignments()
    assert st["n_reads"] == 2 * N_PAIRS and st["n_dropped"] == 0
    assert int(info["n_sets"].sum()) == st["n_sets"] == len(alns)
    full = record_table(info, alns, runs)

    # conservation: (almost) every simulated read gets an alignment; every record belongs to its read's slab range
    assert (info["n_sets"] > 0).mean() > 0.999
    assert (alns["read"] >= 0).all() and (alns["read"] < 2 * N_PAIRS).all()

    # accuracy against the truth of the simulation
    prim = alns[alns["rank_mq"] == 0]
    r = prim["read"].astype(np.int64)
    pair, mate = r // 2, r % 2
    g0 = starts[cid[pair]] + pos[pair]
    is_left = (mate == 0) != rev[pair]  # the mate drawn from the fragment's 5' end on the forward strand
    expect = np.where(is_left, g0, 2 * fwd - g0 - flen[pair])
    got = prim["begin_ref"] - prim["begin_q"]
    ok = np.abs(got - expect) <= 10
    assert len(prim) > 0.999 * 2 * N_PAIRS
    assert ok.mean() > 0.98, ok.mean()

    # determinism: the same batch again
    st2 = ctx.align_run(api.STAGE_MAPQ)
    info2, alns2, runs2 = ctx.download_alignments()
    assert st2["n_sets"] == st["n_sets"] and st2["dp_cells"] == st["dp_cells"]
    assert np.array_equal(record_table(info2, alns2, runs2), full)
    ctx.close()

    # sharding invariance: pairs [lo, hi) on their own context (index replicated, no exchange between shards)
    lo, hi = 2 * 400_000, 2 * 500_000
    c2 = make(SRAND + lo)
    c2.index_build(np.concatenate(genome), starts, lens.tolist())
    d2, o2 = api.pack_reads(reads[lo:hi])
    c2.align_upload(d2, o2)
    c2.align_run(api.STAGE_MAPQ)
    i2, a2, r2 = c2.download_alignments()
    shard = record_table(i2, a2, r2)
    shard[:, 0] += lo
    sel = (full[:, 0] >= lo) & (full[:, 0] < hi)
    assert np.array_equal(shard, full[sel])
    c2.close()


@pytest.mark.gpu
def test_full_size_sample_against_oracle(workload, tmp_path):
    """Bit-exact parity ON the BASELINE configuration: the 100 Mbp index (built on the GPU, stored in the reference's
    file formats) and the first 3 000 reads of the 2 M-read batch go through the oracle; every stage of those reads
    must equal what the device produced for them inside the full batch (a 200 Mbp text is above the 10 M threshold of
    the large-genome heuristics and has a different ambiguity regime than the small fixtures)."""
    import helpers as H
    import pipeline_common as PC
    from ma_b200 import index
    genome, reads, _, lens, starts = workload
    n = 3000
    ctx = api.Context(0, "illumina_paired")
    p = api.preset("illumina_paired")
    p.srand_base = PC.SRAND
    ctx.set_params(p)
    ctx.index_build(np.concatenate(genome), starts, lens.tolist())
    index.store_index(ctx.index_download(["chr%d" % (i + 1) for i in range(len(lens))]), str(tmp_path / "g"))
    synth.write_reads_txt(str(tmp_path / "r.txt"), reads[:n])
    exp = H.oracle_align_dump(str(tmp_path / "g"), str(tmp_path / "r.txt"), "illuminapaired", str(tmp_path / "o.dump"),
                              PC.SRAND, 5)
    got = PC.gpu_stage_dump(ctx, reads[:n])
    PC.assert_same_stages(got, exp, keys=["seg_off", "seg", "seed_off", "seed"], what="full-size seeding")
    bad = PC.mismatching_reads(got, exp)
    assert bad == 0, "%d of %d reads differ" % (bad, n)
    PC.assert_same_stages(got, exp, what="full-size sample")
    mq = PC.gpu_mapq_dump(ctx, reads[:n], p)
    for k in ("mq_off", "mq", "pr_off", "pr"):
        assert np.array_equal(mq[k], exp[k]), k
    ctx.close()
    # ... and, without the oracle in between, what the UNMODIFIED reference produced for these reads on the index of its
    # own builder (tests/golden/make_golden_full_size.py): SHA-1 of every stage's dump
    import hashlib
    import json
    import os
    pin = json.load(open(os.path.join(H.GOLDEN, "full_size_sample_sha1.json")))
    assert pin["n_reads"] == n and pin["srand_base"] == PC.SRAND
    print("full-size sample: %d of %d reads differ from the oracle" % (bad, n))
    if bad == 0:
        got.update(mq)
        for k, h in pin["sha1"].items():
            mine = hashlib.sha1(np.ascontiguousarray(np.asarray(got[k], dtype=np.int64)).tobytes()).hexdigest()
            assert mine == h, "stage %s differs from the reference's dump" % k
        print("full-size sample: all %d stage dumps hash-identical to the unmodified reference" % len(pin["sha1"]))


@pytest.mark.gpu
def test_full_size_100k_reads_against_reference_hashes(workload):
    """The first 100 000 reads (5 %) of the BASELINE batch: every stage's dump, mapping qualities and pairing hash-identical
    to what the UNMODIFIED reference wrote for them on the index of its own builder
    (tests/golden/make_golden_full_size.py, key "large_sample"). No tolerance: one read with a different record fails."""
    import hashlib
    import json
    import os
    import helpers as H
    import pipeline_common as PC
    genome, reads, _, lens, starts = workload
    pin = json.load(open(os.path.join(H.GOLDEN, "full_size_sample_sha1.json")))
    if "large_sample" not in pin:
        pytest.skip("golden hashes of the large sample not generated")
    n = pin["large_sample"]["n_reads"]
    ctx = api.Context(0, "illumina_paired")
    p = api.preset("illumina_paired")
    p.srand_base = PC.SRAND
    ctx.set_params(p)
    ctx.index_build(np.concatenate(genome), starts, lens.tolist())
    got = PC.gpu_stage_dump(ctx, reads[:n], keep_segments=1024)
    got.update(PC.gpu_mapq_dump(ctx, reads[:n], p))
    ctx.close()
    for k, h in pin["large_sample"]["sha1"].items():
        mine = hashlib.sha1(np.ascontiguousarray(np.asarray(got[k], dtype=np.int64)).tobytes()).hexdigest()
        assert mine == h, "stage %s of the %d-read sample differs from the reference's dump" % (k, n)


@pytest.mark.gpu
def test_full_size_pacbio_sample_against_reference(workload):
    """configs[3]-shaped long reads (200 x 10 kbp, 12 % error, PacBio preset) on the 100 Mbp index: every stage and the
    mapping qualities hash-identical to the unmodified reference's dumps (tests/golden/full_size_sample_sha1.json)."""
    import hashlib
    import json
    import os
    import helpers as H
    import pipeline_common as PC
    genome, _, _, lens, starts = workload
    pin = json.load(open(os.path.join(H.GOLDEN, "full_size_sample_sha1.json")))["pacbio"]
    reads, *_ = synth.simulate_long_reads(genome, pin["n_reads"], pin["read_len"], pin["seed"])
    ctx = api.Context(0, "pacbio")
    p = api.preset("pacbio")
    p.srand_base = PC.SRAND
    ctx.set_params(p)
    ctx.index_build(np.concatenate(genome), starts, lens.tolist())
    got = PC.gpu_stage_dump(ctx, reads, keep_segments=16384)
    got.update(PC.gpu_mapq_dump(ctx, reads, p))
    ctx.close()
    diff = [k for k, h in pin["sha1"].items()
            if hashlib.sha1(np.ascontiguousarray(np.asarray(got[k], dtype=np.int64)).tobytes()).hexdigest() != h]
    assert not diff, diff


@pytest.mark.gpu
def test_full_size_cli_sam_against_reference(workload, tmp_path):
    """End to end on the BASELINE configuration: index files written from the GPU-built index, the first 3 000 reads as
    FASTA, maCMD_b200 -> SAM; the file is byte-identical (SHA-1) to the one the unmodified reference's FileReader-free
    harness + PairedFileWriter wrote for the index of its own builder."""
    import hashlib
    import json
    import os
    import subprocess
    import helpers as H
    import pipeline_common as PC
    from ma_b200 import index
    genome, reads, _, lens, starts = workload
    pin = json.load(open(os.path.join(H.GOLDEN, "full_size_sample_sha1.json")))
    n = pin["n_reads"]
    ctx = api.Context(0, "illumina_paired")
    ctx.index_build(np.concatenate(genome), starts, lens.tolist())
    index.store_index(ctx.index_download(["chr%d" % (i + 1) for i in range(len(lens))]), str(tmp_path / "g"))
    ctx.close()
    with open(tmp_path / "r.fa", "w") as f:
        for i in range(n):
            f.write(">r%d\n%s\n" % (i, "".join("ACGT"[c] for c in reads[i])))
    cli = os.path.join(H.ROOT, "ma_b200", "cli", "maCMD_b200")
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(cli)])
    out = subprocess.check_output([cli, "-x", str(tmp_path / "g"), "-i", str(tmp_path / "r.fa"), "-p", "Illumina_Paired",
                                   "--Interleaved", "--Srand", str(pin["srand_base"]), "--Batch", "1024"])
    assert out.count(b"\n") == pin["sam_lines"]
    assert hashlib.sha1(out).hexdigest() == pin["sam_sha1"]
