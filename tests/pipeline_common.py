"""Helpers that turn the C-ABI outputs into the same per-stage arrays as the reference dump (oracle/ref_dump.cpp)."""
import os

import numpy as np

import helpers as H
from ma_b200 import api, index, synth

SRAND = 1000
GOLD_PREFIX = os.path.join(H.GOLDEN, "gold")


def read_reads_txt(path):
    with open(path) as f:
        return [synth.text_to_codes(line.strip()) for line in f if line.strip()]


def load_gold(preset):
    g = np.load(os.path.join(H.GOLDEN, "gold_%s.npz" % preset))
    return {k: g[k].astype(np.int64) for k in g.files}


def gold_reads(preset):
    name = ("gold_reads_pairs.txt" if preset == "illuminapaired" else
            "gold_reads_short.txt" if preset in ("illumina", "default") else "gold_reads_long.txt")
    return os.path.join(H.GOLDEN, name)


def gpu_stage_dump(ctx, reads, keep_segments=4096):
    """Runs the path through the C ABI and returns a dict with the reference dump's keys (where exposed)."""
    data, off = api.pack_reads(reads)
    n = len(off) - 1
    out = {}
    ctx.align_upload(data, off)
    st1 = ctx.align_run(api.STAGE_SEEDS, keep_segments)
    info = ctx.download_info()
    segs, nseg = ctx.download_segments()
    assert (nseg <= keep_segments).all()
    seg_rows = []
    for i in range(n):
        s = segs[i, :nseg[i]]
        seg_rows.append(np.stack([s["start"], s["size"], s["sa_start"], s["sa_rev"], s["sa_size"]], axis=1)
                        .astype(np.int64).reshape(-1))
    out["seg_off"] = np.concatenate([[0], np.cumsum(nseg)]).astype(np.int64)
    out["seg"] = np.concatenate(seg_rows) if seg_rows else np.zeros(0, np.int64)
    seeds = ctx.download_seeds()
    rows = []
    for i in range(n):
        s = seeds[info["seed_off"][i]:info["seed_off"][i] + info["n_seeds"][i]]
        rows.append(np.stack([s["q"], s["len"], s["r"], s["amb"], s["fw"], s["delta"]], axis=1).astype(np.int64)
                    .reshape(-1))
    out["seed_off"] = np.concatenate([[0], np.cumsum(info["n_seeds"])]).astype(np.int64)
    out["seed"] = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    st3 = ctx.align_run(api.STAGE_ALIGN, 0)
    sets, sseeds = ctx.download_sets()
    info, alns, runs = ctx.download_alignments()
    harm_off, harmseed_off, harmseed = [0], [0], []
    aln_off, aln, alndata_off, alndata = [0], [], [0], []
    for i in range(n):
        so, ns = int(info["set_off"][i]), int(info["n_sets"][i])
        for h in sets[so:so + ns]:
            s = sseeds[h["seed_off"]:h["seed_off"] + h["n"]]
            harmseed.append(np.stack([s["q"], s["len"], s["r"], s["fw"], np.full(len(s), h["soc_index"])], axis=1)
                            .astype(np.int64).reshape(-1))
            harmseed_off.append(harmseed_off[-1] + len(s))
        harm_off.append(len(harmseed_off) - 1)
        a = alns[so:so + ns]
        for k in np.argsort(a["rank"], kind="stable"):
            x = a[k]
            aln.append([x["begin_q"], x["end_q"], x["begin_ref"], x["end_ref"], x["score"], x["soc_index"],
                        x["length"], x["n_runs"]])
            r = runs[x["run_off"]:x["run_off"] + x["n_runs"]]
            alndata.append(np.stack([r & 7, r >> 3], axis=1).astype(np.int64).reshape(-1))
            alndata_off.append(alndata_off[-1] + len(r))
        aln_off.append(len(aln))
    out["harm_off"] = np.array(harm_off, dtype=np.int64)
    out["harmseed_off"] = np.array(harmseed_off, dtype=np.int64)
    out["harmseed"] = np.concatenate(harmseed) if harmseed else np.zeros(0, np.int64)
    out["aln_off"] = np.array(aln_off, dtype=np.int64)
    out["aln"] = np.array(aln, dtype=np.int64).reshape(-1)
    out["alndata_off"] = np.array(alndata_off, dtype=np.int64)
    out["alndata"] = np.concatenate(alndata) if alndata else np.zeros(0, np.int64)
    out["_stats"] = (st1, st3)
    return out


def gpu_mapq_dump(ctx, reads, params):
    """MappingQuality rows (run without pairing) and PairedReads rows (run with use_paired_reads) in the layout of
    oracle/ref_dump.cpp: mq = {NW index, secondary | supplementary << 1, mapq bits}, pr = {mate, NW index, flags,
    mapq bits}."""
    data, off = api.pack_reads(reads)
    n = len(off) - 1
    out = {}
    saved = params.use_paired_reads
    for paired in (0, 1):
        params.use_paired_reads = paired
        ctx.set_params(params)
        ctx.align_upload(data, off)
        ctx.align_run(api.STAGE_MAPQ, 0)
        info, alns, runs = ctx.download_alignments()
        bits = alns["mapq"].view(np.int64)
        if not paired:
            rows, offs = [], [0]
            for i in range(n):
                so, ns = int(info["set_off"][i]), int(info["n_sets"][i])
                a = alns[so:so + ns]
                keep = np.nonzero(a["rank_mq"] >= 0)[0]
                for k in keep[np.argsort(a["rank_mq"][keep], kind="stable")]:
                    rows.append([a["rank"][k], a["flags"][k] & 3, bits[so + k]])
                offs.append(len(rows))
            out["mq_off"] = np.array(offs, dtype=np.int64)
            out["mq"] = np.array(rows, dtype=np.int64).reshape(-1)
        else:
            rows, offs = [], [0]
            for p in range(n // 2):
                cand = []
                for mate in (0, 1):
                    i = 2 * p + mate
                    so, ns = int(info["set_off"][i]), int(info["n_sets"][i])
                    a = alns[so:so + ns]
                    for k in np.nonzero(a["pair_rank"] >= 0)[0]:
                        assert bool(a["flags"][k] & api.ALN_FIRST_MATE) == (mate == 0)
                        cand.append((int(a["pair_rank"][k]), mate, int(a["rank"][k]), int(a["flags"][k] & 3),
                                     int(bits[so + k])))
                for c in sorted(cand):
                    rows.append(list(c[1:]))
                offs.append(len(rows))
            out["pr_off"] = np.array(offs, dtype=np.int64)
            out["pr"] = np.array(rows, dtype=np.int64).reshape(-1)
    params.use_paired_reads = saved
    ctx.set_params(params)
    return out


STAGE_KEYS = ["seg_off", "seg", "seed_off", "seed", "harm_off", "harmseed_off", "harmseed", "aln_off", "aln",
              "alndata_off", "alndata"]


def assert_same_stages(got, exp, keys=STAGE_KEYS, what=""):
    for k in keys:
        a, b = np.asarray(got[k]), np.asarray(exp[k])
        assert a.shape == b.shape, "%s %s: shape %s vs %s" % (what, k, a.shape, b.shape)
        if not np.array_equal(a, b):
            d = np.nonzero(a != b)[0]
            raise AssertionError("%s %s: %d mismatching entries, first at %d: %s vs %s" %
                                 (what, k, len(d), d[0], a[d[:6]], b[d[:6]]))


def mismatching_reads(got, exp):
    """Number of reads whose alignment records differ (used where a floating-point tolerance is stated)."""
    n = len(exp["aln_off"]) - 1
    bad = 0
    for i in range(n):
        ga = got["aln"][got["aln_off"][i] * 8:got["aln_off"][i + 1] * 8]
        ea = exp["aln"][exp["aln_off"][i] * 8:exp["aln_off"][i + 1] * 8]
        if len(ga) != len(ea) or not np.array_equal(ga, ea):
            bad += 1
    return bad


# Parameter variations of the presets (ma_b200_params field names) that change at least one stage's output on the
# golden reads. test_pipeline_cpu.py pins the oracle to the LIVE reference for each of them (where /root/reference
# exists), test_pipeline_gpu.py compares the device path with the oracle. "genome_size_disable": 0 switches the
# large-genome heuristics on (they are off below 10 M bases).
_H = {"genome_size_disable": 0}
PARAM_VARIATIONS = [
    ("illumina", {"min_seed_length": 10}), ("illumina", {"min_seed_length": 25}),
    ("illumina", {"max_num_soc": 1, "min_num_soc": 1}), ("illumina", {"seeding_technique": 0}),
    ("pacbio", {"seeding_technique": 1}), ("illumina", dict(_H, harm_score_min=40)),
    ("illumina", dict(_H, harm_score_min_rel=0.2)), ("illumina", dict(_H, seed_drop_min_size=30, seed_drop_factor=0.05)),
    ("illumina", dict(_H, disable_heuristics=1)), ("illumina", dict(_H, max_ambiguity=1)),
    ("pacbio", dict(_H, max_ambiguity=1)), ("pacbio", dict(_H, min_num_soc=1, max_num_soc=1)),
    ("pacbio", dict(_H, harm_score_min=200)), ("pacbio", {"max_gap_area": 200}), ("illumina", {"max_gap_area": 0}),
    ("illumina", {"padding": 50}), ("pacbio", {"bandwidth_ext": 64, "zdrop": 50}), ("illumina", {"report_n": 1}),
    ("illumina", {"min_alignment_score": 10, "max_supplementary_per_prim": 3, "max_overlap_supplementary": 0.5}),
    ("illuminapaired", {"paired_mean": 300.0, "paired_std": 20.0, "paired_bonus": 2.0}),
    ("illumina", dict(_H, soc_score_drop=0.9, switch_qlen=10, score_diff_tolerance=0.5, max_score_lookahead=1)),
    ("pacbio", dict(_H, max_delta_dist=0.01, min_delta_dist=100, gap_cost_cutting=0, optimistic_gap_estimation=0)),
]


def repeat_rich_genome(seed=31):
    """Two contigs full of repeats: a tandem array of 60 diverged 300 bp units, 40 interspersed exact copies (and
    reverse complements) of a 120 bp unit, a dinucleotide repeat and a homopolymer. Reads from it carry hundreds of
    ambiguous seeds: long std::sort / heap runs with many equal keys, overlapping SoC windows."""
    rng = np.random.Generator(np.random.PCG64(seed))

    def mutate(u, rate):
        u = u.copy()
        m = rng.random(len(u)) < rate
        u[m] = (u[m] + rng.integers(1, 4, m.sum())) & 3
        return u

    unit = rng.integers(0, 4, 300).astype(np.uint8)
    c1 = [rng.integers(0, 4, 5000).astype(np.uint8)] + [mutate(unit, 0.02) for _ in range(60)]
    c1.append(rng.integers(0, 4, 5000).astype(np.uint8))
    unit2 = rng.integers(0, 4, 120).astype(np.uint8)
    c2 = []
    for k in range(40):
        c2.append(rng.integers(0, 4, int(rng.integers(200, 600))).astype(np.uint8))
        c2.append(unit2 if k % 3 else (3 - unit2[::-1]).astype(np.uint8))
    c2.append(np.tile(np.array([0, 1], dtype=np.uint8), 200))
    c2.append(np.zeros(300, dtype=np.uint8))
    c2.append(rng.integers(0, 4, 3000).astype(np.uint8))
    return [np.concatenate(c1), np.concatenate(c2)]


REPEAT_RUNS = [("illumina", False), ("default", False), ("pacbio", True), ("nanopore", True)]


def repeat_rich_reads(genome, long_reads):
    from ma_b200 import synth
    if long_reads:
        return synth.simulate_long_reads(genome, 10, 3000, 33)[0]
    return synth.simulate_reads(genome, 500, 150, 32, sub_rate=0.01, ins_rate=0.002, del_rate=0.002)[0]


RAGGED_LENGTHS = [17, 18, 20, 35, 50, 75, 100, 151, 250, 400, 799, 800, 801, 1200, 2000]


def ragged_reads(fwd, seed=41):
    """Reads of many lengths from the golden genome (both strands, 2 % substitutions, a few indels): minimal seed
    length, the 800-base switch of the harmonization heuristics, reads longer than the DP padding."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for L in RAGGED_LENGTHS:
        for k in range(6):
            p = int(rng.integers(0, len(fwd) - L - 5))
            r = fwd[p:p + L].copy()
            m = rng.random(L) < 0.02
            r[m] = (r[m] + rng.integers(1, 4, m.sum())) & 3
            if L > 40 and k % 3 == 1:
                c = int(rng.integers(10, L - 10))
                r = np.delete(r, c)
            if L > 40 and k % 3 == 2:
                c = int(rng.integers(10, L - 10))
                r = np.insert(r, c, rng.integers(0, 4, 2))
            if k % 2:
                r = (3 - r[::-1]).astype(np.uint8)
            out.append(r.astype(np.uint8))
    return out


def write_ragged_txt(path, reads):
    with open(path, "w") as f:
        for r in reads:
            f.write("".join("ACGTN"[c] for c in r) + "\n")
