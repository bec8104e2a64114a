"""GPU parity: the sm_100a banded-DP kernel through the C ABI (ma_b200_ksw_batch) vs golden vectors and the oracle.

Bit-exact bar: every kswcpp_extz_t field and every CIGAR word (integer work)."""
import os

import numpy as np
import pytest

import dpgen
import helpers as H
from ma_b200 import api

pytestmark = pytest.mark.gpu
FIELDS = ["max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar", "reach_end"]


def check_against_oracle(ctx, pairs):
    tasks, seq = api.pack_ksw_tasks(pairs)
    res, cig = ctx.ksw_batch(tasks, seq)
    assert (res["status"] == 0).all()
    cells = 0
    for i, (w, zd, fl, q, t) in enumerate(pairs):
        exp, ecig, ecells = H.oracle_ksw(q, t, w, zd, fl)
        for k in FIELDS:
            assert int(res[k][i]) == exp[k], (i, k, int(res[k][i]), exp[k], len(q), len(t), w, zd, fl)
        got = cig[res["cigar_off"][i]:res["cigar_off"][i] + res["n_cigar"][i]]
        assert np.array_equal(got, ecig), (i, got[:6], ecig[:6])
        assert int(res["cells"][i]) == ecells
        cells += ecells
    return cells


def test_golden_reference_vectors(gpu_ctx):
    g = np.load(os.path.join(H.GOLDEN, "ksw_golden.npz"))
    d = {"ksw_calls": g["calls"].astype(np.int64), "ksw_seq": g["seq"], "ksw_cigar": g["cigar"]}
    calls = list(H.split_ksw_dump(d))
    tasks, seq = api.pack_ksw_tasks([(f["w"], f["zdrop"], f["flag"], q, t) for f, q, t, c in calls])
    res, cig = gpu_ctx.ksw_batch(tasks, seq)
    for i, (f, q, t, c) in enumerate(calls):
        for k in FIELDS:
            assert int(res[k][i]) == f[k], (i, k, int(res[k][i]), f[k], f)
        got = cig[res["cigar_off"][i]:res["cigar_off"][i] + res["n_cigar"][i]]
        assert np.array_equal(got, c), (i, f)


def test_random_mixed_vs_oracle(gpu_ctx):
    check_against_oracle(gpu_ctx, dpgen.random_pairs(1500, seed=4242))


@pytest.mark.parametrize("mode", [dpgen.GLOBAL, dpgen.EXT, dpgen.EXT_RIGHT])
@pytest.mark.parametrize("length,w", [(100, 16), (300, 64), (1000, 128), (1366, 32), (3000, 256), (5000, 512)])
def test_sweep_points_vs_oracle(gpu_ctx, mode, length, w):
    n = 24 if length <= 1000 else 6
    check_against_oracle(gpu_ctx, dpgen.sweep_pairs(n, length, w, mode, 0.05, seed=length * 7 + w))


@pytest.mark.parametrize("mode", [dpgen.GLOBAL, dpgen.EXT, dpgen.EXT_RIGHT])
@pytest.mark.parametrize("length,w,div", [(10000, 512, 0.05), (20000, 512, 0.05), (20000, 64, 0.01), (10000, 256, 0.15),
                                          (3000, 128, 0.01), (3000, 128, 0.15), (300, 16, 0.15), (1000, 32, 0.01)])
def test_sweep_long_and_divergent_points_vs_oracle(gpu_ctx, mode, length, w, div):
    """The upper end of BASELINE configs[4] (10 and 20 kbp, int32 score mode, two-megabyte traceback per problem) and its
    1 % / 15 % divergence levels: every kswcpp_extz_t field, the CIGAR and the band-cell count against the oracle."""
    n = 3 if length >= 10000 else 8
    check_against_oracle(gpu_ctx, dpgen.sweep_pairs(n, length, w, mode, div, seed=length * 11 + w))


def test_illumina_like_end_extensions(gpu_ctx):
    """The dominant DP shape of the Illumina preset: ~50 bp read tail against a 1000 bp padded window, w=512."""
    rng = np.random.Generator(np.random.PCG64(31))
    pairs = []
    for i in range(200):
        t = rng.integers(0, 4, size=int(rng.integers(950, 1060)), dtype=np.uint8)
        q = dpgen.mutate(rng, t[:int(rng.integers(1, 90))], 0.03)
        if len(q) == 0:
            q = t[:1].copy()
        pairs.append((512, 200, dpgen.EXT if i % 2 else dpgen.EXT_RIGHT, q, t))
    check_against_oracle(gpu_ctx, pairs)


def test_edge_cases(gpu_ctx):
    e = np.zeros(0, np.uint8)
    one = np.array([2], np.uint8)
    allN = np.full(40, 4, np.uint8)
    t = np.arange(64, dtype=np.uint8) & 3
    pairs = [(10, -1, 0, e, one), (10, -1, 0, one, e), (10, -1, 0, one, one), (10, 200, dpgen.EXT, one, t),
             (20, -1, 0, allN, t[:40]), (512, 200, dpgen.EXT_RIGHT, t, t), (0, -1, 0, t[:20], t[:20]),
             (-1, -1, 0, t[:30], t[:50])]
    check_against_oracle(gpu_ctx, pairs)


def test_three_step_form_matches_batch(gpu_ctx):
    pairs = dpgen.random_pairs(200, seed=5)
    tasks, seq = api.pack_ksw_tasks(pairs)
    r1, c1 = gpu_ctx.ksw_batch(tasks, seq)
    gpu_ctx.ksw_upload(tasks, seq)
    ms = gpu_ctx.ksw_run()
    r2, c2 = gpu_ctx.ksw_download()
    assert ms > 0
    for k in FIELDS:
        assert np.array_equal(r1[k], r2[k])
    for i in range(len(pairs)):  # cigar slab order is allocation order (non-deterministic); compare per task
        a = c1[r1["cigar_off"][i]:r1["cigar_off"][i] + r1["n_cigar"][i]]
        b = c2[r2["cigar_off"][i]:r2["cigar_off"][i] + r2["n_cigar"][i]]
        assert np.array_equal(a, b)


def test_linearity_property_large_batch(gpu_ctx):
    """Size-independent property at bench scale: identical sequences score 2*len, CIGAR = one M run; duplicating a
    batch gives identical per-task results."""
    rng = np.random.Generator(np.random.PCG64(8))
    pairs = []
    for i in range(20000):
        t = rng.integers(0, 4, size=int(rng.integers(20, 200)), dtype=np.uint8)
        pairs.append((20, -1, 0, t, t))
    tasks, seq = api.pack_ksw_tasks(pairs)
    res, cig = gpu_ctx.ksw_batch(tasks, seq)
    assert np.array_equal(res["score"], 2 * tasks["qlen"])
    assert (res["n_cigar"] == 1).all()
    assert np.array_equal(cig[res["cigar_off"]], (tasks["qlen"].astype(np.uint32) << 4))


# ---- early-termination mode of extension tasks (what the alignment path runs; packed half2 kernel when it applies)
EXT_FIELDS = ["max", "max_q", "max_t", "n_cigar"]


def check_extension_only(ctx, pairs):
    """max / max_q / max_t / CIGAR must equal the reference's full computation; returns the cells the GPU processed."""
    tasks, seq = api.pack_ksw_tasks(pairs)
    ctx.ksw_set_extension_only(True)
    try:
        res, cig = ctx.ksw_batch(tasks, seq)
    finally:
        ctx.ksw_set_extension_only(False)
    assert (res["status"] == 0).all()
    full = 0
    for i, (w, zd, fl, q, t) in enumerate(pairs):
        exp, ecig, ecells = H.oracle_ksw(q, t, w, zd, fl)
        if exp["zdropped"] == 0 and exp["mqe"] > exp["max"]:
            continue  # reach_end back-trace: not an early-stop caller's case (never occurs: mqe <= max)
        for k in EXT_FIELDS:
            assert int(res[k][i]) == exp[k], (i, k, int(res[k][i]), exp[k], len(q), len(t), w, zd, fl)
        got = cig[res["cigar_off"][i]:res["cigar_off"][i] + res["n_cigar"][i]]
        assert np.array_equal(got, ecig), (i, got[:6], ecig[:6], len(q), len(t), w, zd, fl)
        full += ecells
    return int(res["cells"].sum()), full


def _ext_pairs(n, seed, qmax, tmin, tmax, err, w=512, zdrop=200, with_n=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    pairs = []
    for i in range(n):
        t = rng.integers(0, 4, size=int(rng.integers(tmin, tmax)), dtype=np.uint8)
        q = dpgen.mutate(rng, t[:int(rng.integers(1, qmax))], err)
        if len(q) == 0:
            q = t[:1].copy()
        if with_n and i % 3 == 0:
            q = q.copy()
            q[rng.integers(0, len(q), size=max(1, len(q) // 10))] = 4
            t = t.copy()
            t[rng.integers(0, min(len(t), 2 * len(q)), size=3)] = 4
        pairs.append((w, zdrop, dpgen.EXT if i % 2 else dpgen.EXT_RIGHT, q, t))
    return pairs


def test_extension_only_illumina_shapes(gpu_ctx):
    done, full = check_extension_only(gpu_ctx, _ext_pairs(400, 77, 150, 950, 1060, 0.03))
    assert done < full  # early termination really cuts rows


def test_extension_only_noisy_and_n(gpu_ctx):
    check_extension_only(gpu_ctx, _ext_pairs(300, 78, 200, 300, 1100, 0.15, with_n=True))
    check_extension_only(gpu_ctx, _ext_pairs(200, 79, 60, 20, 120, 0.3, with_n=True))


def test_extension_only_unrelated_and_tiny(gpu_ctx):
    rng = np.random.Generator(np.random.PCG64(80))
    pairs = []
    for i in range(200):
        q = rng.integers(0, 4, size=int(rng.integers(1, 120)), dtype=np.uint8)
        t = rng.integers(0, 4, size=int(rng.integers(1, 400)), dtype=np.uint8)
        pairs.append((512, 200 if i % 4 else -1, dpgen.EXT if i % 2 else dpgen.EXT_RIGHT, q, t))
    check_extension_only(gpu_ctx, pairs)


def test_extension_only_band_limited_falls_back(gpu_ctx):
    """w < qlen or a band that starts to limit: the packed/fast modes must hand over to the exact mode."""
    check_extension_only(gpu_ctx, _ext_pairs(120, 81, 300, 350, 700, 0.05, w=32))
    check_extension_only(gpu_ctx, _ext_pairs(60, 82, 400, 1500, 2500, 0.05, w=100))


def test_extension_only_long_int32_mode(gpu_ctx):
    check_extension_only(gpu_ctx, _ext_pairs(6, 83, 3000, 20000, 21000, 0.05))


@pytest.mark.parametrize("name", ["bwa_like", "swapped", "large", "early_return"])
def test_golden_reference_vectors_with_other_scores(name):
    """Non-default KswCppParam<5> (ParameterSetManager: Match / Mismatch / Gap / Extend / Gap2 / Extend2) against vectors
    written by the compiled reference: BWA-like costs, the q2 + e2 < q + e swap (whose row-0 H keeps the pre-swap
    q + e, kswcpp_core.h:338/247), scores beyond the 8-bit range of the packed mode, the early return. All fields and
    CIGARs, plus the early-stop mode of the extensions (max, position, CIGAR)."""
    g = np.load(os.path.join(H.GOLDEN, "ksw_golden_%s.npz" % name))
    d = {"ksw_calls": g["calls"].astype(np.int64), "ksw_seq": g["seq"], "ksw_cigar": g["cigar"]}
    sc = [int(x) for x in g["score"]]
    calls = list(H.split_ksw_dump(d))
    ctx = api.Context(0)
    p = api.preset("default")
    p.match, p.mismatch, p.gap, p.extend, p.gap2, p.extend2 = sc
    ctx.set_params(p)
    tasks, seq = api.pack_ksw_tasks([(f["w"], f["zdrop"], f["flag"], q, t) for f, q, t, c in calls])
    for ext_only in (False, True):
        ctx.ksw_set_extension_only(ext_only)
        res, cig = ctx.ksw_batch(tasks, seq)
        for i, (f, q, t, c) in enumerate(calls):
            is_ext = bool(f["flag"] & 0x40)
            if ext_only and not is_ext:
                continue
            for k in (["max", "max_q", "max_t", "n_cigar"] if ext_only else FIELDS):
                assert int(res[k][i]) == f[k], (name, ext_only, i, k, int(res[k][i]), f[k], f)
            got = cig[res["cigar_off"][i]:res["cigar_off"][i] + res["n_cigar"][i]]
            assert np.array_equal(got, c), (name, ext_only, i, f)
    ctx.close()


def test_random_score_sets_vs_oracle():
    """Random scoring parameters (the oracle is pinned to the reference for such sets in test_ksw_oracle.py): every field
    and CIGAR, and the early-stop mode of the extensions."""
    from test_ksw_oracle import random_score_sets
    for i, sc in enumerate(random_score_sets(14, 77)):
        pairs = dpgen.random_pairs(50, seed=3000 + i, lengths=(1, 2, 3, 8, 17, 33, 50, 100, 150, 300, 700))
        pairs += dpgen.sweep_pairs(4, 110, 512, dpgen.EXT, 0.03, seed=i) + dpgen.sweep_pairs(4, 110, 512, dpgen.EXT_RIGHT, 0.03, seed=50 + i)
        score = H.OracleScore(*sc)
        ctx = api.Context(0)
        p = api.preset("default")
        p.match, p.mismatch, p.gap, p.extend, p.gap2, p.extend2 = sc
        ctx.set_params(p)
        tasks, seq = api.pack_ksw_tasks(pairs)
        exp = [H.oracle_ksw(q, t, w, zd, fl, score) for w, zd, fl, q, t in pairs]
        for ext_only in (False, True):
            ctx.ksw_set_extension_only(ext_only)
            res, cig = ctx.ksw_batch(tasks, seq)
            for k, ((w, zd, fl, q, t), (e, ecig, _)) in enumerate(zip(pairs, exp)):
                if ext_only and not (fl & 0x40):
                    continue
                for name in (["max", "max_q", "max_t", "n_cigar"] if ext_only else FIELDS):
                    assert int(res[name][k]) == e[name], (sc, ext_only, k, name, int(res[name][k]), e[name], len(q), len(t), w, zd, fl)
                got = cig[res["cigar_off"][k]:res["cigar_off"][k] + res["n_cigar"][k]]
                assert np.array_equal(got, ecig), (sc, ext_only, k)
        ctx.close()
