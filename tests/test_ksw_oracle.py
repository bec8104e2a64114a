"""CPU: the DP oracle (oracle/ksw_oracle.cpp) against golden vectors produced by the compiled reference."""
import os

import numpy as np
import pytest

import dpgen
import helpers as H


def golden_calls():
    g = np.load(os.path.join(H.GOLDEN, "ksw_golden.npz"))
    d = {"ksw_calls": g["calls"].astype(np.int64), "ksw_seq": g["seq"], "ksw_cigar": g["cigar"]}
    return list(H.split_ksw_dump(d))


def test_oracle_matches_reference_golden():
    calls = golden_calls()
    assert len(calls) > 500
    for f, q, t, c in calls:
        res, cig, _ = H.oracle_ksw(q, t, f["w"], f["zdrop"], f["flag"])
        for k, v in res.items():
            assert v == f[k], (k, f)
        assert np.array_equal(cig, c)


def test_lane_blocked_argmax_witnesses_are_covered():
    """the reference's arg-max is per SSE lane and records the block base, not the true column, so
    some extensions end in an I/D run (impossible for an exact arg-max). The golden set must contain such calls and
    the oracle must reproduce them."""
    n = 0
    for f, q, t, c in golden_calls():
        if f["flag"] == dpgen.EXT and len(c) and (int(c[-1]) & 15) != 0 and not f["reach_end"]:
            res, cig, _ = H.oracle_ksw(q, t, f["w"], f["zdrop"], f["flag"])
            assert (res["max_q"], res["max_t"]) == (f["max_q"], f["max_t"]) and np.array_equal(cig, c)
            n += 1
    assert n >= 3


def test_empty_inputs():
    res, cig, cells = H.oracle_ksw(np.zeros(0, np.uint8), np.array([1, 2], np.uint8), 10, -1, 0)
    assert res["n_cigar"] == 0 and res["max"] == 0 and res["max_q"] == -1 and cells == 0


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_oracle_matches_live_reference(tmp_path):
    pairs = dpgen.random_pairs(300, seed=99)
    H.write_pairs(str(tmp_path / "p.txt"), pairs)
    H.run_ref("ksw", tmp_path / "p.txt", tmp_path / "k.dump")
    for f, q, t, c in H.split_ksw_dump(H.load_dump(str(tmp_path / "k.dump"))):
        res, cig, _ = H.oracle_ksw(q, t, f["w"], f["zdrop"], f["flag"])
        assert all(res[k] == f[k] for k in res) and np.array_equal(cig, c)


SCORE_SETS = ["bwa_like", "swapped", "large", "early_return"]


def golden_calls_for(name):
    g = np.load(os.path.join(H.GOLDEN, "ksw_golden_%s.npz" % name))
    d = {"ksw_calls": g["calls"].astype(np.int64), "ksw_seq": g["seq"], "ksw_cigar": g["cigar"]}
    return [int(x) for x in g["score"]], list(H.split_ksw_dump(d))


@pytest.mark.parametrize("name", SCORE_SETS)
def test_oracle_matches_reference_golden_with_other_scores(name):
    """Non-default KswCppParam<5> (tests/golden/make_golden_ksw.py SCORE_SETS, written by the compiled reference):
    BWA-like one-piece costs, the q2 + e2 < q + e swap, scores beyond the 8-bit difference range, the early return."""
    sc, calls = golden_calls_for(name)
    score = H.OracleScore(*sc)
    assert len(calls) > 250
    for f, q, t, c in calls:
        res, cig, _ = H.oracle_ksw(q, t, f["w"], f["zdrop"], f["flag"], score)
        for k, v in res.items():
            assert v == f[k], (name, k, v, f)
        assert np.array_equal(cig, c)


def random_score_sets(n, seed):
    rng = np.random.default_rng(seed)
    return [(int(rng.integers(1, 7)), int(rng.integers(1, 11)), int(rng.integers(0, 13)), int(rng.integers(1, 5)),
             int(rng.integers(0, 41)), int(rng.integers(1, 4))) for _ in range(n)]


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_oracle_matches_live_reference_random_scores(tmp_path):
    """Random KswCppParam<5> sets (gap 0, e == e2, q2 < q, match up to 6 ...) against the reference run here."""
    for i, sc in enumerate(random_score_sets(10, 5)):
        pairs = dpgen.random_pairs(40, seed=1000 + i, lengths=(1, 2, 3, 8, 17, 33, 50, 100, 150, 300))
        H.write_pairs(str(tmp_path / "p.txt"), pairs)
        H.run_ref("ksw", tmp_path / "p.txt", tmp_path / "k.dump", *sc)
        score = H.OracleScore(*sc)
        for f, q, t, c in H.split_ksw_dump(H.load_dump(str(tmp_path / "k.dump"))):
            res, cig, _ = H.oracle_ksw(q, t, f["w"], f["zdrop"], f["flag"], score)
            assert all(res[k] == f[k] for k in res) and np.array_equal(cig, c), (sc, f)
