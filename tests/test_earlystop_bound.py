"""CPU: the early-termination bound used by the DP kernel for extension problems is checked against
the oracle: after the bound fires, NO later anti-diagonal of the reference computation raises ez.max (so max, max_q,
max_t and the CIGAR from that position are final)."""
import ctypes
import os

import numpy as np

import dpgen
import helpers as H


def check(q, t, w, zd, fl):
    lib = H.oracle_lib()
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    sr, vi, rows = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    lib.ma_oracle_ksw_earlystop_check(len(q), q.ctypes.data_as(ctypes.c_void_p), len(t),
                                      t.ctypes.data_as(ctypes.c_void_p), ctypes.byref(H.DEFAULT_SCORE), w, zd, fl,
                                      ctypes.byref(sr), ctypes.byref(vi), ctypes.byref(rows))
    return sr.value, vi.value, rows.value


def test_bound_never_violated_on_reference_dp_calls():
    """Every extension call the reference issued for the golden Illumina + PacBio reads."""
    n = saved = total = 0
    for preset in ("illumina", "pacbio"):
        g = np.load(os.path.join(H.GOLDEN, "gold_%s.npz" % preset))
        d = {"ksw_calls": g["ksw_calls"].astype(np.int64), "ksw_seq": g["ksw_seq"], "ksw_cigar": g["ksw_cigar"]}
        for f, q, t, c in H.split_ksw_dump(d):
            if f["flag"] & 0x40:
                sr, vi, rows = check(q, t, f["w"], f["zdrop"], f["flag"])
                assert vi == 0
                n += 1
                total += rows
                saved += rows - (sr + 1 if sr >= 0 else rows)
    assert n > 500 and saved > total // 2  # it also has to be worth it


def test_bound_never_violated_on_adversarial_cases():
    rng = np.random.Generator(np.random.PCG64(2026))
    for it in range(3000):
        kind = it % 6
        ql = int(rng.integers(1, 120)) if kind < 5 else int(rng.integers(200, 900))
        tl = int(rng.integers(ql, ql + 700))
        t = rng.integers(0, 2 if kind == 3 else 4, size=tl, dtype=np.uint8)
        if kind == 1:  # tandem repeat
            t = np.resize(rng.integers(0, 4, size=int(rng.integers(1, 12)), dtype=np.uint8), tl)
        q = t[:ql].copy()
        if kind == 2:  # a long deletion pays off: prefix + far-away copy
            cut = int(rng.integers(0, ql)) + 1
            far = int(rng.integers(0, max(1, tl - ql)))
            q = np.concatenate([t[:cut], t[far:far + ql - cut]])[:ql]
        if kind == 4:
            q = rng.integers(0, 4, size=ql, dtype=np.uint8)
        m = rng.random(len(q)) < rng.choice([0, 0.02, 0.1, 0.3])
        q = np.where(m, (q + 1) & 3, q).astype(np.uint8)
        if rng.random() < 0.1:
            q[rng.integers(0, len(q))] = 4
        w = int(rng.choice([16, 64, 512, 512]))
        zd = int(rng.choice([200, 20, 1000]))
        fl = int(rng.choice([dpgen.EXT, dpgen.EXT_RIGHT]))
        sr, vi, rows = check(q, t, w, zd, fl)
        assert vi == 0, (kind, ql, tl, w, zd, fl, sr, rows)
