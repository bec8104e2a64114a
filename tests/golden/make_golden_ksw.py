"""Generates tests/golden/ksw_golden.npz by running the compiled, UNMODIFIED reference (oracle/_ref/ref_dump ksw).

Run in the build container (needs /root/reference -> `make -C oracle ref`):  python tests/golden/make_golden_ksw.py
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import dpgen  # noqa: E402
import helpers as H  # noqa: E402


SCORE_SETS = {"bwa_like": (1, 4, 6, 1, 24, 1), "swapped": (2, 4, 4, 2, 3, 1), "large": (5, 9, 12, 3, 40, 2),
              "early_return": (2, 20, 4, 2, 24, 1)}


def main():
    pairs = dpgen.random_pairs(500, seed=20261017)
    # both score widths around the int16/int32 switch (max(qlen,tlen) = 1365 | 1366), SURVEY.md A-3
    for L in (1364, 1365, 1366, 1367):
        pairs += dpgen.sweep_pairs(2, L, 64, dpgen.EXT, 0.05, seed=L)
        pairs += dpgen.sweep_pairs(2, L, 64, dpgen.GLOBAL, 0.05, seed=L + 7)
    # A-1 lane-blocked arg-max witness: 25 identical bases, then random
    rng = np.random.Generator(np.random.PCG64(1))
    t = rng.integers(0, 4, size=300, dtype=np.uint8)
    q = np.concatenate([t[:25], rng.integers(0, 4, size=40, dtype=np.uint8)])
    pairs.append((512, 200, dpgen.EXT, q, t))
    pairs.append((512, 200, dpgen.EXT_RIGHT, q, t))
    with tempfile.TemporaryDirectory() as d:
        H.write_pairs(os.path.join(d, "p.txt"), pairs)
        H.run_ref("ksw", os.path.join(d, "p.txt"), os.path.join(d, "k.dump"))
        dump = H.load_dump(os.path.join(d, "k.dump"))
    np.savez_compressed(os.path.join(H.GOLDEN, "ksw_golden.npz"), calls=dump["ksw_calls"].astype(np.int32),
                        seq=dump["ksw_seq"].astype(np.uint8), cigar=dump["ksw_cigar"].astype(np.uint32))
    print("wrote", len(pairs), "calls")
    # non-default scoring (match mismatch gap extend gap2 extend2): one-piece-like costs, the q2 + e2 < q + e swap
    # (kswcpp_core.h:367-375), scores too large for 8-bit difference arithmetic without wrap-around, and the early
    # return of -min_sc > 2 (q + e); see SCORE_SETS in tests/test_ksw_oracle.py
    for name, sc in SCORE_SETS.items():
        pairs = dpgen.random_pairs(260, seed=777 + sum(sc), lengths=(1, 2, 3, 8, 17, 33, 50, 100, 150, 300, 700, 1400))
        pairs += dpgen.sweep_pairs(6, 120, 512, dpgen.EXT, 0.03, seed=5) + dpgen.sweep_pairs(6, 120, 512, dpgen.EXT_RIGHT, 0.03, seed=6)
        with tempfile.TemporaryDirectory() as d:
            H.write_pairs(os.path.join(d, "p.txt"), pairs)
            H.run_ref("ksw", os.path.join(d, "p.txt"), os.path.join(d, "k.dump"), *sc)
            dump = H.load_dump(os.path.join(d, "k.dump"))
        np.savez_compressed(os.path.join(H.GOLDEN, "ksw_golden_%s.npz" % name), calls=dump["ksw_calls"].astype(np.int32),
                            seq=dump["ksw_seq"].astype(np.uint8), cigar=dump["ksw_cigar"].astype(np.uint32),
                            score=np.array(sc, dtype=np.int32))
        print(name, sc, len(pairs), "calls")


if __name__ == "__main__":
    main()
