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


if __name__ == "__main__":
    main()
