"""Generates the FASTA / FASTQ golden fixtures (reader side) with the compiled, UNMODIFIED
reference (oracle/_ref/ref_dump):

  tests/golden/gold_reads.fq              20 pairs of gold_reads_pairs.txt as FASTQ with the irregularities the
                                          reference's FileReader accepts: descriptions after the name, lower case,
                                          CR LF line ends, records wrapped over two lines, blank lines between records,
                                          an all-N mate. (No quality string shorter than its sequence: the
                                          reference then returns uninitialised bytes for the rest.)
  tests/golden/gold_reads.fa              the first 20 of those reads as FASTA wrapped at 60 columns, CR LF on every
                                          third record, IUPAC codes, blank lines
  tests/golden/gold_reads_<fq|fa>.parsed  name / sequence / quality of every read as FileReader::execute returns them
  tests/golden/gold_fq_illuminapaired.sam FileReader -> alignment path -> PairedFileWriter (srand(1000 + read index))
  tests/golden/gold_fa_illumina.sam       FileReader -> alignment path -> FileWriter

  tests/golden/gold_reads_inv.fa          reads with short inverted segments that hold no seed (a substitution every
                                          12 bases inside the inverted part), both strands, 250 and 1000 bases
  tests/golden/gold_inv_default_z20.sam   ... -> MappingQuality -> SmallInversions -> FileWriter, Default presetting
                                          with "Z Drop Inversions" 20 (14 inversion records)
  tests/golden/gold_inv_illumina.sam      the same with the Illumina presetting and the default threshold of 100

Run in the build container after make_golden_pipeline.py (needs /root/reference -> `make -C oracle ref`).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import helpers as H  # noqa: E402

SRAND = 1000


def main():
    lines = [l.strip() for l in open(os.path.join(H.GOLDEN, "gold_reads_pairs.txt")) if l.strip()]
    rng = np.random.Generator(np.random.PCG64(11))

    def qual(n):
        return "".join(chr(int(c)) for c in rng.integers(35, 74, n))

    with open(os.path.join(H.GOLDEN, "gold_reads.fq"), "wb") as f:
        for i in range(40):
            pair, mate = divmod(i, 2)
            s = lines[i]
            q = qual(len(s))
            if i in (2, 7):
                s = s.lower()
            if s.strip("N") == "":
                s = s.lower()
            eol = "\r\n" if i in (3, 10) else "\n"
            name = "@pair%d/%d some description" % (pair, mate + 1)
            if i in (4, 10):  # record wrapped over two lines
                rec = [name, s[:70], s[70:], "+", q[:70], q[70:]]
            else:
                rec = [name, s, "+", q]
            f.write((eol.join(rec) + eol).encode())
            if i == 8:
                f.write(b"\n")
    with open(os.path.join(H.GOLDEN, "gold_reads.fa"), "wb") as f:
        for i in range(20):
            s = lines[i]
            if s.strip("N") == "":
                continue
            if i in (1, 9):
                s = "R" + s[1:]
            if i == 5:
                s = s[:10] + "R" + s[11:]
            eol = "\r\n" if i % 3 == 0 else "\n"
            rec = [">read%d desc" % i] + [s[k:k + 60] for k in range(0, len(s), 60)]
            f.write((eol.join(rec) + eol).encode())
            if i in (0, 5):
                f.write(b"\n")
    for name in ("fq", "fa"):
        out = H.run_ref("reads", os.path.join(H.GOLDEN, "gold_reads.%s" % name))
        open(os.path.join(H.GOLDEN, "gold_reads_%s.parsed" % name), "w").write(out)
    H.run_ref("sam", os.path.join(H.GOLDEN, "gold"), os.path.join(H.GOLDEN, "gold_reads.fq"), "illuminapaired",
              os.path.join(H.GOLDEN, "gold_fq_illuminapaired.sam"), SRAND)
    H.run_ref("sam", os.path.join(H.GOLDEN, "gold"), os.path.join(H.GOLDEN, "gold_reads.fa"), "illumina",
              os.path.join(H.GOLDEN, "gold_fa_illumina.sam"), SRAND)


def inversions():
    from ma_b200 import index as maindex
    ix = maindex.load_index(os.path.join(H.GOLDEN, "gold"))
    g = np.asarray(ix.forward_codes() if callable(ix.forward_codes) else ix.forward_codes, dtype=np.uint8)
    rng = np.random.Generator(np.random.PCG64(22))

    def rc(x):
        return (3 - x[::-1]).astype(np.uint8)

    def make(L, inv_len, contig, revstrand, sub=0.0, n_inv=1):
        s0, cl = int(ix.contig_start[contig]), int(ix.contig_len[contig])
        p = int(rng.integers(s0 + 10, s0 + cl - L - 10))
        r = g[p:p + L].copy()
        for k in range(n_inv):
            a = (k + 1) * L // (n_inv + 1) - inv_len // 2
            seg = rc(r[a:a + inv_len])
            for j in range(int(rng.integers(6, 10)), inv_len, 12):  # break every seed inside the inverted part
                seg[j] = (seg[j] + 1 + int(rng.integers(0, 3))) & 3
            r[a:a + inv_len] = seg
        if sub > 0:
            m = rng.random(L) < sub
            r[m] = (r[m] + rng.integers(1, 4, m.sum())) & 3
        return rc(r) if revstrand else r

    reads = [make(250, [36, 44, 52, 60, 70, 80][k % 6], k % 3, k % 2 == 1, 0.0 if k < 6 else 0.01) for k in range(12)]
    reads += [make(1000, [40, 50, 60, 70, 80, 90][k], k % 3, k % 2 == 0, 0.01, n_inv=1 + k % 3) for k in range(6)]
    reads.append(g[1000:1250].copy())  # no inversion
    reads.append(rng.integers(0, 4, 250).astype(np.uint8))  # does not align
    fa = os.path.join(H.GOLDEN, "gold_reads_inv.fa")
    with open(fa, "w") as f:
        for i, r in enumerate(reads):
            f.write(">inv%d\n%s\n" % (i, "".join("ACGT"[c] for c in r)))
    H.run_ref("sam", os.path.join(H.GOLDEN, "gold"), fa, "default", os.path.join(H.GOLDEN, "gold_inv_default_z20.sam"),
              SRAND, "inv=20")
    H.run_ref("sam", os.path.join(H.GOLDEN, "gold"), fa, "illumina", os.path.join(H.GOLDEN, "gold_inv_illumina.sam"),
              SRAND, "inv")
    # pairs (mates of 250 bases, fragments of ~500) in which one or both mates carry a seedless inversion:
    # MappingQuality -> SmallInversions per mate -> PairedReads -> PairedFileWriter (export.cpp:176-184)
    pairs = []
    for k in range(16):
        contig = k % 3
        s0, cl = int(ix.contig_start[contig]), int(ix.contig_len[contig])
        p = int(rng.integers(s0 + 10, s0 + cl - 700))
        frag = g[p:p + 520].copy()
        m1, m2 = frag[:250].copy(), rc(frag[-250:])
        for m, has in ((m1, k % 4 != 3), (m2, k % 2 == 0)):
            if has:
                n = [44, 52, 60, 70][k % 4]
                a = 125 - n // 2
                seg = rc(m[a:a + n])
                for j in range(int(rng.integers(6, 10)), n, 12):
                    seg[j] = (seg[j] + 1 + int(rng.integers(0, 3))) & 3
                m[a:a + n] = seg
        if k % 5 == 4:
            m1, m2 = m2, m1
        pairs += [m1, m2]
    fq = os.path.join(H.GOLDEN, "gold_reads_inv_pairs.fa")
    with open(fq, "w") as f:
        for i, r in enumerate(pairs):
            f.write(">invpair%d/%d\n%s\n" % (i // 2, i % 2 + 1, "".join("ACGT"[c] for c in r)))
    H.run_ref("sam", os.path.join(H.GOLDEN, "gold"), fq, "illuminapaired",
              os.path.join(H.GOLDEN, "gold_inv_illuminapaired_z20.sam"), SRAND, "inv=20")


if __name__ == "__main__":
    main()
    inversions()
