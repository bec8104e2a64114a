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

Run in the build container after make_golden_pipeline.py (needs /root/reference -> `make -C oracle ref`).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
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


if __name__ == "__main__":
    main()
