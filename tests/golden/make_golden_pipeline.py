"""Generates the pipeline golden fixtures by running the compiled, UNMODIFIED reference (oracle/_ref/ref_dump):

  tests/golden/gold.{bwt,sa,pac,ann,amb}   index of a 3-contig 60 kbp synthetic genome (reference's own builder)
  tests/golden/gold_reads_short.txt / gold_reads_long.txt / gold_reads_pairs.txt (mates interleaved: 2k, 2k+1)
  tests/golden/gold_<preset>.sam            SAM of the reference's own writers (illumina, illuminapaired, pacbio)
  tests/golden/gold_<preset>.npz            per-stage dumps (segments, seeds, SoC pops, harmonized sets, DP calls,
                                            alignments, MappingQuality results, PairedReads results of consecutive
                                            reads) with srand(1000 + read index) before Harmonization::execute

Run in the build container (needs /root/reference -> `make -C oracle ref`).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import helpers as H  # noqa: E402
from ma_b200 import synth  # noqa: E402

SRAND = 1000
SAM_OPTION_RUNS = [("illumina", "gold_reads_short.txt", "x_soft", "Use M in CIGAR=false;Soft clip=true"),
                   ("illuminapaired", "gold_reads_pairs.txt", "nosec_soft",
                    "Omit Secondary Alignments=true;Omit Supplementary Alignments=true;Soft clip=true"),
                   ("pacbio", "gold_reads_long.txt", "x_nosupp", "Use M in CIGAR=false;Omit Supplementary Alignments=true")]
HEURISTIC_RUNS = [("illumina", "gold_reads_short.txt"), ("illuminapaired", "gold_reads_pairs.txt"),
                  ("pacbio", "gold_reads_long.txt")]


def main():
    g = synth.random_genome([30_000, 20_000, 10_000], 42)
    # a repeat: copy 400 bp of contig 1 into contig 2 (ambiguous seeds, multiple SoCs)
    g[1][5000:5400] = g[0][1000:1400]
    gt = os.path.join(H.GOLDEN, "gold_genome.txt")
    synth.write_genome_txt(gt, g)
    H.run_ref("index", gt, os.path.join(H.GOLDEN, "gold"))
    os.remove(gt)
    short, _, _, _ = synth.simulate_reads(g, 400, 150, 7, sub_rate=0.02, ins_rate=0.004, del_rate=0.004)
    short[3, 60] = 4          # an N inside a read
    short[4, :] = 4           # all-N read
    short[5, 0] = 4
    short[6, 149] = 4
    short[7] = np.random.Generator(np.random.PCG64(5)).integers(0, 4, 150)  # unalignable random read
    short[8, :75] = g[0][2000:2075]
    short[8, 75:] = g[2][3000:3075]  # chimeric read
    synth.write_reads_txt(os.path.join(H.GOLDEN, "gold_reads_short.txt"), short)
    long_, _, _, _ = synth.simulate_long_reads(g, 12, 2500, 8)
    synth.write_reads_txt(os.path.join(H.GOLDEN, "gold_reads_long.txt"), long_)
    m1, m2, _, _, _, _ = synth.simulate_pairs(g, 150, 150, 9, sub_rate=0.02, indel_rate=0.01, ins_mean=400, ins_sd=60)
    pairs = np.empty((300, 150), dtype=np.uint8)
    pairs[0::2], pairs[1::2] = m1, m2
    pairs[5] = np.random.Generator(np.random.PCG64(6)).integers(0, 4, 150)  # a mate that does not align
    pairs[8] = pairs[20]  # a mate from elsewhere (unpaired distance)
    pairs[12, :] = 4
    pairs[15] = synth.revcomp(pairs[15]) if hasattr(synth, "revcomp") else (3 - pairs[15][::-1])  # same strand pair
    synth.write_reads_txt(os.path.join(H.GOLDEN, "gold_reads_pairs.txt"), pairs)
    for preset, rf in [("illumina", "gold_reads_short.txt"), ("default", "gold_reads_short.txt"),
                       ("pacbio", "gold_reads_long.txt"), ("nanopore", "gold_reads_long.txt"),
                       ("illuminapaired", "gold_reads_pairs.txt")]:
        out = os.path.join(H.GOLDEN, "tmp.dump")
        H.run_ref("align", os.path.join(H.GOLDEN, "gold"), os.path.join(H.GOLDEN, rf), preset, out, SRAND)
        d = H.load_dump(out)
        os.remove(out)
        comp = {}
        for k, v in d.items():
            if k == "ksw_seq":
                comp[k] = v.astype(np.uint8)
            elif k in ("ksw_cigar",):
                comp[k] = v.astype(np.uint32)
            else:
                comp[k] = v
        np.savez_compressed(os.path.join(H.GOLDEN, "gold_%s.npz" % preset), **comp)
        if preset in ("illumina", "illuminapaired", "pacbio"):
            # SAM text written by the reference's own FileWriter / PairedFileWriter
            H.run_ref("sam", os.path.join(H.GOLDEN, "gold"), os.path.join(H.GOLDEN, rf), preset,
                      os.path.join(H.GOLDEN, "gold_%s.sam" % preset), SRAND)
        print(preset, {k: len(v) for k, v in d.items() if k.endswith("_off")})
    # SAM text with non-default writer options (MA_REF_SET in oracle/ref_dump.cpp sets them on the presetting)
    for preset, rf, tag, opts in SAM_OPTION_RUNS:
        os.environ["MA_REF_SET"] = opts
        try:
            H.run_ref("sam", os.path.join(H.GOLDEN, "gold"), os.path.join(H.GOLDEN, rf), preset,
                      os.path.join(H.GOLDEN, "gold_%s_%s.sam" % (preset, tag)), SRAND)
        finally:
            del os.environ["MA_REF_SET"]
    # the heuristics for large genomes (seeding drop-off binarySeeding.cpp:172-175, SoC minimal length
    # stripOfConsideration.cpp:21-23) are off for a genome below "Minimum Genome Size for Heuristics" (10 M): these
    # sets switch them on for the small golden genome, which is how BASELINE's 100 Mbp configuration runs
    for preset, rf in HEURISTIC_RUNS:
        out = os.path.join(H.GOLDEN, "tmp.dump")
        os.environ["MA_REF_MIN_GENOME_SIZE"] = "0"
        try:
            H.run_ref("align", os.path.join(H.GOLDEN, "gold"), os.path.join(H.GOLDEN, rf), preset, out, SRAND)
        finally:
            del os.environ["MA_REF_MIN_GENOME_SIZE"]
        d = H.load_dump(out)
        os.remove(out)
        comp = {k: (v.astype(np.uint8) if k == "ksw_seq" else v.astype(np.uint32) if k == "ksw_cigar" else v)
                for k, v in d.items()}
        np.savez_compressed(os.path.join(H.GOLDEN, "gold_%s_heur.npz" % preset), **comp)
        print(preset, "heuristics on", {k: len(v) for k, v in d.items() if k.endswith("_off")})


if __name__ == "__main__":
    main()
