"""Pins the BASELINE.json configs[1] workload to the UNMODIFIED reference: builds the index of the 100 Mbp synthetic
genome with the reference's own builder (bwtLarge, ~2 min), runs its modules over the first 3 000 reads of the
1 M simulated pairs (Illumina_Paired preset, srand(1000 + read index)) and over 200 simulated 10 kbp PacBio reads
(PacBio preset) and stores the SHA-1 of every stage's dump in
tests/golden/full_size_sample_sha1.json. tests/test_full_size_gpu.py compares the device path with these hashes.

Run in the build container (needs /root/reference -> `make -C oracle ref`):  python tests/golden/make_golden_full_size.py
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import helpers as H  # noqa: E402
import pipeline_common as PC  # noqa: E402
from ma_b200 import synth  # noqa: E402

N_READS = 3000
N_LARGE = 100_000  # second, larger sample of the same batch: hashes only (key "large_sample")


def sha1(a):
    return hashlib.sha1(np.ascontiguousarray(np.asarray(a, dtype=np.int64)).tobytes()).hexdigest()


def main():
    genome = synth.random_genome([10_000_000] * 10, 2)
    m1, m2, *_ = synth.simulate_pairs(genome, 1_000_000, 150, 2017)
    reads = np.empty((N_READS, 150), dtype=np.uint8)
    reads[0::2], reads[1::2] = m1[:N_READS // 2], m2[:N_READS // 2]
    with tempfile.TemporaryDirectory() as d:
        synth.write_genome_txt(os.path.join(d, "g.txt"), genome)
        H.run_ref("index", os.path.join(d, "g.txt"), os.path.join(d, "g"))
        synth.write_reads_txt(os.path.join(d, "r.txt"), reads)
        H.run_ref("align", os.path.join(d, "g"), os.path.join(d, "r.txt"), "illuminapaired", os.path.join(d, "r.dump"),
                  PC.SRAND)
        r = H.load_dump(os.path.join(d, "r.dump"))
        # the SAM file of the reference's PairedFileWriter for the same reads (named r<i>)
        H.run_ref("sam", os.path.join(d, "g"), os.path.join(d, "r.txt"), "illuminapaired", os.path.join(d, "r.sam"),
                  PC.SRAND)
        sam = open(os.path.join(d, "r.sam"), "rb").read()
        # configs[3]-shaped: 200 simulated PacBio reads of 10 kbp (12 % error) on the same index, PacBio preset
        long_reads, *_ = synth.simulate_long_reads(genome, 200, 10000, 4)
        synth.write_reads_txt(os.path.join(d, "l.txt"), long_reads)
        H.run_ref("align", os.path.join(d, "g"), os.path.join(d, "l.txt"), "pacbio", os.path.join(d, "l.dump"), PC.SRAND)
        rl = H.load_dump(os.path.join(d, "l.dump"))
        # the first 100 000 reads of the batch (5 % of it): the reference is single-threaded here, about a minute
        big = np.empty((N_LARGE, 150), dtype=np.uint8)
        big[0::2], big[1::2] = m1[:N_LARGE // 2], m2[:N_LARGE // 2]
        synth.write_reads_txt(os.path.join(d, "b.txt"), big)
        H.run_ref("align", os.path.join(d, "g"), os.path.join(d, "b.txt"), "illuminapaired", os.path.join(d, "b.dump"),
                  PC.SRAND)
        rb = H.load_dump(os.path.join(d, "b.dump"))
    out = {"n_reads": N_READS, "srand_base": PC.SRAND,
           "sha1": {k: sha1(r[k]) for k in PC.STAGE_KEYS + ["mq_off", "mq", "pr_off", "pr"]},
           "sam_sha1": hashlib.sha1(sam).hexdigest(), "sam_lines": sam.count(b"\n"),
           "large_sample": {"n_reads": N_LARGE,
                            "sha1": {k: sha1(rb[k]) for k in PC.STAGE_KEYS + ["mq_off", "mq", "pr_off", "pr"]}},
           "pacbio": {"n_reads": 200, "read_len": 10000, "seed": 4,
                      "sha1": {k: sha1(rl[k]) for k in PC.STAGE_KEYS + ["mq_off", "mq"]}}}
    with open(os.path.join(H.GOLDEN, "full_size_sample_sha1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main()
