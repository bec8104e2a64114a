"""Pins the human-sized configurations (BASELINE.json configs[2] / configs[3]) to the UNMODIFIED reference: builds the
index of the 31 x 100 Mbp synthetic genome (seed 3; 3.1 Gbp, 6.2 G BWT symbols) with the reference's own builder
(FMIndex(pPack) -> bwtLarge, fMIndex.cpp:316-391; about an hour and a half on one core), stores the SHA-1 of the
index arrays, then runs the reference's modules over the first 3 000 reads of 100 000 simulated pairs
(Illumina_Paired preset, srand(1000 + read index)) and over 200 simulated 10 kbp PacBio reads (PacBio preset) and
stores the SHA-1 of every stage's dump in tests/golden/human_size_sha1.json.
tests/test_human_size_gpu.py compares the GPU index builder and the device path with these hashes.

Run in the build container (needs /root/reference -> `make -C oracle ref`):
    python tests/golden/make_golden_human_size.py [work_dir]        (work_dir keeps genome text + index, ~8 GB)
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import helpers as H  # noqa: E402
import pipeline_common as PC  # noqa: E402
from ma_b200 import index as IX  # noqa: E402
from ma_b200 import synth  # noqa: E402

N_CONTIGS, CONTIG_LEN, GENOME_SEED = 31, 100_000_000, 3
N_READS, N_PAIRS_SIM, READ_SEED = 3000, 100_000, 3
N_LONG, LONG_LEN, LONG_SEED = 200, 10000, 4


def sha1(a):
    return hashlib.sha1(np.ascontiguousarray(a)).hexdigest()


def sha1_i64(a):
    return hashlib.sha1(np.ascontiguousarray(np.asarray(a, dtype=np.int64)).tobytes()).hexdigest()


def main():
    d = sys.argv[1] if len(sys.argv) > 1 else "/tmp/human_ref"
    os.makedirs(d, exist_ok=True)
    t0 = time.time()
    genome = synth.random_genome([CONTIG_LEN] * N_CONTIGS, GENOME_SEED)
    if not os.path.exists(os.path.join(d, "g.sa")):
        synth.write_genome_txt(os.path.join(d, "g.txt"), genome)
        print("genome text written %.0f s" % (time.time() - t0), flush=True)
        H.run_ref("index", os.path.join(d, "g.txt"), os.path.join(d, "g"))
        os.remove(os.path.join(d, "g.txt"))
    t_index = time.time() - t0
    print("reference index built %.0f s" % t_index, flush=True)
    ix = IX.load_index(os.path.join(d, "g"))
    out = {"n_contigs": N_CONTIGS, "contig_len": CONTIG_LEN, "genome_seed": GENOME_SEED,
           "reference_index_build_s": round(t_index),
           "index": {"primary": int(ix.primary), "L2": [int(x) for x in ix.L2], "n_words": int(ix.bwt.size),
                     "n_sa": int(ix.sa.size),
                     "sha1": {"bwt": sha1(ix.bwt), "sa": sha1(ix.sa), "pac": sha1(ix.pac[:(ix.fwd_len + 3) // 4])}}}
    del ix
    print(json.dumps(out), flush=True)
    m1, m2, *_ = synth.simulate_pairs(genome, N_PAIRS_SIM, 150, READ_SEED)
    reads = np.empty((N_READS, 150), dtype=np.uint8)
    reads[0::2], reads[1::2] = m1[:N_READS // 2], m2[:N_READS // 2]
    synth.write_reads_txt(os.path.join(d, "r.txt"), reads)
    H.run_ref("align", os.path.join(d, "g"), os.path.join(d, "r.txt"), "illuminapaired", os.path.join(d, "r.dump"),
              PC.SRAND)
    r = H.load_dump(os.path.join(d, "r.dump"))
    H.run_ref("sam", os.path.join(d, "g"), os.path.join(d, "r.txt"), "illuminapaired", os.path.join(d, "r.sam"), PC.SRAND)
    sam = open(os.path.join(d, "r.sam"), "rb").read()
    out.update({"n_reads": N_READS, "n_pairs_simulated": N_PAIRS_SIM, "read_seed": READ_SEED, "srand_base": PC.SRAND,
                "sha1": {k: sha1_i64(r[k]) for k in PC.STAGE_KEYS + ["mq_off", "mq", "pr_off", "pr"]},
                "sam_sha1": hashlib.sha1(sam).hexdigest(), "sam_lines": sam.count(b"\n")})
    print("illumina sample done %.0f s" % (time.time() - t0), flush=True)
    long_reads, *_ = synth.simulate_long_reads(genome, N_LONG, LONG_LEN, LONG_SEED)
    synth.write_reads_txt(os.path.join(d, "l.txt"), long_reads)
    H.run_ref("align", os.path.join(d, "g"), os.path.join(d, "l.txt"), "pacbio", os.path.join(d, "l.dump"), PC.SRAND)
    rl = H.load_dump(os.path.join(d, "l.dump"))
    out["pacbio"] = {"n_reads": N_LONG, "read_len": LONG_LEN, "seed": LONG_SEED,
                     "sha1": {k: sha1_i64(rl[k]) for k in PC.STAGE_KEYS + ["mq_off", "mq"]}}
    with open(os.path.join(H.GOLDEN, "human_size_sha1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out), flush=True)
    print("done %.0f s" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
