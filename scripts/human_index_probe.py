"""GPU box probe for the human-sized configuration (BASELINE configs[2]/[3]): builds the index of the 31 x 100 Mbp
synthetic genome (seed 3) with the bucketed GPU builder, hashes the index arrays (to be compared with the reference's
bwtLarge output, tests/golden/human_size_sha1.json), and aligns a sample of simulated pairs against it.
Usage: python scripts/human_index_probe.py [n_contigs] [n_pairs]  ->  gpurun_out/human_index_probe.json"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ma_b200 import api, synth  # noqa: E402


def sha1(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    n_contigs = int(sys.argv[1]) if len(sys.argv) > 1 else 31
    n_pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
    out = {"n_contigs": n_contigs, "contig_len": 100_000_000, "seed": 3}
    t = time.time()
    genome = synth.random_genome([100_000_000] * n_contigs, 3)
    lens = np.array([len(c) for c in genome], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    fwd = np.concatenate(genome)
    out["gen_s"] = time.time() - t
    ctx = api.Context(0, "illumina_paired")
    p = api.preset("illumina_paired")
    p.srand_base = 1000
    ctx.set_params(p)
    t = time.time()
    ctx.index_build(fwd, starts, lens)
    out["index_build_s"] = time.time() - t
    print("index built in %.1f s" % out["index_build_s"], flush=True)
    t = time.time()
    ix = ctx.index_download()
    out["download_s"] = time.time() - t
    out["primary"], out["L2"] = int(ix.primary), [int(x) for x in ix.L2]
    out["sha1"] = {"bwt": sha1(ix.bwt), "sa": sha1(ix.sa), "pac": sha1(ix.pac[:(len(fwd) + 3) // 4])}
    out["n_words"], out["n_sa"] = int(ix.bwt.size), int(ix.sa.size)
    print(json.dumps(out), flush=True)
    del ix
    m1, m2, cid, pos, flen, rev = synth.simulate_pairs(genome, n_pairs, 150, 3)
    reads = np.empty((2 * n_pairs, 150), dtype=np.uint8)
    reads[0::2], reads[1::2] = m1, m2
    data, off = api.pack_reads(reads)
    ctx.align_upload(data, off)
    for it in range(2):
        st = ctx.align_run(api.STAGE_MAPQ)
    info, alns, runs = ctx.download_alignments()
    prim = alns[alns["rank_mq"] == 0]
    r = prim["read"].astype(np.int64)
    pair, mate = r // 2, r % 2
    g0 = starts[cid[pair]] + pos[pair]
    is_left = (mate == 0) != rev[pair]
    expect = np.where(is_left, g0, 2 * len(fwd) - g0 - flen[pair])
    got = prim["begin_ref"] - prim["begin_q"]
    out["align"] = {"n_reads": 2 * n_pairs, "primaries": int(len(prim)),
                    "within_10": float((np.abs(got - expect) <= 10).mean()),
                    "stats": {k: (float(v) if isinstance(v, float) else int(v)) for k, v in st.items()}}
    out["gather_GBs"] = {"index_size": ctx.gather_probe(int(out["n_words"]) * 4)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "human_index_probe.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
