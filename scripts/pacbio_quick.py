"""Config-4-like probe: PacBio preset, long reads with 12 % error on the 100 Mbp synthetic genome."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ma_b200 import api, synth
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
genome = synth.random_genome([10_000_000] * 10, 2)
reads, *_ = synth.simulate_long_reads(genome, n_reads, L, 4)
ctx = api.Context(0, "pacbio")
lens = [len(c) for c in genome]
ctx.index_build(np.concatenate(genome), np.cumsum([0] + lens[:-1]), lens)
data, off = api.pack_reads(reads)
ctx.align_upload(data, off)
for it in range(2):
    t = time.time()
    st = ctx.align_run()
    dt = time.time() - t
    print(json.dumps({"reads": n_reads, "len": L, "wall_s": round(dt, 3), "reads_per_s": round(n_reads / dt, 1),
                      "Mbp_per_s": round(n_reads * L / dt / 1e6, 2),
                      **{k: (round(v, 1) if isinstance(v, float) else v) for k, v in st.items()}}))
info, alns, runs = ctx.download_alignments()
print("aligned reads", int((info["n_sets"] > 0).sum()), "alignments", len(alns))
