#!/usr/bin/env python
"""bench.py — aligned reads/s of the B200-native MA hot path (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|dp_sweep]
                  [--pairs P] [--genome-mbp G]

--config 2: configs[2] — 31 x 100 Mbp genome (seed 3; index replicated per GPU, built on the GPU), ONE fixed set of
10 M 2x150 pairs split into contiguous shards over the N ranks (strong scaling), sub-batches of 1 M pairs.
--config 3: configs[3] — the same genome, ONE fixed set of 100 k simulated 10 kbp PacBio reads (12 % error), PacBio
preset, sharded the same way.  --config dp_sweep: configs[4] — kswcpp problems only (scripts/dp_sweep_bench.py).

Default workload (--config 1, N = 1): BASELINE.json configs[1] — synthetic 100 Mbp genome (10 contigs x 10 Mbp, seed 2), 1 M simulated
2x150 bp Illumina pairs (1 % substitutions + 1 % indels, seed 2) = 2 M reads (mates interleaved), Illumina_Paired
preset; every mate is aligned by the path (BinarySeeding -> SoC -> Harmonization -> NeedlemanWunsch) and then goes
through MappingQuality and PairedReads.  A step = one pass over the 2 M reads.
With N > 1 (torchrun, one rank per GPU) the index is replicated and every rank aligns its own 2 M reads (weak scaling,
no collective on the path); value = reads of all ranks / max-over-ranks device time.

  value     reads/s with reads + index resident in HBM (CUDA-event time of ma_b200_align_run)
  e2e       reads/s through ma_b200_align_batch with pinned HOST buffers: H2D of the reads, all kernels, D2H of the
            alignment records inside the timed region
  roofline  the dominant kernel of the step (by device time) against its bound
  cpu_baseline / --impl reference: the UNMODIFIED reference (oracle/_ref, compiled from /root/reference by
            oracle/Makefile) on the box's host cores over a bounded sample of the same reads
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from ma_b200 import index as maindex  # noqa: E402
from ma_b200 import synth  # noqa: E402

REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
CACHE = os.environ.get("MA_B200_CACHE", "/tmp/ma_b200_cache")
METRIC = "aligned reads/sec (2x150 Illumina)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


def make_genome(genome_mbp, seed):
    n_contigs = 10 if genome_mbp >= 10 else 1
    if genome_mbp >= 3000:
        n_contigs = genome_mbp // 100  # configs[2] / [3]: 31 x 100 Mbp
    contig = genome_mbp * 1_000_000 // n_contigs
    return synth.random_genome([contig] * n_contigs, seed)


def make_workload(genome_mbp, n_pairs, seed, rank=0):
    genome = make_genome(genome_mbp, seed)
    m1, m2, *_ = synth.simulate_pairs(genome, n_pairs, 150, seed * 1000 + 17 + rank)
    reads = np.empty((2 * n_pairs, 150), dtype=np.uint8)  # mates interleaved: read 2i, 2i+1 = pair i
    reads[0::2], reads[1::2] = m1, m2
    return genome, reads


def index_prefix(genome_mbp, seed):
    tag = hashlib.sha1(("g%d_s%d" % (genome_mbp, seed)).encode()).hexdigest()[:12]
    return os.path.join(CACHE, "idx_" + tag)


def ensure_index_files(genome, genome_mbp, seed, ctx=None):
    """Index in the reference's file formats (for the reference arm / cpu_baseline). ctx is None (the reference arm):
    built by the reference's own builder (ref_dump index), cached as <prefix>_ref. With a context (our arm's
    cpu_baseline): that cached index if it exists, else the GPU-built one (bit-identical to the reference builder's,
    tests/test_pipeline_gpu.py, tests/test_cli.py)."""
    prefix = index_prefix(genome_mbp, seed)
    exts = (".bwt", ".sa", ".pac", ".ann", ".amb")
    if all(os.path.exists(prefix + "_ref" + e) for e in exts):
        return prefix + "_ref", "cached, reference builder (ref_dump index)"
    os.makedirs(CACHE, exist_ok=True)
    if ctx is not None:
        if all(os.path.exists(prefix + e) for e in exts):
            return prefix, "cached, ma_b200_index_build"
        ix = ctx.index_download()
        tmp = prefix + ".tmp%d" % os.getpid()
        maindex.store_index(ix, tmp)
        for e in exts:
            os.replace(tmp + e, prefix + e)
        return prefix, "ma_b200_index_build (bit-identical to the reference builder)"
    gt = prefix + ".genome.txt"
    tmp = prefix + "_ref.tmp%d" % os.getpid()
    synth.write_genome_txt(gt, genome)
    subprocess.check_call([REF_DUMP, "index", gt, tmp], stdout=subprocess.DEVNULL)
    os.remove(gt)
    for e in exts:
        os.replace(tmp + e, prefix + "_ref" + e)
    return prefix + "_ref", "reference builder (ref_dump index)"


def run_reference(prefix, reads, threads, srand=-1, preset="illuminapaired"):
    """The cpu_baseline leg: the unmodified reference's modules (oracle/_ref/ref_dump bench) over a sample of reads."""
    os.makedirs(CACHE, exist_ok=True)
    rf = os.path.join(CACHE, "sample_%d_%d.txt" % (os.getpid(), len(reads)))
    synth.write_reads_txt(rf, reads)
    try:
        out = subprocess.check_output([REF_DUMP, "bench", prefix, rf, preset, str(threads)]).decode()
    finally:
        os.remove(rf)
    return json.loads(out.strip().splitlines()[-1])


def run_reference_ksw(pairs_file, threads, repeat):
    """cpu_baseline leg of the DP-only sweep (scripts/dp_sweep_bench.py --cpu): the unmodified kswcpp_dispatch."""
    return json.loads(subprocess.check_output([REF_DUMP, "kswbench", pairs_file, str(threads), str(repeat)]))


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


SM_COUNT, LANES_PER_SM, INT_OPS_PER_CELL = 148, 128, 44  # 25 score + 16 traceback + 3 H row / max


def config_of(args, world):
    """Workload description of the alignment configs."""
    c = args.config
    if c == "1":
        return {"id": 1, "genome_mbp": args.genome_mbp, "seed": 2, "preset": "illumina_paired", "ref_preset": "illuminapaired",
                "kind": "pairs", "read_len": 150, "scaling": "weak", "total_reads": None, "sub_batch": 2 * args.pairs,
                "metric": METRIC,
                "workload": "configs[1]: synthetic %d Mbp genome (10 contigs), %d simulated 2x150 bp Illumina pairs per "
                            "GPU and step (1%% subst + 1%% indel), Illumina_Paired preset: every mate aligned, "
                            "MappingQuality + PairedReads" % (args.genome_mbp, args.pairs)}
    if c == "2":
        pairs = args.pairs if args.pairs_given else 10_000_000
        return {"id": 2, "genome_mbp": 3100, "seed": 3, "preset": "illumina_paired", "ref_preset": "illuminapaired",
                "kind": "pairs", "read_len": 150, "scaling": "strong", "total_reads": 2 * pairs, "sub_batch": 2_000_000,
                "metric": METRIC,
                "workload": "configs[2]: synthetic 3.1 Gbp genome (31 contigs x 100 Mbp, seed 3), index replicated per "
                            "GPU; ONE fixed set of %d simulated 2x150 bp pairs (1%% subst + 1%% indel, seed 3) split into "
                            "contiguous shards over the %d rank(s), sub-batches of 1 M pairs; Illumina_Paired preset, "
                            "MappingQuality + PairedReads" % (pairs, world)}
    n = args.long_reads
    return {"id": 3, "genome_mbp": 3100, "seed": 3, "preset": "pacbio", "ref_preset": "pacbio", "kind": "long",
            "read_len": args.long_len, "scaling": "strong", "total_reads": n, "sub_batch": args.long_batch,
            "metric": "aligned reads/sec (%d bp PacBio, 12%% error)" % args.long_len,
            "workload": "configs[3]: synthetic 3.1 Gbp genome (31 contigs x 100 Mbp, seed 3), ONE fixed set of %d "
                        "simulated reads of %d bp (4%% subst + 4%% ins + 4%% del, seed 4) split over the %d rank(s), "
                        "sub-batches of %d reads; PacBio preset, MappingQuality" % (n, args.long_len, world, args.long_batch)}


def rank_reads(cfg, genome, flat, rank, world, args, workers):
    """This rank's reads [n, L] and the global index of its first read."""
    from ma_b200 import dist as madist
    if cfg["id"] == 1:
        m1, m2, *_ = synth.simulate_pairs(genome, args.pairs, 150, cfg["seed"] * 1000 + 17 + rank, flat=flat)
        reads = np.empty((2 * args.pairs, 150), dtype=np.uint8)  # mates interleaved: read 2i, 2i+1 = pair i
        reads[0::2], reads[1::2] = m1, m2
        return reads, 0
    if cfg["kind"] == "pairs":
        lo, hi = madist.shard_pairs(cfg["total_reads"] // 2, rank, world)
        return synth.pair_set_reads(genome, lo, hi, 150, cfg["seed"], workers, flat), lo
    lo, hi = madist.shard_pairs(cfg["total_reads"], rank, world)  # contiguous shard of single reads
    lo, hi = lo // 2, hi // 2
    return synth.long_set_reads(genome, lo, hi, cfg["read_len"], 4, workers, flat), lo


def reference_arm(args, cfg, world):
    """--impl reference: the UNMODIFIED reference on the host cores over a bounded sample of the same workload."""
    if not os.path.exists(REF_DUMP):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_dump not built"}))
        return 0
    sample_n = args.cpu_sample if cfg["kind"] == "pairs" else args.cpu_sample_long
    prefix = index_prefix(cfg["genome_mbp"], cfg["seed"])
    exts = (".bwt", ".sa", ".pac", ".ann", ".amb")
    if cfg["id"] != 1 and not any(all(os.path.exists(prefix + t + e) for e in exts) for t in ("_ref", "")):
        print(json.dumps({"impl": "reference", "unavailable": "no 3.1 Gbp index files under %s: the reference's own "
                          "builder needs about 1.5 h for this genome; run the default arm with --config %d once, it stores "
                          "the (bit-identical, tests/test_human_size_gpu.py) GPU-built index there" % (CACHE, cfg["id"])}))
        return 0
    genome = make_genome(cfg["genome_mbp"], cfg["seed"])
    flat = np.concatenate(genome)
    if cfg["id"] == 1:
        m1, m2, *_ = synth.simulate_pairs(genome, max(sample_n // 2, 1), 150, cfg["seed"] * 1000 + 17, flat=flat)
        sample = np.empty((2 * len(m1), 150), dtype=np.uint8)
        sample[0::2], sample[1::2] = m1, m2
        # nothing of this repository's engine on this arm: the index comes from the reference's own builder (about two
        # minutes for 100 Mbp, cached under MA_B200_CACHE for the runs that follow on the same box)
        prefix, how = ensure_index_files(genome, cfg["genome_mbp"], cfg["seed"], None)
    else:
        sample = (synth.pair_set_reads(genome, 0, sample_n, 150, cfg["seed"], 1, flat) if cfg["kind"] == "pairs"
                  else synth.long_set_reads(genome, 0, sample_n, cfg["read_len"], 4, os.cpu_count() or 1, flat))
        how = ("cached, reference builder" if os.path.exists(prefix + "_ref.bwt")
               else "cached index files written from ma_b200_index_build (bit-identical to the reference builder's)")
        prefix = prefix + "_ref" if os.path.exists(prefix + "_ref.bwt") else prefix
    del flat
    threads = os.cpu_count() or 1
    sample = sample[:sample_n]
    for _ in range(max(args.warmup, 0) and 1):
        run_reference(prefix, sample[:max(8, len(sample) // 30)], threads, preset=cfg["ref_preset"])
    t, n, aligned = 0.0, 0, 0
    for _ in range(args.steps):
        r = run_reference(prefix, sample, threads, preset=cfg["ref_preset"])
        t += r["seconds"]
        n += r["reads"]
        aligned += r["aligned"]
    v = aligned / t
    config = {"workload": cfg["workload"], "preset": cfg["preset"], "read_len": cfg["read_len"],
              "genome_bp": cfg["genome_mbp"] * 1_000_000,
              "sample": "each step = the first %d reads of the workload (a rate: the full workload would take hours "
                        "on the host)" % len(sample)}
    line = {"impl": "reference", "metric": cfg["metric"], "value": v, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * t / args.steps,
            "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "u8/int64",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "reference",
                             "sample": "%d reads of the same workload per step, ref_dump bench (the reference's seven "
                                       "modules incl. MappingQuality and PairedReads through their stock execute(), "
                                       "index preloaded), index: %s" % (len(sample), how)},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="1", choices=["1", "2", "3", "dp_sweep"],
                    help="BASELINE.json configs[1] (default), configs[2], configs[3], or the DP-only sweep configs[4]")
    ap.add_argument("--pairs", type=int, default=None,
                    help="config 1: read pairs per GPU and step (1 M); config 2: pairs of the whole fixed set (10 M)")
    ap.add_argument("--genome-mbp", type=int, default=100)
    ap.add_argument("--long-reads", type=int, default=100_000, help="config 3: reads of the whole fixed set")
    ap.add_argument("--long-len", type=int, default=10_000)
    ap.add_argument("--long-batch", type=int, default=100_000, help="config 3: reads per sub-batch (measured: 25 k 25.3, 50 k 26.3, 100 k 27.1 k reads/s)")
    ap.add_argument("--cpu-sample", type=int, default=60_000, help="reads of the bounded CPU-reference sample")
    ap.add_argument("--cpu-sample-long", type=int, default=320, help="... of long reads (config 3)")
    ap.add_argument("--split", type=int, default=0, help="reads per sub-batch of the pipelined align_batch (0: default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--all-records", action="store_true", help="e2e downloads every alignment record, not only the reported ones")
    args, rest = ap.parse_known_args()
    args.pairs_given = args.pairs is not None
    if args.pairs is None:
        args.pairs = 1_000_000
    rank, local_rank, world = dist_env()
    if args.config == "dp_sweep":
        if rank != 0:
            return 0
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import dp_sweep_bench
        return dp_sweep_bench.main(rest + (["--impl", "reference"] if args.impl == "reference" else []))
    cfg = config_of(args, world)

    if args.impl == "reference":
        return reference_arm(args, cfg, world) if rank == 0 else 0

    # ------------------------------------------------------------------ our arm
    # the workload is generated before CUDA is touched (forked worker processes)
    t0 = time.time()
    genome = make_genome(cfg["genome_mbp"], cfg["seed"])
    fwd = np.concatenate(genome)
    workers = max(1, (os.cpu_count() or 1) // world)
    reads, first_read = rank_reads(cfg, genome, fwd, rank, world, args, workers)
    t_gen = time.time() - t0
    import torch
    import torch.distributed as dist
    from ma_b200 import api
    from ma_b200 import dist as madist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = api.Context(local_rank, cfg["preset"])
    if args.split > 0:
        ctx.set_batch_split(args.split)
    if not args.all_records:
        # the end-to-end call hands back what the reference's chain hands to its writer: the records MappingQuality /
        # PairedReads return (ma_b200_set_reported_only); --all-records downloads every alignment NeedlemanWunsch computed
        ctx.set_reported_only(True)
    t0 = time.time()
    lens = [len(c) for c in genome]
    ctx.index_build(fwd, np.cumsum([0] + lens[:-1]), lens)
    t_index = time.time() - t0
    del fwd
    L = cfg["read_len"]
    n_reads = len(reads)
    base_params = api.preset(cfg["preset"])
    # sub-batches of this rank's shard: pinned host buffers (reads in, records out) per sub-batch
    subs = []
    for lo in range(0, n_reads, cfg["sub_batch"]):
        n = min(cfg["sub_batch"], n_reads - lo)
        pin = torch.empty(n * L, dtype=torch.uint8, pin_memory=True)
        pin.numpy()[:] = reads[lo:lo + n].reshape(-1)
        subs.append({"n": n, "lo": lo, "reads": pin, "offsets": np.arange(n + 1, dtype=np.int64) * L})
    max_n = max(s["n"] for s in subs)
    per_read_alns, per_read_runs = (3, 40) if cfg["kind"] == "pairs" else (8, 12 * L // 10)
    out = {"cap_alns": per_read_alns * max_n + 1024, "cap_runs": per_read_runs * max_n + 4096}

    def alloc_out(o):
        o["info"] = torch.empty(max_n * api.INFO_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
        o["alns"] = torch.empty(o["cap_alns"] * api.ALN_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
        o["runs"] = torch.empty(o["cap_runs"], dtype=torch.int32, pin_memory=True)

    alloc_out(out)

    def set_shard_params(sub, c=None):
        # RANSAC streams follow the GLOBAL read index: results do not depend on the sharding
        p = api.preset(cfg["preset"])
        p.srand_base = madist.shard_srand_base(base_params.srand_base, first_read + sub["lo"])
        (c or ctx).set_params(p)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def e2e_sub(sub, count=False, c=None, o=None):
        """ma_b200_align_batch: pinned host reads in, alignment records in pinned host memory out."""
        c, o = c or ctx, o or out
        set_shard_params(sub, c)
        st = api.AlignStats()
        rc = c.lib.ma_b200_align_batch(c.h, sub["n"], sub["reads"].data_ptr(), sub["offsets"].ctypes.data,
                                       o["info"].data_ptr(), o["alns"].data_ptr(), o["cap_alns"],
                                       o["runs"].data_ptr(), o["cap_runs"], ctypes.byref(st))
        if rc == -3:  # MA_B200_ENOMEM: the record buffers of this bench were too small; grow them from the counts
            o["cap_alns"] = max(o["cap_alns"], int(max(st.n_sets if args.all_records else 0, st.n_reported) * 1.3) + 1024)
            o["cap_runs"] = max(o["cap_runs"], int(st.n_runs * 1.3) + 4096)
            alloc_out(o)
            return e2e_sub(sub, count, c, o)
        c._check(rc)
        if not count:
            return st, 0
        info = np.frombuffer(o["info"].numpy(), dtype=api.INFO_DTYPE)[:sub["n"]]
        return st, int((info["n_sets"] > 0).sum())

    def device_pass(collect=None):
        """One step with the reads resident in HBM: per sub-batch upload (untimed), then the CUDA-event time of
        ma_b200_align_run. Returns the summed device ms."""
        ms = 0.0
        for sub in subs:
            set_shard_params(sub)
            if len(subs) > 1 or not out.get("uploaded"):
                ctx.align_upload(sub["reads"].numpy(), sub["offsets"])
                out["uploaded"] = len(subs) == 1
            st = ctx.align_run(api.STAGE_MAPQ)
            ms += st["ms_total"]
            if collect is not None:
                for k, v in st.items():
                    collect[k] = collect.get(k, 0.0) + v
        return ms

    # ---- device-resident timing: `value`
    for _ in range(args.warmup):
        device_pass()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count
    dev_ms, work = 0.0, {}
    for _ in range(args.steps):
        dev_ms += device_pass(work)
    barrier()
    launches = ctx.launch_count - launches0
    K = args.steps
    last = {k: v / K for k, v in work.items()}  # per step (sum over the sub-batches)
    # ---- end to end timing: `e2e`. Every step goes through ma_b200_align_batch with pinned HOST buffers (reads in,
    # records out). Two batches are kept in flight — two host threads, each with its own context (the second one a sibling
    # that shares the index) and its own record buffers — so that the copies of one batch run under the kernels of the
    # other: the way a caller with a stream of batches uses the call (maCMD_b200 does it per GPU). The figure of ONE
    # call at a time is measured as well (e2e.single_call).
    aligned = 0
    for sub in subs:  # warm-up pass of the end-to-end path; the aligned reads are counted here, outside the timed region
        aligned += e2e_sub(sub, True)[1]
    out["uploaded"] = False
    barrier()
    t = time.perf_counter()
    e2e_stage = {}
    for _ in range(args.steps):
        for sub in subs:
            est, _a = e2e_sub(sub)
            for k in ("ms_seed", "ms_locate", "ms_socharm", "ms_plan", "ms_dp", "ms_assemble", "ms_total"):
                e2e_stage[k] = e2e_stage.get(k, 0.0) + getattr(est, k) / args.steps
    torch.cuda.synchronize()
    e2e_single_s = time.perf_counter() - t
    ctx2 = ctx.sibling()
    out2 = {"cap_alns": out["cap_alns"], "cap_runs": out["cap_runs"]}
    alloc_out(out2)
    items = [sub for _ in range(args.steps) for sub in subs]
    lanes = [(ctx, out, items[0::2]), (ctx2, out2, items[1::2])]
    for c, o, mine in lanes:  # warm-up of both pipelines (slab sizes of the sibling)
        for sub in mine[:max(1, len(subs) // 2)] or items[:1]:
            e2e_sub(sub, False, c, o)
    counts = [[0, 0], [0, 0]]
    stage2 = [{kk: 0.0 for kk in ("ms_seed", "ms_locate", "ms_socharm", "ms_plan", "ms_dp", "ms_assemble", "ms_total")} for _ in range(2)]
    errors = []

    def lane(k):
        c, o, mine = lanes[k]
        try:
            torch.cuda.set_device(local_rank)
            for sub in mine:
                est, _a = e2e_sub(sub, False, c, o)
                counts[k][0] += est.n_sets if args.all_records else est.n_reported
                counts[k][1] += est.n_runs
                for kk in stage2[k]:
                    stage2[k][kk] += getattr(est, kk) / args.steps
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    barrier()
    t = time.perf_counter()
    threads = [threading.Thread(target=lane, args=(k,)) for k in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t
    if errors:
        raise errors[0]
    barrier()
    n_sets = (counts[0][0] + counts[1][0]) // args.steps
    n_runs = (counts[0][1] + counts[1][1]) // args.steps
    ctx2.close()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = n_reads * L + (n_reads + len(subs)) * 8
    d2h = n_reads * api.INFO_DTYPE.itemsize + n_sets * api.ALN_DTYPE.itemsize + n_runs * 4

    times = torch.tensor([dev_ms, e2e_s * 1000.0, float(aligned), float(h2d), float(d2h), float(n_reads), e2e_single_s * 1000.0],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        mx = times.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = times.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms_max, e2e_ms_max, e2e_single_ms_max = float(mx[0]), float(mx[1]), float(mx[6])
        aligned_all, h2d_all, d2h_all, reads_all = float(sm[2]), float(sm[3]), float(sm[4]), float(sm[5])
    else:
        dev_ms_max, e2e_ms_max, e2e_single_ms_max = dev_ms, e2e_s * 1000.0, e2e_single_s * 1000.0
        aligned_all, h2d_all, d2h_all, reads_all = float(aligned), float(h2d), float(d2h), float(n_reads)
    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_src = measured_peaks()
    value = aligned_all * K / (dev_ms_max / 1000.0)
    e2e_value = aligned_all * K / (e2e_ms_max / 1000.0)
    per = {k: last[k] for k in last if k.startswith("ms_")}
    sm_mhz = peaks.get("sm_max_mhz", 1965.0)
    gcups_peak = SM_COUNT * LANES_PER_SM * sm_mhz * 1e6 / INT_OPS_PER_CELL / 1e9
    occ_bytes = cfg["genome_mbp"] * 1_000_000  # 64 B per 128 BWT symbols of the 2 x genome text
    gather_gbs = ctx.gather_probe(occ_bytes)
    gather_hbm_gbs = ctx.gather_probe(8 << 30)
    ref_call_bytes = 128.0 * last["n_ext"]
    table_bytes = 128.0 * last["n_lookup"]
    locate_bytes = 64.0 * last["n_invpsi"] + 8.0 * last["n_seeds"]
    dp_cells = float(last["dp_cells"])

    def rate(b, ms):
        return b / ms / 1e6 if ms > 0 else None

    kernels = {
        "seed_kernel": {"ms": per["ms_seed"], "bound": "hbm-gather", "algorithmic_bytes": table_bytes,
                        "GB/s": rate(table_bytes, per["ms_seed"]), "peak": gather_gbs,
                        "frac": rate(table_bytes, per["ms_seed"]) / gather_gbs,
                        "reference_call_bytes": ref_call_bytes, "reference_call_GB/s": rate(ref_call_bytes, per["ms_seed"])},
        "locate_kernel": {"ms": per["ms_locate"], "bound": "hbm-gather", "algorithmic_bytes": locate_bytes,
                          "GB/s": rate(locate_bytes, per["ms_locate"]), "peak": gather_gbs,
                          "frac": rate(locate_bytes, per["ms_locate"]) / gather_gbs},
        "socharm_kernel": {"ms": per["ms_socharm"], "bound": "instruction issue (no byte / flop figure, SURVEY.md §8(d))"},
        "nwplan_kernel": {"ms": per["ms_plan"]},
        "ksw_kernels": {"ms": per["ms_dp"], "bound": "int-alu", "cells": dp_cells, "GCUPS": rate(dp_cells, per["ms_dp"]),
                        "peak": gcups_peak, "frac": rate(dp_cells, per["ms_dp"]) / gcups_peak,
                        "algorithmic_bytes": dp_cells * 1.0 + 2.0 * last["n_tasks"] * 600},
        "nwasm+alnsort+mapq+pair_kernel": {"ms": per["ms_assemble"]},
    }
    dom = max(("seed_kernel", "locate_kernel", "ksw_kernels"), key=lambda k: kernels[k]["ms"])
    traffic, traffic_src = None, None  # DRAM bytes per step of the dominant kernel, ncu --set full of this configuration
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        ent = cap.get("config%d" % cfg["id"], {})
        if ent.get("reads_per_step") == int(reads_all) and dom in ent:
            traffic = ent[dom]["dram_read_bytes"] + ent[dom]["dram_write_bytes"]
            traffic_src = ent.get("source")
    except (OSError, ValueError, KeyError):
        pass
    if dom == "ksw_kernels":
        roofline = {"kernel": "ksw_qs_kernel + ksw_batch_kernel (all DP launches of the step)", "bound": "int-alu",
                    "achieved": kernels[dom]["GCUPS"], "peak": gcups_peak, "unit": "GCUPS", "frac": kernels[dom]["frac"],
                    "peak_source": "SURVEY.md §8(d): %d SMs x %d INT32 lanes x %.0f MHz (%s) / %d integer ops per cell"
                                   % (SM_COUNT, LANES_PER_SM, sm_mhz, peak_src, INT_OPS_PER_CELL),
                    "algorithmic_cells_per_step": dp_cells,
                    "hbm_view": {"achieved_GBs": rate(kernels[dom]["algorithmic_bytes"], per["ms_dp"]),
                                 "peak_GBs": peaks["hbm_gbs"],
                                 "frac": rate(kernels[dom]["algorithmic_bytes"], per["ms_dp"]) / peaks["hbm_gbs"],
                                 "note": "1 B traceback per cell + sequences: far below the HBM bound, the kernel is "
                                         "integer-issue bound"}}
    else:
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["GB/s"], "peak": kernels[dom]["peak"],
                    "unit": "GB/s", "frac": kernels[dom]["frac"],
                    "peak_source": "ma_b200_gather_probe, this run: independent random 64-byte reads over a buffer of "
                                   "the occurrence table's size (%d MB); streaming peak %s GB/s (%s)"
                                   % (occ_bytes // 1_000_000, peaks["hbm_gbs"], peak_src),
                    "algorithmic_bytes_per_step": kernels[dom]["algorithmic_bytes"]}
    roofline.update({"traffic": traffic, "traffic_source": traffic_src, "share_of_step": kernels[dom]["ms"] / per["ms_total"]})
    residency = "mostly L2-resident (126 MB L2)" if occ_bytes <= 126_000_000 else "HBM-resident (larger than the 126 MB L2)"
    roofline_seeding = {"kernel": "seed_kernel", "bound": "hbm (random 64-byte blocks; the %d MB occ table is %s on "
                        "this configuration)" % (occ_bytes // 1_000_000, residency),
                        "achieved": kernels["seed_kernel"]["GB/s"], "peak": gather_gbs, "unit": "GB/s",
                        "frac": kernels["seed_kernel"]["frac"], "peak_hbm_resident_buffer": gather_hbm_gbs,
                        "reference_call_GBs": kernels["seed_kernel"]["reference_call_GB/s"],
                        "lookup_share": last["n_lookup"] / max(1.0, last["n_ext"]),
                        "note": "achieved = 128 B x the occurrence-table lookups the kernel really issues / kernel time "
                                "(entries with the same SA interval are merged: lookup_share of the reference's "
                                "extend_backward calls read the table); reference_call_GBs = 128 B x the reference's call "
                                "count / time (SURVEY.md §8(d)'s unit); ncu DRAM bytes and L2 hit rate: profiles/",
                        "peak_source": "ma_b200_gather_probe: independent random 64-byte reads over a buffer of the "
                                       "occ table's size (and over 8 GiB for the HBM-resident figure), this run"}

    cpu_baseline = None
    if not args.no_cpu_baseline and os.path.exists(REF_DUMP):
        prefix, how = ensure_index_files(genome, cfg["genome_mbp"], cfg["seed"], ctx)
        threads = os.cpu_count() or 1
        sample = reads[:args.cpu_sample if cfg["kind"] == "pairs" else args.cpu_sample_long]
        r = run_reference(prefix, sample, threads, preset=cfg["ref_preset"])
        cpu_baseline = {"value": r["aligned"] / r["seconds"], "unit": "reads/s", "cores": threads,
                        "kind": "reference",
                        "sample": "first %d reads of the same workload, oracle/_ref/ref_dump bench (the reference's "
                                  "seven modules (incl. MappingQuality, PairedReads) on %d host threads, index preloaded; "
                                  "index files: %s); stage cpu-seconds %s"
                                  % (len(sample), threads, how, json.dumps(r["stage_cpu_s"]))}
    config = {"workload": cfg["workload"], "preset": cfg["preset"], "read_len": L,
              "reads_per_step": int(reads_all), "reads_per_step_per_gpu": n_reads, "sub_batches_per_gpu": len(subs),
              "genome_bp": cfg["genome_mbp"] * 1_000_000, "parallelism": "index replicated, reads sharded x%d" % world,
              "timed_region": "value: CUDA-event time of ma_b200_align_run over the sub-batches, reads resident in HBM; "
                              "e2e: wall time of ma_b200_align_batch over the same sub-batches, pinned host buffers, two batches in flight, "
                              + ("every alignment record" if args.all_records else
                                 "the records MappingQuality / PairedReads return (what the writer consumes) + all run words"),
              "l2": "inputs larger than L2 (reads %d MB + index %d MB per step, no flush needed)"
                    % (n_reads * L // 1_000_000, cfg["genome_mbp"] * 7 // 4)}
    line = {"metric": cfg["metric"], "value": value, "unit": "reads/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
            "dtype": "u8/int64 (f64 in Harmonization)", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d_all),
                    "d2h_bytes_per_step": int(d2h_all), "ms_per_step": e2e_ms_max / K, "batches_in_flight": 2,
                    "device_stage_ms": {kk: round(stage2[0][kk] + stage2[1][kk], 2) for kk in stage2[0]},
                    "single_call": {"value": aligned_all * K / (e2e_single_ms_max / 1000.0), "ms_per_step": e2e_single_ms_max / K,
                                    "device_stage_ms": {k: round(v, 2) for k, v in e2e_stage.items()},
                                    "note": "one ma_b200_align_batch at a time (upload under the seeding kernel, run "
                                            "words down under stage 4, records down after the last kernel)"},
                    "note": "K steps through ma_b200_align_batch from two host threads on two contexts of the GPU (the "
                            "second a sibling sharing the index): the host<->device copies of one batch run under the "
                            "kernels of the other"},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline,
            "roofline_seeding": roofline_seeding, "kernels": kernels, "cpu_baseline": cpu_baseline,
            "aligned_reads_per_step": aligned_all, "Mbp_per_s": value * L / 1e6, "index_build_s": t_index,
            "workload_generation_s": t_gen,
            "work_per_step_rank0": {k: int(last[k]) for k in ("n_reads", "n_seeds", "n_sets", "n_tasks", "n_ext",
                                                              "n_lookup", "n_invpsi", "dp_cells", "n_dropped")}}
    print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
