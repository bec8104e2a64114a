#!/usr/bin/env python
"""bench.py — aligned reads/s of the B200-native MA hot path (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P] [--genome-mbp G]

Workload (N = 1): BASELINE.json configs[1] — synthetic 100 Mbp genome (10 contigs x 10 Mbp, seed 2), 1 M simulated
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


def make_workload(genome_mbp, n_pairs, seed, rank=0):
    n_contigs = 10 if genome_mbp >= 10 else 1
    contig = genome_mbp * 1_000_000 // n_contigs
    genome = synth.random_genome([contig] * n_contigs, seed)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="read pairs per GPU and step (configs[1]: 1 M)")
    ap.add_argument("--genome-mbp", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=60_000, help="reads of the bounded CPU-reference sample")
    ap.add_argument("--split", type=int, default=0, help="reads per sub-batch of the pipelined align_batch (0: default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    seed = 2
    config = {"workload": "configs[1]: synthetic %d Mbp genome (10 contigs), %d simulated 2x150 bp Illumina pairs "
                          "per GPU and step (1%% subst + 1%% indel), Illumina_Paired preset: every mate aligned, MappingQuality + "
                          "PairedReads"
                          % (args.genome_mbp, args.pairs),
              "preset": "illumina_paired", "reads_per_step_per_gpu": 2 * args.pairs, "read_len": 150,
              "genome_bp": args.genome_mbp * 1_000_000, "parallelism": "index replicated, reads sharded x%d" % world,
              "l2": "inputs larger than L2 (reads %d MB + index %d MB per step, no flush needed)"
                    % (2 * args.pairs * 150 // 1_000_000, args.genome_mbp * 7 // 4)}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        if not os.path.exists(REF_DUMP):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_dump not built"}))
            return 0
        genome, reads = make_workload(args.genome_mbp, max(args.cpu_sample // 2, 1), seed)
        # nothing of this repository's engine on this arm: the index comes from the reference's own builder (about two
        # minutes for 100 Mbp, cached under MA_B200_CACHE for the runs that follow on the same box)
        prefix, how = ensure_index_files(genome, args.genome_mbp, seed, None)
        threads = os.cpu_count() or 1
        sample = reads[:args.cpu_sample]
        for _ in range(max(args.warmup, 0) and 1):
            run_reference(prefix, sample[:2000], threads)
        t, n, aligned = 0.0, 0, 0
        for _ in range(args.steps):
            r = run_reference(prefix, sample, threads)
            t += r["seconds"]
            n += r["reads"]
            aligned += r["aligned"]
        v = aligned / t
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * t / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "reference",
                                 "sample": "%d reads of the same workload per step, ref_dump bench (five reference "
                                           "modules, index preloaded), index: %s" % (len(sample), how)},
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from ma_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    genome, reads = make_workload(args.genome_mbp, args.pairs, seed, rank)
    ctx = api.Context(local_rank, "illumina_paired")
    if args.split > 0:
        ctx.set_batch_split(args.split)
    t0 = time.time()
    fwd = np.concatenate(genome)
    lens = [len(c) for c in genome]
    ctx.index_build(fwd, np.cumsum([0] + lens[:-1]), lens)
    t_index = time.time() - t0
    n_reads = len(reads)
    # pinned host buffers for the e2e path
    pin_reads = torch.empty(n_reads * 150, dtype=torch.uint8, pin_memory=True)
    pin_reads.numpy()[:] = reads.reshape(-1)
    offsets = np.arange(n_reads + 1, dtype=np.int64) * 150
    cap_alns, cap_runs = 3 * n_reads + 1024, 40 * n_reads + 4096
    pin_info = torch.empty(n_reads * api.INFO_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
    pin_alns = torch.empty(cap_alns * api.ALN_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
    pin_runs = torch.empty(cap_runs, dtype=torch.int32, pin_memory=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def e2e_step():
        st = api.AlignStats()
        t = time.perf_counter()
        rc = ctx.lib.ma_b200_align_batch(ctx.h, n_reads, pin_reads.data_ptr(), offsets.ctypes.data,
                                         pin_info.data_ptr(), pin_alns.data_ptr(), cap_alns, pin_runs.data_ptr(),
                                         cap_runs, ctypes.byref(st))
        ctx._check(rc)
        return time.perf_counter() - t, st

    # ---- device-resident timing: `value`
    ctx.align_upload(pin_reads.numpy(), offsets)
    for _ in range(args.warmup):
        ctx.align_run(api.STAGE_MAPQ)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count
    dev_ms, stage = 0.0, {}
    for _ in range(args.steps):
        st = ctx.align_run(api.STAGE_MAPQ)
        dev_ms += st["ms_total"]
        for k, v in st.items():
            if k.startswith("ms_"):
                stage[k] = stage.get(k, 0.0) + v
    barrier()
    launches = ctx.launch_count - launches0
    last = st
    # ---- end to end timing: `e2e`
    e2e_step()
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        dt, est = e2e_step()
        e2e_s += dt
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    info = np.frombuffer(pin_info.numpy(), dtype=api.INFO_DTYPE)
    aligned = int((info["n_sets"] > 0).sum())
    h2d = n_reads * 150 + (n_reads + 1) * 8
    d2h = n_reads * api.INFO_DTYPE.itemsize + est.n_sets * api.ALN_DTYPE.itemsize + est.n_runs * 4

    times = torch.tensor([dev_ms, e2e_s * 1000.0, float(aligned)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = times.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = times.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms_max, e2e_ms_max, aligned_all = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        dev_ms_max, e2e_ms_max, aligned_all = dev_ms, e2e_s * 1000.0, float(aligned)
    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_src = measured_peaks()
    K = args.steps
    value = aligned_all * K / (dev_ms_max / 1000.0)
    e2e_value = aligned_all * K / (e2e_ms_max / 1000.0)
    per = {k: v / K for k, v in stage.items()}
    # dominant kernel of the step
    seed_bytes = 128.0 * last["n_ext"]
    locate_bytes = 64.0 * last["n_invpsi"] + 8.0 * last["n_seeds"]
    dp_cells = float(last["dp_cells"])
    kernels = {
        "seed_kernel": {"ms": per["ms_seed"], "algorithmic_bytes": seed_bytes,
                        "GB/s": seed_bytes / per["ms_seed"] / 1e6 if per["ms_seed"] > 0 else None},
        "locate_kernel": {"ms": per["ms_locate"], "algorithmic_bytes": locate_bytes,
                          "GB/s": locate_bytes / per["ms_locate"] / 1e6 if per["ms_locate"] > 0 else None},
        "socharm_kernel": {"ms": per["ms_socharm"]},
        "nwplan_kernel": {"ms": per["ms_plan"]},
        "ksw_batch_kernel": {"ms": per["ms_dp"], "cells": dp_cells,
                             "GCUPS": dp_cells / per["ms_dp"] / 1e6 if per["ms_dp"] > 0 else None,
                             "algorithmic_bytes": dp_cells * 1.0 + 2.0 * last["n_tasks"] * 600},
        "nwasm+alnsort+mapq+pair_kernel": {"ms": per["ms_assemble"]},
    }
    dom = max(("seed_kernel", "locate_kernel", "ksw_batch_kernel"), key=lambda k: kernels[k]["ms"])
    ach = kernels[dom]["algorithmic_bytes"] / kernels[dom]["ms"] / 1e6
    traffic = None  # DRAM bytes per step of the dominant kernel from the committed ncu capture of this configuration
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if cap["workload"] == {"genome_mbp": args.genome_mbp, "pairs": args.pairs} and dom in cap:
            traffic = cap[dom]["dram_read_bytes"] + cap[dom]["dram_write_bytes"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic,
                "traffic_source": "profiles/ncu_traffic.json (ncu --set full, same workload; bytes per step = all "
                                  "launches of the kernel)" if traffic else None,
                "algorithmic_bytes_per_step": kernels[dom]["algorithmic_bytes"], "peak_source": peak_src,
                "share_of_step": kernels[dom]["ms"] / per["ms_total"],
                "note": "algorithmic bytes: 128 B per extend_backward, 64 B per invPsi step + 8 B per SA sample, "
                        "1 B traceback per DP cell + sequences (SURVEY.md §8(d)); the DP kernel is integer-ALU "
                        "bound: see kernels.ksw_batch_kernel.GCUPS and dp_int_roofline"}
    # seeding roofline: measured random 64-byte gather bandwidth over a buffer of the occ table's size
    occ_bytes = args.genome_mbp * 1_000_000
    gather_gbs = ctx.gather_probe(occ_bytes)
    gather_hbm_gbs = ctx.gather_probe(8 << 30)
    seed_ach = kernels["seed_kernel"]["GB/s"]
    table_gbs = 128.0 * last["n_lookup"] / per["ms_seed"] / 1e6 if per["ms_seed"] > 0 else None
    residency = "mostly L2-resident (126 MB L2)" if occ_bytes <= 126_000_000 else "HBM-resident (larger than the 126 MB L2)"
    roofline_seeding = {"kernel": "seed_kernel", "bound": "hbm (random 64-byte blocks; the %d MB occ table is %s on "
                        "this configuration)" % (occ_bytes // 1_000_000, residency),
                        "achieved": seed_ach, "peak": gather_gbs, "unit": "GB/s", "frac": seed_ach / gather_gbs,
                        "peak_hbm_resident_buffer": gather_hbm_gbs,
                        "table_reads_GBs": table_gbs,
                        "frac_table_reads": table_gbs / gather_gbs if table_gbs else None,
                        "lookup_share": last["n_lookup"] / max(1, last["n_ext"]),
                        "note": "achieved = 128 B x extend_backward calls of the reference algorithm / kernel time; "
                                "the kernel reads the table for lookup_share of them and reuses the previous result "
                                "for entries with the same SA interval (table_reads_GBs is that real traffic, "
                                "frac_table_reads its share of the gather rate; a frac above 1 means that reuse and "
                                "the cached top of the table beat independent random gathers)",
                        "peak_source": "ma_b200_gather_probe: independent random 64-byte reads over a buffer of the "
                                       "occ table's size (and over 8 GiB for the HBM-resident figure), this run"}
    sm_mhz = peaks.get("sm_max_mhz", 1965.0)
    int_ops_peak = 148 * 128 * sm_mhz * 1e6  # INT32 lane-ops/s
    dp_roof = {"gcups": kernels["ksw_batch_kernel"]["GCUPS"], "int32_lane_ops_per_s_peak": int_ops_peak,
               "gcups_at_44_ops_per_cell": int_ops_peak / 44 / 1e9}

    cpu_baseline = None
    if not args.no_cpu_baseline and os.path.exists(REF_DUMP):
        prefix, how = ensure_index_files(genome, args.genome_mbp, seed, ctx)
        threads = os.cpu_count() or 1
        sample = reads[:args.cpu_sample]
        r = run_reference(prefix, sample, threads)
        cpu_baseline = {"value": r["aligned"] / r["seconds"], "unit": "reads/s", "cores": threads,
                        "kind": "reference",
                        "sample": "first %d reads of the same workload, oracle/_ref/ref_dump bench (the reference's "
                                  "seven modules (incl. MappingQuality, PairedReads) on %d host threads, index preloaded); stage cpu-seconds %s"
                                  % (len(sample), threads, json.dumps(r["stage_cpu_s"]))}
    line = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int64 (f64 in Harmonization)", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms_max / K},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline,
            "roofline_seeding": roofline_seeding,
            "dp_int_roofline": dp_roof, "kernels": kernels, "cpu_baseline": cpu_baseline,
            "aligned_reads_per_step": aligned_all, "index_build_s": t_index,
            "work_per_step": {k: int(last[k]) for k in ("n_reads", "n_seeds", "n_sets", "n_tasks", "n_ext",
                                                       "n_lookup", "n_invpsi", "dp_cells", "n_dropped")}}
    print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
