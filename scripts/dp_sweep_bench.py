#!/usr/bin/env python
"""BASELINE.json configs[4] — the DP-only sweep: kswcpp-equivalent banded DP on synthetic query/ref pairs
(SURVEY.md §8(d) config 5).  Driver-runnable as  python bench.py --config dp_sweep [--impl reference]  or directly:

  python scripts/dp_sweep_bench.py [--out profiles/dp_sweep.json] [--pairs 100000] [--max-cells 5e10] [--quick]

Grid: length 100 .. 20 000 x band 16 .. 512 x {global, extension, reversed extension} x divergence {1, 5, 15} %;
query = mutated copy of the target. Per point >= --pairs problems (64 distinct pairs, repeated) unless that exceeds
--max-cells band cells, then as many as fit (stated per point). Every problem is computed with ALL kswcpp_extz_t fields
and its CIGAR (the exact mode; extension points additionally in the extension-only mode the alignment path uses).
Unit: the band cell the reference processes (sum over rows of en0 - st0 + 1 up to the row where it stops).
  value     cells of the whole grid / summed CUDA-event kernel time, inputs resident in HBM (ma_b200_ksw_upload / _run)
  e2e       the same cells / wall time of ma_b200_ksw_batch (host buffers in, results + CIGARs out)
  cpu_baseline / --impl reference: the UNMODIFIED kswcpp_dispatch (oracle/_ref/ref_dump kswbench) on the 64 distinct
            pairs of the 5 % extension points, all host threads, in the same run
Roofline: 148 SMs x 128 INT32 lanes x 1.965 GHz / 44 integer ops per cell = 846 GCUPS.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import dpgen  # noqa: E402

REF_DUMP = bench.REF_DUMP
MODES = {"global": dpgen.GLOBAL, "ext": dpgen.EXT, "ext_right": dpgen.EXT_RIGHT}
LENGTHS = (100, 300, 1000, 3000, 10000, 20000)
BANDS = (16, 32, 64, 128, 256, 512)
DIVS = (0.01, 0.05, 0.15)
METRIC = "DP GCUPS (kswcpp global / extension, band cells of the reference per second)"


def write_pairs(path, pairs):
    with open(path, "w") as f:
        for w, zd, fl, q, t in pairs:
            f.write("%d %d %d %s %s\n" % (w, zd, fl, "".join(str(int(c)) for c in q) or "-",
                                           "".join(str(int(c)) for c in t) or "-"))


def grid(quick):
    for length in LENGTHS:
        for w in BANDS:
            if w > 2 * length:
                continue
            for mode in MODES:
                for div in DIVS:
                    if quick and not (div == 0.05 and mode != "ext_right"):
                        continue
                    yield length, w, mode, div


def base_pairs(length, w, mode, div):
    return dpgen.sweep_pairs(64, length, w, MODES[mode], div, seed=length * 31 + w + int(div * 1000))


def cpu_point(base, threads, cells_per_base_pass):
    """GCUPS of the unmodified kswcpp_dispatch on the distinct pairs of a point (bounded: about 2e8 cells per thread)."""
    with tempfile.TemporaryDirectory() as d:
        pf = os.path.join(d, "p.txt")
        write_pairs(pf, base)
        rep = max(1, int(2e8 / max(cells_per_base_pass, 1)))
        o = bench.run_reference_ksw(pf, threads, rep * threads)  # bench.py's cpu_baseline leg
        return cells_per_base_pass * (o["calls"] / len(base)) / o["seconds"] / 1e9, o["seconds"]


def reference_arm(args):
    """--impl reference: kswcpp_dispatch on all host threads over the distinct pairs of every 5 % point."""
    import helpers as H
    threads = os.cpu_count() or 1
    cells, secs, pts = 0.0, 0.0, 0
    t0 = time.time()
    for length, w, mode, div in grid(True):
        base = base_pairs(length, w, mode, div)
        per = sum(H.oracle_ksw(q, t, ww, zd, fl)[2] for ww, zd, fl, q, t in base[:4]) / 4 * len(base)
        g, s = cpu_point(base, threads, per)
        cells += g * s * 1e9
        secs += s
        pts += 1
        if time.time() - t0 > 240:
            break
    v = cells / secs / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "GCUPS", "n_gpus": 1, "steps": 1, "warmup": 0,
            "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8/int16/int32",
            "data": "synthetic", "config": {"workload": "configs[4]: DP-only sweep, %d points at 5 %% divergence (64 distinct "
                                                        "pairs each, repeated to about 2e8 cells per thread)" % pts},
            "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": threads, "kind": "reference",
                             "sample": "ref_dump kswbench (unmodified kswcpp_dispatch) on %d host threads, %d points" % (threads, pts)},
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=100_000, help="problems per point (SURVEY.md §8(d): >= 1e5)")
    ap.add_argument("--max-cells", type=float, default=5e10, help="cap of band cells per point (long x wide points)")
    ap.add_argument("--quick", action="store_true", help="5 %% divergence, global + extension only")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args(argv)
    if args.impl == "reference":
        return reference_arm(args)
    import torch
    from ma_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("dp_sweep: no CUDA device — the path has no CPU fallback")
    ctx = api.Context(0)
    threads = os.cpu_count() or 1
    peaks, peak_src = bench.measured_peaks()
    gcups_peak = bench.SM_COUNT * bench.LANES_PER_SM * peaks.get("sm_max_mhz", 1965.0) * 1e6 / bench.INT_OPS_PER_CELL / 1e9
    sampler = bench.ClockSampler(0)
    sampler.start()
    launches0 = ctx.launch_count
    points, tot_cells, tot_ms, tot_e2e_s, h2d, d2h = [], 0.0, 0.0, 0.0, 0, 0
    cpu_cells, cpu_secs = 0.0, 0.0
    for length, w, mode, div in grid(args.quick):
        base = base_pairs(length, w, mode, div)
        per = sum(min(len(q), len(t), 2 * w + 1) * (len(q) + len(t)) for _, _, _, q, t in base) / len(base)  # upper bound
        n = int(max(64, min(args.pairs, args.max_cells / per)))
        reps = (n + 63) // 64
        btasks, bseq = api.pack_ksw_tasks(base)
        # the 64 distinct pairs repeated: every repetition is its own task on its own copy of the sequences
        tasks = np.tile(btasks, reps)
        shift = np.repeat(np.arange(reps, dtype=np.int64) * len(bseq), len(btasks))
        tasks["qoff"] += shift
        tasks["toff"] += shift
        seq = np.tile(bseq, reps)
        row = {"len": length, "w": w, "mode": mode, "div": div, "tasks": int(len(tasks)), "int16_mode": bool(length <= 1365)}
        ctx.ksw_set_extension_only(False)
        ctx.ksw_upload(tasks, seq)
        ctx.ksw_run()  # warm-up: sizes the traceback / CIGAR slabs
        ms = ctx.ksw_run()
        res, _ = ctx.ksw_download()
        assert (res["status"] == 0).all()
        cells = int(res["cells"].sum())
        row.update(cells=cells, ms=ms, gcups=cells / ms / 1e6, frac=cells / ms / 1e6 / gcups_peak)
        t = time.perf_counter()
        r2, c2 = ctx.ksw_batch(tasks, seq)
        dt = time.perf_counter() - t
        row.update(e2e_ms=dt * 1e3, e2e_gcups=cells / dt / 1e9)
        h2d += tasks.nbytes + seq.nbytes
        d2h += r2.nbytes + int(r2["n_cigar"].sum()) * 4
        if mode != "global":
            ctx.ksw_set_extension_only(True)
            ctx.ksw_upload(tasks, seq)
            ctx.ksw_run()
            ms2 = ctx.ksw_run()
            res2, _ = ctx.ksw_download()
            ctx.ksw_set_extension_only(False)
            row.update(ext_only_ms=ms2, ext_only_gcups_equiv=cells / ms2 / 1e6, ext_only_cells_done=int(res2["cells"].sum()))
        if not args.no_cpu and os.path.exists(REF_DUMP) and div == 0.05 and mode == "ext":
            g, s = cpu_point(base, threads, cells / reps)
            row["cpu_gcups_%dt" % threads] = g
            cpu_cells += g * s * 1e9
            cpu_secs += s
        tot_cells += cells
        tot_ms += ms
        tot_e2e_s += dt
        points.append(row)
        print(json.dumps(row), file=sys.stderr, flush=True)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    value = tot_cells / tot_ms / 1e6
    wide = [p for p in points if p["w"] >= 128]
    line = {"metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": 1, "steps": 1, "warmup": 1, "ms_per_step": tot_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8 differences, int16/int32 H (kswcpp's modes)",
            "data": "synthetic",
            "config": {"workload": "configs[4]: DP-only sweep, %d points (length %s x band %s x {global, ext, ext_right} x "
                                   "divergence %s), >= %d problems per point capped at %.0e band cells, all kswcpp_extz_t "
                                   "fields + CIGAR" % (len(points), list(LENGTHS), list(BANDS), list(DIVS), args.pairs, args.max_cells),
                       "l2": "traceback slabs of a point exceed the L2 (1 byte per cell)"},
            "e2e": {"value": tot_cells / tot_e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": tot_e2e_s * 1e3},
            "gpu_launches": int(ctx.launch_count - launches0), "clocks": sampler.summary(),
            "roofline": {"kernel": "ksw_batch_kernel (exact mode)", "bound": "int-alu", "achieved": value, "peak": gcups_peak,
                         "unit": "GCUPS", "frac": value / gcups_peak, "traffic": None,
                         "peak_source": "SURVEY.md §8(d): 148 SMs x 128 INT32 lanes x %.0f MHz (%s) / 44 integer ops per cell"
                                        % (peaks.get("sm_max_mhz", 1965.0), peak_src),
                         "band_ge_128": {"min_gcups": min(p["gcups"] for p in wide), "max_gcups": max(p["gcups"] for p in wide),
                                         "mean_gcups": sum(p["cells"] for p in wide) / sum(p["ms"] for p in wide) / 1e6}},
            "cpu_baseline": ({"value": cpu_cells / cpu_secs / 1e9, "unit": "GCUPS", "cores": threads, "kind": "reference",
                              "sample": "ref_dump kswbench (unmodified kswcpp_dispatch) on the 64 distinct pairs of every 5 % "
                                        "extension point, about 2e8 cells per thread and point, same run"}
                             if cpu_secs > 0 else None),
            "points": points}
    print(json.dumps(line))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(line, f, indent=1)
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
