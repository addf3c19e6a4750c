#!/usr/bin/env python
"""bench.py — the POA window-correction path on BASELINE.json's config[1] workload.

  python bench.py [--gpus N --steps K --warmup W]                      our arm (CUDA engine through the C-ABI)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   one rank per GPU, weak scaling
  python bench.py --impl reference [--steps K --warmup W]              the reference's own CPU path (oracle/_ref)

Workload ("step" = one pass of the hot path over one batch): synthetic PacBio CLR of SURVEY.md §8(d) config 2
(10k reads x 10 kb over a 3.33 Mb genome, 15 % error 9:4.5:1.5, Q~N(12,2), ground-truth overlaps, 500 bp
windows, m=3 x=-5 g=-4 -p -d 0.2 -s 0.2 -k 3).  One batch = all windows of `--targets` consecutive target reads
(default 1600 reads = 32000 windows, depth ~30, ~100 GB of DP scratch) per GPU; rank r takes T targets starting at
r*min(T, 10000/N) (weak scaling, no data-path collective; the gather of corrected reads to rank 0 is part of the e2e leg).

Legs of our arm:
  value : windows/s with the batch already resident in HBM (vgc_upload once; each timed step = kernels + D2H of
          the corrected windows), timed on the device with CUDA events on the engine's launch stream, max over ranks
  e2e   : windows/s through the public calls (vechat_b200.polisher.Polisher.submit_shard / collect_shard ->
          vgc_submit / vgc_collect): pinned HOST buffers in, host buffers out; every step's host prep + packing + H2D +
          kernels + D2H + stitch (+ NCCL gather at N > 1) inside the timed region, steps pipelined two deep
  roofline / cpu_baseline: see DESIGN.md §Measurement.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import vechat_b200  # noqa: E402,F401  (sets CUDA_DEVICE_MAX_CONNECTIONS before torch creates the CUDA context)

WORKLOAD = "pb_clr_10k_x_10kb"
METRIC = "poa_windows_per_sec"
UNIT = "windows/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--targets", type=int, default=1600, help="target reads per GPU per step (20 windows each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work budget of the cpu_baseline leg")
    ap.add_argument("--workload", default="pb", choices=["pb", "ont", "hap2"],
                    help="pb = BASELINE configs[1] (the benchmark); ont / hap2 = configs[2] / configs[3], extra evidence only")
    ap.add_argument("--hap2-leg", action="store_true", help="also run the hap2 (configs[3] shape) leg at N = 1 (always run at N > 1)")
    ap.add_argument("--no-hap2-leg", action="store_true", help="experiments: skip the hap2 leg at N > 1")
    ap.add_argument("--num-prune", type=int, default=3, help="diagnostics only: -k of the haplotype path (3 = the benchmark)")
    return ap.parse_args()


WORKLOADS = {
    "pb": ("pb_clr_10k_x_10kb", "synthetic PacBio CLR 10k reads x 10 kb, 15% error"),
    "ont": ("ont_10k_x_20kb", "synthetic ONT 10k reads x 20 kb, 10% error"),
    "hap2": ("hap2_50k_x_12kb", "2-haplotype 50:50 mix, 50k reads x 12 kb, 15% error"),
}


def workload_config(args, world):
    name, text = WORKLOADS[args.workload]
    return {"workload": "%s: %s, 500 bp windows, haplotype mode "
                        "(-p -d 0.2 -s 0.2 -k 3, m=3 x=-5 g=-4); batch = all windows of %d target reads per GPU "
                        "per step" % (name, text, args.targets),
            "targets_per_gpu": args.targets, "window_length": 500, "ranks": world,
            "sharding": "whole target reads per rank, no data-path collective; gather of corrected reads in e2e",
            "l2": "inputs per step (~1 GB at 1600 targets) and DP scratch (~100 GB) exceed the 126 MB L2; no explicit flush"}


def make_batch(args, rank, world=1):
    """Rank r corrects the windows of targets [r*s, r*s + T): T = --targets, s = min(T, n_reads // world) — disjoint
    ranges while world * T fits the config's 10k reads, overlapping ones beyond that (8 x 1600 > 10k); either way
    every rank carries the same amount of work (weak scaling)."""
    from vechat_b200.sim import Simulator
    sim = Simulator(WORKLOADS[args.workload][0])
    T = min(args.targets, sim.n_reads)
    stride = min(T, sim.n_reads // max(world, 1))
    t0 = min(rank * stride, sim.n_reads - T)
    b = sim.windows(t0, t0 + T)
    return sim, b


def pin_batch(batch):
    """Re-home the batch arrays in pinned host memory (the e2e leg copies from pinned memory)."""
    import torch
    from vechat_b200._ffi import WindowBatch
    keep = []

    def pin(a):
        t = torch.empty(max(a.nbytes, 8), dtype=torch.uint8, pin_memory=True)
        v = t.numpy()[:a.nbytes].view(a.dtype)
        v[...] = a
        keep.append(t)
        return v

    nb = WindowBatch(pin(batch.bases), pin(batch.quals), pin(batch.seq_off), pin(batch.has_qual), pin(batch.begin),
                     pin(batch.end), pin(batch.win_first), pin(batch.win_flags))
    for attr in ("win_target", "win_rank", "target_coverage"):
        setattr(nb, attr, getattr(batch, attr))
    nb._pinned = keep
    return nb


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def cpu_leg(batch, params, seconds, threads=None):
    """The reference's CPU path (oracle/_ref if it travelled, else the oracle port) on a bounded sample of the
    same batch: every k-th window, sized by a pilot so the run takes about `seconds`."""
    from oracle import checker
    kind = "reference" if checker.have_ref() else "port"
    fn = checker.ref_polish if kind == "reference" else checker.oracle_polish
    cores = threads or os.cpu_count() or 1
    nw = batch.n_windows
    pilot_idx = np.linspace(0, nw - 1, num=min(nw, 2 * cores)).astype(int)
    pb = batch.select(pilot_idx)
    t0 = time.perf_counter()
    fn(pb, params, threads=cores)
    rate = len(pilot_idx) / max(time.perf_counter() - t0, 1e-6)
    n = int(max(cores, min(nw, rate * seconds)))
    idx = np.linspace(0, nw - 1, num=n).astype(int)
    sb = batch.select(idx)
    sb.sample_index = idx
    return kind, cores, sb, fn


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from vechat_b200._ffi import make_params
    _, batch = make_batch(args, 0)
    params = make_params()
    # sample sized so that (warmup + steps) passes stay within a few minutes
    per_step = max(2.0, min(10.0, 150.0 / max(1, args.steps + args.warmup)))
    kind, cores, sb, fn = cpu_leg(batch, params, per_step)
    bases = 0
    for _ in range(args.warmup):
        fn(sb, params, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = fn(sb, params, threads=cores)
        bases += r.total_bases()
    dt = time.perf_counter() - t0
    value = sb.n_windows * args.steps / dt
    sample = "%d of the batch's %d windows (every k-th), %d host threads, %s" % (
        sb.n_windows, batch.n_windows, cores,
        "unmodified reference window.cpp + spoa compiled -O3 -msse4.1" if kind == "reference" else "oracle port")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": workload_config(args, args.gpus), "corrected_bases_per_sec": bases / dt,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def measure(args, torch, dist, world, rank, local, cpu_seconds):
    """One workload through both legs on every rank; returns the JSON line (rank 0) or None."""
    from vechat_b200._ffi import make_params
    from vechat_b200.engine import Engine
    from vechat_b200.polisher import Polisher, stitch, _gather_records

    sim, batch = make_batch(args, rank, world)
    batch = pin_batch(batch)
    params = make_params(num_prune=args.num_prune)
    eng = Engine(local, num_prune=args.num_prune)
    pol = Polisher(local, rank=rank, world=world, engine=eng)
    dev = torch.device("cuda", local)
    names = lambda t: "read%d" % t  # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from concurrent.futures import ThreadPoolExecutor

    def finish(result):
        recs = stitch(result, batch.win_target, batch.win_rank, names, batch.target_coverage)
        if world > 1:
            recs, _ = _gather_records(recs, rank, world, None, dev)
        return recs

    def e2e_run(steps):
        """`steps` batches through the public submit / collect pair (Polisher.submit_shard / collect_shard =
        vgc_submit / vgc_collect): every step stages its own copy of the inputs from pinned host memory (host
        prepare + packing + H2D), runs the kernels, copies the corrected windows back and stitches them; the
        staging of step k + 1 and the stitch of step k - 1 overlap the kernels of step k."""
        stats, result, recs = [], None, None
        trace = os.environ.get("BENCH_TRACE") and rank == 0
        pol.submit_shard(batch)
        with ThreadPoolExecutor(1) as ex:
            prev = None
            for k in range(steps):
                ta = time.perf_counter()
                fut = ex.submit(pol.collect_shard)
                if k + 1 < steps:
                    pol.submit_shard(batch)
                tb = time.perf_counter()
                if prev is not None:
                    recs = finish(prev)
                tc = time.perf_counter()
                prev = fut.result()
                if trace:
                    print("[bench] step %d: submit %.1f ms, finish(prev) %.1f ms, waited %.1f ms more for collect" % (
                        k, (tb - ta) * 1e3, (tc - tb) * 1e3, (time.perf_counter() - tc) * 1e3), file=sys.stderr)
                stats.append(pol.last_stats)
            result = prev
            recs = finish(prev)
        return result, recs, stats

    # ---- warm-up (both legs) ----------------------------------------------------------------------
    e2e_run(args.warmup)
    sampler = ClockSampler(local)
    sampler.start()

    # ---- e2e leg: host buffers through the public call ---------------------------------------------
    barrier()
    t0 = time.perf_counter()
    result, recs, e2e_stats = e2e_run(args.steps)
    barrier()
    e2e_dt = time.perf_counter() - t0
    corrected = result.total_bases()

    # ---- resident leg: inputs already in HBM; device-event timing ------------------------------------
    eng.upload(batch)
    for _ in range(args.warmup):
        eng.polish_resident()
    barrier()
    t0 = time.perf_counter()
    res_stats = []
    for _ in range(args.steps):
        _, st = eng.polish_resident()
        res_stats.append(st)
    barrier()
    res_wall = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)

    dev_ms = sum(s["device_ms"] for s in res_stats)
    kern_ms = sum(s["kernel_ms"] for s in res_stats)
    launches = sum(s["kernel_launches"] for s in res_stats) + sum(s["kernel_launches"] for s in e2e_stats)
    cells = res_stats[-1]["cells"]
    # e2e diagnostics: device time of the passes inside the e2e leg (mean over the steps; max over ranks below) — a gap
    # to the resident leg's kernel time means something shared the GPU or starved the launches, not the copies
    e2e_kern_mean = sum(s["kernel_ms"] for s in e2e_stats) / max(1, len(e2e_stats))
    e2e_wait_mean = sum(s.get("host_prep_ms", 0.0) + s.get("host_pack_ms", 0.0) for s in e2e_stats) / max(1, len(e2e_stats))
    agg = torch.tensor([dev_ms, e2e_dt * 1e3, res_wall * 1e3, kern_ms, e2e_kern_mean, e2e_wait_mean],
                       dtype=torch.float64, device=dev)
    tot = torch.tensor([batch.n_windows, corrected, launches, cells], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max, res_wall_max, kern_ms_max, e2e_kern_max, e2e_stage_max = [float(x) for x in agg.tolist()]
    n_windows, n_bases, n_launch, n_cells = [float(x) for x in tot.tolist()]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # roofline of the dominant kernel (fill_kernel: it alone writes the DP cells).  A pass is thousands of small
        # launches of four kernels on dozens of streams, so the unit here is one PASS (= one step): algorithmic bytes
        # = 2 B x sum over the pass's alignments of (R_a+1) x L_a of THIS rank, divided by the device time of the
        # whole pass (CUDA events on the launch stream, in the library) — i.e. the fill is charged with the traceback /
        # update / sort kernels it waits for.  The fill kernel's own figure (ncu launch list) is in profiles/.
        k_launches = len(res_stats)
        alg_bytes_per_launch = 2.0 * cells
        k_avg_s = kern_ms / 1e3 / max(k_launches, 1)
        achieved = alg_bytes_per_launch / k_avg_s / 1e9
        # DRAM traffic of the align kernels per pass: bytes per DP cell measured with ncu over every align_kernel
        # launch of whole passes of this workload on the final binary (dram__bytes_read.sum + dram__bytes_write.sum,
        # profiles/traffic.json says how) x the cells of this pass
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = float(json.load(open(tpath))["align_dram_bytes_per_cell"]) * cells
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": n_windows * args.steps / (dev_ms_max / 1e3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": workload_config(args, world),
            "corrected_bases_per_sec": n_bases * args.steps / (dev_ms_max / 1e3),
            "windows_per_step": n_windows, "ms_per_step_wall": res_wall_max / args.steps,
            "e2e": {"value": n_windows * args.steps / (e2e_ms_max / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": int(e2e_stats[-1]["input_bytes"]),
                    "d2h_bytes_per_step": int(e2e_stats[-1]["output_bytes"]),
                    "ms_per_step": e2e_ms_max / args.steps,
                    "corrected_bases_per_sec": n_bases * args.steps / (e2e_ms_max / 1e3),
                    "pipeline": "vgc_submit / vgc_collect: staging of step k+1 (host prepare, 2-bit packing, H2D) and "
                                "stitch of step k-1 overlap the kernels of step k; every step copies its own inputs",
                    "kernel_ms_mean_max_over_ranks": e2e_kern_max, "staging_host_ms_mean_max_over_ranks": e2e_stage_max,
                    "host_prep_ms": e2e_stats[-1]["host_prep_ms"], "host_pack_ms": e2e_stats[-1]["host_pack_ms"],
                    "h2d_ms": e2e_stats[-1]["h2d_ms"],
                    "kernel_ms": e2e_stats[-1]["kernel_ms"], "d2h_ms": e2e_stats[-1]["d2h_ms"]},
            "gpu_launches": int(n_launch),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "align_kernel (DP fill + traceback), charged with the whole lockstep pass (update / sort / align)",
                         "launch_unit": "one pass = one step (all kernel launches of a vgc_polish_resident call)",
                         "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                         "dp_cells_per_launch": cells, "kernel_ms_per_launch": k_avg_s * 1e3,
                         "gcups": cells / k_avg_s / 1e9},
            "clocks": sampler.summary(),
        }
        ph = eng.phase_profile()
        # per-window latency shares (leader-lane / walker cycles summed over windows): diagnostics, not device time.
        # trace_refill is a sub-interval of traceback and trace_refills a count: both are left out of the total
        from vechat_b200.engine import PHASE_NOT_CYCLES
        tot_ph = sum(v for k, v in ph.items() if k not in PHASE_NOT_CYCLES) or 1.0
        line["phase_share"] = {k: round(v / tot_ph, 4) for k, v in ph.items() if k not in PHASE_NOT_CYCLES}
        line["phase_raw"] = {k: float(v) for k, v in ph.items()}
        line["alignments"] = int(res_stats[-1]["alignments"])
        line["relaunched_windows"] = int(res_stats[-1]["relaunched_windows"])
        if cpu_seconds > 0:
            kind, cores, sb, fn = cpu_leg(batch, params, cpu_seconds)
            t0 = time.perf_counter()
            r = fn(sb, params, threads=cores)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {
                "value": sb.n_windows / dt, "unit": UNIT, "cores": cores, "kind": kind,
                "corrected_bases_per_sec": r.total_bases() / dt,
                "sample": "%d of the batch's %d windows (every k-th) in %.1f s on %d host threads" % (
                    sb.n_windows, batch.n_windows, dt, cores)}
            # the checker's windows against the engine's (last e2e step), byte for byte: parity at the bench's size
            bad = sum(1 for i, w in enumerate(sb.sample_index)
                      if result.window(int(w)) != r.window(i) or int(result.polished[int(w)]) != int(r.polished[i]))
            line["parity"] = {"windows_checked": int(sb.n_windows), "mismatches": int(bad),
                              "against": "compiled reference" if kind == "reference" else "oracle port"}
            # two more CPU figures on a smaller sample: the same reference on ONE core, and its
            # -mavx2 build (spoa's 16-lane engine) on all cores
            if kind == "reference" and args.workload == "pb":
                from oracle import checker
                variants = {}
                small = batch.select(sb.sample_index[:: max(1, len(sb.sample_index) // max(8, sb.n_windows // 40))])
                t0 = time.perf_counter()
                fn(small, params, threads=1)
                variants["sse41_1_core"] = {"value": small.n_windows / (time.perf_counter() - t0), "unit": UNIT,
                                            "cores": 1, "windows": int(small.n_windows)}
                if checker.have_ref_avx2():
                    mid = batch.select(sb.sample_index[::3])
                    t0 = time.perf_counter()
                    r2 = checker.ref_avx2_polish(mid, params, threads=cores)
                    variants["avx2_all_cores"] = {"value": mid.n_windows / (time.perf_counter() - t0), "unit": UNIT,
                                                  "cores": cores, "windows": int(mid.n_windows),
                                                  "same_output_as_sse41": all(
                                                      r2.window(i) == r.window(3 * i) for i in range(mid.n_windows))}
                line["cpu_baseline"]["variants"] = variants
        eng.close()
        return line
    eng.close()
    return None


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL only carries the per-step gather of the corrected reads to rank 0 (and the timing reductions); it runs
        # while the next pass owns the GPU.  Its kernels take one CTA per channel and spin there until the peers
        # arrive, and at 8 ranks the default channel count cost the root's pass 8 % (987 vs 911 ms, round-2 run):
        # two channels move the 0.5 GB in ~15 ms and occupy two SMs.
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "2")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    line = measure(args, torch, dist, world, rank, local,
                   args.cpu_seconds if (world == 1 and not args.no_cpu_baseline) else 0.0)
    if (world > 1 or args.hap2_leg) and args.workload == "pb" and not args.no_hap2_leg:
        # the shape the 8-GPU target is quoted on (BASELINE configs[3]: 2-haplotype mix, depth ~60), on disjoint
        # target ranges (50k reads): same legs, fewer steps, with a parity sample against the CPU reference on rank 0
        import copy
        a2 = copy.copy(args)
        a2.workload = "hap2"
        a2.steps = max(2, min(args.steps, 8))  # enough steps to amortise the pipeline fill of the e2e leg
        a2.warmup = 3
        a2.targets = min(args.targets, 800)
        h = measure(a2, torch, dist, world, rank, local, 0.0 if args.no_cpu_baseline else 8.0)
        if rank == 0:
            keep = ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "config", "windows_per_step", "e2e",
                    "roofline", "cpu_baseline", "parity", "alignments", "relaunched_windows")
            line["hap2"] = {k: h[k] for k in keep if k in h}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
