#!/usr/bin/env python
"""bench.py — exchange-factor rays/s on BASELINE.json's metric configuration (config 3: 101x101 grey scattering
enclosure, kappa = sigma_s = 0.5, 1e10 rays per GPU per step).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference on the host cores

A "step" is one full exchange-factor trace: zero the UInt64 count matrix, trace rays_total rays (every emitter
row, all bands), reduce the per-GPU matrices to rank 0 (N > 1).  `value` is device-resident whole-job rays/s;
`e2e` is the same trace through the C ABI with HOST buffers (mesh upload + trace + device->host copy of the
counts inside the timed region).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

METRIC = "exchange_factor_rays_per_sec"
UNIT = "rays/s"
# algorithmic FP64 flop per ray by (emitter kind, ending), SURVEY.md §8(d) / BASELINE.md §4
FLOP_SG, FLOP_SW, FLOP_VG, FLOP_VW, FLOP_CROSS = 90.0, 138.0, 98.0, 146.0, 65.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rthx", choices=["rthx", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--rays", type=float, default=None, help="rays per GPU per step (default: 1e10 for cfg3)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-sample-rays", type=float, default=1.0e9,
                    help="rays of the bounded CPU-baseline sample (about 15 s on 16 cores); the reference arm takes a fifth of it per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-smoothing", action="store_true")
    ap.add_argument("--block-threads", type=int, default=0)
    ap.add_argument("--row-chunks", type=int, default=0)
    ap.add_argument("--locator", default="auto", choices=["auto", "generic"],
                    help="generic = the reference-faithful grid + point-in-polygon locator everywhere (the path arbitrary, non-affine meshes take)")
    ap.add_argument("--mode", default="first_interaction", choices=["first_interaction", "multi_bounce", "multi_bounce_specular"],
                    help="first_interaction = what method=:exchange computes (the headline); multi_bounce = total-exchange mode, rays followed "
                         "through scattering / wall reflection until absorbed (rthx.h RTHX_MULTI_BOUNCE)")
    ap.add_argument("--reduce", default="fused", choices=["fused", "nccl"],
                    help="N > 1: fused peer-memory flush over NVLink (default) or a private matrix per rank + NCCL reduce")
    return ap.parse_args()


# DRAM bytes per launch of the trace kernel measured by ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum)
NCU_DRAM_BYTES_PER_LAUNCH = {"cfg3": (894.04e6 + 833.20e6, "profiles/r1y_trace_exchange_sq_metrics.csv"),
                             "cfg5": (None, "profiles/r1y_trace_exchange_queue_metrics.csv")}

# executed warp instructions per 32 rays of the trace kernel (ncu source page, profiles/<capture>_sass_mix.csv, TOTAL row)
NCU_WARP_INSTR_PER_32_RAYS = {"cfg3": (250.5, "profiles/r1y_trace_exchange_sq_sass_mix.csv"),
                              "cfg5": (680.9, "profiles/r1y_trace_exchange_queue_sass_mix.csv")}

DEFAULT_RAYS = {"cfg1": 1e6, "cfg2": 1e8, "cfg3": 1e10, "cfg4": 1e8, "cfg5": 1e9}


def build_workload(name):
    import rthx
    rtm = getattr(rthx.meshes, name)()
    flat = rthx.flatten_domain(rtm)
    if rtm.spectral_mode == "spectral_variable":
        groups, reps, nonuniform = rthx.group_uniform_bins(rtm.uniform_across_bin)
        bins = [b - 1 for b in nonuniform] + [g[0] - 1 for g in groups]
    else:
        bins = [0]
    return rtm, flat, bins


def flop_per_ray(stats):
    n = stats["n_surface_gas"] + stats["n_surface_wall"] + stats["n_volume_gas"] + stats["n_volume_wall"]
    if n == 0:
        return 100.0
    return (FLOP_SG * stats["n_surface_gas"] + FLOP_SW * stats["n_surface_wall"] + FLOP_VG * stats["n_volume_gas"]
            + FLOP_VW * stats["n_volume_wall"] + FLOP_CROSS * stats["n_crossings"]) / n


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_oracle_sample(flat, bins, rays_total, seed):
    """Time the CPU oracle (all host threads) on a bounded sample of the workload."""
    from oracle import oracle
    N = flat.n_elements
    rpe = max(1, int(rays_total) // (N * len(bins)))
    t0 = time.perf_counter()
    # every host core this process may run on, set explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers
    n_threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    out = oracle.trace(flat, rpe, seed=seed, bins=bins, n_threads=n_threads)
    dt = time.perf_counter() - t0
    traced = rpe * N * len(bins)
    return dict(rays_per_s=traced / dt, seconds=dt, rays=traced, rpe=rpe, stats=out["stats"])


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self._stop.is_set():
                    break
                self.rows.append((time.monotonic(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def stop(self):
        self._stop.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self, t0=None, t1=None):
        """Median SM clock and active throttle reasons over the samples taken inside [t0, t1] (the timed region)."""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if (t0 is not None and ts < t0) or (t1 is not None and ts > t1):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_julia_reference(args, sample):
    """If a `julia` binary and baseline/_ref exist, time the UNMODIFIED reference (baseline/run_julia_ref.jl).
    Neither exists in this image, so this returns None and the oracle port is timed instead."""
    import shutil
    julia = shutil.which("julia")
    ref = os.path.join(_ROOT, "baseline", "_ref")
    if julia is None or not os.path.isdir(ref):
        return None
    ndim, kappa, sig = {"cfg1": (11, 1.0, 0.0), "cfg2": (41, 1.0, 0.0), "cfg3": (101, 0.5, 0.5)}[args.workload]
    try:
        r = subprocess.run([julia, "-t", "auto", f"--project={ref}", os.path.join(_ROOT, "baseline", "run_julia_ref.jl"),
                            str(ndim), str(kappa), str(sig), str(int(sample)), str(args.steps)],
                           capture_output=True, text=True, timeout=1500)
        line = next(l for l in r.stdout.splitlines() if l.startswith("RTHX_JULIA_REF"))
        kv = dict(x.split("=") for x in line.split()[1:])
    except Exception:
        return None
    value = float(kv["rays_per_s"])
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(kv["rays"]) / value * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {workload_desc(args.workload)}"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(kv["threads"]), "kind": "reference",
                             "sample": f"{kv['rays']} rays per step, parallelRayTracing of the unmodified Julia reference"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def main_reference(args):
    """Reference arm: the reference's own algorithm on the host cores.  Julia is absent from the image, so this is
    the CPU oracle (oracle/rthx_oracle.c, kind "port") with every host thread, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rtm, flat, bins = build_workload(args.workload)
    sample = args.cpu_sample_rays / 5.0
    julia_line = run_julia_reference(args, sample) if args.workload in ("cfg1", "cfg2", "cfg3") else None
    if julia_line is not None:
        print(json.dumps(julia_line))
        return 0
    for _ in range(max(0, args.warmup)):
        run_oracle_sample(flat, bins, sample / 10.0, seed=7)
    t0 = time.perf_counter()
    rays = 0
    last = None
    for s in range(args.steps):
        last = run_oracle_sample(flat, bins, sample, seed=100 + s)
        rays += last["rays"]
    dt = time.perf_counter() - t0
    value = rays / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(1, args.steps) * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {workload_desc(args.workload)}", "elements": flat.n_elements,
                   "bands": len(bins)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["stats"]["n_threads"] if last else 0, "kind": "port",
                         "sample": f"{int(sample):d} rays per step of the same workload (C restatement of the reference "
                                   f"algorithm, OpenMP; Julia absent); cpu: {cpu_model()}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_desc(name):
    return {
        "cfg1": "README Example 1, 1x1 m square, 11x11 grey gas kappa=1, black walls",
        "cfg2": "2D grey reflecting enclosure 41x41, kappa=1",
        "cfg3": "2D grey absorbing+scattering medium kappa=0.5 sigma_s=0.5, 101x101 mesh",
        "cfg4": "2D spectral multi-band gas 51x51, 8 bands batched in the grid",
        "cfg5": "16-wedge circle with transparent interfaces and triangle sub-meshes (11,11)",
    }[name]


def main():
    args = parse_args()
    if args.impl == "reference":
        return main_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import rthx
    from rthx.dist import ShardedTracer, reduce_counts

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the rthx path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    rtm, flat, bins = build_workload(args.workload)
    N = flat.n_elements
    nb = len(bins)
    rays_per_gpu = float(args.rays) if args.rays else DEFAULT_RAYS[args.workload]
    rays_total = rays_per_gpu * (world if args.scaling == "weak" else 1)
    rpe = int(rays_total) // N      # rays_total applies per traced band (parallelRayTracing.jl:22-25,34-37)
    traced_per_step = rpe * N * nb

    sh = ShardedTracer(flat, device=local_rank, rank=rank, world=world, n_bins=nb, mode=args.reduce)
    kw = dict(bins=bins, block_threads=args.block_threads, row_chunks=args.row_chunks, locator=1 if args.locator == "generic" else 0,
              mode={"first_interaction": 0, "multi_bounce": 1, "multi_bounce_specular": 2}[args.mode])
    stream = torch.cuda.current_stream(dev)

    def step(seed, time_kernel=None):
        if time_kernel is not None:
            time_kernel[0].record(stream)
        st = sh.enqueue(rpe, seed=seed, **kw)        # zero (own rows) + trace kernel (+ fused peer flush)
        if time_kernel is not None:
            time_kernel[1].record(stream)
        sh.finish()                                  # barrier (fused) or NCCL reduce
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing ------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()                      # nvidia-smi needs ~0.5 s to deliver its first sample: start before warm-up
    for w in range(args.warmup):
        st = step(1000 + w)
    barrier()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_win0 = time.monotonic()
    e0.record(stream)
    for s in range(args.steps):
        st = step(2000 + s, kev[s])
    e1.record(stream)
    barrier()
    t_win1 = time.monotonic()
    if sampler:
        time.sleep(0.15)
        sampler.stop()
    total_ms = e0.elapsed_time(e1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / max(1, args.steps)
    t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / max(1, args.steps)
    value = traced_per_step / (ms_per_step * 1e-3)
    lost_total = int(sh.lost.sum().item()) if rank == 0 else 0
    tallied = int(sh.counts.sum().item()) if rank == 0 else 0

    # ---- end-to-end through the C ABI with host buffers ------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        counts_host = torch.empty((nb, N, N), dtype=torch.int64, pin_memory=True) if world == 1 else None
        mesh_bytes = sum(getattr(flat, n).nbytes for n in (
            "coarse_nv", "coarse_vx", "coarse_vy", "coarse_solid", "fine_off", "cell_nv", "cell_vx", "cell_vy", "cell_mid",
            "cell_volume", "cell_surf_id", "kappa", "sigma_s", "epsilon", "uniform_beta"))

        e2e_dev_ms = []
        e2e_phases = []   # N > 1: [create, trace + own-row D2H, barrier, close] ms per step

        def e2e_step(seed):
            if world == 1:
                tr = rthx.DeviceTracer(flat, device=local_rank)       # mesh flattening output -> device (H2D)
                out = tr.trace(rpe, counts_out=counts_host.numpy().view(np.uint64), seed=seed, **kw)
                tr.close()
                e2e_dev_ms.append((out["stats"]["kernel_ms"], out["stats"]["total_ms"]))
                return int(out["lost"].sum())
            # N > 1: every rank re-uploads the mesh and copies ITS rows into one page-locked matrix in POSIX shared
            # memory over its own PCIe link (overlapped with tracing); a barrier makes the matrix complete on rank 0.
            tp = [time.perf_counter()]
            if shared_host is not None:
                tr = rthx.DeviceTracer(flat, device=local_rank)
                tp.append(time.perf_counter())
                tr.trace(rpe, counts_out=shared_host.array, seed=seed, emitter_rank=rank, emitter_world=world, **kw)
                tp.append(time.perf_counter())
                dist.barrier(device_ids=[local_rank])
                tp.append(time.perf_counter())
                tr.close()
            else:
                # fallback when /dev/shm cannot hold the matrix: device-side flush / reduce, then D2H on rank 0
                old = sh.tracer
                sh.tracer = rthx.DeviceTracer(flat, device=local_rank)
                tp.append(time.perf_counter())
                sh.trace(rpe, seed=seed, **kw)
                tp.append(time.perf_counter())
                if rank == 0:
                    fallback_host.copy_(sh.counts, non_blocking=True)
                torch.cuda.synchronize(dev)
                tp.append(time.perf_counter())
                old.close()
            tp.append(time.perf_counter())
            e2e_phases.append([1e3 * (b - a) for a, b in zip(tp[:-1], tp[1:])])
            return 0

        shared_host = fallback_host = None
        if world > 1:
            from rthx._lib import SharedHostMatrix
            shm_name = f"rthx_bench_{os.environ.get('MASTER_PORT', '0')}"
            ok = torch.ones(1, dtype=torch.int32, device=dev)
            if rank == 0:
                try:
                    shared_host = SharedHostMatrix(shm_name, (nb, N, N), create=True)
                    shared_host.array[0, 0, :8] = 0          # touch
                except Exception as ex:                       # e.g. a 64 MB /dev/shm
                    print(f"bench.py: shared host matrix unavailable ({ex}); e2e falls back to D2H on rank 0", file=sys.stderr)
                    shared_host = None
                    ok.zero_()
            dist.broadcast(ok, src=0)
            if int(ok.item()) == 1:
                if rank != 0:
                    shared_host = SharedHostMatrix(shm_name, (nb, N, N), create=False)
            elif rank == 0:
                fallback_host = torch.empty((nb, N, N), dtype=torch.int64, pin_memory=True)
        e2e_step(3000)
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            e2e_step(3100 + s)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        e2e = {"value": traced_per_step * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(mesh_bytes + 4 * nb),
               "d2h_bytes_per_step": int(8 * nb * N * N + 8 * nb * N),
               "ms_per_step": dt / args.steps * 1e3,
               "device_ms_per_step": {"kernel": float(np.mean([a for a, _ in e2e_dev_ms[1:]])),
                                      "zero+kernel+d2h": float(np.mean([b for _, b in e2e_dev_ms[1:]]))} if e2e_dev_ms else None,
               "phases_ms": [float(x) for x in np.mean(np.array(e2e_phases[1:]), axis=0)] if len(e2e_phases) > 1 else None,
               "path": "rthx_create + rthx_trace_exchange (pinned host count matrix)" if world == 1 else
                       "per rank: rthx_create + rthx_trace_exchange(own rows -> page-locked shared host matrix); barrier"
                       if shared_host is not None else
                       f"rthx_create + rthx_trace_exchange_device ({args.reduce}) + D2H of the matrix on rank 0"}
        if world > 1 and shared_host is not None:
            e2e["check_total"] = int(shared_host.array.sum()) if rank == 0 else None
            dist.barrier(device_ids=[local_rank])
            shared_host.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- next-stage kernel: dense reciprocity smoothing of the traced matrix on the device (HBM-bound) -----------
    smoothing = solve = None
    if world == 1 and not args.no_smoothing:
        try:
            hbm_peak = json.load(open(os.path.join(_ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth of this pool's B200)"
        except Exception:
            hbm_peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        tr = rthx.DeviceTracer(flat, device=local_rank)
        tr.trace(max(1, rpe // 10), counts_out=counts_host.numpy().view(np.uint64) if not args.no_e2e else None, seed=4000, **kw)
        w = rthx.get_w(rtm)
        F_host = torch.empty((N, N), dtype=torch.float64, pin_memory=True)
        _, ss = tr.smooth(w / w.min(), max_iters=1000, measure_pass=True, out=F_host.numpy())
        # grey equilibrium solve on the F_smooth that is still resident on the device (equilibriumGrey2D.jl:148-194)
        from rthx import equilibrium as _eq
        _, _b, coeff, h = _eq._system_vectors(rtm, _eq.populateWorkspace(rtm))
        _, _, sv = tr.solve_grey(coeff, h, measure_pass=True)
        solve = {"kernel": "matvec_t_partial_kernel (y = F'x on the resident dense F_smooth: one read of F per Krylov step)",
                 "bound": "hbm", "achieved": sv["matvec_gbs"], "peak": hbm_peak, "unit": "GB/s",
                 "frac": sv["matvec_gbs"] / hbm_peak, "bytes_per_pass": sv["matvec_bytes"], "pass_ms": sv["matvec_ms"],
                 "iterations": sv["iterations"], "restarts": sv["restarts"], "matvecs": sv["matvecs"],
                 "converged": sv["converged"], "residual": sv["residual"], "rhs_norm": sv["rhs_norm"],
                 "total_ms": sv["total_ms"], "launches": sv["launches"], "peak_source": peak_src,
                 "traffic": 900086016 + 6328576,
                 "traffic_note": "dram bytes read + written per launch, ncu --set full, profiles/r1g_solve_matvec_metrics.csv"}
        tr.close()
        smoothing = {"kernel": "scale_rows_kernel (one alternating-projection iteration: X *= (u_i+u_j)/2 with fused row sums)",
                     "bound": "hbm", "achieved": ss["pass_gbs"], "peak": hbm_peak, "unit": "GB/s",
                     "frac": ss["pass_gbs"] / hbm_peak, "bytes_per_pass": 16 * N * N, "pass_ms": ss["pass_ms"],
                     "iterations": ss["iterations"], "delta_init": ss["delta_init"], "delta": ss["delta"],
                     "total_ms": ss["total_ms"], "peak_source": peak_src}

    # ---- roofline (FP64 pipe) and CPU baseline ---------------------------------------------------------------
    fp64_peak = sh.tracer.measure_fp64_peak()
    cpu = None
    A = 100.0
    if world == 1 and not args.no_cpu_baseline:
        r = run_oracle_sample(flat, bins, args.cpu_sample_rays, seed=7)
        A = flop_per_ray(r["stats"])
        cpu = {"value": r["rays_per_s"], "unit": UNIT, "cores": r["stats"]["n_threads"], "kind": "port",
               "sample": f"{r['rays']} rays of the same workload in {r['seconds']:.1f} s (C restatement of the reference "
                         f"algorithm, OpenMP, Julia absent); cpu: {cpu_model()}"}
    kernel_rays_per_s = (traced_per_step / world) / (kernel_ms * 1e-3)
    achieved = kernel_rays_per_s * A / 1e12
    info = sh.tracer.info
    kernel_name = ("trace_exchange_kernel<MULTI> (multi-bounce)" if args.mode != "first_interaction" else
                   "trace_exchange_kernel (generic locator)" if args.locator == "generic" else
                   "trace_exchange_sq_kernel" if info["n_coarse"] == 1 and info["n_affine_faces"] == 1 else
                   "trace_exchange_queue_kernel" if info["n_affine_faces"] == info["n_coarse"] else "trace_exchange_kernel")
    hbm_bytes_per_ray = 8.0 * N * N * nb / world / max(1, traced_per_step / world)
    roofline = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak if fp64_peak else None,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(args.workload, (None, None))[0],
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture "
                                f"{NCU_DRAM_BYTES_PER_LAUNCH.get(args.workload, (None, 'none for this workload'))[1]} (the reductions "
                                "read-modify-write the zeroed 8*N*N-byte count matrix once, independent of the ray count)",
                "kernel": kernel_name, "kernel_ms": kernel_ms, "kernel_rays_per_s": kernel_rays_per_s,
                "flop_per_ray": A,
                "peak_source": "FP64 DFMA-chain micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "hbm": {"algorithmic_bytes_per_ray": hbm_bytes_per_ray,
                        "achieved_gbs": kernel_rays_per_s * hbm_bytes_per_ray / 1e9, "peak_gbs": 6559.4,
                        "note": "count-matrix write-out only; the path is not HBM-bound"}}
    # secondary explanation: the kernel is bound by instruction issue (three half-rate pipes share one issue port per SM
    # sub-partition), so state the issue-slot utilisation its measured rate implies
    clocks = sampler.summary(t_win0, t_win1 + 0.1) if sampler else None
    wi = NCU_WARP_INSTR_PER_32_RAYS.get(args.workload)
    if wi and clocks and clocks.get("sm_mhz"):
        slots = info["sm_count"] * 4 * clocks["sm_mhz"] * 1e6
        roofline["issue"] = {"warp_instr_per_32_rays": wi[0], "source": wi[1], "issue_slots_per_s": slots,
                             "frac": kernel_rays_per_s / 32.0 * wi[0] / slots,
                             "note": "executed warp instructions (ncu) x measured ray rate / (SMs x 4 sub-partitions x SM clock)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {workload_desc(args.workload)}", "elements": N, "bands": nb,
                   "rays_per_step": traced_per_step, "rays_per_emitter": rpe,
                   "partition": (f"emitter rows e % {world} == rank; " + ("rows flushed into rank 0's matrix over NVLink peer memory "
                                 "inside the trace kernel, then a barrier" if sh.mode == "fused" else
                                 "one NCCL reduce of the u64 count matrix to rank 0" if sh.mode == "nccl" else "single GPU")),
                   "l2": "count matrix (8*N*N bytes) exceeds the 126 MB L2 for cfg3; a fresh Philox seed every step",
                   "launch": {k: st[k] for k in ("n_blocks", "block_threads", "row_chunks", "smem_bytes", "hist_in_smem")}},
        "roofline": roofline, "smoothing": smoothing, "solve": solve, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": args.steps * world,
        "clocks": clocks,
        "check": {"tallied_last_step": tallied, "lost_last_step": lost_total},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
