#!/usr/bin/env python
"""bench.py -- SU(3) Wilson HMC MD steps/s on B200 (BASELINE.json metric) with roofline + CPU baseline.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W     # CPU arm (oracle port)

A "step" is one QPQ molecular-dynamics step (md_step!, src/molecular_dynamics.jl:611-616) of the
quenched Wilson action over the whole lattice.  The timed region is ONE gfb_md_trajectory call of
exactly K steps (diagnostics off), device-timed with CUDA events on the library's compute stream,
bracketed by barrier + device synchronize, max over ranks.  Workload (BASELINE.json configs[4], the lattice the
metric and the north-star target are quoted on): synthetic hot start (seed 1234), Gaussian momenta (seed 0x5678),
beta = 6.2, GLOBAL lattice 64^4 split into t-slabs over the N GPUs (strong scaling; 64^4 fits one B200: 24 GB).
`--scaling weak --lattice X,Y,Z,T` keeps X*Y*Z*T sites PER GPU instead.  The link field (9.7 GB / N per GPU) is larger
than L2 (126 MB), so no flush is needed.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gaugefields.jl_b200"))

METRIC = "SU(3) Wilson HMC MD steps/s"
UNIT = "MD steps/s"
U_BYTES, P_BYTES = 576, 256  # per site


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lattice", default="64,64,64,64", help="NX,NY,NZ,NT: the global lattice (strong) or the per-GPU lattice (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--beta", type=float, default=6.2)
    ap.add_argument("--tau", type=float, default=1.0)
    ap.add_argument("--unfused", action="store_true", help="issue the reference's op sequence (link, kick, link) per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons of this rank's GPU, polled through NVML from a thread every ~2 ms so that even a
    timed region of a few hundred ms is covered (nvidia-smi -lms cannot go that fast)."""

    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index=0):
        self.index, self.samples, self.stop_flag, self.thread, self.err = index, [], False, None, None

    def _run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = None
                self.samples.append((self.window, sm, rs, pw))
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def start(self):
        import threading

        self.window = False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def mark(self, on):
        self.window = on

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2.0)
        inwin = [x for x in self.samples if x[0]] or self.samples
        if not inwin:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML samples: %s" % self.err]}
        sm = sorted(x[1] for x in inwin)
        reasons = set()
        for _, _, rs, _ in inwin:
            for n, b in self.BITS.items():
                if rs & b:
                    reasons.add(n)
        pw = [x[3] for x in inwin if x[3] is not None]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": getattr(self, "smax", None), "reasons": sorted(reasons), "samples": len(inwin),
                "power_w_max": max(pw) if pw else None, "how": "NVML polled every 2 ms during the timed region"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port timed on the host cores (the reference is pure Julia with un-vendored
# dependencies and cannot be built here; DESIGN.md "Reference arm")
# ------------------------------------------------------------------------------------------------
def cpu_md_steps_per_s(dims_global, beta, tau, steps, warmup, budget_s):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gf_oracle as oracle

    cores = os.cpu_count() or 1
    v_full = 1
    for d in dims_global:
        v_full *= d
    # probe on a 16^3 x 8 sample to decide whether the full lattice fits the time budget
    probe = (16, 16, 16, 8)
    Up = oracle.hot_start_philox(probe, 1234)
    Pp = oracle.gaussian_momenta(probe, 0x5678, 0)
    t0 = time.time()
    oracle.md_step(Up, Pp, probe, beta, tau / max(steps, 1), 0)
    t_probe = time.time() - t0
    v_probe = 16 * 16 * 16 * 8
    est_full = t_probe * v_full / v_probe
    if est_full * (steps + warmup) <= budget_s:
        dims, scale, sample = tuple(dims_global), 1.0, "full lattice %s, %d+%d QPQ steps" % ("x".join(map(str, dims_global)), warmup, steps)
    else:
        # bounded sample: a sub-volume with the same arithmetic per site, throughput scaled by the site ratio
        dims = (16, 16, 16, 16)
        while 16 * 16 * 16 * dims[3] * 2 <= v_full and t_probe * (16 * 16 * 16 * dims[3] * 2) / v_probe * (steps + warmup) <= budget_s:
            dims = (16, 16, 16, dims[3] * 2)
        v_s = dims[0] * dims[1] * dims[2] * dims[3]
        scale = v_s / v_full
        sample = "sub-volume %s of %s (per-site work identical; steps/s scaled by %d/%d sites), %d+%d QPQ steps" % (
            "x".join(map(str, dims)), "x".join(map(str, dims_global)), v_s, v_full, warmup, steps)
    U = oracle.hot_start_philox(dims, 1234)
    P = oracle.gaussian_momenta(dims, 0x5678, 0)
    eps = tau / steps
    for _ in range(warmup):
        oracle.md_step(U, P, dims, beta, eps, 0)
    t0 = time.time()
    for _ in range(steps):
        oracle.md_step(U, P, dims, beta, eps, 0)
    dt = time.time() - t0
    value = steps / dt * scale
    return value, dt / steps * 1e3 / scale, {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}


def workload_name(args, dims_global, dims_local):
    return "SU(3) Wilson QPQ HMC, beta=%g, hot start seed 1234, global lattice %s, %s scaling (%s per GPU, t-slabs)" % (
        args.beta, "x".join(map(str, dims_global)), args.scaling, "x".join(map(str, dims_local)))


def run_reference(args, dims_global, dims_local, rank, world):
    if rank != 0:
        return
    value, ms, cb = cpu_md_steps_per_s(dims_global, args.beta, args.tau, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, dims_global, dims_local),
                   "note": "CPU oracle port of the reference's serial math, all host cores (OpenMP); the reference is Julia + un-vendored "
                           "LatticeMatrices.jl and cannot be built in this image (DESIGN.md section 7)"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nx, ny, nz, nt = [int(v) for v in args.lattice.split(",")]
    ngp = max(args.gpus, 1)
    if args.scaling == "weak":
        dims_global, tl = (nx, ny, nz, nt * ngp), nt
    else:
        if nt % ngp:
            raise SystemExit("NT must be divisible by the number of GPUs")
        dims_global, tl = (nx, ny, nz, nt), nt // ngp
    dims_local = (nx, ny, nz, tl)

    if args.impl == "reference":
        run_reference(args, dims_global, dims_local, rank, world)
        return

    import numpy as np
    import torch

    if args.gpus != world and world == 1 and args.gpus > 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import gfb200

    backend = gfb200.B200Backend(ngpu=1, devices=[local_rank], distributed=(world > 1))
    K, W = args.steps, max(args.warmup, 3)
    sites_global = dims_global[0] * dims_global[1] * dims_global[2] * dims_global[3]
    sites_local = nx * ny * nz * tl

    U = gfb200.gauge_configuration(dims_global, backend=backend, start="hot", seed=1234)
    P = gfb200.gaussian_momenta(U, seed=0x5678, sweep=0)
    loops = gfb200.make_loops_fromname("plaquette")
    action = gfb200.GaugeAction(U).push(args.beta / 2, loops + loops.adjoint())
    md = gfb200.md_driver(U, action, steps=K, trajectory_length=args.tau, integrator=gfb200.QPQ, fused=not args.unfused)
    md_w = gfb200.md_driver(U, action, steps=1, trajectory_length=args.tau / K, integrator=gfb200.QPQ, fused=not args.unfused)

    def barrier():
        backend.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        gfb200.md_trajectory_(U, P, md_w, diagnostics=False)  # one untimed MD step each
    barrier()
    # ---- timed region: exactly K MD steps --------------------------------------------------------
    n0 = backend.kernel_launches()
    barrier()
    sampler.mark(True)
    backend.tic()
    gfb200.md_trajectory_(U, P, md, diagnostics=False)
    ms = backend.toc()
    barrier()
    sampler.mark(False)
    launches = backend.kernel_launches() - n0
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    # ---- dominant kernel: average duration of the fused kick+drift launch --------------------------
    # the K-step trajectory is K fused passes plus one half-drift (update_links); time that one
    # alone and subtract, so achieved = algorithmic bytes per launch / average launch duration
    t_extra = 0.0
    if not args.unfused:
        reps = 5
        backend.tic()
        for _ in range(reps):
            gfb200.update_gaugefields_(U, P, 1e-9)
        t_extra = backend.toc() / reps
        barrier()
    peak, peak_src = measured_peak()
    if args.unfused:
        kern, kern_bytes = "k_update_links + k_force_fused<kick> + k_update_links (whole QPQ step)", (2 * (P_BYTES + 2 * U_BYTES) + U_BYTES + 2 * P_BYTES)
        kern_ms = ms / K
    else:
        tiled = nx % 8 == 0 and ny % 4 == 0 and nz % 2 == 0 and os.environ.get("GFB200_TMARCH", "1") != "0"
        if world > 1 and tl - 2 < 8 and os.environ.get("GFB200_TMARCH", "1") != "2":
            tiled = False  # short slab interiors stay in k_force_fused (csrc/tmarch.cu, launch_tmarch_fused)
        kern = ("k_tmarch_fused<READ_Z,WRITE_Z,DO_EXP>" if tiled else "k_force_fused<READ_Z,WRITE_Z,DO_EXP>") + " (staple->TA force->momentum kick->exp(eps P) U)"
        kern_bytes = 2 * U_BYTES + 2 * P_BYTES
        kern_ms = max(ms - t_extra, 1e-9) / K
    achieved = kern_bytes * sites_local / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            key = "k_tmarch_fused_bytes_per_site" if kern.startswith("k_tmarch") and "k_tmarch_fused_bytes_per_site" in tj else "k_force_fused_bytes_per_site"
            traffic = tj[key] * sites_local  # ncu dram bytes per site (measured at tj["lattice"]) x this launch's sites
        except Exception:
            traffic = None

    # ---- e2e: the md_trajectory! call with HOST buffers (pinned), H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        Uh = gfb200.pinned_empty(U.local_shape(), np.complex128)  # each rank holds its own t-slab on the host
        Ph = gfb200.pinned_empty(P.local_shape(), np.float64)
        U.to_host(Uh, local=True); P.to_host(Ph, local=True)
        md_e = gfb200.md_driver(U, action, steps=K, trajectory_length=args.tau, integrator=gfb200.QPQ, fused=not args.unfused)

        def traj_e2e():
            U.upload(Uh, local=True); P.upload(Ph, local=True)
            res = gfb200.md_trajectory_(U, P, md_e, diagnostics=True)
            U.to_host(Uh, local=True); P.to_host(Ph, local=True)
            return res

        traj_e2e()
        barrier()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            res = traj_e2e()
        barrier()
        dt = (time.perf_counter() - t0) / reps
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        bytes_dir = sites_local * (U_BYTES + P_BYTES) * world
        e2e = {"value": K / dt, "unit": UNIT, "h2d_bytes_per_step": bytes_dir / K, "d2h_bytes_per_step": (bytes_dir + 16) / K,
               "call": "upload U,P (pinned host, reference gathered layout) -> md_trajectory!(%d QPQ steps, diagnostics) -> download U,P" % K,
               "delta_hamiltonian": res.delta_hamiltonian}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, _, cpu_baseline = cpu_md_steps_per_s(dims_global, args.beta, args.tau, 2, 1, budget_s=25.0)

    if rank == 0:
        step_bytes = (2 * U_BYTES + 2 * P_BYTES) if not args.unfused else 3904
        line = {
            "metric": METRIC, "value": K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(args, dims_global, dims_local),
                "integrator": "QPQ, %s" % ("reference op sequence (link, kick, link)" if args.unfused else "fused kick+drift kernel, adjacent half drifts merged"),
                "l2": "inputs larger than L2 (links %.0f MB per GPU), no flush" % (sites_local * U_BYTES / 1e6),
                "link_updates_per_s": 4.0 * sites_global * K / (ms * 1e-3),
                "algorithmic_bytes_per_site_per_step": step_bytes,
                "hbm_roofline_frac_of_8TBs_per_gpu": step_bytes * sites_local / (ms / K * 1e-3) / 8e12,
            },
            "roofline": {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": kern_bytes * sites_local,
                         "note": "FP64 issue and shared-memory reads co-limit this kernel (about 1.5 k FP64 instructions and 141 LDS.128 per link); DESIGN.md section 3a, profiles/r1_tmarch.md"},
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
