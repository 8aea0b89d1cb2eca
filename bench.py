#!/usr/bin/env python
"""bench.py -- the SU(3) Wilson update loop on B200: MD steps/s (BASELINE.json metric) with roofline, CPU baseline, e2e.

    python bench.py --gpus 1 --steps 20 --warmup 3                       # default workload md64 (BASELINE.json configs[4])
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W     # CPU arm (oracle port)
    python bench.py --workload flow32  [--gpus N]                        # configs[2]: 32^4 RK3 Wilson flow with E(t) every step
    python bench.py --workload stout48 [--gpus N]                        # configs[3]: 48^3x96 HMC, 2-layer stout-smeared Wilson action
    python bench.py --workload md16                                      # configs[1]: 16^4 beta = 6.0 (L2-resident, no roofline claim)

Workloads (a "step" is one pass of the hot path over the whole lattice):
  md64 / md16  one QPQ molecular-dynamics step (md_step!, src/molecular_dynamics.jl:611-616) of the quenched Wilson action.
               The timed region is ONE gfb_md_trajectory call of exactly K steps (diagnostics off).
  flow32       one Luescher RK3 flow step (flow!, src/smearing/gradientflow.jl:171-238) followed by the clover E(t)
               measurement (samples/measurements/energydensity.jl:4-78), eps = 0.01.
  stout48      one QPQ MD step of HMC with the action evaluated on 2-layer stout-smeared links, rho = 0.1, composed exactly as
               the reference's user script does (test/HMCstout_test_nowing.jl:99-118): calc_smearedU, calc_dSdUmu, back_prop,
               momentum kick, link update.
All are device-timed with CUDA events on the library's compute stream, bracketed by barrier + device synchronize, max over
ranks.  Synthetic hot start (seed 1234), Gaussian momenta (seed 0x5678); the global lattice is split into t-slabs over the N
GPUs (strong scaling); `--scaling weak --lattice X,Y,Z,T` keeps X*Y*Z*T sites PER GPU.  The fields are larger than L2 (126 MB)
except for md16, so no flush is needed (md16 says so in its config and makes no roofline claim).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gaugefields.jl_b200"))

U_BYTES, P_BYTES = 576, 256  # per site

WORKLOADS = {
    # name: (default lattice, default beta, metric, unit)
    "md64": ("64,64,64,64", 6.2, "SU(3) Wilson HMC MD steps/s", "MD steps/s"),
    "md16": ("16,16,16,16", 6.0, "SU(3) Wilson HMC MD steps/s", "MD steps/s"),
    "flow32": ("32,32,32,32", 6.0, "SU(3) RK3 Wilson flow steps/s (with E(t))", "flow steps/s"),
    "stout48": ("48,48,48,96", 6.0, "SU(3) stout-smeared (rho=0.1, 2 layers) Wilson HMC MD steps/s", "MD steps/s"),
}
FLOW_EPS, STOUT_RHO, STOUT_LAYERS = 0.01, 0.1, 2
# algorithmic bytes per site of one step as executed: each field element moved once per kernel
MD_STEP_BYTES = 2 * U_BYTES + 2 * P_BYTES                      # fused kick+drift: U r/w, P r/w = 1664
FLOW_STEP_BYTES = (2 * U_BYTES + P_BYTES) + (2 * U_BYTES + 2 * P_BYTES) + (2 * U_BYTES + P_BYTES)  # 3 fused stages = 4480
STOUT_FWD_BYTES = 2 * U_BYTES                                   # per layer, no tape
STOUT_BWD_BYTES = 8 * U_BYTES                                   # per layer: (U, d_out -> Lambda, d_in) + (U, Lambda, d_in -> d_in)
STOUT_STEP_BYTES = (STOUT_LAYERS * STOUT_FWD_BYTES + 2 * U_BYTES + STOUT_LAYERS * STOUT_BWD_BYTES + (2 * U_BYTES + 2 * P_BYTES)
                    + (2 * U_BYTES + P_BYTES))                  # forward, dSdU, back_prop, kick, merged link update


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="md64", choices=sorted(WORKLOADS))
    ap.add_argument("--lattice", default=None, help="NX,NY,NZ,NT: the global lattice (strong) or the per-GPU lattice (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--beta", type=float, default=None)
    ap.add_argument("--tau", type=float, default=1.0)
    ap.add_argument("--unfused", action="store_true", help="md*: issue the reference's op sequence (link, kick, link) per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    args = ap.parse_args()
    lat, beta, _, _ = WORKLOADS[args.workload]
    if args.lattice is None:
        args.lattice = lat
    if args.beta is None:
        args.beta = beta
    return args


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
# CPU arm: the oracle port timed on the host cores.  The reference is pure Julia with un-vendored
# dependencies and cannot be built here, so kind = "port".
# ------------------------------------------------------------------------------------------------
def _oracle(threads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gf_oracle as oracle

    oracle.build()
    # torch.distributed.run exports OMP_NUM_THREADS=1; the oracle's thread count is set explicitly and read back
    return oracle, oracle.threads(threads)


def _oracle_step(oracle, workload, dims, beta, eps):
    """One step of `workload` on the oracle: returns a closure over freshly initialised fields."""
    U = oracle.hot_start_philox(dims, 1234)
    if workload == "flow32":
        def step():
            oracle.flow_step(U, dims, FLOW_EPS)
            oracle.energy_density_clover(U, dims)
        return step
    P = oracle.gaussian_momenta(dims, 0x5678, 0)
    if workload == "stout48":
        def step():
            U[...] = oracle.update_links(U, P, dims, eps / 2)
            tape, cur = [], U
            for _ in range(STOUT_LAYERS):
                tape.append(cur)
                cur = oracle.stout_forward(cur, dims, STOUT_RHO)
            d = oracle.wilson_dSdU(cur, dims, beta)
            for inp in reversed(tape):
                d = oracle.stout_backward(d, inp, dims, STOUT_RHO)
            oracle.kick_from_dSdU(P, U, d, dims, -eps / 3.0)
            U[...] = oracle.update_links(U, P, dims, eps / 2)
        return step
    return lambda: oracle.md_step(U, P, dims, beta, eps, 0)


def cpu_steps_per_s(workload, dims_global, beta, tau, steps, warmup, budget_s):
    """`steps` timed steps (after `warmup`) of the oracle on a BOUNDED sample of the workload: the full lattice when
    (steps + warmup) of them fit `budget_s`, otherwise a t-sub-volume NX x NY x NZ x T_s (identical arithmetic per site) whose
    throughput is scaled by the site ratio -- stated in `sample` and flagged by `same_config`."""
    cores = os.cpu_count() or 1
    oracle, threads = _oracle(cores)
    nx, ny, nz, nt = dims_global
    v_full = nx * ny * nz * nt
    eps = tau / max(steps, 1)
    probe = (16, 16, 16, 4)
    t0 = time.time()
    _oracle_step(oracle, workload, probe, beta, eps)()
    per_site = (time.time() - t0) / (16 * 16 * 16 * 4)
    n = steps + warmup
    if per_site * v_full * n <= budget_s:
        dims = tuple(dims_global)
    else:
        ts = nt
        while ts > 2 and per_site * nx * ny * nz * ts * n > budget_s:
            ts //= 2
        dims = (nx, ny, nz, max(ts, 2))
        while per_site * dims[0] * dims[1] * dims[2] * dims[3] * n > budget_s and dims[0] > 8:
            dims = (dims[0] // 2, dims[1] // 2, dims[2] // 2, dims[3])
    v_s = dims[0] * dims[1] * dims[2] * dims[3]
    same = v_s == v_full
    step = _oracle_step(oracle, workload, dims, beta, eps)
    for _ in range(warmup):
        step()
    t0 = time.time()
    for _ in range(steps):
        step()
    dt = time.time() - t0
    scale = v_s / v_full
    value = steps / dt * scale
    sample = ("full lattice %s, %d+%d steps" % ("x".join(map(str, dims_global)), warmup, steps)) if same else (
        "sub-volume %s of %s (identical per-site arithmetic; steps/s scaled by %d/%d sites), %d+%d steps, %.1f s of CPU time" % (
            "x".join(map(str, dims)), "x".join(map(str, dims_global)), v_s, v_full, warmup, steps, dt))
    cb = {"value": value, "unit": WORKLOADS[workload][3], "cores": threads, "kind": "port", "sample": sample, "same_config": same,
          "sample_ms_per_step": dt / steps * 1e3, "sample_sites": v_s,
          "how": "oracle/gf_oracle.cpp (C++17/OpenMP restatement of the reference's serial math), %d OpenMP threads of %d host cores" % (threads, cores)}
    return value, dt / steps * 1e3 / scale, cb


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned e2e buffers are
    allocated: cudaHostAlloc places pages on the node of the allocating thread, and with eight ranks on a two-socket box
    unbound ranks send half of their H2D/D2H traffic across the socket link (round 1: 15 GB/s per rank at N = 8 against
    52 GB/s at N = 1).  Returns what it did for the JSON line; never fatal."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return {"numa_node": None, "note": "the platform reports no NUMA affinity for %s" % bus}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"numa_node": node, "note": "no allowed CPU on that node"}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001
        return {"numa_node": None, "note": repr(e)[:120]}


def workload_name(args, dims_global, dims_local):
    what = {"md64": "SU(3) Wilson QPQ HMC", "md16": "SU(3) Wilson QPQ HMC", "flow32": "SU(3) RK3 Wilson flow (eps=%g) + clover E(t) per step" % FLOW_EPS,
            "stout48": "SU(3) QPQ HMC, Wilson action on %d-layer stout links (rho=%g)" % (STOUT_LAYERS, STOUT_RHO)}[args.workload]
    return "%s, beta=%g, hot start seed 1234, global lattice %s, %s scaling (%s per GPU, t-slabs)" % (
        what, args.beta, "x".join(map(str, dims_global)), args.scaling, "x".join(map(str, dims_local)))


def run_reference(args, dims_global, dims_local, rank, world):
    if rank != 0:
        return
    _, _, metric, unit = WORKLOADS[args.workload]
    value, ms, cb = cpu_steps_per_s(args.workload, dims_global, args.beta, args.tau, args.steps, args.warmup, budget_s=120.0)
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, dims_global, dims_local),
                   "note": "CPU oracle port of the reference's serial math on all host cores (OpenMP); the reference is Julia + un-vendored "
                           "LatticeMatrices.jl and cannot be built in this image (DESIGN.md section 7).  value/ms_per_step are the sample's "
                           "throughput scaled to the full lattice when cpu_baseline.same_config is false"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# multi-GPU parity check (outside the timed region): the same slab code path as the benchmark on a small lattice the oracle
# finishes in a blink; every rank compares ITS slab, the worst deviation over ranks is printed in the JSON line
# ------------------------------------------------------------------------------------------------
def parity_check(gfb200, backend, world, dist, torch):
    import numpy as np

    oracle, _ = _oracle(2)
    dims, beta = (8, 4, 4, 4 * world), 5.7
    Uh = oracle.hot_start_philox(dims, 1234)
    for _ in range(2):
        oracle.flow_step(Uh, dims, 0.02)
    Ph = oracle.gaussian_momenta(dims, 0x5678, 2)
    t0, t1 = backend.t_range(dims[3])
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    P = gfb200.gauge_momenta(U).upload(Ph)
    loops = gfb200.make_loops_fromname("plaquette")
    action = gfb200.GaugeAction(U).push(beta / 2, loops + loops.adjoint())
    F = gfb200.gauge_momenta(U)
    gfb200.md_force_(F, action, U)
    Fw = oracle.force(Uh, dims, beta)
    e_force = float(np.abs(F.to_host(local=True) - Fw[:, t0:t1]).max() / np.abs(Fw).max())
    plaq, plaq_w = gfb200.calculate_Plaquette(U), oracle.plaquette_sum(Uh, dims)
    md = gfb200.md_driver(U, action, steps=6, trajectory_length=0.3, integrator=gfb200.QPQ, fused=True)
    res = gfb200.md_trajectory_(U, P, md)
    Uo, Po = Uh.copy(), Ph.copy()
    H0, H1 = oracle.md_trajectory(Uo, Po, dims, beta, 6, 0.3, 0)
    e_links = float(np.abs(U.to_host(local=True) - Uo[:, t0:t1]).max())
    U.upload(Uh)
    gfb200.flow_(U, gfb200.gradient_flow(U, steps=2, step_size=0.01))
    Uf = Uh.copy()
    for _ in range(2):
        oracle.flow_step(Uf, dims, 0.01)
    e_flow = float(np.abs(U.to_host(local=True) - Uf[:, t0:t1]).max())
    ee, ee_w = gfb200.energy_density(U), oracle.energy_density_clover(Uf, dims)
    errs = [e_force, abs(plaq - plaq_w) / abs(plaq_w), abs(res.delta_hamiltonian - (H1 - H0)), e_links, e_flow, abs(ee - ee_w) / max(1.0, abs(ee_w))]
    if dist is not None:
        t = torch.tensor(errs, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        errs = [float(v) for v in t.tolist()]
    tol = [1e-12, 1e-12, 1e-9, 1e-11, 1e-12, 1e-11]
    names = ["force_rel", "plaquette_rel", "delta_H_abs", "links_after_trajectory_abs", "links_after_flow_abs", "energy_density_rel"]
    return {"lattice": "x".join(map(str, dims)), "ranks": world, "checker": "oracle/gf_oracle.cpp", "max_over_ranks": dict(zip(names, errs)),
            "tolerance": dict(zip(names, tol)), "ok": all(e <= t for e, t in zip(errs, tol))}


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
    _, _, metric, unit = WORKLOADS[args.workload]

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

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    backend = gfb200.B200Backend(ngpu=1, devices=[local_rank], distributed=(world > 1))
    K, W = args.steps, max(args.warmup, 3)
    sites_global = dims_global[0] * dims_global[1] * dims_global[2] * dims_global[3]
    sites_local = nx * ny * nz * tl
    wl = args.workload
    is_md = wl in ("md64", "md16")

    U = gfb200.gauge_configuration(dims_global, backend=backend, start="hot", seed=1234)
    P = gfb200.gaussian_momenta(U, seed=0x5678, sweep=0) if wl != "flow32" else None
    loops = gfb200.make_loops_fromname("plaquette")
    action = gfb200.GaugeAction(U).push(args.beta / 2, loops + loops.adjoint())
    eps = args.tau / K

    def barrier():
        backend.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the step of each workload -----------------------------------------------------------------
    energies = []
    if is_md:
        md = gfb200.md_driver(U, action, steps=K, trajectory_length=args.tau, integrator=gfb200.QPQ, fused=not args.unfused)
        md_w = gfb200.md_driver(U, action, steps=1, trajectory_length=args.tau / K, integrator=gfb200.QPQ, fused=not args.unfused)

        def warm():
            gfb200.md_trajectory_(U, P, md_w, diagnostics=False)

        def timed():
            gfb200.md_trajectory_(U, P, md, diagnostics=False)
    elif wl == "flow32":
        g1 = gfb200.gradient_flow(U, steps=1, step_size=FLOW_EPS)

        def warm():
            gfb200.flow_(U, g1)
            gfb200.energy_density(U)

        def timed():
            for _ in range(K):
                gfb200.flow_(U, g1)
                energies.append(gfb200.energy_density(U))
    else:
        smearing = gfb200.stout_smearing(U, rho=STOUT_RHO, layers=STOUT_LAYERS)
        ws = gfb200.StoutWorkspace(U, smearing)

        def stout_step():
            # QPQ with adjacent half drifts merged over the trajectory, like the fused Wilson trajectory
            gfb200.stout_force_(P, U, action, smearing, eps, workspace=ws)
            gfb200.update_gaugefields_(U, P, eps)

        def warm():
            stout_step()

        def timed():
            gfb200.update_gaugefields_(U, P, eps / 2)
            for k in range(K):
                gfb200.stout_force_(P, U, action, smearing, eps, workspace=ws)
                gfb200.update_gaugefields_(U, P, eps if k + 1 < K else eps / 2)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        warm()
    barrier()
    # ---- timed region: exactly K steps ---------------------------------------------------------------
    n0 = backend.kernel_launches()
    barrier()
    sampler.mark(True)
    backend.tic()
    timed()
    ms = backend.toc()
    barrier()
    sampler.mark(False)
    launches = backend.kernel_launches() - n0
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel: average launch duration measured live with CUDA events on the compute stream ----------
    peak, peak_src = measured_peak()

    def timed_ms(fn, reps):
        backend.tic()
        for _ in range(reps):
            fn()
        t = backend.toc() / reps
        barrier()
        return t

    extra = {}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    tj = {}
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
        except Exception:
            tj = {}
    if is_md:
        if args.unfused:
            kern, kern_bytes, kern_ms = "k_update_links + fused kick + k_update_links (whole QPQ step)", 3904, ms / K
        else:
            t_extra = timed_ms(lambda: gfb200.update_gaugefields_(U, P, 1e-9), 5)  # the trajectory's single half drift
            tiled = nx % 8 == 0 and ny % 4 == 0 and nz % 2 == 0 and os.environ.get("GFB200_TMARCH", "1") != "0"
            kern = ("k_tmarch_ws<READ_Z,WRITE_Z,DO_EXP>" if tiled else "k_force_fused<READ_Z,WRITE_Z,DO_EXP>") + " (staple->TA force->momentum kick->exp(eps P) U)"
            kern_bytes, kern_ms = MD_STEP_BYTES, max(ms - t_extra, 1e-9) / K
            key = "k_tmarch_ws" if tiled else "k_force_fused"
            ent = tj.get(key, {}).get("x".join(map(str, dims_local)))
            if ent:
                traffic = ent["dram_bytes_per_site"] * sites_local
                extra["traffic_source"] = ent.get("source")
        step_bytes = MD_STEP_BYTES if not args.unfused else 3904
    elif wl == "flow32":
        gK = gfb200.gradient_flow(U, steps=K, step_size=FLOW_EPS)
        t_flow = timed_ms(lambda: gfb200.flow_(U, gK), 1) / K
        kern = "k_tmarch_ws<0,1,0> + <1,1,1> + <1,1,1> (the three RK3 stages: staple->TA->Z update->exp(c Z) U)"
        kern_bytes, kern_ms = FLOW_STEP_BYTES, t_flow
        extra["energy_density_ms"] = ms / K - t_flow
        step_bytes = FLOW_STEP_BYTES + U_BYTES
    else:
        parts = gfb200.stout_force_(P, U, action, smearing, 1e-12, workspace=ws, timer=lambda fn: timed_ms(fn, 1))
        t_link = timed_ms(lambda: gfb200.update_gaugefields_(U, P, 1e-12), 3)
        parts["link_update"] = t_link
        extra["parts_ms"] = parts
        kern = "k_stout_lambda + k_stout_backward (back_prop through one stout layer)"
        kern_bytes, kern_ms = STOUT_BWD_BYTES, parts["back_prop"] / STOUT_LAYERS
        step_bytes = STOUT_STEP_BYTES
    achieved = kern_bytes * sites_local / (kern_ms * 1e-3) / 1e9

    # ---- e2e: the same call with HOST buffers (pinned), H2D + D2H inside the timed region ---------------------------
    e2e = None
    if not args.no_e2e:
        Uh = gfb200.pinned_empty(U.local_shape(), np.complex128)  # each rank holds its own t-slab on the host
        U.to_host(Uh, local=True)
        Ph = None
        if P is not None:
            Ph = gfb200.pinned_empty(P.local_shape(), np.float64)
            P.to_host(Ph, local=True)
        if is_md:
            md_e = gfb200.md_driver(U, action, steps=K, trajectory_length=args.tau, integrator=gfb200.QPQ, fused=not args.unfused)

        def traj_e2e():
            U.upload(Uh, local=True)
            if P is not None:
                P.upload(Ph, local=True)
            if is_md:
                res = gfb200.md_trajectory_(U, P, md_e, diagnostics=True).delta_hamiltonian
            elif wl == "flow32":
                for _ in range(K):
                    gfb200.flow_(U, g1)
                    res = gfb200.energy_density(U)
            else:
                h0 = gfb200.stout_hamiltonian(U, P, action, smearing, workspace=ws)
                timed()
                res = gfb200.stout_hamiltonian(U, P, action, smearing, workspace=ws) - h0
            U.to_host(Uh, local=True)
            if P is not None:
                P.to_host(Ph, local=True)
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
        # where the end-to-end time goes (one more repetition, synchronised between the phases; not part of the reported rate)
        def phase(fn):
            t1 = time.perf_counter()
            fn()
            backend.sync()
            return (time.perf_counter() - t1) * 1e3

        def up():
            U.upload(Uh, local=True)
            if P is not None:
                P.upload(Ph, local=True)

        def down():
            U.to_host(Uh, local=True)
            if P is not None:
                P.to_host(Ph, local=True)

        breakdown = {"upload_ms": phase(up), "download_ms": phase(down)}
        barrier()
        bytes_dir = sites_local * (U_BYTES + (P_BYTES if P is not None else 0)) * world
        e2e = {"value": K / dt, "unit": unit, "h2d_bytes_per_step": bytes_dir / K, "d2h_bytes_per_step": (bytes_dir + 16) / K,
               "call": "upload fields (pinned host, reference gathered layout) -> %d steps with diagnostics -> download fields" % K,
               "result": res, "host_binding_rank0": numa, "phases_rank0": breakdown,
               "pcie_gbs_rank0": {k[:-3]: sites_local * (U_BYTES + (P_BYTES if P is not None else 0)) / (v * 1e-3) / 1e9 for k, v in breakdown.items()}}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, _, cpu_baseline = cpu_steps_per_s(wl, dims_global, args.beta, args.tau, 2, 1, budget_s=20.0)

    parity = None
    if world > 1 and not args.no_parity_check:
        parity = parity_check(gfb200, backend, world, dist, torch)

    if rank == 0:
        if is_md and not args.unfused and kern_ms and clocks and clocks.get("sm_mhz"):
            # The fused pass is bound by FP64 issue, not by HBM: 1506 FP64 warp instructions per link and launch
            # (ncu, profiles/r2_ncu_tmarch_ws_64x4.csv: 197.4 M per 32^4 pass) against one FP64 warp instruction per 2 cycles per SM
            # sub-partition (148 SMs x 4).  Informative: the contract's roofline above stays the HBM one.
            fp64_instr = 1506.0 * 4.0 * sites_local / 32.0
            fp64_peak = 148 * 4 * clocks["sm_mhz"] * 1e6 / 2.0
            extra = dict(extra, fp64_pipe={"warp_instructions_per_launch": fp64_instr, "peak_warp_instructions_per_s": fp64_peak,
                                           "frac": fp64_instr / (kern_ms * 1e-3) / fp64_peak,
                                           "hbm_frac_at_full_fp64_pipe": (kern_bytes * sites_local / (fp64_instr / fp64_peak)) / 1e9 / peak})
        line = {
            "metric": metric, "value": K / (ms * 1e-3), "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(args, dims_global, dims_local),
                "integrator": ("QPQ, %s" % ("reference op sequence (link, kick, link)" if args.unfused else "fused kick+drift kernel, adjacent half drifts merged")) if is_md
                              else ("RK3 (W1, W2, W3 stages fused with their exponentials)" if wl == "flow32" else "QPQ, adjacent half drifts merged; smearing recomputed every force"),
                "l2": ("inputs larger than L2 (links %.0f MB per GPU), no flush" % (sites_local * U_BYTES / 1e6)) if sites_local * U_BYTES > 2.6e8
                      else "links %.0f MB per GPU fit L2: no roofline claim for this configuration" % (sites_local * U_BYTES / 1e6),
                "link_updates_per_s": 4.0 * sites_global * K / (ms * 1e-3),
                "algorithmic_bytes_per_site_per_step": step_bytes,
                "hbm_roofline_frac_of_8TBs_per_gpu": step_bytes * sites_local / (ms / K * 1e-3) / 8e12,
                "halo": os.environ.get("GFB200_HALO", "peer stores inside the fused kernels (NCCL send/recv when peer access is unavailable)") if world > 1 else None,
            },
            "roofline": dict({"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                              "traffic": traffic, "peak_source": peak_src, "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": kern_bytes * sites_local,
                              "note": "FP64 issue and shared-memory reads co-limit the fused kernel (about 1.5 k FP64 instructions and 137 LDS.128 per link); DESIGN.md section 3a, profiles/r2_tmarch.md"},
                             **extra),
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
        }
        if parity is not None:
            line["parity_check"] = parity
        if energies:
            line["config"]["E_first_last"] = [energies[0], energies[-1]]
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
