#!/usr/bin/env python
"""other_kernels.py [lattice] -- one call of every path outside the Wilson MD / flow / stout benchmarks (general action, topological
charge, heatbath, overrelaxation, plaquette, Polyakov loop, kinetic energy) on one GPU, so that `ncu` can list their kernels
(scripts/gpu_ncu_other.sh) and `--time` prints device-synchronised wall times per call."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaugefields.jl_b200"))
import gfb200  # noqa: E402

dims = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "32,32,32,32").split(","))
timing = "--time" in sys.argv
backend = gfb200.B200Backend(ngpu=1)
U = gfb200.gauge_configuration(dims, backend=backend, start="hot", seed=1234)
gfb200.flow_(U, gfb200.gradient_flow(U, steps=2, step_size=0.02))  # a smoother field: the heatbath accepts, the exponentials are typical
P = gfb200.gaussian_momenta(U, seed=0x5678, sweep=0)


def run(name, fn, reps=3):
    fn()
    backend.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    backend.sync()
    if timing:
        print("%-44s %8.3f ms" % (name, (time.perf_counter() - t0) / reps * 1e3), flush=True)


plaq = gfb200.make_loops_fromname("plaquette")
rect = gfb200.make_loops_fromname("rectangular")
symanzik = gfb200.GaugeAction(U).push(6.0 / 2 * (1 + 8 / 12.0), plaq + plaq.adjoint()).push(-6.0 / 2 / 12.0, rect + rect.adjoint())
F = gfb200.gauge_momenta(U)
run("md_force (Symanzik: plaquette + rectangle)", lambda: gfb200.md_force_(F, symanzik, U))
md = gfb200.md_driver(U, symanzik, steps=2, trajectory_length=0.02, integrator=gfb200.QPQ, fused=True)
run("md_trajectory, 2 steps (Symanzik)", lambda: gfb200.md_trajectory_(U, P, md, diagnostics=False))
run("md_hamiltonian (Symanzik)", lambda: gfb200.md_hamiltonian(U, P, md))
gen = gfb200.Gradientflow_general(U, ["plaquette", "rectangular"], [1 + 8 / 12.0, -1 / 12.0], Nflow=1, eps=0.01)
run("flow step, Gradientflow_general (Symanzik)", lambda: gfb200.flow_(U, gen))
for m in ("plaquette", "clover", "improved"):
    run("topological_charge(%s)" % m, lambda m=m: gfb200.topological_charge(U, method=m))
h = gfb200.Heatbath(U, 6.0, seed=7)
run("heatbath sweep", lambda: gfb200.heatbath_(U, h))
run("overrelaxation sweep", lambda: gfb200.overrelaxation_(U, h))
run("calculate_Plaquette", lambda: gfb200.calculate_Plaquette(U))
run("Polyakov loop", lambda: gfb200.measure_polyakov_loop(U))
run("energy_density(clover)", lambda: gfb200.energy_density(U))
run("kinetic energy p*p", lambda: P.dot())
run("gaussian_momenta", lambda: gfb200.gaussian_momenta_(P, seed=1, sweep=1))
backend.finalize()
