"""C3-shaped full trajectories (BlackScholes Milstein) with a free number of steps: prof_c3.py <f64|f32> <npaths> <nsteps> [layouts]
Prints kernel ms and GB/s written per layout (the normal stream follows SDEB200_RNG)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg, torch
S = _pkg.load()
S._lib.check(S._lib.lib().sdeb200_init(0))
prec, npaths, nsteps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
lays = sys.argv[4].split(",") if len(sys.argv) > 4 else ["reference"]
m, sch = S.BlackScholes(0.01, 0.3), S.Milstein(1 / nsteps)
for lay in lays:
    shape = (nsteps + 1, npaths) if lay == "soa" else (npaths, nsteps + 1)
    out = torch.empty(shape, dtype=torch.float64 if prec == "f64" else torch.float32, device="cuda")
    best = None
    for seed in range(3):
        S.simulate(m, sch, 0.0, 100.0, nsteps, npaths, seed=seed, precision=prec, out=out, layout=lay)
        ms = S.last_kernel_ms()
        best = ms if best is None or (seed > 0 and ms < best) else best
    print(prec, lay, nsteps, "%.3f ms" % best, "%.0f GB/s" % (out.numel() * out.element_size() / best / 1e6), "%.1f G path-steps/s" % (npaths * nsteps / best / 1e6))
    del out
