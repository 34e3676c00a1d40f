"""One warm-up + one profiled launch of a secondary kernel (for ncu): mode in {sample32, simsoa64, simsoa32, simref64, simref32, simw64, mlmc64, logt64}.
The normal stream follows SDEB200_RNG (0/1 = v1, 2 = v2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg, torch
S = _pkg.load()
S._lib.check(S._lib.lib().sdeb200_init(0))
mode = sys.argv[1]
he, bs = S.Heston(0.01, 10.0, 0.3, 0.1, 0.4), S.BlackScholes(0.01, 0.3)
for rep in range(2):
    if mode == "sample32":
        n = 148 * 256 * 64
        S.estimate(he, S.EulerMaruyama(1 / 1024), S.Moments(), 0.0, [100.0, 0.4], 1024, n, seed=rep, precision="f32")
        work = n * 1024
    elif mode in ("simsoa64", "simsoa32"):
        n, prec = 2_000_000, ("f64" if mode.endswith("64") else "f32")
        out = torch.empty((513, n), dtype=torch.float64 if prec == "f64" else torch.float32, device="cuda")
        S.simulate(bs, S.Milstein(1 / 512), 0.0, 100.0, 512, n, seed=rep, precision=prec, out=out, layout="soa")
        work = n * 512
    elif mode in ("simref64", "simref32"):
        n, prec = 2_000_000, ("f64" if mode.endswith("64") else "f32")
        out = torch.empty((n, 513), dtype=torch.float64 if prec == "f64" else torch.float32, device="cuda")
        S.simulate(bs, S.Milstein(1 / 512), 0.0, 100.0, 512, n, seed=rep, precision=prec, out=out)
        work = n * 512
    elif mode == "logt64":
        n = 2_000_000
        w = torch.empty((n, 513), dtype=torch.float64, device="cuda")
        S.simulate(bs, S.EulerMaruyama(1 / 512), 0.0, 100.0, 512, n, seed=rep, out=w)
        lt = torch.empty((n, 512), dtype=torch.float64, device="cuda")
        S.logtransition(bs, S.EulerMaruyama(1 / 512), 0.0, w, out=lt)
        work = n * 512
    elif mode == "simw64":
        n = 2_000_000
        w = torch.empty((n, 513), dtype=torch.float64, device="cuda")
        S.wiener_(w, 1 / 512, seed=rep)
        S.simulate_(w, bs, S.EulerMaruyama(1 / 512), 0.0, 100.0, w)
        work = n * 512
    elif mode == "mlmc64":
        n = 148 * 128 * 16
        S.multilevel_sample_level(he, S.EulerMaruyama(1 / 256), S.EulerMaruyama(1 / 512), 2, 0.0, [100.0, 0.4], 256, n, seed=rep, device=True)
        work = n * 512
    print(mode, rep, S.last_kernel_ms(), work / S.last_kernel_ms() / 1e6, "G path-steps/s")
