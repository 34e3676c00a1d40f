"""Short driver for ncu: one warm-up + one profiled launch of the headline kernel (Heston EM FP64 sample)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg

S = _pkg.load()
prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
npaths = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 256 * 64
kind = sys.argv[3] if len(sys.argv) > 3 else "sample"
S._lib.check(S._lib.lib().sdeb200_init(0))
if kind == "sample":
    m, sch = S.Heston(0.01, 10.0, 0.3, 0.1, 0.4), S.EulerMaruyama(1 / 1024)
    for seed in range(2):
        e = S.estimate(m, sch, S.Moments(), 0.0, [100.0, 0.4], 1024, npaths, seed=seed, precision=prec)
        print(seed, e.mean, S.last_kernel_ms(), npaths * 1024 / S.last_kernel_ms() / 1e6, "G path-steps/s")
elif kind == "bench":  # the bench.py step: fused sample + reduce WITH the terminal states kept in HBM
    import torch
    m, sch = S.Heston(0.01, 10.0, 0.3, 0.1, 0.4), S.EulerMaruyama(1 / 1024)
    term = torch.empty((npaths, 2), dtype=torch.float64 if prec == "f64" else torch.float32, device="cuda")
    for seed in range(2):
        e = S.estimate(m, sch, S.Moments(), 0.0, [100.0, 0.4], 1024, npaths, seed=seed, precision=prec, terminal=term)
        print(seed, e.mean, S.last_kernel_ms(), npaths * 1024 / S.last_kernel_ms() / 1e6, "G path-steps/s", float(term[:, 0].mean()))
elif kind == "wiener":  # Wiener-driven in-place trajectories: the HBM-bound entry point (one load + one store per element)
    import torch
    tdt = torch.float64 if prec == "f64" else torch.float32
    for name, m, x0, D in (("BlackScholes", S.BlackScholes(0.01, 0.3), 100.0, 1), ("Heston", S.Heston(0.01, 10.0, 0.3, 0.1, 0.4), [100.0, 0.4], 2)):
        n = npaths // D
        w = torch.empty((n, 513) if D == 1 else (n, 513, D), dtype=tdt, device="cuda")
        sch = S.EulerMaruyama(1 / 512)
        for rep in range(3):
            S.wiener_(w, 1 / 512, seed=rep)
            t_w = S.last_kernel_ms()
            S.simulate_(w, m, sch, 0.0, x0, w)
            ms = S.last_kernel_ms()
            nbytes = 2 * w.numel() * w.element_size()
            print(name, prec, rep, "wiener! %.3f ms (%.0f GB/s written)" % (t_w, nbytes / 2 / t_w / 1e6),
                  "simulate!(x===w) %.3f ms" % ms, "%.0f GB/s read+write" % (nbytes / ms / 1e6), "mean S_T %.4f" % float(w[:, -1].double().mean() if D == 1 else w[:, -1, 0].double().mean()))
else:  # full-path BlackScholes Milstein (C3 shape, fewer paths)
    import torch
    m, sch = S.BlackScholes(0.01, 0.3), S.Milstein(1 / 512)
    for lay in ("soa", "reference"):
        shape = (513, npaths) if lay == "soa" else (npaths, 513)
        out = torch.empty(shape, dtype=torch.float64 if prec == "f64" else torch.float32, device="cuda")
        for seed in range(2):
            S.simulate(m, sch, 0.0, 100.0, 512, npaths, seed=seed, precision=prec, out=out, layout=lay)
            ms = S.last_kernel_ms()
            print(lay, seed, ms, npaths * 512 / ms / 1e6, "G path-steps/s", out.numel() * out.element_size() / ms / 1e6, "GB/s")
