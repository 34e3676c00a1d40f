"""ncu driver: BlackScholes Milstein trajectories in the reference's order (C3 shape, 2e6 paths), FP32 and FP64"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg, torch
S = _pkg.load()
S._lib.check(S._lib.lib().sdeb200_init(0))
m, sch = S.BlackScholes(0.01, 0.3), S.Milstein(1 / 512)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
for prec, tdt in (("f32", torch.float32), ("f64", torch.float64)):
    out = torch.empty((n, 513), dtype=tdt, device="cuda")
    for seed in range(2):
        S.simulate(m, sch, 0.0, 100.0, 512, n, seed=seed, precision=prec, out=out)
        ms = S.last_kernel_ms()
        print(prec, seed, ms, out.numel() * out.element_size() / ms / 1e6, "GB/s")
    del out
