import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg, torch
S = _pkg.load()
S._lib.check(S._lib.lib().sdeb200_init(0))
lay = sys.argv[1] if len(sys.argv) > 1 else "reference"
n = 2_000_000
m, sch = S.BlackScholes(0.01, 0.3), S.Milstein(1 / 512)
out = torch.empty((n, 513) if lay == "reference" else (513, n), dtype=torch.float64, device="cuda")
for seed in range(2):
    S.simulate(m, sch, 0.0, 100.0, 512, n, seed=seed, out=out, layout=lay)
    print(lay, S.last_kernel_ms(), out.numel() * 8 / S.last_kernel_ms() / 1e6, "GB/s")
