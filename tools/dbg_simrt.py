import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, _pkg
S = _pkg.load()
S._lib.check(S._lib.lib().sdeb200_init(0))
bs = S.BlackScholes(0.01, 0.3)
for (npaths, nrows, dtype) in [(257, 513, torch.float64), (4097, 65, torch.float64), (257, 513, torch.float32), (100000, 513, torch.float64)]:
    dt = 1.0 / (nrows - 1)
    for tag, fn in (("wiener1", lambda v: S.wiener_(v, dt, seed=8)),
                    ("bs-em", lambda v: S.simulate(bs, S.EulerMaruyama(dt), 0.0, 100.0, nrows - 1, npaths, seed=5, precision="f64" if dtype == torch.float64 else "f32", out=v))):
        for rep in range(3):
            res = []
            for tma in (True, False):
                S.set_tma(tma)
                v = torch.full((npaths, nrows), -7.0, dtype=dtype, device="cuda")
                fn(v)
                torch.cuda.synchronize()
                res.append(v.cpu().numpy())
            S.set_tma(True)
            bad = np.argwhere(res[0] != res[1])
            print(tag, npaths, nrows, dtype, "rep", rep, "mismatches", len(bad), "first", bad[:6].tolist(), "paths", sorted(set(bad[:, 0].tolist()))[:8], "rows", sorted(set(bad[:, 1].tolist()))[:12])
            if len(bad):
                i, j = bad[0]
                print("   tma", res[0][i, max(j - 2, 0):j + 3], "old", res[1][i, max(j - 2, 0):j + 3])
