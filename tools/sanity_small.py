"""Tiny, ragged invocations of every kernel family — run under compute-sanitizer (one tool per gpurun call)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg

S = _pkg.load()
S._lib.check(S._lib.lib().sdeb200_init(0))
bs, he = S.BlackScholes(0.01, 0.3), S.Heston(0.01, 10.0, 0.3, 0.1, 0.4)
gen = S.sde_model("CIRs", S.REGISTRY["CoxIngersollRoss"])(1.3, 0.09, 0.1)
tm = S.sde_model("TMs", "dX = Y*a*dt + b*dW1 + d*dW2\ndY = Z*c*dt + dW2 + dW3\ndZ = X*e*dt + f*dW3 - dW4")(1, 2, 3, 4, 5, 6)
for prec in ("f64", "f32"):
    for n, steps in ((1, 1), (33, 7), (1000, 37), (4097, 5)):
        for m, x0, schs in ((bs, 100.0, (S.EulerMaruyama, S.Milstein)), (he, [100.0, 0.4], (S.EulerMaruyama,)),
                            (gen, 0.09, (S.EulerMaruyama, S.Milstein)), (tm, [0.1, 0.2, 0.3], (S.EulerMaruyama,))):
            for sc in schs:
                sch = sc(1.0 / 64)
                a = S.sample(m, sch, 0.0, x0, steps, n, seed=1, precision=prec)
                x, _ = S.simulate(m, sch, 0.0, x0, steps, n, seed=1, precision=prec)
                xs, _ = S.simulate(m, sch, 0.0, x0, steps, n, seed=1, precision=prec, layout="soa")
                e = S.estimate(m, sch, S.Moments(), 0.0, x0, steps, n, seed=1, precision=prec)
                assert np.array_equal(np.asarray(x).reshape(n, steps + 1, -1)[:, -1, :], np.asarray(a).reshape(n, -1))
                assert e.n == n
                if m._D == max(m._M, 1):
                    w = np.zeros(x.shape, dtype=x.dtype)
                    S.wiener_(w, 1.0 / 64, seed=2)
                    S.simulate_(w, m, sch, 0.0, x0, w)
                    xc, xf = S.multilevel_sample_level(m, sc(1 / 16), sc(1 / 32), 2, 0.0, x0, max(steps // 2, 1), n, seed=3, precision=prec)
                    S.multilevel_simulate_level(m, sc(1 / 16), sc(1 / 64), 4, 0.0, x0, 3, n, seed=3, precision=prec)
    ml = S.MultilevelScheme(S.EulerMaruyama(1 / 8), 2, range(0, 4))
    S.ml_estimate(he, ml, S.Call(100.0, 0.99), 0.0, [100.0, 0.4], 8, [1000, 300, 77, 5], seed=4, precision=prec)
    S.sample(he, ml, 0.0, [100.0, 0.4], 8, [10, 9, 8, 7], seed=4, precision=prec)
print("sanity ok, launches:", S.launch_count())
