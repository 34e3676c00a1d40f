"""Kernel time of the C4 workload (Heston MLMC call, levels 0:8, N_l = 2^(26-2l), all levels concurrent)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
S = _pkg.load()
S._lib.check(S._lib.lib().sdeb200_init(0))
m = S.Heston(0.01, 10.0, 0.3, 0.1, 0.4)
ml = S.MultilevelScheme(S.EulerMaruyama(1 / 16), 2, range(0, 9))
nl = [2 ** (26 - 2 * l) for l in range(9)]
t = []
for rep in range(5):
    e = S.ml_estimate(m, ml, S.Call(100.0, 0.99004983374916811), 0.0, [100.0, 0.4], 16, nl, seed=rep)
    t.append(round(S.last_kernel_ms(), 3))
print("C4 ms per estimate:", t, "estimate %.4f +- %.4f" % (e.mean, e.stderr))
