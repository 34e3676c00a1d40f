"""Kernel time of small terminal-value launches (Heston EM, 1024 steps): the one-path-per-thread form against the
four-paths-per-thread form (SDEB200_SAMPLE1_MAX=0 python tools/small_launch_timing.py forces the latter)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
S = _pkg.load()
S._lib.check(S._lib.lib().sdeb200_init(0))
m, sch = S.Heston(0.01, 10.0, 0.3, 0.1, 0.4), S.EulerMaruyama(1 / 1024)
for prec in ("f64", "f32"):
    for n in (100000, 75776, 151552, 1000000):
        t = []
        for rep in range(8):
            S.estimate(m, sch, S.Moments(), 0.0, [100.0, 0.4], 1024, n, seed=rep, precision=prec)
            t.append(round(S.last_kernel_ms(), 3))
        print(prec, n, t)
