"""Built-in vs @sde_model-generated functors on the headline workload shape (kernel ms, same seed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
S = _pkg.load()
S._lib.check(S._lib.lib().sdeb200_init(0))
n = 148 * 256 * 64
for name, params, x0, nst in (("Heston", (0.01, 10.0, 0.3, 0.1, 0.4), [100.0, 0.4], 1024), ("BlackScholes", (0.01, 0.3), 100.0, 512)):
    for kind, cls in (("built-in", getattr(S, name)), ("generated", S.sde_model(name + "G", S.REGISTRY[name]))):
        m = cls(*params)
        for prec in ("f64", "f32"):
            for rep in range(2):
                e = S.estimate(m, S.EulerMaruyama(1 / nst), S.Moments(), 0.0, x0, nst, n, seed=1, precision=prec)
            print(f"{name:13s} {kind:9s} {prec}: {S.last_kernel_ms():8.3f} ms  {n * nst / S.last_kernel_ms() / 1e6:8.1f} G path-steps/s  mean {e.mean[0]:.5f}")
