"""logtransition along stored trajectories (K7 / K7t): BlackScholes and Heston, FP64 and FP32, TMA-load form on / off.
Prints kernel ms and GB/s read+written (algorithmic bytes).  Usage: prof_logt.py [npaths] [only: bs64|bs32|he64|he32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg, torch
S = _pkg.load()
S._lib.check(S._lib.lib().sdeb200_init(0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
only = sys.argv[2] if len(sys.argv) > 2 else None
tmas = (True, False) if os.environ.get("LOGT_BOTH", "1") == "1" else (True,)
he, bs = S.Heston(0.01, 10.0, 0.3, 0.1, 0.4), S.BlackScholes(0.01, 0.3)
for tag, m, x0, D, prec in (("bs64", bs, 100.0, 1, "f64"), ("bs32", bs, 100.0, 1, "f32"), ("he64", he, [100.0, 0.4], 2, "f64"),
                            ("he32", he, [100.0, 0.4], 2, "f32")):
    if only and tag != only:
        continue
    tdt = torch.float64 if prec == "f64" else torch.float32
    np_ = n // D
    w = torch.empty((np_, 513) if D == 1 else (np_, 513, D), dtype=tdt, device="cuda")
    S.simulate(m, S.EulerMaruyama(1 / 512), 0.0, x0, 512, np_, seed=1, precision=prec, out=w)
    lt = torch.empty((np_, 512), dtype=tdt, device="cuda")
    nbytes = (w.numel() + lt.numel()) * w.element_size()
    for tma in tmas:
        S.set_tma(tma)
        best = None
        for rep in range(3):
            S.logtransition(m, S.EulerMaruyama(1 / 512), 0.0, w, out=lt)
            t = S.last_kernel_ms()
            best = t if best is None or (rep > 0 and t < best) else best
        print(tag, "tma" if tma else "plain", "%.3f ms" % best, "%.0f GB/s" % (nbytes / best / 1e6), "sum %.6e" % float(lt.double().sum()))
    S.set_tma(True)
    del w, lt
