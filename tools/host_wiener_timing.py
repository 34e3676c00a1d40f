"""End-to-end time of simulate!(x, model, scheme, t0, x0, w) with x === w in PINNED HOST memory (the library stages
slabs through the device: host -> device, kernel, device -> host; two slabs in flight)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg, torch
S = _pkg.load()
S._lib.check(S._lib.lib().sdeb200_init(0))
n, rows = 1_000_000, 513
bs = S.BlackScholes(0.01, 0.3)
w = torch.empty((n, rows), dtype=torch.float64, pin_memory=True)
wd = torch.empty((n, rows), dtype=torch.float64, device="cuda")
S.wiener_(wd, 1 / 512, seed=1)
w.copy_(wd)
torch.cuda.synchronize()
for rep in range(3):
    w.copy_(wd); torch.cuda.synchronize()
    t = time.perf_counter()
    S.simulate_(w.numpy(), bs, S.EulerMaruyama(1 / 512), 0.0, 100.0, w.numpy())
    dt = time.perf_counter() - t
    print("host in place: %.1f ms, %.1f GB/s each way (%.1f GB in + %.1f GB out)" % (dt * 1e3, w.numel() * 8 / dt / 1e9, w.numel() * 8 / 1e9, w.numel() * 8 / 1e9))
S.simulate_(wd, bs, S.EulerMaruyama(1 / 512), 0.0, 100.0, wd)
print("matches the device-resident run:", bool(torch.equal(w, wd.cpu())))
