#!/usr/bin/env python
"""Generate sdemodels.jl_b200/csrc/sde_coeffs.inc: polynomial coefficients and the log table used by the
fused Philox -> normal transform (sde_device.cuh).  Run here (needs mpmath); the output is committed.

  sin(2 pi x) = x * S(x^2),  cos(2 pi x) = C(x^2)      for |x| <= 1/8   (x in turns)
  -2 log1p(r) = r * Q(r)                                for |r| <= 2^-8 (1 + 2^-7)
  log table: c_j = 1 + (j + 1/2)/128, inv_c[j] = rn(1/c_j), tl[j] = rn(2 ln inv_c[j] + 2^-50)   j = 0..127

Fits are Chebyshev-node interpolants (near-minimax); reported errors are of the double-rounded
coefficients evaluated in 60-digit arithmetic.
"""
import os
import sys

import mpmath as mp

mp.mp.dps = 60


def cheb_fit(f, a, b, n):
    """Degree n-1 interpolant of f on [a,b] at n Chebyshev nodes; returns monomial coefficients."""
    nodes = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, t in enumerate(nodes):
        for j in range(n):
            A[i, j] = t ** j
        y[i] = f(t)
    c = mp.lu_solve(A, y)
    return [c[i] for i in range(n)]


def rnd(c):
    return [mp.mpf(float(v)) for v in c]


def poly(c, t):
    r = mp.mpf(0)
    for v in reversed(c):
        r = r * t + v
    return r


def max_err(f, c, a, b, rel=False, n=4001):
    e = mp.mpf(0)
    for k in range(n):
        t = a + (b - a) * k / (n - 1)
        ft = f(t)
        d = abs(poly(c, t) - ft)
        if rel:
            d = d / abs(ft)
        e = max(e, d)
    return e


def hexf(v):
    return float(v).hex()


def main():
    out = []
    U = mp.mpf(1) / 64  # x^2 <= 1/64

    def fs(u):
        if u == 0:
            return 2 * mp.pi
        x = mp.sqrt(u)
        return mp.sin(2 * mp.pi * x) / x

    def fc(u):
        return mp.cos(2 * mp.pi * mp.sqrt(u))

    report = []
    kidx = []
    for name, f, n, rel in (("SIN", fs, 7, True), ("COS", fc, 7, False)):
        c = rnd(cheb_fit(f, mp.mpf(0), U, n))
        e = max_err(f, c, mp.mpf(0), U, rel)
        report.append(f"{name}: {n} coefficients, max {'rel' if rel else 'abs'} err {mp.nstr(e, 3)}")
        out.append(f"// {report[-1]}")
        out.append(f"#define SDE_{name}_N {n}")
        for i, v in enumerate(c):
            # argument is T = t*t with t = 2^64 x: sin(2 pi x) = t * sum_k S_k 2^(-64(2k+1)) T^k, cos = sum_k C_k 2^(-128k) T^k
            sc = mp.mpf(2) ** (-64 * (2 * i + 1)) if name == "SIN" else mp.mpf(2) ** (-128 * i)
            kidx.append((f"SDE_{name}_C{i}", hexf(v * sc)))

    R = mp.mpf(2) ** -8 * (1 + mp.mpf(2) ** -7)

    def fq(r):
        if r == 0:
            return mp.mpf(-2)
        return -2 * mp.log1p(r) / r

    n = 5
    c = rnd(cheb_fit(fq, -R, R, n))
    e = max_err(lambda r: fq(r) * r, [mp.mpf(0)] + c, -R, R)
    report.append(f"LOG1P: {n} coefficients, max abs err of r*Q(r) {mp.nstr(e, 3)}")
    out.append(f"// {report[-1]}")
    out.append(f"#define SDE_LOG_N {n}")
    for i, v in enumerate(c):
        kidx.append((f"SDE_LOG_C{i}", hexf(v)))
    kidx.append(("SDE_TWO_LN2", hexf(2 * mp.log(2))))

    # float versions (FP32 normal transform): sin/cos(2 pi x), |x| <= 1/8
    for name, f, n, rel in (("SINF", fs, 4, True), ("COSF", fc, 5, False)):
        cf = cheb_fit(f, mp.mpf(0), U, n)
        import struct
        c32 = [mp.mpf(struct.unpack("f", struct.pack("f", float(v)))[0]) for v in cf]
        e = max_err(f, c32, mp.mpf(0), U, rel)
        report.append(f"{name}: {n} coefficients, max {'rel' if rel else 'abs'} err {mp.nstr(e, 3)}")
        out.append(f"// {report[-1]}")
        out.append(f"#define SDE_{name}_N {n}")
        for i, v in enumerate(c32):
            out.append(f"#define SDE_{name}_C{i} {float(v).hex()}f")

    inv, tl = [], []
    for j in range(128):
        cj = 1 + (mp.mpf(j) + mp.mpf(1) / 2) / 128
        ic = mp.mpf(float(1 / cj))
        inv.append(ic)
        # + 2^-50: keeps the computed -2 ln u1 strictly positive when u1 rounds to 1 (no clamp in the kernel)
        tl.append(mp.mpf(float(2 * mp.log(ic) + mp.mpf(2) ** -50)))
    out.append("// log table: {rn(1/c_j), rn(2 ln rn(1/c_j))}, c_j = 1 + (j + 1/2)/128")
    out.append("#define SDE_LOG_TAB_INIT { \\")
    for j in range(128):
        out.append(f"  {hexf(inv[j])}, {hexf(tl[j])}, \\")
    out.append("}")

    # table-driven sincos: angle = (i + 1/2 + d) / 1024 turns, |d| <= 1/2;  t = 2^54 d is the signed integer offset
    #   sin(2 pi d / 1024) = t * (S0 + S1 T + S2 T^2),  cos = 1 + C1 T + C2 T^2,   T = t^2   (Taylor; |phi| <= 3.07e-3)
    a = 2 * mp.pi / 1024 / mp.mpf(2) ** 54
    for nm, v in (("SDE_ROT_S0", a), ("SDE_ROT_S1", -a ** 3 / 6), ("SDE_ROT_S2", a ** 5 / 120),
                  ("SDE_ROT_C1", -a ** 2 / 2), ("SDE_ROT_C2", a ** 4 / 24)):
        kidx.append((nm, hexf(v)))
    # stream v2 (32-bit angle word c): the same 1024 cells, offset t = ((c & 0x3FFFFF) - 2^21) * 2^10 = 2^32 d as a 32-bit
    # signed integer: exact power-of-two rescalings of the coefficients above (a2 = a * 2^22)
    a2 = 2 * mp.pi / 1024 / mp.mpf(2) ** 32
    for nm, v in (("SDE_ROT2_S0", a2), ("SDE_ROT2_S1", -a2 ** 3 / 6), ("SDE_ROT2_S2", a2 ** 5 / 120),
                  ("SDE_ROT2_C1", -a2 ** 2 / 2), ("SDE_ROT2_C2", a2 ** 4 / 24)):
        kidx.append((nm, hexf(v)))
    # stream v2, finer tables (fewer polynomial terms): 2048 rotation cells, idx = c >> 21, t = ((c & 0x1FFFFF) - 2^20) * 2^11
    a3 = 2 * mp.pi / 2048 / mp.mpf(2) ** 32
    for nm, v in (("SDE_ROT3_S0", a3), ("SDE_ROT3_S1", -a3 ** 3 / 6), ("SDE_ROT3_C1", -a3 ** 2 / 2), ("SDE_ROT3_C2", a3 ** 4 / 24)):
        kidx.append((nm, hexf(v)))
    phi3 = mp.pi / 2048
    report.append(f"ROT3 (2048 cells): truncation errors sin {mp.nstr(phi3 ** 5 / 120, 3)}, cos {mp.nstr(phi3 ** 6 / 720, 3)}")
    out.append("// " + report[-1])
    out.append("// rotation table of stream v2: {cos, sin}(2 pi (i + 1/2) / 2048), i = 0..2047")
    out.append("#define SDE_ROT3_TAB_INIT { \\")
    for i in range(2048):
        th = 2 * mp.pi * (mp.mpf(i) + mp.mpf(1) / 2) / 2048
        out.append(f"  {hexf(mp.cos(th))}, {hexf(mp.sin(th))}, \\")
    out.append("}")
    # stream v2 log table: 512 entries, c_j = 1 + (j + 1/2)/512, |r| <= 2^-10 (1 + 2^-9): -2 log1p(r) = r * Q2(r), 4 coefficients
    R2 = mp.mpf(2) ** -10 * (1 + mp.mpf(2) ** -9)
    c2 = rnd(cheb_fit(fq, -R2, R2, 4))
    e2 = max_err(lambda r: fq(r) * r, [mp.mpf(0)] + c2, -R2, R2)
    report.append(f"LOG1P v2 (512-entry table): 4 coefficients, max abs err of r*Q2(r) {mp.nstr(e2, 3)}")
    out.append(f"// {report[-1]}")
    for i, v in enumerate(c2):
        kidx.append((f"SDE_LOG2_C{i}", hexf(v)))
    out.append("// log table of stream v2: {rn(1/c_j), rn(2 ln rn(1/c_j) + 2^-50)}, c_j = 1 + (j + 1/2)/512")
    out.append("#define SDE_LOG2_TAB_INIT { \\")
    for j in range(512):
        cj = 1 + (mp.mpf(j) + mp.mpf(1) / 2) / 512
        ic = mp.mpf(float(1 / cj))
        out.append(f"  {hexf(ic)}, {hexf(mp.mpf(float(2 * mp.log(ic) + mp.mpf(2) ** -50)))}, \\")
    out.append("}")
    phi = mp.pi / 1024
    report.append(f"ROT: truncation errors sin {mp.nstr(phi ** 7 / 5040, 3)}, cos {mp.nstr(phi ** 6 / 720, 3)}")
    out.append("// " + report[-1])
    out.append("// rotation table: {cos, sin}(2 pi (i + 1/2) / 1024), i = 0..1023")
    out.append("#define SDE_ROT_TAB_INIT { \\")
    for i in range(1024):
        th = 2 * mp.pi * (mp.mpf(i) + mp.mpf(1) / 2) / 1024
        out.append(f"  {hexf(mp.cos(th))}, {hexf(mp.sin(th))}, \\")
    out.append("}")

    # FP32 twin: 256 cells, t = 2^24 d (24-bit signed offset), sin = t (S0 + S1 T), cos - 1 = C1 T
    af = 2 * mp.pi / 256 / mp.mpf(2) ** 24
    import struct as _st
    f32 = lambda v: _st.unpack("f", _st.pack("f", float(v)))[0]
    out.append(f"#define SDE_ROTF_S0 {float(f32(af)).hex()}f")
    out.append(f"#define SDE_ROTF_S1 {float(f32(-af ** 3 / 6)).hex()}f")
    out.append(f"#define SDE_ROTF_C1 {float(f32(-af ** 2 / 2)).hex()}f")
    out.append(f"#define SDE_ROTF_C2 {float(f32(af ** 4 / 24)).hex()}f")
    out.append("// FP32 rotation table: {cos, sin}(2 pi (i + 1/2) / 256), i = 0..255")
    out.append("#define SDE_ROTF_TAB_INIT { \\")
    for i in range(256):
        th = 2 * mp.pi * (mp.mpf(i) + mp.mpf(1) / 2) / 256
        out.append(f"  {float(f32(mp.cos(th))).hex()}f, {float(f32(mp.sin(th))).hex()}f, \\")
    out.append("}")

    out.append("// FP64 constants live in one __constant__ array (uniform-register operands of DFMA on sm_100a);")
    out.append("// SDE_K(name) selects the element on the device and the literal on the host.")
    for i, (nm, hx) in enumerate(kidx):
        out.append(f"#define {nm}_I {i}")
        out.append(f"#define {nm}_V {hx}")
    out.append(f"#define SDE_K_N {len(kidx)}")
    out.append("#define SDE_K_INIT { " + ", ".join(hx for _, hx in kidx) + " }")

    hdr = ["// GENERATED by tools/gen_coeffs.py — do not edit.", "#pragma once"]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sdemodels.jl_b200", "csrc", "sde_coeffs.inc")
    with open(path, "w") as fh:
        fh.write("\n".join(hdr + out) + "\n")
    print("\n".join(report))
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    sys.exit(main())
