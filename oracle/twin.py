"""twin.py — an INDEPENDENT second restatement of the SDEModels.jl hot path, in pure Python.

TEST INFRASTRUCTURE ONLY (see the header of sde_oracle.c).  Nothing here shares code with oracle/sde_oracle.c (hand-written
C functors per model, hand-coded derivatives) or with the product's C++ front end (sdemodels.jl_b200/csrc/sde_dsl.cpp:
symbolic simplification, symbolic differentiation).  It goes the way the REFERENCE goes:

  * `@sde_model` text -> drift / diffusion by SUBSTITUTION, as src/models/codegen.jl:178-183 does (`factor_extract`: the
    differential of interest := 1, all other differentials := 0), evaluated numerically with exact symbolic 0 / 1 so that
    the surviving arithmetic is what `Calculus.simplify` leaves (x*1 -> x, x*0 -> 0, x+0 -> x); the correlation Cholesky
    factor by the recurrences of `symbolic_chol` (codegen.jl:221-237) and the product of `symbolic_mul` (:213-219);
    parameter order by order of appearance, diffusion traversed column-major (codegen.jl:116-118, cat_expressions :185-189);
  * Milstein's diffusion derivative by FORWARD-MODE DUAL NUMBERS, as src/schemes/milstein.jl:1-2 does through
    ForwardDiff (derivative / jacobian), with ForwardDiff's scalar rules (sqrt -> inv(2 sqrt x), product rule, ...);
  * the `step` formulas of euler_maruyama.jl:1-6, milstein.jl:4-10, runge_kutta.jl:4-11 with Julia's evaluation order
    (n-ary + and * fold left, literal powers x^2 = x*x, x^3 = x*x*x, static mat-vec row by row left to right);
  * the drivers of simulation.jl:39-48 (sample), :119-134 (Wiener-driven simulate!) and multilevel.jl:74-89.

Python floats are IEEE doubles without FMA contraction, so every result is comparable BIT FOR BIT with the C oracle
(tests/test_twin.py does that for all 13 models).  Speed is irrelevant: small cases only.
"""
import ast
import math
import re

EM, MILSTEIN, RUNGE_KUTTA = 0, 1, 2


class _Zero:
    """Exact symbolic 0 (a differential that was substituted away): 0*x -> 0, x+0 -> x, without touching x."""
    def __repr__(self):
        return "0"


class _One:
    def __repr__(self):
        return "1"


ZERO, ONE = _Zero(), _One()


# ------------------------------------------------------------------------------------------------------------
# forward-mode dual numbers (ForwardDiff.Dual: value + a tuple of partials), ForwardDiff 0.10 scalar rules
# ------------------------------------------------------------------------------------------------------------
class Dual:
    __slots__ = ("v", "p")

    def __init__(self, v, p):
        self.v, self.p = v, tuple(p)


def _num(x):
    """numeric value of an operand that is not symbolic"""
    if x is ONE:
        return 1.0
    if x is ZERO:
        return 0.0
    return x


def add(a, b):
    if a is ZERO:
        return b
    if b is ZERO:
        return a
    a, b = _num(a), _num(b)
    if isinstance(a, Dual) or isinstance(b, Dual):
        if not isinstance(a, Dual):
            return Dual(a + b.v, b.p)
        if not isinstance(b, Dual):
            return Dual(a.v + b, a.p)
        return Dual(a.v + b.v, [x + y for x, y in zip(a.p, b.p)])
    return a + b


def neg(a):
    if a is ZERO:
        return ZERO
    a = _num(a)
    if isinstance(a, Dual):
        return Dual(-a.v, [-x for x in a.p])
    return -a


def sub(a, b):
    if b is ZERO:
        return a
    if a is ZERO:
        return neg(b)
    a, b = _num(a), _num(b)
    if isinstance(a, Dual) or isinstance(b, Dual):
        if not isinstance(a, Dual):
            return Dual(a - b.v, [-y for y in b.p])
        if not isinstance(b, Dual):
            return Dual(a.v - b, a.p)
        return Dual(a.v - b.v, [x - y for x, y in zip(a.p, b.p)])
    return a - b


def mul(a, b):
    if a is ZERO or b is ZERO:
        return ZERO
    if a is ONE:
        return b
    if b is ONE:
        return a
    if isinstance(a, Dual) or isinstance(b, Dual):
        if not isinstance(a, Dual):
            return Dual(a * b.v, [a * y for y in b.p])            # Real * Dual
        if not isinstance(b, Dual):
            return Dual(a.v * b, [x * b for x in a.p])            # Dual * Real
        return Dual(a.v * b.v, [(b.v * x) + (a.v * y) for x, y in zip(a.p, b.p)])   # product rule (mul_tuples)
    return a * b


def div(a, b):
    if a is ZERO:
        return ZERO
    if b is ONE:
        return a
    a, b = _num(a), _num(b)
    if isinstance(a, Dual) or isinstance(b, Dual):
        if not isinstance(b, Dual):
            return Dual(a.v / b, [x / b for x in a.p])
        if not isinstance(a, Dual):
            a = Dual(a, [0.0] * len(b.p))
        ib, f = 1.0 / b.v, -(a.v / (b.v * b.v))                   # _div_partials
        return Dual(a.v / b.v, [(ib * x) + (f * y) for x, y in zip(a.p, b.p)])
    return a / b


def _pow_float(x, n):
    """x^n for an integer literal n as Julia evaluates it on Float64: literal_pow for n = 0..3, power_by_squaring above."""
    if n < 0:
        return 1.0 / _pow_float(x, -n)
    if n == 0:
        return 1.0
    if n == 1:
        return x
    if n == 2:
        return x * x
    if n == 3:
        return x * x * x
    p = n
    t = (p & -p).bit_length()          # trailing_zeros(p) + 1
    p >>= t
    t -= 1
    while t > 0:
        x *= x
        t -= 1
    y = x
    while p > 0:
        t = (p & -p).bit_length()
        p >>= t
        t -= 1
        while t >= 0:
            x *= x
            t -= 1
        y *= x
    return y


def power(a, n):
    if not isinstance(n, int):
        raise ValueError("only integer literal exponents are supported")
    if a is ZERO:
        return ZERO if n > 0 else ONE
    if a is ONE:
        return ONE
    if isinstance(a, Dual):                                        # ForwardDiff: d/dx x^n = n x^(n-1)
        return Dual(_pow_float(a.v, n), [(n * _pow_float(a.v, n - 1)) * x for x in a.p])
    return _pow_float(a, n)


def _unary(f, df):
    def g(a):
        a = _num(a)
        if isinstance(a, Dual):
            val, d = f(a.v), df(a.v)
            return Dual(val, [d * x for x in a.p])                 # Dual(val, deriv * partials)
        return f(a)
    return g


def _sqrt(x):
    if x < 0:
        raise ValueError("DomainError: sqrt of a negative number")   # Julia throws
    return math.sqrt(x)


FUNCS = {
    "sqrt": _unary(_sqrt, lambda x: 1.0 / (2.0 * _sqrt(x))),          # DiffRules: inv(2 * sqrt(x))
    "exp": _unary(math.exp, math.exp),
    "log": _unary(math.log, lambda x: 1.0 / x),
    "sin": _unary(math.sin, math.cos),
    "cos": _unary(math.cos, lambda x: -math.sin(x)),
    "abs": _unary(abs, lambda x: math.copysign(1.0, x)),
}


# ------------------------------------------------------------------------------------------------------------
# @sde_model text -> model
# ------------------------------------------------------------------------------------------------------------
def _parse(expr):
    return ast.parse(expr.strip().replace("^", "**"), mode="eval").body


def _evaluate(node, env):
    if isinstance(node, ast.Constant):
        return float(node.value)
    if isinstance(node, ast.Name):
        return env[node.id]
    if isinstance(node, ast.UnaryOp):
        v = _evaluate(node.operand, env)
        return neg(v) if isinstance(node.op, ast.USub) else v
    if isinstance(node, ast.BinOp):
        if isinstance(node.op, ast.Pow):
            if not (isinstance(node.right, ast.Constant) and isinstance(node.right.value, int)):
                raise ValueError("only integer literal exponents are supported")
            return power(_evaluate(node.left, env), node.right.value)
        a, b = _evaluate(node.left, env), _evaluate(node.right, env)
        op = {ast.Add: add, ast.Sub: sub, ast.Mult: mul, ast.Div: div}[type(node.op)]
        return op(a, b)
    if isinstance(node, ast.Call):
        return FUNCS[node.func.id](*[_evaluate(a, env) for a in node.args])
    raise ValueError(f"unsupported syntax: {ast.dump(node)}")


def _symbols(node, zero, out):
    """ordered symbols of the expression that survives the substitution (names in `zero` := 0): returns False when the
    whole subtree is annihilated.  Mirrors what `symbols(simplify(replace_symbols(ex, ...)))` keeps (codegen.jl:116-118)."""
    if isinstance(node, ast.Constant):
        return True
    if isinstance(node, ast.Name):
        if node.id in zero:
            return False
        out.append(node.id)
        return True
    if isinstance(node, ast.UnaryOp):
        return _symbols(node.operand, zero, out)
    if isinstance(node, ast.BinOp):
        if isinstance(node.op, (ast.Mult, ast.Div, ast.Pow)):
            l, r = [], []
            okl = _symbols(node.left, zero, l)
            okr = _symbols(node.right, zero, r) if not isinstance(node.op, ast.Pow) else True
            if not okl or (isinstance(node.op, ast.Mult) and not okr):
                return False
            out.extend(l + r)
            return True
        l, r = [], []
        okl, okr = _symbols(node.left, zero, l), _symbols(node.right, zero, r)
        if okl:
            out.extend(l)
        if okr:
            out.extend(r)
        return okl or okr
    if isinstance(node, ast.Call):
        for a in node.args:
            _symbols(a, zero, out)
        return True
    raise ValueError("unsupported syntax")


class Model:
    """What `@sde_model Name <equations>` defines: state variables, Wiener processes, parameters (in the reference's
    order) and callable drift / diffusion."""

    def __init__(self, name, text):
        self.name = name
        eqs = [e.strip() for line in text.split("\n") for e in line.split(";") if e.strip()]
        self.state, self.rhs, corr = [], [], []
        for e in eqs:
            if "~" in e:
                raise ValueError("jump diffusions are outside the accelerated path")
            lhs, rhs = e.split("=", 1)
            lhs = lhs.strip()
            if re.fullmatch(r"d\w+", lhs):
                self.state.append(lhs[1:])
                self.rhs.append(_parse(rhs))
            else:
                a, b = [s.strip() for s in lhs.split("*")]
                corr.append((a, b, _parse(rhs)))
        names = []
        for r in self.rhs:
            for n in ast.walk(r):
                if isinstance(n, ast.Name):
                    names.append((n.lineno, n.col_offset, n.id))
        # process variables in order of first appearance, equation by equation (matchdict over e.args[2], codegen.jl:102)
        self.wiener = []
        for r in self.rhs:
            for _, _, s in sorted((n.lineno, n.col_offset, n.id) for n in ast.walk(r) if isinstance(n, ast.Name)):
                if re.match(r"d[wW]", s) and s not in self.wiener:
                    self.wiener.append(s)
                if re.match(r"dN", s):
                    raise ValueError("jump diffusions are outside the accelerated path")
        self.D, self.M = len(self.state), len(self.wiener)
        self.diffs = ["dt"] + self.wiener
        self.corr = corr
        self.params = self._parameter_order()

    # -- parameters: symbols of drift, then of the diffusion matrix column-major, minus state variables and t ---------
    def _structural_chol(self):
        """which entries of chol(correlation) are structurally 1 / 0 / an expression (symbols of that expression)"""
        n = self.M
        sym = [[[] for _ in range(n)] for _ in range(n)]
        kind = [["one" if i == j else "zero" for j in range(n)] for i in range(n)]
        for a, b, rhs in self.corr:
            i, j = self.wiener.index(a), self.wiener.index(b)
            s = []
            _symbols(rhs, {d for d in self.diffs if d != "dt"}, s)
            s = [x for x in s if x != "dt"]
            kind[i][j] = kind[j][i] = "expr"
            sym[i][j] = sym[j][i] = s
        # symbolic_chol, symbols only: B[k][k] ~ A[k][k], B[k][j<k]; B[i][k] ~ A[i][k], B[i][j<k], B[k][j<k], B[k][k]
        B = [[None] * n for _ in range(n)]
        for k in range(n):
            dk = []
            for j in range(k):
                dk += B[k][j][1] * 1
            offdiag = any(B[k][j][0] != "zero" for j in range(k))
            B[k][k] = ("expr", dk) if offdiag else ("one", [])
            for i in range(n):
                if i == k:
                    continue
                if i < k:
                    B[i][k] = ("zero", [])      # (A[i,k] - sum) simplifies to x - x = 0 for two processes; see module note
                    continue
                terms = list(sym[i][k]) if kind[i][k] == "expr" else []
                for j in range(k):
                    if B[i][j][0] != "zero" and B[k][j][0] != "zero":
                        terms += B[i][j][1] + B[k][j][1]
                if kind[i][k] == "zero" and not terms:
                    B[i][k] = ("zero", [])
                else:
                    B[i][k] = ("expr", terms + (B[k][k][1] if B[k][k][0] == "expr" else []))
        return B

    def _parameter_order(self):
        order = []
        zero_all = set(self.diffs)
        for r in self.rhs:                                          # drift vector
            s = []
            _symbols(r, zero_all - {"dt"}, s)
            order += [x for x in s if x != "dt"]
        if self.M:
            B = self._structural_chol()
            raw = [[None] * self.M for _ in range(self.D)]
            for i, r in enumerate(self.rhs):
                for k, w in enumerate(self.wiener):
                    s = []
                    ok = _symbols(r, zero_all - {w}, s)
                    raw[i][k] = [x for x in s if x != w] if ok and w in s else None
            for j in range(self.M):                                 # column-major traversal of raw * chol
                for i in range(self.D):
                    for k in range(self.M):
                        if raw[i][k] is not None and B[k][j][0] != "zero":
                            order += raw[i][k] + B[k][j][1]
        skip = set(self.state) | {"t"}
        out = []
        for s in order:
            if s not in skip and s not in out:
                out.append(s)
        return out

    # -- numeric evaluation -------------------------------------------------------------------------------------------
    def _env(self, params, t, x, one):
        env = dict(zip(self.params, params))
        env["t"] = t
        for name, v in zip(self.state, x):
            env[name] = v
        for d in self.diffs:
            env[d] = ONE if d == one else ZERO
        return env

    def drift(self, params, t, x):
        """drift(model, t, x): D values (factor_extract(rhs, :dt, differentials), codegen.jl:105)"""
        env = self._env(params, t, x, "dt")
        return [_num(_evaluate(r, env)) for r in self.rhs]

    def chol(self, params):
        """lower Cholesky factor of the Wiener correlation matrix by the recurrences of symbolic_chol (codegen.jl:221-237),
        numerically, with exact 0 / 1 where the symbolic version has them."""
        n = self.M
        A = [[ONE if i == j else ZERO for j in range(n)] for i in range(n)]
        env = dict(zip(self.params, params))
        env.update({d: ZERO for d in self.diffs})
        env["dt"] = ONE
        for a, b, rhs in self.corr:
            i, j = self.wiener.index(a), self.wiener.index(b)
            A[i][j] = A[j][i] = _evaluate(rhs, env)
        B = [[ZERO] * n for _ in range(n)]
        for k in range(n):
            acc = ZERO                                              # Expr(:call, :+, 0, B[k,j]^2 ...)
            for j in range(k):
                acc = add(acc, power(B[k][j], 2))
            tmp = sub(A[k][k], acc)
            B[k][k] = ONE if (tmp is ONE or tmp == 1.0) else FUNCS["sqrt"](tmp)
            for i in range(n):
                if i == k:
                    continue
                acc = ZERO
                for j in range(k):
                    acc = add(acc, mul(B[i][j], B[k][j]))
                num = sub(A[i][k], acc)
                if not (num is ZERO) and not isinstance(num, (_One,)) and _num(num) == 0.0:
                    num = ZERO                                      # x - x -> 0 (Calculus.simplify)
                B[i][k] = div(num, B[k][k])
        return B

    def diffusion(self, params, t, x):
        """diffusion(model, t, x): D x M with the correlation Cholesky folded in (symbolic_mul(raw, chol), codegen.jl:111-112);
        entries may be Dual when x is."""
        if self.M == 0:
            return [[0.0] for _ in range(self.D)]
        raw = []
        for r in self.rhs:
            row = []
            for w in self.wiener:
                row.append(_evaluate(r, self._env(params, t, x, w)))
            raw.append(row)
        B = self.chol(params) if self.corr else None
        out = []
        for i in range(self.D):
            row = []
            for j in range(self.M):
                if B is None:
                    v = raw[i][j]
                else:
                    v = ZERO
                    for k in range(self.M):                         # +(A[i,1]*B[1,j], A[i,2]*B[2,j], ...)
                        v = add(v, mul(raw[i][k], B[k][j]))
                row.append(v)
            out.append(row)
        return out

    def diffusion_values(self, params, t, x):
        return [[_num(v) if not isinstance(v, Dual) else v.v for v in row] for row in self.diffusion(params, t, x)]

    def diffusion_jacobian(self, params, t, x):
        """ForwardDiff.derivative / jacobian of x -> diffusion(model, t, x) for M == 1 (milstein.jl:1-2): d sigma_i / d x_k."""
        D = self.D
        xd = [Dual(x[k], [1.0 if q == k else 0.0 for q in range(D)]) for k in range(D)]
        sig = self.diffusion(params, t, xd)
        jac = []
        for i in range(D):
            v = sig[i][0]
            jac.append(list(v.p) if isinstance(v, Dual) else [0.0] * D)   # a Real result has zero partials
        return jac


# ------------------------------------------------------------------------------------------------------------
# schemes and drivers
# ------------------------------------------------------------------------------------------------------------
def step(model, scheme, params, dt, t, x0, dw):
    """step(model, scheme, t0, x0, Δw): euler_maruyama.jl:1-6, milstein.jl:4-10, runge_kutta.jl:4-11"""
    D, M = model.D, model.M
    mu = model.drift(params, t, x0)
    sig = model.diffusion_values(params, t, x0)
    base = []
    for i in range(D):
        a = x0[i] + mu[i] * dt
        if M > 0:
            s = sig[i][0] * dw[0]
            for j in range(1, M):
                s = s + sig[i][j] * dw[j]
        else:
            s = 0.0 * dw[0]
        base.append(a + s)
    if scheme == EM:
        return base
    if M != 1:
        raise ValueError("no method: Milstein / RungeKutta are defined for AbstractSDE{D,1} only")
    c = dw[0] * dw[0] - dt
    if scheme == MILSTEIN:
        jac = model.diffusion_jacobian(params, t, x0)
        out = []
        for i in range(D):
            js = jac[i][0] * sig[0][0]
            for k in range(1, D):
                js = js + jac[i][k] * sig[k][0]
            out.append(base[i] + (js * c) / 2.0)
        return out
    if scheme == RUNGE_KUTTA:
        sq = math.sqrt(dt)
        xh = [(x0[i] + mu[i] * dt) + sig[i][0] * sq for i in range(D)]
        sp = model.diffusion_values(params, t, xh)
        den = 2.0 * sq
        return [base[i] + ((sp[i][0] - sig[i][0]) * c) / den for i in range(D)]
    raise ValueError("unknown scheme")


def sample_z(model, scheme, params, t0, dt, x0, z):
    """sample(model, scheme, t0, x0, nsteps) for every path (simulation.jl:39-48, 59-67); z[npaths][nsteps][max(M,1)]"""
    sdt = math.sqrt(dt)
    out = []
    for zk in z:
        x = list(x0)
        for i, zi in enumerate(zk, start=1):
            ti = t0 + (i - 1) * dt
            dw = [sdt * (zi[j] if model.M > 0 else 0.0) for j in range(max(model.M, 1))]
            x = step(model, scheme, params, dt, ti, x, dw)
        out.append(x)
    return out


def simulate_wiener(model, scheme, params, t0, dt, x0, w):
    """simulate!(x, model, scheme, t0, x0, w) (simulation.jl:119-134); w[npaths][nrows][max(M,1)] -> x[npaths][nrows][D]"""
    out = []
    for wk in w:
        xs = [list(x0)]
        wprev = list(wk[0])
        for i in range(2, len(wk) + 1):
            t = t0 + (i - 2) * dt
            wnext = list(wk[i - 1])
            dw = [a - b for a, b in zip(wnext, wprev)]
            xs.append(step(model, scheme, params, dt, t, xs[-1], dw))
            wprev = wnext
        out.append(xs)
    return out


def multilevel_sample_z(model, scheme, params, t0, dt_coarse, dt_fine, substeps, x0, z):
    """_multilevel_sample (multilevel.jl:74-89); z[npaths][ncoarse*substeps][M] -> (xc, xf)"""
    s = math.sqrt(dt_fine)
    M = max(model.M, 1)
    xcs, xfs = [], []
    for zk in z:
        xc, xf = list(x0), list(x0)
        nsteps = len(zk) // substeps
        for n in range(1, nsteps + 1):
            DW = [0.0] * M
            tc = t0 + (n - 1) * dt_coarse
            for i in range(1, substeps + 1):
                tf = tc + (i - 1) * dt_fine
                zi = zk[(n - 1) * substeps + (i - 1)]
                dw = [s * (zi[j] if model.M > 0 else 0.0) for j in range(M)]
                DW = [a + b for a, b in zip(DW, dw)]
                xf = step(model, scheme, params, dt_fine, tf, xf, dw)
            xc = step(model, scheme, params, dt_coarse, tc, xc, DW)
        xcs.append(xc)
        xfs.append(xf)
    return xcs, xfs
