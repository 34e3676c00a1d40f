"""Static checks of the Julia binding sdemodels.jl_b200/julia/SDEModelsB200.jl (Julia is not installed in the image, so
the file cannot be executed here — VERDICT r1 item 2):

 (i)   every `ccall` names a symbol include/sdeb200.h declares, with the declared arity and compatible C types;
 (ii)  the `Opts` / `Payoff` mirrors have the header structs' field order, types and offsets;
 (iii) every method the shim adds to a generic function of the reference (`sample`, `sample!`, `simulate`, `simulate!`,
       `wiener!`, `logtransition`, `subdivide`) is checked against every reference method of that function
       (parsed from /root/reference/src when present, else from the committed table below) and against its shim siblings:
       two methods whose signatures intersect must be ordered by specificity, or a third method must cover the
       intersection — otherwise Julia raises an ambiguity MethodError at the call (round 1's binding did exactly that for
       every vector-x0 call);
 (iv)  the documented calls of the header comment / INTEGRATION.md resolve to a shim method;
 (v)   every registry model of the reference has its equation text registered, and model handles carry finalizers.

The type lattice is modelled by finite sets of representative leaf types ("atoms") per argument position: T1 <: T2 iff
atoms(T1) ⊆ atoms(T2); signatures intersect iff every position shares an atom.  That is Julia's rule for the non-parametric
types used here."""
import itertools
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "sdemodels.jl_b200", "julia", "SDEModelsB200.jl")
HEADER = os.path.join(ROOT, "include", "sdeb200.h")
REF = "/root/reference/src"

# reference methods of the generic functions the shim extends: (function, [argument types]) — "" = untyped (Any).
# Parsed from the reference when it is present (this container); this committed copy serves the GPU box.
REFERENCE_METHODS = [
    ("wiener!", ["AbstractVecOrMat", ""]),                                                   # simulation.jl:22
    ("sample", ["", "", "", "", ""]),                                                        # :39
    ("sample", ["", "Exact", "", "", ""]),                                                   # :50
    ("sample", ["", "", "", "", "", ""]),                                                    # :59
    ("sample!", ["", "", "", "", "", ""]),                                                   # :62
    ("simulate", ["", "", "", "", "", "Vararg{Integer}"]),                                   # :71
    ("simulate!", ["", "", "Exact", "", ""]),                                                # :78
    ("simulate!", ["", "JumpSDE", "Exact", "", ""]),                                         # :91
    ("simulate!", ["", "", "", "", ""]),                                                     # :104
    ("simulate!", ["", "", "", "", "", ""]),                                                 # :119
    ("sample", ["", "AbstractScheme", "", "Array{Float64,1}", ""]),                          # :138
    ("sample", ["", "AbstractScheme", "", "Array{Float64,1}", "", "Vararg{Integer}"]),       # :143
    ("simulate", ["", "AbstractScheme", "", "Array{Float64,1}", "", "Vararg{Integer}"]),     # :149
    ("simulate!", ["", "JumpSDE", "", "", ""]),                                              # :167
    ("simulate", ["", "MultilevelScheme", "", "", "", "Int64"]),                             # multilevel.jl:31
    ("sample", ["", "MultilevelScheme", "", "", "", "Int64"]),                               # :34
    ("simulate", ["", "MultilevelScheme", "", "Union{Float64,SVector}", "", "Array{Int64,1}"]),   # :37
    ("sample", ["", "MultilevelScheme", "", "Union{Float64,SVector}", "", "Array{Int64,1}"]),     # :52
    ("sample", ["", "MultilevelScheme", "", "Array{Float64,1}", "", "Array{Int64,1}"]),      # :104
    ("simulate", ["", "MultilevelScheme", "", "Array{Float64,1}", "", "Array{Int64,1}"]),    # :110
    ("logtransition", ["AbstractSDE", "NormalScheme", "", "", ""]),                          # schemes.jl:64
    ("logtransition", ["", "", "", ""]),                                                     # :74
    ("subdivide", ["ConditionalScheme", ""]),                                                # :51
    ("subdivide", ["UnconditionalScheme", ""]),                                              # :53
]

# ---- the finite type model -------------------------------------------------------------------------------------------
ATOMS = {
    # values
    "Float64", "Float32", "Int64", "Int32", "BigInt", "Bool", "String", "Symbol",
    "Vector{Float64}", "Vector{Float32}", "Vector{Int64}", "Matrix{Float64}", "Array3{Float64}", "Vector{SVector}", "Matrix{SVector}",
    "SVector", "UnitRange", "StepRangeLen", "DeviceBuffer", "Payoff",
    # models
    "Model", "JumpModel",
    # schemes
    "EulerMaruyama", "Milstein", "Exact", "ModifiedBridge", "ImplicitEM", "B200{EulerMaruyama}", "B200{Milstein}",
    "ML{EulerMaruyama}", "ML{B200}",
}
REALS = {"Float64", "Float32", "Int64", "Int32", "BigInt", "Bool"}
INTEGERS = {"Int64", "Int32", "BigInt", "Bool"}
ARRAYS = {"Vector{Float64}", "Vector{Float32}", "Vector{Int64}", "Matrix{Float64}", "Array3{Float64}", "Vector{SVector}", "Matrix{SVector}", "SVector",
          "UnitRange", "StepRangeLen"}
SCHEMES = {"EulerMaruyama", "Milstein", "Exact", "ModifiedBridge", "ImplicitEM", "B200{EulerMaruyama}", "B200{Milstein}"}
TYPES = {
    "": ATOMS, "Any": ATOMS,
    "Real": REALS, "Number": REALS, "Integer": INTEGERS, "Int64": {"Int64"}, "Int": {"Int64"}, "Float64": {"Float64"},
    "Array{Float64,1}": {"Vector{Float64}"}, "Vector{Float64}": {"Vector{Float64}"}, "Array{Int64,1}": {"Vector{Int64}"},
    "AbstractArray": ARRAYS, "AbstractVector": {"Vector{Float64}", "Vector{Float32}", "Vector{Int64}", "Vector{SVector}", "SVector", "UnitRange", "StepRangeLen"},
    "AbstractVecOrMat": {"Vector{Float64}", "Vector{Float32}", "Vector{Int64}", "Matrix{Float64}", "Vector{SVector}", "Matrix{SVector}", "SVector", "UnitRange", "StepRangeLen"},
    "AbstractVecOrMat{T<:Real}": {"Vector{Float64}", "Vector{Float32}", "Vector{Int64}", "Matrix{Float64}", "UnitRange", "StepRangeLen"},
    "AbstractVecOrMat{SVector}": {"Vector{SVector}", "Matrix{SVector}"},
    "AbstractArray{T,3}": {"Array3{Float64}"},
    "StaticVector": {"SVector"}, "SVector": {"SVector"}, "Union{Float64,SVector}": {"Float64", "SVector"},
    "Array": {"Vector{Float64}", "Vector{Float32}", "Vector{Int64}", "Matrix{Float64}", "Array3{Float64}", "Vector{SVector}", "Matrix{SVector}"},
    "AbstractSDE": {"Model", "JumpModel"}, "JumpSDE": {"JumpModel"},
    "AbstractScheme": SCHEMES, "Exact": {"Exact"}, "UnconditionalScheme": {"EulerMaruyama", "Milstein", "Exact"},
    "ConditionalScheme": {"ModifiedBridge"}, "NormalScheme": {"EulerMaruyama", "ImplicitEM", "ModifiedBridge"},
    "B200": {"B200{EulerMaruyama}", "B200{Milstein}"}, "B200{EulerMaruyama}": {"B200{EulerMaruyama}"},
    "MultilevelScheme": {"ML{EulerMaruyama}", "ML{B200}"}, "MultilevelScheme{<:B200}": {"ML{B200}"},
    "Payoff": {"Payoff"}, "DeviceBuffer": {"DeviceBuffer"},
}


def normalize(t):
    """a Julia type annotation -> a key of TYPES"""
    t = t.strip()
    t = re.sub(r"AbstractSDE\{[^}]*\}", "AbstractSDE", t)
    t = re.sub(r"JumpSDE\{[^}]*\}", "JumpSDE", t)
    t = re.sub(r"StaticVector\{[^}]*\}", "StaticVector", t)
    t = re.sub(r"SVector\{D,\s*Float64\}", "SVector", t)
    t = re.sub(r"Vararg\{Integer,\s*N\}", "Vararg{Integer}", t)
    t = re.sub(r"AbstractVecOrMat\{T\}", "AbstractVecOrMat", t)
    t = re.sub(r"AbstractVecOrMat\{SVector\{D,\s*T\}\}", "AbstractVecOrMat{SVector}", t)
    t = re.sub(r"SDEModels\.", "", t)
    if re.fullmatch(r"T\d?|S", t):               # a bare type parameter: unconstrained unless the where-clause says otherwise
        return ""
    return t


def split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        if ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def parse_methods(text, functions):
    """[(function, [type keys], source line)] of `function f(args) ... ` and `f(args) = ...` definitions"""
    methods = []
    pat = re.compile(r"^(?:@inline\s+)?(?:function\s+)?(?:SDEModels\.)?(" + "|".join(re.escape(f) for f in functions) + r")\(", re.M)
    for m in pat.finditer(text):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        args = text[m.end():i - 1]
        tail = text[i:i + 200]
        is_def = text[m.start():m.end()].lstrip().startswith(("function", "@inline")) or re.match(r"\s*(where\s*\{?[^=\n]*\}?\s*)?=(?!=)", tail)
        if not is_def:
            continue
        where = re.match(r"\s*where\s*(\{[^}]*\}|\S+)", tail)
        bounds = {}
        if where:
            for b in split_args(where.group(1).strip("{}")):
                if "<:" in b:
                    k, v = b.split("<:", 1)
                    bounds[k.strip()] = v.strip()
        types = []
        pos = args.split(";")[0]
        for a in split_args(pos):
            a = a.strip()
            if "=" in a and "::" not in a.split("=")[0] and not a.split("=")[0].strip().isidentifier():
                continue
            a = a.split("=")[0].strip() if re.match(r"^[\w!]+(::[^=]+)?\s*=", a) else a      # optional argument: typed part only
            if a.endswith("..."):
                t = a[:-3].split("::")[1].strip() if "::" in a else ""
                types.append("Vararg{" + (t or "Any") + "}")
                continue
            t = a.split("::", 1)[1].strip() if "::" in a else ""
            if t in bounds:
                t = bounds[t]
                t = "Union{Float64,SVector}" if t.startswith("Union{Float64") else ("AbstractVecOrMat{T<:Real}" if False else t)
            if re.fullmatch(r"AbstractVecOrMat\{T\}", t) and bounds.get("T", "").strip() == "Real":
                t = "AbstractVecOrMat{T<:Real}"
            if re.fullmatch(r"AbstractArray\{T,\s*3\}", t):
                t = "AbstractArray{T,3}"
            types.append(normalize(t))
        line = text.count("\n", 0, m.start()) + 1
        methods.append((m.group(1), types, line))
    return methods


def expand(types, n):
    """argument type sets of a method for a call with n arguments, or None if the arity does not fit"""
    if types and types[-1].startswith("Vararg{"):
        inner = types[-1][7:-1]
        if n < len(types) - 1:
            return None
        types = types[:-1] + [inner if inner != "Any" else ""] * (n - len(types) + 1)
    if len(types) != n:
        return None
    return [TYPES[t] for t in types]


def more_specific(a, b, n):
    A, B = expand(a, n), expand(b, n)
    return all(x <= y for x, y in zip(A, B))


def ambiguities(methods):
    bad = []
    for (f1, a, l1), (f2, b, l2) in itertools.combinations(methods, 2):
        if f1 != f2 or (l1[0] != "shim" and l2[0] != "shim"):
            continue      # other function, or two reference methods (the reference's own ambiguity, e.g. Exact with a Vector x0, is not ours)
        for n in range(1, 9):
            A, B = expand(a, n), expand(b, n)
            if A is None or B is None:
                continue
            inter = [x & y for x, y in zip(A, B)]
            if not all(inter):
                continue
            if more_specific(a, b, n) or more_specific(b, a, n):
                continue
            # a third method that covers the intersection and is at least as specific as both resolves it
            cover = False
            for f3, c, _ in methods:
                Cn = expand(c, n) if f3 == f1 else None
                if Cn and c is not a and c is not b and all(i <= z for i, z in zip(inter, Cn)) and \
                        all(z <= x and z <= y for z, x, y in zip(Cn, A, B)):
                    cover = True
            if not cover:
                bad.append((f1, n, a, l1, b, l2))
    return bad


FUNCS = ["sample!", "sample", "simulate!", "simulate", "wiener!", "logtransition", "subdivide"]


def shim_methods():
    return [(f, t, ("shim", l)) for f, t, l in parse_methods(open(SHIM, encoding="utf-8").read(), FUNCS)]


def reference_methods():
    if os.path.isdir(REF):
        out = []
        for fn in ("simulation.jl", "multilevel.jl", "schemes/schemes.jl"):
            txt = open(os.path.join(REF, fn), encoding="utf-8").read()
            txt = "\n".join(l for l in txt.split("\n") if not l.lstrip().startswith("#"))
            out += [(f, t, (fn, l)) for f, t, l in parse_methods(txt, FUNCS)]
        return out
    return [(f, t, ("table", i)) for i, (f, t) in enumerate(REFERENCE_METHODS)]


def test_reference_method_table_is_current():
    """the committed table equals what the reference sources declare (checked wherever the reference is present)"""
    if not os.path.isdir(REF):
        pytest.skip("reference sources not on this machine: the committed table is used")
    got = sorted((f, tuple(t)) for f, t, _ in reference_methods() if not (f in ("sample", "simulate!") and "JumpSDE" in t and False))
    want = sorted((f, tuple(t)) for f, t in REFERENCE_METHODS)
    assert got == want


def test_the_checker_finds_round_one_ambiguity():
    """self-test: the round-1 signature `sample(model::AbstractSDE{D,M}, b::B200, t0, x0, nsteps::Integer, npaths::Integer)`
    (x0 untyped) IS ambiguous against simulation.jl:143 — the analysis must say so."""
    old = [("sample", ["AbstractSDE", "B200", "", "", "Integer", "Integer"], ("shim", 108))]
    bad = ambiguities(old + reference_methods())
    assert any(f == "sample" and n == 6 for f, n, *_ in bad)


def test_no_dispatch_ambiguity_between_shim_and_reference():
    ms = shim_methods()
    assert len(ms) >= 25, ms
    bad = ambiguities(ms + reference_methods())
    assert not bad, "\n".join(f"{f}/{n}: {a} @{l1}  vs  {b} @{l2}" for f, n, a, l1, b, l2 in bad)


def resolve(methods, f, call):
    """the unique most specific applicable method for a call given as atoms, or None"""
    app = []
    for g, t, l in methods:
        T = expand(t, len(call)) if g == f else None
        if T and all(c in s for c, s in zip(call, T)):
            app.append((t, l))
    best = [m for m in app if all(more_specific(m[0], o[0], len(call)) for o in app)]
    return best[0][1] if len(best) >= 1 else None


def test_documented_calls_resolve_to_the_shim():
    ms = shim_methods() + reference_methods()
    calls = [
        ("sample", ["Model", "B200{EulerMaruyama}", "Float64", "Vector{Float64}", "Int64", "Int64"]),
        ("sample", ["Model", "B200{EulerMaruyama}", "Float64", "Float64", "Int64", "Int64"]),
        ("sample", ["Model", "B200{EulerMaruyama}", "Float64", "Vector{Float64}", "Int64"]),
        ("sample", ["Model", "B200{EulerMaruyama}", "Float64", "SVector", "Int64", "Int64"]),
        ("sample!", ["Matrix{Float64}", "Model", "B200{EulerMaruyama}", "Float64", "Vector{Float64}", "Int64"]),
        ("simulate", ["Model", "B200{Milstein}", "Float64", "Float64", "Int64", "Int64"]),
        ("simulate", ["Model", "B200{EulerMaruyama}", "Float64", "Vector{Float64}", "Int64", "Int64"]),
        ("simulate", ["Model", "B200{EulerMaruyama}", "Float64", "Vector{Float64}", "Int64"]),
        ("simulate!", ["Matrix{Float64}", "Model", "B200{EulerMaruyama}", "Float64", "Float64", "Matrix{Float64}"]),
        ("simulate!", ["Matrix{Float64}", "Model", "B200{EulerMaruyama}", "Float64", "Float64"]),
        ("simulate!", ["DeviceBuffer", "Model", "B200{EulerMaruyama}", "Float64", "Float64"]),
        ("wiener!", ["Matrix{Float64}", "B200{EulerMaruyama}"]),
        ("wiener!", ["Matrix{SVector}", "B200{EulerMaruyama}"]),
        ("logtransition", ["Model", "B200{EulerMaruyama}", "StepRangeLen", "Vector{Float64}"]),
        ("sample", ["Model", "ML{B200}", "Float64", "Vector{Float64}", "Int64", "Vector{Int64}"]),
        ("sample", ["Model", "ML{B200}", "Float64", "Float64", "Int64", "Vector{Int64}"]),
        ("simulate", ["Model", "ML{B200}", "Float64", "Vector{Float64}", "Int64", "Vector{Int64}"]),
        ("simulate", ["Model", "ML{B200}", "Float64", "Float64", "Int64", "Vector{Int64}"]),
        ("subdivide", ["B200{EulerMaruyama}", "Int64"]),
    ]
    for f, call in calls:
        where = resolve(ms, f, call)
        assert where is not None and where[0] == "shim", (f, call, where)
    # npaths::Int64 for a MultilevelScheme goes through the reference's own broadcast (multilevel.jl:31-35), then to the shim
    assert resolve(ms, "sample", ["Model", "ML{B200}", "Float64", "Vector{Float64}", "Int64", "Int64"])[0] != "shim"
    # the reference's own calls are untouched
    assert resolve(ms, "sample", ["Model", "EulerMaruyama", "Float64", "Vector{Float64}", "Int64", "Int64"])[0] != "shim"
    assert resolve(ms, "simulate", ["Model", "ML{EulerMaruyama}", "Float64", "Vector{Float64}", "Int64", "Vector{Int64}"])[0] != "shim"


# ---- ccalls against the header --------------------------------------------------------------------------------------
def header_prototypes():
    txt = open(HEADER, encoding="utf-8").read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"^\s*#.*$", "", txt, flags=re.M)          # preprocessor lines
    protos = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(sdeb200_\w+)\s*\(([^)]*)\)\s*;", txt):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        alist = [] if args in ("", "void") else [re.sub(r"\s+", " ", a.strip()) for a in args.split(",")]
        protos[name] = (ret, alist)
    return protos


def c_kind(decl):
    """C parameter declaration -> coarse kind"""
    d = decl.replace("const ", "").strip()
    d = re.sub(r"\b\w+$", "", d).strip() if not d.endswith("*") and " " in d else d     # drop the parameter name
    d = d.replace(" ", "")
    if d.startswith("char*") and d.endswith("**"):
        return "ptr_cstring"
    if d.endswith("**"):
        return "ptrptr"
    if d in ("char*",):
        return "cstring"
    if d.startswith("char*"):
        return "ptr"
    if d.endswith("*"):
        base = d[:-1]
        return {"double": "ptr_f64", "int64_t": "ptr_i64", "int": "ptr_i32", "sdeb200_opts": "ptr_opts", "sdeb200_payoff": "ptr_payoff"}.get(base, "ptr")
    return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "double": "f64", "uint64_t": "u64", "uint32_t": "u32", "void": "void"}.get(d, d)


JL_KIND = {
    "Cint": {"i32"}, "Int32": {"i32"}, "Int64": {"i64"}, "Float64": {"f64"}, "Cdouble": {"f64"}, "UInt64": {"u64"}, "Cvoid": {"void"},
    "Cstring": {"cstring"}, "Ptr{Cvoid}": {"ptr", "ptr_f64", "ptr_i64", "ptr_opts", "ptr_payoff", "ptrptr"}, "Ptr{Float64}": {"ptr_f64"},
    "Ptr{Int64}": {"ptr_i64"}, "Ref{Opts}": {"ptr_opts"}, "Ref{Payoff}": {"ptr_payoff"}, "Ref{Ptr{Cvoid}}": {"ptrptr"},
    "Ptr{Cstring}": {"ptr_cstring"}, "Ptr{UInt8}": {"ptr"}, "Ref{Cint}": {"ptr_i32"},
}


def test_every_ccall_matches_the_header():
    txt = open(SHIM, encoding="utf-8").read()
    protos = header_prototypes()
    calls = list(re.finditer(r"ccall\(\(:(\w+),\s*lib\),\s*([\w{}]+),\s*\(([^)]*)\)", txt))
    assert len(calls) >= 20
    seen = set()
    for m in calls:
        name, ret, args = m.group(1), m.group(2), [a.strip() for a in split_args(m.group(3)) if a.strip()]
        assert name in protos, f"ccall of {name}: not declared in include/sdeb200.h"
        cret, cargs = protos[name]
        assert len(args) == len(cargs), f"{name}: {len(args)} ccall arguments, header has {len(cargs)}: {cargs}"
        for i, (ja, ca) in enumerate(zip(args, cargs)):
            assert ja in JL_KIND, f"{name}: unknown Julia ccall type {ja}"
            assert c_kind(ca) in JL_KIND[ja], f"{name} argument {i}: Julia {ja} vs C `{ca}`"
        rk = "cstring" if "char" in cret else ("ptr" if cret.endswith("*") else c_kind(cret))
        assert rk in JL_KIND[ret] or (rk == "ptr" and ret == "Ptr{Cvoid}"), f"{name}: return {ret} vs C `{cret}`"
        seen.add(name)
    # the drivers of the path are all bound
    for need in ("sdeb200_sample", "sdeb200_simulate", "sdeb200_simulate_wiener", "sdeb200_wiener", "sdeb200_logtransition",
                 "sdeb200_mlmc_sample_level", "sdeb200_mlmc_simulate_level", "sdeb200_estimate", "sdeb200_mlmc_estimate",
                 "sdeb200_model_create", "sdeb200_model_builtin", "sdeb200_model_free", "sdeb200_init", "sdeb200_init_devices",
                 "sdeb200_buffer_alloc", "sdeb200_buffer_free", "sdeb200_buffer_copy_to_host"):
        assert need in seen, f"{need} is not bound by the shim"


def test_struct_mirrors_match_the_header():
    h = re.sub(r"/\*.*?\*/", "", open(HEADER, encoding="utf-8").read(), flags=re.S)
    j = open(SHIM, encoding="utf-8").read()
    size = {"int32_t": 4, "uint32_t": 4, "double": 8, "uint64_t": 8, "int64_t": 8}
    jl = {"Int32": ("int32_t", 4), "UInt32": ("uint32_t", 4), "Float64": ("double", 8), "UInt64": ("uint64_t", 8), "Int64": ("int64_t", 8)}
    for cname, jname in (("sdeb200_opts", "Opts"), ("sdeb200_payoff", "Payoff")):
        body = re.search(r"typedef struct \{([^}]*)\}\s*" + cname + ";", re.sub(r"^\s*#.*$", "", h, flags=re.M)).group(1)
        cf = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ty, names = decl.split(None, 1)
            cf += [(n.strip(), ty) for n in names.split(",")]
        jbody = re.search(r"struct " + jname + r"\n(.*?)\nend", j, flags=re.S).group(1)
        jf = [(n.strip(), t.strip()) for n, t in re.findall(r"(\w+)::(\w+)", jbody)]
        assert [n for n, _ in cf] == [n for n, _ in jf], (cname, cf, jf)
        off_c = off_j = 0
        for (n, ct), (_, jt) in zip(cf, jf):
            assert jl[jt][0] == ct, f"{cname}.{n}: C {ct} vs Julia {jt}"
            a = size[ct]
            off_c = (off_c + a - 1) // a * a
            off_j = (off_j + a - 1) // a * a
            assert off_c == off_j
            off_c += a
            off_j += a
    # the Python mirror used by the tests has the same layout
    import _pkg
    S = _pkg.load()
    assert [f[0] for f in S._lib.Opts._fields_] == [n for n, _ in re.findall(r"(\w+)::(\w+)", re.search(r"struct Opts\n(.*?)\nend", j, flags=re.S).group(1))]


def test_registry_models_finalizers_and_macro():
    j = open(SHIM, encoding="utf-8").read()
    import _pkg
    S = _pkg.load()
    for name, eq in S.REGISTRY.items():
        m = re.search(r"equations\[SDEModels\." + name + r"\] = \"(.*?)\"\s", j)
        assert m, f"{name}: no equation text registered in the shim"
        assert m.group(1).replace("\\n", "\n") == eq, name
    assert "finalizer(x -> ccall((:sdeb200_model_free, lib)" in j
    assert "finalizer(x -> ccall((:sdeb200_buffer_free, lib)" in j
    # correlation equations parse with a block on the right-hand side (codegen.jl:252): the macro flattens them
    assert "flatten_equation" in j and "e.args[2].args[end]" in j
    # documented example calls
    doc = open(os.path.join(ROOT, "INTEGRATION.md"), encoding="utf-8").read()
    assert "sample(m, B200(EulerMaruyama(1/1024); seed = 0), 0.0, [100.0, 0.4], 1024, 10^8)" in doc
