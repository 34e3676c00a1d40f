# SDEModelsB200.jl — the Julia side of the drop-in: pure `ccall` glue over include/sdeb200.h.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image (SURVEY F9).  What CAN be checked without
# Julia is checked by tests/test_julia_shim.py: every `ccall` against the header (symbol, arity, C types), the `Opts` /
# `Payoff` mirrors against the header structs (field order, types, offsets), and every method signature below against
# the reference's methods of the same generic function (src/simulation.jl, src/multilevel.jl, src/schemes/schemes.jl) for
# dispatch ambiguities.  Every entry point bound here is exercised through the identical C ABI by the Python ctypes tests.
#
# Usage (no change to SDEModels.jl itself — new methods of its generic functions, selected by a wrapper scheme):
#
#     using SDEModels, SDEModelsB200
#     b200_init(devices = 8)                                    # one process, all GPUs (or b200_init(device = 0))
#     m  = Heston(0.01, 10.0, 0.3, 0.1, 0.4)
#     x  = sample(m, B200(EulerMaruyama(1/1024); seed=0), 0.0, [100.0, 0.4], 1024, 10^8)     # (2, 10^8) Matrix
#     x, t = simulate(BlackScholes(0.01, 0.3), B200(Milstein(1/512)), 0.0, 100.0, 512, 10^6)
#     e  = estimate(m, B200(EulerMaruyama(1/1024)), Call(100.0, exp(-0.01)), 0.0, [100.0, 0.4], 1024, 10^8)
#
# Dispatch rule that keeps the new methods unambiguous against the reference's `Vector`-x0 wrappers
# (simulation.jl:138-154, multilevel.jl:104-115): EVERY method below types `x0` — `Array{Float64,1}`, `Real` or
# `StaticVector` — so that it is either a strict subtype of a reference signature or disjoint from it.
#
# Models defined with `@sde_model` are forwarded as text: `@sde_model_b200 Name begin ... end` defines the model
# with the reference macro AND registers its equations with the device library (which re-derives drift/diffusion
# and compiles the functor with NVRTC).
module SDEModelsB200

using SDEModels
using SDEModels: AbstractSDE, JumpSDE, AbstractScheme, EulerMaruyama, Milstein, RungeKutta, MultilevelScheme, subdivide, dim
using StaticArrays
import SDEModels: sample, sample!, simulate, simulate!, wiener!, logtransition, subdivide

export B200, estimate, Call, Put, Moments, Component, @sde_model_b200, b200_init, DeviceBuffer

const lib = get(ENV, "SDEB200_LIB", joinpath(@__DIR__, "..", "libsdeb200.so"))

# -- mirror of sdeb200_opts / sdeb200_payoff (include/sdeb200.h) --------------------------------------------
struct Opts
  scheme::Int32; precision::Int32; math::Int32; rng::Int32
  t0::Float64; dt::Float64
  seed::UInt64; stream::UInt32; reserved2::UInt32
  path_offset::Int64
end
struct Payoff
  kind::Int32; comp::Int32; strike::Float64; discount::Float64
end
Moments()                          = Payoff(0, 0, 0.0, 1.0)
Call(K, disc = 1.0; comp = 1)      = Payoff(1, comp - 1, K, disc)
Put(K, disc = 1.0; comp = 1)       = Payoff(2, comp - 1, K, disc)
Component(comp = 1)                = Payoff(3, comp - 1, 0.0, 1.0)

"Wrapper scheme selecting the B200 backend: `B200(EulerMaruyama(Δt); seed, precision, strict, stream, path_offset, rng)`."
struct B200{S<:AbstractScheme} <: AbstractScheme
  scheme::S
  seed::UInt64
  precision::Type          # Float64 | Float32
  strict::Bool             # reference operation order, no FMA contraction
  stream::UInt32
  path_offset::Int64
  rng::Int32               # normal stream version: 0/1 = v1 (default), 2 = v2 (96 Philox bits per Float64 pair)
end
B200(s::AbstractScheme; seed = 0, precision = Float64, strict = false, stream = 0, path_offset = 0, rng = 0) =
  B200(s, UInt64(seed), precision, strict, UInt32(stream), Int64(path_offset), Int32(rng))
subdivide(b::B200, n) = B200(subdivide(b.scheme, n), b.seed, b.precision, b.strict, b.stream, b.path_offset, b.rng)
Base.getproperty(b::B200, s::Symbol) = s === :Δt ? getfield(b, :scheme).Δt : getfield(b, s)

scheme_id(::EulerMaruyama) = Int32(0)
scheme_id(::Milstein)      = Int32(1)
scheme_id(::RungeKutta)    = Int32(2)
scheme_id(::SDEModels.Exact) = Int32(3)   # built-in BlackScholes / OrnsteinUhlenbeck only
opts(b::B200, t0; dt = b.scheme.Δt, stream = b.stream, path_offset = b.path_offset) =
  Opts(scheme_id(b.scheme), b.precision === Float32 ? 1 : 0, b.strict ? 1 : 0, b.rng, Float64(t0), Float64(dt),
       b.seed, stream, 0, path_offset)
level_stream(stream, i) = UInt32(stream) + (UInt32(i) << 24)          # SDEB200_LEVEL_STREAM

function check(rc::Cint)
  rc == 0 && return
  msg = unsafe_string(ccall((:sdeb200_last_error, lib), Cstring, ()))
  rc == -6 && throw(DomainError(NaN, msg))     # SDEB200_E_DOMAIN: sqrt of a negative state, as in the reference
  error("sdeb200 error $rc: $msg")
end

"`b200_init(device = 0)` binds the process to one GPU; `b200_init(devices = n)` (n <= 0: all) lets every call on host arrays use n GPUs."
function b200_init(; device = nothing, devices = nothing)
  if devices !== nothing
    n = Ref{Cint}(0)
    check(ccall((:sdeb200_init_devices, lib), Cint, (Cint, Ref{Cint}), devices, n))
    return Int(n[])
  end
  check(ccall((:sdeb200_init, lib), Cint, (Cint,), device === nothing ? 0 : device))
  1
end

# -- model handles --------------------------------------------------------------------------------------------
mutable struct Handle
  p::Ptr{Cvoid}
end
const handles = Dict{DataType,Handle}()
const equations = Dict{DataType,String}()

flatten_equation(e::Expr) =          # `dW1*dW2 = ρ*dt` parses with a `begin #= line =# ... end` block on the right (codegen.jl:252)
  (e.head == :(=) && e.args[2] isa Expr && e.args[2].head == :block) ?
    string(e.args[1], " = ", e.args[2].args[end]) : string(e)

"`@sde_model_b200 Name ex [params...]`: the reference macro plus registration of the equation text."
macro sde_model_b200(name, ex, params...)
  txt = join((flatten_equation(e) for e in (ex.head == :block ? filter(e -> e isa Expr, ex.args) : [ex])), "\n")
  quote
    SDEModels.@sde_model $(esc(name)) $(esc(ex)) $(map(esc, params)...)
    SDEModelsB200.equations[$(esc(name))] = $txt
  end
end
# the registry models of the reference, equation text verbatim
equations[SDEModels.BlackScholes] = "dS = r*S*dt + σ*S*dW"                                            # black_scholes.jl:1
equations[SDEModels.Heston] = "dS = r*S*dt + sqrt(V)*S*dW1\ndV = κ*(θ-V)*dt + σ*sqrt(V)*dW2\ndW1*dW2 = ρ*dt"  # models.jl:44-48
equations[SDEModels.CoxIngersollRoss] = "dr = κ*(θ-r)*dt + σ*sqrt(r)*dW"                              # cox_ingersoll_ross.jl:1
equations[SDEModels.OrnsteinUhlenbeck] = "dx = θ*(μ-x)*dt + σ*dW"                                     # ornstein_uhlenbeck.jl:1
equations[SDEModels.FitzHughNagumo] = "dX = ɛ*(X-X^3-Y+s)*dt + σ1*dW1\ndY = (γ*X-Y+β)*dt + σ2*dW2"     # models.jl:39-42
equations[SDEModels.Lorenz] = "dx = σ*(y-x)*dt + σ1*dw1\ndy = (x*(ρ-z)-y)*dt + σ2*dw2\ndz = (x*y-β*z)*dt + σ3*dw3"  # models.jl:50-54

function handle(::Type{T}) where {T<:AbstractSDE}
  h = get!(handles, T) do
    p = Ref{Ptr{Cvoid}}(C_NULL)
    nm = string(nameof(T))
    if T === SDEModels.BlackScholes || T === SDEModels.Heston || T === SDEModels.OrnsteinUhlenbeck   # ahead-of-time functors
      check(ccall((:sdeb200_model_builtin, lib), Cint, (Cstring, Ref{Ptr{Cvoid}}), nm, p))
    else
      haskey(equations, T) || error("no equation text registered for $T: define it with @sde_model_b200")
      names = [string(f) for f in fieldnames(T)]                   # explicit order = the struct's field order
      check(ccall((:sdeb200_model_create, lib), Cint, (Cstring, Cstring, Ptr{Cstring}, Cint, Ref{Ptr{Cvoid}}),
                  nm, equations[T], names, length(names), p))
    end
    hh = Handle(p[])
    finalizer(x -> ccall((:sdeb200_model_free, lib), Cvoid, (Ptr{Cvoid},), x.p), hh)
    hh
  end
  h.p
end
params(m::T) where {T<:AbstractSDE} = Float64[getfield(m, f) for f in fieldnames(T)]
state(x0::Real) = Float64[x0]
state(x0::AbstractVector) = Vector{Float64}(x0)

# -- sample: src/simulation.jl:39-48, 59-67 ------------------------------------------------------------------------
function _sample!(out, model::AbstractSDE, b::B200, t0, x0, nsteps, npaths)
  o = Ref(opts(b, t0))
  check(ccall((:sdeb200_sample, lib), Cint,
              (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ptr{Cvoid}),
              handle(typeof(model)), o, params(model), state(x0), nsteps, npaths, out))
  out
end
# scalar x0 -> Vector{T}(npaths); vector x0 -> (D, npaths) like the working multilevel wrapper (multilevel.jl:104-108)
sample(model::AbstractSDE, b::B200, t0, x0::Real, nsteps::Integer, npaths::Integer) =
  _sample!(Vector{b.precision}(undef, npaths), model, b, t0, x0, nsteps, npaths)
sample(model::AbstractSDE{D,M}, b::B200, t0, x0::Array{Float64,1}, nsteps::Integer, npaths::Integer) where {D,M} =
  _sample!(Matrix{b.precision}(undef, D, npaths), model, b, t0, x0, nsteps, npaths)
sample(model::AbstractSDE{D,M}, b::B200, t0, x0::StaticVector{D,Float64}, nsteps::Integer, npaths::Integer) where {D,M} =
  reinterpret(SVector{D,b.precision}, vec(_sample!(Matrix{b.precision}(undef, D, npaths), model, b, t0, x0, nsteps, npaths)))
sample(model::AbstractSDE, b::B200, t0, x0::Real, nsteps::Integer) = sample(model, b, t0, x0, nsteps, 1)[1]
sample(model::AbstractSDE, b::B200, t0, x0::Array{Float64,1}, nsteps::Integer) = sample(model, b, t0, x0, nsteps, 1)[:, 1]
sample(model::AbstractSDE{D,M}, b::B200, t0, x0::StaticVector{D,Float64}, nsteps::Integer) where {D,M} = sample(model, b, t0, x0, nsteps, 1)[1]
# sample!(x, model, scheme, t0, x0, nsteps): src/simulation.jl:62-67 — fills the caller's array (Vector{T} or (D, npaths))
sample!(x::AbstractArray, model::AbstractSDE{D,M}, b::B200, t0, x0, nsteps::Integer) where {D,M} =
  (_sample!(x, model, b, t0, x0, nsteps, x0 isa Real ? length(x) : length(x) ÷ D); x)

# -- simulate: src/simulation.jl:71-76, 149-154 ---------------------------------------------------------------------
function _simulate(model::AbstractSDE{D,M}, b::B200, t0, x0, nsteps, npaths::Integer, shape) where {D,M}
  x = Array{b.precision}(undef, shape...)
  o = Ref(opts(b, t0))
  check(ccall((:sdeb200_simulate, lib), Cint,
              (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Cint, Ptr{Cvoid}),
              handle(typeof(model)), o, params(model), state(x0), nsteps, npaths, 0, x))
  x, t0 .+ b.scheme.Δt * (0:nsteps)
end
simulate(model::AbstractSDE, b::B200, t0, x0::Real, nsteps::Integer) = _simulate(model, b, t0, x0, nsteps, 1, (nsteps + 1,))
simulate(model::AbstractSDE, b::B200, t0, x0::Real, nsteps::Integer, npaths::Integer) =
  _simulate(model, b, t0, x0, nsteps, npaths, (nsteps + 1, npaths))
simulate(model::AbstractSDE{D,M}, b::B200, t0, x0::Array{Float64,1}, nsteps::Integer) where {D,M} =
  _simulate(model, b, t0, x0, nsteps, 1, (D, nsteps + 1))
simulate(model::AbstractSDE{D,M}, b::B200, t0, x0::Array{Float64,1}, nsteps::Integer, npaths::Integer) where {D,M} =
  _simulate(model, b, t0, x0, nsteps, npaths, (D, nsteps + 1, npaths))
function simulate(model::AbstractSDE{D,M}, b::B200, t0, x0::StaticVector{D,Float64}, nsteps::Integer, npaths::Integer) where {D,M}
  x, t = _simulate(model, b, t0, x0, nsteps, npaths, (D, nsteps + 1, npaths))
  reshape(reinterpret(SVector{D,b.precision}, vec(x)), (nsteps + 1, npaths)), t
end

# -- simulate!(x, model, scheme, t0, x0): src/simulation.jl:104-117 — on-device noise into the caller's array ----------
function simulate!(x::AbstractArray, model::AbstractSDE{D,M}, b::B200, t0, x0) where {D,M}
  E = eltype(x) <: StaticVector ? eltype(eltype(x)) : eltype(x)
  E === b.precision || error("array eltype $E does not match the backend precision $(b.precision)")
  nrows = (x0 isa Real || eltype(x) <: StaticVector) ? size(x, 1) : size(x, 2)
  npaths = length(x) * (eltype(x) <: StaticVector ? D : 1) ÷ (nrows * D)
  o = Ref(opts(b, t0))
  check(ccall((:sdeb200_simulate, lib), Cint,
              (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Cint, Ptr{Cvoid}),
              handle(typeof(model)), o, params(model), state(x0), nrows - 1, npaths, 0, x))
  x
end
simulate!(x::AbstractArray, model::JumpSDE, b::B200, t0, x0) = error("jump diffusions are outside the accelerated path")

# -- simulate!(x, model, scheme, t0, x0, w): src/simulation.jl:119-134 (x === w allowed) ----------------------------
function simulate!(x::AbstractArray, model::AbstractSDE{D,M}, b::B200, t0, x0, w::AbstractArray) where {D,M}
  E = eltype(x) <: StaticVector ? eltype(eltype(x)) : eltype(x)
  E === b.precision || error("array eltype $E does not match the backend precision $(b.precision)")
  nrows = (x0 isa Real || eltype(x) <: StaticVector) ? size(x, 1) : size(x, 2)
  npaths = length(x) * (eltype(x) <: StaticVector ? D : 1) ÷ (nrows * D)
  o = Ref(opts(b, t0))
  check(ccall((:sdeb200_simulate_wiener, lib), Cint,
              (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(typeof(model)), o, params(model), state(x0), nrows, npaths, w, x))
  x
end

# -- wiener!(w, Δt): src/simulation.jl:22-35 — cumulative Wiener paths from the device stream; Δt = b.scheme.Δt ----------
function _wiener!(w, b::B200, dim, nrows, npaths)
  o = Ref(opts(b, 0.0))
  check(ccall((:sdeb200_wiener, lib), Cint, (Ref{Opts}, Cint, Int64, Int64, Ptr{Cvoid}), o, dim, nrows, npaths, w))
  w
end
wiener!(w::AbstractVecOrMat{T}, b::B200) where {T<:Real} = _wiener!(w, b, 1, size(w, 1), size(w, 2))
wiener!(w::AbstractVecOrMat{SVector{D,T}}, b::B200) where {D,T<:Real} = _wiener!(w, b, D, size(w, 1), size(w, 2))
wiener!(w::AbstractArray{T,3}, b::B200) where {T<:Real} = _wiener!(w, b, size(w, 1), size(w, 2), size(w, 3))   # (D, nrows, npaths)

# -- logtransition(model, scheme, times, data): src/schemes/schemes.jl:64-76 (EulerMaruyama, uniform grid) -------------------
function logtransition(model::AbstractSDE{D,M}, b::B200{EulerMaruyama}, times, data::AbstractArray) where {D,M}
  nrows = (D == 1 || eltype(data) <: StaticVector) ? size(data, 1) : size(data, 2)
  npaths = length(data) * (eltype(data) <: StaticVector ? D : 1) ÷ (nrows * D)
  out = npaths == 1 ? Vector{b.precision}(undef, nrows - 1) : Matrix{b.precision}(undef, nrows - 1, npaths)
  o = Ref(opts(b, first(times)))
  check(ccall((:sdeb200_logtransition, lib), Cint,
              (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(typeof(model)), o, params(model), nrows, npaths, data, out))
  out
end

# -- multilevel: src/multilevel.jl:52-63 — level loop in Julia, one coupled kernel per level pair -------------------------
function _ml_sample(model::AbstractSDE{D,M}, ml::MultilevelScheme{<:B200}, t0, x0, nsteps, npaths::Array{Int64,1}) where {D,M}
  b = ml.scheme
  T = b.precision
  ml_steps = ml.substeps .^ ml.levels
  mk(n) = x0 isa Real ? Vector{T}(undef, n) : Matrix{T}(undef, D, n)
  x = Any[_sample!(mk(npaths[1]), model, subdivide(b, ml_steps[1]), t0, x0, nsteps * ml_steps[1], npaths[1])]
  for i in 2:length(ml.levels)
    coarse, fine = subdivide(b, ml_steps[i-1]), subdivide(b, ml_steps[i])
    xc, xf = mk(npaths[i]), mk(npaths[i])
    o = Ref(opts(coarse, t0; stream = level_stream(b.stream, i - 1)))
    check(ccall((:sdeb200_mlmc_sample_level, lib), Cint,
                (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Float64, Int64, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                handle(typeof(model)), o, params(model), state(x0), fine.scheme.Δt, ml.substeps,
                nsteps * ml_steps[i-1], npaths[i], xc, xf))
    push!(x, xc, xf)
  end
  x
end
sample(model::AbstractSDE, ml::MultilevelScheme{<:B200}, t0, x0::Float64, nsteps::Integer, npaths::Array{Int64,1}) =
  _ml_sample(model, ml, t0, x0, nsteps, npaths)
sample(model::AbstractSDE, ml::MultilevelScheme{<:B200}, t0, x0::Array{Float64,1}, nsteps::Integer, npaths::Array{Int64,1}) =
  _ml_sample(model, ml, t0, x0, nsteps, npaths)
sample(model::AbstractSDE{D,M}, ml::MultilevelScheme{<:B200}, t0, x0::SVector{D,Float64}, nsteps::Integer, npaths::Array{Int64,1}) where {D,M} =
  [reinterpret(SVector{D,ml.scheme.precision}, vec(xx)) for xx in _ml_sample(model, ml, t0, x0, nsteps, npaths)]
# npaths::Int64 reaches these through the reference's own broadcast methods (multilevel.jl:31-35)

# simulate(model, ::MultilevelScheme, ...): src/multilevel.jl:37-50, 91-100 — coupled full trajectories per level pair
function _ml_simulate(model::AbstractSDE{D,M}, ml::MultilevelScheme{<:B200}, t0, x0, nsteps, npaths::Array{Int64,1}) where {D,M}
  b = ml.scheme
  T = b.precision
  ml_steps = ml.substeps .^ ml.levels
  mk(rows, n) = x0 isa Real ? Matrix{T}(undef, rows, n) : Array{T}(undef, D, rows, n)
  sh(rows, n) = x0 isa Real ? (rows, n) : (D, rows, n)
  x1, t1 = _simulate(model, subdivide(b, ml_steps[1]), t0, x0, nsteps * ml_steps[1], npaths[1], sh(nsteps * ml_steps[1] + 1, npaths[1]))
  x, t = Any[x1], Any[t1]
  for i in 2:length(ml.levels)
    coarse, fine = subdivide(b, ml_steps[i-1]), subdivide(b, ml_steps[i])
    nc = nsteps * ml_steps[i-1]
    xc, xf = mk(nc + 1, npaths[i]), mk(nc * ml.substeps + 1, npaths[i])
    o = Ref(opts(coarse, t0; stream = level_stream(b.stream, i - 1)))
    check(ccall((:sdeb200_mlmc_simulate_level, lib), Cint,
                (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Float64, Int64, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                handle(typeof(model)), o, params(model), state(x0), fine.scheme.Δt, ml.substeps, nc, npaths[i], xc, xf))
    tf = t0 .+ fine.scheme.Δt * (0:nc*ml.substeps)
    push!(x, xc, xf)
    push!(t, tf[1:ml.substeps:end], tf)
  end
  x, t
end
simulate(model::AbstractSDE, ml::MultilevelScheme{<:B200}, t0, x0::Float64, nsteps::Integer, npaths::Array{Int64,1}) =
  _ml_simulate(model, ml, t0, x0, nsteps, npaths)
simulate(model::AbstractSDE, ml::MultilevelScheme{<:B200}, t0, x0::Array{Float64,1}, nsteps::Integer, npaths::Array{Int64,1}) =
  _ml_simulate(model, ml, t0, x0, nsteps, npaths)
simulate(model::AbstractSDE{D,M}, ml::MultilevelScheme{<:B200}, t0, x0::SVector{D,Float64}, nsteps::Integer, npaths::Array{Int64,1}) where {D,M} =
  _ml_simulate(model, ml, t0, Vector{Float64}(x0), nsteps, npaths)

# -- fused estimators (new; the reference returns raw states) ---------------------------------------------------------------
"`estimate(model, B200(scheme), payoff, t0, x0, nsteps, npaths)` -> (mean, stderr, n, nonfinite) per feature."
function estimate(model::AbstractSDE{D,M}, b::B200, p::Payoff, t0, x0, nsteps::Integer, npaths::Integer) where {D,M}
  nf = p.kind == 0 ? D : 1
  res = zeros(2nf + 2)
  o = Ref(opts(b, t0))
  check(ccall((:sdeb200_estimate, lib), Cint,
              (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ref{Payoff}, Ptr{Cvoid}, Ptr{Float64}),
              handle(typeof(model)), o, params(model), state(x0), nsteps, npaths, Ref(p), C_NULL, res))
  n, bad = res[2nf+1], res[2nf+2]
  good = max(n - bad, 1)
  s = reshape(res[1:2nf], 2, nf)
  μ = s[1, :] ./ good
  v = max.(s[2, :] ./ good .- μ .^ 2, 0) .* (good / max(good - 1, 1))
  (mean = μ, stderr = sqrt.(v ./ good), n = Int(n), nonfinite = Int(bad))
end

"Multilevel estimator: all levels concurrently, one all-reduce per level.  Returns per-level sums (8 x nlevels)."
function estimate(model::AbstractSDE, ml::MultilevelScheme{<:B200}, p::Payoff, t0, x0, nsteps::Integer, npaths::Array{Int64,1})
  b = ml.scheme
  res = zeros(8, length(ml.levels))
  o = Ref(opts(b, t0))
  check(ccall((:sdeb200_mlmc_estimate, lib), Cint,
              (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Cint, Int64, Ptr{Int64}, Ref{Payoff}, Ptr{Float64}),
              handle(typeof(model)), o, params(model), state(x0), ml.substeps, first(ml.levels), length(ml.levels),
              nsteps, npaths, Ref(p), res))
  res
end

# -- library-owned device buffers: results that stay on the GPU (no CUDA.jl needed) -------------------------------------------
"`DeviceBuffer(T, dims...)`: device memory in Julia's column-major order; pass it wherever an output array goes."
mutable struct DeviceBuffer
  h::Ptr{Cvoid}
  T::Type
  dims::Tuple
end
function DeviceBuffer(T::Type, dims::Integer...)
  shape = Int64[reverse(dims)...]                                   # the library counts extents in C order
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:sdeb200_buffer_alloc, lib), Cint, (Cint, Cint, Ptr{Int64}, Ref{Ptr{Cvoid}}), T === Float32 ? 1 : 0, length(shape), shape, h))
  b = DeviceBuffer(h[], T, dims)
  finalizer(x -> ccall((:sdeb200_buffer_free, lib), Cint, (Ptr{Cvoid},), x.h), b)
  b
end
Base.unsafe_convert(::Type{Ptr{Cvoid}}, b::DeviceBuffer) = ccall((:sdeb200_buffer_ptr, lib), Ptr{Cvoid}, (Ptr{Cvoid},), b.h)
function Base.copyto!(host::Array, b::DeviceBuffer, offset::Integer = 0)
  check(ccall((:sdeb200_buffer_copy_to_host, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}), b.h, offset, length(host), host))
  host
end
"`simulate!(buf, model, b, t0, x0)` with a DeviceBuffer: trajectories stay on the device (C3: 41 GB never cross PCIe)."
function simulate!(buf::DeviceBuffer, model::AbstractSDE{D,M}, b::B200, t0, x0) where {D,M}
  nrows = x0 isa Real ? buf.dims[1] : buf.dims[2]
  npaths = prod(buf.dims) ÷ (nrows * D)
  o = Ref(opts(b, t0))
  check(ccall((:sdeb200_simulate, lib), Cint,
              (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Cint, Ptr{Cvoid}),
              handle(typeof(model)), o, params(model), state(x0), nrows - 1, npaths, 0, buf))
  buf
end

simulate!(buf::DeviceBuffer, model::JumpSDE, b::B200, t0, x0) = error("jump diffusions are outside the accelerated path")

# -- multi-GPU, one process per GPU (Distributed.addprocs(ngpus)); rank 0 creates the id, all ranks join.  The simpler
#    route is ONE process: b200_init(devices = n). ------------------------------------------------------------------------
comm_unique_id() = (id = zeros(UInt8, 128); check(ccall((:sdeb200_comm_unique_id, lib), Cint, (Ptr{UInt8},), id)); id)
comm_init(id::Vector{UInt8}, nranks, rank) = check(ccall((:sdeb200_comm_init, lib), Cint, (Ptr{UInt8}, Cint, Cint), id, nranks, rank))

end # module
