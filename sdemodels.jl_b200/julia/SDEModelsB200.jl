# SDEModelsB200.jl — the Julia side of the drop-in: pure `ccall` glue over include/sdeb200.h.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image (SURVEY F9).  The file is kept
# deliberately small so it can be reviewed against the header line by line; every entry point it binds is
# exercised through the identical C ABI by the Python ctypes tests (tests/test_gpu_*.py).
#
# Usage (no change to SDEModels.jl itself — new methods of its generic functions, selected by a wrapper scheme):
#
#     using SDEModels, SDEModelsB200
#     m  = Heston(0.01, 10.0, 0.3, 0.1, 0.4)
#     x  = sample(m, B200(EulerMaruyama(1/1024); seed=0), 0.0, [100.0, 0.4], 1024, 10^8)     # (2, 10^8) Matrix
#     x, t = simulate(BlackScholes(0.01, 0.3), B200(Milstein(1/512)), 0.0, 100.0, 512, 10^6)
#     e  = estimate(m, B200(EulerMaruyama(1/1024)), Call(100.0, exp(-0.01)), 0.0, [100.0, 0.4], 1024, 10^8)
#
# Models defined with `@sde_model` are forwarded as text: `@sde_model_b200 Name begin ... end` defines the model
# with the reference macro AND registers its equations with the device library (which re-derives drift/diffusion
# and compiles the functor with NVRTC).
module SDEModelsB200

using SDEModels
using SDEModels: AbstractSDE, AbstractScheme, EulerMaruyama, Milstein, RungeKutta, MultilevelScheme, subdivide, dim
using StaticArrays
import SDEModels: sample, simulate, simulate!, logtransition

export B200, estimate, Call, Put, Moments, Component, @sde_model_b200, b200_init

const lib = get(ENV, "SDEB200_LIB", joinpath(@__DIR__, "..", "libsdeb200.so"))

# -- mirror of sdeb200_opts / sdeb200_payoff (include/sdeb200.h) --------------------------------------------
struct Opts
  scheme::Int32; precision::Int32; math::Int32; reserved::Int32
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

"Wrapper scheme selecting the B200 backend: `B200(EulerMaruyama(Δt); seed, precision, math, stream, path_offset)`."
struct B200{S<:AbstractScheme} <: AbstractScheme
  scheme::S
  seed::UInt64
  precision::Type          # Float64 | Float32
  strict::Bool             # reference operation order, no FMA contraction
  stream::UInt32
  path_offset::Int64
end
B200(s::AbstractScheme; seed = 0, precision = Float64, strict = false, stream = 0, path_offset = 0) =
  B200(s, UInt64(seed), precision, strict, UInt32(stream), Int64(path_offset))
SDEModels.subdivide(b::B200, n) = B200(subdivide(b.scheme, n), b.seed, b.precision, b.strict, b.stream, b.path_offset)

scheme_id(::EulerMaruyama) = Int32(0)
scheme_id(::Milstein)      = Int32(1)
scheme_id(::RungeKutta)    = Int32(2)
scheme_id(::SDEModels.Exact) = Int32(3)   # built-in BlackScholes / OrnsteinUhlenbeck only
opts(b::B200, t0; dt = b.scheme.Δt, stream = b.stream) =
  Opts(scheme_id(b.scheme), b.precision === Float32 ? 1 : 0, b.strict ? 1 : 0, 0, Float64(t0), Float64(dt),
       b.seed, stream, 0, b.path_offset)

function check(rc::Cint)
  rc == 0 && return
  msg = unsafe_string(ccall((:sdeb200_last_error, lib), Cstring, ()))
  rc == -6 && throw(DomainError(NaN, msg))     # SDEB200_E_DOMAIN: sqrt of a negative state, as in the reference
  error("sdeb200 error $rc: $msg")
end

b200_init(device = 0) = check(ccall((:sdeb200_init, lib), Cint, (Cint,), device))

# -- model handles --------------------------------------------------------------------------------------------
const handles = Dict{DataType,Ptr{Cvoid}}()
const equations = Dict{DataType,String}()

"`@sde_model_b200 Name ex [params...]`: the reference macro plus registration of the equation text."
macro sde_model_b200(name, ex, params...)
  txt = join((string(e) for e in (ex.head == :block ? filter(e -> e isa Expr, ex.args) : [ex])), "\n")
  quote
    SDEModels.@sde_model $(esc(name)) $(esc(ex)) $(map(esc, params)...)
    SDEModelsB200.equations[$(esc(name))] = $txt
  end
end
equations[SDEModels.BlackScholes] = "dS = r*S*dt + σ*S*dW"                       # black_scholes.jl:1
equations[SDEModels.Heston] = "dS = r*S*dt + sqrt(V)*S*dW1\ndV = κ*(θ-V)*dt + σ*sqrt(V)*dW2\ndW1*dW2 = ρ*dt"  # models.jl:44-48

function handle(::Type{T}) where {T<:AbstractSDE}
  get!(handles, T) do
    h = Ref{Ptr{Cvoid}}(C_NULL)
    nm = string(nameof(T))
    if T === SDEModels.BlackScholes || T === SDEModels.Heston || T === SDEModels.OrnsteinUhlenbeck   # ahead-of-time functors
      check(ccall((:sdeb200_model_builtin, lib), Cint, (Cstring, Ref{Ptr{Cvoid}}), nm, h))
    else
      names = [string(f) for f in fieldnames(T)]                   # explicit order = the struct's field order
      check(ccall((:sdeb200_model_create, lib), Cint, (Cstring, Cstring, Ptr{Cstring}, Cint, Ref{Ptr{Cvoid}}),
                  nm, equations[T], names, length(names), h))
    end
    h[]
  end
end
params(m::T) where {T<:AbstractSDE} = Float64[getfield(m, f) for f in fieldnames(T)]
state(x0::Number) = Float64[x0]
state(x0::AbstractVector) = Vector{Float64}(x0)

# -- sample: src/simulation.jl:59-67 ------------------------------------------------------------------------------
function sample(model::AbstractSDE{D,M}, b::B200, t0, x0, nsteps::Integer, npaths::Integer) where {D,M}
  T = b.precision
  out = x0 isa Number ? Vector{T}(undef, npaths) : Matrix{T}(undef, D, npaths)   # (D, npaths) like multilevel.jl:104-108
  o = Ref(opts(b, t0))
  check(ccall((:sdeb200_sample, lib), Cint,
              (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ptr{Cvoid}),
              handle(typeof(model)), o, params(model), state(x0), nsteps, npaths, out))
  out
end
sample(model::AbstractSDE, b::B200, t0, x0, nsteps::Integer) = (x = sample(model, b, t0, x0, nsteps, 1); x0 isa Number ? x[1] : x[:, 1])

# -- simulate: src/simulation.jl:71-76,149-154 -----------------------------------------------------------------------
function simulate(model::AbstractSDE{D,M}, b::B200, t0, x0, nsteps::Integer, npaths::Integer = 1) where {D,M}
  T = b.precision
  x = x0 isa Number ? Array{T}(undef, nsteps + 1, npaths) : Array{T}(undef, D, nsteps + 1, npaths)
  o = Ref(opts(b, t0))
  check(ccall((:sdeb200_simulate, lib), Cint,
              (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Cint, Ptr{Cvoid}),
              handle(typeof(model)), o, params(model), state(x0), nsteps, npaths, 0, x))
  x, t0 .+ b.scheme.Δt * (0:nsteps)
end

# -- simulate!(x, model, scheme, t0, x0, w): src/simulation.jl:119-134 (x === w allowed) ----------------------------
function simulate!(x::AbstractArray, model::AbstractSDE, b::B200, t0, x0, w::AbstractArray)
  E = eltype(x) <: SVector ? eltype(eltype(x)) : eltype(x)
  E === b.precision || error("array eltype $E does not match the backend precision $(b.precision)")
  nrows, npaths = size(x, 1), size(x, 2)
  o = Ref(opts(b, t0))
  check(ccall((:sdeb200_simulate_wiener, lib), Cint,
              (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(typeof(model)), o, params(model), state(x0), nrows, npaths, w, x))
  x
end

# -- logtransition(model, scheme, times, data): src/schemes/schemes.jl:64-76 (EulerMaruyama, uniform grid) -------------------
function logtransition(model::AbstractSDE{D,M}, b::B200{EulerMaruyama}, times, data::AbstractVector) where {D,M}
  nrows = length(data)
  out = Vector{b.precision}(undef, nrows - 1)
  o = Ref(opts(b, first(times)))
  check(ccall((:sdeb200_logtransition, lib), Cint,
              (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(typeof(model)), o, params(model), nrows, 1, data, out))
  out
end

# -- multilevel: src/multilevel.jl:52-63 — level loop in Julia, one coupled kernel per level pair -------------------------
function sample(model::AbstractSDE{D,M}, ml::MultilevelScheme{<:B200}, t0, x0, nsteps::Integer, npaths::Vector{Int}) where {D,M}
  b = ml.scheme
  T = b.precision
  ml_steps = ml.substeps .^ ml.levels
  mk(n) = x0 isa Number ? Vector{T}(undef, n) : Matrix{T}(undef, D, n)
  x = Any[sample(model, subdivide(b, ml_steps[1]), t0, x0, nsteps * ml_steps[1], npaths[1])]
  for i in 2:length(ml.levels)
    coarse, fine = subdivide(b, ml_steps[i-1]), subdivide(b, ml_steps[i])
    xc, xf = mk(npaths[i]), mk(npaths[i])
    o = Ref(opts(coarse, t0; stream = b.stream + UInt32(i - 1)))
    check(ccall((:sdeb200_mlmc_sample_level, lib), Cint,
                (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Float64, Int64, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                handle(typeof(model)), o, params(model), state(x0), fine.scheme.Δt, ml.substeps,
                nsteps * ml_steps[i-1], npaths[i], xc, xf))
    push!(x, xc, xf)
  end
  x
end
sample(model::AbstractSDE, ml::MultilevelScheme{<:B200}, t0, x0, nsteps::Integer, npaths::Int) =
  sample(model, ml, t0, x0, nsteps, fill(npaths, length(ml.levels)))

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

"Multilevel estimator: all levels concurrently, one all-reduce per level.  Returns per-level sums (nlevels x 8)."
function estimate(model::AbstractSDE, ml::MultilevelScheme{<:B200}, p::Payoff, t0, x0, nsteps::Integer, npaths::Vector{Int})
  b = ml.scheme
  res = zeros(8, length(ml.levels))
  o = Ref(opts(b, t0))
  check(ccall((:sdeb200_mlmc_estimate, lib), Cint,
              (Ptr{Cvoid}, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Cint, Int64, Ptr{Int64}, Ref{Payoff}, Ptr{Float64}),
              handle(typeof(model)), o, params(model), state(x0), ml.substeps, first(ml.levels), length(ml.levels),
              nsteps, npaths, Ref(p), res))
  res
end

# -- multi-GPU: one Julia worker per GPU (Distributed.addprocs(ngpus)); rank 0 creates the id, all ranks join ------------------
comm_unique_id() = (id = zeros(UInt8, 128); check(ccall((:sdeb200_comm_unique_id, lib), Cint, (Ptr{UInt8},), id)); id)
comm_init(id::Vector{UInt8}, nranks, rank) = check(ccall((:sdeb200_comm_init, lib), Cint, (Ptr{UInt8}, Cint, Cint), id, nranks, rank))

end # module
