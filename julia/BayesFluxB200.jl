# BayesFluxB200.jl -- the reference-side binding of libbfx (include/bfx.h).
#
# This is the stub a BayesFlux maintainer would add: it walks the reference's own objects
# (BNN, likelihood, prior, sampler -- field names as in enweg/BayesFlux.jl v0.2.4) into the plain-C
# descriptors of include/bfx.h and `ccall`s the entry points.  One `ccall` per `mcmc` call: the
# whole loop of src/inference/mcmc/abstract.jl:117-124 runs on the device.
#
# Julia is not installed in the build image, so this file is syntax-reviewed only; every symbol
# it binds is exercised by the Python ctypes mirror (bayesflux.jl_b200/_lib.py), which uses the
# same structs field for field.
#
# Surface (BASELINE.json north_star): DeviceBNN(bnn), mcmc (SGLD / SGNHT / SGNHTS / GGMC / HMC, incl.
# continue_sampling on the kept device state), find_mode, loglikeprior, ∇loglikeprior, predict /
# sample_posterior_predict, diagnostics; multi-GPU through `ngpus` / `shard_mode` and the communicator
# entries.  Reference tests expected to pass unchanged with `bnn` replaced by `DeviceBNN(bnn)`:
# test/sgld.jl, test/sgnht.jl, test/sgnht-s.jl, test/ggmc.jl (DiagCov / RMSProp / Fixed mass adapters),
# test/hmc.jl (with a diagonal mass adapter: FullCovMassAdapter is rejected), test/modes.jl; the
# calculate_epochs / resume assertions (test/sgld.jl:33-35) read sampler.samples / nsampled, which are
# written back below.
module BayesFluxB200

using BayesFlux
using Flux
using Distributions
using LinearAlgebra: Diagonal

const libbfx = get(ENV, "LIBBFX", "libbfx.so")

# ---- include/bfx.h structs (isbits, C layout) ----------------------------------------
struct LayerDesc
    kind::Int32   # 0 Dense, 1 RNN, 2 LSTM
    nin::Int32
    nout::Int32
    act::Int32    # 0 identity, 1 relu, 2 tanh, 3 sigmoid
end

struct LikeDesc
    kind::Int32              # 0 Normal, 1 TDist
    sigma_prior_kind::Int32  # 0 Gamma(shape, scale), 1 InverseGamma(shape, scale)
    nu::Float64
    sigma_prior_a::Float64
    sigma_prior_b::Float64
end

struct PriorDesc
    kind::Int32   # 0 GaussianPrior, 1 MixtureScalePrior
    _pad::Int32
    sigma0::Float64
    sigma1::Float64
    sigma2::Float64
    pi1::Float64
end

struct SamplerDesc
    kind::Int32; steps::Int32; path_len::Int32; sadapter_kind::Int32; madapter_kind::Int32
    da_t0::Int32; da_adapt_steps::Int32; ma_adapt_steps::Int32; ma_windowlength::Int32; _pad::Int32
    stepsize_a::Float32; stepsize_b::Float32; stepsize_gamma::Float32; min_stepsize::Float32
    maxnorm::Float32; l::Float32; A::Float32; sigmaA::Float32; xi::Float32; mu::Float32; beta::Float32
    da_target_accept::Float32; da_gamma::Float32; da_kappa::Float32
    ma_kappa::Float32; ma_epsilon::Float32; ma_lambda::Float32; ma_alpha::Float32
    # BFX_MODE (find_mode): Flux optimiser rule + FluxModeFinder window
    opt_kind::Int32; mode_windowlength::Int32
    opt_eta::Float32; opt_rho::Float32; opt_beta2::Float32; opt_eps::Float32
    mode_abstol::Float32; _pad2::Float32
end
# the five samplers do not use the find_mode fields
SamplerDesc(a::Vararg{Any,28}) = SamplerDesc(a..., Int32(3), Int32(100), 0.001f0, 0.9f0, 0.999f0, 1f-8, 1f-6, 0f0)

# FluxModeFinder (src/inference/mode/flux.jl:8-31) with Flux.ADAM / RMSProp / Momentum / Descent
function samplerdesc(f::FluxModeFinder)
    o = f.opt
    kind, eta, rho, b2, eps =
        o isa Flux.Optimise.ADAM ? (Int32(3), o.eta, o.beta[1], o.beta[2], o.epsilon) :
        o isa Flux.Optimise.RMSProp ? (Int32(2), o.eta, o.rho, 0.999, o.epsilon) :
        o isa Flux.Optimise.Momentum ? (Int32(1), o.eta, o.rho, 0.999, 1e-8) :
        o isa Flux.Optimise.Descent ? (Int32(0), o.eta, 0.9, 0.999, 1e-8) :
        error("optimiser $(typeof(o)) is not supported by the B200 engine")
    SamplerDesc(5, 1, 1, 0, 0, 10, 0, 0, 2, 0, 0f0, 0f0, 0f0, -Inf32, 0f0, 0f0, 1f0, 1f0, 0f0, 1f0, 0.55f0,
                0.65f0, 0.05f0, 0.75f0, 0f0, 0f0, 0f0, 0f0,
                kind, Int32(f.windowlength), Float32(eta), Float32(rho), Float32(b2), Float32(eps), Float32(f.ϵ), 0f0)
end

struct RunDesc
    batchsize::Int64
    numsamples::Int64
    shuffle::Int32
    partial::Int32
    continue_sampling::Int32
    keep_every::Int32
    batch_mode::Int32
    _pad::Int32
    seed::UInt64
    chain_offset::Int64
    ngpus::Int32        # ranks / devices taking part
    shard_mode::Int32   # 0 none, 1 chains sharded (no collective), 2 one chain, minibatch sharded + all-reduce
end
RunDesc(b, n, sh, pa, cs, ke, bm, pad, seed, off) = RunDesc(b, n, sh, pa, cs, ke, bm, pad, seed, off, Int32(1), Int32(0))

# state fields of bfx_sampler_get_state / set_state
const STATE_THETA, STATE_MOMENTUM, STATE_MINV, STATE_XI, STATE_STEPSIZE, STATE_T, STATE_NSAMPLED, STATE_STATUS,
      STATE_CONVERGED = Int32.(0:8)

function check(code::Int32)
    code == 0 && return
    buf = Vector{UInt8}(undef, 512)
    ccall((:bfx_last_error, libbfx), Int32, (Ptr{UInt8}, Csize_t), buf, length(buf))
    msg = unsafe_string(pointer(buf))
    # the reference raises ErrorException("NaN in g") / ("NaN in θ") (sgnht.jl:70-72,81-83)
    error(msg)
end

# ---- object translation ----------------------------------------------------------------
actcode(f) = f === identity ? 0 : f === Flux.relu ? 1 : (f === tanh || f === Flux.tanh_fast) ? 2 :
             (f === Flux.sigmoid || f === Flux.sigmoid_fast) ? 3 :
             error("activation $f is not supported by the B200 engine")

function layerdesc(l)
    if l isa Flux.Dense
        return LayerDesc(0, size(l.weight, 2), size(l.weight, 1), actcode(l.σ))
    elseif l isa Flux.Recur && l.cell isa Flux.RNNCell
        return LayerDesc(1, size(l.cell.Wi, 2), size(l.cell.Wh, 2), actcode(l.cell.σ))
    elseif l isa Flux.Recur && l.cell isa Flux.LSTMCell
        return LayerDesc(2, size(l.cell.Wi, 2), size(l.cell.Wh, 2), 2)
    end
    error("layer $(typeof(l)) is not supported by the B200 engine")
end

sigmaprior(d::Gamma) = (Int32(0), Float64(shape(d)), Float64(scale(d)))
sigmaprior(d::InverseGamma) = (Int32(1), Float64(shape(d)), Float64(scale(d)))

likedesc(l::Union{FeedforwardNormal,SeqToOneNormal}) = LikeDesc(0, sigmaprior(l.prior_σ)[1], 0.0, sigmaprior(l.prior_σ)[2:3]...)
likedesc(l::Union{FeedforwardTDist,SeqToOneTDist}) = LikeDesc(1, sigmaprior(l.prior_σ)[1], Float64(l.ν), sigmaprior(l.prior_σ)[2:3]...)

priordesc(p::GaussianPrior) = PriorDesc(0, 0, Float64(p.σ0), 1.0, 1.0, 0.5)
priordesc(p::MixtureScalePrior) = PriorDesc(1, 0, 1.0, Float64(p.σ1), Float64(p.σ2), Float64(p.π1))

madapter_fields(::FixedMassAdapter) = (Int32(0), Int32(0), Int32(2), 0f0, 0f0, 0f0, 0f0)
madapter_fields(m::DiagCovMassAdapter) = (Int32(1), Int32(m.adapt_steps), Int32(m.windowlength), Float32(m.kappa), Float32(m.epsilon), 0f0, 0f0)
madapter_fields(m::RMSPropMassAdapter) = (Int32(2), Int32(m.adapt_steps), Int32(2), 0f0, 0f0, Float32(m.λ), Float32(m.α))
madapter_fields(::FullCovMassAdapter) = (Int32(3), Int32(0), Int32(2), 0f0, 0f0, 0f0, 0f0)  # rejected by the library

sadapter_fields(s::ConstantStepsize) = (Int32(0), Int32(10), Int32(0), 0.65f0, 0.05f0, 0.75f0, Float32(s.l))
sadapter_fields(s::DualAveragingStepSize) = (Int32(1), Int32(s.t), Int32(s.adapt_steps), Float32(s.target_accept), Float32(s.gamma), Float32(s.kappa), Float32(s.l))

function samplerdesc(s::SGLD)
    SamplerDesc(0, 1, 1, 0, 0, 10, 0, 0, 2, 0, s.stepsize_a, s.stepsize_b, s.stepsize_γ, s.min_stepsize, s.maxnorm,
                0f0, 1f0, 1f0, 0f0, 1f0, 0.55f0, 0.65f0, 0.05f0, 0.75f0, 0f0, 0f0, 0f0, 0f0)
end
function samplerdesc(s::SGNHT)
    SamplerDesc(1, 1, 1, 0, 0, 10, 0, 0, 2, 0, 0f0, 0f0, 0f0, -Inf32, 0f0, s.l, s.A, 1f0, s.xi, 1f0, 0.55f0,
                0.65f0, 0.05f0, 0.75f0, 0f0, 0f0, 0f0, 0f0)
end
function samplerdesc(s::SGNHTS)
    mk, mas, mw, kap, eps, lam, alp = madapter_fields(s.madapter)
    SamplerDesc(2, 1, 1, 0, mk, 10, 0, mas, mw, 0, 0f0, 0f0, 0f0, -Inf32, 0f0, s.l, 1f0, s.σA, s.xi, s.μ, 0.55f0,
                0.65f0, 0.05f0, 0.75f0, kap, eps, lam, alp)
end
function samplerdesc(s::GGMC)
    mk, mas, mw, kap, eps, lam, alp = madapter_fields(s.madapter)
    sk, t0, sas, tgt, gam, kp, l = sadapter_fields(s.sadapter)
    SamplerDesc(3, s.steps, 1, sk, mk, t0, sas, mas, mw, 0, 0f0, 0f0, 0f0, -Inf32, s.maxnorm, l, 1f0, 1f0, 0f0, 1f0,
                s.β, tgt, gam, kp, kap, eps, lam, alp)
end
function samplerdesc(s::HMC)
    mk, mas, mw, kap, eps, lam, alp = madapter_fields(s.madapter)
    sk, t0, sas, tgt, gam, kp, l = sadapter_fields(s.sadapter)
    SamplerDesc(4, 1, s.path_len, sk, mk, t0, sas, mas, mw, 0, 0f0, 0f0, 0f0, -Inf32, s.maxnorm, l, 1f0, 1f0, 0f0, 1f0,
                0.55f0, tgt, gam, kp, kap, eps, lam, alp)
end

# ---- device model (destruct + BNN on the GPU) -------------------------------------------
mutable struct DeviceBNN
    handle::Ptr{Cvoid}
    bnn::BNN
end

function DeviceBNN(bnn::BNN; device::Integer=0)
    net = bnn.like.nc(bnn.like.nc.θ)                      # the Flux.Chain destruct() walked
    layers = [layerdesc(l) for l in net.layers]
    like, prior = Ref(likedesc(bnn.like)), Ref(priordesc(bnn.prior))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:bfx_model_create, libbfx), Int32,
                (Ptr{LayerDesc}, Int32, Ref{LikeDesc}, Ref{PriorDesc}, Int32, Ref{Ptr{Cvoid}}),
                layers, length(layers), like, prior, device, h))
    x = Array{Float32}(bnn.x); y = Vector{Float32}(vec(bnn.y))
    dims = Int64[size(x)...]
    check(ccall((:bfx_model_set_data, libbfx), Int32,
                (Ptr{Cvoid}, Ptr{Float32}, Ptr{Int64}, Int32, Ptr{Float32}, Int64),
                h[], x, dims, ndims(x), y, length(y)))
    m = DeviceBNN(h[], bnn)
    finalizer(m -> ccall((:bfx_model_destroy, libbfx), Int32, (Ptr{Cvoid},), m.handle), m)
    return m
end

# ∇loglikeprior(bnn, θ, x, y; num_batches) on the observations `idx` (1-based) -- src/model/BNN.jl:61-69
function BayesFlux.∇loglikeprior(m::DeviceBNN, θ::AbstractVector{Float32}, idx::AbstractVector{<:Integer}; num_batches=1)
    v = Ref{Float64}(0.0); g = similar(θ); idx0 = Int64.(idx) .- 1
    check(ccall((:bfx_grad_loglikeprior, libbfx), Int32,
                (Ptr{Cvoid}, Ptr{Float32}, Int32, Ptr{Int64}, Int64, Int64, Float32, Ref{Float64}, Ptr{Float32}),
                m.handle, θ, 1, idx0, length(idx0), 0, Float32(num_batches), v, g))
    return v[], g
end

# loglikeprior(bnn, θ, x, y; num_batches) on the observations `idx` (1-based) -- src/model/BNN.jl:50-56
function BayesFlux.loglikeprior(m::DeviceBNN, θ::AbstractVector{Float32}, idx::AbstractVector{<:Integer}; num_batches=1)
    v = Ref{Float64}(0.0); idx0 = Int64.(idx) .- 1
    check(ccall((:bfx_loglikeprior, libbfx), Int32,
                (Ptr{Cvoid}, Ptr{Float32}, Int32, Ptr{Int64}, Int64, Int64, Float32, Ref{Float64}),
                m.handle, θ, 1, idx0, length(idx0), 0, Float32(num_batches), v))
    return v[]
end
# the whole data set (what GGMC / HMC evaluate in their accept steps, ggmc.jl:147,176, hmc.jl:135-136)
function BayesFlux.loglikeprior(m::DeviceBNN, θ::AbstractVector{Float32}; num_batches=1)
    v = Ref{Float64}(0.0)
    check(ccall((:bfx_loglikeprior, libbfx), Int32,
                (Ptr{Cvoid}, Ptr{Float32}, Int32, Ptr{Int64}, Int64, Int64, Float32, Ref{Float64}),
                m.handle, θ, 1, C_NULL, 0, 0, Float32(num_batches), v))
    return v[]
end

# ---- the device state of a sampler -------------------------------------------------------
# The reference's samplers are mutable structs without a spare field, so the handle of the device-side state lives
# in a side table keyed by the sampler object.  It is what makes `continue_sampling=true` work (the chain state,
# adapters and counters stay on the device between calls) and it is destroyed by its finalizer.
mutable struct SamplerHandle
    handle::Ptr{Cvoid}
    model::DeviceBNN      # keeps the model alive as long as the sampler state
    nchains::Int
end
const HANDLES = IdDict{Any,SamplerHandle}()

function sampler_handle!(m::DeviceBNN, sampler, nchains::Int; fresh::Bool)
    h = get(HANDLES, sampler, nothing)
    if h !== nothing && (h.model !== m || h.nchains != nchains)
        fresh || error("continue_sampling: the sampler was run on another model / number of chains")
        finalize(h); h = nothing
    end
    if h === nothing
        fresh || error("continue_sampling needs a sampler that has already sampled on this DeviceBNN")
        d = Ref(samplerdesc(sampler)); p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:bfx_sampler_create, libbfx), Int32, (Ref{SamplerDesc}, Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}),
                    d, m.handle, nchains, p))
        h = SamplerHandle(p[], m, nchains)
        finalizer(x -> (x.handle != C_NULL && ccall((:bfx_sampler_destroy, libbfx), Int32, (Ptr{Cvoid},), x.handle);
                        x.handle = C_NULL), h)
        HANDLES[sampler] = h
    end
    return h
end

function get_state(h::SamplerHandle, field::Int32, ::Type{T}, n::Int) where {T}
    out = Vector{T}(undef, n)
    check(ccall((:bfx_sampler_get_state, libbfx), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int64),
                h.handle, field, out, sizeof(out)))
    return out
end

# what the reference's structs carry after a run and its tests / callers read back (chain 1 of the handle):
# nsampled, t, l (and the adapter's l), xi, the momentum (p / momentum) and the diagonal of Minv
function write_back!(sampler, h::SamplerHandle, P::Int)
    C = h.nchains
    sampler.nsampled = Int(get_state(h, STATE_NSAMPLED, Int64, C)[1])
    hasproperty(sampler, :t) && (sampler.t = Int(get_state(h, STATE_T, Int64, C)[1]))
    if hasproperty(sampler, :l)
        l = get_state(h, STATE_STEPSIZE, Float32, C)[1]
        sampler.l = l                                            # ggmc.jl:177, hmc.jl:140
        hasproperty(sampler, :sadapter) && hasproperty(sampler.sadapter, :l) && (sampler.sadapter.l = l)  # dualaveragestepsize.jl:59
    end
    hasproperty(sampler, :xi) && (sampler.xi = get_state(h, STATE_XI, Float32, C)[1])
    if hasproperty(sampler, :p) || hasproperty(sampler, :momentum)
        mom = reshape(get_state(h, STATE_MOMENTUM, Float32, P * C), P, C)[:, 1]
        hasproperty(sampler, :p) ? (sampler.p = mom) : (sampler.momentum = mom)
    end
    if hasproperty(sampler, :Minv)
        sampler.Minv = Diagonal(reshape(get_state(h, STATE_MINV, Float32, P * C), P, C)[:, 1])   # ggmc.jl:178-181, hmc.jl:141
    end
    return sampler
end

# mcmc(bnn, batchsize, numsamples, sampler; ...) -- src/inference/mcmc/abstract.jl:74-127
# Engine extensions: nchains (θstart may be P x C), keep_every, seed, and the multi-GPU partitioning of the run:
#   shard_mode = :chains     this process runs chains chain_offset+1 ... chain_offset+nchains of `ngpus` ranks (no collective)
#   shard_mode = :minibatch  ONE chain whose minibatch is sharded over `ngpus` ranks; every rank holds its shard in its
#                            DeviceBNN and has attached the communicator (comm_init!); batchsize is the local one
function BayesFlux.mcmc(m::DeviceBNN, batchsize::Int, numsamples::Int, sampler::MCMCState;
                        shuffle=true, partial=true, showprogress=false, continue_sampling=false,
                        θstart=vcat(m.bnn.init()...), nchains::Int=size(θstart, 2), keep_every::Int=1,
                        seed::Integer=rand(UInt64), chain_offset::Integer=0, ngpus::Integer=1, shard_mode::Symbol=:none)
    P = m.bnn.num_total_params
    h = sampler_handle!(m, sampler, nchains; fresh=!continue_sampling)
    mode = shard_mode === :minibatch ? 2 : shard_mode === :chains ? 1 : 0
    run = Ref(RunDesc(batchsize, numsamples, shuffle, partial, continue_sampling, keep_every, 0, 0, UInt64(seed),
                      Int64(chain_offset), Int32(ngpus), Int32(mode)))
    nnew = Ref{Int64}(0); nkept = Ref{Int64}(0)
    check(ccall((:bfx_mcmc_num_kept, libbfx), Int32, (Ptr{Cvoid}, Ref{RunDesc}, Ref{Int64}, Ref{Int64}), h.handle, run, nnew, nkept))
    θ0 = Matrix{Float32}(reshape(θstart, P, :))
    size(θ0, 2) == 1 && nchains > 1 && (θ0 = repeat(θ0, 1, nchains))
    samples = Array{Float32}(undef, P, nkept[], nchains)
    kinetic = sampler isa Union{SGNHT,SGNHTS} ? zeros(Float32, nnew[], nchains) : Float32[]
    accepted = sampler isa Union{GGMC,HMC} ? zeros(Int32, nnew[], nchains) : Int32[]
    code = GC.@preserve θ0 samples kinetic accepted ccall((:bfx_mcmc_run, libbfx), Int32,
        (Ptr{Cvoid}, Ref{RunDesc}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}),
        h.handle, run, continue_sampling ? C_NULL : pointer(θ0), samples,
        isempty(kinetic) ? C_NULL : pointer(kinetic), isempty(accepted) ? C_NULL : pointer(accepted))
    # write back what the reference's tests read.  initialise!(...; continue_sampling=true) keeps the drawn prefix
    # (sgld.jl:77-87): the new columns are appended to it
    if nchains == 1
        new = samples[:, :, 1]
        sampler.samples = continue_sampling ? hcat(sampler.samples[:, 1:(numsamples - nnew[])], new) : new
        if sampler isa Union{SGNHT,SGNHTS}
            sampler.kinetic = continue_sampling ? vcat(sampler.kinetic, vec(kinetic)) : vec(kinetic)
        elseif sampler isa GGMC
            acc = Int.(vec(accepted))
            sampler.accepted = continue_sampling ? vcat(sampler.accepted, acc) : acc          # ggmc.jl:186
        elseif sampler isa HMC
            acc = vec(accepted) .!= 0
            sampler.accepted = continue_sampling ? vcat(sampler.accepted, acc) : acc          # hmc.jl:146
        end
    end
    write_back!(sampler, h, P)
    check(code)    # "NaN in g" / "NaN in θ" like the reference (sgnht.jl:70-72,81-83); per-chain codes via STATE_STATUS
    return nchains == 1 ? sampler.samples : samples
end

# find_mode(bnn, batchsize, epochs, optimiser; shuffle, partial, showprogress) -- src/inference/mode/abstract.jl:36-86
# with FluxModeFinder + Descent / Momentum / RMSProp / ADAM (src/inference/mode/flux.jl:8-54).  `nrestarts` runs that
# many independent restarts at once (θstart P x nrestarts or one draw of bnn.init() per restart) and returns a matrix.
function BayesFlux.find_mode(m::DeviceBNN, batchsize::Int, epochs::Int, optimiser::FluxModeFinder;
                             shuffle=true, partial=true, showprogress=false, nrestarts::Int=1,
                             θstart=hcat([vcat(m.bnn.init()...) for _ in 1:nrestarts]...), seed::Integer=rand(UInt64))
    P = m.bnn.num_total_params
    n = length(m.bnn.y)
    nb = partial ? cld(n, min(batchsize, n)) : fld(n, min(batchsize, n))      # length(DataLoader)
    h = sampler_handle!(m, optimiser, nrestarts; fresh=true)
    run = Ref(RunDesc(batchsize, epochs * nb, shuffle, partial, false, 1, 0, 0, UInt64(seed), 0, Int32(1), Int32(0)))
    θ0 = Matrix{Float32}(reshape(θstart, P, :))
    check(GC.@preserve θ0 ccall((:bfx_mcmc_run, libbfx), Int32,
        (Ptr{Cvoid}, Ref{RunDesc}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}),
        h.handle, run, pointer(θ0), C_NULL, C_NULL, C_NULL))
    θ = reshape(get_state(h, STATE_THETA, Float32, P * nrestarts), P, nrestarts)
    conv = get_state(h, STATE_CONVERGED, Int32, nrestarts) .!= 0
    all(conv) ? (@info "Converged!") : (@info "Failed to converge.")            # mode/abstract.jl:77,84
    return nrestarts == 1 ? θ[:, 1] : θ
end

# net(x) for every posterior draw (columns of θs) and σ per draw: the deterministic part of posterior_predict
# (src/likelihoods/feedforward.jl:45-55,99-110, seq_to_one.jl:46-57,102-114); the random draw around it stays here
function predict(m::DeviceBNN, θs::AbstractMatrix{Float32}; x=nothing)
    C = size(θs, 2)
    n = x === nothing ? length(m.bnn.y) : size(x)[end]
    yhat = Matrix{Float32}(undef, n, C); σ = Vector{Float32}(undef, C)
    xs = x === nothing ? Float32[] : Array{Float32}(x)
    dims = x === nothing ? Int64[] : Int64[size(xs)...]
    check(GC.@preserve xs ccall((:bfx_predict, libbfx), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Int32, Ptr{Float32}, Ptr{Int64}, Int32, Ptr{Float32}, Ptr{Float32}),
        m.handle, Matrix{Float32}(θs), C, x === nothing ? C_NULL : pointer(xs), dims, length(dims), yhat, σ))
    return yhat, σ
end
# sample_posterior_predict(bnn, ch; x) -- src/model/BNN.jl:129-135
function BayesFlux.sample_posterior_predict(m::DeviceBNN, ch::AbstractMatrix{Float32}; x=nothing)
    yhat, σ = predict(m, ch; x=x)
    like = m.bnn.like
    noise = like isa Union{FeedforwardTDist,SeqToOneTDist} ? rand(TDist(like.ν), size(yhat)...) : randn(size(yhat)...)
    return yhat .+ Float32.(noise) .* σ'
end

# ---- multi-GPU (include/bfx.h: bfx_comm_unique_id / bfx_sampler_comm_init) --------------------------------------
# One Julia process per GPU (Distributed.jl, MPI.jl ...): rank 0 creates the id, the caller ships its 128 bytes to the
# other ranks, every rank builds its DeviceBNN on its shard of the data, initialises the sampler state with a zero-
# sample run and attaches the communicator; then mcmc(...; ngpus, shard_mode=:minibatch, continue_sampling=true).
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:bfx_comm_unique_id, libbfx), Int32, (Ptr{UInt8},), id))
    return id
end
function comm_init!(m::DeviceBNN, sampler, id::Vector{UInt8}, nranks::Integer, rank::Integer)
    h = sampler_handle!(m, sampler, 1; fresh=!haskey(HANDLES, sampler))
    check(ccall((:bfx_sampler_comm_init, libbfx), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), h.handle, id, nranks, rank))
end
comm_destroy!(sampler) = haskey(HANDLES, sampler) &&
    check(ccall((:bfx_sampler_comm_destroy, libbfx), Int32, (Ptr{Cvoid},), HANDLES[sampler].handle))
# ranks whose exchange buffers this sampler reads over NVLink peer memory (0: the NCCL all-reduce is in use)
function peer_ranks(sampler)
    n = Ref{Int32}(0)
    check(ccall((:bfx_sampler_peer_ranks, libbfx), Int32, (Ptr{Cvoid}, Ref{Int32}), HANDLES[sampler].handle, n))
    return Int(n[])
end

# R̂ / ESS / pooled moments of the samples `mcmc` returned (P × n or P × n × C, Float32), computed on the device.
# The reference hands `samples'` to MCMCChains for this (README.md:191-193).
function diagnostics(samples::Array{Float32}; max_lag::Integer=0, device::Integer=0)
    P, n = size(samples, 1), size(samples, 2)
    C = ndims(samples) == 3 ? size(samples, 3) : 1
    mean, var, rhat, ess = (Vector{Float32}(undef, P) for _ in 1:4)
    check(GC.@preserve samples ccall((:bfx_chain_diagnostics, libbfx), Int32,
        (Int32, Ptr{Float32}, Int32, Int64, Int64, Int64, Int64, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
        device, samples, 0, P, n, C, max_lag, mean, var, rhat, ess))
    return (mean=mean, var=var, rhat=rhat, ess=ess)
end

end # module
