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
module BayesFluxB200

using BayesFlux
using Flux
using Distributions

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
end

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

# mcmc(bnn, batchsize, numsamples, sampler; ...) -- src/inference/mcmc/abstract.jl:74-127
# Engine extensions: nchains (θstart may be P x C), keep_every, seed.
function BayesFlux.mcmc(m::DeviceBNN, batchsize::Int, numsamples::Int, sampler::MCMCState;
                        shuffle=true, partial=true, showprogress=false, continue_sampling=false,
                        θstart=vcat(m.bnn.init()...), nchains::Int=size(θstart, 2), keep_every::Int=1,
                        seed::Integer=rand(UInt64), state=nothing)
    P = m.bnn.num_total_params
    sh = state
    if sh === nothing
        d = Ref(samplerdesc(sampler)); h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:bfx_sampler_create, libbfx), Int32, (Ref{SamplerDesc}, Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}),
                    d, m.handle, nchains, h))
        sh = h[]
    end
    run = Ref(RunDesc(batchsize, numsamples, shuffle, partial, continue_sampling, keep_every, 0, 0, UInt64(seed), 0))
    nnew = Ref{Int64}(0); nkept = Ref{Int64}(0)
    check(ccall((:bfx_mcmc_num_kept, libbfx), Int32, (Ptr{Cvoid}, Ref{RunDesc}, Ref{Int64}, Ref{Int64}), sh, run, nnew, nkept))
    θ0 = Matrix{Float32}(reshape(θstart, P, :))
    size(θ0, 2) == 1 && nchains > 1 && (θ0 = repeat(θ0, 1, nchains))
    samples = Array{Float32}(undef, P, nkept[], nchains)
    kinetic = sampler isa Union{SGNHT,SGNHTS} ? zeros(Float32, nnew[], nchains) : Float32[]
    accepted = sampler isa Union{GGMC,HMC} ? zeros(Int32, nnew[], nchains) : Int32[]
    code = GC.@preserve θ0 samples kinetic accepted ccall((:bfx_mcmc_run, libbfx), Int32,
        (Ptr{Cvoid}, Ref{RunDesc}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}),
        sh, run, continue_sampling ? C_NULL : pointer(θ0), samples,
        isempty(kinetic) ? C_NULL : pointer(kinetic), isempty(accepted) ? C_NULL : pointer(accepted))
    # write back what the reference's tests read: samples, nsampled, t, kinetic, accepted
    if nchains == 1
        new = samples[:, :, 1]
        sampler.samples = continue_sampling ? hcat(sampler.samples[:, 1:sampler.nsampled], new) : new
        sampler isa Union{SGNHT,SGNHTS} && (sampler.kinetic = vec(kinetic))
    end
    sampler.nsampled = numsamples
    check(code)
    return nchains == 1 ? sampler.samples : samples
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
