# NextGPB200.jl — Julia-side shim that puts libngp.so (B200, sm_100a) behind NextGP.jl's marker-set sampler.
#
# UNTESTED IN THIS REPOSITORY: the build image has no Julia (SURVEY F5).  The same C ABI (include/ngp.h) is exercised
# end to end from Python (tests/test_gpu_parity.py); this file is the binding a NextGP.jl maintainer would add.
#
# Two levels, as in INTEGRATION.md:
#   sweep level : b200_funct(h, set)  -> a function with the signature of M[mSet].funct (functions.jl:118,157,197)
#   run level   : runLMEM_b200(...)   -> replaces MCMC.runLMEM (MCMC.jl:31-41) for intercept + SNP(...) models:
#                 prep / getMME! stay NextGP's own; only runSampler! is replaced.
module NextGPB200

using NextGP, DelimitedFiles

const libngp = get(ENV, "LIBNGP", joinpath(@__DIR__, "..", "nextgp.jl_b200", "libngp.so"))

const NGP_BAYESPR, NGP_BAYESB, NGP_BAYESC, NGP_BAYESR = Cint(0), Cint(1), Cint(2), Cint(3)
const NGP_GENO_F64, NGP_STORE_I8 = Cint(1), Cint(0)
const NGP_MAX_SETS = 8

struct NgpPrior            # mirrors struct ngp_prior
    method::Cint; est_pi::Cint
    df::Cdouble; scale::Cdouble; var_init::Cdouble; pi_in::Cdouble
    n_regions::Int64
    region_off::Ptr{Int64}; lhs0::Ptr{Cdouble}; rhs0::Ptr{Cdouble}
    n_class::Cint; pad::Cint                         # BayesR only (mme.jl:374-383)
    v_class::Ptr{Cdouble}; pi_class::Ptr{Cdouble}
end

struct NgpJointPrior       # mirrors struct ngp_joint_prior: (:M1,:M2) => BayesPR(r, V), mme.jl:448-489
    k::Cint
    set_id::NTuple{NGP_MAX_SETS,Cint}
    pad::Cint
    df::Cdouble
    scale::Ptr{Cdouble}; var_init::Ptr{Cdouble}
    n_regions::Int64
    region_off::Ptr{Int64}
end

mutable struct NgpState    # mirrors struct ngp_state
    n::Int64; n_sets::Cint; pad::Cint
    e::Ptr{Cdouble}; mu::Cdouble; varE::Cdouble; iter::Int64
    beta::NTuple{NGP_MAX_SETS,Ptr{Cdouble}}
    delta::NTuple{NGP_MAX_SETS,Ptr{Int64}}
    varBeta::NTuple{NGP_MAX_SETS,Ptr{Cdouble}}
    pi::NTuple{2 * NGP_MAX_SETS,Cdouble}
end

function check(h, rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:ngp_last_error, libngp), Cstring, (Ptr{Cvoid},), h))
    error("libngp error $rc: $msg")          # Julia exceptions, like mme.jl:77,343
end

function create(device::Integer = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:ngp_create, libngp), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h)
    rc == 0 || error("libngp: " * unsafe_string(ccall((:ngp_last_error, libngp), Cstring, (Ptr{Cvoid},), C_NULL)))
    return h[]
end
destroy(h) = ccall((:ngp_destroy, libngp), Cint, (Ptr{Cvoid},), h)

"""Upload M[pSet].data of getMME! (centred Matrix{Float64}): the codes are data .+ column means of the raw file;
pass the RAW 0/1/2 matrix (before prepMatVec.jl:129) — the library centres on device."""
function upload!(h, set::Integer, raw::Matrix{Float64})
    n, p = size(raw)
    check(h, ccall((:ngp_upload_genotypes, libngp), Cint,
                   (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Cint, Int64, Cint),
                   h, set, n, p, raw, NGP_GENO_F64, n, NGP_STORE_I8))
end

function method_id(name::String)
    name == "BayesPR" && return NGP_BAYESPR
    name == "BayesB" && return NGP_BAYESB
    name == "BayesC" && return NGP_BAYESC
    name == "BayesR" && return NGP_BAYESR
    name == "BayesLV" && return NGP_BAYESPR          # the single-site loop of BayesLV is BayesPR with one region per locus (b200_bayeslv_funct)
    error("$name is not on the B200 hot path (stays in Julia)")
end

"""Translate one entry of the NamedTuple dictionary M returned by getMME! (mme.jl:598-601) into ngp_set_prior."""
function set_prior!(h, set::Integer, Mset, v0::Float64)
    offs = Int64[first(r) - 1 for r in Mset.regionArray]; push!(offs, last(Mset.regionArray[end]))
    isPR = Mset.method == "BayesPR" || Mset.method == "BayesLV"
    isR = Mset.method == "BayesR"
    pi_in = (isPR || isR) ? 0.0 : Mset.piHat[2]
    vclass = isR ? Vector{Float64}(Mset.vClass) : Float64[]
    piclass = isR ? Vector{Float64}(Mset.piHat) : Float64[]
    GC.@preserve offs vclass piclass begin
        pr = NgpPrior(method_id(Mset.method), isPR ? 0 : Cint(Mset.estPi), Mset.df, Mset.scale, v0, pi_in,
                      isPR ? length(offs) - 1 : 0, isPR ? pointer(offs) : C_NULL, pointer(Mset.lhs), pointer(Mset.rhs),
                      Cint(length(vclass)), Cint(0), isR ? pointer(vclass) : C_NULL, isR ? pointer(piclass) : C_NULL)
        check(h, ccall((:ngp_set_prior, libngp), Cint, (Ptr{Cvoid}, Cint, Ref{NgpPrior}), h, set, pr))
    end
end

"""E.str == "D" (mme.jl:70-73): hand E.iVarStr = inv.(D) to the device once, after the uploads; also needed for the sweep-level
drop-in, whose Mp / mpm are the weighted ones of mme.jl:299-303.  `nothing` returns to "I"."""
function set_residual_weights!(h, iVarStr)
    if iVarStr === nothing || isempty(iVarStr)
        check(h, ccall((:ngp_set_residual_weights, libngp), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64), h, C_NULL, 0))
    else
        w = Vector{Float64}(iVarStr)
        check(h, ccall((:ngp_set_residual_weights, libngp), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64), h, w, length(w)))
    end
end

"""Sweep-level drop-in: returns a function with the signature of M[mSet].funct (samplers.jl:52)."""
function b200_funct(h, set::Integer)
    return function (mSet, M, beta, delta, ycorr, varE, varBeta)
        pos = M[mSet].pos
        b = vec(beta[pos]); d = vec(delta[pos]); vb = varBeta[mSet]
        piHat = haskey(M[mSet], :piHat) ? vec(M[mSet].piHat) : C_NULL
        check(h, ccall((:ngp_sweep, libngp), Cint,
                       (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Int64}, Ptr{Cdouble}, Ptr{Cdouble}),
                       h, set, ycorr, varE, b, d, vb, piHat))
        haskey(M[mSet], :logPi) && (M[mSet].logPi .= log.(M[mSet].piHat))
        nothing
    end
end

"""BayesLV (functions.jl:421-486): the single-site loop (:431-443) is the BayesPR sweep with one region per locus — getMME! already
sets M[pSet].regionArray = [r:r ...] (mme.jl:421), so `set_prior!` needs `method_id("BayesLV") == NGP_BAYESPR` — and runs on the device;
the model of the log-variances (:446-485) stays the reference's own code, copied here verbatim in spirit: the closure calls
`NextGP`'s sampleBayesLV! tail through `lv_tail!`, which the caller supplies (e.g. the body of functions.jl:446-485 cut into a function)."""
function b200_bayeslv_funct(h, set::Integer, lv_tail!::Function)
    return function (mSet, M, beta, delta, ycorr, varE, varBeta)
        pos = M[mSet].pos
        b = vec(beta[pos]); d = vec(delta[pos])
        scratch = copy(varBeta[mSet])                 # the device appends its own scaled-inverse-chi-square draw to the sweep: discarded
        check(h, ccall((:ngp_sweep, libngp), Cint,
                       (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Int64}, Ptr{Cdouble}, Ptr{Cdouble}),
                       h, set, ycorr, varE, b, d, scratch, C_NULL))
        lv_tail!(mSet, M, beta, varBeta)              # functions.jl:446-485: slice update of varBeta[mSet][locus] / logVar, c, SNPVARRESID, varZeta
        nothing
    end
end

"""Run-level BayesLV: after every `ngp_run(h, 1)` push the host's variances back (the device drew its own)."""
set_var_beta!(h, set::Integer, vb::Vector{Float64}) =
    check(h, ccall((:ngp_set_var_beta, libngp), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), h, set, vb))

"""sampleΛ2!(Λ2,Xc,yCorr,σ2τ,σ2ϵ,pMeans) of the GRN sampler (GRN.jl:150-164) on the device.  Upload X' once (individuals x SNPs, RAW
codes: the device centres like GRN.jl:23) as a marker set with a one-region BayesPR prior; per gene the right-hand-side offset
α·pMeans[g] becomes rhs0 = pMeans[g]/σ2τ[g] and the improper effect prior (the reference's LHS holds no prior precision) varBeta = Inf."""
function sample_lambda2_b200!(h, set::Integer, Λ2::Matrix{Float64}, yCorr::Matrix{Float64}, σ2τ, σ2ϵ::Float64, pMeans)
    nGenes, nSNPs = size(Λ2)
    ones64 = ones(Int64, nSNPs)
    for g in 1:nGenes
        rhs0 = fill(pMeans[g] / σ2τ[g], nSNPs)
        check(h, ccall((:ngp_set_marker_summary, libngp), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), h, set, C_NULL, rhs0))
        y = yCorr[g, :]; b = Λ2[g, :]; vb = [Inf]    # rows of column-major matrices: copied (keep yCorr individuals x genes to avoid it)
        check(h, ccall((:ngp_sweep, libngp), Cint,
                       (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Int64}, Ptr{Cdouble}, Ptr{Cdouble}),
                       h, set, y, σ2ϵ, b, ones64, vb, C_NULL))
        yCorr[g, :] .= y; Λ2[g, :] .= b
    end
    nothing
end

"""Tuple of correlated marker sets (multi-breed): sets 0..k-1 of the handle are the members, in the order of the tuple key.
`Mt` is M[(:M1,:M2,...)] of getMME! (mme.jl:448-489), `V0` the k x k prior covariance (priorVCV[pSet].v)."""
function set_joint_prior!(h, Mt, V0::Matrix{Float64})
    k = size(V0, 1)
    offs = Int64[first(r) - 1 for r in Mt.regionArray]; push!(offs, last(Mt.regionArray[end]))
    sc = Matrix{Float64}(permutedims(Mt.scale)); v0 = Matrix{Float64}(permutedims(V0))       # row-major for the C side
    GC.@preserve offs sc v0 begin
        pr = NgpJointPrior(Cint(k), ntuple(i -> Cint(i <= k ? i - 1 : 0), NGP_MAX_SETS), Cint(0), Mt.df, pointer(sc), pointer(v0),
                           length(offs) - 1, pointer(offs))
        check(h, ccall((:ngp_set_joint_prior, libngp), Cint, (Ptr{Cvoid}, Ref{NgpJointPrior}), h, pr))
    end
end

"""Sweep-level drop-in for sampleBayesPR!(mSet::Tuple, ...) (functions.jl:140-154): same 7 arguments."""
function b200_joint_funct(h)
    return function (mSet, M, beta, delta, ycorr, varE, varBeta)
        pos = M[mSet].pos
        k = length(pos); p = length(beta[pos[1]])
        B = Matrix{Float64}(undef, p, k)                       # column b = breed b  ==  k x p row-major on the C side
        for (b, ps) in enumerate(pos); B[:, b] .= vec(beta[ps]); end
        R = length(varBeta[mSet])
        VB = Array{Float64}(undef, k, k, R)
        for r in 1:R; VB[:, :, r] .= permutedims(varBeta[mSet][r]); end
        check(h, ccall((:ngp_joint_sweep, libngp), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
                       h, ycorr, varE, B, VB))
        for (b, ps) in enumerate(pos); beta[ps] .= reshape(B[:, b], size(beta[ps])); end
        for r in 1:R; varBeta[mSet][r] = permutedims(VB[:, :, r]); end
        nothing
    end
end

"""Run-level replacement of samplers.runSampler! (samplers.jl:23-106) for X = intercept only, Z = empty.
Same positional arguments; writes the same <outPut>/*Out rows in the order of samplers.jl:57-103."""
function runSampler_b200!(h, ycorr, nData, E, X, b, Z, u, varU, M, beta, varBeta, delta, chainLength, burnIn, outputFreq, outPut;
                          raw::Dict, seed::UInt64 = UInt64(0), chain::UInt32 = UInt32(0))
    isempty(Z) || error("random effects stay in Julia: use b200_funct at sweep level")
    sets = collect(keys(M))
    for (s, mSet) in enumerate(sets)
        upload!(h, s - 1, raw[mSet]); set_prior!(h, s - 1, M[mSet], varBeta[mSet][1])
    end
    check(h, ccall((:ngp_set_phenotype, libngp), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64), h, ycorr, nData))
    check(h, ccall((:ngp_set_residual_prior, libngp), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), h, E.df, E.scale))
    if E.str == "D"      # mme.jl:70-73: iVarStr = inv.(D)
        check(h, ccall((:ngp_set_residual_weights, libngp), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64), h, Vector{Float64}(E.iVarStr), nData))
    end
    check(h, ccall((:ngp_set_intercept, libngp), Cint, (Ptr{Cvoid}, Cint, Cdouble, Cdouble), h, isempty(X) ? 0 : 1, 0.0, 0.0))
    check(h, ccall((:ngp_set_rng, libngp), Cint, (Ptr{Cvoid}, UInt64, UInt32), h, seed, chain))
    these2Keep = collect((burnIn + outputFreq):outputFreq:chainLength)
    done = 0
    if burnIn > 0     # the device's posterior sums (ngp_get_posterior) count EVERY iteration since the last reset: start them after burn-in
        check(h, ccall((:ngp_run, libngp), Cint, (Ptr{Cvoid}, Int32), h, burnIn)); done = burnIn
        check(h, ccall((:ngp_reset_posterior, libngp), Cint, (Ptr{Cvoid},), h))
    end
    for it in these2Keep
        check(h, ccall((:ngp_run, libngp), Cint, (Ptr{Cvoid}, Int32), h, it - done)); done = it
        st = NgpState(nData, length(sets), 0, pointer(ycorr), 0.0, 0.0, 0,
                      ntuple(i -> i <= length(sets) ? pointer(beta[M[sets[i]].pos]) : Ptr{Cdouble}(C_NULL), NGP_MAX_SETS),
                      ntuple(i -> i <= length(sets) ? pointer(delta[M[sets[i]].pos]) : Ptr{Int64}(C_NULL), NGP_MAX_SETS),
                      ntuple(i -> i <= length(sets) ? pointer(varBeta[sets[i]]) : Ptr{Cdouble}(C_NULL), NGP_MAX_SETS),
                      ntuple(_ -> 0.0, 2 * NGP_MAX_SETS))
        check(h, ccall((:ngp_get_state, libngp), Cint, (Ptr{Cvoid}, Ref{NgpState}), h, st))
        isempty(b) || (b[1] = st.mu)
        NextGP.IO.outMCMC(outPut, "b", b'); NextGP.IO.outMCMC(outPut, "varE", st.varE)
        for (s, mSet) in enumerate(sets)
            NextGP.IO.outMCMC(outPut, "beta$mSet", beta[M[mSet].pos]); NextGP.IO.outMCMC(outPut, "delta$mSet", delta[M[mSet].pos])
            if M[mSet].method in ("BayesB", "BayesC")
                M[mSet].piHat .= [st.pi[2s-1] st.pi[2s]]; NextGP.IO.outMCMC(outPut, "pi$mSet", [M[mSet].piHat])
            end
        end
        for mSet in sets
            NextGP.IO.outMCMC(outPut, "var$mSet", hcat(reduce(hcat, varBeta[mSet])...))
        end
    end
    done < chainLength && check(h, ccall((:ngp_run, libngp), Cint, (Ptr{Cvoid}, Int32), h, chainLength - done))
    nothing
end

end # module
