# record_variates.jl — runs the UNMODIFIED NextGP.jl sampler (intercept + SNP marker sets: BayesPR / BayesB / BayesC) and writes
#   (i)  the variates it consumed, in the replay-log format of this repository (SURVEY §8c; include/ngp.h: ngp_replay), and
#   (ii) the chain it produced (per-iteration beta, delta, varE, mu, varBeta, pi and the final ycorr),
# so that REAL reference goldens can be produced wherever Julia + NextGP.jl are installed:
#
#     julia julia/record_variates.jl <genotypes.txt> <phenotypes.csv> <method> <out_dir> [iters=20] [seed=1] [pi=0.05] [v=<var(y)/2/p>]
#         genotypes.txt   the reference's marker file (prepMatVec.jl:116: space-delimited 0/1/2, no header)
#         phenotypes.csv  a column `y`
#         method          BayesPR | BayesB | BayesC
#     python tests/golden/import_julia_log.py <out_dir> tests/golden/julia/<case>      # -> the fixture the tests replay
#
# UNTESTED HERE (the build image has no Julia, SURVEY F5).  How it works — nothing of the sampler is restated:
# the reference draws from Julia's task-local default RNG.  Before every call of a reference function the script takes
# `copy(Random.default_rng())`; after the call it re-draws, from that copy, exactly the variates the call consumed, in the order the
# reference consumes them (functions.jl:118-137, 157-195, 197-236, 39-47, 523-525), using the call's OUTPUT to know the branches taken
# (a normal is drawn only for an included locus: delta == 1).  The copy must then be in the same state as the live RNG — that
# equality is asserted after every call, so a wrong assumption about the draw order cannot produce a log silently.
using NextGP, Distributions, DelimitedFiles, CSV, DataFrames, StatsModels, LinearAlgebra, Statistics
using Random: default_rng, seed!          # (not `using Random`: NextGP exports its own `Random(str, v)` prior constructor, runTime.jl:135-139)

const F = NextGP.MCMC.samplers.functions

same_state(a, b) = (a.s0, a.s1, a.s2, a.s3) == (b.s0, b.s1, b.s2, b.s3)       # Xoshiro256++ state words
live_rng() = copy(default_rng())

function main(args)
    geno, phen, method, out = args[1], args[2], args[3], args[4]
    iters = length(args) >= 5 ? parse(Int, args[5]) : 20
    seed = length(args) >= 6 ? parse(Int, args[6]) : 1
    piIn = length(args) >= 7 ? parse(Float64, args[7]) : 0.05
    data = CSV.read(phen, DataFrame)
    p = size(readdlm(geno, ' '; header = false), 2)
    v = length(args) >= 8 ? parse(Float64, args[8]) : var(data.y) / 2 / p
    prior = method == "BayesPR" ? BayesPR(9999, v) : method == "BayesB" ? BayesB(piIn, v; estimatePi = true) :
            method == "BayesC" ? BayesC(piIn, v; estimatePi = true) : error("method $method")
    priorVCV = Dict(:M => prior, :e => Random("I", var(data.y) / 2))
    f = eval(Meta.parse("@formula(y ~ 1 + SNP(M,\"$geno\"))"))
    mkpath(out)
    yVec, X, Z, M = NextGP.MCMC.prepMatVec.prep(f, data; userHints = Dict{Symbol,Any}(), path2ped = [], priorVCV = priorVCV)
    ycorr, nData, E, X, b, Z, u, varU, M, beta, varBeta, delta =
        NextGP.MCMC.mme.getMME!(yVec, X, Z, M, [], priorVCV, Dict{Any,Any}(), out)
    @assert isempty(Z) "random effects other than marker sets are not on the recorded path"
    mSet = first(keys(M)); Ms = M[mSet]; pos = Ms.pos
    nvar = length(varBeta[mSet])
    seed!(seed)
    io = Dict(k => open(joinpath(out, k * ".f64"), "w") for k in
              ("chi2_e", "z_mu", "u", "z", "chi2_b", "beta_pi", "varE", "mu", "beta", "varBeta", "pi"))
    iod = open(joinpath(out, "delta.i64"), "w")
    for iter in 1:iters
        # ---- residual variance (samplers.jl:32-35; functions.jl:523-525): one Chisq(df + n)
        r = live_rng()
        varE = F.sampleVarE(E.df, E.scale, ycorr, nData)
        chi2_e = rand(r, Chisq(E.df + nData))
        @assert same_state(r, live_rng()) "sampleVarE consumed something else than one Chisq draw"
        @assert isapprox(varE, (E.df * E.scale + dot(ycorr, ycorr)) / chi2_e; rtol = 1e-12)
        # ---- intercept (samplers.jl:37-39; functions.jl:39-47): one normal = mean + sd * randn()
        z_mu = 0.0
        for xSet in keys(X)
            @assert length(b[X[xSet].pos]) == 1 "only the intercept is recorded"
            r = live_rng()
            F.sampleX!(xSet, X, b, ycorr, varE)
            z_mu = randn(r)
            @assert same_state(r, live_rng()) "sampleX! consumed something else than one normal draw"
        end
        # ---- the marker sweep: M[mSet].funct (samplers.jl:52)
        r = live_rng()
        Ms.funct(mSet, M, beta, delta, ycorr, varE, varBeta)
        uu = zeros(p); zz = zeros(p); cb = zeros(nvar); bpi = 0.0
        d = vec(delta[pos]); bt = vec(beta[pos])
        if Ms.method == "BayesPR"                       # functions.jl:118-137: per region: a normal per locus, then one Chisq
            for (rg, loci) in enumerate(Ms.regionArray)
                for j in loci; zz[j] = randn(r); end
                cb[rg] = rand(r, Chisq(Ms.df + length(loci)))
            end
        elseif Ms.method == "BayesC"                    # functions.jl:197-236: rand() per locus, a normal if included; Chisq(df + nLoci); Beta
            for j in 1:p
                uu[j] = rand(r)
                if d[j] == 1; zz[j] = randn(r); end
            end
            nLoci = sum(d)
            cb[1] = rand(r, Chisq(Ms.df + nLoci))
            if Ms.estPi; bpi = rand(r, Beta(nLoci + 1, p - nLoci + 1)); end
        elseif Ms.method == "BayesB"                    # functions.jl:157-195: rand(); if included a normal and Chisq(df + 1); Beta
            for j in 1:p
                uu[j] = rand(r)
                if d[j] == 1; zz[j] = randn(r); cb[j] = rand(r, Chisq(Ms.df + 1)); end
            end
            nLoci = sum(d)
            if Ms.estPi; bpi = rand(r, Beta(nLoci + 1, p - nLoci + 1)); end
        else
            error("method $(Ms.method) is not recorded")
        end
        @assert same_state(r, live_rng()) "the sweep consumed its variates in another order than assumed (iteration $iter)"
        write(io["chi2_e"], chi2_e); write(io["z_mu"], z_mu); write(io["u"], uu); write(io["z"], zz); write(io["chi2_b"], cb)
        write(io["beta_pi"], bpi); write(io["varE"], varE); write(io["mu"], Float64(b[1][1])); write(io["beta"], bt)
        write(io["varBeta"], Float64.(varBeta[mSet])); write(io["pi"], Float64.(vec(Ms.piHat))); write(iod, Int64.(d))
    end
    foreach(close, values(io)); close(iod)
    open(joinpath(out, "ycorr_final.f64"), "w") do fh; write(fh, ycorr); end
    open(joinpath(out, "header.txt"), "w") do fh
        println(fh, "format ngp-replay-log 1")
        println(fh, "n $nData\np $p\niters $iters\nnvar $nvar\nmethod $(Ms.method)\nest_pi $(Ms.estPi ? 1 : 0)")
        println(fh, "df $(Ms.df)\nscale $(Ms.scale)\nv $v\npi $piIn\ndf_e $(E.df)\nscale_e $(E.scale)\nseed $seed")
        println(fh, "n_regions $(length(Ms.regionArray))")
        println(fh, "genotypes $(abspath(geno))\nphenotypes $(abspath(phen))")
    end
    println("recorded $iters iterations of $(Ms.method) on $nData x $p into $out")
end

main(ARGS)
