"""Second, independent restatement of the hot path in plain numpy (small cases only).

TEST INFRASTRUCTURE ONLY (see oracle/ngp_oracle.c).  PARITY UNPINNED.
Written directly from the reference text, statement by statement, so that the C
oracle can be cross-checked against something that shares no code with it.
All draws are explicit inputs (the replay-log format of oracle.OracleChain).

  sample_varE        /root/reference/src/functions.jl:523-525
  sample_intercept   functions.jl:39-47
  bayes_pr           functions.jl:118-137
  bayes_b            functions.jl:157-195
  bayes_c            functions.jl:197-236
  iteration          samplers.jl:32-53
  bayes_lv           functions.jl:421-486 (sampleBayesLV!), set-up mme.jl:418-440
  grn_sample_lambda2 GRN.jl:150-164 (sampleΛ2!), sampleBeta GRN.jl:234-236
"""
import numpy as np


def sample_varE(df_e, S_e, ycorr, n, chi2):
    return (df_e * S_e + ycorr @ ycorr) / chi2


def sample_intercept(ycorr, b, varE, z, lhs0=0.0, rhs0=0.0):
    iVarE = 1.0 / varE
    ones = np.ones_like(ycorr)
    ycorr += ones * b
    rhs = (ones @ ycorr) * iVarE + rhs0
    lhs = (ones @ ones) * iVarE + lhs0
    meanMu = rhs / lhs
    b = meanMu + np.sqrt(1.0 / lhs) * z
    ycorr -= ones * b
    return b


def sample_beta(meanBeta, lhs, z):
    with np.errstate(divide="ignore"):
        return meanBeta + np.sqrt(1.0 / lhs) * z


def sample_var_beta_pr(scalem, dfm, which, chi2):
    return (scalem * dfm + which @ which) / chi2


def bayes_pr(X, mpm, lhs0, rhs0, regions, scale, df, beta, ycorr, varE, varBeta, z, chi2_b):
    iVarE = 1.0 / varE
    for r in range(len(regions) - 1):
        loci = range(regions[r], regions[r + 1])
        iVarBeta = 1.0 / varBeta[r]
        for j in loci:
            ycorr += beta[j] * X[:, j]
            rhs = (X[:, j] @ ycorr) * iVarE + rhs0[j]
            lhs = mpm[j] * iVarE + lhs0[j] + iVarBeta
            beta[j] = sample_beta(rhs / lhs, lhs, z[j])
            ycorr += -1.0 * beta[j] * X[:, j]
        varBeta[r] = sample_var_beta_pr(scale, df, beta[regions[r]:regions[r + 1]], chi2_b[r])


def _prob_delta1(mpm_j, rrr, varE, varBeta_j, logPi):
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        v0 = mpm_j * varE
        v1 = (mpm_j ** 2) * varBeta_j + v0
        logDelta0 = -0.5 * (np.log(v0) + (rrr ** 2) / v0) + logPi[0]
        logDelta1 = -0.5 * (np.log(v1) + (rrr ** 2) / v1) + logPi[1]
        return 1.0 / (1.0 + np.exp(logDelta0 - logDelta1))


def bayes_b(X, mpm, lhs0, rhs0, scale, df, est_pi, beta, delta, ycorr, varE, varBeta, piHat, logPi, u, z, chi2_b, beta_pi):
    p = X.shape[1]
    nLoci = 0
    for j in range(p):
        iVarE = 1.0 / varE
        with np.errstate(divide="ignore"):
            iVarBeta = np.float64(1.0) / np.float64(varBeta[j])
        ycorr += beta[j] * X[:, j]
        rrr = X[:, j] @ ycorr
        if u[j] < _prob_delta1(mpm[j], rrr, varE, varBeta[j], logPi):
            delta[j] = 1
            nLoci += 1
            rhs = (X[:, j] @ ycorr) * iVarE + rhs0[j]
            lhs = mpm[j] * iVarE + lhs0[j] + iVarBeta
            beta[j] = sample_beta(rhs / lhs, lhs, z[j])
            ycorr += -1.0 * beta[j] * X[:, j]
            varBeta[j] = sample_var_beta_pr(scale, df, beta[j:j + 1], chi2_b[j])
        else:
            beta[j] = 0.0
            delta[j] = 0
            varBeta[j] = 0.0
    if est_pi:
        piHat[:] = [1.0 - beta_pi, beta_pi]
        logPi[:] = np.log(piHat)
    return nLoci


def bayes_c(X, mpm, lhs0, scale, df, est_pi, beta, delta, ycorr, varE, varBeta, piHat, logPi, u, z, chi2_b, beta_pi):
    p = X.shape[1]
    nLoci = 0
    iVarE = 1.0 / varE
    iVarBeta = 1.0 / varBeta[0]
    for j in range(p):
        ycorr += beta[j] * X[:, j]
        rrr = X[:, j] @ ycorr
        if u[j] < _prob_delta1(mpm[j], rrr, varE, varBeta[0], logPi):
            delta[j] = 1
            nLoci += 1
            rhs = (X[:, j] @ ycorr) * iVarE            # rhs0 is commented out in the reference
            lhs = mpm[j] * iVarE + lhs0[j] + iVarBeta
            beta[j] = sample_beta(rhs / lhs, lhs, z[j])
            ycorr += -1.0 * beta[j] * X[:, j]
        else:
            beta[j] = 0.0
            delta[j] = 0
    varBeta[0] = sample_var_beta_pr(scale, df, beta, chi2_b[0])
    if est_pi:
        piHat[:] = [1.0 - beta_pi, beta_pi]
        logPi[:] = np.log(piHat)
    return nLoci


# --------------------------------------------------------------------------- weighted residuals, E.str == "D"
#   iVarStr = inv.(D)                          mme.jl:73
#   sampleVarE(E, ycorr, n)                    functions.jl:526-528 (samplers.jl:32-33)
#   xpx = X'(iVarStr .* X), Xp = (X .* iVarStr)'   mme.jl:133-136
#   mpm_j = sum(c .* iVarStr .* c), Mp_j = (c .* iVarStr)'   mme.jl:299-303
# The sweeps below spell out which dot takes Mp (weighted) and which takes view(data,:,locus) (unweighted).
def sample_varE_w(df_e, S_e, iVarStr, ycorr, n, chi2):
    return (df_e * S_e + np.sum(iVarStr * (ycorr ** 2))) / chi2


def sample_intercept_w(ycorr, iVarStr, b, varE, z, lhs0=0.0, rhs0=0.0):
    iVarE = 1.0 / varE
    ones = np.ones_like(ycorr)
    xpx = ones @ (iVarStr * ones)
    Xp = ones * iVarStr
    ycorr += ones * b
    rhs = (Xp @ ycorr) * iVarE + rhs0
    lhs = xpx * iVarE + lhs0
    b = rhs / lhs + np.sqrt(1.0 / lhs) * z
    ycorr -= ones * b
    return b


def weighted_setup(X, iVarStr):
    mpm = np.array([np.sum(X[:, j] * iVarStr * X[:, j]) for j in range(X.shape[1])])
    Mp = [X[:, j] * iVarStr for j in range(X.shape[1])]
    return mpm, Mp


def bayes_pr_w(X, Mp, mpm, lhs0, rhs0, regions, scale, df, beta, ycorr, varE, varBeta, z, chi2_b):
    iVarE = 1.0 / varE
    for r in range(len(regions) - 1):
        iVarBeta = 1.0 / varBeta[r]
        for j in range(regions[r], regions[r + 1]):
            ycorr += beta[j] * X[:, j]                       # :128 data
            rhs = (Mp[j] @ ycorr) * iVarE + rhs0[j]          # :129 Mp
            lhs = mpm[j] * iVarE + lhs0[j] + iVarBeta
            beta[j] = sample_beta(rhs / lhs, lhs, z[j])
            ycorr += -1.0 * beta[j] * X[:, j]                # :133 data
        varBeta[r] = sample_var_beta_pr(scale, df, beta[regions[r]:regions[r + 1]], chi2_b[r])


def bayes_bc_w(method_b, X, Mp, mpm, lhs0, rhs0, scale, df, est_pi, beta, delta, ycorr, varE, varBeta, piHat, logPi, u, z, chi2_b, beta_pi):
    p = X.shape[1]
    iVarE = 1.0 / varE
    for j in range(p):
        vb = varBeta[j] if method_b else varBeta[0]
        with np.errstate(divide="ignore"):
            iVarBeta = np.float64(1.0) / np.float64(vb)
        ycorr += beta[j] * X[:, j]                           # :167 / :207 data
        rrr = X[:, j] @ ycorr                                # :168 / :208 data (unweighted), with the weighted mpm below
        if u[j] < _prob_delta1(mpm[j], rrr, varE, vb, logPi):
            delta[j] = 1
            rhs = (Mp[j] @ ycorr) * iVarE + (rhs0[j] if method_b else 0.0)     # :177 / :219 Mp
            lhs = mpm[j] * iVarE + lhs0[j] + iVarBeta
            beta[j] = sample_beta(rhs / lhs, lhs, z[j])
            ycorr += -1.0 * beta[j] * X[:, j]
            if method_b:
                varBeta[j] = sample_var_beta_pr(scale, df, beta[j:j + 1], chi2_b[j])
        else:
            beta[j] = 0.0
            delta[j] = 0
            if method_b:
                varBeta[j] = 0.0
    if not method_b:
        varBeta[0] = sample_var_beta_pr(scale, df, beta, chi2_b[0])
    if est_pi:
        piHat[:] = [1.0 - beta_pi, beta_pi]
        logPi[:] = np.log(piHat)


def bayes_rc(X, mpm, e, varE, beta, delta, annot_cat, varBeta, piHat, logPi, annot, annot_prob, v_class, df, scale, est_pi, plus, v):
    """Literal numpy restatement of sampleBayesRCpi! (functions.jl:291-360, plus=False) / sampleBayesRCplus! (functions.jl:362-419,
    plus=True) with explicit variates v = {u_annot, dirp, u, z, chi2_b, dir_pi} (layouts of oracle/ngp_oracle.c: ngo_rc_sweep).
    Mutates e, beta, delta, annot_cat, varBeta, piHat, logPi, annot_prob in place."""
    p = X.shape[1]
    nA, nc = annot.shape[1], len(v_class)
    nLoci = np.zeros((nA, nc), dtype=np.int64)
    nNonZero = np.zeros(nA, dtype=np.int64)
    varc = [vb * np.asarray(v_class) for vb in varBeta]
    sumS = np.zeros(nA)
    iVarE = 1.0 / varE
    for j in range(p):
        x = X[:, j]
        e += beta[j] * x
        nz = [a for a in range(nA) if annot[j, a] != 0]
        if not plus:
            rhs = (x @ e) * iVarE
            lhs = np.zeros((nA, nc)); Ex = np.zeros((nA, nc))
            for a in nz:
                for c in range(nc):
                    lhs[a, c] = 0.0 if varc[a][c] == 0.0 else mpm[j] * iVarE + 1.0 / varc[a][c]
                    ll = logPi[a, c] if varc[a][c] == 0.0 else -0.5 * (np.log(varc[a][c] * lhs[a, c]) - rhs ** 2 / lhs[a, c]) + logPi[a, c]
                    Ex[a, c] = np.exp(ll)
            pa1 = annot_prob[j, :] * Ex.sum(1)
            pa = pa1 / pa1.sum()
            A, cp = 0, pa[0]                                  # rand(Categorical(pa)) with the uniform u_annot[j]
            while cp <= v["u_annot"][j] and A < nA - 1:
                A += 1
                cp += pa[A]
            assert A in nz
            annot_prob[j, nz] = v["dirp"][j, nz]              # sampleProb: the Dirichlet draw itself is the variate
            pv = Ex[A, :] / Ex[A, :].sum()
            cum = np.cumsum(pv)
            cls = next(c for c in range(nc) if cum[c] >= v["u"][j, c])
            delta[j] = cls + 1
            annot_cat[j] = A + 1
            nLoci[A, cls] += 1
            if varc[A][cls] != 0.0:
                nNonZero[A] += 1
                b = rhs / lhs[A, cls] + np.sqrt(1.0 / lhs[A, cls]) * v["z"][j]
                beta[j] = b
                e -= b * x
                sumS[A] += b * b / v_class[cls]
            else:
                beta[j] = 0.0
        else:
            temp = 0.0
            for a in nz:
                rhs = (x @ e) * iVarE
                lhs = np.zeros(nc); Ex = np.zeros(nc)
                for c in range(nc):
                    lhs[c] = 0.0 if varc[a][c] == 0.0 else mpm[j] * iVarE + 1.0 / varc[a][c]
                    ll = logPi[a, c] if varc[a][c] == 0.0 else -0.5 * (np.log(varc[a][c] * lhs[c]) - rhs ** 2 / lhs[c]) + logPi[a, c]
                    Ex[c] = np.exp(ll)
                cum = np.cumsum(Ex / Ex.sum())
                cls = next(c for c in range(nc) if cum[c] >= v["u"][j, a, c])
                delta[j] = cls + 1
                nLoci[a, cls] += 1
                b = 0.0
                if varc[a][cls] != 0.0:
                    nNonZero[a] += 1
                    b = rhs / lhs[cls] + np.sqrt(1.0 / lhs[cls]) * v["z"][j, a]
                    sumS[a] += b * b / v_class[cls]
                temp += b
                e -= b * x
            beta[j] = temp
    for a in range(nA):
        varBeta[a] = (scale * df + sumS[a]) / v["chi2_b"][a]
    if est_pi:
        for a in range(nA):
            piHat[a, :] = v["dir_pi"][a, :]
            logPi[a, :] = np.log(v["dir_pi"][a, :])
    return nLoci, nNonZero


def grn_sample_lambda2(Lambda2, Xc, yCorr, var_tau, varE, pMeans, z):
    """sampleΛ2!(Λ2,Xc,yCorr,σ2τ,σ2ϵ,pMeans), GRN.jl:150-164, statement by statement.  Xc (SNPs, individuals) row-centred (GRN.jl:23);
    yCorr (genes, individuals); z (genes, SNPs) the standard normals behind rand(Normal(meanBeta, sqrt(lhs\σ2E))) (GRN.jl:234-236)."""
    nGenes, nSNPs = yCorr.shape[0], Xc.shape[0]
    for g in range(nGenes):
        alpha = varE / var_tau[g]
        for q in range(nSNPs):
            yCorr[g, :] += Lambda2[g, q] * Xc[q, :]
            RHS = Xc[q, :] @ yCorr[g, :] + alpha * pMeans[g]
            LHS = Xc[q, :] @ Xc[q, :]
            meanBeta = RHS / LHS
            nowBeta = meanBeta + np.sqrt(varE / LHS) * z[g, q]
            Lambda2[g, q] = nowBeta
            yCorr[g, :] -= nowBeta * Xc[q, :]


class _JuliaMath:
    @staticmethod
    def log(x):
        with np.errstate(all="ignore"):
            return float(np.log(np.float64(x)))

    @staticmethod
    def exp(x):
        with np.errstate(all="ignore"):
            return float(np.exp(np.float64(x)))

    @staticmethod
    def sqrt(x):
        with np.errstate(all="ignore"):
            return float(np.sqrt(np.float64(x)))


def bayes_lv(X, mpm, lhs0, rhs0, beta, ycorr, varE, varBeta, M, z, u, zc):
    """sampleBayesLV!, functions.jl:421-486, statement by statement.  M: dict with logVar, SNPVARRESID, covariates, iCpC, c, varZeta (1-list),
    estVarZeta (the fields of mme.jl:418-440).  z (p) normals of sampleBeta, u (p, 4) the rand() calls per locus in the order of the text
    (the 4th only when the slice is not trapped), zc (k) normals of the MvNormal draw.  Returns the trapped count."""
    math = _JuliaMath          # log(0) = -Inf, exp overflow = Inf like Julia (no exceptions)
    var_var = M["varZeta"][0]
    iVarE = 1.0 / varE
    p = len(beta)
    for locus in range(p):                                            # regionArray = [r:r for r in 1:p]
        ycorr += beta[locus] * X[:, locus]
        rhs = (X[:, locus] @ ycorr) * iVarE + rhs0[locus]
        lhs = mpm[locus] * iVarE + lhs0[locus] + 1.0 / varBeta[locus]
        beta[locus] = rhs / lhs + math.sqrt(1.0 / lhs) * z[locus]
        ycorr += -1.0 * beta[locus] * X[:, locus]
    trapped = 0
    for locus in range(p):
        vari = float(varBeta[locus])
        bi = float(beta[locus])
        log_vari = float(M["logVar"][locus])
        zeta = float(M["SNPVARRESID"][locus])
        var_mui = log_vari - zeta
        c1 = float(np.float64(vari) ** -1.5) * u[locus, 0]
        c2 = math.exp(-0.5 * bi * bi / vari) * u[locus, 1]
        c3 = math.exp(-0.5 * zeta * zeta / var_var) * u[locus, 2]
        temp = math.sqrt(-2.0 * var_var * math.log(c3))
        lbound = math.exp(var_mui - temp)
        rbound = math.exp(var_mui + temp)
        if math.exp((-2.0 / 3.0) * math.log(c1)) < rbound:
            rbound = math.exp((-2.0 / 3.0) * math.log(c1))
        with np.errstate(all="ignore"):
            l1 = float(np.float64(-0.5 * bi * bi) / np.float64(math.log(c2)))
        if l1 > lbound:
            lbound = l1
        if lbound >= rbound:
            trapped += 1
        else:
            vari = lbound + u[locus, 3] * (rbound - lbound)
            varBeta[locus] = vari
            M["logVar"][locus] = math.log(vari)
    rhsC = M["covariates"].T @ M["logVar"]
    meanC = M["iCpC"] @ rhsC
    cov = M["iCpC"] * var_var
    cov = np.triu(cov) + np.triu(cov, 1).T                            # Symmetric(): the upper triangle
    M["c"][:] = meanC + np.linalg.cholesky(cov) @ zc                  # rand(MvNormal(mean, cov)) = mean + chol(cov).L z
    M["SNPVARRESID"][:] = M["logVar"] - M["covariates"] @ M["c"]
    est = M["estVarZeta"]
    if isinstance(est, float):
        M["varZeta"][0] = est * np.var(M["logVar"], ddof=1)
    elif est is False:
        pass
    elif est is True:
        M["varZeta"][0] = np.var(M["SNPVARRESID"], ddof=1)
    return trapped
