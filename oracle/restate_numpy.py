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
