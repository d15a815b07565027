// ngp_small_la.cuh — k x k dense algebra (k <= 8, row-major, one thread) of the tuple sampler: Cholesky, SPD inverse through the
// Cholesky factor, inverse-Wishart draw by Bartlett's decomposition.  The definitions follow the CPU oracle (oracle/ngp_oracle.c:
// chol_lower, inv_spd, inv_wishart_bartlett) so that replayed variates reproduce its chain.
#pragma once
#include <math.h>

namespace ngp {

constexpr int kMaxK = 8;

// ---- small dense algebra, k <= 8, row-major, one thread
template <int k>
__device__ __forceinline__ bool jt_chol(const double* A, double* L)
{
    for (int i = 0; i < k * k; ++i) L[i] = 0.0;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = A[i * k + j];
            for (int t = 0; t < j; ++t) s -= L[i * k + t] * L[j * k + t];
            if (i == j) { if (!(s > 0.0)) return false; L[i * k + i] = sqrt(s); }
            else L[i * k + j] = s / L[j * k + j];
        }
    return true;
}

// inverse of an SPD matrix through its Cholesky factor: A^-1 = L^-T L^-1
template <int k>
__device__ __forceinline__ bool jt_inv_spd(const double* A, double* Ainv)
{
    double L[k * k], Li[k * k];
    if (!jt_chol<k>(A, L)) return false;
    for (int i = 0; i < k * k; ++i) Li[i] = 0.0;
    for (int c = 0; c < k; ++c) {
        Li[c * k + c] = 1.0 / L[c * k + c];
        for (int i = c + 1; i < k; ++i) {
            double s = 0.0;
            for (int t = c; t < i; ++t) s -= L[i * k + t] * Li[t * k + c];
            Li[i * k + c] = s / L[i * k + i];
        }
    }
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            double s = 0.0;
            for (int t = (i > j ? i : j); t < k; ++t) s += Li[t * k + i] * Li[t * k + j];
            Ainv[i * k + j] = s;
        }
    return true;
}

// Sigma ~ InvWishart(df, Psi) by Bartlett: W = (L A)(L A)' with L = chol(Psi^-1), A lower, A_ii = sqrt(chi2_i), A_ij = z_ij
template <int k>
__device__ __forceinline__ bool jt_inv_wishart(const double* Psi, const double* chi2, const double* zl, double* Sigma)
{
    double Pinv[k * k], L[k * k], LA[k * k], W[k * k];
    if (!jt_inv_spd<k>(Psi, Pinv) || !jt_chol<k>(Pinv, L)) return false;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            double s = 0.0;
            for (int t = j; t <= i; ++t) s += L[i * k + t] * ((t == j) ? sqrt(chi2[j]) : zl[t * k + j]);     // A is lower: A[t][j], t >= j
            LA[i * k + j] = s;
        }
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            double s = 0.0;
            for (int t = 0; t < k; ++t) s += LA[i * k + t] * LA[j * k + t];
            W[i * k + j] = s;
        }
    return jt_inv_spd<k>(W, Sigma);
}

}  // namespace ngp
