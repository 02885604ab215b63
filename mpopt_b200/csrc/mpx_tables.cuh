// mpx_tables.cuh -- K0/K1: collocation nodes and D / w / C tables on the device.
//
// Replaces CollocationRoots (/root/reference/mpopt/mpopt.py:4134-4276: scipy.special.j_roots
// = Golub-Welsch + Newton polish, closed-form CGL) and Collocation.get_diff_matrix /
// get_quadrature_weights / get_interpolation_matrix (:3815-3905: CasADi symbolic
// differentiation of the Lagrange product form, IDAS integration).  One CTA per degree:
//   nodes   : Sturm-sequence bisection on the Jacobi matrix of P^(alpha,beta)_{d-1}, one thread
//             per root, then Newton on the monic three-term recurrence;
//   D       : product-form derivative at the nodes (closed form of what ca.gradient yields);
//   w       : exact Gauss-Legendre quadrature of l_j (the reference integrates with IDAS at
//             default tolerances -- SURVEY.md quirk Q2);
//   Cmid    : l_j at the mid-points between nodes (mpopt.py:350-359).
#pragma once
#include "mpx_kernels.cuh"

#define MPX_TAB_THREADS 128
#define MPX_MAX_DEG 200

// Jacobi-matrix recurrence coefficients of P^(al,be): diagonal a_k, squared off-diagonal b2_k (k>=1)
__host__ __device__ __forceinline__ double mpx_jac_a(int k, double al, double be) {
  if (k == 0) return (be - al) / (al + be + 2.0);
  const double s = 2.0 * k + al + be;
  return (be * be - al * al) / (s * (s + 2.0));
}
__host__ __device__ __forceinline__ double mpx_jac_b2(int k, double al, double be) {
  const double s = 2.0 * k + al + be;
  return 4.0 * k * (k + al) * (k + be) * (k + al + be) / (s * s * (s - 1.0) * (s + 1.0));
}
// number of eigenvalues of the n x n Jacobi matrix below x (Sturm count via LDL^T pivots)
__host__ __device__ __forceinline__ int mpx_sturm(int n, double al, double be, double x) {
  int cnt = 0;
  double q = mpx_jac_a(0, al, be) - x;
  if (q < 0.0) ++cnt;
  for (int k = 1; k < n; ++k) {
    if (q == 0.0) q = 1e-300;
    q = mpx_jac_a(k, al, be) - x - mpx_jac_b2(k, al, be) / q;
    if (q < 0.0) ++cnt;
  }
  return cnt;
}
// j-th (ascending) root of P^(al,be)_n
__host__ __device__ double mpx_jacobi_root(int n, int j, double al, double be) {
  double lo = -1.0, hi = 1.0;
  for (int it = 0; it < 64; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (mpx_sturm(n, al, be, mid) > j) hi = mid;
    else lo = mid;
  }
  double x = 0.5 * (lo + hi);
  for (int it = 0; it < 4; ++it) {  // Newton on the monic recurrence p_k = (x-a_{k-1}) p_{k-1} - b2_{k-1} p_{k-2}
    double pm = 0.0, p = 1.0, dpm = 0.0, dp = 0.0;
    for (int k = 1; k <= n; ++k) {
      const double a = mpx_jac_a(k - 1, al, be);
      const double b2 = k >= 2 ? mpx_jac_b2(k - 1, al, be) : 0.0;
      const double pn = (x - a) * p - b2 * pm;
      const double dpn = p + (x - a) * dp - b2 * dpm;
      pm = p, p = pn, dpm = dp, dp = dpn;
    }
    const double dx = p / dp;
    x -= dx;
    if (fabs(dx) < 1e-17) break;
  }
  return x;
}

// nodes of one degree into shared memory R[n1] (mapped to [tmin, tmax]); all threads of the CTA call it
__device__ void mpx_nodes(int scheme, int d, double tmin, double tmax, double* R) {
  const int n1 = d + 1;
  const double half = (tmax - tmin) / 2.0;
  for (int i = threadIdx.x; i < n1; i += blockDim.x) {
    double r;
    if (d == 1) {
      R[i] = i == 0 ? tmin : tmax;  // mpopt.py:4226-4227
      continue;
    }
    if (scheme == 2) r = cos(3.14159265358979323846 * (double)(d - i) / (double)d);  // mpopt.py:4271
    else if (i == 0) r = -1.0;
    else if (i == d) r = 1.0;
    else r = mpx_jacobi_root(d - 1, i - 1, 1.0, scheme == 1 ? 1.0 : 0.0);  // mpopt.py:4220, :4246
    R[i] = tmin + half * (r + 1.0);  // mpopt.py:4224
  }
  __syncthreads();
}

// Gauss-Legendre rule with nq points on [-1,1] into shared xq/wq
__device__ void mpx_gauss_legendre(int nq, double* xq, double* wq) {
  for (int i = threadIdx.x; i < nq; i += blockDim.x) {
    double x = mpx_jacobi_root(nq, i, 0.0, 0.0);
    double p0 = 1.0, p1 = x;  // standard Legendre: (k+1) P_{k+1} = (2k+1) x P_k - k P_{k-1}
    for (int k = 1; k < nq; ++k) {
      const double p2 = ((2.0 * k + 1.0) * x * p1 - k * p0) / (k + 1.0);
      p0 = p1, p1 = p2;
    }
    const double dp = nq * (x * p1 - p0) / (x * x - 1.0);
    xq[i] = x;
    wq[i] = 2.0 / ((1.0 - x * x) * dp * dp);
  }
  __syncthreads();
}

// w_j = int_{ta}^{tb} l_j, exact (integrand degree d <= 2 nq - 1)
__device__ void mpx_weights(const double* R, int n1, double ta, double tb, const double* xq, const double* wq, int nq,
                            double* w) {
  for (int j = threadIdx.x; j < n1; j += blockDim.x) {
    double acc = 0.0;
    for (int q = 0; q < nq; ++q) acc += wq[q] * mpx_lagrange(R, n1, j, ta + (tb - ta) / 2.0 * (xq[q] + 1.0));
    w[j] = (tb - ta) / 2.0 * acc;
  }
}

// full table record (layout MpxTab) for each degree in degs[]; grid = number of degrees
__global__ void __launch_bounds__(MPX_TAB_THREADS) mpx_tables_kernel(int scheme, const int* degs, const int* rec_off,
                                                                    double tmin, double tmax, double* recs) {
  extern __shared__ __align__(16) double sm[];
  const int d = degs[blockIdx.x], n1 = d + 1, nq = n1 / 2 + 1;
  double* R = sm;
  double* xq = R + n1;
  double* wq = xq + nq;
  double* rec = recs + rec_off[blockIdx.x];
  mpx_nodes(scheme, d, tmin, tmax, R);
  mpx_gauss_legendre(nq, xq, wq);
  for (int i = threadIdx.x; i < n1; i += blockDim.x) rec[MpxTab::off_roots(n1) + i] = R[i];
  mpx_weights(R, n1, tmin, tmax, xq, wq, nq, rec + MpxTab::off_w(n1));
  // D[i][j] = l_j'(R_i): i != j -> 1/(R_j - R_i) prod_{m != i,j} (R_i - R_m)/(R_j - R_m); i == j -> sum_k 1/(R_i - R_k)
  double* D = rec + MpxTab::off_D(n1);
  for (int e = threadIdx.x; e < n1 * n1; e += blockDim.x) {
    const int i = e / n1, j = e - i * n1;
    double v;
    if (i == j) {
      v = 0.0;
      for (int k = 0; k < n1; ++k)
        if (k != i) v += 1.0 / (R[i] - R[k]);
    } else {
      v = 1.0 / (R[j] - R[i]);
      for (int m = 0; m < n1; ++m)
        if (m != i && m != j) v *= (R[i] - R[m]) / (R[j] - R[m]);
    }
    D[e] = v;
    rec[MpxTab::off_Dt(n1) + j * n1 + i] = v;
  }
  double* C = rec + MpxTab::off_C(n1);
  for (int e = threadIdx.x; e < d * n1; e += blockDim.x) {
    const int m = e / n1, j = e - m * n1;
    const double v = mpx_lagrange(R, n1, j, (R[m] + R[m + 1]) / 2.0);  // mpopt.py:350-352
    C[e] = v;
    rec[MpxTab::off_Ct(n1) + j * d + m] = v;
  }
  // replicated constant blocks (one bulk store per work unit of the warp kernels): smax copies of Cmid and of D[1:, :]
  const int smax = MpxTab::smax(n1);
  if (smax > 1) {
    __syncthreads();  // D and C of this record are complete (same CTA wrote them)
    for (int e = threadIdx.x; e < smax * d * n1; e += blockDim.x) {
      const int w = e % (d * n1);
      rec[MpxTab::off_Cr(n1) + e] = C[w];
      rec[MpxTab::off_Dr(n1) + e] = D[n1 + w];
    }
  }
}

// basis (order 0) or its derivatives (order 1|2) at arbitrary points: out[n_taus][n1]
__global__ void __launch_bounds__(MPX_TAB_THREADS) mpx_basis_at_kernel(int scheme, int d, double tmin, double tmax,
                                                                      int order, int n_taus, const double* taus,
                                                                      double* out) {
  extern __shared__ __align__(16) double sm[];
  const int n1 = d + 1;
  double* R = sm;
  mpx_nodes(scheme, d, tmin, tmax, R);
  for (int e = threadIdx.x; e < n_taus * n1; e += blockDim.x) {
    const int i = e / n1, j = e - i * n1;
    out[e] = order == 0 ? mpx_lagrange(R, n1, j, taus[i]) : mpx_lagrange_der(R, n1, j, taus[i], order);
  }
}

// quadrature weights over a sub-interval [ta, tb] (mpopt.py:3851-3882 with tau0/tau1 given)
__global__ void __launch_bounds__(MPX_TAB_THREADS) mpx_weights_kernel(int scheme, int d, double tmin, double tmax,
                                                                     double ta, double tb, double* w) {
  extern __shared__ __align__(16) double sm[];
  const int n1 = d + 1, nq = n1 / 2 + 1;
  double* R = sm;
  double* xq = R + n1;
  double* wq = xq + nq;
  mpx_nodes(scheme, d, tmin, tmax, R);
  mpx_gauss_legendre(nq, xq, wq);
  mpx_weights(R, n1, ta, tb, xq, wq, nq, w);
}
