// oracle/cpu_ref/cpu_ref.cpp -- compiled CPU competitor for the fused g + jac_g evaluation (TEST INFRASTRUCTURE).
//
// This is the "honest CPU baseline" of SURVEY.md 8(d) / BASELINE.md 3-2b: what a careful C++17 -O3 -march=native
// OpenMP implementation of the reference's transcription (mpopt/mpopt.py:154-462, evaluated per iteration by CasADi's
// single-threaded SX virtual machine at mpopt.py:804) costs on the host cores of the GPU box.  It is NOT part of the
// product: only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.build() (which compiles
// it) touch it; mpopt_b200/ never loads it.  It is checked against the numpy oracle (oracle/nlp.py) to 1e-13 in
// tests/test_cpu_ref.py, which in turn is pinned to the reference's known answers.
//
// Scope: single-phase OCPs without path constraints, parameters or explicit time dependence, unit scales -- the
// three problems BASELINE.json's configs 1-4 are built on (moon-lander tests/test_mpopt.py:113-144, van-der-Pol
// :205-227, the seeded synthetic 6/3 quadratic dynamics of SURVEY.md 8d); uniform or mixed degrees; rows
// [F | mU | TC] (mpopt.py:458), CSR with sorted columns, state-major variables (mpopt.py:537-543).
//   F(s,i)  = sum_j D_k[loc(i), j] X(s_k + j, s) - h_k f_s(x_i, u_i)             mpopt.py:201, :232
//   mU(c,m) = sum_j Cmid_k[m_loc, j] U(s_k + j, c)                                mpopt.py:357-360
//   TC(r)   = xf[tc_state(r)]                                                     mpopt.py:277-292
// h_k = (tf - t0) / (tau1 - tau0) * w_k (mpopt.py:184); node ownership: a shared node belongs to the earlier
// segment (mpopt.py:189-195).  Jacobian values follow SURVEY.md Appendix A.
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Synthetic63 {  // f_s = A_s.x + B_s.u + x_s (C_s.x)
  static constexpr int NX = 6, NU = 3, NTC = 0;
  const double *A, *B, *C;
  static bool pat(int, int) { return true; }
  static int tc_state(int) { return 0; }
  inline void dyn(const double* x, const double* u, double* f, double* J) const {
    for (int s = 0; s < NX; ++s) {
      double ax = 0, cx = 0, bu = 0;
      for (int j = 0; j < NX; ++j) ax += A[s * NX + j] * x[j], cx += C[s * NX + j] * x[j];
      for (int c = 0; c < NU; ++c) bu += B[s * NU + c] * u[c];
      f[s] = ax + bu + x[s] * cx;
      for (int j = 0; j < NX; ++j) J[s * (NX + NU) + j] = A[s * NX + j] + x[s] * C[s * NX + j] + (j == s ? cx : 0.0);
      for (int c = 0; c < NU; ++c) J[s * (NX + NU) + NX + c] = B[s * NU + c];
    }
  }
};
struct MoonLander {  // f = [x1, u0 - 1.5], TC = [xf0, xf1]
  static constexpr int NX = 2, NU = 1, NTC = 2;
  static bool pat(int s, int v) { return s == 0 ? v == 1 : v == 2; }
  static int tc_state(int r) { return r; }
  inline void dyn(const double* x, const double* u, double* f, double* J) const {
    f[0] = x[1], f[1] = u[0] - 1.5;
    J[0] = 0, J[1] = 1, J[2] = 0, J[3] = 0, J[4] = 0, J[5] = 1;
  }
};
struct VanDerPol {  // f = [(1 - x1^2) x0 - x1 + u0, x0]
  static constexpr int NX = 2, NU = 1, NTC = 0;
  static bool pat(int s, int v) { return s == 0 ? true : v == 0; }
  static int tc_state(int) { return 0; }
  inline void dyn(const double* x, const double* u, double* f, double* J) const {
    f[0] = (1 - x[1] * x[1]) * x[0] - x[1] + u[0], f[1] = x[0];
    J[0] = 1 - x[1] * x[1], J[1] = -2 * x[1] * x[0] - 1, J[2] = 1, J[3] = 1, J[4] = 0, J[5] = 0;
  }
};

struct Base {
  virtual ~Base() {}
  virtual void structure(int64_t* rowptr, int64_t* colind) const = 0;
  virtual void eval(const double* z, const double* w, double* g, double* vals) const = 0;
  int64_t n_z = 0, n_g = 0, nnz = 0;
};

template <class PH>
struct Plan final : Base {
  static constexpr int NX = PH::NX, NU = PH::NU, NV = NX + NU;
  PH ph;
  int K, N;
  bool midu;
  std::vector<int> po, s0;                 // degree and first node per segment
  std::vector<const double*> Dk, Ck;       // table of the segment's degree
  std::vector<std::vector<double>> Dt, Ct; // storage per unique degree
  std::vector<int64_t> dpre, ipre;         // D / mid-point non-zeros before the segment (per state / control)
  int next[NX], npre[NX];                  // entries of an F row outside / before the D block
  int64_t vF[NX], vmU, vTC, gmU, gTC, nnzD, nnzI;

  Plan(const PH& f, int K_, const int* po_, int n_deg, const int* degs, const double* const* D, const double* const* C, bool midu_)
      : ph(f), K(K_), midu(midu_), po(po_, po_ + K_) {
    Dt.resize(n_deg), Ct.resize(n_deg);
    for (int i = 0; i < n_deg; ++i) {
      const int n1 = degs[i] + 1;
      Dt[i].assign(D[i], D[i] + n1 * n1), Ct[i].assign(C[i], C[i] + (n1 - 1) * n1);
    }
    s0.resize(K + 1), dpre.resize(K), ipre.resize(K), Dk.resize(K), Ck.resize(K);
    int64_t dp = 0, ip = 0;
    s0[0] = 0;
    for (int k = 0; k < K; ++k) {
      const int d = po[k];
      s0[k + 1] = s0[k] + d;
      dpre[k] = dp, ipre[k] = ip;
      dp += (int64_t)(k == 0 ? d + 1 : d) * (d + 1), ip += (int64_t)d * (d + 1);
      for (int i = 0; i < n_deg; ++i)
        if (degs[i] == d) Dk[k] = Dt[i].data(), Ck[k] = Ct[i].data();
    }
    N = s0[K] + 1, nnzD = dp, nnzI = ip;
    n_z = (int64_t)NV * N + 2;
    int64_t v = 0;
    for (int s = 0; s < NX; ++s) {
      next[s] = 2, npre[s] = 0;  // T0, TF: f_s is not identically zero in any of these problems
      for (int j = 0; j < NV; ++j)
        if (j != s && PH::pat(s, j)) ++next[s], npre[s] += j < s;
      vF[s] = v, v += nnzD + (int64_t)N * next[s];
    }
    gmU = (int64_t)NX * N;
    vmU = v;
    if (midu) v += (int64_t)NU * nnzI;
    gTC = gmU + (midu ? (int64_t)NU * (N - 1) : 0);
    vTC = v, v += PH::NTC;
    n_g = gTC + PH::NTC, nnz = v;
  }

  void structure(int64_t* rp, int64_t* ci) const override {
    int64_t e = 0, r = 0;
    for (int s = 0; s < NX; ++s)
      for (int k = 0; k < K; ++k)
        for (int j = (k == 0 ? 0 : 1); j <= po[k]; ++j) {
          const int i = s0[k] + j;
          rp[r++] = e;
          for (int v = 0; v < s; ++v)
            if (PH::pat(s, v)) ci[e++] = (int64_t)v * N + i;
          for (int c = 0; c <= po[k]; ++c) ci[e++] = (int64_t)s * N + s0[k] + c;
          for (int v = s + 1; v < NV; ++v)
            if (PH::pat(s, v)) ci[e++] = (int64_t)v * N + i;
          ci[e++] = (int64_t)NV * N, ci[e++] = (int64_t)NV * N + 1;
        }
    if (midu)
      for (int c = 0; c < NU; ++c)
        for (int k = 0; k < K; ++k)
          for (int m = 0; m < po[k]; ++m) {
            rp[r++] = e;
            for (int j = 0; j <= po[k]; ++j) ci[e++] = (int64_t)(NX + c) * N + s0[k] + j;
          }
    for (int t = 0; t < PH::NTC; ++t) rp[r++] = e, ci[e++] = (int64_t)PH::tc_state(t) * N + N - 1;
    rp[r] = e;
  }

  void eval(const double* z, const double* w, double* g, double* vals) const override {
    const double t0 = z[(int64_t)NV * N], tf = z[(int64_t)NV * N + 1];
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; ++k) {
      const int d = po[k], n1 = d + 1, rb = k == 0 ? 0 : 1, sk = s0[k];
      const double* D = Dk[k];
      const double h = (tf - t0) * 0.5 * w[k], gk = 0.5 * w[k];  // tau1 - tau0 = 2, scale_t = 1
      const int64_t rowpre = k == 0 ? 0 : sk + 1;
      for (int j = rb; j <= d; ++j) {
        const int i = sk + j;
        double x[NX], u[NU > 0 ? NU : 1], f[NX], J[NX * NV];
        for (int s = 0; s < NX; ++s) x[s] = z[(int64_t)s * N + i];
        for (int c = 0; c < NU; ++c) u[c] = z[(int64_t)(NX + c) * N + i];
        ph.dyn(x, u, f, J);
        for (int s = 0; s < NX; ++s) {
          const int L = n1 + next[s];
          double* row = vals + vF[s] + dpre[k] + rowpre * next[s] + (int64_t)(j - rb) * L;
          const double* xs = z + (int64_t)s * N + sk;
          const double* Dr = D + j * n1;
          int e = 0;
          for (int v = 0; v < s; ++v)
            if (PH::pat(s, v)) row[e++] = -h * J[s * NV + v];
          double acc = 0;
#pragma omp simd reduction(+ : acc)
          for (int c = 0; c < n1; ++c) {
            row[e + c] = Dr[c];
            acc += Dr[c] * xs[c];
          }
          if (PH::pat(s, s)) row[e + j] = Dr[j] - h * J[s * NV + s];
          e += n1;
          for (int v = s + 1; v < NV; ++v)
            if (PH::pat(s, v)) row[e++] = -h * J[s * NV + v];
          row[e++] = gk * f[s], row[e++] = -gk * f[s];
          g[(int64_t)s * N + i] = acc - h * f[s];
        }
      }
      if (midu) {
        const double* Cm = Ck[k];
        for (int c = 0; c < NU; ++c) {
          const double* us = z + (int64_t)(NX + c) * N + sk;
          double* blk = vals + vmU + (int64_t)c * nnzI + ipre[k];
          std::memcpy(blk, Cm, sizeof(double) * d * n1);
          for (int m = 0; m < d; ++m) {
            double acc = 0;
#pragma omp simd reduction(+ : acc)
            for (int jj = 0; jj < n1; ++jj) acc += Cm[m * n1 + jj] * us[jj];
            g[gmU + (int64_t)c * (N - 1) + sk + m] = acc;
          }
        }
      }
    }
    for (int t = 0; t < PH::NTC; ++t) {
      g[gTC + t] = z[(int64_t)PH::tc_state(t) * N + N - 1];
      vals[vTC + t] = 1.0;
    }
  }
};
}  // namespace

extern "C" {
// problem: 0 synthetic 6/3 (params = A[36] | B[18] | C[36], kept by the caller), 1 moon-lander, 2 van-der-Pol
void* cpu_ref_create(int problem, const double* params, int K, const int* po, int n_deg, const int* degs,
                     const double* const* D, const double* const* C, int midu) {
  if (problem == 0) return new Plan<Synthetic63>(Synthetic63{params, params + 36, params + 54}, K, po, n_deg, degs, D, C, midu);
  if (problem == 1) return new Plan<MoonLander>(MoonLander{}, K, po, n_deg, degs, D, C, midu);
  if (problem == 2) return new Plan<VanDerPol>(VanDerPol{}, K, po, n_deg, degs, D, C, midu);
  return nullptr;
}
void cpu_ref_destroy(void* h) { delete static_cast<Base*>(h); }
void cpu_ref_sizes(void* h, int64_t* n_z, int64_t* n_g, int64_t* nnz) {
  const Base* b = static_cast<Base*>(h);
  *n_z = b->n_z, *n_g = b->n_g, *nnz = b->nnz;
}
void cpu_ref_structure(void* h, int64_t* rowptr, int64_t* colind) { static_cast<Base*>(h)->structure(rowptr, colind); }
void cpu_ref_eval(void* h, const double* z, const double* w, double* g, double* vals) { static_cast<Base*>(h)->eval(z, w, g, vals); }
int cpu_ref_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void cpu_ref_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
}
