// mpx_shims.cu -- solver-facing adapters over the C ABI of include/mpx.h (host code only).
//
//   * IPOPT's C interface (IpStdCInterface.h callback shapes): what CasADi's Nlpsol/Ipopt plugin calls inside the
//     single solver call of the reference (/root/reference/mpopt/mpopt.py:804);
//   * CasADi's external-function ABI for nlp_f / nlp_g / nlp_grad_f / nlp_jac_g, the functions ca.nlpsol derives at
//     mpopt.py:757 (names as printed in the reference's stored timing tables, e.g.
//     docs/source/notebooks/moon_lander.ipynb:204-209).
//   * nlp_hess_l (x, p, lam_f, lam_g) -> triu Hessian of the Lagrangian, and IPOPT's Eval_H_CB.
// Neither IPOPT nor CasADi is present in this image, so these are exercised through their C calling conventions
// by tests/test_shims.py, not by the real solvers.
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mpx.h"

namespace {
// the parameter vector may be NULL when the NLP has none (adaptive NLP: the widths are part of x, n_p = 0)
bool p_missing(mpx_plan* plan, const double* p) {
  int64_t n_p = 0;
  return !p && (mpx_sizes(plan, nullptr, &n_p, nullptr, nullptr) != MPX_OK || n_p != 0);
}
// one fused evaluation per distinct x: stage everything on new_x, then only fetch
int ensure_staged(mpx_ipopt_data* d, const double* x, int new_x, int need) {
  if (!d || !d->plan || p_missing(d->plan, d->p)) return MPX_EINVAL;
  if ((new_x & 0xff) || (mpx_staged(d->plan) & need) != need)
    return mpx_stage(d->plan, x, d->p, MPX_STAGE_F | MPX_STAGE_GRAD | MPX_STAGE_G | MPX_STAGE_JAC);
  return MPX_OK;
}
bool sizes_ok(mpx_plan* plan, int n, int m, int nele) {
  int64_t n_z, n_p, n_g, nnz;
  if (mpx_sizes(plan, &n_z, &n_p, &n_g, &nnz) != MPX_OK) return false;
  return n == n_z && (m < 0 || m == n_g) && (nele < 0 || nele == nnz);
}
}  // namespace

extern "C" int mpx_ipopt_eval_f(int n, const double* x, int new_x, double* obj_value, void* user_data) {
  mpx_ipopt_data* d = static_cast<mpx_ipopt_data*>(user_data);
  if (!d || !obj_value || !sizes_ok(d->plan, n, -1, -1)) return 0;
  if (ensure_staged(d, x, new_x, MPX_STAGE_F) != MPX_OK) return 0;
  return mpx_fetch(d->plan, MPX_STAGE_F, obj_value) == MPX_OK;
}

extern "C" int mpx_ipopt_eval_grad_f(int n, const double* x, int new_x, double* grad_f, void* user_data) {
  mpx_ipopt_data* d = static_cast<mpx_ipopt_data*>(user_data);
  if (!d || !grad_f || !sizes_ok(d->plan, n, -1, -1)) return 0;
  if (ensure_staged(d, x, new_x, MPX_STAGE_GRAD) != MPX_OK) return 0;
  return mpx_fetch(d->plan, MPX_STAGE_GRAD, grad_f) == MPX_OK;
}

extern "C" int mpx_ipopt_eval_g(int n, const double* x, int new_x, int m, double* g, void* user_data) {
  mpx_ipopt_data* d = static_cast<mpx_ipopt_data*>(user_data);
  if (!d || !g || !sizes_ok(d->plan, n, m, -1)) return 0;
  if (ensure_staged(d, x, new_x, MPX_STAGE_G) != MPX_OK) return 0;
  return mpx_fetch(d->plan, MPX_STAGE_G, g) == MPX_OK;
}

extern "C" int mpx_ipopt_eval_jac_g(int n, const double* x, int new_x, int m, int nele_jac, int* iRow, int* jCol,
                                    double* values, void* user_data) {
  mpx_ipopt_data* d = static_cast<mpx_ipopt_data*>(user_data);
  if (!d || !sizes_ok(d->plan, n, m, nele_jac)) return 0;
  if (!values) {  // structure request: CSR rows expanded to triplets, C index style
    if (!iRow || !jCol) return 0;
    std::vector<int64_t> rp((size_t)m + 1), ci((size_t)nele_jac);
    if (mpx_jac_structure(d->plan, rp.data(), ci.data()) != MPX_OK) return 0;
    for (int r = 0; r < m; ++r)
      for (int64_t e = rp[r]; e < rp[r + 1]; ++e) iRow[e] = r, jCol[e] = (int)ci[e];
    return 1;
  }
  if (ensure_staged(d, x, new_x, MPX_STAGE_JAC) != MPX_OK) return 0;
  return mpx_fetch(d->plan, MPX_STAGE_JAC, values) == MPX_OK;
}

// Eval_H_CB: lower triangle of sigma * hess f + sum lambda_i hess g_i.  values == NULL: the pattern as triplets.
extern "C" int mpx_ipopt_eval_h(int n, const double* x, int new_x, double obj_factor, int m, const double* lambda,
                                int new_lambda, int nele_hess, int* iRow, int* jCol, double* values, void* user_data) {
  (void)new_lambda;  // the Hessian kernel is cheap next to the Jacobian: always evaluated from (x, lambda)
  mpx_ipopt_data* d = static_cast<mpx_ipopt_data*>(user_data);
  if (!d || !d->plan || p_missing(d->plan, d->p) || !sizes_ok(d->plan, n, m, -1)) return 0;
  int64_t nnz = 0;
  if (mpx_hess_structure(d->plan, &nnz, nullptr, nullptr) != MPX_OK || nnz != nele_hess) return 0;
  if (!values) {
    if (!iRow || !jCol) return 0;
    std::vector<int64_t> rp((size_t)n + 1), ci((size_t)nnz);
    if (mpx_hess_structure(d->plan, nullptr, rp.data(), ci.data()) != MPX_OK) return 0;
    for (int r = 0; r < n; ++r)
      for (int64_t e = rp[r]; e < rp[r + 1]; ++e) iRow[e] = r, jCol[e] = (int)ci[e];
    return 1;
  }
  if (!x || !lambda) return 0;
  // eval_h may be the FIRST callback at a new iterate: stage f / grad_f / g / jac_g for this x like the other callbacks
  // do, so that the callbacks that follow with new_x = 0 fetch results of the same x; then the Hessian at the staged x
  if (ensure_staged(d, x, new_x, MPX_STAGE_F) != MPX_OK) return 0;
  return mpx_hess_l_staged(d->plan, obj_factor, lambda, values) == MPX_OK;
}

// ------------------------------------------------------------------ CasADi external functions
namespace {
struct Bound {
  mpx_plan* plan = nullptr;
  int64_t n_z = 0, n_p = 0, n_g = 0, nnz = 0;
  // compact CCS patterns [nrow, ncol, colind[ncol+1], row[nnz]] (CasADi's casadi_int = long long)
  std::vector<long long> sp_x, sp_p, sp_f, sp_g, sp_jac, sp_hess;
} B;

std::vector<long long> dense_col(int64_t n) {
  std::vector<long long> s{(long long)n, 1, 0, (long long)n};
  for (int64_t i = 0; i < n; ++i) s.push_back(i);
  return s;
}
}  // namespace

extern "C" int mpx_casadi_bind(mpx_plan* plan) {
  if (!plan) {
    B = Bound();
    return MPX_OK;
  }
  Bound b;
  b.plan = plan;
  int rc = mpx_sizes(plan, &b.n_z, &b.n_p, &b.n_g, &b.nnz);
  if (rc) return rc;
  b.sp_x = dense_col(b.n_z), b.sp_p = dense_col(b.n_p), b.sp_f = dense_col(1), b.sp_g = dense_col(b.n_g);
  std::vector<int64_t> cp((size_t)b.n_z + 1), ri((size_t)b.nnz);
  rc = mpx_jac_structure_ccs(plan, cp.data(), ri.data(), nullptr);
  if (rc) return rc;
  b.sp_jac.reserve(2 + cp.size() + ri.size());
  b.sp_jac.push_back(b.n_g), b.sp_jac.push_back(b.n_z);
  for (int64_t v : cp) b.sp_jac.push_back(v);
  for (int64_t v : ri) b.sp_jac.push_back(v);
  // Hessian: the lower triangle in CSR is the upper triangle in CCS (CasADi's hess_gamma_x_x); optional
  int64_t hn = 0;
  if (mpx_hess_structure(plan, &hn, nullptr, nullptr) == MPX_OK) {
    std::vector<int64_t> rp((size_t)b.n_z + 1), ci((size_t)hn);
    if (mpx_hess_structure(plan, nullptr, rp.data(), ci.data()) == MPX_OK) {
      b.sp_hess.push_back(b.n_z), b.sp_hess.push_back(b.n_z);
      for (int64_t v : rp) b.sp_hess.push_back(v);
      for (int64_t v : ci) b.sp_hess.push_back(v);
    }
  }
  B = std::move(b);
  return MPX_OK;
}

namespace {
// kinds: 0 nlp_f, 1 nlp_g, 2 nlp_grad_f, 3 nlp_jac_g
int ca_eval(int kind, const double** arg, double** res) {
  if (!B.plan || !arg || !res || !arg[0] || p_missing(B.plan, arg[1])) return 1;
  const int what = kind == 0 ? MPX_STAGE_F : kind == 1 ? MPX_STAGE_G : kind == 2 ? (MPX_STAGE_F | MPX_STAGE_GRAD)
                                                                                 : (MPX_STAGE_G | MPX_STAGE_JAC);
  if (mpx_stage(B.plan, arg[0], arg[1], what) != MPX_OK) return 1;
  int rc = MPX_OK;
  if (kind == 0 && res[0]) rc = mpx_fetch(B.plan, MPX_STAGE_F, res[0]);
  if (kind == 1 && res[0]) rc = mpx_fetch(B.plan, MPX_STAGE_G, res[0]);
  if (kind == 2) {
    if (res[0]) rc = mpx_fetch(B.plan, MPX_STAGE_F, res[0]);
    if (!rc && res[1]) rc = mpx_fetch(B.plan, MPX_STAGE_GRAD, res[1]);
  }
  if (kind == 3) {
    if (res[0]) rc = mpx_fetch(B.plan, MPX_STAGE_G, res[0]);
    if (!rc && res[1]) rc = mpx_fetch(B.plan, MPX_FETCH_JAC_CCS, res[1]);
  }
  return rc == MPX_OK ? 0 : 1;  // non-zero: CasADi reports an evaluation failure, IPOPT backtracks
}
const long long* ca_sp_out(int kind, long long i) {
  if (!B.plan) return nullptr;
  if (kind == 0) return i == 0 ? B.sp_f.data() : nullptr;
  if (kind == 1) return i == 0 ? B.sp_g.data() : nullptr;
  if (kind == 2) return i == 0 ? B.sp_f.data() : i == 1 ? B.sp_x.data() : nullptr;
  return i == 0 ? B.sp_g.data() : i == 1 ? B.sp_jac.data() : nullptr;
}
const char* ca_name_out(int kind, long long i) {
  static const char* names[4][2] = {{"f", nullptr}, {"g", nullptr}, {"f", "grad_f_x"}, {"g", "jac_g_x"}};
  return (i == 0 || i == 1) ? names[kind][i] : nullptr;
}
}  // namespace

#define MPX_CASADI_DEFINE(NAME, KIND, NOUT)                                                                        \
  extern "C" int NAME(const double** arg, double** res, long long*, double*, int) { return ca_eval(KIND, arg, res); } \
  extern "C" long long NAME##_n_in(void) { return 2; }                                                             \
  extern "C" long long NAME##_n_out(void) { return NOUT; }                                                         \
  extern "C" double NAME##_default_in(long long) { return 0.0; }                                                   \
  extern "C" const char* NAME##_name_in(long long i) { return i == 0 ? "x" : i == 1 ? "p" : nullptr; }             \
  extern "C" const char* NAME##_name_out(long long i) { return i < NOUT ? ca_name_out(KIND, i) : nullptr; }        \
  extern "C" const long long* NAME##_sparsity_in(long long i) {                                                    \
    return !B.plan ? nullptr : i == 0 ? B.sp_x.data() : i == 1 ? B.sp_p.data() : nullptr;                          \
  }                                                                                                                \
  extern "C" const long long* NAME##_sparsity_out(long long i) { return i < NOUT ? ca_sp_out(KIND, i) : nullptr; } \
  extern "C" int NAME##_work(long long* sz_arg, long long* sz_res, long long* sz_iw, long long* sz_w) {            \
    if (sz_arg) *sz_arg = 2;                                                                                       \
    if (sz_res) *sz_res = NOUT;                                                                                    \
    if (sz_iw) *sz_iw = 0;                                                                                         \
    if (sz_w) *sz_w = 0;                                                                                           \
    return 0;                                                                                                      \
  }                                                                                                                \
  extern "C" int NAME##_alloc_mem(void) { return 0; }                                                              \
  extern "C" int NAME##_init_mem(int) { return 0; }                                                                \
  extern "C" void NAME##_free_mem(int) {}                                                                          \
  extern "C" int NAME##_checkout(void) { return 0; }                                                               \
  extern "C" void NAME##_release(int) {}                                                                           \
  extern "C" void NAME##_incref(void) {}                                                                           \
  extern "C" void NAME##_decref(void) {}

MPX_CASADI_DEFINE(nlp_f, 0, 1)
MPX_CASADI_DEFINE(nlp_g, 1, 1)
MPX_CASADI_DEFINE(nlp_grad_f, 2, 2)
MPX_CASADI_DEFINE(nlp_jac_g, 3, 2)

// nlp_hess_l (x, p, lam_f, lam_g) -> (hess_gamma_x_x): triu of the Lagrangian Hessian in CCS
extern "C" int nlp_hess_l(const double** arg, double** res, long long*, double*, int) {
  if (!B.plan || B.sp_hess.empty() || !arg || !res || !arg[0] || p_missing(B.plan, arg[1]) || !arg[2] || !arg[3]) return 1;
  if (!res[0]) return 0;
  return mpx_eval_hess_l(B.plan, arg[0], arg[1], arg[2][0], arg[3], res[0]) == MPX_OK ? 0 : 1;
}
extern "C" long long nlp_hess_l_n_in(void) { return 4; }
extern "C" long long nlp_hess_l_n_out(void) { return 1; }
extern "C" double nlp_hess_l_default_in(long long) { return 0.0; }
extern "C" const char* nlp_hess_l_name_in(long long i) {
  static const char* n[4] = {"x", "p", "lam_f", "lam_g"};
  return i >= 0 && i < 4 ? n[i] : nullptr;
}
extern "C" const char* nlp_hess_l_name_out(long long i) { return i == 0 ? "hess_gamma_x_x" : nullptr; }
extern "C" const long long* nlp_hess_l_sparsity_in(long long i) {
  if (!B.plan) return nullptr;
  return i == 0 ? B.sp_x.data() : i == 1 ? B.sp_p.data() : i == 2 ? B.sp_f.data() : i == 3 ? B.sp_g.data() : nullptr;
}
extern "C" const long long* nlp_hess_l_sparsity_out(long long i) {
  return (B.plan && i == 0 && !B.sp_hess.empty()) ? B.sp_hess.data() : nullptr;
}
extern "C" int nlp_hess_l_work(long long* sz_arg, long long* sz_res, long long* sz_iw, long long* sz_w) {
  if (sz_arg) *sz_arg = 4;
  if (sz_res) *sz_res = 1;
  if (sz_iw) *sz_iw = 0;
  if (sz_w) *sz_w = 0;
  return 0;
}
extern "C" int nlp_hess_l_alloc_mem(void) { return 0; }
extern "C" int nlp_hess_l_init_mem(int) { return 0; }
extern "C" void nlp_hess_l_free_mem(int) {}
extern "C" int nlp_hess_l_checkout(void) { return 0; }
extern "C" void nlp_hess_l_release(int) {}
extern "C" void nlp_hess_l_incref(void) {}
extern "C" void nlp_hess_l_decref(void) {}
