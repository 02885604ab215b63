/*
 * mpx.h -- C ABI of the B200-native collocation-transcription hot path of mpopt.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * The reference (pure Python over CasADi, /root/reference/mpopt/mpopt.py) has no FFI
 * of its own; each entry point below names the reference interface it replaces.
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - all values are float64, all indices int64 (CasADi's casadi_int), row/col 0-based;
 *   - decision vector z, parameters p (segment-width fractions), constraints g and the
 *     Jacobian follow the reference's layout exactly (mpopt.py:537-543, :631, :458,
 *     :617-621); the Jacobian is CSR with sorted column indices (CCS adapter provided);
 *   - every function returns 0 on success, a negative MPX_E* code otherwise and never
 *     throws; mpx_last_error() gives the message of the last failure on this thread;
 *   - host-pointer entry points copy z/p to the device, run the kernels, copy results
 *     back; *_dev entry points take device pointers and a cudaStream_t (as void*) and
 *     do no host work beyond the launch;
 *   - a plan owns its device buffers and one CUDA stream; entry points are not
 *     re-entrant on the same plan (ONE evaluation per plan in flight: the launches share
 *     the plan's argument blocks and scratch buffers, also through the *_dev entry points on
 *     caller streams), different plans may be used concurrently.  *_dev entry points make
 *     the plan's device current on the calling thread.
 *   - there is NO CPU fallback: without a CUDA device mpx_plan_create fails with
 *     MPX_ENODEVICE.
 */
#ifndef MPX_H_
#define MPX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPX_VERSION 100

#define MPX_OK 0
#define MPX_EINVAL (-1)    /* bad argument / inconsistent description           */
#define MPX_ENODEVICE (-2) /* no usable CUDA device                             */
#define MPX_ECUDA (-3)     /* CUDA runtime / driver / NVRTC failure             */
#define MPX_ENOPROGRAM (-4)/* no compiled node functors for this problem        */
#define MPX_ELIMIT (-5)    /* problem exceeds a compiled-in limit (smem, sizes) */

/* Collocation schemes: CollocationRoots.get_collocation_points, mpopt.py:4157-4188 */
#define MPX_LGR 0
#define MPX_LGL 1
#define MPX_CGL 2

#define MPX_MAX_DIM 16 /* max of n_states, n_controls, n_params, path rows per node */

/* One phase of the OCP: structural information obtained by tracing the user's Python
 * callables once (what CasADi's SX graph + AD sparsity give the reference at mpopt.py:757).
 * Pattern matrices are uint8, row-major. Node variables are ordered x_0.., u_0.., a_0..;
 * terminal variables xf_0.., x0_0.., tf, t0, a_0.. */
typedef struct mpx_phase_desc {
  int32_t n_path;          /* path-constraint rows per node (0: none), mpopt.py:239-262    */
  int32_t n_term;          /* terminal-constraint rows, mpopt.py:264-300                   */
  const uint8_t* pat_f;    /* [nx][nx+nu+na]   d f_s / d var  structurally nonzero         */
  const uint8_t* f_nz;     /* [nx]             f_s is not identically zero                 */
  const uint8_t* f_t;      /* [nx]             f_s depends on t                            */
  const uint8_t* pat_c;    /* [n_path][nx+nu+na]                                           */
  const uint8_t* c_t;      /* [n_path]                                                     */
  const uint8_t* pat_tc;   /* [n_term][2nx+2+na]                                           */
  int32_t diff_u;          /* control-slope rows present, mpopt.py:302-328                 */
  int32_t midu;            /* mid-point control rows present, mpopt.py:330-377             */
  int32_t du_continuity;   /* slope-continuity rows present (needs n_segments>1), :379-413 */
  int32_t cost_t;          /* running cost depends on t explicitly                         */
  /* second-derivative patterns for the Hessian of the Lagrangian (may be NULL: then mpx_*hess* fail);
   * lower triangle used, row-major, variables (x.., u.., a.., t, h) resp. (xf.., x0.., tf, t0, a..) */
  const uint8_t* pat_hw;   /* [nv+2][nv+2]  node Lagrangian h (sw L - sum lamF Sx f) + sum lamC c     */
  const uint8_t* pat_ht;   /* [2nx+2+na][2nx+2+na]  sw M + sum lamT tc                                */
  /* adaptive NLP only (mpx_problem_desc.adaptive): mid-point rows of the SW block, mpopt.py:3062-3082 */
  int32_t sw_u;            /* any control bound finite: compI.U rows present                */
  int32_t sw_x;            /* any state bound finite: compI.X rows present                  */
  /* Hessian of the adaptive NLP (may be NULL / 0 otherwise): second-derivative pattern of sum_s mu_s f_s w.r.t.
   * (x.., u.., a..), lower triangle, row-major [nv][nv] (the mid-point residual rows, mpopt.py:3084-3136), and whether
   * the coefficient of h in the node Lagrangian, sw L - sum lamF Sx f, is not identically zero                        */
  const uint8_t* pat_hf;
  int32_t phi_nz;
} mpx_phase_desc;

typedef struct mpx_problem_desc {
  int32_t nx, nu, na, n_phases;
  const mpx_phase_desc* phases;   /* [n_phases]                                            */
  int32_t n_segments;             /* per phase, mpopt.py:70                                */
  const int32_t* poly_orders;     /* [n_segments], shared by all phases, mpopt.py:73-75    */
  int32_t scheme;                 /* MPX_LGR / MPX_LGL / MPX_CGL                           */
  double tau_min, tau_max;        /* CollocationRoots._TAU_MIN/_TAU_MAX, mpopt.py:4144-4145*/
  const double* scale_x;          /* [nx]  OCP.scale_x, mpopt.py:3445-3448                 */
  const double* scale_u;          /* [nu]                                                  */
  const double* scale_a;          /* [na]                                                  */
  double scale_t;
  int32_t n_links;                /* phase links, mpopt.py:464-521                         */
  const int32_t* links;           /* [n_links][2] = (phase_i, phase_j)                     */
  int32_t drop_exact_zeros;       /* SX folds 0*x: exact-zero table entries leave the pattern */
  const char* program_key;        /* key of the AOT-compiled node functors (may be NULL)   */
  const char* program_source;     /* generated CUDA source of the node functors for the
                                     NVRTC path (may be NULL if program_key is registered) */
  int32_t device;                 /* CUDA device ordinal                                   */
  int32_t seg_begin, seg_end;     /* shard: evaluate segments [seg_begin, seg_end) of every
                                     phase; 0,0 means all. Tail rows (terminal, events)
                                     belong to the shard that owns the last segment.       */
  int32_t adaptive;               /* 1: the NLP of mpopt_adaptive (mpopt.py:2877-3375): the segment
                                     widths are decision variables appended to every phase of z
                                     ([X | U | t0 | tf | a | w], :2938-2945), n_p = 0, rows per phase
                                     [F | C | DU | TC | SW] (:3169; midu / du_continuity are ignored).
                                     Not available for shards, peers or the Hessian.       */
  int32_t mid_residuals;          /* adaptive: mid-point residual rows present (:2918, :3084-3130) */
} mpx_problem_desc;

typedef struct mpx_plan mpx_plan;

int mpx_version(void);
const char* mpx_last_error(void);

/* -- collocation tables: replaces CollocationRoots._taus_fn, Collocation.get_diff_matrix,
 *    get_quadrature_weights, get_interpolation_matrix (mpopt.py:4208-4276, :3815-3905).
 *    Computed on the device. roots[deg+1], D[(deg+1)^2] row-major, w[deg+1] (integral over
 *    [tau_min,tau_max]), Cmid[deg*(deg+1)] = basis at the mid-points (mpopt.py:350-359).
 *    Any output pointer may be NULL. */
int mpx_collocation_tables(int32_t scheme, int32_t deg, double tau_min, double tau_max, int32_t device,
                           double* roots, double* D, double* w, double* Cmid);

/* -- basis at arbitrary points: get_diff_matrix(key, taus, order) / get_interpolation_matrix
 *    (mpopt.py:3815-3849, :3884-3905) and get_quadrature_weights(key, tau0, tau1) (:3851-3882).
 *    order 0 -> C[n_taus][deg+1] = l_j(tau_i); order 1|2 -> derivatives. */
int mpx_collocation_basis_at(int32_t scheme, int32_t deg, double tau_min, double tau_max, int32_t device,
                             int32_t order, int32_t n_taus, const double* taus, double* out);
int mpx_collocation_weights(int32_t scheme, int32_t deg, double tau_min, double tau_max, int32_t device,
                            double tau0, double tau1, double* w);

/* -- plan: replaces mpopt.create_nlp + create_solver's ca.nlpsol(...) function derivation
 *    (mpopt.py:574-639, :725-758). */
int mpx_plan_create(const mpx_problem_desc* desc, mpx_plan** out);
void mpx_plan_destroy(mpx_plan* plan);

int mpx_sizes(const mpx_plan* plan, int64_t* n_z, int64_t* n_p, int64_t* n_g, int64_t* nnz_jac);
/* CSR pattern of jac_g: rowptr[n_g+1], colind[nnz] (sorted within each row). */
int mpx_jac_structure(const mpx_plan* plan, int64_t* rowptr, int64_t* colind);
/* CCS pattern (CasADi's native order) + permutation: ccs_values[i] = csr_values[perm[i]]. */
int mpx_jac_structure_ccs(const mpx_plan* plan, int64_t* colptr, int64_t* rowind, int64_t* perm);
/* tables the plan uses for one degree (same layout as mpx_collocation_tables) */
int mpx_plan_tables(const mpx_plan* plan, int32_t deg, double* roots, double* D, double* w, double* Cmid);
/* contiguous runs of g / values written by this plan's shard: pairs (offset, count).
 * Call with runs == NULL to get the count. kind 0: g, 1: jac values, 2: grad_f. */
int mpx_shard_runs(const mpx_plan* plan, int32_t kind, int64_t* runs, int64_t* n_runs);

/* -- evaluators: replace CasADi's nlp_f / nlp_grad_f / nlp_g / nlp_jac_g (derived at
 *    mpopt.py:757, called inside mpopt.py:804). z[n_z], p[n_p] host pointers. */
int mpx_eval_f(mpx_plan* plan, const double* z, const double* p, double* f);
int mpx_eval_grad_f(mpx_plan* plan, const double* z, const double* p, double* f_or_null, double* grad);
int mpx_eval_g(mpx_plan* plan, const double* z, const double* p, double* g);
int mpx_eval_jac_g(mpx_plan* plan, const double* z, const double* p, double* g_or_null, double* values);

/* -- device-pointer variants (stream: cudaStream_t, NULL = the plan's own stream).
 *    f_dev: 1 double; they enqueue work and return without synchronising. */
int mpx_eval_f_grad_dev(mpx_plan* plan, const double* d_z, const double* d_p, double* d_f, double* d_grad_or_null,
                        void* stream);
int mpx_eval_g_jac_dev(mpx_plan* plan, const double* d_z, const double* d_p, double* d_g, double* d_values_or_null,
                       void* stream);
int mpx_sync(mpx_plan* plan);
/* -- the host hop. The reference's consumers (IPOPT through CasADi, mpopt.py:804) own PAGEABLE host buffers, so every
 *    evaluation ends with a device-to-host copy that dwarfs the kernel (105 MB of Jacobian values at the headline
 *    size: 2 ms over PCIe 5 against a 19 us kernel).
 *    - Every host-pointer entry point moves pageable buffers through a plan-owned pinned staging ring, copied in / out
 *      by a small worker pool while the next chunk is in flight (link speed instead of the driver's bounce buffers);
 *      MPX_HOST_THREADS sets the pool size. Pinned or registered buffers are used directly.
 *    - mpx_host_register pins a caller-owned range (cudaHostRegister) for the life of the plan or until
 *      mpx_host_unregister; the caller must unregister before freeing the memory.
 *    - mpx_eval_jac_g_dynamic moves only the n_dynamic entries that depend on z or p (the off-block partials, the
 *      merged D diagonals, d/dT0, d/dTF, path and terminal rows): the first call on a given `values`
 *      buffer is a full mpx_eval_jac_g, later calls ON THE SAME BUFFER rewrite just those entries in place (packed
 *      copy + parallel scatter). The caller must not modify `values` in between.
 *    - mpx_eval_jac_g_packed returns the dynamic entries packed (n_dynamic doubles, positions from
 *      mpx_jac_dynamic_positions, ascending CSR order) for callers that keep their own copy of the constants. */
int mpx_host_register(mpx_plan* plan, void* ptr, int64_t bytes);
int mpx_host_unregister(mpx_plan* plan, void* ptr);
int mpx_jac_dynamic_count(mpx_plan* plan, int64_t* n_dynamic);
int mpx_jac_dynamic_positions(mpx_plan* plan, int32_t* positions /* n_dynamic */);
int mpx_eval_jac_g_dynamic(mpx_plan* plan, const double* z, const double* p, double* g_or_null, double* values);
int mpx_eval_jac_g_packed(mpx_plan* plan, const double* z, const double* p, double* g_or_null, double* packed);

/* -- measurement aids (bench.py, profiles/tools): mpx_gate occupies `stream` for usec microseconds so that a timed
 *    region can be enqueued completely before the device starts on it (host launch latency stays outside the CUDA
 *    events); mpx_trace_read returns the per-warp timeline of the last g + jac_g launch of a plan created under
 *    MPX_TRACE=1 (n_warps records of `slots` 64-bit stamps; out == NULL: only the counts). */
int mpx_gate(void* stream, double usec);
int mpx_trace_read(mpx_plan* plan, int64_t* n_warps, int64_t* slots, unsigned long long* out);
/*    mpx_hess_zero_fill: whether the Hessian of an adaptive plan (mpopt_adaptive's NLP) zero-fills its output before
 *    the kernels run: -1 not evaluated yet, 0 no (the first evaluation ran on a NaN-filled buffer and every entry of
 *    the pattern turned out to have a writer), 1 yes.  Always 0 for a plan that is not adaptive. */
int mpx_hess_zero_fill(const mpx_plan* plan, int* state);

/* -- fused evaluation + all-gather over peer memory (multi-GPU, one process per GPU): every store of the g + jac_g
 *    kernel is issued to this GPU's buffers AND to the same offsets of n_peers peer buffers (other GPUs' allocations
 *    mapped through CUDA IPC, reached over NVLink), so that once all ranks have evaluated their shard every rank
 *    holds the whole g / Jacobian -- the all-gather of BASELINE.json's north_star without a separate collective.
 *    Buffers that peers write into must come from mpx_peer_alloc (cudaMalloc + IPC handle); the 64-byte handle is
 *    sent to the peers (any transport), which map it with mpx_peer_open. The caller orders the ranks before reading. */
typedef struct mpx_ipc_handle { unsigned char bytes[64]; } mpx_ipc_handle;
int mpx_peer_alloc(int32_t device, int64_t bytes, void** dptr, mpx_ipc_handle* handle);
int mpx_peer_open(int32_t device, const mpx_ipc_handle* handle, void** dptr);
int mpx_peer_close(void* dptr);
int mpx_peer_free(void* dptr);
int mpx_eval_g_jac_dev_peers(mpx_plan* plan, const double* d_z, const double* d_p, double* d_g, double* d_values,
                             int32_t n_peers, double* const* peer_g, double* const* peer_values, void* stream);

/* -- Hessian of the Lagrangian lam_f * f + lam_g . g: replaces CasADi's nlp_hess_l(x, p, lam_f, lam_g) (derived at
 *    mpopt.py:757; IPOPT's eval_h). LOWER triangle in CSR with sorted columns -- which is also the upper triangle in
 *    CCS, the form CasADi returns. Built on first use. */
int mpx_hess_structure(mpx_plan* plan, int64_t* nnz_hess, int64_t* rowptr /* n_z+1 */, int64_t* colind);
int mpx_eval_hess_l(mpx_plan* plan, const double* z, const double* p, double lam_f, const double* lam_g, double* values);
int mpx_eval_hess_l_dev(mpx_plan* plan, const double* d_z, const double* d_p, double lam_f, const double* d_lam_g,
                        double* d_values, void* stream);

/* -- interpolation of a solution and the dynamics residual at arbitrary points: replaces, per phase,
 *    mpopt.interpolate_single_phase + get_dynamics_residuals_single_phase (mpopt.py:1428-1542), the step after every
 *    solve (process_results) and the inner loop of mpopt_h_adaptive. Point i lies in segment seg[i] at the local
 *    abscissa taus[i] in [tau_min, tau_max] (the reference's per-segment tau lists, flattened). Outputs, any of which
 *    may be NULL, row-major: xi[n][nx], ui[n][nu] (interpolated, scaled variables), ti[n] (time), dxi[n][nx],
 *    dui[n][nu] (d/dtau through the segment's Lagrange basis), res[n][nx] = dxi - h_seg Sx f(xi/Sx, ui/Su, ti, a/Sa). */
int mpx_eval_residuals(mpx_plan* plan, const double* z, const double* p, int32_t phase, int64_t n_points,
                       const int32_t* seg, const double* taus, double* xi, double* ui, double* ti, double* dxi,
                       double* dui, double* res);
/* -- second derivatives of the interpolants at arbitrary points: replaces mpopt.get_state_second_derivative_single_phase
 *    (mpopt.py:1285-1358; composite get_diff_matrix(order=2) times X and U). Same point list as above; ddxi[n][nx],
 *    ddui[n][nu] are d2/dtau2 through the segment's Lagrange basis (either may be NULL, not both), ti[n] may be NULL. */
int mpx_eval_second_derivatives(mpx_plan* plan, const double* z, const double* p, int32_t phase, int64_t n_points,
                                const int32_t* seg, const double* taus, double* ti, double* ddxi, double* ddui);
/* -- state residual by quadrature: replaces mpopt.compute_states_from_solution_dynamics / get_states_residuals
 *    (mpopt.py:989-1150). Per segment, h Sx f evaluated at the segment's target points is interpolated through those
 *    points and integrated from tau_min to every target point; xint[n][nx] = x(segment start) + integral (scaled
 *    states), res_x[n][nx] = interpolated state - xint. Points must be listed segment by segment (as the reference's
 *    per-segment tau lists are). ui[n][nu], ti[n] as in mpx_eval_residuals; any output may be NULL except that one of
 *    xint / res_x is required. */
int mpx_eval_state_residuals(mpx_plan* plan, const double* z, const double* p, int32_t phase, int64_t n_points,
                             const int32_t* seg, const double* taus, double* xint, double* ui, double* ti, double* res_x);

/* -- staged evaluation: ONE upload and ONE fused evaluation per distinct x, results kept in the plan's device
 *    buffers; the pieces are copied out when asked for. This is how the solver-facing shims below honour IPOPT's
 *    new_x flag (eval_g and eval_jac_g of the same x share one kernel launch) and how CasADi's nlp_jac_g gets its
 *    values in column-compressed order (device gather through the static permutation of mpx_jac_structure_ccs). */
#define MPX_STAGE_F 1
#define MPX_STAGE_GRAD 2
#define MPX_STAGE_G 4
#define MPX_STAGE_JAC 8
#define MPX_FETCH_JAC_CCS 16 /* mpx_fetch only: Jacobian values in CCS order */
int mpx_stage(mpx_plan* plan, const double* z, const double* p, int32_t what /* MPX_STAGE_* bits */);
int mpx_staged(const mpx_plan* plan);                       /* bits valid for the x of the last mpx_stage */
int mpx_fetch(mpx_plan* plan, int32_t what, double* out);   /* F: 1, GRAD: n_z, G: n_g, JAC / JAC_CCS: nnz doubles */
/* Hessian of the Lagrangian at the x of the last mpx_stage (no upload of x; staged results stay valid). Any host
 * entry point that uploads another x (mpx_eval_*) invalidates what is staged: mpx_staged() returns 0 afterwards. */
int mpx_hess_l_staged(mpx_plan* plan, double lam_f, const double* lam_g, double* values);

/* -- IPOPT C interface (IpStdCInterface.h: Eval_F_CB, Eval_Grad_F_CB, Eval_G_CB, Eval_Jac_G_CB): pass these four
 *    functions to CreateIpoptProblem and a filled mpx_ipopt_data as user_data. Index = int, Number = double,
 *    Bool = int (only the low byte of new_x is read, so a C99 bool works too); index_style 0 (C). eval_jac_g / eval_h
 *    with values == NULL write the pattern as triplets. */
typedef struct mpx_ipopt_data {
  mpx_plan* plan;
  const double* p;  /* segment-width fractions (the NLP parameter vector, mpopt.py:631) */
} mpx_ipopt_data;
int mpx_ipopt_eval_f(int n, const double* x, int new_x, double* obj_value, void* user_data);
int mpx_ipopt_eval_grad_f(int n, const double* x, int new_x, double* grad_f, void* user_data);
int mpx_ipopt_eval_g(int n, const double* x, int new_x, int m, double* g, void* user_data);
int mpx_ipopt_eval_jac_g(int n, const double* x, int new_x, int m, int nele_jac, int* iRow, int* jCol, double* values,
                         void* user_data);
/* Eval_H_CB: lower triangle of obj_factor * hess f + sum lambda_i hess g_i (nele_hess from mpx_hess_structure) */
int mpx_ipopt_eval_h(int n, const double* x, int new_x, double obj_factor, int m, const double* lambda, int new_lambda,
                     int nele_hess, int* iRow, int* jCol, double* values, void* user_data);

/* -- CasADi external functions (the ABI of CasADi's generated C code, as loaded by ca.external(name, lib) and by
 *    ca.nlpsol(name, plugin, lib)): nlp_f (x,p)->(f), nlp_g (x,p)->(g), nlp_grad_f (x,p)->(f, grad_f_x),
 *    nlp_jac_g (x,p)->(g, jac_g_x in CCS) -- the functions ca.nlpsol derives at mpopt.py:757. The symbols are fixed
 *    by that ABI, so they evaluate the plan bound with mpx_casadi_bind (one plan at a time per process). Each
 *    NAME comes with NAME_n_in/_n_out/_name_in/_name_out/_sparsity_in/_sparsity_out/_work/_incref/_decref/
 *    _alloc_mem/_init_mem/_free_mem/_checkout/_release/_default_in (declared in mpx_casadi.h style below). */
int mpx_casadi_bind(mpx_plan* plan);
#define MPX_CASADI_DECLARE(NAME)                                                                      \
  int NAME(const double** arg, double** res, long long* iw, double* w, int mem);                      \
  long long NAME##_n_in(void);                                                                        \
  long long NAME##_n_out(void);                                                                       \
  double NAME##_default_in(long long i);                                                              \
  const char* NAME##_name_in(long long i);                                                            \
  const char* NAME##_name_out(long long i);                                                           \
  const long long* NAME##_sparsity_in(long long i);                                                   \
  const long long* NAME##_sparsity_out(long long i);                                                  \
  int NAME##_work(long long* sz_arg, long long* sz_res, long long* sz_iw, long long* sz_w);           \
  int NAME##_alloc_mem(void);                                                                         \
  int NAME##_init_mem(int mem);                                                                       \
  void NAME##_free_mem(int mem);                                                                      \
  int NAME##_checkout(void);                                                                          \
  void NAME##_release(int mem);                                                                       \
  void NAME##_incref(void);                                                                           \
  void NAME##_decref(void);
MPX_CASADI_DECLARE(nlp_f)
MPX_CASADI_DECLARE(nlp_g)
MPX_CASADI_DECLARE(nlp_grad_f)
MPX_CASADI_DECLARE(nlp_jac_g)
MPX_CASADI_DECLARE(nlp_hess_l) /* (x, p, lam_f, lam_g) -> (hess_gamma_x_x): upper triangle, CCS */

/* number of kernel launches issued by this plan so far (bench.py's gpu_launches) */
int64_t mpx_launch_count(const mpx_plan* plan);
/* human-readable description of how the node functors were obtained ("aot:<key>" / "nvrtc:<key>") */
const char* mpx_program_origin(const mpx_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* MPX_H_ */
