// mpx_plan.cu -- the C ABI of include/mpx.h: plan creation (layout, CSR structure, device
// tables), the program registry, and the evaluators that launch the kernels of
// mpx_kernels.cuh.  Host code here is index logic only; every floating-point value of the
// hot path is produced on the device.
//
// Layout being reproduced (reference: /root/reference/mpopt/mpopt.py):
//   variables  :537-543, :627      z = per phase [X(:,0..nx-1) | U(:,0..nu-1) | t0 | tf | a]
//   rows       :458, :617-621      per phase [F | C | DU | mU | dU | TC], then the event blocks
//   ownership  :189-195            a shared segment-boundary node belongs to the earlier segment
//   staircase  :4015-4039          composite D; :4066-4096 composite mid-point interpolation
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/mpx.h"
#include "mpx_kernels.cuh"
#include "mpx_program.h"
#include "mpx_tables.cuh"

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                                 \
  do {                                                                                                 \
    cudaError_t e_ = (expr);                                                                           \
    if (e_ != cudaSuccess)                                                                             \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? MPX_ENODEVICE : MPX_ECUDA, \
                  std::string(#expr) + ": " + cudaGetErrorString(e_));                                 \
  } while (0)

extern "C" int mpx_version(void) { return MPX_VERSION; }
extern "C" const char* mpx_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------ program registry
static MpxProgramEntry* g_programs = nullptr;
extern "C" void mpx_register_program(MpxProgramEntry* e) {
  e->next = g_programs;
  g_programs = e;
}
const MpxProgramEntry* mpx_find_program(const char* key) {
  if (!key) return nullptr;
  for (MpxProgramEntry* e = g_programs; e; e = e->next)
    if (strcmp(e->key, key) == 0) return e;
  return nullptr;
}

// ------------------------------------------------------------------ small generic kernels
// slope-continuity rows (mpopt.py:379-413): row (c,k) = D_k(tau1) - D_{k+1}(tau0) applied to U(:,c)
struct MpxDuArgs {
  const double* z;
  const double* tabs;
  const int32_t* seg_tab;
  const int32_t* seg_start;
  const int64_t* seg_spre;
  double* g;
  double* vals;
  int32_t K, N, nx, nu;
  int64_t zoff, gdU, vdU, nnzS;
};
__global__ void mpx_ducont_kernel(const MpxDuArgs A, int k_begin, int k_end, int jac) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nk = k_end - k_begin;
  if (i >= nk * A.nu) return;
  const int c = i / nk, k = k_begin + (i - c * nk);
  const int s0 = A.seg_start[k], s1 = A.seg_start[k + 1], s2 = A.seg_start[k + 2];
  const int da = s1 - s0, db = s2 - s1, na1 = da + 1, nb1 = db + 1;
  const double* Da = A.tabs + A.seg_tab[k] + MpxTab::off_D(na1) + da * na1;  // last row of D_k
  const double* Db = A.tabs + A.seg_tab[k + 1] + MpxTab::off_D(nb1);        // first row of D_{k+1}
  const double* U = A.z + A.zoff + (int64_t)(A.nx + c) * A.N;
  double* v = A.vals + A.vdU + (int64_t)c * A.nnzS + A.seg_spre[k];
  double acc = 0.0;
  for (int j = 0; j < da; ++j) {
    acc = fma(Da[j], U[s0 + j], acc);
    if (jac) v[j] = Da[j];
  }
  const double mid = Da[da] - Db[0];
  acc = fma(mid, U[s1], acc);
  if (jac) v[da] = mid;
  for (int j = 1; j <= db; ++j) {
    acc = fma(-Db[j], U[s1 + j], acc);
    if (jac) v[da + j] = -Db[j];
  }
  A.g[A.gdU + (int64_t)c * (A.K - 1) + k] = acc;
}

// phase-link rows (mpopt.py:464-521): one thread per row (MpxEvArgs / mpx_event_row: csrc/mpx_kernels.cuh)
__global__ void mpx_events_kernel(const MpxEvArgs A, int jac) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < A.rows) mpx_event_row(A, r, jac);
}

// drop masked entries (exact-zero table values, SX folding -- SURVEY.md Q10)
__global__ void mpx_compact_kernel(const double* __restrict__ full, const int64_t* __restrict__ map, double* out,
                                   int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = full[map[i]];
}

// dynamic fetch: the z- / p-dependent Jacobian entries gathered into a contiguous device buffer (ONE device-to-host copy)
// how many entries of v are NaN (one-time coverage probe of the adaptive Hessian, see launch_hess)
__global__ void mpx_count_nan_kernel(const double* __restrict__ v, int64_t n, unsigned long long* count) {
  unsigned long long c = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c += v[i] != v[i];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

__global__ void mpx_gather_dyn_kernel(const double* __restrict__ vals, const int32_t* __restrict__ pos,
                                      double* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = vals[pos[i]];
}

// state residual by quadrature (mpopt.compute_states_from_solution_dynamics, mpopt.py:989-1076): per segment, the
// Lagrange interpolant of F = h Sx f through the segment's own target points is integrated from tau0 to every target
// point (Gauss-Legendre, exact for the interpolant), x_int = x(segment start) + integral, residual = x_I - x_int.
// F comes from the residual kernel (dxi - res).  One CTA per segment, one thread per target point.
struct MpxSrArgs {
  const double* z;          // phase slice of the decision vector (state-major)
  const int32_t* seg_start; // [K+1]
  const int32_t* pt_off;    // [K+1] first point of every segment
  const double* tau;        // [n] local abscissae
  const double* xi;         // [n][nx] interpolated states
  const double* dxi;        // [n][nx]
  const double* res;        // [n][nx] dynamics residual (dxi - h Sx f)
  double* xint;             // [n][nx] out
  double* rx;               // [n][nx] out
  int32_t N, nx, max_pts;
  double tau0;
};
#define MPX_SR_THREADS 64
__global__ void __launch_bounds__(MPX_SR_THREADS) mpx_state_resid_kernel(const MpxSrArgs A) {
  extern __shared__ __align__(16) double sm[];
  const int k = blockIdx.x, tid = threadIdx.x;
  const int p0 = A.pt_off[k], nt = A.pt_off[k + 1] - p0;
  if (nt == 0) return;
  const int nx = A.nx, nq = nt / 2 + 1;
  double* sT = sm;                    // target abscissae (the custom roots)
  double* sB = sT + A.max_pts;        // 1 / prod_{i != m} (r_m - r_i)
  double* sF = sB + A.max_pts;        // F[m][s]
  double* xq = sF + A.max_pts * nx;
  double* wq = xq + A.max_pts;
  for (int i = tid; i < nt; i += MPX_SR_THREADS) sT[i] = A.tau[p0 + i];
  for (int i = tid; i < nt * nx; i += MPX_SR_THREADS) sF[i] = A.dxi[(int64_t)p0 * nx + i] - A.res[(int64_t)p0 * nx + i];
  __syncthreads();
  for (int m = tid; m < nt; m += MPX_SR_THREADS) {
    double b = 1.0;
    for (int i = 0; i < nt; ++i)
      if (i != m) b *= sT[m] - sT[i];
    sB[m] = 1.0 / b;
  }
  mpx_gauss_legendre(nq, xq, wq);  // ends with a barrier
  for (int i = tid; i < nt; i += MPX_SR_THREADS) {
    const double ta = A.tau0, tb = sT[i], half = 0.5 * (tb - ta);
    double acc[MPX_MAXS];
    for (int s = 0; s < nx; ++s) acc[s] = 0.0;
    for (int q = 0; q < nq; ++q) {
      const double xi_q = ta + half * (xq[q] + 1.0);
      for (int m = 0; m < nt; ++m) {
        double l = sB[m];
        for (int j = 0; j < nt; ++j)
          if (j != m) l *= xi_q - sT[j];
        l *= wq[q];
        for (int s = 0; s < nx; ++s) acc[s] = fma(l, sF[m * nx + s], acc[s]);
      }
    }
    for (int s = 0; s < nx; ++s) {
      const double xstart = A.z[(int64_t)s * A.N + A.seg_start[k]];
      const double v = xstart + half * acc[s];
      A.xint[(int64_t)(p0 + i) * nx + s] = v;
      A.rx[(int64_t)(p0 + i) * nx + s] = A.xi[(int64_t)(p0 + i) * nx + s] - v;
    }
  }
}

// exclusive prefix sum of the segment widths of every phase (time grid, mpopt.py:192)
#define MPX_SCAN_THREADS 1024
__global__ void __launch_bounds__(MPX_SCAN_THREADS) mpx_scan_widths_kernel(const double* w, double* sig0, int K,
                                                                           int64_t stride) {
  // block-wide scan, one CTA per phase: every thread sums a contiguous chunk, the chunk totals are scanned with
  // shuffles (fixed association, so the result is deterministic), then each thread writes its chunk's prefixes
  __shared__ double wsum[MPX_SCAN_THREADS / 32];
  const double* wp = w + (int64_t)blockIdx.x * stride;  // stride K: parameter vector; nvar: widths inside z (adaptive)
  double* sp = sig0 + (int64_t)blockIdx.x * K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = (K + MPX_SCAN_THREADS - 1) / MPX_SCAN_THREADS;
  const int k0 = min(tid * chunk, K), k1 = min(k0 + chunk, K);
  double s = 0.0;
  for (int k = k0; k < k1; ++k) s += wp[k];
  double incl = s;
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  double excl = __shfl_up_sync(0xffffffffu, incl, 1);  // exclusive prefix of this thread's chunk within its warp
  if (lane == 0) excl = 0.0;
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    double v = wsum[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    wsum[lane] = v;  // inclusive totals of the warps
  }
  __syncthreads();
  double acc = excl + (warp > 0 ? wsum[warp - 1] : 0.0);
  for (int k = k0; k < k1; ++k) {
    sp[k] = acc;
    acc += wp[k];
  }
}

// ------------------------------------------------------------------ host side of the host hop
// A small persistent worker pool: parallel_for(n, fn) runs fn(i) for i in [0, n) on the workers and the caller.
// Used to move bytes between plan-owned pinned staging and caller-owned PAGEABLE buffers at memory speed (one thread
// copies ~10 GB/s; the PCIe link delivers 54 GB/s) and to scatter the packed dynamic Jacobian entries.
class MpxPool {
 public:
  explicit MpxPool(int n) {
    for (int t = 0; t < n; ++t) workers_.emplace_back([this] { loop(); });
  }
  ~MpxPool() {
    {
      std::lock_guard<std::mutex> l(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  int size() const { return (int)workers_.size() + 1; }
  void parallel_for(int n, const std::function<void(int)>& fn) {
    if (n <= 0) return;
    {
      std::lock_guard<std::mutex> l(m_);
      fn_ = &fn, next_ = 0, total_ = n, pending_ = n, ++epoch_;
      epoch_a_.store(epoch_, std::memory_order_release);
    }
    cv_.notify_all();
    run();
    std::unique_lock<std::mutex> l(m_);
    done_.wait(l, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void run() {
    for (;;) {
      int i;
      const std::function<void(int)>* f;
      {
        std::lock_guard<std::mutex> l(m_);
        if (!fn_ || next_ >= total_) return;
        i = next_++, f = fn_;
      }
      (*f)(i);
      std::lock_guard<std::mutex> l(m_);
      if (--pending_ == 0) done_.notify_all();
    }
  }
  void loop() {
    unsigned long seen = 0;
    for (;;) {
      // the chunks of one transfer arrive ~100 us apart: spin that long before sleeping on the condition variable
      // (a wake-up through the futex costs about as much as copying a worker's share of a chunk)
      const auto t0 = std::chrono::steady_clock::now();
      bool got = false;
      while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(150)) {
        if (epoch_a_.load(std::memory_order_acquire) != seen) {
          got = true;
          break;
        }
      }
      {
        std::unique_lock<std::mutex> l(m_);
        if (!got) cv_.wait(l, [&] { return stop_ || epoch_ != seen; });
        if (stop_) return;
        seen = epoch_;
      }
      run();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int next_ = 0, total_ = 0, pending_ = 0;
  unsigned long epoch_ = 0;
  std::atomic<unsigned long> epoch_a_{0};
  bool stop_ = false;
};

// ------------------------------------------------------------------ plan
struct PhaseLayout {
  int nc = 0, ntc = 0;
  bool has_DU = false, has_mU = false, has_dU = false;
  std::vector<int> f_next, c_len, tc_len;
  std::vector<uint8_t> pat_f, f_nz, f_t, pat_c, c_t, pat_tc, pat_hw, pat_ht, pat_hf;
  bool phi_nz = false;
  bool has_hess = false;
  int64_t zoff = 0, n_g = 0;
  int64_t gF = 0, gC = 0, gDU = 0, gmU = 0, gdU = 0, gTC = 0;
  int64_t vF[MPX_MAXS], vC[MPX_MAXS];
  int64_t vDU = 0, vmU = 0, vdU = 0, vTC = 0;
  bool uses_t = false, cost_t = false;
  // adaptive NLP (mpopt_adaptive): SW block and the offsets of the entries mpx_adapt_kernel writes (into the ext section)
  bool sw_u = false, sw_x = false;
  int64_t gSW = 0, eSum = 0, eUi = 0, eXi = 0, eRes = 0;
  int64_t eF[MPX_MAXS], eC[MPX_MAXS];
  std::vector<int> res_nblk, res_na;  // per state: (d+1)-wide column blocks / parameter entries of a residual row
  std::vector<int64_t> seg_rpre;      // [K] residual-row entries before the segment
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr, o.bytes = 0; }
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t ensure(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr, bytes = 0;
    cudaError_t e = cudaMalloc(&p, n ? n : 8);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

struct mpx_plan {
  int nx, nu, na, P, K, N, scheme, device;
  double tau_min, tau_max, st;
  std::vector<int> po, seg_start;
  std::vector<double> sx, su, sa;
  std::vector<int> links;
  int seg_begin, seg_end;
  bool drop, uniform;
  bool adaptive = false, mid_res = false;  // widths-as-variables NLP (mpopt.py:2877-3375)
  bool wcol = false;         // adaptive: the base kernel writes the trailing d/dw_k column of the F rows (rows in place)
  bool base_direct = false;  // adaptive: every CSR entry is written at its final position (no gather pass)
  int64_t n_base = 0, n_ext = 0;           // staged values: [base kernels | mpx_adapt_kernel]
  std::vector<int> dmid_off;               // per unique degree: record offset in h_dmid
  std::vector<double> h_dmid;              // D at the mid points, [d][d+1] per degree
  DevBuf d_dmid, d_seg_dmid, d_wpart, d_seg_rpre;  // d_seg_rpre: [P][K]
  DevBuf d_ticket;                                 // [P] arrival counters of the single-launch f + grad_f kernel
  DevBuf d_queue;                                  // [P][4] work counters: K2's dynamic unit queue, mpx_adapt_kernel's segment queue (MPX_QUEUE=0: none)
  bool k2_queue = false;                           // K2 deals its units through the counter (MPX_K2_QUEUE=1; measured neutral, off by default)
  DevBuf d_rseg, d_rtau, d_rout;                   // mpx_eval_residuals: point list and outputs
  DevBuf d_sr_off, d_sr_out;                       // mpx_eval_state_residuals: point offsets per segment, outputs
  bool v2_spread = true;
  int smem_adapt = 0;
  int adapt_grid = 1 << 30;  // CTAs of the persistent mpx_adapt_kernel (MPX_ADAPT_CTAS per SM; segments when there is no queue)
  std::vector<int> adapt_img;              // per phase: doubles of a staged residual-row image (0: direct stores)
  std::vector<int64_t> sw_direct;          // per phase: CSR position of the SW block when it is written in place, else -1
  std::vector<int64_t> gather_runs;        // (first, count) CSR ranges that go through the gather (empty: everything)
  int64_t nvar, n_z, n_p, n_g, nnz_full, nnz;
  int64_t nnzD, nnzI, nnzS;
  int64_t g_events, v_events;
  std::vector<PhaseLayout> ph;
  // unique degrees and their table records
  std::vector<int> degs, rec_off;
  std::vector<double> h_tabs;
  int tab_doubles;
  // structure (compact = what the caller sees)
  std::vector<int64_t> rowptr, colind, gather;  // gather: compact -> full index (empty when identical)
  // device
  cudaStream_t stream = nullptr;
  DevBuf d_tabs, d_seg_tab, d_seg_start, d_seg_dpre, d_seg_ipre, d_seg_spre, d_z, d_p, d_sig0, d_g, d_vals, d_full,
      d_grad, d_partial, d_f, d_gather, d_evcols, d_unit_k, d_unit_n;
  // v2 (persistent-warp) launch geometry; v2_warps == 0 -> v1 kernel (one CTA per segment)
  int v2_warps = 0, v2_grid = 0, v2_units = 0, v2_stage_cap = 0, v2_smem_jac = 0, v2_smem_g = 0, v2_nbuf = 1;
  // v4 (row-block warps, persistent images): per-phase grid, images per warp, shared memory
  int v4 = 0, num_sms = 0;
  std::vector<int> v4_grid, v4_threads, v4_smem, v4_smem_g, v4_stage, v4_nbuf;
  int spec_deg = 0;                // uniform degree handed to gjac2 / gjac4 (0: generic instance)
  const void* rt_spec = nullptr;   // RtSpec*: run-time compiled degree-specialised kernels
  std::vector<double> h_p_cache;
  bool p_valid = false;
  // Hessian of the Lagrangian (built on first use): lower triangle, CSR
  bool hess_built = false;
  int ah_zero_fill = -1;  // adaptive Hessian: 1 = the evaluation needs a zero-filled buffer, 0 = every entry has a writer, -1 = not probed yet
  std::vector<int64_t> h_rowptr, h_colind;
  struct HessPhase {
    DevBuf pos_yy, pos_ay, pos_ty, pos_corner, pos_term, term_assign, part;
    MpxHessLin lin;  // affine positions of the interior nodes' entries, passed to the node kernel by value
    DevBuf ah_pos, ah_off;  // adaptive NLP: positions of what a segment adds (mpx_adapt_hess_kernel), [K + 1] offsets
    DevBuf ah_sync;         // [3 + K] queue, epoch and per-segment flags of the persistent launch
    bool ah_persist = true; // one persistent launch (MPX_AHESS_PERSIST=0 at plan creation: one launch per segment parity)
    DevBuf hnl;             // [NRH][N] staging of the node-local block-diagonal entries (MpxPhaseArgs::hnl); empty: off
    int blocks = 0, n_corner = 0;
  };
  std::vector<HessPhase> hess_ph;
  DevBuf d_node_seg, d_lam, d_hvals;
  DevBuf d_hpart2;  // adaptive NLP: per-segment corner partials
  std::vector<int64_t> h_tail_runs[2];  // rows of the small tail kernels (mpx_eval_g_jac_dev_peers)
  std::vector<int64_t> h_runs[3];  // shard plans: (offset, count) runs of g / values / grad_f written by this shard
  int staged = 0;                  // MPX_STAGE_* results currently valid in the device buffers (mpx_stage / mpx_fetch)
  DevBuf d_ccs_perm, d_ccs_vals;   // CCS order of the Jacobian values, built on first use
  // host buffers the caller has registered (mpx_host_register): pinned, so that copies run at full PCIe speed without
  // the staging ring
  struct HostReg { char* base; size_t bytes; bool primed; };
  std::vector<HostReg> regs;
  DevBuf d_dyn_pos, d_dyn_vals;    // positions (CSR order, int32) of the z- / p-dependent Jacobian entries; gathered values
  std::vector<int32_t> h_dyn_pos;
  int64_t n_dyn = -1;
  const double* dyn_primed = nullptr;  // unregistered buffer that holds the constants (last full mpx_eval_jac_g_dynamic)
  // pinned staging ring (MPX_STAGE_SLOTS x MPX_STAGE_BYTES) + worker pool for caller buffers that are pageable
  std::unique_ptr<MpxPool> pool;
  char* h_ring = nullptr;
  cudaEvent_t ring_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  DevBuf d_trace;                  // MPX_TRACE=1: timeline records of the K2 kernel (diagnostics), ring of MPX_TRACE_RING launches
  int64_t trace_seq = 0;
  bool fuse_phases = true;         // MPX_FUSE_PHASES=0: one launch per phase + the events kernel (measurements)
  const MpxProgramEntry* prog = nullptr;
  std::string origin;
  std::vector<MpxPhaseArgs> args;
  int64_t launches = 0;
  int smem_gjac = 0, smem_g = 0, smem_fgrad = 0;
  bool smem_too_big = false;
  ~mpx_plan() {
    for (auto& r : regs) cudaHostUnregister(r.base);
    if (h_ring) cudaFreeHost(h_ring);
    for (auto& e : ring_ev)
      if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
  }
};

static int compute_tables(int scheme, const std::vector<int>& degs, const std::vector<int>& rec_off, int total,
                          double tmin, double tmax, double* d_recs, cudaStream_t st) {
  DevBuf dd, doff;
  CUDA_TRY(dd.ensure(degs.size() * sizeof(int)));
  CUDA_TRY(doff.ensure(degs.size() * sizeof(int)));
  CUDA_TRY(cudaMemcpyAsync(dd.p, degs.data(), degs.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(doff.p, rec_off.data(), degs.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(d_recs, 0, (size_t)total * sizeof(double), st));
  int dmax = *std::max_element(degs.begin(), degs.end());
  size_t smem = (size_t)(2 * (dmax + 1) + 8) * sizeof(double);
  mpx_tables_kernel<<<(int)degs.size(), MPX_TAB_THREADS, smem, st>>>(scheme, dd.as<int>(), doff.as<int>(), tmin, tmax,
                                                                    d_recs);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(st));
  return MPX_OK;
}

extern "C" int mpx_pdl_enabled(void) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MPX_PDL");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v;
}

static int check_degree(int scheme, int deg) {
  if (scheme < MPX_LGR || scheme > MPX_CGL) return fail(MPX_EINVAL, "unknown collocation scheme");
  if (deg < 1 || deg > MPX_MAX_DEG) return fail(MPX_ELIMIT, "polynomial degree must be in [1, 200]");
  return MPX_OK;
}

extern "C" int mpx_collocation_tables(int32_t scheme, int32_t deg, double tau_min, double tau_max, int32_t device,
                                      double* roots, double* D, double* w, double* Cmid) {
  int rc = check_degree(scheme, deg);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(device));
  const int n1 = deg + 1, total = MpxTab::size(n1);
  DevBuf rec;
  CUDA_TRY(rec.ensure((size_t)total * sizeof(double)));
  std::vector<int> degs{deg}, off{0};
  rc = compute_tables(scheme, degs, off, total, tau_min, tau_max, rec.as<double>(), 0);
  if (rc) return rc;
  std::vector<double> h(total);
  CUDA_TRY(cudaMemcpy(h.data(), rec.p, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost));
  if (roots) memcpy(roots, h.data() + MpxTab::off_roots(n1), n1 * sizeof(double));
  if (w) memcpy(w, h.data() + MpxTab::off_w(n1), n1 * sizeof(double));
  if (D) memcpy(D, h.data() + MpxTab::off_D(n1), (size_t)n1 * n1 * sizeof(double));
  if (Cmid) memcpy(Cmid, h.data() + MpxTab::off_C(n1), (size_t)deg * n1 * sizeof(double));
  return MPX_OK;
}

extern "C" int mpx_collocation_basis_at(int32_t scheme, int32_t deg, double tau_min, double tau_max, int32_t device,
                                        int32_t order, int32_t n_taus, const double* taus, double* out) {
  int rc = check_degree(scheme, deg);
  if (rc) return rc;
  if (order < 0 || order > 2 || n_taus < 0 || (n_taus && (!taus || !out))) return fail(MPX_EINVAL, "bad basis_at arguments");
  if (n_taus == 0) return MPX_OK;
  CUDA_TRY(cudaSetDevice(device));
  const int n1 = deg + 1;
  DevBuf dt, dout;
  CUDA_TRY(dt.ensure((size_t)n_taus * sizeof(double)));
  CUDA_TRY(dout.ensure((size_t)n_taus * n1 * sizeof(double)));
  CUDA_TRY(cudaMemcpy(dt.p, taus, (size_t)n_taus * sizeof(double), cudaMemcpyHostToDevice));
  mpx_basis_at_kernel<<<1, MPX_TAB_THREADS, (size_t)(n1 + 2) * sizeof(double)>>>(scheme, deg, tau_min, tau_max, order,
                                                                                n_taus, dt.as<double>(), dout.as<double>());
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(out, dout.p, (size_t)n_taus * n1 * sizeof(double), cudaMemcpyDeviceToHost));
  return MPX_OK;
}

extern "C" int mpx_collocation_weights(int32_t scheme, int32_t deg, double tau_min, double tau_max, int32_t device,
                                       double tau0, double tau1, double* w) {
  int rc = check_degree(scheme, deg);
  if (rc) return rc;
  if (!w) return fail(MPX_EINVAL, "w is NULL");
  CUDA_TRY(cudaSetDevice(device));
  const int n1 = deg + 1;
  DevBuf dw;
  CUDA_TRY(dw.ensure((size_t)n1 * sizeof(double)));
  mpx_weights_kernel<<<1, MPX_TAB_THREADS, (size_t)(2 * n1 + 8) * sizeof(double)>>>(scheme, deg, tau_min, tau_max, tau0,
                                                                                   tau1, dw.as<double>());
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(w, dw.p, (size_t)n1 * sizeof(double), cudaMemcpyDeviceToHost));
  return MPX_OK;
}


// ------------------------------------------------------------------ run-time compiled programs (NVRTC)
// Problems whose generated functors were not compiled ahead of time are compiled here from the SAME kernel header
// (csrc/mpx_kernels.cuh, read from disk next to libmpx.so) plus the generated source handed over the C ABI.
// libnvrtc and libcuda are opened lazily so that libmpx.so itself loads on machines without a driver.
namespace {
typedef int nvrtcResult_t;
typedef struct _nvrtcProgram* nvrtcProgram_t;
typedef int CUresult_t;
typedef struct CUmod_st* CUmodule_t;
typedef struct CUfunc_st* CUfunction_t;

struct RtApi {
  bool ok = false;
  std::string err;
  nvrtcResult_t (*CreateProgram)(nvrtcProgram_t*, const char*, const char*, int, const char* const*, const char* const*);
  nvrtcResult_t (*AddNameExpression)(nvrtcProgram_t, const char*);
  nvrtcResult_t (*CompileProgram)(nvrtcProgram_t, int, const char* const*);
  nvrtcResult_t (*GetProgramLogSize)(nvrtcProgram_t, size_t*);
  nvrtcResult_t (*GetProgramLog)(nvrtcProgram_t, char*);
  nvrtcResult_t (*GetCUBINSize)(nvrtcProgram_t, size_t*);
  nvrtcResult_t (*GetCUBIN)(nvrtcProgram_t, char*);
  nvrtcResult_t (*GetLoweredName)(nvrtcProgram_t, const char*, const char**);
  nvrtcResult_t (*DestroyProgram)(nvrtcProgram_t*);
  CUresult_t (*ModuleLoadData)(CUmodule_t*, const void*);
  CUresult_t (*ModuleGetFunction)(CUfunction_t*, CUmodule_t, const char*);
  CUresult_t (*FuncSetAttribute)(CUfunction_t, int, int);
  CUresult_t (*LaunchKernel)(CUfunction_t, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*,
                             void**, void**);
  CUresult_t (*LaunchKernelEx)(const void* /*CUlaunchConfig*/, CUfunction_t, void**, void**);  // may be null
};

// mirrors of CUlaunchConfig / CUlaunchAttribute (cuda.h, CUDA 12): only the fields used here
struct RtLaunchAttr {
  int id;              // CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION = 6
  char pad[8 - sizeof(int)];
  union {
    char raw[64];
    int programmaticStreamSerializationAllowed;
  } value;
};
struct RtLaunchConfig {
  unsigned gridDimX, gridDimY, gridDimZ, blockDimX, blockDimY, blockDimZ, sharedMemBytes;
  void* hStream;
  RtLaunchAttr* attrs;
  unsigned numAttrs;
};

template <class F>
bool load_sym(void* h, const char* name, F& f, std::string& err) {
  f = reinterpret_cast<F>(dlsym(h, name));
  if (!f) err = std::string("missing symbol ") + name;
  return f != nullptr;
}

RtApi& rt_api() {
  static RtApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* hn = nullptr;
  // the toolkit's own NVRTC first (same release as the nvcc that built the AOT instances), then whatever the loader finds
  for (const char* n : {"/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so"})
    if ((hn = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  void* hc = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!hn || !hc) {
    api.err = !hn ? "libnvrtc.so.12 not found" : "libcuda.so.1 not found";
    return api;
  }
  api.ok = load_sym(hn, "nvrtcCreateProgram", api.CreateProgram, api.err) &&
           load_sym(hn, "nvrtcAddNameExpression", api.AddNameExpression, api.err) &&
           load_sym(hn, "nvrtcCompileProgram", api.CompileProgram, api.err) &&
           load_sym(hn, "nvrtcGetProgramLogSize", api.GetProgramLogSize, api.err) &&
           load_sym(hn, "nvrtcGetProgramLog", api.GetProgramLog, api.err) &&
           load_sym(hn, "nvrtcGetCUBINSize", api.GetCUBINSize, api.err) &&
           load_sym(hn, "nvrtcGetCUBIN", api.GetCUBIN, api.err) &&
           load_sym(hn, "nvrtcGetLoweredName", api.GetLoweredName, api.err) &&
           load_sym(hn, "nvrtcDestroyProgram", api.DestroyProgram, api.err) &&
           load_sym(hc, "cuModuleLoadData", api.ModuleLoadData, api.err) &&
           load_sym(hc, "cuModuleGetFunction", api.ModuleGetFunction, api.err) &&
           load_sym(hc, "cuFuncSetAttribute", api.FuncSetAttribute, api.err) &&
           load_sym(hc, "cuLaunchKernel", api.LaunchKernel, api.err);
  api.LaunchKernelEx = reinterpret_cast<decltype(api.LaunchKernelEx)>(dlsym(hc, "cuLaunchKernelEx"));
  return api;
}

// launches through the driver API; same interface as the AOT phases
struct MpxRtPhase final : MpxPhaseKernels {
  CUfunction_t f_gjac[2] = {nullptr, nullptr}, f_gjac2[2] = {nullptr, nullptr}, f_gjac4[2] = {nullptr, nullptr},
               f_fgrad[2] = {nullptr, nullptr}, f_final[2] = {nullptr, nullptr}, f_resid[2] = {nullptr, nullptr}, f_hess[2] = {nullptr, nullptr},
               f_adapt[3] = {nullptr, nullptr, nullptr},  // [0] unused, [1] adapt_grad, [2] suffix
               f_adaptk[2] = {nullptr, nullptr}, f_ahess[2] = {nullptr, nullptr};
  int nx_ = 0, nu_ = 0, na_ = 0;  // for the shared-memory bound of the adaptive Hessian kernel
  static cudaError_t go(CUfunction_t f, const MpxPhaseArgs& a, int grid, int threads, size_t smem, cudaStream_t st,
                        bool pdl = false, const void* arg2 = nullptr) {
    RtApi& R = rt_api();
    if (smem > 48 * 1024) {  // opt in to large dynamic shared memory once per function, not once per launch
      static std::vector<CUfunction_t> configured;
      if (std::find(configured.begin(), configured.end(), f) == configured.end()) {
        if (R.FuncSetAttribute(f, 8 /*CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES*/, 227 * 1024) != 0)
          return cudaErrorInvalidValue;
        configured.push_back(f);
      }
    }
    void* params[] = {const_cast<MpxPhaseArgs*>(&a), const_cast<void*>(arg2)};
    if (pdl && R.LaunchKernelEx && mpx_pdl_enabled()) {
      RtLaunchAttr at;
      memset(&at, 0, sizeof at);
      at.id = 6;  // CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION
      at.value.programmaticStreamSerializationAllowed = 1;
      RtLaunchConfig cfg = {(unsigned)grid, 1, 1, (unsigned)threads, 1, 1, (unsigned)smem, st, &at, 1};
      return R.LaunchKernelEx(&cfg, f, params, nullptr) == 0 ? cudaSuccess : cudaErrorLaunchFailure;
    }
    return R.LaunchKernel(f, grid, 1, 1, threads, 1, 1, (unsigned)smem, st, params, nullptr) == 0 ? cudaSuccess
                                                                                                : cudaErrorLaunchFailure;
  }
  cudaError_t gjac(const MpxPhaseArgs& a, bool jac, int grid, size_t smem, cudaStream_t st) const override {
    return go(f_gjac[jac], a, grid, MPX_THREADS, smem, st);
  }
  cudaError_t gjac2(const MpxPhaseArgs& a, bool jac, int, int grid, int threads, size_t smem, cudaStream_t st) const override {
    return go(f_gjac2[jac], a, grid, threads, smem, st, true);  // generic instance; degree-specialised ones: RtSpec
  }
  cudaError_t gjac4(const MpxPhaseArgs& a, bool jac, int, int grid, int threads, size_t smem, cudaStream_t st) const override {
    return go(f_gjac4[jac], a, grid, threads, smem, st);  // generic instance; degree-specialised ones: RtSpec
  }
  bool has_degree(int) const override { return false; }
  cudaError_t fgrad(const MpxPhaseArgs& a, bool grad, int grid, size_t smem, cudaStream_t st) const override {
    return go(f_fgrad[grad], a, grid, MPX_THREADS, smem, st);
  }
  cudaError_t fgrad_final(const MpxPhaseArgs& a, bool grad, cudaStream_t st) const override {
    return go(f_final[grad], a, 1, 256, 0, st);
  }
  cudaError_t residual(const MpxPhaseArgs& a, bool deriv, int grid, cudaStream_t st) const override {
    return go(f_resid[deriv], a, grid, 128, 0, st);
  }
  cudaError_t hess(const MpxPhaseArgs& a, const MpxHessLin& hl, int grid, cudaStream_t st) const override {
    cudaError_t e = go(f_hess[0], a, grid, MPX_HESS_THREADS, 0, st, false, &hl);
    return e != cudaSuccess || a.ticket ? e : go(f_hess[1], a, 1, MPX_HESS_FINAL_THREADS, 0, st);
  }
  cudaError_t adapt(const MpxPhaseArgs& a, int grid, size_t smem, cudaStream_t st) const override {
    return go(f_adaptk[a.ad_jac ? 1 : 0], a, grid, MPX_THREADS, smem, st);
  }
  cudaError_t adapt_grad(const MpxPhaseArgs& a, int grid, bool suffix, cudaStream_t st) const override {
    cudaError_t e = go(f_adapt[1], a, grid, MPX_THREADS, 0, st);
    return e != cudaSuccess || !suffix ? e : go(f_adapt[2], a, 1, 32, 0, st);
  }
  cudaError_t adapt_hess(const MpxPhaseArgs& a, int grid, int dmax, cudaStream_t st) const override {
    // the functor's NRG / NRH are not known on the host: bound them by the number of node variables
    const int n1 = dmax + 1, ny = nx_ + nu_, nv = ny + na_, nr = 1 + nv + nv * (nv + 1) / 2;
    const size_t dbl = 2 * (size_t)MpxTab::pad2(dmax * n1) + MpxTab::pad2(n1) + MpxTab::pad2(ny * n1) + MpxTab::pad2(ny * dmax) +
                       MpxTab::pad2(dmax * nr) + MpxTab::pad2(n1 * (1 + nv)) + MpxTab::pad2(nx_ * dmax) + 4 +
                       (size_t)(1 + MPX_THREADS / 32) * n1 * mpx_ahess_stride(dmax) + (size_t)(3 * ny + nv * (nv + 1) / 2) * n1 + na_ + 3 +
                       (size_t)(2 * ny + nv * (nv + 1) / 2) * n1 +  // + the prefetched old values (mpx_adapt_hess_smem_doubles)
                       (size_t)(3 * ny + nv * (nv + 1) / 2) * n1 + na_ + 3 + MpxTab::pad2(ny * n1) + MpxTab::pad2(nx_ * dmax);
    return go(f_ahess[0], a, grid, MPX_THREADS, dbl * sizeof(double), st);
  }
  cudaError_t adapt_hess_final(const MpxPhaseArgs& a, cudaStream_t st) const override { return go(f_ahess[1], a, 1, 64, 0, st); }
};

// run-time compiled all-phases launch: the kernel takes one MpxMultiArgs<P> by value; the host builds its image in a
// byte buffer (P is a run-time number here): a[P] | ev | grid_per_phase | pad
struct MpxRtAll final : MpxProgramKernels {
  CUfunction_t f[2] = {nullptr, nullptr};
  int P = 0;
  cudaError_t gjac2_all(const MpxPhaseArgs* args, int n_phases, const MpxEvArgs& ev, bool jac, int, int grid_per_phase,
                        int threads, size_t smem, cudaStream_t st) const override {
    if (n_phases != P) return cudaErrorInvalidValue;
    RtApi& R = rt_api();
    if (smem > 48 * 1024) {
      static std::vector<CUfunction_t> configured;
      if (std::find(configured.begin(), configured.end(), f[jac]) == configured.end()) {
        if (R.FuncSetAttribute(f[jac], 8 /*CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES*/, 227 * 1024) != 0)
          return cudaErrorInvalidValue;
        configured.push_back(f[jac]);
      }
    }
    std::vector<char> blob((size_t)P * sizeof(MpxPhaseArgs) + sizeof(MpxEvArgs) + 2 * sizeof(int32_t) + 8, 0);
    memcpy(blob.data(), args, (size_t)P * sizeof(MpxPhaseArgs));
    memcpy(blob.data() + (size_t)P * sizeof(MpxPhaseArgs), &ev, sizeof(MpxEvArgs));
    const int32_t gp = grid_per_phase;
    memcpy(blob.data() + (size_t)P * sizeof(MpxPhaseArgs) + sizeof(MpxEvArgs), &gp, sizeof gp);
    void* params[] = {blob.data()};
    if (R.LaunchKernelEx && mpx_pdl_enabled()) {
      RtLaunchAttr at;
      memset(&at, 0, sizeof at);
      at.id = 6;  // CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION
      at.value.programmaticStreamSerializationAllowed = 1;
      RtLaunchConfig cfg = {(unsigned)(grid_per_phase * P), 1, 1, (unsigned)threads, 1, 1, (unsigned)smem, st, &at, 1};
      return R.LaunchKernelEx(&cfg, f[jac], params, nullptr) == 0 ? cudaSuccess : cudaErrorLaunchFailure;
    }
    return R.LaunchKernel(f[jac], grid_per_phase * P, 1, 1, threads, 1, 1, (unsigned)smem, st, params, nullptr) == 0
               ? cudaSuccess
               : cudaErrorLaunchFailure;
  }
};
static_assert(sizeof(MpxMultiArgs<2>) == 2 * sizeof(MpxPhaseArgs) + sizeof(MpxEvArgs) + 2 * sizeof(int32_t) &&
                  offsetof(MpxMultiArgs<3>, ev) == 3 * sizeof(MpxPhaseArgs) &&
                  offsetof(MpxMultiArgs<3>, grid_per_phase) == 3 * sizeof(MpxPhaseArgs) + sizeof(MpxEvArgs),
              "MpxRtAll builds the kernel argument by offset");

struct RtProgram {
  std::string key;
  std::vector<std::unique_ptr<MpxRtPhase>> phases;
  std::vector<const MpxPhaseKernels*> ptrs;
  std::unique_ptr<MpxRtAll> all;
  MpxProgramEntry entry;
};
std::vector<std::unique_ptr<RtProgram>> g_rt_programs;  // kept for the life of the process (modules stay loaded)

std::string lib_dir() {
  Dl_info info;
  if (dladdr(reinterpret_cast<void*>(&mpx_version), &info) && info.dli_fname) {
    std::string p(info.dli_fname);
    size_t k = p.rfind('/');
    return k == std::string::npos ? "." : p.substr(0, k);
  }
  return ".";
}

int compile_program(const char* key, const char* source, int n_phases, int nx, int nu, int na, const MpxProgramEntry** out) {
  RtApi& R = rt_api();
  if (!R.ok) return fail(MPX_ENOPROGRAM, "program '" + std::string(key) + "' is not compiled in and NVRTC is unavailable: " + R.err);
  const std::string hdr_path = lib_dir() + "/csrc/mpx_kernels.cuh";
  FILE* fh = fopen(hdr_path.c_str(), "rb");
  if (!fh) return fail(MPX_ENOPROGRAM, "cannot read " + hdr_path + " for run-time compilation");
  std::string hdr;
  char buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, fh)) > 0) hdr.append(buf, n);
  fclose(fh);
  std::string src = hdr + "\n#define MPX_PHASE_NAME(k) MpxPhRt_##k\n" + source + "\n";
  nvrtcProgram_t prog = nullptr;
  if (R.CreateProgram(&prog, src.c_str(), "mpx_rt.cu", 0, nullptr, nullptr) != 0) return fail(MPX_ECUDA, "nvrtcCreateProgram failed");
  std::vector<std::string> names;
  for (int ph = 0; ph < n_phases; ++ph)
    for (const char* k : {"mpx_gjac_kernel", "mpx_gjac2_kernel", "mpx_gjac4_kernel", "mpx_fgrad_kernel", "mpx_fgrad_final",
                          "mpx_residual_kernel", "mpx_adapt_kernel"})
      for (const char* b : {"false", "true"})
        names.push_back(std::string(k) + "<MpxPhRt_" + std::to_string(ph) + ", " + b +
                        (strcmp(k, "mpx_gjac4_kernel") == 0 || strcmp(k, "mpx_gjac2_kernel") == 0 ? ", 0>" : ">"));
  std::vector<std::string> names1;  // kernels with the phase functor as their only template argument
  for (int ph = 0; ph < n_phases; ++ph)
    for (const char* k : {"mpx_hess_kernel", "mpx_hess_final", "mpx_adapt_grad_kernel", "mpx_adapt_grad_suffix",
                          "mpx_adapt_hess_kernel", "mpx_adapt_hess_final"})
      names1.push_back(std::string(k) + "<MpxPhRt_" + std::to_string(ph) + ">");
  std::vector<std::string> names_all;  // one g + jac_g launch for all phases
  if (n_phases > 1)
    for (const char* b : {"false", "true"}) {
      std::string nm = std::string("mpx_gjac2_multi_kernel<") + b + ", 0";
      for (int ph = 0; ph < n_phases; ++ph) nm += ", MpxPhRt_" + std::to_string(ph);
      names_all.push_back(nm + ">");
    }
  for (auto& nm : names) R.AddNameExpression(prog, nm.c_str());
  for (auto& nm : names1) R.AddNameExpression(prog, nm.c_str());
  for (auto& nm : names_all) R.AddNameExpression(prog, nm.c_str());
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
  const nvrtcResult_t rc = R.CompileProgram(prog, 3, opts);
  if (rc != 0) {
    size_t ls = 0;
    R.GetProgramLogSize(prog, &ls);
    std::string log(ls, '\0');
    if (ls) R.GetProgramLog(prog, &log[0]);
    R.DestroyProgram(&prog);
    return fail(MPX_ECUDA, "NVRTC compilation of the node functors failed:\n" + log.substr(0, 4000));
  }
  size_t cs = 0;
  R.GetCUBINSize(prog, &cs);
  std::vector<char> cubin(cs);
  R.GetCUBIN(prog, cubin.data());
  cudaFree(0);  // make sure the primary context is current for the driver calls below
  CUmodule_t mod = nullptr;
  if (R.ModuleLoadData(&mod, cubin.data()) != 0) {
    R.DestroyProgram(&prog);
    return fail(MPX_ECUDA, "cuModuleLoadData failed for the run-time compiled program");
  }
  std::unique_ptr<RtProgram> rp(new RtProgram());
  rp->key = key;
  size_t idx = 0;
  for (int ph = 0; ph < n_phases; ++ph) {
    std::unique_ptr<MpxRtPhase> P(new MpxRtPhase());
    P->nx_ = nx, P->nu_ = nu, P->na_ = na;
    CUfunction_t* slots[7] = {P->f_gjac, P->f_gjac2, P->f_gjac4, P->f_fgrad, P->f_final, P->f_resid, P->f_adaptk};
    for (int k = 0; k < 7; ++k)
      for (int b = 0; b < 2; ++b, ++idx) {
        const char* lowered = nullptr;
        if (R.GetLoweredName(prog, names[idx].c_str(), &lowered) != 0 || !lowered ||
            R.ModuleGetFunction(&slots[k][b], mod, lowered) != 0) {
          R.DestroyProgram(&prog);
          return fail(MPX_ECUDA, "kernel " + names[idx] + " not found in the run-time compiled module");
        }
      }
    for (int b = 0; b < 6; ++b) {
      const char* lowered = nullptr;
      const std::string& nm = names1[(size_t)ph * 6 + b];
      CUfunction_t* slot = b < 2 ? &P->f_hess[b] : (b < 4 ? &P->f_adapt[b - 1] : &P->f_ahess[b - 4]);
      if (R.GetLoweredName(prog, nm.c_str(), &lowered) != 0 || !lowered || R.ModuleGetFunction(slot, mod, lowered) != 0) {
        R.DestroyProgram(&prog);
        return fail(MPX_ECUDA, "kernel " + nm + " not found in the run-time compiled module");
      }
    }
    rp->ptrs.push_back(P.get());
    rp->phases.push_back(std::move(P));
  }
  if (n_phases > 1) {
    rp->all.reset(new MpxRtAll());
    rp->all->P = n_phases;
    for (int b = 0; b < 2; ++b) {
      const char* lowered = nullptr;
      if (R.GetLoweredName(prog, names_all[b].c_str(), &lowered) != 0 || !lowered ||
          R.ModuleGetFunction(&rp->all->f[b], mod, lowered) != 0) {
        R.DestroyProgram(&prog);
        return fail(MPX_ECUDA, "kernel " + names_all[b] + " not found in the run-time compiled module");
      }
    }
  }
  R.DestroyProgram(&prog);
  rp->entry = MpxProgramEntry{rp->key.c_str(), n_phases, rp->ptrs.data(), nullptr, rp->all.get()};
  *out = &rp->entry;
  g_rt_programs.push_back(std::move(rp));
  return MPX_OK;
}

// ---- degree-specialised g+jac kernels (mpx_gjac2_kernel / mpx_gjac4_kernel<PH, JAC, DEG>) compiled on demand
struct RtSpec {
  std::string key;  // "<program key>:<kernel>:d<deg>"
  std::vector<std::array<CUfunction_t, 2>> f;  // per phase: [jac]
};
std::vector<std::unique_ptr<RtSpec>> g_rt_specs;

const RtSpec* find_rt_spec(const std::string& key) {
  for (auto& sp : g_rt_specs)
    if (sp->key == key) return sp.get();
  return nullptr;
}

int read_kernel_header(std::string& hdr) {
  const std::string hdr_path = lib_dir() + "/csrc/mpx_kernels.cuh";
  FILE* fh = fopen(hdr_path.c_str(), "rb");
  if (!fh) return fail(MPX_ENOPROGRAM, "cannot read " + hdr_path + " for run-time compilation");
  char buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, fh)) > 0) hdr.append(buf, n);
  fclose(fh);
  return MPX_OK;
}

int compile_spec(const char* key, const char* source, int n_phases, int deg, const char* kernel, const RtSpec** out) {
  RtApi& R = rt_api();
  if (!R.ok) return fail(MPX_ENOPROGRAM, "NVRTC is unavailable: " + R.err);
  std::string hdr;
  int rc0 = read_kernel_header(hdr);
  if (rc0) return rc0;
  std::string src = hdr + "\n#define MPX_PHASE_NAME(k) MpxPhRt_##k\n" + source + "\n";
  nvrtcProgram_t prog = nullptr;
  if (R.CreateProgram(&prog, src.c_str(), "mpx_rt_spec.cu", 0, nullptr, nullptr) != 0) return fail(MPX_ECUDA, "nvrtcCreateProgram failed");
  std::vector<std::string> names;
  for (int ph = 0; ph < n_phases; ++ph)
    for (const char* b : {"false", "true"})
      names.push_back(std::string(kernel) + "<MpxPhRt_" + std::to_string(ph) + ", " + b + ", " + std::to_string(deg) + ">");
  for (auto& nm : names) R.AddNameExpression(prog, nm.c_str());
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
  if (R.CompileProgram(prog, 3, opts) != 0) {
    size_t ls = 0;
    R.GetProgramLogSize(prog, &ls);
    std::string log(ls, '\0');
    if (ls) R.GetProgramLog(prog, &log[0]);
    R.DestroyProgram(&prog);
    return fail(MPX_ECUDA, "NVRTC compilation of the degree-specialised kernel failed:\n" + log.substr(0, 4000));
  }
  size_t cs = 0;
  R.GetCUBINSize(prog, &cs);
  std::vector<char> cubin(cs);
  R.GetCUBIN(prog, cubin.data());
  cudaFree(0);
  CUmodule_t mod = nullptr;
  if (R.ModuleLoadData(&mod, cubin.data()) != 0) {
    R.DestroyProgram(&prog);
    return fail(MPX_ECUDA, "cuModuleLoadData failed for the degree-specialised kernel");
  }
  std::unique_ptr<RtSpec> sp(new RtSpec());
  sp->key = std::string(key) + ":" + kernel + ":d" + std::to_string(deg);
  sp->f.resize(n_phases);
  size_t idx = 0;
  for (int ph = 0; ph < n_phases; ++ph)
    for (int b = 0; b < 2; ++b, ++idx) {
      const char* lowered = nullptr;
      if (R.GetLoweredName(prog, names[idx].c_str(), &lowered) != 0 || !lowered ||
          R.ModuleGetFunction(&sp->f[ph][b], mod, lowered) != 0) {
        R.DestroyProgram(&prog);
        return fail(MPX_ECUDA, "kernel " + names[idx] + " not found in the run-time compiled module");
      }
    }
  R.DestroyProgram(&prog);
  *out = sp.get();
  g_rt_specs.push_back(std::move(sp));
  return MPX_OK;
}

const MpxProgramEntry* find_rt_program(const char* key) {
  for (auto& rp : g_rt_programs)
    if (rp->key == key) return &rp->entry;
  return nullptr;
}
}  // namespace

// ------------------------------------------------------------------ structure
namespace {
struct Builder {
  mpx_plan& P;
  std::vector<int64_t> rowptr, colind;
  std::vector<uint8_t> keep;
  std::vector<int64_t> src;  // position of the entry in the staged value buffer the kernels write
  int64_t n_base = 0;        // entries of the base kernels so far (their write order = the order of add())
  explicit Builder(mpx_plan& p) : P(p) { rowptr.push_back(0); }
  void add(int64_t col, bool k = true) {
    colind.push_back(col);
    keep.push_back(k ? 1 : 0);
    src.push_back(n_base++);
  }
  void add_ext(int64_t col, int64_t ext_pos, bool k = true) {  // entry written by mpx_adapt_kernel
    colind.push_back(col);
    keep.push_back(k ? 1 : 0);
    src.push_back(-1 - ext_pos);
  }
  void end_row() { rowptr.push_back((int64_t)colind.size()); }
};
}  // namespace

static const double* tab_of(const mpx_plan& p, int deg) {
  for (size_t i = 0; i < p.degs.size(); ++i)
    if (p.degs[i] == deg) return p.h_tabs.data() + p.rec_off[i];
  return nullptr;
}

static void build_structure(mpx_plan& p) {
  Builder B(p);
  const int nx = p.nx, nu = p.nu, na = p.na, N = p.N, K = p.K;
  const int nv = nx + nu + na;
  // node -> (owner segment, local index)
  std::vector<int> nseg(N), nloc(N);
  for (int k = 0; k < K; ++k)
    for (int r = (k == 0 ? 0 : 1); r <= p.po[k]; ++r) nseg[p.seg_start[k] + r] = k, nloc[p.seg_start[k] + r] = r;
  for (int ph = 0; ph < p.P; ++ph) {
    const PhaseLayout& L = p.ph[ph];
    const int64_t zo = L.zoff;
    auto colX = [&](int i, int s) { return zo + (int64_t)s * N + i; };
    auto colU = [&](int i, int c) { return zo + (int64_t)(nx + c) * N + i; };
    const int64_t cT0 = zo + (int64_t)(nx + nu) * N, cTF = cT0 + 1;
    auto colA = [&](int m) { return cT0 + 2 + m; };
    auto colW = [&](int m) { return cT0 + 2 + na + m; };  // adaptive NLP only (mpopt.py:2938-2945)
    // F rows
    for (int s = 0; s < nx; ++s) {
      const uint8_t* pat = L.pat_f.data() + (size_t)s * nv;
      for (int i = 0; i < N; ++i) {
        const int k = nseg[i], r = nloc[i], d = p.po[k], n1 = d + 1;
        const double* D = tab_of(p, d) + MpxTab::off_D(n1);
        for (int sp = 0; sp < s; ++sp)
          if (pat[sp]) B.add(colX(i, sp));
        for (int j = 0; j <= d; ++j) {
          bool keep = true;
          if (p.drop && D[r * n1 + j] == 0.0 && !(j == r && pat[s])) keep = false;
          B.add(colX(p.seg_start[k] + j, s), keep);
        }
        for (int sp = s + 1; sp < nx; ++sp)
          if (pat[sp]) B.add(colX(i, sp));
        for (int c = 0; c < nu; ++c)
          if (pat[nx + c]) B.add(colU(i, c));
        if (L.f_nz[s]) B.add(cT0), B.add(cTF);
        for (int m = 0; m < na; ++m)
          if (pat[nx + nu + m]) B.add(colA(m));
        if (p.adaptive && L.f_nz[s] && p.wcol) {
          B.add(colW(k));  // written by the base kernel, in place
        } else if (p.adaptive && L.f_nz[s]) {  // h_k = (tf - t0)/delta * w_k; t_i also moves with every earlier width
          if (L.f_t[s])
            for (int m = 0; m <= k; ++m) B.add_ext(colW(m), L.eF[s] + (int64_t)i * K + m);
          else
            B.add_ext(colW(k), L.eF[s] + i);
        }
        B.end_row();
      }
    }
    // path rows
    for (int q = 0; q < L.nc; ++q) {
      const uint8_t* pat = L.pat_c.data() + (size_t)q * nv;
      for (int i = 0; i < N; ++i) {
        for (int s = 0; s < nx; ++s)
          if (pat[s]) B.add(colX(i, s));
        for (int c = 0; c < nu; ++c)
          if (pat[nx + c]) B.add(colU(i, c));
        if (L.c_t[q]) B.add(cT0), B.add(cTF, i != 0);  // node 0: t = t0 + h*0.0 folds to t0 (mpopt.py:198)
        for (int m = 0; m < na; ++m)
          if (pat[nx + nu + m]) B.add(colA(m));
        if (p.adaptive && L.c_t[q] && i != 0)
          for (int m = 0; m <= nseg[i]; ++m) B.add_ext(colW(m), L.eC[q] + (int64_t)i * K + m);
        B.end_row();
      }
    }
    // control slope rows
    if (L.has_DU)
      for (int c = 0; c < nu; ++c)
        for (int i = 0; i < N; ++i) {
          const int k = nseg[i], r = nloc[i], d = p.po[k], n1 = d + 1;
          const double* D = tab_of(p, d) + MpxTab::off_D(n1);
          for (int j = 0; j <= d; ++j) B.add(colU(p.seg_start[k] + j, c), !(p.drop && D[r * n1 + j] == 0.0));
          B.end_row();
        }
    // mid-point rows
    if (L.has_mU)
      for (int c = 0; c < nu; ++c)
        for (int k = 0; k < K; ++k) {
          const int d = p.po[k], n1 = d + 1;
          const double* C = tab_of(p, d) + MpxTab::off_C(n1);
          for (int m = 0; m < d; ++m) {
            for (int j = 0; j <= d; ++j) B.add(colU(p.seg_start[k] + j, c), !(p.drop && C[m * n1 + j] == 0.0));
            B.end_row();
          }
        }
    // slope continuity rows
    if (L.has_dU)
      for (int c = 0; c < nu; ++c)
        for (int k = 0; k + 1 < K; ++k) {
          const int da = p.po[k], db = p.po[k + 1], na1 = da + 1, nb1 = db + 1;
          const double* Da = tab_of(p, da) + MpxTab::off_D(na1) + da * na1;
          const double* Db = tab_of(p, db) + MpxTab::off_D(nb1);
          for (int j = 0; j < da; ++j) B.add(colU(p.seg_start[k] + j, c), !(p.drop && Da[j] == 0.0));
          B.add(colU(p.seg_start[k + 1], c), !(p.drop && Da[da] - Db[0] == 0.0));
          for (int j = 1; j <= db; ++j) B.add(colU(p.seg_start[k + 1] + j, c), !(p.drop && Db[j] == 0.0));
          B.end_row();
        }
    // terminal rows: x0_s (col s*N) then xf_s (col s*N+N-1) per state, T0, TF, a
    const int ntv = 2 * nx + 2 + na;
    for (int r = 0; r < L.ntc; ++r) {
      const uint8_t* pat = L.pat_tc.data() + (size_t)r * ntv;
      for (int s = 0; s < nx; ++s) {
        if (pat[nx + s]) B.add(colX(0, s));
        if (pat[s]) B.add(colX(N - 1, s));
      }
      if (pat[2 * nx + 1]) B.add(cT0);
      if (pat[2 * nx]) B.add(cTF);
      for (int m = 0; m < na; ++m)
        if (pat[2 * nx + 2 + m]) B.add(colA(m));
      B.end_row();
    }
    // SW block of the adaptive NLP (mpopt.py:3034-3136)
    if (p.adaptive) {
      for (int m = 0; m < K; ++m) B.add_ext(colW(m), L.eSum + m);
      B.end_row();
      for (int blk = 0; blk < 2; ++blk) {  // compI.U then compI.X
        if (!(blk == 0 ? L.sw_u : L.sw_x)) continue;
        const int nvv = blk == 0 ? nu : nx;
        for (int c = 0; c < nvv; ++c) {
          int64_t pos = (blk == 0 ? L.eUi : L.eXi) + (int64_t)c * p.nnzI;
          for (int k = 0; k < K; ++k) {
            const int d = p.po[k], n1 = d + 1;
            const double* C = tab_of(p, d) + MpxTab::off_C(n1);
            for (int m = 0; m < d; ++m) {
              for (int j = 0; j <= d; ++j, ++pos) {
                const int64_t col = blk == 0 ? colU(p.seg_start[k] + j, c) : colX(p.seg_start[k] + j, c);
                B.add_ext(col, pos, !(p.drop && C[m * n1 + j] == 0.0));
              }
              B.end_row();
            }
          }
        }
      }
      if (p.mid_res) {
        for (int k = 0; k < K; ++k) {
          const int d = p.po[k], n1 = d + 1;
          const double* DI = p.h_dmid.data() + p.dmid_off[std::lower_bound(p.degs.begin(), p.degs.end(), d) - p.degs.begin()];
          int64_t pos = L.eRes + L.seg_rpre[k];
          for (int s = 0; s < nx; ++s) {
            const uint8_t* pat = L.pat_f.data() + (size_t)s * nv;
            for (int m = 0; m < d; ++m) {
              for (int v = 0; v < nx + nu; ++v) {
                if (!(v == s || pat[v])) continue;
                for (int j = 0; j <= d; ++j, ++pos) {
                  const int64_t col = v < nx ? colX(p.seg_start[k] + j, v) : colU(p.seg_start[k] + j, v - nx);
                  B.add_ext(col, pos, !(p.drop && v == s && !pat[s] && DI[m * n1 + j] == 0.0));
                }
              }
              if (L.f_nz[s]) B.add_ext(cT0, pos++), B.add_ext(cTF, pos++);
              for (int m2 = 0; m2 < na; ++m2)
                if (pat[nx + nu + m2]) B.add_ext(colA(m2), pos++);
              if (L.f_t[s]) {
                for (int m2 = 0; m2 < K; ++m2, ++pos)
                  if (m2 <= k) B.add_ext(colW(m2), pos);
              } else {
                B.add_ext(colW(k), pos++);
              }
              B.end_row();
            }
          }
        }
      }
    }
  }
  // event rows: state block, control block, time block (mpopt.py:484-519)
  const int nl = (int)p.links.size() / 2;
  std::vector<int64_t> ca, cb;
  auto ev = [&](int64_t a, int64_t b) {
    ca.push_back(a), cb.push_back(b);
    B.add(std::min(a, b)), B.add(std::max(a, b));
    B.end_row();
  };
  for (int l = 0; l < nl; ++l)
    for (int s = 0; s < nx; ++s)
      ev(p.ph[p.links[2 * l + 1]].zoff + (int64_t)s * N, p.ph[p.links[2 * l]].zoff + (int64_t)s * N + N - 1);
  for (int l = 0; l < nl; ++l)
    for (int c = 0; c < nu; ++c)
      ev(p.ph[p.links[2 * l + 1]].zoff + (int64_t)(nx + c) * N, p.ph[p.links[2 * l]].zoff + (int64_t)(nx + c) * N + N - 1);
  for (int l = 0; l < nl; ++l)
    ev(p.ph[p.links[2 * l + 1]].zoff + (int64_t)(nx + nu) * N, p.ph[p.links[2 * l]].zoff + (int64_t)(nx + nu) * N + 1);
  // event columns for the device kernel
  if (nl) {
    std::vector<int64_t> both(ca);
    both.insert(both.end(), cb.begin(), cb.end());
    p.d_evcols.ensure(both.size() * sizeof(int64_t));
    cudaMemcpy(p.d_evcols.p, both.data(), both.size() * sizeof(int64_t), cudaMemcpyHostToDevice);
  }
  // compact
  p.n_base = B.n_base;
  p.nnz_full = p.n_base + p.n_ext;
  for (int64_t& v : B.src)
    if (v < 0) v = p.n_base + (-1 - v);  // ext section follows the base kernels' entries
  bool all = !p.adaptive;
  for (uint8_t k : B.keep)
    if (!k) {
      all = false;
      break;
    }
  if (all) {
    p.rowptr.swap(B.rowptr);
    p.colind.swap(B.colind);
    p.gather.clear();
  } else {
    p.rowptr.assign(1, 0);
    p.colind.clear();
    p.gather.clear();
    for (size_t r = 0; r + 1 < B.rowptr.size(); ++r) {
      for (int64_t e = B.rowptr[r]; e < B.rowptr[r + 1]; ++e)
        if (B.keep[e]) p.colind.push_back(B.colind[e]), p.gather.push_back(B.src[e]);
      p.rowptr.push_back((int64_t)p.colind.size());
    }
  }
  p.nnz = (int64_t)p.colind.size();
  // adaptive NLP: an SW block whose CSR entries are exactly its ext entries in order is written in place
  p.sw_direct.assign(p.P, -1), p.gather_runs.clear();
  if (p.adaptive) {
    int64_t done = 0;  // CSR entries before this point are covered by gather runs or in-place blocks
    for (int ph = 0; ph < p.P; ++ph) {
      const PhaseLayout& L = p.ph[ph];
      const int64_t rb = L.gSW, re = ph + 1 < p.P ? p.ph[ph + 1].gF : p.g_events;
      const int64_t a = p.rowptr[rb], b = p.rowptr[re];
      bool direct = b > a && p.gather[a] == p.n_base + L.eSum;
      for (int64_t e = a + 1; e < b && direct; ++e) direct = p.gather[e] == p.gather[e - 1] + 1;
      if (direct) {
        p.sw_direct[ph] = a;
        if (a > done) p.gather_runs.push_back(done), p.gather_runs.push_back(a - done);
        done = b;
      }
    }
    if (p.nnz > done) p.gather_runs.push_back(done), p.gather_runs.push_back(p.nnz - done);
    bool any = false, all_sw = true;
    for (int64_t v : p.sw_direct) any |= v >= 0, all_sw = all_sw && v >= 0;
    if (!any) p.gather_runs.clear();
    // every entry outside the in-place SW blocks already sits at its CSR position (single phase, width column written
    // by the base kernel, nothing folded away): the base kernels write straight into the caller's array
    p.base_direct = all_sw && any;
    for (size_t i = 0; p.base_direct && i + 1 < p.gather_runs.size(); i += 2)
      for (int64_t e = p.gather_runs[i]; e < p.gather_runs[i] + p.gather_runs[i + 1]; ++e)
        if (p.gather[e] != e) {
          p.base_direct = false;
          break;
        }
    if (p.base_direct) p.gather_runs.clear();
  }
}

// ------------------------------------------------------------------ plan creation
static void copy_pat(std::vector<uint8_t>& dst, const uint8_t* src, size_t n) {
  dst.assign(n, 0);
  if (src)
    for (size_t i = 0; i < n; ++i) dst[i] = src[i] ? 1 : 0;
}

extern "C" int mpx_plan_create(const mpx_problem_desc* d, mpx_plan** out) {
  if (!d || !out) return fail(MPX_EINVAL, "NULL argument");
  *out = nullptr;
  if (d->nx < 0 || d->nu < 0 || d->na < 0 || d->nx > MPX_MAXS || d->nu > MPX_MAXS || d->na > MPX_MAXS)
    return fail(MPX_ELIMIT, "nx, nu, na must be in [0, 16]");
  if (d->n_phases < 1 || !d->phases) return fail(MPX_EINVAL, "n_phases must be >= 1");
  if (d->n_segments < 1 || !d->poly_orders) return fail(MPX_EINVAL, "n_segments must be >= 1");
  if (d->scheme < MPX_LGR || d->scheme > MPX_CGL) return fail(MPX_EINVAL, "unknown collocation scheme");
  if (!(d->tau_max > d->tau_min)) return fail(MPX_EINVAL, "tau_max must exceed tau_min");
  for (int k = 0; k < d->n_segments; ++k)
    if (d->poly_orders[k] < 1 || d->poly_orders[k] > MPX_MAX_DEG)
      return fail(MPX_ELIMIT, "every poly_order must be in [1, 200]");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (ndev <= 0 || d->device < 0 || d->device >= ndev) return fail(MPX_ENODEVICE, "no such CUDA device");
  CUDA_TRY(cudaSetDevice(d->device));

  std::unique_ptr<mpx_plan> pp(new mpx_plan());
  mpx_plan& p = *pp;
  p.nx = d->nx, p.nu = d->nu, p.na = d->na, p.P = d->n_phases, p.K = d->n_segments, p.scheme = d->scheme;
  p.device = d->device, p.tau_min = d->tau_min, p.tau_max = d->tau_max, p.st = d->scale_t;
  p.drop = d->drop_exact_zeros != 0;
  p.po.assign(d->poly_orders, d->poly_orders + p.K);
  p.seg_start.assign(p.K + 1, 0);
  for (int k = 0; k < p.K; ++k) p.seg_start[k + 1] = p.seg_start[k] + p.po[k];
  p.N = p.seg_start[p.K] + 1;
  p.uniform = std::all_of(p.po.begin(), p.po.end(), [&](int v) { return v == p.po[0]; });
  p.sx.assign(p.nx, 1.0), p.su.assign(p.nu, 1.0), p.sa.assign(p.na, 1.0);
  for (int i = 0; i < p.nx && d->scale_x; ++i) p.sx[i] = d->scale_x[i];
  for (int i = 0; i < p.nu && d->scale_u; ++i) p.su[i] = d->scale_u[i];
  for (int i = 0; i < p.na && d->scale_a; ++i) p.sa[i] = d->scale_a[i];
  if (p.st == 0.0) return fail(MPX_EINVAL, "scale_t must be non-zero");
  p.seg_begin = d->seg_begin, p.seg_end = d->seg_end;
  if (p.seg_begin == 0 && p.seg_end == 0) p.seg_end = p.K;
  p.adaptive = d->adaptive != 0, p.mid_res = p.adaptive && d->mid_residuals != 0;
  if (p.adaptive && !(p.seg_begin == 0 && p.seg_end == p.K))
    return fail(MPX_EINVAL, "the adaptive NLP (widths as variables) cannot be sharded: its rows couple all segments");
  if (p.seg_begin < 0 || p.seg_end > p.K || p.seg_begin >= p.seg_end) return fail(MPX_EINVAL, "bad segment shard");
  if (d->n_links < 0 || (d->n_links && !d->links)) return fail(MPX_EINVAL, "bad phase links");
  p.links.assign(d->links, d->links + 2 * d->n_links);
  for (int v : p.links)
    if (v < 0 || v >= p.P) return fail(MPX_EINVAL, "phase link out of range");

  // ---- program
  p.prog = mpx_find_program(d->program_key);
  p.origin = "aot:";
  if (!p.prog && d->program_key) {
    p.origin = "nvrtc:";
    p.prog = find_rt_program(d->program_key);
    if (!p.prog) {
      if (!d->program_source)
        return fail(MPX_ENOPROGRAM, std::string("no compiled node functors registered for program key '") +
                                        d->program_key + "' and no program_source given");
      int rc_ = compile_program(d->program_key, d->program_source, p.P, p.nx, p.nu, p.na, &p.prog);
      if (rc_) return rc_;
    }
  }
  if (!p.prog) return fail(MPX_ENOPROGRAM, "program_key is NULL");
  if (p.prog->n_phases != p.P) return fail(MPX_EINVAL, "program/phase count mismatch");
  p.origin += p.prog->key;
  const bool rt_prog = p.origin[0] == 'n';

  // ---- sizes and offsets
  const int nx = p.nx, nu = p.nu, na = p.na, N = p.N, K = p.K, nv = nx + nu + na;
  p.nvar = (int64_t)N * (nx + nu) + 2 + na + (p.adaptive ? K : 0);  // adaptive: widths appended (mpopt.py:2938-2945)
  p.n_z = p.nvar * p.P;
  p.n_p = p.adaptive ? 0 : (int64_t)K * p.P;                        // :3190-3191
  p.nnzD = (int64_t)(p.po[0] + 1) * (p.po[0] + 1);
  p.nnzI = 0, p.nnzS = 0;
  for (int k = 0; k < K; ++k) {
    if (k) p.nnzD += (int64_t)p.po[k] * (p.po[k] + 1);
    p.nnzI += (int64_t)p.po[k] * (p.po[k] + 1);
    if (k + 1 < K) p.nnzS += p.po[k] + p.po[k + 1] + 1;
  }
  p.ph.resize(p.P);
  // widths-as-variables NLP: when no dynamics row depends on t explicitly, every F row gains exactly ONE width column
  // (its own segment's w_k, last in the row), which the persistent-warp kernel writes itself -- the rows then need no
  // gather pass.  (With explicit time dependence the row is dense in all earlier widths: mpx_adapt_kernel + gather.)
  if (p.adaptive) {
    const char* fk = getenv("MPX_KERNEL");
    const char* we = getenv("MPX_ADAPT_WCOL");
    bool ok = !(fk && (strcmp(fk, "v1") == 0 || strcmp(fk, "v4") == 0)) && !(we && atoi(we) == 0);
    for (int k = 0; k < p.K; ++k) ok = ok && p.po[k] <= 31;
    for (int ph = 0; ph < p.P && ok; ++ph)
      for (int s = 0; s < nx && d->phases[ph].f_t; ++s) ok = ok && !d->phases[ph].f_t[s];
    p.wcol = ok;
  }
  int64_t row = 0, val = 0;
  for (int ph = 0; ph < p.P; ++ph) {
    const mpx_phase_desc& q = d->phases[ph];
    PhaseLayout& L = p.ph[ph];
    if (q.n_path < 0 || q.n_path > MPX_MAXS || q.n_term < 0) return fail(MPX_ELIMIT, "n_path must be in [0,16]");
    L.nc = q.n_path, L.ntc = q.n_term;
    copy_pat(L.pat_f, q.pat_f, (size_t)nx * nv);
    copy_pat(L.f_nz, q.f_nz, nx);
    copy_pat(L.f_t, q.f_t, nx);
    copy_pat(L.pat_c, q.pat_c, (size_t)L.nc * nv);
    copy_pat(L.c_t, q.c_t, L.nc);
    copy_pat(L.pat_tc, q.pat_tc, (size_t)L.ntc * (2 * nx + 2 + na));
    L.has_hess = q.pat_hw != nullptr && q.pat_ht != nullptr;
    copy_pat(L.pat_hw, q.pat_hw, (size_t)(nv + 2) * (nv + 2));
    copy_pat(L.pat_ht, q.pat_ht, (size_t)(2 * nx + 2 + na) * (2 * nx + 2 + na));
    copy_pat(L.pat_hf, q.pat_hf, (size_t)nv * nv);
    L.phi_nz = q.phi_nz != 0;
    L.has_DU = q.diff_u != 0, L.has_mU = q.midu != 0, L.has_dU = q.du_continuity != 0 && K > 1;
    if (p.adaptive) L.has_mU = L.has_dU = false, L.sw_u = q.sw_u != 0, L.sw_x = q.sw_x != 0;  // rows [F C DU TC SW], :3169
    L.zoff = p.nvar * ph;
    L.cost_t = q.cost_t != 0;
    L.f_next.assign(nx, 0), L.c_len.assign(L.nc, 0), L.tc_len.assign(L.ntc, 0);
    for (int s = 0; s < nx; ++s) {
      int n = 2 * L.f_nz[s];
      for (int v = 0; v < nv; ++v)
        if (v != s && L.pat_f[(size_t)s * nv + v]) ++n;
      L.f_next[s] = n + ((p.wcol && L.f_nz[s]) ? 1 : 0);  // + the w_k column of the widths-as-variables NLP
      L.uses_t |= L.f_t[s] != 0;
    }
    for (int c = 0; c < L.nc; ++c) {
      int n = 2 * L.c_t[c];
      for (int v = 0; v < nv; ++v) n += L.pat_c[(size_t)c * nv + v];
      L.c_len[c] = n;
      L.uses_t |= L.c_t[c] != 0;
    }
    for (int r = 0; r < L.ntc; ++r)
      for (int v = 0; v < 2 * nx + 2 + na; ++v) L.tc_len[r] += L.pat_tc[(size_t)r * (2 * nx + 2 + na) + v];
    L.gF = row, row += (int64_t)nx * N;
    L.gC = row, row += (int64_t)L.nc * N;
    L.gDU = row, row += L.has_DU ? (int64_t)nu * N : 0;
    L.gmU = row, row += L.has_mU ? (int64_t)nu * (N - 1) : 0;
    L.gdU = row, row += L.has_dU ? (int64_t)nu * (K - 1) : 0;
    L.gTC = row, row += L.ntc;
    if (p.adaptive) {
      L.gSW = row;
      row += 1 + (L.sw_u ? (int64_t)nu * (N - 1) : 0) + (L.sw_x ? (int64_t)nx * (N - 1) : 0) +
             (p.mid_res ? (int64_t)nx * (N - 1) : 0);
      // ext section: what mpx_adapt_kernel writes, in its own regular layout
      int64_t e = p.n_ext;
      for (int s = 0; s < nx; ++s) {
        L.eF[s] = (L.f_nz[s] && !p.wcol) ? e : -1;
        if (L.f_nz[s] && !p.wcol) e += L.f_t[s] ? (int64_t)N * K : N;
      }
      for (int c = 0; c < L.nc; ++c) {
        L.eC[c] = L.c_t[c] ? e : -1;
        if (L.c_t[c]) e += (int64_t)N * K;
      }
      L.eSum = e, e += K;
      L.eUi = e, e += L.sw_u ? (int64_t)nu * p.nnzI : 0;
      L.eXi = e, e += L.sw_x ? (int64_t)nx * p.nnzI : 0;
      L.eRes = e;
      L.res_nblk.assign(nx, 0), L.res_na.assign(nx, 0), L.seg_rpre.assign(K, 0);
      for (int s = 0; s < nx; ++s) {
        for (int v = 0; v < nx + nu; ++v) L.res_nblk[s] += (v == s || L.pat_f[(size_t)s * nv + v]) ? 1 : 0;
        for (int m = 0; m < na; ++m) L.res_na[s] += L.pat_f[(size_t)s * nv + nx + nu + m] ? 1 : 0;
      }
      if (p.mid_res) {
        int64_t acc = 0;
        for (int k = 0; k < K; ++k) {
          L.seg_rpre[k] = acc;
          for (int s = 0; s < nx; ++s)
            acc += (int64_t)p.po[k] * ((int64_t)(p.po[k] + 1) * L.res_nblk[s] + 2 * L.f_nz[s] + L.res_na[s] + (L.f_t[s] ? K : 1));
        }
        e += acc;
      }
      p.n_ext = e;
    }
    for (int s = 0; s < nx; ++s) L.vF[s] = val, val += p.nnzD + (int64_t)N * L.f_next[s];
    for (int c = 0; c < L.nc; ++c) L.vC[c] = val, val += (int64_t)N * L.c_len[c];
    L.vDU = val, val += L.has_DU ? (int64_t)nu * p.nnzD : 0;
    L.vmU = val, val += L.has_mU ? (int64_t)nu * p.nnzI : 0;
    L.vdU = val, val += L.has_dU ? (int64_t)nu * p.nnzS : 0;
    L.vTC = val;
    for (int r = 0; r < L.ntc; ++r) val += L.tc_len[r];
  }
  p.g_events = row, p.v_events = val;
  const int nl = d->n_links;
  p.n_g = row + (int64_t)nl * (nx + nu + 1);

  // ---- tables on the device (K0/K1), copied back once for the exact-zero mask
  CUDA_TRY(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
  p.degs = p.po;
  std::sort(p.degs.begin(), p.degs.end());
  p.degs.erase(std::unique(p.degs.begin(), p.degs.end()), p.degs.end());
  p.rec_off.clear();
  int total = 0;
  for (int dg : p.degs) p.rec_off.push_back(total), total += MpxTab::size(dg + 1);
  p.tab_doubles = total;
  CUDA_TRY(p.d_tabs.ensure((size_t)total * sizeof(double)));
  int rc = compute_tables(p.scheme, p.degs, p.rec_off, total, p.tau_min, p.tau_max, p.d_tabs.as<double>(), p.stream);
  if (rc) return rc;
  p.h_tabs.resize(total);
  CUDA_TRY(cudaMemcpy(p.h_tabs.data(), p.d_tabs.p, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost));

  // ---- per-segment index tables
  std::vector<int32_t> seg_tab(K);
  std::vector<int64_t> dpre(K), ipre(K), spre(K);
  int64_t accD = 0, accI = 0, accS = 0;
  int dmax = 0;
  for (int k = 0; k < K; ++k) {
    const int dg = p.po[k];
    dmax = std::max(dmax, dg);
    seg_tab[k] = p.rec_off[std::lower_bound(p.degs.begin(), p.degs.end(), dg) - p.degs.begin()];
    dpre[k] = accD, ipre[k] = accI, spre[k] = accS;
    accD += k == 0 ? (int64_t)(dg + 1) * (dg + 1) : (int64_t)dg * (dg + 1);
    accI += (int64_t)dg * (dg + 1);
    if (k + 1 < K) accS += dg + p.po[k + 1] + 1;
  }
  auto up = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = b.ensure(bytes);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice);
  };
  CUDA_TRY(up(p.d_seg_tab, seg_tab.data(), K * sizeof(int32_t)));
  CUDA_TRY(up(p.d_seg_start, p.seg_start.data(), (K + 1) * sizeof(int32_t)));
  CUDA_TRY(up(p.d_seg_dpre, dpre.data(), K * sizeof(int64_t)));
  CUDA_TRY(up(p.d_seg_ipre, ipre.data(), K * sizeof(int64_t)));
  CUDA_TRY(up(p.d_seg_spre, spre.data(), K * sizeof(int64_t)));

  if (p.adaptive) {  // D at the mid points of every unique degree (mpopt.py:3055-3060), on the device like the other tables
    int tot = 0;
    for (int dg : p.degs) p.dmid_off.push_back(tot), tot += dg * (dg + 1);
    p.h_dmid.resize(tot);
    CUDA_TRY(p.d_dmid.ensure((size_t)tot * sizeof(double)));
    for (size_t i = 0; i < p.degs.size(); ++i) {
      const int dg = p.degs[i], n1 = dg + 1;
      const double* R = p.h_tabs.data() + p.rec_off[i] + MpxTab::off_roots(n1);
      std::vector<double> mid(dg);
      for (int m = 0; m < dg; ++m) mid[m] = (R[m] + R[m + 1]) / 2.0;  // :3040-3046
      DevBuf dt;
      CUDA_TRY(dt.ensure(dg * sizeof(double)));
      CUDA_TRY(cudaMemcpy(dt.p, mid.data(), dg * sizeof(double), cudaMemcpyHostToDevice));
      mpx_basis_at_kernel<<<1, MPX_TAB_THREADS, (size_t)(n1 + 2) * sizeof(double), p.stream>>>(
          p.scheme, dg, p.tau_min, p.tau_max, 1, dg, dt.as<double>(), p.d_dmid.as<double>() + p.dmid_off[i]);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaStreamSynchronize(p.stream));
    }
    CUDA_TRY(cudaMemcpy(p.h_dmid.data(), p.d_dmid.p, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<int32_t> seg_dmid(K);
    for (int k = 0; k < K; ++k)
      seg_dmid[k] = p.dmid_off[std::lower_bound(p.degs.begin(), p.degs.end(), p.po[k]) - p.degs.begin()];
    CUDA_TRY(up(p.d_seg_dmid, seg_dmid.data(), K * sizeof(int32_t)));
    {
      std::vector<int64_t> rpre;
      for (auto& L : p.ph) rpre.insert(rpre.end(), L.seg_rpre.begin(), L.seg_rpre.end());
      CUDA_TRY(up(p.d_seg_rpre, rpre.data(), rpre.size() * sizeof(int64_t)));
    }
    CUDA_TRY(p.d_wpart.ensure((size_t)K * sizeof(double)));
    int njf = 0;
    for (auto& L : p.ph) {
      int n = 0;
      for (uint8_t b : L.pat_f) n += b;
      njf = std::max(njf, n);
    }
    const int n1 = dmax + 1;  // same formula as mpx_adapt_smem_doubles
    p.smem_adapt = 8 * (MpxTab::pad2(n1) + 2 * MpxTab::pad2(dmax * n1) + 2 * MpxTab::pad2((nx + nu) * n1) + 2 +
                        MpxTab::pad2(dmax * (3 * nx + njf + 2)) + MpxTab::pad2(dmax * (2 * nx + nu)) + MPX_THREADS);
    if (p.smem_adapt > 227 * 1024) return fail(MPX_ELIMIT, "polynomial degree too large for the adaptive NLP kernel");
    // residual-row images for the bulk-copy engine: two per CTA, when no row carries the K-wide width block of
    // time-dependent dynamics and a CTA stays below ~1/3 of an SM's shared memory (MPX_ADAPT_STAGE=0: direct stores)
    p.adapt_img.assign(p.P, 0);
    const char* se = getenv("MPX_ADAPT_STAGE");
    for (int ph = 0; ph < p.P && p.mid_res && !(se && atoi(se) == 0); ++ph) {
      const PhaseLayout& L = p.ph[ph];
      bool ft = false;
      int img = 0;
      for (int s = 0; s < nx; ++s) {
        ft |= L.f_t[s] != 0;
        img = std::max(img, dmax * ((dmax + 1) * L.res_nblk[s] + 2 * L.f_nz[s] + L.res_na[s] + 1));
      }
      img = MpxTab::pad2(img);
      if (!ft && p.smem_adapt + 2 * (img + 2) * 8 <= 76 * 1024) p.adapt_img[ph] = img;
    }
  }
  build_structure(p);
  if (!p.gather.empty()) CUDA_TRY(up(p.d_gather, p.gather.data(), p.gather.size() * sizeof(int64_t)));

  // ---- work buffers
  CUDA_TRY(p.d_z.ensure((size_t)p.n_z * sizeof(double)));
  CUDA_TRY(p.d_p.ensure((size_t)p.n_p * sizeof(double)));
  CUDA_TRY(p.d_sig0.ensure((size_t)K * p.P * sizeof(double)));
  CUDA_TRY(cudaMemset(p.d_sig0.p, 0, (size_t)K * p.P * sizeof(double)));  // read (times 0) even when nothing depends on t
  CUDA_TRY(p.d_g.ensure((size_t)p.n_g * sizeof(double)));
  CUDA_TRY(p.d_vals.ensure((size_t)p.nnz * sizeof(double)));
  if (!p.gather.empty()) CUDA_TRY(p.d_full.ensure((size_t)p.nnz_full * sizeof(double)));
  CUDA_TRY(p.d_grad.ensure((size_t)p.n_z * sizeof(double)));
  CUDA_TRY(p.d_partial.ensure(((size_t)(p.N + MPX_THREADS - 1) / MPX_THREADS + 1) * p.P * MPX_NPART * sizeof(double)));
  {
    std::vector<int32_t> node_seg((size_t)p.N, 0);
    for (int k = 0; k < K; ++k)
      for (int j = (k == 0 ? 0 : 1); j <= p.po[k]; ++j) node_seg[(size_t)p.seg_start[k] + j] = k;  // shared node: earlier segment
    CUDA_TRY(up(p.d_node_seg, node_seg.data(), node_seg.size() * sizeof(int32_t)));
  }
  CUDA_TRY(p.d_f.ensure(sizeof(double)));
  {
    const char* qe = getenv("MPX_QUEUE");  // 0: units dealt round-robin (measurements)
    if (!qe || atoi(qe)) {
      if (const char* k2 = getenv("MPX_K2_QUEUE")) p.k2_queue = atoi(k2) != 0;
      CUDA_TRY(p.d_queue.ensure((size_t)4 * p.P * sizeof(unsigned int)));
      CUDA_TRY(cudaMemset(p.d_queue.p, 0, (size_t)4 * p.P * sizeof(unsigned int)));
    }
  }
  {
    const char* fe = getenv("MPX_FGRAD_FUSED");  // 0: node kernel + separate final sum (two launches), for measurements
    if (!fe || atoi(fe)) {
      CUDA_TRY(p.d_ticket.ensure((size_t)2 * p.P * sizeof(unsigned int)));  // [P] f + grad_f, [P] hess_l
      CUDA_TRY(cudaMemset(p.d_ticket.p, 0, (size_t)2 * p.P * sizeof(unsigned int)));
    }
  }

  // ---- shared-memory budget (same formulas as the kernels)
  {
    const int n1 = dmax + 1;
    int base = MpxTab::size(n1) + MpxTab::pad2((nx + nu) * n1) + 2 * MpxTab::pad2(nx * n1) + 2;
    int stage = 0;
    for (int ph = 0; ph < p.P; ++ph) {
      int s_ = 0;
      for (int s = 0; s < nx; ++s) s_ += n1 * (n1 + p.ph[ph].f_next[s]);
      for (int c = 0; c < p.ph[ph].nc; ++c) s_ += n1 * p.ph[ph].c_len[c];
      stage = std::max(stage, s_);
    }
    p.smem_g = base * 8;
    p.smem_gjac = (base + MpxTab::pad2(stage)) * 8;
    p.smem_fgrad = (MpxTab::size(n1) + MpxTab::pad2((nx + nu) * n1) + 4 * MPX_NPART + 2) * 8;
    p.smem_too_big = p.smem_gjac > 227 * 1024;
  }

  // ---- v2 work units: runs of adjacent equal-degree segments, at most 32/pow2ceil(d+1) per unit
  {
    const char* force = getenv("MPX_KERNEL");
    const bool want_v2 = !(force && strcmp(force, "v1") == 0);
    const bool want_v4 = force && strcmp(force, "v4") == 0;  // row-block teams: opt-in until it beats v2 on the headline
    auto lw_of = [](int dg) { int lw = 2; while (lw < dg + 1) lw <<= 1; return lw; };
    int stage_cap = 0;  // widest single row-block image over degrees / phases
    for (int dg : p.degs) {
      const int n1 = dg + 1, rows_cap = (32 / std::max(lw_of(dg), 1)) * dg + 1;
      for (int ph = 0; ph < p.P && dg <= 31; ++ph) {
        for (int s = 0; s < nx; ++s) stage_cap = std::max(stage_cap, rows_cap * (n1 + p.ph[ph].f_next[s]));
        for (int c = 0; c < p.ph[ph].nc; ++c) stage_cap = std::max(stage_cap, rows_cap * p.ph[ph].c_len[c]);
      }
    }
    stage_cap = MpxTab::pad2(stage_cap);
    const long avail = 227L * 1024 - (long)(p.tab_doubles + 2) * 8;
    long per_warp = (long)(32 * (nx + nu) + stage_cap + 2) * 8;  // node values + one row-block image
    int warps = (int)std::min<long>(MPX2_MAX_THREADS / 32, avail > 0 ? avail / per_warp : 0);
    if (const char* wenv = getenv("MPX_V2_WARPS")) warps = std::max(1, std::min(warps, atoi(wenv)));
    if (want_v2 && dmax <= 31 && warps >= 1 && p.tab_doubles * 8L <= 96L * 1024) {
      std::vector<int32_t> uk, un;
      for (int k = p.seg_begin; k < p.seg_end;) {
        const int dg = p.po[k], cap = 32 / lw_of(dg);
        int n = 1;
        while (n < cap && k + n < p.seg_end && p.po[k + n] == dg) ++n;
        uk.push_back(k), un.push_back(n);
        k += n;
      }
      if (!p.uniform) {  // heavier units first so the static round-robin over warps stays balanced
        std::vector<int> idx(uk.size());
        for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
          return (long)un[a] * p.po[uk[a]] * (p.po[uk[a]] + 1) > (long)un[b] * p.po[uk[b]] * (p.po[uk[b]] + 1);
        });
        std::vector<int32_t> k2, n2;
        for (int i : idx) k2.push_back(uk[i]), n2.push_back(un[i]);
        uk.swap(k2), un.swap(n2);
      }
      int nsm = 0;
      CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, p.device));
      p.num_sms = nsm;
      struct { int multiProcessorCount; } prop{nsm};
      // v4: CTA = T teams of RT = nx + nc warps; per warp 32 doubles of node values + nbuf images of its row block
      p.v4_threads.assign(p.P, 0), p.v4_grid.assign(p.P, 0), p.v4_smem.assign(p.P, 0), p.v4_smem_g.assign(p.P, 0);
      p.v4_stage.assign(p.P, 0), p.v4_nbuf.assign(p.P, 1);
      bool ok4 = want_v4;
      for (int ph = 0; ph < p.P && ok4; ++ph) {
        const int RT = nx + p.ph[ph].nc;
        int img = 0;  // widest row-block image over degrees
        for (int dg : p.degs) {
          const int n1 = dg + 1, rows_cap = (32 / lw_of(dg)) * dg + 1;
          for (int s = 0; s < nx; ++s) img = std::max(img, rows_cap * (n1 + p.ph[ph].f_next[s]));
          for (int c = 0; c < p.ph[ph].nc; ++c) img = std::max(img, rows_cap * p.ph[ph].c_len[c]);
        }
        img = MpxTab::pad2(img);
        const int max_warps = MPX4_MAX_THREADS / 32;
        if (RT < 1 || RT > max_warps) { ok4 = false; break; }
        const long fixed = (long)(p.tab_doubles + 2) * 8, lim = 227L * 1024;
        auto smem_of = [&](int warps, int nbuf) {  // per warp: two ring slots of node values + nbuf images
          return fixed + (long)warps * (2L * (nx + nu) * 32 + (long)nbuf * (img + 2)) * 8;
        };
        const int units_per_sm = (int)((uk.size() + p.num_sms - 1) / p.num_sms);
        int T = std::max(1, std::min(max_warps / RT, units_per_sm));
        if (const char* tenv = getenv("MPX_V4_TEAMS")) T = std::max(1, std::min(max_warps / RT, atoi(tenv)));
        while (T > 1 && smem_of(RT * T, 1) > lim) --T;
        if (smem_of(RT * T, 1) > lim) { ok4 = false; break; }
        int nbuf = smem_of(RT * T, 2) <= lim ? 2 : 1;
        // two images per warp beat more teams when both do not fit: drop teams while that keeps >= 8 warps
        if (nbuf == 1 && !getenv("MPX_V4_TEAMS")) {
          int T2 = T;
          while (T2 > 1 && smem_of(RT * T2, 2) > lim) --T2;
          if (smem_of(RT * T2, 2) <= lim && RT * T2 >= 8) T = T2, nbuf = 2;
        }
        if (const char* benv = getenv("MPX_V4_NBUF")) {
          const int want = atoi(benv) >= 2 ? 2 : 1;
          if (smem_of(RT * T, want) <= lim) nbuf = want;
        }
        const int warps = RT * T;
        p.v4_threads[ph] = warps * 32, p.v4_stage[ph] = img, p.v4_nbuf[ph] = nbuf;
        p.v4_smem[ph] = (int)smem_of(warps, nbuf);
        p.v4_smem_g[ph] = (int)smem_of(warps, 0);
        p.v4_grid[ph] = (int)std::max<long>(1, std::min<long>(p.num_sms, ((long)uk.size() + T - 1) / T));
      }
      p.v4 = ok4 ? 1 : 0;
      // degree-specialised instance: compiled in (AOT list of the problem), or through NVRTC for large plans
      if (p.uniform && !getenv("MPX_NOSPEC")) {
        const int dg = p.po[0];
        bool aot = true;
        for (int ph = 0; ph < p.P; ++ph) aot = aot && p.prog->phases[ph]->has_degree(dg);
        const char* jit = getenv("MPX_JIT");
        const bool want_jit = jit ? atoi(jit) != 0 : (long)(p.seg_end - p.seg_begin) * dg >= 16384;
        const char* kname = ok4 ? "mpx_gjac4_kernel" : "mpx_gjac2_kernel";
        if (aot && !(jit && atoi(jit) < 0)) {
          p.spec_deg = dg;
        } else if (want_jit && d->program_source && d->program_key) {
          const std::string skey = std::string(d->program_key) + ":" + kname + ":d" + std::to_string(dg);
          const RtSpec* sp = find_rt_spec(skey);
          if (!sp) {
            int rc_ = compile_spec(d->program_key, d->program_source, p.P, dg, kname, &sp);
            if (rc_) return rc_;
          }
          p.rt_spec = sp;
        }
      }
      p.v2_units = (int)uk.size();
      p.v2_warps = std::max(1, std::min(warps, (p.v2_units + prop.multiProcessorCount - 1) / prop.multiProcessorCount));
      {
        // consecutive units on different SMs (MPX_F_SPREAD) when the units differ in cost (mixed degrees, sorted longest
        // first: config 3 8.65 -> 7.68 us) or do not fill the machine (config 5 6.76 -> 6.04 us); a full uniform plan keeps
        // CTA b on units [b * warps, (b + 1) * warps) (headline 18.13 vs 18.3 us).  MPX_V2_SPREAD=0/1 forces either.
        const char* sp = getenv("MPX_V2_SPREAD");
        p.v2_spread = sp ? atoi(sp) != 0 : (!p.uniform || 2L * p.v2_units <= (long)prop.multiProcessorCount * warps);
      }
      p.v2_grid = p.v2_spread ? std::min(prop.multiProcessorCount, p.v2_units)
                              : std::min(prop.multiProcessorCount, (p.v2_units + p.v2_warps - 1) / p.v2_warps);
      p.v2_stage_cap = stage_cap;
      // a second image per warp (the engine drains one while the next is assembled) when it fits
      p.v2_nbuf = (long)p.v2_warps * (per_warp + (long)(stage_cap + 2) * 8) <= avail ? 2 : 1;
      if (const char* benv = getenv("MPX_V2_NBUF")) p.v2_nbuf = (atoi(benv) >= 2 && p.v2_nbuf == 2) ? 2 : 1;
      if (p.v2_nbuf == 2) per_warp += (long)(stage_cap + 2) * 8;
      p.v2_smem_jac = (int)((p.tab_doubles + 2) * 8 + p.v2_warps * per_warp);
      p.v2_smem_g = (int)((p.tab_doubles + 2) * 8 + p.v2_warps * (long)(32 * (nx + nu)) * 8);
      CUDA_TRY(up(p.d_unit_k, uk.data(), uk.size() * sizeof(int32_t)));
      CUDA_TRY(up(p.d_unit_n, un.data(), un.size() * sizeof(int32_t)));
    }
  }

  if (p.wcol && (p.v2_warps == 0 || p.v4))
    return fail(MPX_ELIMIT, "adaptive NLP: the in-place width column needs the persistent-warp kernel (set MPX_ADAPT_WCOL=0)");
  if (p.v2_warps == 0 && p.smem_too_big)
    return fail(MPX_ELIMIT, "segment too large for the shared-memory staged kernels (degree x states)");
  p.origin += p.v4 ? ";gjac=v4" : (p.v2_warps > 0 ? ";gjac=v2" : ";gjac=v1");
  if ((p.v4 || p.v2_warps > 0) && p.spec_deg) p.origin += "/d" + std::to_string(p.spec_deg);
  if ((p.v4 || p.v2_warps > 0) && p.rt_spec) p.origin += "/jit-d" + std::to_string(p.po[0]);

  // ---- kernel arguments per phase (pointers filled per call)
  p.args.resize(p.P);
  for (int ph = 0; ph < p.P; ++ph) {
    MpxPhaseArgs& a = p.args[ph];
    const PhaseLayout& L = p.ph[ph];
    memset(&a, 0, sizeof(a));
    a.tabs = p.d_tabs.as<double>();
    a.seg_tab = p.d_seg_tab.as<int32_t>();
    a.seg_start = p.d_seg_start.as<int32_t>();
    a.seg_dpre = p.d_seg_dpre.as<int64_t>();
    a.seg_ipre = p.d_seg_ipre.as<int64_t>();
    a.K = K, a.N = N, a.seg_begin = p.seg_begin, a.seg_end = p.seg_end;
    a.uniform_deg = p.uniform ? p.po[0] : 0;
    a.unit_k = p.d_unit_k.as<int32_t>(), a.unit_n = p.d_unit_n.as<int32_t>();
    a.n_units = p.v2_units, a.tab_doubles = p.tab_doubles, a.stage_cap = p.v2_stage_cap;
    a.flags = (L.has_DU ? MPX_F_DU : 0) | (L.has_mU ? MPX_F_MU : 0) | (p.seg_end == K ? MPX_F_TAIL : 0);
    // constant blocks before the row blocks: the copy engine has work from the moment the tables land, one memory
    // round trip before the first image is ready (MPX_CONST_FIRST=0 restores the old order, for measurements)
    const char* cf = getenv("MPX_CONST_FIRST");
    if (!cf || atoi(cf)) a.flags |= MPX_F_CONST_FIRST;
    if (p.v2_spread) a.flags |= MPX_F_SPREAD;
    a.accumulate_f = ph > 0;
    a.zoff = L.zoff;
    a.gF = L.gF, a.gC = L.gC, a.gDU = L.gDU, a.gmU = L.gmU, a.gTC = L.gTC;
    for (int s = 0; s < nx; ++s) a.vF[s] = L.vF[s];
    for (int c = 0; c < L.nc; ++c) a.vC[c] = L.vC[c];
    a.vDU = L.vDU, a.vmU = L.vmU, a.vTC = L.vTC;
    a.nnzD = p.nnzD, a.nnzI = p.nnzI;
    for (int s = 0; s < nx; ++s) a.sx[s] = p.sx[s], a.isx[s] = 1.0 / p.sx[s];
    for (int c = 0; c < nu; ++c) a.isu[c] = 1.0 / p.su[c];
    for (int m = 0; m < na; ++m) a.isa[m] = 1.0 / p.sa[m];
    a.st = p.st, a.delta = p.tau_max - p.tau_min, a.tau0 = p.tau_min;
    a.ist = 1.0 / a.st, a.idelta = 1.0 / a.delta;
    if (p.adaptive) {
      a.ad_wcol = p.wcol ? 1 : 0;
      a.ad_sw_u = L.sw_u, a.ad_sw_x = L.sw_x, a.ad_res = p.mid_res;
      a.ad_img = p.adapt_img[ph];
      a.dmid = p.d_dmid.as<double>(), a.seg_dmid = p.d_seg_dmid.as<int32_t>();
      a.seg_rpre = p.d_seg_rpre.as<int64_t>() + (int64_t)ph * K;
      for (int s = 0; s < nx; ++s) a.eF[s] = L.eF[s];
      for (int c = 0; c < L.nc; ++c) a.eC[c] = L.eC[c];
      a.eSum = L.eSum, a.eUi = L.eUi, a.eXi = L.eXi, a.eRes = L.eRes, a.gSW = L.gSW;
      a.wpart = p.d_wpart.as<double>();
    }
  }
  if (p.adaptive) p.origin += p.base_direct ? ";adaptive/in-place" : ";adaptive";
  if (const char* fe = getenv("MPX_FUSE_PHASES")) p.fuse_phases = atoi(fe) != 0;
  if (p.P > 1 && p.prog->all && p.v2_warps > 0 && !p.v4 && !p.adaptive && !p.rt_spec && p.fuse_phases) p.origin += ";phases=fused";
  if (const char* te = getenv("MPX_TRACE")) {  // K2 timeline records, read back with mpx_trace_read
    if (atoi(te) && p.v2_warps > 0 && !p.v4) {
      const size_t nb = (size_t)MPX_TRACE_RING * p.v2_grid * p.v2_warps * MPX_TRACE_SLOTS * sizeof(unsigned long long);
      CUDA_TRY(p.d_trace.ensure(nb));
      CUDA_TRY(cudaMemset(p.d_trace.p, 0, nb));
    }
  }
  *out = pp.release();
  return MPX_OK;
}

// diagnostics: the timeline records of the last MPX_TRACE_RING K2 launches (MPX_TRACE=1 at plan creation), oldest
// first: n_warps records (= ring x warps per launch) of MPX_TRACE_SLOTS 64-bit stamps. out == NULL: only the count.
extern "C" int mpx_hess_zero_fill(const mpx_plan* p, int* state) {
  if (!p || !state) return fail(MPX_EINVAL, "NULL argument");
  *state = p->adaptive ? p->ah_zero_fill : 0;
  return MPX_OK;
}

extern "C" int mpx_trace_read(mpx_plan* p, int64_t* n_warps, int64_t* slots, unsigned long long* out) {
  if (!p) return fail(MPX_EINVAL, "NULL plan");
  const int64_t nw = p->d_trace.p ? (int64_t)MPX_TRACE_RING * p->v2_grid * p->v2_warps : 0;
  if (n_warps) *n_warps = nw;
  if (slots) *slots = MPX_TRACE_SLOTS;
  if (out && nw) {
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(out, p->d_trace.p, (size_t)nw * MPX_TRACE_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  }
  return MPX_OK;
}

// measurement aid: a kernel that occupies the stream for `usec` microseconds, so that a caller can enqueue a whole
// timed region behind it and keep host launch latency out of the CUDA-event pair (bench.py)
__global__ void mpx_gate_kernel(unsigned long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    __nanosleep(200);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
}
extern "C" int mpx_gate(void* stream, double usec) {
  if (usec < 0 || usec > 1e6) return fail(MPX_EINVAL, "mpx_gate: 0 <= usec <= 1e6");
  mpx_gate_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>((unsigned long long)(usec * 1e3));
  CUDA_TRY(cudaGetLastError());
  return MPX_OK;
}

extern "C" void mpx_plan_destroy(mpx_plan* plan) {
  if (!plan) return;
  cudaSetDevice(plan->device);
  delete plan;
}

extern "C" int mpx_sizes(const mpx_plan* p, int64_t* n_z, int64_t* n_p, int64_t* n_g, int64_t* nnz) {
  if (!p) return fail(MPX_EINVAL, "NULL plan");
  if (n_z) *n_z = p->n_z;
  if (n_p) *n_p = p->n_p;
  if (n_g) *n_g = p->n_g;
  if (nnz) *nnz = p->nnz;
  return MPX_OK;
}

extern "C" int mpx_jac_structure(const mpx_plan* p, int64_t* rowptr, int64_t* colind) {
  if (!p) return fail(MPX_EINVAL, "NULL plan");
  if (rowptr) memcpy(rowptr, p->rowptr.data(), p->rowptr.size() * sizeof(int64_t));
  if (colind) memcpy(colind, p->colind.data(), p->colind.size() * sizeof(int64_t));
  return MPX_OK;
}

extern "C" int mpx_jac_structure_ccs(const mpx_plan* p, int64_t* colptr, int64_t* rowind, int64_t* perm) {
  if (!p) return fail(MPX_EINVAL, "NULL plan");
  std::vector<int64_t> cnt(p->n_z + 1, 0);
  for (int64_t c : p->colind) ++cnt[c + 1];
  for (int64_t c = 0; c < p->n_z; ++c) cnt[c + 1] += cnt[c];
  if (colptr) memcpy(colptr, cnt.data(), (p->n_z + 1) * sizeof(int64_t));
  std::vector<int64_t> next(cnt.begin(), cnt.end() - 1);
  for (int64_t r = 0; r < p->n_g; ++r)
    for (int64_t e = p->rowptr[r]; e < p->rowptr[r + 1]; ++e) {
      const int64_t dst = next[p->colind[e]]++;
      if (rowind) rowind[dst] = r;
      if (perm) perm[dst] = e;
    }
  return MPX_OK;
}

extern "C" int mpx_plan_tables(const mpx_plan* p, int32_t deg, double* roots, double* D, double* w, double* Cmid) {
  if (!p) return fail(MPX_EINVAL, "NULL plan");
  const double* h = tab_of(*p, deg);
  if (!h) return fail(MPX_EINVAL, "degree not used by this plan");
  const int n1 = deg + 1;
  if (roots) memcpy(roots, h + MpxTab::off_roots(n1), n1 * sizeof(double));
  if (w) memcpy(w, h + MpxTab::off_w(n1), n1 * sizeof(double));
  if (D) memcpy(D, h + MpxTab::off_D(n1), (size_t)n1 * n1 * sizeof(double));
  if (Cmid) memcpy(Cmid, h + MpxTab::off_C(n1), (size_t)deg * n1 * sizeof(double));
  return MPX_OK;
}

// ------------------------------------------------------------------ shard runs
extern "C" int mpx_shard_runs(const mpx_plan* p, int32_t kind, int64_t* runs, int64_t* n_runs) {
  if (!p || !n_runs) return fail(MPX_EINVAL, "NULL argument");
  if (!p->gather.empty() && kind == 1) return fail(MPX_EINVAL, "shard runs need an unmasked Jacobian pattern");
  std::vector<int64_t> r;
  const int kb = p->seg_begin, ke = p->seg_end, N = p->N, K = p->K, nx = p->nx, nu = p->nu;
  const int64_t node_b = kb == 0 ? 0 : p->seg_start[kb] + 1, node_e = p->seg_start[ke] + 1;  // owned nodes [b, e)
  const int64_t mid_b = p->seg_start[kb], mid_e = p->seg_start[ke];
  int64_t dpre_b = 0, dpre_e = 0, ipre_b = 0, ipre_e = 0, spre_b = 0, spre_e = 0;
  for (int k = 0; k < ke; ++k) {
    const int64_t dd = k == 0 ? (int64_t)(p->po[k] + 1) * (p->po[k] + 1) : (int64_t)p->po[k] * (p->po[k] + 1);
    const int64_t ii = (int64_t)p->po[k] * (p->po[k] + 1);
    const int64_t ss = k + 1 < K ? p->po[k] + p->po[k + 1] + 1 : 0;
    if (k < kb) dpre_b += dd, ipre_b += ii, spre_b += ss;
    dpre_e += dd, ipre_e += ii, spre_e += ss;
  }
  const int kdu_b = kb, kdu_e = std::min(ke, K - 1);
  const bool tail = ke == K;
  auto push = [&](int64_t off, int64_t cnt) {
    if (cnt > 0) r.push_back(off), r.push_back(cnt);
  };
  for (int ph = 0; ph < p->P; ++ph) {
    const PhaseLayout& L = p->ph[ph];
    if (kind == 0) {
      for (int s = 0; s < nx; ++s) push(L.gF + (int64_t)s * N + node_b, node_e - node_b);
      for (int c = 0; c < L.nc; ++c) push(L.gC + (int64_t)c * N + node_b, node_e - node_b);
      if (L.has_DU)
        for (int c = 0; c < nu; ++c) push(L.gDU + (int64_t)c * N + node_b, node_e - node_b);
      if (L.has_mU)
        for (int c = 0; c < nu; ++c) push(L.gmU + (int64_t)c * (N - 1) + mid_b, mid_e - mid_b);
      if (L.has_dU)
        for (int c = 0; c < nu; ++c) push(L.gdU + (int64_t)c * (K - 1) + kdu_b, kdu_e - kdu_b);
      if (tail) push(L.gTC, L.ntc);
    } else if (kind == 1) {
      for (int s = 0; s < nx; ++s)
        push(L.vF[s] + node_b * L.f_next[s] + dpre_b, (node_e - node_b) * L.f_next[s] + dpre_e - dpre_b);
      for (int c = 0; c < L.nc; ++c) push(L.vC[c] + node_b * L.c_len[c], (node_e - node_b) * L.c_len[c]);
      if (L.has_DU)
        for (int c = 0; c < nu; ++c) push(L.vDU + (int64_t)c * p->nnzD + dpre_b, dpre_e - dpre_b);
      if (L.has_mU)
        for (int c = 0; c < nu; ++c) push(L.vmU + (int64_t)c * p->nnzI + ipre_b, ipre_e - ipre_b);
      if (L.has_dU)
        for (int c = 0; c < nu; ++c) push(L.vdU + (int64_t)c * p->nnzS + spre_b, spre_e - spre_b);
      if (tail) {
        int64_t n = 0;
        for (int v : L.tc_len) n += v;
        push(L.vTC, n);
      }
    } else {
      for (int v = 0; v < nx + nu; ++v) push(L.zoff + (int64_t)v * N + node_b, node_e - node_b);
    }
  }
  if (tail && kind == 0) push(p->g_events, p->n_g - p->g_events);
  if (tail && kind == 1) push(p->v_events, p->nnz_full - p->v_events);
  if (runs) memcpy(runs, r.data(), r.size() * sizeof(int64_t));
  *n_runs = (int64_t)r.size() / 2;
  return MPX_OK;
}

// ------------------------------------------------------------------ evaluation
// segment widths of a phase: the NLP parameters p, or (adaptive NLP) decision variables at the end of the phase's z
static const double* widths_of(const mpx_plan& p, const double* d_z, const double* d_p, int ph) {
  return p.adaptive ? d_z + p.ph[ph].zoff + (p.nvar - p.K) : d_p + (int64_t)ph * p.K;
}
static void scan_widths(mpx_plan& p, const double* d_z, const double* d_p, cudaStream_t st) {
  if (p.adaptive)
    mpx_scan_widths_kernel<<<p.P, MPX_SCAN_THREADS, 0, st>>>(d_z + (p.nvar - p.K), p.d_sig0.as<double>(), p.K, p.nvar);
  else
    mpx_scan_widths_kernel<<<p.P, MPX_SCAN_THREADS, 0, st>>>(d_p, p.d_sig0.as<double>(), p.K, p.K);
  ++p.launches;
}

// The *_dev entry points launch on the caller's stream from the caller's thread: make the plan's device current when it
// is not (a thread whose current device differs would otherwise launch on the wrong GPU).  One evaluation per plan may
// be in flight at a time: the launches share the plan's argument blocks and scratch buffers (include/mpx.h, Threading).
static inline cudaError_t use_plan_device(const mpx_plan& p) {
  int cur = -1;
  cudaError_t e = cudaGetDevice(&cur);
  if (e != cudaSuccess || cur == p.device) return e;
  return cudaSetDevice(p.device);
}

static int launch_g_jac(mpx_plan& p, const double* d_z, const double* d_p, double* d_g, double* d_vals, cudaStream_t st) {
  const bool jac = d_vals != nullptr;
  double* target = jac ? ((p.gather.empty() || p.base_direct) ? d_vals : p.d_full.as<double>()) : nullptr;
  const int grid = p.seg_end - p.seg_begin;
  bool need_sig = false;
  for (auto& L : p.ph) need_sig |= L.uses_t;
  if (need_sig) scan_widths(p, d_z, d_p, st);
  // multi-phase NLP: ONE launch for all phases and the phase-link rows (mpx_gjac2_multi_kernel)
  const int nl_ = (int)p.links.size() / 2;
  const bool fused = p.P > 1 && p.prog->all && p.v2_warps > 0 && !p.v4 && !p.adaptive && !p.rt_spec && p.fuse_phases &&
                     true;
  if (fused) {
    for (int ph = 0; ph < p.P; ++ph) {
      MpxPhaseArgs& a = p.args[ph];
      a.z = d_z, a.w = widths_of(p, d_z, d_p, ph), a.sig0 = p.d_sig0.as<double>() + (int64_t)ph * p.K;
      a.g = d_g, a.vals = target, a.v4_nbuf = p.v2_nbuf;
      a.queue = (p.d_queue.p && p.k2_queue) ? p.d_queue.as<unsigned int>() + 4 * ph : nullptr;
      if (p.d_trace.p) a.trace = nullptr;
    }
    MpxEvArgs ev{};
    if (nl_ && p.seg_end == p.K) {
      const int rows = nl_ * (p.nx + p.nu + 1);
      ev = MpxEvArgs{d_z, d_g, target, p.d_evcols.as<int64_t>(), p.d_evcols.as<int64_t>() + rows, p.g_events, p.v_events, rows};
    }
    CUDA_TRY(p.prog->all->gjac2_all(p.args.data(), p.P, ev, jac, p.spec_deg, p.v2_grid, p.v2_warps * 32,
                                    jac ? p.v2_smem_jac : p.v2_smem_g, st));
    ++p.launches;
  }
  for (int ph = 0; ph < p.P; ++ph) {
    MpxPhaseArgs& a = p.args[ph];
    a.z = d_z, a.w = widths_of(p, d_z, d_p, ph), a.sig0 = p.d_sig0.as<double>() + (int64_t)ph * p.K;
    a.g = d_g, a.vals = target;
    if (fused) {
      // done above
    } else if (p.v4) {
      a.stage_cap = p.v4_stage[ph], a.v4_nbuf = p.v4_nbuf[ph];
      const size_t sm4 = jac ? p.v4_smem[ph] : p.v4_smem_g[ph];
      if (p.rt_spec)
        CUDA_TRY(MpxRtPhase::go(static_cast<const RtSpec*>(p.rt_spec)->f[ph][jac], a, p.v4_grid[ph], p.v4_threads[ph], sm4, st));
      else
        CUDA_TRY(p.prog->phases[ph]->gjac4(a, jac, p.spec_deg, p.v4_grid[ph], p.v4_threads[ph], sm4, st));
    } else if (p.v2_warps > 0) {
      a.v4_nbuf = p.v2_nbuf;
      a.queue = (p.d_queue.p && p.k2_queue) ? p.d_queue.as<unsigned int>() + 4 * ph : nullptr;
      if (p.d_trace.p)
        a.trace = p.d_trace.as<unsigned long long>() + (size_t)(p.trace_seq++ % MPX_TRACE_RING) * p.v2_grid * p.v2_warps * MPX_TRACE_SLOTS;
      const size_t sm2 = jac ? p.v2_smem_jac : p.v2_smem_g;
      if (p.rt_spec)
        CUDA_TRY(MpxRtPhase::go(static_cast<const RtSpec*>(p.rt_spec)->f[ph][jac], a, p.v2_grid, p.v2_warps * 32, sm2, st, true));
      else
        CUDA_TRY(p.prog->phases[ph]->gjac2(a, jac, p.spec_deg, p.v2_grid, p.v2_warps * 32, sm2, st));
    }
    else
      CUDA_TRY(p.prog->phases[ph]->gjac(a, jac, grid, jac ? p.smem_gjac : p.smem_g, st));
    if (!fused) ++p.launches;
    const PhaseLayout& L = p.ph[ph];
    if (p.adaptive) {  // SW rows and the d/dw entries, staged behind the base kernels' values
      a.ad_jac = jac ? 1 : 0, a.ext = jac ? p.d_full.as<double>() + p.n_base : nullptr;
      a.ext_sw = !jac ? nullptr : (p.sw_direct[ph] >= 0 ? d_vals + p.sw_direct[ph] - L.eSum : a.ext);
      if (p.d_trace.p) a.trace = (size_t)p.K * MPX_TRACE_SLOTS <= (size_t)MPX_TRACE_RING * p.v2_grid * p.v2_warps * MPX_TRACE_SLOTS
                                     ? p.d_trace.as<unsigned long long>() : nullptr;  // MPX_TRACE=1: per-CTA timeline
      a.queue = p.d_queue.p ? p.d_queue.as<unsigned int>() + 4 * ph : nullptr;
      if (p.adapt_grid == (1 << 30) && a.queue) {  // persistent CTAs: a few per SM, segments dealt out through the counter
        int nsm = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, p.device));
        const char* ce = getenv("MPX_ADAPT_CTAS");
        p.adapt_grid = std::max(1, nsm * (ce && atoi(ce) > 0 ? atoi(ce) : 4));
      }
      CUDA_TRY(p.prog->phases[ph]->adapt(a, std::min(p.K, p.adapt_grid), (size_t)p.smem_adapt + (jac ? 2 * (size_t)(a.ad_img + 2) * 8 : 0), st));
      ++p.launches;
    }
    if (L.has_dU) {
      const int kb = p.seg_begin, ke = std::min(p.seg_end, p.K - 1);
      if (ke > kb) {
        MpxDuArgs du{d_z, p.d_tabs.as<double>(), p.d_seg_tab.as<int32_t>(), p.d_seg_start.as<int32_t>(),
                     p.d_seg_spre.as<int64_t>(), d_g, target, p.K, p.N, p.nx, p.nu, L.zoff, L.gdU, L.vdU, p.nnzS};
        const int n = (ke - kb) * p.nu;
        mpx_ducont_kernel<<<(n + 127) / 128, 128, 0, st>>>(du, kb, ke, jac ? 1 : 0);
        CUDA_TRY(cudaGetLastError());
        ++p.launches;
      }
    }
  }
  const int nl = (int)p.links.size() / 2;
  if (nl && p.seg_end == p.K && !fused) {
    const int rows = nl * (p.nx + p.nu + 1);
    MpxEvArgs ev{d_z, d_g, target, p.d_evcols.as<int64_t>(), p.d_evcols.as<int64_t>() + rows, p.g_events, p.v_events, rows};
    mpx_events_kernel<<<(rows + 127) / 128, 128, 0, st>>>(ev, jac ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    ++p.launches;
  }
  if (jac && !p.gather.empty() && p.gather_runs.empty() && !p.base_direct) {
    mpx_compact_kernel<<<(unsigned)((p.nnz + 255) / 256), 256, 0, st>>>(p.d_full.as<double>(), p.d_gather.as<int64_t>(),
                                                                       d_vals, p.nnz);
    CUDA_TRY(cudaGetLastError());
    ++p.launches;
  }
  for (size_t i = 0; jac && i + 1 < p.gather_runs.size(); i += 2) {  // adaptive NLP: everything but the in-place SW blocks
    const int64_t off = p.gather_runs[i], n = p.gather_runs[i + 1];
    mpx_compact_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.d_full.as<double>(), p.d_gather.as<int64_t>() + off,
                                                                   d_vals + off, n);
    CUDA_TRY(cudaGetLastError());
    ++p.launches;
  }
  return MPX_OK;
}

static int launch_f_grad(mpx_plan& p, const double* d_z, const double* d_p, double* d_f, double* d_grad, cudaStream_t st) {
  const bool grad = d_grad != nullptr;
  const int grid = p.seg_end - p.seg_begin;
  bool need_sig = false;
  for (auto& L : p.ph) need_sig |= L.cost_t;
  if (need_sig) scan_widths(p, d_z, d_p, st);
  for (int ph = 0; ph < p.P; ++ph) {
    MpxPhaseArgs& a = p.args[ph];
    a.z = d_z, a.w = widths_of(p, d_z, d_p, ph), a.sig0 = p.d_sig0.as<double>() + (int64_t)ph * p.K;
    a.node_begin = p.seg_begin == 0 ? 0 : p.seg_start[p.seg_begin] + 1, a.node_end = p.seg_start[p.seg_end] + 1;
    a.f_blocks = (a.node_end - a.node_begin + MPX_THREADS - 1) / MPX_THREADS;
    a.node_seg = p.d_node_seg.as<int32_t>();
    a.grad = d_grad, a.partial = p.d_partial.as<double>() + (int64_t)ph * ((p.N + MPX_THREADS - 1) / MPX_THREADS + 1) * MPX_NPART;
    a.fout = d_f;
    a.ticket = p.d_ticket.p ? p.d_ticket.as<unsigned int>() + ph : nullptr;
    CUDA_TRY(p.prog->phases[ph]->fgrad(a, grad, a.f_blocks, 0, st));
    ++p.launches;
    if (!a.ticket) {  // otherwise the last CTA of the node kernel has done the final sum
      CUDA_TRY(p.prog->phases[ph]->fgrad_final(a, grad, st));
      ++p.launches;
    }
    if (p.adaptive && grad) {  // d f / d w (mpopt.py:2945: the widths are part of x)
      CUDA_TRY(p.prog->phases[ph]->adapt_grad(a, (p.K + MPX_THREADS / 32 - 1) / (MPX_THREADS / 32), p.ph[ph].cost_t, st));
      p.launches += p.ph[ph].cost_t ? 2 : 1;
    }
  }
  return MPX_OK;
}

#define MPX_STAGE_SLOTS 4
#define MPX_STAGE_BYTES ((size_t)8 << 20)
static int ensure_staging(mpx_plan& p) {
  if (p.h_ring) return MPX_OK;
  CUDA_TRY(cudaHostAlloc((void**)&p.h_ring, MPX_STAGE_SLOTS * MPX_STAGE_BYTES, cudaHostAllocDefault));
  for (auto& e : p.ring_ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  if (!p.pool) {
    // copies and scatters are memory-bound: eight threads saturate a socket, more only fight the caller's own threads
    int n = std::min(8, std::max(1, (int)std::thread::hardware_concurrency() / 2));
    if (const char* e = getenv("MPX_HOST_THREADS")) n = atoi(e);
    p.pool.reset(new MpxPool(std::max(0, std::min(n, 64) - 1)));
  }
  return MPX_OK;
}
// true when the driver can DMA straight from / into `ptr` (cudaHostAlloc'ed or registered memory)
static bool is_pinned(const void* ptr) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}
static void pool_memcpy(MpxPool& pool, char* dst, const char* src, size_t bytes) {
  const int parts = (int)std::min<size_t>((size_t)pool.size(), (bytes + (256 << 10) - 1) / (256 << 10));
  if (parts <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  const size_t step = ((bytes / parts) + 63) & ~(size_t)63;
  pool.parallel_for(parts, [&](int i) {
    const size_t a = (size_t)i * step, b = std::min(bytes, a + step);
    if (a < b) memcpy(dst + a, src + a, b - a);
  });
}
// device -> PAGEABLE host: chunks go through the pinned ring (DMA at link speed) and are copied out by the pool while
// the next chunks are in flight.  `consume(slot_ptr, first_elem, n_elem)` handles a landed chunk (memcpy or scatter).
static int staged_d2h(mpx_plan& p, const double* src, size_t n, const std::function<void(const double*, size_t, size_t)>& consume) {
  int rc = ensure_staging(p);
  if (rc) return rc;
  const size_t per = MPX_STAGE_BYTES / sizeof(double), chunks = (n + per - 1) / per;
  auto issue = [&](size_t c) -> cudaError_t {
    const size_t a = c * per, cnt = std::min(per, n - a);
    cudaError_t e = cudaMemcpyAsync(p.h_ring + (c % MPX_STAGE_SLOTS) * MPX_STAGE_BYTES, src + a, cnt * sizeof(double),
                                    cudaMemcpyDeviceToHost, p.stream);
    return e != cudaSuccess ? e : cudaEventRecord(p.ring_ev[c % MPX_STAGE_SLOTS], p.stream);
  };
  for (size_t c = 0; c < std::min<size_t>(chunks, MPX_STAGE_SLOTS - 1); ++c) CUDA_TRY(issue(c));
  for (size_t c = 0; c < chunks; ++c) {
    if (c + MPX_STAGE_SLOTS - 1 < chunks) CUDA_TRY(issue(c + MPX_STAGE_SLOTS - 1));  // its slot was consumed at c - 1
    CUDA_TRY(cudaEventSynchronize(p.ring_ev[c % MPX_STAGE_SLOTS]));
    const size_t a = c * per, cnt = std::min(per, n - a);
    consume(reinterpret_cast<const double*>(p.h_ring + (c % MPX_STAGE_SLOTS) * MPX_STAGE_BYTES), a, cnt);
  }
  return MPX_OK;
}
static int d2h_any(mpx_plan& p, double* dst, const double* src, size_t n) {
  if (n * sizeof(double) < ((size_t)1 << 20) || is_pinned(dst)) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
    return MPX_OK;
  }
  return staged_d2h(p, src, n, [&](const double* s, size_t a, size_t cnt) {
    pool_memcpy(*p.pool, (char*)(dst + a), (const char*)s, cnt * sizeof(double));
  });
}
// PAGEABLE host -> device through the same ring
static int h2d_any(mpx_plan& p, double* dst, const double* src, size_t n) {
  if (n * sizeof(double) < ((size_t)1 << 20) || is_pinned(src)) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, p.stream));
    return MPX_OK;
  }
  int rc = ensure_staging(p);
  if (rc) return rc;
  const size_t per = MPX_STAGE_BYTES / sizeof(double), chunks = (n + per - 1) / per;
  for (size_t c = 0; c < chunks; ++c) {
    const size_t a = c * per, cnt = std::min(per, n - a);
    char* slot = p.h_ring + (c % MPX_STAGE_SLOTS) * MPX_STAGE_BYTES;
    if (c >= MPX_STAGE_SLOTS) CUDA_TRY(cudaEventSynchronize(p.ring_ev[c % MPX_STAGE_SLOTS]));  // the slot's last DMA is done
    pool_memcpy(*p.pool, slot, (const char*)(src + a), cnt * sizeof(double));
    CUDA_TRY(cudaMemcpyAsync(dst + a, slot, cnt * sizeof(double), cudaMemcpyHostToDevice, p.stream));
    CUDA_TRY(cudaEventRecord(p.ring_ev[c % MPX_STAGE_SLOTS], p.stream));
  }
  // a later staged transfer reuses the ring from slot 0: everything issued here has to be out of it first
  for (size_t c = chunks > MPX_STAGE_SLOTS ? chunks - MPX_STAGE_SLOTS : 0; c < chunks; ++c)
    CUDA_TRY(cudaEventSynchronize(p.ring_ev[c % MPX_STAGE_SLOTS]));
  return MPX_OK;
}

// device -> host copy of a result vector: whole for a full plan, only the runs this shard writes for a shard plan
// (the caller's array is full-size; other shards fill the rest -- each GPU moves its part over its own PCIe link)
static int download(mpx_plan& p, int kind, double* dst, const double* src, size_t n_full) {
  const bool shard = p.seg_begin != 0 || p.seg_end != p.K;
  if (!shard || (kind == 1 && !p.gather.empty())) return d2h_any(p, dst, src, n_full);
  std::vector<int64_t>& runs = p.h_runs[kind];
  if (runs.empty()) {
    int64_t n = 0;
    int rc = mpx_shard_runs(&p, kind, nullptr, &n);
    if (rc) return rc;
    runs.resize((size_t)2 * n);
    rc = mpx_shard_runs(&p, kind, runs.data(), &n);
    if (rc) return rc;
  }
  for (size_t i = 0; i + 1 < runs.size(); i += 2)
    CUDA_TRY(cudaMemcpyAsync(dst + runs[i], src + runs[i], (size_t)runs[i + 1] * sizeof(double), cudaMemcpyDeviceToHost,
                             p.stream));
  return MPX_OK;
}

static int upload_inputs(mpx_plan& p, const double* z, const double* pw) {
  if (!z || (!pw && p.n_p)) return fail(MPX_EINVAL, "z and p must not be NULL");
  CUDA_TRY(cudaSetDevice(p.device));
  p.staged = 0;  // every host entry point overwrites d_z (and then d_g / d_vals / d_f / d_grad): nothing staged survives
  if (p.seg_begin == 0 && p.seg_end == p.K) {
    int rc = h2d_any(p, p.d_z.as<double>(), z, (size_t)p.n_z);
    if (rc) return rc;
  } else {  // a shard reads only its own nodes (plus the node it shares with the previous segment) and t0 / tf / a
    // (one segment more at the end: the slope-continuity row of the last boundary reads the next segment's nodes)
    const int64_t nb = p.seg_start[p.seg_begin], cnt = p.seg_start[std::min(p.seg_end + 1, p.K)] - nb + 1, nv = p.nx + p.nu;
    double* dz = p.d_z.as<double>();
    for (int ph = 0; ph < p.P; ++ph) {
      const int64_t off = p.ph[ph].zoff;
      for (int64_t v = 0; v < nv; ++v)
        CUDA_TRY(cudaMemcpyAsync(dz + off + v * p.N + nb, z + off + v * p.N + nb, (size_t)cnt * sizeof(double),
                                 cudaMemcpyHostToDevice, p.stream));
      CUDA_TRY(cudaMemcpyAsync(dz + off + nv * p.N, z + off + nv * p.N, (size_t)(2 + p.na) * sizeof(double),
                               cudaMemcpyHostToDevice, p.stream));
      if (p.seg_begin > 0)  // terminal rows (last shard) and the slope-continuity rows also read the first node
        for (int64_t v = 0; v < nv; ++v)
          CUDA_TRY(cudaMemcpyAsync(dz + off + v * p.N, z + off + v * p.N, sizeof(double), cudaMemcpyHostToDevice, p.stream));
    }
  }
  if (p.n_p && (!p.p_valid || memcmp(p.h_p_cache.data(), pw, (size_t)p.n_p * sizeof(double)) != 0)) {
    p.h_p_cache.assign(pw, pw + p.n_p);
    CUDA_TRY(cudaMemcpyAsync(p.d_p.p, p.h_p_cache.data(), (size_t)p.n_p * sizeof(double), cudaMemcpyHostToDevice, p.stream));
    p.p_valid = true;
  }
  return MPX_OK;
}

extern "C" int mpx_eval_g(mpx_plan* p, const double* z, const double* pw, double* g) {
  if (!p || !g) return fail(MPX_EINVAL, "NULL argument");
  int rc = upload_inputs(*p, z, pw);
  if (rc) return rc;
  rc = launch_g_jac(*p, p->d_z.as<double>(), p->d_p.as<double>(), p->d_g.as<double>(), nullptr, p->stream);
  if (rc) return rc;
  rc = download(*p, 0, g, p->d_g.as<double>(), (size_t)p->n_g);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MPX_OK;
}

extern "C" int mpx_eval_jac_g(mpx_plan* p, const double* z, const double* pw, double* g, double* values) {
  if (!p || !values) return fail(MPX_EINVAL, "NULL argument");
  int rc = upload_inputs(*p, z, pw);
  if (rc) return rc;
  rc = launch_g_jac(*p, p->d_z.as<double>(), p->d_p.as<double>(), p->d_g.as<double>(), p->d_vals.as<double>(), p->stream);
  if (rc) return rc;
  if (g && (rc = download(*p, 0, g, p->d_g.as<double>(), (size_t)p->n_g))) return rc;
  if ((rc = download(*p, 1, values, p->d_vals.as<double>(), (size_t)p->nnz))) return rc;
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MPX_OK;
}

// ------------------------------------------------------------------ host hop: registered buffers, dynamic fetch
extern "C" int mpx_host_register(mpx_plan* p, void* ptr, int64_t bytes) {
  if (!p || !ptr || bytes <= 0) return fail(MPX_EINVAL, "mpx_host_register: NULL pointer or empty range");
  CUDA_TRY(cudaSetDevice(p->device));
  for (auto& r : p->regs)
    if (r.base == (char*)ptr && r.bytes == (size_t)bytes) return MPX_OK;
  CUDA_TRY(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
  p->regs.push_back({(char*)ptr, (size_t)bytes, false});
  return MPX_OK;
}
extern "C" int mpx_host_unregister(mpx_plan* p, void* ptr) {
  if (!p || !ptr) return fail(MPX_EINVAL, "mpx_host_unregister: NULL argument");
  for (size_t i = 0; i < p->regs.size(); ++i)
    if (p->regs[i].base == (char*)ptr) {
      CUDA_TRY(cudaSetDevice(p->device));
      CUDA_TRY(cudaStreamSynchronize(p->stream));
      CUDA_TRY(cudaHostUnregister(ptr));
      p->regs.erase(p->regs.begin() + i);
      return MPX_OK;
    }
  return fail(MPX_EINVAL, "mpx_host_unregister: that pointer is not registered with this plan");
}
static mpx_plan::HostReg* find_reg(mpx_plan& p, const void* ptr, size_t bytes) {
  for (auto& r : p.regs)
    if ((const char*)ptr >= r.base && (const char*)ptr + bytes <= r.base + r.bytes) return &r;
  return nullptr;
}

// positions of the Jacobian entries that depend on z or p (everything else is a table entry or +-1): in the defect
// rows F(s, i) every entry outside the row's own D block, plus the block's diagonal entry when d f_s / d x_s is not
// identically zero; all of the path and terminal rows.  Slope, mid-point, slope-continuity and event rows are constant
// (mpopt.py:321, :357-360, :408, :484-519).
static int build_dynamic(mpx_plan& p) {
  if (p.n_dyn >= 0) return MPX_OK;
  if (p.adaptive) return fail(MPX_EINVAL, "the dynamic fetch is not available for the adaptive NLP (its D blocks scale with the widths)");
  if (p.nnz >= (int64_t)1 << 31) return fail(MPX_ELIMIT, "the dynamic fetch indexes the Jacobian with 32 bits");
  const int nx = p.nx, nu = p.nu, nv = nx + nu + p.na, N = p.N;
  std::vector<int> nseg(N);
  for (int k = 0; k < p.K; ++k)
    for (int r = (k == 0 ? 0 : 1); r <= p.po[k]; ++r) nseg[p.seg_start[k] + r] = k;
  std::vector<int32_t>& pos = p.h_dyn_pos;
  pos.clear();
  for (int ph = 0; ph < p.P; ++ph) {
    const PhaseLayout& L = p.ph[ph];
    for (int s = 0; s < nx; ++s) {
      const bool diag = L.pat_f[(size_t)s * nv + s] != 0;
      for (int i = 0; i < N; ++i) {
        const int k = nseg[i];
        const int64_t c0 = L.zoff + (int64_t)s * N + p.seg_start[k], c1 = c0 + p.po[k], cd = L.zoff + (int64_t)s * N + i;
        const int64_t r = L.gF + (int64_t)s * N + i;
        for (int64_t e = p.rowptr[r]; e < p.rowptr[r + 1]; ++e) {
          const int64_t c = p.colind[e];
          if (c < c0 || c > c1 || (diag && c == cd)) pos.push_back((int32_t)e);
        }
      }
    }
    for (int64_t r = L.gC; r < L.gC + (int64_t)L.nc * N; ++r)
      for (int64_t e = p.rowptr[r]; e < p.rowptr[r + 1]; ++e) pos.push_back((int32_t)e);
    for (int64_t r = L.gTC; r < L.gTC + L.ntc; ++r)
      for (int64_t e = p.rowptr[r]; e < p.rowptr[r + 1]; ++e) pos.push_back((int32_t)e);
  }
  std::sort(pos.begin(), pos.end());
  CUDA_TRY(cudaSetDevice(p.device));
  CUDA_TRY(p.d_dyn_pos.ensure(std::max<size_t>(pos.size(), 1) * sizeof(int32_t)));
  CUDA_TRY(cudaMemcpy(p.d_dyn_pos.p, pos.data(), pos.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  p.n_dyn = (int64_t)pos.size();
  return MPX_OK;
}

extern "C" int mpx_jac_dynamic_count(mpx_plan* p, int64_t* n_dynamic) {
  if (!p || !n_dynamic) return fail(MPX_EINVAL, "NULL argument");
  int rc = build_dynamic(*p);
  if (rc) return rc;
  *n_dynamic = p->n_dyn;
  return MPX_OK;
}
extern "C" int mpx_jac_dynamic_positions(mpx_plan* p, int32_t* pos) {
  if (!p || !pos) return fail(MPX_EINVAL, "NULL argument");
  int rc = build_dynamic(*p);
  if (rc) return rc;
  memcpy(pos, p->h_dyn_pos.data(), p->h_dyn_pos.size() * sizeof(int32_t));
  return MPX_OK;
}

// g + jac_g into caller-owned host buffers, moving only what changed.  The first call on a given `values` buffer is a
// full mpx_eval_jac_g, which also writes the constant entries; later calls on the SAME buffer evaluate on the device,
// gather the n_dynamic z- / p-dependent entries (32 % of the bytes at the headline size), bring them over in chunks
// through the pinned ring and scatter them into place with the worker pool while the next chunk is in flight.  (Storing
// them from a kernel straight into mapped host memory was measured too: 2.07 ms against 2.12 ms for the full copy --
// 40-byte PCIe writes -- so the packed copy + host scatter is what ships.)  The caller must not modify `values` between
// calls (IPOPT does not).  Works with any host memory; registration only speeds up the first, full call.
extern "C" int mpx_eval_jac_g_dynamic(mpx_plan* p, const double* z, const double* pw, double* g, double* values) {
  if (!p || !values) return fail(MPX_EINVAL, "NULL argument");
  const bool shard = p->seg_begin != 0 || p->seg_end != p->K;
  mpx_plan::HostReg* reg = find_reg(*p, values, (size_t)p->nnz * sizeof(double));
  const bool primed = (reg && reg->primed) || p->dyn_primed == values;
  if (shard || p->adaptive || !primed) {
    int rc = mpx_eval_jac_g(p, z, pw, g, values);
    if (rc == MPX_OK && !shard && !p->adaptive) {
      if (reg) reg->primed = true;
      else p->dyn_primed = values;
    }
    return rc;
  }
  int rc = build_dynamic(*p);
  if (rc) return rc;
  rc = upload_inputs(*p, z, pw);
  if (rc) return rc;
  rc = launch_g_jac(*p, p->d_z.as<double>(), p->d_p.as<double>(), p->d_g.as<double>(), p->d_vals.as<double>(), p->stream);
  if (rc) return rc;
  CUDA_TRY(p->d_dyn_vals.ensure(std::max<int64_t>(p->n_dyn, 1) * sizeof(double)));
  if (p->n_dyn > 0) {
    mpx_gather_dyn_kernel<<<(unsigned)((p->n_dyn + 255) / 256), 256, 0, p->stream>>>(p->d_vals.as<double>(),
                                                                                  p->d_dyn_pos.as<int32_t>(),
                                                                                  p->d_dyn_vals.as<double>(), p->n_dyn);
    CUDA_TRY(cudaGetLastError());
    ++p->launches;
    const int32_t* pos = p->h_dyn_pos.data();
    rc = staged_d2h(*p, p->d_dyn_vals.as<double>(), (size_t)p->n_dyn, [&](const double* src, size_t a, size_t cnt) {
      const int parts = p->pool->size();
      const size_t step = (cnt + parts - 1) / parts;
      p->pool->parallel_for(parts, [&](int t) {
        const size_t b = (size_t)t * step, e = std::min(cnt, b + step);
        for (size_t i = b; i < e; ++i) values[pos[a + i]] = src[i];
      });
    });
    if (rc) return rc;
  }
  if (g && (rc = download(*p, 0, g, p->d_g.as<double>(), (size_t)p->n_g))) return rc;
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MPX_OK;
}

// the dynamic entries packed (n_dynamic doubles, order of mpx_jac_dynamic_positions), for callers that keep their own
// copy of the constants or consume the entries in packed form
extern "C" int mpx_eval_jac_g_packed(mpx_plan* p, const double* z, const double* pw, double* g, double* packed) {
  if (!p || !packed) return fail(MPX_EINVAL, "NULL argument");
  int rc = build_dynamic(*p);
  if (rc) return rc;
  rc = upload_inputs(*p, z, pw);
  if (rc) return rc;
  rc = launch_g_jac(*p, p->d_z.as<double>(), p->d_p.as<double>(), p->d_g.as<double>(), p->d_vals.as<double>(), p->stream);
  if (rc) return rc;
  if (g && (rc = download(*p, 0, g, p->d_g.as<double>(), (size_t)p->n_g))) return rc;
  CUDA_TRY(p->d_dyn_vals.ensure(std::max<int64_t>(p->n_dyn, 1) * sizeof(double)));
  if (p->n_dyn > 0) {
    mpx_gather_dyn_kernel<<<(unsigned)((p->n_dyn + 255) / 256), 256, 0, p->stream>>>(p->d_vals.as<double>(),
                                                                                  p->d_dyn_pos.as<int32_t>(),
                                                                                  p->d_dyn_vals.as<double>(), p->n_dyn);
    CUDA_TRY(cudaGetLastError());
    ++p->launches;
    if ((rc = d2h_any(*p, packed, p->d_dyn_vals.as<double>(), (size_t)p->n_dyn))) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MPX_OK;
}

extern "C" int mpx_eval_f(mpx_plan* p, const double* z, const double* pw, double* f) {
  if (!p || !f) return fail(MPX_EINVAL, "NULL argument");
  int rc = upload_inputs(*p, z, pw);
  if (rc) return rc;
  rc = launch_f_grad(*p, p->d_z.as<double>(), p->d_p.as<double>(), p->d_f.as<double>(), nullptr, p->stream);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(f, p->d_f.p, sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MPX_OK;
}

extern "C" int mpx_eval_grad_f(mpx_plan* p, const double* z, const double* pw, double* f, double* grad) {
  if (!p || !grad) return fail(MPX_EINVAL, "NULL argument");
  int rc = upload_inputs(*p, z, pw);
  if (rc) return rc;
  rc = launch_f_grad(*p, p->d_z.as<double>(), p->d_p.as<double>(), p->d_f.as<double>(), p->d_grad.as<double>(), p->stream);
  if (rc) return rc;
  if (f) CUDA_TRY(cudaMemcpyAsync(f, p->d_f.p, sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaMemcpyAsync(grad, p->d_grad.p, (size_t)p->n_z * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MPX_OK;
}

// ------------------------------------------------------------------ Hessian of the Lagrangian
namespace {
// entry order and categories of the node Hessian: identical to mpopt_b200/program.py::PhaseProgram.hess_layout
struct HessEntry { int a, b, cat, slot; };

int build_hessian(mpx_plan& p) {
  if (p.hess_built) return MPX_OK;
  if (p.seg_begin != 0 || p.seg_end != p.K) return fail(MPX_EINVAL, "the Hessian needs a plan over all segments");
  for (auto& L : p.ph)
    if (!L.has_hess) return fail(MPX_ENOPROGRAM, "the problem description carries no Hessian patterns (pat_hw / pat_ht)");
  if (p.adaptive)
    for (auto& L : p.ph) {
      if (L.uses_t || L.cost_t)
        return fail(MPX_EINVAL, "the Hessian of the adaptive NLP is not available for problems with explicit time "
                                "dependence (every earlier width moves t; use a quasi-Newton Hessian)");
      if (L.pat_hf.empty()) return fail(MPX_ENOPROGRAM, "the problem description carries no pat_hf (adaptive Hessian)");
    }
  const int nx = p.nx, nu = p.nu, na = p.na, ny = nx + nu, nv = ny + na, N = p.N, NW = nv + 2, NT = 2 * nx + 2 + na;
  const int it = nv, ih = nv + 1;
  std::vector<std::pair<int64_t, int64_t>> trip;
  std::vector<std::vector<HessEntry>> ents(p.P);
  std::vector<std::vector<int>> tys(p.P);
  auto colv = [&](const PhaseLayout& L, int q, int64_t i) { return L.zoff + (int64_t)q * N + i; };  // node variable q < ny
  auto colT0 = [&](const PhaseLayout& L) { return L.zoff + (int64_t)ny * N; };
  auto colA = [&](const PhaseLayout& L, int m) { return L.zoff + (int64_t)ny * N + 2 + m; };
  auto termcol = [&](const PhaseLayout& L, int v) -> int64_t {  // xf.., x0.., tf, t0, a..
    if (v < nx) return colv(L, v, N - 1);
    if (v < 2 * nx) return colv(L, v - nx, 0);
    if (v == 2 * nx) return colT0(L) + 1;
    if (v == 2 * nx + 1) return colT0(L);
    return colA(L, v - 2 * nx - 2);
  };
  auto corner_cols = [&](const PhaseLayout& L, int c, int64_t& r, int64_t& cc) {
    if (c == 0) r = colT0(L), cc = colT0(L);
    else if (c == 1) r = colT0(L) + 1, cc = colT0(L);
    else if (c == 2) r = colT0(L) + 1, cc = colT0(L) + 1;
    else if (c < 3 + 2 * na) r = colA(L, (c - 3) / 2), cc = colT0(L) + ((c - 3) & 1);
    else {
      int m = 0, rem = c - 3 - 2 * na;
      while (rem > m) rem -= m + 1, ++m;
      r = colA(L, m), cc = colA(L, rem);
    }
  };
  const int n_corner = 3 + 2 * na + na * (na + 1) / 2;
  std::vector<std::vector<uint8_t>> corner_on(p.P, std::vector<uint8_t>(n_corner, 0));
  for (int ph = 0; ph < p.P; ++ph) {
    const PhaseLayout& L = p.ph[ph];
    std::vector<int>& ty = tys[ph];
    for (int b = 0; b < ny; ++b)
      if (L.pat_hw[(size_t)it * NW + b] || L.pat_hw[(size_t)ih * NW + b]) ty.push_back(b);
    int n_yy = 0, n_ay = 0;
    for (int a = 0; a < NW; ++a)
      for (int b = 0; b <= a; ++b) {
        if (!L.pat_hw[(size_t)a * NW + b]) continue;
        HessEntry e{a, b, 0, 0};
        if (a < ny) e.cat = 0, e.slot = n_yy++;
        else if (a < nv && b < ny) e.cat = 1, e.slot = n_ay++;
        else if (a < nv) e.cat = 2, e.slot = 3 + 2 * na + (a - ny) * (a - ny + 1) / 2 + (b - ny);
        else if (b < ny) e.cat = a == it ? 3 : 4, e.slot = (int)(std::find(ty.begin(), ty.end(), b) - ty.begin());
        else if (b < nv) e.cat = a == it ? 5 : 6, e.slot = b - ny;
        else if (a == it) e.cat = 7;
        else e.cat = 8;
        ents[ph].push_back(e);
        // structural entries
        if (e.cat == 0)
          for (int64_t i = 0; i < N; ++i) trip.emplace_back(colv(L, a, i), colv(L, b, i));
        else if (e.cat == 1)
          for (int64_t i = 0; i < N; ++i) trip.emplace_back(colA(L, a - ny), colv(L, b, i));
        else if (e.cat == 2) corner_on[ph][e.slot] = 1;
        else if (e.cat == 5 || e.cat == 6) corner_on[ph][3 + 2 * e.slot] = corner_on[ph][3 + 2 * e.slot + 1] = 1;
        else if (e.cat == 7 || e.cat == 8) corner_on[ph][0] = corner_on[ph][1] = corner_on[ph][2] = 1;
      }
    for (int b : ty)
      for (int64_t i = 0; i < N; ++i) trip.emplace_back(colT0(L), colv(L, b, i)), trip.emplace_back(colT0(L) + 1, colv(L, b, i));
    for (int c = 0; c < n_corner; ++c)
      if (corner_on[ph][c]) {
        int64_t r, cc;
        corner_cols(L, c, r, cc);
        trip.emplace_back(r, cc);
      }
    for (int a = 0; a < NT; ++a)
      for (int b = 0; b <= a; ++b)
        if (L.pat_ht[(size_t)a * NT + b]) {
          const int64_t ca = termcol(L, a), cb = termcol(L, b);
          trip.emplace_back(std::max(ca, cb), std::min(ca, cb));
        }
    if (p.adaptive) {  // what the widths add (mpx_adapt_hess_kernel): see the comment block there
      auto colW = [&](int k) { return colT0(L) + 2 + na + k; };
      std::vector<uint8_t> gy(nv, 0);  // psi = sum mu Sx f depends on node variable v
      bool psi_nz = false;
      for (int s = 0; s < nx; ++s) {
        psi_nz |= L.f_nz[s] != 0;
        for (int v = 0; v < nv; ++v) gy[v] |= L.pat_f[(size_t)s * nv + v];
      }
      for (int k = 0; k < p.K; ++k) {
        const int s0 = p.seg_start[k], d = p.po[k], rb0 = k == 0 ? 0 : 1;
        // (1) h_k bilinear in (T0 | TF, w_k)
        for (int b = 0; b < ny; ++b)
          if (L.pat_hw[(size_t)ih * NW + b])
            for (int j = rb0; j <= d; ++j) trip.emplace_back(colW(k), colv(L, b, s0 + j));
        for (int m = 0; m < na; ++m)
          if (L.pat_hw[(size_t)ih * NW + ny + m]) trip.emplace_back(colW(k), colA(L, m));
        if (L.phi_nz) trip.emplace_back(colW(k), colT0(L)), trip.emplace_back(colW(k), colT0(L) + 1);
        if (!p.mid_res) continue;
        // (2) mid-point residual rows
        for (int b = 0; b < ny; ++b)
          if (b < nx || gy[b])
            for (int j = 0; j <= d; ++j) trip.emplace_back(colW(k), colv(L, b, s0 + j));
        for (int b = 0; b < ny; ++b)
          if (gy[b])
            for (int j = 0; j <= d; ++j)
              trip.emplace_back(colT0(L), colv(L, b, s0 + j)), trip.emplace_back(colT0(L) + 1, colv(L, b, s0 + j));
        for (int m = 0; m < na; ++m)
          if (gy[ny + m]) {
            trip.emplace_back(colW(k), colA(L, m));
            corner_on[ph][3 + 2 * m] = corner_on[ph][3 + 2 * m + 1] = 1;
            trip.emplace_back(colA(L, m), colT0(L)), trip.emplace_back(colA(L, m), colT0(L) + 1);
          }
        if (psi_nz) {
          trip.emplace_back(colW(k), colT0(L)), trip.emplace_back(colW(k), colT0(L) + 1);
          trip.emplace_back(colW(k), colW(k));
        }
        for (int a = 0; a < nv; ++a)
          for (int b = 0; b <= a; ++b) {
            if (!L.pat_hf[(size_t)a * nv + b]) continue;
            if (a < ny) {
              for (int j = 0; j <= d; ++j)
                for (int jp = 0; jp <= (a == b ? j : d); ++jp) trip.emplace_back(colv(L, a, s0 + j), colv(L, b, s0 + jp));
            } else if (b < ny) {
              for (int j = 0; j <= d; ++j) trip.emplace_back(colA(L, a - ny), colv(L, b, s0 + j));
            } else {
              const int c = 3 + 2 * na + (a - ny) * (a - ny + 1) / 2 + (b - ny);
              if (!corner_on[ph][c]) {
                corner_on[ph][c] = 1;
                int64_t r, cc;
                corner_cols(L, c, r, cc);
                trip.emplace_back(r, cc);
              }
            }
          }
      }
    }
  }
  std::sort(trip.begin(), trip.end());
  trip.erase(std::unique(trip.begin(), trip.end()), trip.end());
  p.h_rowptr.assign((size_t)p.n_z + 1, 0);
  p.h_colind.resize(trip.size());
  for (size_t e = 0; e < trip.size(); ++e) ++p.h_rowptr[(size_t)trip[e].first + 1], p.h_colind[e] = trip[e].second;
  for (int64_t r = 0; r < p.n_z; ++r) p.h_rowptr[(size_t)r + 1] += p.h_rowptr[(size_t)r];
  auto pos_of = [&](int64_t r, int64_t c) -> int64_t {
    const int64_t* b = p.h_colind.data() + p.h_rowptr[(size_t)r];
    const int64_t* e = p.h_colind.data() + p.h_rowptr[(size_t)r + 1];
    return std::lower_bound(b, e, c) - p.h_colind.data();
  };
  // which terminal entries share their position with a node / corner entry?  mark the node-owned positions
  std::vector<uint8_t> owned(trip.size(), 0);
  p.hess_ph.resize(p.P);
  auto upload = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = b.ensure(bytes ? bytes : 8);
    if (e != cudaSuccess || !bytes) return e;
    return cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice);
  };
  CUDA_TRY(cudaSetDevice(p.device));
  for (int ph = 0; ph < p.P; ++ph) {
    const PhaseLayout& L = p.ph[ph];
    auto& H = p.hess_ph[ph];
    std::vector<int64_t> pyy, pay, pty, pcorner(n_corner, -1);
    const size_t nty = tys[ph].size();
    pty.resize(2 * nty * (size_t)N);
    for (const HessEntry& e : ents[ph]) {
      if (e.cat == 0) {
        pyy.resize(pyy.size() + N);
        for (int64_t i = 0; i < N; ++i) owned[pyy[(size_t)e.slot * N + i] = pos_of(colv(L, e.a, i), colv(L, e.b, i))] = 1;
      } else if (e.cat == 1) {
        pay.resize(pay.size() + N);
        for (int64_t i = 0; i < N; ++i) owned[pay[(size_t)e.slot * N + i] = pos_of(colA(L, e.a - ny), colv(L, e.b, i))] = 1;
      }
    }
    for (size_t q = 0; q < nty; ++q)
      for (int64_t i = 0; i < N; ++i) {
        owned[pty[q * N + i] = pos_of(colT0(L), colv(L, tys[ph][q], i))] = 1;
        owned[pty[(nty + q) * N + i] = pos_of(colT0(L) + 1, colv(L, tys[ph][q], i))] = 1;
      }
    for (int c = 0; c < n_corner; ++c)
      if (corner_on[ph][c]) {
        int64_t r, cc;
        corner_cols(L, c, r, cc);
        owned[pcorner[c] = pos_of(r, cc)] = 1;
      }
    std::vector<int64_t> pterm;
    std::vector<int32_t> tassign;
    for (int a = 0; a < NT; ++a)
      for (int b = 0; b <= a; ++b)
        if (L.pat_ht[(size_t)a * NT + b]) {
          const int64_t ca = termcol(L, a), cb = termcol(L, b);
          const int64_t pos = pos_of(std::max(ca, cb), std::min(ca, cb));
          // who else writes this position?  (the terminal variables are x0, xf, t0, tf, a: node 0, node N-1 or the corner)
          int own = 3;
          if (owned[pos]) {
            own = -1;
            for (int c = 0; c < n_corner && own < 0; ++c)
              if (pcorner[c] == pos) own = 2;
            auto at_node = [&](const std::vector<int64_t>& tab, int64_t node) {
              for (size_t q = 0; q * (size_t)N < tab.size(); ++q)
                if (tab[q * (size_t)N + (size_t)node] == pos) return true;
              return false;
            };
            if (own < 0 && (at_node(pyy, 0) || at_node(pay, 0) || at_node(pty, 0))) own = 0;
            if (own < 0 && (at_node(pyy, N - 1) || at_node(pay, N - 1) || at_node(pty, N - 1))) own = 1;
            if (own < 0) return fail(MPX_EINVAL, "Hessian pattern: a terminal entry shares its position with an interior node");
          }
          pterm.push_back(pos), tassign.push_back(own);
        }
    // rows of interior nodes are regular: position = base + i * stride (checked, not assumed)
    std::vector<int64_t> lin;
    bool affine = N >= 4;
    auto add_lin = [&](const std::vector<int64_t>& pos, size_t n_ent) {
      for (size_t e = 0; e < n_ent && affine; ++e) {
        const int64_t* q = pos.data() + e * (size_t)N;
        const int64_t stride = q[2] - q[1], base = q[1] - stride;
        for (int64_t i = 1; i < N - 1; ++i)
          if (q[i] != base + i * stride) { affine = false; break; }
        lin.push_back(base), lin.push_back(stride);
      }
    };
    add_lin(pyy, pyy.size() / (size_t)N), add_lin(pay, pay.size() / (size_t)N), add_lin(pty, 2 * nty);
    memset(&H.lin, 0, sizeof H.lin);
    if (affine && lin.size() / 2 <= MPX_HLIN_MAX) {
      for (size_t e = 0; e < lin.size() / 2 && affine; ++e) {
        H.lin.base[e] = lin[2 * e];
        if (lin[2 * e + 1] < 0 || lin[2 * e + 1] > INT32_MAX) affine = false;
        H.lin.stride[e] = (int32_t)lin[2 * e + 1];
      }
      H.lin.n = affine ? (int32_t)(lin.size() / 2) : 0;
    }
    // node-diagonal rows: are the cat-0 entries of row a adjacent, in entry order, with the rows of consecutive nodes
    // back to back?  (true for the base NLP; the kernel then writes each row block as one contiguous run per warp)
    H.lin.rows = H.lin.n > 0 ? 1 : 0;
    for (int a = 0; a < ny && H.lin.rows; ++a) {
      int len = 0, first = -1;
      for (const HessEntry& e : ents[ph])
        if (e.cat == 0 && e.a == a) first = first < 0 ? e.slot : first, ++len;
      int rank = 0;
      for (const HessEntry& e : ents[ph])
        if (e.cat == 0 && e.a == a) {
          if (H.lin.stride[e.slot] != len || H.lin.base[e.slot] != H.lin.base[first] + rank) H.lin.rows = 0;
          ++rank;
        }
    }
    if (const char* re = getenv("MPX_HESS_ROWS")) H.lin.rows = H.lin.rows && atoi(re) != 0;  // 0: scattered stores (measurements)
    CUDA_TRY(upload(H.pos_yy, pyy.data(), pyy.size() * sizeof(int64_t)));
    CUDA_TRY(upload(H.pos_ay, pay.data(), pay.size() * sizeof(int64_t)));
    CUDA_TRY(upload(H.pos_ty, pty.data(), pty.size() * sizeof(int64_t)));
    CUDA_TRY(upload(H.pos_corner, pcorner.data(), pcorner.size() * sizeof(int64_t)));
    CUDA_TRY(upload(H.pos_term, pterm.data(), pterm.size() * sizeof(int64_t)));
    CUDA_TRY(upload(H.term_assign, tassign.data(), tassign.size() * sizeof(int32_t)));
    H.blocks = (N + MPX_HESS_THREADS - 1) / MPX_HESS_THREADS, H.n_corner = n_corner;
    CUDA_TRY(H.part.ensure((size_t)H.blocks * n_corner * sizeof(double)));
  }
  if (p.adaptive) {
    // Positions of everything a segment adds, looked up ONCE here (the kernel used to find each entry by binary search
    // in a 178 MB column-index array).  Per segment, with n1 = degree + 1 and e over the second derivatives of psi:
    //   [ny][n1] (w_k, Y_jb) | [ny][n1] (T0, Y_jb) | [ny][n1] (TF, Y_jb) | [NRH][n1]: first entry of the run
    //   (Y_j va, Y_0 vb) of a dense block, or the single entry (a_m, Y_j vb) | [na] (w_k, a_m) | (w_k, T0), (w_k, TF), (w_k, w_k)
    // -1: not in the pattern (the value is an exact zero).
    auto find_pos = [&](int64_t r, int64_t c) -> int64_t {
      const int64_t* b = p.h_colind.data() + p.h_rowptr[(size_t)r];
      const int64_t* e = p.h_colind.data() + p.h_rowptr[(size_t)r + 1];
      const int64_t* q = std::lower_bound(b, e, c);
      return (q < e && *q == c) ? (int64_t)(q - p.h_colind.data()) : -1;
    };
    for (int ph = 0; ph < p.P; ++ph) {
      const PhaseLayout& L = p.ph[ph];
      std::vector<std::pair<int, int>> rh;
      for (int a = 0; a < nv; ++a)
        for (int b = 0; b <= a; ++b)
          if (L.pat_hf[(size_t)a * nv + b]) rh.emplace_back(a, b);
      std::vector<int64_t> off((size_t)p.K + 1, 0), pos;
      for (int k = 0; k < p.K; ++k) {
        const int s0 = p.seg_start[k], n1 = p.po[k] + 1;
        const int64_t cW = colT0(L) + 2 + na + k;
        off[(size_t)k] = (int64_t)pos.size();
        for (int sec = 0; sec < 3; ++sec)
          for (int b = 0; b < ny; ++b)
            for (int j = 0; j < n1; ++j)
              pos.push_back(find_pos(sec == 0 ? cW : colT0(L) + (sec - 1), colv(L, b, s0 + j)));
        for (auto& ab : rh)
          for (int j = 0; j < n1; ++j) {
            if (ab.first < ny) pos.push_back(find_pos(colv(L, ab.first, s0 + j), colv(L, ab.second, s0)));
            else if (ab.second < ny) pos.push_back(find_pos(colA(L, ab.first - ny), colv(L, ab.second, s0 + j)));
            else pos.push_back(-1);
          }
        for (int m = 0; m < na; ++m) pos.push_back(find_pos(cW, colA(L, m)));
        pos.push_back(find_pos(cW, colT0(L))), pos.push_back(find_pos(cW, colT0(L) + 1)), pos.push_back(find_pos(cW, cW));
      }
      off[(size_t)p.K] = (int64_t)pos.size();
      CUDA_TRY(upload(p.hess_ph[ph].ah_pos, pos.data(), pos.size() * sizeof(int64_t)));
      CUDA_TRY(upload(p.hess_ph[ph].ah_off, off.data(), off.size() * sizeof(int64_t)));
      // staging of the node-local entries under the block diagonals: only if every block pair is also a pair of the
      // node Lagrangian (it is, by construction: the Lagrangian contains every f_s with its own multiplier)
      bool stage = !(getenv("MPX_AHESS_STAGE") && atoi(getenv("MPX_AHESS_STAGE")) == 0) && !p.hess_ph[ph].lin.rows &&
                   p.args[ph].ad_res;  // without the residual rows nothing would pick the staged entries up
      size_t n_blocks = 0;
      for (auto& ab : rh) {
        if (ab.first < ny && !L.pat_hw[(size_t)ab.first * NW + ab.second]) stage = false;
        ++n_blocks;
      }
      if (stage && n_blocks) CUDA_TRY(p.hess_ph[ph].hnl.ensure(n_blocks * (size_t)N * sizeof(double)));
      p.hess_ph[ph].ah_persist = !(getenv("MPX_AHESS_PERSIST") && atoi(getenv("MPX_AHESS_PERSIST")) == 0);
      std::vector<unsigned int> sync0((size_t)p.K + 3, 0u);
      sync0[2] = 1u;  // epoch of the first launch (flags start at 0)
      CUDA_TRY(upload(p.hess_ph[ph].ah_sync, sync0.data(), sync0.size() * sizeof(unsigned int)));
    }
    CUDA_TRY(p.d_hpart2.ensure((size_t)p.K * (3 + 2 * na + na * (na + 1) / 2) * sizeof(double)));
  }
  CUDA_TRY(p.d_lam.ensure((size_t)p.n_g * sizeof(double)));
  CUDA_TRY(p.d_hvals.ensure(p.h_colind.size() * sizeof(double)));
  p.hess_built = true;
  return MPX_OK;
}

int launch_hess(mpx_plan& p, const double* d_z, const double* d_p, double lam_f, const double* d_lam, double* d_vals,
                cudaStream_t st) {
  bool need_sig = false;  // sigma only enters through explicit time dependence
  for (auto& L : p.ph) need_sig |= L.uses_t || L.cost_t;
  if (need_sig) scan_widths(p, d_z, d_p, st);
  // The widths' contributions are ADDED on top of the node-local part (mpx_adapt_hess_kernel), so the buffer used to be
  // zero-filled first: 179 MB = 28 us at the headline size, although every entry of the pattern normally has a writer
  // (plain store) ahead of whatever is added to it.  Probed once per plan: the first evaluation runs on a buffer filled
  // with NaN; if no NaN survives, nothing leans on the initial content and the fill is skipped from then on.
  bool probing = false;
  if (p.adaptive) {
    if (p.ah_zero_fill < 0) {
      const char* ze = getenv("MPX_AHESS_ZERO");
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      CUDA_TRY(cudaStreamIsCapturing(st, &cs));
      if (ze && atoi(ze) != 0) p.ah_zero_fill = 1;
      else if (cs == cudaStreamCaptureStatusNone) probing = true;
    }
    if (probing) CUDA_TRY(cudaMemsetAsync(d_vals, 0xFF, p.h_colind.size() * sizeof(double), st));
    else if (p.ah_zero_fill != 0) CUDA_TRY(cudaMemsetAsync(d_vals, 0, p.h_colind.size() * sizeof(double), st));
  }
  for (int ph = 0; ph < p.P; ++ph) {
    MpxPhaseArgs a = p.args[ph];
    auto& H = p.hess_ph[ph];
    a.ticket = p.d_ticket.p ? p.d_ticket.as<unsigned int>() + p.P + ph : nullptr;
    a.z = d_z, a.w = widths_of(p, d_z, d_p, ph), a.sig0 = p.d_sig0.as<double>() + (int64_t)ph * p.K;
    a.lam = d_lam, a.lam_f = lam_f, a.node_seg = p.d_node_seg.as<int32_t>();
    a.hp_yy = H.pos_yy.as<int64_t>(), a.hp_ay = H.pos_ay.as<int64_t>(), a.hp_ty = H.pos_ty.as<int64_t>();
    a.hp_corner = H.pos_corner.as<int64_t>(), a.hp_term = H.pos_term.as<int64_t>();
    a.hp_term_assign = H.term_assign.as<int32_t>();
    a.hvals = d_vals, a.hpart = H.part.as<double>(), a.h_blocks = H.blocks;
    a.hnl = p.adaptive && H.hnl.p ? H.hnl.as<double>() : nullptr;
    a.trace = p.d_trace.p ? p.d_trace.as<unsigned long long>() : nullptr;  // MPX_TRACE=1: per-warp timeline stamps
    CUDA_TRY(p.prog->phases[ph]->hess(a, H.lin, H.blocks, st));
    p.launches += a.ticket ? 1 : 2;
    if (p.adaptive) {
      a.hpart2 = p.d_hpart2.as<double>();
      a.ah_pos = H.ah_pos.as<int64_t>(), a.ah_off = H.ah_off.as<int64_t>();
      // MPX_TRACE=1: per-segment timeline behind the records of mpx_hess_kernel (one per warp of nodes)
      const size_t trace_cap = (size_t)MPX_TRACE_RING * p.v2_grid * p.v2_warps, trace_base = (size_t)p.N / 32 + 1;
      a.ah_trace = (p.d_trace.p && trace_base + (size_t)p.K <= trace_cap)
                       ? p.d_trace.as<unsigned long long>() + trace_base * MPX_TRACE_SLOTS : nullptr;
      const int dmax = *std::max_element(p.po.begin(), p.po.end());
      if (H.ah_persist) {  // one launch: persistent CTAs, even segments first, flags at the shared nodes (see the kernel)
        a.ah_parity = -1;
        a.ah_sync = H.ah_sync.as<unsigned int>();
        if (p.num_sms <= 0) CUDA_TRY(cudaDeviceGetAttribute(&p.num_sms, cudaDevAttrMultiProcessorCount, p.device));
        CUDA_TRY(p.prog->phases[ph]->adapt_hess(a, std::min(p.K, 4 * std::max(1, p.num_sms)), dmax, st));
        ++p.launches;
      } else {
        for (int par = 0; par < 2; ++par) {
          const int grid = (p.K - par + 1) / 2;
          if (grid <= 0) continue;
          a.ah_parity = par;
          CUDA_TRY(p.prog->phases[ph]->adapt_hess(a, grid, dmax, st));
          ++p.launches;
        }
      }
      CUDA_TRY(p.prog->phases[ph]->adapt_hess_final(a, st));
      ++p.launches;
    }
  }
  if (probing) {
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(p.d_hpart2.p);  // free once the corner sum has run
    CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), st));
    mpx_count_nan_kernel<<<592, 256, 0, st>>>(d_vals, (int64_t)p.h_colind.size(), cnt);
    CUDA_TRY(cudaGetLastError());
    unsigned long long h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    p.ah_zero_fill = h ? 1 : 0;
    return launch_hess(p, d_z, d_p, lam_f, d_lam, d_vals, st);  // the evaluation proper
  }
  return MPX_OK;
}
}  // namespace

extern "C" int mpx_hess_structure(mpx_plan* p, int64_t* nnz, int64_t* rowptr, int64_t* colind) {
  if (!p) return fail(MPX_EINVAL, "NULL plan");
  int rc = build_hessian(*p);
  if (rc) return rc;
  if (nnz) *nnz = (int64_t)p->h_colind.size();
  if (rowptr) memcpy(rowptr, p->h_rowptr.data(), p->h_rowptr.size() * sizeof(int64_t));
  if (colind) memcpy(colind, p->h_colind.data(), p->h_colind.size() * sizeof(int64_t));
  return MPX_OK;
}

extern "C" int mpx_eval_hess_l_dev(mpx_plan* p, const double* d_z, const double* d_p, double lam_f, const double* d_lam_g,
                                   double* d_values, void* stream) {
  if (!p || !d_z || (!d_p && p->n_p) || !d_lam_g || !d_values) return fail(MPX_EINVAL, "NULL argument");
  int rc = build_hessian(*p);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(p->device));
  return launch_hess(*p, d_z, d_p, lam_f, d_lam_g, d_values, stream ? static_cast<cudaStream_t>(stream) : p->stream);
}

extern "C" int mpx_eval_hess_l(mpx_plan* p, const double* z, const double* pw, double lam_f, const double* lam_g,
                               double* values) {
  if (!p || !lam_g || !values) return fail(MPX_EINVAL, "NULL argument");
  int rc = build_hessian(*p);
  if (rc) return rc;
  rc = upload_inputs(*p, z, pw);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(p->d_lam.p, lam_g, (size_t)p->n_g * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  rc = launch_hess(*p, p->d_z.as<double>(), p->d_p.as<double>(), lam_f, p->d_lam.as<double>(), p->d_hvals.as<double>(), p->stream);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(values, p->d_hvals.p, p->h_colind.size() * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MPX_OK;
}

// Hessian at the x of the last mpx_stage (already on the device): no upload of x, staged results stay valid
extern "C" int mpx_hess_l_staged(mpx_plan* p, double lam_f, const double* lam_g, double* values) {
  if (!p || !lam_g || !values) return fail(MPX_EINVAL, "NULL argument");
  if (!p->staged) return fail(MPX_EINVAL, "mpx_hess_l_staged: nothing is staged (call mpx_stage first)");
  int rc = build_hessian(*p);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(p->device));
  CUDA_TRY(cudaMemcpyAsync(p->d_lam.p, lam_g, (size_t)p->n_g * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  rc = launch_hess(*p, p->d_z.as<double>(), p->d_p.as<double>(), lam_f, p->d_lam.as<double>(), p->d_hvals.as<double>(), p->stream);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(values, p->d_hvals.p, p->h_colind.size() * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MPX_OK;
}

// ------------------------------------------------------------------ interpolation / dynamics residual at arbitrary points
static int eval_points(mpx_plan* p, const double* z, const double* pw, int32_t phase, int64_t n_points, const int32_t* seg,
                       const double* taus, double* xi, double* ui, double* ti, double* dxi, double* dui, double* res,
                       double* ddxi, double* ddui, bool force_deriv = false) {
  if (!p) return fail(MPX_EINVAL, "NULL plan");
  if (phase < 0 || phase >= p->P) return fail(MPX_EINVAL, "phase out of range");
  if (n_points < 0 || (n_points && (!seg || !taus))) return fail(MPX_EINVAL, "bad point list");
  if (p->seg_begin != 0 || p->seg_end != p->K) return fail(MPX_EINVAL, "residuals need a plan over all segments");
  for (int64_t i = 0; i < n_points; ++i)
    if (seg[i] < 0 || seg[i] >= p->K) return fail(MPX_EINVAL, "segment index out of range in the point list");
  int rc = upload_inputs(*p, z, pw);
  if (rc || n_points == 0) return rc;
  const int nx = p->nx, nu = p->nu;
  const bool deriv = dxi || dui || res || ddxi || ddui || force_deriv;
  DevBuf &dseg = p->d_rseg, &dtau = p->d_rtau, &dout = p->d_rout;  // plan-owned, grow-only: the h-adaptive loop calls this every pass
  const size_t per = (size_t)(2 * nx + 2 * nu + 1 + nx) + (size_t)(nx + nu);
  CUDA_TRY(dseg.ensure((size_t)n_points * sizeof(int32_t)));
  CUDA_TRY(dtau.ensure((size_t)n_points * sizeof(double)));
  CUDA_TRY(dout.ensure((size_t)n_points * per * sizeof(double)));
  CUDA_TRY(cudaMemcpyAsync(dseg.p, seg, (size_t)n_points * sizeof(int32_t), cudaMemcpyHostToDevice, p->stream));
  CUDA_TRY(cudaMemcpyAsync(dtau.p, taus, (size_t)n_points * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  scan_widths(*p, p->d_z.as<double>(), p->d_p.as<double>(), p->stream);
  MpxPhaseArgs a = p->args[phase];
  a.z = p->d_z.as<double>(), a.w = widths_of(*p, p->d_z.as<double>(), p->d_p.as<double>(), phase);
  a.sig0 = p->d_sig0.as<double>() + (int64_t)phase * p->K;
  a.pt_seg = dseg.as<int32_t>(), a.pt_tau = dtau.as<double>(), a.n_points = n_points;
  double* o = dout.as<double>();
  a.r_xi = o, o += n_points * nx;
  a.r_ui = o, o += n_points * nu;
  a.r_ti = o, o += n_points;
  a.r_dxi = o, o += n_points * nx;
  a.r_dui = o, o += n_points * nu;
  a.r_res = o, o += n_points * nx;
  a.r_ddxi = ddxi ? o : nullptr, o += n_points * nx;
  a.r_ddui = ddui ? o : nullptr;
  CUDA_TRY(p->prog->phases[phase]->residual(a, deriv, (int)((n_points + 127) / 128), p->stream));
  ++p->launches;
  auto back = [&](double* dst, const double* src, size_t n) -> cudaError_t {
    return dst && n ? cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, p->stream) : cudaSuccess;
  };
  CUDA_TRY(back(xi, a.r_xi, (size_t)n_points * nx));
  CUDA_TRY(back(ui, a.r_ui, (size_t)n_points * nu));
  CUDA_TRY(back(ti, a.r_ti, (size_t)n_points));
  CUDA_TRY(back(dxi, a.r_dxi, (size_t)n_points * nx));
  CUDA_TRY(back(dui, a.r_dui, (size_t)n_points * nu));
  CUDA_TRY(back(res, a.r_res, (size_t)n_points * nx));
  if (ddxi) CUDA_TRY(back(ddxi, a.r_ddxi, (size_t)n_points * nx));
  if (ddui) CUDA_TRY(back(ddui, a.r_ddui, (size_t)n_points * nu));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MPX_OK;
}

extern "C" int mpx_eval_residuals(mpx_plan* p, const double* z, const double* pw, int32_t phase, int64_t n_points,
                                  const int32_t* seg, const double* taus, double* xi, double* ui, double* ti, double* dxi,
                                  double* dui, double* res) {
  return eval_points(p, z, pw, phase, n_points, seg, taus, xi, ui, ti, dxi, dui, res, nullptr, nullptr);
}

extern "C" int mpx_eval_second_derivatives(mpx_plan* p, const double* z, const double* pw, int32_t phase, int64_t n_points,
                                           const int32_t* seg, const double* taus, double* ti, double* ddxi, double* ddui) {
  if (!ddxi && !ddui) return fail(MPX_EINVAL, "ddxi and ddui are both NULL");
  return eval_points(p, z, pw, phase, n_points, seg, taus, nullptr, nullptr, ti, nullptr, nullptr, nullptr, ddxi, ddui);
}

extern "C" int mpx_eval_state_residuals(mpx_plan* p, const double* z, const double* pw, int32_t phase, int64_t n_points,
                                        const int32_t* seg, const double* taus, double* xint, double* ui, double* ti,
                                        double* res_x) {
  if (!p) return fail(MPX_EINVAL, "NULL plan");
  if (!xint && !res_x) return fail(MPX_EINVAL, "xint and res_x are both NULL");
  for (int64_t i = 1; i < n_points; ++i)
    if (seg && seg[i] < seg[i - 1]) return fail(MPX_EINVAL, "points must be listed segment by segment");
  // interpolation + dynamics residual of every point stay on the device (same buffer layout as eval_points)
  int rc = eval_points(p, z, pw, phase, n_points, seg, taus, nullptr, ui, ti, nullptr, nullptr, nullptr, nullptr, nullptr, true);
  if (rc || n_points == 0) return rc;
  const int nx = p->nx, nu = p->nu, K = p->K;
  std::vector<int32_t> off((size_t)K + 1, 0);
  for (int64_t i = 0; i < n_points; ++i) ++off[(size_t)seg[i] + 1];
  int max_pts = 0;
  for (int k = 0; k < K; ++k) max_pts = std::max(max_pts, off[k + 1]), off[k + 1] += off[k];
  CUDA_TRY(p->d_sr_off.ensure(off.size() * sizeof(int32_t)));
  CUDA_TRY(p->d_sr_out.ensure((size_t)2 * n_points * nx * sizeof(double)));
  CUDA_TRY(cudaMemcpyAsync(p->d_sr_off.p, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, p->stream));
  const double* o = p->d_rout.as<double>();
  MpxSrArgs a;
  a.z = p->d_z.as<double>() + p->ph[phase].zoff, a.seg_start = p->d_seg_start.as<int32_t>();
  a.pt_off = p->d_sr_off.as<int32_t>(), a.tau = p->d_rtau.as<double>();
  a.xi = o, a.dxi = o + n_points * (nx + nu + 1), a.res = a.dxi + n_points * (nx + nu);
  a.xint = p->d_sr_out.as<double>(), a.rx = a.xint + n_points * nx;
  a.N = p->N, a.nx = nx, a.max_pts = max_pts, a.tau0 = p->tau_min;
  const size_t smem = (size_t)(4 * max_pts + max_pts * nx) * sizeof(double);
  if (smem > 200 * 1024) return fail(MPX_ELIMIT, "too many target points in one segment for the state-residual kernel");
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(mpx_state_resid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mpx_state_resid_kernel<<<K, MPX_SR_THREADS, smem, p->stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  ++p->launches;
  if (xint) CUDA_TRY(cudaMemcpyAsync(xint, a.xint, (size_t)n_points * nx * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (res_x) CUDA_TRY(cudaMemcpyAsync(res_x, a.rx, (size_t)n_points * nx * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MPX_OK;
}

// ------------------------------------------------------------------ staged evaluation (one upload, one fused
// evaluation per distinct x; the callers -- IPOPT / CasADi shims -- fetch the pieces they are asked for)
extern "C" int mpx_stage(mpx_plan* p, const double* z, const double* pw, int32_t what) {
  if (!p) return fail(MPX_EINVAL, "NULL plan");
  if (what & ~(MPX_STAGE_F | MPX_STAGE_GRAD | MPX_STAGE_G | MPX_STAGE_JAC)) return fail(MPX_EINVAL, "unknown MPX_STAGE_* bit");
  p->staged = 0;
  int rc = upload_inputs(*p, z, pw);
  if (rc) return rc;
  if (what & (MPX_STAGE_G | MPX_STAGE_JAC)) {
    rc = launch_g_jac(*p, p->d_z.as<double>(), p->d_p.as<double>(), p->d_g.as<double>(),
                      (what & MPX_STAGE_JAC) ? p->d_vals.as<double>() : nullptr, p->stream);
    if (rc) return rc;
  }
  if (what & (MPX_STAGE_F | MPX_STAGE_GRAD)) {
    rc = launch_f_grad(*p, p->d_z.as<double>(), p->d_p.as<double>(), p->d_f.as<double>(),
                       (what & MPX_STAGE_GRAD) ? p->d_grad.as<double>() : nullptr, p->stream);
    if (rc) return rc;
  }
  p->staged = what | ((what & MPX_STAGE_JAC) ? MPX_STAGE_G : 0) | ((what & MPX_STAGE_GRAD) ? MPX_STAGE_F : 0);
  return MPX_OK;
}

extern "C" int mpx_staged(const mpx_plan* p) { return p ? p->staged : 0; }

extern "C" int mpx_fetch(mpx_plan* p, int32_t what, double* out) {
  if (!p || !out) return fail(MPX_EINVAL, "NULL argument");
  const int base = what == MPX_FETCH_JAC_CCS ? MPX_STAGE_JAC : what;
  if (base != MPX_STAGE_F && base != MPX_STAGE_GRAD && base != MPX_STAGE_G && base != MPX_STAGE_JAC)
    return fail(MPX_EINVAL, "mpx_fetch takes exactly one MPX_STAGE_* bit (or MPX_FETCH_JAC_CCS)");
  if (!(p->staged & base)) return fail(MPX_EINVAL, "mpx_fetch: that result has not been staged for the current x");
  CUDA_TRY(cudaSetDevice(p->device));
  const void* src = nullptr;
  size_t n = 0;
  if (what == MPX_STAGE_F) src = p->d_f.p, n = 1;
  else if (what == MPX_STAGE_GRAD) src = p->d_grad.p, n = (size_t)p->n_z;
  else if (what == MPX_STAGE_G) src = p->d_g.p, n = (size_t)p->n_g;
  else if (what == MPX_STAGE_JAC) src = p->d_vals.p, n = (size_t)p->nnz;
  else {  // CasADi's column-compressed order: gather on the device through the static permutation
    if (!p->d_ccs_perm.p) {
      std::vector<int64_t> perm((size_t)p->nnz);
      int rc = mpx_jac_structure_ccs(p, nullptr, nullptr, perm.data());
      if (rc) return rc;
      CUDA_TRY(p->d_ccs_perm.ensure(perm.size() * sizeof(int64_t)));
      CUDA_TRY(p->d_ccs_vals.ensure(perm.size() * sizeof(double)));
      CUDA_TRY(cudaMemcpy(p->d_ccs_perm.p, perm.data(), perm.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    }
    mpx_compact_kernel<<<(unsigned)((p->nnz + 255) / 256), 256, 0, p->stream>>>(p->d_vals.as<double>(), p->d_ccs_perm.as<int64_t>(),
                                                                              p->d_ccs_vals.as<double>(), p->nnz);
    CUDA_TRY(cudaGetLastError());
    ++p->launches;
    src = p->d_ccs_vals.p, n = (size_t)p->nnz;
  }
  CUDA_TRY(cudaMemcpyAsync(out, src, n * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MPX_OK;
}

extern "C" int mpx_eval_g_jac_dev(mpx_plan* p, const double* d_z, const double* d_p, double* d_g, double* d_values,
                                  void* stream) {
  if (!p || !d_z || (!d_p && p->n_p) || !d_g) return fail(MPX_EINVAL, "NULL argument");
  CUDA_TRY(use_plan_device(*p));
  return launch_g_jac(*p, d_z, d_p, d_g, d_values, stream ? (cudaStream_t)stream : p->stream);
}

// ------------------------------------------------------------------ fused evaluation + all-gather over peer memory
// Every store of the g + jac_g kernel is issued once per destination: this GPU's buffers and the same offsets of each
// peer's buffers (device pointers of other GPUs' allocations, opened through CUDA IPC and reached over NVLink).  When
// all ranks have run their shard this way every rank holds the whole g / Jacobian without a separate collective;
// the caller orders the ranks (a barrier or a tiny all-reduce) before it reads.
extern "C" int mpx_peer_alloc(int32_t device, int64_t bytes, void** dptr, mpx_ipc_handle* handle) {
  if (!dptr || !handle || bytes <= 0) return fail(MPX_EINVAL, "bad mpx_peer_alloc arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(mpx_ipc_handle), "mpx_ipc_handle too small");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaMalloc(dptr, (size_t)bytes));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, *dptr));
  memset(handle, 0, sizeof *handle);
  memcpy(handle, &h, sizeof h);
  return MPX_OK;
}
extern "C" int mpx_peer_open(int32_t device, const mpx_ipc_handle* handle, void** dptr) {
  if (!dptr || !handle) return fail(MPX_EINVAL, "bad mpx_peer_open arguments");
  CUDA_TRY(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  CUDA_TRY(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MPX_OK;
}
extern "C" int mpx_peer_close(void* dptr) {
  if (dptr) CUDA_TRY(cudaIpcCloseMemHandle(dptr));
  return MPX_OK;
}
extern "C" int mpx_peer_free(void* dptr) {
  if (dptr) CUDA_TRY(cudaFree(dptr));
  return MPX_OK;
}

extern "C" int mpx_eval_g_jac_dev_peers(mpx_plan* p, const double* d_z, const double* d_p, double* d_g, double* d_values,
                                        int32_t n_peers, double* const* peer_g, double* const* peer_values, void* stream) {
  if (!p || !d_z || !d_p || !d_g || !d_values) return fail(MPX_EINVAL, "NULL device pointer");
  if (p->adaptive) return fail(MPX_EINVAL, "peer replication is not available for the adaptive NLP");
  if (n_peers < 0 || n_peers > MPX_MAX_PEERS || (n_peers && (!peer_g || !peer_values)))
    return fail(MPX_EINVAL, "n_peers must be in [0, 7] with both pointer arrays given");
  if (p->v4 || p->v2_warps == 0 || !p->gather.empty())
    return fail(MPX_ELIMIT, "peer stores need the default g + jac kernel and an unmasked Jacobian pattern");
  CUDA_TRY(cudaSetDevice(p->device));
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : p->stream;
  for (auto& a : p->args) {
    a.n_peers = n_peers;
    for (int r = 0; r < n_peers; ++r) a.peer_g[r] = peer_g[r], a.peer_vals[r] = peer_values[r];
  }
  int rc = launch_g_jac(*p, d_z, d_p, d_g, d_values, st);
  for (auto& a : p->args) a.n_peers = 0;
  if (rc) return rc;
  // rows written by the small tail kernels (slope continuity, phase links): a few values, copied peer by peer
  bool tails = !p->links.empty() && p->seg_end == p->K;
  for (auto& L : p->ph) tails |= L.has_dU;
  if (tails && n_peers) {
    for (int kind = 0; kind < 2; ++kind) {
      std::vector<int64_t>& runs = p->h_tail_runs[kind];
      if (runs.empty()) {
        const int kb = p->seg_begin, ke = std::min(p->seg_end, p->K - 1);
        for (auto& L : p->ph)
          if (L.has_dU && ke > kb)
            for (int c = 0; c < p->nu; ++c) {
              if (kind == 0) runs.push_back(L.gdU + (int64_t)c * (p->K - 1) + kb), runs.push_back(ke - kb);
              else {
                int64_t sb = 0, se = 0;
                for (int k = 0; k < ke; ++k) {
                  const int64_t ss = p->po[k] + p->po[k + 1] + 1;
                  if (k < kb) sb += ss;
                  se += ss;
                }
                runs.push_back(L.vdU + (int64_t)c * p->nnzS + sb), runs.push_back(se - sb);
              }
            }
        if (!p->links.empty() && p->seg_end == p->K) {
          if (kind == 0) runs.push_back(p->g_events), runs.push_back(p->n_g - p->g_events);
          else runs.push_back(p->v_events), runs.push_back(p->nnz_full - p->v_events);
        }
        if (runs.empty()) runs.push_back(0), runs.push_back(0);
      }
      for (size_t i = 0; i + 1 < runs.size(); i += 2)
        for (int r = 0; r < n_peers && runs[i + 1] > 0; ++r) {
          double* dst = (kind == 0 ? peer_g[r] : peer_values[r]) + runs[i];
          const double* src = (kind == 0 ? d_g : d_values) + runs[i];
          CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)runs[i + 1] * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
    }
  }
  return MPX_OK;
}

extern "C" int mpx_eval_f_grad_dev(mpx_plan* p, const double* d_z, const double* d_p, double* d_f, double* d_grad,
                                   void* stream) {
  if (!p || !d_z || (!d_p && p->n_p) || !d_f) return fail(MPX_EINVAL, "NULL argument");
  CUDA_TRY(use_plan_device(*p));
  return launch_f_grad(*p, d_z, d_p, d_f, d_grad, stream ? (cudaStream_t)stream : p->stream);
}

extern "C" int mpx_sync(mpx_plan* p) {
  if (!p) return fail(MPX_EINVAL, "NULL plan");
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MPX_OK;
}

extern "C" int64_t mpx_launch_count(const mpx_plan* p) { return p ? p->launches : 0; }
extern "C" const char* mpx_program_origin(const mpx_plan* p) { return p ? p->origin.c_str() : ""; }
