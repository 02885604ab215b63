// mpx_kernels.cuh -- hand-written sm_100a kernels of the collocation hot path.
//
// One CTA per collocation segment.  The kernels are templates over a generated
// "phase functor" PH (mpopt_b200/program.py) that evaluates the user's dynamics /
// path / cost / terminal functions and their packed partials at ONE node; everything
// else -- table staging (1-D bulk TMA), D.X defects, row layout, CSR assembly in shared
// memory and the coalesced write-out -- is here.
//
// Replaces, per evaluation, what CasADi's SX virtual machine does for the reference
// inside mpopt.solve (/root/reference/mpopt/mpopt.py:804) on the functions derived at
// :757: nlp_g + nlp_jac_g (mpx_gjac_kernel) and nlp_f + nlp_grad_f (mpx_fgrad_kernel).
// The transcription being evaluated is mpopt.py:154-462.
//
// Compiles under nvcc (AOT, csrc/mpx_aot_*.cu) and under NVRTC (csrc/mpx_plan.cu), so
// it includes no host headers.
#pragma once

#ifdef __CUDACC_RTC__
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
#else
#include <stdint.h>
#endif

#define MPX_HD __host__ __device__ __forceinline__
#define MPX_MAXS 16
#define MPX_THREADS 128

// flags
#define MPX_F_DU 1     // control-slope rows present
#define MPX_F_MU 2     // mid-point control rows present
#define MPX_F_TAIL 4   // this launch also evaluates the terminal rows (last segment's CTA)

struct MpxPhaseArgs {
  const double* z;          // whole decision vector
  const double* w;          // segment widths of this phase [K]
  const double* sig0;       // sum_{m<k} w_m per segment (read only by time-dependent functors)
  const double* tabs;       // packed table records, see MpxTab
  const int32_t* seg_tab;   // [K] record offset (doubles) per segment; unused when uniform_deg > 0
  const int32_t* seg_start; // [K+1] first node of each segment
  const int64_t* seg_dpre;  // [K] D-nonzeros in the F rows before the segment's first owned row
  const int64_t* seg_ipre;  // [K] mid-point nonzeros before the segment
  double* g;
  double* vals;
  double* grad;             // fgrad only
  double* partial;          // fgrad only: [K][MPX_NPART] per-segment partial sums
  double* fout;             // fgrad final: objective
  const int32_t* unit_k;    // v2: first segment of each work unit (adjacent segments of equal degree)
  const int32_t* unit_n;    // v2: number of segments in the unit (<= 32 / pow2ceil(degree+1))
  int32_t K, N, seg_begin, seg_end;
  int32_t uniform_deg;
  int32_t flags;
  int32_t accumulate_f;     // fgrad final: add to *fout instead of overwriting (phases > 0)
  int32_t n_units;          // v2
  int32_t tab_doubles;      // v2: size of all table records (copied to shared memory once per CTA)
  int32_t stage_cap;        // v2: staging doubles per warp
  int64_t zoff;             // offset of this phase in z
  int64_t gF, gC, gDU, gmU, gTC;       // first row of each block
  int64_t vF[MPX_MAXS], vC[MPX_MAXS];  // value offset of the first F row of state s / path row q
  int64_t vDU, vmU, vTC;
  int64_t nnzD, nnzI;                  // D / mid-point nonzeros per state (per control)
  double sx[MPX_MAXS], isx[MPX_MAXS], isu[MPX_MAXS], isa[MPX_MAXS];
  double st, delta, tau0;
};

// table record of one degree d (n1 = d+1), all sections padded to an even number of doubles so
// the whole record is one 16-byte-granular bulk copy:
//   roots[n1] | w[n1] | D[n1*n1] row-major | Cmid[d*n1] row-major | Dt[n1*n1] = D^T | Ct[n1*d] = Cmid^T
// (the transposed copies give conflict-free shared-memory reads when lane = row)
struct MpxTab {
  MPX_HD static int pad2(int n) { return (n + 1) & ~1; }
  MPX_HD static int off_roots(int) { return 0; }
  MPX_HD static int off_w(int n1) { return pad2(n1); }
  MPX_HD static int off_D(int n1) { return 2 * pad2(n1); }
  MPX_HD static int off_C(int n1) { return 2 * pad2(n1) + pad2(n1 * n1); }
  MPX_HD static int off_Dt(int n1) { return off_C(n1) + pad2((n1 - 1) * n1); }
  MPX_HD static int off_Ct(int n1) { return off_Dt(n1) + pad2(n1 * n1); }
  MPX_HD static int size(int n1) { return off_Ct(n1) + pad2((n1 - 1) * n1); }
};

#define MPX_NPART (3 + MPX_MAXS)  // J, dJ/dT0, dJ/dTF, dJ/da_m

#ifdef __CUDACC__

// ------------------------------------------------------------------ PTX helpers (mbarrier + bulk TMA)
__device__ __forceinline__ uint32_t mpx_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mpx_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mpx_smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mpx_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mpx_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mpx_tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   mpx_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(mpx_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mpx_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = mpx_smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

// streaming (evict-first) 16-byte store: the Jacobian is written once and never re-read here
__device__ __forceinline__ void mpx_st_cs_v2(double* p, double a, double b) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void mpx_st_cs(double* p, double a) {
  asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(a) : "memory");
}

// one warp copies n doubles from shared to global with 16-byte stores (8-byte peel at the ends)
__device__ __forceinline__ void mpx_warp_copy_out(double* __restrict__ dst, const double* __restrict__ src, int n,
                                                  int lane) {
  if (n <= 0) return;
  const int head = (int)((((uintptr_t)dst) >> 3) & 1);
  if (head) {
    if (lane == 0) mpx_st_cs(dst, src[0]);
    dst += 1, src += 1, n -= 1;
  }
  const int nv = n >> 1;
  if ((((uint32_t)mpx_smem_u32(src)) & 15u) == 0) {
    const double2* s2 = reinterpret_cast<const double2*>(src);
    for (int i = lane; i < nv; i += 32) {
      const double2 v = s2[i];
      mpx_st_cs_v2(dst + 2 * i, v.x, v.y);
    }
  } else {
    for (int i = lane; i < nv; i += 32) mpx_st_cs_v2(dst + 2 * i, src[2 * i], src[2 * i + 1]);
  }
  if ((n & 1) && lane == 0) mpx_st_cs(dst + n - 1, src[n - 1]);
}

template <class PH>
MPX_HD constexpr double mpx_dummy() { return 0.0; }

// 1/scale of node variable v (x.., u.., a..)
template <class PH>
__device__ __forceinline__ double mpx_iscale(const MpxPhaseArgs& A, int v) {
  return v < PH::NX ? A.isx[v] : (v < PH::NX + PH::NU ? A.isu[v - PH::NX] : A.isa[v - PH::NX - PH::NU]);
}
// 1/scale of terminal variable v (xf.., x0.., tf, t0, a..)
template <class PH>
__device__ __forceinline__ double mpx_iscale_term(const MpxPhaseArgs& A, int v) {
  return v < PH::NX ? A.isx[v]
                    : (v < 2 * PH::NX ? A.isx[v - PH::NX] : (v < 2 * PH::NX + 2 ? 1.0 / A.st : A.isa[v - 2 * PH::NX - 2]));
}

struct MpxSeg {
  int k, d, n1, s0, rb, nrow;
  int64_t dpre, ipre, rowpre;
  const double* tab;
};

__device__ __forceinline__ MpxSeg mpx_segment(const MpxPhaseArgs& A, int k) {
  MpxSeg S;
  S.k = k;
  if (A.uniform_deg > 0) {
    const int d = A.uniform_deg;
    S.d = d;
    S.s0 = k * d;
    S.dpre = k == 0 ? 0 : (int64_t)(d + 1) * (d + 1) + (int64_t)(k - 1) * d * (d + 1);
    S.ipre = (int64_t)k * d * (d + 1);
    S.tab = A.tabs;
  } else {
    S.s0 = A.seg_start[k];
    S.d = A.seg_start[k + 1] - S.s0;
    S.dpre = A.seg_dpre[k];
    S.ipre = A.seg_ipre[k];
    S.tab = A.tabs + A.seg_tab[k];
  }
  S.n1 = S.d + 1;
  S.rb = k == 0 ? 0 : 1;      // node ownership: a shared node belongs to the earlier segment (mpopt.py:189-195)
  S.nrow = S.n1 - S.rb;
  S.rowpre = k == 0 ? 0 : S.s0 + 1;
  return S;
}

// shared-memory doubles needed by mpx_gjac_kernel for degree d (host uses the same formula)
template <class PH>
MPX_HD int mpx_gjac_smem_doubles(int d, bool jac) {
  const int n1 = d + 1;
  int n = MpxTab::size(n1);                 // table record
  n += MpxTab::pad2((PH::NX + PH::NU) * n1);  // XU
  n += MpxTab::pad2(PH::NX * n1) * 2;         // DX, Fv
  if (jac) {
    int st = 0;
    for (int s = 0; s < PH::NX; ++s) st += n1 * (n1 + PH::f_next(s));
    for (int q = 0; q < PH::NC; ++q) st += n1 * PH::c_len(q);
    n += MpxTab::pad2(st);
  }
  return n + 2;  // mbarrier
}

// =====================================================================================
// K2: fused g + jac_g for one phase.  grid = segments of the shard, block = MPX_THREADS.
// =====================================================================================
template <class PH, bool JAC>
__global__ void __launch_bounds__(MPX_THREADS) mpx_gjac_kernel(const __grid_constant__ MpxPhaseArgs A) {
  extern __shared__ __align__(16) double smem[];
  constexpr int NX = PH::NX, NU = PH::NU, NA = PH::NA, NC = PH::NC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const MpxSeg S = mpx_segment(A, A.seg_begin + (int)blockIdx.x);
  const int n1 = S.n1, d = S.d, rb = S.rb, nrow = S.nrow;

  // ---- shared memory carve-up
  double* sTab = smem;
  const double* sRoots = sTab + MpxTab::off_roots(n1);
  const double* sD = sTab + MpxTab::off_D(n1);
  const double* sC = sTab + MpxTab::off_C(n1);
  double* sXU = sTab + MpxTab::size(n1);
  double* sDX = sXU + MpxTab::pad2((NX + NU) * n1);
  double* sFv = sDX + MpxTab::pad2(NX * n1);
  double* stage = sFv + MpxTab::pad2(NX * n1);
  int stage_doubles = 0;
  if (JAC) {
#pragma unroll
    for (int s = 0; s < NX; ++s) stage_doubles += nrow * (n1 + PH::f_next(s));
#pragma unroll
    for (int q = 0; q < NC; ++q) stage_doubles += nrow * PH::c_len(q);
  }
  uint64_t* bar = reinterpret_cast<uint64_t*>(stage + MpxTab::pad2(JAC ? stage_doubles : 0));

  // ---- stage the D / C / root tables of this degree through shared memory with one bulk TMA copy
  if (tid == 0) {
    mpx_mbar_init(bar, 1);
    const uint32_t bytes = (uint32_t)MpxTab::size(n1) * 8u;
    mpx_mbar_expect_tx(bar, bytes);
    mpx_tma_load_1d(sTab, S.tab, bytes, bar);
  }
  // ---- coalesced loads of the segment's slice of the state-major decision vector
  const double* zp = A.z + A.zoff;
  for (int i = tid; i < (NX + NU) * n1; i += MPX_THREADS) {
    const int v = i / n1, j = i - v * n1;
    sXU[i] = zp[(int64_t)v * A.N + S.s0 + j];
  }
  const int64_t zt = (int64_t)(NX + NU) * A.N;
  const double T0 = zp[zt], TF = zp[zt + 1];
  const double t0 = T0 / A.st, tf = TF / A.st;      // mpopt.py:175-176
  const double wk = A.w[S.k];
  const double h = (tf - t0) / A.delta * wk;        // mpopt.py:184
  const double gk = wk / (A.delta * A.st);          // d h / d TF  (= - d h / d T0)
  __syncthreads();       // sXU visible, barrier initialised
  mpx_mbar_wait(bar, 0); // tables landed

  // ---- per-node functor evaluation: one thread per owned node
  for (int r = tid + rb; r < n1; r += MPX_THREADS) {
    double x[NX > 0 ? NX : 1], u[NU > 0 ? NU : 1], a[NA > 0 ? NA : 1];
#pragma unroll
    for (int s = 0; s < NX; ++s) x[s] = sXU[s * n1 + r] * A.isx[s];
#pragma unroll
    for (int c = 0; c < NU; ++c) u[c] = sXU[(NX + c) * n1 + r] * A.isu[c];
#pragma unroll
    for (int m = 0; m < NA; ++m) a[m] = zp[zt + 2 + m] * A.isa[m];
    double sigma = 0.0, t = t0;
    if (PH::F_T || PH::C_T) {
      const double dt = sRoots[r] - A.tau0;
      sigma = A.sig0[S.k] + wk * dt / A.delta;
      t = (t0 + (tf - t0) * A.sig0[S.k]) + h * dt;  // mpopt.py:192, :198
    }
    const int row = r - rb;
    {
      double f[NX > 0 ? NX : 1], jf[PH::NJF > 0 ? PH::NJF : 1], ft[NX > 0 ? NX : 1];
      PH::dyn(x, u, t, a, f, jf, ft);
#pragma unroll
      for (int s = 0; s < NX; ++s) sFv[s * n1 + r] = h * A.sx[s] * f[s];  // mpopt.py:201
      if (JAC) {
        int off = 0;
#pragma unroll
        for (int s = 0; s < NX; ++s) {
          const int Ls = n1 + PH::f_next(s);
          double* rowp = stage + off + row * Ls;
#pragma unroll
          for (int e = 0; e < PH::NJF; ++e) {
            if (PH::jf_row(e) != s) continue;
            const double val = -h * (A.sx[s] * mpx_iscale<PH>(A, PH::jf_var(e))) * jf[e];
            const int pos = PH::jf_pos(e);
            if (pos < 0) rowp[PH::f_npre(s) + r] = sD[r * n1 + r] + val;   // merges with the D diagonal
            else rowp[pos < PH::f_npre(s) ? pos : pos + n1] = val;
          }
          if (PH::f_nz(s)) {
            const int tp = PH::f_tpos(s) + n1;
            const double a1 = gk * A.sx[s] * f[s];
            const double a2 = PH::F_T ? h * A.sx[s] * ft[s] / A.st : 0.0;
            rowp[tp] = a1 - a2 * (1.0 - sigma);   // d/dT0
            rowp[tp + 1] = -a1 - a2 * sigma;      // d/dTF
          }
          off += nrow * Ls;
        }
      }
    }
    if (NC > 0) {
      double c[NC > 0 ? NC : 1], jc[PH::NJC > 0 ? PH::NJC : 1], ct[NC > 0 ? NC : 1];
      PH::path(x, u, t, a, c, jc, ct);
#pragma unroll
      for (int q = 0; q < NC; ++q) A.g[A.gC + (int64_t)q * A.N + S.s0 + r] = c[q];  // mpopt.py:204
      if (JAC) {
        int off = 0;
#pragma unroll
        for (int s = 0; s < NX; ++s) off += nrow * (n1 + PH::f_next(s));
#pragma unroll
        for (int q = 0; q < NC; ++q) {
          const int Lq = PH::c_len(q);
          double* rowp = stage + off + row * Lq;
#pragma unroll
          for (int e = 0; e < PH::NJC; ++e) {
            if (PH::jc_row(e) != q) continue;
            rowp[PH::jc_pos(e)] = jc[e] * mpx_iscale<PH>(A, PH::jc_var(e));
          }
          if (PH::c_tpos(q) >= 0) {
            rowp[PH::c_tpos(q)] = ct[q] * (1.0 - sigma) / A.st;
            rowp[PH::c_tpos(q) + 1] = ct[q] * sigma / A.st;
          }
          off += nrow * Lq;
        }
      }
    }
  }

  // ---- constant D blocks into the staged rows (all threads)
  if (JAC) {
    int off = 0;
#pragma unroll
    for (int s = 0; s < NX; ++s) {
      const int Ls = n1 + PH::f_next(s);
      for (int i = tid; i < nrow * n1; i += MPX_THREADS) {
        const int row = i / n1, j = i - row * n1, r = row + rb;
        if (PH::f_diag(s) >= 0 && j == r) continue;  // written by the node thread
        stage[off + row * Ls + PH::f_npre(s) + j] = sD[r * n1 + j];
      }
      off += nrow * Ls;
    }
  }
  // ---- D.X defects; skewed column order keeps the shared-memory reads conflict-free
  for (int i = tid; i < NX * n1; i += MPX_THREADS) {
    const int s = i / n1, r = i - s * n1;
    double acc = 0.0;
    int j = r;
    for (int jj = 0; jj < n1; ++jj) {
      acc = fma(sD[r * n1 + j], sXU[s * n1 + j], acc);
      j = (j + 1 == n1) ? 0 : j + 1;
    }
    sDX[i] = acc;
  }
  // ---- mid-point control rows (mpopt.py:350-369) and control slope rows (:315-324): g values
  if (A.flags & MPX_F_MU) {
    for (int i = tid; i < NU * d; i += MPX_THREADS) {
      const int c = i / d, m = i - c * d;
      double acc = 0.0;
      int j = m;
      for (int jj = 0; jj < n1; ++jj) {
        acc = fma(sC[m * n1 + j], sXU[(NX + c) * n1 + j], acc);
        j = (j + 1 == n1) ? 0 : j + 1;
      }
      A.g[A.gmU + (int64_t)c * (A.N - 1) + S.s0 + m] = acc;
    }
  }
  if (A.flags & MPX_F_DU) {
    for (int i = tid; i < NU * n1; i += MPX_THREADS) {
      const int c = i / n1, r = i - c * n1;
      if (r < rb) continue;
      double acc = 0.0;
      int j = r;
      for (int jj = 0; jj < n1; ++jj) {
        acc = fma(sD[r * n1 + j], sXU[(NX + c) * n1 + j], acc);
        j = (j + 1 == n1) ? 0 : j + 1;
      }
      A.g[A.gDU + (int64_t)c * A.N + S.s0 + r] = acc;
    }
  }
  __syncthreads();

  // ---- defect rows  F = D.X - h Sx f   (mpopt.py:232)
  for (int i = tid; i < NX * n1; i += MPX_THREADS) {
    const int s = i / n1, r = i - s * n1;
    if (r >= rb) A.g[A.gF + (int64_t)s * A.N + S.s0 + r] = sDX[i] - sFv[i];
  }

  // ---- write the CSR value chunks: each (state, segment) block of rows is contiguous in CSR order
  if (JAC) {
    int chunk = 0, off = 0;
#pragma unroll
    for (int s = 0; s < NX; ++s) {
      const int n = nrow * (n1 + PH::f_next(s));
      if ((chunk++ & 3) == warp)
        mpx_warp_copy_out(A.vals + A.vF[s] + S.rowpre * PH::f_next(s) + S.dpre, stage + off, n, lane);
      off += n;
    }
#pragma unroll
    for (int q = 0; q < NC; ++q) {
      const int n = nrow * PH::c_len(q);
      if ((chunk++ & 3) == warp) mpx_warp_copy_out(A.vals + A.vC[q] + S.rowpre * PH::c_len(q), stage + off, n, lane);
      off += n;
    }
    if (A.flags & MPX_F_DU) {
      for (int c = 0; c < NU; ++c)
        if ((chunk++ & 3) == warp)
          mpx_warp_copy_out(A.vals + A.vDU + (int64_t)c * A.nnzD + S.dpre, sD + rb * n1, nrow * n1, lane);
    }
    if (A.flags & MPX_F_MU) {
      for (int c = 0; c < NU; ++c)
        if ((chunk++ & 3) == warp)
          mpx_warp_copy_out(A.vals + A.vmU + (int64_t)c * A.nnzI + S.ipre, sC, d * n1, lane);
    }
  }

  // ---- terminal constraint rows (mpopt.py:277-292), by the CTA that owns the last node
  if ((A.flags & MPX_F_TAIL) && PH::NTC > 0 && S.k == A.K - 1 && tid == 0) {
    double xf[NX > 0 ? NX : 1], x0[NX > 0 ? NX : 1], a[NA > 0 ? NA : 1];
#pragma unroll
    for (int s = 0; s < NX; ++s) {
      xf[s] = sXU[s * n1 + d] * A.isx[s];
      x0[s] = zp[(int64_t)s * A.N] * A.isx[s];
    }
#pragma unroll
    for (int m = 0; m < NA; ++m) a[m] = zp[zt + 2 + m] * A.isa[m];
    double tc[PH::NTC > 0 ? PH::NTC : 1], jtc[PH::NJTC > 0 ? PH::NJTC : 1], M[1], gm[PH::NGM > 0 ? PH::NGM : 1];
    PH::term(xf, tf, x0, t0, a, tc, jtc, M, gm);
    int off = 0;
#pragma unroll
    for (int r = 0; r < PH::NTC; ++r) {
      A.g[A.gTC + r] = tc[r];
      if (JAC) {
#pragma unroll
        for (int e = 0; e < PH::NJTC; ++e)
          if (PH::jtc_row(e) == r) A.vals[A.vTC + off + PH::jtc_pos(e)] = jtc[e] * mpx_iscale_term<PH>(A, PH::jtc_var(e));
      }
      off += PH::tc_len(r);
    }
  }
}

// =====================================================================================
// K2 (v2): persistent warps.  Each warp repeatedly takes a work unit = up to 32/LW adjacent
// segments of equal degree d (LW = pow2ceil(d+1)); lane = (segment q, local node r).  The
// tables of every degree stay resident in shared memory (one bulk TMA copy per CTA), the
// constant D blocks are pre-staged once per (degree) and only the z-dependent entries are
// rewritten per unit; the rows of a unit are contiguous in CSR order, so each state is one
// coalesced 16-byte-store stream.  No block-level barrier after start-up.
// =====================================================================================
#define MPX2_MAX_THREADS 256

template <class PH>
__device__ __forceinline__ int mpx2_region(int rows_cap, int n1, int s_end, int q_end) {
  int off = 0;
#pragma unroll
  for (int s = 0; s < PH::NX; ++s)
    if (s < s_end) off += rows_cap * (n1 + PH::f_next(s));
#pragma unroll
  for (int q = 0; q < PH::NC; ++q)
    if (q < q_end) off += rows_cap * PH::c_len(q);
  return off;
}

template <class PH, bool JAC>
__global__ void __launch_bounds__(MPX2_MAX_THREADS, 1) mpx_gjac2_kernel(const __grid_constant__ MpxPhaseArgs A) {
  extern __shared__ __align__(16) double smem[];
  constexpr int NX = PH::NX, NU = PH::NU, NA = PH::NA, NC = PH::NC, NV = PH::NX + PH::NU;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double* sTab = smem;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sTab + A.tab_doubles);
  double* wbase = sTab + A.tab_doubles + 2 + (size_t)warp * (32 * NV + (JAC ? A.stage_cap : 0));
  double* sXU = wbase;            // [q][v][LW]
  double* stage = wbase + 32 * NV;
  if (threadIdx.x == 0) {
    mpx_mbar_init(bar, 1);
    const uint32_t bytes = (uint32_t)A.tab_doubles * 8u;
    mpx_mbar_expect_tx(bar, bytes);
    mpx_tma_load_1d(sTab, A.tabs, bytes, bar);
  }
  const double* zp = A.z + A.zoff;
  const int64_t zt = (int64_t)NV * A.N;
  const double T0 = zp[zt], TF = zp[zt + 1];
  const double t0 = T0 / A.st, tf = TF / A.st;  // mpopt.py:175-176
  double a[NA > 0 ? NA : 1];
#pragma unroll
  for (int m = 0; m < NA; ++m) a[m] = zp[zt + 2 + m] * A.isa[m];
  __syncthreads();
  mpx_mbar_wait(bar, 0);

  int cur_sig = -1;
  const int tw = gridDim.x * nwarps;
  for (int u = blockIdx.x * nwarps + warp; u < A.n_units; u += tw) {
    const int k0 = A.unit_k[u], cnt = A.unit_n[u];
    const MpxSeg S0 = mpx_segment(A, k0);
    const double* tab = sTab + (S0.tab - A.tabs);
    const int d = S0.d, n1 = d + 1;
    const int shift = 32 - __clz(d);          // LW = pow2ceil(d+1)
    const int LW = 1 << shift, smax = 32 >> shift;
    const int q = lane >> shift, r = lane & (LW - 1);
    const int has0 = k0 == 0 ? 1 : 0;
    const bool inseg = q < cnt && r <= d;
    const int rb = (has0 && q == 0) ? 0 : 1;
    const bool owner = inseg && r >= rb;       // a shared node belongs to the earlier segment (mpopt.py:189-195)
    const int row = (q == 0 ? 0 : q * d + has0) + r - rb;
    const int rows_total = cnt * d + has0, rows_cap = smax * d + 1;
    const int s0q = S0.s0 + q * d;
    const double* sD = tab + MpxTab::off_D(n1);
    const double* sDt = tab + MpxTab::off_Dt(n1);

    // ---- constant D blocks: staged once per (degree, first-unit) signature
    if (JAC && cur_sig != d * 2 + has0) {
      __syncwarp();
#pragma unroll
      for (int s = 0; s < NX; ++s) {
        const int Ls = n1 + PH::f_next(s);
        double* reg = stage + mpx2_region<PH>(rows_cap, n1, s, 0) + PH::f_npre(s);
        for (int qq = 0; qq < smax; ++qq) {
          const int rbq = (has0 && qq == 0) ? 0 : 1;
          const int rowb = (qq == 0 ? 0 : qq * d + has0) - rbq;
          for (int rr = rbq + q; rr <= d; rr += smax)
            if (r <= d && !(PH::f_diag(s) >= 0 && r == rr)) reg[(rowb + rr) * Ls + r] = sD[rr * n1 + r];
        }
      }
      cur_sig = d * 2 + has0;
    }

    // ---- this lane's node: coalesced loads from the state-major decision vector
    double x[NX > 0 ? NX : 1], uu[NU > 0 ? NU : 1];
    if (inseg) {
#pragma unroll
      for (int s = 0; s < NX; ++s) {
        const double val = zp[(int64_t)s * A.N + s0q + r];
        sXU[((q * NV + s) << shift) + r] = val;
        x[s] = val * A.isx[s];
      }
#pragma unroll
      for (int c = 0; c < NU; ++c) {
        const double val = zp[(int64_t)(NX + c) * A.N + s0q + r];
        sXU[((q * NV + NX + c) << shift) + r] = val;
        uu[c] = val * A.isu[c];
      }
    }
    __syncwarp();
    const double* xq = sXU + ((q * NV) << shift);
    const int kk = k0 + (q < cnt ? q : 0);
    const double wk = A.w[kk];
    const double h = (tf - t0) / A.delta * wk;  // mpopt.py:184
    const double gk = wk / (A.delta * A.st);

    if (owner) {
      double sigma = 0.0, t = t0;
      if (PH::F_T || PH::C_T) {
        const double dt = tab[MpxTab::off_roots(n1) + r] - A.tau0;
        sigma = A.sig0[kk] + wk * dt / A.delta;
        t = (t0 + (tf - t0) * A.sig0[kk]) + h * dt;  // mpopt.py:192, :198
      }
      {
        double f[NX > 0 ? NX : 1], jf[PH::NJF > 0 ? PH::NJF : 1], ft[NX > 0 ? NX : 1];
        PH::dyn(x, uu, t, a, f, jf, ft);
        if (JAC) {
#pragma unroll
          for (int s = 0; s < NX; ++s) {
            const int Ls = n1 + PH::f_next(s);
            double* rowp = stage + mpx2_region<PH>(rows_cap, n1, s, 0) + row * Ls;
#pragma unroll
            for (int e = 0; e < PH::NJF; ++e) {
              if (PH::jf_row(e) != s) continue;
              const double val = -h * (A.sx[s] * mpx_iscale<PH>(A, PH::jf_var(e))) * jf[e];
              const int pos = PH::jf_pos(e);
              if (pos < 0) rowp[PH::f_npre(s) + r] = sD[r * n1 + r] + val;
              else rowp[pos < PH::f_npre(s) ? pos : pos + n1] = val;
            }
            if (PH::f_nz(s)) {
              const int tp = PH::f_tpos(s) + n1;
              const double a1 = gk * A.sx[s] * f[s];
              const double a2 = PH::F_T ? h * A.sx[s] * ft[s] / A.st : 0.0;
              rowp[tp] = a1 - a2 * (1.0 - sigma);
              rowp[tp + 1] = -a1 - a2 * sigma;
            }
          }
        }
        // defect rows F = D.X - h Sx f (mpopt.py:232); lane = row, D^T gives conflict-free reads
        double acc[NX > 0 ? NX : 1];
#pragma unroll
        for (int s = 0; s < NX; ++s) acc[s] = 0.0;
        for (int j = 0; j < n1; ++j) {
          const double dj = sDt[j * n1 + r];
#pragma unroll
          for (int s = 0; s < NX; ++s) acc[s] = fma(dj, xq[(s << shift) + j], acc[s]);
        }
#pragma unroll
        for (int s = 0; s < NX; ++s) A.g[A.gF + (int64_t)s * A.N + s0q + r] = acc[s] - h * A.sx[s] * f[s];
      }
      if (NC > 0) {
        double c[NC > 0 ? NC : 1], jc[PH::NJC > 0 ? PH::NJC : 1], ct[NC > 0 ? NC : 1];
        PH::path(x, uu, t, a, c, jc, ct);
#pragma unroll
        for (int qc = 0; qc < NC; ++qc) A.g[A.gC + (int64_t)qc * A.N + s0q + r] = c[qc];  // mpopt.py:204
        if (JAC) {
#pragma unroll
          for (int qc = 0; qc < NC; ++qc) {
            const int Lq = PH::c_len(qc);
            double* rowp = stage + mpx2_region<PH>(rows_cap, n1, NX, qc) + row * Lq;
#pragma unroll
            for (int e = 0; e < PH::NJC; ++e) {
              if (PH::jc_row(e) != qc) continue;
              rowp[PH::jc_pos(e)] = jc[e] * mpx_iscale<PH>(A, PH::jc_var(e));
            }
            if (PH::c_tpos(qc) >= 0) {
              rowp[PH::c_tpos(qc)] = ct[qc] * (1.0 - sigma) / A.st;
              rowp[PH::c_tpos(qc) + 1] = ct[qc] * sigma / A.st;
            }
          }
        }
      }
      if (A.flags & MPX_F_DU) {  // control slope rows (mpopt.py:315-324)
        double acc[NU > 0 ? NU : 1];
#pragma unroll
        for (int c = 0; c < NU; ++c) acc[c] = 0.0;
        for (int j = 0; j < n1; ++j) {
          const double dj = sDt[j * n1 + r];
#pragma unroll
          for (int c = 0; c < NU; ++c) acc[c] = fma(dj, xq[((NX + c) << shift) + j], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < NU; ++c) A.g[A.gDU + (int64_t)c * A.N + s0q + r] = acc[c];
      }
      // terminal constraint rows (mpopt.py:277-292) by the lane that owns the last node
      if ((A.flags & MPX_F_TAIL) && PH::NTC > 0 && kk == A.K - 1 && r == d) {
        double x0[NX > 0 ? NX : 1];
#pragma unroll
        for (int s = 0; s < NX; ++s) x0[s] = zp[(int64_t)s * A.N] * A.isx[s];
        double tc[PH::NTC > 0 ? PH::NTC : 1], jtc[PH::NJTC > 0 ? PH::NJTC : 1], M[1], gm[PH::NGM > 0 ? PH::NGM : 1];
        PH::term(x, tf, x0, t0, a, tc, jtc, M, gm);
        int off = 0;
#pragma unroll
        for (int rr = 0; rr < PH::NTC; ++rr) {
          A.g[A.gTC + rr] = tc[rr];
          if (JAC) {
#pragma unroll
            for (int e = 0; e < PH::NJTC; ++e)
              if (PH::jtc_row(e) == rr)
                A.vals[A.vTC + off + PH::jtc_pos(e)] = jtc[e] * mpx_iscale_term<PH>(A, PH::jtc_var(e));
          }
          off += PH::tc_len(rr);
        }
      }
    }
    // mid-point control rows (mpopt.py:350-369): lane r < d is mid-point r of its segment
    if ((A.flags & MPX_F_MU) && inseg && r < d) {
      const double* sCt = tab + MpxTab::off_Ct(n1);
      double acc[NU > 0 ? NU : 1];
#pragma unroll
      for (int c = 0; c < NU; ++c) acc[c] = 0.0;
      for (int j = 0; j < n1; ++j) {
        const double cj = sCt[j * d + r];
#pragma unroll
        for (int c = 0; c < NU; ++c) acc[c] = fma(cj, xq[((NX + c) << shift) + j], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < NU; ++c) A.g[A.gmU + (int64_t)c * (A.N - 1) + s0q + r] = acc[c];
    }

    // ---- stream the unit's CSR value blocks out: per state / path row they are contiguous
    if (JAC) {
      __syncwarp();
#pragma unroll
      for (int s = 0; s < NX; ++s)
        mpx_warp_copy_out(A.vals + A.vF[s] + S0.rowpre * PH::f_next(s) + S0.dpre,
                          stage + mpx2_region<PH>(rows_cap, n1, s, 0), rows_total * (n1 + PH::f_next(s)), lane);
#pragma unroll
      for (int qc = 0; qc < NC; ++qc)
        mpx_warp_copy_out(A.vals + A.vC[qc] + S0.rowpre * PH::c_len(qc), stage + mpx2_region<PH>(rows_cap, n1, NX, qc),
                          rows_total * PH::c_len(qc), lane);
      if (A.flags & MPX_F_DU) {
        for (int c = 0; c < NU; ++c)
          for (int qq = 0; qq < cnt; ++qq) {
            const int rbq = (has0 && qq == 0) ? 0 : 1;
            const int rowb = (qq == 0 ? 0 : qq * d + has0);
            mpx_warp_copy_out(A.vals + A.vDU + (int64_t)c * A.nnzD + S0.dpre + (int64_t)rowb * n1, sD + rbq * n1,
                              (n1 - rbq) * n1, lane);
          }
      }
      if (A.flags & MPX_F_MU) {
        const double* sC = tab + MpxTab::off_C(n1);
        for (int c = 0; c < NU; ++c)
          for (int qq = 0; qq < cnt; ++qq)
            mpx_warp_copy_out(A.vals + A.vmU + (int64_t)c * A.nnzI + S0.ipre + (int64_t)qq * d * n1, sC, d * n1, lane);
      }
      __syncwarp();
    }
  }
}

// =====================================================================================
// K3: objective + gradient.  Same grid; per-segment partial sums are reduced with warp
// shuffles and written to A.partial[k][*]; mpx_fgrad_final sums them in a fixed order.
// =====================================================================================
__device__ __forceinline__ double mpx_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <class PH>
MPX_HD int mpx_fgrad_smem_doubles(int d) {
  const int n1 = d + 1;
  return MpxTab::size(n1) + MpxTab::pad2((PH::NX + PH::NU) * n1) + 4 * MPX_NPART + 2;
}

template <class PH, bool GRAD>
__global__ void __launch_bounds__(MPX_THREADS) mpx_fgrad_kernel(const __grid_constant__ MpxPhaseArgs A) {
  extern __shared__ __align__(16) double smem[];
  constexpr int NX = PH::NX, NU = PH::NU, NA = PH::NA;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const MpxSeg S = mpx_segment(A, A.seg_begin + (int)blockIdx.x);
  const int n1 = S.n1, rb = S.rb;
  double* sTab = smem;
  const double* sRoots = sTab + MpxTab::off_roots(n1);
  const double* sW = sTab + MpxTab::off_w(n1);
  double* sXU = sTab + MpxTab::size(n1);
  double* sRed = sXU + MpxTab::pad2((NX + NU) * n1);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sRed + 4 * MPX_NPART);
  if (tid == 0) {
    mpx_mbar_init(bar, 1);
    const uint32_t bytes = (uint32_t)MpxTab::size(n1) * 8u;
    mpx_mbar_expect_tx(bar, bytes);
    mpx_tma_load_1d(sTab, S.tab, bytes, bar);
  }
  const double* zp = A.z + A.zoff;
  for (int i = tid; i < (NX + NU) * n1; i += MPX_THREADS) {
    const int v = i / n1, j = i - v * n1;
    sXU[i] = zp[(int64_t)v * A.N + S.s0 + j];
  }
  const int64_t zt = (int64_t)(NX + NU) * A.N;
  const double T0 = zp[zt], TF = zp[zt + 1];
  const double t0 = T0 / A.st, tf = TF / A.st;
  const double wk = A.w[S.k];
  const double h = (tf - t0) / A.delta * wk;
  const double gk = wk / (A.delta * A.st);
  __syncthreads();
  mpx_mbar_wait(bar, 0);

  double acc[3 + (NA > 0 ? NA : 1)];
#pragma unroll
  for (int i = 0; i < 3 + NA; ++i) acc[i] = 0.0;
  for (int r = tid + rb; r < n1; r += MPX_THREADS) {
    double x[NX > 0 ? NX : 1], u[NU > 0 ? NU : 1], a[NA > 0 ? NA : 1];
#pragma unroll
    for (int s = 0; s < NX; ++s) x[s] = sXU[s * n1 + r] * A.isx[s];
#pragma unroll
    for (int c = 0; c < NU; ++c) u[c] = sXU[(NX + c) * n1 + r] * A.isu[c];
#pragma unroll
    for (int m = 0; m < NA; ++m) a[m] = zp[zt + 2 + m] * A.isa[m];
    double sigma = 0.0, t = t0;
    if (PH::L_T) {
      const double dt = sRoots[r] - A.tau0;
      sigma = A.sig0[S.k] + wk * dt / A.delta;
      t = (t0 + (tf - t0) * A.sig0[S.k]) + h * dt;
    }
    double L[2], gl[PH::NGL > 0 ? PH::NGL : 1];
    PH::cost(x, u, t, a, L, gl);
    // composite weight: node r>=1 takes w[r] of its owner; w[0] of segments k>=1 is dropped (mpopt.py:4060-4062)
    const double W = sW[r];
    acc[0] += W * (h * L[0]);                                  // mpopt.py:206, :455
    if (GRAD) {
      double gv[NX + NU > 0 ? NX + NU : 1];
#pragma unroll
      for (int v = 0; v < NX + NU; ++v) gv[v] = 0.0;
#pragma unroll
      for (int e = 0; e < PH::NGL; ++e) {
        const int v = PH::gl_var(e);
        const double val = W * h * gl[e] * mpx_iscale<PH>(A, v);
        if (v < NX + NU) gv[v] = val;
        else acc[3 + (v - NX - NU)] += val;
      }
#pragma unroll
      for (int v = 0; v < NX + NU; ++v) A.grad[A.zoff + (int64_t)v * A.N + S.s0 + r] = gv[v];
      if (PH::L_NZ) {
        const double b = PH::L_T ? h * L[1] / A.st : 0.0;
        acc[1] += W * (-gk * L[0] + b * (1.0 - sigma));
        acc[2] += W * (gk * L[0] + b * sigma);
      }
    }
  }
  // block reduction: warp shuffles, then warp 0 sums the 4 per-warp values in a fixed order
#pragma unroll
  for (int i = 0; i < 3 + NA; ++i) {
    const double v = mpx_warp_sum(acc[i]);
    if (lane == 0) sRed[warp * MPX_NPART + i] = v;
  }
  __syncthreads();
  if (tid < 3 + NA) {
    double v = 0.0;
#pragma unroll
    for (int wq = 0; wq < MPX_THREADS / 32; ++wq) v += sRed[wq * MPX_NPART + tid];
    A.partial[(int64_t)S.k * MPX_NPART + tid] = v;
  }
}

// one CTA: deterministic sum of the per-segment partials + Mayer term (mpopt.py:296-298)
template <class PH, bool GRAD>
__global__ void __launch_bounds__(256) mpx_fgrad_final(const __grid_constant__ MpxPhaseArgs A) {
  constexpr int NX = PH::NX, NA = PH::NA;
  __shared__ double red[8][3 + (NA > 0 ? NA : 1)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double acc[3 + (NA > 0 ? NA : 1)];
#pragma unroll
  for (int i = 0; i < 3 + NA; ++i) acc[i] = 0.0;
  for (int k = A.seg_begin + tid; k < A.seg_end; k += 256)
#pragma unroll
    for (int i = 0; i < 3 + NA; ++i) acc[i] += A.partial[(int64_t)k * MPX_NPART + i];
#pragma unroll
  for (int i = 0; i < 3 + NA; ++i) {
    const double v = mpx_warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (tid == 0) {
    double tot[3 + (NA > 0 ? NA : 1)];
#pragma unroll
    for (int i = 0; i < 3 + NA; ++i) {
      tot[i] = 0.0;
      for (int wq = 0; wq < 8; ++wq) tot[i] += red[wq][i];
    }
    const double* zp = A.z + A.zoff;
    const int64_t zt = (int64_t)(NX + PH::NU) * A.N;
    double J = tot[0];
    if (GRAD) {
      A.grad[A.zoff + zt] = tot[1];
      A.grad[A.zoff + zt + 1] = tot[2];
#pragma unroll
      for (int m = 0; m < NA; ++m) A.grad[A.zoff + zt + 2 + m] = tot[3 + m];
    }
    if (A.flags & MPX_F_TAIL) {
      double xf[NX > 0 ? NX : 1], x0[NX > 0 ? NX : 1], a[NA > 0 ? NA : 1];
#pragma unroll
      for (int s = 0; s < NX; ++s) {
        xf[s] = zp[(int64_t)s * A.N + A.N - 1] * A.isx[s];
        x0[s] = zp[(int64_t)s * A.N] * A.isx[s];
      }
#pragma unroll
      for (int m = 0; m < NA; ++m) a[m] = zp[zt + 2 + m] * A.isa[m];
      double tc[PH::NTC > 0 ? PH::NTC : 1], jtc[PH::NJTC > 0 ? PH::NJTC : 1], M[1], gm[PH::NGM > 0 ? PH::NGM : 1];
      PH::term(xf, zp[zt + 1] / A.st, x0, zp[zt] / A.st, a, tc, jtc, M, gm);
      J += M[0];
      if (GRAD) {
#pragma unroll
        for (int e = 0; e < PH::NGM; ++e) {
          const int v = PH::gm_var(e);
          const double val = gm[e] * mpx_iscale_term<PH>(A, v);
          int64_t col;
          if (v < NX) col = (int64_t)v * A.N + A.N - 1;
          else if (v < 2 * NX) col = (int64_t)(v - NX) * A.N;
          else if (v == 2 * NX) col = zt + 1;
          else if (v == 2 * NX + 1) col = zt;
          else col = zt + 2 + (v - 2 * NX - 2);
          A.grad[A.zoff + col] += val;
        }
      }
    }
    if (A.accumulate_f) *A.fout += J;
    else *A.fout = J;
  }
}

#endif  // __CUDACC__
