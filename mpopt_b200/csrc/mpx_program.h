// mpx_program.h -- registry of compiled node-functor programs.
//
// A "program" is the set of kernels of csrc/mpx_kernels.cuh instantiated for the phase
// functors generated from one traced OCP (mpopt_b200/program.py).  Programs compiled ahead
// of time (csrc/gen/*.cu, built by __graft_entry__.build) register themselves here under
// the hash of their generated source; programs compiled at run time through NVRTC
// (csrc/mpx_plan.cu) implement the same interface.
#pragma once
#include <cuda_runtime.h>

#include "mpx_kernels.cuh"

struct MpxPhaseKernels {
  virtual ~MpxPhaseKernels() {}
  virtual cudaError_t gjac(const MpxPhaseArgs& a, bool jac, int grid, size_t smem, cudaStream_t st) const = 0;
  // deg > 0: use the instance specialised for that uniform degree if there is one (see has_degree)
  virtual cudaError_t gjac2(const MpxPhaseArgs& a, bool jac, int deg, int grid, int threads, size_t smem,
                            cudaStream_t st) const = 0;
  virtual cudaError_t gjac4(const MpxPhaseArgs& a, bool jac, int deg, int grid, int threads, size_t smem,
                            cudaStream_t st) const = 0;
  virtual bool has_degree(int deg) const = 0;
  virtual cudaError_t fgrad(const MpxPhaseArgs& a, bool grad, int grid, size_t smem, cudaStream_t st) const = 0;
  virtual cudaError_t fgrad_final(const MpxPhaseArgs& a, bool grad, cudaStream_t st) const = 0;
  virtual cudaError_t residual(const MpxPhaseArgs& a, bool deriv, int grid, cudaStream_t st) const = 0;
  virtual cudaError_t hess(const MpxPhaseArgs& a, const MpxHessLin& hl, int grid, cudaStream_t st) const = 0;  // node kernel + final
  // widths-as-variables NLP (mpopt_adaptive): extra rows / columns, and d f / d w
  virtual cudaError_t adapt(const MpxPhaseArgs& a, int grid, size_t smem, cudaStream_t st) const = 0;
  virtual cudaError_t adapt_grad(const MpxPhaseArgs& a, int grid, bool suffix, cudaStream_t st) const = 0;
  // Hessian of the widths-as-variables NLP: the part the widths add (one launch per segment parity, then the corner sum)
  virtual cudaError_t adapt_hess(const MpxPhaseArgs& a, int grid, int dmax, cudaStream_t st) const = 0;
  virtual cudaError_t adapt_hess_final(const MpxPhaseArgs& a, cudaStream_t st) const = 0;
};

// kernels that span all phases of a program: ONE g + jac_g launch for a multi-phase NLP (mpx_gjac2_multi_kernel)
struct MpxProgramKernels {
  virtual ~MpxProgramKernels() {}
  // args: n_phases MpxPhaseArgs; deg as in MpxPhaseKernels::gjac2 (0 = generic); grid = CTAs per phase
  virtual cudaError_t gjac2_all(const MpxPhaseArgs* args, int n_phases, const MpxEvArgs& ev, bool jac, int deg,
                                int grid_per_phase, int threads, size_t smem, cudaStream_t st) const = 0;
};

struct MpxProgramEntry {
  const char* key;
  int n_phases;
  const MpxPhaseKernels* const* phases;
  MpxProgramEntry* next;
  const MpxProgramKernels* all;  // NULL: single-phase program (or not generated): one launch per phase
};

extern "C" void mpx_register_program(MpxProgramEntry* e);
extern "C" int mpx_pdl_enabled(void);  // MPX_PDL=0 turns programmatic dependent launch off
const MpxProgramEntry* mpx_find_program(const char* key);

template <int... DEGS>
struct MpxDegs {};

// AOT implementation of the all-phases launch
template <class D, class... PHS>
struct MpxAotProgram;
template <int... DEGS, class... PHS>
struct MpxAotProgram<MpxDegs<DEGS...>, PHS...> final : MpxProgramKernels {
  static constexpr int P = sizeof...(PHS);
  template <bool JAC, int DEG>
  static cudaError_t launch(const MpxMultiArgs<P>& m, int threads, size_t smem, cudaStream_t st) {
    static bool done = false;
    auto kern = mpx_gjac2_multi_kernel<JAC, DEG, PHS...>;
    if (smem > 48 * 1024 && !done) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return e;
      done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(m.grid_per_phase * P), cfg.blockDim = dim3(threads), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = mpx_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, m);
  }
  template <int D0, int... REST>
  static cudaError_t pick(int deg, const MpxMultiArgs<P>& m, bool jac, int threads, size_t smem, cudaStream_t st) {
    if constexpr (sizeof...(REST) == 0) {
      return jac ? launch<true, D0>(m, threads, smem, st) : launch<false, D0>(m, threads, smem, st);  // list ends with 0
    } else {
      if (deg == D0) return jac ? launch<true, D0>(m, threads, smem, st) : launch<false, D0>(m, threads, smem, st);
      return pick<REST...>(deg, m, jac, threads, smem, st);
    }
  }
  cudaError_t gjac2_all(const MpxPhaseArgs* args, int n_phases, const MpxEvArgs& ev, bool jac, int deg, int grid_per_phase,
                        int threads, size_t smem, cudaStream_t st) const override {
    if (n_phases != P) return cudaErrorInvalidValue;
    MpxMultiArgs<P> m;
    for (int i = 0; i < P; ++i) m.a[i] = args[i];
    m.ev = ev, m.grid_per_phase = grid_per_phase, m.pad_ = 0;
    return pick<DEGS..., 0>(deg, m, jac, threads, smem, st);
  }
};

// AOT implementation: direct <<<>>> launches of the template instantiations
template <class PH, int... DEGS>
struct MpxAotPhase final : MpxPhaseKernels {
  template <class K>
  static cudaError_t allow_smem(K kern, size_t smem, bool& done) {
    if (smem > 48 * 1024 && !done) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return e;
      done = true;
    }
    return cudaSuccess;
  }
  cudaError_t gjac(const MpxPhaseArgs& a, bool jac, int grid, size_t smem, cudaStream_t st) const override {
    static bool d0 = false, d1 = false;
    cudaError_t e;
    if (jac) {
      if ((e = allow_smem(mpx_gjac_kernel<PH, true>, smem, d1)) != cudaSuccess) return e;
      mpx_gjac_kernel<PH, true><<<grid, MPX_THREADS, smem, st>>>(a);
    } else {
      if ((e = allow_smem(mpx_gjac_kernel<PH, false>, smem, d0)) != cudaSuccess) return e;
      mpx_gjac_kernel<PH, false><<<grid, MPX_THREADS, smem, st>>>(a);
    }
    return cudaGetLastError();
  }
  // launch with the programmatic-stream-serialization attribute (see mpx_pdl_wait in mpx_kernels.cuh)
  template <class K>
  static cudaError_t launch_pdl(K kern, const MpxPhaseArgs& a, int grid, int threads, size_t smem, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(threads), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = mpx_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, a);
  }
  template <int DEG>
  static cudaError_t launch2(const MpxPhaseArgs& a, bool jac, int grid, int threads, size_t smem, cudaStream_t st) {
    static bool d0 = false, d1 = false;
    cudaError_t e;
    if (jac) {
      if ((e = allow_smem(mpx_gjac2_kernel<PH, true, DEG>, smem, d1)) != cudaSuccess) return e;
      return launch_pdl(mpx_gjac2_kernel<PH, true, DEG>, a, grid, threads, smem, st);
    }
    if ((e = allow_smem(mpx_gjac2_kernel<PH, false, DEG>, smem, d0)) != cudaSuccess) return e;
    return launch_pdl(mpx_gjac2_kernel<PH, false, DEG>, a, grid, threads, smem, st);
  }
  template <int D0, int... REST>
  static cudaError_t pick2(int deg, const MpxPhaseArgs& a, bool jac, int grid, int threads, size_t smem, cudaStream_t st) {
    if constexpr (sizeof...(REST) == 0) {
      return launch2<D0>(a, jac, grid, threads, smem, st);  // the list ends with 0 = generic
    } else {
      if (deg == D0) return launch2<D0>(a, jac, grid, threads, smem, st);
      return pick2<REST...>(deg, a, jac, grid, threads, smem, st);
    }
  }
  cudaError_t gjac2(const MpxPhaseArgs& a, bool jac, int deg, int grid, int threads, size_t smem,
                    cudaStream_t st) const override {
    return pick2<DEGS..., 0>(deg, a, jac, grid, threads, smem, st);
  }
  template <int DEG>
  static cudaError_t launch4(const MpxPhaseArgs& a, bool jac, int grid, int threads, size_t smem, cudaStream_t st) {
    static bool d0 = false, d1 = false;
    cudaError_t e;
    if (jac) {
      if ((e = allow_smem(mpx_gjac4_kernel<PH, true, DEG>, smem, d1)) != cudaSuccess) return e;
      mpx_gjac4_kernel<PH, true, DEG><<<grid, threads, smem, st>>>(a);
    } else {
      if ((e = allow_smem(mpx_gjac4_kernel<PH, false, DEG>, smem, d0)) != cudaSuccess) return e;
      mpx_gjac4_kernel<PH, false, DEG><<<grid, threads, smem, st>>>(a);
    }
    return cudaGetLastError();
  }
  template <int D0, int... REST>
  static cudaError_t pick4(int deg, const MpxPhaseArgs& a, bool jac, int grid, int threads, size_t smem, cudaStream_t st) {
    if constexpr (sizeof...(REST) == 0) {
      return launch4<D0>(a, jac, grid, threads, smem, st);  // the list ends with 0 = generic
    } else {
      if (deg == D0) return launch4<D0>(a, jac, grid, threads, smem, st);
      return pick4<REST...>(deg, a, jac, grid, threads, smem, st);
    }
  }
  cudaError_t gjac4(const MpxPhaseArgs& a, bool jac, int deg, int grid, int threads, size_t smem,
                    cudaStream_t st) const override {
    return pick4<DEGS..., 0>(deg, a, jac, grid, threads, smem, st);
  }
  bool has_degree(int deg) const override { return deg > 0 && (... || (deg == DEGS)); }
  cudaError_t fgrad(const MpxPhaseArgs& a, bool grad, int grid, size_t smem, cudaStream_t st) const override {
    static bool d0 = false, d1 = false;
    cudaError_t e;
    if (grad) {
      if ((e = allow_smem(mpx_fgrad_kernel<PH, true>, smem, d1)) != cudaSuccess) return e;
      mpx_fgrad_kernel<PH, true><<<grid, MPX_THREADS, smem, st>>>(a);
    } else {
      if ((e = allow_smem(mpx_fgrad_kernel<PH, false>, smem, d0)) != cudaSuccess) return e;
      mpx_fgrad_kernel<PH, false><<<grid, MPX_THREADS, smem, st>>>(a);
    }
    return cudaGetLastError();
  }
  cudaError_t fgrad_final(const MpxPhaseArgs& a, bool grad, cudaStream_t st) const override {
    if (grad) mpx_fgrad_final<PH, true><<<1, 256, 0, st>>>(a);
    else mpx_fgrad_final<PH, false><<<1, 256, 0, st>>>(a);
    return cudaGetLastError();
  }
  cudaError_t residual(const MpxPhaseArgs& a, bool deriv, int grid, cudaStream_t st) const override {
    if (deriv) mpx_residual_kernel<PH, true><<<grid, 128, 0, st>>>(a);
    else mpx_residual_kernel<PH, false><<<grid, 128, 0, st>>>(a);
    return cudaGetLastError();
  }
  cudaError_t hess(const MpxPhaseArgs& a, const MpxHessLin& hl, int grid, cudaStream_t st) const override {
    mpx_hess_kernel<PH><<<grid, MPX_HESS_THREADS, 0, st>>>(a, hl);
    if (!a.ticket) mpx_hess_final<PH><<<1, MPX_HESS_FINAL_THREADS, 0, st>>>(a);  // else done by the node kernel's last CTA
    return cudaGetLastError();
  }
  template <int DEG>
  static cudaError_t launch_ad(const MpxPhaseArgs& a, int grid, size_t smem, cudaStream_t st) {
    static bool d0 = false, d1 = false;
    cudaError_t e;
    if (a.ad_jac) {
      if ((e = allow_smem(mpx_adapt_kernel<PH, true, DEG>, smem, d1)) != cudaSuccess) return e;
      mpx_adapt_kernel<PH, true, DEG><<<grid, MPX_THREADS, smem, st>>>(a);
    } else {  // g only: a separate instance, so that the row-assembly code does not set its register count
      if ((e = allow_smem(mpx_adapt_kernel<PH, false, DEG>, smem, d0)) != cudaSuccess) return e;
      mpx_adapt_kernel<PH, false, DEG><<<grid, MPX_THREADS, smem, st>>>(a);
    }
    return cudaGetLastError();
  }
  template <int D0, int... REST>
  static cudaError_t pick_ad(int deg, const MpxPhaseArgs& a, int grid, size_t smem, cudaStream_t st) {
    if constexpr (sizeof...(REST) == 0) {
      return launch_ad<D0>(a, grid, smem, st);  // the list ends with 0 = generic
    } else {
      if (deg == D0) return launch_ad<D0>(a, grid, smem, st);
      return pick_ad<REST...>(deg, a, grid, smem, st);
    }
  }
  cudaError_t adapt(const MpxPhaseArgs& a, int grid, size_t smem, cudaStream_t st) const override {
    return pick_ad<DEGS..., 0>(a.uniform_deg > 0 ? a.uniform_deg : 0, a, grid, smem, st);
  }
  cudaError_t adapt_grad(const MpxPhaseArgs& a, int grid, bool suffix, cudaStream_t st) const override {
    mpx_adapt_grad_kernel<PH><<<grid, MPX_THREADS, 0, st>>>(a);
    if (suffix) mpx_adapt_grad_suffix<PH><<<1, 32, 0, st>>>(a);  // time-dependent running cost only
    return cudaGetLastError();
  }
  template <int DEG>
  static cudaError_t launch_ah(const MpxPhaseArgs& a, int grid, size_t smem, cudaStream_t st) {
    static bool done = false;
    cudaError_t e = allow_smem(mpx_adapt_hess_kernel<PH, DEG>, smem, done);
    if (e != cudaSuccess) return e;
    mpx_adapt_hess_kernel<PH, DEG><<<grid, MPX_THREADS, smem, st>>>(a);
    return cudaGetLastError();
  }
  template <int D0, int... REST>
  static cudaError_t pick_ah(int deg, const MpxPhaseArgs& a, int grid, size_t smem, cudaStream_t st) {
    if constexpr (sizeof...(REST) == 0) {
      return launch_ah<D0>(a, grid, smem, st);  // the list ends with 0 = generic
    } else {
      if (deg == D0) return launch_ah<D0>(a, grid, smem, st);
      return pick_ah<REST...>(deg, a, grid, smem, st);
    }
  }
  cudaError_t adapt_hess(const MpxPhaseArgs& a, int grid, int dmax, cudaStream_t st) const override {
    const size_t smem = (size_t)mpx_adapt_hess_smem_doubles<PH>(dmax) * sizeof(double);
    return pick_ah<DEGS..., 0>(a.uniform_deg > 0 ? a.uniform_deg : 0, a, grid, smem, st);  // degree-specialised when the plan is uniform
  }
  cudaError_t adapt_hess_final(const MpxPhaseArgs& a, cudaStream_t st) const override {
    mpx_adapt_hess_final<PH><<<1, 64, 0, st>>>(a);
    return cudaGetLastError();
  }
};
