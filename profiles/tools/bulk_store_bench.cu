// Microbenchmark: how fast can one SM (and the chip) drain shared-memory images to HBM?
//   mode 0: cp.async.bulk.global.shared::cta, one bulk store of `chunk` bytes per warp per iteration, nbuf images per warp
//   mode 1: 16-byte st.global.cs thread stores of the same bytes straight from registers
//   mode 2: like 0, but the warp rewrites `touch` doubles per lane of the image before every store (generic-proxy writes
//           + fence.proxy.async), which is what the Jacobian kernel does
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_store_bench bulk_store_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

//   mode 3: like 2, plus a distance-1 prefetch of 9 doubles per lane from a 4.4 MB input (the decision vector) that the
//           next iteration consumes -- does read latency grow when the memory system is saturated with writes?
template <int NBUF>
__global__ void __launch_bounds__(1024, 1) k_bulk(double* out, long total_chunks, int chunk_doubles, int touch,
                                                  const double* __restrict__ zin, long zn) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double* img = sm + (size_t)warp * NBUF * chunk_doubles;
  for (int i = lane; i < NBUF * chunk_doubles; i += 32) img[i] = 1.0 + i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const long tw = (long)gridDim.x * nwarps;
  int turn = 0;
  double xr[9], acc = 0.0;
  long c0 = (long)warp * gridDim.x + blockIdx.x;
  if (zin)
    for (int v = 0; v < 9; ++v) xr[v] = zin[(v * (zn / 9) + (c0 * 32) % (zn / 9) + lane)];
  for (long c = c0; c < total_chunks; c += tw, ++turn) {
    double* buf = img + (size_t)(NBUF == 2 ? (turn & 1) : 0) * chunk_doubles;
    if (zin) {
      for (int v = 0; v < 9; ++v) acc += xr[v];
      const long cn = c + tw < total_chunks ? c + tw : c;
      for (int v = 0; v < 9; ++v) xr[v] = zin[(v * (zn / 9) + (cn * 32) % (zn / 9) + lane)];
      if (lane == 31) buf[0] = acc;
    }
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
    __syncwarp();
    if (touch) {
      for (int t = 0; t < touch; ++t) buf[(lane * 26 + t) % chunk_doubles] = (double)(c + t);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
    }
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk_doubles),
                   "r"(smem_u32(buf)), "r"((uint32_t)chunk_doubles * 8u)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__global__ void __launch_bounds__(1024, 1) k_thread(double* out, long total_chunks, int chunk_doubles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const long tw = (long)gridDim.x * nwarps;
  for (long c = (long)warp * gridDim.x + blockIdx.x; c < total_chunks; c += tw) {
    double* dst = out + c * chunk_doubles;
    const double a = (double)c, b = a + 1.0;
    for (int i = lane; i < chunk_doubles / 2; i += 32)
      asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(dst + 2 * i), "d"(a), "d"(b) : "memory");
  }
}

int main(int argc, char** argv) {
  const long total_bytes = 100L << 20;
  const int R = 4;
  double* out;
  cudaMalloc(&out, R * total_bytes);
  cudaMemset(out, 0, R * total_bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(k_bulk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  printf("mode nbuf warps chunkB touch  us  GB/s\n");
  const int chunks[] = {1920, 6240, 12480, 24960};
  const int warps[] = {4, 8, 12, 16, 24, 32};
  double* zin;
  const long zn = 552960;
  cudaMalloc(&zin, zn * 8);
  cudaMemset(zin, 0, zn * 8);
  for (int mode = 2; mode < 4; ++mode)
    for (int nbuf = 1; nbuf <= (mode == 1 ? 1 : 2); ++nbuf)
      for (int ci = 0; ci < 4; ++ci)
        for (int wi = 0; wi < 6; ++wi) {
          const int cd = chunks[ci] / 8, w = warps[wi];
          const size_t smem = (size_t)w * nbuf * cd * 8;
          if (mode != 1 && smem > 220 * 1024) continue;
          const long nchunks = total_bytes / (cd * 8);
          const int touch = mode >= 2 ? 11 : 0;
          if (cd * 8 != 6240) continue;
          float best = 1e9f;
          for (int rep = 0; rep < 12; ++rep) {
            double* o = out + (size_t)(rep % R) * (total_bytes / 8);
            cudaEventRecord(e0);
            if (mode == 1) k_thread<<<nsm, w * 32>>>(o, nchunks, cd);
            else if (nbuf == 1) k_bulk<1><<<nsm, w * 32, smem>>>(o, nchunks, cd, touch, mode == 3 ? zin : nullptr, zn);
            else k_bulk<2><<<nsm, w * 32, smem>>>(o, nchunks, cd, touch, mode == 3 ? zin : nullptr, zn);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 2 && ms < best) best = ms;
          }
          cudaError_t err = cudaGetLastError();
          if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
          printf("%d %d %2d %5d %2d %7.2f %7.0f\n", mode, nbuf, w, cd * 8, touch, best * 1e3, nchunks * cd * 8.0 / (best * 1e-3) / 1e9);
        }
  return 0;
}
