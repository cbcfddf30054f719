// Dev probe (not product): raw tcgen05.mma issue rate from resident shared-memory operands, no TMA in the loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
// Prints cycles per MMA for kind::f16 / kind::f8f6f4, N = 128 / 256, cta_group 1 / 2, all SMs busy.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../buddy_b200/csrc/common.cuh"
namespace buddy { void set_last_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } }
using namespace buddy;

template <bool kPair, bool kF8>
__global__ void __launch_bounds__(128, 1) probe(int N, int iters, long long* out, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&bar2[i], 1); fence_barrier_init(); }
  if (warp == 0) { if (kPair) { tmem_alloc_2sm(&slot, 512); tmem_relinquish_2sm(); } else { tmem_alloc(&slot, 512); tmem_relinquish(); } }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  tc_fence_after();
  const uint32_t d = slot;
  if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t idesc = kF8 ? make_idesc_e4m3(kPair ? 256 : 128, N) : make_idesc_f16(kPair ? 256 : 128, N);
    // mode & 8: A operand = K-major SWIZZLE_NONE halo patch [k-group][18 x 10 pixels][16 B], window shifted by 11 pixels
    uint64_t da = make_sw128_kmajor_desc(smem_u32(smem));
    uint64_t da_step = 2;   // per K=16 MMA: +32 bytes inside the 128-byte swizzled row
    if (mode & 8) {
      const uint32_t lbo = 180 * 16, sbo = 160;
      da = static_cast<uint64_t>(((smem_u32(smem) + 11 * 16) & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) |
           (static_cast<uint64_t>(sbo >> 4) << 32) | (static_cast<uint64_t>(1) << 46);
      da_step = (2 * lbo) >> 4;   // next two k-groups
    }
    const uint64_t db = make_sw128_kmajor_desc(smem_u32(smem) + 16384);
    const long long t0 = clock64();
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      if (mode & 2) tc_fence_after();
      if ((mode & 4) && i >= 8) { mbar_wait(&bar2[st], ph ^ 1); }   // wait for the commit of 8 groups ago (always done)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (kPair) { if (kF8) umma_f8_2sm(d, da + da_step * k, db + 2 * k, idesc, 1u); else umma_f16_2sm(d, da + da_step * k, db + 2 * k, idesc, 1u); }
        else { if (kF8) umma_f8(d, da + da_step * k, db + 2 * k, idesc, 1u); else umma_f16(d, da + da_step * k, db + 2 * k, idesc, 1u); }
      }
      if (mode & 1) { if (kPair) umma_commit_2sm(&bar2[st]); else umma_commit(&bar2[st]); }
      if (++st == 8) { st = 0; ph ^= 1; }
    }
    if (kPair) umma_commit_2sm(&bar); else umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (kPair && warp == 1 && lane == 0) {
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  if (warp == 0) { tc_fence_after(); if (kPair) tmem_dealloc_2sm(d, 512); else tmem_dealloc(d, 512); }
}

template <bool kPair, bool kF8>
void run(const char* name, int N, int grid, int mode = 0) {
  long long* out;
  cudaMalloc(&out, 8);
  const int iters = 4000;
  auto k = probe<kPair, kF8>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 48 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = kPair ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, N, iters, out, mode);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc; cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
    if (rep == 1) {
      const double per = double(cyc) / (iters * 4.0);
      const double flop = 2.0 * (kPair ? 256 : 128) * N * (kF8 ? 32 : 16) * iters * 4.0 * (kPair ? grid / 2 : grid);
      printf("mode %d %-24s N=%3d grid %3d: %7.1f cycles/MMA, %.3f ms, %.0f TFLOP/s (%s) err=%d\n", mode, name, N, grid, per, ms,
             flop / (ms * 1e-3) / 1e12, kF8 ? "fp8" : "fp16", (int)e);
    }
  }
  cudaFree(out);
}

int main() {
  for (int mode : {0, 8}) {
    run<false, false>("cta_group::1 f16 M128", 256, 148, mode);
    run<false, false>("cta_group::1 f16 M128", 128, 148, mode);
    run<true, false>("cta_group::2 f16 M256", 256, 148, mode);
    run<true, false>("cta_group::2 f16 M256", 128, 148, mode);
    run<true, true>("cta_group::2 f8  M256", 128, 148, mode);
  }
  return 0;
}
