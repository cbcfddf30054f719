// Dev probe (not product) for the round-2 conv design: does a K-major SWIZZLE_NONE UMMA shared-memory descriptor accept
// a start address that is only 16-byte aligned, i.e. can ONE activation patch stored as [k-group][pixel][8 fp16] serve
// all nine taps of a 3x3 conv by shifting the descriptor start by whole pixels (16 B) — including column shifts, which
// the SWIZZLE_128B layout cannot express?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_noswizzle_shift_probe umma_noswizzle_shift_probe.cu
// A[pixel p][k] = p (k < 8) / 1000 + p (k >= 8), 256 pixels; B = 16x16 identity; D[m][n] must equal A[m + shift][n].
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../buddy_b200/csrc/common.cuh"
namespace buddy { void set_last_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } }
using namespace buddy;

__device__ __forceinline__ uint64_t make_noswizzle_kmajor_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;   // leading byte offset: core matrices adjacent in K
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;   // stride byte offset: 8-row groups adjacent in M / N
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version 1
  return d;                                                  // layout type 0 = no swizzle
}

constexpr int kPix = 256;
__global__ void __launch_bounds__(128, 1) probe(int shift, int sbo, float* out) {
  __shared__ __align__(1024) __half sA[2 * kPix * 8];   // [kgroup 2][pixel 256][8]
  __shared__ __align__(1024) __half sB[2 * 16 * 8];     // [kgroup 2][n 16][8]
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * kPix * 8; i += blockDim.x) {
    const int kg = i / (kPix * 8), p = (i / 8) % kPix;
    sA[i] = __float2half(kg == 0 ? float(p) : float(1000 + p));
  }
  for (int i = threadIdx.x; i < 2 * 16 * 8; i += blockDim.x) {
    const int kg = i / (16 * 8), n = (i / 8) % 16, kk = i % 8;
    sB[i] = __float2half((kg * 8 + kk) == n ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 32); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t d = slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_f16(128, 16);
    const uint64_t da = make_noswizzle_kmajor_desc(smem_u32(sA) + shift * 16, kPix * 16, sbo);
    const uint64_t db = make_noswizzle_kmajor_desc(smem_u32(sB), 16 * 16, 128);
    umma_f16(d, da, db, idesc, 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[32];
  // 32x32b.x32 reads 32 columns; only 16 are meaningful (N = 16) but the allocation is 32 columns wide
  tmem_ld_32x32(d + (static_cast<uint32_t>(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(d, 32); }
}

int main() {
  float* out;
  cudaMalloc(&out, 128 * 16 * 4);
  float h[128 * 16];
  int bad_total = 0;
  for (int sbo : {128, 160}) {   // 128: rows dense; 160: 8-pixel rows of a patch with pitch 10 pixels (16x8 tile + halo)
  for (int shift : {0, 1, 2, 7, 8, 9, 10, 11, 12, 21, 22}) {
    cudaMemset(out, 0, sizeof(h));
    probe<<<1, 128>>>(shift, sbo, out);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 16; ++n) {
        const int pix = (m / 8) * (sbo / 16) + (m % 8) + shift;
        const float want = n < 8 ? float(pix) : float(1000 + pix);
        if (h[m * 16 + n] != want) ++bad;
      }
    printf("SBO %3d  shift %2d pixels (start + %3d B): %s (%d mismatches; D[0][0]=%g D[0][8]=%g D[127][0]=%g) err=%d\n", sbo, shift,
           shift * 16, bad ? "MISMATCH" : "exact", bad, h[0], h[8], h[127 * 16], (int)e);
    bad_total += bad;
  }
  }
  printf(bad_total ? "RESULT: shifted SWIZZLE_NONE descriptors do NOT behave linearly\n"
                   : "RESULT: a 16-byte-aligned start shifts the A window by whole pixels — one patch can serve all 9 taps\n");
  return 0;
}
