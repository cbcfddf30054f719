// Dev probe (not product): L2->SM throughput of the activation-patch loads of the conv kernel, current layout vs the
// round-2 layout, with no consumer (pure TMA + mbarrier ring, 4 stages, 148 persistent CTAs).
//   (a) channels-last [H][W][256] fp16, SWIZZLE_128B, three column-shifted (18 x 8 px x 64 ch) boxes per chunk (55 KB)
//   (b) interleaved [32][H][W][8] fp16, SWIZZLE_NONE, ONE (18 x 10 px x 8 k-groups) box per chunk (23 KB)
// Both serve the same 16 x 8 output tile x 64 channels; reports useful-tile throughput (tiles*chunks per second).
#include <cstdio>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../buddy_b200/csrc/common.cuh"
namespace buddy { void set_last_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } }
using namespace buddy;
constexpr int H = 256, W = 528, C = 256, kStages = 4;

template <bool kInterleaved>
__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, int tiles_real, int reps, long long* out) {
  const int tiles = tiles_real * reps;   // the tile set is walked `reps` times (tile index modulo tiles_real)
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kStages];
  constexpr int stage_bytes = kInterleaved ? 8 * 18 * 10 * 16 : 3 * 18 * 8 * 128;
  constexpr int stage_pitch = ((stage_bytes + 1023) / 1024) * 1024;
  if (threadIdx.x == 0) { for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    int issued = 0, waited = 0;
    uint32_t phase_bits = 0;
    const int per_cta = (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x * (C / 64);
    for (int it = 0; it < per_cta + kStages; ++it) {
      if (it >= kStages || it >= per_cta) {
        if (waited < per_cta) {
          const int s = waited % kStages;
          mbar_wait(&full[s], (phase_bits >> s) & 1u);
          phase_bits ^= 1u << s;
          ++waited;
        }
      }
      if (issued < per_cta) {
        const int s = issued % kStages;
        const int t = (blockIdx.x + (issued / (C / 64)) * gridDim.x) % tiles_real, kc = issued % (C / 64);
        const int h0 = (t / (W / 8)) * 16, w0 = (t % (W / 8)) * 8;
        mbar_expect_tx(&full[s], stage_bytes);
        if (kInterleaved) {
          tma_load_4d(&tm, smem + s * stage_pitch, &full[s], (w0 - 1) * 8, h0 - 1, kc * 8, 0);
        } else {
          for (int j = 0; j < 3; ++j)
            tma_load_4d(&tm, smem + s * stage_pitch + j * 18 * 8 * 128, &full[s], kc * 64, w0 + j - 1, h0 - 1, 0);
        }
        ++issued;
      }
    }
    while (waited < per_cta) {
      const int s = waited % kStages;
      mbar_wait(&full[s], (phase_bits >> s) & 1u);
      phase_bits ^= 1u << s;
      ++waited;
    }
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  __half* d; cudaMalloc(&d, size_t(H) * W * C * 2); cudaMemset(d, 0, size_t(H) * W * C * 2);
  long long* out; cudaMalloc(&out, 8);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUtensorMap tmA, tmB;
  { cuuint64_t dims[4] = {C, W, H, 1}; cuuint64_t st[3] = {cuuint64_t(C) * 2, cuuint64_t(C) * W * 2, cuuint64_t(C) * W * H * 2};
    cuuint32_t box[4] = {64, 8, 18, 1};
    printf("encode a: %d\n", (int)enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
  { cuuint64_t dims[4] = {cuuint64_t(8) * W, H, C / 8, 1}; cuuint64_t st[3] = {cuuint64_t(8) * W * 2, cuuint64_t(8) * W * H * 2, cuuint64_t(8) * W * H * (C / 8) * 2};
    cuuint32_t box[4] = {80, 18, 8, 1};
    printf("encode b: %d\n", (int)enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
  const int tiles = (H / 16) * (W / 8);
  const int smemA = kStages * 56 * 1024, smemB = kStages * 23 * 1024;
  cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemA);
  cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemB);
  for (int rep = 0; rep < 3; ++rep) {
    for (int mode = 0; mode < 2; ++mode) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      const int reps = 64;
      if (mode == 0) probe<false><<<148, 64, smemA>>>(tmA, tiles, reps, out);
      else probe<true><<<148, 64, smemB>>>(tmB, tiles, reps, out);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double chunks = double(tiles) * reps * (C / 64);
      const double bytes = chunks * (mode == 0 ? 3 * 18 * 8 * 128 : 8 * 18 * 10 * 16);
      printf("%s: %.3f ms, %.2f M tile-chunks/s, %.2f TB/s moved L2->SM, err=%d\n",
             mode == 0 ? "(a) channels-last SW128, 3 shifted patches" : "(b) interleaved no-swizzle, 1 halo patch     ", ms,
             chunks / ms / 1e3, bytes / ms / 1e9, (int)e);
    }
  }
  return 0;
}
