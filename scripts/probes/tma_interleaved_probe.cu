// Dev probe (not product) for the round-2 conv design: load ONE (bh+2) x (bw+2) halo patch of a 16-byte-interleaved
// activation tensor [B][C/8][H][W][8 fp16] with a single TMA request into the K-major SWIZZLE_NONE "core matrix"
// layout [k-group][patch row][patch pixel][8] that umma_noswizzle_shift_probe.cu reads with shifted descriptors.
// The tensor map merges (8, W) into one contiguous inner dimension of 8*W elements: box = (8*(bw+2), bh+2, C/8 chunk).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_interleaved_probe tma_interleaved_probe.cu -lcuda
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../buddy_b200/csrc/common.cuh"
namespace buddy { void set_last_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } }
using namespace buddy;

constexpr int H = 20, W = 24, CG = 4, BH = 16, BW = 8, KG = 2;   // tensor [1][CG][H][W][8]; patch 18 x 10, 2 k-groups
constexpr int PH = BH + 2, PW = BW + 2;

__device__ __forceinline__ void tma_load_4d_plain(const CUtensorMap* m, void* smem, uint64_t* bar, int c0, int c1, int c2, int c3) {
  tma_load_4d(m, smem, bar, c0, c1, c2, c3);
}

__global__ void probe(const __grid_constant__ CUtensorMap tm, int w0, int h0, int cg0, __half* out) {
  __shared__ __align__(1024) __half s[KG * PH * PW * 8];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, sizeof(s));
    tma_load_4d_plain(&tm, s, &bar, (w0 - 1) * 8, h0 - 1, cg0, 0);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < KG * PH * PW * 8; i += blockDim.x) out[i] = s[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  // value of element (cg, h, w, c) = +-(cg * 512 + h * 24 + w + 1), negative for c == 7: integers < 2048 are exact in fp16
  std::vector<__half> hx(size_t(CG) * H * W * 8);
  for (int cg = 0; cg < CG; ++cg)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w)
        for (int c = 0; c < 8; ++c)
          hx[((size_t(cg) * H + h) * W + w) * 8 + c] = __float2half((c == 7 ? -1.f : 1.f) * float(cg * 512 + h * 24 + w + 1));
  __half* dx; cudaMalloc(&dx, hx.size() * 2); cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  __half* dout; cudaMalloc(&dout, KG * PH * PW * 8 * 2);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  CUtensorMap tm;
  cuuint64_t dims[4] = {cuuint64_t(8 * W), H, CG, 1};
  cuuint64_t strides[3] = {cuuint64_t(8 * W * 2), cuuint64_t(8 * W * H * 2), cuuint64_t(8 * W * H * CG * 2)};
  cuuint32_t box[4] = {8 * PW, PH, KG, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d (box inner = %d bytes)\n", (int)r, 8 * PW * 2);
  if (r != CUDA_SUCCESS) return 1;
  std::vector<__half> ho(KG * PH * PW * 8);
  int bad_total = 0;
  const int cases[4][3] = {{0, 0, 0}, {8, 0, 2}, {16, 16, 1}, {8, 16, 2}};   // (w0, h0, cg0): corners incl. halo outside the image
  for (auto& cs : cases) {
    const int w0 = cs[0], h0 = cs[1], cg0 = cs[2];
    cudaMemset(dout, 0xff, ho.size() * 2);
    probe<<<1, 128>>>(tm, w0, h0, cg0, dout);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(ho.data(), dout, ho.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int kg = 0; kg < KG; ++kg)
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw)
          for (int c = 0; c < 8; ++c) {
            const int h = h0 - 1 + ph, w = w0 - 1 + pw, cg = cg0 + kg;
            float want = 0.f;
            if (h >= 0 && h < H && w >= 0 && w < W && cg < CG) want = (c == 7 ? -1.f : 1.f) * float(cg * 512 + h * 24 + w + 1);
            const float got = __half2float(ho[((size_t(kg) * PH + ph) * PW + pw) * 8 + c]);
            if (got != want) ++bad;
          }
    printf("patch at (w0 %2d, h0 %2d, cg0 %d): %s (%d mismatches) err=%d\n", w0, h0, cg0, bad ? "MISMATCH" : "exact incl. zero-filled halo", bad, (int)e);
    bad_total += bad;
  }
  printf(bad_total ? "RESULT: layout differs from [k-group][row][pixel][8]\n"
                   : "RESULT: one TMA request fills [k-group][row][pixel][8] (pitch 10 pixels) with hardware zero padding\n");
  return 0;
}
