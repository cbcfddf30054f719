// Implicit-GEMM 3x3 / 1x1 convolution and batched GEMM for sm_100a.
//
//   * operands: fp16 (+ optional e4m3 correction pair), channels-last; fp32 accumulation in TMEM
//     (tcgen05.mma kind::f16 K=16 / kind::f8f6f4 K=32, N <= 256)
//   * CTA pairs: two CTAs of a 2-CTA cluster share a 256-row tile pair (cta_group::2, M = 256): each loads its own
//     128-pixel activation patch and HALF of the weight tile; the leader CTA issues the MMAs
//   * A ring : per 64-channel chunk ONE halo patch of (bh+2) x (bw+2) pixels (one TMA 4-D tiled load; out-of-image
//              coordinates are zero-filled by TMA == the conv's "same" padding, no im2col buffer, no halo code).
//              The SWIZZLE_128B XOR is a function of the absolute shared-memory address (verified on B200,
//              scripts/probes/umma_sw128_shift_probe.cu: any 128-byte line shift, 8-row groups 1280 bytes apart), so
//              tap (dy, dx) is the descriptor start offset (dy * (bw+2) + dx) * 128 B with SBO = (bw+2) * 128 B:
//              the chunk is read from L2 1.41x instead of 9x (round 1: three column-shifted patches, 3.4x)
//   * B ring : TMA 3-D loads of the packed weights per (chunk, tap) — a whole kernel row of three taps per stage where
//              it fits (N <= 128)
//   * both land in SWIZZLE_128B K-major layout, exactly what the UMMA shared-memory descriptors expect
//   * warp roles: warp0 = TMA producer, warp1 = MMA issuer (+ TMEM alloc), warps 2..5 = epilogue; producer and
//     issuer loops are warp-uniform (descriptor math on the uniform datapath), one elected lane issues
//   * persistent CTAs (one per SM), double-buffered TMEM accumulator so the epilogue of tile i overlaps
//     the mainloop of tile i+1
//   * epilogue: TMEM -> registers -> (+bias row, +TMA-loaded residual) * scale -> swizzled smem chunk -> TMA store;
//               fused GroupNorm bundle statistics (column sums of the staged chunk, fp64 atomics); a direct
//               register -> global epilogue serves fp16 / strided / narrow outputs
//
// Replaces nn.Conv2d / NIN / attention einsums of the reference network
// (networks/ncsnpp_utils/layers.py:100-126,548-557; layerspp.py:75-91) and their data-gradients.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/buddy_b200.h"
#include "common.cuh"

namespace buddy {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;            // fp16 elements per k-block = 128 bytes = one swizzle row
constexpr int kMaxStagesA = 4;         // A ring (activation patches)
constexpr int kMaxStagesB = 12;        // B ring (weight tiles)
constexpr int kMaxAcc = 4;             // TMEM accumulator stages: 2 x 256 columns, or 4 x 128 when n_tile <= 128
constexpr int kThreads = 192;

struct GemmParams {
  int batch, H, W;
  int bh, bw, tiles_h, tiles_w;
  int n_tiles, n_tile, n_total;
  int taps, kchunks1, kchunks2, b_batched;
  int a_wrap1, a_wrap2;  // A-side channel chunk = k-chunk % a_wrap (split-precision operands re-read the hi half)
  int kchunks8_1, kchunks8_2;  // fp8 correction phases (128-byte chunks per tap / for the skip conv); 0 = none
  // A ring: one stage = the activation patch of ONE 64-channel chunk.  3x3 phases (halo = 1): one patch of
  // (bh + 2) x (bw + 2) pixels (patch_bytes; 128-byte lines, a_pitch bytes between patch rows); the nine taps are
  // MMA-descriptor start offsets into it.  1x1 phases: bh x bw pixels (patch1_bytes).  B ring: weight tiles per
  // (chunk, tap).
  int halo, patch_bytes, patch1_bytes, a_pitch, a_stage_bytes, stages_a, stages_b;
  int acc_stages, acc_stride;   // TMEM accumulator ring
  int mt;   // pixel tiles (of bh x bw) stacked vertically per CTA and work item: 1, or 2 for N <= 128 3x3 launches —
            // every weight tile then feeds two accumulators (M = 256 per CTA), halving the weight stream that
            // dominates the L2->SM traffic of those layers; the two halves share one (2*bh + 2)-row halo patch
  int tpb;  // taps per B-ring stage (3x3 phases): 3 = one kernel row of weights per stage (one TMA load, one wait, one
            // commit per 12 MMAs — the issue loop, not the tensor pipe, bounds layers with N <= 128), else 1
  float* out32;
  __half* out16;
  long long ldc;
  int col_off;
  const float* bias;
  const float* bias_b;
  const float* resid;
  long long ld_res;
  float scale;
  double* stats;
  int dbg;         // profiling experiments only: 1 = epilogue only hands the accumulator back, 2 = + TMEM loads
  // fused GroupNorm-backward statistics (see buddy_gemm_desc::gnb_*); gnb_x == nullptr: off
  const float* gnb_x;
  const double* gnb_stats;
  const float* gnb_gamma;
  const float* gnb_beta;
  double* gnb_gsum;
  int gnb_groups, gnb_cpg, gnb_silu;
  float gnb_eps;
  int staged;      // 1: epilogue stages 128x32 fp32 chunks in shared memory and writes them with TMA stores
  int res_staged;  // 1: the residual is TMA-loaded into the staging buffer (needs staged)
};

constexpr int kChunkBytes = kTileM * 128;   // one staged epilogue chunk: 128 rows x 32 fp32
constexpr int kEpiThreads = 128;

// kPair: two CTAs of a cluster work on one 256-row tile pair with tcgen05 cta_group::2 — each CTA loads its own 128
// pixel rows of A and HALF of the weight tile, the leader CTA issues M=256 MMAs that read both halves, each CTA
// drains its own 128 accumulator lanes.  Halves the L2->SM weight traffic and the shared-memory reads per MMA.
template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ CUtensorMap tmA8, const __grid_constant__ CUtensorMap tmB8,
                 const __grid_constant__ CUtensorMap tmA82, const __grid_constant__ CUtensorMap tmB82,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                 const GemmParams p) {
  // SWIZZLE_128B atoms need a 1024-byte aligned base: the window is rounded up at run time (the host allocates 1 KB
  // of slack), so nothing depends on how much shared memory the driver reserves in front of the dynamic window.
  // Both CTAs of a pair see the same window offset, so the rounded offsets agree (cta_group::2 descriptors and the
  // multicast commits address the peer's shared memory by offset).
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;   // 0 = leader CTA of the pair
  const int b_rows = kPair ? (p.n_tile >> 1) : p.n_tile;   // weight-tile rows this CTA loads
  const int a_stage_bytes = p.a_stage_bytes;
  const int b_tile_bytes = b_rows * 128;              // one weight tile (one tap, one 64-wide chunk)
  const int b_stage_bytes = p.tpb * b_tile_bytes;     // one B-ring stage
  // [A ring][B ring][2 staged epilogue chunks (if staged)][bias row 1 KB (if staged)][barriers]
  uint8_t* smem_b = smem + p.stages_a * a_stage_bytes;
  uint8_t* stage_out = smem_b + p.stages_b * b_stage_bytes;
  float* bias_s = reinterpret_cast<float*>(stage_out + 2 * kChunkBytes);
  float* gnb_s = bias_s + 256;   // [4][n_tile]: rstd, -mean*rstd, gamma, beta per output channel of this tile (gnb)
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out +
                                               (p.staged ? 2 * kChunkBytes + 1024 + (p.gnb_x ? 4096 : 0) : 0));
  uint64_t* a_full = bars;                       // [kMaxStagesA]
  uint64_t* a_empty = a_full + kMaxStagesA;      // [kMaxStagesA]
  uint64_t* b_full = a_empty + kMaxStagesA;      // [kMaxStagesB]
  uint64_t* b_empty = b_full + kMaxStagesB;      // [kMaxStagesB]
  uint64_t* tmem_full = b_empty + kMaxStagesB;   // [kMaxAcc]
  uint64_t* tmem_empty = tmem_full + kMaxAcc;    // [kMaxAcc]
  uint64_t* res_bar = tmem_empty + kMaxAcc;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.kchunks2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    for (int s = 0; s < p.stages_a; ++s) {
      mbar_init(&a_full[s], kPair ? 2 : 1);   // pair: one expect_tx arrival per CTA, on the leader's barrier
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < p.stages_b; ++s) {
      mbar_init(&b_full[s], kPair ? 2 : 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < kMaxAcc; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kPair ? 8 : 4);  // one arrival per epilogue warp (of both CTAs, on the leader's)
    }
    for (int s = 0; s < 2; ++s) mbar_init(&res_bar[s], 1);
    if (p.staged) {
      tma_prefetch_desc(&tmOut);
      if (p.res_staged) tma_prefetch_desc(&tmRes);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) {
      tmem_alloc_2sm(tmem_slot, 512);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();   // the peer's barriers are initialised before anything signals them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int pix_tiles = p.batch * tiles_per_img;
  // work items: (pixel tile | pixel-tile pair) x n-tile; a pair's second tile may lie past the end (b == batch):
  // its TMA loads are zero-filled, its stores clipped, its statistics skipped
  const int total_tiles = (kPair ? ((pix_tiles + 1) >> 1) : pix_tiles) * p.n_tiles;
  // blocked assignment: CTA (pair) k works on the consecutive work items [t_begin, t_end) — it stays inside one image
  // as long as possible, so the epilogue's running GroupNorm sums are flushed (fp64 atomics) a few times per launch
  const int n_workers = kPair ? (gridDim.x >> 1) : gridDim.x;
  const int worker = kPair ? (blockIdx.x >> 1) : blockIdx.x;
  const int t_begin = static_cast<int>(static_cast<long long>(total_tiles) * worker / n_workers);
  const int t_end = static_cast<int>(static_cast<long long>(total_tiles) * (worker + 1) / n_workers);
  const int th = p.bh * p.mt;   // pixel rows per work item and CTA
  // contraction order: phases [fp8 conv | fp8 skip conv | fp16 conv | fp16 skip conv]; inside a phase: channel chunk
  // outer (one A stage each), tap inner (one B stage each)

  if (warp == 0) {
    // ================================================================ TMA producer
    // The whole warp runs the loop (warp-uniform control flow and addresses stay on the uniform datapath); only the
    // TMA / expect_tx instructions themselves are issued by one elected lane.
    if (!(p.dbg & 4)) {   // dbg 4 (profiling): no loads at all, the MMA loop runs on stale smem
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int nt = t % p.n_tiles;
        const int pt = kPair ? 2 * (t / p.n_tiles) + static_cast<int>(rank) : t / p.n_tiles;
        const int b = pt / tiles_per_img;
        const int r = pt - b * tiles_per_img;
        const int h0 = (r / p.tiles_w) * th;
        const int w0 = (r % p.tiles_w) * p.bw;
        const int brow = nt * p.n_tile + static_cast<int>(rank) * b_rows;   // first weight row this CTA loads
        // one k-chunk of phase ph (0 fp8 conv, 1 fp8 skip conv, 2 fp16 conv, 3 fp16 skip conv): its A patch, then its
        // weight stages
        auto load_chunk = [&](int ph, int kc) {
          const CUtensorMap* ma = ph == 0 ? &tmA8 : (ph == 1 ? &tmA82 : (ph == 2 ? &tmA : &tmA2));
          const CUtensorMap* mb = ph == 0 ? &tmB8 : (ph == 1 ? &tmB82 : (ph == 2 ? &tmB : &tmB2));
          const int ptaps = (ph & 1) ? 1 : p.taps;
          const int unit = ph < 2 ? 128 : kBlockK;                       // coordinate units per chunk (128 B)
          const int wrap = ph == 2 ? p.a_wrap1 : (ph == 3 ? p.a_wrap2 : 0x7fffffff);
          // A stage = the patch of ONE chunk: (bh+2) x (bw+2) pixels for a 3x3 phase (one load serves all nine taps),
          // bh x bw pixels for a 1x1 phase (the tensor maps of the two kinds carry the two box shapes)
          const bool tap9 = ptaps == 9;
          const int a_bytes = tap9 ? p.patch_bytes : p.patch1_bytes;
          const int hh = tap9 ? p.halo : 0;
          mbar_wait(&a_empty[sa], pa ^ 1);
          uint8_t* abase = smem + sa * a_stage_bytes;
          if (elect_one()) {
            const int ac = (kc % wrap) * unit;
            if (kPair) {
              const uint32_t fb = mapa_u32(&a_full[sa], 0);   // pair: the leader's barrier collects both CTAs' bytes
              mbar_expect_tx_cluster(fb, a_bytes);
              tma_load_4d_2sm(ma, abase, fb, ac, w0 - hh, h0 - hh, b);
            } else {
              mbar_expect_tx(&a_full[sa], a_bytes);
              tma_load_4d(ma, abase, &a_full[sa], ac, w0 - hh, h0 - hh, b);
            }
          }
          __syncwarp();
          if (++sa == p.stages_a) {
            sa = 0;
            pa ^= 1;
          }
          const int nsteps = tap9 ? 9 : 1;
          const int tstep = tap9 ? p.tpb : 1;   // taps covered by one B stage (one TMA box over the tap dim)
          for (int st = 0; st < nsteps; st += tstep) {
            const int tap = tap9 ? st : 0;
            mbar_wait(&b_empty[sb], pb ^ 1);
            uint8_t* bbase = smem_b + sb * b_stage_bytes;
            const int b3 = (ph & 1) ? 0 : (p.b_batched ? b : tap);
            if (elect_one()) {
              if (kPair) {
                const uint32_t fb = mapa_u32(&b_full[sb], 0);
                mbar_expect_tx_cluster(fb, tstep * b_tile_bytes);
                tma_load_3d_2sm(mb, bbase, fb, kc * unit, brow, b3);
              } else {
                mbar_expect_tx(&b_full[sb], tstep * b_tile_bytes);
                tma_load_3d(mb, bbase, &b_full[sb], kc * unit, brow, b3);
              }
            }
            __syncwarp();
            if (++sb == p.stages_b) {
              sb = 0;
              pb ^= 1;
            }
          }
        };
        // contraction schedule (the MMA warp walks the same one): the e4m3 group, then the fp16 group; inside a group
        // the short k-chunks of a fused 1x1 skip conv (one MMA group each) are spread evenly between the long 3x3
        // chunks (nine taps each), so the patch ring is never asked for several short stages in a row
        for (int g = 0; g < 2; ++g) {
          const int nm = g == 0 ? p.kchunks8_1 : p.kchunks1;
          const int ns = g == 0 ? p.kchunks8_2 : p.kchunks2;
          int js = 0;
          for (int i = 0; i < nm; ++i) {
            load_chunk(2 * g, i);
            for (const int je = (i + 1) * ns / nm; js < je; ++js) load_chunk(2 * g + 1, js);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (pair: leader CTA only)
    // All 32 lanes run the loop so that barrier indices, descriptors and addresses are warp-uniform (uniform
    // registers feed tcgen05.mma directly — issued from a divergent single-lane branch every descriptor needs a
    // vector->uniform move and the issue loop, not the tensor pipe, bounds N = 128 layers); the MMAs and commits of a
    // k-block are issued by one elected lane.
    if (rank == 0) {
      const uint32_t idesc = make_idesc_f16(kPair ? 2 * kTileM : kTileM, p.n_tile);
      const uint32_t idesc8 = make_idesc_e4m3(kPair ? 2 * kTileM : kTileM, p.n_tile);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
        bool fresh = true;       // no MMA issued into this accumulator yet
        bool unscaled8 = false;  // fp8 corrections accumulated (x 2^14) and not yet folded
        auto mma_chunk = [&](int ph) {
          {
            const int ptaps = (ph & 1) ? 1 : p.taps;
            const bool f8 = ph < 2;
            // 8-row groups of the A tile: a patch row apart in a halo patch, dense (1024 B) otherwise
            const uint32_t a_sbo = ptaps == 9 ? static_cast<uint32_t>(p.a_pitch) : 1024u;
            const int nsteps = ptaps == 9 ? 9 : 1;
            if (!(p.dbg & 4)) mbar_wait(&a_full[sa], pa);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + sa * a_stage_bytes);
            const int tstep = ptaps == 9 ? p.tpb : 1;
            for (int st = 0; st < nsteps; st += tstep) {
              if (!(p.dbg & 4)) mbar_wait(&b_full[sb], pb);
              tc_fence_after();
              const uint32_t b_addr = smem_u32(smem_b + sb * b_stage_bytes);
              const bool last_step = (st + tstep >= nsteps);
              if (elect_one()) {
                for (int tt = 0; tt < tstep; ++tt) {
                  const int s1 = st + tt;
                  // 3x3 phase, tap s1 = 3 * ky + kx: the halo patch shifted by ky rows and kx pixels (whole 128-byte
                  // lines: the swizzle XOR follows the absolute address); 1x1 phase: the patch as loaded
                  const int a_off = ptaps == 9 ? (s1 / 3) * p.a_pitch + (s1 % 3) * 128 : 0;
                  const uint64_t da = make_sw128_kmajor_desc(a_addr + a_off, a_sbo);
                  const uint64_t db = make_sw128_kmajor_desc(b_addr + tt * b_tile_bytes);
                  const bool first = fresh && tt == 0;
                  if (p.mt == 2) {
                    // two stacked pixel tiles per CTA: the same weight tile feeds accumulator halves 0 / 1 (rows
                    // 0..bh-1 / bh..2bh-1 of the shared halo patch)
                    const uint64_t da1 = make_sw128_kmajor_desc(a_addr + a_off + p.bh * p.a_pitch, a_sbo);
                    const uint32_t d1 = d_tmem + (p.acc_stride >> 1);
                    if (f8) {
#pragma unroll
                      for (int k = 0; k < 4; ++k) {
                        const uint32_t accu = (first && k == 0) ? 0u : 1u;
                        if (kPair) {
                          umma_f8_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc8, accu);
                          umma_f8_2sm(d1, da1 + 2 * k, db + 2 * k, idesc8, accu);
                        } else {
                          umma_f8(d_tmem, da + 2 * k, db + 2 * k, idesc8, accu);
                          umma_f8(d1, da1 + 2 * k, db + 2 * k, idesc8, accu);
                        }
                      }
                    } else {
                      const bool fold = unscaled8 && tt == 0;
#pragma unroll
                      for (int k = 0; k < 4; ++k) {
                        const uint32_t accu = (first && k == 0) ? 0u : 1u;
                        if (fold && k == 0) {
                          if (kPair) {
                            umma_f16_scale_d14_2sm(d_tmem, da, db, idesc);
                            umma_f16_scale_d14_2sm(d1, da1, db, idesc);
                          } else {
                            umma_f16_scale_d14(d_tmem, da, db, idesc);
                            umma_f16_scale_d14(d1, da1, db, idesc);
                          }
                        } else if (kPair) {
                          umma_f16_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, accu);
                          umma_f16_2sm(d1, da1 + 2 * k, db + 2 * k, idesc, accu);
                        } else {
                          umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, accu);
                          umma_f16(d1, da1 + 2 * k, db + 2 * k, idesc, accu);
                        }
                      }
                    }
                  } else
                  // each MMA consumes 32 bytes of K per row (16 fp16 or 32 e4m3): +2 in (addr >> 4) units
                  if (f8) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      const uint32_t accu = (first && k == 0) ? 0u : 1u;
                      if (kPair) umma_f8_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc8, accu);
                      else umma_f8(d_tmem, da + 2 * k, db + 2 * k, idesc8, accu);
                    }
                  } else if (unscaled8 && tt == 0) {
                    // first fp16 block after the corrections: fold their 2^14 scale away (D = A*B + D * 2^-14)
                    if (kPair) umma_f16_scale_d14_2sm(d_tmem, da, db, idesc);
                    else umma_f16_scale_d14(d_tmem, da, db, idesc);
#pragma unroll
                    for (int k = 1; k < 4; ++k) {
                      if (kPair) umma_f16_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, 1u);
                      else umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, 1u);
                    }
                  } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      const uint32_t accu = (first && k == 0) ? 0u : 1u;
                      if (kPair) umma_f16_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, accu);
                      else umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, accu);
                    }
                  }
                }
                // frees the weight stage (in both CTAs of a pair) once the MMAs above have read it
                if (kPair) umma_commit_2sm(&b_empty[sb]);
                else umma_commit(&b_empty[sb]);
                // every step of this patch stage issued: it is free once they complete
                if (last_step) {
                  if (kPair) umma_commit_2sm(&a_empty[sa]);
                  else umma_commit(&a_empty[sa]);
                }
              }
              __syncwarp();
              unscaled8 = f8;
              fresh = false;
              if (++sb == p.stages_b) {
                sb = 0;
                pb ^= 1;
              }
            }
            if (++sa == p.stages_a) {
              sa = 0;
              pa ^= 1;
            }
          }
        };
        // same contraction schedule as the producer warp: e4m3 group, then fp16 group; skip-conv chunks interleaved
        for (int g = 0; g < 2; ++g) {
          const int nm = g == 0 ? p.kchunks8_1 : p.kchunks1;
          const int ns = g == 0 ? p.kchunks8_2 : p.kchunks2;
          int js = 0;
          for (int i = 0; i < nm; ++i) {
            mma_chunk(2 * g);
            for (const int je = (i + 1) * ns / nm; js < je; ++js) mma_chunk(2 * g + 1);
          }
        }
        // accumulator complete -> epilogue (of both CTAs of a pair)
        if (elect_one()) {
          if (kPair) umma_commit_2sm(&tmem_full[acc]);
          else umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++acc == p.acc_stages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (p.staged) {
    // ================================================================ epilogue, staged (4 warps, 128 threads)
    // Per 32-column chunk: TMEM -> registers -> (+bias row, +residual, *scale) -> swizzled smem chunk
    // [128 rows][128 B] -> one TMA store (hardware clips the image border); GroupNorm bundle statistics are taken
    // column-wise from the staged chunk (no warp shuffles over rows).  Two chunk buffers: the store of chunk i and
    // the residual load of chunk i+1 overlap the arithmetic.
    const int wq = warp & 3;
    const int et = wq * 32 + lane;  // 0..127 = accumulator row handled by this thread
    const int m = et;
    const int hl = m / p.bw;
    const int wl = m - hl * p.bw;
    const bool elected = (et == 0);
    const int nchunks = p.n_tile >> 5;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t cc = 0;                 // running chunk counter -> staging buffer parity
    uint32_t res_phase = 0;          // bit s = parity to wait for on res_bar[s]
    const int sw = (m & 7);          // swizzle phase of this thread's staging row
    const uint32_t tmem_empty_leader = kPair ? mapa_u32(tmem_empty, 0) : 0u;
    // GroupNorm bundle statistics: every thread keeps partial sums (bundle et & 7 of each 32-column chunk, its 8
    // rows of every tile) across the consecutive tiles this CTA works on and hands them to the global fp64
    // accumulators only when the (image, column block) changes — a persistent CTA stays inside one image for many
    // tiles, so the same-address fp64 atomics (all CTAs work on the same image at the same time) drop by that factor.
    // Round 1 issued them per chunk per warp: they serialised in L2 and cost 15 % of the 256->256 layer.  The
    // running sums are fp64 (of fp32 partials over a fixed 8-row x 4-channel footprint), so the result does not
    // depend on which tiles a CTA happens to get, i.e. not on the batch size or the grid.
    double st_s[8], st_q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) st_s[i] = st_q[i] = 0.0;
    int st_b = -1, st_col = 0;       // (image, first column) the partial sums belong to; -1: none
    auto flush_stats = [&]() {
      if (st_b >= 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < nchunks) {
            double s = st_s[c], q = st_q[c];
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            q += __shfl_xor_sync(0xffffffffu, q, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            q += __shfl_xor_sync(0xffffffffu, q, 16);
            if (lane < 8) {
              double* sp = p.stats + (static_cast<long long>(st_b) * (p.n_total >> 2) + ((st_col + c * 32) >> 2) + lane) * 2;
              atomicAdd(sp, s);
              atomicAdd(sp + 1, q);
            }
            st_s[c] = st_q[c] = 0.0;
          }
        }
      }
    };
    for (int t = t_begin; t < t_end; ++t) {
      const int nt = t % p.n_tiles;
      const int pt = kPair ? 2 * (t / p.n_tiles) + static_cast<int>(rank) : t / p.n_tiles;
      const int b = pt / tiles_per_img;
      const int r = pt - b * tiles_per_img;
      const int h0 = (r / p.tiles_w) * th;
      const int w0 = (r % p.tiles_w) * p.bw;
      const bool in_batch = b < p.batch;   // false only for the padding tile of an odd pair count
      const int ncol_base = nt * p.n_tile;
      if (p.stats && in_batch && (b != st_b || ncol_base != st_col)) {
        flush_stats();
        st_b = b;
        st_col = ncol_base;
      }
      // bias row of this tile (bias + per-image bias): every reader of the previous tile's row is past its last
      // chunk barrier, so it can be overwritten; the chunk-0 barrier below publishes it
      for (int i = et; i < p.n_tile; i += kEpiThreads) {
        float bv = 0.f;
        if (p.bias) bv += __ldg(p.bias + ncol_base + i);
        if (p.bias_b && in_batch) bv += __ldg(p.bias_b + static_cast<long long>(b) * p.n_total + ncol_base + i);
        bias_s[i] = bv;
        if (p.gnb_x) {
          // GroupNorm statistics of x for this channel's group, from the fp64 bundle sums (as gn_apply does)
          float rstd = 0.f, nmr = 0.f;
          if (in_batch) {
            const int ch = ncol_base + i;
            const int g0 = (ch / p.gnb_cpg) * p.gnb_cpg;
            double S = 0.0, Q = 0.0;
            for (int c = g0; c < g0 + p.gnb_cpg; c += 4) {
              const double* sp = p.gnb_stats + (static_cast<long long>(b) * (p.n_total >> 2) + (c >> 2)) * 2;
              S += sp[0];
              Q += sp[1];
            }
            const double n = static_cast<double>(p.gnb_cpg) * p.H * p.W;
            const double mean = S / n;
            double var = Q / n - mean * mean;
            if (var < 0.0) var = 0.0;
            rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.gnb_eps)));
            nmr = -static_cast<float>(mean) * rstd;
          }
          gnb_s[i] = rstd;
          gnb_s[p.n_tile + i] = nmr;
          gnb_s[2 * p.n_tile + i] = __ldg(p.gnb_gamma + ncol_base + i);
          gnb_s[3 * p.n_tile + i] = __ldg(p.gnb_beta + ncol_base + i);
        }
      }
      if (p.res_staged && elected) {
        // staging buffer cc&1 is free: the store that last read it was waited for before the previous chunk barrier
        mbar_expect_tx(&res_bar[cc & 1], kChunkBytes);
        tma_load_4d(&tmRes, stage_out + (cc & 1) * kChunkBytes, &res_bar[cc & 1], ncol_base, w0, h0, b);
      }
      named_bar_sync(1, kEpiThreads);

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * p.acc_stride;
      const int half_cols = p.acc_stride >> 1;      // TMEM column offset of the second stacked tile (mt == 2)
      const int nvc = nchunks * p.mt;               // virtual chunks of this work item: [half][32-column chunk]
      if (p.dbg & 3) {   // profiling experiment: no stores / statistics; dbg 2 still reads the accumulator
        if (p.dbg & 2) {
          uint32_t dd[32];
          for (int c = 0; c < nvc; ++c) {
            tmem_ld_32x32(taddr + (c >= nchunks ? half_cols + (c - nchunks) * 32 : c * 32), dd);
            tmem_ld_wait_dep(dd);
            if (dd[0] == 0x7fc12345u && dd[1] == 0x12345u) bias_s[0] = 1.f;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kPair) mbar_arrive_cluster(tmem_empty_leader + acc * 8);
          else mbar_arrive(&tmem_empty[acc]);
        }
        if (++acc == p.acc_stages) {
          acc = 0;
          acc_phase ^= 1;
        }
        continue;
      }
      uint32_t rr[32];
      tmem_ld_32x32(taddr, rr);
      tmem_ld_wait_dep(rr);
      for (int vc = 0; vc < nvc; ++vc, ++cc) {
        const int hf = vc >= nchunks ? 1 : 0;        // which stacked tile
        const int c = vc - hf * nchunks;             // 32-column chunk inside it
        const int h0v = h0 + hf * p.bh;
        const bool valid = in_batch && (h0v + hl < p.H) && (w0 + wl < p.W);
        const int sbuf = cc & 1;
        uint8_t* srow = stage_out + sbuf * kChunkBytes + m * 128;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
        if (vc + 1 < nvc) {
          // next virtual chunk's load in flight while this one is finished
          const int nv = vc + 1;
          tmem_ld_32x32(taddr + (nv >= nchunks ? half_cols + (nv - nchunks) * 32 : nv * 32), rr);
        } else {
          // every TMEM read of this accumulator has landed in registers: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kPair) mbar_arrive_cluster(tmem_empty_leader + acc * 8);
            else mbar_arrive(&tmem_empty[acc]);
          }
        }
        const float4* bp = reinterpret_cast<const float4*>(bias_s + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 q = bp[j];
          v[4 * j] += q.x;
          v[4 * j + 1] += q.y;
          v[4 * j + 2] += q.z;
          v[4 * j + 3] += q.w;
        }
        if (p.res_staged) {
          mbar_wait(&res_bar[sbuf], (res_phase >> sbuf) & 1u);
          res_phase ^= 1u << sbuf;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 q = *reinterpret_cast<const float4*>(srow + ((j ^ sw) << 4));
            v[4 * j] += q.x;
            v[4 * j + 1] += q.y;
            v[4 * j + 2] += q.z;
            v[4 * j + 3] += q.w;
          }
        }
        const float sc = valid ? p.scale : 0.f;   // rows outside the image: zero (clipped by the store, inert in the statistics)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(srow + ((j ^ sw) << 4)) =
              make_float4(v[4 * j] * sc, v[4 * j + 1] * sc, v[4 * j + 2] * sc, v[4 * j + 3] * sc);
        if (p.gnb_x && in_batch) {
          // GroupNorm-backward pass 0 for the tensor this launch differentiates through: per 4-channel bundle of this
          // thread's pixel, (sum dxh, sum dxh*xh); butterfly-reduced over the warp's 32 pixels (16 shuffles for the
          // 16 values), one fp64 atomic pair per bundle per warp.
          float part[16];
          {
            const long long pix = (static_cast<long long>(b) * p.H + (h0v + hl)) * p.W + (w0 + wl);
            const float4* xp = reinterpret_cast<const float4*>(p.gnb_x + pix * p.n_total + ncol_base + c * 32);
            const float* rs = gnb_s + c * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 x4 = valid ? __ldg(xp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
              float s1 = 0.f, s2 = 0.f;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int cc = 4 * j + q;
                const float xh = fmaf(xv[q], rs[cc], rs[p.n_tile + cc]);
                const float gam = rs[2 * p.n_tile + cc];
                float dz = v[cc] * sc;
                if (p.gnb_silu) {
                  const float z = fmaf(xh, gam, rs[3 * p.n_tile + cc]);
                  const float sg = __fdividef(1.f, 1.f + __expf(-z));
                  dz *= sg * fmaf(z, 1.f - sg, 1.f);
                }
                const float dxh = dz * gam;
                s1 += dxh;
                s2 = fmaf(dxh, xh, s2);
              }
              part[2 * j] = s1;
              part[2 * j + 1] = s2;
            }
          }
          // transpose-reduce: after the steps lane l holds the warp sum of value index (l >> 1) & 15
#pragma unroll
          for (int step = 0; step < 4; ++step) {
            const int half = 8 >> step;              // values kept after this step
            const int bit = 16 >> step;              // lane bit deciding which half is kept
            const bool upper = (lane & bit) != 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (i < half) {
                const float keep = upper ? part[i + half] : part[i];
                const float send = upper ? part[i] : part[i + half];
                part[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
              }
            }
          }
          part[0] += __shfl_xor_sync(0xffffffffu, part[0], 1);
          if ((lane & 1) == 0) {
            const int vi = lane >> 1;                // value index 0..15 = 2 * bundle + {0: sum dxh, 1: sum dxh*xh}
            const int ch = ncol_base + c * 32 + (vi >> 1) * 4;
            atomicAdd(p.gnb_gsum + (static_cast<long long>(b) * p.gnb_groups + ch / p.gnb_cpg) * 2 + (vi & 1),
                      static_cast<double>(part[0]));
          }
        }
        fence_proxy_async();
        if (elected) tma_store_wait_read0();   // the previous chunk's store no longer reads the other buffer
        named_bar_sync(1, kEpiThreads);
        if (elected) {
          if (!(p.dbg & 8)) {   // dbg 8 (profiling): everything but the output stores
            tma_store_4d(&tmOut, stage_out + sbuf * kChunkBytes, ncol_base + c * 32, w0, h0v, b);
            tma_store_commit();
          }
          if (p.res_staged && vc + 1 < nvc) {
            const int nv = vc + 1, nhf = nv >= nchunks ? 1 : 0;
            mbar_expect_tx(&res_bar[sbuf ^ 1], kChunkBytes);
            tma_load_4d(&tmRes, stage_out + (sbuf ^ 1) * kChunkBytes, &res_bar[sbuf ^ 1],
                        ncol_base + (nv - nhf * nchunks) * 32, w0, h0 + nhf * p.bh, b);
          }
        }
        if (p.stats && in_batch) {
          // thread -> (bundle = et & 7, rows 8*(et>>3) .. +8): column sums of the staged chunk
          const int bun = et & 7;
          const uint8_t* sbase = stage_out + sbuf * kChunkBytes + (et >> 3) * 1024;
          float s = 0.f, q = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 x = *reinterpret_cast<const float4*>(sbase + i * 128 + ((bun ^ i) << 4));
            s += (x.x + x.y) + (x.z + x.w);
            q = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, q))));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i == c) {
              st_s[i] += static_cast<double>(s);
              st_q[i] += static_cast<double>(q);
            }
        }
        if (vc + 1 < nvc) tmem_ld_wait_dep(rr);
      }
      if (++acc == p.acc_stages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (p.stats) flush_stats();
    if (elected) tma_store_wait_all();
  } else {
    // ================================================================ epilogue, direct (4 warps, 128 threads)
    const int wq = warp & 3;  // TMEM lane quarter this warp may access
    const int m = wq * 32 + lane;
    const int hl = m / p.bw;
    const int wl = m - hl * p.bw;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t tmem_empty_leader = kPair ? mapa_u32(tmem_empty, 0) : 0u;
    for (int t = t_begin; t < t_end; ++t) {
      const int nt = t % p.n_tiles;
      const int pt = kPair ? 2 * (t / p.n_tiles) + static_cast<int>(rank) : t / p.n_tiles;
      const int b = pt / tiles_per_img;
      const int r = pt - b * tiles_per_img;
      const int h = (r / p.tiles_w) * p.bh + hl;
      const int w = (r % p.tiles_w) * p.bw + wl;
      const bool valid = (b < p.batch) && (h < p.H) && (w < p.W);
      const long long pix = (static_cast<long long>(b) * p.H + h) * p.W + w;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * p.acc_stride;
      for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
        uint32_t rr[32];
        tmem_ld_32x32(taddr + c0, rr);
        tmem_ld_wait();
        const int ncol0 = nt * p.n_tile + c0;  // first global output column of this chunk
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
        const int nvalid = min(32, min(p.n_tile - c0, p.n_total - ncol0));
        if (nvalid <= 0) continue;
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) v[j] += __ldg(p.bias + ncol0 + j);
        }
        if (p.bias_b && b < p.batch) {
          const float* bb = p.bias_b + static_cast<long long>(b) * p.n_total + ncol0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) v[j] += __ldg(bb + j);
        }
        if (p.resid && valid) {
          const float* rp = p.resid + pix * p.ld_res + ncol0;
          if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(rp + j));
              v[j] += q.x;
              v[j + 1] += q.y;
              v[j + 2] += q.z;
              v[j + 3] += q.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) v[j] += __ldg(rp + j);
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.scale;

        if (valid) {
          if (p.out16) {
            __half* op = p.out16 + pix * p.ldc + p.col_off + ncol0;
            if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                __half2 h0 = __floats2half2_rn(v[j], v[j + 1]);
                __half2 h1 = __floats2half2_rn(v[j + 2], v[j + 3]);
                __half2 h2 = __floats2half2_rn(v[j + 4], v[j + 5]);
                __half2 h3 = __floats2half2_rn(v[j + 6], v[j + 7]);
                uint4 q;
                q.x = *reinterpret_cast<uint32_t*>(&h0);
                q.y = *reinterpret_cast<uint32_t*>(&h1);
                q.z = *reinterpret_cast<uint32_t*>(&h2);
                q.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(op + j) = q;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) op[j] = __float2half_rn(v[j]);
            }
          } else {
            float* op = p.out32 + pix * p.ldc + p.col_off + ncol0;
            if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) op[j] = v[j];
            }
          }
        }
        if (p.stats && b < p.batch) {
          // GroupNorm partial statistics of what was just written: per 4-channel bundle, over this warp's
          // 32 pixels, then one fp64 atomic pair per bundle per warp.
          double* sp = p.stats + (static_cast<long long>(b) * (p.n_total >> 2) + (ncol0 >> 2)) * 2;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float s = 0.f, q = 0.f;
            if (valid && g * 4 < nvalid) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float x = v[g * 4 + j];
                s += x;
                q += x * x;
              }
            }
            s = warp_sum(s);
            q = warp_sum(q);
            if (lane == 0 && g * 4 < nvalid) {
              atomicAdd(sp + g * 2, static_cast<double>(s));
              atomicAdd(sp + g * 2 + 1, static_cast<double>(q));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(tmem_empty_leader + acc * 8);
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == p.acc_stages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();   // no CTA leaves while its peer may still signal its barriers / read its smem
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_2sm(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}

// fp16 (esize 2), fp32 (esize 4) or byte (esize 1) tensor map with `rank` dims (dim 0 contiguous), 128B swizzle, zero OOB fill.
static int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_el,
                      const uint32_t* box, int esize = 2) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return BUDDY_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_el[i] * esize;  // bytes
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_last_error("tensor map base not 16B aligned");
    return BUDDY_ERR_INVALID;
  }
  for (int i = 0; i + 1 < rank; ++i)
    if (gstr[i] % 16 != 0) {
      set_last_error("tensor map stride %d (%llu bytes) not a multiple of 16", i + 1, (unsigned long long)gstr[i]);
      return BUDDY_ERR_INVALID;
    }
  const CUtensorMapDataType dt = esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                 : (esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
  CUresult r = fn(m, dt, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return BUDDY_ERR_CUDA;
  }
  return 0;
}

static void choose_patch(int H, int W, int* bh, int* bw) {
  long best = -1;
  for (int w = 128; w >= 1; w >>= 1) {
    const int h = 128 / w;
    const long tiles = static_cast<long>((H + h - 1) / h) * ((W + w - 1) / w);
    if (best < 0 || tiles < best) {
      best = tiles;
      *bh = h;
      *bw = w;
    }
  }
}

std::atomic<long long> g_launches{0};

// Shared-memory budget of one conv CTA (bytes).  Default: everything (227 KB).  BUDDY_CONV_SMEM_KB = n leaves
// 227 - n KB of every SM to co-resident CTAs of HBM-bound kernels launched on another stream (GroupNorm of another
// micro-batch: sampler.n_streams = 2), at the price of shallower pipeline rings.
static int conv_smem_budget() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("BUDDY_CONV_SMEM_KB");
    int kb = e ? atoi(e) : 227;
    if (kb < 128 || kb > 227) kb = 227;
    v = kb * 1024;
  }
  return v;
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

}  // namespace buddy

using namespace buddy;

extern "C" int buddy_conv_gemm(const buddy_gemm_desc* d, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!d || !d->a || !d->b || !d->out) {
    set_last_error("buddy_conv_gemm: null pointer");
    return BUDDY_ERR_INVALID;
  }
  if (d->a_c <= 0 || d->a_c % kBlockK != 0 || (d->a2 && (d->a2_c <= 0 || d->a2_c % kBlockK != 0))) {
    set_last_error("buddy_conv_gemm: channel count must be a positive multiple of 64 (got %d / %d)", d->a_c,
                   d->a2_c);
    return BUDDY_ERR_UNSUPPORTED;
  }
  if (d->n_tile < 16 || d->n_tile > 256 || d->n_tile % 16 != 0) {
    set_last_error("buddy_conv_gemm: n_tile must be a multiple of 16 in [16,256] (got %d)", d->n_tile);
    return BUDDY_ERR_UNSUPPORTED;
  }
  if (d->taps != 1 && d->taps != 9) {
    set_last_error("buddy_conv_gemm: taps must be 1 or 9 (got %d)", d->taps);
    return BUDDY_ERR_UNSUPPORTED;
  }
  if (d->batch <= 0 || d->H <= 0 || d->W <= 0 || d->n_total <= 0) {
    set_last_error("buddy_conv_gemm: empty problem (batch %d H %d W %d n %d)", d->batch, d->H, d->W, d->n_total);
    return BUDDY_ERR_INVALID;
  }
  if (d->stats && (d->n_total % 4 != 0 || d->out_fp16)) {
    set_last_error("buddy_conv_gemm: fused stats need fp32 output and n_total %% 4 == 0");
    return BUDDY_ERR_UNSUPPORTED;
  }
  if ((d->a2 != nullptr) != (d->b2 != nullptr)) {
    set_last_error("buddy_conv_gemm: a2 and b2 must be given together");
    return BUDDY_ERR_INVALID;
  }

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.batch = d->batch;
  p.H = d->H;
  p.W = d->W;
  // 3x3: fixed 16 x 8 pixel tiles with a one-pixel halo: ONE (16+2) x (8+2)-pixel patch per 64-channel chunk, 128-byte
  // lines, the nine taps are descriptor start offsets; 1x1 / GEMM: the patch shape that wastes least
  p.halo = d->taps == 9 ? 1 : 0;
  if (p.halo) {
    p.bh = 16;
    p.bw = 8;
  } else {
    choose_patch(d->H, d->W, &p.bh, &p.bw);
  }
  // staged epilogue (TMA stores of 128x32 fp32 chunks): dense fp32 output whose tile columns are whole chunks
  p.staged = (!d->out_fp16 && d->ldc == d->n_total && d->col_off == 0 && d->n_tile % 32 == 0 &&
              d->n_total % d->n_tile == 0 && (!d->resid || d->ld_res == d->n_total) && !d->no_staged_epilogue)
                 ? 1
                 : 0;
  p.res_staged = (p.staged && d->resid) ? 1 : 0;
  // two stacked pixel tiles per CTA (see GemmParams::mt): plain 3x3 launches with N <= 128 and the staged epilogue
  // ... when there is enough work: stacking halves the number of work items (small launches keep one tile per CTA)
  {
    const long long tiles1 = (long long)d->batch * ((d->H + 15) / 16) * ((d->W + 7) / 8);
    const long long items2 = ((tiles1 / 2 + 1) / 2) * ((d->n_total + d->n_tile - 1) / d->n_tile);
    p.mt = (p.halo && p.staged && d->n_tile <= 128 && !d->a2 && !d->gnb_x && d->single_tile_per_cta != 1 && d->H > 16 &&
            (items2 >= 2 * (num_sms() / 2) || d->single_tile_per_cta == 2))
               ? 2
               : 1;
  }
  const int patch_rows = p.bh * p.mt + 2 * p.halo;
  const int patch_cols = p.bw + 2 * p.halo;
  p.a_pitch = patch_cols * 128;
  p.patch_bytes = patch_rows * patch_cols * 128;
  p.patch1_bytes = p.bh * p.bw * 128;
  p.a_stage_bytes = (p.patch_bytes + 1023) & ~1023;
  p.tiles_h = (d->H + p.bh * p.mt - 1) / (p.bh * p.mt);
  p.tiles_w = (d->W + p.bw - 1) / p.bw;
  p.n_tile = d->n_tile;
  p.n_total = d->n_total;
  p.n_tiles = (d->n_total + d->n_tile - 1) / d->n_tile;
  p.taps = d->taps;
  const int k1 = d->k_total > 0 ? d->k_total : d->a_c;
  const int k2 = d->a2 ? (d->k2_total > 0 ? d->k2_total : d->a2_c) : 0;
  if (k1 % kBlockK || k2 % kBlockK) {
    set_last_error("buddy_conv_gemm: k_total must be a multiple of 64");
    return BUDDY_ERR_UNSUPPORTED;
  }
  p.kchunks1 = k1 / kBlockK;
  p.kchunks2 = k2 / kBlockK;
  p.a_wrap1 = d->a_c / kBlockK;
  p.a_wrap2 = d->a2 ? d->a2_c / kBlockK : 1;
  p.b_batched = d->b_batched;
  // CTA pairs (cta_group::2): whenever the weight tile can be split in two swizzle-aligned halves
  const long long pix_tiles = (long long)p.batch * p.tiles_h * p.tiles_w;
  const bool pair = !d->b_batched && !d->no_cta_pairs && d->n_tile >= 32 && d->n_tile % 16 == 0 && pix_tiles >= 2;
  const int b_box_rows = pair ? d->n_tile / 2 : d->n_tile;
  const int a_stage_bytes = p.a_stage_bytes;
  const int b_stage_bytes = b_box_rows * 128;
  if (d->gnb_x) {
    if (!p.staged || !d->gnb_stats || !d->gnb_gamma || !d->gnb_beta || !d->gnb_gsum || d->gnb_groups <= 0 ||
        d->n_total % d->gnb_groups || (d->n_total / d->gnb_groups) % 4) {
      set_last_error("buddy_conv_gemm: gnb_* needs the staged epilogue (dense fp32 output, n_tile %% 32 == 0) and "
                     "channels-per-group a multiple of 4");
      return BUDDY_ERR_UNSUPPORTED;
    }
  }
  // TMEM: 512 columns = 2 accumulators of up to 256 columns, or 4 of up to 128 (more slack for the epilogue of the
  // short N <= 128 mainloops)
  p.acc_stages = (d->n_tile <= 128 && p.mt == 1) ? 4 : 2;
  p.acc_stride = (d->n_tile <= 128 && p.mt == 1) ? 128 : 256;
  const int epi_bytes = p.staged ? 2 * kChunkBytes + 1024 + (d->gnb_x ? 4096 : 0) : 0;
  const int ring_bytes = conv_smem_budget() - 1024 /*alignment slack*/ - 512 /*barriers*/ - epi_bytes;
  p.tpb = 1;
  if (p.halo) {
    // a patch stage lasts nine weight tiles: three of them, the rest of the shared memory goes to the weight ring —
    // as kernel rows of three taps per stage when at least three such stages fit (N <= 128 in pair mode)
    // (a fused 1x1 skip phase drains one patch stage per k-block: one more stage keeps the ring ahead of it)
    p.stages_a = d->a2 ? 4 : (p.mt == 2 ? 2 : 3);
    if (!d->one_tap_per_stage && (ring_bytes - p.stages_a * a_stage_bytes) / (3 * b_stage_bytes) >= 3) p.tpb = 3;
    p.stages_b = (ring_bytes - p.stages_a * a_stage_bytes) / (p.tpb * b_stage_bytes);
  } else {
    p.stages_a = p.stages_b = ring_bytes / (a_stage_bytes + b_stage_bytes);
  }
  if (p.stages_a > kMaxStagesA) p.stages_a = kMaxStagesA;
  if (p.stages_b > kMaxStagesB) p.stages_b = kMaxStagesB;
  if (p.stages_a < 2 || p.stages_b < 2) {
    set_last_error("buddy_conv_gemm: not enough shared memory for 2 pipeline stages (n_tile %d)", d->n_tile);
    return BUDDY_ERR_UNSUPPORTED;
  }
  p.out32 = d->out_fp16 ? nullptr : static_cast<float*>(d->out);
  p.out16 = d->out_fp16 ? static_cast<__half*>(d->out) : nullptr;
  p.ldc = d->ldc;
  p.col_off = d->col_off;
  p.bias = d->bias;
  p.bias_b = d->bias_b;
  p.resid = d->resid;
  p.ld_res = d->ld_res;
  p.scale = d->scale;
  p.stats = d->stats;
  p.dbg = d->debug_flags;
  p.gnb_x = d->gnb_x;
  p.gnb_stats = d->gnb_stats;
  p.gnb_gamma = d->gnb_gamma;
  p.gnb_beta = d->gnb_beta;
  p.gnb_gsum = d->gnb_gsum;
  p.gnb_groups = d->gnb_groups;
  p.gnb_cpg = d->gnb_groups > 0 ? d->n_total / d->gnb_groups : 0;
  p.gnb_silu = d->gnb_silu;
  p.gnb_eps = d->gnb_eps;

  CUtensorMap tmA, tmB, tmA2, tmB2, tmA8, tmB8, tmA82, tmB82, tmOut, tmRes;
  p.kchunks8_1 = 0;
  p.kchunks8_2 = 0;
  if (d->a8) {
    if (!d->b8 || d->a8_c <= 0 || d->a8_c % 128 || d->b_batched) {
      set_last_error("buddy_conv_gemm: fp8 correction operands need b8, a8_c %% 128 == 0 and no batched B");
      return BUDDY_ERR_UNSUPPORTED;
    }
    p.kchunks8_1 = d->a8_c / 128;
    uint64_t dims[4] = {(uint64_t)d->a8_c, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->batch};
    uint64_t str[4] = {1, (uint64_t)d->a8_stride_w, (uint64_t)d->a8_stride_h, (uint64_t)d->a8_stride_b};
    uint32_t box[4] = {128, (uint32_t)patch_cols, (uint32_t)patch_rows, 1};
    uint32_t box1[4] = {128, (uint32_t)p.bw, (uint32_t)p.bh, 1};
    int e = encode_map(&tmA8, d->a8, 4, dims, str, box, 1);
    if (e) return e;
    uint64_t dimsb[3] = {(uint64_t)d->a8_c, (uint64_t)d->b_rows, (uint64_t)d->b_t};
    uint64_t strb[3] = {1, (uint64_t)d->a8_c, (uint64_t)d->a8_c * (uint64_t)d->b_rows};
    uint32_t boxb8[3] = {128, (uint32_t)b_box_rows, (uint32_t)p.tpb};
    uint32_t boxb[3] = {128, (uint32_t)b_box_rows, 1};
    e = encode_map(&tmB8, d->b8, 3, dimsb, strb, boxb8, 1);
    if (e) return e;
    if (d->a2) {
      if (!d->a8_2 || !d->b8_2 || d->a8_2_c % 128) {
        set_last_error("buddy_conv_gemm: fp8 correction of the skip conv needs a8_2 / b8_2");
        return BUDDY_ERR_INVALID;
      }
      p.kchunks8_2 = d->a8_2_c / 128;
      uint64_t dims2[4] = {(uint64_t)d->a8_2_c, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->batch};
      uint64_t str2[4] = {1, (uint64_t)d->a8_2_stride_w, (uint64_t)d->a8_2_stride_h, (uint64_t)d->a8_2_stride_b};
      e = encode_map(&tmA82, d->a8_2, 4, dims2, str2, box1, 1);
      if (e) return e;
      uint64_t dimsb2[3] = {(uint64_t)d->a8_2_c, (uint64_t)d->b2_rows, 1};
      uint64_t strb2[3] = {1, (uint64_t)d->a8_2_c, (uint64_t)d->a8_2_c * (uint64_t)d->b2_rows};
      e = encode_map(&tmB82, d->b8_2, 3, dimsb2, strb2, boxb, 1);
      if (e) return e;
    }
  }
  {
    uint64_t dims[4] = {(uint64_t)d->a_c, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->batch};
    uint64_t str[4] = {1, (uint64_t)d->a_stride_w, (uint64_t)d->a_stride_h, (uint64_t)d->a_stride_b};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)patch_cols, (uint32_t)patch_rows, 1};
    int e = encode_map(&tmA, d->a, 4, dims, str, box);
    if (e) return e;
  }
  {
    uint64_t dims[3] = {(uint64_t)k1, (uint64_t)d->b_rows, (uint64_t)d->b_t};
    uint64_t str[3] = {1, (uint64_t)d->b_stride_n, (uint64_t)d->b_stride_t};
    uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)b_box_rows, (uint32_t)p.tpb};
    int e = encode_map(&tmB, d->b, 3, dims, str, box);
    if (e) return e;
  }
  if (d->a2) {
    uint64_t dims[4] = {(uint64_t)d->a2_c, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->batch};
    uint64_t str[4] = {1, (uint64_t)d->a2_stride_w, (uint64_t)d->a2_stride_h, (uint64_t)d->a2_stride_b};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.bw, (uint32_t)p.bh, 1};
    int e = encode_map(&tmA2, d->a2, 4, dims, str, box);
    if (e) return e;
    uint64_t dimsb[3] = {(uint64_t)k2, (uint64_t)d->b2_rows, 1};
    uint64_t strb[3] = {1, (uint64_t)d->b2_stride_n, (uint64_t)d->b2_stride_n * (uint64_t)d->b2_rows};
    uint32_t boxb[3] = {(uint32_t)kBlockK, (uint32_t)b_box_rows, 1};
    e = encode_map(&tmB2, d->b2, 3, dimsb, strb, boxb);
    if (e) return e;
  } else {
    tmA2 = tmA;
    tmB2 = tmB;
  }
  if (!d->a8) {
    tmA8 = tmA;
    tmB8 = tmB;
  }
  if (p.kchunks8_2 == 0) {
    tmA82 = tmA;
    tmB82 = tmB;
  }
  if (p.staged) {
    uint64_t dims[4] = {(uint64_t)d->n_total, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->batch};
    uint64_t str[4] = {1, (uint64_t)d->n_total, (uint64_t)d->n_total * d->W, (uint64_t)d->n_total * d->W * d->H};
    uint32_t box[4] = {32, (uint32_t)p.bw, (uint32_t)p.bh, 1};
    int e = encode_map(&tmOut, d->out, 4, dims, str, box, 4);
    if (e) return e;
    if (p.res_staged) {
      e = encode_map(&tmRes, d->resid, 4, dims, str, box, 4);
      if (e) return e;
    } else {
      tmRes = tmOut;
    }
  } else {
    tmOut = tmA;
    tmRes = tmA;
  }

  const size_t smem_bytes = 1024 /*run-time alignment slack*/ + (size_t)p.stages_a * a_stage_bytes +
                            (size_t)p.stages_b * p.tpb * b_stage_bytes + epi_bytes + 512 /*barriers*/;
  static bool attr_set = false;
  if (!attr_set) {
    int e = check_cuda(
        cudaFuncSetAttribute(conv_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
        "cudaFuncSetAttribute(conv_gemm_kernel)");
    if (e) return e;
    e = check_cuda(
        cudaFuncSetAttribute(conv_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
        "cudaFuncSetAttribute(conv_gemm_kernel<pair>)");
    if (e) return e;
    attr_set = true;
  }
  if (!pair) {
    const long long total_tiles = pix_tiles * p.n_tiles;
    int grid = num_sms();
    if (d->max_ctas > 0 && d->max_ctas < grid) grid = d->max_ctas;
    if (total_tiles < grid) grid = (int)total_tiles;
    conv_gemm_kernel<false><<<grid, kThreads, smem_bytes, stream>>>(tmA, tmB, tmA2, tmB2, tmA8, tmB8, tmA82, tmB82,
                                                                    tmOut, tmRes, p);
  } else {
    const long long work = ((pix_tiles + 1) / 2) * p.n_tiles;
    // persistent pairs: as many 2-CTA clusters as can be co-resident (74 on a full B200; fewer if a GPC has an odd
    // number of usable SMs) — queried once with the largest shared-memory footprint
    static int max_clusters = 0;
    if (max_clusters == 0) {
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3(num_sms(), 1, 1);
      q.blockDim = dim3(kThreads, 1, 1);
      q.dynamicSmemBytes = 227 * 1024;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, conv_gemm_kernel<true>, &q) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        n = num_sms() / 2;
      }
      max_clusters = n < num_sms() / 2 ? n : num_sms() / 2;
    }
    int clusters = max_clusters;
    if (d->max_ctas > 0 && d->max_ctas / 2 < clusters) clusters = d->max_ctas / 2 > 0 ? d->max_ctas / 2 : 1;
    if (work < clusters) clusters = (int)work;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int e = check_cuda(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<true>, tmA, tmB, tmA2, tmB2, tmA8, tmB8, tmA82, tmB82,
                                          tmOut, tmRes, p),
                       "cudaLaunchKernelEx(conv_gemm_kernel<pair>)");
    if (e) return e;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  BUDDY_CHECK_LAUNCH("conv_gemm_kernel");
  return 0;
}
