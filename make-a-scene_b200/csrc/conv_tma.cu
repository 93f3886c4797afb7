// 3x3 stride-1 convolution (fprop / data gradient) as a pure TMA + tcgen05 kernel: the A operand comes from an fp16 "shadow"
// of the activation (written channels-last by the producing kernel: GroupNorm(+SiLU) apply in the forward pass, GroupNorm
// backward in the backward pass) instead of being converted by producer warps.
//
//   * one persistent CTA per SM walks work items (pair of 128-pixel tiles x 128 output channels), as shift_gemm_p16;
//   * A: ONE cp.async.bulk.tensor (4-D map over [N][H][W][C] halves, box 1 x 18 x 10 x 64, 128-byte swizzle, out-of-image
//     halo pixels zero-filled by the copy engine) per tile per 64-channel chunk.  The staged halo is [180 pixels][128 B];
//     the nine taps and the four K = 16 steps of the chunk are descriptor start-address shifts ((ty*10+tx)*128 + k*32 bytes)
//     over that one copy: the 8-row core group is eight horizontally adjacent pixels, the group stride (SBO) one staged image
//     row = 1280 B.  Row-shifted starts under the 128-byte swizzle are address-exact on sm_100 (the XOR is taken from the
//     absolute shared-memory address bits; tools/probe_sw128.py pins it);
//   * B (weights): the same pre-packed no-swizzle stages as shift_gemm_tc (mas_pack_conv3x3_tc16), one bulk copy per
//     16-channel step, ring of three;
//   * operand roles are swapped (D^T = W x X^T: the packed weights are the M-side operand, the pixels the N side), so a TMEM
//     lane is an output channel and a column a pixel: the epilogue's 32 lanes store 32 consecutive channels of one pixel -
//     a full 128-byte line per instruction without a shared-memory transpose; bias and GroupNorm statistics are per-thread;
//   * two accumulator sets in tensor memory (2 x 256 columns): eight epilogue warps drain set b (bias / residual /
//     GroupNorm-statistics epilogue) while the MMAs of the next item fill set b^1;
//   * warps 0-7 epilogue, warp 8 MMA issuer, warp 9 copy issuer: no thread of the CTA touches the operands.
//
// Reference call sites replaced: nn.Conv2d 3x3 stride 1 (modules.py:93-104) forward and its data gradient.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "mas_common.cuh"
#include "tc_ptx.cuh"

namespace mas {

PFN_cuTensorMapEncodeTiled tensor_map_encoder();   // contract_tc.cu

namespace tc {

constexpr int T_EPI_WARPS = 8;
constexpr int T_THREADS = (T_EPI_WARPS + 2) * 32;
constexpr int T_ASTAGES = 2, T_BSTAGES = 3;
constexpr int T_ATILE = 23 * 1024;                  // 18 x 10 halo pixels x 128 B = 23040, padded to the 1024-byte swizzle atom
constexpr int T_ASTAGE = 2 * T_ATILE;               // pair of 16 x 8 tiles, or one 32 x 8 tile (34 x 10 halo = 43520 B)
constexpr int T_BSTAGE = 9 * 2 * BN * 16;           // nine taps x 16 channels x 128 output channels (fp16)

struct HParams {
  const void* wpk;    // mas_pack_conv3x3_tc16 packing
  const float* bias;  // [Cout] or null
  const float* res;   // NHWC fp32 like y, or null
  float* y;
  int N, H, W, Cin, Cout, Cstore;
  int64_t ldy;
  int tiles_x, tiles_y;   // 16 x 8 tiles per image row / column
  int64_t units;          // work units of 256 pixels: 32 x 8 tiles (TALL) or pairs of consecutive 16 x 8 tiles
  float* stats_part;   // GroupNorm-statistics epilogue (see shift_gemm_tc), or null
  const float* x_amax; // amax the shadow's power-of-two scale was derived from (null: unscaled shadow)
};

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}

// 16 x 8 tile (n, ty, tx) of half `hf` of work unit `u`; false when the unit's second tile does not exist (odd tile count)
template <bool TALL>
__device__ __forceinline__ bool unit_tile(const HParams& p, int64_t u, int hf, int& n, int& ty, int& tx) {
  if (TALL) {
    const int ty2 = p.tiles_y >> 1;
    tx = (int)(u % p.tiles_x);
    ty = (int)((u / p.tiles_x) % ty2) * 2 + hf;
    n = (int)(u / ((int64_t)p.tiles_x * ty2));
    return true;
  }
  const int64_t total = (int64_t)p.N * p.tiles_x * p.tiles_y;
  int64_t t = u * 2 + hf;
  const bool live = t < total;
  if (!live) t = total - 1;
  tx = (int)(t % p.tiles_x);
  ty = (int)((t / p.tiles_x) % p.tiles_y);
  n = (int)(t / ((int64_t)p.tiles_x * p.tiles_y));
  return live;
}

// TALL: the unit is one 32 x 8 tile whose staged halo (34 x 10 pixels, uniform 1280-byte row pitch) is ONE N = 256 operand:
// per K = 16 step and tap a single M128 x N256 MMA reads 4 KB of weights + 8 KB of pixels instead of 2 x (4 + 4) KB - the
// SS-mode kernel is bound by exactly that operand traffic (measured: no change with the weight copies stubbed out).
template <bool TALL>
__global__ void __launch_bounds__(T_THREADS, 1) shift_gemm_t16(const HParams p, const __grid_constant__ CUtensorMap x_map) {
  constexpr int TAPS = 9;
  constexpr int LBO_B = BN * 16, B_TAP = 2 * LBO_B;
  constexpr int A_BYTES = TALL ? 34 * 10 * 128 : 2 * 18 * 10 * 128;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_base = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_base + 1023u) & ~1023u;           // swizzle atoms are 1024-byte aligned
  uint8_t* smem = smem_raw + (smem_base - raw_base);
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + T_ASTAGES * T_ASTAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T_ASTAGES * T_ASTAGE + T_BSTAGES * T_BSTAGE);
  constexpr int NBARS = 2 * T_ASTAGES + 2 * T_BSTAGES + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
  const uint32_t bar_base = smem_u32(bars);
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (T_ASTAGES + s); };
  auto bfull = [&](int s) { return bar_base + 8u * (2 * T_ASTAGES + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (2 * T_ASTAGES + T_BSTAGES + s); };
  auto accf = [&](int b) { return bar_base + 8u * (2 * T_ASTAGES + 2 * T_BSTAGES + b); };
  auto acce = [&](int b) { return bar_base + 8u * (2 * T_ASTAGES + 2 * T_BSTAGES + 2 + b); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int achunks = p.Cin / 64;
  const int n_tiles = p.Cout / BN;
  const int64_t nitems = p.units * n_tiles;       // channel tile fastest: the halo of a unit is re-read from L2

  if (tid == 0) {
    for (int s = 0; s < T_ASTAGES; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < T_BSTAGES; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(accf(b), 1); mbar_init(acce(b), T_EPI_WARPS * 32); }
    fence_barrier_init();
  }
  if (warp == T_EPI_WARPS) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < T_EPI_WARPS) {
    // ===================== epilogue warps =====================
    // The accumulator is D^T: TMEM lane = output channel, column = pixel (weights are the M-side operand).  A warp owns 32
    // consecutive channels (lane quarter warp % 4) of one 16 x 8 half of the unit (warp / 4); for every pixel its 32 lanes
    // store 32 consecutive floats = one full 128-byte line: coalesced without a shared-memory transpose, bias and GroupNorm
    // statistics are per-thread scalars.
    float inv_scale = 1.f;
    operand_scale(p.x_amax, &inv_scale);
    const float alpha = inv_scale;
    const int quarter = warp & 3, hf = warp >> 2;
    int buf = 0;
    uint32_t ph[2] = {0u, 0u};
    for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
      int n_img, ty_, tx_;
      const bool live = unit_tile<TALL>(p, item / n_tiles, hf, n_img, ty_, tx_);   // warp-uniform
      const int ch = (int)(item % n_tiles) * BN + quarter * 32 + lane;
      const bool st_ok = ch < p.Cstore;
      const float bv = (p.bias && st_ok) ? __ldg(p.bias + ch) : 0.f;
      const int64_t pix0 = ((int64_t)n_img * p.H + ty_ * 16) * p.W + tx_ * 8;
      // residual: the 32 loads of a 32-pixel batch are issued ONE BATCH AHEAD (the first before the accumulator is even
      // waited for), so 32 KB per SM are in flight - the epilogue was bound by this latency, not by the stores.  (An L2
      // prefetch of the next item's residual tile on top of this measured SLOWER - 0.66 vs 0.61 ms - and fetched 44 % of the
      // residual twice: the output stream evicts the prefetched lines.)
      const int rs = p.W * (int)p.ldy, ps = (int)p.ldy;      // element strides of an image row / a pixel (tile-local: fits int)
      const int64_t base0 = pix0 * p.ldy + ch;
      const bool use_res = p.res != nullptr && live && st_ok;
      float ra[32], rb[32];
      auto rload = [&](int cb, float* r) {
        const float* rp = p.res + base0 + (int64_t)(cb * 4) * rs;
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __ldg(rp + (j >> 3) * rs + (j & 7) * ps);
      };
      if (use_res) rload(0, ra);
      mbar_wait(accf(buf), ph[buf]);
      ph[buf] ^= 1u;
      tc_fence_after();
      auto batch = [&](int cb, const float* r) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 256 + hf * 128 + cb * 32), v);
        if (cb == 3) {
          tc_fence_before();
          mbar_arrive(acce(buf));     // this warp's share of the accumulator set is in registers
        }
        if (!live) return;
        float st_s = 0.f, st_q = 0.f;
        if (st_ok) {
          float* yp = p.y + base0 + (int64_t)(cb * 4) * rs;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float o = fmaf(v[j], alpha, bv);
            if (use_res) o += r[j];
            yp[(j >> 3) * rs + (j & 7) * ps] = o;
            st_s += o;
            st_q = fmaf(o, o, st_q);
          }
        }
        if (p.stats_part) {   // same partial layout as shift_gemm_tc: [16 x 8 tile][32-pixel group][channel quad][sum, sumsq]
          st_s += __shfl_xor_sync(0xffffffffu, st_s, 1);
          st_q += __shfl_xor_sync(0xffffffffu, st_q, 1);
          st_s += __shfl_xor_sync(0xffffffffu, st_s, 2);
          st_q += __shfl_xor_sync(0xffffffffu, st_q, 2);
          if ((lane & 3) == 0) {
            const size_t tile = ((size_t)n_img * p.tiles_y + ty_) * p.tiles_x + tx_;
            float* sp = p.stats_part + ((tile * 4 + cb) * (p.Cout >> 2) + (ch >> 2)) * 2;
            sp[0] = st_s;
            sp[1] = st_q;
          }
        }
      };
      if (use_res) rload(1, rb);
      batch(0, ra);
      if (use_res) rload(2, ra);
      batch(1, rb);
      if (use_res) rload(3, rb);
      batch(2, ra);
      batch(3, rb);
      buf ^= 1;
    }
  } else if (warp == T_EPI_WARPS) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop (warp-uniform control flow keeps the operand descriptors in uniform registers: one add
    // per descriptor per MMA instead of a 64-bit vector add + five R2UR moves - the single-thread form spent ~23 instructions
    // per MMA and could not run ahead of the tensor pipe); one elected lane issues the MMAs and commits.
    {
      constexpr uint32_t idesc = make_idesc_f16(TALL ? 256 : BN);
      int as = 0, bs = 0, buf = 0;
      uint32_t aph = 0, bph = 0, eph[2] = {0u, 0u};
      for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        mbar_wait(acce(buf), eph[buf] ^ 1);     // the epilogue warps have read this accumulator set (first use: passes)
        eph[buf] ^= 1u;
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * 256);
        for (int c = 0; c < achunks; ++c) {
          mbar_wait(afull(as), aph);
          tc_fence_after();
          const uint64_t xd0 = make_desc_sw128(a_base + (uint32_t)as * T_ASTAGE, 1280);
#pragma unroll 1
          for (int sub = 0; sub < 4; ++sub) {
            mbar_wait(bfull(bs), bph);
            tc_fence_after();
            const uint64_t wd0 = make_desc(b_base + (uint32_t)bs * T_BSTAGE, LBO_B, 128);
            const uint64_t xds = xd0 + (uint64_t)((sub * 32) >> 4);
            const uint32_t acc0 = (c > 0 || sub > 0) ? 1u : 0u;
            if (elect_one()) {
#pragma unroll
              for (int t = 0; t < TAPS; ++t) {
                const uint32_t tapoff = (uint32_t)(((t / 3) * 10 + (t % 3)) * 128);
                const uint64_t wd = wd0 + (uint64_t)((t * B_TAP) >> 4);
                const uint64_t xd = xds + (uint64_t)(tapoff >> 4);
                // D^T = W x X^T: weights on the M side, pixels on the N side
                mma_f16_ss(acc, wd, xd, idesc, t > 0 ? 1u : acc0);
                if (!TALL) mma_f16_ss(acc + 128u, wd, xd + (uint64_t)(T_ATILE >> 4), idesc, t > 0 ? 1u : acc0);
              }
              mma_commit(bempty(bs));
              if (sub == 3) mma_commit(aempty(as));
              if (sub == 3 && c == achunks - 1) mma_commit(accf(buf));
            }
            __syncwarp();
            if (++bs == T_BSTAGES) { bs = 0; bph ^= 1; }
          }
          if (++as == T_ASTAGES) { as = 0; aph ^= 1; }
        }
        buf ^= 1;
      }
    }
  } else {
    // ===================== copy issuer (one thread): halos by tensor map, weight stages by bulk copy =====================
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      const int kchunks = p.Cin / 16;
      for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpk) + (size_t)(item % n_tiles) * kchunks * T_BSTAGE;
        for (int c = 0; c < achunks; ++c) {
          mbar_wait(aempty(as), aph ^ 1);
          mbar_expect_tx(afull(as), A_BYTES);
#pragma unroll
          for (int hf = 0; hf < (TALL ? 1 : 2); ++hf) {
            int n, ty_, tx_;
            unit_tile<TALL>(p, item / n_tiles, hf, n, ty_, tx_);   // a missing second tile re-reads the last one (never stored)
            tma_load_4d(a_base + (uint32_t)(as * T_ASTAGE + hf * T_ATILE), &x_map, c * 64, tx_ * 8 - 1, ty_ * 16 - 1, n, afull(as));
          }
          if (++as == T_ASTAGES) { as = 0; aph ^= 1; }
          for (int sub = 0; sub < 4; ++sub) {
            mbar_wait(bempty(bs), bph ^ 1);
            mbar_expect_tx(bfull(bs), T_BSTAGE);
            bulk_g2s(b_base + (uint32_t)bs * T_BSTAGE, wsrc + (size_t)(c * 4 + sub) * T_BSTAGE, T_BSTAGE, bfull(bs));
            if (++bs == T_BSTAGES) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == T_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

constexpr size_t t16_smem_bytes() {
  return 1024 + (size_t)T_ASTAGES * T_ASTAGE + (size_t)T_BSTAGES * T_BSTAGE + (2 * T_ASTAGES + 2 * T_BSTAGES + 4) * 8 + 16;
}

// ------------------------------------------------------------------------------------------------------------
// Weight gradient of the 3x3 stride-1 convolution from the two fp16 shadows (activation x16, output gradient dy16 - already
// scaled), dW[tap][co][ci] = sum_pixels dy[p][co] * x[p + tap][ci]; the reduction runs over PIXELS, so both operands are
// "MN-major" in memory (channels contiguous, pixels strided):
//   * A = dy^T in tensor memory (TS mode): the dy tile of a unit (8 x 8 pixels x 128 co, one tensor-map copy) is moved
//     shared -> registers -> TMEM by four loader warps, two pixels per 32-bit column, lane = co (the transpose is free);
//   * B = the activation halo exactly as the copy engine lands it under the 128-byte swizzle: [row][pixel][64 ci] with 128-byte
//     pixel rows, read as an MN-major operand (K groups = image rows of 8 pixels, SBO = the 1280-byte halo row; probe:
//     tools/probe_mn128.py).  A tap is a descriptor start shift (dx * 128 B); no producer warps, no dx copies;
//   * a CTA owns one kernel ROW (3 horizontal taps) of a 128 co x NCI ci block: 3 x NCI accumulator columns, N = 64 MMAs
//     (two per tap and K step for NCI = 128), K = 16 pixels = two image rows of the unit;
//   * split-K over the units; partial sums (and the bias gradient from the dy loaders) go to the caller's workspace in the
//     layout conv_wgrad_reduce expects.
constexpr int WT_STAGES = 4;
constexpr int WT_XATOM = 8 * 10 * 128;      // 8 halo rows x 10 pixels x 128 B (64 channels)
constexpr int WT_DY = 64 * 128 * 2;         // 64 pixels x 128 co halves
constexpr int WT_THREADS = 6 * 32;

struct WTParams {
  float* part;      // [splits][9][Cout][Cin]
  float* bpart;     // [splits][Cout] or null
  int N, H, W, Cin, Cout;
  int units_x, units_y;
  int64_t total_units, units_per_split;
  const float* dy_amax;   // the magnitude dy16's power-of-two scale was derived from
};

template <int NCI>
__global__ void __launch_bounds__(WT_THREADS, 1) wgrad_t16(const WTParams p, const __grid_constant__ CUtensorMap x_map,
                                                          const __grid_constant__ CUtensorMap dy_map) {
  constexpr int XB = (NCI / 64) * WT_XATOM, STAGE = XB + WT_DY;
  constexpr uint32_t ACC_COLS = 3 * NCI, A_COLS = 32;
  static_assert(ACC_COLS + WT_STAGES * A_COLS <= 512, "tensor memory budget");
  constexpr uint32_t idesc = make_idesc_f16(64) | (1u << 16);   // B operand MN-major

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_base = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_base + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_base);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WT_STAGES * STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * WT_STAGES + 1);
  const uint32_t bar_base = smem_u32(bars);
  auto fullD = [&](int s) { return bar_base + 8u * s; };                      // copies of the stage have landed
  auto fullA = [&](int s) { return bar_base + 8u * (WT_STAGES + s); };        // dy^T of the stage is in tensor memory
  auto empty = [&](int s) { return bar_base + 8u * (2 * WT_STAGES + s); };    // the MMAs of the stage have completed
  const uint32_t accum_bar = bar_base + 8u * (3 * WT_STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dyy = blockIdx.x % 3, ci0 = (blockIdx.x / 3) * NCI, co0 = blockIdx.y * BM, split = blockIdx.z;
  const int64_t u0 = (int64_t)split * p.units_per_split;
  const int64_t u1 = min(p.total_units, u0 + p.units_per_split);

  if (tid == 0) {
    for (int s = 0; s < WT_STAGES; ++s) { mbar_init(fullD(s), 1); mbar_init(fullA(s), 128); mbar_init(empty(s), 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ============ dy loaders, then epilogue ============
    const int cl = warp * 32 + lane;          // channel within the co tile = TMEM lane
    float a_inv;
    operand_scale(p.dy_amax, &a_inv);
    float bsum = 0.f;
    const bool want_bias = p.bpart != nullptr && blockIdx.x == 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t u = u0; u < u1; ++u) {
      mbar_wait(fullD(stage), phase);
      const unsigned short* dh = reinterpret_cast<const unsigned short*>(smem + (size_t)stage * STAGE + XB) + cl;
      float w[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const uint32_t lo = dh[(2 * j) * 128], hi = dh[(2 * j + 1) * 128];
        w[j] = __uint_as_float(lo | (hi << 16));
        if (want_bias) bsum += __half2float(__ushort_as_half((unsigned short)lo)) + __half2float(__ushort_as_half((unsigned short)hi));
      }
      tc_fence_after();
      tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16) + ACC_COLS + (uint32_t)(stage * A_COLS), w);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(fullA(stage));
      if (++stage == WT_STAGES) { stage = 0; phase ^= 1; }
    }
    if (want_bias) p.bpart[(size_t)split * p.Cout + co0 + cl] = bsum * a_inv;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int dx = 0; dx < 3; ++dx) {
      float* o = p.part + (((size_t)split * 9 + dyy * 3 + dx) * p.Cout + co0 + cl) * p.Cin + ci0;
#pragma unroll 1
      for (int cb = 0; cb < NCI / 32; ++cb) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(dx * NCI + cb * 32), v);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(o + cb * 32 + q * 4) =
              make_float4(v[4 * q] * a_inv, v[4 * q + 1] * a_inv, v[4 * q + 2] * a_inv, v[4 * q + 3] * a_inv);
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ============ MMA issuer (warp-uniform loop, one elected lane issues) ============
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t u = u0; u < u1; ++u) {
      mbar_wait(fullA(stage), phase);     // implies fullD: the halo of the stage has landed too
      tc_fence_after();
      const uint32_t xs = smem_base + (uint32_t)stage * STAGE;
      const uint32_t a_t = tmem_base + ACC_COLS + (uint32_t)(stage * A_COLS);
      // MN-major, 128-byte swizzle: K groups (8 pixels = one image row of the unit) are SBO = 1280 B apart
      const uint64_t xd0 = (uint64_t)((xs >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1280 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      const uint32_t acc0 = (u > u0) ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int r = 0; r < 8; r += 2) {          // K = 16 pixels: image rows r, r + 1 of the unit
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
            for (int hf = 0; hf < NCI / 64; ++hf) {
              const uint64_t xd = xd0 + (uint64_t)((hf * WT_XATOM + r * 1280 + dx * 128) >> 4);
              mma_f16_ts(tmem_base + (uint32_t)(dx * NCI + hf * 64), a_t + (uint32_t)(r * 4), xd, idesc, r > 0 ? 1u : acc0);
            }
          }
        }
        mma_commit(empty(stage));
        if (u + 1 == u1) mma_commit(accum_bar);
      }
      __syncwarp();
      if (++stage == WT_STAGES) { stage = 0; phase ^= 1; }
    }
    if (u0 >= u1 && elect_one()) mma_commit(accum_bar);
    __syncwarp();
  } else {
    // ============ copy issuer (one thread): dy tile + activation halo rows of this kernel row ============
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t u = u0; u < u1; ++u) {
        const int ux = (int)(u % p.units_x), uy = (int)((u / p.units_x) % p.units_y);
        const int n = (int)(u / ((int64_t)p.units_x * p.units_y));
        const uint32_t dst = smem_base + (uint32_t)stage * STAGE;
        mbar_wait(empty(stage), phase ^ 1);
        mbar_expect_tx(fullD(stage), STAGE);
#pragma unroll
        for (int hf = 0; hf < NCI / 64; ++hf)
          tma_load_4d(dst + (uint32_t)(hf * WT_XATOM), &x_map, ci0 + hf * 64, ux * 8 - 1, uy * 8 + dyy - 1, n, fullD(stage));
        tma_load_4d(dst + XB, &dy_map, co0, ux * 8, uy * 8, n, fullD(stage));
        if (++stage == WT_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NCI>
constexpr size_t wt_smem_bytes() {
  return 1024 + (size_t)WT_STAGES * ((NCI / 64) * WT_XATOM + WT_DY) + (3 * WT_STAGES + 1) * 8 + 16;
}

// fp32 -> fp16 shadow (optionally scaled by the power-of-two operand scale of *amax): plain vectorised copy
__global__ void to_half_kernel(const float4* __restrict__ x, uint2* __restrict__ y, int64_t n4, const float* __restrict__ amax) {
  float inv;
  const float s = operand_scale(amax, &inv);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    y[i] = make_uint2(pack_h2(v.x * s, v.y * s), pack_h2(v.z * s, v.w * s));
  }
}

}  // namespace tc

static bool dense_nhwc4(const mas_tensor4& t) {
  return t.sc == 1 && t.sw == t.c && t.sh == t.w * t.c && t.sn == t.h * t.w * t.c;
}

bool conv3x3_tma16_ok(mas_tensor4 xs, mas_tensor4 ys) {
  return dense_nhwc4(xs) && dense_nhwc4(ys) && xs.c % 64 == 0 && xs.c >= 64 && ys.c % 4 == 0 && ys.h % 16 == 0 && ys.w % 8 == 0 &&
         xs.h == ys.h && xs.w == ys.w && xs.n == ys.n;
}

// x16: fp16 NHWC shadow of the (activated) input, scaled by operand_scale(*x_amax) when x_amax is given.
int conv3x3_fprop_tma16_launch(const void* x16, mas_tensor4 xs, const void* w_tc16, const float* bias, const float* res, float* y,
                               mas_tensor4 ys, float* stats_part, const float* x_amax, cudaStream_t st) {
  if (!conv3x3_tma16_ok(xs, ys)) return fail(MAS_ERR_UNSUPPORTED, "tma conv: shape/layout not eligible (Cin=%lld Cout=%lld H=%lld W=%lld)",
                                              (long long)xs.c, (long long)ys.c, (long long)ys.h, (long long)ys.w);
  const int Cstore = (int)ys.c, Cout = (int)cdiv(ys.c, tc::BN) * tc::BN;
  if (Cstore != Cout && (res || stats_part)) return fail(MAS_ERR_UNSUPPORTED, "tma conv: residual / statistics epilogues need Cout %% 128 == 0");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(x16) || !al16(y) || !al16(w_tc16) || (res && !al16(res)) || (bias && !al16(bias)))
    return fail(MAS_ERR_INVALID_ARG, "tma conv: pointers must be 16-byte aligned");
  tc::HParams p;
  p.wpk = w_tc16; p.bias = bias; p.res = res; p.y = y;
  p.N = (int)xs.n; p.H = (int)xs.h; p.W = (int)xs.w; p.Cin = (int)xs.c; p.Cout = Cout; p.Cstore = Cstore; p.ldy = Cstore;
  p.tiles_x = (int)(ys.w / 8); p.tiles_y = (int)(ys.h / 16);
  p.stats_part = stats_part; p.x_amax = x_amax;
  const bool tall = ys.h % 32 == 0;
  const int64_t tiles = (int64_t)p.N * p.tiles_x * p.tiles_y;
  p.units = tall ? tiles / 2 : cdiv(tiles, 2);

  PFN_cuTensorMapEncodeTiled enc = tensor_map_encoder();
  if (!enc) return fail(MAS_ERR_LAUNCH, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap map;
  cuuint64_t dims[4] = {(cuuint64_t)xs.c, (cuuint64_t)xs.w, (cuuint64_t)xs.h, (cuuint64_t)xs.n};
  cuuint64_t strides[3] = {(cuuint64_t)xs.c * 2, (cuuint64_t)xs.w * xs.c * 2, (cuuint64_t)xs.h * xs.w * xs.c * 2};
  cuuint32_t box[4] = {64, 10, tall ? 34u : 18u, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x16), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MAS_ERR_LAUNCH, "cuTensorMapEncodeTiled (conv halo map) failed (%d)", (int)r);

  constexpr size_t smem = tc::t16_smem_bytes();
  static std::atomic<uint64_t> configured{0};
  static int sm_count = 148;
  if (first_on_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(tc::shift_gemm_t16<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::shift_gemm_t16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(MAS_ERR_LAUNCH, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    mark_device(configured);
  }
  const int64_t nitems = p.units * (Cout / tc::BN);
  const unsigned g = (unsigned)(nitems < sm_count ? nitems : sm_count);
  if (tall) tc::shift_gemm_t16<true><<<g, tc::T_THREADS, smem, st>>>(p, map);
  else tc::shift_gemm_t16<false><<<g, tc::T_THREADS, smem, st>>>(p, map);
  return launched_tc(tall ? "shift_gemm_t16<tall>" : "shift_gemm_t16<pair>");
}

void conv_wgrad_reduce_launch(const float* part, int splits, int ntap, int Cout, int Cin, float* dw, const float* bpart, float* dbias,
                              cudaStream_t st);   // contract_simt.cu

static int wt_splits(int64_t cps, int64_t units) {
  int64_t s = 148 / cps;
  if (s < 1) s = 1;
  if (s > units) s = units;
  const int64_t ups = cdiv(units, s);
  return (int)cdiv(units, ups);
}
bool conv_wgrad_t16_ok(mas_tensor4 xs, mas_tensor4 dys) {
  return dense_nhwc4(xs) && dense_nhwc4(dys) && xs.c % 64 == 0 && dys.c % 8 == 0 && dys.h % 8 == 0 && dys.w % 8 == 0 && xs.h == dys.h &&
         xs.w == dys.w && xs.n == dys.n;
}
size_t conv_wgrad_t16_ws(mas_tensor4 xs, mas_tensor4 dys) {
  if (!conv_wgrad_t16_ok(xs, dys)) return 0;
  const int64_t coutk = cdiv(dys.c, tc::BM) * tc::BM;
  const int nci = xs.c % 128 == 0 ? 128 : 64;
  const size_t splits = wt_splits((coutk / tc::BM) * (xs.c / nci) * 3, dys.n * (dys.h / 8) * (dys.w / 8));
  return splits * 9 * (size_t)coutk * xs.c * sizeof(float) + splits * (size_t)coutk * sizeof(float) + 256;
}
// x16: fp16 activation shadow; dy16: fp16 output-gradient shadow scaled by operand_scale(*dy_amax); dw/dbias sized for
// round_up(dys.c, 128) output channels (the copy engine zero-fills the channels dy does not have).
int conv_wgrad_t16_launch(const void* x16, mas_tensor4 xs, const void* dy16, mas_tensor4 dys, float* dw, float* dbias,
                          const float* dy_amax, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!conv_wgrad_t16_ok(xs, dys) || !dy_amax) return fail(MAS_ERR_UNSUPPORTED, "tma wgrad: shape/layout not eligible");
  if (ws_bytes < conv_wgrad_t16_ws(xs, dys)) return fail(MAS_ERR_WORKSPACE, "tma wgrad: workspace too small");
  tc::WTParams p;
  p.N = (int)xs.n; p.H = (int)xs.h; p.W = (int)xs.w; p.Cin = (int)xs.c; p.Cout = (int)(cdiv(dys.c, tc::BM) * tc::BM);
  p.units_x = (int)(dys.w / 8); p.units_y = (int)(dys.h / 8);
  p.total_units = (int64_t)p.N * p.units_x * p.units_y;
  p.dy_amax = dy_amax;
  const int nci = p.Cin % 128 == 0 ? 128 : 64;
  const int splits = wt_splits((int64_t)(p.Cout / tc::BM) * (p.Cin / nci) * 3, p.total_units);
  p.units_per_split = cdiv(p.total_units, splits);
  p.part = (float*)ws;
  p.bpart = dbias ? (float*)ws + (size_t)splits * 9 * p.Cout * p.Cin : nullptr;

  PFN_cuTensorMapEncodeTiled enc = tensor_map_encoder();
  if (!enc) return fail(MAS_ERR_LAUNCH, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap xmap, dmap;
  {
    cuuint64_t dims[4] = {(cuuint64_t)xs.c, (cuuint64_t)xs.w, (cuuint64_t)xs.h, (cuuint64_t)xs.n};
    cuuint64_t strides[3] = {(cuuint64_t)xs.c * 2, (cuuint64_t)xs.w * xs.c * 2, (cuuint64_t)xs.h * xs.w * xs.c * 2};
    cuuint32_t box[4] = {64, 10, 8, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x16), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MAS_ERR_LAUNCH, "cuTensorMapEncodeTiled (wgrad halo map) failed (%d)", (int)r);
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)dys.c, (cuuint64_t)dys.w, (cuuint64_t)dys.h, (cuuint64_t)dys.n};
    cuuint64_t strides[3] = {(cuuint64_t)dys.c * 2, (cuuint64_t)dys.w * dys.c * 2, (cuuint64_t)dys.h * dys.w * dys.c * 2};
    cuuint32_t box[4] = {128, 8, 8, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&dmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(dy16), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MAS_ERR_LAUNCH, "cuTensorMapEncodeTiled (wgrad dy map) failed (%d)", (int)r);
  }
  static std::atomic<uint64_t> configured{0};
  if (first_on_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(tc::wgrad_t16<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::wt_smem_bytes<128>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::wgrad_t16<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::wt_smem_bytes<64>());
    if (e != cudaSuccess) return fail(MAS_ERR_LAUNCH, "cudaFuncSetAttribute(wgrad_t16): %s", cudaGetErrorString(e));
    mark_device(configured);
  }
  dim3 grid((unsigned)((p.Cin / nci) * 3), (unsigned)(p.Cout / tc::BM), (unsigned)splits);
  if (nci == 128) tc::wgrad_t16<128><<<grid, tc::WT_THREADS, tc::wt_smem_bytes<128>(), st>>>(p, xmap, dmap);
  else tc::wgrad_t16<64><<<grid, tc::WT_THREADS, tc::wt_smem_bytes<64>(), st>>>(p, xmap, dmap);
  if (int e = launched_tc("wgrad_t16")) return e;
  conv_wgrad_reduce_launch((const float*)ws, splits, 9, p.Cout, p.Cin, dw, p.bpart, dbias, st);
  return launched("conv_wgrad_reduce");
}

int to_half_launch(const float* x, void* y, int64_t n, const float* amax, cudaStream_t st) {
  if (n % 4 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 7))
    return fail(MAS_ERR_INVALID_ARG, "to_half: n %% 4 == 0 and aligned pointers required");
  const int64_t n4 = n / 4;
  const int64_t blocks = cdiv(n4, 256);
  tc::to_half_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(reinterpret_cast<const float4*>(x),
                                                                                       reinterpret_cast<uint2*>(y), n4, amax);
  return launched("to_half");
}

}  // namespace mas

extern "C" {

int mas_conv3x3_tc16h_eligible(mas_tensor4 xs, mas_tensor4 ys) { return mas::conv3x3_tma16_ok(xs, ys) ? 1 : 0; }

int mas_conv3x3_fprop_tc16h(const void* x_f16, mas_tensor4 xs, const void* w_tc16, const float* bias, const float* residual, float* y,
                            mas_tensor4 ys, float* stats_part, const float* x_amax, void* stream) {
  MAS_REQUIRE(x_f16 && w_tc16 && y, "conv3x3_fprop_tc16h: null pointer");
  return mas::conv3x3_fprop_tma16_launch(x_f16, xs, w_tc16, bias, residual, y, ys, stats_part, x_amax, mas::S(stream));
}

int mas_to_half(const float* x, void* y_f16, int64_t n, const float* amax, void* stream) {
  MAS_REQUIRE(x && y_f16 && n > 0, "to_half: bad arguments");
  return mas::to_half_launch(x, y_f16, n, amax, mas::S(stream));
}

}  // extern "C"
