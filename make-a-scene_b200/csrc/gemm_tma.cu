// Row GEMM  y[M, N] = alpha * x[M, K] . W[N, K]^T (+ bias, + residual)  as a pure TMA + tcgen05 kernel on fp16 operands: the
// Linear layers of the token transformer (transformer.py:17-56, forward and data gradient) and, in general, any 1x1
// convolution whose input exists as a dense fp16 [M, K] matrix.  It is shift_gemm_t16 (conv_tma.cu) with one tap:
//   * A side of the problem (activations): ONE 2-D tensor-map copy per 64-channel chunk, box 64 halves x 256 rows under the
//     128-byte swizzle = the K-major SW128 operand image (rows at a 128-byte pitch, 8-row groups 1024 bytes apart); rows past
//     M are zero-filled by the copy engine;
//   * weights: mas_pack_gemm_tc16 image [n_tile][k/16][k/8 % 2][128][8 halves] (no swizzle): the four K = 16 steps of a chunk are
//     16 KB contiguous -> one cp.async.bulk per chunk, into the same ring stage as the activation tile (one barrier pair);
//   * operand roles swapped (D^T = W x X^T: weights on the M side, 256 rows on the N side of ONE M128 x N256 x K16 MMA): a TMEM
//     lane is an output feature, a column a row of x - the epilogue's 32 lanes store 32 consecutive floats of one output row
//     (a full 128-byte line per instruction, no shared-memory transpose), bias is a per-thread scalar;
//   * one persistent CTA per SM over (256-row tile x 128-feature tile) items, feature tile fastest (the row tile is re-read
//     from L2), 4-stage ring of 48 KB, two 256-column accumulator sets: 8 epilogue warps drain one while the MMAs fill the other;
//   * warps 0-7 epilogue, warp 8 MMA issuer (warp-uniform loop, elected lane), warp 9 copy issuer: no thread touches an operand.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

#include "mas_common.cuh"
#include "tc_ptx.cuh"

namespace mas {

PFN_cuTensorMapEncodeTiled tensor_map_encoder();   // contract_tc.cu

namespace tc {

constexpr int G_EPI_WARPS = 8;
constexpr int G_THREADS = (G_EPI_WARPS + 2) * 32;
constexpr int G_STAGES = 4;
constexpr int G_ROWS = 256;                        // rows of x per work item (N of the MMA)
constexpr int G_A = G_ROWS * 128;                  // 256 rows x 64 halves
constexpr int G_BSUB = 2 * BN * 16;                // one K = 16 step of the packed weights: [2][128][8 halves] = 4 KB
constexpr int G_B = 4 * G_BSUB;                    // the chunk's four steps
constexpr int G_STAGE = G_A + G_B;                 // 48 KB

struct GParams {
  const void* wpk;     // mas_pack_gemm_tc16 image
  const float* bias;   // [N] or null
  const float* res;    // [M, ldy] like y, or null
  float* y;
  int64_t M, ldy;
  int K, N, Nstore;    // N: padded to 128 (rows of the packed image), Nstore: columns actually stored
  int64_t units;       // 256-row tiles
  const float* x_amax; // amax the fp16 copy of x was scaled from (null: unscaled)
  float alpha;
};

__device__ __forceinline__ uint64_t make_desc_sw128g(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(G_THREADS, 1) rows_gemm_t16(const GParams p, const __grid_constant__ CUtensorMap x_map) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_base = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_base + 1023u) & ~1023u;           // swizzle atoms are 1024-byte aligned
  uint8_t* smem = smem_raw + (smem_base - raw_base);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)G_STAGES * G_STAGE);
  constexpr int NBARS = 2 * G_STAGES + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
  const uint32_t bar_base = smem_u32(bars);
  auto full = [&](int s) { return bar_base + 8u * s; };
  auto empty = [&](int s) { return bar_base + 8u * (G_STAGES + s); };
  auto accf = [&](int b) { return bar_base + 8u * (2 * G_STAGES + b); };
  auto acce = [&](int b) { return bar_base + 8u * (2 * G_STAGES + 2 + b); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunks = p.K / 64;
  const int n_tiles = p.N / BN;
  const int64_t nitems = p.units * n_tiles;

  if (tid == 0) {
    for (int s = 0; s < G_STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(accf(b), 1); mbar_init(acce(b), G_EPI_WARPS * 32); }
    fence_barrier_init();
  }
  if (warp == G_EPI_WARPS) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < G_EPI_WARPS) {
    // ===================== epilogue warps: lane = output feature, column = row of x =====================
    float inv_scale = 1.f;
    operand_scale(p.x_amax, &inv_scale);
    const float alpha = inv_scale * p.alpha;
    const int quarter = warp & 3, hf = warp >> 2;
    int buf = 0;
    uint32_t ph[2] = {0u, 0u};
    for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int64_t unit = item / n_tiles;
      const int ch = (int)(item % n_tiles) * BN + quarter * 32 + lane;
      const bool st_ok = ch < p.Nstore;
      const float bv = (p.bias && st_ok) ? __ldg(p.bias + ch) : 0.f;
      const int64_t row0 = unit * G_ROWS + hf * 128;
      mbar_wait(accf(buf), ph[buf]);
      ph[buf] ^= 1u;
      tc_fence_after();
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 256 + hf * 128 + cb * 32), v);
        if (cb == 3) {
          tc_fence_before();
          mbar_arrive(acce(buf));     // this warp's share of the accumulator set is in registers
        }
        const int64_t r0 = row0 + cb * 32;
        if (st_ok) {
          float* yp = p.y + r0 * p.ldy + ch;
          const float* rp = p.res ? p.res + r0 * p.ldy + ch : nullptr;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (r0 + j < p.M) {
              float o = fmaf(v[j], alpha, bv);
              if (rp) o += __ldg(rp + (int64_t)j * p.ldy);
              yp[(int64_t)j * p.ldy] = o;
            }
          }
        }
      }
      buf ^= 1;
    }
  } else if (warp == G_EPI_WARPS) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_f16(G_ROWS);
    int s = 0, buf = 0;
    uint32_t sph = 0, eph[2] = {0u, 0u};
    for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
      mbar_wait(acce(buf), eph[buf] ^ 1);     // the epilogue warps have read this accumulator set (first use: passes)
      eph[buf] ^= 1u;
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * 256);
      for (int c = 0; c < chunks; ++c) {
        mbar_wait(full(s), sph);
        tc_fence_after();
        const uint32_t st = smem_base + (uint32_t)s * G_STAGE;
        const uint64_t xd0 = make_desc_sw128g(st, 1024);
        const uint64_t wd0 = make_desc(st + G_A, BN * 16, 128);
        if (elect_one()) {
#pragma unroll
          for (int sub = 0; sub < 4; ++sub) {
            // D^T = W x X^T: weights on the M side, the 256 rows on the N side
            mma_f16_ss(acc, wd0 + (uint64_t)((sub * G_BSUB) >> 4), xd0 + (uint64_t)((sub * 32) >> 4), idesc, (c > 0 || sub > 0) ? 1u : 0u);
          }
          mma_commit(empty(s));
          if (c == chunks - 1) mma_commit(accf(buf));
        }
        __syncwarp();
        if (++s == G_STAGES) { s = 0; sph ^= 1; }
      }
      buf ^= 1;
    }
  } else {
    // ===================== copy issuer (one thread) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t sph = 0;
      const int ksteps = p.K / 16;
      for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int64_t unit = item / n_tiles;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpk) + (size_t)(item % n_tiles) * ksteps * G_BSUB;
        for (int c = 0; c < chunks; ++c) {
          mbar_wait(empty(s), sph ^ 1);
          const uint32_t st = smem_base + (uint32_t)s * G_STAGE;
          mbar_expect_tx(full(s), G_STAGE);
          tma_load_2d(st, &x_map, c * 64, (int)(unit * G_ROWS), full(s));
          bulk_g2s(st + G_A, wsrc + (size_t)c * G_B, G_B, full(s));
          if (++s == G_STAGES) { s = 0; sph ^= 1; }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == G_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

constexpr size_t g16_smem_bytes() { return 1024 + (size_t)G_STAGES * G_STAGE + (2 * G_STAGES + 4) * 8 + 16; }

}  // namespace tc

bool gemm_rows_f16_ok(int64_t M, int N, int K) { return M > 0 && N > 0 && K >= 64 && K % 64 == 0 && M < (1ll << 31) - 256; }

int gemm_rows_f16_launch(const void* x16, int64_t M, int K, const void* w_tc16, float* y, int64_t ldy, int N, const float* bias,
                         const float* res, const float* x_amax, float alpha, cudaStream_t st) {
  if (!gemm_rows_f16_ok(M, N, K)) return fail(MAS_ERR_UNSUPPORTED, "gemm_rows_f16: needs K %% 64 == 0 (M=%lld N=%d K=%d)", (long long)M, N, K);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(x16) || !al16(w_tc16)) return fail(MAS_ERR_INVALID_ARG, "gemm_rows_f16: x16 / packed weights must be 16-byte aligned");
  tc::GParams p;
  p.wpk = w_tc16; p.bias = bias; p.res = res; p.y = y; p.M = M; p.ldy = ldy; p.K = K;
  p.N = (int)cdiv(N, tc::BN) * tc::BN; p.Nstore = N;
  p.units = cdiv(M, tc::G_ROWS); p.x_amax = x_amax; p.alpha = alpha;

  PFN_cuTensorMapEncodeTiled enc = tensor_map_encoder();
  if (!enc) return fail(MAS_ERR_LAUNCH, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)tc::G_ROWS}, es[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x16), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MAS_ERR_LAUNCH, "cuTensorMapEncodeTiled (row GEMM map) failed (%d)", (int)r);

  constexpr size_t smem = tc::g16_smem_bytes();
  static std::atomic<uint64_t> configured{0};
  static int sm_count = 148;
  if (first_on_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(tc::rows_gemm_t16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(MAS_ERR_LAUNCH, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    mark_device(configured);
  }
  const int64_t nitems = p.units * (p.N / tc::BN);
  const unsigned g = (unsigned)(nitems < sm_count ? nitems : sm_count);
  tc::rows_gemm_t16<<<g, tc::G_THREADS, smem, st>>>(p, map);
  return launched_tc("rows_gemm_t16");
}

}  // namespace mas

extern "C" {

int mas_gemm_rows_f16(const void* x_f16, int64_t M, int K, const void* w_tc16, float* y, int64_t ldy, int N, const float* bias,
                      const float* residual, const float* x_amax, float alpha, void* stream) {
  MAS_REQUIRE(x_f16 && w_tc16 && y && ldy >= N, "gemm_rows_f16: bad arguments");
  return mas::gemm_rows_f16_launch(x_f16, M, K, w_tc16, y, ldy, N, bias, residual, x_amax, alpha, mas::S(stream));
}

}  // extern "C"
