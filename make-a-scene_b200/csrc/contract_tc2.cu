// STAGED (next round; not selected unless MAS_CONV_2CTA=1): the 3x3 convolution forward / data-gradient kernel of
// contract_tc.cu (`shift_gemm_tc<9,8,2,2>`) with tcgen05 `cta_group::2` CTA pairs.
//
// Why: the SS-mode 128x128x8 TF32 MMA fetches 8 KB of shared-memory operands per ~65-cycle instruction - the operand port is
// ~100 % busy at full tensor rate, and ncu shows 67 % tensor-core + 20 % LSU wavefronts on it at 62 % tensor-pipe active
// (profiles/r01_ncu_conv_b32.md).  With a CTA pair one MMA covers M = 256 pixels (128 from each CTA's own halo tile) x N = 128
// channels and each CTA supplies only HALF of the weight tile (64 rows): 6 KB per CTA per instruction.
//
// Differences from the single-CTA kernel (everything else - shift-GEMM halo staging, GroupNorm+SiLU prologue, statistics
// epilogue, store patch - is identical):
//   * cluster (2,1,1); both CTAs allocate tensor memory with .cta_group::2; only rank 0 issues tcgen05.mma.cta_group::2
//   * full barriers live in the leader: the producers of BOTH CTAs arrive there (mapa + mbarrier.arrive.shared::cluster), and
//     each CTA's weight thread waits for its own 18 x 1 KB bulk copies (local barrier) and then relays one arrive
//   * tcgen05.commit.cta_group::2 ... multicast::cluster (mask 0b11) frees the stage / signals the accumulator in both CTAs
//   * cluster barriers bracket the kernel (remote arrives need initialised barriers; nobody may exit while the pair's MMAs
//     still read its shared memory)
#include "mas_common.cuh"

namespace mas {
namespace tc2 {

constexpr int BM = 128;        // pixels per M tile (16 x 8)
constexpr int BN = 128;        // output channels per CTA
constexpr int NPROD = 256;     // producer threads (warps 0-7)
constexpr int NTHREADS = 320;  // + MMA warp + bulk-copy warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {   // both CTAs of the pair execute this
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // waits for arrivals from both CTAs
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 in, fp32 accumulate, issued by ONE thread
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (when the pair's MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, no swizzle ("interleaved"), sm_100 version field = 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (1ull << 46);
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, M=256 (the CTA pair), N
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
}

enum { MAP_S1 = 0, MAP_UP = 2, MAP_ZS = 3, MAP_ROWS = 4 };

struct Params {
  const float* x;    // NHWC input (dense) or row matrix
  const float* wpk;  // packed weights [n_tile][k_chunk][tap][k/4][BN][4]
  const float* bias; // [Cout] or null
  const float* res;  // same layout as y, or null
  float* y;
  int N, Hin, Win, Cin, Hout, Wout, Cout;  // for MAP_ROWS: Hout*Wout*N = rows, Win unused
  int map;
  int64_t ldx, ldy;  // row pitches (elements) of x pixels and y pixels
  int tiles_x, tiles_y;  // tiles per image row / column (image maps)
  int64_t total_tiles;
  float alpha;
  // fused GroupNorm(+SiLU) PROLOGUE on the A operand: a = act(x * sc + sh) with (sc, sh) per (image, input channel) in
  // gn_table [N][Cin][2] (null = plain input).  Padding pixels stay exactly zero (the reference pads the ACTIVATED tensor).
  const float* gn_table;
  int gn_silu;
  // fused GroupNorm-statistics EPILOGUE for the NEXT layer's norm: per (tile, 32-row lane group, channel quad) sum and
  // sum of squares of the stored output, [total_tiles][4][Cout/4][2] floats (null = off); reduced deterministically
  // per (image, group) by mas_gn_finalize_partials.
  float* stats_part;
};

// One CTA = TILES M-tiles x BN output channels, full K; two CTAs of a cluster share every MMA (M = 256).
constexpr int TAPS = 9, KC = 8, STAGES = 2, TILES = 2;
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 2) shift_gemm_tc2(const Params p) {
  constexpr int SLOTS = (TAPS == 9) ? 180 : 132;        // staged pixels per tile (18x10 halo | 128 rows + pad)
  constexpr int ROWP = (TAPS == 9) ? 10 : 8;            // staged pixels per image row
  constexpr int LBO_A = SLOTS * 16;                     // bytes between k-quads of A
  constexpr int SBO_A = ROWP * 16;                      // bytes between 8-pixel groups of A
  constexpr int A_TILE = (KC / 4) * LBO_A;              // bytes per tile per stage
  constexpr int A_STAGE = TILES * A_TILE;
  constexpr int LBO_B = (BN / 2) * 16;                  // this CTA holds HALF of the weight rows (N/2 = 64) of every k-quad
  constexpr int B_TAP = (KC / 4) * LBO_B;
  constexpr int B_STAGE = TAPS * B_TAP;
  constexpr int STAGE = A_STAGE + B_STAGE;
  constexpr int QUADS = KC / 4;
  constexpr int ITEMS = TILES * SLOTS * QUADS;          // float4 items staged per K chunk
  constexpr int PER_THREAD = (ITEMS + NPROD - 1) / NPROD;
  static_assert((SLOTS % 8) == 4, "A plane pitch must be 64 mod 128 bytes for conflict-free producer stores");

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE);
  // bars[0..S) full (the LEADER's are the ones the MMA thread waits on: both CTAs arrive there), bars[S..2S) empty,
  // bars[2S] accumulator ready, bars[2S+1..3S+1) local completion of this CTA's weight bulk copies; then the TMEM base word
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * STAGES);
  auto wfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 1 + s); };
  const uint32_t rank = cluster_ctarank();   // 0 = leader (issues the MMAs of the pair)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tile0 = (int64_t)blockIdx.x * TILES;
  const int n0 = blockIdx.y * BN;
  const int nchunks = p.Cin / KC;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 2 * (NPROD + 1));   // producers + weight relay of BOTH CTAs (only the leader's is waited on)
      mbar_init(empty_bar(s), 1);
      mbar_init(wfull_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), TILES * BN);
  tc_fence_before();
  cluster_sync_all();   // both CTAs' barriers are initialised before anyone arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ===================== producers: stage A (input pixels) =====================
    const float* src[PER_THREAD];
    const float* tab[(TAPS == 9) ? PER_THREAD : 1];   // prologue table pointers (3x3 convolutions only)
    uint32_t dst[PER_THREAD];
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      const int item = tid + i * NPROD;
      src[i] = nullptr;
      if (TAPS == 9) tab[i] = nullptr;
      dst[i] = 0xFFFFFFFFu;
      if (item < ITEMS) {
        const int q = item % QUADS, rest = item / QUADS, slot = rest % SLOTS, tl = rest / SLOTS;
        dst[i] = (uint32_t)(tl * A_TILE + q * LBO_A + slot * 16);
        const int64_t tile = tile0 + tl;
        if (tile < p.total_tiles) {
          if (TAPS == 9) {
            const int tx_ = (int)(tile % p.tiles_x), ty_ = (int)((tile / p.tiles_x) % p.tiles_y);
            const int n = (int)(tile / ((int64_t)p.tiles_x * p.tiles_y));
            const int r = slot / 10, c = slot % 10;
            const int vy = ty_ * 16 - 1 + r, vx = tx_ * 8 - 1 + c;  // coordinates in the (virtual) conv input image
            int iy = vy, ix = vx;
            bool ok;
            if (p.map == MAP_S1) {
              ok = (unsigned)vy < (unsigned)p.Hin && (unsigned)vx < (unsigned)p.Win;
            } else if (p.map == MAP_UP) {
              ok = (unsigned)vy < (unsigned)(2 * p.Hin) && (unsigned)vx < (unsigned)(2 * p.Win);
              iy = vy >> 1; ix = vx >> 1;
            } else {  // MAP_ZS
              ok = vy >= 0 && vx >= 0 && (vy & 1) && (vx & 1) && (vy >> 1) < p.Hin && (vx >> 1) < p.Win;
              iy = vy >> 1; ix = vx >> 1;
            }
            if (ok) {
              src[i] = p.x + ((int64_t)(n * p.Hin + iy) * p.Win + ix) * p.ldx + q * 4;
              if (p.gn_table) tab[i] = p.gn_table + ((size_t)n * p.Cin + q * 4) * 2;
            }
          } else {
            const int64_t row = tile * BM + slot;
            if (slot >= BM) dst[i] = 0xFFFFFFFFu;  // pad slots are never read by the MMA
            else if (row < (int64_t)p.N * p.Hout * p.Wout) src[i] = p.x + row * p.ldx + q * 4;
          }
        }
      }
    }
    int stage = 0;
    uint32_t phase = 0;
    float4 vn[PER_THREAD];
    auto gload = [&](int kc, float4* v) {
#pragma unroll
      for (int i = 0; i < PER_THREAD; ++i) {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src[i]) v[i] = __ldg(reinterpret_cast<const float4*>(src[i] + (size_t)kc * KC));
      }
    };
    gload(0, vn);
    for (int kc = 0; kc < nchunks; ++kc) {
      float4 v[PER_THREAD];
#pragma unroll
      for (int i = 0; i < PER_THREAD; ++i) v[i] = vn[i];
      if (kc + 1 < nchunks) gload(kc + 1, vn);  // next chunk's global loads fly while this chunk is stored / consumed
      if (TAPS == 9) {
        if (p.gn_table) {
          // fused GroupNorm (+SiLU) prologue, applied at CONSUME time so the prefetch above stays asynchronous;
          // the (scale, shift) pairs are L1-resident
#pragma unroll
          for (int i = 0; i < PER_THREAD; ++i) {
            if (tab[i]) {
              const float4 t0 = __ldg(reinterpret_cast<const float4*>(tab[i] + (size_t)kc * KC * 2));      // sc0 sh0 sc1 sh1
              const float4 t1 = __ldg(reinterpret_cast<const float4*>(tab[i] + (size_t)kc * KC * 2) + 1);  // sc2 sh2 sc3 sh3
              float a0 = fmaf(v[i].x, t0.x, t0.y), a1 = fmaf(v[i].y, t0.z, t0.w);
              float a2 = fmaf(v[i].z, t1.x, t1.y), a3 = fmaf(v[i].w, t1.z, t1.w);
              if (p.gn_silu) { a0 = silu_f(a0); a1 = silu_f(a1); a2 = silu_f(a2); a3 = silu_f(a3); }
              v[i] = make_float4(a0, a1, a2, a3);
            }
          }
        }
      }
      mbar_wait(empty_bar(stage), phase ^ 1);
      uint8_t* a_st = smem + (size_t)stage * STAGE;
#pragma unroll
      for (int i = 0; i < PER_THREAD; ++i)
        if (dst[i] != 0xFFFFFFFFu) *reinterpret_cast<float4*>(a_st + dst[i]) = v[i];
      fence_proxy_async();  // make the generic-proxy stores visible to the tensor core (async proxy)
      mbar_arrive_remote(full_bar(stage), 0);   // the leader's barrier collects both CTAs' producers
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }

    // ===================== epilogue: TMEM -> registers -> smem transpose -> coalesced global stores =====================
    // A thread owns one pixel row of the accumulator (32 consecutive channels per tcgen05.ld); writing that directly
    // makes every store instruction touch 32 different 128-byte lines with 16 bytes each.  Each warp instead bounces
    // its 32x32 block through a private shared-memory patch (the pipeline stages are idle by now) so that 8 lanes
    // cover one full line: 4 lines per store instruction, and the residual is read the same way.
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int lane_grp = warp & 3;           // TMEM lanes [32*lane_grp, +32)
    const int chalf = warp >> 2;             // column half of the 128-wide tile
    constexpr int EP_LD = 36;                // floats per staged row (144 B: conflict-free 16-byte accesses)
    float* patch = reinterpret_cast<float*>(smem) + warp * (32 * EP_LD);
    const int sub_r = lane >> 3, sub_c = lane & 7;
#pragma unroll 1
    for (int tl = 0; tl < TILES; ++tl) {
      const int64_t tile = tile0 + tl;
      if (tile >= p.total_tiles) break;     // warp-uniform
      int64_t pix_base = 0;                 // pixel index of accumulator row 0 of this tile (image maps: per-row formula)
      int tx_ = 0, ty_ = 0, n_img = 0;
      if (TAPS == 9) {
        tx_ = (int)(tile % p.tiles_x); ty_ = (int)((tile / p.tiles_x) % p.tiles_y);
        n_img = (int)(tile / ((int64_t)p.tiles_x * p.tiles_y));
      } else {
        pix_base = tile * BM;
      }
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int col = chalf * 64 + cc * 32;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(tl * BN + col), v);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(patch + lane * EP_LD + j) =
              make_float4(v[j] * p.alpha, v[j + 1] * p.alpha, v[j + 2] * p.alpha, v[j + 3] * p.alpha);
        __syncwarp();
        float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bq = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + col + sub_c * 4));
        float st_s = 0.f, st_q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = i * 4 + sub_r;               // accumulator row within this warp's 32
          const int m = lane_grp * 32 + row;
          int64_t pix;
          if (TAPS == 9) pix = ((int64_t)n_img * p.Hout + ty_ * 16 + (m >> 3)) * p.Wout + tx_ * 8 + (m & 7);
          else pix = pix_base + m;
          if (TAPS == 9 || pix < (int64_t)p.N * p.Hout * p.Wout) {
            float4 o = *reinterpret_cast<const float4*>(patch + row * EP_LD + sub_c * 4);
            o.x += bq.x; o.y += bq.y; o.z += bq.z; o.w += bq.w;
            const int64_t off = pix * p.ldy + n0 + col + sub_c * 4;
            if (p.res) {
              const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.res + off));
              o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
            }
            *reinterpret_cast<float4*>(p.y + off) = o;
            st_s += (o.x + o.y) + (o.z + o.w);
            st_q = fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, fmaf(o.w, o.w, st_q))));
          }
        }
        if (p.stats_part) {  // fixed-order combine of the four row sub-groups, one (sum, sumsq) per channel quad
          st_s += __shfl_xor_sync(0xffffffffu, st_s, 8);
          st_q += __shfl_xor_sync(0xffffffffu, st_q, 8);
          st_s += __shfl_xor_sync(0xffffffffu, st_s, 16);
          st_q += __shfl_xor_sync(0xffffffffu, st_q, 16);
          if (lane < 8) {
            float* sp = p.stats_part + (((size_t)tile * 4 + lane_grp) * (p.Cout >> 2) + ((n0 + col) >> 2) + sub_c) * 2;
            sp[0] = st_s;
            sp[1] = st_q;
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp == 8) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < nchunks; ++kc) {
        mbar_wait_cluster(full_bar(stage), phase);
        tc_fence_after();
        // one base descriptor per operand per stage; every MMA of the stage is (base + compile-time constant): the
        // start-address field is the low 14 bits (address >> 4) and never carries out for < 256 KB of shared memory, so
        // the single issuing thread spends one add per operand per MMA instead of re-encoding descriptors.
        const uint32_t a_st = smem_base + (uint32_t)stage * STAGE;
        const uint64_t a_base = make_desc(a_st, LBO_A, SBO_A);
        const uint64_t b_base = make_desc(a_st + A_STAGE, LBO_B, 128);   // each CTA's half: 64 rows per k-quad plane
        const uint32_t acc0 = (kc > 0) ? 1u : 0u;
#pragma unroll
        for (int tl = 0; tl < TILES; ++tl) {
#pragma unroll
          for (int t = 0; t < TAPS; ++t) {
            const uint32_t tapoff = (TAPS == 9) ? (uint32_t)(((t / 3) * 10 + (t % 3)) * 16) : 0u;
#pragma unroll
            for (int k8 = 0; k8 < KC / 8; ++k8) {
              const uint64_t ad = a_base + (uint64_t)((tl * A_TILE + tapoff + k8 * 2 * LBO_A) >> 4);
              const uint64_t bd = b_base + (uint64_t)((t * B_TAP + k8 * 2 * LBO_B) >> 4);
              mma_tf32_ss(tmem_base + (uint32_t)(tl * BN), ad, bd, idesc, (t > 0 || k8 > 0) ? 1u : acc0);
            }
          }
        }
        mma_commit(empty_bar(stage));  // frees the smem stage when these MMAs have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      mma_commit(accum_bar);  // all accumulators complete
    }
    __syncwarp();
  } else {
    // ===================== weight bulk-copy issuer (one thread) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // packed weights: [n_tile][k_chunk][tap][k/4][128][4]; this CTA takes rows 64*rank .. 64*rank+63 of every (tap, k-quad)
      // plane: 18 pieces of 1 KB per stage
      constexpr int FULL_STAGE_FLOATS = TAPS * (KC / 4) * BN * 4;
      const float* wsrc = p.wpk + (size_t)blockIdx.y * nchunks * FULL_STAGE_FLOATS + rank * (BN / 2) * 4;
      for (int kc = 0; kc < nchunks; ++kc) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_expect_tx(wfull_bar(stage), B_STAGE);
#pragma unroll
        for (int pc = 0; pc < TAPS * (KC / 4); ++pc)
          bulk_g2s(smem_base + (uint32_t)stage * STAGE + A_STAGE + pc * LBO_B, wsrc + (size_t)kc * FULL_STAGE_FLOATS + (size_t)pc * BN * 4,
                   LBO_B, wfull_bar(stage));
        mbar_wait(wfull_bar(stage), phase);        // landed in THIS CTA's shared memory ...
        mbar_arrive_remote(full_bar(stage), 0);    // ... tell the leader's MMA thread
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  cluster_sync_all();   // the pair's MMAs read both CTAs' shared memory and write both tensor memories: leave together
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TILES * BN);
  }
}


constexpr size_t SMEM_BYTES = (size_t)STAGES * (TILES * (KC / 4) * 180 * 16 + TAPS * (KC / 4) * (BN / 2) * 16) + (3 * STAGES + 1) * 8 + 16;

}  // namespace tc2

static bool dense_nhwc2(const mas_tensor4& t) { return t.sc == 1 && t.sw == t.c && t.sh == t.w * t.c && t.sn == t.h * t.w * t.c; }
static inline bool al16r(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Same contract as conv3x3_fprop_tc_launch (contract_tc.cu); w_tc is the SAME packed image (mas_pack_conv3x3_tc).
int conv3x3_fprop_tc2_launch(const float* x, mas_tensor4 xs, const float* w_tc, const float* bias, const float* res, float* y,
                             mas_tensor4 ys, int mode, const float* gn_table, int gn_silu, float* stats_part, cudaStream_t st) {
  const int Cin = (int)xs.c, Cout = (int)ys.c;
  if (!(mode == MAS_CONV_S1 || mode == MAS_CONV_UP || mode == MAS_CONV_ZS)) return fail(MAS_ERR_UNSUPPORTED, "tc2 conv: mode %d", mode);
  if (!dense_nhwc2(xs) || !dense_nhwc2(ys) || Cin % 8 || Cout % tc2::BN || ys.h % 16 || ys.w % 8 || !al16r(x) || !al16r(y) ||
      (res && !al16r(res)) || (bias && !al16r(bias)) || !al16r(w_tc) || (gn_table && !al16r(gn_table)))
    return fail(MAS_ERR_UNSUPPORTED, "tc2 conv: shape/layout not eligible");
  int64_t eh = (mode == MAS_CONV_S1) ? xs.h : 2 * xs.h, ew = (mode == MAS_CONV_S1) ? xs.w : 2 * xs.w;
  if (ys.h != eh || ys.w != ew || xs.n != ys.n) return fail(MAS_ERR_INVALID_ARG, "tc2 conv: output extent mismatch");
  tc2::Params p;
  p.x = x; p.wpk = w_tc; p.bias = bias; p.res = res; p.y = y;
  p.N = (int)xs.n; p.Hin = (int)xs.h; p.Win = (int)xs.w; p.Cin = Cin; p.Hout = (int)ys.h; p.Wout = (int)ys.w; p.Cout = Cout;
  p.map = (mode == MAS_CONV_S1) ? tc2::MAP_S1 : (mode == MAS_CONV_UP ? tc2::MAP_UP : tc2::MAP_ZS);
  p.ldx = Cin; p.ldy = Cout;
  p.tiles_x = (int)(ys.w / 8); p.tiles_y = (int)(ys.h / 16);
  p.total_tiles = (int64_t)p.N * p.tiles_x * p.tiles_y;
  p.alpha = 1.0f;
  p.gn_table = gn_table; p.gn_silu = gn_silu; p.stats_part = stats_part;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc2::shift_gemm_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2::SMEM_BYTES);
    if (e != cudaSuccess) return fail(MAS_ERR_LAUNCH, "cudaFuncSetAttribute(smem=%zu): %s", tc2::SMEM_BYTES, cudaGetErrorString(e));
    configured = true;
  }
  const int64_t ctas = cdiv(p.total_tiles, tc2::TILES);
  dim3 grid((unsigned)(2 * cdiv(ctas, 2)), (unsigned)(Cout / tc2::BN));   // whole CTA pairs; surplus tiles are masked in the kernel
  tc2::shift_gemm_tc2<<<grid, tc2::NTHREADS, tc2::SMEM_BYTES, st>>>(p);
  return launched("shift_gemm_tc2");
}

}  // namespace mas
