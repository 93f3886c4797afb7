// Register-tiled edge convolutions (conv_in 3 -> C, conv_out C -> 3 and their gradients, modules.py:219,364) for
// C % 128 == 0.  These layers move one C-channel fp32 tensor (1.07 GB at 32 x 256 x 256 x 128) for 3456 MACs per pixel,
// so the target is the HBM floor with the FFMA pipe close behind; what limited the first kernels (edge.cu) was neither
// but the load/shared-memory instruction rate (one broadcast LDS per FMA).  Here every FMA takes both operands from
// registers:
//   * lane = four wide-side channels (one 16-byte access per pixel, a warp covers 128 channels = 512 contiguous bytes)
//   * warp = a strip of 2 image rows, walked left to right with a sliding window of 4 rows x 3 columns held in registers
//     (three named column buffers rotate through the roles left / centre / right, so nothing is copied)
//   * the 27 x 4 weights (forward / data gradient) or 27 x 4 accumulators (weight gradient) of the lane's channels live in
//     registers for the whole strip; the 3-channel side is read through warp-uniform loads (one L1 wavefront each)
//   * conv_out's forward needs a sum over the 128 channels = across the lanes: six values per pixel column are reduced
//     with a recursive-halving exchange (4+2+1+1+1 shuffles instead of 6 x 5)
// Weight/bias gradients are deterministic: static strip assignment, warps folded in order, fixed-order final reduction.
#include "edge.cuh"

namespace mas {
namespace {

constexpr int QT_H = 16, QT_W = 32;  // block tile: 8 warps x 2 rows, 32 columns

__device__ __forceinline__ float4 ld4(const float* p, bool ok) {
  return ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& v) {  // acc += s * v
  acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float s) {
  s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); return fmaf(a.w, b.w, s);
}
// Sum v[0..7] over the 32 lanes; on return lane l holds the total of v[idx(l)], idx = 4*bit4 + 2*bit3 + bit2 of l.
__device__ __forceinline__ float reduce8(const float (&v)[8], int lane, int& idx) {
  const unsigned FULL = 0xffffffffu;
  float r4[4], r2[2], r1;
  bool b = lane & 16;
#pragma unroll
  for (int i = 0; i < 4; ++i) r4[i] = (b ? v[i + 4] : v[i]) + __shfl_xor_sync(FULL, b ? v[i] : v[i + 4], 16);
  b = lane & 8;
#pragma unroll
  for (int i = 0; i < 2; ++i) r2[i] = (b ? r4[i + 2] : r4[i]) + __shfl_xor_sync(FULL, b ? r4[i] : r4[i + 2], 8);
  b = lane & 4;
  r1 = (b ? r2[1] : r2[0]) + __shfl_xor_sync(FULL, b ? r2[0] : r2[1], 4);
  r1 += __shfl_xor_sync(FULL, r1, 2);
  r1 += __shfl_xor_sync(FULL, r1, 1);
  idx = ((lane & 16) ? 4 : 0) + ((lane & 8) ? 2 : 0) + ((lane & 4) ? 1 : 0);
  return r1;
}

// ------------------------------------------------------------------------------------------------ C -> 3 forward
// ys[n,co,oy,ox] = bias[co] + sum_{ci,tap} a[n,oy+ty-1,ox+tx-1,ci] * w[(co*128+ci)*9+tap]      (Cbig == 128)
__global__ void __launch_bounds__(256, 1) small_cout_fprop_q(const float* __restrict__ a, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ ys, EdgeGeom g,
                                                             int tiles_x, int tiles_y) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 wr[SC][9];
#pragma unroll
  for (int co = 0; co < SC; ++co)
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float* p = w + ((size_t)co * 128 + 4 * lane) * 9 + tap;
      wr[co][tap] = make_float4(__ldg(p), __ldg(p + 9), __ldg(p + 18), __ldg(p + 27));
    }
  const float bv = bias ? __ldg(bias + (((lane & 16) ? 4 : 0) + ((lane & 8) ? 2 : 0) + ((lane & 4) ? 1 : 0)) % 3) : 0.f;
  const int64_t ntiles = (int64_t)g.N * tiles_x * tiles_y;
#pragma unroll 1
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx_ = (int)(tile % tiles_x), ty_ = (int)((tile / tiles_x) % tiles_y), n = (int)(tile / ((int64_t)tiles_x * tiles_y));
    const int oy0 = ty_ * QT_H + warp * 2, x0 = tx_ * QT_W;
    if (oy0 >= g.H) continue;
    const float* row[4];
    bool rok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int iy = oy0 - 1 + j;
      rok[j] = (unsigned)iy < (unsigned)g.H;
      row[j] = a + ((size_t)(n * g.H + (rok[j] ? iy : 0)) * g.W) * 128 + 4 * lane;
    }
    float4 A[4], B[4], C[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      A[j] = ld4(row[j] + (int64_t)(x0 - 1) * 128, rok[j] && x0 > 0);
      B[j] = ld4(row[j] + (int64_t)x0 * 128, rok[j]);
      C[j] = ld4(row[j] + (int64_t)(x0 + 1) * 128, rok[j] && x0 + 1 < g.W);
    }
    // one column: the left buffer is consumed first and immediately refilled with column c+2, so that load has the rest of
    // this step and two thirds of the next one to land (HBM latency, 8 warps per SM)
    auto step = [&](float4 (&Lc)[4], const float4 (&Cc)[4], const float4 (&Rc)[4], int c) {
      float v[8];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int co = 0; co < SC; ++co) {
          float s = 0.f;
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) s = dot4(Lc[r + ty], wr[co][ty * 3 + 0], s);
          v[r * 3 + co] = s;
        }
      const int ix = x0 + c + 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) Lc[j] = ld4(row[j] + (int64_t)ix * 128, rok[j] && ix < g.W);
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int co = 0; co < SC; ++co) {
          float s = v[r * 3 + co];
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) {
            s = dot4(Cc[r + ty], wr[co][ty * 3 + 1], s);
            s = dot4(Rc[r + ty], wr[co][ty * 3 + 2], s);
          }
          v[r * 3 + co] = s;
        }
      v[6] = v[7] = 0.f;
      int idx;
      const float tot = reduce8(v, lane, idx);
      const int oy = oy0 + idx / 3, ox = x0 + c;
      if ((lane & 3) == 0 && idx < 6 && oy < g.H && ox < g.W) ys[n * g.sn + (idx % 3) * g.sc + oy * g.sh + ox * g.sw] = tot + bv;
    };
#pragma unroll 1
    for (int c = 0; c < QT_W; c += 3) {
      step(A, B, C, c);
      if (c + 1 < QT_W) step(B, C, A, c + 1);
      if (c + 2 < QT_W) step(C, A, B, c + 2);
    }
  }
}

// ------------------------------------------------------------------------------------------------ C -> 3 weight gradient
// dW[co][ci][tap] = sum_p dys[n,co,p] * a[p+tap][ci], db[co] = sum_p dys;  part[block][27+3][Cbig]
__global__ void __launch_bounds__(256, 1) small_cout_wgrad_q(const float* __restrict__ a, const float* __restrict__ dys,
                                                             float* __restrict__ part, EdgeGeom g, int tiles_x, int tiles_y) {
  __shared__ __align__(16) float red[SC * 9 + SC][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cb = blockIdx.y;
  float4 acc[SC][9];
  float bs[SC];
#pragma unroll
  for (int co = 0; co < SC; ++co) {
    bs[co] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) acc[co][tap] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int64_t ntiles = (int64_t)g.N * tiles_x * tiles_y;
#pragma unroll 1
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx_ = (int)(tile % tiles_x), ty_ = (int)((tile / tiles_x) % tiles_y), n = (int)(tile / ((int64_t)tiles_x * tiles_y));
    const int oy0 = ty_ * QT_H + warp * 2, x0 = tx_ * QT_W;
    if (oy0 >= g.H) continue;
    const float* row[4];
    bool rok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int iy = oy0 - 1 + j;
      rok[j] = (unsigned)iy < (unsigned)g.H;
      row[j] = a + ((size_t)(n * g.H + (rok[j] ? iy : 0)) * g.W) * g.Cbig + cb * 128 + 4 * lane;
    }
    const float* dbase = dys + n * g.sn;
    float4 A[4], B[4], C[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      A[j] = ld4(row[j] + (int64_t)(x0 - 1) * g.Cbig, rok[j] && x0 > 0);
      B[j] = ld4(row[j] + (int64_t)x0 * g.Cbig, rok[j]);
      C[j] = ld4(row[j] + (int64_t)(x0 + 1) * g.Cbig, rok[j] && x0 + 1 < g.W);
    }
    float dn[2][SC];  // gradient values of the NEXT column (prefetched one step ahead)
    auto lddy = [&](int ox) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int co = 0; co < SC; ++co)
          dn[r][co] = (ox < g.W && oy0 + r < g.H) ? __ldg(dbase + co * g.sc + (oy0 + r) * g.sh + ox * g.sw) : 0.f;
    };
    lddy(x0);
    auto step = [&](float4 (&Lc)[4], const float4 (&Cc)[4], const float4 (&Rc)[4], int c) {
      float d[2][SC];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int co = 0; co < SC; ++co) {
          d[r][co] = dn[r][co];
          bs[co] += d[r][co];
        }
      lddy(x0 + c + 1);
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int co = 0; co < SC; ++co)
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) fma4(acc[co][ty * 3 + 0], d[r][co], Lc[r + ty]);
      const int ix = x0 + c + 2;  // refill the consumed left buffer two columns ahead
#pragma unroll
      for (int j = 0; j < 4; ++j) Lc[j] = ld4(row[j] + (int64_t)ix * g.Cbig, rok[j] && ix < g.W);
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int co = 0; co < SC; ++co)
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) {
            fma4(acc[co][ty * 3 + 1], d[r][co], Cc[r + ty]);
            fma4(acc[co][ty * 3 + 2], d[r][co], Rc[r + ty]);
          }
    };
#pragma unroll 1
    for (int c = 0; c < QT_W; c += 3) {
      step(A, B, C, c);
      if (c + 1 < QT_W) step(B, C, A, c + 1);
      if (c + 2 < QT_W) step(C, A, B, c + 2);
    }
  }
  // fold the eight warps in order, then one row of partials per block
  for (int i = threadIdx.x; i < (SC * 9 + SC) * 128; i += 256) (&red[0][0])[i] = 0.f;
  __syncthreads();
  for (int wv = 0; wv < 8; ++wv) {
    if (warp == wv) {
#pragma unroll
      for (int co = 0; co < SC; ++co)
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          float4* p = reinterpret_cast<float4*>(&red[co * 9 + tap][4 * lane]);
          float4 t = *p;
          const float4 s = acc[co][tap];
          t.x += s.x; t.y += s.y; t.z += s.z; t.w += s.w;
          *p = t;
        }
      if (lane == 0)
#pragma unroll
        for (int co = 0; co < SC; ++co) red[SC * 9 + co][0] += bs[co];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < (SC * 9 + SC) * 128; i += 256) {
    const int r = i >> 7, col = i & 127;
    part[((size_t)blockIdx.x * (SC * 9 + SC) + r) * g.Cbig + cb * 128 + col] = red[r][col];
  }
}

// ------------------------------------------------------------------------------------------------ 3 -> C forward
// y[n,oy,ox,co] = bias[co] + sum_{ci<3,tap} xs[n,ci,oy+ty-1,ox+tx-1] * W(co,ci,tap)   (W as in edge.cu: flipT selects the
// conv_in weight [C][3][3][3] or conv_out's [3][C][3][3] flipped = conv_out's data gradient)
struct Win3 {
  float v[4][SC];
};
__global__ void __launch_bounds__(256, 1) small_cin_fprop_q(const float* __restrict__ xs, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y, EdgeGeom g, int flipT,
                                                            int tiles_x, int tiles_y) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cb = blockIdx.y;
  const int co0 = cb * 128 + 4 * lane;
  float4 wr[SC * 9];
#pragma unroll
  for (int ci = 0; ci < SC; ++ci)
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      float e[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        e[k] = flipT ? __ldg(w + ((size_t)ci * g.Cbig + co0 + k) * 9 + (8 - tap)) : __ldg(w + ((size_t)(co0 + k) * SC + ci) * 9 + tap);
      wr[ci * 9 + tap] = make_float4(e[0], e[1], e[2], e[3]);
    }
  const float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias + co0)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t ntiles = (int64_t)g.N * tiles_x * tiles_y;
#pragma unroll 1
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx_ = (int)(tile % tiles_x), ty_ = (int)((tile / tiles_x) % tiles_y), n = (int)(tile / ((int64_t)tiles_x * tiles_y));
    const int oy0 = ty_ * QT_H + warp * 2, x0 = tx_ * QT_W;
    if (oy0 >= g.H) continue;
    const float* xb = xs + n * g.sn;
    int64_t roff[4];
    bool rok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int iy = oy0 - 1 + j;
      rok[j] = (unsigned)iy < (unsigned)g.H;
      roff[j] = (rok[j] ? iy : 0) * g.sh;
    }
    auto ldcol = [&](Win3& wn, int ix) {
      const bool cok = (unsigned)ix < (unsigned)g.W;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int ci = 0; ci < SC; ++ci) wn.v[j][ci] = (cok && rok[j]) ? __ldg(xb + ci * g.sc + roff[j] + ix * g.sw) : 0.f;
    };
    Win3 A, B, C;
    ldcol(A, x0 - 1);
    ldcol(B, x0);
    ldcol(C, x0 + 1);
    auto step = [&](Win3& Lc, const Win3& Cc, const Win3& Rc, int c) {
      const int ox = x0 + c;
      float4 o[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        o[r] = b4;
#pragma unroll
        for (int ci = 0; ci < SC; ++ci)
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) fma4(o[r], Lc.v[r + ty][ci], wr[ci * 9 + ty * 3 + 0]);
      }
      ldcol(Lc, ox + 2);  // refill the consumed left buffer two columns ahead
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int ci = 0; ci < SC; ++ci)
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) {
            fma4(o[r], Cc.v[r + ty][ci], wr[ci * 9 + ty * 3 + 1]);
            fma4(o[r], Rc.v[r + ty][ci], wr[ci * 9 + ty * 3 + 2]);
          }
        const int oy = oy0 + r;
        if (oy < g.H && ox < g.W) *reinterpret_cast<float4*>(y + ((size_t)(n * g.H + oy) * g.W + ox) * g.Cbig + co0) = o[r];
      }
    };
#pragma unroll 1
    for (int c = 0; c < QT_W; c += 3) {
      step(A, B, C, c);
      if (c + 1 < QT_W) step(B, C, A, c + 1);
      if (c + 2 < QT_W) step(C, A, B, c + 2);
    }
  }
}

// ------------------------------------------------------------------------------------------------ 3 -> C weight gradient
// dW[co][ci][tap] = sum_p dy[p][co] * xs[p+tap][ci], db[co] = sum_p dy[p][co];  part[block][27+1][Cbig], row i = ci*9+tap
__global__ void __launch_bounds__(256, 1) small_cin_wgrad_q(const float* __restrict__ xs, const float* __restrict__ dy,
                                                            float* __restrict__ part, EdgeGeom g, int tiles_x, int tiles_y) {
  __shared__ __align__(16) float red[SC * 9 + 1][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cb = blockIdx.y;
  const int co0 = cb * 128 + 4 * lane;
  float4 acc[SC * 9 + 1];
#pragma unroll
  for (int i = 0; i < SC * 9 + 1; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t ntiles = (int64_t)g.N * tiles_x * tiles_y;
#pragma unroll 1
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx_ = (int)(tile % tiles_x), ty_ = (int)((tile / tiles_x) % tiles_y), n = (int)(tile / ((int64_t)tiles_x * tiles_y));
    const int oy0 = ty_ * QT_H + warp * 2, x0 = tx_ * QT_W;
    if (oy0 >= g.H) continue;
    const float* xb = xs + n * g.sn;
    int64_t roff[4];
    bool rok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int iy = oy0 - 1 + j;
      rok[j] = (unsigned)iy < (unsigned)g.H;
      roff[j] = (rok[j] ? iy : 0) * g.sh;
    }
    auto ldcol = [&](Win3& wn, int ix) {
      const bool cok = (unsigned)ix < (unsigned)g.W;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int ci = 0; ci < SC; ++ci) wn.v[j][ci] = (cok && rok[j]) ? __ldg(xb + ci * g.sc + roff[j] + ix * g.sw) : 0.f;
    };
    Win3 A, B, C;
    ldcol(A, x0 - 1);
    ldcol(B, x0);
    ldcol(C, x0 + 1);
    float4 dn[2];  // output-gradient quads of the NEXT column (prefetched one step ahead: they stream from HBM)
    auto lddy = [&](int ox) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int oy = oy0 + r;
        dn[r] = ld4(dy + ((size_t)(n * g.H + (oy < g.H ? oy : 0)) * g.W + (ox < g.W ? ox : 0)) * g.Cbig + co0, oy < g.H && ox < g.W);
      }
    };
    lddy(x0);
    auto step = [&](Win3& Lc, const Win3& Cc, const Win3& Rc, int c) {
      const int ox = x0 + c;
      const float4 d0 = dn[0], d1 = dn[1];
      lddy(ox + 1);
      acc[SC * 9].x += d0.x + d1.x; acc[SC * 9].y += d0.y + d1.y; acc[SC * 9].z += d0.z + d1.z; acc[SC * 9].w += d0.w + d1.w;
#pragma unroll
      for (int ci = 0; ci < SC; ++ci)
#pragma unroll
        for (int ty = 0; ty < 3; ++ty) {
          fma4(acc[ci * 9 + ty * 3 + 0], Lc.v[ty][ci], d0);
          fma4(acc[ci * 9 + ty * 3 + 0], Lc.v[1 + ty][ci], d1);
        }
      ldcol(Lc, ox + 2);  // refill the consumed left buffer two columns ahead
#pragma unroll
      for (int ci = 0; ci < SC; ++ci)
#pragma unroll
        for (int ty = 0; ty < 3; ++ty) {
          fma4(acc[ci * 9 + ty * 3 + 1], Cc.v[ty][ci], d0);
          fma4(acc[ci * 9 + ty * 3 + 1], Cc.v[1 + ty][ci], d1);
          fma4(acc[ci * 9 + ty * 3 + 2], Rc.v[ty][ci], d0);
          fma4(acc[ci * 9 + ty * 3 + 2], Rc.v[1 + ty][ci], d1);
        }
    };
#pragma unroll 1
    for (int c = 0; c < QT_W; c += 3) {
      step(A, B, C, c);
      if (c + 1 < QT_W) step(B, C, A, c + 1);
      if (c + 2 < QT_W) step(C, A, B, c + 2);
    }
  }
  for (int i = threadIdx.x; i < (SC * 9 + 1) * 128; i += 256) (&red[0][0])[i] = 0.f;
  __syncthreads();
  for (int wv = 0; wv < 8; ++wv) {
    if (warp == wv) {
#pragma unroll
      for (int i = 0; i < SC * 9 + 1; ++i) {
        float4* p = reinterpret_cast<float4*>(&red[i][4 * lane]);
        float4 t = *p;
        t.x += acc[i].x; t.y += acc[i].y; t.z += acc[i].z; t.w += acc[i].w;
        *p = t;
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < (SC * 9 + 1) * 128; i += 256) {
    const int r = i >> 7, col = i & 127;
    part[((size_t)blockIdx.x * (SC * 9 + 1) + r) * g.Cbig + cb * 128 + col] = red[r][col];
  }
}

}  // namespace

int small_cin_fprop_q_launch(const float* xs, const float* w, const float* bias, float* y, const EdgeGeom& g, int flipT, cudaStream_t st) {
  const int tx = (int)cdiv(g.W, QT_W), ty = (int)cdiv(g.H, QT_H);
  small_cin_fprop_q<<<dim3(EDGE_Q_BLOCKS, g.Cbig / 128), 256, 0, st>>>(xs, w, bias, y, g, flipT, tx, ty);
  return launched("small_cin_fprop_q");
}
int small_cout_fprop_q_launch(const float* a, const float* w, const float* bias, float* ys, const EdgeGeom& g, cudaStream_t st) {
  const int tx = (int)cdiv(g.W, QT_W), ty = (int)cdiv(g.H, QT_H);
  small_cout_fprop_q<<<EDGE_Q_BLOCKS, 256, 0, st>>>(a, w, bias, ys, g, tx, ty);
  return launched("small_cout_fprop_q");
}
int small_cin_wgrad_q_launch(const float* xs, const float* dy, float* part, const EdgeGeom& g, cudaStream_t st) {
  const int tx = (int)cdiv(g.W, QT_W), ty = (int)cdiv(g.H, QT_H);
  small_cin_wgrad_q<<<dim3(EDGE_Q_BLOCKS, g.Cbig / 128), 256, 0, st>>>(xs, dy, part, g, tx, ty);
  return launched("small_cin_wgrad_q");
}
int small_cout_wgrad_q_launch(const float* a, const float* dys, float* part, const EdgeGeom& g, cudaStream_t st) {
  const int tx = (int)cdiv(g.W, QT_W), ty = (int)cdiv(g.H, QT_H);
  small_cout_wgrad_q<<<dim3(EDGE_Q_BLOCKS, g.Cbig / 128), 256, 0, st>>>(a, dys, part, g, tx, ty);
  return launched("small_cout_wgrad_q");
}

}  // namespace mas
