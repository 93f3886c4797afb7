// Autoregressive sampling of the token transformer: KV cache + single-token decode kernels.
// The reference has no working implementation of this path (models/transformer.py:73-115 vs :176-210 disagree on the
// cache layout and train.py never samples); the specification is the non-cached forward (transformer.py:77-103,
// 216-244): a decode step must reproduce the logits the full causal forward gives at that position.
//   * mas_linear_small : y[r,n] = act(sum_k x[r,k] W[n,k] + b[n]) for a handful of rows (cond + uncond streams): a
//     weight-streaming kernel — every weight is read exactly once per token, so it is HBM-bound by construction
//     (371 M parameters = 1.48 GB per token for the 24-layer / 1024-wide model); strict fp32.
//   * mas_kv_append    : scatters the k / v thirds of a fused qkv activation into the [R, heads, Tmax, hd] caches.
//   * mas_attn_decode  : one query per (row, head) against the cache: scores, softmax, weighted sum of v.
//   * mas_cfg_mix      : classifier-free guidance, out = uncond + scale * (cond - uncond).
#include "mas_common.cuh"

using namespace mas;

namespace {

constexpr int LS_MAXR = 8;

// one warp per output column n; lanes stride over k in float4 steps (W row n is contiguous: 512-byte warp transactions)
template <int R>
__global__ void __launch_bounds__(256) linear_small_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ W,
                                                           const float* __restrict__ bias, float* __restrict__ y, int64_t ldy, int N,
                                                           int K, int act) {
  const int lane = threadIdx.x & 31, n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)n * K);
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  const int K4 = K >> 2;
  int k = lane;
  for (; k + 96 < K4; k += 128) {  // four weight quads in flight per lane
    float4 wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) wv[u] = __ldg(w4 + k + 32 * u);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * ldx) + k + 32 * u);
        acc[r] = fmaf(wv[u].x, xv.x, acc[r]); acc[r] = fmaf(wv[u].y, xv.y, acc[r]);
        acc[r] = fmaf(wv[u].z, xv.z, acc[r]); acc[r] = fmaf(wv[u].w, xv.w, acc[r]);
      }
  }
  for (; k < K4; k += 32) {
    const float4 wv = __ldg(w4 + k);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * ldx) + k);
      acc[r] = fmaf(wv.x, xv.x, acc[r]); acc[r] = fmaf(wv.y, xv.y, acc[r]);
      acc[r] = fmaf(wv.z, xv.z, acc[r]); acc[r] = fmaf(wv.w, xv.w, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = warp_sum(acc[r]);
  if (lane == 0) {
    const float b = bias ? __ldg(bias + n) : 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float v = acc[r] + b;
      if (act == 1) {  // OpenAI tanh-GELU, transformer.py:11-14
        const float u = 0.7978845608028654f * v * (1.f + 0.044715f * v * v);
        v = 0.5f * v * (1.f + tanhf(u));
      }
      y[(size_t)r * ldy + n] = v;
    }
  }
}

// qkv [R, T, 3H] -> kcache / vcache [R, heads, Tmax, hd] at positions pos0 .. pos0+T-1
__global__ void kv_append_kernel(const float* __restrict__ qkv, int R, int T, int heads, int hd, float* __restrict__ kc,
                                 float* __restrict__ vc, int Tmax, int pos0) {
  const int H = heads * hd, Q = hd >> 2;
  const int64_t total = (int64_t)R * T * heads * Q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r_ = i;
    const int q = (int)(r_ % Q); r_ /= Q;
    const int h = (int)(r_ % heads); r_ /= heads;
    const int t = (int)(r_ % T);
    const int r = (int)(r_ / T);
    const float* src = qkv + ((size_t)r * T + t) * 3 * H + h * hd + q * 4;
    const size_t dst = (((size_t)r * heads + h) * Tmax + pos0 + t) * hd + q * 4;
    *reinterpret_cast<float4*>(kc + dst) = __ldg(reinterpret_cast<const float4*>(src + H));
    *reinterpret_cast<float4*>(vc + dst) = __ldg(reinterpret_cast<const float4*>(src + 2 * H));
  }
}

// block per (row, head), 128 threads; q = this token's query (row r of qkv [R, 3H]); len cached positions (incl. this one)
__global__ void __launch_bounds__(128) attn_decode_kernel(const float* __restrict__ qkv, const float* __restrict__ kc,
                                                          const float* __restrict__ vc, float* __restrict__ ctx, int heads, int hd,
                                                          int Tmax, int len) {
  extern __shared__ float sm[];  // [hd] q, [len] scores, [128] scratch, [128] partial outputs
  float* qs = sm;
  float* sc = sm + hd;
  float* red = sc + len;
  float* po = red + 128;
  const int r = blockIdx.x / heads, h = blockIdx.x % heads, H = heads * hd, t0 = threadIdx.x;
  for (int d = t0; d < hd; d += 128) qs[d] = qkv[(size_t)r * 3 * H + h * hd + d];
  __syncthreads();
  const float alpha = rsqrtf((float)hd);
  const float* kb = kc + ((size_t)r * heads + h) * Tmax * hd;
  const float* vb = vc + ((size_t)r * heads + h) * Tmax * hd;
  float mx = -INFINITY;
  for (int t = t0; t < len; t += 128) {
    const float4* kr = reinterpret_cast<const float4*>(kb + (size_t)t * hd);
    float s = 0.f;
    for (int d4 = 0; d4 < (hd >> 2); ++d4) {
      const float4 kv = __ldg(kr + d4);
      s = fmaf(qs[d4 * 4 + 0], kv.x, s); s = fmaf(qs[d4 * 4 + 1], kv.y, s);
      s = fmaf(qs[d4 * 4 + 2], kv.z, s); s = fmaf(qs[d4 * 4 + 3], kv.w, s);
    }
    s *= alpha;
    sc[t] = s;
    mx = fmaxf(mx, s);
  }
  red[t0] = mx;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (t0 < o) red[t0] = fmaxf(red[t0], red[t0 + o]);
    __syncthreads();
  }
  mx = red[0];
  __syncthreads();
  float sum = 0.f;
  for (int t = t0; t < len; t += 128) {
    const float e = expf(sc[t] - mx);
    sc[t] = e;
    sum += e;
  }
  red[t0] = sum;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (t0 < o) red[t0] += red[t0 + o];
    __syncthreads();
  }
  const float inv = 1.f / red[0];
  // ctx[d] = sum_t p_t v[t][d]: the 128 threads form 128/hd groups that split the t range; reads coalesced over d
  const int groups = 128 / hd, d = t0 % hd, gidx = t0 / hd;
  float o_ = 0.f;
  for (int t = gidx; t < len; t += groups) o_ = fmaf(sc[t], __ldg(vb + (size_t)t * hd + d), o_);
  po[gidx * hd + d] = o_;
  __syncthreads();
  if (t0 < hd) {
    float v = 0.f;
    for (int g2 = 0; g2 < groups; ++g2) v += po[g2 * hd + t0];
    ctx[(size_t)r * H + h * hd + t0] = v * inv;
  }
}

__global__ void cfg_mix_kernel(const float* __restrict__ cond, const float* __restrict__ uncond, float* __restrict__ out, int64_t n,
                               float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float u = uncond[i];
    out[i] = fmaf(scale, cond[i] - u, u);
  }
}

template <int R>
int linear_small_run(const float* x, int64_t ldx, const float* W, const float* bias, float* y, int64_t ldy, int N, int K, int act,
                     cudaStream_t st) {
  linear_small_kernel<R><<<(int)cdiv(N, 8), 256, 0, st>>>(x, ldx, W, bias, y, ldy, N, K, act);
  return launched("linear_small");
}

}  // namespace

extern "C" {

int mas_linear_small(const float* x, int64_t ldx, const float* W, const float* bias, float* y, int64_t ldy, int R, int N, int K, int act,
                     void* stream) {
  MAS_REQUIRE(x && W && y && R > 0 && N > 0 && K > 0, "linear_small: bad arguments");
  if (R > LS_MAXR) return fail(MAS_ERR_UNSUPPORTED, "linear_small: at most %d rows (got %d); use mas_gemm_rows_packed", LS_MAXR, R);
  if (K % 4 || ldx % 4 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return fail(MAS_ERR_UNSUPPORTED, "linear_small: K and ldx must be multiples of 4 and x / W 16-byte aligned");
  if (act != 0 && act != 1) return fail(MAS_ERR_INVALID_ARG, "linear_small: act must be 0 (none) or 1 (tanh-GELU)");
  cudaStream_t st = S(stream);
  switch (R) {
    case 1: return linear_small_run<1>(x, ldx, W, bias, y, ldy, N, K, act, st);
    case 2: return linear_small_run<2>(x, ldx, W, bias, y, ldy, N, K, act, st);
    case 3: return linear_small_run<3>(x, ldx, W, bias, y, ldy, N, K, act, st);
    case 4: return linear_small_run<4>(x, ldx, W, bias, y, ldy, N, K, act, st);
    case 5: return linear_small_run<5>(x, ldx, W, bias, y, ldy, N, K, act, st);
    case 6: return linear_small_run<6>(x, ldx, W, bias, y, ldy, N, K, act, st);
    case 7: return linear_small_run<7>(x, ldx, W, bias, y, ldy, N, K, act, st);
    default: return linear_small_run<8>(x, ldx, W, bias, y, ldy, N, K, act, st);
  }
}

int mas_kv_append(const float* qkv, int R, int T, int heads, int hd, float* kcache, float* vcache, int Tmax, int pos0, void* stream) {
  MAS_REQUIRE(qkv && kcache && vcache && R > 0 && T > 0 && heads > 0 && hd > 0, "kv_append: bad arguments");
  if (hd % 4) return fail(MAS_ERR_UNSUPPORTED, "kv_append: head dim %% 4 != 0");
  if (pos0 < 0 || pos0 + T > Tmax) return fail(MAS_ERR_INVALID_ARG, "kv_append: positions %d..%d outside the cache (%d)", pos0, pos0 + T, Tmax);
  const int64_t total = (int64_t)R * T * heads * (hd / 4);
  kv_append_kernel<<<(int)(cdiv(total, 256) < 2368 ? cdiv(total, 256) : 2368), 256, 0, S(stream)>>>(qkv, R, T, heads, hd, kcache, vcache,
                                                                                                 Tmax, pos0);
  return launched("kv_append");
}

int mas_attn_decode(const float* qkv, const float* kcache, const float* vcache, float* ctx, int R, int heads, int hd, int Tmax, int len,
                    void* stream) {
  MAS_REQUIRE(qkv && kcache && vcache && ctx && R > 0 && heads > 0, "attn_decode: bad arguments");
  if (hd % 4 || hd > 128 || 128 % hd) return fail(MAS_ERR_UNSUPPORTED, "attn_decode: head dim must divide 128 and be a multiple of 4 (got %d)", hd);
  if (len <= 0 || len > Tmax) return fail(MAS_ERR_INVALID_ARG, "attn_decode: cache length %d outside (0, %d]", len, Tmax);
  const size_t smem = (size_t)(hd + len + 128 + 128) * sizeof(float);
  if (smem > 48 * 1024) return fail(MAS_ERR_UNSUPPORTED, "attn_decode: sequence too long for the single-pass kernel (%d)", len);
  attn_decode_kernel<<<R * heads, 128, smem, S(stream)>>>(qkv, kcache, vcache, ctx, heads, hd, Tmax, len);
  return launched("attn_decode");
}

int mas_cfg_mix(const float* cond, const float* uncond, float* out, int64_t n, float scale, void* stream) {
  MAS_REQUIRE(cond && uncond && out && n > 0, "cfg_mix: bad arguments");
  cfg_mix_kernel<<<(int)(cdiv(n, 256) < 1184 ? cdiv(n, 256) : 1184), 256, 0, S(stream)>>>(cond, uncond, out, n, scale);
  return launched("cfg_mix");
}

}  // extern "C"
