// Codebook (vector quantiser) kernels — replaces modules.py:501-517 (distance matrix, argmin, gather,
// loss, straight-through) with one fused kernel; the [R,K] distance matrix is never written.
//
// Arithmetic contract: d[r,k] = fl( fl(|z_r|^2 + |e_k|^2) - 2*dot(z_r,e_k) ) with the
// dot product accumulated in fp32 FMA over k ascending, strict '<' scan over codes ascending =>
// first-index tie-break, i.e. the reference's formula, association and argmin semantics.  Pure fp32 FFMA:
// no TF32/BF16 on this contraction, so indices agree with a strict-fp32 oracle except on rows whose two best
// distances are within a few ulp (classified by the tests with an fp64 gap).
#include "mas_common.cuh"

namespace mas {

constexpr int VQ_BM = 64;    // latent rows per CTA
constexpr int VQ_BN = 128;   // codes per tile
constexpr int VQ_BK = 32;    // dims per smem stage
constexpr int VQ_LDE = VQ_BK + 4;
constexpr int VQ_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// |e_k|^2 per code: one warp per code, lanes stride the row, xor-tree combine (deterministic)
__global__ void vq_code_norms(const float* __restrict__ E, int K, int D, float* __restrict__ ee) {
  int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= K) return;
  int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) {
    float v = E[(size_t)k * D + d];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) ee[k] = s;
}

__global__ void __launch_bounds__(VQ_THREADS, 2)
vq_forward_kernel(const float* __restrict__ z, const float* __restrict__ E, const float* __restrict__ ee, int64_t R, int K, int D,
                  int64_t* __restrict__ idx_out, float* __restrict__ zq_out, double* __restrict__ loss_part) {
  extern __shared__ __align__(16) float smem[];
  const int LDZ = D + 4;
  float* Zs = smem;                             // [VQ_BM][LDZ]
  float* Es = Zs + VQ_BM * LDZ;                 // [2][VQ_BN][VQ_LDE]
  float* zz_s = Es + 2 * VQ_BN * VQ_LDE;        // [VQ_BM]
  int* idx_s = reinterpret_cast<int*>(zz_s + VQ_BM);  // [VQ_BM]
  __shared__ double red_s[VQ_THREADS / 32];

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int tx = t & 15, ty = t >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * VQ_BM;
  const int nkc = D / VQ_BK, ntile = (K + VQ_BN - 1) / VQ_BN, nstage = nkc * ntile;

  auto load_stage = [&](int s) {
    const int ct = s / nkc, kc = s % nkc;
    float* dst = Es + (s & 1) * VQ_BN * VQ_LDE;
#pragma unroll
    for (int i = 0; i < (VQ_BN * VQ_BK / 4) / VQ_THREADS; ++i) {
      int f = t + i * VQ_THREADS, code = f >> 3, kq = f & 7;
      int gcode = min(ct * VQ_BN + code, K - 1);
      cp_async16(dst + code * VQ_LDE + kq * 4, E + (size_t)gcode * D + kc * VQ_BK + kq * 4);
    }
    cp_async_commit();
  };
  load_stage(0);

  // z tile -> shared (zero rows beyond R), and |z|^2 per row
  for (int f = t; f < VQ_BM * (D / 4); f += VQ_THREADS) {
    int r = f / (D / 4), q = f % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < R) v = __ldg(reinterpret_cast<const float4*>(z + (size_t)(row0 + r) * D) + q);
    *reinterpret_cast<float4*>(Zs + r * LDZ + q * 4) = v;
  }
  __syncthreads();
  for (int r = warp * (VQ_BM / 8); r < (warp + 1) * (VQ_BM / 8); ++r) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) {
      float v = Zs[r * LDZ + d];
      s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) zz_s[r] = s;
  }

  float acc[4][8];
  float best[4];
  int bidx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    best[i] = INFINITY;
    bidx[i] = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  }

  for (int s = 0; s < nstage; ++s) {
    if (s + 1 < nstage) {
      load_stage(s + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int ct = s / nkc, kc = s % nkc;
    const float* es = Es + (s & 1) * VQ_BN * VQ_LDE;
    const float* zs = Zs + kc * VQ_BK;
#pragma unroll
    for (int k4 = 0; k4 < VQ_BK / 4; ++k4) {
      float4 a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(zs + (ty * 4 + i) * LDZ + k4 * 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const float4*>(es + (j * 16 + tx) * VQ_LDE + k4 * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    }
    if (kc == nkc - 1) {  // tile finished: distances + running argmin (codes ascending inside a thread)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int code = ct * VQ_BN + j * 16 + tx;
        if (code < K) {
          float e2 = __ldg(ee + code);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float d = (zz_s[ty * 4 + i] + e2) - 2.0f * acc[i][j];
            if (d < best[i]) {
              best[i] = d;
              bidx[i] = code;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = 0.f;
      }
    }
    __syncthreads();  // stage buffer (s&1) is refilled by the prefetch issued in iteration s+1
  }

  // combine the 16 tx-lanes of each row: smaller distance wins, ties -> smaller index (first occurrence)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best[i], o);
      int oi = __shfl_xor_sync(0xffffffffu, bidx[i], o);
      if (ov < best[i] || (ov == best[i] && oi < bidx[i])) {
        best[i] = ov;
        bidx[i] = oi;
      }
    }
    if (tx == 0) idx_s[ty * 4 + i] = bidx[i];
  }
  __syncthreads();

  // gather + straight-through value + loss partial (modules.py:506-512)
  float ls = 0.f;
  for (int f = t; f < VQ_BM * (D / 4); f += VQ_THREADS) {
    int r = f / (D / 4), q = f % (D / 4);
    if (row0 + r < R) {
      int k = idx_s[r];
      float4 e = __ldg(reinterpret_cast<const float4*>(E + (size_t)k * D) + q);
      float4 zv = *reinterpret_cast<const float4*>(Zs + r * LDZ + q * 4);
      float4 df = make_float4(e.x - zv.x, e.y - zv.y, e.z - zv.z, e.w - zv.w);
      ls += df.x * df.x + df.y * df.y + df.z * df.z + df.w * df.w;
      // z + (z_q - z).detach(): the forward VALUE carries these two roundings in the reference too
      reinterpret_cast<float4*>(zq_out + (size_t)(row0 + r) * D)[q] =
          make_float4(zv.x + df.x, zv.y + df.y, zv.z + df.z, zv.w + df.w);
    }
  }
  for (int r = t; r < VQ_BM; r += VQ_THREADS)
    if (row0 + r < R) idx_out[row0 + r] = (int64_t)idx_s[r];
  float w = warp_sum(ls);
  if (lane == 0) red_s[warp] = (double)w;
  __syncthreads();
  if (t == 0) {
    double a = 0;
    for (int k = 0; k < VQ_THREADS / 32; ++k) a += red_s[k];
    loss_part[blockIdx.x] = a;
  }
}

__global__ void vq_loss_final(const double* __restrict__ part, int n, double inv_count, float beta, float* __restrict__ out) {
  __shared__ double sh[256];
  double a = 0;
  for (int i = threadIdx.x; i < n; i += 256) a += part[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float m = (float)(sh[0] * inv_count);
    out[0] = m + beta * m;  // mean((zq.detach()-z)^2) + beta*mean((zq-z.detach())^2), modules.py:509
  }
}

__global__ void vq_backward_kernel(const float* __restrict__ g_zq, const float* __restrict__ g_loss, const float* __restrict__ z,
                                   const float* __restrict__ E, const int64_t* __restrict__ idx, int64_t R, int D, float beta,
                                   float* __restrict__ grad_z, float* __restrict__ grad_E) {
  const int D4 = D >> 2;
  const float gl = g_loss ? g_loss[0] : 0.f;
  const float cz = gl * 2.0f / (float)((double)R * D), ce = cz * beta;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R * D4; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / D4;
    int q = (int)(i % D4);
    int64_t k = idx[r];
    float4 zv = __ldg(reinterpret_cast<const float4*>(z) + i);
    float4 e = __ldg(reinterpret_cast<const float4*>(E + (size_t)k * D) + q);
    float4 g = g_zq ? __ldg(reinterpret_cast<const float4*>(g_zq) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 df = make_float4(zv.x - e.x, zv.y - e.y, zv.z - e.z, zv.w - e.w);
    if (grad_z) reinterpret_cast<float4*>(grad_z)[i] = make_float4(g.x + cz * df.x, g.y + cz * df.y, g.z + cz * df.z, g.w + cz * df.w);
    if (grad_E && gl != 0.f) {
      float* ge = grad_E + (size_t)k * D + q * 4;
      atomicAdd(ge + 0, -ce * df.x);
      atomicAdd(ge + 1, -ce * df.y);
      atomicAdd(ge + 2, -ce * df.z);
      atomicAdd(ge + 3, -ce * df.w);
    }
  }
}

__global__ void vq_gather_kernel(const float* __restrict__ E, const int64_t* __restrict__ idx, int64_t R, int K, int D,
                                 float* __restrict__ out) {
  const int D4 = D >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R * D4; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / D4, k = idx[r];
    k = k < 0 ? 0 : (k >= K ? K - 1 : k);
    reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(E + (size_t)k * D) + (i % D4));
  }
}

}  // namespace mas

using namespace mas;

extern "C" {

size_t mas_vq_ws_bytes(int64_t R, int K, int D) {
  (void)D;
  return (size_t)K * sizeof(float) + 256 + (size_t)cdiv(R, VQ_BM) * sizeof(double);
}

int mas_vq_forward(const float* z, const float* E, int64_t R, int K, int D, float beta, int64_t* idx_out, float* zq_out,
                   float* loss_out, void* ws, size_t ws_bytes, void* stream) {
  MAS_REQUIRE(R > 0 && K > 0 && D > 0, "vq_forward: bad shape R=%lld K=%d D=%d", (long long)R, K, D);
  if (D % VQ_BK != 0) return fail(MAS_ERR_UNSUPPORTED, "vq_forward: D=%d must be a multiple of %d", D, VQ_BK);
  if (ws_bytes < mas_vq_ws_bytes(R, K, D)) return fail(MAS_ERR_WORKSPACE, "vq_forward: workspace too small");
  float* ee = (float*)ws;
  double* part = (double*)((char*)ws + (((size_t)K * sizeof(float) + 255) / 256) * 256);
  int blocks = (int)cdiv(R, VQ_BM);
  size_t smem = ((size_t)VQ_BM * (D + 4) + 2 * VQ_BN * VQ_LDE + VQ_BM) * sizeof(float) + VQ_BM * sizeof(int);
  static int configured_smem = 0;
  if ((int)smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(vq_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(MAS_ERR_LAUNCH, "vq_forward: smem attr: %s", cudaGetErrorString(e));
    configured_smem = (int)smem;
  }
  vq_code_norms<<<(int)cdiv(K, 8), 256, 0, S(stream)>>>(E, K, D, ee);
  if (int e = launched("vq_code_norms")) return e;
  vq_forward_kernel<<<blocks, VQ_THREADS, smem, S(stream)>>>(z, E, ee, R, K, D, idx_out, zq_out, part);
  if (int e = launched("vq_forward")) return e;
  vq_loss_final<<<1, 256, 0, S(stream)>>>(part, blocks, 1.0 / ((double)R * D), beta, loss_out);
  return launched("vq_loss_final");
}

int mas_vq_backward(const float* g_zq, const float* g_loss, const float* z, const float* E, const int64_t* idx, int64_t R, int K,
                    int D, float beta, float* grad_z, float* grad_E, void* stream) {
  (void)K;
  MAS_REQUIRE(R > 0 && D > 0 && D % 4 == 0, "vq_backward: bad shape");
  int64_t n = R * (D / 4);
  int grid = (int)(cdiv(n, 256) < 148 * 16 ? cdiv(n, 256) : 148 * 16);
  vq_backward_kernel<<<grid, 256, 0, S(stream)>>>(g_zq, g_loss, z, E, idx, R, D, beta, grad_z, grad_E);
  return launched("vq_backward");
}

int mas_vq_gather(const float* E, const int64_t* idx, int64_t R, int K, int D, float* out, void* stream) {
  MAS_REQUIRE(R > 0 && D > 0 && D % 4 == 0, "vq_gather: bad shape");
  int64_t n = R * (D / 4);
  int grid = (int)(cdiv(n, 256) < 148 * 16 ? cdiv(n, 256) : 148 * 16);
  vq_gather_kernel<<<grid, 256, 0, S(stream)>>>(E, idx, R, K, D, out);
  return launched("vq_gather");
}

}  // extern "C"
