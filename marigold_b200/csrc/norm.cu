// GroupNorm (+SiLU) and LayerNorm over NHWC / token-major fp32 activations -> bf16 GEMM operands.
// Both are HBM-bound streaming kernels: 128-bit loads, fp32 statistics, one read of x per pass.
//
// GroupNorm is two launches so that both are fully parallel over pixels:
//   gn_stats : grid (chunks, NB): per-(image, chunk, group) partial sum / sum of squares
//   gn_apply : grid (chunks, NB): combine the partials of its image (double), normalise, affine,
//              optional SiLU, cast to bf16 (and optionally also emit a raw bf16 copy of x, the
//              operand of a ResnetBlock's 1x1 shortcut conv).
// Semantics: torch.nn.GroupNorm (biased variance) as used by diffusers ResnetBlock2D /
// Transformer2DModel / VAE blocks; SURVEY.md App. A.1-A.2.
#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace mgb {

constexpr int kGnThreads = 256;
constexpr int kGnMaxK = 4;      // channel-quads per thread
constexpr int kGnMaxChunks = 592;   // partial-statistics slots per image
constexpr int kGnMaxImages = 256;   // arrival counters

struct GnGeom {
  int Q;        // C / 4
  int Tq, Tp;   // thread grid: Tq channel-quad lanes x Tp pixel lanes (Tq * Tp <= 256)
  int Kq;       // Q / Tq  (<= kGnMaxK)
  int chunks, P;  // pixel chunks per image, pixels per chunk
};

static bool gn_geometry(int HW, int C, GnGeom* g) {
  if (C % 4) return false;
  g->Q = C / 4;
  int best = -1, bestTq = 0;
  for (int tq = 1; tq <= 256 && tq <= g->Q; ++tq) {
    if (g->Q % tq) continue;
    if (g->Q / tq > kGnMaxK) continue;
    const int tp = 256 / tq;
    if (tq * tp > best) { best = tq * tp; bestTq = tq; }
  }
  if (best < 0) return false;
  g->Tq = bestTq;
  g->Tp = 256 / bestTq;
  g->Kq = g->Q / bestTq;
  // ~8K elements (32 KB fp32) per chunk: enough CTAs to saturate HBM on the big VAE tensors, few enough
  // partials that combining them stays negligible on the small UNet ones
  long long want = ((long long)HW * C + 8191) / 8192;
  if (want < 1) want = 1;
  if (want > kGnMaxChunks) want = kGnMaxChunks;
  if (want > HW) want = HW;
  g->chunks = int(want);
  g->P = (HW + g->chunks - 1) / g->chunks;
  g->chunks = (HW + g->P - 1) / g->P;
  return true;
}

size_t groupnorm_ws_bytes(int NB, int HW, int C, int G) {
  (void)HW; (void)C;
  // [arrival counters: kGnMaxImages, FIXED offset 0 so that they stay valid (zero) across calls with
  //  different NB][mean|rstd NB x 2G][partials NB x kGnMaxChunks x G x 2]
  return (size_t(kGnMaxImages) + size_t(NB) * 2 * G + size_t(NB) * kGnMaxChunks * G * 2) * sizeof(float);
}

__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const float* __restrict__ x, float* __restrict__ ws,
                                                              float* __restrict__ stat, unsigned* __restrict__ counters,
                                                              int HW, int C, int G, float eps, GnGeom g) {
  extern __shared__ float s_acc[];  // [2 * C]
  pdl_launch_dependents();
  pdl_wait();
  const int img = blockIdx.y, chunk = blockIdx.x;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int tq = threadIdx.x % g.Tq, tp = threadIdx.x / g.Tq;
  if (tp < g.Tp) {
    float sum[kGnMaxK][4], sq[kGnMaxK][4];
#pragma unroll
    for (int k = 0; k < kGnMaxK; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) { sum[k][j] = 0.f; sq[k][j] = 0.f; }
    const int p0 = chunk * g.P, p1 = min(HW, p0 + g.P);
    const float4* xi = reinterpret_cast<const float4*>(x + (size_t)img * HW * C);
#pragma unroll
    for (int k = 0; k < kGnMaxK; ++k) {
      if (k >= g.Kq) break;
      const size_t qoff = size_t(tq) + size_t(k) * g.Tq;
      int p = p0 + tp;
      for (; p + 3 * g.Tp < p1; p += 4 * g.Tp) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(xi + (size_t)(p + u * g.Tp) * g.Q + qoff);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          sum[k][0] += v[u].x; sq[k][0] += v[u].x * v[u].x;
          sum[k][1] += v[u].y; sq[k][1] += v[u].y * v[u].y;
          sum[k][2] += v[u].z; sq[k][2] += v[u].z * v[u].z;
          sum[k][3] += v[u].w; sq[k][3] += v[u].w * v[u].w;
        }
      }
      for (; p < p1; p += g.Tp) {
        const float4 v = __ldg(xi + (size_t)p * g.Q + qoff);
        sum[k][0] += v.x; sq[k][0] += v.x * v.x;
        sum[k][1] += v.y; sq[k][1] += v.y * v.y;
        sum[k][2] += v.z; sq[k][2] += v.z * v.z;
        sum[k][3] += v.w; sq[k][3] += v.w * v.w;
      }
    }
#pragma unroll
    for (int k = 0; k < kGnMaxK; ++k) {
      if (k < g.Kq) {
        const int c0 = 4 * (tq + k * g.Tq);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          atomicAdd(&s_acc[c0 + j], sum[k][j]);
          atomicAdd(&s_acc[C + c0 + j], sq[k][j]);
        }
      }
    }
  }
  __syncthreads();
  const int cpg = C / G;
  for (int gi = threadIdx.x; gi < G; gi += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int c = gi * cpg; c < (gi + 1) * cpg; ++c) { s += s_acc[c]; q += s_acc[C + c]; }
    float* dst = ws + (((size_t)img * kGnMaxChunks + chunk) * G + gi) * 2;
    dst[0] = s; dst[1] = q;
  }
  // last CTA of this image combines the partials into mean / rstd (so gn_apply starts with 2G floats)
  __shared__ bool s_last;
  __shared__ double s_part[2][kGnThreads];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counters + img, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    const int slices = kGnThreads / G;
    const int gi = threadIdx.x % G, sl = threadIdx.x / G;
    double s = 0.0, q = 0.0;
    if (sl < slices) {
#pragma unroll 4
      for (int ch = sl; ch < int(gridDim.x); ch += slices) {
        const float2 v = __ldcg(reinterpret_cast<const float2*>(ws + (((size_t)img * kGnMaxChunks + ch) * G + gi) * 2));
        s += double(v.x); q += double(v.y);
      }
    }
    s_part[0][threadIdx.x] = s; s_part[1][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.x < G) {
      s = 0.0; q = 0.0;
      for (int k = 0; k < slices; ++k) { s += s_part[0][k * G + threadIdx.x]; q += s_part[1][k * G + threadIdx.x]; }
      const double n = double(HW) * cpg;
      const double mean = s / n;
      double var = q / n - mean * mean;
      if (var < 0.0) var = 0.0;
      stat[(size_t)img * 2 * G + threadIdx.x] = float(mean);
      stat[(size_t)img * 2 * G + G + threadIdx.x] = rsqrtf(float(var) + eps);
    }
    if (threadIdx.x == 0) counters[img] = 0;   // self-resetting
  }
}

__global__ void __launch_bounds__(kGnThreads)
    gn_apply_kernel(const float* __restrict__ x, bf16* __restrict__ y, bf16* __restrict__ raw,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stat,
                    int HW, int C, int G, int silu, GnGeom g) {
  extern __shared__ float s_stat[];  // mean[G], rstd[G]
  pdl_launch_dependents();
  pdl_wait();
  const int img = blockIdx.y, chunk = blockIdx.x;
  const int cpg = C / G;
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) s_stat[i] = __ldcg(stat + (size_t)img * 2 * G + i);
  __syncthreads();
  const int tq = threadIdx.x % g.Tq, tp = threadIdx.x / g.Tq;
  if (tp >= g.Tp) return;
  const int p0 = chunk * g.P, p1 = min(HW, p0 + g.P);
  const float4* xi = reinterpret_cast<const float4*>(x + (size_t)img * HW * C);
  uint2* yo = reinterpret_cast<uint2*>(y + (size_t)img * HW * C);
  uint2* ro = raw ? reinterpret_cast<uint2*>(raw + (size_t)img * HW * C) : nullptr;
  float sc[kGnMaxK][4], sh[kGnMaxK][4];
#pragma unroll
  for (int k = 0; k < kGnMaxK; ++k) {
    if (k < g.Kq) {
      const int c0 = 4 * (tq + k * g.Tq);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j, gi = c / cpg;
        const float ga = gamma ? __ldg(gamma + c) : 1.f, be = beta ? __ldg(beta + c) : 0.f;
        sc[k][j] = s_stat[G + gi] * ga;
        sh[k][j] = be - s_stat[gi] * s_stat[G + gi] * ga;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kGnMaxK; ++k) {
    if (k >= g.Kq) break;
    const float s0 = sc[k][0], s1 = sc[k][1], s2 = sc[k][2], s3 = sc[k][3];
    const float h0 = sh[k][0], h1 = sh[k][1], h2 = sh[k][2], h3 = sh[k][3];
    const size_t qoff = size_t(tq) + size_t(k) * g.Tq;
    int p = p0 + tp;
    // 4 pixels per trip: the loads are issued back to back, then consumed
    for (; p + 3 * g.Tp < p1; p += 4 * g.Tp) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(xi + (size_t)(p + u * g.Tp) * g.Q + qoff);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t idx = (size_t)(p + u * g.Tp) * g.Q + qoff;
        float o0 = v[u].x * s0 + h0, o1 = v[u].y * s1 + h1, o2 = v[u].z * s2 + h2, o3 = v[u].w * s3 + h3;
        if (silu) { o0 = silu_f(o0); o1 = silu_f(o1); o2 = silu_f(o2); o3 = silu_f(o3); }
        yo[idx] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
        if (ro) ro[idx] = make_uint2(pack_bf16x2(v[u].x, v[u].y), pack_bf16x2(v[u].z, v[u].w));
      }
    }
    for (; p < p1; p += g.Tp) {
      const size_t idx = (size_t)p * g.Q + qoff;
      const float4 v = __ldg(xi + idx);
      float o0 = v.x * s0 + h0, o1 = v.y * s1 + h1, o2 = v.z * s2 + h2, o3 = v.w * s3 + h3;
      if (silu) { o0 = silu_f(o0); o1 = silu_f(o1); o2 = silu_f(o2); o3 = silu_f(o3); }
      yo[idx] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
      if (ro) ro[idx] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
  }
}

int launch_groupnorm(const float* x, bf16* y, bf16* raw_copy, const float* gamma, const float* beta, float* ws,
                     int NB, int HW, int C, int G, float eps, int silu, cudaStream_t stream) {
  GnGeom g;
  if (C % G != 0 || G > kGnThreads || kGnThreads % G != 0 || NB > kGnMaxImages || !gn_geometry(HW, C, &g)) {
    set_error("groupnorm: unsupported C=%d G=%d", C, G);
    return MGB_ERR_INVALID;
  }
  dim3 grid(g.chunks, NB);
  unsigned* counters = reinterpret_cast<unsigned*>(ws);
  float* stat = ws + kGnMaxImages;
  float* partials = stat + size_t(NB) * 2 * G;
  launch_k(gn_stats_kernel, grid, kGnThreads, 2 * C * sizeof(float), stream, x, partials, stat, counters, HW, C, G, eps, g);
  launch_k(gn_apply_kernel, grid, kGnThreads, 2 * G * sizeof(float), stream, x, y, raw_copy, gamma, beta,
           static_cast<const float*>(stat), HW, C, G, silu, g);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("groupnorm launch: %s", cudaGetErrorString(e));
    return MGB_ERR_CUDA;
  }
  return MGB_OK;
}

// -------------------------------------------------------------------------------------------------
// LayerNorm: one warp per token, values held in registers (C <= 1280 -> <= 10 float4 per lane).
// Two-pass (mean, then centred variance) like torch.nn.LayerNorm.
// -------------------------------------------------------------------------------------------------
constexpr int kLnMaxQ = 10;
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, bf16* __restrict__ y,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int M, int C, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int Q = C / 4;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)warp * C);
  float4 v[kLnMaxQ];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kLnMaxQ; ++k) {
    const int q = lane + 32 * k;
    if (q < Q) {
      v[k] = __ldg(xr + q);
      s += v[k].x + v[k].y + v[k].z + v[k].w;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < kLnMaxQ; ++k) {
    const int q = lane + 32 * k;
    if (q < Q) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
      ss += a * a + b * b + c * c + d * d;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / C + eps);
  uint2* yr = reinterpret_cast<uint2*>(y + (size_t)warp * C);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int k = 0; k < kLnMaxQ; ++k) {
    const int q = lane + 32 * k;
    if (q < Q) {
      const float4 ga = __ldg(g4 + q), be = __ldg(b4 + q);
      const float o0 = (v[k].x - mean) * rstd * ga.x + be.x, o1 = (v[k].y - mean) * rstd * ga.y + be.y;
      const float o2 = (v[k].z - mean) * rstd * ga.z + be.z, o3 = (v[k].w - mean) * rstd * ga.w + be.w;
      yr[q] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
    }
  }
}

int launch_layernorm(const float* x, bf16* y, const float* gamma, const float* beta, int M, int C, float eps,
                     cudaStream_t stream) {
  if (C % 4 != 0 || C / 4 > 32 * kLnMaxQ) {
    set_error("layernorm: unsupported C=%d", C);
    return MGB_ERR_INVALID;
  }
  const int warps_per_block = 8;
  const int blocks = (M + warps_per_block - 1) / warps_per_block;
  launch_k(layernorm_kernel, blocks, 256, 0, stream, x, y, gamma, beta, M, C, eps);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("layernorm launch: %s", cudaGetErrorString(e));
    return MGB_ERR_CUDA;
  }
  return MGB_OK;
}

}  // namespace mgb
