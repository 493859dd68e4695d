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
constexpr int kGnMaxChunks = 592;   // CTAs per image

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
  (void)HW; (void)G;
  return size_t(NB) * C * 2 * sizeof(float);   // per-channel (sum, sum of squares)
}

// -------------------------------------------------------------------------------------------------
// Per-channel statistics of x [NB, HW, C]: cs[(img * C + c) * 2 + {0,1}] += (sum, sum of squares).
// Only used where the producer could not emit them from its epilogue (VAE-sized tensors, ragged token
// tiles); cs must be zero on entry.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGnThreads) chan_stats_kernel(const float* __restrict__ x, float* __restrict__ cs,
                                                                int HW, int C, GnGeom g) {
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
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(cs + ((size_t)img * C + c) * 2, s_acc[c]);
    atomicAdd(cs + ((size_t)img * C + c) * 2 + 1, s_acc[C + c]);
  }
}

// -------------------------------------------------------------------------------------------------
// GroupNorm apply over the channel concatenation [a | b] (b optional): group statistics come from the
// per-channel sums of each source; y = act((x - mean) * rstd * gamma + beta) as bf16 [NB, HW, Ca + Cb];
// optionally also the raw bf16 copy of [a | b] (operand of a ResnetBlock's 1x1 shortcut conv).
// This is torch.cat(dim=1) + GroupNorm (+SiLU) of diffusers' up-block resnets in one pass.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGnThreads)
    gn_apply2_kernel(const float* __restrict__ xa, const float* __restrict__ csa, int Ca, const float* __restrict__ xb,
                     const float* __restrict__ csb, int Cb, bf16* __restrict__ y, bf16* __restrict__ raw,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int HW, int G, float eps, int silu,
                     GnGeom g) {
  extern __shared__ float s_stat[];  // mean[G], rstd[G]
  pdl_launch_dependents();
  pdl_wait();
  const int img = blockIdx.y, chunk = blockIdx.x;
  const int C = Ca + Cb, cpg = C / G;
  {
    // 8 lanes per group, then a shuffle reduction (G * 8 <= 256 threads)
    const int gi = threadIdx.x >> 3, part = threadIdx.x & 7;
    double s = 0.0, q = 0.0;
    if (gi < G) {
      for (int j = part; j < cpg; j += 8) {
        const int c = gi * cpg + j;
        const float2 v = c < Ca ? __ldcg(reinterpret_cast<const float2*>(csa + ((size_t)img * Ca + c) * 2))
                                : __ldcg(reinterpret_cast<const float2*>(csb + ((size_t)img * Cb + (c - Ca)) * 2));
        s += double(v.x); q += double(v.y);
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (gi < G && part == 0) {
      const double n = double(HW) * cpg;
      const double mean = s / n;
      double var = q / n - mean * mean;
      if (var < 0.0) var = 0.0;
      s_stat[gi] = float(mean);
      s_stat[G + gi] = rsqrtf(float(var) + eps);
    }
  }
  __syncthreads();
  const int tq = threadIdx.x % g.Tq, tp = threadIdx.x / g.Tq;
  if (tp >= g.Tp) return;
  const int p0 = chunk * g.P, p1 = min(HW, p0 + g.P);
  const int Qa = Ca / 4, Qb = Cb / 4;
  const float4* xai = reinterpret_cast<const float4*>(xa + (size_t)img * HW * Ca);
  const float4* xbi = xb ? reinterpret_cast<const float4*>(xb + (size_t)img * HW * Cb) : nullptr;
  uint2* yo = reinterpret_cast<uint2*>(y + (size_t)img * HW * C);
  uint2* ro = raw ? reinterpret_cast<uint2*>(raw + (size_t)img * HW * C) : nullptr;
#pragma unroll
  for (int k = 0; k < kGnMaxK; ++k) {
    if (k >= g.Kq) break;
    const int qd = tq + k * g.Tq;                 // quad index in the concatenated channel space
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 4 * qd + j, gi = c / cpg;
      const float ga = gamma ? __ldg(gamma + c) : 1.f, be = beta ? __ldg(beta + c) : 0.f;
      sc[j] = s_stat[G + gi] * ga;
      sh[j] = be - s_stat[gi] * s_stat[G + gi] * ga;
    }
    const bool from_a = qd < Qa;
    const float4* src = from_a ? xai + qd : xbi + (qd - Qa);
    const size_t sstride = from_a ? size_t(Qa) : size_t(Qb);
    int p = p0 + tp;
    for (; p + 3 * g.Tp < p1; p += 4 * g.Tp) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(src + (size_t)(p + u * g.Tp) * sstride);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t idx = (size_t)(p + u * g.Tp) * g.Q + qd;
        float o0 = v[u].x * sc[0] + sh[0], o1 = v[u].y * sc[1] + sh[1], o2 = v[u].z * sc[2] + sh[2],
              o3 = v[u].w * sc[3] + sh[3];
        if (silu) { o0 = silu_f(o0); o1 = silu_f(o1); o2 = silu_f(o2); o3 = silu_f(o3); }
        yo[idx] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
        if (ro) ro[idx] = make_uint2(pack_bf16x2(v[u].x, v[u].y), pack_bf16x2(v[u].z, v[u].w));
      }
    }
    for (; p < p1; p += g.Tp) {
      const size_t idx = (size_t)p * g.Q + qd;
      const float4 v = __ldg(src + (size_t)p * sstride);
      float o0 = v.x * sc[0] + sh[0], o1 = v.y * sc[1] + sh[1], o2 = v.z * sc[2] + sh[2], o3 = v.w * sc[3] + sh[3];
      if (silu) { o0 = silu_f(o0); o1 = silu_f(o1); o2 = silu_f(o2); o3 = silu_f(o3); }
      yo[idx] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
      if (ro) ro[idx] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
  }
}

int launch_chan_stats(const float* x, float* cs, int NB, int HW, int C, cudaStream_t stream) {
  GnGeom g;
  if (!gn_geometry(HW, C, &g)) { set_error("chan_stats: unsupported C=%d", C); return MGB_ERR_INVALID; }
  dim3 grid(g.chunks, NB);
  cudaError_t e = launch_k(chan_stats_kernel, grid, kGnThreads, 2 * C * sizeof(float), stream, x, cs, HW, C, g);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("chan_stats launch: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  return MGB_OK;
}

int launch_gn_apply2(const float* xa, const float* csa, int Ca, const float* xb, const float* csb, int Cb, bf16* y,
                     bf16* raw_copy, const float* gamma, const float* beta, int NB, int HW, int G, float eps, int silu,
                     cudaStream_t stream) {
  GnGeom g;
  const int C = Ca + Cb;
  if (C % G != 0 || G * 8 > kGnThreads || (Ca & 3) || (Cb & 3) || !gn_geometry(HW, C, &g)) {
    set_error("groupnorm: unsupported C=%d+%d G=%d", Ca, Cb, G);
    return MGB_ERR_INVALID;
  }
  dim3 grid(g.chunks, NB);
  cudaError_t e = launch_k(gn_apply2_kernel, grid, kGnThreads, 2 * G * sizeof(float), stream, xa, csa, Ca, xb, csb, Cb, y,
                           raw_copy, gamma, beta, HW, G, eps, silu, g);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("groupnorm launch: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  return MGB_OK;
}

// Stand-alone GroupNorm (operator-level ABI): ws = per-channel stats scratch [NB, C, 2] (zeroed here).
int launch_groupnorm(const float* x, bf16* y, bf16* raw_copy, const float* gamma, const float* beta, float* ws,
                     int NB, int HW, int C, int G, float eps, int silu, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(ws, 0, size_t(NB) * C * 2 * sizeof(float), stream);
  if (e != cudaSuccess) { set_error("groupnorm memset: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  int rc = launch_chan_stats(x, ws, NB, HW, C, stream);
  if (rc) return rc;
  return launch_gn_apply2(x, ws, C, nullptr, nullptr, 0, y, raw_copy, gamma, beta, NB, HW, G, eps, silu, stream);
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
