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
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace mgb {

constexpr int kGnThreads = 256;
constexpr int kGnLoads = 8;          // float4 loads in flight per thread and round
constexpr int kGnMaxChunks = 1184;   // CTAs per image (8 per SM)

// Thread geometry: Tq lanes along channel quads x Tp lanes along pixels; a thread owns Kq quads (Kq in {1, 2, 4}) and
// R = 8 / Kq pixels per round, so that ALL of a round's loads are issued before anything is consumed. These kernels
// run between two GEMMs on tensors that mostly sit in L2: they are bound by dependent load latency, not bandwidth
// (the first version walked pixels with 2-3 dependent round trips per quad and took 9-15 us on 0.7-12 MB).
struct GnGeom {
  int Q;          // C / 4
  int Tq, Tp;     // Tq * Tp <= 256
  int Kq, R;      // quads per thread, pixels per thread and round (Kq * R == kGnLoads)
  int chunks, P;  // pixel chunks per image, pixels per chunk (a multiple of Tp * R)
};

static bool gn_geometry(int HW, int C, GnGeom* g, int max_chunks = kGnMaxChunks) {
  if (C % 4) return false;
  g->Q = C / 4;
  int best = -1;
  for (int kq = 1; kq <= 4; kq *= 2) {
    if (g->Q % kq) continue;
    const int tq = g->Q / kq;
    if (tq > kGnThreads) continue;
    const int tp = kGnThreads / tq;
    if (tq * tp > best) { best = tq * tp; g->Tq = tq; g->Tp = tp; g->Kq = kq; }
  }
  if (best < 0) return false;
  g->R = kGnLoads / g->Kq;
  const int per_round = g->Tp * g->R;
  long long rounds_total = (HW + per_round - 1) / per_round;
  long long rounds = (rounds_total + max_chunks - 1) / max_chunks;
  g->P = int(rounds) * per_round;
  g->chunks = (HW + g->P - 1) / g->P;
  return true;
}

size_t groupnorm_ws_bytes(int NB, int HW, int C, int G) {
  (void)HW; (void)G;
  return size_t(NB) * C * 2 * sizeof(float);   // per-channel (sum, sum of squares)
}

// -------------------------------------------------------------------------------------------------
// Per-channel statistics of x [NB, HW, C]: cs[(img * C + c) * 2 + {0,1}] += (sum, sum of squares).
// cs must be zero on entry (the network zeroes its whole statistics slab once per forward).
// -------------------------------------------------------------------------------------------------
template <int KQ>
__global__ void __launch_bounds__(kGnThreads) chan_stats_kernel(const float* __restrict__ x, float* __restrict__ cs,
                                                                int HW, int C, GnGeom g) {
  constexpr int R = kGnLoads / KQ;
  extern __shared__ float s_acc[];  // [2 * C]
  pdl_launch_dependents();
  const int img = blockIdx.y, chunk = blockIdx.x;
  const bool use_smem = g.Tp > 1;
  if (use_smem) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
  }
  pdl_wait();
  const int tq = threadIdx.x % g.Tq, tp = threadIdx.x / g.Tq;
  const bool active = tp < g.Tp;
  float sum[KQ][4], sq[KQ][4];
#pragma unroll
  for (int k = 0; k < KQ; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) { sum[k][j] = 0.f; sq[k][j] = 0.f; }
  const int p0 = chunk * g.P, p1 = min(HW, p0 + g.P);
  const float4* xi = reinterpret_cast<const float4*>(x + (size_t)img * HW * C);
  for (int pb = p0 + tp; pb < p1; pb += g.Tp * R) {
    float4 v[KQ][R];
#pragma unroll
    for (int k = 0; k < KQ; ++k)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int p = pb + r * g.Tp;
        v[k][r] = (active && p < p1) ? __ldg(xi + (size_t)p * g.Q + tq + k * g.Tq) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int k = 0; k < KQ; ++k)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        sum[k][0] += v[k][r].x; sq[k][0] = fmaf(v[k][r].x, v[k][r].x, sq[k][0]);
        sum[k][1] += v[k][r].y; sq[k][1] = fmaf(v[k][r].y, v[k][r].y, sq[k][1]);
        sum[k][2] += v[k][r].z; sq[k][2] = fmaf(v[k][r].z, v[k][r].z, sq[k][2]);
        sum[k][3] += v[k][r].w; sq[k][3] = fmaf(v[k][r].w, v[k][r].w, sq[k][3]);
      }
  }
  if (!use_smem) {
    if (active) {
#pragma unroll
      for (int k = 0; k < KQ; ++k) {
        float* dst = cs + ((size_t)img * C + 4 * (tq + k * g.Tq)) * 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) { atomicAdd(dst + 2 * j, sum[k][j]); atomicAdd(dst + 2 * j + 1, sq[k][j]); }
      }
    }
    return;
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < KQ; ++k) {
      const int c0 = 4 * (tq + k * g.Tq);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&s_acc[2 * (c0 + j)], sum[k][j]);
        atomicAdd(&s_acc[2 * (c0 + j) + 1], sq[k][j]);
      }
    }
  }
  __syncthreads();
  float* dst = cs + (size_t)img * C * 2;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(dst + i, s_acc[i]);
}

// -------------------------------------------------------------------------------------------------
// GroupNorm apply over the channel concatenation [a | b] (b optional): group statistics come from the
// per-channel sums of each source; y = act((x - mean) * rstd * gamma + beta) as bf16 [NB, HW, Ca + Cb];
// optionally also the raw bf16 copy of [a | b] (operand of a ResnetBlock's 1x1 shortcut conv).
// This is torch.cat(dim=1) + GroupNorm (+SiLU) of diffusers' up-block resnets in one pass.
// -------------------------------------------------------------------------------------------------
template <int KQ>
__global__ void __launch_bounds__(kGnThreads)
    gn_apply2_kernel(const float* __restrict__ xa, const float* __restrict__ csa, int Ca, const float* __restrict__ xb,
                     const float* __restrict__ csb, int Cb, bf16* __restrict__ y, bf16* __restrict__ raw,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int HW, int G, float eps, int silu,
                     GnGeom g) {
  constexpr int R = kGnLoads / KQ;
  extern __shared__ float s_stat[];  // mean[G], rstd[G]
  pdl_launch_dependents();
  pdl_wait();
  const int img = blockIdx.y, chunk = blockIdx.x;
  const int C = Ca + Cb, cpg = C / G;
  const int tq = threadIdx.x % g.Tq, tp = threadIdx.x / g.Tq;
  const bool active = tp < g.Tp;
  const int p0 = chunk * g.P, p1 = min(HW, p0 + g.P);
  const int Qa = Ca / 4, Qb = Cb / 4;
  const float4* xai = reinterpret_cast<const float4*>(xa + (size_t)img * HW * Ca);
  const float4* xbi = xb ? reinterpret_cast<const float4*>(xb + (size_t)img * HW * Cb) : nullptr;
  uint2* yo = reinterpret_cast<uint2*>(y + (size_t)img * HW * C);
  uint2* ro = raw ? reinterpret_cast<uint2*>(raw + (size_t)img * HW * C) : nullptr;

  // (1) first round of pixel loads + the affine parameters: in flight while the statistics are reduced
  const float4* src[KQ];
  size_t sstride[KQ];
  float4 ga4[KQ], be4[KQ];
  float4 v[KQ][R];
#pragma unroll
  for (int k = 0; k < KQ; ++k) {
    const int qd = tq + k * g.Tq;   // quad index in the concatenated channel space
    const bool from_a = qd < Qa;
    src[k] = from_a ? xai + qd : xbi + (qd - Qa);
    sstride[k] = from_a ? size_t(Qa) : size_t(Qb);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int p = p0 + tp + r * g.Tp;
      v[k][r] = (active && p < p1) ? __ldg(src[k] + (size_t)p * sstride[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ga4[k] = gamma ? __ldg(reinterpret_cast<const float4*>(gamma) + qd) : make_float4(1.f, 1.f, 1.f, 1.f);
    be4[k] = beta ? __ldg(reinterpret_cast<const float4*>(beta) + qd) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // (2) group statistics: 8 lanes per group (G * 8 <= 256 threads), up to 4 independent loads per lane and pass
  {
    const int gi = threadIdx.x >> 3, part = threadIdx.x & 7;
    double s = 0.0, q = 0.0;
    if (gi < G) {
      for (int j0 = part; j0 < cpg; j0 += 32) {
        float2 t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + 8 * u, c = gi * cpg + j;
          t[u] = make_float2(0.f, 0.f);
          if (j < cpg)
            t[u] = c < Ca ? __ldcg(reinterpret_cast<const float2*>(csa + ((size_t)img * Ca + c) * 2))
                          : __ldcg(reinterpret_cast<const float2*>(csb + ((size_t)img * Cb + (c - Ca)) * 2));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { s += double(t[u].x); q += double(t[u].y); }
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
  if (!active) return;
  float sc[KQ][4], sh[KQ][4];
#pragma unroll
  for (int k = 0; k < KQ; ++k) {
    const int qd = tq + k * g.Tq;
    const float gav[4] = {ga4[k].x, ga4[k].y, ga4[k].z, ga4[k].w}, bev[4] = {be4[k].x, be4[k].y, be4[k].z, be4[k].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gi = (4 * qd + j) / cpg;
      const float rstd = s_stat[G + gi];
      sc[k][j] = rstd * gav[j];
      sh[k][j] = bev[j] - s_stat[gi] * rstd * gav[j];
    }
  }
  // (3) apply; further rounds (only tensors too large for one round per CTA) reload in the same batched way
  for (int pb = p0 + tp;;) {
#pragma unroll
    for (int k = 0; k < KQ; ++k)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int p = pb + r * g.Tp;
        if (p < p1) {
          const size_t idx = (size_t)p * g.Q + tq + k * g.Tq;
          const float4 t = v[k][r];
          float o0 = fmaf(t.x, sc[k][0], sh[k][0]), o1 = fmaf(t.y, sc[k][1], sh[k][1]),
                o2 = fmaf(t.z, sc[k][2], sh[k][2]), o3 = fmaf(t.w, sc[k][3], sh[k][3]);
          if (silu) { o0 = silu_f(o0); o1 = silu_f(o1); o2 = silu_f(o2); o3 = silu_f(o3); }
          yo[idx] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
          if (ro) ro[idx] = make_uint2(pack_bf16x2(t.x, t.y), pack_bf16x2(t.z, t.w));
        }
      }
    pb += g.Tp * R;
    if (pb >= p1) break;
#pragma unroll
    for (int k = 0; k < KQ; ++k)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int p = pb + r * g.Tp;
        v[k][r] = p < p1 ? __ldg(src[k] + (size_t)p * sstride[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
  }
}

int launch_chan_stats(const float* x, float* cs, int NB, int HW, int C, cudaStream_t stream) {
  GnGeom g;
  // every CTA ends with 2*C same-address global REDs, which serialise in L2: fewer, longer CTAs than the apply pass
  static const int stat_chunks = getenv("MGB_GN_STAT_CHUNKS") ? atoi(getenv("MGB_GN_STAT_CHUNKS")) : kGnMaxChunks;   // 148 / 296 / 1184 measured identical (r01)
  if (!gn_geometry(HW, C, &g, std::max(1, stat_chunks / std::max(1, NB)))) { set_error("chan_stats: unsupported C=%d", C); return MGB_ERR_INVALID; }
  dim3 grid(g.chunks, NB);
  const size_t smem = g.Tp > 1 ? 2 * C * sizeof(float) : 0;
  cudaError_t e;
  if (g.Kq == 1) e = launch_k(chan_stats_kernel<1>, grid, kGnThreads, smem, stream, x, cs, HW, C, g);
  else if (g.Kq == 2) e = launch_k(chan_stats_kernel<2>, grid, kGnThreads, smem, stream, x, cs, HW, C, g);
  else e = launch_k(chan_stats_kernel<4>, grid, kGnThreads, smem, stream, x, cs, HW, C, g);
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
  const size_t smem = 2 * G * sizeof(float);
  cudaError_t e;
  if (g.Kq == 1)
    e = launch_k(gn_apply2_kernel<1>, grid, kGnThreads, smem, stream, xa, csa, Ca, xb, csb, Cb, y, raw_copy, gamma, beta, HW,
                 G, eps, silu, g);
  else if (g.Kq == 2)
    e = launch_k(gn_apply2_kernel<2>, grid, kGnThreads, smem, stream, xa, csa, Ca, xb, csb, Cb, y, raw_copy, gamma, beta, HW,
                 G, eps, silu, g);
  else
    e = launch_k(gn_apply2_kernel<4>, grid, kGnThreads, smem, stream, xa, csa, Ca, xb, csb, Cb, y, raw_copy, gamma, beta, HW,
                 G, eps, silu, g);
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
