// Run-to-run deterministic GroupNorm statistics (opt-in: MGB_GN_DETERMINISTIC=1; see net.cu::groupnorm).
//
// Same kernels as norm.cu, but the per-channel (sum, sum of squares) are accumulated as 64-bit FIXED POINT
// (value * 2^30, two's complement through unsigned atomics): integer addition is associative, so the result does
// not depend on the order in which CTAs reach the atomics, and with it the whole denoising path becomes bit-
// reproducible. A thread's own partial sums stay fp32 in a fixed order; only the cross-thread / cross-CTA
// combination is integer. Range: |sum| < 2^33 = 8.6e9 (a 768 x 768 VAE plane of |x| ~ 100 reaches 5.9e9);
// resolution 2^-30 = 9.3e-10 per contribution.
// STATUS (round 1): compiled, NOT yet validated on the GPU (the round's GPU budget was spent); off by default.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "launch.h"
#include "norm_geom.h"

namespace mgb {

constexpr double kFxScale = 1073741824.0;        // 2^30
constexpr double kFxInv = 1.0 / kFxScale;
__device__ __forceinline__ unsigned long long to_fx(float v) {
  return static_cast<unsigned long long>(__double2ll_rn(double(v) * kFxScale));
}

// -------------------------------------------------------------------------------------------------
// Per-channel statistics of x [NB, HW, C]: cs[(img * C + c) * 2 + {0,1}] += (sum, sum of squares).
// cs must be zero on entry (the network zeroes its whole statistics slab once per forward).
// -------------------------------------------------------------------------------------------------
template <int KQ>
__global__ void __launch_bounds__(kGnThreads) chan_stats_fx_kernel(const float* __restrict__ x, long long* __restrict__ cs,
                                                                int HW, int C, GnGeom g) {
  constexpr int R = kGnLoads / KQ;
  extern __shared__ unsigned long long s_acc[];  // [2 * C] fixed point
  pdl_launch_dependents();
  const int img = blockIdx.y, chunk = blockIdx.x;
  const bool use_smem = g.Tp > 1;
  if (use_smem) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0ull;
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
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(cs) + ((size_t)img * C + 4 * (tq + k * g.Tq)) * 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) { atomicAdd(dst + 2 * j, to_fx(sum[k][j])); atomicAdd(dst + 2 * j + 1, to_fx(sq[k][j])); }
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
        atomicAdd(&s_acc[2 * (c0 + j)], to_fx(sum[k][j]));
        atomicAdd(&s_acc[2 * (c0 + j) + 1], to_fx(sq[k][j]));
      }
    }
  }
  __syncthreads();
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(cs) + (size_t)img * C * 2;
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
    gn_apply2_fx_kernel(const float* __restrict__ xa, const long long* __restrict__ csa, int Ca, const float* __restrict__ xb,
                     const long long* __restrict__ csb, int Cb, bf16* __restrict__ y, bf16* __restrict__ raw,
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
    long long si = 0, qi = 0;
    if (gi < G) {
      for (int j0 = part; j0 < cpg; j0 += 32) {
        longlong2 t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + 8 * u, c = gi * cpg + j;
          t[u] = make_longlong2(0, 0);
          if (j < cpg)
            t[u] = c < Ca ? __ldcg(reinterpret_cast<const longlong2*>(csa + ((size_t)img * Ca + c) * 2))
                          : __ldcg(reinterpret_cast<const longlong2*>(csb + ((size_t)img * Cb + (c - Ca)) * 2));
        }
        // integer accumulation: the group sums are exact and independent of any order
#pragma unroll
        for (int u = 0; u < 4; ++u) { si += t[u].x; qi += t[u].y; }
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      si += __shfl_xor_sync(0xffffffffu, si, o);
      qi += __shfl_xor_sync(0xffffffffu, qi, o);
    }
    const double s = double(si) * kFxInv, q = double(qi) * kFxInv;
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

int launch_chan_stats_fx(const float* x, long long* cs, int NB, int HW, int C, cudaStream_t stream) {
  GnGeom g;
  // every CTA ends with 2*C same-address global REDs, which serialise in L2: fewer, longer CTAs than the apply pass
  const int stat_chunks = kGnMaxChunks;
  if (!gn_geometry(HW, C, &g, std::max(1, stat_chunks / std::max(1, NB)))) { set_error("chan_stats_fx: unsupported C=%d", C); return MGB_ERR_INVALID; }
  dim3 grid(g.chunks, NB);
  const size_t smem = g.Tp > 1 ? 2 * C * sizeof(unsigned long long) : 0;
  cudaError_t e;
  if (g.Kq == 1) e = launch_k(chan_stats_fx_kernel<1>, grid, kGnThreads, smem, stream, x, cs, HW, C, g);
  else if (g.Kq == 2) e = launch_k(chan_stats_fx_kernel<2>, grid, kGnThreads, smem, stream, x, cs, HW, C, g);
  else e = launch_k(chan_stats_fx_kernel<4>, grid, kGnThreads, smem, stream, x, cs, HW, C, g);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("chan_stats_fx launch: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  return MGB_OK;
}

int launch_gn_apply2_fx(const float* xa, const long long* csa, int Ca, const float* xb, const long long* csb, int Cb, bf16* y,
                     bf16* raw_copy, const float* gamma, const float* beta, int NB, int HW, int G, float eps, int silu,
                     cudaStream_t stream) {
  GnGeom g;
  const int C = Ca + Cb;
  if (C % G != 0 || G * 8 > kGnThreads || (Ca & 3) || (Cb & 3) || !gn_geometry(HW, C, &g)) {
    set_error("groupnorm_fx: unsupported C=%d+%d G=%d", Ca, Cb, G);
    return MGB_ERR_INVALID;
  }
  dim3 grid(g.chunks, NB);
  const size_t smem = 2 * G * sizeof(float);
  cudaError_t e;
  if (g.Kq == 1)
    e = launch_k(gn_apply2_fx_kernel<1>, grid, kGnThreads, smem, stream, xa, csa, Ca, xb, csb, Cb, y, raw_copy, gamma, beta, HW,
                 G, eps, silu, g);
  else if (g.Kq == 2)
    e = launch_k(gn_apply2_fx_kernel<2>, grid, kGnThreads, smem, stream, xa, csa, Ca, xb, csb, Cb, y, raw_copy, gamma, beta, HW,
                 G, eps, silu, g);
  else
    e = launch_k(gn_apply2_fx_kernel<4>, grid, kGnThreads, smem, stream, xa, csa, Ca, xb, csb, Cb, y, raw_copy, gamma, beta, HW,
                 G, eps, silu, g);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("groupnorm_fx launch: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  return MGB_OK;
}

}  // namespace mgb
