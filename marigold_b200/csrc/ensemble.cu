// Test-time ensembling on the device (reference marigold/util/ensemble.py).
//
//  * ens_depth_cost   : the BFGS objective of ensemble_depth (ensemble.py:138-152) in one pass over
//                       the E aligned maps and ONE host synchronisation (the reference does C(E,2)+2
//                       `.item()` syncs per evaluation).
//  * ens_depth_reduce : align (ensemble.py:107-118) + median/mean (+MAD/std) (:120-136) + min-max
//                       renormalisation (:184-194), plus the index of the member the lower median picks.
//  * ens_normals      : ensemble_normals (:199-249): mean -> normalise -> cosine -> clamp -> argmax -> gather.
//
// These are HBM-bound streaming kernels: each reads the E maps exactly once (E*4 bytes / pixel).
// Arithmetic that decides an index (median / argmax) uses explicitly un-fused fp32 ops
// (__fmul_rn/__fadd_rn) in the reference's operation order so that ties break identically.
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace mgb {

constexpr int kEnsThreads = 256;
constexpr int kEnsMaxBlocks = 148 * 4;
constexpr int kEnsMaxE = 16;

__device__ __forceinline__ float align1(float d, float s, float t, int shift) {
  // torch: depth * s + t  (two roundings; no FMA)
  const float m = __fmul_rn(d, s);
  return shift ? __fadd_rn(m, t) : m;
}

template <int E>
__device__ __forceinline__ void sort_small(float (&v)[E], int (&idx)[E]) {
#pragma unroll
  for (int i = 1; i < E; ++i) {
#pragma unroll
    for (int j = i; j > 0; --j) {
      // stable: only swap on strict greater, so equal values keep ascending member index
      if (v[j - 1] > v[j]) {
        const float tv = v[j]; v[j] = v[j - 1]; v[j - 1] = tv;
        const int ti = idx[j]; idx[j] = idx[j - 1]; idx[j - 1] = ti;
      }
    }
  }
}

struct CostPartial {
  double pair_sum[kEnsMaxE * (kEnsMaxE - 1) / 2];
  float pmin, pmax;
};

template <int E>
__global__ void __launch_bounds__(kEnsThreads)
    ens_cost_kernel(const float* __restrict__ depth, const float* __restrict__ st, long long HW, int shift, int median,
                    CostPartial* __restrict__ partials) {
  constexpr int NP = E * (E - 1) / 2;
  float s[E], t[E];
#pragma unroll
  for (int e = 0; e < E; ++e) { s[e] = st[e]; t[e] = st[E + e]; }
  float acc[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) acc[k] = 0.f;
  float pmin = FLT_MAX, pmax = -FLT_MAX;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    float a[E];
#pragma unroll
    for (int e = 0; e < E; ++e) a[e] = align1(__ldg(depth + (long long)e * HW + p), s[e], t[e], shift);
    int k = 0;
#pragma unroll
    for (int i = 0; i < E; ++i)
#pragma unroll
      for (int j = i + 1; j < E; ++j) {
        const float d = a[i] - a[j];
        acc[k] = fmaf(d, d, acc[k]);
        ++k;
      }
    float pred;
    if (median) {
      int idx[E];
#pragma unroll
      for (int e = 0; e < E; ++e) idx[e] = e;
      sort_small<E>(a, idx);
      pred = a[(E - 1) / 2];
    } else {
      float sm = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) sm += a[e];
      pred = sm / float(E);
    }
    pmin = fminf(pmin, pred);
    pmax = fmaxf(pmax, pred);
  }
  // block reduction (double for the pair sums)
  __shared__ double sh[kEnsThreads / 32];
  __shared__ float shf[2][kEnsThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  CostPartial* out = partials + blockIdx.x;
#pragma unroll 1
  for (int k = 0; k < NP; ++k) {
    double v = double(acc[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < kEnsThreads / 32; ++w) tot += sh[w];
      out->pair_sum[k] = tot;
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pmin = fminf(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
    pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
  }
  if (lane == 0) { shf[0][warp] = pmin; shf[1][warp] = pmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kEnsThreads / 32; ++w) { pmin = fminf(pmin, shf[0][w]); pmax = fmaxf(pmax, shf[1][w]); }
    out->pmin = pmin; out->pmax = pmax;
  }
}

__global__ void ens_cost_final_kernel(const CostPartial* __restrict__ partials, int nblocks, int E, long long HW,
                                      double reg, double* __restrict__ out) {
  // single block; thread k owns pair k
  const int NP = E * (E - 1) / 2;
  __shared__ double sh[128];
  double c = 0.0;
  for (int k = threadIdx.x; k < NP; k += blockDim.x) {
    double tot = 0.0;
    for (int b = 0; b < nblocks; ++b) tot += partials[b].pair_sum[k];
    // reference: (diff**2).mean().sqrt() evaluated in fp32
    c += double(sqrtf(float(tot / double(HW))));
  }
  sh[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    double cost = 0.0;
    for (int i = 0; i < int(blockDim.x); ++i) cost += sh[i];
    float pmin = FLT_MAX, pmax = -FLT_MAX;
    for (int b = 0; b < nblocks; ++b) { pmin = fminf(pmin, partials[b].pmin); pmax = fmaxf(pmax, partials[b].pmax); }
    if (reg > 0.0) cost += (double(fabsf(0.0f - pmin)) + double(fabsf(1.0f - pmax))) * reg;
    out[0] = cost;
    out[1] = double(pmin);
    out[2] = double(pmax);
  }
}

template <int E>
static void launch_cost_t(const float* depth, const float* st, long long HW, int shift, int median,
                          CostPartial* partials, int blocks, cudaStream_t stream) {
  ens_cost_kernel<E><<<blocks, kEnsThreads, 0, stream>>>(depth, st, HW, shift, median, partials);
}

size_t ens_ws_bytes() { return sizeof(CostPartial) * kEnsMaxBlocks + 64 * sizeof(float) + 64; }

// ws layout: [CostPartial x kEnsMaxBlocks][st: 2*kEnsMaxE floats][out: 4 doubles]
int launch_ens_depth_cost(const float* depth, const float* st_host, int E, long long HW, int shift, int median,
                          double reg, void* ws, double* out_host_pinned, cudaStream_t stream) {
  if (E < 2 || E > kEnsMaxE) { set_error("ensemble size %d outside [2, %d]", E, kEnsMaxE); return MGB_ERR_UNSUPPORTED; }
  CostPartial* partials = reinterpret_cast<CostPartial*>(ws);
  float* st = reinterpret_cast<float*>(partials + kEnsMaxBlocks);
  double* out = reinterpret_cast<double*>(st + 2 * kEnsMaxE + 2);
  cudaError_t e = cudaMemcpyAsync(st, st_host, sizeof(float) * 2 * E, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) { set_error("ens cost H2D: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  const int blocks = int(std::min<long long>((HW + kEnsThreads - 1) / kEnsThreads, kEnsMaxBlocks));
  switch (E) {
#define CASE(n) case n: launch_cost_t<n>(depth, st, HW, shift, median, partials, blocks, stream); break;
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
    CASE(15) CASE(16)
#undef CASE
  }
  ens_cost_final_kernel<<<1, 128, 0, stream>>>(partials, blocks, E, HW, reg, out);
  e = cudaMemcpyAsync(out_host_pinned, out, 3 * sizeof(double), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("ens cost: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  return MGB_OK;
}

// ---- per-member min / max (init_param, ensemble.py:91-105) ----------------------------------------
__global__ void __launch_bounds__(kEnsThreads) ens_minmax_kernel(const float* __restrict__ depth, long long HW,
                                                                 float* __restrict__ out /* [E, gridDim.x, 2] */) {
  const int e = blockIdx.y;
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(depth + (long long)e * HW + p);
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  __shared__ float sh[2][kEnsThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = mn; sh[1][threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kEnsThreads / 32; ++w) { mn = fminf(mn, sh[0][w]); mx = fmaxf(mx, sh[1][w]); }
    out[((long long)e * gridDim.x + blockIdx.x) * 2 + 0] = mn;
    out[((long long)e * gridDim.x + blockIdx.x) * 2 + 1] = mx;
  }
}

int launch_ens_minmax(const float* depth, int E, long long HW, float* ws, float* host_pinned, int* blocks_out,
                      cudaStream_t stream) {
  const int blocks = int(std::min<long long>((HW + kEnsThreads - 1) / kEnsThreads, 64));
  dim3 grid(blocks, E);
  ens_minmax_kernel<<<grid, kEnsThreads, 0, stream>>>(depth, HW, ws);
  cudaError_t e = cudaMemcpyAsync(host_pinned, ws, sizeof(float) * 2 * blocks * E, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("ens minmax: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  *blocks_out = blocks;
  return MGB_OK;
}

// ---- reduce: align + (median | mean) (+ uncertainty) ; then global min-max renormalisation ------
template <int E>
__global__ void __launch_bounds__(kEnsThreads)
    ens_reduce_kernel(const float* __restrict__ depth, const float* __restrict__ st, long long HW, int shift,
                      int median, float* __restrict__ pred_out, float* __restrict__ unc_out,
                      int* __restrict__ idx_out, float* __restrict__ block_minmax) {
  float s[E], t[E];
#pragma unroll
  for (int e = 0; e < E; ++e) { s[e] = st[e]; t[e] = st[E + e]; }
  float pmin = FLT_MAX, pmax = -FLT_MAX;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    float a[E];
#pragma unroll
    for (int e = 0; e < E; ++e) a[e] = align1(__ldg(depth + (long long)e * HW + p), s[e], t[e], shift);
    float pred, unc = 0.f;
    int pick = 0;
    if (median) {
      float v[E]; int idx[E];
#pragma unroll
      for (int e = 0; e < E; ++e) { v[e] = a[e]; idx[e] = e; }
      sort_small<E>(v, idx);
      pred = v[(E - 1) / 2];            // torch.median: LOWER median for even E
      pick = idx[(E - 1) / 2];
      if (unc_out) {
        float dv[E]; int di[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { dv[e] = fabsf(a[e] - pred); di[e] = e; }
        sort_small<E>(dv, di);
        unc = dv[(E - 1) / 2];          // MAD
      }
    } else {
      float sm = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) sm += a[e];
      pred = sm / float(E);
      if (unc_out) {                    // torch.std: unbiased
        float q = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) { const float d = a[e] - pred; q += d * d; }
        unc = sqrtf(q / float(E - 1));
      }
      pick = -1;
    }
    pred_out[p] = pred;
    if (unc_out) unc_out[p] = unc;
    if (idx_out) idx_out[p] = pick;
    pmin = fminf(pmin, pred); pmax = fmaxf(pmax, pred);
  }
  __shared__ float sh[2][kEnsThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pmin = fminf(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
    pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = pmin; sh[1][threadIdx.x >> 5] = pmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kEnsThreads / 32; ++w) { pmin = fminf(pmin, sh[0][w]); pmax = fmaxf(pmax, sh[1][w]); }
    block_minmax[2 * blockIdx.x] = pmin; block_minmax[2 * blockIdx.x + 1] = pmax;
  }
}

__global__ void __launch_bounds__(kEnsThreads)
    ens_renorm_kernel(float* __restrict__ pred, float* __restrict__ unc, long long HW, const float* __restrict__ bmm,
                      int nblocks, int use_min) {
  __shared__ float s_min, s_rng;
  if (threadIdx.x == 0) {
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int b = 0; b < nblocks; ++b) { mn = fminf(mn, bmm[2 * b]); mx = fmaxf(mx, bmm[2 * b + 1]); }
    if (!use_min) mn = 0.f;                       // scale-only alignment: depth_min = 0 (ensemble.py:187-188)
    s_min = mn;
    s_rng = fmaxf(mx - mn, 1e-6f);                // .clamp(min=1e-6)
  }
  __syncthreads();
  const float mn = s_min, rng = s_rng;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    pred[p] = (pred[p] - mn) / rng;
    if (unc) unc[p] = unc[p] / rng;
  }
}

int launch_ens_depth_reduce(const float* depth, const float* st_host, int E, long long HW, int shift, int median,
                            int use_min, float* pred, float* unc, int* idx, void* ws, cudaStream_t stream) {
  if (E < 2 || E > kEnsMaxE) { set_error("ensemble size %d outside [2, %d]", E, kEnsMaxE); return MGB_ERR_UNSUPPORTED; }
  CostPartial* partials = reinterpret_cast<CostPartial*>(ws);
  float* st = reinterpret_cast<float*>(partials + kEnsMaxBlocks);
  float* bmm = reinterpret_cast<float*>(ws);  // reuse the partial area for block min/max
  cudaError_t e = cudaMemcpyAsync(st, st_host, sizeof(float) * 2 * E, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) { set_error("ens reduce H2D: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  const int blocks = int(std::min<long long>((HW + kEnsThreads - 1) / kEnsThreads, kEnsMaxBlocks));
  switch (E) {
#define CASE(n) case n: ens_reduce_kernel<n><<<blocks, kEnsThreads, 0, stream>>>(depth, st, HW, shift, median, pred, unc, idx, bmm); break;
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
    CASE(15) CASE(16)
#undef CASE
  }
  ens_renorm_kernel<<<blocks, kEnsThreads, 0, stream>>>(pred, unc, HW, bmm, blocks, use_min);
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("ens reduce: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  return MGB_OK;
}

// ---- ensemble_normals -------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEnsThreads)
    ens_normals_kernel(const float* __restrict__ nrm, int E, long long HW, int closest, float* __restrict__ out,
                       float* __restrict__ unc, int* __restrict__ idx_out) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    // mean over members: sequential fp32 sum in member order, then one division (torch CPU mean)
    float mx = 0.f, my = 0.f, mz = 0.f;
    for (int e = 0; e < E; ++e) {
      const float* q = nrm + ((long long)e * 3) * HW + p;
      mx = __fadd_rn(mx, __ldg(q)); my = __fadd_rn(my, __ldg(q + HW)); mz = __fadd_rn(mz, __ldg(q + 2 * HW));
    }
    mx = __fdiv_rn(mx, float(E)); my = __fdiv_rn(my, float(E)); mz = __fdiv_rn(mz, float(E));
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz));
    const float nn = fmaxf(__fsqrt_rn(n2), 1e-6f);
    mx = __fdiv_rn(mx, nn); my = __fdiv_rn(my, nn); mz = __fdiv_rn(mz, nn);
    float best = -FLT_MAX, acc_unc = 0.f;
    int bi = 0;
    for (int e = 0; e < E; ++e) {
      const float* q = nrm + ((long long)e * 3) * HW + p;
      float sim = __fadd_rn(__fadd_rn(__fmul_rn(mx, __ldg(q)), __fmul_rn(my, __ldg(q + HW))),
                            __fmul_rn(mz, __ldg(q + 2 * HW)));
      sim = fminf(fmaxf(sim, -1.f), 1.f);
      if (sim > best) { best = sim; bi = e; }      // first maximum wins (torch.argmax)
      if (unc) acc_unc += acosf(sim);
    }
    if (unc) unc[p] = (acc_unc / float(E)) / 3.14159265358979323846f;
    if (closest) {
      const float* q = nrm + ((long long)bi * 3) * HW + p;
      out[p] = __ldg(q); out[HW + p] = __ldg(q + HW); out[2 * HW + p] = __ldg(q + 2 * HW);
    } else {
      out[p] = mx; out[HW + p] = my; out[2 * HW + p] = mz;
    }
    if (idx_out) idx_out[p] = bi;
  }
}

int launch_ens_normals(const float* nrm, int E, long long HW, int closest, float* out, float* unc, int* idx,
                       cudaStream_t stream) {
  if (E < 1) { set_error("ensemble size %d", E); return MGB_ERR_INVALID; }
  const int blocks = int(std::min<long long>((HW + kEnsThreads - 1) / kEnsThreads, kEnsMaxBlocks));
  ens_normals_kernel<<<blocks, kEnsThreads, 0, stream>>>(nrm, E, HW, closest, out, unc, idx);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("ens normals: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  return MGB_OK;
}

}  // namespace mgb
