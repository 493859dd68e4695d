// Flash self-attention for head_dim 64 on sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces F.scaled_dot_product_attention under diffusers' Attention (attn1 of every
// BasicTransformerBlock) reached from reference marigold/marigold_depth_pipeline.py:461-463.
//
//   qkv : bf16 [NB * T, 3C]   (Q | K | V column blocks; head h owns columns h*64 .. h*64+63)
//   out : bf16 [NB * T, C]
//
// One CTA = 128 queries of one (image, head); 192 threads:
//   warp 0     TMA producer: Q tile once, then K / V tiles (128 tokens x 64) through 2-stage rings
//   warp 1     TMEM allocator + MMA issuer:  S = Q K^T  (M128 N128 K64, both K-major)
//                                            PV = P V   (M128 N64 K128, A = P K-major from smem,
//                                                        B = V in its natural [token, d] layout
//                                                        = MN-major operand, no transpose pass)
//   warps 2-5  softmax: one query row per thread. S is read from TMEM twice (row max, then
//              exp2 / sum / bf16 P written to a SWIZZLE_128B smem tile); the PV partial product is
//              read back from TMEM and accumulated into registers with the online-softmax rescale.
// TMEM: S fp32 128x128 at columns [0,128), PV fp32 128x64 at [128,192) -> 256 columns per CTA,
// so two CTAs share an SM (smem ~113 KB each) and overlap each other's softmax and MMA phases.
#include "common.cuh"
#include "kernels.h"

namespace mgb {

constexpr int kAttnThreads = 192;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KB: 128 rows x 128 B
constexpr int kKvStages = 2;

struct AttnParams {
  CUtensorMap tmap;  // 3D {3C, T, NB}, box {64, 128, 1}
  bf16* out;
  int T, C;
  float scale_log2;
};

__global__ void __launch_bounds__(kAttnThreads, 2) flash_attn64_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTileBytes;
  uint8_t* sV = sK + kKvStages * kTileBytes;
  uint8_t* sP = sV + kKvStages * kTileBytes;  // 2 x 16 KB (kv columns 0-63 | 64-127)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;              // [2]
  uint64_t* k_empty = bars + 3;             // [2]
  uint64_t* v_full = bars + 5;              // [2]
  uint64_t* v_empty = bars + 7;             // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* p_full = bars + 10;
  uint64_t* pv_full = bars + 11;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, img = blockIdx.z;
  const int nkv = (p.T + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem_base, tmem_pv = tmem_base + 128;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, kTileBytes);
      tma_load_3d(sQ, &p.tmap, q_full, head * 64, q0, img);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % kKvStages;
        const uint32_t ph = (j / kKvStages) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], kTileBytes);
        tma_load_3d(sK + s * kTileBytes, &p.tmap, &k_full[s], p.C + head * 64, j * 128, img);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], kTileBytes);
        tma_load_3d(sV + s * kTileBytes, &p.tmap, &v_full[s], 2 * p.C + head * 64, j * 128, img);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, true);
    mbar_wait(q_full, 0);
    for (int j = 0; j < nkv; ++j) {
      const int s = j % kKvStages;
      const uint32_t ph = (j / kKvStages) & 1;
      mbar_wait(&k_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dq = umma_desc_sw128(smem_u32(sQ));
        const uint64_t dk = umma_desc_sw128(smem_u32(sK + s * kTileBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, dq + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_s, k > 0);
        umma_commit(&k_empty[s]);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, j & 1);
      mbar_wait(&v_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A: P rows K-major; 64-column halves are 16 KB apart, 32 B per K=16 step inside a half
          const uint64_t dp = umma_desc_sw128(smem_u32(sP + (k >> 2) * kTileBytes) + (k & 3) * 32);
          // B: V [kv, d], d contiguous (MN-major): 16 kv rows per step = 2048 B
          const uint64_t dv = umma_desc_sw128(smem_u32(sV + s * kTileBytes) + k * 2048);
          umma_bf16(tmem_pv, dp, dv, idesc_pv, k > 0);
        }
        umma_commit(&v_empty[s]);
        umma_commit(pv_full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kv_valid = min(128, p.T - j * 128);
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_off + c * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float alpha = exp2f(m_run - m_new);
      m_run = m_new;
      // pass 2: P = exp2(S * scale_log2 - m), row sum, bf16 to swizzled smem
      float lsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_off + c * 32, r);
        tmem_wait_ld();
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = exp2f(__uint_as_float(r[2 * i]) * p.scale_log2 - m_new);
          float p1 = exp2f(__uint_as_float(r[2 * i + 1]) * p.scale_log2 - m_new);
          if (c * 32 + 2 * i >= kv_valid) p0 = 0.f;
          if (c * 32 + 2 * i + 1 >= kv_valid) p1 = 0.f;
          // sum what the tensor core will actually multiply (bf16-rounded P)
          __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
          lsum += __bfloat162float(pb.x) + __bfloat162float(pb.y);
          packed[i] = *reinterpret_cast<uint32_t*>(&pb);
        }
        uint8_t* half_base = sP + (c >> 1) * kTileBytes + row * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = (c & 1) * 4 + ch;  // 16 B chunk index inside the 128 B row
          uint4* dst = reinterpret_cast<uint4*>(half_base + ((chunk ^ (row & 7)) << 4));
          *dst = make_uint4(packed[4 * ch], packed[4 * ch + 1], packed[4 * ch + 2], packed[4 * ch + 3]);
        }
      }
      l_run = l_run * alpha + lsum;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      // accumulate the PV partial product
      mbar_wait(pv_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_pv + lane_off + c * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * alpha + __uint_as_float(r[i]);
      }
    }
    tc_fence_before();
    const int qrow = q0 + row;
    if (qrow < p.T) {
      const float inv = 1.f / l_run;
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)img * p.T + qrow) * p.C + head * 64);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        dst[i] = make_uint4(pack_bf16x2(o[8 * i] * inv, o[8 * i + 1] * inv), pack_bf16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv),
                            pack_bf16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv), pack_bf16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv));
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int launch_flash_attn64(const bf16* qkv, bf16* out, int NB, int T, int C, float scale, cudaStream_t stream) {
  if (C % 64 != 0 || T <= 0) {
    set_error("flash_attn64: C %% 64 != 0 or bad T");
    return MGB_ERR_INVALID;
  }
  AttnParams p;
  const uint64_t dims[3] = {uint64_t(3 * C), uint64_t(T), uint64_t(NB)};
  const uint64_t strides[2] = {uint64_t(3 * C) * 2, uint64_t(T) * 3 * C * 2};
  const uint32_t box[3] = {64, 128, 1};
  int rc = make_tmap_3d(&p.tmap, qkv, dims, strides, box);
  if (rc) return rc;
  p.out = out; p.T = T; p.C = C;
  p.scale_log2 = scale * 1.4426950408889634f;
  const size_t smem = 1024 + size_t(1 + 2 * kKvStages + 2) * kTileBytes + 128;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) { set_error("flash_attn64 attr: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
    attr_set = true;
  }
  dim3 grid((T + 127) / 128, C / 64, NB);
  flash_attn64_kernel<<<grid, kAttnThreads, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("flash_attn64 launch: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  return MGB_OK;
}

}  // namespace mgb
