// Flash self-attention for head_dim 64 on sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces F.scaled_dot_product_attention under diffusers' Attention (attn1 of every
// BasicTransformerBlock) reached from reference marigold/marigold_depth_pipeline.py:461-463.
//
//   qkv : bf16 [NB * T, 3C]   (Q | K | V column blocks; head h owns columns h*64 .. h*64+63)
//   out : bf16 [NB * T, C]
//
// One CTA = 128 queries of one (image, head), KV processed in blocks of 64 tokens; 192 threads:
//   warp 0     TMA producer: Q tile (128 x 64) once, K / V tiles (64 x 64) through 3-stage rings
//   warp 1     TMEM allocator + MMA issuer
//                S_b = Q K_j^T          M128 N64 K64, both operands K-major, b = j & 1 (double-buffered)
//                O  += P_b V_j          M128 N64 K64, A = P (bf16, K-major, written by the softmax warps
//                                       into a SWIZZLE_128B smem tile), B = V in its natural [token, d]
//                                       layout = MN-major operand (no transpose pass); O stays in TMEM
//   warps 2-5  softmax, one query row per thread, ONE TMEM read of S per block:
//                lazy rescaling: the row keeps a reference max m_ref; P = exp2(S*c - m_ref). Only when
//                a block's max exceeds m_ref by more than 8 (P could exceed 2^8) is O in TMEM
//                rescaled (tcgen05.ld / st), which after the first blocks is rare; otherwise the
//                softmax warps never wait on the PV MMA.
// TMEM: S0 [0,64) S1 [64,128) O [128,192) -> 256 columns, two CTAs per SM (smem ~98 KB each) so the
// exp (MUFU) phase of one CTA overlaps the load/convert/store phase of the other.
#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace mgb {

constexpr int kAttnThreads = 192;
constexpr int kQBytes = 128 * 128;       // 128 rows x 64 bf16
constexpr int kKvBytes = 64 * 128;       // 64 rows x 64 bf16
constexpr int kPBytes = 128 * 128;       // 128 rows x 64 bf16
constexpr int kKvStages = 3;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
// P (bf16) goes back to tensor memory and feeds the PV MMA as a TMEM A-operand: no smem round trip and no
// generic->async proxy fence in the softmax loop.
constexpr bool kPInTmem = true;

struct AttnParams {
  CUtensorMap tmap_q;   // 3D {3C, T, NB}, box {64, 128, 1}
  CUtensorMap tmap_kv;  // 3D {3C, T, NB}, box {64, 64, 1}
  bf16* out;
  int T, C;
  float scale_log2;
};

__global__ void __launch_bounds__(kAttnThreads, 2) flash_attn64_kernel(const __grid_constant__ AttnParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;
  uint8_t* sV = sK + kKvStages * kKvBytes;
  uint8_t* sP = sV + kKvStages * kKvBytes;   // 2 buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
  uint64_t* q_full = bars;                   // 1
  uint64_t* k_full = bars + 1;               // [3]
  uint64_t* k_empty = bars + 4;              // [3]
  uint64_t* v_full = bars + 7;               // [3]
  uint64_t* v_empty = bars + 10;             // [3]
  uint64_t* s_full = bars + 13;              // [2]
  uint64_t* p_full = bars + 15;              // [2]  (128 arrivals)
  uint64_t* p_empty = bars + 17;             // [2]  PV(j) complete: P buffer free, O updated
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, img = blockIdx.z;
  const int nkv = (p.T + 63) / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_q);
    tma_prefetch_desc(&p.tmap_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], 128);
      mbar_init(&p_empty[b], 1);
    }
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
  const uint32_t tmem_o = tmem_base + 128;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_3d(sQ, &p.tmap_q, q_full, head * 64, q0, img);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % kKvStages;
        const uint32_t ph = (j / kKvStages) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], kKvBytes);
        tma_load_3d(sK + s * kKvBytes, &p.tmap_kv, &k_full[s], p.C + head * 64, j * 64, img);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], kKvBytes);
        tma_load_3d(sV + s * kKvBytes, &p.tmap_kv, &v_full[s], 2 * p.C + head * 64, j * 64, img);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, false);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, true);
    auto issue_s = [&](int j) {
      const int s = j % kKvStages;
      mbar_wait(&k_full[s], (j / kKvStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dq = umma_desc_sw128(smem_u32(sQ));
        const uint64_t dk = umma_desc_sw128(smem_u32(sK + s * kKvBytes));
        const uint32_t ts = tmem_base + (j & 1) * 64;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(ts, dq + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_s, k > 0);
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[j & 1]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    for (int j = 0; j < nkv; ++j) {
      // S(j+1) goes into the other S buffer: free because softmax(j-1) signalled p_full(j-1), which
      // this warp waited for before PV(j-1)
      if (j + 1 < nkv) issue_s(j + 1);
      const int b = j & 1, s = j % kKvStages;
      mbar_wait(&p_full[b], (j >> 1) & 1);
      mbar_wait(&v_full[s], (j / kKvStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t pbase = smem_u32(sP + b * kPBytes), vbase = smem_u32(sV + s * kKvBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // B: V [kv, d] d-contiguous (MN-major): 16 kv rows = 2048 B per K=16 step
          if constexpr (kPInTmem) {
            // A: P in TMEM, 16 bf16 = 8 columns per K=16 step
            umma_bf16_ts(tmem_o, tmem_base + 192 + b * 32 + k * 8, umma_desc_sw128(vbase + k * 2048), idesc_pv,
                         (j > 0 || k > 0) ? 1u : 0u);
          } else {
            // A: P rows K-major in smem, 32 B per K=16 step
            umma_bf16(tmem_o, umma_desc_sw128(pbase + k * 32), umma_desc_sw128(vbase + k * 2048), idesc_pv,
                      (j > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&v_empty[s]);
        umma_commit(&p_empty[b]);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    float m_ref = 0.f, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int b = j & 1, u = j >> 1;
      mbar_wait(&s_full[b], u & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld32(tmem_base + lane_off + b * 64, r0);
      tmem_ld32(tmem_base + lane_off + b * 64 + 32, r1);
      tmem_wait_ld();
      const int kv_valid = p.T - j * 64;   // >= 64 except possibly in the last block
      if (kv_valid < 64) {                 // ragged tail (T % 64 != 0): mask once, then share the fast path
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= kv_valid) r0[i] = 0xff800000u;        // -inf
          if (32 + i >= kv_valid) r1[i] = 0xff800000u;
        }
      }
      // row max with 4 independent chains (one softmax warp per scheduler: dependent chains are the cost)
      float mxa = __uint_as_float(r0[0]), mxb = __uint_as_float(r0[1]), mxc = __uint_as_float(r1[0]),
            mxd = __uint_as_float(r1[1]);
#pragma unroll
      for (int i = 2; i < 32; i += 2) {
        mxa = fmaxf(mxa, __uint_as_float(r0[i]));
        mxb = fmaxf(mxb, __uint_as_float(r0[i + 1]));
        mxc = fmaxf(mxc, __uint_as_float(r1[i]));
        mxd = fmaxf(mxd, __uint_as_float(r1[i + 1]));
      }
      const float mx = fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd));
      const float m_blk = mx * p.scale_log2;
      if (j == 0) {
        m_ref = m_blk;
      } else {
        const bool need = m_blk > m_ref + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // rescale O (and l) of the rows that need it; other rows multiply by 1
          const float m_new = need ? m_blk : m_ref;
          const float alpha = ex2_approx(m_ref - m_new);
          m_ref = m_new;
          l_run *= alpha;
          mbar_wait(&p_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);   // every PV issued so far has completed
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem_o + lane_off + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tmem_o + lane_off + c * 32, o);
          }
          tmem_wait_st();
        }
      }
      // P = exp2(S * c - m_ref) (exp2(-inf) = 0 masks the tail); bf16 pairs; 4 partial row sums
      uint32_t pk[32];
      float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
      const float nm = -m_ref;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a0 = ex2_approx(fmaf(__uint_as_float(r0[2 * i]), p.scale_log2, nm));
        const float a1 = ex2_approx(fmaf(__uint_as_float(r0[2 * i + 1]), p.scale_log2, nm));
        const float b0 = ex2_approx(fmaf(__uint_as_float(r1[2 * i]), p.scale_log2, nm));
        const float b1 = ex2_approx(fmaf(__uint_as_float(r1[2 * i + 1]), p.scale_log2, nm));
        ls0 += a0; ls1 += a1; ls2 += b0; ls3 += b1;
        pk[i] = pack_bf16x2(a0, a1);
        pk[16 + i] = pack_bf16x2(b0, b1);
      }
      const float lsum = (ls0 + ls1) + (ls2 + ls3);
      l_run += lsum;
      // P buffer b was last read by PV(j-2)
      mbar_wait(&p_empty[b], (u & 1) ^ 1);
      if constexpr (kPInTmem) {
        tmem_st32(tmem_base + 192 + b * 32 + lane_off, pk);
        tmem_wait_st();
      } else {
        uint8_t* prow = sP + b * kPBytes + row * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint4* dst = reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4));
          *dst = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        }
        fence_proxy_async_smem();
      }
      tc_fence_before();
      mbar_arrive(&p_full[b]);
    }
    // epilogue: O / l
    mbar_wait(&p_empty[(nkv - 1) & 1], ((nkv - 1) >> 1) & 1);
    tc_fence_after();
    const int qrow = q0 + row;
    const float inv = 1.f / l_run;
    uint32_t o0[32], o1[32];
    tmem_ld32(tmem_o + lane_off, o0);
    tmem_ld32(tmem_o + lane_off + 32, o1);
    tmem_wait_ld();
    if (qrow < p.T) {
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)img * p.T + qrow) * p.C + head * 64);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        dst[i] = make_uint4(pack_bf16x2(__uint_as_float(o0[8 * i]) * inv, __uint_as_float(o0[8 * i + 1]) * inv),
                            pack_bf16x2(__uint_as_float(o0[8 * i + 2]) * inv, __uint_as_float(o0[8 * i + 3]) * inv),
                            pack_bf16x2(__uint_as_float(o0[8 * i + 4]) * inv, __uint_as_float(o0[8 * i + 5]) * inv),
                            pack_bf16x2(__uint_as_float(o0[8 * i + 6]) * inv, __uint_as_float(o0[8 * i + 7]) * inv));
#pragma unroll
      for (int i = 0; i < 4; ++i)
        dst[4 + i] = make_uint4(pack_bf16x2(__uint_as_float(o1[8 * i]) * inv, __uint_as_float(o1[8 * i + 1]) * inv),
                                pack_bf16x2(__uint_as_float(o1[8 * i + 2]) * inv, __uint_as_float(o1[8 * i + 3]) * inv),
                                pack_bf16x2(__uint_as_float(o1[8 * i + 4]) * inv, __uint_as_float(o1[8 * i + 5]) * inv),
                                pack_bf16x2(__uint_as_float(o1[8 * i + 6]) * inv, __uint_as_float(o1[8 * i + 7]) * inv));
    }
    tc_fence_before();
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
  const uint32_t box_q[3] = {64, 128, 1}, box_kv[3] = {64, 64, 1};
  int rc = make_tmap_3d(&p.tmap_q, qkv, dims, strides, box_q);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmap_kv, qkv, dims, strides, box_kv);
  if (rc) return rc;
  p.out = out; p.T = T; p.C = C;
  p.scale_log2 = scale * 1.4426950408889634f;
  const size_t smem = 1024 + kQBytes + 2 * kKvStages * kKvBytes + 2 * kPBytes + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) { set_error("flash_attn64 attr: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
    attr_set = true;
  }
  dim3 grid((T + 127) / 128, C / 64, NB);
  cudaError_t e = launch_k(flash_attn64_kernel, grid, kAttnThreads, smem, stream, p);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("flash_attn64 launch: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  return MGB_OK;
}

}  // namespace mgb
