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
//   warps 2-9  softmax. A query row is shared by TWO threads (warps w and w+4 own the same TMEM lane quarter):
//                each handles 32 of the block's 64 scores, ONE TMEM read of S per block; the row max is
//                exchanged through shared memory (named barrier per lane quarter). Halving the per-thread
//                dependent chain and doubling the warps per scheduler is worth more than the exchange costs:
//                the loop is bound by one warp's LDTM -> max -> exp2 -> STTM latency chain, not by a pipe.
//                Lazy rescaling: the row keeps a reference max m_ref; P = exp2(S*c - m_ref). Only when
//                a block's max exceeds m_ref by more than 8 (P could exceed 2^8) is O in TMEM
//                rescaled (tcgen05.ld / st), which after the first blocks is rare; otherwise the
//                softmax warps never wait on the PV MMA.
// TMEM: S0 [0,64) S1 [64,128) O [128,192) -> 256 columns, two CTAs per SM (smem ~98 KB each) so the
// exp (MUFU) phase of one CTA overlaps the load/convert/store phase of the other.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace mgb {

constexpr int attn_threads(int rt) { return 64 + 128 * rt; }   // producer + MMA warps, 4 RT softmax warps
constexpr int kQBytes = 128 * 128;       // 128 rows x 64 bf16
constexpr int kKvBytes = 64 * 128;       // 64 rows x 64 bf16
constexpr int kPBytes = 128 * 128;       // 128 rows x 64 bf16
constexpr int kKvStages = 3;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
constexpr int kAttnDefaultRT = 2;
// P (bf16) goes back to tensor memory and feeds the PV MMA as a TMEM A-operand: no smem round trip and no
// generic->async proxy fence in the softmax loop.
// RT = softmax threads per query row (2 or 4: warps w, w+4, .. own the same TMEM lane quarter and split a block's 64
// scores). The loop is a per-warp dependent chain (LDTM -> max -> exchange -> exp2 -> STTM), not a saturated pipe: ncu
// reads the XU pipe (MUFU.EX2 + F2FP) at 102 %, yet converting on the integer ALU and evaluating 25-50 % of the
// exponentials as an FMA-pipe polynomial made the kernel 7-18 % SLOWER (r01 and r02). More, shorter chains per row are
// the lever: RT = 4 halves every thread's share and doubles the warps per scheduler.
struct AttnParams {
  CUtensorMap tmap_q;   // 3D {3C, T, NB}, box {64, 128, 1}
  CUtensorMap tmap_kv;  // 3D {3C, T, NB}, box {64, 64, 1}
  bf16* out;
  int T, C;
  float scale_log2;
  // split-KV (balances the last wave: 72 x 5 = 360 tiles on 296 CTA slots is 2 rounds of full tiles, but 1.25 rounds
  // of quarter tiles): blockIdx.z = img * splits + split; split s covers KV blocks [s * nkv / splits, (s+1) * ...).
  // With splits > 1 the CTA writes un-normalised fp32 O plus (m, l) per row; attn_combine_kernel merges them.
  int splits;
  float* part_o;    // [splits][NB][C/64][T][64]
  float* part_ml;   // [splits][NB][C/64][T][2]   (m in log2 units incl. the softmax scale, l)
};

// 64-thread named barrier of one TMEM lane quarter (constant ids: a register id makes ptxas reserve all 16)
template <int N>
__device__ __forceinline__ void quarter_barrier(int q) {
  switch (q) {
    case 0: asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); break;
    case 1: asm volatile("bar.sync 2, %0;" ::"n"(N) : "memory"); break;
    case 2: asm volatile("bar.sync 3, %0;" ::"n"(N) : "memory"); break;
    default: asm volatile("bar.sync 4, %0;" ::"n"(N) : "memory"); break;
  }
}

template <int RT>
__global__ void __launch_bounds__(attn_threads(RT), 2) flash_attn64_kernel(const __grid_constant__ AttnParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;
  uint8_t* sV = sK + kKvStages * kKvBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kKvStages * kKvBytes);
  uint64_t* q_full = bars;                   // 1
  uint64_t* k_full = bars + 1;               // [3]
  uint64_t* k_empty = bars + 4;              // [3]
  uint64_t* v_full = bars + 7;               // [3]
  uint64_t* v_empty = bars + 10;             // [3]
  uint64_t* s_full = bars + 13;              // [2]
  uint64_t* p_full = bars + 15;              // [2]  (128 arrivals)
  uint64_t* p_empty = bars + 17;             // [2]  PV(j) complete: P buffer free, O updated
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 19);
  float* s_xch = reinterpret_cast<float*>(bars + 20);   // [2 slots][4 quarters][2 halves][32 lanes] row-max exchange (+ final l)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, img = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const int nkv_all = (p.T + 63) / 64;
  const int jb0 = split * nkv_all / p.splits;                 // first KV block of this CTA
  const int nkv = (split + 1) * nkv_all / p.splits - jb0;     // its number of KV blocks (>= 1: host keeps splits <= nkv_all)

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
      mbar_init(&p_full[b], 128 * RT);
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

  // Producer and MMA issuer are single-thread latency chains (see gemm_tc.cu): whole loop inside one elected
  // thread, shared-window addresses and descriptor words precomputed, counters instead of % and /. The first
  // version (elect + warp sync + generic addressing per phase) took ~1300 cycles per KV block and bounded the
  // whole kernel (384k cycles for T = 9216 = 2 waves x 144 blocks x 1333).
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const uint32_t kfull = smem_u32(k_full), kempty = smem_u32(k_empty), vfull = smem_u32(v_full),
                     vempty = smem_u32(v_empty);
      const uint32_t sK_a = smem_u32(sK), sV_a = smem_u32(sV);
      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_3d(sQ, &p.tmap_q, q_full, head * 64, q0, img);
      const int ck = p.C + head * 64, cv = 2 * p.C + head * 64;
      uint32_t s = 0, ph = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait_a(kempty + s * 8, ph ^ 1);
        mbar_expect_tx_a(kfull + s * 8, kKvBytes);
        tma_load_3d_a(sK_a + s * kKvBytes, &p.tmap_kv, kfull + s * 8, ck, (jb0 + j) * 64, img);
        mbar_wait_a(vempty + s * 8, ph ^ 1);
        mbar_expect_tx_a(vfull + s * 8, kKvBytes);
        tma_load_3d_a(sV_a + s * kKvBytes, &p.tmap_kv, vfull + s * 8, cv, (jb0 + j) * 64, img);
        if (++s == kKvStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, true);
      constexpr uint32_t kHi = uint32_t(kDescSw128Hi >> 32), kLbo = 1u << 16;
      const uint32_t kfull = smem_u32(k_full), kempty = smem_u32(k_empty), vfull = smem_u32(v_full),
                     vempty = smem_u32(v_empty), sfull = smem_u32(s_full), pfull = smem_u32(p_full),
                     pempty = smem_u32(p_empty);
      const uint32_t dq_lo = (smem_u32(sQ) >> 4) | kLbo, dk_lo0 = (smem_u32(sK) >> 4) | kLbo,
                     dv_lo0 = (smem_u32(sV) >> 4) | kLbo;
      uint32_t ks = 0, kph = 0, vs = 0, vph = 0;
      auto issue_s = [&](int j) {
        mbar_wait_a(kfull + ks * 8, kph);
        const uint32_t dk_lo = dk_lo0 + ks * uint32_t(kKvBytes >> 4);
        const uint32_t ts = tmem_base + uint32_t(j & 1) * 64;
        umma_bf16(ts, make_u64(dq_lo, kHi), make_u64(dk_lo, kHi), idesc_s, 0u);
        umma_bf16(ts, make_u64(dq_lo + 2, kHi), make_u64(dk_lo + 2, kHi), idesc_s, 1u);
        umma_bf16(ts, make_u64(dq_lo + 4, kHi), make_u64(dk_lo + 4, kHi), idesc_s, 1u);
        umma_bf16(ts, make_u64(dq_lo + 6, kHi), make_u64(dk_lo + 6, kHi), idesc_s, 1u);
        umma_commit_a(kempty + ks * 8);
        umma_commit_a(sfull + uint32_t(j & 1) * 8);
        if (++ks == kKvStages) { ks = 0; kph ^= 1; }
      };
      mbar_wait_a(smem_u32(q_full), 0);
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        // S(j+1) goes into the other S buffer: free because softmax(j-1) signalled p_full(j-1), which
        // this thread waited for before PV(j-1)
        if (j + 1 < nkv) issue_s(j + 1);
        const uint32_t b = uint32_t(j & 1);
        mbar_wait_a(pfull + b * 8, uint32_t(j >> 1) & 1u);
        mbar_wait_a(vfull + vs * 8, vph);
        tc_fence_after();
        // A: P (bf16) in TMEM, 16 bf16 = 8 columns per K=16 step; B: V [kv, d] d-contiguous (MN-major):
        // 16 kv rows = 2048 B per K=16 step
        const uint32_t tp = tmem_base + 192 + b * 32;
        const uint32_t dv_lo = dv_lo0 + vs * uint32_t(kKvBytes >> 4);
        umma_bf16_ts(tmem_o, tp, make_u64(dv_lo, kHi), idesc_pv, j > 0 ? 1u : 0u);
        umma_bf16_ts(tmem_o, tp + 8, make_u64(dv_lo + 128, kHi), idesc_pv, 1u);
        umma_bf16_ts(tmem_o, tp + 16, make_u64(dv_lo + 256, kHi), idesc_pv, 1u);
        umma_bf16_ts(tmem_o, tp + 24, make_u64(dv_lo + 384, kHi), idesc_pv, 1u);
        umma_commit_a(vempty + vs * 8);
        umma_commit_a(pempty + b * 8);
        if (++vs == kKvStages) { vs = 0; vph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax =====================
    constexpr int NS = 64 / RT;             // scores (and O columns) per thread
    const int q = warp & 3;                 // TMEM lane quarter
    const int h = (warp - 2) >> 2;          // which part of the block's 64 scores / of O's 64 columns
    const int row = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    float m_ref = 0.f, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int b = j & 1, u = j >> 1;
      mbar_wait(&s_full[b], u & 1);
      tc_fence_after();
      uint32_t r[NS];
      TmemIO<NS>::ld(tmem_base + lane_off + b * 64 + h * NS, r);
      tmem_wait_ld();
      const int kv_valid = p.T - (jb0 + j) * 64 - h * NS;   // >= NS except possibly in the last block
      if (kv_valid < NS) {                          // ragged tail (T % 64 != 0): mask once, then share the fast path
#pragma unroll
        for (int i = 0; i < NS; ++i)
          if (i >= kv_valid) r[i] = 0xff800000u;    // -inf
      }
      // partial row max with 4 independent chains, then the exchange with the warps owning the other scores of the row
      float mxa = __uint_as_float(r[0]), mxb = __uint_as_float(r[1]), mxc = __uint_as_float(r[2]),
            mxd = __uint_as_float(r[3]);
#pragma unroll
      for (int i = 4; i < NS; i += 4) {
        mxa = fmaxf(mxa, __uint_as_float(r[i]));
        mxb = fmaxf(mxb, __uint_as_float(r[i + 1]));
        mxc = fmaxf(mxc, __uint_as_float(r[i + 2]));
        mxd = fmaxf(mxd, __uint_as_float(r[i + 3]));
      }
      float mx = fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd));
      float* xs = s_xch + ((b * 4 + q) * RT) * 32;
      xs[h * 32 + lane] = mx;
      quarter_barrier<32 * RT>(q);
#pragma unroll
      for (int k = 1; k < RT; ++k) mx = fmaxf(mx, xs[((h + k) % RT) * 32 + lane]);
      const float m_blk = mx * p.scale_log2;
      if (j == 0) {
        m_ref = m_blk;
      } else {
        const bool need = m_blk > m_ref + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {      // identical decision in every warp of the quarter
          // rescale O (and l) of the rows that need it; other rows multiply by 1
          const float m_new = need ? m_blk : m_ref;
          const float alpha = ex2_approx(m_ref - m_new);
          m_ref = m_new;
          l_run *= alpha;
          mbar_wait(&p_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);   // every PV issued so far has completed
          tc_fence_after();
          uint32_t o[NS];
          TmemIO<NS>::ld(tmem_o + lane_off + h * NS, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < NS; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          TmemIO<NS>::st(tmem_o + lane_off + h * NS, o);
          tmem_wait_st();
        }
      }
      // P = exp2(S * c - m_ref) (exp2(-inf) = 0 masks the tail); bf16 pairs; 4 partial row sums
      // (the scale-and-shift and the row sums go through FFMA2 / FADD2: two scores per instruction)
      uint32_t pk[NS / 2];
      f2 ls01 = f2_splat(0.f), ls23 = ls01;
      const f2 sc2 = f2_splat(p.scale_log2), nm2 = f2_splat(-m_ref);
#pragma unroll
      for (int i = 0; i < NS / 4; ++i) {
        float t0, t1, t2, t3;
        f2_split(f2_fma(f2_make(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), sc2, nm2), t0, t1);
        f2_split(f2_fma(f2_make(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), sc2, nm2), t2, t3);
        const float a0 = ex2_approx(t0), a1 = ex2_approx(t1), a2 = ex2_approx(t2), a3 = ex2_approx(t3);
        ls01 = f2_add(ls01, f2_make(a0, a1));
        ls23 = f2_add(ls23, f2_make(a2, a3));
        pk[2 * i] = pack_bf16x2(a0, a1);
        pk[2 * i + 1] = pack_bf16x2(a2, a3);
      }
      {
        float ls0, ls1, ls2, ls3;
        f2_split(ls01, ls0, ls1);
        f2_split(ls23, ls2, ls3);
        l_run += (ls0 + ls1) + (ls2 + ls3);
      }
      // P buffer b was last read by PV(j-2)
      mbar_wait(&p_empty[b], (u & 1) ^ 1);
      TmemIO<NS / 2>::st(tmem_base + 192 + b * 32 + h * (NS / 2) + lane_off, pk);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[b]);
    }
    // epilogue: O / l. The parts of a row add their partial sums through the exchange buffer
    // (slot (nkv & 1): not the one the last block's max exchange used).
    float* xl = s_xch + (((nkv & 1) * 4 + q) * RT) * 32;
    xl[h * 32 + lane] = l_run;
    quarter_barrier<32 * RT>(q);
    float l_tot = 0.f;
#pragma unroll
    for (int k = 0; k < RT; ++k) l_tot += xl[k * 32 + lane];      // same order in every part: identical l_tot
    mbar_wait(&p_empty[(nkv - 1) & 1], ((nkv - 1) >> 1) & 1);
    tc_fence_after();
    const int qrow = q0 + row;
    uint32_t o0[NS];
    TmemIO<NS>::ld(tmem_o + lane_off + h * NS, o0);
    tmem_wait_ld();
    if (qrow < p.T) {
      if (p.splits == 1) {
        const float inv = 1.f / l_tot;
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)img * p.T + qrow) * p.C + head * 64 + h * NS);
#pragma unroll
        for (int i = 0; i < NS / 8; ++i)
          dst[i] = make_uint4(pack_bf16x2(__uint_as_float(o0[8 * i]) * inv, __uint_as_float(o0[8 * i + 1]) * inv),
                              pack_bf16x2(__uint_as_float(o0[8 * i + 2]) * inv, __uint_as_float(o0[8 * i + 3]) * inv),
                              pack_bf16x2(__uint_as_float(o0[8 * i + 4]) * inv, __uint_as_float(o0[8 * i + 5]) * inv),
                              pack_bf16x2(__uint_as_float(o0[8 * i + 6]) * inv, __uint_as_float(o0[8 * i + 7]) * inv));
      } else {
        const size_t prow = ((size_t(split) * gridDim.z / p.splits + img) * gridDim.y + head) * p.T + qrow;
        uint4* dst = reinterpret_cast<uint4*>(p.part_o + prow * 64 + h * NS);
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) dst[i] = make_uint4(o0[4 * i], o0[4 * i + 1], o0[4 * i + 2], o0[4 * i + 3]);
        if (h == 0) *reinterpret_cast<float2*>(p.part_ml + prow * 2) = make_float2(m_ref, l_tot);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// Merge the split-KV partials: out[row, :] = sum_s w_s O_s / sum_s w_s l_s, w_s = 2^(m_s - max m). One thread per
// (row, 8 columns).
__global__ void __launch_bounds__(256) attn_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml,
                                                           bf16* __restrict__ out, int splits, int NB, int heads, int T, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t rows = size_t(NB) * heads * T;
  const size_t gid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t r = gid >> 3;
  const int c8 = int(gid & 7) * 8;
  if (r >= rows) return;
  float m = -INFINITY;
  for (int s = 0; s < splits; ++s) m = fmaxf(m, __ldg(part_ml + (size_t(s) * rows + r) * 2));
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, l = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float2 ml = __ldg(reinterpret_cast<const float2*>(part_ml + (size_t(s) * rows + r) * 2));
    const float w = ex2_approx(ml.x - m);
    l = fmaf(ml.y, w, l);
    const float4* po = reinterpret_cast<const float4*>(part_o + (size_t(s) * rows + r) * 64 + c8);
    const float4 a = __ldg(po), b = __ldg(po + 1);
    acc[0] = fmaf(a.x, w, acc[0]); acc[1] = fmaf(a.y, w, acc[1]); acc[2] = fmaf(a.z, w, acc[2]); acc[3] = fmaf(a.w, w, acc[3]);
    acc[4] = fmaf(b.x, w, acc[4]); acc[5] = fmaf(b.y, w, acc[5]); acc[6] = fmaf(b.z, w, acc[6]); acc[7] = fmaf(b.w, w, acc[7]);
  }
  const float inv = 1.f / l;
  const size_t t = r % T, ih = r / T;
  const size_t img = ih / heads, head = ih % heads;
  uint4* dst = reinterpret_cast<uint4*>(out + (img * T + t) * C + head * 64 + c8);
  *dst = make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                    pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
}

// KV splits that minimise the number of CTA rounds (2 CTAs per SM) weighted by the split's length
int flash_attn64_splits(int NB, int T, int C) {
  const int units = ((T + 127) / 128) * (C / 64) * NB, nkv = (T + 63) / 64, slots = 148 * 2;
  int best = 1;
  double best_t = 1e30;
  for (int s = 1; s <= 8; ++s) {
    if (s > 1 && nkv / s < 6) break;
    const double rounds = double((units * s + slots - 1) / slots);
    // measured (r01, T = 9216): a CTA's fixed cost (prologue, Q load, first S, epilogue) is worth ~15 KV blocks,
    // the combine pass ~8
    const double t = rounds * (double(nkv) / s + 15.0) + (s > 1 ? 8.0 : 0.0);
    if (t < best_t - 1e-9) { best_t = t; best = s; }
  }
  return best;
}
size_t flash_attn64_ws_bytes(int NB, int T, int C) {
  const int s = flash_attn64_splits(NB, T, C);
  if (s == 1) return 0;
  return size_t(s) * NB * (C / 64) * T * (64 + 2) * sizeof(float);
}

int launch_flash_attn64(const bf16* qkv, bf16* out, int NB, int T, int C, float scale, float* ws, size_t ws_bytes,
                        cudaStream_t stream) {
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
  p.splits = 1; p.part_o = nullptr; p.part_ml = nullptr;
  {
    const int sp = flash_attn64_splits(NB, T, C);
    if (sp > 1 && ws != nullptr && ws_bytes >= flash_attn64_ws_bytes(NB, T, C)) {
      p.splits = sp;
      p.part_o = ws;
      p.part_ml = ws + size_t(sp) * NB * (C / 64) * T * 64;
    }
  }
  static const size_t dbg_pad = getenv("MGB_ATTN_SMEM_PAD") ? size_t(atoi(getenv("MGB_ATTN_SMEM_PAD"))) : 0;   // debug: force 1 CTA/SM
  const size_t smem = 1024 + kQBytes + 2 * kKvStages * kKvBytes + 256 + 4096 + dbg_pad;
  // softmax threads per row: MGB_ATTN_RT = 2 | 4
  static const int rt = getenv("MGB_ATTN_RT") ? atoi(getenv("MGB_ATTN_RT")) : kAttnDefaultRT;
  void (*kern)(AttnParams) = rt == 4 ? flash_attn64_kernel<4> : flash_attn64_kernel<2>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) { set_error("flash_attn64 attr: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
    attr_set = true;
  }
  dim3 grid((T + 127) / 128, C / 64, NB * p.splits);
  cudaError_t e = launch_k(kern, grid, attn_threads(rt == 4 ? 4 : 2), smem, stream, p);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("flash_attn64 launch: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  if (p.splits > 1) {
    const size_t threads = size_t(NB) * (C / 64) * T * 8;
    e = launch_k(attn_combine_kernel, dim3(unsigned((threads + 255) / 256)), 256, 0, stream, (const float*)p.part_o,
                 (const float*)p.part_ml, out, p.splits, NB, C / 64, T, C);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("attn_combine launch: %s", cudaGetErrorString(e)); return MGB_ERR_CUDA; }
  }
  return MGB_OK;
}

}  // namespace mgb
