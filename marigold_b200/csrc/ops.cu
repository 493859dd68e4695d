// Host-side operator builders: turn "linear" / "conv2d" requests into GemmParams (tensor maps, tile
// geometry, tap tables) and pick tile shapes. Used by the network composition (net.cu) and by the
// operator-level C ABI (api.cu).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "ops.h"

namespace mgb {

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches += n; }
long long launch_count() { return g_launches.load(); }

int conv_halo_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MGB_CONV_HALO");
    v = e ? atoi(e) : 0;   // measured r01: no gain once the MMA issue loop runs at the tcgen05 floor (6.71 vs 6.54 ms/step)
    if (v == 2) v = 1;     // (variant 2, descriptor base offset set, is numerically WRONG: the swizzle is address based)
    if (v < 0 || v > 3) v = 0;
  }
  return v;
}

constexpr int kHaloTileW = 8, kHaloTileH = 16;
static int halo_slots() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MGB_HALO_SLOTS"); v = e ? atoi(e) : 2; if (v < 2 || v > 4) v = 2; }
  return v;
}
#define kHaloSlots halo_slots()
static int halo_copy_bytes(int variant) {
  const int w = variant == 3 ? kHaloTileW : kHaloTileW + 2;
  return ((w * (kHaloTileH + 2) * 128 + 1023) / 1024) * 1024;
}
static int halo_slot_bytes(int variant) { return halo_copy_bytes(variant) * (variant == 3 ? 3 : 1); }
int conv_halo_ring_bytes(int kind) {
  const int v = conv_halo_variant();
  return (kind == 0 && v > 0) ? kHaloSlots * halo_slot_bytes(v) : 0;
}

void conv_tile_shape(int Hout, int Wout, int* tile_w, int* tile_h, int kind) {
  if (kind == 0 && conv_halo_variant() > 0) {
    *tile_w = kHaloTileW;
    *tile_h = kHaloTileH;
    return;
  }
  int best_w = 16, best_tiles = 1 << 30;
  const int cands[5] = {128, 64, 32, 16, 8};
  for (int tw : cands) {
    const int th = 128 / tw;
    const int tiles = ((Wout + tw - 1) / tw) * ((Hout + th - 1) / th);
    if (tiles < best_tiles) { best_tiles = tiles; best_w = tw; }
  }
  *tile_w = best_w;
  *tile_h = 128 / best_w;
}

int fill_linear_params(GemmParams* p, const bf16* a, const bf16* w, int M, int N, int K, int block_n, int splits,
                       int stages, const bf16* a2, int K2) {
  // a2 (optional): second row-major operand [M, K2]; the weight is [N, K + K2] (K concatenation)
  memset(p, 0, sizeof(*p));
  if (K % 64 != 0 || M <= 0 || N <= 0 || (a2 && (K2 % 64 != 0 || K2 <= 0))) {
    set_error("linear: need K %% 64 == 0 (got M=%d N=%d K=%d K2=%d)", M, N, K, K2);
    return MGB_ERR_INVALID;
  }
  if (!a2) K2 = 0;
  p->mode = 0;
  p->M = M; p->N = N;
  p->num_kb1 = K / 64;
  p->num_kb = (K + K2) / 64;
  if (splits < 1) splits = 1;
  splits = std::min(splits, p->num_kb);
  p->kb_per_split = (p->num_kb + splits - 1) / splits;
  p->stages = stages;
  int rc = make_tmap_2d(&p->tmap_a, a, uint64_t(K), uint64_t(M), uint64_t(K) * 2, 64, 128);
  if (rc) return rc;
  if (a2) {
    rc = make_tmap_2d(&p->tmap_a2, a2, uint64_t(K2), uint64_t(M), uint64_t(K2) * 2, 64, 128);
    if (rc) return rc;
  }
  rc = make_tmap_2d(&p->tmap_b, w, uint64_t(K + K2), uint64_t(N), uint64_t(K + K2) * 2, 64, uint32_t(block_n));
  return rc;
}

int fill_conv_params(GemmParams* p, const bf16* x, const bf16* w, int NB, int Hout, int Wout, int Cin, int Cout,
                     int kind, int block_n, int splits, int stages, int Hsrc, int Wsrc, const bf16* x2, int Cin2) {
  // x2 (optional): bf16 NHWC [NB, Hout, Wout, Cin2], the operand of a 1x1 convolution over the same output pixels whose
  // weight columns follow the taps in w ([Cout, ntaps * Cin + Cin2]): K concatenation (stride-1 kinds only)
  // Hsrc x Wsrc: spatial extent of the tensor the taps address (the image for stride 1; one parity plane for
  // stride 2, i.e. ceil(Hin / 2) x ceil(Win / 2), which exceeds the output by one for the VAE's pad-(0,1,0,1) conv
  // on an odd input). <= 0: same as the output.
  if (Hsrc <= 0) Hsrc = Hout;
  if (Wsrc <= 0) Wsrc = Wout;
  memset(p, 0, sizeof(*p));
  if (Cin % 64 != 0 || NB <= 0 || Hout <= 0 || Wout <= 0) {
    set_error("conv2d: need Cin %% 64 == 0 (got NB=%d H=%d W=%d Cin=%d)", NB, Hout, Wout, Cin);
    return MGB_ERR_INVALID;
  }
  p->mode = 1;
  p->M = NB * Hout * Wout;
  p->N = Cout;
  p->H = Hout; p->W = Wout;
  conv_tile_shape(Hout, Wout, &p->tile_w, &p->tile_h, kind);
  p->tile_w_shift = 0;
  while ((1 << p->tile_w_shift) < p->tile_w) ++p->tile_w_shift;
  p->tiles_x = (Wout + p->tile_w - 1) / p->tile_w;
  p->tiles_y = (Hout + p->tile_h - 1) / p->tile_h;
  p->cblocks = Cin / 64;
  int planes = 1;
  if (kind == 0) {
    p->ntaps = 9;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const int t = kh * 3 + kw;
        p->tap_p[t] = 0; p->tap_dy[t] = int8_t(kh - 1); p->tap_dx[t] = int8_t(kw - 1);
      }
  } else if (kind == 1) {
    p->ntaps = 1;
  } else if (kind == 2 || kind == 3) {
    // stride-2 over the 4 parity planes p = (h & 1) * 2 + (w & 1) of the input
    //   kind 2 (pad 1):          input row 2*oh + kh - 1 -> kh=0: (odd, -1)  kh=1: (even, 0)  kh=2: (odd, 0)
    //   kind 3 (pad (0,1,0,1)):  input row 2*oh + kh     -> kh=0: (even, 0)  kh=1: (odd, 0)   kh=2: (even, +1)
    planes = 4;
    p->ntaps = 9;
    const int par2[3] = {1, 0, 1}, off2[3] = {-1, 0, 0};
    const int par3[3] = {0, 1, 0}, off3[3] = {0, 0, 1};
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const int t = kh * 3 + kw;
        const int ph = kind == 2 ? par2[kh] : par3[kh], pw = kind == 2 ? par2[kw] : par3[kw];
        p->tap_p[t] = int8_t(ph * 2 + pw);
        p->tap_dy[t] = int8_t(kind == 2 ? off2[kh] : off3[kh]);
        p->tap_dx[t] = int8_t(kind == 2 ? off2[kw] : off3[kw]);
      }
  } else {
    set_error("conv2d: unknown kind %d", kind);
    return MGB_ERR_INVALID;
  }
  p->num_kb1 = p->ntaps * p->cblocks;
  if (x2 != nullptr && (Cin2 % 64 != 0 || Cin2 <= 0 || (kind != 0 && kind != 1))) {
    set_error("conv2d: second operand needs Cin2 %% 64 == 0 and a stride-1 kind (got %d, kind %d)", Cin2, kind);
    return MGB_ERR_INVALID;
  }
  p->num_kb = p->num_kb1 + (x2 ? Cin2 / 64 : 0);
  if (splits < 1) splits = 1;
  splits = std::min(splits, p->num_kb);
  p->kb_per_split = (p->num_kb + splits - 1) / splits;
  p->stages = stages;
  const int hv = (kind == 0 && !x2) ? conv_halo_variant() : 0;
  uint32_t box_w = uint32_t(p->tile_w), box_h = uint32_t(p->tile_h);
  if (hv > 0) {
    // operand-reuse path: K order (channel block, tap), splits in whole channel blocks
    p->mode = 2;
    splits = std::min(splits, p->cblocks);
    p->kb_per_split = 9 * ((p->cblocks + splits - 1) / splits);
    p->halo_copies = hv == 3 ? 3 : 1;
    p->halo_w = hv == 3 ? p->tile_w : p->tile_w + 2;
    p->halo_copy_bytes = halo_copy_bytes(hv);
    p->halo_slot_bytes = halo_slot_bytes(hv);
    p->halo_slots = kHaloSlots;
    p->halo_base_off = hv == 2;
    box_w = uint32_t(p->halo_w);
    box_h = uint32_t(p->tile_h + 2);
  }

  const uint64_t C2 = uint64_t(Cin) * 2;
  const uint64_t dims[5] = {uint64_t(Cin), uint64_t(Wsrc), uint64_t(Hsrc), uint64_t(planes), uint64_t(NB)};
  const uint64_t strides[4] = {C2, C2 * Wsrc, C2 * Wsrc * Hsrc, C2 * Wsrc * Hsrc * planes};
  const uint32_t box[5] = {64, box_w, box_h, 1, 1};
  int rc = make_tmap_5d(&p->tmap_a, x, dims, strides, box);
  if (rc) return rc;
  if (x2) {
    const uint64_t C22 = uint64_t(Cin2) * 2;
    const uint64_t dims2[5] = {uint64_t(Cin2), uint64_t(Wout), uint64_t(Hout), 1, uint64_t(NB)};
    const uint64_t strides2[4] = {C22, C22 * Wout, C22 * Wout * Hout, C22 * Wout * Hout};
    const uint32_t box2[5] = {64, uint32_t(p->tile_w), uint32_t(p->tile_h), 1, 1};
    rc = make_tmap_5d(&p->tmap_a2, x2, dims2, strides2, box2);
    if (rc) return rc;
  }
  const uint64_t Ktot = uint64_t(p->ntaps) * Cin + (x2 ? uint64_t(Cin2) : 0);
  rc = make_tmap_2d(&p->tmap_b, w, Ktot, uint64_t(Cout), Ktot * 2, 64, uint32_t(block_n));
  return rc;
}

int effective_splits(const GemmParams& p) { return (p.num_kb + p.kb_per_split - 1) / p.kb_per_split; }

int run_gemm(GemmParams& p, int block_n, float* splitk_ws, cudaStream_t stream) {
  const int splits = effective_splits(p);
  if ((p.epi.flags & EPI_GEGLU) && (block_n % 64 != 0)) {
    set_error("GEGLU epilogue needs block_n %% 64 == 0 (got %d)", block_n);
    return MGB_ERR_INVALID;
  }
  // Multi-wave grids run two CTAs per SM (gemm_tc.cu, MINB = 2): shallow operand rings of <= 113 KB, the epilogue of
  // one tile under the K loop of its neighbour. Single-wave grids keep one CTA per SM with a deep ring.
  static const int two_cta_env = getenv("MGB_GEMM_2CTA") ? atoi(getenv("MGB_GEMM_2CTA")) : 1;
  int ctas_per_sm = 1;
  {
    const long long m_tiles = p.mode == 0 ? (p.M + 127) / 128 : (long long)(p.M / (p.H * p.W)) * p.tiles_x * p.tiles_y;
    const long long ctas = m_tiles * ((p.N + block_n - 1) / block_n) * splits;
    if (two_cta_env && block_n >= 64 && p.mode != 2 && ctas > 148) {
      const int stage_bytes = 16384 + block_n * 128;
      const int st2 = std::min(p.stages, (113 * 1024 - 1280) / stage_bytes);
      const int need = 8 * ((p.epi.flags & EPI_GEGLU) ? 9216 : 4608);
      if (st2 >= 2 && st2 * stage_bytes >= need) { p.stages = st2; ctas_per_sm = 2; }
    }
    // experiment switch (MGB_GEMM_2CTA=2): single-wave grids keep their deep ring but run the 96-register binary, so that
    // small successor kernels launched early (PDL) can become resident beside the tail of this one
    if (two_cta_env == 2 && ctas_per_sm == 1 && block_n >= 64 && p.mode != 2) ctas_per_sm = -2;
  }
  const int kernel_minb = ctas_per_sm == 1 ? 1 : 2;
  if (ctas_per_sm == -2) ctas_per_sm = 1;
  if (block_n > 16 && ctas_per_sm == 1) {
    // the drained operand ring doubles as the epilogue's transpose scratch (8 warps x 4.5 KB, x2 for GEGLU): deepen
    // the pipeline until that fits
    const int need = 8 * ((p.epi.flags & EPI_GEGLU) ? 9216 : 4608);
    for (;;) {
      const int ring = (p.mode == 2 ? p.halo_slots * p.halo_slot_bytes : p.stages * 16384) + p.stages * block_n * 128;
      if (ring >= need || p.stages >= 16) break;
      ++p.stages;
    }
  }
  if (p.stages < 2 || p.stages > 8 ||
      gemm_smem_bytes(block_n, p.stages, p.mode == 2 ? p.halo_slots * p.halo_slot_bytes : -1) > 227 * 1024) {
    set_error("gemm: stages=%d does not fit shared memory for block_n=%d", p.stages, block_n);
    return MGB_ERR_INVALID;
  }
  if (splits > 1) {
    if (!splitk_ws) {
      set_error("split-K requested without a workspace");
      return MGB_ERR_INVALID;
    }
    if (p.epi.flags & (EPI_SCHED | EPI_DEPTH | EPI_NORMALS | EPI_NCHW | EPI_GEGLU)) {
      set_error("split-K is not supported with GEGLU or the small-N special epilogues");
      return MGB_ERR_INVALID;
    }
    p.partial = splitk_ws;
  } else {
    p.partial = nullptr;
  }
  int e = launch_gemm_tc(p, block_n, splits, kernel_minb, stream);
  if (e) {
    set_error("gemm_tc launch failed: %s", cudaGetErrorString(cudaError_t(e)));
    return MGB_ERR_CUDA;
  }
  count_launch(1);
  if (splits > 1) {
    e = launch_splitk_epilogue(p, block_n, splits, stream);
    if (e) {
      set_error("splitk epilogue launch failed: %s", cudaGetErrorString(cudaError_t(e)));
      return MGB_ERR_CUDA;
    }
    count_launch(1);
  }
  return MGB_OK;
}

static int tile_model() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MGB_TILE_MODEL"); v = e ? atoi(e) : 1; }
  return v;
}

// Tile-shape heuristic. Cost model (cycles): per CTA  num_kb * 2*BN (tcgen05 floor at M=128)
// + epilogue ~ 6*BN + fixed 3000; CTAs run in waves of 148 (1 CTA/SM).
void choose_tile(int m_tiles, int N, int num_kb, bool geglu, bool allow_split, int* block_n, int* splits,
                 int* stages, int a_ring_bytes) {
  const int cands[6] = {256, 160, 128, 64, 32, 16};
  double best = 1e30;
  int bbn = 128, bsp = 1;
  for (int bn : cands) {
    if (geglu && (bn % 64 != 0)) continue;
    if (bn > 64 && N < bn / 2 + 1) continue;  // mostly padding
    if (bn < 64 && N >= 64) continue;         // 16 / 32 wide tiles are for the tiny heads only (N <= 32)
    const int n_tiles = (N + bn - 1) / bn;
    const double waste = double(n_tiles) * bn / N;  // MMA work on padded columns is still paid
    (void)waste;
    for (int sp = 1; sp <= (allow_split ? 16 : 1); ++sp) {
      if (sp > 1 && num_kb / sp < 4) break;
      if (a_ring_bytes > 0 && sp > num_kb / 9) break;
      const long long ctas = (long long)m_tiles * n_tiles * sp;
      const long long waves = (ctas + 147) / 148;
      const int kb = a_ring_bytes > 0 ? 9 * ((num_kb / 9 + sp - 1) / sp) : (num_kb + sp - 1) / sp;
      double t;
      if (tile_model() == 0) {
        double cta_cycles = double(kb) * 2.0 * bn + 6.0 * bn + 3000.0;
        // small tiles are smem-bandwidth bound: A (16 KB) + B per k-block at 128 B/cycle
        const double smem_cycles = double(kb) * ((a_ring_bytes > 0 ? 2560.0 : 16384.0) + bn * 128.0) / 128.0 + 6.0 * bn + 3000.0;
        cta_cycles = std::max(cta_cycles, smem_cycles);
        t = waves * cta_cycles;
        if (sp > 1) t += 4000.0 + double(m_tiles) * 128.0 * N * sp * 4.0 / (148.0 * 64.0);  // reduce pass
      } else {
        // constants measured with tools/conv_phases.py (r01): K block = max(tcgen05 floor 2*BN + 40, issue / smem floor
        // ~260) cycles; epilogue = ceil(BN / 64) * 2000 (8 warps, 2000 cycles per 32-column chunk per warp);
        // prologue + first operand latency 3300; a split-K reduce launch costs ~9000 cycles + its traffic
        const double per_kb = std::max(2.0 * bn + 40.0, 260.0);
        const double cta_cycles = double(kb) * per_kb + double((bn + 63) / 64) * 2000.0 + 3300.0;
        t = waves * cta_cycles;
        if (sp > 1) t += 9000.0 + double(m_tiles) * 128.0 * N * sp * 4.0 / (148.0 * 64.0);
      }
      if (t < best) { best = t; bbn = bn; bsp = sp; }
    }
  }
  *block_n = bbn;
  *splits = bsp;
  const int stage_bytes = (a_ring_bytes > 0 ? 0 : 16384) + bbn * 128;
  const int kb = (num_kb + bsp - 1) / bsp;
  // Deep pipelines (the whole 200 KB) only pay off for long K loops. Short-K GEMMs are dominated by
  // prologue / first-load / epilogue latency: cap them near 110 KB so that the NEXT kernel's CTA (launched
  // early through PDL) can become resident on the same SM and overlap its prologue and first operand loads
  // with this kernel's epilogue.
  const int budget = (kb <= 12 ? 110 * 1024 : 200 * 1024) - a_ring_bytes;
  int st = int((budget - 2048) / stage_bytes);
  st = std::max(2, std::min(st, 8));
  st = std::min(st, std::max(2, kb));
  *stages = st;
}

}  // namespace mgb
