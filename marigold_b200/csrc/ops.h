// Host-side operator builders (ops.cu).
#pragma once
#include <atomic>

#include "kernels.h"

namespace mgb {

void count_launch(int n);
long long launch_count();

// Halo variant used for 3x3 stride-1 convolutions (kind 0): 0 = off (mode 1, default), 1 = one (tile_h+2) x (tile_w+2)
// box, 3 = three dx-shifted (tile_h+2) x tile_w boxes. Env MGB_CONV_HALO. Both are bit-compatible with mode 1.
int conv_halo_variant();
int conv_halo_ring_bytes(int kind);   // A ring bytes of the halo path (0 when the conv kind does not use it)
void conv_tile_shape(int Hout, int Wout, int* tile_w, int* tile_h, int kind = -1);

// A [M, K] bf16 row-major, W [N, K] bf16 row-major.
int fill_linear_params(GemmParams* p, const bf16* a, const bf16* w, int M, int N, int K, int block_n, int splits,
                       int stages, const bf16* a2 = nullptr, int K2 = 0);
// x NHWC bf16 ([NB, Hout, Wout, Cin] for kind 0/1, [NB, 4, Hout, Wout, Cin] parity planes for kind 2/3);
// w [Cout, taps * Cin] bf16 tap-major.
int fill_conv_params(GemmParams* p, const bf16* x, const bf16* w, int NB, int Hout, int Wout, int Cin, int Cout,
                     int kind, int block_n, int splits, int stages, int Hsrc = 0, int Wsrc = 0, const bf16* x2 = nullptr,
                     int Cin2 = 0);
int effective_splits(const GemmParams& p);
// Launch (plus the deferred epilogue when split-K is active). p.epi must be filled by the caller.
int run_gemm(GemmParams& p, int block_n, float* splitk_ws, cudaStream_t stream);
// a_ring_bytes > 0: halo conv (fixed A ring, B-only pipeline stages, split-K in units of 9 K blocks)
void choose_tile(int m_tiles, int N, int num_kb, bool geglu, bool allow_split, int* block_n, int* splits,
                 int* stages, int a_ring_bytes = 0);

}  // namespace mgb
