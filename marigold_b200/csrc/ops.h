// Host-side operator builders (ops.cu).
#pragma once
#include <atomic>

#include "kernels.h"

namespace mgb {

void count_launch(int n);
long long launch_count();

void conv_tile_shape(int Hout, int Wout, int* tile_w, int* tile_h);

// A [M, K] bf16 row-major, W [N, K] bf16 row-major.
int fill_linear_params(GemmParams* p, const bf16* a, const bf16* w, int M, int N, int K, int block_n, int splits,
                       int stages);
// x NHWC bf16 ([NB, Hout, Wout, Cin] for kind 0/1, [NB, 4, Hout, Wout, Cin] parity planes for kind 2/3);
// w [Cout, taps * Cin] bf16 tap-major.
int fill_conv_params(GemmParams* p, const bf16* x, const bf16* w, int NB, int Hout, int Wout, int Cin, int Cout,
                     int kind, int block_n, int splits, int stages);
int effective_splits(const GemmParams& p);
// Launch (plus the deferred epilogue when split-K is active). p.epi must be filled by the caller.
int run_gemm(GemmParams& p, int block_n, float* splitk_ws, cudaStream_t stream);
void choose_tile(int m_tiles, int N, int num_kb, bool geglu, bool allow_split, int* block_n, int* splits,
                 int* stages);

}  // namespace mgb
