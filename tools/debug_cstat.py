import ctypes as C, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import torch
from marigold_b200 import _lib, ops
lib = _lib.load(); raw = C.CDLL(str(_lib.lib_path())); raw.mgb_debug_gemm_cstat.argtypes = [C.c_void_p, C.c_int]
for (NB, H, W, Cin, Cout, bn, splits) in [(1, 16, 16, 64, 64, 64, 0), (2, 16, 16, 64, 128, 128, 0), (1, 96, 96, 320, 320, 160, 0), (1, 24, 24, 256, 256, 128, 3)]:
    x = torch.randn(NB, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = ops.pack_conv_weight((torch.randn(Cout, Cin, 3, 3, device="cuda") / (9 * Cin) ** 0.5).to(torch.bfloat16))
    b = torch.randn(Cout, device="cuda")
    cs = torch.zeros(NB, Cout, 2, device="cuda")
    ws = torch.empty(8 * NB * H * W * Cout, device="cuda") if splits else None
    raw.mgb_debug_gemm_cstat(C.c_void_p(cs.data_ptr()), H * W)
    of, _ = ops.conv2d(x, w, b, NB, H, W, Cin, Cout, block_n=bn, splits=splits, ws=ws)
    torch.cuda.synchronize()
    raw.mgb_debug_gemm_cstat(None, 0)
    ref = torch.stack([of.reshape(NB, H * W, Cout).sum(1), (of.reshape(NB, H * W, Cout) ** 2).sum(1)], dim=-1)
    err = ((cs - ref).abs() / (ref.abs() + 1)).max().item()
    print((NB, H, W, Cin, Cout, bn, splits), "cstat max rel err", err, "nonzero", (cs != 0).float().mean().item(), flush=True)
