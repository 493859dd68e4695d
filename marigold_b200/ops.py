"""Operator-level Python wrappers over the C ABI's mgb_op_* entry points (used by layer parity tests
and by tools/bringup.py). Tensors are torch CUDA tensors; layouts are the library's internal ones
(NHWC / token-major, bf16 operands, fp32 trunk)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def pack_conv_weight(w: torch.Tensor, cin_pad: int | None = None) -> torch.Tensor:
    """[Cout, Cin, kh, kw] (PyTorch) -> bf16 [Cout, kh*kw*Cin_pad] tap-major (tap = kh*3 + kw)."""
    cout, cin, kh, kw = w.shape
    cp = cin_pad or cin
    out = torch.zeros(cout, kh * kw, cp, dtype=torch.float32, device=w.device)
    out[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin).float()
    return out.reshape(cout, kh * kw * cp).to(torch.bfloat16).contiguous()


def linear(a, w, bias=None, residual=None, out_f32=True, out_bf16=False, flags=0, block_n=0, splits=0, stages=0,
           ws=None):
    lib = _lib.load()
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if flags & _lib.EPI_GEGLU else N
    of = torch.empty(M, n_out, dtype=torch.float32, device=a.device) if out_f32 else None
    ob = torch.empty(M, n_out, dtype=torch.bfloat16, device=a.device) if out_bf16 else None
    check(lib.mgb_op_linear(ptr(a), ptr(w), ptr(bias), ptr(residual), ptr(of), ptr(ob), M, N, K, flags, block_n,
                            splits, stages, ptr(ws), stream_ptr()), "mgb_op_linear")
    return of, ob


def conv2d(x, w_packed, bias, NB, Hout, Wout, Cin, Cout, kind=0, residual=None, out_f32=True, out_bf16=False, flags=0,
           block_n=0, splits=0, stages=0, ws=None):
    lib = _lib.load()
    of = torch.empty(NB, Hout, Wout, Cout, dtype=torch.float32, device=x.device) if out_f32 else None
    ob = torch.empty(NB, Hout, Wout, Cout, dtype=torch.bfloat16, device=x.device) if out_bf16 else None
    check(lib.mgb_op_conv2d(ptr(x), ptr(w_packed), ptr(bias), ptr(residual), ptr(of), ptr(ob), NB, Hout, Wout, Cin,
                            Cout, kind, flags, block_n, splits, stages, ptr(ws), stream_ptr()), "mgb_op_conv2d")
    return of, ob


def flash_attn64(qkv, NB, T, C, scale):
    lib = _lib.load()
    out = torch.empty(NB * T, C, dtype=torch.bfloat16, device=qkv.device)
    check(lib.mgb_op_flash_attn64(ptr(qkv), ptr(out), NB, T, C, float(scale), stream_ptr()), "mgb_op_flash_attn64")
    return out


def groupnorm(x, gamma, beta, NB, HW, C, G, eps, silu):
    lib = _lib.load()
    y = torch.empty(NB, HW, C, dtype=torch.bfloat16, device=x.device)
    nbytes = int(lib.mgb_op_groupnorm_ws_bytes(NB, HW, C, G))
    if nbytes == 0:
        raise _lib.MgbError(f"groupnorm: unsupported shape NB={NB} HW={HW} C={C} G={G}")
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=x.device)
    check(lib.mgb_op_groupnorm(ptr(x), ptr(y), ptr(gamma), ptr(beta), ptr(ws), NB, HW, C, G, float(eps), int(silu),
                               stream_ptr()), "mgb_op_groupnorm")
    return y


def xattn2(x, ln2_g, ln2_b, ln3_g, ln3_b, GU, c1, H, scale, eps=1e-5):
    """Collapsed cross-attention against the fixed 2-token context, fused with norm2 / norm3 (see include/marigold_b200.h)."""
    lib = _lib.load()
    M, Cc = x.shape
    y = torch.empty(M, Cc, dtype=torch.bfloat16, device=x.device)
    a = torch.empty(M, Cc, dtype=torch.bfloat16, device=x.device)
    check(lib.mgb_op_xattn2(ptr(x), ptr(y), ptr(a), ptr(ln2_g), ptr(ln2_b), ptr(ln3_g), ptr(ln3_b), ptr(GU), ptr(c1), M, Cc,
                            H, float(scale), float(eps), stream_ptr()), "mgb_op_xattn2")
    return y, a


def layernorm(x, gamma, beta, eps=1e-5):
    lib = _lib.load()
    M, Cc = x.shape
    y = torch.empty(M, Cc, dtype=torch.bfloat16, device=x.device)
    check(lib.mgb_op_layernorm(ptr(x), ptr(y), ptr(gamma), ptr(beta), M, Cc, float(eps), stream_ptr()),
          "mgb_op_layernorm")
    return y


def space_to_depth(x):
    lib = _lib.load()
    NB, H, W, Cc = x.shape
    y = torch.empty(NB, 4, (H + 1) // 2, (W + 1) // 2, Cc, dtype=torch.bfloat16, device=x.device)
    check(lib.mgb_op_space_to_depth(ptr(x), ptr(y), NB, H, W, Cc, stream_ptr()), "mgb_op_space_to_depth")
    return y


def upsample2x(x):
    lib = _lib.load()
    NB, H, W, Cc = x.shape
    y = torch.empty(NB, 2 * H, 2 * W, Cc, dtype=torch.bfloat16, device=x.device)
    check(lib.mgb_op_upsample2x(ptr(x), ptr(y), NB, H, W, Cc, stream_ptr()), "mgb_op_upsample2x")
    return y
