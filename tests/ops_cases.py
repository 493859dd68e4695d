"""Operator-level parity cases at PRODUCTION shapes: every case calls ONE C-ABI operator (through marigold_b200.ops)
and compares it with a plain torch fp32 computation on the same bf16-rounded inputs. Used by tests/test_ops_gpu.py
(driver-run, `-m gpu`) and by tools/bringup.py (crash-isolating battery with timings).

Covers the tile shapes the SD-2-size network actually launches (gemm_tc_kernel<160>/<128>/<256>, split-K + its
deferred epilogue, split-KV flash attention + attn_combine at T = 9216, the space-to-depth stride-2 convs, the
small-N special epilogues), which the tiny-model graph tests do not reach."""
from __future__ import annotations


def cases():
    cases = []

    def add(name, fn, **kw):
        cases.append((name, fn, kw))

    # ---- linear -----------------------------------------------------------------------------
    for bn in (128, 64, 256, 160, 32, 16):
        add(f"linear_bn{bn}_small", case_linear, M=256, N=320 if bn == 160 else 256, K=128, block_n=bn)
    add("linear_qkv_96", case_linear, M=9216, N=960, K=320, block_n=160)
    add("linear_auto_96", case_linear, M=9216, N=320, K=1280, block_n=0)
    add("linear_bias_res_bf16", case_linear, M=2304, N=640, K=640, block_n=128, bias=True, residual=True, bf16out=True)
    add("linear_ragged_m144", case_linear, M=144, N=1280, K=1280, block_n=128, bias=True)
    add("linear_ragged_m576", case_linear, M=576, N=1280, K=1280, block_n=256, bias=True, residual=True)
    add("linear_geglu", case_linear, M=2304, N=5120, K=640, block_n=256, bias=True, geglu=True, bf16out=True)
    add("linear_geglu_bn128", case_linear, M=300, N=512, K=128, block_n=128, bias=True, geglu=True)
    add("linear_splitk4", case_linear, M=144, N=1280, K=5120, block_n=128, splits=4, bias=True, residual=True)
    add("linear_stages2", case_linear, M=512, N=256, K=1024, block_n=128, stages=2)
    add("linear_silu_scale", case_linear, M=256, N=128, K=256, block_n=128, bias=True, silu=True)
    add("linear_n_ragged", case_linear, M=256, N=200, K=128, block_n=128, bias=True)
    # ---- conv -------------------------------------------------------------------------------
    add("conv3_16x16_c64", case_conv, NB=1, H=16, W=16, Cin=64, Cout=64, kind=0, block_n=64)
    add("conv3_96_c320", case_conv, NB=1, H=96, W=96, Cin=320, Cout=320, kind=0, block_n=160, bias=True)
    add("conv3_24_nb2", case_conv, NB=2, H=24, W=24, Cin=128, Cout=256, kind=0, block_n=128, bias=True, residual=True)
    add("conv3_12_splitk", case_conv, NB=1, H=12, W=12, Cin=1280, Cout=1280, kind=0, block_n=128, splits=6, bias=True)
    add("conv3_rect_40x72", case_conv, NB=1, H=40, W=72, Cin=64, Cout=128, kind=0, block_n=128, bias=True)
    add("conv3_s2_pad1", case_conv, NB=2, H=24, W=24, Cin=128, Cout=128, kind=2, block_n=128, bias=True)
    add("conv3_s2_asym", case_conv, NB=1, H=48, W=48, Cin=128, Cout=128, kind=3, block_n=128, bias=True)
    add("conv3_cout4_nchw", case_conv, NB=2, H=32, W=32, Cin=64, Cout=4, kind=0, block_n=16, bias=True, special="nchw")
    add("conv3_cout3_depth", case_conv, NB=2, H=32, W=32, Cin=128, Cout=3, kind=0, block_n=16, bias=True,
        special="depth")
    add("conv3_cout3_normals", case_conv, NB=1, H=32, W=32, Cin=128, Cout=3, kind=0, block_n=16, bias=True,
        special="normals")
    add("conv3_auto_48", case_conv, NB=1, H=48, W=48, Cin=640, Cout=640, kind=0, block_n=0, bias=True)
    # ---- attention --------------------------------------------------------------------------
    add("attn_t128_h1", case_attn, NB=1, T=128, C=64)
    add("attn_t256_h2", case_attn, NB=1, T=256, C=128)
    add("attn_t576_nb2", case_attn, NB=2, T=576, C=128)
    add("attn_t144", case_attn, NB=1, T=144, C=1280)
    add("attn_t2304", case_attn, NB=1, T=2304, C=640)
    add("attn_t9216", case_attn, NB=1, T=9216, C=320)
    # ---- streaming kernels ------------------------------------------------------------------
    add("groupnorm_320", case_groupnorm, NB=2, HW=2304, C=320, G=32, eps=1e-5, silu=1)
    add("groupnorm_1920", case_groupnorm, NB=1, HW=576, C=1920, G=32, eps=1e-5, silu=1)
    add("groupnorm_2560", case_groupnorm, NB=1, HW=144, C=2560, G=32, eps=1e-6, silu=0)
    add("groupnorm_128_big", case_groupnorm, NB=1, HW=147456, C=128, G=32, eps=1e-6, silu=1)
    # collapsed cross-attention (+ norm2 / norm3): one warp per token at three register sizes, four warps per token
    add("xattn2_c320", case_xattn2, M=9216, C=320)
    add("xattn2_c640", case_xattn2, M=2304, C=640)
    add("xattn2_c1280_wide", case_xattn2, M=576, C=1280)
    add("xattn2_c1280_wide_odd", case_xattn2, M=145, C=1280)
    add("xattn2_c1280_batched", case_xattn2, M=2 * 576, C=1280)
    add("layernorm_320", case_layernorm, M=9216, C=320)
    add("layernorm_1280", case_layernorm, M=576, C=1280)
    add("s2d", case_s2d, NB=2, H=24, W=16, C=128)
    add("s2d_odd", case_s2d, NB=1, H=27, W=13, C=64)
    add("upsample", case_upsample, NB=2, H=12, W=8, C=64)
    return cases


# -------------------------------------------------------------------------------------------------
def _timeit(fn, iters=10):
    import torch

    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _err(out, ref):
    out, ref = out.float(), ref.float()
    d = (out - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    return {"max_abs": d.max().item(), "rel_to_max": d.max().item() / scale, "ref_max": scale,
            "mean_abs": d.mean().item(), "nan": bool(torch_isnan(out))}


def torch_isnan(t):
    import torch

    return torch.isnan(t).any().item()


def case_linear(M, N, K, block_n, bias=False, residual=False, bf16out=False, geglu=False, splits=0, stages=0,
                silu=False):
    import torch
    from marigold_b200 import _lib, ops

    g = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(M, K, device="cuda", generator=g)).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    n_out = N // 2 if geglu else N
    r = torch.randn(M, n_out, device="cuda", generator=g) if residual else None
    flags = (_lib.EPI_GEGLU if geglu else 0) | (_lib.EPI_SILU if silu else 0)
    ws = torch.empty(max(splits, 16) * M * N, device="cuda") if (splits > 1 or block_n == 0) else None
    # reference
    ref = a.float() @ w.float().t()
    w_used, b_used = w, b
    if geglu:
        # library expects [value | gate] interleaved per block_n tile; emulate finalize_weights' packing
        bn = block_n
        half = bn // 2
        nt = N // bn
        val_rows = torch.arange(N // 2, device="cuda").reshape(nt, half)
        gate_rows = val_rows + N // 2
        perm = torch.cat([val_rows, gate_rows], dim=1).reshape(-1)
        w_used = w[perm].contiguous()
        b_used = b[perm].contiguous() if b is not None else None
        full = ref + (b if b is not None else 0)
        ref = full[:, : N // 2] * torch.nn.functional.gelu(full[:, N // 2:])
    else:
        if b is not None:
            ref = ref + b
        if silu:
            ref = torch.nn.functional.silu(ref)
    if r is not None:
        ref = ref + r
    run = lambda: ops.linear(a, w_used, b_used, r, out_f32=True, out_bf16=bf16out, flags=flags, block_n=block_n,
                             splits=splits, stages=stages, ws=ws)
    of, ob = run()
    torch.cuda.synchronize()
    res = {"f32": _err(of, ref)}
    if ob is not None:
        res["bf16"] = _err(ob, ref)
    res["ms"] = _timeit(run)
    res["tflops"] = 2.0 * M * N * K / res["ms"] / 1e9
    tol = 2e-3 if not geglu else 4e-3
    res["ok"] = (res["f32"]["rel_to_max"] < tol) and not res["f32"]["nan"]
    return res


def case_conv(NB, H, W, Cin, Cout, kind, block_n, bias=False, residual=False, splits=0, special=None):
    import torch
    import torch.nn.functional as F
    from marigold_b200 import _lib, ops

    g = torch.Generator(device="cuda").manual_seed(2)
    stride = 2 if kind in (2, 3) else 1
    Hin, Win = H * stride, W * stride
    x = torch.randn(NB, Hin, Win, Cin, device="cuda", generator=g)  # NHWC fp32
    wt = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (9 * Cin) ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g) if bias else None
    r = torch.randn(NB, H, W, Cout, device="cuda", generator=g) if residual else None
    xb = x.to(torch.bfloat16)
    wb = wt.to(torch.bfloat16)
    x_nchw = xb.float().permute(0, 3, 1, 2)
    if kind == 0:
        ref = F.conv2d(x_nchw, wb.float(), b, stride=1, padding=1)
        x_in = xb.contiguous()
    elif kind == 2:
        ref = F.conv2d(x_nchw, wb.float(), b, stride=2, padding=1)
        x_in = ops.space_to_depth(x)
    else:
        ref = F.conv2d(F.pad(x_nchw, (0, 1, 0, 1)), wb.float(), b, stride=2, padding=0)
        x_in = ops.space_to_depth(x)
    flags = 0
    ref_out = ref.permute(0, 2, 3, 1)
    if r is not None:
        ref_out = ref_out + r
    if special == "nchw":
        flags = _lib.EPI_NCHW
        ref_out = ref  # NCHW
    elif special == "depth":
        flags = _lib.EPI_DEPTH
        ref_out = (ref.mean(dim=1, keepdim=True).clip(-1, 1) + 1) / 2
    elif special == "normals":
        flags = _lib.EPI_NORMALS
        c = ref.clip(-1, 1)
        ref_out = c / torch.norm(c, dim=1, keepdim=True).clamp(min=1e-6)
    wp = ops.pack_conv_weight(wb)
    ws = torch.empty(max(splits, 16) * NB * H * W * Cout, device="cuda") if (splits > 1 or block_n == 0) else None
    run = lambda: ops.conv2d(x_in, wp, b, NB, H, W, Cin, Cout, kind=kind, residual=r, flags=flags, block_n=block_n,
                             splits=splits, ws=ws)
    of, _ = run()
    torch.cuda.synchronize()
    if special == "nchw":
        out = of.reshape(-1)[: ref_out.numel()].reshape(ref_out.shape)
    elif special == "depth":
        out = of.reshape(-1)[: ref_out.numel()].reshape(ref_out.shape)
    elif special == "normals":
        out = of.reshape(-1)[: ref_out.numel()].reshape(ref_out.shape)
    else:
        out = of
    res = {"f32": _err(out, ref_out)}
    res["ms"] = _timeit(run)
    res["tflops"] = 2.0 * NB * H * W * Cout * Cin * 9 / res["ms"] / 1e9
    res["ok"] = res["f32"]["rel_to_max"] < 3e-3 and not res["f32"]["nan"]
    return res


def case_attn(NB, T, C):
    import torch
    import torch.nn.functional as F
    from marigold_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = torch.randn(NB * T, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
    heads = C // 64
    q, k, v = [t.float().reshape(NB, T, heads, 64).permute(0, 2, 1, 3) for t in qkv.split(C, dim=1)]
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(NB * T, C)
    run = lambda: ops.flash_attn64(qkv, NB, T, C, 0.125)
    out = run()
    torch.cuda.synchronize()
    res = {"bf16": _err(out, ref)}
    res["ms"] = _timeit(run)
    res["tflops"] = 4.0 * NB * heads * T * T * 64 / res["ms"] / 1e9
    res["ok"] = res["bf16"]["rel_to_max"] < 2e-2 and not res["bf16"]["nan"]
    return res


def case_groupnorm(NB, HW, C, G, eps, silu):
    import torch
    import torch.nn.functional as F
    from marigold_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(NB, HW, C, device="cuda", generator=g) * 2 + 0.5
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    ref = F.group_norm(x.permute(0, 2, 1), G, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1)
    run = lambda: ops.groupnorm(x, gamma, beta, NB, HW, C, G, eps, silu)
    out = run()
    torch.cuda.synchronize()
    res = {"bf16": _err(out, ref)}
    res["ms"] = _timeit(run)
    res["gbs"] = NB * HW * C * (4 + 4 + 2) / res["ms"] / 1e6
    res["ok"] = res["bf16"]["rel_to_max"] < 6e-3 and not res["bf16"]["nan"]
    return res


def case_xattn2(M, C):
    import torch
    import torch.nn.functional as F
    from marigold_b200 import ops

    H = C // 64
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(M, C, device="cuda", generator=g) * 1.5 + 0.2
    p = [torch.randn(C, device="cuda", generator=g) * s + o for s, o in ((0.3, 1.0), (0.3, 0.0), (0.3, 1.0), (0.3, 0.0))]
    G = (torch.randn(H, C, device="cuda", generator=g) / C ** 0.5 * 4).to(torch.bfloat16)
    U = (torch.randn(H, C, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    c1 = torch.randn(C, device="cuda", generator=g) * 0.5
    GU = torch.stack([G, U]).contiguous()
    z = F.layer_norm(x.double(), (C,), p[0].double(), p[1].double(), 1e-5)
    w = torch.sigmoid(0.125 * z @ G.double().t())
    yf = x.double() + c1.double() + w @ U.double()
    af = F.layer_norm(yf, (C,), p[2].double(), p[3].double(), 1e-5)
    run = lambda: ops.xattn2(x, p[0], p[1], p[2], p[3], GU, c1, H, 0.125)
    y, a = run()
    torch.cuda.synchronize()
    res = {"y": _err(y, yf.float()), "a": _err(a, af.float())}
    res["ms"] = _timeit(run)
    res["ok"] = res["y"]["rel_to_max"] < 6e-3 and res["a"]["rel_to_max"] < 6e-3 and not res["y"]["nan"] and not res["a"]["nan"]
    return res


def case_layernorm(M, C):
    import torch
    import torch.nn.functional as F
    from marigold_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(M, C, device="cuda", generator=g) * 3 - 1
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    ref = F.layer_norm(x, (C,), gamma, beta, 1e-5)
    run = lambda: ops.layernorm(x, gamma, beta)
    out = run()
    torch.cuda.synchronize()
    res = {"bf16": _err(out, ref)}
    res["ms"] = _timeit(run)
    res["ok"] = res["bf16"]["rel_to_max"] < 6e-3 and not res["bf16"]["nan"]
    return res


def case_s2d(NB, H, W, C):
    import torch
    from marigold_b200 import ops

    x = torch.randn(NB, H, W, C, device="cuda")
    out = ops.space_to_depth(x)
    # odd sizes: planes hold ceil(H/2) x ceil(W/2) entries, zero where the source pixel does not exist
    xp = torch.nn.functional.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    ref = torch.stack([xp[:, a::2, b::2] for a in (0, 1) for b in (0, 1)], dim=1).to(torch.bfloat16)
    return {"ok": bool(torch.equal(out, ref))}


def case_upsample(NB, H, W, C):
    import torch
    from marigold_b200 import ops

    x = torch.randn(NB, H, W, C, device="cuda")
    out = ops.upsample2x(x)
    ref = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).to(torch.bfloat16)
    return {"ok": bool(torch.equal(out, ref))}


