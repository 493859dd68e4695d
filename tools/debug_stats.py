"""Compare the GroupNorm statistics slab produced by the fused epilogue path with the stats-kernel path."""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from marigold_b200 import _lib  # noqa: E402
from marigold_b200.schedulers import DDIMScheduler  # noqa: E402
from tests.helpers import engine_from_oracle, oracle_models  # noqa: E402

raw = C.CDLL(str(_lib.lib_path()))
raw.mgb_debug_set_fuse_stats.argtypes = [C.c_void_p, C.c_int]
raw.mgb_debug_stat_slab.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
raw.mgb_debug_stat_slab.restype = C.c_void_p
unet, vae, text = oracle_models("tiny")
eng = engine_from_oracle(unet, vae, text)
s = DDIMScheduler(); s.set_timesteps(2)
eng.set_schedule(s.timesteps, *s.coefficients())
g = torch.Generator().manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lh, lw = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (16, 16)
rgb = torch.randn(B, 4, lh, lw, generator=g).cuda(); x = torch.randn(B, 4, lh, lw, generator=g).cuda()
slabs = {}
outs = {}
for mode in (0, 1):
    raw.mgb_debug_set_fuse_stats(eng._h, mode)
    t = x.clone()
    mo = eng.unet_step(rgb, t, 0, want_model_out=True)
    torch.cuda.synchronize()
    n = C.c_size_t()
    p = raw.mgb_debug_stat_slab(eng._h, C.byref(n))
    buf = (C.c_float * n.value).from_address(0)  # placeholder
    host = torch.empty(n.value, dtype=torch.float32)
    torch.cuda.synchronize()
    C.cdll.LoadLibrary("libcudart.so.12") if False else None
    dev = torch.empty(0)
    # copy via torch: wrap the raw pointer
    import numpy as np
    arr = torch.zeros(n.value, dtype=torch.float32, device="cuda")
    C.CDLL("libcudart.so.12").cudaMemcpy(C.c_void_p(arr.data_ptr()), C.c_void_p(p), C.c_size_t(n.value * 4), 3)
    slabs[mode] = arr.cpu()
    outs[mode] = mo.cpu()
    print("mode", mode, "nan in model out:", torch.isnan(mo).sum().item(), "slab floats", n.value)
a, b = slabs[0], slabs[1]
print("model-out rel diff between modes:", ((outs[0]-outs[1]).norm()/outs[0].norm()).item())
# allocation order of unet_forward (tiny config: ch = [64,128,256,256], L = 2)
ch, L = [64, 128, 256, 256], 2
recs = [("conv_in", ch[0])]
for i in range(4):
    for j in range(L):
        recs.append((f"down{i}.res{j}.y", ch[i])); recs.append((f"down{i}.res{j}.h", ch[i]))
        if i < 3:
            recs.append((f"down{i}.xf{j}.y", ch[i]))
    if i < 3:
        recs.append((f"down{i}.ds", ch[i]))
recs += [("mid.res0.y", ch[3]), ("mid.res0.h", ch[3]), ("mid.xf.y", ch[3]), ("mid.res1.y", ch[3]), ("mid.res1.h", ch[3])]
for i in range(4):
    co = ch[3 - i]
    for j in range(L + 1):
        recs.append((f"up{i}.res{j}.y", co)); recs.append((f"up{i}.res{j}.h", co))
        if i > 0:
            recs.append((f"up{i}.xf{j}.y", co))
    if i < 3:
        recs.append((f"up{i}.us", co))
off = 0
nprint = 0
for name, C_ in recs:
    n = B * C_ * 2
    ra, rb = a[off:off + n], b[off:off + n]
    rel = ((ra - rb).abs().max() / (ra.abs().max() + 1e-6)).item() if n else 0
    flag = "" if rel < 1e-3 else "   <-- MISMATCH"
    if (flag or ra.abs().max() == 0) and nprint < 6:
        nprint += 1
        print(f"{name:16s} C={C_:4d} off={off:6d} max_rel={rel:.3e} ref_absmax={ra.abs().max().item():.3g} fused_absmax={rb.abs().max().item():.3g}{flag}")
    off += n
print("slab used", off, "of", a.numel())
# timing of both modes
import time
for mode in (0, 1):
    raw.mgb_debug_set_fuse_stats(eng._h, mode)
    t = x.clone(); eng.unet_step(rgb, t, 0); torch.cuda.synchronize()
eng.close()
