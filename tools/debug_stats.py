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
B, lh, lw = int(sys.argv[1]) if len(sys.argv) > 1 else 1, 16, 16
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
d = (a - b).abs() / (a.abs() + 1e-3)
bad = (d > 1e-3).nonzero().flatten()
print("first mismatching slab indices:", bad[:20].tolist(), "count", bad.numel(), "of", a.numel())
if bad.numel():
    i = int(bad[0])
    print("around first mismatch (kernel-path vs fused):", a[i:i + 8].tolist(), b[i:i + 8].tolist())
nz = (b != 0)
chg = (nz[1:] != nz[:-1]).nonzero().flatten().tolist()
print('fused slab: nonzero fraction', nz.float().mean().item(), 'segment boundaries', chg[:40])
nz0 = (a != 0)
chg0 = (nz0[1:] != nz0[:-1]).nonzero().flatten().tolist()
print('kernel-path slab: nonzero fraction', nz0.float().mean().item(), 'segment boundaries', chg0[:40])
eng.close()
