"""Hugging Face *diffusers* checkpoint directory -> Engine + scheduler + pipeline.

The reference builds its pipelines with `MarigoldDepthPipeline.from_pretrained(checkpoint_path, variant=...,
torch_dtype=...)` (script/depth/run.py:213-222, script/normals/run.py likewise) on the layout the README documents
(README.md:261-290):

    <root>/model_index.json                      default_denoising_steps, default_processing_resolution, ...
    <root>/unet/config.json + diffusion_pytorch_model[.<variant>].safetensors
    <root>/vae/config.json  + diffusion_pytorch_model[.<variant>].safetensors
    <root>/scheduler/scheduler_config.json       DDIMScheduler | LCMScheduler
    <root>/text_encoder, <root>/tokenizer        CLIP, only ever evaluated on the empty prompt

Everything here is host code: it parses the files, checks that the architecture is the one the CUDA library
implements (SD-2 UNet / SD VAE family) and hands the tensors to `mgb_load_tensor` under their diffusers names.
The empty-prompt embedding (marigold_depth_pipeline.py:381-394) is a constant of the checkpoint: it is taken from
`<root>/empty_text_embed.{safetensors,pt,npy}` when present, else computed once on the CPU with `transformers`.
"""
from __future__ import annotations

import json
import os
import struct
import warnings
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .schedulers import DDIMScheduler, LCMScheduler

_ST_DTYPES = {
    "F64": torch.float64, "F32": torch.float32, "F16": torch.float16, "BF16": torch.bfloat16,
    "I64": torch.int64, "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8, "BOOL": torch.bool,
}
_ST_NAMES = {v: k for k, v in _ST_DTYPES.items()}


class CheckpointError(RuntimeError):
    pass


# -------------------------------------------------------------------------------------------------
# safetensors (format: u64 little-endian header length, JSON header, raw little-endian tensor data)
# -------------------------------------------------------------------------------------------------
def read_safetensors(path) -> Dict[str, torch.Tensor]:
    """Zero-copy views into a read-only memory map of `path` (tensors are copied when they are uploaded)."""
    path = Path(path)
    size = path.stat().st_size
    with open(path, "rb") as f:
        head = f.read(8)
        if len(head) != 8:
            raise CheckpointError(f"{path}: not a safetensors file (shorter than its length prefix)")
        (n,) = struct.unpack("<Q", head)
        if n > size - 8 or n > (100 << 20):
            raise CheckpointError(f"{path}: implausible safetensors header length {n}")
        try:
            header = json.loads(f.read(n).decode("utf-8"))
        except (UnicodeDecodeError, json.JSONDecodeError) as e:
            raise CheckpointError(f"{path}: corrupt safetensors header ({e})") from None
    base = 8 + n
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    out: Dict[str, torch.Tensor] = {}
    for name, info in header.items():
        if name == "__metadata__":
            continue
        dt = _ST_DTYPES.get(info["dtype"])
        if dt is None:
            raise CheckpointError(f"{path}: tensor {name!r} has unsupported dtype {info['dtype']}")
        b0, b1 = info["data_offsets"]
        shape = [int(s) for s in info["shape"]]
        nbytes = int(np.prod(shape, dtype=np.int64)) * torch.empty((), dtype=dt).element_size()
        if b1 - b0 != nbytes or base + b1 > size:
            raise CheckpointError(f"{path}: tensor {name!r} data range {b0}:{b1} does not match shape {shape}")
        if nbytes == 0:
            out[name] = torch.empty(shape, dtype=dt)
            continue
        buf = mm[base + b0: base + b1]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)        # read-only buffer: consumers only read / copy
            out[name] = torch.frombuffer(buf, dtype=dt).reshape(shape)
    return out


def write_safetensors(path, tensors: Dict[str, torch.Tensor], metadata: Optional[Dict[str, str]] = None) -> None:
    """Minimal writer (tests and `empty_text_embed.safetensors`)."""
    header, blobs, off = {}, [], 0
    for name in sorted(tensors):
        t = tensors[name].detach().cpu().contiguous()
        if t.dtype not in _ST_NAMES:
            raise CheckpointError(f"cannot store dtype {t.dtype}")
        raw = t.view(torch.uint8).numpy().tobytes() if t.numel() else b""
        header[name] = {"dtype": _ST_NAMES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    if metadata:
        header["__metadata__"] = dict(metadata)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


_LEGACY_ATTN = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


def _modernise_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """SD-era VAE checkpoints name the mid-block attention projections query / key / value / proj_attn (sometimes as
    1x1 convolutions); diffusers renames them on load (`_convert_deprecated_attention_blocks`). Do the same here."""
    out = {}
    for k, v in sd.items():
        parts = k.split(".")
        if "attentions" in parts and len(parts) >= 2 and parts[-2] in _LEGACY_ATTN:
            parts[-2] = _LEGACY_ATTN[parts[-2]]
            k = ".".join(parts)
            if parts[-1] == "weight" and v.dim() == 4 and v.shape[-2:] == (1, 1):
                v = v.reshape(v.shape[0], v.shape[1])
        elif "attentions" in parts and parts[-1] == "weight" and v.dim() == 4 and v.shape[-2:] == (1, 1) and \
                any(p.startswith("to_") for p in parts):
            v = v.reshape(v.shape[0], v.shape[1])
        out[k] = v
    return out


def read_weights(component_dir, variant: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """diffusers' file naming: diffusion_pytorch_model[.<variant>].safetensors, else the .bin pickle."""
    d = Path(component_dir)
    stems = ([f"diffusion_pytorch_model.{variant}"] if variant else []) + ["diffusion_pytorch_model"]
    for stem in stems:
        p = d / f"{stem}.safetensors"
        if p.is_file():
            return _modernise_keys(read_safetensors(p))
    for stem in stems:
        p = d / f"{stem}.bin"
        if p.is_file():
            return _modernise_keys(torch.load(p, map_location="cpu", weights_only=True))
    raise CheckpointError(f"no diffusion_pytorch_model[.{variant or '<variant>'}].safetensors|.bin under {d}")


def _read_json(path) -> dict:
    try:
        with open(path, "r", encoding="utf-8") as f:
            return json.load(f)
    except FileNotFoundError:
        raise CheckpointError(f"missing {path}") from None


# -------------------------------------------------------------------------------------------------
# configs
# -------------------------------------------------------------------------------------------------
def engine_config_from_diffusers(unet_cfg: dict, vae_cfg: dict):
    """Map unet/config.json + vae/config.json to EngineConfig; refuse architectures the kernels do not implement
    instead of loading them wrongly."""
    from .engine import EngineConfig

    def need(cond, msg):
        if not cond:
            raise CheckpointError("unsupported checkpoint architecture: " + msg)

    boc = list(unet_cfg.get("block_out_channels", []))
    need(len(boc) == 4, f"UNet needs 4 resolution levels, got block_out_channels={boc}")
    need(unet_cfg.get("down_block_types", ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"]) ==
         ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"], f"down_block_types={unet_cfg.get('down_block_types')}")
    need(unet_cfg.get("up_block_types", ["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3) ==
         ["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3, f"up_block_types={unet_cfg.get('up_block_types')}")
    need(unet_cfg.get("use_linear_projection", False), "use_linear_projection must be true (SD-2 style transformer blocks)")
    need(unet_cfg.get("act_fn", "silu") == "silu", f"act_fn={unet_cfg.get('act_fn')}")
    need(unet_cfg.get("norm_num_groups", 32) == vae_cfg.get("norm_num_groups", 32), "UNet / VAE group counts differ")
    heads = unet_cfg.get("attention_head_dim", 8)     # SD-2 configs store the NUMBER OF HEADS under this name
    heads = [heads] * 4 if isinstance(heads, int) else list(heads)
    need(all(c % h == 0 and c // h == 64 for c, h in zip(boc, heads)),
         f"self-attention head size must be 64 (block_out_channels={boc}, attention_head_dim={heads})")
    need(unet_cfg.get("transformer_layers_per_block", 1) == 1, "one transformer layer per block")
    need(not unet_cfg.get("class_embed_type") and not unet_cfg.get("addition_embed_type"), "no class / addition embeddings")
    vboc = list(vae_cfg.get("block_out_channels", []))
    need(len(vboc) == 4, f"VAE needs 4 levels, got {vboc}")
    need(vae_cfg.get("in_channels", 3) == 3 and vae_cfg.get("out_channels", 3) == 3, "VAE is RGB in / RGB out")
    return EngineConfig(
        unet_in_channels=int(unet_cfg.get("in_channels", 8)), unet_out_channels=int(unet_cfg.get("out_channels", 4)),
        unet_block_channels=boc, unet_layers_per_block=int(unet_cfg.get("layers_per_block", 2)),
        unet_cross_dim=int(unet_cfg.get("cross_attention_dim", 1024)),
        vae_block_channels=vboc, vae_layers_per_block=int(vae_cfg.get("layers_per_block", 2)),
        vae_latent_channels=int(vae_cfg.get("latent_channels", 4)), norm_groups=int(unet_cfg.get("norm_num_groups", 32)),
        latent_scale=float(vae_cfg.get("scaling_factor", 0.18215)))


_SCHED_KEYS = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "prediction_type", "timestep_spacing",
               "rescale_betas_zero_snr", "set_alpha_to_one", "steps_offset", "original_inference_steps", "timestep_scaling")


_DIFFUSERS_SCHED_DEFAULTS = {
    "DDIMScheduler": dict(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                          clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
                          thresholding=False, timestep_spacing="leading", rescale_betas_zero_snr=False),
    "LCMScheduler": dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                         original_inference_steps=50, clip_sample=False, set_alpha_to_one=True, steps_offset=0,
                         prediction_type="epsilon", thresholding=False, timestep_spacing="leading", timestep_scaling=10.0,
                         rescale_betas_zero_snr=False),
}


def scheduler_from_config(cfg: dict):
    """scheduler/scheduler_config.json -> host scheduler mirror (marigold_depth_pipeline.py:340-379 accepts exactly
    these two classes)."""
    name = cfg.get("_class_name", "DDIMScheduler")
    # keys a config omits take DIFFUSERS' constructor defaults (not Marigold's values), as from_pretrained would
    d = dict(_DIFFUSERS_SCHED_DEFAULTS.get(name, {}))
    d.update(cfg)
    cfg = d
    kw = {k: cfg[k] for k in _SCHED_KEYS if k in cfg}
    if cfg.get("clip_sample", False) or cfg.get("thresholding", False):
        raise CheckpointError("clip_sample / thresholding schedulers are not supported (Marigold ships them disabled; a "
                              "config that omits `clip_sample` means diffusers' default, True)")
    if name == "DDIMScheduler":
        kw.pop("original_inference_steps", None)
        kw.pop("timestep_scaling", None)
        return DDIMScheduler(**kw)
    if name == "LCMScheduler":
        return LCMScheduler(**kw)
    raise CheckpointError(f"Unsupported scheduler type: {name}")


# -------------------------------------------------------------------------------------------------
# empty-prompt embedding
# -------------------------------------------------------------------------------------------------
def empty_text_embedding(root, cross_dim: int, variant: Optional[str] = None) -> torch.Tensor:
    """[1, 2, cross_dim] fp32: CLIP hidden states of the empty prompt with padding="do_not_pad" (BOS, EOS)."""
    root = Path(root)
    for name in ("empty_text_embed.safetensors", "empty_text_embed.pt", "empty_text_embed.npy"):
        p = root / name
        if not p.is_file():
            continue
        if p.suffix == ".safetensors":
            t = next(iter(read_safetensors(p).values()))
        elif p.suffix == ".pt":
            t = torch.load(p, map_location="cpu", weights_only=True)
        else:
            t = torch.from_numpy(np.load(p))
        t = t.float().reshape(1, -1, t.shape[-1]).clone()
        if t.shape[-1] != cross_dim:
            raise CheckpointError(f"{p}: embedding width {t.shape[-1]} != cross_attention_dim {cross_dim}")
        return t
    if (root / "text_encoder").is_dir() and (root / "tokenizer").is_dir():
        try:
            from transformers import CLIPTextModel, CLIPTokenizer
        except Exception as e:  # noqa: BLE001
            raise CheckpointError(f"transformers is needed to encode the empty prompt ({e}); or provide "
                                  f"{root}/empty_text_embed.safetensors") from None
        tok = CLIPTokenizer.from_pretrained(str(root / "tokenizer"))
        try:
            enc = CLIPTextModel.from_pretrained(str(root / "text_encoder"), variant=variant).eval()
        except Exception:  # noqa: BLE001  (no weights of that variant: fall back to the default file names)
            enc = CLIPTextModel.from_pretrained(str(root / "text_encoder")).eval()
        enc = enc.float()
        ids = tok("", padding="do_not_pad", max_length=tok.model_max_length, truncation=True, return_tensors="pt").input_ids
        with torch.no_grad():
            t = enc(ids)[0].float()
        if t.shape[-1] != cross_dim:
            raise CheckpointError(f"text encoder width {t.shape[-1]} != cross_attention_dim {cross_dim}")
        return t
    raise CheckpointError(f"{root}: neither empty_text_embed.* nor text_encoder/ + tokenizer/ found")


# -------------------------------------------------------------------------------------------------
# the whole directory
# -------------------------------------------------------------------------------------------------
def inspect_checkpoint(root, variant: Optional[str] = None) -> dict:
    """Parse everything that needs no GPU: configs, scheduler, pipeline defaults, weight file names."""
    root = Path(os.fspath(root))
    if not root.is_dir():
        raise CheckpointError(f"checkpoint directory not found: {root} (there is no network: pass a local path)")
    index = _read_json(root / "model_index.json") if (root / "model_index.json").is_file() else {}
    unet_cfg, vae_cfg = _read_json(root / "unet" / "config.json"), _read_json(root / "vae" / "config.json")
    sched_cfg = _read_json(root / "scheduler" / "scheduler_config.json")
    return {
        "root": root, "variant": variant, "index": index, "unet_cfg": unet_cfg, "vae_cfg": vae_cfg,
        "engine_config": engine_config_from_diffusers(unet_cfg, vae_cfg), "scheduler": scheduler_from_config(sched_cfg),
        "defaults": {k: index.get(k) for k in ("default_denoising_steps", "default_processing_resolution",
                                                "scale_invariant", "shift_invariant", "target_properties") if k in index},
    }


def load_pipeline(cls, root, variant: Optional[str] = None, torch_dtype=None, device=None, **overrides):
    """`cls.from_pretrained(root, variant=, torch_dtype=)`: torch_dtype is accepted for call-site compatibility; the
    kernels always run bf16 operands with fp32 accumulation."""
    from .engine import Engine

    info = inspect_checkpoint(root, variant)
    cfg = info["engine_config"]
    unet_sd, vae_sd = read_weights(info["root"] / "unet", variant), read_weights(info["root"] / "vae", variant)
    text = empty_text_embedding(info["root"], cfg.unet_cross_dim, variant)
    eng = Engine(cfg, device=device)
    eng.load_state_dict("unet", unet_sd)
    eng.load_state_dict("vae", vae_sd)
    eng.finalize()
    kw = dict(info["defaults"])
    kw.update(overrides)
    import inspect

    accepted = set(inspect.signature(cls.__init__).parameters)
    kw = {k: v for k, v in kw.items() if k in accepted}
    return cls(eng, info["scheduler"], text, **kw)


def export_checkpoint(root, unet_sd: Dict[str, torch.Tensor], vae_sd: Dict[str, torch.Tensor], unet_cfg: dict,
                      vae_cfg: dict, scheduler_cfg: dict, empty_text_embed: torch.Tensor,
                      index: Optional[dict] = None, variant: Optional[str] = None) -> Tuple[Path, Path]:
    """Write the layout above (used by the tests to fabricate a checkpoint from the seeded oracle networks)."""
    root = Path(root)
    stem = f"diffusion_pytorch_model.{variant}" if variant else "diffusion_pytorch_model"
    for sub in ("unet", "vae", "scheduler"):
        (root / sub).mkdir(parents=True, exist_ok=True)
    (root / "unet" / "config.json").write_text(json.dumps(unet_cfg, indent=1))
    (root / "vae" / "config.json").write_text(json.dumps(vae_cfg, indent=1))
    (root / "scheduler" / "scheduler_config.json").write_text(json.dumps(scheduler_cfg, indent=1))
    (root / "model_index.json").write_text(json.dumps(index or {}, indent=1))
    pu, pv = root / "unet" / f"{stem}.safetensors", root / "vae" / f"{stem}.safetensors"
    write_safetensors(pu, unet_sd, {"format": "pt"})
    write_safetensors(pv, vae_sd, {"format": "pt"})
    write_safetensors(root / "empty_text_embed.safetensors", {"empty_text_embed": empty_text_embed.float()})
    return pu, pv
