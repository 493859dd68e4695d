"""Checkpoint loader: the HF diffusers directory layout the reference's
`from_pretrained` consumes (script/depth/run.py:213-222, README.md:261-290), fabricated here from the seeded
oracle networks because no real checkpoint exists offline."""
import json

import numpy as np
import pytest
import torch

from marigold_b200 import checkpoint as ck
from marigold_b200.schedulers import DDIMScheduler, LCMScheduler
from tests.helpers import oracle_models


def _configs(unet, vae):
    boc = list(unet.cfg.block_out_channels)
    unet_cfg = {"_class_name": "UNet2DConditionModel", "in_channels": 8, "out_channels": 4, "block_out_channels": boc,
                "layers_per_block": 2, "cross_attention_dim": unet.cfg.cross_attention_dim,
                "attention_head_dim": [c // 64 for c in boc], "use_linear_projection": True, "norm_num_groups": 32,
                "act_fn": "silu",
                "down_block_types": ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"],
                "up_block_types": ["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3}
    vae_cfg = {"_class_name": "AutoencoderKL", "in_channels": 3, "out_channels": 3,
               "block_out_channels": list(vae.cfg.block_out_channels), "layers_per_block": 2, "latent_channels": 4,
               "norm_num_groups": 32, "scaling_factor": 0.18215}
    sched_cfg = {"_class_name": "DDIMScheduler", "num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012,
                 "beta_schedule": "scaled_linear", "prediction_type": "v_prediction", "timestep_spacing": "trailing",
                 "rescale_betas_zero_snr": True, "set_alpha_to_one": False, "steps_offset": 1, "clip_sample": False}
    index = {"_class_name": "MarigoldDepthPipeline", "default_denoising_steps": 4, "default_processing_resolution": 64,
             "scale_invariant": True, "shift_invariant": True}
    return unet_cfg, vae_cfg, sched_cfg, index


@pytest.fixture(scope="module")
def fake_checkpoint(tmp_path_factory):
    unet, vae, text = oracle_models("tiny")
    root = tmp_path_factory.mktemp("ckpt")
    unet_cfg, vae_cfg, sched_cfg, index = _configs(unet, vae)
    # the fp16 variant file next to the plain one, like the published repositories
    ck.export_checkpoint(root, unet.state_dict(), vae.state_dict(), unet_cfg, vae_cfg, sched_cfg, text, index)
    ck.export_checkpoint(root, {k: v.half() for k, v in unet.state_dict().items()},
                         {k: v.half() for k, v in vae.state_dict().items()}, unet_cfg, vae_cfg, sched_cfg, text, index,
                         variant="fp16")
    return root, unet, vae, text


def test_safetensors_roundtrip_and_official_reader(tmp_path):
    g = torch.Generator().manual_seed(3)
    tensors = {"a.weight": torch.randn(5, 7, generator=g), "b": torch.randn(3, generator=g).to(torch.bfloat16),
               "c": torch.arange(6, dtype=torch.int64).reshape(2, 3), "empty": torch.empty(0, 4), "h": torch.randn(2, 2).half()}
    p = tmp_path / "t.safetensors"
    ck.write_safetensors(p, tensors, {"format": "pt"})
    back = ck.read_safetensors(p)
    assert set(back) == set(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == v.dtype and tuple(back[k].shape) == tuple(v.shape)
        assert torch.equal(back[k].clone(), v)
    # the file is a valid safetensors file for the reference's own stack, and theirs parses with our reader
    st = pytest.importorskip("safetensors.torch")
    theirs = st.load_file(str(p))
    for k, v in tensors.items():
        assert torch.equal(theirs[k], v)
    q = tmp_path / "u.safetensors"
    st.save_file({k: v.contiguous() for k, v in tensors.items() if v.numel()}, str(q))
    ours = ck.read_safetensors(q)
    for k, v in tensors.items():
        if v.numel():
            assert torch.equal(ours[k].clone(), v)


def test_safetensors_rejects_corrupt_files(tmp_path):
    p = tmp_path / "bad.safetensors"
    p.write_bytes(b"\x01\x02")
    with pytest.raises(ck.CheckpointError):
        ck.read_safetensors(p)
    p.write_bytes((10 ** 12).to_bytes(8, "little") + b"{}")
    with pytest.raises(ck.CheckpointError):
        ck.read_safetensors(p)
    hdr = json.dumps({"w": {"dtype": "F32", "shape": [4], "data_offsets": [0, 8]}}).encode()
    p.write_bytes(len(hdr).to_bytes(8, "little") + hdr + b"\0" * 8)
    with pytest.raises(ck.CheckpointError):      # 4 floats need 16 bytes
        ck.read_safetensors(p)


def test_inspect_checkpoint_maps_configs(fake_checkpoint):
    root, unet, vae, text = fake_checkpoint
    info = ck.inspect_checkpoint(root)
    cfg = info["engine_config"]
    assert cfg.unet_block_channels == list(unet.cfg.block_out_channels)
    assert cfg.vae_block_channels == list(vae.cfg.block_out_channels)
    assert cfg.unet_cross_dim == unet.cfg.cross_attention_dim and cfg.latent_scale == pytest.approx(0.18215)
    assert isinstance(info["scheduler"], DDIMScheduler)
    assert info["scheduler"].config.timestep_spacing == "trailing" and info["scheduler"].config.rescale_betas_zero_snr
    assert info["defaults"] == {"default_denoising_steps": 4, "default_processing_resolution": 64,
                                "scale_invariant": True, "shift_invariant": True}
    sd = ck.read_weights(root / "unet")
    assert set(sd) == set(unet.state_dict())
    k = "conv_in.weight"
    assert torch.equal(sd[k].clone(), unet.state_dict()[k])
    sd16 = ck.read_weights(root / "unet", variant="fp16")
    assert sd16[k].dtype == torch.float16
    assert ck.read_weights(root / "unet", variant="nonexistent")[k].dtype == torch.float32   # falls back like diffusers
    assert torch.equal(ck.empty_text_embedding(root, unet.cfg.cross_attention_dim), text.float())


def test_unsupported_architectures_and_schedulers_are_refused(fake_checkpoint):
    root, unet, vae, _ = fake_checkpoint
    unet_cfg, vae_cfg, sched_cfg, _ = _configs(unet, vae)
    for patch in ({"use_linear_projection": False}, {"attention_head_dim": 8}, {"act_fn": "gelu"},
                  {"block_out_channels": [64, 128, 256]}, {"class_embed_type": "timestep"}):
        with pytest.raises(ck.CheckpointError):
            ck.engine_config_from_diffusers({**unet_cfg, **patch}, vae_cfg)
    assert isinstance(ck.scheduler_from_config({**sched_cfg, "_class_name": "LCMScheduler", "timestep_spacing": "leading",
                                                "rescale_betas_zero_snr": False}), LCMScheduler)
    with pytest.raises(ck.CheckpointError, match="Unsupported scheduler type"):
        ck.scheduler_from_config({**sched_cfg, "_class_name": "EulerDiscreteScheduler"})
    with pytest.raises(ck.CheckpointError):
        ck.scheduler_from_config({**sched_cfg, "clip_sample": True})
    # a sparse scheduler config means DIFFUSERS' defaults (clip_sample=True -> refused; epsilon / leading / alpha_to_one),
    # never Marigold's values silently
    with pytest.raises(ck.CheckpointError):
        ck.scheduler_from_config({"_class_name": "DDIMScheduler"})
    sparse = ck.scheduler_from_config({"_class_name": "DDIMScheduler", "clip_sample": False})
    assert (sparse.config.prediction_type, sparse.config.timestep_spacing, sparse.config.set_alpha_to_one,
            sparse.config.rescale_betas_zero_snr, sparse.config.beta_schedule) == ("epsilon", "leading", True, False, "linear")
    with pytest.raises(ck.CheckpointError, match="not found"):
        ck.inspect_checkpoint(root / "missing")
    with pytest.raises(ck.CheckpointError):
        ck.empty_text_embedding(root, 999)


@pytest.mark.gpu
def test_from_pretrained_matches_engine_built_from_state_dicts(fake_checkpoint):
    """`from_pretrained` (fp32 file and fp16 variant) == the pipeline the other GPU tests build by hand."""
    from marigold_b200.pipeline import MarigoldDepthPipeline, MarigoldNormalsPipeline
    from tests.helpers import engine_from_oracle, synthetic_image

    root, unet, vae, text = fake_checkpoint
    img = synthetic_image(64)
    noise = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(2024))
    pipe = MarigoldDepthPipeline.from_pretrained(str(root), variant=None, torch_dtype=torch.bfloat16)
    assert pipe.default_denoising_steps == 4 and pipe.default_processing_resolution == 64
    assert pipe.scale_invariant and pipe.shift_invariant
    out = pipe(img, ensemble_size=1, noise=noise, show_progress_bar=False)     # defaults from model_index.json
    ref_pipe = MarigoldDepthPipeline(engine_from_oracle(unet, vae, text), DDIMScheduler(), text,
                                     default_denoising_steps=4, default_processing_resolution=64)
    ref = ref_pipe(img, ensemble_size=1, noise=noise, show_progress_bar=False)
    # same weights, kernels and schedule, and every reduction on the path has a fixed order: bit-equal
    np.testing.assert_array_equal(out.depth_np, ref.depth_np)
    pipe16 = MarigoldDepthPipeline.from_pretrained(str(root), variant="fp16")
    out16 = pipe16(img, ensemble_size=1, noise=noise, show_progress_bar=False)
    d16 = np.abs(out16.depth_np - ref.depth_np)
    assert d16.max() < 0.15 and d16.mean() < 1e-2, (d16.max(), d16.mean())   # DIFFERENT weights: fp16-rounded first
    npipe = MarigoldNormalsPipeline.from_pretrained(str(root))
    nout = npipe(img, denoising_steps=2, ensemble_size=1, noise=noise, show_progress_bar=False)
    assert nout.normals_np.shape == (3, 64, 64) and np.isfinite(nout.normals_np).all()
    for p in (pipe, pipe16, npipe, ref_pipe):
        p.engine.close()


def test_legacy_vae_attention_keys_are_renamed(tmp_path):
    """SD-era VAE files: mid_block.attentions.0.{query,key,value,proj_attn}, sometimes stored as 1x1 convolutions
    (diffusers converts them on load)."""
    sd = {"encoder.mid_block.attentions.0.query.weight": torch.randn(8, 8, 1, 1),
          "encoder.mid_block.attentions.0.query.bias": torch.randn(8),
          "encoder.mid_block.attentions.0.proj_attn.weight": torch.randn(8, 8),
          "decoder.mid_block.attentions.0.to_k.weight": torch.randn(8, 8, 1, 1),
          "encoder.conv_in.weight": torch.randn(8, 3, 3, 3)}
    (tmp_path / "vae").mkdir()
    ck.write_safetensors(tmp_path / "vae" / "diffusion_pytorch_model.safetensors", sd)
    out = ck.read_weights(tmp_path / "vae")
    assert set(out) == {"encoder.mid_block.attentions.0.to_q.weight", "encoder.mid_block.attentions.0.to_q.bias",
                        "encoder.mid_block.attentions.0.to_out.0.weight", "decoder.mid_block.attentions.0.to_k.weight",
                        "encoder.conv_in.weight"}
    assert out["encoder.mid_block.attentions.0.to_q.weight"].shape == (8, 8)
    assert out["decoder.mid_block.attentions.0.to_k.weight"].shape == (8, 8)
    assert torch.equal(out["encoder.mid_block.attentions.0.to_q.weight"],
                       sd["encoder.mid_block.attentions.0.query.weight"].reshape(8, 8))
    assert out["encoder.conv_in.weight"].shape == (8, 3, 3, 3)
