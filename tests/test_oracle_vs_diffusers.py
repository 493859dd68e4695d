"""The escape hatch of SURVEY.md §8c(3): the network / scheduler arithmetic of the hot path lives in `diffusers`
(requirements.txt:2, reached from marigold/marigold_depth_pipeline.py:35-42), which is not installed offline, so
oracle/{unet,vae,schedulers}.py are restatements whose parity is UNPINNED. Wherever `diffusers` IS importable this
test pins them: the oracle modules use diffusers' state-dict key names, so the same seeded weights load into the real
classes and the outputs must agree to fp32 rounding. Skipped automatically when diffusers is absent."""
import pytest
import torch

diffusers = pytest.importorskip("diffusers", reason="diffusers is not installed (offline image): oracle parity stays unpinned")

from oracle.schedulers import DDIMSchedulerOracle, LCMSchedulerOracle, SchedulerConfig  # noqa: E402
from oracle.unet import UNet2DConditionOracle, UNetConfig  # noqa: E402
from oracle.vae import AutoencoderKLOracle, VAEConfig  # noqa: E402


def _real_unet(cfg: UNetConfig):
    boc = list(cfg.block_out_channels)
    return diffusers.UNet2DConditionModel(
        in_channels=cfg.in_channels, out_channels=cfg.out_channels, block_out_channels=boc,
        layers_per_block=cfg.layers_per_block, cross_attention_dim=cfg.cross_attention_dim,
        attention_head_dim=[c // cfg.head_dim for c in boc],      # SD-2 configs store head COUNTS under this name
        down_block_types=["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"],
        up_block_types=["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3, use_linear_projection=True,
        norm_num_groups=cfg.norm_num_groups, norm_eps=cfg.norm_eps, act_fn="silu", sample_size=16).eval()


@pytest.mark.parametrize("h,w", [(16, 16), (27, 12), (7, 9)])
def test_unet_oracle_equals_diffusers(h, w):
    torch.manual_seed(0)
    cfg = UNetConfig.tiny()
    ora = UNet2DConditionOracle(cfg).eval()
    real = _real_unet(cfg)
    missing, unexpected = real.load_state_dict(ora.state_dict(), strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, cfg.in_channels, h, w, generator=g)
    ctx = torch.randn(2, 2, cfg.cross_attention_dim, generator=g)
    with torch.no_grad():
        for t in (999, 499, 19):
            a = ora(x, t, ctx)
            b = real(x, t, encoder_hidden_states=ctx).sample
            assert torch.allclose(a, b, atol=2e-5, rtol=1e-4), (t, (a - b).abs().max())


@pytest.mark.parametrize("H,W", [(64, 64), (100, 50), (77, 131)])
def test_vae_oracle_equals_diffusers(H, W):
    torch.manual_seed(0)
    cfg = VAEConfig.tiny()
    ora = AutoencoderKLOracle(cfg).eval()
    real = diffusers.AutoencoderKL(
        in_channels=3, out_channels=3, block_out_channels=list(cfg.block_out_channels), layers_per_block=cfg.layers_per_block,
        latent_channels=cfg.latent_channels, norm_num_groups=cfg.norm_num_groups,
        down_block_types=["DownEncoderBlock2D"] * 4, up_block_types=["UpDecoderBlock2D"] * 4).eval()
    missing, unexpected = real.load_state_dict(ora.state_dict(), strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    img = torch.rand(1, 3, H, W, generator=torch.Generator().manual_seed(2)) * 2 - 1
    with torch.no_grad():
        m_o = ora.quant_conv(ora.encoder(img))
        m_r = real.quant_conv(real.encoder(img))          # the sub-module calls of marigold_depth_pipeline.py:491-492
        assert torch.allclose(m_o, m_r, atol=2e-5, rtol=1e-4)
        z = m_o[:, :4]
        d_o = ora.decoder(ora.post_quant_conv(z))
        d_r = real.decoder(real.post_quant_conv(z))       # :512-513
        assert torch.allclose(d_o, d_r, atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("n", [1, 4, 10, 50, 7])
def test_ddim_oracle_equals_diffusers(n):
    kw = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
              prediction_type="v_prediction", timestep_spacing="trailing", rescale_betas_zero_snr=True,
              set_alpha_to_one=False, steps_offset=1, clip_sample=False)
    real = diffusers.DDIMScheduler(**kw)
    ora = DDIMSchedulerOracle(SchedulerConfig())
    real.set_timesteps(n)
    ora.set_timesteps(n)
    assert real.timesteps.tolist() == ora.timesteps.tolist()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 4, 8, 8, generator=g)
    for t in real.timesteps:
        v = torch.randn(1, 4, 8, 8, generator=g)
        a = ora.step(v, int(t), x)
        b = real.step(v, t, x).prev_sample
        assert torch.allclose(a, b, atol=1e-6, rtol=1e-5), (int(t), (a - b).abs().max())
        x = b


@pytest.mark.parametrize("n", [1, 4])
def test_lcm_oracle_equals_diffusers(n):
    real = diffusers.LCMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                  prediction_type="v_prediction", timestep_spacing="leading", rescale_betas_zero_snr=False,
                                  set_alpha_to_one=False, steps_offset=1, clip_sample=False, original_inference_steps=50)
    ora = LCMSchedulerOracle(SchedulerConfig(timestep_spacing="leading", rescale_betas_zero_snr=False))
    real.set_timesteps(n)
    ora.set_timesteps(n)
    assert real.timesteps.tolist() == ora.timesteps.tolist()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 4, 8, 8, generator=g)
    for i, t in enumerate(real.timesteps):
        v = torch.randn(1, 4, 8, 8, generator=g)
        gen = torch.Generator().manual_seed(100 + i)
        b = real.step(v, t, x, generator=gen).prev_sample
        z = torch.randn(x.shape, generator=torch.Generator().manual_seed(100 + i)) if i < n - 1 else None
        a = ora.step(v, int(t), x, noise=z)
        assert torch.allclose(a, b, atol=1e-6, rtol=1e-5), (int(t), (a - b).abs().max())
        x = b
