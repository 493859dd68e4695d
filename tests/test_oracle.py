"""CPU tests that pin the oracle: architecture known-answers (parameter counts, timestep lists,
zero-terminal-SNR) and the reference-generated golden vectors for the ensembling functions."""
import numpy as np
import pytest
import torch

from oracle.ensemble import ensemble_depth, ensemble_iid, ensemble_normals
from oracle.schedulers import DDIMSchedulerOracle, LCMSchedulerOracle
from oracle.unet import UNet2DConditionOracle, UNetConfig
from oracle.vae import AutoencoderKLOracle, VAEConfig
from tests.golden.cases import (DEPTH_CASES, IID_CASES, NORMALS_CASES, RESIZE_CASES, depth_input, iid_input, normals_input,
                                resize_input)

GOLD = np.load(__file__.rsplit("/", 1)[0] + "/golden/ensemble_golden.npz")


def _count(m):
    return sum(p.numel() for p in m.parameters())


def test_sd2_parameter_counts():
    # the only offline evidence that the restated architecture is structurally right
    with torch.device("meta"):
        u = UNet2DConditionOracle(UNetConfig())
        v = AutoencoderKLOracle(VAEConfig())
    assert _count(u) == 865_922_244
    assert _count(v.encoder) + _count(v.quant_conv) == 34_163_664
    assert _count(v.decoder) + _count(v.post_quant_conv) == 49_490_199


def test_unet_shape_walk_tiny():
    torch.manual_seed(0)
    u = UNet2DConditionOracle(UNetConfig.tiny()).eval()
    with torch.no_grad():
        y = u(torch.randn(2, 8, 16, 24), 999, torch.randn(2, 2, 128))
    assert y.shape == (2, 4, 16, 24) and torch.isfinite(y).all()


def test_ddim_timesteps_and_zero_snr():
    d = DDIMSchedulerOracle()
    want = {1: [999], 4: [999, 749, 499, 249], 10: list(range(999, 0, -100))}
    for n, ts in want.items():
        d.set_timesteps(n)
        assert d.timesteps.tolist() == ts
    d.set_timesteps(50)
    assert d.timesteps[:3].tolist() == [999, 979, 959] and int(d.timesteps[-1]) == 19
    assert float(d.alphas_cumprod[999]) == 0.0
    # t = 999, v-prediction, zero SNR: x0 = -v, no division
    d.set_timesteps(1)
    x, v = torch.randn(1, 4, 4, 4), torch.randn(1, 4, 4, 4)
    a_prev = d.final_alpha_cumprod
    assert torch.allclose(d.step(v, 999, x), a_prev.sqrt() * (-v) + (1 - a_prev).sqrt() * x, atol=1e-6)


def test_lcm_timesteps():
    l = LCMSchedulerOracle()
    l.set_timesteps(4)
    assert l.timesteps.tolist() == [999, 759, 499, 259]
    l.set_timesteps(1)
    assert l.timesteps.tolist() == [999]


@pytest.mark.parametrize("name", list(DEPTH_CASES))
def test_ensemble_depth_oracle_matches_reference_golden(name):
    cfg = DEPTH_CASES[name]
    pred, unc, param = ensemble_depth(depth_input(cfg), return_param=True, **cfg.get("kwargs", {}))
    np.testing.assert_array_equal(pred.numpy(), GOLD[f"depth/{name}/pred"])      # bit-exact
    if f"depth/{name}/unc" in GOLD:
        np.testing.assert_array_equal(unc.numpy(), GOLD[f"depth/{name}/unc"])
    np.testing.assert_array_equal(param, GOLD[f"depth/{name}/x"])


@pytest.mark.parametrize("name", list(NORMALS_CASES))
def test_ensemble_normals_oracle_matches_reference_golden(name):
    cfg = NORMALS_CASES[name]
    pred, unc = ensemble_normals(normals_input(cfg), **cfg.get("kwargs", {}))
    np.testing.assert_array_equal(pred.numpy(), GOLD[f"normals/{name}/pred"])
    if f"normals/{name}/unc" in GOLD:
        np.testing.assert_array_equal(unc.numpy(), GOLD[f"normals/{name}/unc"])


def test_ensemble_error_behaviour():
    with pytest.raises(ValueError):
        ensemble_depth(torch.rand(2, 3, 4, 4))
    with pytest.raises(ValueError):
        ensemble_depth(torch.rand(2, 1, 4, 4), reduction="mode")
    with pytest.raises(ValueError):
        ensemble_depth(torch.rand(2, 1, 4, 4), scale_invariant=False, shift_invariant=True)
    with pytest.raises(ValueError):
        ensemble_normals(torch.rand(2, 1, 4, 4))
    with pytest.raises(ValueError):
        ensemble_normals(torch.rand(2, 3, 4, 4), reduction="median")


@pytest.mark.parametrize("name", [k for k in RESIZE_CASES if f"resize/{k}" in GOLD.files])
def test_resize_max_res_matches_reference_golden(name):
    from oracle.pipeline import resize_max_res

    cfg = RESIZE_CASES[name]
    mode = {"bilinear": "bilinear", "nearest": "nearest-exact"}[cfg["method"]]
    out = resize_max_res(resize_input(cfg), cfg["max_edge"], mode)
    g = GOLD[f"resize/{name}"]
    assert out.shape == g.shape
    assert np.abs(out.numpy().astype(np.int32) - g.astype(np.int32)).max() <= 1   # uint8 rounding of antialias


def test_oracle_network_outputs_are_frozen():
    """tests/golden/network_golden.npz pins the oracle to itself (tests/golden/make_network_golden.py): UNet at three
    timesteps, VAE encode / decode, 4-step DDIM and LCM depth, 2-step DDIM normals on the seeded tiny model. Tolerance
    covers oneDNN thread-count / ISA differences between the build container and the GPU box."""
    from pathlib import Path

    from tests.golden.make_network_golden import compute

    gold = np.load(Path(__file__).resolve().parent / "golden" / "network_golden.npz")
    got = compute()
    assert set(got) == set(gold.files)
    for k in gold.files:
        a, b = got[k], gold[k]
        assert a.shape == b.shape, k
        err = np.abs(a - b).max() / (np.abs(b).max() + 1e-12)
        assert err < 2e-4, (k, err)


@pytest.mark.parametrize("name", sorted(IID_CASES))
def test_ensemble_iid_oracle_matches_reference_golden(name):
    """ensemble_iid (ensemble.py:250-270) against outputs of the reference's own function (tests/golden/make_golden.py):
    bit-exact, including the lower median on even E, exact ties, E = 1 (std of one member is NaN, like the reference)."""
    from pathlib import Path

    gold = np.load(Path(__file__).resolve().parent / "golden" / "iid_golden.npz")
    cfg = IID_CASES[name]
    pred, unc = ensemble_iid(iid_input(cfg), **dict(cfg.get("kwargs", {})))
    assert np.array_equal(pred.numpy(), gold[f"iid/{name}/pred"])
    if f"iid/{name}/unc" in gold.files:
        assert np.array_equal(unc.numpy(), gold[f"iid/{name}/unc"], equal_nan=True)
    else:
        assert unc is None
    with pytest.raises(ValueError):
        ensemble_iid(iid_input(cfg), reduction="max")


def test_oracle_iid_pipeline_control_flow():
    """OracleIIDPipeline (marigold_iid_pipeline.py:239-411,467-585) on a seeded 2-target tiny model: conv_in takes
    4 * (n + 1) channels, conv_out gives 4 * n, every target is decoded separately, E > 1 goes through ensemble_iid."""
    from oracle.pipeline import OracleIIDPipeline
    from tests.helpers import synthetic_image

    torch.manual_seed(0)
    ucfg = UNetConfig.tiny()
    ucfg.in_channels, ucfg.out_channels = 12, 8
    unet, vae = UNet2DConditionOracle(ucfg).eval(), AutoencoderKLOracle(VAEConfig.tiny()).eval()
    text = torch.randn(1, 2, ucfg.cross_attention_dim, generator=torch.Generator().manual_seed(7))
    pipe = OracleIIDPipeline(unet, vae, DDIMSchedulerOracle(), text, ["albedo", "material"], 2, 128)
    img = synthetic_image(256)[:, :, :128, :]                     # 128 x 256 input, processed at 64 x 128
    z = torch.randn(3, 8, 8, 16, generator=torch.Generator().manual_seed(2024))
    one, unc1, m1 = pipe(img, ensemble_size=1, noise=z[:1])
    assert one.shape == (1, 6, 128, 256) and unc1 is None and m1.shape == (1, 6, 64, 128)
    assert 0.0 <= float(one.min()) and float(one.max()) <= 1.0
    ens, unc, mem = pipe(img, ensemble_size=3, noise=z, batch_size=2, ensemble_kwargs=dict(output_uncertainty=True))
    assert ens.shape == (1, 6, 128, 256) and unc.shape == (1, 6, 64, 128) and mem.shape == (3, 6, 64, 128)
    assert torch.allclose(mem[:1], m1, atol=1e-5)                  # member k only depends on noise row k, not on the batching
    med = mem.median(dim=0, keepdim=True).values
    assert torch.allclose(pipe(img, ensemble_size=3, noise=z, batch_size=2, match_input_res=False)[0], med, atol=1e-6)
    parts = pipe.split(ens, None)
    assert list(parts) == ["albedo", "material"] and parts["material"][0].shape == (1, 3, 128, 256)
    with pytest.raises(AssertionError):
        pipe(img, ensemble_size=1, noise=z[:1, :4])                # wrong latent width
