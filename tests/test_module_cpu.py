"""CPU: host-side logic of the drop-in module -- state_dict layout, checkpoint compatibility, config mapping,
loss-handler parsing -- and that the compute path fails loudly instead of falling back."""
from types import SimpleNamespace

import pytest
import torch

from oracle import swinv2_oracle as O
from swin_v2_weather_b200.networks.helpers import MultiStepWrapper, SingleStepWrapper, get_model
from swin_v2_weather_b200.networks.swinv2_global import SwinTransformerV2Cr, swinv2net
from swin_v2_weather_b200.utils.losses import LossHandler


def small(rel_pos=False, **kw):
    return SwinTransformerV2Cr(img_size=(72, 144), patch_size=4, depths=(2,), num_heads=(2,), in_chans=7, out_chans=5,
                               embed_dim=192, img_window_ratio=8, full_pos_embed=True, rel_pos=rel_pos, **kw)


@pytest.mark.parametrize("rel_pos", [False, True])
def test_state_dict_layout_matches_reference_tree(rel_pos):
    cfg = O.SwinConfig(img_size=(72, 144), depth=2, num_heads=2, in_chans=7, out_chans=5, embed_dim=192, window_ratio=8,
                       rel_pos=rel_pos)
    want = O.init_state_dict(cfg)           # key tree validated against the live reference in test_oracle_vs_reference
    got = small(rel_pos).state_dict()
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert got[k].shape == want[k].shape and got[k].dtype == torch.float32, k


def test_full_config_parameter_count():
    """depth 12, C=768 -> 162 keys, 136,617,312 elements (meta device: no memory)."""
    with torch.device("meta"):
        m = SwinTransformerV2Cr(img_size=(720, 1440), patch_size=4, depths=(12,), num_heads=(8,), in_chans=73, out_chans=73,
                                embed_dim=768, img_window_ratio=80, full_pos_embed=True, rel_pos=False, drop_path_rate=0.1)
    sd = m.state_dict()
    assert len(sd) == 162
    assert sum(v.numel() for v in sd.values()) == 136_617_312
    assert m.window_size == (9, 18)
    blocks = m.stages[0].blocks
    assert [b.shift_size for b in blocks] == [(0, 0), (4, 9)] * 6
    dpr = torch.linspace(0, 0.1, 12).tolist()
    for b, p in zip(blocks, dpr):
        assert (p == 0.0) == isinstance(b.drop_path1, torch.nn.Identity)
        if p > 0:
            assert abs(b.drop_path1.drop_prob - p) < 1e-9


def test_checkpoint_with_ddp_prefix_loads():
    """train.py:377,386-388 -- checkpoints carry `module.` (DDP) + `model.` (wrapper) prefixes."""
    params = SimpleNamespace(nettype='swin', n_future=0, img_size=[72, 144], patch_size=4, depth=2, num_heads=2,
                             n_in_channels=7, n_out_channels=5, embed_dim=192, window_ratio=8, drop_path_rate=0.0,
                             full_pos_embed=True, rel_pos=False, mlp_ratio=4, activation_ckpt=False, residual=False)
    src = get_model(params)
    assert isinstance(src, SingleStepWrapper)
    ckpt = {"module." + k: v.clone() for k, v in src.state_dict().items()}
    assert all(k.startswith("module.model.") for k in ckpt)
    dst = get_model(params)
    stripped = {k[7:]: v for k, v in ckpt.items()}      # restore_checkpoint strips the first 7 chars
    dst.load_state_dict(stripped)
    for (ka, va), (kb, vb) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)


def test_swinv2net_reads_the_reference_hyperparameters():
    params = SimpleNamespace(img_size=[72, 144], patch_size=4, depth=3, num_heads=2, n_in_channels=9, n_out_channels=5,
                             embed_dim=192, window_ratio=8, drop_path_rate=0.2, full_pos_embed=False, rel_pos=True, mlp_ratio=2,
                             activation_ckpt=True, residual=True)
    m = swinv2net(params)
    assert not hasattr(m, "pos_embed") and m.residual and m.stages[0].grad_checkpointing
    assert m.patch_embed.proj.weight.shape == (192, 9, 4, 4)
    assert m.stages[0].blocks[0].mlp.fc1.weight.shape == (384, 192)
    assert m.head.weight.shape == (5 * 16, 192)
    assert hasattr(m.stages[0].blocks[0].attn, "meta_mlp")
    m.set_grad_checkpointing(False)
    assert not m.stages[0].grad_checkpointing


def test_multistep_wrapper_kept():
    params = SimpleNamespace(nettype='swin', n_future=2, add_orography=True, add_landmask=True, img_size=[72, 144], patch_size=4,
                             depth=1, num_heads=1, n_in_channels=8, n_out_channels=5, embed_dim=96, window_ratio=8,
                             drop_path_rate=0.0, full_pos_embed=True, rel_pos=False, mlp_ratio=4, activation_ckpt=False,
                             residual=True)
    m = get_model(params)
    assert isinstance(m, MultiStepWrapper) and m.invar == 3


def test_no_cpu_fallback():
    m = small()
    with pytest.raises(Exception) as e:
        m(torch.randn(1, 7, 72, 144))
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)


def test_optimizer_mirrors_torch_adam_surface_and_has_no_cpu_fallback():
    from swin_v2_weather_b200._lib import SwinB200Error
    from swin_v2_weather_b200.optim import Adam
    w = torch.nn.Parameter(torch.randn(4, 3))
    opt = Adam([w], lr=1e-3, betas=(0.9, 0.95), fused=True)              # the reference's call (train.py:175-176)
    ref = torch.optim.Adam([torch.nn.Parameter(w.detach().clone())], lr=1e-3, betas=(0.9, 0.95))
    assert set(ref.state_dict()["param_groups"][0]) <= set(opt.state_dict()["param_groups"][0])
    assert opt._step_supports_amp_scaling
    w.grad = torch.randn(4, 3)
    with pytest.raises(SwinB200Error, match="CUDA"):
        opt.step()
    with pytest.raises(NotImplementedError):
        Adam([w], amsgrad=True)


def test_wrong_image_size_raises_like_reference():
    m = small()
    with pytest.raises(AssertionError, match="doesn't match model"):
        m(torch.randn(1, 7, 64, 144))


def test_unsupported_configs_fail_loudly():
    with pytest.raises(NotImplementedError):
        SwinTransformerV2Cr(img_size=(72, 144), patch_size=8, depths=(1,), num_heads=(1,))
    with pytest.raises(NotImplementedError):
        SwinTransformerV2Cr(img_size=(72, 144), depths=(1, 1), num_heads=(1, 1))
    with pytest.raises(ValueError):
        small(compute_mode="fp8")


def test_loss_handler_parsing():
    base = dict(n_future=1, img_shape_x=72, img_shape_y=144, channel_weights='none', n_out_channels=4,
                channel_names=['a', 'b', 'c', 'd'], out_channels=[0, 1, 2, 3], dt=1, model_grid_type='equiangular')
    h = LossHandler(SimpleNamespace(loss='squared geometric l2', **base))
    assert not h.loss_obj.absolute and h.loss_obj.squared
    assert torch.allclose(h.channel_weights.reshape(-1), torch.full((4,), 0.25))
    assert h.multistep_weight.reshape(-1).tolist() == [0.5, 0.5]
    h = LossHandler(SimpleNamespace(loss='l2', **base))
    assert not h.loss_obj.absolute and not h.loss_obj.squared
    h = LossHandler(SimpleNamespace(loss='pole-masked geometric l1', **base))
    assert h.loss_obj.p == 1 and h.loss_obj.pole_mask == 1 and not h.loss_obj.absolute
    qw = h.loss_obj.quadrature.quad_row_weight
    assert float(qw[0]) == 0.0 and float(qw[-1]) == 0.0 and float(qw[1]) > 0.0
    with pytest.raises(NotImplementedError):
        LossHandler(SimpleNamespace(loss='geometric h1', **base))
    with pytest.raises(ValueError):
        LossHandler(SimpleNamespace(loss='huber', **base))
