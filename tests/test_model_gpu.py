"""End-to-end parity of the drop-in model + loss against the CPU oracle (and, through it, the reference).

fp32 mode: outputs and every parameter gradient within 1e-5 relative L2 (north-star fp32 bound; a few
ill-conditioned tensors get 5e-5).  bf16 modes: within 1e-2 (north-star bf16 bound); `logit_scale` (8 numbers
per block, a sum of softmax-gradient x cosine products that largely cancel) gets 5e-2 at this 4-window test size:
the reference's own bf16 autocast is 2.7-3.2e-2 away from its fp32 there and the noise averages
down with the window count (400 per sample at full size).
`meta_mlp.fc2.bias` has an analytically zero gradient (softmax shift invariance) -> absolute check.
`meta_mlp.fc{1,2}.*` (CPB) gradients are sums of softmax gradients that cancel row-wise; at this test size
(4 windows) bf16 storage noise does not average out: 3e-2, the reference autocast's own distance (F9).
"""
import os

import pytest
import torch

from oracle import swinv2_oracle as O
from swin_v2_weather_b200 import ops
from swin_v2_weather_b200.functional import LatWeightedL2Fn
from swin_v2_weather_b200.networks.swinv2_global import SwinTransformerV2Cr

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def build(cfg: O.SwinConfig, sd, mode: str, **kw):
    m = SwinTransformerV2Cr(img_size=cfg.img_size, patch_size=cfg.patch_size, depths=(cfg.depth,), num_heads=(cfg.num_heads,),
                            in_chans=cfg.in_chans, out_chans=cfg.out_chans, embed_dim=cfg.embed_dim,
                            img_window_ratio=cfg.window_ratio, drop_path_rate=cfg.drop_path_rate,
                            full_pos_embed=cfg.full_pos_embed, rel_pos=cfg.rel_pos, mlp_ratio=cfg.mlp_ratio,
                            residual=cfg.residual, compute_mode=mode, **kw)
    m.load_state_dict(sd)
    return m.cuda()


def inputs(cfg, batch, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, cfg.in_chans, *cfg.img_size, generator=g)
    tar = torch.randn(batch, cfg.out_chans, *cfg.img_size, generator=g)
    chw = torch.rand(cfg.out_chans, generator=g) + 0.5
    return x, tar, chw / chw.sum()


def run_ours(model, x, tar, chw, relative):
    qw = O.quadrature_row_weights(*x.shape[-2:]).cuda()
    pred = model(x.cuda())
    loss = LatWeightedL2Fn.apply(pred, tar.cuda(), qw, chw.cuda(), relative, True)
    loss.backward()
    grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
    return pred.detach().cpu(), float(loss), grads


def check(pred, loss, grads, pred_ref, loss_ref, grads_ref, tol, tol_scale):
    report = {"pred": O.rel_l2(pred, pred_ref)}
    assert report["pred"] < tol, report
    assert abs(loss - float(loss_ref)) / abs(float(loss_ref)) < tol
    bad = []
    for k, g_ref in grads_ref.items():
        if k.endswith("meta_mlp.fc2.bias"):
            if grads[k].abs().max() > 1e-3 * max(1.0, float(max(v.abs().max() for v in grads_ref.values()))):
                bad.append((k, "abs", float(grads[k].abs().max())))
            continue
        e = O.rel_l2(grads[k], g_ref)
        lim = tol_scale if ("logit_scale" in k or "meta_mlp" in k) else tol
        if not e < lim:
            bad.append((k, e))
    assert not bad, bad


CFGS = {
    "nopos": dict(img_size=(72, 144), depth=3, num_heads=2, in_chans=7, out_chans=5, embed_dim=192, window_ratio=8, rel_pos=False),
    "cpb_residual": dict(img_size=(72, 144), depth=2, num_heads=2, in_chans=7, out_chans=5, embed_dim=192, window_ratio=8,
                         rel_pos=True, residual=True),
}


@pytest.mark.parametrize("name", list(CFGS))
@pytest.mark.parametrize("mode", ["fp32", "bf16_simt", "bf16"])
def test_model_parity_vs_oracle(name, mode):
    cfg = O.SwinConfig(**CFGS[name])
    sd = O.init_state_dict(cfg, seed=3)
    x, tar, chw = inputs(cfg, 2)
    relative = name == "nopos"
    pred_ref, loss_ref, grads_ref = O.loss_and_grads(x, tar, sd, cfg, chw, relative=relative)
    model = build(cfg, sd, mode).eval()
    pred, loss, grads = run_ours(model, x, tar, chw, relative)
    tol, tol_scale = (1e-5, 5e-5) if mode == "fp32" else (1e-2, 5e-2)
    check(pred, loss, grads, pred_ref, loss_ref, grads_ref, tol, tol_scale)


@pytest.mark.parametrize("case", ["nopos_rel", "cpb_abs_residual"])
def test_model_fp32_vs_reference_golden(case):
    """fp32 mode against fixtures produced by the unmodified reference (oracle/make_golden.py)."""
    fix = torch.load(os.path.join(GOLDEN, f"model_{case}.pt"), weights_only=False)
    cfg = O.SwinConfig(**fix["config"])
    sd = O.init_state_dict(cfg, seed=1)
    x, tar, chw = inputs(cfg, fix["batch"])
    model = build(cfg, sd, "fp32").eval()
    pred, loss, grads = run_ours(model, x, tar, chw, fix["loss_kind"] == "rel")
    assert O.rel_l2(pred, fix["pred"]) < 1e-5
    assert abs(loss - float(fix["loss"])) / float(fix["loss"]) < 1e-5
    for k, n in fix["grad_norm"].items():
        if k.endswith("meta_mlp.fc2.bias"):
            continue
        assert abs(float(grads[k].double().norm()) - float(n)) / float(n) < 5e-5, k
        idx = torch.randint(0, grads[k].numel(), (min(64, grads[k].numel()),), generator=torch.Generator().manual_seed(grads[k].numel()))
        got = grads[k].reshape(-1)[idx]
        assert (got - fix["grad_sample"][k]).norm() <= 5e-5 * float(n) + 1e-5 * fix["grad_sample"][k].norm(), k


def test_drop_path_and_checkpoint_consistency():
    """DropPath(train) + activation checkpointing: checkpointed and plain runs agree on the loss and on every
    gradient (same RNG draws replayed by torch.utils.checkpoint)."""
    cfg = O.SwinConfig(**dict(CFGS["nopos"], drop_path_rate=0.5))
    sd = O.init_state_dict(cfg, seed=3)
    x, tar, chw = inputs(cfg, 4)
    outs = []
    for ckpt in (False, True):
        model = build(cfg, sd, "bf16_simt", checkpoint_stages=ckpt).train()
        torch.manual_seed(11)
        torch.cuda.manual_seed(11)
        outs.append(run_ours(model, x, tar, chw, True))
    assert abs(outs[0][1] - outs[1][1]) <= 1e-6 * abs(outs[0][1])   # fp32 atomics: reduction order differs run to run
    for k in outs[0][2]:
        assert O.rel_l2(outs[1][2][k], outs[0][2][k]) < 5e-3, k   # fp32-atomic order can flip individual bf16 roundings


def test_drop_path_matches_oracle_with_same_masks():
    cfg = O.SwinConfig(**dict(CFGS["nopos"], drop_path_rate=0.4))
    sd = O.init_state_dict(cfg, seed=3)
    x, tar, chw = inputs(cfg, 4)
    model = build(cfg, sd, "fp32").train()
    torch.manual_seed(5)
    torch.cuda.manual_seed(5)
    pred, loss, grads = run_ours(model, x, tar, chw, True)
    # replay the same CUDA draws for the oracle: blocks with drop_prob > 0 draw (B,1,1,1) then (B,1,1)
    torch.manual_seed(5)
    torch.cuda.manual_seed(5)
    masks = []
    for i in range(cfg.depth):
        p = cfg.drop_path(i)
        if p == 0.0:
            masks.append(None)
            continue
        keep = 1 - p
        m1 = torch.empty((4, 1, 1, 1), device="cuda").bernoulli_(keep).div_(keep).reshape(-1).cpu()
        m2 = torch.empty((4, 1, 1), device="cuda").bernoulli_(keep).div_(keep).reshape(-1).cpu()
        masks.append((m1, m2))
    pred_ref, loss_ref, grads_ref = O.loss_and_grads(x, tar, sd, cfg, chw, relative=True, drop_masks=masks)
    check(pred, loss, grads, pred_ref, loss_ref, grads_ref, 1e-5, 5e-5)


def test_attn_mask_property_bit_exact():
    cfg = O.SwinConfig(**CFGS["nopos"])
    model = build(cfg, O.init_state_dict(cfg, seed=3), "fp32")
    for i, blk in enumerate(model.stages[0].blocks):
        want = O.shift_attention_mask(cfg.grid, cfg.window, cfg.shift(i))
        got = blk.attn_mask
        assert (want is None) == (got is None)
        if want is not None:
            assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_multistep_rollout_gradients(mode):
    """MultiStepWrapper (networks/helpers.py:18-41): the second step consumes the first step's prediction, so gradients
    flow through the skip connection and the PatchEmbed im2col back into the first step."""
    from types import SimpleNamespace
    from swin_v2_weather_b200.networks.helpers import get_model
    cfg = O.SwinConfig(img_size=(72, 144), depth=2, num_heads=2, in_chans=6, out_chans=5, embed_dim=192, window_ratio=8,
                       rel_pos=False, residual=True)
    params = SimpleNamespace(nettype='swin', n_future=1, add_orography=True, add_landmask=False, img_size=[72, 144], patch_size=4,
                             depth=2, num_heads=2, n_in_channels=6, n_out_channels=5, embed_dim=192, window_ratio=8,
                             drop_path_rate=0.0, full_pos_embed=True, rel_pos=False, mlp_ratio=4, activation_ckpt=False,
                             residual=True, compute_mode=mode)
    sd = O.init_state_dict(cfg, seed=3)
    wrapper = get_model(params)
    wrapper.model.load_state_dict(sd)
    wrapper = wrapper.cuda().eval()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 6, 72, 144, generator=g)
    tar = torch.randn(2, 10, 72, 144, generator=g)
    chw = torch.ones(10) / 10
    # oracle rollout
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    p1 = O.model_forward(x, leaves, cfg)
    p2 = O.model_forward(torch.cat([p1, x[:, -1:]], dim=1), leaves, cfg)
    loss_ref = O.geometric_l2(torch.cat([p1, p2], dim=1), tar, chw, True)
    grads_ref = dict(zip(leaves.keys(), torch.autograd.grad(loss_ref, list(leaves.values()))))
    # ours
    pred = wrapper(x.cuda())
    qw = O.quadrature_row_weights(72, 144).cuda()
    loss = LatWeightedL2Fn.apply(pred, tar.cuda(), qw, chw.cuda(), True, True)
    loss.backward()
    grads = {k: p.grad.detach().cpu() for k, p in wrapper.model.named_parameters()}
    tol = 1e-5 if mode == "fp32" else 1e-2
    check(pred.detach().cpu(), float(loss), grads, torch.cat([p1, p2], dim=1).detach(), loss_ref.detach(), grads_ref, tol,
          5e-5 if mode == "fp32" else 5e-2)


def test_full_resolution_batch_independence():
    """BASELINE size (73 channels, 720 x 1440, C = 768, 8 heads), one block: no kernel mixes samples, so sample 0 of a
    two-sample batch must reproduce the single-sample run bit for bit in the forward (no atomics there) and to fp32
    accumulation-order noise in the gradients; the batch loss is the sum of the per-sample losses (losses.py:200-204)."""
    torch.manual_seed(0)
    kw = dict(img_size=(720, 1440), patch_size=4, depths=(1,), num_heads=(8,), in_chans=73, out_chans=73, embed_dim=768,
              img_window_ratio=80, drop_path_rate=0.0, full_pos_embed=True, rel_pos=False, mlp_ratio=4.0, residual=False)
    model = SwinTransformerV2Cr(compute_mode="bf16", **kw).cuda().eval()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 73, 720, 1440, generator=g).cuda()
    t = torch.randn(2, 73, 720, 1440, generator=g).cuda()
    qw = O.quadrature_row_weights(720, 1440).cuda()
    chw = torch.full((73,), 1.0 / 73, device="cuda")

    def run(xb, tb):
        model.zero_grad(set_to_none=True)
        pred = model(xb)
        loss = LatWeightedL2Fn.apply(pred, tb, qw, chw, True, True)
        loss.backward()
        return pred.detach(), float(loss.detach()), {k: p.grad.detach().clone() for k, p in model.named_parameters()}

    p2, l2, g2 = run(x, t)
    p0, l0, g0 = run(x[:1].contiguous(), t[:1].contiguous())
    p1, l1, g1 = run(x[1:].contiguous(), t[1:].contiguous())
    assert torch.equal(p2[:1], p0) and torch.equal(p2[1:], p1)
    assert abs(l2 - (l0 + l1)) / abs(l2) < 1e-6
    for k in g2:
        assert O.rel_l2(g2[k].cpu(), (g0[k] + g1[k]).cpu()) < 2e-3, k


def test_channel_groups_equal_concatenation():
    """Conditioning inputs: feeding [field | zenith | static features] as channel groups must give
    bit-identical predictions and parameter gradients to the reference's torch.cat input -- same im2col values, no copy --
    including the batch-shared static group, the residual skip, and the MultiStepWrapper rollout."""
    from types import SimpleNamespace
    from swin_v2_weather_b200.networks.helpers import MultiStepWrapper
    from swin_v2_weather_b200.utils.preprocess_utils import PreProcessor
    torch.manual_seed(1)
    B, H, W = 2, 72, 144
    kw = dict(img_size=(H, W), patch_size=4, depths=(2,), num_heads=(2,), in_chans=9, out_chans=5, embed_dim=192,
              img_window_ratio=8, drop_path_rate=0.0, full_pos_embed=True, rel_pos=False, residual=True)
    model = SwinTransformerV2Cr(compute_mode="bf16", **kw).cuda().eval()
    g = torch.Generator().manual_seed(7)
    field, zen = torch.randn(B, 5, H, W, generator=g).cuda(), torch.rand(B, 1, H, W, generator=g).cuda()
    lsm = (torch.rand(H, W, generator=g) > 0.7).long()
    oro = torch.randn(H, W, generator=g)
    params = SimpleNamespace(img_size=(H, W), add_landmask=True, add_orography=True, add_zenith=True, landmask=lsm, orography=oro)
    pre = PreProcessor(params, "cuda").cuda()
    tar = torch.randn(B, 5, H, W, generator=g).cuda()
    groups, tar2, tzen = pre((field, tar, zen, zen))
    assert isinstance(groups, tuple) and [t.shape[1] for t in groups] == [5, 1, 3] and groups[2].shape[0] == 1
    params_cat = SimpleNamespace(**{**vars(params), "fuse_conditioning": False})
    cat, _, _ = PreProcessor(params_cat, "cuda").cuda()((field, tar, zen, zen))
    assert cat.shape == (B, 9, H, W)

    def run(inp):
        model.zero_grad(set_to_none=True)
        out = model(inp)
        out.square().mean().backward()
        return out.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}

    o1, g1 = run(groups)
    o2, g2 = run(cat)
    assert torch.equal(o1, o2)
    for k in g1:
        assert O.rel_l2(g1[k].cpu(), g2[k].cpu()) < 1e-5, k      # fp32 atomics reorder; the inputs to every kernel are identical

    # rollout: groups re-appended per step instead of two torch.cat per step
    wrap = MultiStepWrapper(SimpleNamespace(n_future=1, add_orography=True, add_landmask=True), lambda p: model)
    coszen = torch.rand(B, 1, H, W, generator=g).cuda()
    model.zero_grad(set_to_none=True)
    r1 = wrap(groups, coszen)
    r1.square().mean().backward()
    gr1 = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    r2 = wrap(cat, coszen)
    r2.square().mean().backward()
    assert torch.equal(r1, r2)
    for k in gr1:
        assert O.rel_l2(gr1[k].cpu(), model.get_parameter(k).grad.cpu()) < 2e-3, k
