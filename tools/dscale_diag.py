"""bring-up: d(scale) of the attention backward at full geometry -- tcgen05 vs CUDA-core vs fp32 autograd, per head"""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_kernels_gpu import oracle_attention, gen
from swin_v2_weather_b200 import ops
from swin_v2_weather_b200._lib import BACKEND_TCGEN05, BACKEND_SIMT

B, H, W, C, heads, window = 1, 180, 360, 768, 8, (9, 18)
T = B * H * W
for shift in ((0, 0), (4, 9)):
    raw = gen(T, 3 * C, seed=60).to(torch.bfloat16)
    scale = torch.linspace(8.0, 14.0, heads, device="cuda")
    raw_f = raw.float().requires_grad_(True)
    sc_f = scale.clone().requires_grad_(True)
    o_ref, lse_ref = oracle_attention(raw_f, sc_f, None, B, H, W, C, heads, window, shift)
    qkv = raw.clone()
    inv = ops.qk_normalize_(qkv, C, heads)
    d_o = gen(T, C, seed=62).to(torch.bfloat16)
    o_ref.backward(d_o.float())
    res = {}
    for name, be in (("tcgen05", BACKEND_TCGEN05), ("simt", BACKEND_SIMT)):
        o, lse = ops.window_attn_fwd(qkv, scale, None, B, H, W, C, heads, 9, 18, shift[0], shift[1], ops.MODE_BF16, backend=be)
        dqkv, dscale, _ = ops.window_attn_bwd(qkv, inv, scale, None, o, d_o, lse, B, H, W, C, heads, 9, 18, shift[0], shift[1], ops.MODE_BF16, backend=be)
        res[name] = dscale
        print(shift, name, "rel", ((dscale - sc_f.grad).norm() / sc_f.grad.norm()).item(), "ratio", (dscale / sc_f.grad).tolist())
    # mixed: tcgen05 backward fed by the CUDA-core forward (o, lse) and vice versa
    o_s, lse_s = ops.window_attn_fwd(qkv, scale, None, B, H, W, C, heads, 9, 18, shift[0], shift[1], ops.MODE_BF16, backend=BACKEND_SIMT)
    _, ds_mix, _ = ops.window_attn_bwd(qkv, inv, scale, None, o_s, d_o, lse_s, B, H, W, C, heads, 9, 18, shift[0], shift[1], ops.MODE_BF16, backend=BACKEND_TCGEN05)
    print(shift, "tc bwd on simt fwd: rel", ((ds_mix - sc_f.grad).norm() / sc_f.grad.norm()).item())
    print("   ref", sc_f.grad.tolist())
