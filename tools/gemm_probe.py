"""Bring-up probe for the tcgen05 GEMM: runs each operand-major / epilogue case in its own process so that
a trap in one variant does not poison the others, and prints error structure (not just pass/fail)."""
import subprocess
import sys

CASES = {
    "kk_small": (128, 128, 64, 0, 0, 4),
    "kk_k128": (128, 128, 128, 0, 0, 4),
    "kk_n256": (128, 256, 256, 0, 0, 4),
    "kk_multi": (1000, 2304, 768, 0, 0, 0),
    "kn_small": (128, 128, 64, 0, 1, 4),
    "kn_n256": (256, 256, 256, 0, 1, 4),
    "mn_small": (128, 128, 64, 1, 1, 4),
    "mn_n256": (256, 256, 256, 1, 1, 4),
    "mn_wgrad": (768, 3072, 1000, 1, 1, 4),
}


def run_case(name):
    import torch
    sys.path.insert(0, ".")
    from swin_v2_weather_b200 import ops
    from swin_v2_weather_b200._lib import BACKEND_TCGEN05
    M, N, K, am, bm, epi = CASES[name]
    g = torch.Generator().manual_seed(0)
    A = (torch.randn((M, K) if am == 0 else (K, M), generator=g) * 0.5).cuda().bfloat16()
    B = (torch.randn((N, K) if bm == 0 else (K, N), generator=g) * 0.5).cuda().bfloat16()
    bias = torch.zeros(N, device="cuda") if epi == 0 else None
    out = ops.gemm(ops.MODE_BF16, A, am, B, bm, epi, bias=bias, backend=BACKEND_TCGEN05)
    torch.cuda.synchronize()
    Af = A.float() if am == 0 else A.float().t()
    Bf = B.float() if bm == 0 else B.float().t()
    want = Af @ Bf.t()
    got = out.float()
    err = (got - want).abs()
    print(f"{name}: rel={float(err.norm() / want.norm()):.3e} max={float(err.max()):.3e} "
          f"zeros={float((got == 0).float().mean()):.3f} nan={int(torch.isnan(got).sum())}")
    if float(err.norm() / want.norm()) > 2e-2:
        # structure: error by row block of 8 and by column block of 8 (first 64)
        rb = err[:64].view(8, 8, -1).mean(dim=(1, 2))
        cb = err[:, :64].reshape(err.shape[0], 8, 8).mean(dim=(0, 2))
        print("  row-block err:", [f"{float(v):.2f}" for v in rb])
        print("  col-block err:", [f"{float(v):.2f}" for v in cb])
        # does it match a k-permuted / partial-k product?  compare against using only the first 16/32/48 of each 64-k block
        for kk in (16, 32, 48):
            mask = (torch.arange(K, device="cuda") % 64) < kk
            part = (Af * mask) @ Bf.t()
            print(f"  vs first-{kk}-of-64 k: rel={float((got - part).norm() / want.norm()):.3e}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for name in CASES:
            r = subprocess.run(["timeout", "90", sys.executable, __file__, name], capture_output=True, text=True)
            tail = (r.stdout + r.stderr).strip().splitlines()[-12:]
            print(f"--- {name} (exit {r.returncode})")
            print("\n".join(tail))
