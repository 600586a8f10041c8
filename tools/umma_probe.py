"""Runs the un-swizzled / tensor-memory operand probe for every operand form, one process per case."""
import subprocess
import sys


def run_case(a_mode, b_mode, N, K, pad):
    import torch
    sys.path.insert(0, ".")
    from swin_v2_weather_b200 import _lib
    g = torch.Generator().manual_seed(0)
    A = (torch.randn(128, K, generator=g) * 0.5).cuda().bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.5).cuda().bfloat16()
    A_st = A.t().contiguous() if a_mode == 1 else A
    B_st = B.t().contiguous() if b_mode == 1 else B
    D = torch.zeros(128, N, device="cuda")
    _lib.call("swinb200_debug_umma_probe", A_st.data_ptr(), B_st.data_ptr(), D.data_ptr(), N, K, a_mode, b_mode, pad,
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = A.float() @ B.float().t()
    err = float((D - want).norm() / want.norm())
    print(f"a_mode={a_mode} b_mode={b_mode} N={N} K={K} pad={pad}: rel={err:.3e} zeros={float((D == 0).float().mean()):.2f}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(*map(int, sys.argv[1:]))
    else:
        for a_mode in (0, 1, 2):
            for b_mode in (0, 1):
                for (N, K, pad) in ((176, 96, 0), (96, 176, 1), (176, 128, 1)):
                    if a_mode == 1 and K % 8:
                        continue
                    r = subprocess.run(["timeout", "60", sys.executable, __file__, str(a_mode), str(b_mode), str(N), str(K), str(pad)],
                                       capture_output=True, text=True)
                    out = (r.stdout + r.stderr).strip().splitlines()
                    print(f"[exit {r.returncode}]", out[-1] if out else "")
