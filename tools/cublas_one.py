"""cuBLAS reference for the ncu comparison:  python tools/cublas_one.py {fc2|fc1}"""
import sys, torch
T, C, HID = 64800, 768, 3072
bf = lambda *s: (torch.randn(*s, device="cuda") * 0.5).bfloat16()
if sys.argv[1] == "fc2":
    a, b = bf(T, HID), bf(HID, C)
else:
    a, b = bf(T, C), bf(C, HID)
for _ in range(4):
    torch.matmul(a, b)
torch.cuda.synchronize()
