"""Host -> device staging of ERA5 fields (reference: the loaders crop the 721-row files to 720 rows before the model,
utils/data_loader_era5.py / data_loader_era5_dali.py `[:, :img_shape_x]`).

`copy_cropped_async` moves a pinned (B, C, rows_src, W) fp32 host tensor into a (B, C, rows, W) device tensor with ONE
strided cudaMemcpy2DAsync (each plane's first `rows` rows are one contiguous run; the plane pitch differs between source
and destination), instead of B*C plane copies or a host-side repack."""
from __future__ import annotations

import torch


def copy_cropped_async(dst: torch.Tensor, src: torch.Tensor, stream: torch.cuda.Stream) -> None:
    B, C, rows, W = dst.shape
    if not (src.dim() == 4 and src.shape[0] == B and src.shape[1] == C and src.shape[3] == W and src.shape[2] >= rows):
        raise ValueError(f"copy_cropped_async: source {tuple(src.shape)} does not cover destination {tuple(dst.shape)}")
    if not (src.is_pinned() and src.is_contiguous() and dst.is_contiguous() and src.dtype == dst.dtype):
        raise ValueError("copy_cropped_async: the source must be pinned and contiguous, dtypes must match")
    esz = dst.element_size()
    try:
        from cuda.bindings import runtime as cudart
    except ImportError:
        try:
            from cuda import cudart                         # older cuda-python layout
        except ImportError:                                 # no cuda-python: one contiguous copy per plane, same result
            with torch.cuda.stream(stream):
                for b in range(B):
                    for c in range(C):
                        dst[b, c].copy_(src[b, c, :rows], non_blocking=True)
            return
    err, = cudart.cudaMemcpy2DAsync(dst.data_ptr(), rows * W * esz, src.data_ptr(), src.shape[2] * W * esz, rows * W * esz, B * C,
                                    cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, stream.cuda_stream)
    if int(err) != 0:
        raise RuntimeError(f"cudaMemcpy2DAsync failed: {err}")
