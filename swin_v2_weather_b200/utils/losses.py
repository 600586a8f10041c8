"""Drop-in `LossHandler` / `GeometricLpLoss` (reference: utils/losses.py:31-240) backed by the fused
latitude-weighted L2 kernels.  The loss-type string, channel weighting ('auto' / explicit / temp-std) and the
multistep weighting follow the reference; what is computed on the GPU is

    sum_{b,c} chw_c * f( I[(p-t)^2] / I[t^2] )   (relative)    or    sum_{b,c} chw_c * f( I[(p-t)^2] )   (absolute)

with I[x] = sum_{h,w} x * q_h and f = identity ('squared') or sqrt; the 'l1' family (utils/losses.py:116-124) is the
same reduction over |p-t| (and |t|), and 'pole-masked' zeroes the quadrature rows next to the poles.  The spectral H1
loss needs torch_harmonics and is not part of the path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
from torch import nn

from ..functional import LatWeightedL1Fn, LatWeightedL2Fn
from .grids import GridQuadrature


class GeometricLpLoss(nn.Module):
    """reference: utils/losses.py:154-240 (p in {1, 2})."""

    def __init__(self, img_shape: Tuple[int, int], crop_shape: Tuple[int, int], crop_offset: Tuple[int, int],
                 p: Optional[float] = 2., size_average: Optional[bool] = False, reduction: Optional[bool] = True,
                 absolute: Optional[bool] = False, squared: Optional[bool] = False, pole_mask: Optional[int] = 0,
                 jacobian: Optional[str] = 's2', quadrature_rule: Optional[str] = 'naive'):
        super().__init__()
        if p not in (1, 2):
            raise NotImplementedError("GeometricLpLoss: p must be 1 or 2 (the only values LossHandler ever passes)")
        if size_average or not reduction:
            raise NotImplementedError("size_average / reduction=False are never used by the reference's LossHandler")
        self.p = p
        self.img_shape = img_shape
        self.crop_shape = crop_shape
        self.crop_offset = crop_offset
        self.reduction = reduction
        self.size_average = size_average
        self.absolute = absolute
        self.squared = squared
        self.pole_mask = pole_mask
        self.quadrature = GridQuadrature(quadrature_rule, img_shape=self.img_shape, crop_shape=self.crop_shape,
                                         crop_offset=self.crop_offset, normalize=True, pole_mask=self.pole_mask)

    def _run(self, prd, tar, chw, relative):
        C = prd.shape[1]
        chw = chw.reshape(-1).to(device=prd.device, dtype=torch.float32)
        if chw.numel() != C:
            raise ValueError(f"channel weights have {chw.numel()} entries for {C} channels")
        if prd.shape != tar.shape:
            raise ValueError(f"prediction {tuple(prd.shape)} and target {tuple(tar.shape)} shapes differ")
        qw = self.quadrature.quad_row_weight
        if qw.numel() != prd.shape[2]:
            raise ValueError(f"quadrature has {qw.numel()} rows, the prediction {prd.shape[2]}")
        if self.p == 1:   # |x|^(1/1): `squared` has no effect (utils/losses.py:195-196, 218-219)
            return LatWeightedL1Fn.apply(prd.float(), tar.float(), qw, chw.contiguous(), relative)
        return LatWeightedL2Fn.apply(prd.float(), tar.float(), qw, chw.contiguous(), relative, bool(self.squared))

    def abs(self, prd, tar, chw):
        return self._run(prd, tar, chw, False)

    def rel(self, prd, tar, chw):
        return self._run(prd, tar, chw, True)

    def forward(self, prd, tar, chw):
        return self.abs(prd, tar, chw) if self.absolute else self.rel(prd, tar, chw)


class LossHandler(nn.Module):
    """reference: utils/losses.py:31-150."""

    def __init__(self, params):
        super().__init__()
        self.n_future = params.n_future
        self.img_shape = (params.img_shape_x, params.img_shape_y)
        self.crop_shape = (params.img_shape_x, params.img_shape_y)
        self.crop_offset = (0, 0)
        loss_type = self.loss_type = params.loss
        loss_type = set(loss_type.split())
        pole_mask = 1 if 'pole-masked' in loss_type else 0

        if 'weighted' in loss_type:
            if params.channel_weights == 'auto':
                channel_weights = torch.ones(params.n_out_channels, dtype=torch.float32)
                for c, chn in enumerate(params.channel_names):
                    if chn in ['u10m', 'v10m', 'u100m', 'v100m', 'tp', 'sp', 'msl', 'tcwv']:
                        channel_weights[c] = 0.1
                    elif chn in ['t2m', '2d']:
                        channel_weights[c] = 1.0
                    elif chn[0] in ['z', 'u', 'v', 't', 'r', 'q']:
                        channel_weights[c] = 0.001 * float(chn[1:])
                    else:
                        channel_weights[c] = 0.01
            else:
                channel_weights = torch.Tensor(params.channel_weights).float()
        else:
            channel_weights = torch.ones(params.n_out_channels, dtype=torch.float32)
        channel_weights = channel_weights.reshape(1, -1, 1, 1)
        channel_weights = channel_weights / torch.sum(channel_weights)

        absolute = 'absolute' in loss_type
        squared = 'squared' in loss_type

        if 'temp-std' in loss_type:
            eps = 1e-6
            global_stds = torch.from_numpy(np.load(params.global_stds_path)).reshape(1, -1, 1, 1)[:, params.out_channels]
            time_diff_stds = np.sqrt(params.dt) * torch.from_numpy(np.load(params.time_diff_stds_path)).reshape(1, -1, 1, 1)[:, params.out_channels]
            time_var_weights = global_stds / (time_diff_stds + eps)
            if squared:
                time_var_weights = time_var_weights ** 2
            channel_weights = channel_weights * time_var_weights
        self.register_buffer('channel_weights', channel_weights.float())

        quadrature_rule_type = "legendre-gauss" if params.model_grid_type == "legendre_gauss" else "naive"
        if 'l2' in loss_type:
            if 'geometric' in loss_type:
                self.loss_obj = GeometricLpLoss(self.img_shape, self.crop_shape, self.crop_offset, p=2, absolute=absolute,
                                                squared=squared, pole_mask=pole_mask, quadrature_rule=quadrature_rule_type)
            else:
                # the reference passes jacobian='flat' here but GeometricLpLoss ignores it (utils/losses.py:113-114,168)
                self.loss_obj = GeometricLpLoss(self.img_shape, self.crop_shape, self.crop_offset, p=2, absolute=absolute,
                                                pole_mask=pole_mask, jacobian='flat')
        elif 'l1' in loss_type:
            self.loss_obj = GeometricLpLoss(self.img_shape, self.crop_shape, self.crop_offset, p=1, absolute=absolute,
                                            pole_mask=pole_mask,
                                            **(dict(quadrature_rule=quadrature_rule_type) if 'geometric' in loss_type else dict(jacobian='flat')))
        elif 'geometric h1' in self.loss_type:
            raise NotImplementedError("the spectral H1 loss needs torch_harmonics' RealSHT; it is not on the B200 hot path")
        else:
            raise ValueError(f"Unknown loss function: {self.loss_type}")

        multistep_weight = torch.ones(self.n_future + 1, dtype=torch.float32) / float(self.n_future + 1)
        self.register_buffer('multistep_weight', multistep_weight.reshape(-1, 1, 1, 1))

    def forward(self, prd: torch.Tensor, tar: torch.Tensor, inp: torch.Tensor = None):
        chw = self.channel_weights
        if self.training:
            chw = (chw * self.multistep_weight).reshape(1, -1)
        else:
            chw = chw.reshape(1, -1)
        return self.loss_obj(prd, tar, chw)
