"""Latitude quadrature weights for the loss (reference: utils/grids.py:62-117, 'naive' rule only).

The reference builds a full (1, 1, H, W) table whose rows are constant along longitude; the loss kernel
only needs the H row weights.  The table is evaluated with the same fp32 op sequence as the reference so
the weights are reproduced bit for bit (checked in tests against the oracle / golden fixtures).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class GridQuadrature(nn.Module):
    def __init__(self, quadrature_rule, img_shape, crop_shape=None, crop_offset=(0, 0), normalize=False, pole_mask=None):
        super().__init__()
        if quadrature_rule != 'naive':
            raise NotImplementedError(
                f"quadrature rule {quadrature_rule!r}: only 'naive' (equiangular grid) is on the hot path; "
                "clenshaw-curtiss / legendre-gauss need torch_harmonics and are not used by any shipped config")
        nlat, nlon = int(img_shape[0]), int(img_shape[1])
        jacobian = torch.clamp(torch.sin(torch.linspace(0, torch.pi, nlat)), min=0.)
        dA = (2 * torch.pi / nlon) * (torch.pi / nlat)
        w = (dA * jacobian.unsqueeze(1)).tile(1, nlon)
        w = w * (4. * torch.pi) / torch.sum(w)
        if normalize:
            w = w / (4. * torch.pi)
        if (pole_mask is not None) and (pole_mask > 0):
            # reference utils/grids.py:96-99 zeroes the rows next to both poles; upstream indexes an undefined `sizes`
            # there (NameError on any 'pole-masked' loss) -- the evident intent, `img_shape[0]`, is what runs here
            w[:pole_mask, :] = 0.
            w[nlat - pole_mask:, :] = 0.
        if crop_shape is not None:
            w = w[crop_offset[0]:crop_offset[0] + crop_shape[0], crop_offset[1]:crop_offset[1] + crop_shape[1]]
        self.shape = tuple(w.shape)
        self.register_buffer('quad_row_weight', w[:, 0].contiguous())

    @property
    def quad_weight(self) -> torch.Tensor:
        """(1, 1, H, W) view with the reference's layout (rows constant along longitude)."""
        H, W = self.shape
        return self.quad_row_weight.view(1, 1, H, 1).expand(1, 1, H, W)
