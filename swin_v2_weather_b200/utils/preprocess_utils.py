"""Input conditioning, same surface as the reference `utils/preprocess_utils.py:5-68`.

The reference concatenates [field (73 ch) | zenith (1) | one-hot land mask (2) | normalised orography (1)] into a new
(B, 77, 720, 1440) tensor every step.  Here `forward` returns the pieces as a tuple of channel groups (static features
with a leading dimension of 1, shared by the batch); the model's PatchEmbed im2col (`swinb200_patchify_cat`) reads them
where they are.  `params.fuse_conditioning = False` restores the reference's concatenated tensor.

Static fields come from `params.landmask` / `params.orography` (arrays or tensors already in memory) or, as in the
reference, from `params.landmask_path` / `params.orography_path` through h5py / netCDF4 when those are installed
(`utils/conditioning_inputs.py:23-41`; file I/O is outside this package's scope).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def _load(params, attr, path_attr, reader):
    val = getattr(params, attr, None)
    if val is not None:
        return np.asarray(val.cpu() if torch.is_tensor(val) else val)
    return reader(getattr(params, path_attr))


def _read_land_mask(path):
    import h5py                                           # reference: conditioning_inputs.py:35-41
    with h5py.File(path, "r") as f:
        return f["LSM"][0, :, :]


def _read_orography(path):
    from netCDF4 import Dataset as DS                     # reference: conditioning_inputs.py:23-32
    with DS(path, "r") as f:
        oro = f.variables["Z"][0, :, :]
        return (oro - oro.min()) / (oro.max() - oro.min())


class PreProcessor(nn.Module):
    def __init__(self, params, device):
        super().__init__()
        self.params = params
        self.device = device
        imgx, imgy = params.img_size
        static_features = None
        if params.add_landmask:
            with torch.no_grad():
                lsm = torch.tensor(_load(params, "landmask", "landmask_path", _read_land_mask), dtype=torch.long)
                lsm = torch.permute(torch.nn.functional.one_hot(lsm), (2, 0, 1)).to(torch.float32)
                lsm = torch.reshape(lsm, (1, lsm.shape[0], lsm.shape[1], lsm.shape[2]))[:, :, :imgx, :imgy]
                static_features = lsm
        if params.add_orography:
            with torch.no_grad():
                oro = torch.tensor(_load(params, "orography", "orography_path", _read_orography), dtype=torch.float32)
                oro = torch.reshape(oro, (1, 1, oro.shape[0], oro.shape[1]))[:, :, :imgx, :imgy]
                oro = (oro - torch.mean(oro)) / (torch.std(oro) + 1.0e-6)
                static_features = oro if static_features is None else torch.cat([static_features, oro], dim=1)
        self.do_add_static_features = static_features is not None
        if self.do_add_static_features:
            self.register_buffer("static_features", static_features.contiguous(), persistent=False)
        self.fuse = bool(getattr(params, "fuse_conditioning", True))

    def forward(self, data):
        if self.params.add_zenith:
            inp, tar, izen, tzen = map(lambda x: x.to(self.device, dtype=torch.float), data)
            groups = [inp, izen]
        else:
            inp, tar = map(lambda x: x.to(self.device, dtype=torch.float), data)
            tzen = None
            groups = [inp]
        if self.do_add_static_features:
            groups.append(self.static_features)
        if len(groups) == 1:
            return inp, tar, tzen
        if self.fuse:
            return tuple(groups), tar, tzen
        B = inp.shape[0]
        return torch.cat([g.expand(B, -1, -1, -1) for g in groups], dim=1), tar, tzen
