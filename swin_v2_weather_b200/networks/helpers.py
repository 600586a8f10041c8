"""Model glue, same surface as the reference `networks/helpers.py` (:7-55)."""
import torch
import torch.nn as nn

from .swinv2_global import swinv2net


class SingleStepWrapper(nn.Module):
    """reference: networks/helpers.py:7-15"""

    def __init__(self, params, model_handle):
        super().__init__()
        self.model = model_handle(params)

    def forward(self, inp, coszen=None):
        return self.model(inp)


class MultiStepWrapper(nn.Module):
    """reference: networks/helpers.py:18-41 (autoregressive rollout, re-appending zenith / invariant channels)"""

    def __init__(self, params, model_handle):
        super().__init__()
        self.model = model_handle(params)
        self.n_future = params.n_future
        self.invar = 1 * params.add_orography + 2 * params.add_landmask

    def forward(self, inp, coszen=None):
        """`inp` may be a tensor (as in the reference) or a tuple of channel groups from the fused PreProcessor.  The
        re-appended zenith / invariant channels are handed to the model as separate groups: the reference's two
        torch.cat per rollout step (2 x 319 MB written and re-read at 77 x 720 x 1440) never happen."""
        result = []
        inpt = inp
        if isinstance(inp, (tuple, list)):
            flat = None if (self.invar == 0 or inp[-1].shape[1] == self.invar) else torch.cat(
                [t.expand(inp[0].shape[0], -1, -1, -1) for t in inp], dim=1)
            invars = None if not self.invar else (inp[-1] if flat is None else flat[:, -self.invar:, :, :].contiguous())
        else:
            invars = inp[:, -self.invar:, :, :].contiguous() if self.invar else None
        for step in range(self.n_future + 1):
            pred = self.model(inpt)
            result.append(pred)
            if step == self.n_future:
                break
            groups = [pred]
            if coszen is not None:
                groups.append(coszen[:, step:step + 1, :, :].contiguous())
            if self.invar:
                groups.append(invars)
            inpt = groups if len(groups) > 1 else pred
        return torch.cat(result, dim=1)


def get_model(params):
    if params.nettype != 'swin':
        raise Exception(f"model type {params.nettype} not implemented")
    if params.n_future > 0:
        return MultiStepWrapper(params, swinv2net)
    return SingleStepWrapper(params, swinv2net)
