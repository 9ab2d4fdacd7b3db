"""MSELoss with the reference's interface (loss.py:3-11), computed by r4r_mse_fwd / r4r_mse_bwd."""
import torch

from . import ops


class MSELoss(torch.nn.Module):
    def __init__(self, hyper_params=None):
        super().__init__()

    def forward(self, output, y, return_mean=True):
        se = ops.squared_error(output, y)
        return se.mean() if return_mean else se
