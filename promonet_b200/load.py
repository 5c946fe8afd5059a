"""Loading of cached features: promonet.load.ppg (promonet/load.py:172-188)"""
import torch

from promonet_b200 import config
from promonet_b200.edit import grid


def ppg(file, resample_length=None, device=None):
    """Load a PPG file (40, F) and maybe resample it to `resample_length` frames"""
    device = torch.device('cuda', torch.cuda.current_device()) if device is None else device
    result = torch.load(file, map_location='cpu').to(device)
    if resample_length is not None and result.shape[-1] != resample_length:
        # load.py:183-186 renormalises only when REPRESENTATION_KIND == 'ppgs', which the
        # 'ppg' representation of config/promonet.py never is: plain interpolation
        result = grid.sample(result, grid.of_length(result, resample_length), config.PPG_INTERP_METHOD)
    return result
