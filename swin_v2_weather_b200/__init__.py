"""swin_v2_weather_b200 -- B200-native (sm_100a) training hot path of NERSC/swin_v2_weather.

    from swin_v2_weather_b200.networks.swinv2_global import swinv2net, SwinTransformerV2Cr
    from swin_v2_weather_b200.utils.losses import LossHandler

The CUDA kernels live in `csrc/` behind the C ABI of `include/swinb200.h` (libswinb200.so, built in-tree
by `python -m swin_v2_weather_b200.build`).  There is no CPU / PyTorch fallback for the compute path.
"""
__version__ = "0.1.0"
