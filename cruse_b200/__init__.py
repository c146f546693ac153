"""cruse_b200 -- the CRUSE speech-enhancement hot path as hand-written sm_100a CUDA kernels.

Public surface (mirrors the reference's plugin API):
    cruse_b200.cruse_net.unet_2 / GGRU     model/cruse_net.py
    cruse_b200.acoustics.stft / istft / PreProcess   train_base/acoustics/feature.py, utils/utils.py
    cruse_b200.loss.loss_func / wo_male    loss_func/loss.py
    cruse_b200.pipeline.enhance / forward_loss       the fused end-to-end path
All compute goes through ``libcruse_sm100.so`` (C ABI: include/cruse_b200.h); importing the
package does not need a GPU, calling any op does.
"""
from ._lib import LIB_PATH, lib  # noqa: F401

__version__ = "0.1.0"
