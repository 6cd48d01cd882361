"""brever_b200 — B200-native (sm_100a) time-frequency front-end for brever.

Drop-in replacements for the hot-path classes of philgzl/brever:

    from brever_b200.modules import STFT, MelFilterbank, FeatureExtractor
    from brever_b200.criterion import CriterionRegistry, init_criterion, sisnr, snr

backed by hand-written CUDA kernels behind a C ABI (include/brever_b200.h).
CUDA tensors only; there is no CPU fallback.
"""
from . import criterion, ffnn, manner, metrics, modules, transforms
from .criterion import (CriterionRegistry, MultiResYuLoss, apply_mask, init_criterion, mse,
                        sisnr, snr)
from .modules import STFT, ConvSTFT, FeatureExtractor, MelFilterbank
from .registry import Registry
from .transforms import transform_batched

__version__ = '0.1.0'
__all__ = ['STFT', 'ConvSTFT', 'MelFilterbank', 'FeatureExtractor', 'CriterionRegistry',
           'init_criterion', 'sisnr', 'snr', 'mse', 'MultiResYuLoss', 'apply_mask', 'Registry',
           'criterion', 'ffnn', 'manner', 'metrics', 'modules', 'transforms', 'transform_batched']
