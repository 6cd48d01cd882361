"""Mirror of ``brever.modules`` for the hot-path classes (modules/__init__.py:1-32)."""
from .stft import STFT, ConvSTFT, MelFilterbank
from .features import FeatureExtractor

__all__ = ['ConvSTFT', 'STFT', 'MelFilterbank', 'FeatureExtractor']
