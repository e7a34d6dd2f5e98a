"""
amt_tools_b200 -- B200-native (sm_100a) implementation of the `amt_tools.features` front end:
STFT, MelSpec, CQT, VQT, HCQT / HVQT, SignalPower, WaveformWrapper and FeatureCombo behind the
reference's FeatureModule API, computed by libamtfeat.so (hand-written CUDA, C-ABI in include/amtfeat.h).

Import patterns of the reference keep working with the package name swapped:
    import amt_tools_b200.features as ft;  ft.CQT()
    from amt_tools_b200.features import HCQT
"""

from . import _lib  # noqa: F401  (fails loudly when libamtfeat.so has not been built)
from . import features
from .features import (FeatureCombo, FeatureModule, CQT, HCQT, HVQT, MelSpec, SignalPower, STFT, VQT,
                       WaveformWrapper, framify_activations)
from .stream import AudioStream, FeatureStream
from . import ingest
from .ingest import load_normalize_audio, pcm16_to_float, resample, rms_norm, to_mono
from . import longtrack
from .longtrack import process_long_audio

__version__ = '0.1.0'
