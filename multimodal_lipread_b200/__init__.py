"""multimodal_lipread_b200 -- B200-native audio-visual hot path behind the reference's Python surface.

Importing the package loads the in-tree CUDA library (liblipread_b200.so); if it is missing the
import raises -- there is no CPU or eager-PyTorch fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is absent)

__all__ = ["_lib"]
