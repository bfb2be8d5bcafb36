"""whisperseg_b200 -- B200-native implementation of WhisperSeg's segmentation hot path."""
__all__ = ["WhisperSegmenter", "WhisperSegmenterFast", "WhisperSegmenterForEval"]


def __getattr__(name):
    if name in __all__:
        from . import segmenter
        return getattr(segmenter, name)
    raise AttributeError(name)
