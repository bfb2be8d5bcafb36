"""Drop-in replacement for the reference's `model.py` import surface (scripts/segment.py:9,
segment_service.py:10 do `from model import WhisperSegmenter, WhisperSegmenterFast`)."""
from whisperseg_b200.segmenter import (SegmenterBase, WhisperSegmenter, WhisperSegmenterFast,  # noqa: F401
                                       WhisperSegmenterForEval)
