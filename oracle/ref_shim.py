"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference.

Imports `/root/reference/model.py` and `audio_utils.py` as-is (container only; the GPU box
has no /root/reference).  Used by `oracle/gen_golden.py` to produce `tests/golden/*` and by
the CPU tests (when the reference is present) to validate the oracle restatement.

Three shims:
  1. stub out the GUI/IO modules the reference imports at module top (matplotlib, ipywidgets,
     ctranslate2, librosa, mutagen, soundfile, PIL) -- none is touched by the hot path;
  2. `generate(inputs=...)` adapter: reference model.py:609/655 passes `inputs=`, which
     transformers>=4.36 Whisper rejects; map it to `input_features=`;
  3. an offline tokenizer with the real Whisper id layout for the ids the path uses.
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("WHISPERSEG_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_ref_modules = None


def import_reference():
    """Returns (model_module, audio_utils_module) of the unmodified reference."""
    global _ref_modules
    if _ref_modules is not None:
        return _ref_modules
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    # transformers must be fully imported BEFORE stubbing librosa/soundfile (its lazy
    # import machinery probes them with find_spec)
    from transformers import (WhisperFeatureExtractor, WhisperTokenizer,  # noqa: F401
                              WhisperForConditionalGeneration, WhisperConfig)  # noqa: F401
    import transformers.audio_utils  # noqa: F401
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib")
        for sub in ("pyplot", "colors", "patches", "cm"):
            sm = _stub("matplotlib." + sub)
            setattr(mpl, sub, sm)
        sys.modules["matplotlib.patches"].Patch = object
    _stub("ipywidgets", interact=lambda *a, **k: None, fixed=lambda x: x)
    _stub("ctranslate2")
    _stub("librosa")
    _stub("mutagen", File=None)
    _stub("soundfile")
    try:
        import PIL  # noqa: F401
    except Exception:
        pil = _stub("PIL")
        pil.Image = _stub("PIL.Image")
    try:
        import huggingface_hub  # noqa: F401
    except Exception:
        _stub("huggingface_hub", snapshot_download=None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference modules are called `model`, `audio_utils`, `utils` -- import under their
    # own names (they import each other by those names), then hand back the module objects
    saved = {k: sys.modules.get(k) for k in ("model", "audio_utils", "utils")}
    for k in saved:
        sys.modules.pop(k, None)
    import audio_utils as ref_audio_utils
    import model as ref_model
    for k, v in saved.items():
        if v is not None:
            sys.modules[k] = v
        else:
            sys.modules.pop(k, None)
    sys.modules["_whisperseg_ref_model"] = ref_model
    sys.modules["_whisperseg_ref_audio_utils"] = ref_audio_utils
    _ref_modules = (ref_model, ref_audio_utils)
    return _ref_modules


class GenerateAdapter:
    """Shim 2: reference calls hf.generate(inputs=...) (model.py:609)."""

    def __init__(self, hf_model):
        self.hf = hf_model
        self.config = hf_model.config

    def parameters(self):
        return self.hf.parameters()

    def generate(self, inputs=None, **kw):
        return self.hf.generate(input_features=inputs, **kw)
