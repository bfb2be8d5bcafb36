import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 (B200) device")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def tiny_checkpoint(tmp_path_factory):
    """Seeded shaped whisper-tiny-architecture checkpoint directory (HF layout) + the HF model."""
    from tools import synth
    path = str(tmp_path_factory.mktemp("ckpt_tiny"))
    hf = synth.make_hf_model("tiny", seed=0, default_segmentation_config=dict(
        sr=16000, min_frequency=0, spec_time_step=0.01, species="human"))
    synth.save_checkpoint(hf, path)
    return path, hf


@pytest.fixture(scope="session")
def tiny_confident_checkpoint(tmp_path_factory):
    """The same architecture with the "confident" recipe (peaked logits, tools/synth.py: script_vectors): comparisons
    between two bf16 code paths are then about the kernels, not about which near-tie of a Gaussian logit vector flips."""
    from tools import synth
    path = str(tmp_path_factory.mktemp("ckpt_tiny_confident"))
    hf = synth.make_hf_model("tiny", seed=0, confident=True, default_segmentation_config=dict(
        sr=16000, min_frequency=0, spec_time_step=0.01, species="human"))
    synth.save_checkpoint(hf, path)
    return path, hf
