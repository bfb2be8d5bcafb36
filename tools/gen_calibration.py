"""Write tools/calibration/<arch>_seed<k>.npy: the decoder output-bias calibration vector of the synthetic
checkpoint recipe (tools/synth.py:calibrate_output_bias_), so bench.py's GPU arm can build the checkpoint
without executing any oracle code.  Usage: python tools/gen_calibration.py large 0"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "large"
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
path = synth.calibration_path(arch, seed)
if os.path.isfile(path):
    os.unlink(path)
cfg, sd, gen = synth.make_state(arch, seed=seed, calibrate=False, eos_ramp=0.0)
d, L, H, F = synth.ARCHS[arch]
allowed = sorted(set(range(synth.VOCAB_SIZE)) - set(gen["suppress_tokens"]))
mean = synth.calibrate_output_bias_(sd, H, L, allowed, seed)
os.makedirs(os.path.dirname(path), exist_ok=True)
np.save(path, mean.numpy().astype(np.float32))
print("wrote", path, mean.shape, float(mean.abs().max()))
