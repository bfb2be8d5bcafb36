"""TEST INFRASTRUCTURE -- writes tests/golden/tokens_<case>.npz: the fp32 oracle's greedy tokens and top-1/top-2
logit margins for whole parity workloads, so that the GPU box does not have to spend CPU minutes on a 1.5 B-parameter
fp32 network (>= 5000 teacher-forced positions, whisper-large on the bench workload).

    python oracle/gen_golden_tokens.py large_confident      # 240 windows of the bench audio, ~15 min on 8 cores
    python oracle/gen_golden_tokens.py large_stress32       # first 32 windows, bench (stress) recipe
    python oracle/gen_golden_tokens.py base_confident       # cfg1: 60 s @ 16 kHz

The generating inputs are all seeded (tools/synth.py); a test rebuilds the same checkpoint and audio, teacher-forces
the engine with these ids and compares its arg-max position by position."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frontend_np as FO  # noqa: E402
from oracle.whisper_torch import WhisperOracle  # noqa: E402
from tools import synth  # noqa: E402

CASES = {
    # name: (arch, make_state kwargs, seconds, sr, sts, audio seed, first n windows, max_length)
    "large_confident": ("large", dict(confident=True), 600.0, 48000, 0.0025, 2, 240, 64),
    "large_stress32": ("large", dict(), 80.0, 48000, 0.0025, 2, 32, 160),
    "base_confident": ("base", dict(confident=True), 60.0, 16000, 0.01, 1, 6, 64),
    # tiny: script boost 15 and audio seed 12 -- the default boost (14, seed 11) leaves 3 of the 1600 positions with an oracle
    # margin below 0.02, coin flips for any bf16 implementation and 0.19 % of a sample on which one flip is 0.06 %; here the
    # smallest margin is 0.128 (8 positions below 0.2) and the rows still leave the script where the audio says so
    "tiny_confident": ("tiny", dict(confident=True, script_boost=15.0), 640.0, 16000, 0.01, 12, 64, 64),
}


def case_inputs(name):
    arch, kw, seconds, sr, sts, seed, n_win, max_length = CASES[name]
    state = synth.make_state(arch, seed=0, **kw)
    audio = synth.synth_audio(seconds, sr, seed=seed)
    return arch, state, audio, sr, sts, n_win, max_length


def generate(name, chunk=16):
    arch, (cfg, sd, gen), audio, sr, sts, n_win, max_length = case_inputs(name)
    feats = FO.sliced_audio_features(audio, sr, 0, sts, 1, dtype=np.float32)[:n_win]
    x = torch.from_numpy(np.asarray([f[2] for f in feats]))
    orc = WhisperOracle(sd, cfg["encoder_attention_heads"], cfg["encoder_layers"])
    prompt = [synth.ID_SOT, synth.ID_EN, synth.ID_NOTIMESTAMPS]
    n_new = max_length - len(prompt)
    ids = np.full((n_win, n_new), synth.ID_EOT, dtype=np.int32)
    margins = np.zeros((n_win, n_new), dtype=np.float32)
    t0 = time.time()
    for c0 in range(0, n_win, chunk):
        enc = orc.encode(x[c0:c0 + chunk])
        i, m = orc.greedy(enc, prompt, synth.ID_EOT, synth.ID_EOT, max_length, suppress_tokens=gen["suppress_tokens"],
                          return_margins=True)
        ids[c0:c0 + i.shape[0], :i.shape[1]] = i.numpy()
        margins[c0:c0 + i.shape[0], :i.shape[1]] = m.numpy()
        print("%s: windows %d..%d done (%.0f s)" % (name, c0, c0 + i.shape[0], time.time() - t0), flush=True)
    out = os.path.join(ROOT, "tests", "golden", "tokens_%s.npz" % name)
    np.savez_compressed(out, ids=ids, margins=margins.astype(np.float16), n_windows=n_win, max_length=max_length)
    lens = np.array([int((r != synth.ID_EOT).sum()) for r in ids])
    print("wrote %s: %d windows, %d positions, row length mean %.1f max %d, distinct rows %d" %
          (out, n_win, int(lens.sum() + n_win), lens.mean(), lens.max(), len({tuple(r) for r in ids.tolist()})))


if __name__ == "__main__":
    for name in sys.argv[1:] or list(CASES):
        generate(name)
