"""GPU aid: where one decoder position of the persistent kernel (csrc/mega.cu) spends its time, phase by phase
(CTA 0's %globaltimer stamps: work of the phase, then the grid barrier)."""
import ctypes
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth  # noqa: E402
from whisperseg_b200.frontend import FrontendPlan  # noqa: E402
from whisperseg_b200.segmenter import WhisperSegmenter  # noqa: E402

os.environ["WSB_MEGA"] = "1"
arch = sys.argv[1] if len(sys.argv) > 1 else "large"
n_win = int(sys.argv[2]) if len(sys.argv) > 2 else 8
SR, STS = 48000, 0.0025
state = synth.make_state(arch, seed=0, calibrate="file" if arch == "large" else "auto")
tokdir = tempfile.mkdtemp()
synth.token_table_files(tokdir)
seg = WhisperSegmenter.from_state(state, tokdir, device="cuda", device_ids=[0], max_batch=n_win)
eng, tok = seg.engines[0], seg.tokenizer
audio = synth.synth_audio(n_win * 2.5, SR, seed=2)
plan = FrontendPlan(SR, STS, 0)
feats = eng.features(plan, audio, plan.windows(len(audio), 1))
cap = eng.lib.wsb_mega_trace(eng.handle, None, 1)
eng.encode(feats)
eng.generate(n_win, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, 5, use_graph=False)
buf = (ctypes.c_ulonglong * cap)()
eng.lib.wsb_mega_trace(eng.handle, buf, cap)          # clears the stage accumulators
eng.encode(feats)
n_pos = 4
eng.generate(n_win, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, 3 + n_pos - 2, use_graph=False)
n = eng.lib.wsb_mega_trace(eng.handle, buf, cap)
t = np.array(buf[:n], dtype=np.int64)
L = state[0]["decoder_layers"]
names = ["embed"]
for l in range(L):
    names += ["qkv", "self-attn", "so", "cq", "cross-attn", "co", "fc1", "fc2"]
work = (t[1:2 * len(names):2] - t[0:2 * len(names):2]) / 1000.0
wait = (t[2:2 * len(names) + 1:2] - t[1:2 * len(names):2]) / 1000.0
print("position total %.1f us over %d phases (rows %d)" % ((t[2 * len(names)] - t[0]) / 1000.0, len(names), n_win))
for kind in ["embed", "qkv", "self-attn", "so", "cq", "cross-attn", "co", "fc1", "fc2"]:
    idx = [i for i, nm in enumerate(names) if nm == kind]
    print("%-11s work %6.2f us  barrier wait %6.2f us   (CTA 0, mean over %d)" % (kind, work[idx].mean(), wait[idx].mean(), len(idx)))

stages = t[cap - 64:]
labels = ["A/stats loads issued+stats", "(unused)", "weight wait", "MMA", "red write + sync", "reduce + epilogue", "final sync"]
for cls, name in [(0, "folded-LN -> f32 (qkv, cq)"), (1, "folded-LN -> gelu (fc1)"), (4, "bf16 -> residual (so, co, fc2)")]:
    row = stages[8 * cls:8 * cls + 7] / 1000.0
    print("%-32s total %8.1f us over %d positions: %s" % (name, row.sum(), n_pos, ", ".join("%s %.1f" % (l, v) for l, v in zip(labels, row) if l != "(unused)")))
