"""Run the large model for a handful of decode positions (ncu target for the decode kernels)."""
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth  # noqa: E402
from whisperseg_b200.frontend import FrontendPlan  # noqa: E402
from whisperseg_b200.segmenter import WhisperSegmenter  # noqa: E402

n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 240
state = synth.make_state("large", seed=0, calibrate="file")
tokdir = tempfile.mkdtemp()
synth.token_table_files(tokdir)
seg = WhisperSegmenter.from_state(state, tokdir, device="cuda", device_ids=[0], max_batch=n_win)
eng, tok = seg.engines[0], seg.tokenizer
audio = synth.synth_audio(2.5 * n_win, 48000, seed=2)
plan = FrontendPlan(48000, 0.0025, 0)
feats = eng.features(plan, audio, plan.windows(len(audio), 1))
eng.encode(feats)
ids, n = eng.generate(n_win, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, 12, use_graph=False)
torch.cuda.synchronize()
print("done", n)
