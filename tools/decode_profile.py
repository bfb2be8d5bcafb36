"""GPU aid: where the decode time of the bench workload goes, by position.  Times generate() with growing
max_length (cumulative), so the differences give the cost per position in every batch regime."""
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth  # noqa: E402
from whisperseg_b200.frontend import FrontendPlan  # noqa: E402
from whisperseg_b200.segmenter import WhisperSegmenter  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "large"
SR, STS, n_win = 48000, 0.0025, 240
state = synth.make_state(arch, seed=0, calibrate="file")
tokdir = tempfile.mkdtemp()
synth.token_table_files(tokdir)
seg = WhisperSegmenter.from_state(state, tokdir, device="cuda", device_ids=[0], max_batch=n_win)
eng, tok = seg.engines[0], seg.tokenizer
audio = synth.synth_audio(600.0, SR, seed=2)
plan = FrontendPlan(SR, STS, 0)
feats = eng.features(plan, audio, plan.windows(len(audio), 1))
eng.encode(feats)
ids, n_steps = eng.generate(n_win, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, 448)
lens = (ids.cpu().numpy() != tok.eos_token_id).sum(axis=1)
prev_t, prev_l = None, None
for L in [4, 8, 12, 16, 24, 32, 40, 48, 64, 80, 100, 120, 140]:
    best = 1e9
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(eng.stream)
        ids, st = eng.generate(n_win, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, L + 3)
        b.record(eng.stream)
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    live = int((lens + 1 > L).sum())
    extra = "" if prev_t is None else "  -> %.2f ms/position over (%d, %d]" % ((best - prev_t) / max(1, st - prev_l), prev_l, st)
    print("max_new %3d: %7.1f ms, %3d positions, rows still live after them %3d%s" % (L, best, st, live, extra), flush=True)
    prev_t, prev_l = best, st
