"""GPU tuning aid: decode-length distribution / segment counts of the shaped large model as a function
of the EOS and digit embedding scales (rows of the tied embedding are rescaled in place on the GPU)."""
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth  # noqa: E402
from whisperseg_b200 import postprocess as pp  # noqa: E402
from whisperseg_b200.frontend import FrontendPlan  # noqa: E402
from whisperseg_b200.segmenter import WhisperSegmenter  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "large"
SR, STS = 48000, 0.0025
n_win = 240
state = synth.make_state(arch, seed=0)
tokdir = tempfile.mkdtemp()
synth.token_table_files(tokdir)
seg = WhisperSegmenter.from_state(state, tokdir, device="cuda", device_ids=[0], max_batch=n_win)
eng, tok = seg.engines[0], seg.tokenizer
audio = synth.synth_audio(600.0, SR, seed=2)
plan = FrontendPlan(SR, STS, 0)
wins = plan.windows(len(audio), 1)
feats = eng.features(plan, audio, wins)
emb = eng.tensors["dec.emb"]
pos = eng.tensors["dec.pos"]
base_pos = pos.clone()
u_eos = emb[synth.ID_EOT].float()
u_eos = u_eos / u_eos.norm()
ramp_pos = torch.clamp(torch.arange(pos.shape[0], device=pos.device, dtype=torch.float32) - 10.0, min=0.0)
inv = {v: k for k, v in seg.cluster_codebook.items()}
for ramp in [0.0, 0.02, 0.05, 0.1, 0.2, 0.4]:
    pos.copy_(base_pos + ramp * ramp_pos[:, None] * u_eos[None, :])
    eng.encode(feats)
    ids, n_steps = eng.generate(n_win, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, 448)
    ids = ids.cpu().numpy()
    lens = np.array([(row != tok.eos_token_id).sum() for row in ids])
    texts = tok.batch_decode(ids.tolist())
    nseg = sum(len(pp.segments_from_text(t, STS, inv)) for t in texts)
    uniq = len({tuple(r[:8]) for r in ids.tolist()})
    print("eos ramp %.2f: steps %3d  len mean %.1f median %.0f p95 %.0f max %d  zero-len %d  segments %d  distinct prefixes %d"
          % (ramp, n_steps, lens.mean(), np.median(lens), np.percentile(lens, 95), lens.max(), int((lens == 0).sum()), nseg, uniq),
          flush=True)
