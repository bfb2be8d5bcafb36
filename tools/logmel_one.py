"""Run the fused log-mel kernel a few times on the bench workload (target for `ncu --set full`)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisperseg_b200.engine import LogmelRunner  # noqa: E402
from whisperseg_b200.frontend import FrontendPlan  # noqa: E402

sr, sts, secs = (int(sys.argv[1]), float(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (48000, 0.0025, 600.0)
rng = np.random.default_rng(0)
audio = torch.from_numpy((rng.standard_normal(int(sr * secs)) * 0.1).astype(np.float32)).cuda()
plan = FrontendPlan(sr, sts, 0)
wins = plan.windows(audio.numel(), 1)
desc = torch.tensor([[w.start, 0, audio.numel()] for w in wins], dtype=torch.int64, device="cuda")
runner = LogmelRunner(torch.device("cuda", 0))
for _ in range(3):
    out = runner.run(plan, audio, desc, len(wins), torch.cuda.current_stream())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    runner.run(plan, audio, desc, len(wins), torch.cuda.current_stream())
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byt = plan.logmel_bytes_per_window() * len(wins)
print("%d windows n_fft %d hop %d: %.3f ms  %.1f GB/s (algorithmic %d B)" % (len(wins), plan.n_fft, plan.hop, ms, byt / ms / 1e6, byt))
