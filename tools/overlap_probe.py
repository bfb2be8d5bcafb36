"""GPU probe (dev tool): can the latency-bound greedy decode of one chunk of windows run CONCURRENTLY with the
tensor-bound encoder of the next chunk?  Two engines (own workspaces, own streams), two host threads:
  alone      : encoder pass of chunk B; decode of chunk A
  concurrent : both at once, the encoder's persistent GEMMs leaving R SMs free (wsb_set_sm_reserve)
Prints milliseconds for each so that the chunk-pipelined schedule in segmenter.py can be sized.

    python tools/overlap_probe.py [windows_per_chunk=120] [reserves=0,16,32,48]
"""
import os
import sys
import tempfile
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth  # noqa: E402
from whisperseg_b200 import _lib  # noqa: E402
from whisperseg_b200.frontend import FrontendPlan  # noqa: E402
from whisperseg_b200.segmenter import WhisperSegmenter  # noqa: E402

n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 120
reserves = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,16,32,48").split(",")]
SR, STS = 48000, 0.0025
lib = _lib.load()
state = synth.make_state("large", seed=0, calibrate="file")
tokdir = tempfile.mkdtemp()
synth.token_table_files(tokdir)
from whisperseg_b200.engine import Engine  # noqa: E402
segA = WhisperSegmenter.from_state(state, tokdir, device="cuda", device_ids=[0], max_batch=n_win)
engB, tok = segA.engines[0], segA.tokenizer
prio = int(os.environ.get("PROBE_PRIO", "-1"))
engA = Engine(None, engB.device, max_batch=n_win, state=state, tensors=engB.tensors, stream_priority=prio)   # decoder: high priority
plan = FrontendPlan(SR, STS, 0)
audio = synth.synth_audio(2 * n_win * 2.5, SR, seed=2)
wins = plan.windows(len(audio), 1)
feats = engA.features(plan, audio, wins)
fA, fB = feats[:n_win].contiguous(), feats[n_win:2 * n_win].contiguous()


# torch's current stream is thread-local; Engine brackets its work with wait_stream() against it, so each worker
# thread needs a current stream of its own or the two engines would serialise through the default stream
def encode_pass(reps, reserve, out):
    with torch.cuda.stream(sideB):
        lib.wsb_set_sm_reserve(reserve)
        t0 = time.perf_counter()
        for _ in range(reps):
            engB.encode(fB)
        engB.stream.synchronize()
        out["enc_ms"] = (time.perf_counter() - t0) * 1000 / reps
        lib.wsb_set_sm_reserve(0)


def decode_pass(out):
    with torch.cuda.stream(sideA):
        t0 = time.perf_counter()
        ids, n_steps = engA.generate(n_win, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, 448)
        engA.stream.synchronize()
        out["dec_ms"] = (time.perf_counter() - t0) * 1000
        out["steps"] = n_steps
        out["ids"] = ids


sideA, sideB = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()


# warm up (graphs, function attributes) sequentially
engA.encode(fA)
decode_pass({})
encode_pass(1, 0, {})
torch.cuda.synchronize()

r = {}
encode_pass(3, 0, r)
print("encoder alone (%d windows): %.1f ms" % (n_win, r["enc_ms"]), flush=True)
for R in reserves[1:]:
    encode_pass(3, R, r)
    print("encoder alone, %d SMs reserved: %.1f ms" % (R, r["enc_ms"]), flush=True)
engA.encode(fA)
torch.cuda.synchronize()
base = {}
decode_pass(base)
print("decode alone: %.1f ms, %d positions" % (base["dec_ms"], base["steps"]), flush=True)
for R in reserves:
    engA.encode(fA)
    torch.cuda.synchronize()
    ro, rd = {}, {}
    reps = max(1, int(round(base["dec_ms"] / r["enc_ms"])))
    te = threading.Thread(target=encode_pass, args=(reps, R, ro))
    td = threading.Thread(target=decode_pass, args=(rd,))
    t0 = time.perf_counter()
    td.start()
    te.start()
    td.join()
    te.join()
    wall = (time.perf_counter() - t0) * 1000
    same = bool(torch.equal(rd["ids"], base["ids"]))
    print("concurrent, reserve %3d SMs: decode %.1f ms | %d encoder passes at %.1f ms each | wall %.1f ms | tokens identical %s"
          % (R, rd["dec_ms"], reps, ro["enc_ms"], wall, same), flush=True)
