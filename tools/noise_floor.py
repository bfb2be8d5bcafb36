"""CPU tuning aid (dev tool, not product): how often does a bf16-operand implementation flip the greedy arg-max of a
synthetic checkpoint against fp32?  Runs the fp32 oracle (greedy, free-running) and the engine-rounding model
(oracle/bf16_emul.py, teacher-forced on the oracle's tokens) on the same windows and prints agreement, margins,
row-length and diversity statistics.

    python tools/noise_floor.py --arch tiny --windows 64 [--kw key=value ...]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import frontend_np as FO  # noqa: E402
from oracle.bf16_emul import Bf16EngineModel  # noqa: E402
from oracle.whisper_torch import WhisperOracle  # noqa: E402
from tools import synth  # noqa: E402


def measure(arch, n_windows, sr, sts, seed_audio, max_length, chunk=16, verbose=True, **kw):
    cfg, sd, gen = synth.make_state(arch, seed=0, **kw)
    H, L = cfg["encoder_attention_heads"], cfg["encoder_layers"]
    audio = synth.synth_audio(n_windows * 1000 * sts, sr, seed=seed_audio)
    feats = FO.sliced_audio_features(audio, sr, 0, sts, 1, dtype=np.float32)[:n_windows]
    x = torch.from_numpy(np.asarray([f[2] for f in feats]))
    orc = WhisperOracle(sd, H, L)
    emu = Bf16EngineModel(sd, H, L)
    prompt = [synth.ID_SOT, synth.ID_EN, synth.ID_NOTIMESTAMPS]
    sup = torch.tensor(gen["suppress_tokens"], dtype=torch.long)
    tot = agree = 0
    lens, seqs, all_margins, mism_margins = [], [], [], []
    t0 = time.time()
    for c0 in range(0, n_windows, chunk):
        xc = x[c0:c0 + chunk]
        enc = orc.encode(xc)
        ids, margins = orc.greedy(enc, prompt, synth.ID_EOT, synth.ID_EOT, max_length, suppress_tokens=gen["suppress_tokens"],
                                  return_margins=True)
        B, n_new = ids.shape
        full = torch.cat([torch.tensor([prompt] * B), ids], dim=1)
        enc_e = emu.encode(xc)
        lg = emu.decode_logits(full[:, :-1], enc=enc_e)[:, len(prompt) - 1:, :]
        lg[:, :, sup] = float("-inf")
        got = lg.argmax(dim=-1)
        valid = torch.ones_like(ids, dtype=torch.bool)
        for b in range(B):
            eos = (ids[b] == synth.ID_EOT).nonzero()
            if len(eos):
                valid[b, eos[0, 0] + 1:] = False
            lens.append(int(valid[b].sum()))
            seqs.append(tuple(ids[b][valid[b]].tolist()))
        tot += int(valid.sum())
        agree += int(((got == ids) & valid).sum())
        all_margins += margins[valid].tolist()
        mism_margins += margins[valid & (got != ids)].tolist()
        if verbose:
            print("  chunk %d: %d/%d agree so far (%.1fs)" % (c0 // chunk, agree, tot, time.time() - t0), flush=True)
    lens = np.array(lens)
    am = np.array(all_margins)
    res = dict(raw=agree / max(tot, 1), positions=tot, flips=tot - agree, len_mean=float(lens.mean()), len_median=float(np.median(lens)),
               len_p95=float(np.percentile(lens, 95)), len_max=int(lens.max()), distinct=len(set(seqs)), windows=n_windows,
               margin_p01=float(np.percentile(am, 1)), margin_p10=float(np.percentile(am, 10)), margin_median=float(np.median(am)),
               mismatch_margin_max=float(max(mism_margins)) if mism_margins else 0.0)
    return res, seqs


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="tiny")
    ap.add_argument("--windows", type=int, default=32)
    ap.add_argument("--sr", type=int, default=48000)
    ap.add_argument("--sts", type=float, default=0.0025)
    ap.add_argument("--seed-audio", type=int, default=2)
    ap.add_argument("--max-length", type=int, default=160)
    ap.add_argument("--chunk", type=int, default=16)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--show", type=int, default=4)
    ap.add_argument("--kw", nargs="*", default=[])
    a = ap.parse_args()
    if a.threads:
        torch.set_num_threads(a.threads)
    kw = {}
    for item in a.kw:
        k, v = item.split("=")
        kw[k] = (v == "True") if v in ("True", "False") else (float(v) if "." in v or "e" in v else int(v))
    res, seqs = measure(a.arch, a.windows, a.sr, a.sts, a.seed_audio, a.max_length, a.chunk, **kw)
    print(kw)
    print(res)
    tok = {synth.ID_EOT: "E"}
    for s in seqs[:a.show]:
        print(" ".join(tok.get(t, str(t - synth.ID_TS0) if t >= synth.ID_TS0 else "d%d" % (t - synth.ID_DIGIT0)) for t in s))
