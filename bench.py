#!/usr/bin/env python
"""bench.py -- audio-seconds segmented per second on the WhisperSeg hot path (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): whisper-large-architecture WhisperSeg, seeded shaped random-init
weights, 10 min of synthetic 48 kHz audio PER GPU (sts 0.0025 -> 240 windows of 120 000 samples per
GPU), num_trials=1, greedy decode with max_length 448.  One "step" = one pass of the hot path over
that batch: fused log-mel -> encoder -> cross-K/V -> greedy KV-cache decode -> (N>1) one all-gather of
the token lists.  N>1: one process per GPU (torchrun), windows sharded by index (weak scaling).

  value : whole-job audio-s/s with the audio already resident in HBM (device-timed, max over ranks)
  e2e   : the same through the public segmenter API with HOST audio: pinned H2D copy of the samples,
          kernels, D2H of the tokens, host post-processing (wall-clock around synchronised calls)
  roofline      : encoder GEMM class (tensor-bound), CUDA events inside the timed region
  kernels       : the other kernel classes (log-mel HBM GB/s, attention, decode streams ...)
  cpu_baseline  : the oracle port of the reference path timed on the host cores on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, STS, MIN_FREQ = 48000, 0.0025, 0
SECONDS_PER_GPU = 600.0
PROMPT_LEN = 3
CATS = ["conv1", "enc_gemm", "enc_attn", "enc_ln", "crosskv_gemm", "dec_gemm", "dec_logits", "dec_self_attn",
        "dec_cross_attn", "dec_ln", "misc", "dec_graph", "dec_compact"]
# `ncu --set full` summaries (tools/ncu_summary.py) of one representative launch of the roofline kernel class (the qkv
# projection: exactly the class-average FLOPs per launch), newest first: roofline.traffic is read from the first that exists
NCU_TRAFFIC_FILES = ["profiles/r2_ncu_full_gemm_qkv.txt", "profiles/r1b_ncu_full_gemm_qkv.txt", "profiles/r1_ncu_full_gemm_qkv2.txt"]


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the cited capture (MB in the summary files)."""
    for rel in NCU_TRAFFIC_FILES:
        path = os.path.join(ROOT, rel)
        if not os.path.isfile(path):
            continue
        vals = {}
        for line in open(path):
            f = line.split()
            if len(f) >= 2 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[0] not in vals:
                try:
                    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(f[2] if len(f) > 2 else "Mbyte", 1e6)
                    vals[f[0]] = float(f[1]) * scale
                except ValueError:
                    pass
        if len(vals) == 2:
            return sum(vals.values()), rel
    return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


def make_audio(seconds, sr, seed, copies=1):
    from tools import synth
    base = synth.synth_audio(seconds, sr, seed=seed)
    if copies == 1:
        return base
    return np.concatenate([(base * (0.6 + 0.4 * (i + 1) / copies)).astype(np.float32) for i in range(copies)])


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (recipe in B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_reference_sample(state, arch, max_length, n_windows=4, threads=None, **_ignored):
    """The reference path on the host cores, ONE WHOLE reference batch: segment()'s default batch_size=4 windows
    (reference model.py:407) through front-end -> fp32 Whisper encoder -> greedy decode until the longest of the four
    rows emits EOS (or max_length) -- every position is executed and timed, nothing is extrapolated inside the sample.
    kind "reference": the unmodified /root/reference WhisperSegmenterForEval (container only, oracle/ref_shim.py);
    kind "port": the oracle restatement (oracle/frontend_np.py + oracle/whisper_torch.py), which is what exists on the
    GPU box."""
    import torch
    from tools import synth
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg, sd, gen = state
    audio = make_audio(n_windows * 1000 * STS, SR, seed=2)
    audio_s = n_windows * 1000 * STS
    real = None
    if os.environ.get("WSB_BENCH_REAL_REFERENCE", "1") != "0":
        try:
            real = _real_reference_segmenter(state)
        except Exception:  # noqa: BLE001
            real = None
    if real is not None:
        t0 = time.perf_counter()
        real.segment(audio, SR, min_frequency=MIN_FREQ, spec_time_step=STS, num_trials=1, num_beams=1, batch_size=n_windows,
                     max_length=max_length)
        total = time.perf_counter() - t0
        return dict(value=audio_s / total, unit="audio-s/s", cores=threads, kind="reference", seconds=total,
                    sample="%d windows = one reference batch (%s arch) through the unmodified reference's "
                           "WhisperSegmenterForEval.segment (num_beams=1, max_length=%d): %.2f s" %
                           (n_windows, arch, max_length, total))
    from oracle import frontend_np as FO
    from oracle import postprocess_ref as PR
    from oracle.whisper_torch import WhisperOracle
    from whisperseg_b200.frontend import get_n_fft_given_sr
    orc = _oracle_cache.get(id(sd))
    if orc is None:
        orc = _oracle_cache[id(sd)] = WhisperOracle(sd, cfg["encoder_attention_heads"], cfg["encoder_layers"])
    t0 = time.perf_counter()
    feats = FO.sliced_audio_features(audio, SR, MIN_FREQ, STS, 1, dtype=np.float32)
    t_front = time.perf_counter() - t0
    x = torch.from_numpy(np.asarray([f[2] for f in feats]))
    t0 = time.perf_counter()
    enc = orc.encode(x)
    t_enc = time.perf_counter() - t0
    t0 = time.perf_counter()
    ids = orc.greedy(enc, [synth.ID_SOT, synth.ID_EN, synth.ID_NOTIMESTAMPS], synth.ID_EOT, synth.ID_EOT, max_length,
                     suppress_tokens=gen["suppress_tokens"])
    t_dec = time.perf_counter() - t0                     # cross-K/V + the prompt positions + every generated position
    t0 = time.perf_counter()
    tok = _token_table()
    texts = tok.batch_decode(ids.numpy())
    PR.segment_from_texts(texts, [(f[0], f[1], None, f[3]) for f in feats], len(audio), SR, STS, cfg["cluster_codebook"],
                          get_n_fft_given_sr(SR))
    t_post = time.perf_counter() - t0
    total = t_front + t_enc + t_dec + t_post
    return dict(value=audio_s / total, unit="audio-s/s", cores=threads, kind="port", seconds=total,
                sample="%d windows = one WHOLE reference batch (%s arch, oracle port, fp32, %d threads): front-end %.2f s + "
                       "encoder %.2f s + cross-K/V and %d greedy positions %.2f s (decoded until the longest row ended, "
                       "max_length=%d) + post-processing %.3f s; throughput = %.1f audio-s / %.2f s" %
                       (len(feats), arch, threads, t_front, t_enc, ids.shape[1] + PROMPT_LEN - 1, t_dec, max_length, t_post,
                        audio_s, total),
                frontend_s=t_front, encoder_s=t_enc, decode_s=t_dec, positions=int(ids.shape[1]))


_oracle_cache = {}
_tok_cache = []


def _token_table():
    if not _tok_cache:
        from tools import synth
        from whisperseg_b200.tokens import TokenTable
        d = tempfile.mkdtemp(prefix="wsb_tok_")
        synth.token_table_files(d)
        _tok_cache.append(TokenTable.from_pretrained(d))
    return _tok_cache[0]


_real_cache = {}


def _real_reference_segmenter(state):
    """The unmodified reference class over an HF model holding the same weights (only where /root/reference exists)."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        return None
    key = id(state[1])
    if key in _real_cache:
        return _real_cache[key]
    import torch
    from tools import synth
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    ref_model, _ = ref_shim.import_reference()
    cfg, sd, gen = state
    hf_cfg = WhisperConfig(**{k: v for k, v in cfg.items() if k not in ("model_type",)})
    with torch.device("meta"):
        hf = WhisperForConditionalGeneration(hf_cfg)
    hf = hf.to_empty(device="cpu")
    full = dict(sd)
    full["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    hf.load_state_dict({k: v.float() for k, v in full.items()}, strict=False)
    hf.tie_weights()
    hf.eval()
    hf.generation_config.suppress_tokens = gen["suppress_tokens"]
    hf.generation_config.begin_suppress_tokens = None
    hf.config.suppress_tokens = gen["suppress_tokens"]
    hf.config.begin_suppress_tokens = None
    seg = ref_model.WhisperSegmenterForEval(model=ref_shim.GenerateAdapter(hf), tokenizer=synth.build_tokenizer())
    _real_cache[key] = seg
    return seg


def run_reference(args):
    """--impl reference: the reference's CPU path on the box's host cores, all threads, one whole reference batch per
    step (see cpu_reference_sample); `value` = sample audio-seconds / measured step time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tools import synth
    state = synth.make_state(args.arch, seed=0)
    secs = []
    last = None
    for i in range(args.warmup + args.steps):
        r = cpu_reference_sample(state, args.arch, args.max_length, n_windows=args.ref_windows)
        if i >= args.warmup:
            secs.append(r["seconds"])
        last = r
    sample_audio_s = args.ref_windows * 1000 * STS
    step_s = float(np.mean(secs))
    v = sample_audio_s / step_s
    n_win = int(SECONDS_PER_GPU / (1000 * STS))
    cfg = workload_config(args, n_win)
    cfg["reference_step"] = "%d windows (%.1f audio-s) of the workload per step, measured in full" % (args.ref_windows, sample_audio_s)
    line = dict(metric="audio-sec/sec", value=v, unit="audio-s/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000.0 * step_s, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference", config=cfg,
                cpu_baseline=dict(value=v, unit="audio-s/s", cores=last["cores"], kind=last["kind"], sample=last["sample"]),
                e2e=dict(value=v, unit="audio-s/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


def workload_config(args, n_win):
    return dict(workload="configs[1]: whisper-%s architecture, %d s synthetic %d Hz audio per GPU, spec_time_step %g, "
                         "%d windows x %d samples per GPU, num_trials=1, greedy, max_length=%d" %
                         (args.arch, int(SECONDS_PER_GPU), SR, STS, n_win, int(1000 * STS * SR), args.max_length),
                arch=args.arch, windows_per_gpu=n_win, max_length=args.max_length, weights="seeded shaped random-init",
                l2="inputs larger than L2 (weights 3 GB, activations > 1 GB per pass): no flush needed")


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tools import synth
    from whisperseg_b200 import _lib
    from whisperseg_b200.distributed import all_gather_tokens, segment_sharded
    from whisperseg_b200.frontend import FrontendPlan
    from whisperseg_b200.segmenter import WhisperSegmenter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    pk = peaks()

    n_win = int(SECONDS_PER_GPU / (1000 * STS))
    state = synth.make_state(args.arch, seed=0, calibrate="file")   # committed calibration vector: no oracle code on this arm
    tokdir = tempfile.mkdtemp(prefix="wsb_tok_")
    synth.token_table_files(tokdir)
    seg = WhisperSegmenter.from_state(state, tokdir, device="cuda", device_ids=[local_rank], max_batch=n_win)
    eng, tok = seg.engines[0], seg.tokenizer
    plan = FrontendPlan(SR, STS, MIN_FREQ)
    max_new = args.max_length - PROMPT_LEN

    # the logical recording is world x 600 s; every rank owns the windows of its own 600 s piece
    base = make_audio(SECONDS_PER_GPU, SR, seed=2)
    gain = 0.6 + 0.4 * (rank + 1) / world
    piece = (base * gain).astype(np.float32) if world > 1 else base
    wins = plan.windows(len(piece), 1)
    assert len(wins) == n_win
    audio_dev = eng.upload_audio(piece)
    desc_dev = eng.window_descriptors(wins, 0, len(piece))
    torch.cuda.synchronize()

    def device_step():
        feats = eng.features_device(plan, audio_dev, desc_dev, n_win)
        eng.encode(feats)
        ids, n_steps = eng.generate(n_win, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, args.max_length)
        if world > 1:
            ids = all_gather_tokens(ids, n_win, max_new, tok.pad_token_id)
        return ids, n_steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib.wsb_profile_enable(0)
    for _ in range(args.warmup):
        ids, n_steps = device_step()
    barrier()
    lib.wsb_launch_count(1)
    lib.wsb_profile_enable(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.lean:
        torch.cuda.profiler.start()                      # ncu --profile-from-start off: only the timed steps are captured
    ev0.record()
    steps_done = []
    for _ in range(args.steps):
        ids, n_steps = device_step()
        steps_done.append(n_steps)
    ev1.record()
    barrier()
    if args.lean:
        torch.cuda.profiler.stop()
    clocks = sampler.stop()
    launches = int(lib.wsb_launch_count(0))
    dt_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([dt_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt_ms = float(t.item())
    prof = {}
    import ctypes
    for i, name in enumerate(CATS):
        ms, cnt, work = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
        lib.wsb_profile_read(i, ctypes.byref(ms), ctypes.byref(cnt), ctypes.byref(work))
        prof[name] = (ms.value, cnt.value, work.value)
    lib.wsb_profile_enable(0)
    value = world * SECONDS_PER_GPU * args.steps / (dt_ms / 1000.0)
    if args.lean:                                        # profiling runs (ncu launch lists): the timed steps and nothing else
        if rank == 0:
            emit(dict(metric="audio-sec/sec", value=value, unit="audio-s/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                      ms_per_step=dt_ms / args.steps, gpu_launches=launches, note="--lean run (under a profiler this is not a bench value)"))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- kernel-class numbers ------------------------------------------------------------------
    # log-mel: timed alone with CUDA events on its stream (burst HBM peak applies)
    lm = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.stream.wait_stream(torch.cuda.current_stream())
        a.record(eng.stream)
        eng.logmel.run(plan, audio_dev, desc_dev, n_win, eng.stream)
        b.record(eng.stream)
        eng.stream.synchronize()
        lm.append(a.elapsed_time(b))
    lm_ms = float(np.median(lm[1:]))
    lm_bytes = plan.logmel_bytes_per_window() * n_win
    kernels = {"logmel": dict(bound="hbm", achieved=lm_bytes / lm_ms / 1e6, peak=pk["hbm"], unit="GB/s",
                              frac=lm_bytes / lm_ms / 1e6 / pk["hbm"], ms_per_launch=lm_ms, launches=1,
                              algorithmic_bytes_per_launch=lm_bytes,
                              note="compute-co-limited: %.0f FLOP/B with a radix FFT (SURVEY 7.2-5)" %
                                   (plan.logmel_flops_per_window() / plan.logmel_bytes_per_window()))}
    # the same kernel on the 16 kHz / hop-160 / n_fft-512 front-end (cfg1 / cfg4: 14 FLOP/B, the plausibly bandwidth-bound case)
    try:
        plan16 = FrontendPlan(16000, 0.01, MIN_FREQ)
        a16 = torch.from_numpy(np.tile(make_audio(600.0, 16000, seed=4), 6)).to(dev)      # 1 h, the cfg4 front-end workload
        w16 = plan16.windows(a16.numel(), 1)
        d16 = eng.window_descriptors(w16, 0, a16.numel())
        t16 = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(eng.stream)
            eng.logmel.run(plan16, a16, d16, len(w16), eng.stream)
            b.record(eng.stream)
            eng.stream.synchronize()
            t16.append(a.elapsed_time(b))
        ms16 = float(np.median(t16[1:]))
        by16 = plan16.logmel_bytes_per_window() * len(w16)
        kernels["logmel_16k_hop160"] = dict(bound="hbm", achieved=by16 / ms16 / 1e6, peak=pk["hbm"], unit="GB/s",
                                            frac=by16 / ms16 / 1e6 / pk["hbm"], ms_per_launch=ms16, launches=1,
                                            algorithmic_bytes_per_launch=by16, note="%d windows of 160 000 samples" % len(w16))
        del a16, d16
    except Exception as e:  # noqa: BLE001
        kernels["logmel_16k_hop160"] = dict(error=str(e))
    # decode kernel classes: eager (graph-free) teacher-forced pass so every row stays active
    lib.wsb_profile_enable(1)
    forced = torch.full((n_win, args.max_length), tok.eos_token_id, dtype=torch.int32, device=dev)
    forced[:, :PROMPT_LEN] = torch.tensor(tok.prompt_ids, dtype=torch.int32, device=dev)
    forced[:, PROMPT_LEN:] = ids[rank * n_win:(rank + 1) * n_win] if world > 1 else ids
    short = min(args.max_length, PROMPT_LEN + 6)
    eng.generate(n_win, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, short,
                 forced=forced[:, :short].contiguous(), use_graph=False)
    torch.cuda.synchronize()
    dprof = {}
    for i, name in enumerate(CATS):
        ms, cnt, work = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
        lib.wsb_profile_read(i, ctypes.byref(ms), ctypes.byref(cnt), ctypes.byref(work))
        dprof[name] = (ms.value, cnt.value, work.value)
    lib.wsb_profile_enable(0)

    def tensor_entry(p, peak):
        ms, cnt, work = p
        if cnt == 0 or ms <= 0:
            return None
        ach = work / ms / 1e9
        return dict(bound="tensor", achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak, ms_per_launch=ms / cnt,
                    launches=cnt, algorithmic_flops_per_launch=work / cnt)

    def hbm_entry(p, peak):
        ms, cnt, work = p
        if cnt == 0 or ms <= 0 or work <= 0:
            return None
        ach = work / ms / 1e6
        return dict(bound="hbm", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, ms_per_launch=ms / cnt,
                    launches=cnt, algorithmic_bytes_per_launch=work / cnt)

    roof = tensor_entry(prof["enc_gemm"], pk["bf16_sustained"])
    kernels["enc_attn"] = tensor_entry(prof["enc_attn"], pk["bf16_sustained"])
    kernels["crosskv_gemm"] = tensor_entry(prof["crosskv_gemm"], pk["bf16_sustained"])
    kernels["conv1"] = tensor_entry(prof["conv1"], pk["bf16_sustained"])
    kernels["enc_ln"] = hbm_entry(prof["enc_ln"], pk["hbm"])
    kernels["dec_cross_attn"] = hbm_entry(dprof["dec_cross_attn"], pk["hbm"])
    d, L, H, F = synth.ARCHS[args.arch]
    dec_weight_bytes = 2.0 * (L * (3 * d * d + d * d + d * d + d * d + 2 * d * F))
    dg_ms, dg_cnt, _ = dprof["dec_gemm"]
    if dg_cnt:
        n_pos = dg_cnt / (6 * L)
        kernels["dec_gemm"] = dict(bound="hbm", achieved=dec_weight_bytes * n_pos / dg_ms / 1e6, peak=pk["hbm"], unit="GB/s",
                                   frac=dec_weight_bytes * n_pos / dg_ms / 1e6 / pk["hbm"], ms_per_launch=dg_ms / dg_cnt,
                                   launches=dg_cnt, algorithmic_bytes_per_launch=dec_weight_bytes / (6 * L),
                                   note="weight streaming at batch %d" % n_win)
    shares = {k: v[0] for k, v in prof.items() if v[1]}
    shares["logmel"] = lm_ms * args.steps
    # DRAM traffic of one representative launch of the class (the qkv projection: exactly the class-average FLOPs per
    # launch; algorithmic: A 307 MB + W 10 MB + C 922 MB) read from the committed `ncu --set full` summary
    traffic, traffic_src = ncu_traffic_bytes() if args.arch == "large" else (None, None)
    roofline = dict(bound="tensor", achieved=roof["achieved"], peak=roof["peak"], unit="TFLOP/s", frac=roof["frac"],
                    traffic=traffic, traffic_source=traffic_src, kernel="gemm_kernel<BN> (encoder GEMM class: conv2, qkv, out-proj, fc1, fc2)",
                    ms_per_launch=roof["ms_per_launch"], launches=roof["launches"],
                    algorithmic_flops_per_launch=roof["algorithmic_flops_per_launch"],
                    peak_source="%s cuBLAS bf16, sustained figure (kernel timed inside a long step)" % pk["source"])

    # ---- end to end through the public API, host audio ---------------------------------------
    full_audio = None
    if world > 1:
        full_audio = np.concatenate([(base * (0.6 + 0.4 * (r + 1) / world)).astype(np.float32) for r in range(world)])
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        if world > 1:
            return segment_sharded(seg, full_audio, SR, MIN_FREQ, STS, max_length=args.max_length, num_trials=1, num_beams=1)
        return seg.segment(piece, SR, MIN_FREQ, STS, max_length=args.max_length, num_trials=1, num_beams=1)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = e2e_step()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    e2e_value = world * SECONDS_PER_GPU * e2e_steps / t_e2e

    # ---- secondary figure: worst case for the decoder -- the same checkpoint WITHOUT the positional EOS ramp
    # (tools/synth.py:eos_ramp_vectors): ~5 % of the rows then never emit EOS and the batch runs to
    # max_length.  Device-resident, like `value`; the ramp is removed and restored in place in HBM.
    ramp_dev = None
    if getattr(synth, "ARCH_EOS_RAMP", {}).get(args.arch, 0.0):
        ramp_dev = synth.eos_ramp_vectors(state[1]["model.decoder.embed_tokens.weight"],
                                          synth.ARCH_EOS_RAMP[args.arch]).to(dev)
    worst_ms, worst_positions = None, None
    if ramp_dev is not None:
        pos_t = eng.tensors["dec.pos"]
        pos_t.sub_(ramp_dev[:pos_t.shape[0]].to(pos_t.dtype))
        device_step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        _, worst_positions = device_step()
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        worst_ms = float(ts.item())
        pos_t.add_(ramp_dev[:pos_t.shape[0]].to(pos_t.dtype))
    # ---- the reference's DEFAULT decode mode: beam search, num_beams=4 (model.py:409) -- same workload, windows
    # decoded in chunks of max_batch/4 so that windows x beams fills the same 240 decode rows
    beam_ms, beam_positions = None, 0
    if not args.no_beam and n_win >= 4:
        per = n_win // 4

        def beam_pass():
            feats = eng.features_device(plan, audio_dev, desc_dev, n_win)
            total = 0
            for pos in range(0, n_win, per):
                chunk = feats[pos:pos + per].contiguous()
                eng.encode(chunk)
                _, st = eng.generate_beam(chunk.shape[0], 4, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id,
                                          args.max_length, 1.0)
                total += st
            return total
        beam_pass()
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        beam_positions = beam_pass()
        b1.record()
        barrier()
        tb = torch.tensor([b0.elapsed_time(b1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        beam_ms = float(tb.item())
    h2d = len(piece) * 4 + n_win * 24
    d2h = n_win * max_new * 4
    # ---- the decode phase as a whole (graph replays are bracketed as one class): algorithmic HBM bytes = the decoder
    # weights once per position + the cross-attention K/V blocks of every LIVE row at every position it runs
    ids_rank = (ids[rank * n_win:(rank + 1) * n_win] if world > 1 else ids).cpu().numpy()
    live_positions = float(((ids_rank != tok.eos_token_id).sum(axis=1) + 1 + (PROMPT_LEN - 1)).clip(max=max_new + PROMPT_LEN - 1).sum())
    cross_bytes_per_row = 2.0 * L * 2 * H * eng.T * 64
    n_positions = float(np.mean(steps_done)) + PROMPT_LEN - 1
    dec_bytes = dec_weight_bytes * n_positions + cross_bytes_per_row * live_positions
    dec_ms = sum(prof[k][0] for k in ("dec_gemm", "dec_logits", "dec_self_attn", "dec_cross_attn", "dec_ln", "dec_graph", "dec_compact")) / args.steps
    if dec_ms > 0:
        kernels["decode_phase"] = dict(bound="hbm", achieved=dec_bytes / dec_ms / 1e6, peak=pk["hbm"], unit="GB/s",
                                       frac=dec_bytes / dec_ms / 1e6 / pk["hbm"], ms_per_step=dec_ms, positions=n_positions,
                                       live_row_positions=live_positions, algorithmic_bytes_per_step=dec_bytes,
                                       note="all decoder positions of a step (eager prompt positions + CUDA-graph replays + "
                                            "compaction); bytes = decoder weights x positions + cross-K/V x live row-positions")

    ids_host = (ids[rank * n_win:(rank + 1) * n_win] if world > 1 else ids).cpu().numpy()
    row_len = (ids_host != tok.eos_token_id).sum(axis=1)
    per_batch4 = float(np.mean([min(max_new, r.max() + 1) for r in row_len[:len(row_len) // 4 * 4].reshape(-1, 4)]))
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference_sample(state, args.arch, args.max_length, n_windows=args.ref_windows,
                                       )
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = dict(metric="audio-sec/sec", value=value, unit="audio-s/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=dt_ms / args.steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="bf16", data="synthetic", config=workload_config(args, n_win),
                    e2e=dict(value=e2e_value, unit="audio-s/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             steps=e2e_steps, segments=len(res["onset"])),
                    gpu_launches=launches, clocks=clocks, roofline=roofline, kernels=kernels, cpu_baseline=cpu,
                    secondary=None if worst_ms is None else dict(
                        note="worst case: same workload and checkpoint without the positional EOS ramp -- about 5 % of "
                             "the rows never emit EOS and the batch decodes to max_length",
                        ms_per_step=worst_ms, decode_positions_per_step=int(worst_positions),
                        value=world * SECONDS_PER_GPU / (worst_ms / 1000.0), unit="audio-s/s"),
                    beam4=None if beam_ms is None else dict(
                        note="same workload decoded with the reference's default num_beams=4 (HF beam search on the "
                             "device, length_penalty 1.0), 4 calls of %d windows x 4 beams" % (n_win // 4),
                        ms_per_step=beam_ms, decode_positions_per_step=int(beam_positions),
                        value=world * SECONDS_PER_GPU / (beam_ms / 1000.0), unit="audio-s/s"),
                    decode_positions_per_step=float(np.mean(steps_done)),
                    decode_row_lengths=dict(mean=float(row_len.mean()), median=float(np.median(row_len)),
                                            p95=float(np.percentile(row_len, 95)), max=int(row_len.max()),
                                            mean_positions_per_batch_of_4=per_batch4),
                    step_share_ms={k: v / args.steps for k, v in shares.items()})
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------
# The other BASELINE.json configurations, end to end through the public API with HOST audio (python bench.py
# --config cfg3|cfg4|cfg5 [--gpus N]); the default (cfg2) above is what the driver runs.
CONFIGS = {
    # name: (description, sr, sts, seconds, seed, num_trials, segment kwargs)
    "cfg3": ("configs[2]: zebra-finch parameters, 600 s synthetic 32 kHz audio, spec_time_step 0.0025, num_trials=3 "
             "(multi-trial consolidation by clustering included)", 32000, 0.0025, 600.0, 3, 3, dict(eps=0.02, min_segment_length=0.01)),
    "cfg4": ("configs[3]: 1 h synthetic 16 kHz human-VAD-like audio, spec_time_step 0.01, windows sharded across the ranks, "
             "one all-gather of segment tables", 16000, 0.01, 3600.0, 4, 1, dict(min_segment_length=0.1)),
    "cfg5": ("configs[4]: folder mode, variable-length clips U(0.5, 30) s at 32 kHz, spec_time_step 0.0025, windows of all "
             "clips flattened into shared batches and sharded across the ranks", 32000, 0.0025, 600.0, 5, 1, dict()),
}


def run_config(args):
    import torch
    import torch.distributed as dist
    from tools import synth
    from whisperseg_b200 import _lib
    from whisperseg_b200.distributed import LazyClips, segment_many_sharded, segment_sharded
    from whisperseg_b200.frontend import FrontendPlan
    from whisperseg_b200.segmenter import WhisperSegmenter
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    state = synth.make_state(args.arch, seed=0, calibrate="file" if args.arch == "large" else "auto")
    tokdir = tempfile.mkdtemp(prefix="wsb_tok_")
    synth.token_table_files(tokdir)
    seg = WhisperSegmenter.from_state(state, tokdir, device="cuda", device_ids=[local_rank], max_batch=args.max_batch)
    for name in args.config.split(","):                    # several configurations on one model build (one JSON line each)
        _run_one_config(args, name, seg, lib, world, rank, local_rank, dev)
    if world > 1:
        dist.destroy_process_group()


def _run_one_config(args, config, seg, lib, world, rank, local_rank, dev):
    import torch
    import torch.distributed as dist
    from whisperseg_b200.distributed import LazyClips, segment_many_sharded, segment_sharded
    from whisperseg_b200.frontend import FrontendPlan
    desc, sr, sts, seconds, seed, num_trials, kw = CONFIGS[config]
    base = make_audio(seconds, sr, seed=seed)
    plan = FrontendPlan(sr, sts, MIN_FREQ)
    common = dict(min_frequency=MIN_FREQ, spec_time_step=sts, max_length=args.max_length, num_trials=num_trials, num_beams=1, **kw)
    if config == "cfg5":
        rng = np.random.default_rng(seed)
        durations = rng.uniform(0.5, 30.0, size=args.clips)
        lengths = (durations * sr).astype(np.int64)
        offsets = rng.integers(0, len(base) - lengths.max() - 1, size=args.clips)
        load = lambda k: np.ascontiguousarray(base[offsets[k]:offsets[k] + lengths[k]])      # noqa: E731
        audio_seconds = float(lengths.sum()) / sr
        n_windows = int(sum(len(plan.windows(int(n), num_trials)) for n in lengths))

        def step():
            if world > 1:
                return segment_many_sharded(seg, LazyClips(lengths, load), sr, **common)
            return seg.segment_many([load(k) for k in range(args.clips)], sr, **common)
    else:
        audio_seconds = seconds
        n_windows = len(plan.windows(len(base), num_trials))

        def step():
            if world > 1:
                return segment_sharded(seg, base, sr, **common)
            return seg.segment(base, sr, **common)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        res = step()
    barrier()
    lib.wsb_launch_count(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step()
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    n_seg = sum(len(r["onset"]) for r in res) if isinstance(res, list) else len(res["onset"])
    if rank == 0:
        value = audio_seconds * args.steps / dt
        emit(dict(metric="audio-sec/sec", value=value, unit="audio-s/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                  ms_per_step=1000.0 * dt / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="bf16",
                  data="synthetic", config=dict(workload=desc, config=config, arch=args.arch, windows=n_windows,
                                                audio_seconds=audio_seconds, max_length=args.max_length, max_batch=args.max_batch,
                                                clips=args.clips if config == "cfg5" else None,
                                                weights="seeded shaped random-init (stress recipe)",
                                                timing="wall clock around synchronised public-API calls, host audio in, segments out"),
                  e2e=dict(value=value, unit="audio-s/s", h2d_bytes_per_step=int(audio_seconds * sr * 4 / world),
                           d2h_bytes_per_step=int(n_windows * (args.max_length - PROMPT_LEN) * 4 / world), segments=n_seg),
                  gpu_launches=int(lib.wsb_launch_count(0)), clocks=clocks))


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else a library prints (NCCL's version banner,
    torchrun notices) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arch", default="large")
    ap.add_argument("--max-length", type=int, default=448, dest="max_length")
    ap.add_argument("--ref-windows", type=int, default=4, dest="ref_windows")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-beam", action="store_true", dest="no_beam")
    ap.add_argument("--lean", action="store_true", help="only the warm-up and timed steps (for ncu launch lists)")
    ap.add_argument("--config", default="cfg2", help="cfg2 (default, the driver's workload) or a comma-separated list of cfg3, cfg4, cfg5")
    ap.add_argument("--clips", type=int, default=10000, help="cfg5: number of clips in the folder")
    ap.add_argument("--max-batch", type=int, default=240, dest="max_batch")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config != "cfg2":
        run_config(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
