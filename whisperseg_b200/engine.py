"""Thin Python driver over the C ABI: owns device tensors (torch is used for memory and streams
only) and calls libwsb for every piece of arithmetic."""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from .frontend import N_MELS, FrontendPlan
from .weights import load_checkpoint, prepare_tensors


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class LogmelRunner:
    """K1 plans are cached per (sr, spec_time_step, min_frequency) -- the reference rebuilds its
    feature extractor on every segment() call (model.py:128)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.lib = _lib.load()
        self._plans = {}

    def plan_for(self, fp: FrontendPlan):
        key = (fp.sr, fp.hop, fp.n_fft, fp.clip_len, fp.total_spec_columns, float(fp.min_frequency), float(fp.max_frequency))
        h = self._plans.get(key)
        if h is None:
            with torch.cuda.device(self.device):
                filt = np.ascontiguousarray(fp.mel_filters.astype(np.float32))
                handle = ctypes.c_void_p()
                _lib.check(self.lib.wsb_logmel_plan_create(fp.n_fft, fp.hop, fp.clip_len, fp.total_spec_columns,
                                                           filt.ctypes.data_as(ctypes.c_void_p), filt.shape[0],
                                                           ctypes.byref(handle)), "wsb_logmel_plan_create")
            h = handle
            self._plans[key] = h
        return h

    def run(self, fp: FrontendPlan, audio_dev, win_desc_dev, n_windows, stream):
        """audio_dev f32 [N] device; win_desc_dev int64 [n_windows,3] device -> f32 [n_windows,80,n_cols]."""
        out = torch.empty((n_windows, N_MELS, fp.total_spec_columns), dtype=torch.float32, device=self.device)
        if n_windows:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.wsb_logmel_run(self.plan_for(fp), _ptr(audio_dev), _ptr(win_desc_dev), n_windows,
                                                   _ptr(out), ctypes.c_void_p(stream.cuda_stream)), "wsb_logmel_run")
        return out

    def __del__(self):
        try:
            for h in self._plans.values():
                self.lib.wsb_logmel_plan_destroy(h)
        except Exception:  # noqa: BLE001
            pass


def _model_config(cfg, max_batch):
    return _lib.ModelConfig(cfg["d_model"], cfg["encoder_attention_heads"], cfg["encoder_layers"], cfg["encoder_ffn_dim"],
                            cfg["vocab_size"], cfg["num_mel_bins"], 2 * cfg["max_source_positions"], cfg["max_target_positions"],
                            int(max_batch))


def auto_max_batch(cfg, device, lib, limit=240, weights_resident=False):
    """Largest batch of windows (<= `limit`) whose workspace fits in 80 % of the device's free memory after the
    weights: an unchanged scripts/segment.py passes no max_batch and must still get the wide-batch configuration."""
    free, _ = torch.cuda.mem_get_info(device)
    n_params = 2 * cfg["encoder_layers"] * 12 * cfg["d_model"] ** 2 + cfg["vocab_size"] * cfg["d_model"]
    budget = 0.8 * free - (0 if weights_resident else 4.5 * n_params)     # bf16 weights + folded copies + fp32 bits
    for b in (limit, 192, 160, 128, 96, 64, 48, 32, 16, 8, 4, 2, 1):
        if b <= limit and lib.wsb_workspace_bytes_for(ctypes.byref(_model_config(cfg, b))) <= budget:
            return b
    raise _lib.WsbError("not enough free device memory for a single window's workspace")


class Engine:
    """One model replica on one GPU."""

    def __init__(self, model_path, device="cuda:0", max_batch=64, state=None, tensors=None, stream_priority=0):
        if not torch.cuda.is_available():
            raise _lib.WsbError("whisperseg_b200 needs a CUDA (sm_100) device; there is no CPU path")
        self.lib = _lib.load()
        self.device = torch.device(device)
        cfg, sd, gen = state if state is not None else load_checkpoint(model_path)
        self.hf_config = cfg
        self.max_batch = int(max_batch) if max_batch else auto_max_batch(cfg, self.device, self.lib, weights_resident=tensors is not None)
        with torch.cuda.device(self.device):
            # `tensors`: prepared device weights of another Engine on the same device (chunk pipelining: several
            # contexts -- workspaces, streams, CUDA graphs -- share one copy of the weights)
            self.tensors = tensors if tensors is not None else prepare_tensors(cfg, sd, gen, self.device)
            self.stream = torch.cuda.Stream(self.device, priority=stream_priority)
            mc = _model_config(cfg, self.max_batch)
            names = list(self.tensors.keys())
            c_names = (ctypes.c_char_p * len(names))(*[n.encode() for n in names])
            c_ptrs = (ctypes.c_void_p * len(names))(*[self.tensors[n].data_ptr() for n in names])
            handle = ctypes.c_void_p()
            torch.cuda.synchronize(self.device)
            _lib.check(self.lib.wsb_model_create(ctypes.byref(mc), c_names, c_ptrs, len(names), ctypes.byref(handle)),
                       "wsb_model_create")
        self.handle = handle
        self.d_model = cfg["d_model"]
        self.n_cols = 2 * cfg["max_source_positions"]
        self.T = cfg["max_source_positions"]
        self.logmel = LogmelRunner(self.device)
        self._stage, self._stage_evt, self._stage_elems = None, None, 4 << 20     # 2 x 16 MB pinned staging

    def workspace_bytes(self):
        return int(self.lib.wsb_model_workspace_bytes(self.handle))

    @property
    def fold_fallback(self):
        """True once the folded-LayerNorm guard switched this replica to the exact LayerNorm (csrc/engine.cu)."""
        return bool(self.lib.wsb_model_fold_fallback(self.handle))

    def _enter(self):
        self.stream.wait_stream(torch.cuda.current_stream(self.device))

    def _exit(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def encode(self, feats, want_hidden=False):
        B = feats.shape[0]
        assert feats.dtype == torch.float32 and feats.is_contiguous() and feats.device == self.device
        hidden = torch.empty((B, self.T, self.d_model), dtype=torch.float32, device=self.device) if want_hidden else None
        with torch.cuda.device(self.device):
            self._enter()
            _lib.check(self.lib.wsb_encode(self.handle, _ptr(feats), B, _ptr(hidden),
                                           ctypes.c_void_p(self.stream.cuda_stream)), "wsb_encode")
            self._exit()
        return hidden

    def generate(self, batch, prompt_ids, eos_id, pad_id, max_length, forced=None, use_graph=True):
        """Greedy tokens int32 [batch, max_length - len(prompt)] (device) and #positions computed."""
        n_new = max_length - len(prompt_ids)
        tokens = torch.empty((batch, n_new), dtype=torch.int32, device=self.device)
        prompt = (ctypes.c_int32 * len(prompt_ids))(*prompt_ids)
        n_steps = ctypes.c_int(0)
        if forced is not None:
            assert forced.dtype == torch.int32 and forced.shape == (batch, max_length) and forced.is_contiguous()
        with torch.cuda.device(self.device):
            self._enter()
            _lib.check(self.lib.wsb_generate(self.handle, batch, prompt, len(prompt_ids), eos_id, pad_id, max_length,
                                             _ptr(forced), _ptr(tokens), ctypes.byref(n_steps),
                                             (1 if use_graph else 0) | (0 if os.environ.get("WSB_NO_PDL") else 2) | (4 if os.environ.get("WSB_NO_COMPACT") else 0) | (8 if os.environ.get("WSB_NO_GEMV") else 0),
                                             ctypes.c_void_p(self.stream.cuda_stream)), "wsb_generate")
            self._exit()
        return tokens, n_steps.value

    def generate_beam(self, batch, num_beams, prompt_ids, eos_id, pad_id, max_length, length_penalty=1.0,
                      use_graph=True, return_scores=False):
        """HF-equivalent beam search (reference default num_beams=4, model.py:409/662): best hypothesis per
        window, int32 [batch, max_length - len(prompt)] (device), and #positions computed.
        batch * num_beams must not exceed the engine's max_batch."""
        n_new = max_length - len(prompt_ids)
        tokens = torch.empty((batch, n_new), dtype=torch.int32, device=self.device)
        scores = torch.empty((batch,), dtype=torch.float32, device=self.device)
        prompt = (ctypes.c_int32 * len(prompt_ids))(*prompt_ids)
        n_steps = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            self._enter()
            _lib.check(self.lib.wsb_generate_beam(self.handle, batch, num_beams, prompt, len(prompt_ids), eos_id, pad_id,
                                                  max_length, ctypes.c_float(length_penalty), _ptr(tokens), _ptr(scores),
                                                  ctypes.byref(n_steps), 1 if use_graph else 0,
                                                  ctypes.c_void_p(self.stream.cuda_stream)), "wsb_generate_beam")
            self._exit()
        if return_scores:
            return tokens, n_steps.value, scores
        return tokens, n_steps.value

    def features(self, fp: FrontendPlan, audio, windows):
        """Host float32 audio + window list -> device features (H2D through pinned memory)."""
        return self.features_sliced(fp, audio, windows, 0, len(audio))

    def features_sliced(self, fp: FrontendPlan, piece, windows, slice_start, n_total):
        """`piece` = audio[slice_start : slice_start + len(piece)] holds every in-range sample the
        windows touch; window starts are given in whole-recording coordinates."""
        with torch.cuda.device(self.device):
            audio_dev = self.upload_audio(piece)
            desc_dev = self.window_descriptors(windows, slice_start, len(piece))
            return self.features_device(fp, audio_dev, desc_dev, len(windows))

    def upload_audio(self, piece):
        a = np.ascontiguousarray(piece, dtype=np.float32)
        if a.size == 0:
            return torch.zeros(4, dtype=torch.float32, device=self.device)
        # Two persistent pinned staging buffers: the host copy of chunk i + 1 overlaps the H2D copy of chunk i
        # (a fresh pin_memory() of the whole recording costs a page-locking allocation per call).
        with torch.cuda.device(self.device):
            n = a.size
            src = torch.from_numpy(a)
            dev = torch.empty(n, dtype=torch.float32, device=self.device)
            if self._stage is None:
                self._stage = [torch.empty(self._stage_elems, dtype=torch.float32).pin_memory() for _ in range(2)]
                self._stage_evt = [torch.cuda.Event(), torch.cuda.Event()]
            for i, off in enumerate(range(0, n, self._stage_elems)):
                k, m = i & 1, min(self._stage_elems, n - off)
                if i >= 2:
                    self._stage_evt[k].synchronize()           # the copy that last read this buffer has finished
                self._stage[k][:m].copy_(src[off:off + m])
                dev[off:off + m].copy_(self._stage[k][:m], non_blocking=True)
                self._stage_evt[k].record()
            for k in range(min(2, -(-n // self._stage_elems))):    # leave both buffers idle for the next caller
                self._stage_evt[k].synchronize()
            return dev

    def window_descriptors(self, windows, slice_start, n_valid):
        desc = np.array([[w.start - slice_start, 0, n_valid] for w in windows], dtype=np.int64).reshape(-1, 3)
        return torch.from_numpy(desc).to(self.device)

    def features_device(self, fp: FrontendPlan, audio_dev, desc_dev, n_windows):
        with torch.cuda.device(self.device):
            self._enter()
            out = self.logmel.run(fp, audio_dev, desc_dev, n_windows, self.stream)
            self._exit()
            for t in (out, audio_dev, desc_dev):
                t.record_stream(self.stream)
        return out

    def __del__(self):
        try:
            self.lib.wsb_model_destroy(self.handle)
        except Exception:  # noqa: BLE001
            pass
