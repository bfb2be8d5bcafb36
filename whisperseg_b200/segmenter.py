"""Drop-in segmenter classes: the reference's Python API for the hot path, backed by libwsb.

Mirrors reference model.py:
  SegmenterBase            (:118-470)  -> SegmenterBase here (same attributes and method names)
  WhisperSegmenter         (:625-676)  -> WhisperSegmenter
  WhisperSegmenterFast     (:678-746)  -> WhisperSegmenterFast (same engine; the reference's CT2
                                          directory layout `<path>/hf_model/` is accepted)
  WhisperSegmenterForEval  (:572-622)  -> WhisperSegmenterForEval (model_path form)
`segment()` keeps the reference signature and defaults (model.py:398-414) and returns
{"onset": [float], "offset": [float], "cluster": [str]}.

Design differences (B200-first):
  * the front-end runs on the GPU for all windows of a call in one launch (the reference loops
    over windows on the CPU, model.py:146-165); features never leave HBM;
  * windows are independent, so `batch_size` only bounds memory: the engine decodes up to
    `max_batch` windows together whatever `batch_size` says.  A window's tokens do not depend on its neighbours in
    exact arithmetic; in bf16 the decode linear layers switch implementation with the number of live rows (<= 64: fused
    LayerNorm + linear kernels, above: tensor-core split-K), so a near-tie between two tokens (top-1/top-2 margin below
    the bf16 noise floor, tests/test_noise_floor.py) can resolve differently in a different batch;
  * `num_beams=1` decodes greedily (BASELINE.json north_star); `num_beams` 2..4 -- the reference default
    is 4 -- runs HF-equivalent beam search on the device (csrc/beam.cu), `length_penalty` honoured;
    `top_k` / `top_p` only matter for sampling and the reference always passes top_k=1.
"""
import json
import os
import tempfile
import threading

import numpy as np
import torch

from . import postprocess as pp
from .engine import Engine
from .frontend import FrontendPlan, get_n_fft_given_sr
from .tokens import TokenTable
from .weights import load_checkpoint

DEFAULT_MAX_BATCH = 240       # windows decoded together when the caller does not say (shrunk to what fits in free HBM)


def resolve_model_path(model_path):
    """A local checkpoint directory, or a Hugging Face hub id resolved the way the reference's download_model does
    (reference model.py:37-56: snapshot_download into the local cache).  Accepts the CTranslate2 repo layout
    (`<path>/hf_model/` holds the HF config + tokenizer + weights, model.py:694-702)."""
    path = str(model_path)
    if not os.path.isdir(path):
        try:
            from huggingface_hub import snapshot_download
            path = snapshot_download(path, cache_dir=os.environ.get("WHISPERSEG_MODEL_CACHE"))
        except Exception as e:  # noqa: BLE001
            raise FileNotFoundError(
                "model_path %r is neither a local directory nor a Hugging Face hub id that could be resolved here (%s: %s). "
                "Offline: download the checkpoint once (huggingface_hub.snapshot_download) and pass the directory."
                % (model_path, type(e).__name__, e)) from e
    hf_dir = os.path.join(path, "hf_model")
    has_weights = any(os.path.isfile(os.path.join(path, f)) for f in
                      ("model.safetensors", "model.safetensors.index.json", "pytorch_model.bin", "pytorch_model.bin.index.json"))
    if os.path.isdir(hf_dir) and not has_weights:
        if not any(os.path.isfile(os.path.join(hf_dir, f)) for f in ("model.safetensors", "model.safetensors.index.json",
                                                                      "pytorch_model.bin", "pytorch_model.bin.index.json")):
            raise FileNotFoundError(
                "%s is a CTranslate2 export (model.bin): its hf_model/ folder holds config and tokenizer but no HF weights. "
                "whisperseg_b200 loads the HF checkpoint (e.g. nccratliri/whisperseg-large-ms, not the -ct2 repo)." % path)
        return hf_dir
    return path


class SegmenterBase:
    def __init__(self):
        self.total_spec_columns = None
        self.precision_bits = 3
        self.cluster_codebook = None
        self.inverse_cluster_codebook = None
        self.default_segmentation_config = {}
        self.engines = []
        self.tokenizer = None
        self.last_stats = {}

    # ------------------------------------------------------------------ construction helpers
    def _setup(self, model_path, device, device_ids, max_batch, state=None, tokenizer_dir=None):
        if device is None:
            device = "cuda"
        if device == "cpu" or not torch.cuda.is_available():
            raise RuntimeError("whisperseg_b200 has no CPU path: a CUDA sm_100 (B200) device is required")
        if state is None:
            ckpt_dir = resolve_model_path(model_path)
            state = load_checkpoint(ckpt_dir)
            tokenizer_dir = ckpt_dir
        else:
            ckpt_dir = None
        cfg = state[0]
        self.model_config = cfg
        self.total_spec_columns = cfg["total_spec_columns"]                      # model.py:639
        self.cluster_codebook = cfg["cluster_codebook"]                          # model.py:640
        self.inverse_cluster_codebook = {v: k for k, v in self.cluster_codebook.items()}
        if "default_segmentation_config" in cfg:                                 # model.py:643-644
            self.default_segmentation_config.update(cfg["default_segmentation_config"])
        self.tokenizer = TokenTable.from_pretrained(tokenizer_dir)
        self.device_list = [torch.device("cuda", int(g)) for g in device_ids]
        self.engines = [Engine(ckpt_dir, dev, max_batch=max_batch, state=state) for dev in self.device_list]
        self.max_batch = min(e.max_batch for e in self.engines)

    def update_cluster_codebook(self, cluster_codebook):
        self.cluster_codebook = cluster_codebook
        self.inverse_cluster_codebook = {v: k for k, v in cluster_codebook.items()}

    # ------------------------------------------------------------------ hot path
    def get_sliced_audios_features(self, audio, sr, min_frequency, spec_time_step, num_trials, engine=None):
        """Same contract as model.py:127-166, except that item[2] is a device tensor view."""
        plan = FrontendPlan(sr, spec_time_step, min_frequency, total_spec_columns=self.total_spec_columns)
        wins = plan.windows(len(audio), num_trials)
        eng = engine or self.engines[0]
        feats = eng.features(plan, audio, wins)
        return [(w.trial_id, w.offset_time, feats[i], w.clip_seconds) for i, w in enumerate(wins)]

    def _generate_on(self, eng, feats, max_length, status_monitor, texts_out, slot, num_beams=1, length_penalty=1.0):
        """feats: device tensor [n,80,cols] on eng.device -> list of decoded strings.  The ids of chunk i are copied
        to the host and turned into text on a worker thread while the GPU already runs chunk i + 1 (the reference
        decodes text between two generate() calls, model.py:667-668)."""
        from concurrent.futures import ThreadPoolExecutor
        tok = self.tokenizer
        n = feats.shape[0]
        steps = 0
        per_call = eng.max_batch if num_beams == 1 else max(1, eng.max_batch // num_beams)

        def to_text(ids, done):
            done.synchronize()
            return tok.batch_decode(ids.cpu().numpy())

        pending = []
        with ThreadPoolExecutor(max_workers=1) as pool:
            for pos in range(0, n, per_call):
                chunk = feats[pos:pos + per_call].contiguous()
                eng.encode(chunk)
                if num_beams == 1:
                    ids, n_steps = eng.generate(chunk.shape[0], tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, max_length)
                else:
                    ids, n_steps = eng.generate_beam(chunk.shape[0], num_beams, tok.prompt_ids, tok.eos_token_id,
                                                     tok.pad_token_id, max_length, length_penalty)
                steps += n_steps
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(eng.device))
                pending.append(pool.submit(to_text, ids, done))
                if status_monitor is not None:                                       # model.py:672-674
                    status_monitor["progress"] = int(100 * min(1, (pos + per_call) / max(n, 1)))
            texts = [t for fut in pending for t in fut.result()]
        texts_out[slot] = texts
        self.last_stats["decode_steps"] = self.last_stats.get("decode_steps", 0) + steps

    def generate_segment_text(self, sliced_audios_features, batch_size, max_length, num_beams, top_k=1, top_p=1.0,
                              length_penalty=1.0, status_monitor=None):
        """model.py:169-189: contiguous shards of the window list, one worker per device."""
        num_beams = int(num_beams)
        if not 1 <= num_beams <= 4:
            raise ValueError("whisperseg_b200 supports num_beams in [1, 4] (reference default 4), got %d" % num_beams)
        if num_beams > 1 and num_beams > min(e.max_batch for e in self.engines):
            raise ValueError("num_beams exceeds the engine's max_batch")
        n = len(sliced_audios_features)
        if n == 0:
            return []
        feats_all = [item[2] for item in sliced_audios_features]
        n_dev = len(self.engines)
        per = int(np.ceil(n / n_dev))
        outs, threads = {}, []
        for slot, pos in enumerate(range(0, n, per)):
            eng = self.engines[slot]
            shard = feats_all[pos:pos + per]
            stacked = torch.stack([f if torch.is_tensor(f) else torch.from_numpy(np.asarray(f)) for f in shard]).to(eng.device)
            args = (eng, stacked, max_length, status_monitor if slot == 0 else None, outs, slot, num_beams, length_penalty)
            if n_dev == 1:
                self._generate_on(*args)
            else:
                t = threading.Thread(target=self._generate_on, args=args)
                t.start()
                threads.append(t)
        for t in threads:
            t.join()
        return [txt for slot in sorted(outs) for txt in outs[slot]]

    def extract_segments(self, text, spec_time_step):
        return pp.segments_from_text(text, spec_time_step, {v: k for k, v in self.cluster_codebook.items()})

    def parse_generation(self, generated_text_list, sliced_audios_features, min_segment_length, audio_duration,
                         spec_time_step, num_trials, eps, time_per_frame_for_voting, consolidation_method):
        return pp.parse_generation(generated_text_list, sliced_audios_features, min_segment_length, audio_duration,
                                   spec_time_step, num_trials, eps, time_per_frame_for_voting, consolidation_method,
                                   self.cluster_codebook, self.precision_bits)

    def consolidate_trials_by_clustering(self, trials, eps, min_samples):
        return pp.consolidate_trials_by_clustering(trials, eps, min_samples)

    def consolidate_trials_by_voting(self, trials, time_per_frame_for_voting):
        return pp.consolidate_trials_by_voting(trials, time_per_frame_for_voting, self.cluster_codebook)

    @torch.no_grad()
    def segment(self, audio, sr, min_frequency=None, spec_time_step=None, min_segment_length=None, eps=None,
                time_per_frame_for_voting=None, consolidation_method="clustering", max_length=448, batch_size=4,
                num_trials=1, num_beams=4, top_k=1, top_p=1.0, length_penalty=1.0, status_monitor=None):
        if min_frequency is None:
            min_frequency = self.default_segmentation_config.get("min_frequency", 0)
        if spec_time_step is None:
            spec_time_step = self.default_segmentation_config.get("spec_time_step", 0.0025)
        ratio = pp.RATIO_DECODING_TIME_STEP_TO_SPEC_TIME_STEP
        if min_segment_length is None:
            min_segment_length = spec_time_step * ratio
        if eps is None:
            eps = spec_time_step * ratio * 4
        if time_per_frame_for_voting is None:
            time_per_frame_for_voting = spec_time_step
        audio = np.asarray(audio)
        self.last_stats = {}
        sliced = self.get_sliced_audios_features(audio, sr, min_frequency, spec_time_step, num_trials)
        self.last_stats["n_windows"] = len(sliced)
        texts = self.generate_segment_text(sliced, batch_size, max_length, num_beams, top_k, top_p, length_penalty,
                                           status_monitor)
        prediction = self.parse_generation(texts, sliced, min_segment_length, len(audio) / sr, spec_time_step, num_trials,
                                           eps, time_per_frame_for_voting, consolidation_method)
        return pp.correct_fft_blur_and_dedupe(prediction, sr, get_n_fft_given_sr(sr))

    @torch.no_grad()
    def segment_many(self, audios, sr, min_frequency=None, spec_time_step=None, min_segment_length=None, eps=None,
                     time_per_frame_for_voting=None, consolidation_method="clustering", max_length=448, num_trials=1,
                     status_monitor=None, num_beams=4, length_penalty=1.0):
        """Folder mode (reference scripts/segment.py:39-56 loops over files and calls segment() on each, so a
        0.5 s clip runs at batch size 1).  Here the windows of ALL clips are flattened into one list: one H2D
        copy of the concatenated samples, one log-mel launch (per-window [lo, hi) bounds keep clips from
        leaking into each other), encoder/decoder batches filled across clip boundaries, then the token
        streams are regrouped per clip and post-processed exactly like segment().  Returns a list of
        predictions, one per clip: the same windows, features and post-processing as per-clip segment() calls (tokens equal
        up to the near-tie caveat in the module docstring)."""
        if min_frequency is None:
            min_frequency = self.default_segmentation_config.get("min_frequency", 0)
        if spec_time_step is None:
            spec_time_step = self.default_segmentation_config.get("spec_time_step", 0.0025)
        ratio = pp.RATIO_DECODING_TIME_STEP_TO_SPEC_TIME_STEP
        if min_segment_length is None:
            min_segment_length = spec_time_step * ratio
        if eps is None:
            eps = spec_time_step * ratio * 4
        if time_per_frame_for_voting is None:
            time_per_frame_for_voting = spec_time_step
        plan = FrontendPlan(sr, spec_time_step, min_frequency, total_spec_columns=self.total_spec_columns)
        eng = self.engines[0]
        pieces, descs, owners, per_clip_windows = [], [], [], []
        base = 0
        for ci, a in enumerate(audios):
            a = np.ascontiguousarray(np.asarray(a), dtype=np.float32)
            wins = plan.windows(len(a), num_trials)
            per_clip_windows.append(wins)
            for w in wins:
                descs.append([base + w.start, base, base + len(a)])
                owners.append(ci)
            pieces.append(a)
            pad = (-len(a)) % 4                          # keep every clip 16-byte aligned in the device buffer
            if pad:
                pieces.append(np.zeros(pad, dtype=np.float32))
            base += len(a) + pad
        if not descs:
            return []
        audio_dev = eng.upload_audio(np.concatenate(pieces) if pieces else np.zeros(0, np.float32))
        desc_dev = torch.from_numpy(np.asarray(descs, dtype=np.int64)).to(eng.device)
        feats = eng.features_device(plan, audio_dev, desc_dev, len(descs))
        self.last_stats = {"n_windows": len(descs)}
        outs = {}
        if not 1 <= int(num_beams) <= 4:
            raise ValueError("whisperseg_b200 supports num_beams in [1, 4], got %r" % (num_beams,))
        self._generate_on(eng, feats, max_length, status_monitor, outs, 0, int(num_beams), length_penalty)
        texts = outs[0]
        results, pos = [], 0
        n_fft = get_n_fft_given_sr(sr)
        for ci, a in enumerate(audios):
            wins = per_clip_windows[ci]
            clip_texts = texts[pos:pos + len(wins)]
            pos += len(wins)
            pred = self.parse_generation(clip_texts, [w.as_tuple() for w in wins], min_segment_length, len(a) / sr,
                                         spec_time_step, num_trials, eps, time_per_frame_for_voting, consolidation_method)
            results.append(pp.correct_fft_blur_and_dedupe(pred, sr, n_fft))
        return results

    # ------------------------------------------------------------------ scoring (model.py:474-569)
    def compute_syllable_score(self, prediction_on_offset_list, label_on_offset_list, tolerance):
        return pp.compute_syllable_score(prediction_on_offset_list, label_on_offset_list, tolerance)

    def segment_score(self, prediction, label, target_cluster=None, tolerance=None):
        return pp.segment_score(prediction, label, target_cluster, tolerance,
                                self.default_segmentation_config.get("spec_time_step", 0.0025))

    def frame_score(self, prediction, label, target_cluster=None, time_per_frame_for_scoring=None):
        return pp.frame_score(prediction, label, target_cluster, time_per_frame_for_scoring,
                              self.default_segmentation_config.get("spec_time_step", 0.0025))


class WhisperSegmenter(SegmenterBase):
    def __init__(self, model_path, device=None, device_ids=[0, ], max_batch=None):
        """`max_batch` (not in the reference signature): windows encoded/decoded together; None = up to
        DEFAULT_MAX_BATCH, shrunk to what fits in the device's free memory (engine.py: auto_max_batch)."""
        super().__init__()
        self._setup(model_path, device, device_ids, max_batch)

    @classmethod
    def from_state(cls, state, tokenizer_dir, device=None, device_ids=(0,), max_batch=None):
        """Build from an in-memory (config dict, state dict, generation dict) triple -- what
        weights.load_checkpoint returns -- plus a directory holding the tokenizer files."""
        self = cls.__new__(cls)
        SegmenterBase.__init__(self)
        self._setup(None, device, list(device_ids), max_batch, state=state, tokenizer_dir=tokenizer_dir)
        return self


class WhisperSegmenterFast(WhisperSegmenter):
    """The reference's CTranslate2-backed class; here it is the same sm_100a engine."""


class WhisperSegmenterForEval(SegmenterBase):
    def __init__(self, model_path=None, device=None, model=None, tokenizer=None, max_batch=None):
        """reference model.py:573-595: either a checkpoint path, or an in-memory HF `WhisperForConditionalGeneration`
        plus its tokenizer (evaluate.py / train.py pass the model they are training).  The in-memory form snapshots
        the model's weights, config and generation config into an engine: later updates of `model` are not seen."""
        super().__init__()
        dev = torch.device(device) if device is not None else torch.device("cuda", 0)
        if model_path is not None:
            self._setup(model_path, "cuda", [dev.index or 0], max_batch)
        else:
            if model is None or tokenizer is None:
                raise ValueError("WhisperSegmenterForEval needs model_path, or model and tokenizer")
            hf = getattr(model, "module", model)                      # nn.DataParallel wrapper (train.py:132)
            cfg = hf.config.to_dict()
            gen = hf.generation_config.to_dict() if getattr(hf, "generation_config", None) is not None else {}
            sd = {k: v.detach().to("cpu") for k, v in hf.state_dict().items()}
            tokdir = tempfile.mkdtemp(prefix="wsb_tok_")
            tokenizer.save_pretrained(tokdir)
            self._setup(None, "cuda", [dev.index or 0], max_batch, state=(cfg, sd, gen), tokenizer_dir=tokdir)
        self.device = self.device_list[0]


def save_synthetic_config_fields(path, **fields):
    """Utility for tests/bench: add WhisperSeg fields to a config.json."""
    p = os.path.join(path, "config.json")
    cfg = json.load(open(p))
    cfg.update(fields)
    json.dump(cfg, open(p, "w"))
