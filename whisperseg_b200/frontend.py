"""Host-side planning for the fused log-mel kernel (K1).

Only index math and table construction live here; the arithmetic on samples happens in
`csrc/logmel.cu`.  Mirrors, for the hot path only:
  * reference audio_utils.py:32-43   get_n_fft_given_sr
  * reference audio_utils.py:45-76   WhisperSegFeatureExtractor.__init__ (hop, n_fft, mel band)
  * reference model.py:127-166       sliding windows, trial offsets, zero padding
  * HF transformers audio_utils.mel_filter_bank (norm="slaney", mel_scale="slaney") -- the
    reference's third-party dependency; its published formula is re-derived below.
"""
from dataclasses import dataclass

import numpy as np

N_MELS = 80


def get_n_fft_given_sr(sr):
    for limit, n_fft in ((32000, 512), (80000, 1024), (150000, 2048), (300000, 4096)):
        if sr <= limit:
            return n_fft
    return 8192


def _slaney_hz_to_mel(hz):
    hz = np.asarray(hz, dtype=np.float64)
    lin = 3.0 * hz / 200.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log = 15.0 + np.log(hz / 1000.0) * (27.0 / np.log(6.4))
    return np.where(hz >= 1000.0, log, lin)


def _slaney_mel_to_hz(mel):
    mel = np.asarray(mel, dtype=np.float64)
    lin = 200.0 * mel / 3.0
    log = 1000.0 * np.exp((np.log(6.4) / 27.0) * (mel - 15.0))
    return np.where(mel >= 15.0, log, lin)


def slaney_filterbank(n_fft, sr, min_frequency, max_frequency, n_mels=N_MELS):
    """float64 [n_fft//2+1, n_mels]: area-normalised triangles on the slaney mel scale.

    Bin centre frequencies are linspace(0, sr//2, n_freq) -- integer division, as the
    reference's dependency does (matters for odd sampling rates)."""
    n_freq = n_fft // 2 + 1
    edges = _slaney_mel_to_hz(np.linspace(float(_slaney_hz_to_mel(min_frequency)),
                                          float(_slaney_hz_to_mel(max_frequency)), n_mels + 2))
    bins = np.linspace(0, sr // 2, n_freq)
    width = np.diff(edges)
    rel = edges[None, :] - bins[:, None]                    # [n_freq, n_mels+2]
    falling = -rel[:, :-2] / width[:-1]
    rising = rel[:, 2:] / width[1:]
    tri = np.maximum(0.0, np.minimum(falling, rising))
    tri *= (2.0 / (edges[2:] - edges[:-2]))[None, :]
    return tri


@dataclass
class Window:
    trial_id: int
    offset_time: float       # seconds of the window start relative to the un-padded audio
    start: int               # first sample in the audio (negative inside a trial's left padding)
    n_valid: int             # samples of (padded) audio in the window before zero padding
    clip_seconds: float

    def as_tuple(self):
        return (self.trial_id, self.offset_time, None, self.clip_seconds)


class FrontendPlan:
    """Everything K1 needs that depends only on (sr, spec_time_step, min_frequency)."""

    def __init__(self, sr, spec_time_step, min_frequency=None, max_frequency=None, total_spec_columns=1000):
        self.sr = int(sr)
        self.spec_time_step = spec_time_step
        self.total_spec_columns = int(total_spec_columns)
        self.hop = int(spec_time_step * sr)                                 # audio_utils.py:48
        if self.hop < 1:
            raise ValueError("spec_time_step * sr must be >= 1 sample")
        self.n_fft = get_n_fft_given_sr(sr)
        self.min_frequency = 0 if min_frequency is None else min_frequency
        self.max_frequency = sr // 2 if max_frequency is None else max_frequency
        self.clip_duration = total_spec_columns * spec_time_step
        self.clip_len = int(self.clip_duration * sr)                         # model.py:133
        self.n_frames = self.clip_len // self.hop                            # STFT frames kept (last dropped)
        if self.clip_len <= self.n_fft // 2:
            raise ValueError("window shorter than n_fft/2: reflect padding undefined")
        self.mel_filters = slaney_filterbank(self.n_fft, self.sr, self.min_frequency, self.max_frequency)

    def windows(self, n_samples, num_trials):
        """model.py:136-165 without touching samples."""
        out = []
        for trial_id in range(num_trials):
            padding_time = np.round(self.clip_duration * trial_id / num_trials / self.spec_time_step) * self.spec_time_step
            n_pad = int(padding_time * self.sr)
            padded_len = n_pad + n_samples
            for pos in range(0, max(padded_len, 1), self.clip_len):
                n_valid = max(0, min(self.clip_len, padded_len - pos))
                out.append(Window(trial_id, pos / self.sr - padding_time, pos - n_pad, n_valid, n_valid / self.sr))
        return out

    def logmel_bytes_per_window(self):
        """Algorithmic HBM bytes of K1 per window: samples read + f32 features written."""
        return 4 * self.clip_len + 4 * N_MELS * self.total_spec_columns

    def logmel_flops_per_window(self):
        n = self.n_fft
        return self.n_frames * (2.5 * n * np.log2(n) + 3 * (n // 2 + 1) + 2 * 2 * (n // 2 + 1))
