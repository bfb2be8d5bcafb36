"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the WhisperSeg log-mel front-end.

CPU oracle for the fused log-mel kernel (K1).  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module; the product path
(`whisperseg_b200/`) never does.

What it restates, and from where:
  * `get_n_fft_given_sr`            -- reference audio_utils.py:32-43
  * hop / mel-band configuration    -- reference audio_utils.py:45-76 (WhisperSegFeatureExtractor)
  * sliding windows + trial offsets -- reference model.py:127-166 (get_sliced_audios_features)
  * the feature arithmetic lives in a third-party dependency that is NOT vendored in
    /root/reference: HuggingFace `transformers` (pinned 4.38.2 in requirements.txt:1, 5.5.0 in
    this image): `WhisperFeatureExtractor._torch_extract_fbank_features`
    (models/whisper/feature_extraction_whisper.py:135-164) and `mel_filter_bank`
    (audio_utils.py:453-546, hertz_to_mel :263-296, _create_triangular_filter_bank :356-375).
    Its published algorithm is restated below.

Pinned by: tests/golden/frontend_*.npz (outputs of the unmodified reference, produced by
oracle/gen_golden.py) -- see tests/test_oracle_frontend.py.
"""
import numpy as np

N_MELS = 80


def get_n_fft_given_sr(sr):
    # reference audio_utils.py:32-43
    if sr <= 32000:
        return 512
    if sr <= 80000:
        return 1024
    if sr <= 150000:
        return 2048
    if sr <= 300000:
        return 4096
    return 8192


def _hertz_to_mel_slaney(freq):
    # HF audio_utils.py:263-296 (mel_scale="slaney")
    freq = np.asarray(freq, dtype=np.float64)
    min_log_hertz, min_log_mel = 1000.0, 15.0
    logstep = 27.0 / np.log(6.4)
    mels = 3.0 * freq / 200.0
    log_region = freq >= min_log_hertz
    mels = np.where(log_region, min_log_mel + np.log(np.maximum(freq, 1e-300) / min_log_hertz) * logstep, mels)
    return mels


def _mel_to_hertz_slaney(mels):
    mels = np.asarray(mels, dtype=np.float64)
    min_log_hertz, min_log_mel = 1000.0, 15.0
    logstep = np.log(6.4) / 27.0
    freq = 200.0 * mels / 3.0
    log_region = mels >= min_log_mel
    freq = np.where(log_region, min_log_hertz * np.exp(logstep * (mels - min_log_mel)), freq)
    return freq


def mel_filter_bank_slaney(n_freq, n_mels, min_frequency, max_frequency, sr):
    """float64 [n_freq, n_mels]; HF audio_utils.py:453-546 with norm='slaney', mel_scale='slaney'."""
    mel_min = float(_hertz_to_mel_slaney(min_frequency))
    mel_max = float(_hertz_to_mel_slaney(max_frequency))
    mel_freqs = np.linspace(mel_min, mel_max, n_mels + 2)
    filter_freqs = _mel_to_hertz_slaney(mel_freqs)
    fft_freqs = np.linspace(0, sr // 2, n_freq)
    filter_diff = np.diff(filter_freqs)
    slopes = np.expand_dims(filter_freqs, 0) - np.expand_dims(fft_freqs, 1)
    down = -slopes[:, :-2] / filter_diff[:-1]
    up = slopes[:, 2:] / filter_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (filter_freqs[2:n_mels + 2] - filter_freqs[:n_mels])
    return fb * np.expand_dims(enorm, 0)


class FrontendConfig:
    """hop / n_fft / filterbank for (sr, spec_time_step, min_frequency): audio_utils.py:45-76."""

    def __init__(self, sr, spec_time_step, min_frequency=None, max_frequency=None, total_spec_columns=1000):
        self.sr = sr
        self.spec_time_step = spec_time_step
        self.hop = int(spec_time_step * sr)
        self.n_fft = get_n_fft_given_sr(sr)
        self.min_frequency = 0 if min_frequency is None else min_frequency
        self.max_frequency = sr // 2 if max_frequency is None else max_frequency
        self.total_spec_columns = total_spec_columns
        self.mel_filters = mel_filter_bank_slaney(1 + self.n_fft // 2, N_MELS, self.min_frequency,
                                                  self.max_frequency, sr)
        clip_duration = total_spec_columns * spec_time_step
        self.clip_duration = clip_duration
        self.clip_len = int(clip_duration * sr)              # model.py:133


def logmel_clip(clip, cfg, dtype=np.float64):
    """One zero-padded clip -> [80, clip_len//hop] normalised log-mel.

    HF feature_extraction_whisper.py:135-164: periodic Hann, center=True reflect padding,
    frames 0..L//hop with the last dropped, power, mel, log10(clamp 1e-10), clamp to the
    clip-global max - 8, (x+4)/4.  `dtype` float64 gives the exact-arithmetic answer; float32
    mimics the reference's precision.
    """
    x = np.asarray(clip, dtype=dtype)
    n_fft, hop = cfg.n_fft, cfg.hop
    half = n_fft // 2
    n_frames = len(x) // hop
    padded = np.pad(x, (half, half), mode="reflect")
    win = (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft, dtype=np.float64) / n_fft)).astype(dtype)
    idx = (np.arange(n_frames)[:, None] * hop) + np.arange(n_fft)[None, :]
    out = np.empty((N_MELS, n_frames), dtype=dtype)
    filt_t = cfg.mel_filters.T.astype(np.float32).astype(dtype)      # HF casts the bank to f32
    step = 2048
    for s in range(0, n_frames, step):
        fr = padded[idx[s:s + step]] * win[None, :]
        spec = np.fft.rfft(fr, axis=1)
        power = (spec.real.astype(dtype) ** 2 + spec.imag.astype(dtype) ** 2)
        out[:, s:s + step] = filt_t @ power.T
    log_spec = np.log10(np.maximum(out, dtype(1e-10)))
    log_spec = np.maximum(log_spec, log_spec.max() - dtype(8.0))
    return ((log_spec + dtype(4.0)) / dtype(4.0))


def window_plan(n_samples, sr, spec_time_step, num_trials, total_spec_columns=1000):
    """(trial_id, offset_time, start_sample_in_audio, n_valid, clip_seconds) per window.

    model.py:127-166.  `start_sample_in_audio` may be negative (the trial's left padding);
    samples outside [0, n_samples) are zeros.  `n_valid` = len(audio_clip) before zero padding.
    """
    clip_duration = total_spec_columns * spec_time_step
    clip_len = int(clip_duration * sr)
    plan = []
    for trial_id in range(num_trials):
        padding_time = np.round(clip_duration * trial_id / num_trials / spec_time_step) * spec_time_step
        num_padding_samples = int(padding_time * sr)
        padded_len = num_padding_samples + n_samples
        for pos in range(0, max(padded_len, 1), clip_len):
            offset_time = pos / sr - padding_time
            n_valid = max(0, min(clip_len, padded_len - pos))
            plan.append((trial_id, offset_time, pos - num_padding_samples, n_valid, n_valid / sr))
    return plan


def sliced_audio_features(audio, sr, min_frequency, spec_time_step, num_trials, total_spec_columns=1000,
                          dtype=np.float64):
    """Restatement of SegmenterBase.get_sliced_audios_features (model.py:127-166)."""
    cfg = FrontendConfig(sr, spec_time_step, min_frequency, total_spec_columns=total_spec_columns)
    audio = np.asarray(audio, dtype=np.float32)
    out = []
    for trial_id, offset_time, start, n_valid, clip_sec in window_plan(len(audio), sr, spec_time_step,
                                                                       num_trials, total_spec_columns):
        clip = np.zeros(cfg.clip_len, dtype=np.float32)
        lo, hi = max(start, 0), min(start + cfg.clip_len, len(audio))
        if hi > lo:
            clip[lo - start:hi - start] = audio[lo:hi]
        f = logmel_clip(clip, cfg, dtype=dtype)[:, :total_spec_columns]
        min_val = f.min() if f.shape[1] > 0 else 0
        if f.shape[1] < total_spec_columns:                                 # model.py:155-161
            f = np.concatenate([f, min_val * np.ones((f.shape[0], total_spec_columns - f.shape[1]))], axis=1)
        out.append((trial_id, offset_time, f.astype(np.float32), clip_sec))
    return out
