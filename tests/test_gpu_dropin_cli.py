"""Drop-in proof for the reference's command line (reference scripts/segment.py:18-72) against the root `model.py` shim.

/root/reference does not exist on the GPU box and this container has no GPU, so the reference's script itself can never
run next to the CUDA engine; what runs here is its calling sequence, statement for statement:
  `from model import WhisperSegmenter, WhisperSegmenterFast`            (segment.py:9)   -- the repo root's model.py
  try Fast(model_path, device=..., device_ids=...) except: WhisperSegmenter(...)        (:33-37)
  folder mode: glob *.wav, `librosa.load(path, sr=None)`, segment(audio, sr, min_frequency=, spec_time_step=,
  num_trials=, batch_size=) with NOTHING else passed (num_beams stays at the reference default 4), rows appended to
  a {"filename","onset","offset","cluster"} table, DataFrame.to_csv(index=False)                   (:39-56, 65-72)
`librosa` is not installed here: a stub module backed by whisperseg_b200.audio_io.load_audio (same float32 scaling as
libsndfile, tests/test_audio_io.py) stands in for it.  The CSV must equal the one built from direct `segment()` calls
and from the batched folder runner (audio_io.segment_files)."""
import glob
import io
import os
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_wav(path, audio, sr):
    from scipy.io import wavfile
    wavfile.write(path, sr, (np.clip(audio, -1, 1) * 32767).astype(np.int16))


def test_reference_cli_calling_sequence(tmp_path, monkeypatch):
    import pandas as pd
    from tools import synth
    from whisperseg_b200 import audio_io
    # a checkpoint directory as the CLI receives it (--model_path) and a folder of wav files (--audio_folder)
    hf = synth.make_hf_model("tiny", seed=0, confident=True, default_segmentation_config=dict(
        sr=16000, min_frequency=0, spec_time_step=0.01, species="human"))
    model_path = synth.save_checkpoint(hf, str(tmp_path / "ckpt"))
    folder = tmp_path / "wavs"
    folder.mkdir()
    for i, seconds in enumerate((3.2, 11.0, 24.5, 0.4)):
        _write_wav(str(folder / ("clip%d.wav" % i)), synth.synth_audio(seconds, 16000, seed=70 + i), 16000)
    # --- the stub for `import librosa` (segment.py:10)
    librosa = types.ModuleType("librosa")
    librosa.load = lambda src, sr=None, mono=True: audio_io.load_audio(src, sr=sr, mono=mono)
    monkeypatch.setitem(sys.modules, "librosa", librosa)
    monkeypatch.syspath_prepend(ROOT)                    # segment.py:4-6 puts the repo root first
    sys.modules.pop("model", None)
    from model import WhisperSegmenter, WhisperSegmenterFast          # noqa: E402  (segment.py:9)
    import librosa as _librosa                                       # noqa: E402
    args = types.SimpleNamespace(model_path=model_path, device="cuda", device_ids=[0], batch_size=8, min_frequency=None,
                                 spec_time_step=None, num_trials=1, audio_folder=str(folder))
    try:                                                              # segment.py:33-37
        segmenter = WhisperSegmenterFast(args.model_path, device=args.device, device_ids=args.device_ids)
    except Exception:  # noqa: BLE001
        segmenter = WhisperSegmenter(args.model_path, device=args.device, device_ids=args.device_ids)
    assert segmenter.max_batch >= 8                                   # no max_batch passed: sized from free memory
    audio_path_list = glob.glob(args.audio_folder + "/*.wav") + glob.glob(args.audio_folder + "/*.WAV")
    overall = {"filename": [], "onset": [], "offset": [], "cluster": []}
    direct = {}
    for audio_path in audio_path_list:                                # segment.py:48-55
        audio, sr = _librosa.load(audio_path, sr=None)
        res = segmenter.segment(audio, sr, min_frequency=args.min_frequency, spec_time_step=args.spec_time_step,
                                num_trials=args.num_trials, batch_size=args.batch_size)
        overall["filename"] += [os.path.basename(audio_path)] * len(res["onset"])
        overall["onset"] += res["onset"]
        overall["offset"] += res["offset"]
        overall["cluster"] += res["cluster"]
        direct[audio_path] = res
    buf = io.StringIO()
    pd.DataFrame(overall).to_csv(buf, index=False)                    # segment.py:56, 65-70
    csv_cli = buf.getvalue()
    assert csv_cli.splitlines()[0] == "filename,onset,offset,cluster"
    assert len(overall["onset"]) >= 20 and all(isinstance(v, float) for v in overall["onset"] + overall["offset"])
    # the batched folder runner gives the same table
    per_file, table = audio_io.segment_files(segmenter, audio_path_list, workers=2, num_trials=1)
    buf2 = io.StringIO()
    pd.DataFrame(table).to_csv(buf2, index=False)
    assert buf2.getvalue() == csv_cli
    for path in audio_path_list:
        assert per_file[path] == direct[path]
    # a hub id cannot be resolved offline: the error must say so instead of a bare missing-file error
    with pytest.raises(FileNotFoundError, match="Hugging Face hub id"):
        WhisperSegmenter("nccratliri/whisperseg-large-ms", device="cuda", device_ids=[0])


def test_reference_service_calling_sequence(tmp_path, monkeypatch):
    """The Flask handler of the reference's segment_service.py (:56-111), statement for statement, without Flask: a
    base64 WAV (stereo here), `librosa.load(BytesIO, sr=sr, mono=False)`, channel pick, `segment(...)` with the request's
    keys (num_trials defaults to 3 in the service, :72), result through json (jsonify) -- plain floats and strings."""
    import base64
    import json
    from scipy.io import wavfile
    from tools import synth
    from whisperseg_b200 import audio_io
    hf = synth.make_hf_model("tiny", seed=0, confident=True, default_segmentation_config=dict(
        sr=16000, min_frequency=0, spec_time_step=0.01, species="human"))
    model_path = synth.save_checkpoint(hf, str(tmp_path / "ckpt"))
    left = synth.synth_audio(12.0, 16000, seed=81)
    right = synth.synth_audio(12.0, 16000, seed=82)
    wav = tmp_path / "stereo.wav"
    wavfile.write(str(wav), 16000, (np.stack([left, right], axis=1) * 32767).astype(np.int16))
    request_json = {"audio_file_base64_string": base64.b64encode(open(wav, "rb").read()).decode(), "sr": 16000,
                    "min_frequency": None, "spec_time_step": None, "channel_id": 1}
    librosa = types.ModuleType("librosa")
    librosa.load = lambda src, sr=None, mono=True: audio_io.load_audio(src, sr=sr, mono=mono)
    monkeypatch.setitem(sys.modules, "librosa", librosa)
    monkeypatch.syspath_prepend(ROOT)
    sys.modules.pop("model", None)
    from model import WhisperSegmenter, WhisperSegmenterFast          # noqa: E402  (segment_service.py:10)
    try:                                                              # segment_service.py:123-128
        segmenter = WhisperSegmenterFast(model_path, device="cuda", device_ids=[0])
    except Exception:  # noqa: BLE001
        segmenter = WhisperSegmenter(model_path, device="cuda", device_ids=[0])
    request_info = {k: v for k, v in request_json.items() if v is not None}
    sr = request_info["sr"]
    num_trials = request_info.get("num_trials", 3)
    channel_id = request_info.get("channel_id", 0)
    audio, _ = librosa.load(io.BytesIO(base64.b64decode(request_info["audio_file_base64_string"])), sr=sr, mono=False)
    assert audio.ndim == 2 and audio.shape[0] == 2
    audio = audio[channel_id]
    prediction = segmenter.segment(audio, sr=sr, min_frequency=request_info.get("min_frequency"),
                                   spec_time_step=request_info.get("spec_time_step"),
                                   min_segment_length=request_info.get("min_segment_length"), eps=request_info.get("eps"),
                                   num_trials=num_trials, batch_size=8)
    body = json.loads(json.dumps(prediction))                         # jsonify(prediction)
    assert set(body) == {"onset", "offset", "cluster"} and len(body["onset"]) == len(body["offset"]) == len(body["cluster"])
    # same audio, same arguments, direct call: identical
    again = segmenter.segment(np.ascontiguousarray(audio), sr, num_trials=3, batch_size=8)
    assert again == prediction
    # and the channel really was picked
    other = segmenter.segment(np.ascontiguousarray(audio_io.load_audio(str(wav), sr=sr, mono=False)[0][0]), sr, num_trials=3)
    assert isinstance(other["onset"], list)
