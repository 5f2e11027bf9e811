"""Minimal stand-in for `soundfile` built on the stdlib `wave` module.

Oracle import aid only: gives the Python reference's SpatialHRTF._load_ir
(spatial_pe.py:446-459) the same float32 = int16/32768 samples libsndfile
would return for the PCM16 KEMAR WAVs."""
import wave
import numpy as np

def read(path, dtype="float64", always_2d=False):
    with wave.open(str(path), "rb") as w:
        nch, sw, sr, nfr = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(nfr)
    if sw != 2:
        raise RuntimeError("soundfile stub: only PCM16 supported")
    data = np.frombuffer(raw, dtype="<i2").reshape(-1, nch).astype(np.float64) / 32768.0
    data = data.astype(dtype)
    if nch == 1 and not always_2d:
        data = data[:, 0]
    return data, sr

class SoundFile:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("soundfile stub: SoundFile not supported")

def info(*a, **k):  # pragma: no cover
    raise RuntimeError("soundfile stub: info not supported")

def write(*a, **k):  # pragma: no cover
    raise RuntimeError("soundfile stub: write not supported")
