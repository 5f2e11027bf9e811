"""
KEMAR compact HRTF set: geometry, nearest-neighbour lookup and IR table loading.

The reference keeps a literal 368-entry (elev, az, filename) table
(src/pygmu2/spatial_pe.py:318-393) and loads one stereo PCM16 WAV per lookup
(:446-459).  Here the table is generated from the published measurement grid
(Gardner & Martin, MIT Media Lab TR #280, 1994: elevations -40..90 in steps of 10
degrees, equally spaced azimuths per elevation rounded to whole degrees, right
hemisphere) -- ``oracle/gen_golden.py`` asserts it equals the reference's literal
table -- and the 368 IR pairs come from one packed asset
(assets/kemar_compact_i16.npz, int16 exactly as in the WAVs; float32 = int16/32768
is what libsndfile returns for dtype='float32').
"""
from __future__ import annotations

import os

import numpy as np

_AZ_COUNTS = {-40: 56, -30: 60, -20: 72, -10: 72, 0: 72, 10: 72, 20: 72,
              30: 60, 40: 56, 50: 45, 60: 36, 70: 24, 80: 12, 90: 1}
KEMAR_SAMPLE_RATE = 44_100


def _entries():
    out = []
    for elev in sorted(_AZ_COUNTS):
        n_az = _AZ_COUNTS[elev]
        for i in range(n_az):
            az = int(np.floor(i * 360.0 / n_az + 0.5))
            if az > 180:
                break
            out.append((elev, az, f"H{elev}e{az:03d}a.wav"))
    return tuple(out)


KEMAR_HRTF_ENTRIES = _entries()
_ELEV = np.array([e[0] for e in KEMAR_HRTF_ENTRIES], dtype=np.float64)
_AZ = np.array([e[1] for e in KEMAR_HRTF_ENTRIES], dtype=np.float64)


def nearest_index(azimuth: float, elevation: float) -> int:
    """Index of the entry SpatialHRTF.hrtf_filename_for picks (spatial_pe.py:395-426):
    az := min(180, |az|); least squared distance in (elev, az); first minimum in table order."""
    az = min(180.0, abs(float(azimuth)))
    d = (_ELEV - float(elevation)) ** 2 + (_AZ - az) ** 2
    return int(np.argmin(d))  # argmin returns the first minimum, like python's min()


def nearest_indices(azimuth: np.ndarray, elevation: np.ndarray) -> np.ndarray:
    """``nearest_index`` for arrays of directions at once (same arithmetic, same first-minimum rule)."""
    az = np.minimum(180.0, np.abs(np.asarray(azimuth, dtype=np.float64)))[:, None]
    el = np.asarray(elevation, dtype=np.float64)[:, None]
    d = (_ELEV[None, :] - el) ** 2 + (_AZ[None, :] - az) ** 2
    return np.argmin(d, axis=1)


_table_cache = {}


def load_table() -> tuple[np.ndarray, int]:
    """(368, 128, 2) float32 IR pairs in table order, and their sample rate."""
    if "t" not in _table_cache:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "kemar_compact_i16.npz")
        z = np.load(path)
        _table_cache["t"] = (z["ir_i16"].astype(np.float32) / np.float32(32768.0), int(z["sample_rate"]))
    return _table_cache["t"]
