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


def _rings():
    """The table is 14 elevation rings in ascending order, azimuths ascending within a ring."""
    rings, i = [], 0
    while i < len(KEMAR_HRTF_ENTRIES):
        j = i
        while j < len(KEMAR_HRTF_ENTRIES) and KEMAR_HRTF_ENTRIES[j][0] == KEMAR_HRTF_ENTRIES[i][0]:
            j += 1
        rings.append((float(KEMAR_HRTF_ENTRIES[i][0]), _AZ[i:j].copy(), i))
        i = j
    return rings


_RINGS = _rings()
_R_ELEV = np.array([r[0] for r in _RINGS])[:, None]                          # (14, 1)
_R_LEN = np.array([r[1].shape[0] for r in _RINGS])[:, None, None]            # (14, 1, 1)
_R_FIRST = np.array([r[2] for r in _RINGS])[:, None]                         # (14, 1)
_R_STEP = np.array([360.0 / _AZ_COUNTS[int(r[0])] for r in _RINGS])[:, None]  # (14, 1) nominal azimuth spacing
_R_AZ = np.stack([np.pad(r[1], (0, 37 - r[1].shape[0]), mode="edge") for r in _RINGS])   # (14, 37)
_R_ROW = np.arange(len(_RINGS))[:, None, None]


def nearest_indices(azimuth: np.ndarray, elevation: np.ndarray) -> np.ndarray:
    """``nearest_index`` for arrays of directions at once, in the C library (``pgx_nearest_direction``: the plain
    368-entry scan per direction in float64, first minimum in table order) -- the host cost of re-selecting the
    filters of many moving sources before a pull.  ``nearest_indices_numpy`` is the vectorised numpy form of the
    same search (4 candidates per elevation ring), kept as its cross-check."""
    from . import _lib
    az = np.ascontiguousarray(azimuth, dtype=np.float64).reshape(-1)
    el = np.ascontiguousarray(np.broadcast_to(np.asarray(elevation, dtype=np.float64), az.shape)).reshape(-1)
    out = np.empty(az.shape[0], dtype=np.int32)
    _lib.check(_lib.lib().pgx_nearest_direction(_ELEV.ctypes.data, _AZ.ctypes.data, int(_ELEV.shape[0]),
                                                az.ctypes.data, el.ctypes.data, int(az.shape[0]), out.ctypes.data))
    return out.astype(np.int64)


def nearest_indices_numpy(azimuth: np.ndarray, elevation: np.ndarray) -> np.ndarray:
    """``nearest_index`` for arrays of directions at once: same float64 arithmetic, same first-minimum rule.
    On each of the 14 elevation rings only the entries around az/step can be nearest (ring azimuths are
    round(i*step)), so 4 candidates per ring are examined instead of the whole ring: 56 distances per query
    instead of 368, all rings in one vectorised pass."""
    az = np.minimum(180.0, np.abs(np.asarray(azimuth, dtype=np.float64)))[None, :]   # (1, Q)
    el = np.asarray(elevation, dtype=np.float64)[None, :]
    base = np.floor(az / _R_STEP).astype(np.int64)[:, :, None]                       # (14, Q, 1)
    cand = np.clip(base + np.arange(-1, 3)[None, None, :], 0, _R_LEN - 1)            # (14, Q, 4) ascending indices
    d = ((_R_ELEV - el) ** 2)[:, :, None] + (_R_AZ[_R_ROW, cand] - az[:, :, None]) ** 2
    k = np.argmin(d, axis=2)                                                         # first minimum: lowest index
    q = np.arange(az.shape[1])[None, :]
    ring_best_d = d[_R_ROW[:, :, 0], q, k]                                           # (14, Q)
    ring_best_i = _R_FIRST + cand[_R_ROW[:, :, 0], q, k]
    ring = np.argmin(ring_best_d, axis=0)                                            # lowest ring on ties
    return ring_best_i[ring, q[0]]


_table_cache = {}


def load_table() -> tuple[np.ndarray, int]:
    """(368, 128, 2) float32 IR pairs in table order, and their sample rate."""
    if "t" not in _table_cache:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "kemar_compact_i16.npz")
        z = np.load(path)
        _table_cache["t"] = (z["ir_i16"].astype(np.float32) / np.float32(32768.0), int(z["sample_rate"]))
    return _table_cache["t"]
