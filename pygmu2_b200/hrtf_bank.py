"""
HrtfMixBank -- N (moving) sources x per-ear HRTF, summed to one stereo mix on the GPU.

The batched, fused form of ``MixPE(*[SpatialPE(src_i, method=SpatialHRTF(az_i, el_i))])``
(reference call stack SURVEY.md §3.3: spatial_pe.py:465-518 per source, then the MixPE adds
of mix_pe.py:92-94).  The whole HRTF table lives on the device as partition spectra
(2 x 368 filter pairs: entry e as measured, entry 368+e with the ears swapped for negative
azimuth, spatial_pe.py:486-489); a pull re-selects each source's pair from the current
``azimuth`` / ``elevation`` attributes (:446-449), then one fused call does
FFT(source) -> multiply by (H_L, H_R) and accumulate over sources -> one inverse FFT per ear.
"""
from __future__ import annotations

import numpy as np

from . import kemar
from .bank import ConvolveBank, choose_block
from .core import handle_error


class HrtfMixBank:
    def __init__(self, sources, methods, *, table: np.ndarray | None = None, table_sample_rate: int | None = None,
                 pull_hint: int | None = 512, block: int | None = None, device: int = 0):
        self.sources = list(sources)
        self.methods = list(methods)
        if len(self.sources) != len(self.methods) or not self.sources:
            raise ValueError("need one SpatialHRTF-like method (azimuth, elevation) per source")
        if table is None:
            table, table_sample_rate = kemar.load_table()
        table = np.asarray(table, dtype=np.float32)
        if table.ndim != 3 or table.shape[2] != 2:
            raise ValueError(f"SpatialHRTF: expected stereo IR table (E, taps, 2), got shape {table.shape}")
        self.n_entries = table.shape[0]
        self.table_sample_rate = table_sample_rate
        chans = {p.channel_count() for p in self.sources}
        if len(chans) != 1 or None in chans:
            raise ValueError("all sources must declare the same channel count")
        self.c_in = int(chans.pop())
        both = np.concatenate([table, table[:, :, ::-1]], axis=0)  # [e] as measured, [E + e] ears swapped
        B = block or choose_block(table.shape[1], pull_hint)
        self.bank = ConvolveBank(both, len(self.sources), self.c_in, block=B, device=device,
                                 mixdown_input=True, filter_of_stream=self._select())
        self._selected = self._select()
        self._pos = None
        self._warned = False

    def _select(self) -> np.ndarray:
        idx = np.empty(len(self.methods), dtype=np.int32)
        for i, m in enumerate(self.methods):
            e = kemar.nearest_index(m.azimuth, m.elevation) if self.n_entries == len(kemar.KEMAR_HRTF_ENTRIES) \
                else int(getattr(m, "entry", 0))
            idx[i] = e + (self.n_entries if m.azimuth < 0 else 0)
        return idx

    def reset(self) -> None:
        self.bank.reset()
        self._pos = None

    def render(self, start: int, duration: int) -> np.ndarray:
        """One lockstep pull of every source -> (2, duration) float32 stereo mix."""
        sr = self.sources[0].sample_rate
        if self.table_sample_rate is not None and sr != self.table_sample_rate and not self._warned:
            handle_error(
                f"SpatialHRTF: IR sample rate is {self.table_sample_rate} Hz but source is {sr} Hz. "
                "Proceeding without resampling.",
                fatal=False,
            )
            self._warned = True
        sel = self._select()
        if not np.array_equal(sel, self._selected):
            self.bank.set_filter_map(sel)
            self._selected = sel
        if self._pos is None or start != self._pos:
            self.bank.reset()
        x = np.empty((len(self.sources), self.c_in, duration), dtype=np.float32)
        for s, pe in enumerate(self.sources):
            x[s] = pe.render(start, duration).data.T
        self._pos = start + duration
        return self.bank.process_mix(x)
