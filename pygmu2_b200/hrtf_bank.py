"""
HrtfMixBank -- N (moving) sources x per-ear HRTF, summed to one stereo mix on the GPU.

The batched, fused form of ``MixPE(*[SpatialPE(src_i, method=SpatialHRTF(az_i, el_i))])``
(reference call stack SURVEY.md §3.3: spatial_pe.py:465-518 per source, then the MixPE adds
of mix_pe.py:92-94).  The whole HRTF table lives on the device as partition spectra
(2 x 368 filter pairs: entry e as measured, entry 368+e with the ears swapped for negative
azimuth, spatial_pe.py:486-489); a pull re-selects each source's pair from the current
``azimuth`` / ``elevation`` attributes (:446-449), then one fused call does
FFT(source) -> multiply by (H_L, H_R) and accumulate over sources -> one inverse FFT per ear.

Folded in (the graph shape of examples/27_spatial.py:223-234): a per-source integer
``DelayPE`` (the source is simply pulled ``delay`` samples earlier, delay_pe.py:153-160), a per-source constant
``GainPE`` (applied to the source samples; convolution is linear), and sources panned by ``SpatialLinear`` /
``SpatialConstantPower`` with a constant azimuth: a pan law is a 1-tap stereo filter, so it rides in the same
filter table and the same accumulation.
"""
from __future__ import annotations

import numpy as np

from . import kemar
from .bank import ConvolveBank, choose_block
from .core import Extent, handle_error
from .resident import ResidentSources


def _is_pan(m) -> bool:
    from .spatial_pe import _Pan
    return isinstance(m, _Pan)


def _shifted(e: Extent, d: int) -> Extent:
    """Extent of DelayPE(src, d) (delay_pe.py:118-128)."""
    return Extent(None if e.start is None else e.start + d, None if e.end is None else e.end + d)


class HrtfMixBank:
    def __init__(self, sources, methods, *, table: np.ndarray | None = None, table_sample_rate: int | None = None,
                 pull_hint: int | None = 512, block: int | None = None, device: int = 0, delays=None, gains=None):
        self.sources = list(sources)
        self.methods = list(methods)
        n = len(self.sources)
        self.delays = [0] * n if delays is None else [int(d) for d in delays]
        self.gains = [None] * n if gains is None else [None if g is None else np.float32(g) for g in gains]
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
        if None in chans:
            raise ValueError("all sources must declare their channel count")
        # one channel count: the device averages the channels (spatial_pe.py:483); mixed counts: the host does
        self.host_mixdown = len(chans) != 1
        self.c_in = 1 if self.host_mixdown else int(chans.pop())
        self.pan_index = {i: k for k, i in enumerate(i for i, m in enumerate(self.methods) if _is_pan(m))}
        self._pan_loaded = {}
        # methods that can announce a changed direction (SpatialHRTF) let the per-pull re-selection be skipped
        # while nothing moved; anything else (duck-typed methods, pan laws) is re-read on every pull
        self._dirty = [True]
        self._all_watched = all(hasattr(m, "_watchers") for m in self.methods)
        # the directions as arrays: a watched method's setter writes its new value straight into them
        self._az = np.array([float(m.azimuth) for m in self.methods], dtype=np.float64)
        self._el = np.array([float(getattr(m, "elevation", 0.0)) for m in self.methods], dtype=np.float64)
        for i, m in enumerate(self.methods):
            if hasattr(m, "_watchers"):
                m._watchers.append((self._az, self._el, i, self._dirty))
        # [e] as measured, [E + e] ears swapped, [2E] silence: the filter of a source MixPE does not render
        both = np.concatenate([table, table[:, :, ::-1], np.zeros_like(table[:1]),
                               np.zeros((len(self.pan_index),) + table.shape[1:], np.float32)], axis=0)
        B = block or choose_block(table.shape[1], pull_hint)
        self.bank = ConvolveBank(both, len(self.sources), self.c_in, block=B, device=device,
                                 mixdown_input=True, filter_of_stream=self._select())
        self._refresh_pans()
        self._selected = self._select()
        self._pos = None
        self._warned = False
        self._was_active = np.ones(len(self.sources), dtype=bool)
        # extents are static: [lo, hi) of every (delayed) source as arrays, for the per-pull gating
        ext = [_shifted(pe.extent(), d) for pe, d in zip(self.sources, self.delays)]
        self._ext_lo = np.array([-np.inf if e.start is None else e.start for e in ext])
        self._ext_hi = np.array([np.inf if e.end is None else e.end for e in ext])
        self._ext_empty = np.array([e.is_empty() for e in ext], dtype=bool)
        # scalar fast test: a request inside [max lo, min hi) meets every source's extent
        self._lo_max = float(np.max(self._ext_lo)) if not self._ext_empty.any() else np.inf
        self._hi_min = float(np.min(self._ext_hi)) if not self._ext_empty.any() else -np.inf
        self._all_active_last = False
        self._sr_checked = False
        self._traj = None   # (device table handle, n_rows, hop, t0) of set_trajectory
        # plain in-memory sources are uploaded once and stay in HBM (no per-pull host samples)
        self._resident = None
        if ResidentSources.eligible(self.sources, self.delays, self.c_in, self.host_mixdown):
            self._resident = ResidentSources(self.sources, self.delays, self.gains, self.c_in, self.host_mixdown,
                                             device=device)

    def _select(self) -> np.ndarray:
        """Filter-table row of every source for the CURRENT azimuth / elevation attributes (re-resolved on every
        pull like spatial_pe.py:446-449), vectorised over the sources."""
        if self._all_watched and not self._dirty[0] and getattr(self, "_sel_cache", None) is not None:
            return self._sel_cache.copy()
        self._dirty[0] = False
        n = len(self.methods)
        if self._all_watched:
            az = self._az.copy()
        else:
            az = np.fromiter((float(m.azimuth) for m in self.methods), dtype=np.float64, count=n)
        if self.n_entries == len(kemar.KEMAR_HRTF_ENTRIES):
            if self._all_watched:
                el = self._el.copy()
            else:
                el = np.fromiter((float(getattr(m, "elevation", 0.0)) for m in self.methods), dtype=np.float64, count=n)
            last = getattr(self, "_dir_cache", None)   # directions unchanged since the last pull: same rows
            if last is not None and np.array_equal(last[0], az) and np.array_equal(last[1], el):
                e = last[2]
            else:
                e = kemar.nearest_indices(az, el)
                self._dir_cache = (az, el, e)
        else:
            e = np.fromiter((int(getattr(m, "entry", 0)) for m in self.methods), dtype=np.int64, count=n)
        idx = (e + np.where(az < 0, self.n_entries, 0)).astype(np.int32)
        for i, k in self.pan_index.items():
            idx[i] = 2 * self.n_entries + 1 + k
        self._sel_cache = idx.copy()
        return idx

    def _refresh_pans(self) -> None:
        """(Re)load the 1-tap filter of every panned source whose azimuth changed (public attribute, like HRTF's)."""
        for i, k in self.pan_index.items():
            m = self.methods[i]
            gl, gr = m._gains(m._azimuths(0, 1))          # float32, exactly the gains of spatial_pe.py:179-214,250-286
            key = (float(gl[0]), float(gr[0]))
            if self._pan_loaded.get(i) != key:
                h = np.zeros((self.bank.filter_len, 2), dtype=np.float32)
                h[0] = key
                self.bank.load_filter(2 * self.n_entries + 1 + k, h)
                self._pan_loaded[i] = key

    def reset(self) -> None:
        self.bank.reset()
        self._pos = None

    # -- trajectories: moving sources without per-pull host work ---------------------------------------------------
    def set_trajectory(self, azimuth, elevation=None, *, hop: int, start: int = 0) -> None:
        """Directions of every source for a whole run, handed over once (extension; the reference moves a source by
        assigning ``method.azimuth`` between pulls, spatial_pe.py:434-449 -- N attribute writes, N table searches and
        one map upload per pull).  ``azimuth`` is (n_rows, N) degrees, ``elevation`` (n_rows, N), (N,) or None (the
        methods' current elevations); row r applies to the pulls that start in [start + r*hop, start + (r+1)*hop) --
        the same hard switch at pull boundaries -- and the first / last row holds outside the table.  The filter
        choice of every row (nearest KEMAR entry, ears swapped for negative azimuth, spatial_pe.py:395-426,486-489)
        is resolved here, vectorised, and uploaded as one int32 table; a pull then just points the bank at its row
        (pgx_bank_use_filter_map_device).  ``set_trajectory(None)`` returns to the methods' attributes."""
        import ctypes as C

        from . import _lib
        if self._traj is not None:
            self.bank.use_filter_map_device(None)
            _lib.lib().pgx_device_free(self.bank.device, self._traj[0])
            self._traj = None
            self._selected = None
        if azimuth is None:
            return
        if self.pan_index:
            raise ValueError("set_trajectory: every source must be an HRTF source (pan laws keep their attribute)")
        n = len(self.methods)
        az = np.asarray(azimuth, dtype=np.float64)
        if az.ndim != 2 or az.shape[1] != n:
            raise ValueError(f"azimuth must be (n_rows, {n}), got {az.shape}")
        if elevation is None:
            el = np.fromiter((float(getattr(m, "elevation", 0.0)) for m in self.methods), dtype=np.float64, count=n)
        else:
            el = np.asarray(elevation, dtype=np.float64)
        el = np.broadcast_to(el, az.shape)
        if self.n_entries != len(kemar.KEMAR_HRTF_ENTRIES):
            raise ValueError("set_trajectory needs the KEMAR table (directions are resolved against its grid)")
        e = kemar.nearest_indices(az.reshape(-1), el.reshape(-1)).reshape(az.shape)
        table = np.ascontiguousarray(e + np.where(az < 0, self.n_entries, 0), dtype=np.int32)
        ptr = C.c_void_p()
        _lib.check(_lib.lib().pgx_device_alloc(self.bank.device, table.nbytes, C.byref(ptr)))
        _lib.check(_lib.lib().pgx_device_upload(self.bank.device, ptr, table.ctypes.data, table.nbytes))
        self._traj = (ptr, int(table.shape[0]), int(hop), int(start))
        self._traj_host = table
        self._traj_row = None

    def close(self) -> None:
        """Release the device state (bank, resident sources) and stop watching the methods."""
        if self._traj is not None:
            try:
                self.set_trajectory(None)
            except Exception:
                pass
        for m in self.methods:
            w = getattr(m, "_watchers", None)
            if w is not None:
                w[:] = [e for e in w if e[3] is not self._dirty]   # by identity: other banks keep theirs
        if self._resident is not None:
            self._resident.close()
            self._resident = None
        self.bank.close()

    def render(self, start: int, duration: int) -> np.ndarray:
        """One lockstep pull of every source -> (2, duration) float32 stereo mix."""
        if not self._sr_checked:
            sr = self.sources[0].sample_rate
            has_hrtf = len(self.pan_index) < len(self.methods)
            if has_hrtf and self.table_sample_rate is not None and sr != self.table_sample_rate and not self._warned:
                handle_error(
                    f"SpatialHRTF: IR sample rate is {self.table_sample_rate} Hz but source is {sr} Hz. "
                    "Proceeding without resampling.",
                    fatal=False,
                )
                self._warned = True
            self._sr_checked = True
        contiguous = self._pos is not None and start == self._pos
        # MixPE renders only the inputs whose extent meets the request (mix_pe.py:81-85; a SpatialPE's extent is
        # its source's, spatial_pe.py:640-642).  A skipped source contributes nothing - the silent filter - and
        # its next render starts a new run (spatial_pe.py:461-463): its history is cleared when it comes back.
        all_active = start >= self._lo_max and start + duration <= self._hi_min   # scalar test first: the common case
        if all_active and self._all_active_last and contiguous and self._traj is None \
                and self._all_watched and not self._dirty[0] and not self.pan_index and self._selected is not None:
            active = self._was_active    # nothing moved, nothing came or went: the resident selection stands
        else:
            if all_active:
                active = np.ones(len(self.sources), dtype=bool)
            else:
                active = ~self._ext_empty & (self._ext_lo < start + duration) & (self._ext_hi > start)   # Extent.intersects
            if not contiguous:
                self.bank.reset()
            else:
                back = np.flatnonzero(active & ~self._was_active)
                if back.size:
                    self.bank.reset(back)
            if self._traj is not None and all_active:
                ptr, n_rows, hop, t0 = self._traj
                row = min(max((start - t0) // hop, 0), n_rows - 1)
                if row != self._traj_row or self._selected is not None:
                    self.bank.use_filter_map_device(ptr.value + row * len(self.sources) * 4)
                    self._traj_row, self._selected = row, None
            else:
                if self._traj is not None:      # a source is outside its extent: this pull's row, silenced where needed
                    ptr, n_rows, hop, t0 = self._traj
                    sel = self._traj_host[min(max((start - t0) // hop, 0), n_rows - 1)].copy()
                    self._traj_row = None
                else:
                    self._refresh_pans()
                    sel = self._select()
                sel[~active] = 2 * self.n_entries
                if self._selected is None or not np.array_equal(sel, self._selected):
                    if self._traj is not None:
                        self.bank.use_filter_map_device(None)
                    self.bank.set_filter_map(sel)
                    self._selected = sel
            self._was_active = active
            self._all_active_last = all_active
        if self._resident is not None:  # sources live in HBM: a pull is a pointer into the resident buffer
            self._pos = start + duration
            outs, pos = [], 0
            while pos < duration:
                d = min(self.bank.max_pull, self._resident.max_pull, duration - pos)
                outs.append(self.bank.process_device_block(self._resident.device_block(start + pos, d), mix=True))
                pos += d
            return outs[0] if len(outs) == 1 else np.concatenate(outs, axis=-1)
        x = np.zeros((len(self.sources), self.c_in, duration), dtype=np.float32)
        for s, pe in enumerate(self.sources):
            if active[s]:
                data = pe.render(start - self.delays[s], duration).data
                if self.host_mixdown:
                    data = np.mean(data, axis=1, keepdims=True).astype(np.float32)   # spatial_pe.py:483
                if self.gains[s] is not None:
                    data = data * self.gains[s]                                    # gain_pe.py:123-125, float32
                x[s] = data.T
        self._pos = start + duration
        self._was_active = active
        return self.bank.process_mix(x)
