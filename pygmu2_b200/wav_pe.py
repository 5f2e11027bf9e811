"""
WavReaderPE / WavWriterPE / render_to_file -- the file steps either side of the path (SURVEY.md §8f rank 4;
reference wav_reader_pe.py:18-157, wav_writer_pe.py:18-171, utils.py:34-62), with PCM16 <-> float32 staged on
the device.

The reference does its file I/O through ``soundfile`` (libsndfile), which is not in this image; files are read
and written here with the stdlib ``wave`` module (PCM_16 only), and the sample conversions libsndfile would do
are restated:

* read : float32 = int16 / 32768  (libsndfile pcm.c ``s2f_array``; what ``sf.read(dtype="float32")`` returns)
* write: int16 = clip(rint(x * 32768), -32768, 32767), round-half-even  (libsndfile pcm.c ``f2s_clip_array``;
  python-soundfile switches clipping on for every file it opens)

Device staging: a ``WavReaderPE`` feeding a ``ConvolvePE`` hands over raw int16 frames (``render_pcm16``) and
the bank converts them in HBM (``PGX_PULL_X_PCM16``: half the H2D bytes); ``render_to_file`` over a device PE
pulls int16 straight from the device (``PGX_PULL_Y_PCM16``: half the D2H bytes, no float32 array on the host).
Parity for this row is **unpinned** against the reference (soundfile absent): it rests on the formulas above and
on round-trip tests (tests/test_gpu_parity.py::test_wav_*).
"""
from __future__ import annotations

import wave

import numpy as np

from .core import Extent, ProcessingElement, Snippet, get_sample_rate, handle_error
from .renderer import NullRenderer


def f32_to_pcm16(x: np.ndarray) -> np.ndarray:
    """libsndfile's float -> PCM_16 with clipping (host form of csrc/k_osc.cu:k_f32_to_pcm16)."""
    scaled = np.asarray(x, dtype=np.float32) * np.float32(32768.0)
    return np.clip(np.rint(scaled), -32768.0, 32767.0).astype(np.int16)


def pcm16_to_f32(p: np.ndarray) -> np.ndarray:
    return np.asarray(p, dtype=np.int16).astype(np.float32) * np.float32(1.0 / 32768.0)


class WavReaderPE(ProcessingElement):
    """Reads a PCM_16 WAV file; zero outside [0, frames) (wav_reader_pe.py:99-144).  Pure."""

    def __init__(self, path: str):
        self._path = str(path)
        self._frame_count = None
        self._channels = None
        self._file_sample_rate = None

    path = property(lambda self: self._path)

    def _ensure_file_info(self) -> None:
        if self._frame_count is None:
            with wave.open(self._path, "rb") as w:
                if w.getsampwidth() != 2 or w.getcomptype() != "NONE":
                    raise ValueError(f"WavReaderPE: only PCM_16 WAV files are supported here ({self._path})")
                self._frame_count, self._channels = w.getnframes(), w.getnchannels()
                self._file_sample_rate = w.getframerate()

    @property
    def file_sample_rate(self):
        self._ensure_file_info()
        return self._file_sample_rate

    @property
    def sample_rate(self):
        if self._sample_rate is not None:
            return self._sample_rate
        return self.file_sample_rate

    def inputs(self) -> list:
        return []

    def is_pure(self) -> bool:
        return True

    def channel_count(self) -> int:
        self._ensure_file_info()
        return self._channels

    def _compute_extent(self) -> Extent:
        self._ensure_file_info()
        return Extent(0, self._frame_count)

    def render_pcm16(self, start: int, duration: int) -> np.ndarray:
        """The raw frames, (duration, channels) int16, zero outside the file: what the device converts."""
        self._ensure_file_info()
        data = np.zeros((duration, self._channels), dtype=np.int16)
        lo, hi = max(start, 0), min(start + duration, self._frame_count)
        if lo < hi:
            with wave.open(self._path, "rb") as w:     # stateless read, like sf.read(start=, stop=)
                w.setpos(lo)
                raw = w.readframes(hi - lo)
            data[lo - start:hi - start] = np.frombuffer(raw, dtype="<i2").reshape(-1, self._channels)
        return data

    def _render(self, start: int, duration: int) -> Snippet:
        return Snippet(start, pcm16_to_f32(self.render_pcm16(start, duration)))

    def __repr__(self):
        return f"WavReaderPE(path={self._path!r})"


class WavWriterPE(ProcessingElement):
    """Writes what it renders to a PCM_16 WAV file and passes it through (wav_writer_pe.py:18-171).

    ``passthrough=False`` (extension, used by ``render_to_file``): when the source is a device PE offering
    ``render_pcm16_out`` the frames come off the device as int16 and the returned Snippet is the quantised
    audio (int16/32768) instead of the float32 original."""

    def __init__(self, source: ProcessingElement, path: str, sample_rate: int | None = None,
                 subtype: str = "PCM_16", *, passthrough: bool = True):
        if subtype != "PCM_16":
            raise NotImplementedError("pygmu2_b200.WavWriterPE writes PCM_16 only (no libsndfile in this image)")
        self._source, self._path = source, str(path)
        self._output_sample_rate, self._subtype = sample_rate, subtype
        self._passthrough = bool(passthrough)
        self._file = None
        self._frames_written = 0

    path = property(lambda self: self._path)
    frames_written = property(lambda self: self._frames_written)

    def inputs(self) -> list:
        return [self._source]

    def is_pure(self) -> bool:
        return False

    def channel_count(self):
        return self._source.channel_count()

    def _compute_extent(self) -> Extent:
        return self._source.extent()

    def _on_start(self) -> None:
        rate = self._output_sample_rate or self.sample_rate
        channels = self._source.channel_count()
        if channels is None and self._source.inputs():
            channels = self._source.inputs()[0].channel_count()
        if channels is None:
            handle_error(f"Cannot determine channel count for WavWriterPE. Source "
                         f"{self._source.__class__.__name__} returns None for channel_count().", fatal=True)
        self._file = wave.open(self._path, "wb")
        self._file.setnchannels(int(channels))
        self._file.setsampwidth(2)
        self._file.setframerate(int(rate))
        self._frames_written = 0

    def _on_stop(self) -> None:
        if self._file is not None:
            self._file.close()
            self._file = None

    def _render(self, start: int, duration: int) -> Snippet:
        dev = None if self._passthrough else getattr(self._source, "render_pcm16_out", None)
        if dev is not None:
            pcm = dev(start, duration)                          # (n, C) int16 converted on the device
            snippet = Snippet(start, pcm16_to_f32(pcm))
        else:
            snippet = self._source.render(start, duration)
            pcm = f32_to_pcm16(snippet.data)
        if self._file is not None:
            self._file.writeframes(np.ascontiguousarray(pcm).astype("<i2", copy=False).tobytes())
            self._frames_written += snippet.duration
        return snippet

    def __repr__(self):
        return (f"WavWriterPE(source={self._source.__class__.__name__}, path={self._path!r}, "
                f"subtype={self._subtype!r})")


def render_to_file(source: ProcessingElement, out_path: str, *, sample_rate: int | None = None, extent=None,
                   chunk: int = 65536) -> None:
    """Render a finite PE to a PCM_16 WAV as fast as possible (utils.py:34-62).  The reference issues ONE
    render() for the whole extent; here the extent is walked in ``chunk``-sample pulls so device staging stays
    bounded, and the frames leave the device as int16 when the source supports it."""
    sr = int(sample_rate) if sample_rate is not None else get_sample_rate()
    if sr is None:
        raise RuntimeError("Sample rate not set. Call pg.set_sample_rate() or pass sample_rate.")
    if extent is None:
        extent = source.extent()
    if extent.start is None or extent.end is None:
        raise RuntimeError("Cannot render to file: source has infinite extent.")
    writer = WavWriterPE(source, out_path, sample_rate=sr, passthrough=False)
    renderer = NullRenderer(sample_rate=sr)
    renderer.set_source(writer)
    with renderer:
        renderer.start()
        pos = extent.start
        while pos < extent.end:
            n = min(int(chunk), extent.end - pos)
            renderer.render(pos, n)
            pos += n
