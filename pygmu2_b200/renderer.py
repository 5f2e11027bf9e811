"""
Renderers: the block-pull loop that drives the hot path.

``Renderer`` / ``NullRenderer`` mirror the reference's scheduler contract
(renderer.py:226-336 set_source/start/render/stop, :351-421 graph validation,
:423-479 lifecycle walk; null_renderer.py:30-32 discard sink).

``BankRenderer`` is the batched form of that loop: N
independent graphs whose roots are bank-attached PEs are pulled in lockstep, so
one render() is one device call over all N streams instead of N Python pulls.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

from .core import ProcessingElement, Snippet, handle_error, log


class Renderer(ABC):
    def __init__(self, sample_rate: int = 44100):
        self._sample_rate = int(sample_rate)
        self._source = None
        self._started = False
        self._channel_count = None

    sample_rate = property(lambda self: self._sample_rate)
    source = property(lambda self: self._source)
    channel_count = property(lambda self: self._channel_count)
    started = property(lambda self: self._started)

    # -- renderer.py:226-258
    def set_source(self, source: ProcessingElement) -> None:
        if self._started and handle_error("Cannot set source while started. Call stop() first."):
            return
        self._channel_count = self._validate(source, {})
        self._source = source

    # -- renderer.py:260-295
    def start(self) -> None:
        if self._source is None:
            handle_error("No source set. Call set_source() first.", fatal=True)
        if self._started and handle_error("Already started. Call stop() first."):
            return
        self._walk(self._source, set(), bottom_up=True)
        self._started = True

    def stop(self) -> None:
        if not self._started:
            return
        if self._source is not None:
            self._walk(self._source, set(), bottom_up=False)
        self._started = False

    # -- renderer.py:297-327
    def render(self, start: int, duration: int) -> None:
        if self._source is None:
            handle_error("No source set. Call set_source() first.", fatal=True)
        if not self._started:
            handle_error("Not started. Call start() first.", fatal=True)
        if duration < 1:
            handle_error("Renderer.render() requires duration >= 1 to prevent infinite loops.",
                         fatal=True, exception_class=ValueError)
        self._output(self._source.render(start, duration))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.stop()
        return False

    @abstractmethod
    def _output(self, snippet: Snippet) -> None: ...

    # -- renderer.py:351-421: purity (one sink per stateful PE) + channel compatibility
    def _validate(self, pe: ProcessingElement, seen: dict) -> int:
        key = id(pe)
        if key in seen:
            if not pe.is_pure():
                raise ValueError(
                    f"{type(pe).__name__} is not pure but has multiple sinks. "
                    f"Stateful PEs can only connect to one downstream PE."
                )
            return seen[key]
        counts = [self._validate(p, seen) for p in pe.inputs()]
        need = pe.required_input_channels()
        if need is not None:
            for p, got in zip(pe.inputs(), counts):
                if got != need:
                    raise ValueError(f"{type(pe).__name__} requires {need} channel(s), "
                                     f"but {type(p).__name__} outputs {got}")
        out = pe.channel_count()
        if out is None:
            if not counts:
                raise ValueError(f"{type(pe).__name__} has no inputs but channel_count() is None")
            out = pe.resolve_channel_count(counts)
        seen[key] = out
        return out

    # -- renderer.py:423-479: on_start inputs-first, on_stop outputs-first, each PE once
    def _walk(self, pe: ProcessingElement, done: set, bottom_up: bool) -> None:
        if id(pe) in done:
            return
        done.add(id(pe))
        if not bottom_up:
            pe.on_stop()
        for p in pe.inputs():
            self._walk(p, done, bottom_up)
        if bottom_up:
            pe.on_start()


class NullRenderer(Renderer):
    """Discard sink: renders as fast as possible (benchmarks, tests)."""

    def _output(self, snippet: Snippet) -> None:
        pass


class CallbackStop(Exception):
    """Raised by the streaming callback when the requested range has been played (``sounddevice.CallbackStop``
    when PortAudio is the sink; this class for any other ``stream_factory``)."""


class AudioRenderer(Renderer):
    """Blocking playback pull loop (reference audio_renderer.py:91-116,118-181).

    The device PEs are pulled exactly as the reference pulls them -- ``render`` writes one Snippet,
    ``play_extent`` walks the source extent in chunks of ``blocksize * 16`` -- and each float32
    ``(frames, channels)`` array is handed to an output stream with ``start() / write(data) / stop() /
    close()``.  By default that stream is ``sounddevice.OutputStream`` (PortAudio), as in the reference;
    ``stream_factory(samplerate, channels, blocksize)`` substitutes any sink (a file writer, a test
    double).  PortAudio itself is host I/O and out of scope: without sounddevice and without a factory,
    output raises.
    """

    def __init__(self, sample_rate: int = 44100, device=None, blocksize: int = 1024, latency="low",
                 stream_factory=None):
        super().__init__(sample_rate=sample_rate)
        self._device, self._blocksize, self._latency = device, int(blocksize), latency
        self._factory = stream_factory
        self._blocking_stream = None
        self._stream = None            # callback-mode stream (stream_start / stream_stop)
        self._stream_position = 0
        self._stream_end = None

    device = property(lambda self: self._device)
    blocksize = property(lambda self: self._blocksize)

    def _open(self, channels: int, callback=None):
        if self._factory is not None:
            if callback is None:
                return self._factory(self._sample_rate, channels, self._blocksize)
            return self._factory(self._sample_rate, channels, self._blocksize, callback=callback)
        try:
            import sounddevice as sd
        except Exception as exc:  # pragma: no cover - depends on the host
            raise RuntimeError("AudioRenderer needs the sounddevice package (PortAudio) or a stream_factory") from exc
        kw = {"callback": callback} if callback is not None else {}
        return sd.OutputStream(samplerate=self._sample_rate, channels=channels, dtype="float32",
                               device=self._device, blocksize=self._blocksize, latency=self._latency, **kw)

    def _callback_stop(self):
        if self._factory is None:
            try:
                import sounddevice as sd
                return sd.CallbackStop
            except Exception:  # pragma: no cover
                pass
        return CallbackStop

    def _output(self, snippet: Snippet) -> None:
        if self._blocking_stream is None:  # one long-lived stream, opened on the first write
            self._blocking_stream = self._open(snippet.channels)
            self._blocking_stream.start()
        self._blocking_stream.write(snippet.data)

    def play_range(self, start: int, duration: int) -> None:
        self.render(start, duration)

    def play_extent(self, chunk_size: int | None = None) -> None:
        if self._source is None:
            handle_error("No source set. Call set_source() first.", fatal=True)
        ext = self._source.extent()
        if ext.start is None or ext.end is None:
            handle_error("Cannot play_extent() on infinite source. "
                         "Use CropPE to limit the extent, or use play_range().", fatal=True)
        chunk = int(chunk_size) if chunk_size else self._blocksize * 16
        stream = self._open(self._channel_count or 1)
        stream.start()
        try:
            pos = ext.start
            while pos < ext.end:
                n = min(chunk, ext.end - pos)
                stream.write(self._source.render(pos, n).data)
                pos += n
        finally:
            stream.stop()
            stream.close()

    # -- audio_renderer.py:183-262: non-blocking playback, the source is pulled on the sink's callback thread
    def stream_start(self, start: int = 0, end: int | None = None) -> None:
        if not self._started:
            handle_error("Not started. Call start() first.", fatal=True)
        if self._stream is not None:
            handle_error("Already streaming. Call stream_stop() first.", fatal=True)
        if self._source is None:
            handle_error("No source set.", fatal=True)
        self._stream_position, self._stream_end = int(start), end
        stop_exc = self._callback_stop()

        def callback(outdata, frames, time_info, status):
            if status:
                log.warning(f"Stream status: {status}")
            if self._stream_end is not None:
                remaining = self._stream_end - self._stream_position
                if remaining <= 0:
                    outdata.fill(0)
                    raise stop_exc()
                frames = min(frames, remaining)
            snippet = self._source.render(self._stream_position, frames)   # device PEs: any thread may pull
            if snippet.duration < len(outdata):
                outdata[:snippet.duration] = snippet.data
                outdata[snippet.duration:] = 0
            else:
                outdata[:] = snippet.data[:len(outdata)]
            self._stream_position += frames

        self._stream = self._open(self._channel_count or 1, callback=callback)
        self._stream.start()

    def stream_stop(self) -> None:
        if self._stream is not None:
            self._stream.stop()
            self._stream.close()
            self._stream = None

    def stop(self) -> None:
        self.stream_stop()
        if self._blocking_stream is not None:
            self._blocking_stream.stop()
            self._blocking_stream.close()
            self._blocking_stream = None
        super().stop()


class BankRenderer:
    """Batched block-pull loop over a device bank.

    ``bank`` is any object with ``render(start, duration) -> np.ndarray`` covering all
    of its streams at once (``ConvolveBank``, ``HrtfMixBank``); ``sink`` receives each
    result (default: discard, like NullRenderer).  start/stop map onto bank.reset().
    """

    def __init__(self, bank, sink=None, sample_rate: int = 44100):
        self._bank, self._sink = bank, sink
        self._sample_rate = int(sample_rate)
        self._started = False

    def start(self) -> None:
        self._bank.reset()
        self._started = True

    def stop(self) -> None:
        if self._started:
            self._bank.reset()
        self._started = False

    def render(self, start: int, duration: int) -> None:
        if not self._started:
            handle_error("Not started. Call start() first.", fatal=True)
        if duration < 1:
            handle_error("Renderer.render() requires duration >= 1 to prevent infinite loops.",
                         fatal=True, exception_class=ValueError)
        out = self._bank.render(start, duration)
        if self._sink is not None:
            self._sink(out)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.stop()
        return False
