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
