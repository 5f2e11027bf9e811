"""Empty stand-in for `mido` (MIDI I/O is out of scope; oracle import aid only)."""
def get_input_names():
    return []
def open_input(*a, **k):
    raise RuntimeError("mido stub")
class Message:  # pragma: no cover
    pass
