"""Empty stand-in so `import pygmu2` succeeds in a container without PortAudio.
Test infrastructure only (oracle/): lets gen_golden.py import the Python reference."""
class OutputStream:  # pragma: no cover - never opened
    def __init__(self, *a, **k):
        raise RuntimeError("sounddevice stub: no audio device in this container")
def query_devices(*a, **k):
    return []
