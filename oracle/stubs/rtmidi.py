"""Empty stand-in for `rtmidi` (oracle import aid only)."""
