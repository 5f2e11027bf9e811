"""
pygmu2_b200 -- a B200-native (sm_100a) drop-in for one hot path of rdpoor/pygmu2:
ConvolvePE's streaming FFT convolution, the SpatialPE/SpatialHRTF convolution and the
MixPE summation that consume it, behind pygmu2's own ProcessingElement API.

Host side (this package): the PE protocol mirror (``core``), trivial input vehicles
(``sources``), the device-backed PEs and their batched banks, and the renderers.
Device side: ``csrc/libpgx.so`` (hand-written CUDA, C ABI in ``include/pgx.h``), bound
with ctypes in ``_lib``.  There is no CPU fallback: rendering a device-backed PE without
the built library or without a CUDA device raises.
"""
from .core import (ErrorMode, ExtendMode, Extent, ProcessingElement, Snippet, SourcePE,
                   diagnostics_report, enable_diagnostics, get_error_mode, get_sample_rate,
                   handle_error, set_error_mode, set_sample_rate)
from .sources import ArrayPE, CachePE, ConstantPE, CropPE, DelayPE, GainPE, InterpolationMode
from .osc_pe import BlitSawPE, DeviceBlock, OscBank, SinePE, SuperSawPE, VoiceBank
from .renderer import AudioRenderer, BankRenderer, CallbackStop, NullRenderer, Renderer
from .bank import ConvolveBank, choose_block
from .hrtf_bank import HrtfMixBank
from .convolve_pe import ConvolvePE
from .spatial_pe import (SpatialAdapter, SpatialConstantPower, SpatialHRTF, SpatialLinear,
                         SpatialMethod, SpatialPE)
from .mix_pe import MixPE, device_mix_sum
from .reverb_pe import ReverbPE
from .wav_pe import WavReaderPE, WavWriterPE, render_to_file

__all__ = [
    "ErrorMode", "ExtendMode", "Extent", "ProcessingElement", "Snippet", "SourcePE",
    "set_sample_rate", "get_sample_rate", "set_error_mode", "get_error_mode", "handle_error",
    "enable_diagnostics", "diagnostics_report",
    "ArrayPE", "CachePE", "ConstantPE", "CropPE", "DelayPE", "GainPE", "InterpolationMode", "SinePE", "BlitSawPE", "SuperSawPE",
    "OscBank", "VoiceBank", "DeviceBlock",
    "Renderer", "NullRenderer", "AudioRenderer", "BankRenderer", "CallbackStop", "ReverbPE",
    "ConvolveBank", "HrtfMixBank", "choose_block",
    "ConvolvePE", "SpatialPE", "SpatialMethod", "SpatialAdapter", "SpatialLinear",
    "SpatialConstantPower", "SpatialHRTF", "MixPE", "device_mix_sum",
    "WavReaderPE", "WavWriterPE", "render_to_file",
]
__version__ = "0.1.0"
