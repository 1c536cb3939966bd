"""avex_b200 -- B200-native (sm_100a) kernels for avex's embedding hot path, behind avex's plugin surface.

waveform -> Kaldi fbank -> BEATs encoder, as hand-written CUDA reached through the C ABI in include/avexk.h.
No CPU / eager fallback: operations raise when libavexk.so or a CUDA device is missing.
"""
__version__ = "0.1.0"
