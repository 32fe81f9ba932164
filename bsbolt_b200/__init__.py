"""bsbolt_b200 -- B200-native drop-in for the `bsbolt Align` hot path.

Host side mirrors the reference's Python interface (bsbolt/Align/AlignReads.py, bsbolt/Utils/Launcher.py,
bsbolt/Utils/Parser.py); the alignment itself runs in hand-written sm_100a kernels behind the C ABI of
libbsbolt_b200.so (include/bsbolt_b200.h). There is no CPU fallback: importing works anywhere, aligning
requires the built library and a CUDA device.
"""
__version__ = '0.1.0'
