"""wisecondor_b200 - B200-native (sm_100a) implementation of WISECONDOR's within-sample comparison hot path.

Host side mirrors the reference's function interface (wisetools.py / triarray.py); arithmetic runs in
hand-written CUDA kernels behind the C ABI in include/wisecondor_b200.h.
"""
__version__ = "0.1"
