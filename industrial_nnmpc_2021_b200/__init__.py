"""B200-native implementation of the linear-MPC / structured-NN hot path of
pratyushkumar211/industrial_nnmpc_2021.

Host code is Python and mirrors the reference's call signatures; all solves run in
hand-written sm_100a CUDA kernels behind the C ABI declared in ``include/nnmpc.h``
(``csrc/libnnmpc.so``, loaded with ctypes).  There is no CPU fallback: constructing any
solver object without the built library or without a CUDA device raises.
"""
__version__ = "0.1.0"
