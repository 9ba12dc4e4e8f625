"""
shgyield_b200 -- B200-native (sm_100a CUDA, FP64) implementation of the SSHG yield path of
roguephysicist/SHGYield (`shgyield/shg.py`).  The CUDA library `libsshg.so` (C ABI declared in
include/sshg.h) is loaded lazily by `shgyield_b200._lib`; there is no CPU fallback.
"""
__version__ = "0.1.0"
