"""lossyless_b200: the encode / rate-distortion hot path of YannDubs/lossyless on B200 (sm_100a).

Python host over liblossyless_b200.so (C ABI in include/lossyless_b200.h).  Importing the
package does not need a GPU; calling into the kernels does, and there is no CPU fallback.
"""
__version__ = "0.1.0"
