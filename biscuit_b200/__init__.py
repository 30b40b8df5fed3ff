"""biscuit_b200: B200-native implementation of BISCUIT's align / pileup hot paths.

The product is the C ABI in include/bsq.h (biscuit_b200/csrc/libbsq.so, CUDA sm_100a) and the C host
programs in biscuit_b200/host; this Python package is a thin ctypes mirror used by tests and bench.py.
"""
