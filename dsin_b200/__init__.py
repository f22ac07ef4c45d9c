"""dsin_b200 -- B200-native (sm_100a) inference hot path of DSIN.

Host side mirrors the reference's AE / encoder / decoder / SI_full_img / siFinder / siNet
entry points; the arithmetic runs in hand-written CUDA kernels behind the C ABI declared in
include/dsin_b200.h (libdsin_b200.so, built in-tree by __graft_entry__.build()).
"""
__version__ = "0.1.0"
