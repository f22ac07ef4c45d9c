"""Batch-1 latency of the public numpy call (the reference's own batch size), CUDA graphs on / off."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import __graft_entry__ as g
g.build()
import bench
from dsin_b200 import synth

ae = bench.build_ae(0)
for B in (1, 8):
    x, y = synth.make_batch(B, 320, 1224, seed=5)
    x, y = x.astype(np.uint8), y.astype(np.uint8)
    for mode in (True, False):
        ae.use_cuda_graph = mode
        for _ in range(3):
            ae.siNet_get_reconstructed(x, y)
        t0 = time.perf_counter()
        for _ in range(10):
            ae.siNet_get_reconstructed(x, y)
        dt = (time.perf_counter() - t0) / 10
        print("B=%d graph=%s  %.2f ms per call  %.1f Mpix/s" % (B, mode, dt * 1e3, B * 0.39168 / dt), flush=True)
