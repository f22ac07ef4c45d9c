"""Probe: can AE.reconstruct_device be captured in a CUDA graph, and what does it buy at batch 1 / 8?"""
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
    xd, yd = torch.tensor(x).cuda(), torch.tensor(y).cuda()
    for _ in range(3):
        out = ae.reconstruct_device(xd, yd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        out = ae.reconstruct_device(xd, yd)
    torch.cuda.synchronize()
    eager = (time.perf_counter() - t0) / 10
    ref = {k: out[k].clone() for k in ("x_with_si", "bits_sum", "row", "col")}
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            ae.reconstruct_device(xd, yd)
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        gout = ae.reconstruct_device(xd, yd)
    gr.replay()
    torch.cuda.synchronize()
    ok = all(torch.equal(ref[k], gout[k]) for k in ref)
    t0 = time.perf_counter()
    for _ in range(10):
        gr.replay()
    torch.cuda.synchronize()
    graph = (time.perf_counter() - t0) / 10
    print("B=%d eager %.2f ms  graph %.2f ms  identical=%s" % (B, eager * 1e3, graph * 1e3, ok), flush=True)
