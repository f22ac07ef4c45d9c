"""Timing of the PC1 entropy coder at the bench geometry (batch 8 of 32x40x153 symbols)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import __graft_entry__ as g
g.build()
import bench
from dsin_b200 import bitstream, ops, synth

ae = bench.build_ae(0)
pc = ae.pc_imgcomp
centers = torch.from_numpy(ae.ae_imgcomp.centers_host).cuda()
x, y = synth.make_batch(8, 320, 1224, seed=5)
out = ae.reconstruct_device(torch.tensor(x).cuda(), torch.tensor(y).cuda())
sym = out["symbols"]
for ns in (8, 16):
    for _ in range(2):
        b, sizes, status = ops.pc_encode(sym, centers, pc._codec, ns)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    b, sizes, status = ops.pc_encode(sym, centers, pc._codec, ns)
    e1.record()
    back = ops.pc_decode(b, sizes, tuple(sym.shape), centers, pc._codec)
    e2.record()
    torch.cuda.synchronize()
    bits = 8 * int(sizes.sum())
    print("nstreams=%d encode %.2f ms decode %.2f ms (batch 8) | %d payload bits, estimate %.1f | identical=%s" % (
        ns, e0.elapsed_time(e1), e1.elapsed_time(e2), bits, float(out["bits_sum"].sum()), torch.equal(back, sym)), flush=True)
