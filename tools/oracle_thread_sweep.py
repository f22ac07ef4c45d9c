import sys, time, os
sys.path.insert(0, '.')
import torch, numpy as np
from dsin_b200 import synth
from oracle import dsin_oracle as O
W = synth.make_weights(0, residual_gamma=0.25)
x, y = synth.make_batch(1, 320, 1224, seed=1000)
for t in (16, 32, 64):
    torch.set_num_threads(t)
    t0 = time.perf_counter(); enc = O.encode(torch.tensor(x), W); t1 = time.perf_counter()
    xd = O.decode(enc.qbar, W); t2 = time.perf_counter()
    O.si_full_img(xd, torch.tensor(y), xd); t3 = time.perf_counter()
    print("threads", t, "enc %.1f dec %.1f sif %.1f" % (t1 - t0, t2 - t1, t3 - t2), flush=True)
