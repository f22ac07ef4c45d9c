"""How many positions per patch really come within DELTA = 2e-3 of the best masked Pearson score?  (CPU oracle on the
sif-bench inputs and on decoded-like noisy inputs; decides the SI-Finder's candidate bookkeeping.)"""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsin_b200 import synth
from oracle import dsin_oracle as O
import torch.nn.functional as F

def stats(x, y_dec, H, W, npatch=60):
    xi = torch.tensor(x).permute(1, 2, 0); yi = torch.tensor(y_dec).permute(1, 2, 0)
    q = O.rgb_transform(O.sif_normalize_nhwc(O.extract_patches(xi, 20, 24)))
    r = O.rgb_transform(O.sif_normalize_nhwc(yi))
    mask = O.gaussian_masks(H, W, 20, 24)
    dt = q.dtype; P = q.shape[0]; n = 1440.0
    rr = r.permute(2, 0, 1).unsqueeze(0)
    ones = torch.ones(1, 3, 20, 24)
    sum_y = F.conv2d(rr, ones)[0, 0]; sum_y2 = F.conv2d(rr * rr, ones)[0, 0]; y_mean = sum_y / n
    den_y = sum_y2 - 2 * (y_mean * sum_y) + n * (y_mean * y_mean)
    sel = np.linspace(0, P - 1, npatch).astype(int)
    qf = q.reshape(P, -1)
    out = []
    for p in sel:
        filt = q[p].permute(2, 0, 1).unsqueeze(0)
        xy = F.conv2d(rr, filt)[0, 0]
        sx, sx2, xm = qf[p].sum(), (qf[p] ** 2).sum(), qf[p].mean()
        den_x = sx2 - 2 * xm * sx + n * xm * xm
        num = xy - y_mean * sx - sum_y * xm + n * (y_mean * xm)
        s = num / torch.sqrt(den_y * den_x) * torch.tensor(mask[p])
        best = float(s.max())
        within = (s >= best - 2e-3)
        cnt = int(within.sum())
        # per 4-row x 128-col-alternating group: max number within DELTA in one group
        rows, cols = torch.nonzero(within, as_tuple=True)
        grp = (rows // 4) * 2 + ((cols % 256) // 128)
        per_group = int(torch.bincount(grp).max()) if cnt else 0
        out.append((cnt, per_group, best))
    a = np.array(out)
    print("   positions within 2e-3 of the best: median %d, p90 %d, max %d | most in one group: median %d, p90 %d, max %d | best score median %.3f"
          % (np.median(a[:, 0]), np.percentile(a[:, 0], 90), a[:, 0].max(), np.median(a[:, 1]), np.percentile(a[:, 1], 90), a[:, 1].max(), np.median(a[:, 2])))

torch.set_num_threads(8)
H, W = 320, 1224
x, y = synth.make_batch(2, H, W, seed=77)
for n in range(2):
    print("sif-bench style input %d (x as x_dec, y as y_dec):" % n); stats(x[n], y[n], H, W)
Wt = synth.make_weights(0, residual_gamma=0.25)
ref = O.reconstruct(x[:1], y[:1], Wt)
print("decoded images of the random-init network (what the full bench feeds the SI-Finder):"); stats(ref.x_dec[0].numpy(), ref.y_dec[0].numpy(), H, W)
