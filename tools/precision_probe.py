"""Which passes need fp32-class tensor-core arithmetic?  (GPU; compares with the cached oracle cases.)

For each precision policy (dsin_b200/precision.py) on the 320x1224 case:
  * encoder: |z - z64| (rms, max) of the GPU and of the fp32 oracle, symbol mismatches vs fp32 / fp64 oracle and the
    float64 margins of the mismatches
  * full path: bpp, (row, col) agreement, MS-SSIM(x, x_with_si) vs the oracle's, max |d x_dec|, |d x_with_si|
  * device time per step at batch 8 (CUDA events, graph replays)
Writes one JSON document to stdout / --out.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_cache  # noqa: E402
from parity_utils import make_ae  # noqa: E402
from dsin_b200 import ops, precision, synth  # noqa: E402
from oracle import ms_ssim_oracle as M  # noqa: E402


def enc_stats(z_gpu, sym_gpu, d, tag):
    if "z64_" + tag not in d:  # light case: symbols and margins only
        m64 = torch.as_tensor(d["margin64_" + tag])
        s32 = torch.as_tensor(d["sym32_" + tag]).long()
        mm32 = sym_gpu.cpu() != s32
        return {"symbols": int(s32.numel()), "mismatch_vs_fp32": int(mm32.sum()),
                "max_margin64_of_mismatch": float(m64[mm32].max()) if int(mm32.sum()) else 0.0,
                "margins64_of_mismatches_vs_fp32": sorted(float(v) for v in m64[mm32].tolist())[-12:]}
    z64 = torch.as_tensor(d["z64_" + tag])
    z32 = torch.as_tensor(d["z32_" + tag]).double()
    zg = z_gpu.cpu().double()
    m64 = torch.as_tensor(d["margin64_" + tag])
    s32, s64 = torch.as_tensor(d["sym32_" + tag]).long(), torch.as_tensor(d["sym64_" + tag]).long()
    sg = sym_gpu.cpu()
    eg, e32 = (zg - z64).abs(), (z32 - z64).abs()
    mm32, mm64 = sg != s32, sg != s64
    return {
        "z_err_gpu_rms": float(eg.pow(2).mean().sqrt()), "z_err_gpu_max": float(eg.max()),
        "z_err_fp32oracle_rms": float(e32.pow(2).mean().sqrt()), "z_err_fp32oracle_max": float(e32.max()),
        "signed_bias_gpu": float((zg - z64).mean()), "signed_bias_rel_to_abs": float(((zg - z64) * z64.sign()).mean()),
        "symbols": int(sg.numel()),
        "mismatch_vs_fp32": int(mm32.sum()), "mismatch_vs_fp64": int(mm64.sum()),
        "fp32_vs_fp64": int((s32 != s64).sum()),
        "margins64_of_mismatches_vs_fp32": sorted(float(v) for v in m64[mm32].tolist())[-12:],
        "max_margin64_of_mismatch": float(m64[mm32].max()) if int(mm32.sum()) else 0.0,
    }


def full_stats(ae, d, B):
    x, y = d["x"], d["y"]
    y_dec, y_syn, x_dec, x_with_si, bpp = [np.array(a) for a in ae.siNet_get_reconstructed(x, y)]
    keep = d["keep"]
    row, col = ae.last["row"].cpu().numpy(), ae.last["col"].cpu().numpy()
    agree = (row == d["ref_row"]) & (col == d["ref_col"])
    out = {"bpp": float(bpp), "ref_bpp": float(d["ref_bpp"]), "d_bpp": float(bpp) - float(d["ref_bpp"]),
           "rowcol_agree": float(agree.mean()), "rowcol_mismatch": int((~agree).sum()),
           "sym_mismatch_x": int((ae.last["symbols"].cpu().numpy() != d["ref_symbols"]).sum())}
    for k, g in (("y_dec", y_dec), ("x_dec", x_dec), ("y_syn", y_syn), ("x_with_si", x_with_si)):
        diff = np.abs(g[keep] - d["ref_" + k])
        out["max_d_" + k] = float(diff.max())
        out["rms_d_" + k] = float(np.sqrt((diff.astype(np.float64) ** 2).mean()))
    dms, dmc = [], []
    for n in range(B):
        xi = np.transpose(x[n], (1, 2, 0)).astype(np.uint8)
        ri = np.transpose(np.clip(x_with_si[n], 0, 255), (1, 2, 0))
        dms.append(float(M.msssim_standard(xi, ri)) - float(d["ref_msssim_std"][n]))
        dmc.append(float(M.msssim_reference_call(xi, ri)) - float(d["ref_msssim_call"][n]))
    out["d_msssim_std"], out["d_msssim_call"] = dms, dmc
    out["max_abs_d_msssim"] = float(max(np.abs(dms).max(), np.abs(dmc).max()))
    return out


def time_policy(ae, B, steps=6):
    H, W = 320, 1224
    sets = []
    for s in range(3):
        x, y = synth.make_batch(B, H, W, seed=1000 + 17 * s)
        sets.append((torch.tensor(x).cuda(), torch.tensor(y).cuda()))
    for i in range(3):
        ae.replay_device(*sets[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ae.replay_device(*sets[i % 3])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ops.PROF.start()
    ae.reconstruct_device(*sets[0])
    torch.cuda.synchronize()
    ops.PROF.stop()
    kern = {k: round(v["ms"], 3) for k, v in sorted(ops.PROF.summary().items(), key=lambda kv: -kv[1]["ms"])}
    return {"ms_per_step": ms, "mpix_s": B * H * W * 1e-6 / (ms * 1e-3), "kernels_ms": kern}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--cases", default="smoke,full1")
    ap.add_argument("--policies", default="exact,mixed,mixed_y1,fast")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--no-timing", action="store_true")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    res = {"cases": {}, "timing": {}}
    for cname in args.cases.split(","):
        t0 = time.time()
        Wt, d = oracle_cache.get(cname)
        print("case %s loaded in %.1f s" % (cname, time.time() - t0), file=sys.stderr, flush=True)
        c = oracle_cache.CASES[cname]
        res["cases"][cname] = {}
        for pname in args.policies.split(","):
            pol = precision.get(pname)
            ae = make_ae(c["H"], c["W"], Wt, precision=pol)
            r = {}
            xg, yg = torch.tensor(d["x"]).cuda().float(), torch.tensor(d["y"]).cuda().float()
            ex = ae.ae_imgcomp.encode(xg, terms=pol.enc_x)
            r["enc_x"] = enc_stats(ex.z, ex.symbols, d, "x")
            ey = ae.ae_imgcomp.encode(yg, terms=pol.enc_y)
            r["enc_y"] = enc_stats(ey.z, ey.symbols, d, "y")
            # decoder alone, fed the oracle's qbar
            if "qbar32_x" in d and len(d["keep"]) == c["B"]:
                xd = ae.ae_imgcomp.decode(torch.tensor(d["qbar32_x"]).cuda().contiguous(), terms=pol.dec).cpu().numpy()
                diff = np.abs(xd - d["ref_x_dec"])
                r["dec_alone_max"], r["dec_alone_rms"] = float(diff.max()), float(np.sqrt((diff.astype(np.float64) ** 2).mean()))
            r["full"] = full_stats(ae, d, c["B"])
            res["cases"][cname][pname] = r
            print(cname, pname, json.dumps(r), file=sys.stderr, flush=True)
            del ae
            torch.cuda.empty_cache()
    if not args.no_timing:
        Wt = synth.make_weights(0, residual_gamma=0.25)
        for pname in args.policies.split(","):
            ae = make_ae(320, 1224, Wt, precision=precision.get(pname))
            res["timing"][pname] = time_policy(ae, args.batch)
            print("timing", pname, json.dumps(res["timing"][pname]), file=sys.stderr, flush=True)
            del ae
            torch.cuda.empty_cache()
    txt = json.dumps(res, indent=1)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt)
    print(txt)


if __name__ == "__main__":
    main()
