"""Sharding one global batch over N ranks changes nothing, bit for bit (SURVEY 8e / BASELINE configs[4]).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_equivalence.py [--global-batch 8]

Every rank runs AE.siNet_get_reconstructed on its dist.shard_range block of a seeded global batch; rank 0 also runs
the WHOLE batch alone (micro-batches of a different size, so batch composition differs too) and compares, per
image: symbols, per-image bit sums, (row, col), and SHA-256 of y_dec / y_syn / x_dec / x_with_si.  One NCCL
all_gather_object carries the digests.  Prints one JSON line; exit code 1 on any difference."""
import argparse
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def digests(ae, x, y):
    out = []
    res = [np.array(a) for a in ae.siNet_get_reconstructed(x, y)]
    last = ae.last
    sym, row, col = last["symbols"].cpu().numpy(), last["row"].cpu().numpy(), last["col"].cpu().numpy()
    bits = last["bits_sum"].cpu().numpy()
    for n in range(x.shape[0]):
        h = hashlib.sha256()
        for a in (res[0][n], res[1][n], res[2][n], res[3][n]):
            h.update(np.ascontiguousarray(a).tobytes())
        out.append({"images": h.hexdigest(), "symbols": hashlib.sha256(sym[n].tobytes()).hexdigest(),
                    "rowcol": hashlib.sha256(row[n].tobytes() + col[n].tobytes()).hexdigest(), "bits": float(bits[n])})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--global-batch", type=int, default=8)
    ap.add_argument("--hw", default="320x1224")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from dsin_b200 import synth
    from dsin_b200.dist import shard_range
    from parity_utils import make_ae
    H, W = (int(v) for v in args.hw.split("x"))
    G = args.global_batch
    Wt = synth.make_weights(0, residual_gamma=0.25)
    ae = make_ae(H, W, Wt)
    pairs = [synth.make_pair(5000 + g, H, W) for g in range(G)]      # pair g is the same on every rank
    X = np.stack([p[0] for p in pairs]).astype(np.uint8)
    Y = np.stack([p[1] for p in pairs]).astype(np.uint8)
    lo, hi = shard_range(G, rank, world)
    mine = digests(ae, X[lo:hi], Y[lo:hi]) if hi > lo else []
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, (lo, hi, mine))
    else:
        gathered = [(lo, hi, mine)]
    ok = True
    if rank == 0:
        whole = []
        mb = 3  # deliberately not a divisor of the shard sizes
        for s in range(0, G, mb):
            whole += digests(ae, X[s:s + mb], Y[s:s + mb])
        sharded = [None] * G
        for lo_, hi_, d in gathered:
            sharded[lo_:hi_] = d
        diffs = [g for g in range(G) if sharded[g] != whole[g]]
        ok = not diffs
        print(json.dumps({"world_size": world, "global_batch": G, "geometry": [H, W], "identical_per_image": ok,
                          "differing_images": diffs, "shards": [(a, b) for a, b, _ in gathered],
                          "checked": ["symbols", "bits per image", "row/col", "sha256(y_dec,y_syn,x_dec,x_with_si)"]}),
              flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
