"""Data-parallel sharding of image pairs and the path's only collective.

Each (x, y) pair is independent (inference BatchNorm, per-pair SI search), so a batch is split into
contiguous blocks, one per rank, with no data-path collective.  At the end of a run every rank
contributes its metric partials [sum_bits, sum_pixels, sum_msssim, n_images]; one all-gather over
NCCL/NVLink (gloo in CPU tests) gives every rank the global bpp = sum_bits / sum_pixels -- the
batch semantics of bits.bitcost_to_bpp (/root/reference/src/bits_imgcomp.py:13-14) -- and mean MS-SSIM.
The reference itself has no distributed code."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank`; the first n_items % world ranks get one extra."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_metrics(sum_bits, sum_pixels, sum_msssim=0.0, n_images=0, device=None):
    """All-gather the per-rank partials; returns dict(bpp, msssim, n_images, per_rank)."""
    part = torch.tensor([float(sum_bits), float(sum_pixels), float(sum_msssim), float(n_images)],
                        dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parts = [torch.zeros_like(part) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, part)
        allp = torch.stack(parts)
    else:
        allp = part.unsqueeze(0)
    tot = allp.sum(0)
    return {"bpp": float(tot[0] / tot[1]) if float(tot[1]) > 0 else float("nan"),
            "msssim": float(tot[2] / tot[3]) if float(tot[3]) > 0 else float("nan"),
            "n_images": int(tot[3]), "per_rank": allp.cpu().tolist()}
