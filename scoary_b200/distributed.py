"""Gene sharding across the GPUs of one box.

One process per GPU (torchrun); genes are split into contiguous row blocks, the
trait bitsets and the tree are replicated, there is no communication during
compute, and ONE all-gather of fixed-size per-gene records at the end brings
every rank the whole result (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
import numpy as np

from .engine import shard_bounds  # noqa: F401  (re-exported)

RECORD_WORDS = 11   # int32 words: counts[4] | p (f64 as 2 words) | pairs[3] | r | n_done


def pack_records(counts, p, pairs, r, n_done):
    """numpy arrays of one shard -> int32 [G][RECORD_WORDS]."""
    G = len(p)
    rec = np.zeros((G, RECORD_WORDS), dtype=np.int32)
    rec[:, 0:4] = counts
    rec[:, 4:6] = np.ascontiguousarray(p, dtype=np.float64).view(np.int32).reshape(G, 2)
    rec[:, 6:9] = pairs
    rec[:, 9] = r
    rec[:, 10] = n_done
    return rec


def unpack_records(rec):
    rec = np.ascontiguousarray(rec, dtype=np.int32)
    return {"counts": rec[:, 0:4].copy(), "p": np.ascontiguousarray(rec[:, 4:6]).view(np.float64).reshape(-1),
            "pairs": rec[:, 6:9].copy(), "r": rec[:, 9].copy(), "n_done": rec[:, 10].copy()}


def all_gather_records(rec_tensor, n_total, bounds):
    """rec_tensor: torch int32 [g_local][RECORD_WORDS] on this rank's device (or CPU for gloo).
    Shards may differ in size by one row, so every rank pads to the largest shard; one
    collective.  Returns a torch tensor [n_total][RECORD_WORDS] in gene order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    gmax = max(hi - lo for lo, hi in bounds)
    pad = torch.zeros((gmax, RECORD_WORDS), dtype=torch.int32, device=rec_tensor.device)
    pad[: rec_tensor.shape[0]] = rec_tensor
    flat = torch.empty((world * gmax, RECORD_WORDS), dtype=torch.int32, device=rec_tensor.device)
    dist.all_gather_into_tensor(flat, pad)          # output = rank-major concatenation along dim 0
    out = flat.view(world, gmax, RECORD_WORDS)
    parts = [out[r, : bounds[r][1] - bounds[r][0]] for r in range(world)]
    full = torch.cat(parts, dim=0)
    assert full.shape[0] == n_total
    return full
