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
    """rec_tensor: torch int32 [g_local][RECORD_WORDS x traits] on this rank's device (or CPU for gloo).
    Shards may differ in size by one row, so every rank pads to the largest shard; one
    collective.  Returns a torch tensor [n_total][RECORD_WORDS] in gene order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    gmax = max(hi - lo for lo, hi in bounds)
    width = rec_tensor.shape[1]                     # RECORD_WORDS per trait
    pad = torch.zeros((gmax, width), dtype=torch.int32, device=rec_tensor.device)
    pad[: rec_tensor.shape[0]] = rec_tensor
    flat = torch.empty((world * gmax, width), dtype=torch.int32, device=rec_tensor.device)
    dist.all_gather_into_tensor(flat, pad)          # output = rank-major concatenation along dim 0
    out = flat.view(world, gmax, width)
    parts = [out[r, : bounds[r][1] - bounds[r][0]] for r in range(world)]
    full = torch.cat(parts, dim=0)
    assert full.shape[0] == n_total
    return full


# ---------------------------------------------------------------------------- the CLI's N > 1 flow
# `torchrun --nproc-per-node N -m scoary_b200.methods ...` (two-phase):
#   phase 1  contiguous gene shards -> counts / Fisher p / pattern hash -> ONE all-gather ->
#            every rank runs the same deterministic host code (skip rule, collapse, Bonferroni/BH, sort,
#            cut-offs), so every rank knows the surviving genes without a second exchange;
#   phase 2  the survivors are dealt out again, strided like the reference's worker domains
#            (range(t, num_results, threads), scoary/methods.py:1077) so that early-stopping and
#            full-length permutation runs mix evenly -> walks + permutations -> ONE all-gather.
# Rank 0 writes the files.  Host arrays in, host arrays out; NCCL when GPUs are there, gloo on CPU.
_OWN_GROUP = False


def init_from_env():
    """Join the process group torchrun describes (WORLD_SIZE > 1); no-op otherwise."""
    global _OWN_GROUP
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return False
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        cuda = torch.cuda.is_available()
        if cuda:
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl" if cuda else "gloo")
        _OWN_GROUP = True
    return True


def finish():
    global _OWN_GROUP
    if _OWN_GROUP:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
        _OWN_GROUP = False


def world_rank():
    """(world, rank) of the initialised process group, (1, 0) without one."""
    try:
        import sys
        if "torch" not in sys.modules:
            return 1, 0
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
    except ImportError:
        pass
    return 1, 0


def _gather_padded(rows, n_max):
    """int32 [n_r][W] per rank -> numpy int32 [world][n_max][W] (one all-gather of padded blocks)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    W = rows.shape[1]
    pad = torch.zeros((n_max, W), dtype=torch.int32, device=dev)
    if len(rows):
        pad[: len(rows)] = torch.from_numpy(rows).to(dev)
    flat = torch.empty((world * n_max, W), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(flat, pad)
    return flat.view(world, n_max, W).cpu().numpy()


def gather_blocks(rows, n_total):
    """Rank r holds rows shard_bounds(n_total, world)[r] -> every rank gets all n_total rows in order."""
    world, _ = world_rank()
    bounds = shard_bounds(n_total, world)
    parts = _gather_padded(rows, max(hi - lo for lo, hi in bounds))
    return np.concatenate([parts[r][: bounds[r][1] - bounds[r][0]] for r in range(world)], axis=0)


def all_reduce_sum(values):
    """Sum of an int32 array over the ranks (every rank gets the total): the hit counts of the permutation ranges the
    ranks walked (Engine.permute_range) add up to the job's."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(np.ascontiguousarray(values, dtype=np.int32)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def permutation_range(P, world, rank):
    """[first, first + count): rank's share of P permutations (contiguous, sizes differ by at most one)."""
    lo, hi = shard_bounds(P, world)[rank]
    return lo, hi - lo


def split_for(G, P, world, per_tile=768, min_tiles=64, min_perms=64):
    """How N GPUs split an EXHAUSTIVE job: by genes (the reference's own fan-out, methods.py:1076-1097) while a shard is
    still many thread tiles (C5: 163 per GPU), else by permutations -- every GPU walks all genes under its own range of
    the labellings, so launches stay as large as on one GPU, and the hit counts are summed.  Measured on the north_star
    job: gene shards of 25 000 / 6 250 genes scale 1.95x / 6.60x on 2 / 8 GPUs, permutation ranges 2.00x / 3.98x / 7.90x on
    2 / 4 / 8 (profiles/r2_summary.md); what the permutation split repeats per GPU is one Fisher pass and one unpermuted
    walk per gene, i.e. 1 / (P / N) of its work."""
    if world > 1 and G / world < min_tiles * per_tile and P // world >= min_perms:
        return "permutations"
    return "genes"


def gather_strided(rows, n_total):
    """Rank r holds rows r, r + world, r + 2 world, ... -> every rank gets all n_total rows in order."""
    world, _ = world_rank()
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    out = np.zeros((n_total, rows.shape[1]), dtype=np.int32)
    if n_total == 0:
        return out
    parts = _gather_padded(rows, -(-n_total // world))
    for r in range(world):
        n_r = len(range(r, n_total, world))
        out[r::world] = parts[r][:n_r]
    return out
