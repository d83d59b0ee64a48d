"""The N > 1 path on CPU: two gloo ranks shard the genes, compute their shard (oracle-backed
FakeEngine standing in for the GPU), all-gather the per-gene records once, and every rank
ends up with the unsharded result."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, G, N, P, seed, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from fake_engine import FakeEngine
    from scoary_b200 import distributed as D
    from scoary_b200 import synth, tree as treemod
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    traits = synth.make_traits(N, 1, seed)
    bits = synth.make_genes_packed(G, N, seed, traits=traits)
    nested = synth.make_tree(N, seed)
    left, right, names = treemod.flatten(nested)
    col = {n: j for j, n in enumerate(synth.isolate_names(N))}
    cols = np.asarray([col[n] for n in names], dtype=np.int32)
    bounds = D.shard_bounds(G, world)
    lo, hi = bounds[rank]
    e = FakeEngine()
    e.set_genes(bits[lo:hi], N)
    e.set_trait_vector(0, traits[0])
    e.set_tree(0, left, right, cols)
    counts, p, _ = e.contingency_fisher(0)
    pairs, r, nd = e.permute(0, P, seed=seed)
    rec = torch.from_numpy(D.pack_records(counts, p, pairs, r, nd))
    full = D.all_gather_records(rec, G, bounds).numpy()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), full)
    dist.destroy_process_group()


def test_two_rank_gene_sharding(tmp_path):
    G, N, P, seed = 37, 60, 12, 5          # odd G: the shards differ in size
    world = 2
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, G, N, P, seed, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_engine import FakeEngine
    from scoary_b200 import distributed as D
    from scoary_b200 import synth, tree as treemod
    traits = synth.make_traits(N, 1, seed)
    bits = synth.make_genes_packed(G, N, seed, traits=traits)
    left, right, names = treemod.flatten(synth.make_tree(N, seed))
    col = {n: j for j, n in enumerate(synth.isolate_names(N))}
    e = FakeEngine()
    e.set_genes(bits, N)
    e.set_trait_vector(0, traits[0])
    e.set_tree(0, left, right, np.asarray([col[n] for n in names], dtype=np.int32))
    counts, p, _ = e.contingency_fisher(0)
    pairs, r, nd = e.permute(0, P, seed=seed)
    want = D.pack_records(counts, p, pairs, r, nd)
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % rank))
        assert np.array_equal(got, want)
    back = D.unpack_records(want)
    assert np.array_equal(back["p"].view(np.uint64), p.view(np.uint64)) and np.array_equal(back["pairs"], pairs)


def test_shard_bounds_cover_everything():
    from scoary_b200.distributed import shard_bounds
    for n, w in [(10, 3), (8, 8), (5, 8), (100000, 8), (1, 1)]:
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
