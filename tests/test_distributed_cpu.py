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


# ---------------------------------------------------------------------------- the CLI under torchrun-style env
GOLD = os.path.join(ROOT, "tests", "golden")


def _cli_worker(rank, world, port, argv, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "WORLD_SIZE": str(world),
                       "RANK": str(rank), "LOCAL_RANK": str(rank)})
    from fake_engine import FakeEngine
    from scoary_b200 import methods as M
    M._ENGINE = FakeEngine()
    try:
        M.main(argv=argv + ["-o", out_dir])
    except SystemExit as ex:
        assert ex.code == 0, ex.code


def _gunzip_inputs(tmp):
    import gzip
    import shutil
    g = os.path.join(tmp, "Gene_presence_absence.csv")
    with gzip.open(os.path.join(GOLD, "inputs", "Gene_presence_absence.csv.gz"), "rb") as fi, open(g, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    return g, os.path.join(GOLD, "inputs", "Tetracycline_resistance.csv")


def _read(path):
    import gzip
    with (gzip.open if path.endswith(".gz") else open)(path, "rt") as fh:
        return fh.read()


def test_two_rank_cli_matches_reference_goldens(tmp_path):
    """`torchrun --nproc-per-node 2 -m scoary_b200.methods ...`: sharded Fisher pass, one gather, host
    filter on every rank, survivors dealt out strided, second gather; rank 0's files are the reference's."""
    g, t = _gunzip_inputs(str(tmp_path))
    world = 2
    for k, (name, extra) in enumerate({"default": [],
                                       "advanced": ["-p", "0.01", "1E-5", "-c", "B", "EPW", "--collapse", "-m", "50",
                                                    "-u"],
                                       "all": ["-p", "1.0", "-c", "I"]}.items()):
        out = str(tmp_path / name)
        port = 29600 + (os.getpid() % 1000) + k
        mp.spawn(_cli_worker, args=(world, port, ["-g", g, "-t", t, "--no-time"] + extra, out), nprocs=world,
                 join=True)
        for trait in ("Tetracycline_resistance", "Bogus_trait"):
            gold = os.path.join(GOLD, name, trait + ".results.csv")
            gold = gold if os.path.exists(gold) else gold + ".gz"
            assert _read(os.path.join(out, trait + ".results.csv")) == _read(gold), (name, trait)
        assert sorted(f for f in os.listdir(out) if f.endswith(".log")) == ["scoary.log"]     # rank 0 only


def test_three_rank_cli_permutations_match_one_rank(tmp_path):
    """Empirical p does not depend on how the genes are dealt out (labellings are a function of
    (seed, trait, permutation index) only)."""
    g, t = _gunzip_inputs(str(tmp_path))
    argv = ["-g", g, "-t", t, "--no-time", "-e", "60", "-c", "I", "EPW", "-p", "0.05", "0.05"]
    mp.spawn(_cli_worker, args=(3, 29700 + (os.getpid() % 1000), argv, str(tmp_path / "two")), nprocs=3, join=True)
    mp.spawn(_cli_worker, args=(1, 29800 + (os.getpid() % 1000), argv, str(tmp_path / "one")), nprocs=1, join=True)
    for trait in ("Tetracycline_resistance", "Bogus_trait"):
        a = _read(os.path.join(str(tmp_path / "two"), trait + ".results.csv"))
        assert a == _read(os.path.join(str(tmp_path / "one"), trait + ".results.csv"))
        assert "Empirical_p" in a.splitlines()[0]


def _worker_wide_records(rank, world, port, q):
    """records of T traits side by side (what bench.py gathers) and a rank that owns no genes"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from scoary_b200 import distributed as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for G, T in ((11, 3), (2, 2), (1, 1)):             # G < world: some ranks hold an empty shard
        bounds = D.shard_bounds(G, world)
        lo, hi = bounds[rank]
        full = np.arange(G * T * D.RECORD_WORDS, dtype=np.int32).reshape(G, T * D.RECORD_WORDS)
        got = D.all_gather_records(torch.from_numpy(full[lo:hi].copy()), G, bounds).numpy()
        ok = ok and np.array_equal(got, full)
        ok = ok and np.array_equal(D.gather_blocks(full[lo:hi], G), full)
        ok = ok and np.array_equal(D.gather_strided(full[rank::world], G), full)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_gathers_with_several_traits_per_record_and_empty_shards():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + 7
    procs = [ctx.Process(target=_worker_wide_records, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True), (2, True)]


def test_cli_with_fewer_genes_than_ranks(tmp_path):
    """G < world leaves a rank without genes; it must only take part in the gathers (no hang, no error),
    and the files must equal a one-rank run."""
    from scoary_b200.methods import ROARY_COLUMNS
    g, t = str(tmp_path / "g.csv"), str(tmp_path / "t.csv")
    iso = ["i%d" % k for k in range(10)]
    with open(g, "w") as fh:
        fh.write(",".join('"%s"' % c for c in ROARY_COLUMNS[:14]) + "," + ",".join(iso) + "\n")
        fh.write('"gA","","x",1,1,1,1,,,,,,,,' + ",".join("1" if k % 2 else "" for k in range(10)) + "\n")
        fh.write('"gB","","y",1,1,1,1,,,,,,,,' + ",".join("1" if k < 4 else "" for k in range(10)) + "\n")
    with open(t, "w") as fh:
        fh.write(",T\n" + "".join("%s,%d\n" % (n, k < 5) for k, n in enumerate(iso)))
    argv = ["-g", g, "-t", t, "--no-time", "-p", "1.0", "-c", "I", "-e", "20"]
    mp.spawn(_cli_worker, args=(3, 29900 + (os.getpid() % 90), argv, str(tmp_path / "three")), nprocs=3, join=True)
    mp.spawn(_cli_worker, args=(1, 29990 - (os.getpid() % 90), argv, str(tmp_path / "one")), nprocs=1, join=True)
    a = _read(os.path.join(str(tmp_path / "three"), "T.results.csv"))
    assert a == _read(os.path.join(str(tmp_path / "one"), "T.results.csv")) and len(a.splitlines()) == 3


def _worker_perm_split(rank, world, port, q):
    """the other way N GPUs split an exhaustive job: every rank walks all genes under its own range of the
    permutations (Engine.permute_range), the hit counts are summed (distributed.all_reduce_sum)"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from fake_engine import FakeEngine
    from scoary_b200 import distributed as D
    from scoary_b200 import synth, tree as treemod
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G, N, P, seed = 23, 50, 37, 9
    traits = synth.make_traits(N, 1, seed)
    bits = synth.make_genes_packed(G, N, seed, traits=traits)
    left, right, names = treemod.flatten(synth.make_tree(N, seed))
    col = {n: j for j, n in enumerate(synth.isolate_names(N))}
    e = FakeEngine()
    e.set_genes(bits, N)
    e.set_trait_vector(0, traits[0])
    e.set_tree(0, left, right, np.asarray([col[n] for n in names], dtype=np.int32))
    first, count = D.permutation_range(P, world, rank)
    pairs, r_part = e.permute_range(0, first, count, seed=seed)
    r = D.all_reduce_sum(r_part)
    want_pairs, want_r, _ = e.permute(0, P, seed=seed)
    ok = np.array_equal(r, want_r) and np.array_equal(pairs, want_pairs)
    ok = ok and D.split_for(50000, 10000, 8) == "permutations" and D.split_for(1000000, 1000, 8) == "genes"
    ok = ok and D.split_for(50000, 10000, 1) == "genes" and D.split_for(50000, 100, 8) == "genes"
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_permutation_split_adds_up_under_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + 11
    procs = [ctx.Process(target=_worker_perm_split, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True), (2, True)]
