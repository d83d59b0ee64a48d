"""The CUDA walk kernels' own source (scoary_b200/csrc/walk.cuh), compiled for the HOST by
tests/host_emul/walk_emul.cu and run one simulated thread at a time, against the oracle: the CPU tier
then covers the real DP arithmetic (packed 16-bit mode, widening, fused ops, per-gene bonuses, the hit
test), not only a Python model of it.  Needs nvcc (cross-compiles without a GPU); skipped without it."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from scoary_b200 import synth
from scoary_b200 import tree as treemod
from test_tree_program import compile_tree

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")
SRC = os.path.join(HERE, "walk_emul.cu")
LIB = os.path.join(HERE, "libwalk_emul.so")
CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scoary_b200", "csrc")


def _build_walk_emul(lib_path, extra=()):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    deps = [SRC, os.path.join(CSRC, "walk.cuh"), os.path.join(CSRC, "common.cuh")]
    if not os.path.exists(lib_path) or os.path.getmtime(lib_path) < max(os.path.getmtime(d) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler", "-fPIC",
                        "-ccbin", cxx, *extra, "-shared", "-o", lib_path, SRC], check=True)
    lib = ctypes.CDLL(lib_path)
    lib.emul_pairs.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                               ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.emul_permute.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_void_p, ctypes.c_void_p]
    if hasattr(lib, "emul_permute_transposed"):
        lib.emul_permute_transposed.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return lib


@pytest.fixture(scope="module")
def emul():
    return _build_walk_emul(LIB)


@pytest.fixture(scope="module")
def emul_prmt():
    """the experimental -DSB_WALK_PRMT=1 build of the same source (bit-reversed gene windows, masks by PRMT sign
    replication -- emulated here from the PTX definition of prmt.b32)"""
    return _build_walk_emul(os.path.join(HERE, "libwalk_emul_prmt.so"), ("-DSB_WALK_PRMT=1",))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _pack_walk_order(bits_by_leaf, order, W32p):
    """uint8 [R][n_leaves] (leaf-id order) -> uint32 [R][W32p], bit b of word w = the leaf consumed at position 32w+b"""
    walk = np.zeros((bits_by_leaf.shape[0], W32p * 32), dtype=np.uint8)
    real = order >= 0                                     # pad positions (leaf id -1) carry bit 0
    walk[:, np.nonzero(real)[0]] = bits_by_leaf[:, order[real]]
    return np.ascontiguousarray(np.packbits(walk, axis=1, bitorder="little")).view(np.uint32).reshape(-1, W32p)


def _stream_words(order):
    """W32p of a leaf stream: 32-bit words, padded to a multiple of four (engine.cu sb_set_tree)"""
    return ((len(order) + 31) // 32 + 3) // 4 * 4


def _setup(n, G, seed, comb=False):
    rng = np.random.default_rng(seed)
    names = synth.isolate_names(n)
    if comb:
        nested = names[0]
        for nm in names[1:]:
            nested = [nested, nm]
    else:
        nested = synth.make_tree(n, seed)
    ops, order, units, leaves = compile_tree(nested)
    left, right, _ = treemod.flatten(nested)
    f = rng.uniform(0.02, 0.98, size=G)
    m = (rng.random((G, n)) < f[:, None]).astype(np.uint8)
    lab = (rng.random(n) < 0.4).astype(np.uint8)
    m[0] = lab                                            # a perfectly associated gene
    if G > 1:
        m[1] = 1 - lab
    W32p = _stream_words(order)
    shift = 1
    while (1 << shift) <= n // 2:
        shift += 1
    Gs = (G + 31) // 32 * 32
    gw = _pack_walk_order(m, order, W32p)                 # [G][W32p]
    genesT = np.zeros((W32p, Gs), dtype=np.uint32)
    genesT[:, :G] = gw.T
    return dict(ops=np.ascontiguousarray(ops), order=order, units=units, left=left, right=right, m=m, lab=lab,
                W32p=W32p, shift=shift, Gs=Gs, genesT=np.ascontiguousarray(genesT), G=G, n=n)


CASES = [(2, 5, False), (3, 9, False), (5, 40, False), (16, 70, False), (100, 130, False), (127, 33, False),
         (128, 33, False), (129, 600, False), (300, 70, False), (1000, 20, False), (400, 12, True), (150, 520, True),
         (5000, 40, False), (10000, 6, False), (20000, 5, True)]      # C3 / C4 isolate counts, a 19 999-level comb


def _check_pairs(lib, n, G, comb):
    c = _setup(n, G, 100 + n, comb)
    lab0 = _pack_walk_order(c["lab"][None, :], c["order"], c["W32p"])[0]
    pairs = np.full((G, 3), -7, dtype=np.int32)
    rc = lib.emul_pairs(_ptr(c["ops"]), len(c["ops"]), _ptr(lab0), _ptr(c["genesT"]), c["Gs"], G, c["W32p"], c["shift"],
                        c["units"], _ptr(pairs))
    assert rc == 0
    ref = O.permute(c["left"], c["right"], c["m"], c["lab"], P=0)["pairs"]
    assert np.array_equal(pairs, ref)
    _no_carries(lib)


def _check_permute(lib, n, G, comb, ppi):
    c = _setup(n, G, 200 + n, comb)
    P, seed = 9, 77
    ref = O.permute(c["left"], c["right"], c["m"], c["lab"], P=P, seed=seed, trait=0, want_hits=True)
    labs = np.stack([O.shuffle_labels(seed, 0, p, c["lab"]) for p in range(P)])
    labelsW = np.ascontiguousarray(_pack_walk_order(labs, c["order"], c["W32p"]))
    n_chunks = (P + ppi - 1) // ppi
    hits = np.zeros((n_chunks, G), dtype=np.uint8)
    unperm = np.ascontiguousarray(ref["pairs"], dtype=np.int32)
    rc = lib.emul_permute(_ptr(c["ops"]), len(c["ops"]), _ptr(labelsW), P, ppi, _ptr(c["genesT"]), c["Gs"], G, c["W32p"],
                          c["shift"], c["units"], _ptr(unperm), _ptr(hits))
    assert rc == 0
    got = np.zeros((G, P), dtype=np.uint8)
    for p in range(P):
        got[:, p] = (hits[p // ppi] >> (p % ppi)) & 1
    assert np.array_equal(got, ref["hits"])
    assert np.array_equal(got.sum(axis=1), ref["r"])
    _no_carries(lib)


def _check_permute_transposed(lib, n, G, comb, ppi, P=41):
    """the transposed launch (threads = labellings, constant rows = genes): the same hit flags as the oracle's
    Permute -- the DP is symmetric in the gene and the trait bit of a leaf"""
    c = _setup(n, G, 300 + n, comb)
    seed = 78
    ref = O.permute(c["left"], c["right"], c["m"], c["lab"], P=P, seed=seed, trait=0, want_hits=True)
    labs = np.stack([O.shuffle_labels(seed, 0, p, c["lab"]) for p in range(P)])
    labelsW = _pack_walk_order(labs, c["order"], c["W32p"])                 # [P][W32p]
    Ps = (P + 31) // 32 * 32
    labelsT = np.zeros((c["W32p"], Ps), dtype=np.uint32)
    labelsT[:, :P] = labelsW.T
    rowsW = np.ascontiguousarray(c["genesT"][:, :G].T)                      # [G][W32p]
    hits = np.full((G, Ps), 9, dtype=np.uint8)
    unperm = np.ascontiguousarray(ref["pairs"], dtype=np.int32)
    rc = lib.emul_permute_transposed(_ptr(c["ops"]), len(c["ops"]), _ptr(rowsW), G, ppi, _ptr(np.ascontiguousarray(labelsT)),
                                     Ps, P, c["W32p"], c["shift"], c["units"], _ptr(unperm), _ptr(hits))
    assert rc == 0
    assert np.array_equal(hits[:, :P], ref["hits"])
    assert np.all(hits[:, P:] == 9)                                         # nothing written past the labellings
    _no_carries(lib)


@pytest.mark.parametrize("n,G,comb", [(2, 5, False), (5, 9, False), (16, 7, False), (100, 13, False), (127, 6, False),
                                      (129, 30, False), (300, 12, False), (150, 8, True), (1000, 5, False),
                                      (5000, 3, False)])
@pytest.mark.parametrize("ppi", [1, 4])
def test_transposed_permute_kernel_source_on_the_host(emul, n, G, comb, ppi):
    if ppi != 4 and n > 200:
        pytest.skip("one ppi is enough for the large trees")
    _check_permute_transposed(emul, n, G, comb, ppi, P=41 if n < 2000 else 9)


def test_shared_stack_size_reported_by_the_compiler_is_tight(emul):
    """The emulated shared memory ends in guard words: with the compiler's stack_units the kernels stay inside
    (every other test checks rc == 0); with one unit less a 1000-leaf tree tramples the guard (rc == -2)."""
    c = _setup(1000, 20, 4242, False)
    lab0 = _pack_walk_order(c["lab"][None, :], c["order"], c["W32p"])[0]
    pairs = np.full((20, 3), -7, dtype=np.int32)
    args = (_ptr(c["ops"]), len(c["ops"]), _ptr(lab0), _ptr(c["genesT"]), c["Gs"], 20, c["W32p"], c["shift"])
    assert c["units"] >= 2
    assert emul.emul_pairs(*args, c["units"], _ptr(pairs)) == 0
    assert emul.emul_pairs(*args, c["units"] - 1, _ptr(pairs)) == -2


def _no_carries(lib):
    """SB_ADD2_NC (walk.cuh) is a plain 32-bit add on the device: its low 16-bit lane must never carry"""
    lib.emul_carry_violations.restype = ctypes.c_longlong
    assert lib.emul_carry_violations() == 0


@pytest.mark.parametrize("n,G,comb", CASES)
def test_pairs_kernel_source_on_the_host(emul, n, G, comb):
    _check_pairs(emul, n, G, comb)


@pytest.mark.parametrize("n,G,comb", CASES)
@pytest.mark.parametrize("ppi", [1, 2, 4])
def test_permute_kernel_source_on_the_host(emul, n, G, comb, ppi):
    if ppi != 4 and n > 200:
        pytest.skip("one ppi is enough for the large trees")
    _check_permute(emul, n, G, comb, ppi)


@pytest.mark.parametrize("n,comb", [(127, True), (127, False), (126, True), (254, True), (255, False), (1000, False)])
def test_extreme_keys_stay_inside_their_16_bit_lane(emul, n, comb):
    """Labels alternating along the walk and genes that follow / oppose them give the most pairs a subtree can hold
    (63 pairs and 63 pro or anti pairs in 127 leaves: key 4095, the largest drift an unreachable state can pick up).
    Results must equal the oracle's and SB_ADD2_NC must never see a low-lane carry."""
    c = _setup(n, 8, 900 + n, comb)
    pos_of_leaf = np.empty(n, dtype=np.int64)
    pos_of_leaf[c["order"][c["order"] >= 0]] = np.arange(n)        # rank in walk order (pad positions do not count)
    lab = (pos_of_leaf % 2).astype(np.uint8)                      # by leaf id: alternating in walk order
    m = np.stack([lab, 1 - lab, (pos_of_leaf // 2 % 2).astype(np.uint8), 1 - (pos_of_leaf // 2 % 2).astype(np.uint8),
                  np.ones(n, np.uint8), np.zeros(n, np.uint8), (pos_of_leaf % 3 == 0).astype(np.uint8),
                  (pos_of_leaf % 4 < 3).astype(np.uint8)])
    gw = _pack_walk_order(m, c["order"], c["W32p"])
    genesT = np.zeros((c["W32p"], c["Gs"]), dtype=np.uint32)
    genesT[:, :8] = gw.T
    lab0 = _pack_walk_order(lab[None, :], c["order"], c["W32p"])[0]
    pairs = np.full((8, 3), -7, dtype=np.int32)
    assert emul.emul_pairs(_ptr(c["ops"]), len(c["ops"]), _ptr(lab0), _ptr(genesT), c["Gs"], 8, c["W32p"], c["shift"],
                           c["units"], _ptr(pairs)) == 0
    ref = O.permute(c["left"], c["right"], m, lab, P=0)["pairs"]
    assert np.array_equal(pairs, ref)
    P = 4
    labs = np.stack([lab, 1 - lab, np.roll(lab, 1), (pos_of_leaf % 2 == 0).astype(np.uint8)])
    labelsW = np.ascontiguousarray(_pack_walk_order(labs, c["order"], c["W32p"]))
    hits = np.zeros((1, 8), dtype=np.uint8)
    unperm = np.ascontiguousarray(ref, dtype=np.int32)
    assert emul.emul_permute(_ptr(c["ops"]), len(c["ops"]), _ptr(labelsW), P, 4, _ptr(genesT), c["Gs"], 8, c["W32p"],
                             c["shift"], c["units"], _ptr(unperm), _ptr(hits)) == 0
    for p in range(P):                                            # the oracle, labelling by labelling
        want = O.permute(c["left"], c["right"], m, labs[p], P=0)["pairs"].astype(np.int64)
        side = np.where(ref[:, 1] >= ref[:, 2], 1, 2)             # methods.py:1333-1336
        stat, u_stat = want[np.arange(8), side], ref[np.arange(8), side].astype(np.int64)
        hit = stat * ref[:, 0].astype(np.int64) >= u_stat * want[:, 0]
        assert np.array_equal((hits[0] >> p) & 1, hit.astype(np.uint8)), p
    _no_carries(emul)
    assert ref[:, 0].max() >= (n // 2) - 1                        # the construction does reach the maximum


VARIANT_CASES = [(5, 40, False), (129, 600, False), (150, 520, True), (1000, 20, False), (5000, 12, False)]


@pytest.mark.parametrize("n,G,comb", VARIANT_CASES)
def test_prmt_variant_of_the_walk_kernels_on_the_host(emul_prmt, n, G, comb):
    _check_pairs(emul_prmt, n, G, comb)
    _check_permute(emul_prmt, n, G, comb, 4)


@pytest.fixture(scope="module")
def emul_nlab2():
    """the experimental -DSB_WALK_NLAB=2 build: K5 walks two labellings of its genes in lockstep"""
    return _build_walk_emul(os.path.join(HERE, "libwalk_emul_nlab2.so"), ("-DSB_WALK_NLAB=2",))


@pytest.mark.parametrize("n,G,comb", VARIANT_CASES)
@pytest.mark.parametrize("ppi", [2, 4])
def test_lockstep_variant_of_the_permutation_kernel_on_the_host(emul_nlab2, n, G, comb, ppi):
    """9 labellings in blocks of 2 or 4: the odd tail walks its last labelling twice and reports it once"""
    _check_pairs(emul_nlab2, n, G, comb)
    _check_permute(emul_nlab2, n, G, comb, ppi)


# ---------------------------------------------------------------------------- the two leaf-stream layouts
# -DSB_WALK_PADDED=1 (the default: no op crosses a 16-leaf window, the stream holds pad positions) and =0 (ops may
# cross, the kernels keep a window-crossing path), each as an explicit build of the tree compiler and of the kernels
@pytest.fixture(scope="module", params=["1", "0"], ids=["padded", "crossing"])
def padded(request):
    """(tree compiler of a -DSB_WALK_PADDED=<flag> build of the product library, walk kernels of the same build on the
    host, flag)"""
    flag = request.param
    so = os.path.join(HERE, "libscoary_b200_padded%s.so" % flag)
    deps = [os.path.join(CSRC, f) for f in ("engine.cu", "walk.cuh", "common.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-B", "-C", CSRC, "OUT=" + so, "EXTRA=-DSB_WALK_PADDED=" + flag], check=True, env=env,
                       stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(so)
    lib.sb_debug_compile_tree2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32,
                                           ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
    return lib, _build_walk_emul(os.path.join(HERE, "libwalk_emul_padded%s.so" % flag), ("-DSB_WALK_PADDED=" + flag,)), flag


def _compile_padded(lib, nested):
    left, right, names = treemod.flatten(nested)
    n = len(left)
    ops = np.zeros(6 * n + 64, dtype=np.uint16)
    order = np.full(3 * n + 64, -9, dtype=np.int32)
    n_pos, depth = ctypes.c_int32(), ctypes.c_int32()
    k = lib.sb_debug_compile_tree2(_ptr(left), _ptr(right), n, _ptr(ops), len(ops), _ptr(order), len(order),
                                   ctypes.byref(n_pos), ctypes.byref(depth))
    assert k > 0
    return ops[:k].copy(), order[:n_pos.value].copy(), depth.value, left, right


def _pack_stream(bits_by_leaf, order, W32p):
    """as _pack_walk_order, for a stream with pad positions (leaf id -1 -> bit 0)"""
    walk = np.zeros((bits_by_leaf.shape[0], W32p * 32), dtype=np.uint8)
    real = order >= 0
    walk[:, np.nonzero(real)[0]] = bits_by_leaf[:, order[real]]
    return np.ascontiguousarray(np.packbits(walk, axis=1, bitorder="little")).view(np.uint32).reshape(-1, W32p)


def test_padded_stream_program_never_crosses_a_window(padded):
    """Every leaf-consuming op of a padded program lies inside one 16-leaf window under the kernels' rule (an op
    that does not fit opens the next window), every leaf appears exactly once, pads cost at most ~25 % positions."""
    lib, _, flag = padded
    if flag == "0":
        pytest.skip("the crossing layout has no pad positions")
    for n, comb in ((2, False), (17, False), (127, True), (300, True), (1000, False), (5000, False)):
        names = synth.isolate_names(n)
        nested = names[0]
        if comb:
            for nm in names[1:]:
                nested = [nested, nm]
        else:
            nested = synth.make_tree(n, 31 + n)
        ops, order, _, left, _ = _compile_padded(lib, nested)
        assert sorted(order[order >= 0].tolist()) == list(range(n))
        assert len(order) <= 1.25 * n + 16
        room, pos = 0, 0
        for op in ops:
            kind, cnt = int(op) & 63, int(op) >> 6
            leaves = cnt + 2 if kind in (4, 6, 38, 10, 42) else cnt if kind in (2, 17) else 0    # csrc/walk.cuh op numbers
            if leaves == 0:
                continue
            assert leaves <= 16
            if leaves > room:
                assert np.all(order[pos:pos + room] == -1)       # what the kernel skips is padding
                pos += room
                room = 16
            assert np.all(order[pos:pos + leaves] >= 0)
            pos += leaves
            room -= leaves
        assert pos == len(order)


@pytest.mark.parametrize("n,G,comb", [(5, 40, False), (33, 9, False), (129, 300, False), (150, 260, True), (1000, 20, False),
                                      (5000, 12, False)])
def test_padded_variant_of_the_walk_kernels_on_the_host(padded, n, G, comb):
    lib, emul_p, _ = padded
    rng = np.random.default_rng(700 + n)
    names = synth.isolate_names(n)
    nested = names[0]
    if comb:
        for nm in names[1:]:
            nested = [nested, nm]
    else:
        nested = synth.make_tree(n, 700 + n)
    ops, order, units, left, right = _compile_padded(lib, nested)
    m = (rng.random((G, n)) < rng.uniform(0.02, 0.98, size=G)[:, None]).astype(np.uint8)
    lab = (rng.random(n) < 0.4).astype(np.uint8)
    m[0] = lab
    W32p = ((len(order) + 31) // 32 + 3) // 4 * 4
    shift = 1
    while (1 << shift) <= n // 2:
        shift += 1
    Gs = (G + 31) // 32 * 32
    genesT = np.zeros((W32p, Gs), dtype=np.uint32)
    genesT[:, :G] = _pack_stream(m, order, W32p).T
    genesT, ops = np.ascontiguousarray(genesT), np.ascontiguousarray(ops)
    P, seed = 9, 5
    ref = O.permute(left, right, m, lab, P=P, seed=seed, trait=0, want_hits=True)
    pairs = np.full((G, 3), -7, dtype=np.int32)
    lab0 = _pack_stream(lab[None, :], order, W32p)[0]
    assert emul_p.emul_pairs(_ptr(ops), len(ops), _ptr(lab0), _ptr(genesT), Gs, G, W32p, shift, units, _ptr(pairs)) == 0
    assert np.array_equal(pairs, ref["pairs"])
    labs = np.stack([O.shuffle_labels(seed, 0, p, lab) for p in range(P)])
    labelsW = np.ascontiguousarray(_pack_stream(labs, order, W32p))
    unperm = np.ascontiguousarray(ref["pairs"], dtype=np.int32)
    for ppi in (1, 4):
        hits = np.zeros(((P + ppi - 1) // ppi, G), dtype=np.uint8)
        assert emul_p.emul_permute(_ptr(ops), len(ops), _ptr(labelsW), P, ppi, _ptr(genesT), Gs, G, W32p, shift, units,
                                   _ptr(unperm), _ptr(hits)) == 0
        got = np.stack([(hits[p // ppi] >> (p % ppi)) & 1 for p in range(P)], axis=1)
        assert np.array_equal(got, ref["hits"])
    _no_carries(emul_p)


def _balanced(names):
    if len(names) == 1:
        return names[0]
    h = len(names) // 2
    return [_balanced(names[:h]), _balanced(names[h:])]


def test_random_campaign_over_tree_shapes_and_window_offsets(emul, emul_prmt, emul_nlab2):
    """40 random configurations (isolate counts around every 16 / 32 / 127-leaf boundary; random-join, comb and
    balanced trees; gene frequencies from 0 to 1) through the three builds: pairs, per-permutation hit flags, no carry."""
    rng = np.random.default_rng(20261017)
    sizes = [2, 3, 4, 5, 7, 15, 16, 17, 31, 32, 33, 47, 48, 49, 63, 64, 65, 100, 126, 127, 128, 129, 130, 200, 255, 256, 257,
             300, 511, 513]
    for _ in range(40):
        n, G, kind = int(rng.choice(sizes)), int(rng.integers(1, 24)), int(rng.integers(3))
        names = synth.isolate_names(n)
        if kind == 0:
            nested = synth.make_tree(n, int(rng.integers(1 << 30)))
        elif kind == 1:
            nested = names[0]
            for nm in names[1:]:
                nested = [nested, nm]
        else:
            nested = _balanced(names)
        ops, order, units, _ = compile_tree(nested)
        left, right, _ = treemod.flatten(nested)
        m = (rng.random((G, n)) < rng.uniform(0.0, 1.0, size=G)[:, None]).astype(np.uint8)
        lab = (rng.random(n) < rng.uniform(0.05, 0.95)).astype(np.uint8)
        m[0] = lab
        W32p = _stream_words(order)
        shift = 1
        while (1 << shift) <= n // 2:
            shift += 1
        Gs = (G + 31) // 32 * 32
        genesT = np.zeros((W32p, Gs), dtype=np.uint32)
        genesT[:, :G] = _pack_walk_order(m, order, W32p).T
        genesT, ops = np.ascontiguousarray(genesT), np.ascontiguousarray(ops)
        P, seed = int(rng.integers(1, 10)), int(rng.integers(1 << 30))
        ref = O.permute(left, right, m, lab, P=P, seed=seed, trait=0, want_hits=True)
        labs = np.stack([O.shuffle_labels(seed, 0, p, lab) for p in range(P)])
        labelsW = np.ascontiguousarray(_pack_walk_order(labs, order, W32p))
        unperm = np.ascontiguousarray(ref["pairs"], dtype=np.int32)
        lab0 = _pack_walk_order(lab[None, :], order, W32p)[0]
        for lib, ppis in ((emul, (1, 2, 4)), (emul_nlab2, (2, 4)), (emul_prmt, (4,))):
            pairs = np.full((G, 3), -7, dtype=np.int32)
            assert lib.emul_pairs(_ptr(ops), len(ops), _ptr(lab0), _ptr(genesT), Gs, G, W32p, shift, units, _ptr(pairs)) == 0
            assert np.array_equal(pairs, ref["pairs"]), (n, G, kind)
            for ppi in ppis:
                hits = np.zeros(((P + ppi - 1) // ppi, G), dtype=np.uint8)
                assert lib.emul_permute(_ptr(ops), len(ops), _ptr(labelsW), P, ppi, _ptr(genesT), Gs, G, W32p, shift, units,
                                        _ptr(unperm), _ptr(hits)) == 0
                got = np.stack([(hits[p // ppi] >> (p % ppi)) & 1 for p in range(P)], axis=1)
                assert np.array_equal(got, ref["hits"]), (n, G, kind, ppi, P)
            _no_carries(lib)


# ---------------------------------------------------------------------------- Fisher (csrc/fisher.cuh) on the host
FLIB = os.path.join(HERE, "libfisher_emul.so")
FSRC = os.path.join(HERE, "fisher_emul.cu")
FISHER_RTOL = 1e-10


@pytest.fixture(scope="module")
def femul():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    deps = [FSRC, os.path.join(CSRC, "fisher.cuh"), os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "lut.cpp")]
    if not os.path.exists(FLIB) or os.path.getmtime(FLIB) < max(os.path.getmtime(d) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler",
                        "-fPIC,-ffp-contract=off", "-ccbin", cxx, "-shared", "-o", FLIB, FSRC,
                        os.path.join(CSRC, "lut.cpp"), "-lquadmath", "-lpthread"], check=True)
    lib = ctypes.CDLL(FLIB)
    lib.emul_fisher.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]

    def run(tables):
        """tables int [n][4] in the oracle's / C-ABI's order (tpgp, tngp, tpgn, tngn) -> p [n]"""
        t = np.ascontiguousarray(np.asarray(tables, dtype=np.int32)[:, [0, 2, 1, 3]])     # kernel order a, b, c, d
        p = np.full(len(t), -1.0, dtype=np.float64)
        assert lib.emul_fisher(_ptr(t), len(t), _ptr(p)) == 0
        return p
    return run


def test_fisher_source_on_the_host_matches_scipy_goldens(femul):
    """The 600 SciPy tables the oracle is pinned on (tests/golden/fisher.json), through the kernel's own
    warp-cooperative code: same tolerance as the GPU parity tests (1e-10 relative)."""
    import json
    gold = json.load(open(os.path.join(os.path.dirname(HERE), "golden", "fisher.json")))["tables"]
    tabs = np.array([[g["tpgp"], g["tngp"], g["tpgn"], g["tngn"]] for g in gold], dtype=np.int32)
    want = np.array([float.fromhex(g["p"]) for g in gold])
    got = femul(tabs)
    ok = want > 1e-290
    print("max rel err vs SciPy", np.max(np.abs(got[ok] - want[ok]) / want[ok]))
    assert np.max(np.abs(got[ok] - want[ok]) / want[ok]) <= FISHER_RTOL
    assert np.all(got[~ok] <= 2e-290)
    # a batch that does not fill the last warp, in another order: same values per table
    perm = np.random.default_rng(1).permutation(len(tabs))[:-3]
    assert np.array_equal(femul(tabs[perm]).view(np.uint64), got[perm].view(np.uint64))


def test_fisher_source_on_the_host_large_tables_and_symmetry(femul):
    """N up to 20 000 against the oracle (binary128), and the canonical orientation: the eight symmetric
    variants of a table (row swap, column swap, transpose) must give bit-identical p."""
    rng = np.random.default_rng(5)
    tabs = []
    for N in (40, 1000, 5000, 10000, 20000):
        for _ in range(40):
            r1, c1 = int(rng.integers(1, N)), int(rng.integers(1, N))
            if rng.random() < 0.3:
                r1 = N // 2
            lo, hi = max(0, r1 + c1 - N), min(r1, c1)
            sd = max(r1 * c1 / N * (1 - r1 / N) * (1 - c1 / N), 1.0) ** 0.5
            a = int(np.clip(round(r1 * c1 / N + rng.normal() * 3 * sd), lo, hi))
            tabs.append([a, c1 - a, r1 - a, N - r1 - c1 + a])          # tpgp, tngp, tpgn, tngn
    for t in ([0, 0, 3, 5], [4, 0, 0, 6], [0, 7, 0, 2], [1, 0, 0, 0], [1, 1, 1, 1], [0, 1, 1, 0], [2, 0, 0, 1]):
        tabs.append(t)                                                 # empty margins, tiny tables
    tabs = np.array(tabs, dtype=np.int32)
    got = femul(tabs)
    want = O.fisher(tabs)
    ok = want > 1e-290
    print("max rel err vs binary128", np.max(np.abs(got[ok] - want[ok]) / want[ok]))
    assert np.max(np.abs(got[ok] - want[ok]) / want[ok]) <= FISHER_RTOL
    variants = []
    for a, c, b, d in tabs[:60]:                                      # [[a, b], [c, d]]
        for (w, x, y, z) in ((a, b, c, d), (c, d, a, b), (b, a, d, c), (d, c, b, a), (a, c, b, d), (b, d, a, c),
                             (c, a, d, b), (d, b, c, a)):
            variants.append([w, y, x, z])                             # back to tpgp, tngp, tpgn, tngn
    pv = femul(np.array(variants, dtype=np.int32)).reshape(-1, 8)
    assert np.all(pv.view(np.uint64) == pv.view(np.uint64)[:, :1])


def test_fisher_near_mode_complement_path_accuracy(femul):
    """Tables whose observed cell lies within a few standard deviations of the mode take the complement path of
    fisher.cuh (p = 1 - the terms between a and the far-side boundary, kept for p >= 0.01): 3 000 such tables at
    the BASELINE isolate counts against the oracle's binary128 sums -- the switch between the two paths (p around
    0.01, |z| around 2.5) must not be visible at 1e-10."""
    rng = np.random.default_rng(17)
    tabs = []
    for N in (100, 1000, 2000, 5000, 10000):
        for _ in range(600):
            r1 = int(rng.integers(max(2, N // 50), N - 1))
            c1 = int(rng.integers(max(2, N // 50), N - 1)) if rng.random() < 0.7 else int(0.35 * N)
            lo, hi = max(0, r1 + c1 - N), min(r1, c1)
            sd = max(r1 * c1 / N * (1 - r1 / N) * (1 - c1 / N), 0.25) ** 0.5
            z = rng.uniform(-3.2, 3.2)
            a = int(np.clip(round(r1 * c1 / N + z * sd), lo, hi))
            tabs.append([a, c1 - a, r1 - a, N - r1 - c1 + a])          # tpgp, tngp, tpgn, tngn
    tabs = np.array(tabs, dtype=np.int32)
    got = femul(tabs)
    want = O.fisher(tabs)
    rel = np.abs(got - want) / want
    print("near-mode tables: max rel err %.2e; p range %.2e .. 1; %d tables with 0.005 < p < 0.02" % (
        rel.max(), want.min(), int(((want > 0.005) & (want < 0.02)).sum())))
    assert rel.max() <= 1e-11
    assert ((want > 0.005) & (want < 0.02)).sum() >= 50
