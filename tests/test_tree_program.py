"""The native tree compiler (compile_tree in csrc/engine.cu, behind sb_set_tree) checked on the
CPU: a plain-Python interpreter of the compiled stack program -- the same recurrences the CUDA
kernels run, on unpacked integers -- must reproduce the oracle's PhyloTree restatement for every
tree shape, and the program must respect the 16-bit / 32-bit mode rules the kernels rely on."""
import ctypes

import numpy as np
import pytest

from oracle import oracle as O
from scoary_b200 import _lib, synth
from scoary_b200 import tree as treemod

OPS = {0: "END", 32: "MERGE_POP16", 2: "LEAF_A16", 4: "CHERRY_B16", 6: "CHERRY_B16_MERGE", 38: "CHERRY_B16_MERGE", 40: "PUSH16", 10: "CHERRY_A16",
       42: "PUSH_CHERRY_A16", 16: "WIDEN_A", 17: "LEAF_A32", 18: "MERGE_A32_B16", 19: "PUSH32",
       48: "MERGE_POP32", 49: "MERGE_POPW"}      # csrc/walk.cuh
NEG = -(1 << 30)


WINDOW = 16      # csrc/walk.cuh WALK_WINDOW


def compile_tree(nested, lib=None):
    """(ops, stream, stack units, leaf names): stream[pos] = the leaf consumed at position pos, or -1 for a pad
    position (the default build pads the stream so that no op crosses a 16-leaf window, csrc/walk.cuh SB_WALK_PADDED)"""
    left, right, names = treemod.flatten(nested)
    lib = lib or _lib.load()
    n = len(left)
    ops = np.zeros(6 * n + 64, dtype=np.uint16)
    order = np.full(3 * n + 64, -9, dtype=np.int32)
    n_pos, depth = ctypes.c_int32(), ctypes.c_int32()
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)      # noqa: E731
    k = lib.sb_debug_compile_tree2(ptr(left), ptr(right), n, ptr(ops), len(ops), ptr(order), len(order),
                                   ctypes.byref(n_pos), ctypes.byref(depth))
    assert k > 0, lib.sb_last_error(None)
    return ops[:k].copy(), order[:n_pos.value].copy(), depth.value, names


def leaf_state(g, t, K, ps, po):
    """tip (classes.py:580-592) as key vectors p, a over states AB, Ab, aB, ab, 0"""
    s = (1 - g) * 2 + (1 - t)
    v = [NEG] * 5
    v[s] = 0
    return v[:], v[:]


def combine(L, R, K, dual=True):
    """classes.py:268-572 in max-plus key form (walk.cuh merge_pass), both passes"""
    out = []
    for keys_l, keys_r, bs, bo in ((L[0], R[0], K + 1, K), (L[1], R[1], K, K + 1)):
        ml, mr = max(keys_l), max(keys_r)
        o = [max(keys_l[c] + mr, ml + keys_r[c]) for c in range(4)]
        nf = max(keys_l[4] + keys_r[4], max(keys_l[0] + keys_r[3], keys_l[3] + keys_r[0]) + bs,
                 max(keys_l[1] + keys_r[2], keys_l[2] + keys_r[1]) + bo, NEG)
        out.append(o + [nf])
    return out


def run_program(ops, order, gene_bits, labels, n_leaves):
    shift = 1
    while (1 << shift) <= n_leaves // 2:
        shift += 1
    K = 1 << shift
    pos = 0
    room = 0               # leaves left in the current window, as the kernels count them
    padded = len(order) != n_leaves or bool((order < 0).any())
    A = B = None
    mode_a = None          # 16 or 32: which accumulator form A is in
    stack = []
    sizes = {"A": 0, "B": 0}

    def next_leaf():
        nonlocal pos
        leaf = order[pos]
        assert leaf >= 0, "an op consumed a pad position"
        pos += 1
        return leaf_state(int(gene_bits[leaf]), int(labels[leaf]), K, 0, 0)

    for op in ops.tolist():
        kind, cnt = OPS[op & 63], op >> 6
        if kind == "END":
            break
        need = cnt + 2 if "CHERRY" in kind else cnt if kind in ("LEAF_A16", "LEAF_A32") else 0
        if need:
            # the kernels' window rule: an op that does not fit what is left of the 16-leaf window opens the next
            # one; in a padded stream what it skips is padding, and no op is longer than a window
            if padded:
                assert need <= WINDOW
                if need > room:
                    assert np.all(order[pos:pos + room] == -1)
                    pos += room
                    room = WINDOW
                room -= need
            else:
                room = (room - need) % WINDOW
        if kind in ("CHERRY_A16", "PUSH_CHERRY_A16", "PUSH16"):
            if kind != "CHERRY_A16":
                assert mode_a == 16
                stack.append((A, 16, sizes["A"]))
            if kind == "PUSH16":
                continue
            A = combine(next_leaf(), next_leaf(), K)
            sizes["A"] = 2
            mode_a = 16
            for _ in range(cnt):
                A = combine(A, next_leaf(), K)
                sizes["A"] += 1
        elif kind in ("CHERRY_B16", "CHERRY_B16_MERGE"):
            B = combine(next_leaf(), next_leaf(), K)
            sizes["B"] = 2
            for _ in range(cnt):
                B = combine(B, next_leaf(), K)
                sizes["B"] += 1
            assert sizes["B"] <= 127
            if kind == "CHERRY_B16_MERGE":
                assert mode_a == 16
                A = combine(A, B, K)
                sizes["A"] += sizes["B"]
        elif kind == "LEAF_A16":
            assert mode_a == 16
            for _ in range(cnt):
                A = combine(A, next_leaf(), K)
                sizes["A"] += 1
        elif kind == "MERGE_POP16":
            for _ in range(cnt):
                L, m, sz = stack.pop()
                assert m == 16 and mode_a == 16
                A = combine(L, A, K)
                sizes["A"] += sz
        elif kind == "WIDEN_A":
            assert mode_a == 16
            mode_a = 32
        elif kind == "LEAF_A32":
            assert mode_a == 32
            for _ in range(cnt):
                A = combine(A, next_leaf(), K)
                sizes["A"] += 1
        elif kind == "MERGE_A32_B16":
            assert mode_a == 32
            A = combine(A, B, K)
            sizes["A"] += sizes["B"]
        elif kind == "PUSH32":
            assert mode_a == 32
            stack.append((A, 32, sizes["A"]))
        elif kind == "MERGE_POP32":
            for _ in range(cnt):
                L, m, sz = stack.pop()
                assert m == 32 and mode_a == 32
                A = combine(L, A, K)
                sizes["A"] += sz
        elif kind == "MERGE_POPW":
            for _ in range(cnt):
                L, m, sz = stack.pop()
                assert m == 16 and mode_a == 32
                A = combine(L, A, K)
                sizes["A"] += sz
        # the packed 16-bit form is only legal up to 127 leaves (keys < 4096)
        if mode_a == 16:
            assert sizes["A"] <= 127, (kind, sizes["A"])
    assert np.all(order[pos:] == -1) and not stack and mode_a == 32 and sizes["A"] == n_leaves
    p, a = A
    mask = K - 1
    total = max(p) >> shift
    pro = max([k & mask for k in p if k >= 0] or [-1])
    anti = max([k & mask for k in a if k >= 0] or [-1])
    return total, pro, anti


def _shapes(n, seed):
    names = synth.isolate_names(n)
    comb = names[0]
    for x in names[1:]:
        comb = [comb, x]
    level = list(names)
    while len(level) > 1:
        nxt = [[level[i], level[i + 1]] for i in range(0, len(level) - 1, 2)]
        if len(level) % 2:
            nxt.append(level[-1])
        level = nxt
    return names, [comb, level[0], synth.make_tree(n, seed), synth.make_tree(n, seed + 1)]


@pytest.mark.parametrize("n", [2, 3, 4, 5, 16, 100, 126, 127, 128, 129, 255, 256, 257, 400, 1000])
def test_compiled_program_reproduces_the_oracle(n):
    names, shapes = _shapes(n, 1234 + n)
    rng = np.random.default_rng(n)
    for nested in shapes:
        ops, order, depth, leaf_names = compile_tree(nested)
        assert sorted(order[order >= 0].tolist()) == list(range(n))
        left, right, onames = O.flatten_tree(nested)
        assert onames == leaf_names
        for _ in range(3):
            g = (rng.random(n) < rng.uniform(0.1, 0.9)).astype(np.uint8)
            t = (rng.random(n) < rng.uniform(0.1, 0.9)).astype(np.uint8)
            want = O.phylo_walk(left, right, O.leaf_states(g, t))
            assert run_program(ops, order, g, t, n) == want


def test_stack_depth_is_logarithmic_and_ops_are_fused():
    n = 4096
    names, shapes = _shapes(n, 99)
    for nested, max_units in zip(shapes, (0, 2 * 12, 2 * 12, 2 * 12)):
        ops, order, depth, _ = compile_tree(nested)
        assert depth <= max_units
        assert len(ops) <= n            # fused ops: well under one op per internal node + leaf
    with pytest.raises(AssertionError):
        compile_tree_bad()


def compile_tree_bad():
    lib = _lib.load()
    left = np.asarray([1, ~0], dtype=np.int32)      # node 0 refers to node 1: not children-before-parents
    right = np.asarray([~1, ~2], dtype=np.int32)
    ops = np.zeros(16, dtype=np.uint16)
    order = np.zeros(3, dtype=np.int32)
    k = lib.sb_debug_compile_tree(left.ctypes.data_as(ctypes.c_void_p), right.ctypes.data_as(ctypes.c_void_p), 2,
                                  ops.ctypes.data_as(ctypes.c_void_p), 16, order.ctypes.data_as(ctypes.c_void_p), None)
    assert k > 0, lib.sb_last_error(None)


def test_largest_trees_fit_the_constant_pool():
    """sb_set_tree's limit (include/scoary_b200.h: <= 32 766 isolates): program + at least one label vector of the
    padded stream must fit the 15 872-word constant pool (csrc/walk.cuh C_POOL_WORDS) for every tree shape; the
    balanced tree has the longest program (one op per cherry, push and pop), the random-join tree the longest stream."""
    import sys
    sys.setrecursionlimit(1_000_000)
    n = 32766
    names, shapes = _shapes(n, 5)
    for nested in shapes[:3]:
        ops, order, depth, _ = compile_tree(nested)
        w32p = ((len(order) + 31) // 32 + 3) // 4 * 4
        label_base = ((len(ops) + 1) // 2 + 3) // 4 * 4              # walk_label_base
        assert label_base + w32p <= 15872, (len(ops), len(order))
        assert len(order) <= 1.25 * n + 16 and depth <= 8
