#!/usr/bin/env python
"""Warp instructions per walk of walk_permute_kernel (K5), computed here (no GPU) by running the
kernel's SASS control flow (tools/sass_emul.py) on the real tree program and the real label vectors.

  python tools/k5_model.py [--lib scoary_b200/libscoary_b200.so] [--isolates 5000] [--seed 20260903]
                           [--perms 60] [--ppi 2] [--chunks 0,7,15] [--json out.json]

Prints the executed warp instructions of one block-warp per chunk of `ppi` labellings and the average
per 64 (gene, labelling) walks -- the unit of profiles/k5_warp_instructions.json (ncu:
smsp__inst_executed.sum / (genes x permutations / 64)).  The labellings are the ones the engine draws
for `sb_permute(seed=1)` (same Philox stream as the oracle), so for the profiled launch the count can be
compared with ncu directly."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.setrecursionlimit(1_000_000)

import numpy as np  # noqa: E402

import sass_emul  # noqa: E402

def label_base(n_ops):
    """csrc/walk.cuh walk_label_base: first label word behind the program in the constant pool (bank 3)"""
    return ((n_ops + 1) // 2 + 3) // 4 * 4


def workload(n_isolates, seed, n_perms, perm_seed):
    """tree program + label vectors in walk order, exactly as sb_set_tree / sb_permute stage them
    (the tree compiler is the one of the library named by SCOARY_B200_LIB, so variants get their own program)"""
    import ctypes
    from oracle import oracle as O
    from scoary_b200 import _lib, synth
    from scoary_b200 import tree as treemod
    nested = synth.make_tree(n_isolates, seed)
    left, right, names = treemod.flatten(nested)
    lib = _lib.load()
    n = len(left)
    ops = np.zeros(6 * n + 64, dtype=np.uint16)
    order = np.full(3 * n + 64, -9, dtype=np.int32)
    n_pos, depth = ctypes.c_int32(), ctypes.c_int32()
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)      # noqa: E731
    k = lib.sb_debug_compile_tree2(ptr(left), ptr(right), n, ptr(ops), len(ops), ptr(order), len(order),
                                   ctypes.byref(n_pos), ctypes.byref(depth))
    assert k > 0, lib.sb_last_error(None)
    ops, order = ops[:k], order[:n_pos.value]              # order: leaf id per stream position, -1 = pad position
    col = {nm: j for j, nm in enumerate(synth.isolate_names(n_isolates))}
    traits = synth.make_traits(n_isolates, 1, seed)
    lab = (traits[0][np.asarray([col[nm] for nm in names])] == 1).astype(np.uint8)      # by leaf id
    W32p = ((len(order) + 31) // 32 + 3) // 4 * 4
    labs = np.stack([O.shuffle_labels(perm_seed, 0, p, lab) for p in range(n_perms)])
    walk = np.zeros((n_perms, W32p * 32), dtype=np.uint8)
    real = order >= 0
    walk[:, np.nonzero(real)[0]] = labs[:, order[real]]
    labelsW = np.ascontiguousarray(np.packbits(walk, axis=1, bitorder="little")).view(np.uint32).reshape(n_perms, W32p)
    shift = 1
    while (1 << shift) <= n_isolates // 2:
        shift += 1
    return np.ascontiguousarray(ops), labelsW, W32p, shift


def pipe_of(op):
    """issue pipe of a SASS opcode as ncu groups them (sm__inst_executed_pipe_alu / fma / uniform / lsu / ...)"""
    if op.startswith("U") or op in ("R2UR", "S2UR"):
        return "uniform"
    if op in ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IMUL"):
        return "fma"
    if op in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "LD", "ST", "ATOM", "RED"):
        return "lsu"
    if op in ("BRA", "BRX", "EXIT", "BSSY", "BSYNC", "RET", "CALL", "WARPSYNC", "NOP", "BAR"):
        return "branch/control"
    if op in ("LDC", "S2R", "CS2R"):
        return "other (LDC, S2R)"
    if op in ("BREV", "POPC", "FLO", "MUFU"):
        return "xu"
    return "alu"          # LOP3, SHF, SEL, PRMT, ISETP, IADD3, LEA, VIADD, VIADDMNMX, VIMNMX(3), PLOP3, MOV


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "scoary_b200", "libscoary_b200.so"))
    ap.add_argument("--kernel", default="walk_permute_kernelILb0")
    ap.add_argument("--isolates", type=int, default=5000)
    ap.add_argument("--genes", type=int, default=50000)
    ap.add_argument("--seed", type=int, default=20260903)
    ap.add_argument("--perm-seed", type=int, default=1)
    ap.add_argument("--perms", type=int, default=60, help="labellings staged for the launch")
    ap.add_argument("--ppi", type=int, default=2, help="labellings per block")
    ap.add_argument("--chunks", default="all")
    ap.add_argument("--genes-per-thread", type=int, default=4)
    ap.add_argument("--threads", type=int, default=128)
    ap.add_argument("--json")
    ap.add_argument("--profile", type=int, default=0, help="print the N basic blocks with the most executed instructions")
    a = ap.parse_args()
    os.environ["SCOARY_B200_LIB"] = os.path.abspath(a.lib)      # the tree compiler of the library being modelled

    ops, labelsW, W32p, shift = workload(a.isolates, a.seed, a.perms, a.perm_seed)
    base = label_base(len(ops))
    const3 = bytearray(4 * base + 4 * labelsW.size)
    const3[:2 * len(ops)] = ops.astype("<u2").tobytes()
    const3[4 * base:] = labelsW.astype("<u4").tobytes()
    instrs, labels, const2 = sass_emul.extract(a.lib, a.kernel)
    Gs = (a.genes + 31) // 32 * 32
    params = sass_emul.walk_args(Gs, a.genes, a.genes, W32p, shift, a.perms, a.ppi, lab_base=base)
    n_chunks = (a.perms + a.ppi - 1) // a.ppi
    chunks = range(n_chunks) if a.chunks == "all" else [int(c) for c in a.chunks.split(",")]
    per_chunk = {}
    dyn = [0] * len(instrs)
    for c in chunks:
        m = sass_emul.Machine(instrs, labels, const2, params, bytes(const3), tid=0, ctaid=(0, c, 0))
        per_chunk[c] = m.run()
        dyn = [x + y for x, y in zip(dyn, m.counts)]
        if m.assumed:
            print("   per-thread branches assumed not taken:", {hex(k): v for k, v in m.assumed.items()})
        print("chunk %3d: %d warp instructions for %d labellings" % (c, per_chunk[c], min(a.ppi, a.perms - c * a.ppi)))
    walks_per_warp = 32 * a.genes_per_thread
    total_instr = sum(per_chunk.values())
    total_walks = sum(min(a.ppi, a.perms - c * a.ppi) for c in chunks) * walks_per_warp
    per64 = total_instr / total_walks * 64
    print("static instructions in the kernel: %d" % len(instrs))
    print("warp instructions per 64 walks: %.0f   (per internal node and 4-gene thread: %.1f)" % (
        per64, per64 * 2 / max(a.isolates - 1, 1)))
    # executed instructions by issue pipe (ncu: sm__inst_executed_pipe_*), same unit as per64
    pipes = {}
    for ins, n in zip(instrs, dyn):
        pipes[pipe_of(ins.op)] = pipes.get(pipe_of(ins.op), 0) + n
    scale = 64.0 / total_walks
    print("by pipe, per 64 walks: " + ", ".join("%s %.0f" % (k, v * scale) for k, v in sorted(pipes.items(), key=lambda kv: -kv[1])))
    # dynamic instruction footprint: bytes of 128-byte I-cache lines that cover a share of the executed instructions
    lines = {}
    for ins, n in zip(instrs, dyn):
        lines[ins.addr // 128] = lines.get(ins.addr // 128, 0) + n
    acc, cover, want = 0, {}, [0.5, 0.8, 0.9, 0.99]
    for k, n in enumerate(sorted(lines.values(), reverse=True)):
        acc += n
        while want and acc >= want[0] * total_instr:
            cover[want.pop(0)] = (k + 1) * 128
    print("code bytes covering 50 / 80 / 90 / 99 %% of the executed instructions: %s (static %d)" % (
        " / ".join(str(cover.get(q, 0)) for q in (0.5, 0.8, 0.9, 0.99)), 16 * len(instrs)))
    if a.profile:
        import sass_blocks
        count_at = {ins.addr: n for ins, n in zip(instrs, dyn)}
        rows = []
        for b in sass_blocks.blocks(sass_blocks.disassemble(a.lib, a.kernel)):
            ins = [r for r in b if r[0] is not None]
            if ins:
                lines = sorted(set(r[1] for r in ins if r[1]))
                rows.append((sum(count_at.get(r[0], 0) for r in ins), ins[0][0], len(ins), count_at.get(ins[0][0], 0), lines))
        total = sum(r[0] for r in rows)
        print("  share  executed  block  (instructions x entries)  walk.cuh lines")
        for r in sorted(rows, reverse=True)[:a.profile]:
            print("  %5.1f%% %9d  %06x  (%3d x %6d)  %s" % (100.0 * r[0] / total, r[0], r[1], r[2], r[3], r[4][:8]))
    if a.json:
        with open(a.json, "w") as f:
            json.dump({"lib": os.path.relpath(a.lib, ROOT), "kernel": a.kernel, "isolates": a.isolates, "perms": a.perms,
                       "ppi": a.ppi, "chunks": {str(k): v for k, v in per_chunk.items()}, "per_64_walks": per64}, f)
    return per64


if __name__ == "__main__":
    main()
