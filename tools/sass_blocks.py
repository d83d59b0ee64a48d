#!/usr/bin/env python
"""Basic blocks of one kernel's SASS with the source lines they come from (static, no GPU needed).

  python tools/sass_blocks.py [walk_permute_kernel] [path/to/lib.so]

For an issue-bound kernel whose control flow is block-uniform (K5: every branch depends on the tree
program and the label bits only) the dynamic instruction count of a walk is  sum over blocks of
(instructions in the block) x (times the tree program sends a thread through it);  tools/k5_model.py
does that sum.  This tool prints the blocks: address range, instruction count, the walk.cuh lines the
instructions are attributed to (-lineinfo) and the closing branch."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def disassemble(lib, kernel):
    """[(address | None, source line | None, text)] of `kernel`; labels come as (None, None, '.L_x_N:')."""
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, stdout=subprocess.DEVNULL)
        cubins = sorted((os.path.getsize(os.path.join(d, f)), f) for f in os.listdir(d) if f.endswith(".cubin"))
        text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubins[-1][1])], check=True, capture_output=True,
                              text=True).stdout
    rows, cur, inside = [], None, False
    for line in text.splitlines():
        if line.startswith("//---") and ".text." in line:
            inside = kernel in line
            continue
        if not inside:
            continue
        m = re.search(r"line (\d+)", line)
        if "//## File" in line and m:
            cur = int(m.group(1))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*);", line)
        if m:
            rows.append((int(m.group(1), 16), cur, m.group(2).strip()))
        elif re.match(r"\.L_x_\d+:", line):
            rows.append((None, None, line.strip()))
    return rows


def blocks(rows):
    out, b = [], []
    for r in rows:
        if r[0] is None:
            if b:
                out.append(b)
            b = [r]
            continue
        b.append(r)
        if re.search(r"\b(BRA|BRX|EXIT|RET)\b", r[2]):
            out.append(b)
            b = []
    if b:
        out.append(b)
    return out


def main():
    kernel = sys.argv[1] if len(sys.argv) > 1 else "walk_permute_kernel"
    lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "scoary_b200", "libscoary_b200.so")
    rows = disassemble(lib, kernel)
    total = 0
    for b in blocks(rows):
        ins = [r for r in b if r[0] is not None]
        if not ins:
            continue
        total += len(ins)
        label = b[0][2] if b[0][0] is None else ""
        lines = sorted(collections.Counter(r[1] for r in ins).items())
        spill = sum(1 for r in ins if re.search(r"\b(STL|LDL)\b", r[2]))
        print("%-10s %06x-%06x n=%4d%s lines=%s  last=%s" % (label, ins[0][0], ins[-1][0], len(ins),
                                                           " spill=%d" % spill if spill else "", lines[:10], ins[-1][2][:44]))
    print("total instructions:", total)


if __name__ == "__main__":
    main()
