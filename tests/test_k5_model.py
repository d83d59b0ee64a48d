"""tools/sass_emul.py + tools/k5_model.py: the static instruction count of the walk kernels (profiles/
r1_k5_instruction_model.md).  Needs cuobjdump / nvdisasm (CUDA toolkit) but no GPU."""
import os
import shutil
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif(not (shutil.which("nvdisasm") and shutil.which("cuobjdump")),
                                reason="CUDA binary utilities not available")


def _count(kernel, n_isolates, perms, ppi, chunk):
    import k5_model
    import sass_emul
    lib = os.path.join(ROOT, "scoary_b200", "libscoary_b200.so")
    ops, labelsW, W32p, shift = k5_model.workload(n_isolates, 4242, perms, 7)
    base = k5_model.label_base(len(ops))
    const3 = bytearray(4 * base + 4 * labelsW.size)
    const3[:2 * len(ops)] = ops.astype("<u2").tobytes()
    const3[4 * base:] = labelsW.astype("<u4").tobytes()
    instrs, labels, const2 = sass_emul.extract(lib, kernel)
    params = sass_emul.walk_args(1024, 1000, 1000, W32p, shift, perms, ppi, lab_base=base)
    m = sass_emul.Machine(instrs, labels, const2, params, bytes(const3), tid=0, ctaid=(0, chunk, 0))
    return m.run(max_steps=2_000_000), m, instrs


def test_interpreter_follows_the_permutation_kernel_to_its_exit():
    """every branch of K5 resolves from block-uniform data; the count scales with the labellings walked"""
    one, m1, _ = _count("walk_permute_kernelILb0", 300, 4, 1, 0)
    two, m2, instrs = _count("walk_permute_kernelILb0", 300, 4, 2, 1)
    assert len(m1.assumed) <= 4 and all(v == 1 for v in m1.assumed.values())      # the four guarded result stores
    per_walk = two - one                                                          # second labelling of the block
    assert 299 * 25 < per_walk < 299 * 90                                         # 4 genes per thread: ~55 per node
    stores = [i for i, ins in enumerate(instrs) if ins.op == "STG"]
    assert stores and all(m2.counts[i] == 1 for i in stores)


def test_interpreter_follows_the_pairs_kernel():
    n, m, instrs = _count("walk_pairs_kernel", 300, 1, 1, 0)
    assert 299 * 40 < n < 299 * 200                                               # both key sets, same program
    assert m.counts[max(i for i, ins in enumerate(instrs) if ins.op == "EXIT")] == 1      # left through the last EXIT
