#!/usr/bin/env python
"""Wall clock of the whole command line -- files in, results.csv out -- at a BASELINE shape (run on the GPU box).

  python tools/cli_wall.py --shape c3            # 50 000 genes x 5 000 isolates, -e 1000
  python tools/cli_wall.py --shape c5            # 1 000 000 variants x 2 000 isolates, -e 1000
  python tools/cli_wall.py --shape c3 --genes 5000 --isolates 500 --perms 0      # anything smaller

Writes a synthetic Roary-style gene table (the rows of scoary_b200.synth, cells "1" / "0") and a traits file, runs
scoary_b200.methods.main on them (`-p 1.0 -c I -e P --no-time`: every gene is reported, walked and permuted with the
reference's early-stop rule, as the CLI does) and prints one JSON line: total seconds and the seconds spent in each
stage of main (parse + pack, tree, statistics, pairwise + permutations + writing), next to the GPU kernel time."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from scoary_b200 import synth  # noqa: E402


def write_inputs(d, G, N, seed, chunk=4096):
    traits = synth.make_traits(N, 1, seed)
    names = synth.isolate_names(N)
    from scoary_b200.methods import ROARY_COLUMNS
    gpath, tpath = os.path.join(d, "genes.csv"), os.path.join(d, "traits.csv")
    meta = b',"","synthetic",1,1,1,1,,,,,,,,'              # columns 2 .. 14 of a Roary table
    with open(gpath, "wb") as fh:
        fh.write((",".join('"%s"' % c for c in ROARY_COLUMNS[:14]) + "," + ",".join(names) + "\n").encode())
        for lo in range(0, G, chunk):
            hi = min(G, lo + chunk)
            m = synth.unpack_rows(synth.make_genes_rows(lo, hi, G, N, seed, traits=traits), N)
            cells = np.empty((hi - lo, 2 * N), dtype=np.uint8)
            cells[:, 0::2] = m + 48                         # "1" present, "0" absent (methods.py:476-487)
            cells[:, 1::2] = 44
            cells[:, -1] = 10
            rows = cells.tobytes()
            out = bytearray()
            for r in range(hi - lo):
                out += b'"g%07d"' % (lo + r) + meta + rows[r * 2 * N:(r + 1) * 2 * N]
            fh.write(out)
    with open(tpath, "w") as fh:
        fh.write(",T\n" + "".join("%s,%d\n" % (n, v) for n, v in zip(names, traits[0])))
    return gpath, tpath


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="c3", choices=["c3", "c5"])
    ap.add_argument("--genes", type=int, default=0)
    ap.add_argument("--isolates", type=int, default=0)
    ap.add_argument("--perms", type=int, default=-1)
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    G, N, _, P, seed = synth.CONFIGS[a.shape]
    G, N, P = a.genes or G, a.isolates or N, (a.perms if a.perms >= 0 else P)
    d = tempfile.mkdtemp(prefix="scoary_b200_cli_")
    t0 = time.perf_counter()
    gpath, tpath = write_inputs(d, G, N, seed)
    t_write = time.perf_counter() - t0

    from scoary_b200 import methods as M
    stages = {}

    def timed(name):
        fn = getattr(M, name)

        def wrapper(*args, **kw):
            t = time.perf_counter()
            try:
                return fn(*args, **kw)
            finally:
                stages[name] = stages.get(name, 0.0) + time.perf_counter() - t
        setattr(M, name, wrapper)
    for name in ("Csv_to_dic_Roary", "upgma", "Csv_to_dic", "Setup_results", "StoreResults"):
        timed(name)
    argv = ["-g", gpath, "-t", tpath, "-o", os.path.join(d, "out"), "--no-time", "-p", "1.0", "-c", "I"]
    if P >= 10:
        argv += ["-e", str(P)]
    try:
        M.get_engine().set_profiling(True)
    except Exception as ex:       # no GPU: main() will say so
        print("engine:", ex, file=sys.stderr)
    t0 = time.perf_counter()
    try:
        M.main(argv=argv)
    except SystemExit as ex:
        if ex.code not in (0, None):
            raise
    wall = time.perf_counter() - t0
    st = M._ENGINE.stats() if getattr(M, "_ENGINE", None) is not None and hasattr(M._ENGINE, "stats") else {}
    res = os.path.join(d, "out", "T.results.csv")
    line = {"shape": "%s: %d genes x %d isolates, -e %d, -p 1.0 -c I" % (a.shape, G, N, P), "cli_wall_s": wall,
            "stages_s": {k: round(v, 3) for k, v in stages.items()},
            "input_bytes": os.path.getsize(gpath), "result_rows": sum(1 for _ in open(res)) - 1 if os.path.exists(res) else None,
            "gpu_kernel_ms": {k: round(v, 3) for k, v in st.items() if k.startswith("ms_")},
            "gpu_launches": st.get("kernel_launches"), "input_write_s": round(t_write, 2)}
    print(json.dumps(line))
    if not a.keep:
        import shutil
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
