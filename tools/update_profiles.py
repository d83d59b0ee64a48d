#!/usr/bin/env python
"""Refresh the numbers bench.py quotes from ncu (profiles/k5_warp_instructions.json, profiles/k5_dram_bytes.json) and the
per-kernel summaries from one tools/gpu_session.sh run.

  python tools/update_profiles.py gpurun_out/r2h <commit> r2

Reads prof_walk.raw.csv (ncu --set full, one walk_permute_kernel launch of tools/probe.py at the C3 shape: 50 000 genes,
grid = tiles x labellings), step_dram.csv (ncu --cache-control none over bench.py --workload c3: DRAM bytes and executed
instructions of every launch inside running steps), prof_fisher.raw.csv."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, TILE = 50000, 768


def raw(path):
    rows = list(csv.reader(open(path)))
    return dict(zip(rows[0], rows[2]))


def main():
    sess, commit, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    w = raw(os.path.join(sess, "prof_walk.raw.csv"))
    grid, block = int(float(w["launch__grid_size"])), int(float(w["launch__block_size"]))
    tiles = -(-G // (block * 4))
    n_lab = grid // tiles
    inst = float(w["smsp__inst_executed.sum"])
    per64 = inst / (G * n_lab / 64.0)
    wi = {"_comment": "warp instructions executed per 64 (gene, labelling) walks of walk_permute_kernel<false> at N = 5000 "
                      "(random-join tree, 4 genes per thread): ncu smsp__inst_executed.sum of ONE launch / (50 000 genes x "
                      "labellings of that launch / 64)",
          "source": "ncu --set full --clock-control none, session %s, kernels of commit %s (profiles/%s_ncu_walk_permute_%s.txt)"
                    % (os.path.basename(sess), commit, tag, commit),
          "c3": round(per64), "north_star": round(per64),
          "launch": {"grid": grid, "block": block, "labellings": n_lab, "inst_executed": inst,
                     "duration_ms": float(w["gpu__time_duration.sum"]),
                     "issue_slots_busy_pct": float(w["smsp__issue_active.avg.pct_of_peak_sustained_active"])},
          "by_pipe_c3": {"alu_pct_of_peak": float(w["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]),
                         "fma_pct_of_peak": float(w["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]),
                         "uniform_pct_of_peak": float(w["sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"]),
                         "lsu_pct_of_peak": float(w["sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"])},
          "history": {"5af1521 (round 1, profiled)": 152363, "630daac (round 1 HEAD, ncu in round 2)": 133854,
                      "d0be99b (round 2, session r2h)": 131749, "8d63fcb (round 2, session r2j)": 110681,
                      "3a850a5 (round 2, session r2l)": 108611}}
    wi["by_pipe_north_star"] = wi["by_pipe_c3"]
    json.dump(wi, open(os.path.join(ROOT, "profiles", "k5_warp_instructions.json"), "w"), indent=1)
    if not os.path.exists(os.path.join(sess, "step_dram.csv")):      # a tools/final_check.sh session: K5 capture only
        out = os.path.join(ROOT, "profiles", "%s_ncu_walk_permute_%s.txt" % (tag, commit))
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"),
                              os.path.join(sess, "prof_walk.raw.csv"), os.path.join(sess, "prof_walk.source.csv")],
                             capture_output=True, text=True, check=True).stdout
        open(out, "w").write("# ncu --set full --clock-control none --import-source on, one launch of tools/probe.py (C3 shape), "
                             "kernels of commit %s, session %s\n" % (commit, os.path.basename(sess)) + txt)
        print(json.dumps(wi["launch"]), per64)
        return
    # DRAM traffic of the full-size exhaustive launches inside running steps
    rows = list(csv.reader(l for l in open(os.path.join(sess, "step_dram.csv")) if not l.startswith("==")))
    hdr = rows[0]
    ik, im, iv, ii, ig = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size"))
    per = collections.defaultdict(dict)
    for r in rows[1:]:
        if len(r) >= len(hdr) and "walk_permute" in r[ik]:
            per[(r[ii], r[ig])][r[im]] = float(r[iv].replace(",", ""))
    full = [m for (i, g), m in per.items() if int(g.strip("() ").split(",")[0]) == tiles and "dram__bytes_read.sum" in m]
    big = max(int(g.strip("() ").split(",")[1]) for (i, g) in per if int(g.strip("() ").split(",")[0]) == tiles)
    full = [m for (i, g), m in per.items() if g.strip("() ").split(",")[:2] == [str(tiles), " %d" % big] or
            [x.strip() for x in g.strip("() ").split(",")][:2] == [str(tiles), str(big)]]
    rd = sum(m["dram__bytes_read.sum"] for m in full) / len(full)
    wr = sum(m["dram__bytes_write.sum"] for m in full) / len(full)
    dram = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per walk_permute_kernel launch INSIDE running steps "
                        "(ncu --cache-control none --metrics ... over bench.py --workload c3: no cache flush between "
                        "kernels), averaged over the %d full-size launches (%d x %d blocks) of that run; algorithmic bytes "
                        "of such a launch: 0.65 B x 50 000 genes x %d labellings" % (len(full), tiles, big, big),
            "source": "session %s, commit %s, profiles/%s_step_dram_%s.csv" % (os.path.basename(sess), commit, tag, commit),
            "c3": round(rd + wr), "north_star": round(rd + wr), "read": round(rd), "write": round(wr),
            "cold_cache_single_launch": round((float(w["dram__bytes_read.sum"]) + float(w["dram__bytes_write.sum"])) * 1e6)}
    json.dump(dram, open(os.path.join(ROOT, "profiles", "k5_dram_bytes.json"), "w"), indent=1)
    for k, name in (("walk", "walk_permute"), ("fisher", "fisher")):
        if not os.path.exists(os.path.join(sess, "prof_%s.raw.csv" % k)):
            continue
        out = os.path.join(ROOT, "profiles", "%s_ncu_%s_%s.txt" % (tag, name, commit))
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"),
                              os.path.join(sess, "prof_%s.raw.csv" % k), os.path.join(sess, "prof_%s.source.csv" % k)],
                             capture_output=True, text=True, check=True).stdout
        open(out, "w").write("# ncu --set full --clock-control none --import-source on, one launch of tools/probe.py (C3 shape), "
                             "kernels of commit %s, session %s\n" % (commit, os.path.basename(sess)) + txt)
    for f, dst in (("launches.csv", "%s_launches_bench_c3_%s.csv" % (tag, commit)), ("step_dram.csv", "%s_step_dram_%s.csv" % (tag, commit))):
        open(os.path.join(ROOT, "profiles", dst), "w").write(open(os.path.join(sess, f)).read())
    print(json.dumps(wi["launch"]), per64, dram["c3"])


if __name__ == "__main__":
    main()
