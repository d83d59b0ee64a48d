#!/usr/bin/env bash
# Last GPU call of a round: parity tests, one ncu capture of K5 (instruction count for profiles/), the bench line.
#   gpurun --timeout 400 -- 'bash tools/final_check.sh r2m'
set -u
out="gpurun_out/${1:-final}"
mkdir -p "$out"
cd "$(dirname "$0")/.."
timeout 120 python -m pytest tests -m gpu -x -q > "$out/pytest_gpu.log" 2>&1; tail -n 2 "$out/pytest_gpu.log"
timeout 100 python bench.py --steps 3 --no-cpu-baseline --no-cli-wall > "$out/bench.log" 2> "$out/bench.err"; cut -c1-220 "$out/bench.log" | tail -n 1
timeout 100 ncu --set full --clock-control none --import-source on -k regex:walk_permute -c 1 -o "$out/prof_walk" -f \
    python tools/probe.py --perms 60 > "$out/ncu_walk.log" 2>&1
[ -f "$out/prof_walk.ncu-rep" ] && ncu -i "$out/prof_walk.ncu-rep" --page raw --csv > "$out/prof_walk.raw.csv" 2>/dev/null
[ -f "$out/prof_walk.ncu-rep" ] && ncu -i "$out/prof_walk.ncu-rep" --page source --csv --print-source sass > "$out/prof_walk.source.csv" 2>/dev/null
rm -f "$out/prof_walk.ncu-rep"
echo done
