#!/usr/bin/env bash
# One GPU call that answers everything a kernel change needs answered (run under gpurun, one GPU):
#
#   python tools/sweep_variants.py build                      # here, no GPU: variants/*.so travel with the snapshot
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh r2a'  # tag names the artefacts under gpurun_out/
#
# 1. pytest -m gpu                    parity first: nothing below matters if it is red
# 2. tools/sweep_variants.py run      every variant library on the C3 shape, with a checksum against the product library
# 3. ncu launch list                  of `bench.py --steps 2 --warmup 3` (kernel shares of a step)
# 4. ncu --set full                   one walk_permute_kernel launch and one fisher_kernel launch of tools/probe.py
# 5. bench.py                         the number itself (never taken under a profiler), clocks sampled by bench.py
# Each step has its own timeout so a hang cannot eat the box; later steps still run when an earlier one fails.
set -u
tag="${1:-session}"
out="gpurun_out/${tag}"
mkdir -p "$out"
cd "$(dirname "$0")/.."

step() {   # step <seconds> <log> <command...>
    local t="$1" log="$2"; shift 2
    echo "== $* (limit ${t}s)" | tee -a "$out/session.log"
    timeout "$t" "$@" > "$out/$log" 2>&1
    echo "   exit $? ; tail:" | tee -a "$out/session.log"
    tail -n 6 "$out/$log" | tee -a "$out/session.log"
}

mode="${2:-full}"          # quick: parity tests and the bench lines only; walk: everything but the Fisher capture
step 900 pytest_gpu.log python -m pytest tests -m gpu -x -q
if [ "$mode" != quick ]; then
step 1200 sweep.log python tools/sweep_variants.py run
step 300 launches.log ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file "$out/launches.csv" python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline
# the same command without ncu's cache flush between kernels: DRAM bytes and executed instructions of every launch as
# they are inside a running step (the --set full capture below is cold-cache: its dram bytes re-read the gene matrix)
step 300 step_dram.log ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum \
    --cache-control none --clock-control none -c 400 --csv \
    --log-file "$out/step_dram.csv" python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline
step 420 ncu_walk.log ncu --set full --clock-control none --import-source on -k regex:walk_permute -c 1 \
    -o "$out/prof_walk" -f python tools/probe.py --perms 60
if [ "$mode" != walk ]; then     # walk: the Fisher kernel did not change, keep its capture
step 300 ncu_fisher.log ncu --set full --clock-control none --import-source on -k regex:fisher_kernel -c 1 \
    -o "$out/prof_fisher" -f python tools/probe.py --perms 4
fi
for rep in "$out"/prof_walk.ncu-rep "$out"/prof_fisher.ncu-rep; do
    [ -f "$rep" ] && ncu -i "$rep" --page raw --csv > "${rep%.ncu-rep}.raw.csv" 2>/dev/null
    # per-instruction executed counts and stall samples (SASS view): settles the pipe assignment of each opcode
    [ -f "$rep" ] && ncu -i "$rep" --page source --csv --print-source sass > "${rep%.ncu-rep}.source.csv" 2>/dev/null
done
fi
step 600 bench.log python bench.py
step 300 bench_c3.log python bench.py --workload c3 --steps 5 --no-cpu-baseline
if [ "$mode" != walk ]; then
# one GPU's share of the north_star job on 8 GPUs (6 250 genes): what strong scaling asks of the launch planner
step 300 bench_shard8.log python bench.py --genes 6250 --steps 3 --no-cpu-baseline
fi
grep -h '^{' "$out/bench.log" "$out/bench_c3.log" "$out/bench_shard8.log" > "$out/bench_lines.json" 2>/dev/null
echo "== done" | tee -a "$out/session.log"
