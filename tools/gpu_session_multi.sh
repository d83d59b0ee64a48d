#!/usr/bin/env bash
# One multi-GPU call (gpurun --gpus 8): the N > 1 command line on real GPUs, then bench.py at N = 1, 2, 4, 8 on the fixed
# north_star job (strong scaling) and the 8-GPU configurations C4 and C5 through scoary_b200.distributed.
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_session_multi.sh r2m'
set -u
tag="${1:-multi}"
out="gpurun_out/${tag}"
mkdir -p "$out"
cd "$(dirname "$0")/.."
ngpu=$(nvidia-smi -L | wc -l)
echo "GPUs: $ngpu" | tee "$out/session.log"

step() {   # step <seconds> <log> <command...>
    local t="$1" log="$2"; shift 2
    echo "== $* (limit ${t}s)" | tee -a "$out/session.log"
    timeout "$t" "$@" > "$out/$log" 2>&1
    echo "   exit $? ; tail:" | tee -a "$out/session.log"
    tail -n 4 "$out/$log" | cut -c1-600 | tee -a "$out/session.log"
}
if [ "${SKIP_TESTS:-0}" != 1 ]; then
    step 900 pytest_gpu.log python -m pytest tests -m gpu -x -q -rs
    step 900 cli_wall_c3.log python tools/cli_wall.py --shape c3
fi
step 600 cli_two_gpus.log python -m pytest tests/test_gpu_cli_more.py -m gpu -q -k two_gpus -rs
for n in 1 2 4 8; do
    [ "$n" -le "$ngpu" ] || continue
    step 400 "bench_north_star_n$n.log" bash tools/run_bench_n.sh "$n" $((29500 + n)) --steps 5 --warmup 3 $([ "$n" = 1 ] || echo --no-cpu-baseline)
done
if [ "$ngpu" -ge 8 ]; then
    step 400 bench_c4_n8.log bash tools/run_bench_n.sh 8 29611 --workload c4 --steps 2 --warmup 3 --no-cpu-baseline
    step 400 bench_c5_n8.log bash tools/run_bench_n.sh 8 29612 --workload c5 --steps 3 --warmup 3 --no-cpu-baseline
    step 400 bench_c3_n8.log bash tools/run_bench_n.sh 8 29613 --workload c3 --steps 5 --warmup 3 --no-cpu-baseline
    # the split the planner did NOT choose for the north_star job (6 250-gene shards), for comparison
    step 400 bench_north_star_n8_genes.log bash tools/run_bench_n.sh 8 29614 --split genes --steps 5 --warmup 3 --no-cpu-baseline
fi
grep -h '^{' "$out"/bench_*.log > "$out/bench_lines.json" 2>/dev/null
echo "== done" | tee -a "$out/session.log"
