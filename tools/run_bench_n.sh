#!/usr/bin/env bash
# run_bench_n.sh <n> <port> <bench.py args...>: bench.py on n GPUs the way the driver launches it
n="$1"; port="$2"; shift 2
cd "$(dirname "$0")/.."
if [ "$n" = 1 ]; then exec python bench.py --gpus 1 "$@"
else exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port "$port" bench.py --gpus "$n" "$@"; fi
