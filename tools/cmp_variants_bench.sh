#!/usr/bin/env bash
# The north_star step with the product library and with variant libraries built by tools/sweep_variants.py (run on
# the GPU box):   bash tools/cmp_variants_bench.sh <tag> <variant> [<variant> ...]
set -u
tag="${1:-cmp}"; shift || true
cd "$(dirname "$0")/.."
mkdir -p "gpurun_out/$tag"
for v in base "$@"; do
  if [ "$v" = base ]; then lib=scoary_b200/libscoary_b200.so; else lib=variants/$v.so; fi
  [ -f "$lib" ] || { echo "$v: $lib not built"; continue; }
  SCOARY_B200_LIB=$PWD/$lib timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cli-wall \
      > "gpurun_out/$tag/bench_$v.log" 2> "gpurun_out/$tag/bench_$v.err"
  python - "$v" "gpurun_out/$tag/bench_$v.log" <<'P'
import json, sys
for l in open(sys.argv[2]):
    if l.startswith("{"):
        d = json.loads(l)
        print(sys.argv[1], "value %.4g tests/s, %.1f ms/step, e2e %.4g, reference-rule mode %.1f ms" % (
            d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("reference_rule_mode") or {}).get("ms_per_step", float("nan"))))
P
done
