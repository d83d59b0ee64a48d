mkdir -p gpurun_out/r2i
for v in base t128_mb6 t256_mb3; do
  if [ $v = base ]; then lib=scoary_b200/libscoary_b200.so; else lib=variants/$v.so; fi
  SCOARY_B200_LIB=$PWD/$lib timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cli-wall > gpurun_out/r2i/bench_$v.log 2>gpurun_out/r2i/bench_$v.err
  python - <<P
import json
for l in open("gpurun_out/r2i/bench_$v.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$v", d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("reference_rule_mode") or {}).get("ms_per_step"))
P
done
