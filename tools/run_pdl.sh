# PDL experiment (one GPU): on/off bit-identity tests, the whole GPU suite under mode 2, bench lines for modes 0/1/2
set -x
timeout 300 python -m pytest tests/test_gpu_pdl.py -x -q 2>&1 | tail -15
PFC_PDL=2 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "passed|failed|rror" | tail -5
for m in 0 1 2 0 2; do
  timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --pdl $m > gpurun_out/bench_pdl$m.json 2> gpurun_out/bench_pdl$m.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_pdl$m.json").read().strip().splitlines()[-1])
    print("pdl", $m, "ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["config"].get("launch"), d["clocks"])
except Exception as e:
    print("pdl", $m, "FAILED", e); print(open("gpurun_out/bench_pdl$m.err").read()[-1500:])
PY
done
