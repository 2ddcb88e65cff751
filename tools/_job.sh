python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -4
python bench.py --steps 32 --warmup 5 --no-cpu-baseline --no-ref-gpu > gpurun_out/bench_r2_v6.json 2> gpurun_out/bench_r2_v6.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_v6.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"]); print({k:v for k,v in d["kernels_ms"].items() if v>0.15})
PY
tail -3 gpurun_out/bench_r2_v6.err
python tools/prof_step.py config3 3 --stats 2>&1 | grep nodes
