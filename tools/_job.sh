python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -12
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu > gpurun_out/bench_r2_v3.json 2> gpurun_out/bench_r2_v3.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_v3.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]); print({k:v for k,v in d["kernels_ms"].items() if v>0.15})
PY
tail -3 gpurun_out/bench_r2_v3.err
