python -m pytest tests -m gpu -q --tb=short -x -k "trace or tracer or pipeline or golden" 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu > gpurun_out/bench_r2_v4.json 2> gpurun_out/bench_r2_v4.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_v4.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]); print({k:v for k,v in d["kernels_ms"].items() if v>0.15}); print(d["roofline"].get("composited_hits"))
PY
tail -3 gpurun_out/bench_r2_v4.err
