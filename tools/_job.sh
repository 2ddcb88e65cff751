python -m pytest tests/test_gpu_golden.py tests/test_gpu_reference.py tests/test_gpu_raster.py -q --tb=line 2>&1 | grep -v "^$" | tail -14
timeout 600 python tools/optix_probe.py --time > gpurun_out/optix_probe3.log 2>&1; tail -12 gpurun_out/optix_probe3.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_v1.json 2> gpurun_out/bench_r2_v1.err; tail -c 1500 gpurun_out/bench_r2_v1.json
