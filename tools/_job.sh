python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_v2.json 2> gpurun_out/bench_r2_v2.err; tail -c 3000 gpurun_out/bench_r2_v2.json; tail -5 gpurun_out/bench_r2_v2.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_r2_v2_ref.json 2> gpurun_out/bench_r2_v2_ref.err; tail -c 1500 gpurun_out/bench_r2_v2_ref.json; tail -5 gpurun_out/bench_r2_v2_ref.err
