#!/bin/bash
# Round capture on the GPU box: tests, smoke, both bench arms, launch list, ncu full capture of the hot kernels, sanitizer.
# Usage: tools/capture_round.sh <tag>   (outputs under gpurun_out/*_<tag>.*; copy what should be judged into profiles/)
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile --no-ref-gpu > gpurun_out/launches_$TAG.log 2>&1
if [ -z "$SKIP_NCU_FULL" ]; then
ncu --set full --clock-control none --import-source on \
    -k regex:"render_fwd|render_bwd|traverse_kernel|shade_fwd_kernel|shade_bwd_kernel|gather_sorted" -s 8 -c 16 \
    -o gpurun_out/prof_$TAG -f python tools/prof_step.py config3 3 > gpurun_out/ncu_$TAG.log 2>&1
fi
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_raster.py -q -x -k "ragged or empty or sync_free or precomputed_colors" 2>&1 | tail -6 > gpurun_out/sanitizer_$TAG.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_trace.py -q -x -k "two_segment or degenerate or fast_path" 2>&1 | tail -6 >> gpurun_out/sanitizer_$TAG.log
ls -la gpurun_out | tail -12
cat gpurun_out/pytest_$TAG.log gpurun_out/smoke_$TAG.log; tail -c 1500 gpurun_out/bench_$TAG.json; tail -4 gpurun_out/sanitizer_$TAG.log
