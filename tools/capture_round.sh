#!/bin/bash
# Round capture on the GPU box: tests, bench, launch list, ncu full capture of the hot kernels.  Usage: tools/capture_round.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python tools/ramp_diag.py nosync > gpurun_out/ramp_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"render_fwd|render_bwd|traverse_kernel|shade_fwd_kernel|shade_bwd_kernel|preprocess_kernel|preprocess_bwd|sh_features" -s 24 -c 16 \
    -o gpurun_out/prof_$TAG -f python tools/prof_step.py config3 3 > gpurun_out/ncu_$TAG.log 2>&1
ls -la gpurun_out | tail -20
cat gpurun_out/pytest_$TAG.log gpurun_out/smoke_$TAG.log gpurun_out/bench_$TAG.json; tail -8 gpurun_out/ramp_$TAG.log
