python oracle/make_golden.py --trace-only --only=trace_ref_two_segments_stop 2>&1 | tail -2
cp gpurun_out/golden/trace_ref_two_segments_stop.pt tests/golden/
python -m pytest tests/test_gpu_golden.py tests/test_oracle_golden.py -q --tb=line -k "stop" 2>&1 | tail -8
