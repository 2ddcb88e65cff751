python -m pytest tests/test_gpu_densify.py -q --tb=short 2>&1 | tail -30
