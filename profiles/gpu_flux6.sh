timeout 300 python -m pytest tests/test_gpu_flux.py -x -q 2>&1 | tail -3; bash profiles/gpu_flux5.sh
