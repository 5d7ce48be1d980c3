timeout 1500 python -m pytest tests/test_gpu_edge_cases.py -q > gpurun_out/pytest_edge.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_edge.log | cut -c1-220
