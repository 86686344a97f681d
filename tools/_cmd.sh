timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "test_estimates_within_3_sigma_of_reference" > gpurun_out/r2w_pytest.log 2>&1; tail -3 gpurun_out/r2w_pytest.log
