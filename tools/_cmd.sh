set -x
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2t_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r2t_pytest_gpu.log
for s in cfg2 cfg4; do python tools/run_one.py $s 5 | tail -1; done > gpurun_out/r2t_run.txt 2>&1
