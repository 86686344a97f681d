set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2o_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r2o_pytest_gpu.log
bash tools/ab.sh "build/libwost_v16.so build/libwost_v17.so" cfg5 cfg1b > gpurun_out/r2o_ab.txt 2>&1
