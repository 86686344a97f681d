set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2m_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r2m_pytest_gpu.log
bash tools/ab.sh "build/libwost_v15.so build/libwost_v16.so" cfg5 cfg4 cfg5_175e > gpurun_out/r2m_ab.txt 2>&1
