set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2i_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r2i_pytest_gpu.log
( time WOST_JIT=1 timeout 900 python -m pytest tests -m gpu -q -k "not jit" ) > gpurun_out/r2i_pytest_gpu_jit1.log 2>&1; tail -n 3 gpurun_out/r2i_pytest_gpu_jit1.log
( time WOST_JIT=0 timeout 900 python -m pytest tests -m gpu -q -k "not jit" ) > gpurun_out/r2i_pytest_gpu_jit0.log 2>&1; tail -n 3 gpurun_out/r2i_pytest_gpu_jit0.log
( time timeout 900 python bench.py ) > gpurun_out/r2i_bench1.json 2> gpurun_out/r2i_bench1.err; echo rc=$?; tail -c 300 gpurun_out/r2i_bench1.err
python tools/survey_rank_job.py 16384 > gpurun_out/r2i_survey_rank.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.txt 2>&1; tail -2 gpurun_out/r2i_smoke.txt
