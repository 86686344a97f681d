set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2d_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r2d_pytest_gpu.log
bash tools/ab.sh "build/libwost_base.so build/libwost_new.so" cfg5 cfg5_175e cfg4 cfg2 cfg1b cfg1a cfg3 > gpurun_out/r2d_ab.txt 2>&1
for rep in 1 2; do echo -n "lockstep "; WOST_LIB=build/libwost_new.so python tools/run_one.py cfg1b 4 | tail -1; echo -n "no-lockstep "; WOST_JIT_OPTS=-DWOST_NO_LOCKSTEP=1 WOST_LIB=build/libwost_new.so python tools/run_one.py cfg1b 4 | tail -1; done > gpurun_out/r2d_lockstep.txt 2>&1
for L in build/libwost_base.so build/libwost_new.so; do echo "== $L"; WOST_LIB=$L python tools/survey_rank_job.py 16384; done > gpurun_out/r2d_survey_rank.txt 2>&1
echo "== new, WOST_SOURCE_GRID=0" >> gpurun_out/r2d_survey_rank.txt; WOST_SOURCE_GRID=0 WOST_LIB=build/libwost_new.so python tools/survey_rank_job.py 16384 >> gpurun_out/r2d_survey_rank.txt 2>&1
WOST_LIB=build/libwost_new.so python tools/survey_bench.py > gpurun_out/r2d_survey_bench.txt 2>&1
