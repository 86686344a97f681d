set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r2f_pytest_gpu.log
bash tools/ab.sh "build/libwost_v10.so build/libwost_v12.so" cfg5 cfg5_175e cfg4 cfg2 cfg1b cfg1a cfg3 > gpurun_out/r2f_ab.txt 2>&1
for L in build/libwost_v10.so build/libwost_v12.so; do echo "== $L"; WOST_LIB=$L python tools/survey_rank_job.py 16384; done > gpurun_out/r2f_survey_rank.txt 2>&1
echo "== v12, WOST_SOURCE_BLOBS=0" >> gpurun_out/r2f_survey_rank.txt; WOST_SOURCE_BLOBS=0 WOST_LIB=build/libwost_v12.so python tools/survey_rank_job.py 16384 >> gpurun_out/r2f_survey_rank.txt 2>&1
