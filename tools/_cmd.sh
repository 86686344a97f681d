set -x
( time timeout 900 python bench.py ) > gpurun_out/r2n_bench1.json 2> gpurun_out/r2n_bench1.err; echo rc=$?; tail -c 300 gpurun_out/r2n_bench1.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2n_bench_ref.json 2> gpurun_out/r2n_bench_ref.err; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --headline-only --steps 2 --warmup 3 > gpurun_out/r2n_ncu_launches.log 2>&1
for s in cfg5 cfg4 cfg2 cfg1b cfg1a cfg3; do
  timeout 600 ncu --set full --clock-control none -k regex:walk --launch-skip 2 -c 1 -f -o gpurun_out/r2n_full_$s python tools/run_one.py $s 4 > gpurun_out/r2n_ncu_full_$s.log 2>&1
done
python tools/survey_bench.py > gpurun_out/r2n_survey_bench.txt 2>&1
python tools/survey_rank_job.py 16384 > gpurun_out/r2n_survey_rank.txt 2>&1
