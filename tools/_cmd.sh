set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2r_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r2r_pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/r2r_bench1.json 2> gpurun_out/r2r_bench1.err; echo rc=$?; tail -c 300 gpurun_out/r2r_bench1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --headline-only --steps 2 --warmup 3 > gpurun_out/r2r_ncu_launches.log 2>&1
for s in cfg5 cfg4 cfg2 cfg1b cfg1a cfg3; do
  timeout 600 ncu --set full --clock-control none -k regex:walk --launch-skip 2 -c 1 -f -o gpurun_out/r2r_full_$s python tools/run_one.py $s 4 > gpurun_out/r2r_ncu_full_$s.log 2>&1
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.txt 2>&1; tail -2 gpurun_out/r2r_smoke.txt
