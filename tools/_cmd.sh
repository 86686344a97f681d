( time timeout 600 python bench.py ) > gpurun_out/r2u_bench1.json 2> gpurun_out/r2u_bench1.err
for s in cfg2 cfg4; do
  timeout 300 ncu --set full --clock-control none -k regex:walk --launch-skip 2 -c 1 -f -o gpurun_out/r2u_full_$s python tools/run_one.py $s 4 > gpurun_out/r2u_ncu_full_$s.log 2>&1
done
