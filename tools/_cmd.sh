set -x
( time timeout 600 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2_pytest_multi.log 2>&1; tail -n 4 gpurun_out/r2_pytest_multi.log
for N in 8 2; do
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo rc=$?; tail -c 300 gpurun_out/r2_bench_n$N.err
done
