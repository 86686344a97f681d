set -x
for W in 3 40; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$((W%10)) bench.py --gpus 8 --steps 10 --warmup $W --headline-only > gpurun_out/r2_n8_headline_w$W.json 2> gpurun_out/r2_n8_headline_w$W.err; echo rc=$?
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 40 --warmup 3 --headline-only > gpurun_out/r2_n8_headline_k40.json 2> gpurun_out/r2_n8_headline_k40.err; echo rc=$?
