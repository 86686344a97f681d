python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_pytest_multi.log 2>&1; tail -n 5 gpurun_out/r2_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --quick > gpurun_out/r2_bench2_quick.json 2> gpurun_out/r2_bench2_quick.err; echo rc=$?; tail -c 1500 gpurun_out/r2_bench2_quick.err
python bench.py --gpus 1 --steps 5 --warmup 3 --quick --no-cpu-baseline > gpurun_out/r2_bench1_quick.json 2> gpurun_out/r2_bench1_quick.err; echo rc=$?
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench1_quick.json','gpurun_out/r2_bench2_quick.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    print(f, {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e %.3e'%d['e2e']['value'])
    for k,v in d['strong_scaling'].items(): print('  ', k, v['ms'], v['checksum'], v['sharding'])
    for k,v in d['rmse_vs_time'].items(): print('  ', k, [(x['walks'], round(x['wall_s']*1e6), x['checksum']) for x in v])
    if 'configs' in d:
        for k,v in d['configs'].items(): print('  ', k, '%.3e'%v['steps_per_s'], v['shipped_size']['device_resident_us_blocks'], v['shipped_size']['host_buffers_us_blocks'])
PY
