python -m pytest tests/test_gpu_parity.py -x -q -k "jit" > gpurun_out/r2_jit_tests.log 2>&1; tail -n 15 gpurun_out/r2_jit_tests.log
for c in cfg5_175e cfg4 cfg1b cfg2 cfg3 cfg1a; do WOST_JIT=0 python tools/run_one.py $c 3 | tail -1; WOST_JIT=1 python tools/run_one.py $c 3 | tail -1; done 2>&1 | tee gpurun_out/r2_jit_ab.log
WOST_JIT=0 python tools/run_one.py cfg5_175e 3 32768 | tail -1; WOST_JIT=1 python tools/run_one.py cfg5_175e 3 32768 | tail -1
