python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; tail -n 4 gpurun_out/r2_pytest2.log
for c in cfg5_175e cfg4 cfg1b cfg2 cfg3 cfg1a; do WOST_JIT=0 python tools/run_one.py $c 3 | tail -1; WOST_JIT=1 python tools/run_one.py $c 3 | tail -1; done 2>&1 | tee gpurun_out/r2_jit_ab2.log
WOST_JIT=0 python tools/run_one.py cfg5_175e 3 32768 | tail -1; WOST_JIT=1 python tools/run_one.py cfg5_175e 3 32768 | tail -1
python tools/small_solve.py 200 2>&1 | tee gpurun_out/r2_small2.log
