( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2v_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r2v_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.txt 2>&1; tail -2 gpurun_out/r2v_smoke.txt
python bench.py --headline-only --steps 5 --warmup 3 > gpurun_out/r2v_headline.json 2>/dev/null; cat gpurun_out/r2v_headline.json | cut -c1-200
