#!/bin/bash
# Round-2 (session 4) single-GPU verification pass: smoke, whole GPU suite, the bench line, cfg 4 and cfg 5 (+ WALE).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4f_smoke.log 2>&1; echo "rc $?" >> gpurun_out/r4f_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r4f_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r4f_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r4f_bench.json 2> gpurun_out/r4f_bench.err
timeout 300 python tools/bench_configs.py --cases cfg4 cfg5 > gpurun_out/r4f_configs.log 2>&1
tail -3 gpurun_out/r4f_pytest.log; head -c 300 gpurun_out/r4f_bench.json; echo; tail -2 gpurun_out/r4f_smoke.log; grep '"case"' gpurun_out/r4f_configs.log | cut -c1-200
