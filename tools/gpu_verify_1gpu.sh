#!/bin/bash
# Single-GPU verification pass: whole parity suite, bench line, other configs, ncu.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; echo "rc $?" >> gpurun_out/v_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/v_pytest.log 2>&1; echo "rc $?" >> gpurun_out/v_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/v_bench_ref.json 2> gpurun_out/v_bench_ref.err
python tools/bench_configs.py > gpurun_out/v_configs.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/v_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'stage_pair|fft_|pgrad|divergence' -s 48 -c 16 -o gpurun_out/v_full python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/v_ncu_full.log 2>&1
tail -3 gpurun_out/v_pytest.log; cat gpurun_out/v_bench.json | head -c 600
