#!/bin/bash
# Round-2 single-GPU verification pass: smoke, whole GPU suite, the bench line, the other configs, ncu launch list and
# one full capture of the hot kernels of the fused periodic step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "rc $?" >> gpurun_out/r2v_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2v_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2v_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
python tools/bench_configs.py > gpurun_out/r2v_configs.log 2>&1
python tools/bench_poisson.py --grid 512 512 512 > gpurun_out/r2v_poisson_512.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-check > gpurun_out/r2v_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'stage_fused|fft_|project_out' -s 60 -c 14 -o gpurun_out/r2v_full python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-check > gpurun_out/r2v_ncu_full.log 2>&1
tail -3 gpurun_out/r2v_pytest.log; head -c 400 gpurun_out/r2v_bench.json; echo; tail -2 gpurun_out/r2v_smoke.log; grep '"case"' gpurun_out/r2v_configs.log | cut -c1-160
