#!/bin/bash
# Session-8 single-GPU verification: parity suite, bench line, kernel-variant timings, ncu.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s8_smoke.log 2>&1; echo "rc $?" >> gpurun_out/s8_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s8_pytest.log 2>&1; echo "rc $?" >> gpurun_out/s8_pytest.log
python tools/bench_poisson.py --pipeline 0 1 2 3 > gpurun_out/s8_poisson_512.log 2>&1
python tools/bench_poisson.py --grid 1024 512 256 --pipeline 0 3 >> gpurun_out/s8_poisson_512.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err
INSB200_FFT_PIPELINE=3 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/s8_bench_pipe3.json 2> gpurun_out/s8_bench_pipe3.err
INSB200_FFT_PIPELINE=2 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/s8_bench_pipe2.json 2> gpurun_out/s8_bench_pipe2.err
INSB200_NO_FUSED_DIVERGENCE=1 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/s8_bench_nofuse.json 2> gpurun_out/s8_bench_nofuse.err
python tools/bench_configs.py > gpurun_out/s8_configs.log 2>&1
INSB200_BOUSS_DIRECT=1 python tools/bench_configs.py --cases cfg4 > gpurun_out/s8_configs_bouss_direct.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s8_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s8_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'stage_pair|fft_|pgrad' -s 40 -c 14 -o gpurun_out/s8_full python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s8_ncu_full.log 2>&1
ls -la gpurun_out
