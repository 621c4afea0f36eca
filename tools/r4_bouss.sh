#!/bin/bash
# Round-2 (session 4) GPU check: Boussinesq tile pass with the operand prefetch (parity + A/B on cfg 4), eddy-viscosity
# closure after the launch-bound fix.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_les.py tests/test_gpu_variants.py tests/test_gpu_sizes.py -x -q -k "les or closure or bouss or rayleigh" 2>&1 | tail -5 > gpurun_out/r4e_pytest.log
python tools/bench_bouss.py > gpurun_out/r4e_bouss.log 2>&1
python tools/bench_les.py --scale 0.5 --models WALE > gpurun_out/r4e_les_bench.log 2>&1
tail -3 gpurun_out/r4e_pytest.log; cut -c1-420 gpurun_out/r4e_bouss.log; cut -c1-200 gpurun_out/r4e_les_bench.log
