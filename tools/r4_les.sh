#!/bin/bash
# Round-2 (session 4) GPU check of the fused eddy-viscosity closure: parity tests, closure timing fused vs four-pass,
# cfg 5 (+ WALE) at half extents, one ncu --set full capture of the two fused kernels.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_les.py -x -q 2>&1 | tail -8 > gpurun_out/r4c_les_pytest.log
python tools/bench_les.py --scale 0.5 --models WALE Smagorinsky > gpurun_out/r4c_les_bench.log 2>&1
python tools/bench_configs.py --cases cfg5 --scale 0.5 > gpurun_out/r4c_cfg5_half.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'les_tile' -s 2 -c 2 -o gpurun_out/r4c_les_fused \
    python tools/bench_les.py --scale 0.5 --models WALE --reps 1 > gpurun_out/r4c_les_ncu.log 2>&1
tail -3 gpurun_out/r4c_les_pytest.log; cat gpurun_out/r4c_les_bench.log | cut -c1-220; cut -c1-330 gpurun_out/r4c_cfg5_half.log
