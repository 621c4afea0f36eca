#!/bin/bash
# Round-2 (session 4) profile evidence: ncu launch list of the bench command, full capture of the tile kernels of the
# Boussinesq pass and of the eddy-viscosity closure (reduced sizes, tools/profile_secondary.py).
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4g_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-check > gpurun_out/r4g_ncu_launch.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'bouss_tile|les_tile' -c 12 -o gpurun_out/r4g_tiles python tools/profile_secondary.py > gpurun_out/r4g_ncu_tiles.log 2>&1
tail -2 gpurun_out/r4g_ncu_launch.log; tail -2 gpurun_out/r4g_ncu_tiles.log; ls -la gpurun_out/r4g*
