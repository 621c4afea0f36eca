#!/bin/bash
# 8-GPU verification of round 2: slab == single GPU (periodic incl. random_field, channel, Rayleigh-Benard with walls
# in the decomposed z), the config[2] bench line (DecayingTurbulence3D 1024^3), configs 4 and 5 on 8 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
INSB200_HANG_DUMP=100 timeout 150 $TR --master-port 29521 tests/mgpu_worker.py --grid 128 64 128 --dtype f32 --case all > gpurun_out/r2m_mg8.log 2>&1; echo "rc $?" >> gpurun_out/r2m_mg8.log
timeout 240 $TR --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2m_bench8.json 2> gpurun_out/r2m_bench8.err; echo "rc $?" >> gpurun_out/r2m_bench8.err
timeout 200 $TR --master-port 29523 tools/bench_configs.py --cases cfg4 cfg5 --world 8 > gpurun_out/r2m_cfg45_8gpu.log 2>&1; echo "rc $?" >> gpurun_out/r2m_cfg45_8gpu.log
echo "== mg8"; grep -c '"ok": true' gpurun_out/r2m_mg8.log; grep '"ok": false' gpurun_out/r2m_mg8.log; tail -2 gpurun_out/r2m_mg8.log
echo "== bench8"; cut -c1-600 gpurun_out/r2m_bench8.json; tail -3 gpurun_out/r2m_bench8.err
echo "== cfg"; grep '"case"' gpurun_out/r2m_cfg45_8gpu.log | cut -c1-400; tail -2 gpurun_out/r2m_cfg45_8gpu.log
