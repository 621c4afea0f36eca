#!/bin/bash
# 8-GPU verification: slab == single GPU (periodic + channel), the 1024^3 bench line, config 5.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 tests/mgpu_worker.py --grid 128 64 128 --dtype f32 --case both > gpurun_out/v8_mg8.log 2>&1; echo "rc $?" >> gpurun_out/v8_mg8.log
timeout 400 $TR --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/v8_bench8.json 2> gpurun_out/v8_bench8.err
timeout 400 $TR --master-port 29523 tools/bench_configs.py --cases cfg5 --world 8 > gpurun_out/v8_cfg5_8gpu.log 2>&1
nvidia-smi topo -m > gpurun_out/v8_topo.txt 2>&1
