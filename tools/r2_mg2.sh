#!/bin/bash
# 2-GPU check of the fused slab path: slab == single GPU, then the weak-scaling bench line (fused / stage-by-stage)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/mgpu_worker.py --grid 64 32 64 --dtype f64 > gpurun_out/r2e_mg2_f64.log 2>&1; echo "rc $?" >> gpurun_out/r2e_mg2_f64.log
timeout 300 $TR --master-port 29512 tests/mgpu_worker.py --grid 128 64 128 --dtype f32 > gpurun_out/r2e_mg2_f32.log 2>&1; echo "rc $?" >> gpurun_out/r2e_mg2_f32.log
INSB200_NO_P2P=1 timeout 300 $TR --master-port 29513 tests/mgpu_worker.py --grid 64 32 64 --dtype f64 > gpurun_out/r2e_mg2_f64_nccl.log 2>&1; echo "rc $?" >> gpurun_out/r2e_mg2_f64_nccl.log
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2e_bench2.json 2> gpurun_out/r2e_bench2.err
INSB200_NO_FUSED_SLAB=1 timeout 300 $TR --master-port 29515 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2e_bench2_unfused.json 2> gpurun_out/r2e_bench2_unfused.err
tail -3 gpurun_out/r2e_mg2_f64.log gpurun_out/r2e_mg2_f32.log gpurun_out/r2e_mg2_f64_nccl.log
