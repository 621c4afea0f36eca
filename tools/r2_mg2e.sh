#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
INSB200_HANG_DUMP=100 timeout 150 $TR --master-port 29511 tests/mgpu_worker.py --grid 64 32 64 --dtype f64 --case periodic > gpurun_out/r2k_f64.log 2>&1; echo "rc $?" >> gpurun_out/r2k_f64.log
INSB200_HANG_DUMP=100 timeout 150 $TR --master-port 29512 tests/mgpu_worker.py --grid 128 64 128 --dtype f32 --case all > gpurun_out/r2k_f32.log 2>&1; echo "rc $?" >> gpurun_out/r2k_f32.log
INSB200_NO_P2P=1 INSB200_HANG_DUMP=100 timeout 150 $TR --master-port 29513 tests/mgpu_worker.py --grid 64 32 64 --dtype f64 --case periodic > gpurun_out/r2k_f64_nccl.log 2>&1; echo "rc $?" >> gpurun_out/r2k_f64_nccl.log
for f in r2k_f64 r2k_f32 r2k_f64_nccl; do echo "== $f"; grep -c '"ok": true' gpurun_out/$f.log; grep '"ok": false' gpurun_out/$f.log; grep 'random_field' gpurun_out/$f.log | cut -c1-120; grep -v "Warning\|warn\|return func\|^{" gpurun_out/$f.log | tail -6; done
