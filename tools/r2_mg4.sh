#!/bin/bash
mkdir -p gpurun_out
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
INSB200_HANG_DUMP=100 timeout 150 $TR4 --master-port 29511 tests/mgpu_worker.py --grid 128 64 128 --dtype f32 --case all --strict > gpurun_out/r2z_mg4_f32.log 2>&1; echo "rc $?" >> gpurun_out/r2z_mg4_f32.log
INSB200_HANG_DUMP=100 timeout 150 $TR4 --master-port 29512 tests/mgpu_worker.py --grid 64 32 64 --dtype f64 --case all --strict > gpurun_out/r2z_mg4_f64.log 2>&1; echo "rc $?" >> gpurun_out/r2z_mg4_f64.log
timeout 100 $TR2 --master-port 29513 tests/mgpu_worker.py --case timeout > gpurun_out/r2z_timeout.log 2>&1; echo "rc $?" >> gpurun_out/r2z_timeout.log
for f in r2z_mg4_f32 r2z_mg4_f64; do echo "== $f"; grep -c '"ok": true' gpurun_out/$f.log; grep '"ok": false' gpurun_out/$f.log | cut -c1-200; tail -1 gpurun_out/$f.log; done
echo "== timeout"; grep peer_timeout gpurun_out/r2z_timeout.log | cut -c1-300; tail -1 gpurun_out/r2z_timeout.log
timeout 200 python -m pytest tests/test_gpu_variants.py -x -q -m gpu -k "stage_combine" --timeout 100 2>&1 | tail -3
