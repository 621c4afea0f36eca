#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
INSB200_HANG_DUMP=50 timeout 90 $TR --master-port 29511 tests/mgpu_worker.py --grid 64 32 64 --dtype f64 --case rb > gpurun_out/r2i_rb_f64.log 2>&1; echo "rc $?" >> gpurun_out/r2i_rb_f64.log
INSB200_HANG_DUMP=50 timeout 90 $TR --master-port 29512 tests/mgpu_worker.py --grid 64 64 64 --dtype f64 --case rb > gpurun_out/r2i_rb_f64b.log 2>&1; echo "rc $?" >> gpurun_out/r2i_rb_f64b.log
grep -v "Warning\|warn\|return func" gpurun_out/r2i_rb_f64.log | tail -40
grep -v "Warning\|warn\|return func" gpurun_out/r2i_rb_f64b.log | tail -25
