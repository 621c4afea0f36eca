#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
INSB200_HANG_DUMP=60 timeout 90 $TR --master-port 29511 tests/mgpu_worker.py --grid 64 32 64 --dtype f64 --case periodic --strict > gpurun_out/r3c_f64.log 2>&1
INSB200_NO_CHAINED_STEPS=1 INSB200_HANG_DUMP=60 timeout 90 $TR --master-port 29512 tests/mgpu_worker.py --grid 128 64 128 --dtype f32 --case periodic --strict > gpurun_out/r3c_f32_nochain.log 2>&1
grep -h native_loop gpurun_out/r3c_f64.log gpurun_out/r3c_f32_nochain.log | cut -c1-110
