#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
INSB200_HANG_DUMP=60 timeout 100 $TR --master-port 29511 tests/mgpu_worker.py --grid 128 64 128 --dtype f32 --case periodic --strict > gpurun_out/r3d_f32.log 2>&1; echo "rc $?" >> gpurun_out/r3d_f32.log
grep -c '"ok": true' gpurun_out/r3d_f32.log; grep '"ok": false' gpurun_out/r3d_f32.log | cut -c1-160; grep -h native_loop gpurun_out/r3d_f32.log | cut -c1-110; tail -1 gpurun_out/r3d_f32.log
