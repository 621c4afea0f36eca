#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
INSB200_HANG_DUMP=150 timeout 200 $TR --master-port 29511 tests/mgpu_worker.py --grid 64 32 64 --dtype f64 --case all > gpurun_out/r2j_all_f64.log 2>&1; echo "rc $?" >> gpurun_out/r2j_all_f64.log
INSB200_HANG_DUMP=150 timeout 200 $TR --master-port 29512 tests/mgpu_worker.py --grid 128 64 128 --dtype f32 --case all > gpurun_out/r2j_all_f32.log 2>&1; echo "rc $?" >> gpurun_out/r2j_all_f32.log
for f in r2j_all_f64 r2j_all_f32; do echo "== $f"; grep -c '"ok": true' gpurun_out/$f.log; grep '"ok": false' gpurun_out/$f.log; grep -v "Warning\|warn\|return func\|^{" gpurun_out/$f.log | tail -8; done
python -m pytest tests/test_gpu_diagnostics.py -x -q -m gpu -k random_field 2>&1 | tail -15
