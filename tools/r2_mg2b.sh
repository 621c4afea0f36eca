#!/bin/bash
# 2-GPU check: slab == single GPU for the periodic, channel and Rayleigh-Benard (walls in the decomposed z) cases
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tests/mgpu_worker.py --grid 64 32 64 --dtype f64 --case all > gpurun_out/r2h_mg2_f64.log 2>&1; echo "rc $?" >> gpurun_out/r2h_mg2_f64.log
timeout 400 $TR --master-port 29512 tests/mgpu_worker.py --grid 128 64 64 --dtype f32 --case rb > gpurun_out/r2h_mg2_f32.log 2>&1; echo "rc $?" >> gpurun_out/r2h_mg2_f32.log
INSB200_NO_P2P=1 timeout 400 $TR --master-port 29513 tests/mgpu_worker.py --grid 64 32 64 --dtype f64 --case rb > gpurun_out/r2h_mg2_f64_nccl.log 2>&1; echo "rc $?" >> gpurun_out/r2h_mg2_f64_nccl.log
for f in r2h_mg2_f64 r2h_mg2_f32 r2h_mg2_f64_nccl; do echo "== $f"; grep -o '"case": "[^"]*", "err": [^,]*, "tol": [^,]*, "ok": [a-z]*, "world": [0-9]*, "n": [^]]*\], "dtype": "[a-z0-9]*", "problem": "[a-z]*"' gpurun_out/$f.log; tail -n 3 gpurun_out/$f.log; done
