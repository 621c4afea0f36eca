#!/bin/bash
# Session-8 two-GPU verification: parity suite (incl. slab == single GPU, periodic + channel),
# peer-store halo vs NCCL halo, staged fused-divergence pass, pipelined strided passes.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s8b_pytest.log 2>&1; echo "rc $?" >> gpurun_out/s8b_pytest.log
INSB200_NCCL_HALO=1 timeout 300 $TR --master-port 29513 tests/mgpu_worker.py --grid 64 32 64 --dtype f64 > gpurun_out/s8b_mg2_ncclhalo.log 2>&1; echo "rc $?" >> gpurun_out/s8b_mg2_ncclhalo.log
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/s8b_bench2.json 2> gpurun_out/s8b_bench2.err
INSB200_NCCL_HALO=1 timeout 300 $TR --master-port 29515 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/s8b_bench2_ncclhalo.json 2> gpurun_out/s8b_bench2_ncclhalo.err
INSB200_NO_FUSED_DIVERGENCE=1 timeout 300 $TR --master-port 29516 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/s8b_bench2_nofuse.json 2> gpurun_out/s8b_bench2_nofuse.err
timeout 300 $TR --master-port 29517 tools/bench_configs.py --cases cfg5 --world 2 > gpurun_out/s8b_cfg5_2gpu.log 2>&1
INSB200_DEBUG=1 python tools/bench_poisson.py --pipeline 0 1 2 3 > gpurun_out/s8b_poisson_512.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s8b_bench1.json 2> gpurun_out/s8b_bench1.err
INSB200_DIV_STAGED=0 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/s8b_bench1_unstaged.json 2> gpurun_out/s8b_bench1_unstaged.err
python tools/bench_configs.py --cases cfg4 cfg5 > gpurun_out/s8b_configs.log 2>&1
ls -la gpurun_out | tail -15
