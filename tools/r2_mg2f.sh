#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2p_bench2_full.json 2> gpurun_out/r2p_bench2_full.err; echo "rc $?" >> gpurun_out/r2p_bench2_full.err
timeout 200 $TR --master-port 29515 tools/bench_configs.py --cases cfg4 --world 2 > gpurun_out/r2p_cfg4_2gpu.log 2>&1; echo "rc $?" >> gpurun_out/r2p_cfg4_2gpu.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2p_bench2_full.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["config"]["workload"][:60], d["e2e"]["value"] if d["e2e"] else None, d["parity"], d.get("equality_vs_1gpu",{}).get("max_rel_err"), d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
print({k:round(v["ms_per_step"],3) for k,v in d["roofline"]["groups"].items()})
PY
tail -3 gpurun_out/r2p_bench2_full.err
grep '"case"' gpurun_out/r2p_cfg4_2gpu.log | cut -c1-300; tail -2 gpurun_out/r2p_cfg4_2gpu.log
