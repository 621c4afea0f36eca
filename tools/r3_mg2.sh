#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
INSB200_HANG_DUMP=60 timeout 90 $TR --master-port 29511 tests/mgpu_worker.py --grid 128 64 128 --dtype f32 --case periodic --strict > gpurun_out/r3b_mg2_f32.log 2>&1; echo "rc $?" >> gpurun_out/r3b_mg2_f32.log
timeout 150 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r3b_bench2.json 2> gpurun_out/r3b_bench2.err; echo "rc $?" >> gpurun_out/r3b_bench2.err
for f in r3b_mg2_f32; do echo "== $f"; grep -c '"ok": true' gpurun_out/$f.log; grep '"ok": false' gpurun_out/$f.log | cut -c1-200; grep native_loop gpurun_out/$f.log | cut -c1-100; grep -v "Warn\|warn" gpurun_out/$f.log | grep -i "error" | head -5; tail -1 gpurun_out/$f.log; done
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r3b_bench2.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["roofline"]["step"]["frac"], d.get("equality_vs_1gpu",{}).get("max_rel_err"), d["selfcheck"])
print({k:round(v["ms_per_step"],3) for k,v in d["roofline"]["groups"].items() if v["ms_per_step"]})
PY
tail -2 gpurun_out/r3b_bench2.err
