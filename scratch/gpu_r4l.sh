#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --workload interactive --gpus 2 --steps 10 --warmup 3 --bergs-per-gpu 5000000 > gpurun_out/r4l_ia_n2.json 2> gpurun_out/r4l_ia_n2.err; echo "rc=$?"; tail -3 gpurun_out/r4l_ia_n2.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r4l_ia_n2.json").read().strip().splitlines()[-1])
print("IA N=2 ms/step", d["ms_per_step"], "value %.4g" % d["value"], "e2e", d["e2e"]["ms_per_step"], "dyn", d["config"]["momentum_thermo_ms_per_step"], "sort", d["config"]["sort_ms_per_step"], "bergs", d["config"]["bergs_total"], "err", d["config"]["device_error_flags"])
PY
