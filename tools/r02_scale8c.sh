#!/bin/bash
# third 8-GPU batch of round 2: reduction-station breakdown (profile build) and the CG headline on the final kernel
O=gpurun_out
run() { name=$1; shift; n=$1; shift; echo "== $name"; timeout 600 tools/run_dist.sh $n bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err || echo "rc=$? for $name"; tail -2 $O/$name.err | grep -v "^\*\|OMP_NUM" ; }
ATK_LIB=profile timeout 300 tools/run_dist.sh 8 tools/latency_breakdown.py 512 100 2>&1 | grep "^rank" | tee $O/r02_latency_breakdown_8gpu.txt
ATK_LIB=profile timeout 300 tools/run_dist.sh 8 tools/latency_breakdown.py 64 300 2>&1 | grep "^rank" | tee -a $O/r02_latency_breakdown_8gpu.txt
run r02c_n8_s20 8 --steps 20 --warmup 5 --no-extra
run r02c_n8_s300 8 --steps 300 --warmup 5 --no-extra --no-parity
run r02c_n4_s20 4 --steps 20 --warmup 5 --no-extra
python - <<PY
import json
for f in ["r02c_n8_s20","r02c_n8_s300","r02c_n4_s20"]:
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), [round(k["ms_per_launch"],4) for k in d["roofline"]["per_kernel"]], d["e2e"].get("phases_ms"), "parity", d["parity"]["ok"], "traffic", d["roofline"]["traffic"], "alg", d["roofline"]["algorithmic_bytes"])
    except Exception as ex: print(f, "ERR", ex)
PY
