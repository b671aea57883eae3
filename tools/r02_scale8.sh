#!/bin/bash
# 8-GPU measurement batch of round 2 (run under gpurun --gpus 8); everything lands in gpurun_out/
O=gpurun_out
run() { name=$1; shift; n=$1; shift; echo "== $name"; timeout 600 tools/run_dist.sh $n bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err || echo "rc=$? for $name"; tail -2 $O/$name.err | grep -v "^\*\|OMP_NUM" ; }
nvidia-smi topo -m > $O/r02_topo8.txt 2>&1
for n in 8 4; do timeout 300 tools/run_dist.sh $n tools/dist_check.py 24 40 > $O/r02_dist$n.log 2>&1; grep -v "^\*\|OMP_NUM\|^$" $O/r02_dist$n.log | tail -24; done
run r02_n8_s20_full 8 --steps 20 --warmup 5
run r02_n8_s300 8 --steps 300 --warmup 5 --no-extra --no-parity
ATK_PERSIST_SPLIT=2 run r02_n8_s20_split2 8 --steps 20 --warmup 5 --no-extra --no-parity
ATK_CG_PERSISTENT=0 run r02_n8_s20_graph 8 --steps 20 --warmup 5 --no-extra --no-parity
ATK_NUMA_PIN=0 run r02_n8_s20_nonuma 8 --steps 20 --warmup 5 --no-extra --no-parity
run r02_n4_s20_full 4 --steps 20 --warmup 5
python - <<PY
import json
for f in ["r02_n8_s20_full","r02_n8_s300","r02_n8_s20_split2","r02_n8_s20_graph","r02_n8_s20_nonuma","r02_n4_s20_full"]:
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), [round(k["ms_per_launch"],4) for k in d["roofline"]["per_kernel"]], d["e2e"].get("phases_ms"), d["e2e"].get("numa"), "parity", d["parity"]["ok"])
        if "extra" in d: print(json.dumps(d["extra"])[:3000])
    except Exception as e: print(f, "ERR", e)
PY
