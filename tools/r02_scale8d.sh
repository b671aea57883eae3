#!/bin/bash
# fourth 8-GPU batch of round 2 (final multi-GPU code): parity log at 8 ranks, full bench lines at 8 and 4 GPUs
O=gpurun_out
run() { name=$1; shift; n=$1; shift; echo "== $name"; timeout 600 tools/run_dist.sh $n bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err || echo "rc=$? for $name"; tail -2 $O/$name.err | grep -v "^\*\|OMP_NUM" ; }
timeout 300 tools/run_dist.sh 8 tools/dist_check.py 24 40 > $O/r02_dist8_final2.log 2>&1; grep -v "^\*\|OMP_NUM\|^$" $O/r02_dist8_final2.log | tail -8
run r02d_n8_s20_full 8 --steps 20 --warmup 5
run r02d_n4_s20_full 4 --steps 20 --warmup 5
python - <<PY
import json
for f in ["r02d_n8_s20_full","r02d_n4_s20_full"]:
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), [round(k["ms_per_launch"],4) for k in d["roofline"]["per_kernel"]], "parity", d["parity"]["ok"])
        e=d["extra"]
        print("   spmv", {k:(round(v["ms"],4), round(v.get("frac_of_measured_peak_per_gpu",0),3)) for k,v in e.items() if isinstance(v,dict) and "ms" in v})
        g=e.get("gmres300_op27_256")
        if g: print("   gmres", {k:(round(g[k]["seconds_per_restart"],4), g[k]["residual_rel_diff"]) for k in ("batched_mgs","exact_mgs")}, g["modes_rel_diff"])
    except Exception as ex: print(f, "ERR", ex)
PY
