#!/bin/bash
# second 8-GPU batch of round 2 (final code): parity logs, stress, bench at 8 / 4 / 2 GPUs, polling experiments
O=gpurun_out
run() { name=$1; shift; n=$1; shift; echo "== $name"; timeout 600 tools/run_dist.sh $n bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err || echo "rc=$? for $name"; tail -2 $O/$name.err | grep -v "^\*\|OMP_NUM" ; }
for n in 8 4; do timeout 300 tools/run_dist.sh $n tools/dist_check.py 24 40 > $O/r02_dist${n}_final.log 2>&1; grep -v "^\*\|OMP_NUM\|^$" $O/r02_dist${n}_final.log | tail -5; done
ATK_EPOCH_SEED=0xFFFFFE00 timeout 600 tools/run_dist.sh 8 tools/stress_dist.py 1000 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -3 | tee $O/r02_stress8.log
run r02b_n8_s20_full 8 --steps 20 --warmup 5
ATK_PERSIST_POLL=1 run r02b_n8_s20_poll1 8 --steps 20 --warmup 5 --no-extra --no-parity
ATK_PERSIST_POLL_SLEEP=100 run r02b_n8_s20_sleep100 8 --steps 20 --warmup 5 --no-extra --no-parity
ATK_PERSIST_POLL=1 ATK_PERSIST_POLL_SLEEP=100 run r02b_n8_s20_poll1_sleep100 8 --steps 20 --warmup 5 --no-extra --no-parity
run r02b_n8_s300 8 --steps 300 --warmup 5 --no-extra --no-parity
run r02b_n4_s20_full 4 --steps 20 --warmup 5
run r02b_n2_s20_full 2 --steps 20 --warmup 5
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > $O/r02b_n1_s20.json 2> $O/r02b_n1_s20.err
python - <<PY
import json
for f in ["r02b_n1_s20","r02b_n2_s20_full","r02b_n4_s20_full","r02b_n8_s20_full","r02b_n8_s20_poll1","r02b_n8_s20_sleep100","r02b_n8_s20_poll1_sleep100","r02b_n8_s300"]:
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), [round(k["ms_per_launch"],4) for k in d["roofline"]["per_kernel"]], d["e2e"].get("phases_ms"), "parity", d["parity"]["ok"])
        if "extra" in d:
            e=d["extra"]
            print("   spmv", {k:(round(v["ms"],4), round(v.get("frac_of_measured_peak_per_gpu",v.get("frac_of_measured_peak",0)),3)) for k,v in e.items() if isinstance(v,dict) and "ms" in v})
            g=e.get("gmres300_op27_256")
            if g: print("   gmres", {k:(round(g[k]["seconds_per_restart"],4), round(g[k]["projection_frac_of_measured_peak_per_gpu"],3), g[k]["residual_rel_diff"]) for k in ("batched_mgs","exact_mgs")}, g["modes_rel_diff"])
    except Exception as ex: print(f, "ERR", ex)
PY
