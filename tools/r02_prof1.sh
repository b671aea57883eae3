#!/bin/bash
# 1-GPU validation + profiling batch of round 2 (run under gpurun); outputs in gpurun_out/
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r02_gputests4.log; cat $O/r02_gputests4.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_n1_full.json 2> $O/r02_n1_full.err; tail -2 $O/r02_n1_full.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_n1_full.json").read().strip().splitlines()[-1])
print("N=1", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d["parity"]["ok"], d.get("cpu_baseline"))
print(json.dumps(d.get("extra"))[:3500])
PY
# persistent loop forced on one GPU, for the ncu capture of the shipped multi-GPU kernel
ATK_CG_PERSISTENT=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-parity > $O/r02_n1_persist.json 2> $O/r02_n1_persist.err &&
ATK_CG_PERSISTENT=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02_launches_persist.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-parity > $O/ncu_a.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-parity > $O/r02_n1_graph.json 2> $O/r02_n1_graph.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_graph.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-parity > $O/ncu_b.log 2>&1
ATK_CG_PERSISTENT=1 ncu --set full --clock-control none --import-source on -k regex:cg_persistent -c 1 -o $O/r02_persistent_full python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-parity > $O/ncu_c.log 2>&1
timeout 120 python tools/profile_spmv.py 256 > $O/r02_spmv_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:csr_stream2 -s 2 -c 1 -o $O/r02_csr_stream2_full python tools/profile_spmv.py 256 > $O/ncu_d.log 2>&1
ls -la $O/*.ncu-rep | tail -5
