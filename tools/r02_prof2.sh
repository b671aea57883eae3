#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r02_gputests5.log; cat $O/r02_gputests5.log
for st in 20 100; do
  timeout 300 python bench.py --steps $st --warmup 5 --no-cpu-baseline --no-extra --no-parity > $O/r02_ka_ldgsts_s$st.json 2>> $O/r02_ka.err
  ATK_FUSED_TMA=1 timeout 300 python bench.py --steps $st --warmup 5 --no-cpu-baseline --no-extra --no-parity > $O/r02_ka_tma_s$st.json 2>> $O/r02_ka.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --no-config5 > $O/r02_n1_extra.json 2>> $O/r02_ka.err
python - <<PY
import json
for f in ["r02_ka_ldgsts_s20","r02_ka_tma_s20","r02_ka_ldgsts_s100","r02_ka_tma_s100","r02_n1_extra"]:
    d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), [ (k["kernel"][:28],round(k["ms_per_launch"],4)) for k in d["roofline"]["per_kernel"]])
    if "extra" in d: print({k:(round(v["ms"],4),round(v["frac_of_measured_peak"],3)) for k,v in d["extra"].items() if "ms" in v})
PY
ncu --set full --clock-control none --import-source on -k regex:stencil7_cg_fused_smem -s 3 -c 1 -o $O/r02_ka_ldgsts_full python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-parity > $O/ncu_e.log 2>&1
ATK_FUSED_TMA=1 ncu --set full --clock-control none --import-source on -k regex:stencil7_cg_fused_tma -s 3 -c 1 -o $O/r02_ka_tma_full python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-parity > $O/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:csr_stream2 -s 2 -c 1 -f -o $O/r02_csr_stream2_full python tools/profile_spmv.py 256 > $O/ncu_d.log 2>&1
ls -la $O/*.ncu-rep | tail -5
