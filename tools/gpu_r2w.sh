#!/bin/bash
# round-2 GPU call W: warp-instruction counts of the widened rows' kernels (issue-slot rooflines of their side benches)
mkdir -p gpurun_out
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio
ncu --metrics $M --clock-control none -k regex:"k_wf_" --csv --log-file gpurun_out/r02w_wf.csv python - > /dev/null 2>&1 <<'PY'
import sys; sys.path.insert(0, ".")
from forge3d_b200 import wavefront as wf
scene = wf.scene_from_desc(wf.adjudication_scene())
wf.render_pt_reference(scene, 512, 512, 16)
PY
ncu --metrics $M --clock-control none -k regex:"k_smoke_march" -c 3 --csv --log-file gpurun_out/r02w_smoke.csv python tools/bench_smoke.py --steps 1 --warmup 1 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:"k_shadow_mask|k_viewshed" -c 4 --csv --log-file gpurun_out/r02w_vs.csv python tools/bench_viewshed.py --steps 1 > /dev/null 2>&1
python - <<'PY'
import csv, json, collections
def load(f):
    rows=list(csv.reader(open(f))); h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
    c=rows[h]; ki=c.index("Kernel Name"); mi=c.index("Metric Name"); vi=c.index("Metric Value"); ii=c.index("ID")
    out=collections.defaultdict(dict)
    for r in rows[h+1:]:
        if len(r)>vi: out[(r[ii], r[ki].split("(")[0])][r[mi]]=float(r[vi].replace(",",""))
    return out
res={}
wf=load("gpurun_out/r02w_wf.csv")
tot=sum(v["smsp__inst_executed.sum"] for v in wf.values()); t=sum(v["gpu__time_duration.sum"] for v in wf.values())
lanes=sum(v["smsp__inst_executed.sum"]*v["smsp__thread_inst_executed_per_inst_executed.ratio"] for v in wf.values())/tot
res["k_wf_per_frame_512"]=tot/16; res["k_wf_lanes"]=lanes; res["k_wf_launches"]=len(wf); res["k_wf_ncu_ms_per_frame"]=t/16/1e6
sm=load("gpurun_out/r02w_smoke.csv"); last=list(sm.values())[-1]
res["k_smoke_march_1920x1080_128"]=last["smsp__inst_executed.sum"]; res["k_smoke_lanes"]=last["smsp__thread_inst_executed_per_inst_executed.ratio"]
vs=load("gpurun_out/r02w_vs.csv")
for (i,k),v in vs.items():
    res[k+"_1024"]=v["smsp__inst_executed.sum"]; res[k+"_lanes"]=v["smsp__thread_inst_executed_per_inst_executed.ratio"]
json.dump(res, open("gpurun_out/r02w_rows_instr.json","w"), indent=1); print(json.dumps(res))
PY
