#!/bin/bash
mkdir -p gpurun_out
echo "== TC tests"; timeout -k 10 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q --timeout=120 -p no:cacheprovider --tb=short 2>&1 | tail -12 | cut -c1-300
echo "== bench"; timeout -k 10 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_opt.log 2> gpurun_out/bench_opt.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_opt.log').read().strip().splitlines()[-1])
    a=d['roofline_all']
    print("  ms/step %.2f  img/s %.0f  e2e %.2f ms  att %.1f us  conv %.2f ms  dec %.2f ms" % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], a['attention']['us_per_launch'], a['conv']['ms'], a['phases']['decoder_fwd_bwd_ms']))
except Exception as e:
    print("  FAILED", e); print(open('gpurun_out/bench_opt.err').read()[-800:])
PY
echo "== launch list (cache-control none, with dram bytes)"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --cache-control none --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1b.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
tail -1 gpurun_out/prof_step.log
python - <<PY
import csv, re, collections
lines=[l for l in open('gpurun_out/launches_r1b.csv') if l.startswith('"')]
rd=csv.reader(lines); hdr=next(rd)
ki,mi,vi,ui=hdr.index("Kernel Name"),hdr.index("Metric Name"),hdr.index("Metric Value"),hdr.index("Metric Unit")
agg=collections.OrderedDict()
for r in rd:
    name=re.sub(r"\(.*","",r[ki]); name=re.sub(r"^void ","",name)
    v=float(r[vi].replace(",","")); u=r[ui]; m=r[mi]
    a=agg.setdefault(name,{"n":0,"us":0.0,"mb":0.0,"hit":0.0})
    if m.startswith("gpu__time"):
        a["n"]+=1; a["us"]+= v/1e3 if u.startswith("n") else (v if u.startswith("u") else v*1e3)
    elif m.startswith("dram"):
        a["mb"]+= v/1e6 if u=="byte" else (v/1e3 if u=="Kbyte" else (v if u=="Mbyte" else v*1e3))
    else:
        a["hit"]+=v
tot=sum(a["us"] for a in agg.values())
print("total %.2f ms"%(tot/1e3))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1]["us"])[:22]:
    print("%8.2f ms %5.1f%% %5d x %7.1f us  dram %8.1f MB/launch  L2hit %5.1f%%  %s"%(a["us"]/1e3,100*a["us"]/tot,a["n"],a["us"]/a["n"],a["mb"]/a["n"],a["hit"]/a["n"],k[:70]))
PY
