#!/usr/bin/env bash
# Regenerate the text artefacts under profiles/ from the .ncu-rep / csv files a gpurun call left in gpurun_out/.
set -e
cd "$(dirname "$0")/.."
python tools/ncu_summary.py gpurun_out/prof_fused_f32.ncu-rep gpurun_out/prof_fused_f64.ncu-rep gpurun_out/prof_materialised_terminal.ncu-rep gpurun_out/prof_materialised.ncu-rep > profiles/r1_ncu_summary.txt
sed -i '1i # ncu --set full --clock-control none, one launch each (tools/prof_fused.py at config c2 sizes; float64 at B=8192), all captured from the final round-1 build.\n' profiles/r1_ncu_summary.txt
ncu -i gpurun_out/prof_fused_f32.ncu-rep --page raw --csv 2>/dev/null > profiles/r1_fused_f32_ncu_raw.csv
cp gpurun_out/launches_bench.csv profiles/r1_launches_bench.csv
python tools/sass_loop.py _ZN3smc11tile_kernelIfLi0ELi0ELi0ELb0E --dump > profiles/r1_fused_f32_sass_inner_loop.txt 2>/dev/null
python tools/ptxas_summary.py > profiles/r1_ptxas_summary.txt
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open('profiles/r1_launches_bench.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except ValueError: continue
    name=re.sub(r'\(.*','',r[ki])[:90]
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
step=[k for k in agg if any(s in k for s in ('tile_kernel<float, 0, 0, 0','reduce_tiles','cf_finalize','prep_consts'))]
tot=sum(agg[k][1] for k in step)
out=["# ncu launch list of `python bench.py --steps 5 --warmup 3 --no-cpu-baseline` (gpu__time_duration.sum, --clock-control none)",
     "# kernels of the timed step (fused path) and their share of the step; other rows are the calibration / materialised-roofline kernels bench.py runs after the timed region",""]
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    share = f"{100*t/tot:6.2f}% of step" if k in step else ""
    out.append(f"{n:4d} launches {t/n/1e3:10.2f} us avg  {share:18s} {k}")
open('profiles/r1_launches_bench_summary.txt','w').write("\n".join(out)+"\n")
print("\n".join(out[3:16]))
PY
