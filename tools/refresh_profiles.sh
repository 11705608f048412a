#!/usr/bin/env bash
# Regenerate the round-2 text artefacts under profiles/ from what tools/r2_profile.sh left in gpurun_out/.
set -e
cd "$(dirname "$0")/.."
python tools/ncu_summary.py gpurun_out/r2_prof_fused.ncu-rep gpurun_out/r2_prof_fused_f64.ncu-rep gpurun_out/r2_prof_fused_c3.ncu-rep > profiles/r2_ncu_summary.txt
sed -i '1i # ncu --set full --clock-control none, one launch each (tools/prof_fused.py: config c2 for float32, B=8192 for float64, the trainer-test shape 1024 x (T=1, N=16, B=4096) for the short-path kernel), round-2 build.\n' profiles/r2_ncu_summary.txt
ncu -i gpurun_out/r2_prof_fused.ncu-rep --page raw --csv 2>/dev/null > profiles/r2_fused_f32_ncu_raw.csv
ncu -i gpurun_out/r2_prof_fused_f64.ncu-rep --page raw --csv 2>/dev/null > profiles/r2_fused_f64_ncu_raw.csv
ncu -i gpurun_out/r2_prof_fused_c3.ncu-rep --page raw --csv 2>/dev/null > profiles/r2_fused_short_ncu_raw.csv
grep -v "^==" gpurun_out/r2_launches_bench.csv > profiles/r2_launches_bench.csv
python tools/sass_loop.py _ZN3smc11step_kernelIfLi0ELi0ELi0ELi0E --dump 2>/dev/null | awk '/^loop /{n++} n<=1' > profiles/r2_fused_f32_sass_inner_loop.txt
python tools/sass_loop.py _ZN3smc11step_kernelIdLi0ELi0ELi0ELi1E --dump 2>/dev/null | awk '/^loop /{n++} n<=1' > profiles/r2_fused_f64_sass_inner_loop.txt
python tools/ptxas_summary.py > profiles/r2_ptxas_summary.txt
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open('profiles/r2_launches_bench.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except ValueError: continue
    name=re.sub(r'\(.*','',r[ki])[:90]
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
step=[k for k in agg if 'step_kernel<float, 0, 0, 0' in k]
tot=sum(agg[k][1] for k in step)
out=["# ncu launch list of `python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-workloads` (gpu__time_duration.sum, --clock-control none, first 400 launches)",
     "# the RAW step is ONE kernel (smc::step_kernel: tiles, ticket tree, transform) preceded by a memset of its control words;",
     "# the other rows are the L2 flush (torch fill), the calibration and the materialised-roofline kernels bench.py runs after the timed regions",""]
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    share = f"{100*t/tot:6.2f}% of step" if k in step else ""
    out.append(f"{n:4d} launches {t/n/1e3:10.2f} us avg  {share:18s} {k}")
open('profiles/r2_launches_bench_summary.txt','w').write("\n".join(out)+"\n")
print("\n".join(out[4:16]))
PY
