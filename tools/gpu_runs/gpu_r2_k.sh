#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_launches_hd_fwd_b32.csv python tools/hd_fwd_probe.py 32 > gpurun_out/r2k_ncu.log 2>&1
echo "ncu exit=$?"
python tools/launch_shares.py gpurun_out/r2k_launches_hd_fwd_b32.csv gpurun_out/r2k_launch_shares_hd_fwd_b32.txt "second forward of tools/hd_fwd_probe.py 32" | head -30
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2k_launches_hd_fwd_b32.csv',errors='ignore')))
hdr=next(i for i,r in enumerate(rows) if 'Kernel Name' in r); h=rows[hdr]
kn,mv,mu,gs=h.index('Kernel Name'),h.index('Metric Value'),h.index('Metric Unit'),h.index('Grid Size')
rows=rows[hdr+1:]; rows=rows[len(rows)//2:]
def val(r):
    v=float(r[mv].replace(',','')); return v*(1e-3 if r[mu]=='ns' else 1 if r[mu]=='us' else 1e3)
seq=[(i,val(r),r[kn].split('(')[0][:40],r[gs]) for i,r in enumerate(rows)]
top=sorted(seq,key=lambda t:-t[1])[:25]
print("top launches (index in the forward, us, kernel, grid):")
for t in top: print(t)
PY
timeout 300 python tools/tcn_train_bench.py --batch 16 --steps 2 --warmup 1 > gpurun_out/r2k_tcn_train_b16.json 2> gpurun_out/r2k_tcn_train_b16.err; echo "tcn b16 exit=$?"; cat gpurun_out/r2k_tcn_train_b16.json; tail -n 2 gpurun_out/r2k_tcn_train_b16.err
timeout 300 python tools/tcn_train_bench.py --cpu-baseline --batch 1 --cpu-T 32768 --steps 2 > gpurun_out/r2k_tcn_train_cpu.json 2>/dev/null; cat gpurun_out/r2k_tcn_train_cpu.json
