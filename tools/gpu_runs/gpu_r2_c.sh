#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hdemucs_backward.py -m gpu -q --timeout 600 --no-header -p no:cacheprovider -s > gpurun_out/r2c_bwd.log 2>&1
echo "backward exit=$? $(tail -n 1 gpurun_out/r2c_bwd.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r2c_bwd.log | head
for b in 2 16; do
  timeout 600 python tools/hd_train_bench.py --batch $b --steps 3 --warmup 2 > gpurun_out/r2c_hd_train_b$b.json 2> gpurun_out/r2c_hd_train_b$b.err
  echo "train b=$b exit=$?"; cat gpurun_out/r2c_hd_train_b$b.json; tail -n 3 gpurun_out/r2c_hd_train_b$b.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_hd_train_b2.csv python tools/hd_train_bench.py --batch 2 --steps 1 --warmup 1 > gpurun_out/r2c_ncu.log 2>&1
echo "ncu exit=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2c_launches_hd_train_b2.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; kn = h.index('Kernel Name'); mv = h.index('Metric Value'); mu = h.index('Metric Unit')
rows = rows[hdr + 1:]
half = rows[len(rows) // 2:]   # second step only (the first is warm-up)
acc = collections.Counter(); cnt = collections.Counter()
for r in half:
    try: v = float(r[mv].replace(',', ''))
    except Exception: continue
    if r[mu] == 'us': v *= 1e3
    elif r[mu] == 'ms': v *= 1e6
    name = r[kn].split('(')[0][:70]
    acc[name] += v; cnt[name] += 1
tot = sum(acc.values())
with open('gpurun_out/r2c_launch_shares_hd_train_b2.txt', 'w') as fh:
    fh.write(f"total {tot/1e6:.2f} ms over {sum(cnt.values())} launches (second step of tools/hd_train_bench.py --batch 2, ncu serialised)\n")
    for k, v in acc.most_common(30):
        fh.write(f"{v/tot*100:6.2f}%  {v/1e6:8.3f} ms  {cnt[k]:5d}x  {k}\n")
print(open('gpurun_out/r2c_launch_shares_hd_train_b2.txt').read())
PY
