#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_dec_wgrad|k_dec_bn_grads|k_dec_input_grad|k_dec_conv' --launch-skip 120 --launch-count 40 --csv --log-file gpurun_out/dec_launches.csv python tools/bench_decoder.py --iters 1 > gpurun_out/ncu_dec2.log 2>&1
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/dec_launches.csv')))
h=next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
hdr=rows[h]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
d={}
for r in rows[h+1:]:
    if len(r)!=len(hdr): continue
    d.setdefault(r[iid],{'k':r[ik]})[r[im]]=r[iv]
for k,v in list(d.items())[:40]:
    print(k, v['k'][:60], v.get('gpu__time_duration.sum'), v.get('launch__grid_size'), v.get('launch__registers_per_thread'), v.get('sm__warps_active.avg.pct_of_peak_sustained_active'))
P
