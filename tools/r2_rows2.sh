#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 200 --timeout-method thread -k "rows_kernel" > gpurun_out/r2v_rows_tests.log 2>&1; echo "rows kernel tests rc=$?"
tail -n 5 gpurun_out/r2v_rows_tests.log
timeout -s KILL 300 python -u tools/trie_probe.py llama3-8b-int4 2048 > gpurun_out/r2v_trie_probe.json 2> gpurun_out/r2v_trie_probe.err; echo "probe rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/r2v_trie_probe.json')); print(d['decode_timed_ms'], [(r['nodes'], round(r['pass_ms'],2)) for r in d['trie_pass']])
PY
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2v_trie16_launches.csv python tools/trie_short.py llama3-8b-int4 64 16 > gpurun_out/r2v_trie16_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open('gpurun_out/r2v_trie16_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
data=rows[1:]
names=[re.sub(r'\(.*','',r[ki])[:50] for r in data]
# the last trie pass = everything after the second-to-last sampling_argmax launch
idx=[i for i,n in enumerate(names) if 'sampling' in n]
start=idx[-2]+1 if len(idx)>=2 else 0
agg=collections.Counter(); cnt=collections.Counter(); tot=0
for r,n in zip(data[start:],names[start:]):
    v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1000
    elif r[ui]=='ms': v*=1000
    agg[n]+=v; cnt[n]+=1; tot+=v
print(len(data),'launches; last pass:',len(data)-start,'launches,',round(tot,1),'us')
for n,v in agg.most_common(12): print(f'{v:10.1f} us {cnt[n]:5d}  {n}')
# per-launch durations of the rows kernel in one layer (qkv, out, up, down)
rk=[(float(r[vi].replace(',','')), r[ui]) for r,n in zip(data[start:],names[start:]) if 'qmv_rows' in n]
print('rows kernel launches (first 8):', rk[:8])
PY
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:qmv_rows_kernel -s 260 -c 4 -o gpurun_out/r2v_rows_full python tools/trie_short.py llama3-8b-int4 64 16 > gpurun_out/r2v_rows_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r2v_rows_full.ncu-rep
