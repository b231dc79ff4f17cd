#!/bin/bash
mkdir -p gpurun_out
for mm in 2 9; do
UZU_PREFILL_GEMM_MIN_M=$mm timeout -s KILL 300 python -u tools/trie_probe.py llama3-8b-int4 2048 > gpurun_out/r2t_trie_probe_gemm$mm.json 2> gpurun_out/r2t_trie_probe_gemm$mm.err; echo "probe gemm>=$mm rc=$?"
python - <<PY
import json; d=json.load(open('gpurun_out/r2t_trie_probe_gemm$mm.json')); print([(r['nodes'], round(r['pass_ms'],2)) for r in d['trie_pass']])
PY
done
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2t_trie16_launches.csv python tools/trie_short.py llama3-8b-int4 64 16 > gpurun_out/r2t_trie16_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open('gpurun_out/r2t_trie16_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
data=rows[1:]
# last pass = tail: take the final 40% of launches as the two trie passes region; just aggregate the last 700 launches
tail=data[-700:]
agg=collections.Counter(); cnt=collections.Counter()
for r in tail:
    v=float(r[vi].replace(',','')); 
    if r[ui]=='ns': v/=1000
    elif r[ui]=='ms': v*=1000
    n=re.sub(r'\(.*','',r[ki])[:60]; agg[n]+=v; cnt[n]+=1
print(len(data),'launches; last 700:')
for n,v in agg.most_common(12): print(f'{v:10.1f} us {cnt[n]:5d}  {n}')
PY
