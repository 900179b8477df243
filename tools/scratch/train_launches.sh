ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_train1024.csv python bench.py --steps 3 --warmup 3 --no-heads --no-cpu-baseline --train-scenes 1024 --strong-scenes 0 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r2b_launches_train1024.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); iu=h.index('Metric Unit')
seq=[(r[ik], float(r[iv].replace(',',''))*(1e-3 if r[iu]=='ns' else 1.0 if r[iu]=='us' else 1e3)) for r in rows[hdr+1:] if len(r)>iv]
# last training step: find the last occurrence of 'multi_tensor' (AdamW) groups; print the launches between the last two optimizer steps
idx=[i for i,(k,_) in enumerate(seq) if 'enc_fwd_tc_kernel' in k]
a=idx[-1]
step=seq[a-1:]  # from the last enc pack kernel on
tot=collections.OrderedDict()
for k,us in step:
    name=k.split('(')[0][-60:]
    tot[name]=tot.get(name,0)+us
for k,v in sorted(tot.items(), key=lambda x:-x[1])[:25]: print(f'{v:9.1f} us  {k}')
print('sum', sum(tot.values()))
PY
