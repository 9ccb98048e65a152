import csv, collections, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
with open(path) as f:
    lines=[l for l in f if not l.startswith('==')]
rows=list(csv.DictReader(lines))
def us(row):
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    return v/1e3 if u=='ns' else v*1e3 if u=='ms' else v*1e6 if u=='s' else v
# keep the second half (the timed step) by splitting at the last quantile-normalisation hist<0> launch
idx=[i for i,r in enumerate(rows) if 'q_hist_kernel<0>' in r['Kernel Name']]
start = idx[-1] if idx else len(rows)//2
rows=rows[start:]
agg=collections.defaultdict(lambda:[0,0.0])
for row in rows:
    name=re.sub(r'\(.*','',row['Kernel Name']).replace('void ','').replace('adni::','').replace('<unnamed>::','')
    agg[name][0]+=1; agg[name][1]+=us(row)
tot=sum(v[1] for v in agg.values())
print(f"one step: {len(rows)} launches, {tot/1e3:.2f} ms of kernel time (cold-cache, serialised)")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:32]:
    print(f"{v[1]/tot*100:6.2f}%  {v[1]/1e3:8.2f} ms  n={v[0]:4d}  {k[:100]}")
