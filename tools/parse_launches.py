import csv, re, collections, sys
fn = sys.argv[1]
with open(fn) as f:
    lines=[l for l in f if not l.startswith('==')]
rows=[(r['Kernel Name'], float(r['Metric Value'])) for r in csv.DictReader(lines)]
names=[n for n,_ in rows]
starts=[i for i,n in enumerate(names) if 'act_embed_kernel' in n]
print('captured', len(rows), 'step starts', starts[:8])
k=min(2,len(starts)-2); i0=starts[k]; i1=starts[k+1]
step=rows[i0:i1]
agg=collections.OrderedDict()
for n,t in step:
    key=re.sub(r'\(.*','',n.replace('void ',''))
    agg.setdefault(key,[0,0.0]); agg[key][0]+=1; agg[key][1]+=t
tt=sum(t for _,t in step)
print('launches in step', len(step), 'step total (ncu, serialized, cold) us', round(tt/1000,1))
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print('  %-50s x%-3d %8.1f us  %5.1f%%' % (k[:50], c, t/1000, 100*t/tt))
