#!/bin/bash
# per-kernel durations (ncu) for an arbitrary bench.py command line (both passes in Bloom mode)
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e $*"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"^kg_|void kg_" -c 200 --csv $CMD 2>/dev/null | python -c "
import csv,sys,collections
rows=[r for r in csv.reader(sys.stdin) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ki=H.index('Kernel Name'); mi=H.index('Metric Name'); vi=H.index('Metric Value'); ii=H.index('ID')
d=collections.OrderedDict()
for r in rows[hdr+1:]:
    d.setdefault((int(r[ii]),r[ki].split('(')[0]),{})[r[mi]]=float(r[vi].replace(',',''))
items=[x for x in d.items() if 'ceiling' not in x[0][1]]
half=items[len(items)//2:]
agg=collections.OrderedDict()
for (i,n),m in half:
    a=agg.setdefault(n,[0,0.0,0.0,0.0]); a[0]+=1; a[1]+=m['gpu__time_duration.sum']; a[2]+=m['dram__bytes_read.sum']; a[3]+=m['dram__bytes_write.sum']
tot=sum(a[1] for a in agg.values())
for n,a in agg.items():
    print(f'{a[1]/1e6:9.3f} ms {100*a[1]/tot:5.1f}%  x{a[0]:3d}  rd {a[2]/1e9:7.2f} GB  wr {a[3]/1e9:7.2f} GB  {n}')
print(f'{tot/1e6:9.3f} ms total (timed step of the run)')
"
