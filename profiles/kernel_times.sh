#!/bin/bash
# per-kernel durations (ncu, cold-cache serialised: compare shares) of one partitioned counting pass
# usage: profiles/kernel_times.sh <scale> <partitions> <batch_mb>
CMD="python bench.py --scale ${1:-0.5} --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --partitions ${2:-128} --batch-mb ${3:-1024}"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"^kg_|void kg_" -c 40 --csv $CMD 2>/dev/null | python -c "
import csv,sys,collections
rows=[r for r in csv.reader(sys.stdin) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ki=H.index('Kernel Name'); mi=H.index('Metric Name'); vi=H.index('Metric Value'); ii=H.index('ID')
d=collections.OrderedDict()
for r in rows[hdr+1:]:
    d.setdefault((int(r[ii]),r[ki].split('(')[0]),{})[r[mi]]=float(r[vi].replace(',',''))
items=list(d.items())
half=[x for x in items if x[0][0]>=len(items)//2 and 'ceiling' not in x[0][1]]
tot=sum(m['gpu__time_duration.sum'] for _,m in half)
for (i,n),m in half:
    print(f\"{m['gpu__time_duration.sum']/1e6:9.3f} ms {100*m['gpu__time_duration.sum']/tot:5.1f}%  rd {m['dram__bytes_read.sum']/1e9:7.2f} GB  wr {m['dram__bytes_write.sum']/1e9:7.2f} GB  {n}\")
print(f'{tot/1e6:9.3f} ms total (second pass of the run)')
"
