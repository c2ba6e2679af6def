import csv, collections, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0,0.0,0.0])
for row in r:
    name=row['Kernel Name'].split('(')[0][:70]
    v=float(row['Metric Value'].replace(',',''))
    if row['Metric Name']=='gpu__time_duration.sum':
        agg[name][0]+=1; agg[name][1]+= v/1e3 if row['Metric Unit']=='ns' else (v if row['Metric Unit']=='us' else v*1e3)
    elif 'dram__bytes' in row['Metric Name']:
        u=row['Metric Unit']; f={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[u]
        agg[name][2]+=v*f
tot=sum(a[1] for a in agg.values())
print('total us', round(tot,1))
for k,(n,t,b) in sorted(agg.items(), key=lambda x:-x[1][1]):
    print(f'{t:10.1f} us {n:4d} avg {t/n:8.2f} us  dram_read {b/1e6:9.1f} MB  {k}')
