import csv, collections, re, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.OrderedDict(); n=0
for row in r:
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    name=re.sub(r'\(.*','',row['Kernel Name'])
    v=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    v = v/1e3 if unit=='ns' else (v*1e3 if unit=='ms' else (v*1e6 if unit=='s' else v))
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v; n+=1
tot=sum(a[1] for a in agg.values())
print('launches',n,'total us',round(tot,1))
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    print(f'{t:12.1f} us {100*t/tot:5.1f}%  x{c:5d}  avg {t/c:9.2f} us  {k[:120]}')
