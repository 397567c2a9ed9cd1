import csv, collections, sys
path=sys.argv[1]; nsteps=float(sys.argv[2]) if len(sys.argv)>2 else 1
rows=list(csv.reader(open(path)))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; data=rows[hi+1:]
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit'); idi=hdr.index('ID')
per=collections.OrderedDict()
for r in data:
    if len(r)<=vi: continue
    d=per.setdefault(r[idi],{})
    d[r[mi]]=(float(r[vi].replace(',','')), r[ui]); d['name']=r[ki]
agg=collections.defaultdict(lambda:[0,0.0,0.0]); tot=0
tmul={'ns':1e-3,'nsecond':1e-3,'us':1.0,'usecond':1.0,'ms':1e3,'msecond':1e3}
bmul={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}
for k,d in per.items():
    t=d['gpu__time_duration.sum']; tv=t[0]*tmul.get(t[1],1.0)
    by=sum(d[m][0]*bmul.get(d[m][1],1) for m in ('dram__bytes_read.sum','dram__bytes_write.sum') if m in d)
    n=d['name'].split('(')[0]
    agg[n][0]+=1; agg[n][1]+=tv; agg[n][2]+=by; tot+=tv
print(f'{len(per)} launches, total {tot:.1f} us, per step {tot/nsteps:.1f} us')
for n,(c,t,b) in sorted(agg.items(), key=lambda x:-x[1][1]):
    print(f"{n[:58]:58s} n/step={c/nsteps:6.1f} t={t/nsteps:9.1f}us/step {100*t/tot:5.1f}%  dram={b/nsteps/1e6:9.1f}MB/step {b/t/1e3 if t else 0:7.1f}GB/s")
