import csv, json, os
root='/root/repo'
out={}
def grab(f):
    rows=list(csv.reader(open(f)))
    hdr,units,vals=rows[0],rows[1],rows[2]
    d={h:(vals[i],units[i]) for i,h in enumerate(hdr)}
    def bytes_(k):
        v,u=d[k]; v=float(v)
        return v*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[u]
    return bytes_('dram__bytes_read.sum'), bytes_('dram__bytes_write.sum'), d['Kernel Name'][0], float(d['gpu__time_duration.sum'][0])
for name,f,chains,iters in (('cfg2','r2_cfg2_spec_raw.csv',1024,200),('cfg3','r2_cfg3_raw.csv',65536,10),('cfg5','r2_cfg5_raw.csv',131072,10),('cfg4r','r2_cfg4r_raw.csv',16384,1)):
    p=os.path.join(root,'profiles/r2/ncu',f)
    if not os.path.exists(p) or os.path.getsize(p)<1000: continue
    r,w,k,ms=grab(p)
    out[name]={'kernel':k,'chains_per_gpu':chains,'iters_per_launch':iters,'dram_bytes_read':r,'dram_bytes_written':w,'dram_bytes_per_launch':r+w,
               'launch_ms_under_ncu':ms,'source':f'profiles/r2/ncu/{f} (ncu --set full --clock-control none, one launch of the bench workload)'}
json.dump(out,open(os.path.join(root,'profiles/ncu_traffic.json'),'w'),indent=1)
print(json.dumps(out,indent=1))
