import csv, json, sys
R='/root/repo/gpurun_out/'
print(open(R+'pytest.log').read().strip().splitlines()[-2:])
for f in ['bench_val.log','bench_custom.log']:
    try:
        d=json.loads(open(R+f).read().strip().splitlines()[-1]); print(f, 'clips/s',round(d['value']), 'ms/step',round(d['ms_per_step'],3), 'k1_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',open(R+f).read()[-1500:])
try:
    rows=list(csv.reader(open(R+'launches_custom.csv')))
    for i,r in enumerate(rows):
        if r and r[0]=='ID': hdr=r; start=i+1; break
    ci={h:i for i,h in enumerate(hdr)}
    cur={}
    for r in rows[start:]:
        if len(r)<len(hdr): continue
        key=(r[ci['ID']], r[ci['Kernel Name']][:40])
        cur.setdefault(key,{})[r[ci['Metric Name']]]=r[ci['Metric Value']]
    for k,v in cur.items():
        print(k[0], k[1], 'us', float(v['gpu__time_duration.sum'])/1e3, 'rdMB', round(float(v['dram__bytes_read.sum'])/1e6), 'wrMB', round(float(v['dram__bytes_write.sum'])/1e6), 'Minst', round(float(v.get('smsp__inst_executed.sum',0))/1e6,1), 'ipc', v.get('sm__inst_issued.avg.per_cycle_active'))
except Exception as e: print('launch list ERR', e)
