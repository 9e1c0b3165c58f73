#!/bin/bash
# NVLink byte counters (ncu) of the bulk-copy access pattern against a peer's HBM: user vs protocol bytes per direction and kernel
mkdir -p gpurun_out
timeout 280 ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum,gpu__time_duration.sum --clock-control none -k regex:k_rows -c 80 --csv --log-file /tmp/nvl.csv tools/nvlink_bulk_microbench > /tmp/nvl.out 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open('/tmp/nvl.csv') if not l.startswith('=='))]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID'); ui=hdr.index('Metric Unit')
by=collections.OrderedDict()
for r in rows[1:]:
    by.setdefault(r[ii], {'k': r[ki]})[r[mi]] = (float(r[vi].replace(',','')), r[ui])
def b(x):
    v,u=x; return v*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(u,1)
out=['# ncu NVLink counters of tools/nvlink_bulk_microbench (GPU 0 side), launches in program order: per mode local / peer / peer-both-directions, each warm-up then timed',
     '# id kernel  time_us  nvlrx_GB (user_GB)  nvltx_GB (user_GB)  rx_GB/s tx_GB/s']
for i,m in by.items():
    t=m['gpu__time_duration.sum']; tus=t[0]*{'us':1,'ms':1e3,'ns':1e-3,'s':1e6}.get(t[1],1)
    rx,tx=b(m['nvlrx__bytes.sum']),b(m['nvltx__bytes.sum']); rxu,txu=b(m['nvlrx__bytes_data_user.sum']),b(m['nvltx__bytes_data_user.sum'])
    if rx+tx < 1e6: continue
    out.append('%s %s %.0f  rx %.3f (%.3f)  tx %.3f (%.3f)  %.0f %.0f' % (i, m['k'][:60], tus, rx/1e9, rxu/1e9, tx/1e9, txu/1e9, rx/tus/1e3, tx/tus/1e3))
open('gpurun_out/nvlink_ncu_r02.txt','w').write('\n'.join(out)+'\n')
print('\n'.join(out[:40]))
PY
tail -3 /tmp/nvl.out
