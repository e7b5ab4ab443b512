"""Summarises gpurun_out/ncu_conv_r01.csv (per-launch ncu metrics of the tcgen05 conv kernels of one micro-batch)."""
import csv, json, collections, sys
src = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/ncu_conv_r01.csv'
lines = [l for l in open(src) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
by = collections.OrderedDict()
for r in rows:
    by.setdefault(r['ID'], {'kernel': r['Kernel Name'].split('(')[0], 'grid': r['Grid Size']})[r['Metric Name']] = (float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 's': 1, '%': 1, 'Tbyte': 1e12}
def val(d, k):
    v, u = d[k]
    return v * scale.get(u, 1)
T = 'gpu__time_duration.sum'; TP = 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'
tot_t = sum(val(d, T) for d in by.values())
dram = sum(val(d, 'dram__bytes_read.sum') + val(d, 'dram__bytes_write.sum') for d in by.values())
l2 = sum(val(d, 'lts__t_bytes.sum') for d in by.values())
xbar = sum(val(d, 'l1tex__m_xbar2l1tex_read_bytes.sum') for d in by.values())
tp = sum(val(d, TP) * val(d, T) for d in by.values()) / tot_t
out = {'launches': len(by), 'total_ms': tot_t * 1e3, 'dram_bytes_per_launch': dram / len(by), 'dram_gb_total': dram / 1e9,
       'l2_gb_total': l2 / 1e9, 'l2_to_sm_gb_total': xbar / 1e9, 'tensor_pipe_pct_time_weighted': tp,
       'tensor_pipe_pct_max': max(val(d, TP) for d in by.values()), 'tensor_pipe_pct_min': min(val(d, TP) for d in by.values())}
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
    with open(sys.argv[2].replace('.json', '_launches.txt'), 'w') as f:
        f.write("# id kernel grid time_us tensor_pipe_pct dram_MB l2_to_sm_MB\n")
        for i, d in by.items():
            f.write("%s %s %s %.1f %.1f %.1f %.1f\n" % (i, d['kernel'].replace('void ', '').replace('tise::', ''), d['grid'].replace(' ', ''), val(d, T) * 1e6, val(d, TP),
                    (val(d, 'dram__bytes_read.sum') + val(d, 'dram__bytes_write.sum')) / 1e6, val(d, 'l1tex__m_xbar2l1tex_read_bytes.sum') / 1e6))
