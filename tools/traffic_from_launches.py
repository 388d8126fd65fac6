"""profiles/<round>_traffic.json from an ncu launch list of `bench.py --quick` (per-kernel share of a step, DRAM bytes).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file launches.csv python bench.py --steps 2 --warmup 1 --quick --no-e2e --no-cpu
    python tools/traffic_from_launches.py launches.csv profiles/r2_traffic.json <launches of EACH pass kernel per step> [range_replay.txt]

The optional fourth argument is the output of `ncu --replay-mode range` of ONE whole apply (tools/r2_profile.sh): its DRAM
bytes are those of the step under its real concurrency (the launch list serialises the kernels, which keeps every slab in L2).

ncu serialises the launches and flushes caches between them, so absolute times are cold-cache; what bench.py quotes
from here is each kernel's SHARE of a step and the DRAM bytes per step.
"""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import source_hash            # noqa: E402  (ties the profile to the kernel sources it was captured on)

src, dst, per_step = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = list(csv.reader(open(src)))
hdr = None
per = {}
for r in rows:
    if r and r[0] == 'ID':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        k = per.setdefault(d['Kernel Name'], {'ids': set(), 'ns': 0.0, 'rd': 0.0, 'wr': 0.0})
        k['ids'].add(d['ID'])
        v = float(d['Metric Value'].replace(',', ''))
        unit = d['Metric Unit']
        scale = {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)
        if d['Metric Name'] == 'gpu__time_duration.sum':
            k['ns'] += v * scale
        elif d['Metric Name'] == 'dram__bytes_read.sum':
            k['rd'] += v * scale
        elif d['Metric Name'] == 'dram__bytes_write.sum':
            k['wr'] += v * scale
# only the transform kernels of the library (the bench also launches torch's RNG / copy kernels)
ours = {k: v for k, v in per.items() if 'fmb::' in k or 'v32_pass' in k or 'v32t_pass' in k or 'v32p_kernel' in k or 'fast_pass' in k or 'fwht' in k}
tot_ns = sum(v['ns'] for v in ours.values())
N, COLS = 1 << 20, 1024
out = {'source': src + ' (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; '
                   'bench.py --quick; the capture covers part of the run, per-step figures = per-launch averages x %d launches of each pass per step)' % per_step,
       'workload': 'Circulant(2^20).forward, complex64, 1024 columns', 'source_hash': source_hash(), 'kernels': {}}
for k, v in sorted(ours.items(), key=lambda t: -t[1]['ns']):
    n = len(v['ids'])
    out['kernels'][k] = {'launches_captured': n, 'launches_per_step': per_step, 'share_of_step_time': v['ns'] / tot_ns,
                         'avg_us': v['ns'] / n / 1e3, 'dram_read_bytes_per_launch': v['rd'] / n,
                         'dram_write_bytes_per_launch': v['wr'] / n}
out['dram_bytes_per_step'] = sum((v['rd'] + v['wr']) / len(v['ids']) * per_step for v in ours.values())
out['algorithmic_bytes_per_step'] = 2 * 8 * N * COLS
out['traffic_over_algorithmic'] = out['dram_bytes_per_step'] / out['algorithmic_bytes_per_step']
out['dominant_kernel'] = max(ours, key=lambda k: ours[k]['ns'])
if len(sys.argv) > 4:
    rd = wr = None
    for ln in open(sys.argv[4]):
        t = ln.split()
        if len(t) >= 3 and t[0] == 'dram__bytes_read.sum':
            rd = float(t[-1]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[t[1]]
        if len(t) >= 3 and t[0] == 'dram__bytes_write.sum':
            wr = float(t[-1]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[t[1]]
    if rd is not None and wr is not None:
        out['dram_bytes_per_step_concurrent'] = rd + wr
        out['concurrent_note'] = ('ncu --replay-mode range of one whole apply (all slabs, internal streams): %.2f GB read + %.2f GB written = '
                                  '%.2fx the algorithmic bytes' % (rd / 1e9, wr / 1e9, (rd + wr) / out['algorithmic_bytes_per_step']))
json.dump(out, open(dst, 'w'), indent=1)
print(json.dumps({k: out[k] for k in ('dram_bytes_per_step', 'traffic_over_algorithmic', 'dominant_kernel')}))
