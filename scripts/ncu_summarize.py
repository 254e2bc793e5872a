"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
bench.py (with --dump-ops) into per-kernel-class shares and DRAM traffic -> profiles/r1_ncu_step_summary.json and a
compact launch list profiles/r1_ncu_launch_list.csv."""
import collections, csv, json, sys
src, opsf = sys.argv[1], sys.argv[2]
tag = sys.argv[3] if len(sys.argv) > 3 else 'r1'            # round tag of the output file names
outdir = sys.argv[4] if len(sys.argv) > 4 else 'profiles'
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ix = {k: hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID', 'Metric Unit', 'Grid Size', 'Block Size')}
L = collections.OrderedDict()
for r in rows[1:]:
    d = L.setdefault(int(r[ix['ID']]), {'name': r[ix['Kernel Name']], 'grid': r[ix['Grid Size']], 'block': r[ix['Block Size']]})
    v = float(r[ix['Metric Value']].replace(',', '')); u = r[ix['Metric Unit']]
    if r[ix['Metric Name']] == 'gpu__time_duration.sum':
        d['us'] = v / 1000 if u == 'ns' else (v if u == 'us' else v * 1000)
    else:
        d[r[ix['Metric Name']]] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
launches = list(L.values())
ops0 = json.load(open(opsf))
ops = []
C3_DGRAD = ('16->24', '24->40', '40->80')      # one launch (c3_mma.cu); the tcgen05 path takes four (one per parity class)
for op in ops0:
    phases = op[0] == 'conv3x3_dgrad' and 'k3s2' in op[1] and not any(s in op[1] for s in C3_DGRAD)
    ops += [op] * (4 if phases else 1)
per = len(ops)
nsteps = len(launches) // per
assert len(launches) == nsteps * per, (len(launches), per)
step = launches[3 * per:4 * per]          # the timed step (after 3 warm-up steps)
cls = collections.OrderedDict()
for op, l in zip(ops, step):
    c = cls.setdefault(op[0], {'launches': 0, 'us': 0.0, 'dram': 0.0, 'kernels': set()})
    c['launches'] += 1; c['us'] += l['us']
    c['dram'] += l.get('dram__bytes_read.sum', 0) + l.get('dram__bytes_write.sum', 0)
    c['kernels'].add(l['name'].split('(')[0])
tot = sum(c['us'] for c in cls.values())
out = {'command': 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none '
                  '-k regex:_k$ --csv python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline',
       'note': 'per-launch times under ncu are cold-cache and serialised: compare SHARES with the CUDA-event table of '
               'bench.py (profiles/r1_kernel_table_final.json), not absolutes',
       'launches_per_step': per, 'steps_captured': nsteps, 'step_us_sum': round(tot, 1), 'classes': {}}
for k, c in sorted(cls.items(), key=lambda kv: -kv[1]['us']):
    out['classes'][k] = {'sass_kernels': sorted(c['kernels']), 'launches_per_step': c['launches'],
                         'us_per_step': round(c['us'], 1), 'share': round(c['us'] / tot, 4),
                         'dram_GB_per_step': round(c['dram'] / 1e9, 3)}
json.dump(out, open(f'{outdir}/{tag}_ncu_step_summary.json', 'w'), indent=1)
with open(f'{outdir}/{tag}_ncu_launch_list.csv', 'w') as f:
    f.write('id,kernel,grid,block,gpu__time_duration_us,dram_read_bytes,dram_write_bytes\n')
    for i, l in enumerate(step):
        f.write(f'{i},"{l["name"]}","{l["grid"]}","{l["block"]}",{l["us"]:.2f},{l.get("dram__bytes_read.sum",0):.0f},{l.get("dram__bytes_write.sum",0):.0f}\n')
for k, v in list(out['classes'].items())[:16]:
    print(f"{k:20s} {v['launches_per_step']:3d} {v['us_per_step']:8.1f} us {v['share']*100:5.1f}% dram {v['dram_GB_per_step']:6.2f} GB")
print('sum', tot)
