"""Condense the ncu CSVs brought back in gpurun_out/ into profiles/ (tracked): launch-list shares and the
per-launch --set full metrics that the roofline statements in DESIGN.md / bench.py rely on."""
import csv, collections, sys, os, json

tag = sys.argv[1] if len(sys.argv) > 1 else 'r01b'
G = 'gpurun_out'
os.makedirs('profiles', exist_ok=True)
out = []

# ---- launch list (one eager bench step: 2 trunks + tails)
rows = list(csv.reader(open(f'{G}/launches_{tag}.csv')))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.OrderedDict()
n_launch = 0
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0].replace('void ', '').replace('sb::', '')[:70]
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    n_launch += 1
tot = sum(v[1] for v in agg.values())
out.append(f'## Launch list of ONE bench step (B=256, bf16, eager; ncu gpu__time_duration.sum, cold-cache/serialised: compare SHARES)\n')
out.append(f'{n_launch} launches, {tot/1e6:.3f} ms total under ncu\n')
out.append('| kernel | launches | total us | share |\n|---|---:|---:|---:|')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f'| `{k}` | {v[0]} | {v[1]/1e3:.1f} | {100*v[1]/tot:.1f}% |')
is_conv = lambda k: k.lstrip().startswith('conv') or 'bottleneck' in k          # conv_tc*, conv_tcp*, conv_stem7*, conv3x3_halo, bottleneck64
conv_share = sum(v[1] for k, v in agg.items() if is_conv(k)) / tot
out.append(f'\nconv kernels share of the step: **{100*conv_share:.1f}%**\n')

# ---- full capture
rows = list(csv.reader(open(f'{G}/conv_{tag}_raw.csv')))
h = rows[0]
col = {c: i for i, c in enumerate(h)}
want = [('Kernel Name', 'kernel'), ('launch__grid_size', 'grid'), ('gpu__time_duration.sum', 'us'),
        ('dram__bytes_read.sum', 'dram rd MB'), ('dram__bytes_write.sum', 'dram wr MB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'),
        ('launch__registers_per_thread', 'regs'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %')]
want = [(a, b) for a, b in want if a in col]
out.append('## `ncu --set full` of the conv-family launches of ONE step (both trunks)\n')
out.append('| # | ' + ' | '.join(b for _, b in want) + ' |\n|---|' + '---|' * len(want))
def short(n):
    n = n.replace('void sb::', '').replace('__nv_bfloat16', 'bf16')
    return n.split('(')[0]
tens, drams = [], []
for i, r in enumerate(rows[2:]):
    vals = []
    for a, b in want:
        v = r[col[a]]
        if a == 'Kernel Name':
            v = '`' + short(v) + '`'
        else:
            try:
                v = f'{float(v.replace(",", "")):.1f}' if '.' in v else v
            except ValueError:
                pass
        vals.append(v)
    out.append(f'| {i} | ' + ' | '.join(vals) + ' |')
    try:
        tens.append((float(r[col['gpu__time_duration.sum']]), float(r[col['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']]),
                     float(r[col['dram__bytes_read.sum']]) + float(r[col['dram__bytes_write.sum']])))
    except Exception:
        pass
if tens:
    tt = sum(t for t, _, _ in tens)
    unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(rows[1][col['dram__bytes_read.sum']], 1e6)
    total_bytes = sum(b for _, _, b in tens) * unit          # the raw page states one unit per column (row 2 of the CSV)
    out.append(f'\ntime-weighted tensor-pipe utilisation over these launches: **{sum(t*p for t,p,_ in tens)/tt:.1f}%**; '
               f'DRAM traffic {total_bytes/1e9:.2f} GB per step over {len(tens)} conv-family launches\n')
    # bench.py reads roofline.traffic from this file (key = backbone|batch|precision of the captured command)
    key = sys.argv[2] if len(sys.argv) > 2 else 'resnet50|256|bf16'
    tp = 'profiles/ncu_traffic.json'
    d = json.load(open(tp)) if os.path.exists(tp) else {}
    d[key] = {'bytes_per_step': total_bytes, 'launches': len(tens), 'source': f'profiles/ncu_{tag}.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)'}
    json.dump(d, open(tp, 'w'), indent=1)
open(f'profiles/ncu_{tag}.md', 'w').write('\n'.join(out) + '\n')
print('\n'.join(out[:40]))
