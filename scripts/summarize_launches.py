#!/usr/bin/env python
"""profiles/<tag>_launches.csv (ncu launch list of `bench.py --steps 1 --warmup 3 --quick`) ->
profiles/<tag>_launches_summary.md: per-kernel launches / time / share of the LAST step."""
import collections, csv, re, sys
tag = sys.argv[1]
with open(f'profiles/{tag}_launches.csv') as f:
    lines = [l for l in f if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r['ID'] != '']
names = [r['Kernel Name'] for r in rows]
dur = [float(r['Metric Value']) / 1e3 for r in rows]
ks = [i for i, n in enumerate(names) if 'k_starts' in n]
# 13 neighbour calls per pyramid.  With the side-stream prefetcher an iteration is [training of batch i,
# pyramid of batch i+1]; ncu serialises the streams, so one full cycle = from the first launch of the
# second-to-last pyramid up to the first launch of the last one (pyramid + training pass + optimiser).
a, b = ks[-26], ks[-13]
tot = sum(dur[a:b])
def short(n):
    n = n.replace('void ', '').replace('mvk::<unnamed>::', 'mvk::')
    return re.sub(r'\(.*', '', n)[:76]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, d in zip(names[a:b], dur[a:b]):
    agg[short(n)][0] += 1
    agg[short(n)][1] += d
mv = sum(v[1] for k, v in agg.items() if k.startswith('mvk'))
mvn = sum(v[0] for k, v in agg.items() if k.startswith('mvk'))
out = [f"# Round 1, capture {tag}: ncu launch list of the bench command\n",
       f"Command (`scripts/gpu_profile.sh {tag} list`): `ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv "
       f"python bench.py --steps 1 --warmup 3 --quick`; the table covers one full step cycle (launches {a}..{b} of `{tag}_launches.csv`).  "
       "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n",
       f"{b - a} launches in the step, {tot / 1e3:.1f} ms of kernel time; libmvk kernels: {mvn} launches, {mv / 1e3:.1f} ms "
       f"({100 * mv / tot:.0f}%); library kernels (torch: loss, cat, optimizer, fills, gradient accumulation): {(tot - mv) / 1e3:.1f} ms.\n",
       "| kernel | launches | total us | share |\n|---|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    out.append("| `%s` | %d | %.1f | %.1f%% |" % (k, v[0], v[1], 100 * v[1] / tot))
open(f'profiles/{tag}_launches_summary.md', 'w').write("\n".join(out) + "\n")
print("\n".join(out[:30]))
