#!/usr/bin/env python
"""profiles/<tag>_launches.csv -> the -s/-c arguments for the `--set full` windows of scripts/gpu_profile.sh.

    python scripts/ncu_skips.py r1k
prints SKIP_FWD (first kernel of the last training pass), SKIP_BWD (20 matching launches before the last
level-0 stage-A backward) and SKIP_GEMM (first contraction of the last full step cycle), all counted over the
launches that match the -k regex of gpu_profile.sh."""
import csv, re, sys
tag = sys.argv[1]
PAT = re.compile(r'gemm_tc|col_stats|act_bwd|scale_shift|k_query|split_bf16|pool_|kp_fwd|kp_bwd|upcat|xent')
lines = [l for l in open(f'profiles/{tag}_launches.csv') if not l.startswith('==')]
names = [r['Kernel Name'] for r in csv.DictReader(lines) if r['ID'] != '']
match = [bool(PAT.search(n)) for n in names]
ks = [i for i, n in enumerate(names) if 'k_starts' in n]
a, b = ks[-26], ks[-13]                      # one full cycle: pyramid + training pass (13 neighbour calls per pyramid)
first_train = [i for i in range(a, b) if 'kp_fwd_tiny' in names[i] or 'kp_fwd_fast' in names[i]][0]
last_bwd = [i for i in range(a, b) if 'kp_bwd_fast' in names[i]][-1]
mi = [i for i in range(len(names)) if match[i]]
gi = [i for i in range(len(names)) if 'gemm_tc' in names[i]]
print("SKIP_FWD=%d SKIP_BWD=%d SKIP_GEMM=%d CNT=24   (cycle = launches %d..%d, %d matching launches in total)" % (
    sum(match[:first_train]), mi.index(last_bwd) - 19, [k for k, i in enumerate(gi) if i >= a][0], a, b, len(mi)))
