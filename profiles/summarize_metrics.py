"""Summarises an ncu CSV of profiles/profile_step.py taken with
  --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
per kernel and steady-state step: time, warp instructions, DRAM and L2 bytes (cold-cache, serialised
launches: compare shares)."""
import collections, csv, sys

M = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum']

def main(path, first=3):
    lines = [l for l in open(path) if not l.startswith('==')]
    L = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = L.setdefault(int(r['ID']), {'name': r['Kernel Name'].split('(')[0].replace('void ', '')})
        d[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    launch = [v for v in L.values() if 'at::' not in v['name'] and 'elementwise' not in v['name']]
    idx = [i for i, v in enumerate(launch) if v['name'].startswith('prep_kernel')]
    agg = collections.OrderedDict()
    for si in range(first, len(idx)):
        a, b = idx[si], (idx[si + 1] if si + 1 < len(idx) else len(launch))
        for v in launch[a:b]:
            g = agg.setdefault(v['name'], collections.Counter())
            g['n'] += 1
            for m in M:
                g[m] += v.get(m, 0)
    ns = len(idx) - first
    print('steady-state steps %d..%d (batch 32); per step:' % (first, len(idx) - 1))
    print('%-24s %6s %9s %10s %9s %9s %9s' % ('kernel', 'n/stp', 'us', 'Mwarpinst', 'dramR MB', 'dramW MB', 'L2 MB'))
    tot = collections.Counter()
    for k, g in sorted(agg.items(), key=lambda kv: -kv[1][M[1]]):
        row = [g[M[0]] / ns / 1e3, g[M[1]] / ns / 1e6, g[M[2]] / ns / 1e6, g[M[3]] / ns / 1e6, g[M[4]] / ns / 1e6]
        print('%-24s %6.1f %9.1f %10.2f %9.2f %9.2f %9.1f' % ((k, g['n'] / ns) + tuple(row)))
        for i, v in enumerate(row):
            tot[i] += v
    print('%-24s %6s %9.1f %10.2f %9.2f %9.2f %9.1f' % (('TOTAL', '') + tuple(tot[i] for i in range(5))))
    print('issue-limited time of the step at 148 SMs x 4 schedulers x 1.965 GHz: %.0f us' % (tot[1] * 1e6 / (148 * 4 * 1.965e9) * 1e6))

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
