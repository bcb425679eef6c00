"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of profiles/profile_step.py:
per-kernel average duration and share of a steady-state step (steps 4.. of the run)."""
import collections, csv, sys

def load(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') == 'gpu__time_duration.sum':
            rows.append((int(row['ID']), row['Kernel Name'].split('(')[0], float(row['Metric Value'].replace(',', ''))))
    return rows

def main(path, first_step=4):
    rows = [r for r in load(path) if 'at::' not in r[1] and 'elementwise' not in r[1]]
    idx = [i for i, r in enumerate(rows) if r[1].startswith('prep_kernel')]
    per = collections.OrderedDict()
    sums = []
    for si in range(first_step, len(idx)):
        a, b = idx[si], (idx[si + 1] if si + 1 < len(idx) else len(rows))
        sums.append(sum(r[2] for r in rows[a:b]) / 1e3)
        for r in rows[a:b]:
            per.setdefault(r[1], []).append(r[2])
    nsteps = len(sums)
    print('steps %d..%d: sum of kernel durations per step [us]: %s' % (first_step, len(idx) - 1, ' '.join('%.0f' % s for s in sums)))
    tot = sum(sum(v) for v in per.values())
    print('%-26s %5s %10s %10s %7s' % ('kernel', 'n/stp', 'avg_us', 'us/step', 'share'))
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print('%-26s %5.1f %10.1f %10.1f %6.1f%%' % (k, len(v) / nsteps, sum(v) / len(v) / 1e3, sum(v) / nsteps / 1e3, 100 * sum(v) / tot))

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 4)
