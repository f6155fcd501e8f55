"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: total time, share, launches."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        name = re.sub(r'^void ', '', name)[:72]
        v = float(row['Metric Value'].replace(',', ''))
        if row.get('Metric Unit', 'ns') in ('us', 'usecond'):
            v *= 1e3
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f'# {path}: {sum(cnt.values())} launches, {total / 1e6:.3f} ms of kernel time (cold-cache, serialised under ncu)')
    print(f'# {"ms":>9s} {"share":>6s} {"n":>5s}  kernel')
    for k, v in tot.most_common(top):
        print(f'{v / 1e6:10.3f} {100 * v / total:5.1f}% {cnt[k]:5d}  {k}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
