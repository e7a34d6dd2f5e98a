"""Summarise an ncu report's SASS source page: total stall samples by reason and the hottest instructions."""
import csv
import subprocess
import sys


def main(rep, kernel_idx=0, top=25):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(out.splitlines()):
        if row and row[0] == 'Kernel Name':
            cur = {'name': row[1], 'hdr': None, 'rows': []}
            blocks.append(cur)
        elif cur is not None and row and row[0] == 'Address':
            cur['hdr'] = row
        elif cur is not None and cur['hdr'] and row:
            cur['rows'].append(row)
    b = blocks[kernel_idx]
    h = b['hdr']
    print(b['name'], len(b['rows']), 'instructions')
    si = h.index('# Samples')
    stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    tot = {h[i]: 0 for i in stall_cols}
    total = 0
    for r in b['rows']:
        total += int(r[si] or 0)
        for i in stall_cols:
            tot[h[i]] += int(r[i] or 0)
    print('total samples', total)
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
        print('  %-24s %8d %5.1f%%' % (k, v, 100.0 * v / max(total, 1)))
    ei = h.index('Instructions Executed')
    print('total warp instructions executed', sum(int(r[ei] or 0) for r in b['rows']))
    rows = sorted(b['rows'], key=lambda r: -int(r[si] or 0))[:top]
    for r in rows:
        reasons = sorted(((int(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
        print('%6s %-70s %s' % (r[si], r[1][:70], ' '.join('%s=%d' % (n, c) for c, n in reasons)))
    # instruction mix
    mix = {}
    for r in b['rows']:
        op = r[1].split()[0] if r[1] else ''
        if op.startswith('@'):
            op = r[1].split()[1]
        op = op.split('.')[0]
        mix[op] = mix.get(op, 0) + int(r[ei] or 0)
    print('instruction mix (warp-level):')
    for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:14]:
        print('  %-10s %10d' % (k, v))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 25)


def conflicts(rep, kernel_idx=0, top=20):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(out.splitlines()):
        if row and row[0] == 'Kernel Name':
            cur = {'name': row[1], 'hdr': None, 'rows': []}
            blocks.append(cur)
        elif cur is not None and row and row[0] == 'Address':
            cur['hdr'] = row
        elif cur is not None and cur['hdr'] and row:
            cur['rows'].append(row)
    b = blocks[kernel_idx]
    h = b['hdr']
    wi, xi, ii = h.index('L1 Wavefronts Shared'), h.index('L1 Wavefronts Shared Excessive'), h.index('L1 Wavefronts Shared Ideal')
    tot = sum(int(r[wi] or 0) for r in b['rows'])
    exc = sum(int(r[xi] or 0) for r in b['rows'])
    print('shared wavefronts', tot, 'excessive', exc)
    agg = {}
    for r in b['rows']:
        if int(r[wi] or 0) > 0:
            op = r[1].split()[0] if not r[1].startswith('@') else r[1].split()[1]
            a = agg.setdefault(op, [0, 0])
            a[0] += int(r[wi]); a[1] += int(r[xi] or 0)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print('  %-12s wavefronts %10d excessive %10d' % (k, v[0], v[1]))
    for r in sorted(b['rows'], key=lambda r: -int(r[xi] or 0))[:top]:
        print('%10s %10s  %s' % (r[wi], r[xi], r[1][:80]))


if __name__ == '__main__' and len(sys.argv) > 4 and sys.argv[4] == 'conflicts':
    conflicts(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
