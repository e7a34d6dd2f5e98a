"""Per-CUDA-source-line aggregation of an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py REP [launch_skip] [top]"""
import csv
import subprocess
import sys


def main(rep, skip=0, top=45):
    out = subprocess.run(['ncu', '-i', rep, '--launch-skip', str(skip), '--launch-count', '1', '--page', 'source', '--csv',
                          '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    fname, hdr, agg, func = None, None, {}, None
    for row in csv.reader(out.splitlines()):
        if not row:
            continue
        if row[0] == 'File Path':
            fname = row[1].split('/')[-1]
        elif row[0] == 'Function Name':
            func = row[1]
        elif row[0] == 'Line No':
            hdr = row
        elif hdr and row[0] not in ('', '-') and row[0].isdigit():
            def g(name):
                try:
                    return int(row[hdr.index(name)] or 0)
                except ValueError:
                    return 0
            key = (fname, int(row[0]))
            a = agg.setdefault(key, {'src': row[1].strip(), 'inst': 0, 'samples': 0, 'wave': 0, 'ideal': 0})
            a['inst'] += g('Instructions Executed')
            a['samples'] += g('# Samples')
            a['wave'] += g('L1 Wavefronts Shared')
            a['ideal'] += g('L1 Wavefronts Shared Ideal')
    ti = sum(a['inst'] for a in agg.values())
    ts = sum(a['samples'] for a in agg.values())
    tw = sum(a['wave'] for a in agg.values())
    print(func)
    print('total inst %d  samples %d  smem wavefronts %d' % (ti, ts, tw))
    print('%-16s %6s %6s %6s %7s  %s' % ('file:line', 'inst%', 'smpl%', 'wave%', 'w/ideal', 'source'))
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:top]:
        print('%-16s %6.2f %6.2f %6.2f %7.2f  %s' % ('%s:%d' % (f[:10], l), 100.0 * a['inst'] / max(ti, 1), 100.0 * a['samples'] / max(ts, 1),
                                                   100.0 * a['wave'] / max(tw, 1), a['wave'] / max(a['ideal'], 1), a['src'][:110]))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 45)


def regions(rep, skip, spec):
    """spec: list of (label, file prefix, lo, hi)"""
    out = subprocess.run(['ncu', '-i', rep, '--launch-skip', str(skip), '--launch-count', '1', '--page', 'source', '--csv',
                          '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    fname, hdr = None, None
    tot = {s[0]: [0, 0, 0] for s in spec}
    tot['other'] = [0, 0, 0]
    for row in csv.reader(out.splitlines()):
        if not row:
            continue
        if row[0] == 'File Path':
            fname = row[1].split('/')[-1]
        elif row[0] == 'Line No':
            hdr = row
        elif hdr and row[0].isdigit():
            vals = []
            for name in ('Instructions Executed', '# Samples', 'L1 Wavefronts Shared'):
                try:
                    vals.append(int(row[hdr.index(name)] or 0))
                except ValueError:
                    vals.append(0)
            lab = 'other'
            for s in spec:
                if fname.startswith(s[1]) and s[2] <= int(row[0]) <= s[3]:
                    lab = s[0]
                    break
            for i in range(3):
                tot[lab][i] += vals[i]
    sums = [sum(v[i] for v in tot.values()) or 1 for i in range(3)]
    print('%-14s %8s %8s %8s' % ('region', 'inst%', 'samples%', 'smemwave%'))
    for k, v in tot.items():
        print('%-14s %8.1f %8.1f %8.1f' % (k, 100.0 * v[0] / sums[0], 100.0 * v[1] / sums[1], 100.0 * v[2] / sums[2]))
