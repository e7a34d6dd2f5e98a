"""Dev loop helper (run on the GPU box): per-kernel event timings of the bench workloads, compact table."""
import json
import subprocess
import sys

wls = sys.argv[1:] or ['c5', 'c2', 'c4']
for wl in wls:
    out = subprocess.run([sys.executable, 'bench.py', '--workload', wl, '--steps', '10', '--no-e2e', '--no-cpu'],
                         capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception:
        print(wl, 'FAILED', out.stderr[-2000:])
        continue
    print('%s: %.2f audio-h/s  %.3f ms/step  clocks %s' % (wl, j['value'], j['ms_per_step'], j['clocks']))
    for k, v in sorted(j['roofline']['kernels'].items(), key=lambda kv: -kv[1]['ms_total']):
        print('    %-34s %8.4f ms/step  (%d launches/step, %.1f%%)' % (k, v['ms_total'] / j['steps'], v['launches'] // j['steps'], 100 * v['share_of_step']))
