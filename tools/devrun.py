"""Dev loop helper (run on the GPU box): per-kernel event timings of the bench workloads, compact table."""
import json
import subprocess
import sys

wls = sys.argv[1:] or ['c5', 'c2', 'c4']
for wl in wls:
    out = subprocess.run([sys.executable, 'bench.py', '--workload', wl, '--steps', '10', '--no-e2e', '--no-cpu', '--no-extra', '--no-cache'],
                         capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception:
        print(wl, 'FAILED', out.stderr[-2000:])
        continue
    print('%s: %.2f audio-h/s  %.3f ms/step  clocks %s' % (wl, j['value'], j['ms_per_step'], j['clocks']))
    print('    isolated (one launch at a time): %.3f ms/step' % j['roofline']['isolated_ms_per_step'])
    for k, v in sorted(j['roofline']['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step']):
        print('    %-34s %8.4f ms/step isolated  (%d launches/step, %.1f%%)  %s ms/step in the timed region' % (
            k, v['ms_per_step'], round(v['launches_per_step']), 100 * v['share_of_isolated_step'],
            ('%.4f' % v['ms_per_step_in_timed_region']) if v['ms_per_step_in_timed_region'] is not None else '-'))
