"""Extract the headline metrics of every kernel in an ncu report into JSON (merged into profiles/r01_ncu_raw_summary.json under KEY).
usage: python tools/ncu_summary.py REP KEY "workload note" [OUT_JSON]"""
import csv
import json
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_active']


def main(rep, key, note, out='profiles/r01_ncu_raw_summary.json'):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        k = {'Kernel Name': r[h.index('Kernel Name')]}
        for w in WANT:
            if w in h:
                i = h.index(w)
                k[w] = ('%s %s' % (r[i], units[i])).strip()
        kernels.append(k)
    try:
        d = json.load(open(out))
    except Exception:
        d = {}
    d[key] = {'workload': note, 'kernels': kernels}
    json.dump(d, open(out, 'w'), indent=1)
    for k in kernels:
        print(k['Kernel Name'], k.get('gpu__time_duration.sum'), k.get('dram__bytes_read.sum'), k.get('dram__bytes_write.sum'))


if __name__ == '__main__':
    main(*sys.argv[1:])
