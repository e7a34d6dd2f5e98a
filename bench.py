#!/usr/bin/env python
"""
bench.py -- audio-hours/sec of HCQT + log-mel features (BASELINE.json `metric`), one rank per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c3|c4|c1] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic tracks that stays fixed per rank
(weak scaling: every rank processes its own `--batch` tracks; no data-path collective).  Default workload
`c5` is the configuration the metric is quoted on (BASELINE.json configs[4], per-GPU shape):
HCQT(22050, hop 256, 360 bins @ 60 bpo, h = .5,1,2,3,4,5) + MelSpec(16000, n_fft 2048, hop 512, 229 mels)
on 4-minute tracks.  `value` = device-resident throughput; `e2e` = the same through the C-ABI host entry
point (pinned host audio in, pinned host features out, copies inside the timed region).

`--impl reference` times the reference's CPU algorithm (the oracle port: librosa itself cannot be
installed here) on all host cores over a bounded sample of the same workload.
"""

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'audio-hours/sec of HCQT+log-mel features'

WORKLOADS = {
    # name: (description, [(module ctor name, kwargs, sample_rate)], clip seconds, default batch per GPU)
    'c5': ('HCQT(22050,hop256,360bins,60bpo,h=.5,1,2,3,4,5)+MelSpec(16000,2048,512,229) on 240 s tracks (configs[4] per-GPU shape)',
           [('HCQT', dict(sample_rate=22050, hop_length=256, n_bins=360, bins_per_octave=60), 22050),
            ('MelSpec', dict(sample_rate=16000, hop_length=512, n_mels=229, n_fft=2048), 16000)], 240.0, 8),
    'c2': ('MelSpec(16000,2048,512,229) on 64 x 20 s clips (configs[1])',
           [('MelSpec', dict(sample_rate=16000, hop_length=512, n_mels=229, n_fft=2048), 16000)], 20.0, 64),
    'c3': ('HCQT(22050,hop256,360bins,60bpo,6 harmonics) on 30 s clips (configs[2])',
           [('HCQT', dict(sample_rate=22050, hop_length=256, n_bins=360, bins_per_octave=60), 22050)], 30.0, 16),
    'c4': ('STFT(2048,512)+VQT(84,12,gamma auto)+SignalPower @22050 on 240 s tracks (configs[3])',
           [('STFT', dict(sample_rate=22050, hop_length=512, n_fft=2048), 22050),
            ('VQT', dict(sample_rate=22050, hop_length=512), 22050),
            ('SignalPower', dict(sample_rate=22050, hop_length=512), 22050)], 240.0, 4),
    'c1': ('CQT(22050,hop512,192bins,24bpo) on one 30 s clip (configs[0])',
           [('CQT', dict(sample_rate=22050, hop_length=512, n_bins=192, bins_per_octave=24), 22050)], 30.0, 1),
}


def synth_batch(sr, seconds, batch, seed0):
    from amt_tools_b200.synth import piano_like
    n = int(round(sr * seconds))
    uniq = min(batch, 2)  # two distinct tracks, alternated: generation time stays bounded
    tracks = [piano_like(n, sr, seed=seed0 + i) for i in range(uniq)]
    return np.stack([tracks[i % uniq] for i in range(batch)])


# --------------------------------------------------------------------------------------------------
# reference arm: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------

def _oracle_modules(spec, cache=None):
    from oracle import modules as om
    table = {'HCQT': om.OHCQT, 'MelSpec': om.OMelSpec, 'STFT': om.OSTFT, 'VQT': om.OVQT, 'CQT': om.OCQT,
             'SignalPower': om.OSignalPower}
    mods = []
    for name, kw, sr in spec:
        kw = dict(kw, dtype=np.float32)
        if cache is not None and name in ('HCQT', 'VQT', 'CQT'):
            kw['basis_cache'] = cache
        mods.append((table[name](**kw), sr))
    return mods


_AUDIO_CACHE = {}
_BASIS_CACHE = {}
_MODS_CACHE = {}


def _oracle_worker(args):
    """One pass of the oracle (float32, the precision the reference runs at) over one clip; returns the seconds it took.
    cached=True keeps the wavelet bases of the process between calls (librosa rebuilds them on every call: 0.3 s, 1 % of a
    240 s track but 12 % of a 20 s excerpt), so that the per-second cost of an excerpt equals that of a full track."""
    wl, seconds, seed, cached = args
    from threadpoolctl import threadpool_limits
    from amt_tools_b200.synth import piano_like
    key = (wl, cached)
    if key not in _MODS_CACHE:
        _MODS_CACHE[key] = _oracle_modules(WORKLOADS[wl][1], _BASIS_CACHE if cached else None)
    mods = _MODS_CACHE[key]
    for _, sr in mods:  # synthetic audio is generated once per worker process, outside the timed part
        if (sr, seconds) not in _AUDIO_CACHE:
            _AUDIO_CACHE[(sr, seconds)] = piano_like(int(round(sr * seconds)), sr, seed=seed)
    with threadpool_limits(limits=1):
        t = time.perf_counter()
        for m, sr in mods:
            m.process_audio(_AUDIO_CACHE[(sr, seconds)])
        return time.perf_counter() - t


def cpu_sample_seconds(wl):
    # bounded sample of the same workload: excerpts of the same tracks through the same modules (the cost is linear in the
    # duration once the per-call basis construction is taken out, see _oracle_worker)
    return {'c5': 20.0, 'c3': 30.0, 'c4': 30.0, 'c2': 20.0, 'c1': 30.0}[wl]


def workload_config(args, world):
    """The `config` object of the JSON line: the same for both arms (the reference arm runs a bounded sample OF this workload)."""
    desc, spec, seconds, default_batch = WORKLOADS[args.workload]
    B = args.batch or default_batch
    in_b = out_b = 0
    for m, sr in _oracle_modules(spec):          # integer frame arithmetic only (features/common.py:41-66 and overrides)
        n = int(round(sr * seconds))
        in_b += 4 * B * n
        out_b += 4 * B * m.num_channels * m.get_feature_size() * m.get_expected_frames(np.empty(n, dtype=np.float32))
    return {'workload': args.workload + ': ' + desc, 'tracks_per_gpu_per_step': B, 'track_seconds': seconds,
            'audio_hours_per_step': world * B * seconds / 3600.0,
            'l2': 'per-step working set (inputs %.0f MB + outputs %.0f MB per GPU) exceeds the 126 MB L2' % (in_b / 1e6, out_b / 1e6)
                  if in_b + out_b > 126e6 else 'per-step working set (inputs %.0f MB + outputs %.0f MB per GPU) fits the 126 MB L2: '
                  'consecutive steps write fresh output buffers' % (in_b / 1e6, out_b / 1e6),
            'parallelism': 'track-sharded x%d, no collective on the data path' % world}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    world = int(os.environ.get('WORLD_SIZE', '1'))
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))  # one single-threaded oracle process per host core (capped)
    sec = cpu_sample_seconds(args.workload)
    ctx = mp.get_context('fork')
    with ctx.Pool(procs) as pool:
        for _ in range(max(1, args.warmup)):
            pool.map(_oracle_worker, [(args.workload, sec, 100 + i, True) for i in range(procs)], chunksize=1)
        t0 = time.perf_counter()
        for s in range(args.steps):
            pool.map(_oracle_worker, [(args.workload, sec, 100 + i, True) for i in range(procs)], chunksize=1)
        dt = time.perf_counter() - t0
    hours = args.steps * procs * sec / 3600.0
    value = hours / dt
    sample = ('%d excerpts of %.0f s per step (one per host core) of the tracks of this workload, through the oracle float32 port of '
              'the librosa algorithm; wavelet bases kept between calls so that the per-second cost equals that of a full %.0f s track '
              '(librosa rebuilds them per call: 1 %% of a full track)' % (procs, sec, WORKLOADS[args.workload][2]))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'audio-hours/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, world),
        'cpu_baseline': {'value': value, 'unit': 'audio-hours/s', 'cores': procs, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'audio-hours/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------

class ClockSampler(object):
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap,utilization.gpu')

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, sm_busy, reasons, power = [], [], set(), []
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(',')]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                out['sm_max_mhz'] = float(f[2])
                try:
                    power.append(float(f[3]))
                    if len(f) > 9 and float(f[9]) > 0:
                        sm_busy.append(float(f[1]))
                except ValueError:
                    pass
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            use = sm_busy or sm
            out['sm_mhz'] = float(np.median(use))
            out['samples'] = len(sm)
            out['samples_under_load'] = len(sm_busy)
            if power:
                out['power_w_max'] = max(power)
        out['reasons'] = sorted(reasons)
        return out


def algorithmic_bytes(mod, n):
    """4 * N_in + 4 * C * F * T_out per clip."""
    T = mod.get_expected_frames(np.empty(n, dtype=np.float32))
    return 4 * n + 4 * mod.get_num_channels() * mod.get_feature_size() * T


def kernel_algorithmic_bytes(mod, name, n):
    """
    Algorithmic bytes ONE clip contributes to the kernel `name` (DESIGN.md "Kernels"): every input sample of the
    signal(s) it frames read once + every output element it produces written once.
    """
    d = mod.describe()
    T = mod.get_expected_frames(np.empty(n, dtype=np.float32))
    if name.startswith('stft_kernel'):
        return 4 * n + 4 * mod.get_feature_size() * T
    if name.startswith('cqt_kernel_nfft') or name == 'cqt_slide_kernel':
        # FFT-per-frame launches take the items of one n_fft that are not on the sliding-DFT kernel (deep ladder levels)
        nfft = int(name[len('cqt_kernel_nfft'):]) if name != 'cqt_slide_kernel' else None
        total = 0
        for it in d['items']:
            if it.get('alt'):
                continue        # exact-ladder copies of a few rows: first / last frames only (tail_* launches)
            if (nfft is None and it.get('slide')) or (it['n_fft'] == nfft and not it.get('slide')):
                total += 4 * int(np.ceil(n / 2.0 ** it['level'])) + 4 * it['rows'] * T
        return total
    if name in ('decimate_kernel', 'decimate_fft_kernel', 'decimate_fft64_kernel'):
        return sum(4 * int(np.ceil(n / 2.0 ** (l - 1))) + 4 * int(np.ceil(n / 2.0 ** l)) for l in range(1, d['n_levels']))
    if name == 'db_epilogue_kernel':
        return 8 * mod.get_num_channels() * mod.get_feature_size() * T
    if name == 'power_kernel':
        return 4 * n + 4 * T
    return 0


def algorithmic_flops(mod, n):
    """
    SURVEY.md 8(d) flop model for ONE clip of n samples (useful, shared work): real FFT of size m = 2.5 m log2 m, window m,
    power / magnitude 3 F, sparse projections 2 nnz (mel) / 8 nnz (complex wavelet rows, rows shared by octave-related
    harmonics counted once), sliding-DFT items (deep ladder levels) as executed, decimator as executed (fast-convolution form: three 1024-point complex transforms + the
    spectral product per two blocks of 1024 - D outputs, ~111 flop per output sample), dB 6 F.
    """
    d = mod.describe()
    T = mod.get_expected_frames(np.empty(n, dtype=np.float32))
    F, Cn = mod.get_feature_size(), mod.get_num_channels()
    db = 6.0 * Cn * F if mod.decibels else 0.0
    name = mod.features_name()
    if name in ('STFT', 'MelSpec'):
        m = mod.n_fft
        per = 2.5 * m * np.log2(m) + m + 3.0 * (m // 2 + 1) + db
        if name == 'MelSpec':
            per += 2.0 * 2.0 * (m // 2 + 1)     # every FFT bin feeds at most two triangular filters
        return T * per
    if name == 'SignalPower':
        return 2.0 * n + T * db
    if 'items' in d:
        # FFT-per-frame items: the SURVEY model; sliding-DFT items (deep levels) as executed: per band bin and frame
        # 4 hop flops for the entering / leaving samples + 44 for the phase bookkeeping, times the tile lead-in overhead
        fft = 0.0
        main = [it for it in d['items'] if not it.get('alt')]     # the exact-ladder copies cover a few hundred frames of a clip: not counted
        for it in main:
            kb = it['kmax_padded'] - it['kmin'] + 1
            if it.get('slide'):
                tile = min(1024, 4096 // it['hop'])
                fft += kb * (4.0 * it['hop'] + 44.0) * (1.0 + (it['n_fft'] // it['hop']) / float(tile))
            else:
                fft += 2.5 * it['n_fft'] * np.log2(it['n_fft']) + 3.0 * (it['kmax'] - it['kmin'] + 1)
        uniq = sum(it['unique_rows'] for it in main) / max(1, sum(it['rows'] for it in main))
        proj = 8.0 * d['basis_nnz'] * uniq
        dec_out = sum(int(np.ceil(n / 2.0 ** l)) for l in range(1, d['n_levels']))
        # fast-convolution forms as executed: float32 ~111 flop per output, float64 (default) ~107 (2048-point complex forward, spectral
        # product, fold, 1024-point inverse per 2 x 830 outputs) -- float64 operations counted like float32 ones
        dec = dec_out * (111.0 if d.get('decimator') == 'fft32' else 107.0 if d.get('decimator') == 'fft64' else 2.0 * d['decim_taps'])
        return T * (fft + proj + db) + dec
    return 0.0


def device_leg(ab, torch, dist, dev, wl, world, rank, steps, batch=None):
    """Device-resident throughput of one more named workload (BASELINE.json configs[1..3]): same stepping as the main leg
    (one stream per module, consecutive steps alternating between two stream sets), its own synthetic batch per rank."""
    desc, spec, seconds, default_batch = WORKLOADS[wl]
    B = batch or default_batch
    mods, audio = [], []
    for i, (name, kw, sr) in enumerate(spec):
        mods.append(getattr(ab, name)(device=dev, **kw))
        audio.append(torch.from_numpy(synth_batch(sr, seconds, B, seed0=7000 + 1000 * rank + 10 * i)).to(dev))
    n_per = [int(a.shape[1]) for a in audio]
    # L2 hygiene: the inputs rotate over enough copies that a step never finds its audio in the 126 MB L2, and the outputs of the
    # last `ncopy` steps stay alive so that the allocator hands out memory that has left the L2 as well
    in_bytes = sum(4 * a.numel() for a in audio)
    ncopy = max(2, int(np.ceil(3 * 126e6 / max(in_bytes, 1))))
    copies = [[a.clone() for a in audio] for _ in range(min(ncopy, 32))]
    sets = [[torch.cuda.Stream(dev) for _ in mods] for _ in range(2)]
    cur = torch.cuda.current_stream(dev)
    count = [0]
    alive = []

    def step():
        ss = sets[count[0] % 2]
        cp = copies[count[0] % len(copies)]
        count[0] += 1
        outs = []
        for m, a, st in zip(mods, cp, ss):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                outs.append(m.process_audio(a))
        alive.append(outs)
        if len(alive) > len(copies):
            alive.pop(0)
        return outs

    def join():
        for ss in sets:
            for st in ss:
                cur.wait_stream(st)

    t0 = time.perf_counter()
    n = 0
    while n < 3 or time.perf_counter() - t0 < 0.2:
        step()
        n += 1
        if n % 8 == 0:
            torch.cuda.synchronize()
    join()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    # best of three timed repeats of `steps` steps (one-off host hiccups showed up as 4x outliers on these sub-millisecond steps)
    reps = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        join()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        reps.append(float(ms.item()) / steps)
    ms_step = min(reps)
    step_bytes = sum(B * algorithmic_bytes(m, n) for m, n in zip(mods, n_per))
    step_flops = sum(B * algorithmic_flops(m, n) for m, n in zip(mods, n_per))
    return {'workload': wl + ': ' + desc, 'tracks_per_gpu_per_step': B, 'value': world * B * seconds / 3600.0 / (ms_step * 1e-3),
            'ms_per_step': ms_step, 'ms_per_step_repeats': reps, 'algorithmic_bytes_per_step': step_bytes, 'algorithmic_flops_per_step': step_flops,
            'l2': 'inputs rotate over %d device copies (%.0f MB) and the outputs of the last %d steps stay allocated: no step finds '
                  'its data in the 126 MB L2' % (len(copies), len(copies) * in_bytes / 1e6, len(copies))}


def longtrack_leg(ab, torch, dist, dev, world, rank, minutes=30.0):
    """ONE long track (the HCQT of c5) over all N GPUs (amt_tools_b200.longtrack: chunks dealt to the ranks, one all_reduce(MAX) of C
    floats -- the path's only exchange step), next to process_audio of the whole track on one GPU.  Audio resident on every rank."""
    from amt_tools_b200 import longtrack as lt
    name, kw, sr = WORKLOADS['c5'][1][0]
    n = int(sr * 60 * minutes)
    seg = synth_batch(sr, 60.0, 1, seed0=4242)[0]
    y = np.tile(seg, n // len(seg) + 1)[:n].copy()
    y *= (1.0 + 0.1 * np.sin(np.arange(n, dtype=np.float32) * 1e-6)).astype(np.float32)
    yd = torch.from_numpy(y).to(dev)
    m = getattr(ab, name)(device=dev, **kw)
    group = dist.group.WORLD if world > 1 else None

    def timed(fn):
        best = None
        for i in range(4):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if i:           # the first call builds plans / warms the allocator
                best = float(ms.item()) if best is None else min(best, float(ms.item()))
        return best, out

    ms_whole, whole = timed(lambda: m.process_audio(yd))
    T = int(m._out_shape(n)[-1])
    cf = max(16 * lt.ALIGN, -(-T // (4 * world)))      # explicit: on one rank the default is the whole-track call itself
    ms_chunk, own = timed(lambda: lt.process_long_audio(m, yd, chunk_frames=cf, group=group, gather=False))
    ms_user = ms_chunk if world > 1 else timed(lambda: lt.process_long_audio(m, yd, gather=False))[0]
    worst = 0.0
    for f0, f1, part in own.values():
        w = whole[..., f0:f1]
        worst = max(worst, float((part - w).abs()[w > 0.25].max()) * 80.0)
    wt = torch.tensor([worst], device=dev)
    if world > 1:
        dist.all_reduce(wt, op=dist.ReduceOp.MAX)
    hours = minutes / 60.0
    return {'workload': 'one %.0f-minute track, %s of c5, over %d GPU(s)' % (minutes, name, world),
            'value': hours / (ms_user * 1e-3), 'ms': ms_user,
            'chunked_value': hours / (ms_chunk * 1e-3), 'chunked_ms': ms_chunk, 'chunk_frames': cf,
            'whole_track_one_gpu_value': hours / (ms_whole * 1e-3), 'whole_track_one_gpu_ms': ms_whole,
            'note': 'value: process_long_audio as a user calls it on this many GPUs (one GPU: the whole-track call; more: the chunked path)',
            'max_db_diff_vs_whole_track_graded_bins': float(wt.item()), 'halo_frames': lt.halo_frames(m, 4096),
            'collective': 'all_reduce(MAX) of %d floats' % m.get_num_channels() if world > 1 else 'none (one rank)'}


def leg_fractions(r, sm_mhz, hbm_peak):
    """Fractions of the two rooflines for one workload leg: algorithmic bytes over the measured HBM peak, algorithmic flops over
    the FP32 FMA peak at the observed SM clock; the binding one is the larger time."""
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6
    t_hbm = r['algorithmic_bytes_per_step'] / (hbm_peak * 1e9) * 1e3
    t_fp32 = r['algorithmic_flops_per_step'] / fp32_peak * 1e3
    r.update(hbm_frac=t_hbm / r['ms_per_step'], fp32_frac=t_fp32 / r['ms_per_step'], roofline_ms=max(t_hbm, t_fp32),
             roofline_frac=max(t_hbm, t_fp32) / r['ms_per_step'], bound='hbm' if t_hbm >= t_fp32 else 'fp32')
    return r


def cache_leg(ab, torch, dev, spec, seconds, tracks, compressed, writers):
    """The caller BASELINE.json configs[4] names, end to end: TranscriptionDataset's feature precompute (datasets/common.py:212-295:
    compute -> np.savez(_compressed) -> <save_loc>/<Dataset>/<features_name()>/<track>.npz) through precompute_features."""
    import shutil
    from amt_tools_b200 import precompute
    from amt_tools_b200.synth import piano_like
    root = tempfile.mkdtemp(prefix='amtfeat_cache_')
    try:
        t_total, nbytes, stats_all = 0.0, 0, []
        for name, kw, sr in spec:
            m = getattr(ab, name)(device=dev, **kw)
            uniq = [piano_like(int(round(sr * seconds)), sr, seed=9000 + i) for i in range(2)]
            corpus = {'track_%04d' % i: uniq[i % 2] for i in range(tracks)}
            m.process_audio(uniq[0][:sr * 2])          # plan creation / first-launch costs stay outside the timed part
            torch.cuda.synchronize()
            stats = {}
            t = time.perf_counter()
            out = precompute.precompute_features(corpus, m, root, 'Synth', compressed=compressed, writers=writers, stats=stats)
            t_total += time.perf_counter() - t
            nbytes += sum(os.path.getsize(p) for p in out.values())
            stats_all.append(dict(stats, module=name))
        return {'value': tracks * seconds / 3600.0 / t_total, 'unit': 'audio-hours/s', 'tracks': tracks, 'seconds': t_total,
                'bytes_written': nbytes, 'write_GBps': nbytes / t_total / 1e9, 'compressed': bool(compressed), 'writers': writers,
                'dir': root, 'stages': stats_all}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import amt_tools_b200 as ab
    from amt_tools_b200 import _lib

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; amt_tools_b200 has no CPU compute path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    # one rank per GPU: run next to it, so the pinned staging buffers are NUMA-local to its PCIe root port
    from amt_tools_b200 import shard as _shard
    numa_cpus = None if args.no_bind else _shard.bind_host_to_gpu(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    desc, spec, seconds, default_batch = WORKLOADS[args.workload]
    B = args.batch or default_batch
    mods, dev_audio, host_audio = [], [], []
    for i, (name, kw, sr) in enumerate(spec):
        m = getattr(ab, name)(device=dev, **kw)
        a = synth_batch(sr, seconds, B, seed0=5000 + 1000 * rank + 10 * i)
        mods.append(m)
        host_audio.append(torch.from_numpy(a).pin_memory())
        dev_audio.append(host_audio[-1].to(dev))
    n_per = [int(a.shape[1]) for a in host_audio]
    hours_per_step = B * seconds / 3600.0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # one stream per module: the modules of a step are independent (amtfeat_process is re-entrant on distinct stream +
    # workspace), so e.g. the compute-bound mel kernel runs underneath the HBM-bound dB epilogue of the HCQT
    mod_streams = [torch.cuda.Stream(dev) for _ in mods] if len(mods) > 1 and not args.serial_modules else None
    # Consecutive steps are independent batches (amtfeat_process takes caller-owned output and workspace buffers), so they
    # alternate between two sets of streams: the HBM-bound dB epilogue that ends step i runs underneath the FP32-bound
    # kernels that start step i + 1.  --serial-steps joins every step on the current stream instead (A/B).
    step_sets = None
    if not args.serial_steps and not args.serial_modules:
        step_sets = [[torch.cuda.Stream(dev) for _ in mods] for _ in range(2)]
    step_no = [0]

    def join_steps():
        if step_sets is not None:
            cur = torch.cuda.current_stream(dev)
            for ss in step_sets:
                for s in ss:
                    cur.wait_stream(s)

    def step_device():
        outs = []
        if step_sets is not None:
            cur = torch.cuda.current_stream(dev)
            ss = step_sets[step_no[0] % 2]
            step_no[0] += 1
            for m, a, s in zip(mods, dev_audio, ss):
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    outs.append(m.process_audio(a))
            return outs
        if mod_streams is None:
            for m, a in zip(mods, dev_audio):
                outs.append(m.process_audio(a))
            return outs
        cur = torch.cuda.current_stream(dev)
        for m, a, s in zip(mods, dev_audio, mod_streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                outs.append(m.process_audio(a))
        for s in mod_streams:
            cur.wait_stream(s)
        return outs

    # ---------------- device-resident throughput ("value") ----------------
    # the clock sampler runs from the warm-up to the end of the e2e measurement (nvidia-smi answers every ~50 ms, the
    # timed region of a fast workload is shorter than that); the warm-up keeps the GPU loaded for >= 0.3 s
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    tw = time.perf_counter()
    nwarm = 0
    while nwarm < max(args.warmup, 3) or time.perf_counter() - tw < 0.3:
        outs = step_device()
        nwarm += 1
        if nwarm % 8 == 0:
            torch.cuda.synchronize()
    del outs
    join_steps()
    barrier()
    for m in mods:
        _lib.check(_lib.lib.amtfeat_profile_enable(m._dev_plan.handle, 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        outs = step_device()
        del outs
    join_steps()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    def read_profile(modules):
        res = {}
        for m in modules:
            buf = C.create_string_buffer(1 << 16)
            _lib.check(_lib.lib.amtfeat_profile_read(m._dev_plan.handle, buf, len(buf)))
            _lib.check(_lib.lib.amtfeat_profile_enable(m._dev_plan.handle, 0))
            for k, v in json.loads(buf.value.decode()).items():
                res[(m.features_name(), k)] = v
        return res

    prof_overlapped = read_profile(mods)
    # Isolated per-kernel durations for the roofline: in the timed region the kernels of different modules, of the ladder side
    # stream and of consecutive steps share the SMs, so an event pair around one launch also measures its neighbours.  A
    # second set of plans created with AMTFEAT_SERIAL=1 (no side stream) runs the same steps one launch at a time.
    os.environ['AMTFEAT_SERIAL'] = '1'
    smods = [getattr(ab, name)(device=dev, **kw) for (name, kw, sr) in spec]
    for m, a in zip(smods, dev_audio):
        m.process_audio(a)                      # builds the plan while the switch is set
    del os.environ['AMTFEAT_SERIAL']
    for _ in range(2):
        for m, a in zip(smods, dev_audio):
            m.process_audio(a)
    torch.cuda.synchronize()
    for m in smods:
        _lib.check(_lib.lib.amtfeat_profile_enable(m._dev_plan.handle, 1))
    iso_steps = max(3, min(args.steps, 10))
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record()
    for _ in range(iso_steps):
        for m, a in zip(smods, dev_audio):
            m.process_audio(a)
    i1.record()
    torch.cuda.synchronize()
    iso_ms_per_step = i0.elapsed_time(i1) / iso_steps
    prof = read_profile(smods)
    del smods
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * args.steps * hours_per_step / (ms_total / 1e3)

    # ---------------- end to end through the C-ABI host entry point ----------------
    e2e = None
    if not args.no_e2e:
        # amtfeat_pipeline_* (C-ABI): upload / compute / download streams chained with events over NSLOT staging-buffer sets, so
        # the device-to-host copy engine (the bottleneck: the float32 features are 5x the audio) always has a finished batch
        NSLOT = 4
        assert all(n % 4 == 0 for n in n_per)
        subs, max_in, max_out, max_ws = [], 0, 0, 0
        for m, ha, n in zip(mods, host_audio, n_per):
            per = int(np.prod(m._out_shape(n)))
            n_arr = _lib.i64_array([n] * B)
            max_ws = max(max_ws, int(_lib.lib.amtfeat_workspace_bytes(m._dev_plan.handle, B, n_arr)))
            max_in, max_out = max(max_in, B * n), max(max_out, B * per)
            subs.append(dict(m=m, h_in=ha, n_arr=n_arr, in_off=_lib.i64_array([b * n for b in range(B)]),
                             out_off=_lib.i64_array([b * per for b in range(B)]), per=per,
                             # two host result buffers per module: step i+1 must not overwrite what step i is still downloading
                             h_out=[torch.empty(B * per, dtype=torch.float32).pin_memory() for _ in range(2)]))
        pipe = C.c_void_p()
        _lib.check(_lib.lib.amtfeat_pipeline_create(local, NSLOT, max_in, max_out, max_ws, C.byref(pipe)))

        def step_host(i):
            for s in subs:
                _lib.check(_lib.lib.amtfeat_pipeline_submit(
                    pipe, s['m']._dev_plan.handle, s['h_in'].data_ptr(), s['in_off'], s['n_arr'], s['out_off'], B,
                    s['h_out'][i % 2].data_ptr(), s['h_in'].numel(), s['h_out'][i % 2].numel(), None))

        for i in range(max(3, min(args.warmup, 4))):
            step_host(i)
        _lib.check(_lib.lib.amtfeat_pipeline_wait(pipe, -1))
        barrier()
        # device clock: an event on the (idle) default stream before the first submit, one after wait(all) has returned.  The region of
        # exactly `steps` steps is timed twice and the better one is reported (both are listed): this leg runs at the host's PCIe /
        # memory rate, and a busy host showed up as a 2x outlier once (profiles/SUMMARY_r02.md)
        e2e_reps = []
        for _ in range(2):
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            g0.record(torch.cuda.default_stream(dev))
            for i in range(args.steps):
                step_host(i)
            _lib.check(_lib.lib.amtfeat_pipeline_wait(pipe, -1))
            g1.record(torch.cuda.default_stream(dev))
            g1.synchronize()
            wall = 1e3 * (time.perf_counter() - t0)
            barrier()
            rep = torch.tensor([g0.elapsed_time(g1)], device=dev)
            if world > 1:
                dist.all_reduce(rep, op=dist.ReduceOp.MAX)
            e2e_reps.append((float(rep.item()), wall))
        ems_best, wall_ms = min(e2e_reps)
        ems = torch.tensor([ems_best], device=dev)
        checksum = float(subs[0]['h_out'][(args.steps - 1) % 2][:1024].sum())  # the host really holds the features
        e2e = {
            'value': world * args.steps * hours_per_step / (float(ems.item()) / 1e3), 'unit': 'audio-hours/s',
            'h2d_bytes_per_step': int(sum(4 * B * n for n in n_per)),
            'd2h_bytes_per_step': int(sum(4 * s['h_out'][0].numel() for s in subs)),
            'ms_per_step': float(ems.item()) / args.steps, 'wall_ms_per_step_rank0': wall_ms / args.steps,
            'ms_per_step_repeats': [r[0] / args.steps for r in e2e_reps],
            'path': 'amtfeat_pipeline_submit / _wait (C-ABI): pinned host audio -> H2D -> kernels -> D2H of the full float32 '
                    'features, upload / compute / download streams over %d staging slots' % NSLOT, 'checksum': checksum,
        }
        _lib.lib.amtfeat_pipeline_destroy(pipe)
        del subs

        # The other consumer the north_star names: a model on the same GPU (OnsetsFrames / TabCNN pre_proc) takes the
        # features as device tensors.  Public Python API, pinned host audio in, features stay resident, and the step's
        # result read back is one scalar per track (mean feature value).
        NRS = max(2, args.consumer_streams)     # steps in flight: the upload of step i + 1 (+ 2 ...) overlaps the kernels of step i
        rstreams = [torch.cuda.Stream(dev) for _ in range(NRS)]
        rhost = [torch.empty(len(mods), B, dtype=torch.float32).pin_memory() for _ in range(NRS)]

        def step_resident(i):
            with torch.cuda.stream(rstreams[i % NRS]):
                res = []
                for m, ha in zip(mods, host_audio):
                    f = m.process_audio(ha.to(dev, non_blocking=True))
                    res.append(f.reshape(B, -1).mean(dim=1))
                rhost[i % NRS].copy_(torch.stack(res), non_blocking=True)

        for i in range(4):
            step_resident(i)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for st in rstreams:
            st.wait_stream(torch.cuda.current_stream(dev))
        for i in range(args.steps):
            step_resident(i)
        for st in rstreams:
            torch.cuda.current_stream(dev).wait_stream(st)
        r1.record()
        barrier()
        last = rhost[(args.steps - 1) % NRS]
        rms = torch.tensor([r0.elapsed_time(r1)], device=dev)
        if world > 1:
            dist.all_reduce(rms, op=dist.ReduceOp.MAX)
        # the same consumer fed with 16-bit PCM (what the audio files hold): half the upload, converted on the device
        pcm_scale = [float(np.abs(a.numpy()).max()) / 32767.0 for a in host_audio]
        host_pcm = [torch.round(a / sc).to(torch.int16).pin_memory() for a, sc in zip(host_audio, pcm_scale)]

        def step_resident_pcm(i):
            with torch.cuda.stream(rstreams[i % NRS]):
                res = []
                for m, hp, sc in zip(mods, host_pcm, pcm_scale):
                    f = m.process_audio(ab.pcm16_to_float(hp.to(dev, non_blocking=True), sc, device=dev))
                    res.append(f.reshape(B, -1).mean(dim=1))
                rhost[i % NRS].copy_(torch.stack(res), non_blocking=True)

        for i in range(4):
            step_resident_pcm(i)
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for st in rstreams:
            st.wait_stream(torch.cuda.current_stream(dev))
        for i in range(args.steps):
            step_resident_pcm(i)
        for st in rstreams:
            torch.cuda.current_stream(dev).wait_stream(st)
        q1.record()
        barrier()
        qms = torch.tensor([q0.elapsed_time(q1)], device=dev)
        if world > 1:
            dist.all_reduce(qms, op=dist.ReduceOp.MAX)
        e2e['device_consumer_pcm16'] = {
            'value': world * args.steps * hours_per_step / (float(qms.item()) / 1e3), 'unit': 'audio-hours/s',
            'h2d_bytes_per_step': int(sum(2 * B * n for n in n_per)), 'd2h_bytes_per_step': int(4 * B * len(mods)),
            'ms_per_step': float(qms.item()) / args.steps,
            'path': 'pinned host int16 PCM -> H2D -> amtfeat_pcm16_to_float -> kernels; features stay on the device for the model',
        }
        # what the host side of the box gives all ranks at once: every rank copies device -> pinned host concurrently
        cbuf_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        cbuf_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        cbuf_h.copy_(cbuf_d)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(8):
            cbuf_h.copy_(cbuf_d, non_blocking=True)
        c1.record()
        barrier()
        cms = torch.tensor([c0.elapsed_time(c1)], device=dev)
        if world > 1:
            dist.all_reduce(cms, op=dist.ReduceOp.MAX)
        e2e['d2h_ceiling_GBps'] = world * 8 * (256 << 20) / (float(cms.item()) * 1e-3) / 1e9
        e2e['d2h_achieved_GBps'] = world * e2e['d2h_bytes_per_step'] / (e2e['ms_per_step'] * 1e-3) / 1e9
        del cbuf_d, cbuf_h
        e2e['device_consumer'] = {
            'value': world * args.steps * hours_per_step / (float(rms.item()) / 1e3), 'unit': 'audio-hours/s',
            'h2d_bytes_per_step': int(sum(4 * B * n for n in n_per)), 'd2h_bytes_per_step': int(4 * B * len(mods)),
            'ms_per_step': float(rms.item()) / args.steps,
            'path': 'FeatureModule.process_audio (Python API): pinned host audio -> H2D -> kernels; features stay on the '
                    'device for the model, one float per track and module read back; %d steps in flight (one stream each)' % NRS, 'checksum': float(last.sum()),
        }
    # ---------------- the other named shapes (BASELINE.json configs[1..3]), device resident, same run ----------------
    extra = None
    if args.workload == 'c5' and not args.no_extra:
        extra = {}
        for wl in ('c2', 'c3', 'c4'):
            extra[wl] = device_leg(ab, torch, dist, dev, wl, world, rank, max(5, min(args.steps, 20)))
    longtrack = None
    if args.workload == 'c5' and not args.no_extra:
        try:
            longtrack = longtrack_leg(ab, torch, dist, dev, world, rank)
        except Exception as e:      # the line must still go out; the error is the same on every rank (no rank is left in a collective)
            longtrack = {'error': '%s: %s' % (type(e).__name__, e)}
    # ---------------- the configs[4] caller end to end: precompute + npz cache (one GPU only: every rank would hit one disk) ----
    cache = None
    if args.workload == 'c5' and world == 1 and not args.no_cache and not args.no_e2e:
        writers = max(1, min(16, (os.cpu_count() or 2) - 1))
        cache = {'savez': cache_leg(ab, torch, dev, spec, seconds, args.cache_tracks, False, writers),
                 'savez_compressed': cache_leg(ab, torch, dev, spec, seconds, max(2, args.cache_tracks // 4), True, writers),
                 'note': 'precompute_features (amtfeat_pipeline_* + writer threads): every track through HCQT and MelSpec, one .npz per '
                         'track and module in the reference cache layout; np.savez_compressed is the reference default and is bound by zlib on the host cores'}
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks['window'] = 'warm-up + device-timed steps + e2e steps + the other workloads (nvidia-smi -lms 50)'

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel + CPU baseline ----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
    kern = {}
    for (mname, k), v in prof.items():
        m = mods[[x.features_name() for x in mods].index(mname)]
        n = n_per[mods.index(m)]
        per_launch_ms = v['ms'] / max(v['launches'], 1)
        # kernel_algorithmic_bytes covers everything this kernel name does in one step; a step may spread it over several
        # launches (one per ladder level; one per ladder-depth class of an n_fft), so average over the launches of a step
        bytes_per_launch = B * kernel_algorithmic_bytes(m, k, n) / max(1.0, v['launches'] / float(iso_steps))
        ov = prof_overlapped.get((mname, k))
        kern[mname + '.' + k] = {'ms_per_step': v['ms'] / iso_steps, 'launches_per_step': v['launches'] / float(iso_steps),
                                 'avg_ms': per_launch_ms, 'share_of_isolated_step': v['ms'] / (iso_ms_per_step * iso_steps),
                                 'ms_per_step_in_timed_region': (ov['ms'] / args.steps) if ov else None,
                                 'algorithmic_GBps': bytes_per_launch / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else None}
    top = max(kern.items(), key=lambda kv: kv[1]['ms_per_step'])
    # measured DRAM traffic per launch of that kernel (ncu --set full capture of this workload, profiles/r01_traffic.json)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json' if os.path.exists(os.path.join(ROOT, 'profiles', 'r02_traffic.json')) else 'r01_traffic.json')))['traffic_bytes_per_launch']
        ncu_name = {'cqt_kernel_nfft1024': 'void cqt_kernel<512, 1>(CqtParams)', 'cqt_kernel_nfft512': 'void cqt_kernel<256, 1>(CqtParams)',
                    'cqt_slide_kernel': 'cqt_slide_kernel(SlideParams)',
                    'stft_kernel_mel': 'void stft_kernel<1024, 1>(StftParams)'}.get(top[0].split('.')[1])
        cap_b = tj.get('captured_tracks_per_step', {}).get(args.workload)
        traffic = tj.get(args.workload, {}).get(ncu_name)
        if traffic is not None and cap_b:
            # the capture holds the per-step DRAM total of this kernel for `cap_b` tracks per step; every byte belongs to one
            # track (audio in, features out), so the total scales with the tracks of a step
            traffic = traffic * (B / float(cap_b)) / max(1.0, top[1]['launches_per_step'])
        else:
            traffic = None
    except Exception:
        pass
    roofline = {'kernel': top[0], 'bound': 'hbm', 'achieved': top[1]['algorithmic_GBps'], 'peak': hbm_peak, 'unit': 'GB/s',
                'frac': (top[1]['algorithmic_GBps'] or 0.0) / hbm_peak, 'traffic': traffic, 'peak_source': peak_src,
                'note': 'FP32-SIMT/shared-memory bound kernel (SURVEY.md 8d): the HBM fraction is reported as required; '
                        'roofline.fp32 is the compute roofline. Per-kernel durations are CUDA-event pairs from an isolated pass '
                        '(same steps, plans created with AMTFEAT_SERIAL=1: one launch at a time on one stream); in the timed '
                        'region the same kernels overlap across modules, the ladder side stream and consecutive steps '
                        '(ms_per_step_in_timed_region). `traffic` is the ncu per-step DRAM total of that kernel divided by '
                        'its launches per step',
                'isolated_ms_per_step': iso_ms_per_step, 'kernels': kern}
    step_bytes = sum(B * algorithmic_bytes(m, n) for m, n in zip(mods, n_per))
    roofline['step_algorithmic_GBps'] = step_bytes / (ms_total / args.steps * 1e-3) / 1e9
    # HBM-bound floor of the whole step (every input read once, every output written once, plus the dB epilogue's
    # second pass over the output) next to the measured step: how far the FP32-bound kernels sit above it
    out_bytes = step_bytes - sum(4 * B * n for n in n_per)
    hbm_floor_ms = (step_bytes + 2 * out_bytes) / (hbm_peak * 1e9) * 1e3
    # measured DRAM traffic of the whole step over its algorithmic bytes (in + out): producers from the ncu capture, plus the dB
    # second pass -- 2 x out in place (device-resident consumers), 1 x out when it is fused into the download (pipelined executor)
    try:
        tfile = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))
        prod = tfile['step_total_producers'][args.workload] * (B / float(tfile['traffic_bytes_per_launch']['captured_tracks_per_step'][args.workload]))
        db_out = sum(4 * B * m.get_num_channels() * m.get_feature_size() * m.get_expected_frames(np.zeros(n, dtype=np.float32))
                     for m, n in zip(mods, n_per) if getattr(m, 'decibels', False))
        roofline['step_traffic_ratio'] = (prod + 2 * db_out) / step_bytes
        roofline['e2e_traffic_ratio'] = (prod + db_out) / step_bytes
        roofline['producers_traffic_ratio'] = prod / step_bytes
    except Exception:
        pass
    roofline['step_hbm_floor_ms'] = hbm_floor_ms
    roofline['step_frac_of_hbm_floor'] = hbm_floor_ms / (ms_total / args.steps)
    # the relevant compute roofline (every configuration is FP32-bound): algorithmic flops of the step over
    # the FP32 FMA peak of the part at the SM clock observed during the run
    step_flops = sum(B * algorithmic_flops(m, n) for m, n in zip(mods, n_per))
    sm_mhz = clocks.get('sm_mhz') or clocks.get('sm_max_mhz') or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    fp32_ach = step_flops / (ms_total / args.steps * 1e-3) / 1e12
    roofline['fp32'] = {'achieved': fp32_ach, 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': fp32_ach / fp32_peak,
                        'flops_per_step': step_flops,
                        'model': 'SURVEY.md 8(d) algorithmic flops (FFT 2.5 n log2 n, sparse projections; sliding-DFT items '
                                 'and decimator as executed); peak = 148 SMs x 128 FMA lanes x 2 x %.0f MHz' % sm_mhz}

    if extra:
        for r in extra.values():
            leg_fractions(r, sm_mhz, hbm_peak)
    cpu = None
    if not args.no_cpu:
        sec = seconds            # one full track / clip of this very workload, bases rebuilt per call like librosa does
        t = _oracle_worker((args.workload, sec, 77, False))
        cpu = {'value': sec / 3600.0 / t, 'unit': 'audio-hours/s', 'cores': 1, 'kind': 'port',
               'sample': 'one full %.0f s track of this workload through the oracle float32 port of the same modules (%.1f s of CPU)' % (sec, t)}

    launches = sum(int(_lib.lib.amtfeat_launch_count(m._dev_plan.handle, B, _lib.i64_array([n] * B)))
                   for m, n in zip(mods, n_per))
    flat = {'fp32_frac': roofline['fp32']['frac'], 'hbm_frac': roofline['frac'],
            'step_traffic_ratio': roofline.get('step_traffic_ratio'), 'e2e_traffic_ratio': roofline.get('e2e_traffic_ratio'),
            'e2e_device_value': (e2e or {}).get('device_consumer', {}).get('value') if e2e else None,
            'e2e_device_pcm16_value': (e2e or {}).get('device_consumer_pcm16', {}).get('value') if e2e else None,
            'e2e_d2h_ceiling_gbs': (e2e or {}).get('d2h_ceiling_GBps') if e2e else None,
            'e2e_d2h_achieved_gbs': (e2e or {}).get('d2h_achieved_GBps') if e2e else None}
    if extra:
        for wl, r in extra.items():
            for k in ('value', 'ms_per_step', 'hbm_frac', 'fp32_frac', 'roofline_frac'):
                flat['%s_%s' % (wl, k)] = r[k]
    if longtrack and 'value' in longtrack:
        flat['longtrack_value'] = longtrack['value']
        flat['longtrack_whole_track_one_gpu_value'] = longtrack['whole_track_one_gpu_value']
    if cache is not None:
        flat['e2e_cache_value'] = cache['savez']['value']
        flat['e2e_cache_compressed_value'] = cache['savez_compressed']['value']
    line = {
        'metric': METRIC, 'value': value, 'unit': 'audio-hours/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, world),
        'run_config': {'host_binding': ('rank pinned to the %d CPU cores next to its GPU (NVML affinity)' % len(numa_cpus)) if numa_cpus else 'none',
                       'streams': ('one CUDA stream per module, consecutive steps alternate between two stream sets (joined at the end of the timed region)'
                                   if step_sets else 'one CUDA stream per module, joined every step' if mod_streams else 'single stream')},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches * args.steps, 'roofline': roofline, 'cpu_baseline': cpu,
        'workloads': extra, 'longtrack': longtrack, 'cache': cache,
    }
    line.update(flat)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) was diverted to stderr."""
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)   # libraries that print to fd 1 (e.g. "NCCL version ...") must not pollute the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c5', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0, help='tracks per GPU per step (0 = workload default)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--serial-modules', action='store_true', help='run the modules of a step on one stream (A/B)')
    ap.add_argument('--serial-steps', action='store_true', help='join every step on the current stream (A/B for the step pipelining)')
    ap.add_argument('--no-bind', action='store_true', help='do not pin the rank to the CPU cores next to its GPU (A/B)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--consumer-streams', type=int, default=4, help='steps in flight of the device-consumer legs (one CUDA stream each)')
    ap.add_argument('--no-extra', action='store_true', help='skip the device-resident legs of the other named workloads (c2, c3, c4)')
    ap.add_argument('--no-cache', action='store_true', help='skip the precompute + npz cache leg')
    ap.add_argument('--cache-tracks', type=int, default=16, help='tracks of the precompute + cache leg (np.savez; a quarter of them compressed)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
