"""
One long track computed as chunks (amt_tools_b200/longtrack.py): the chunk geometry, the one exchange step
(the per-channel maximum over all chunks, all_reduce(MAX): world-size-2 gloo on the CPU with the oracle standing in for the
kernels) and -- on the GPU -- the native path against `process_audio` of the whole track and against the oracle.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import amt_tools_b200 as ab
from amt_tools_b200 import longtrack as lt
from amt_tools_b200.synth import piano_like
from oracle import modules as om


@pytest.mark.parametrize('n,T,hop,chunk,halo', [(1323000, 5168, 256, 1024, 448), (320000, 626, 512, 128, 64), (1000, 2, 512, 64, 64),
                                                 (5292000, 20672, 256, 2000, 448), (0, 0, 512, 64, 64)])
def test_chunk_plan_geometry(n, T, hop, chunk, halo):
    plan = lt.chunk_plan(n, T, hop, chunk, halo)
    assert (not plan and T == 0) or ([p[0] for p in plan] == [0] + [p[1] for p in plan[:-1]] and plan[-1][1] == T)   # a partition of the frames
    for f0, f1, a, b, k0 in plan:
        assert f0 % lt.ALIGN == 0 and a % hop == 0 and a % (lt.ALIGN * hop) == 0       # cuts on multiples of every level's hop
        assert 0 <= a <= b <= n and k0 == f0 - a // hop
        assert a == 0 or k0 >= halo                                                      # a full halo, or the track's own edge
        assert b == n or b - f1 * hop >= halo * hop
    assert all(p[1] - p[0] >= min(halo, T) for p in plan)                                # no chunk that is all halo


def test_chunk_plan_invariants_on_random_geometry():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(T=st.integers(0, 400000), hop=st.sampled_from([64, 256, 512, 1024]), chunk=st.integers(1, 100000),
           halo=st.integers(1, 16).map(lambda k: k * lt.ALIGN), extra=st.integers(0, 1023))
    def check(T, hop, chunk, halo, extra):
        n = max(0, (T - 1) * hop + extra) if T else 0        # a centred module: T = 1 + n // hop
        plan = lt.chunk_plan(n, T, hop, chunk, halo)
        covered = 0
        for f0, f1, a, b, k0 in plan:
            assert f0 == covered and f1 > f0 and f0 % lt.ALIGN == 0
            covered = f1
            assert a % (lt.ALIGN * hop) == 0 and 0 <= a <= b <= n and k0 == f0 - a // hop
            assert (a == 0 and k0 == f0) or k0 == halo
            assert b == n or b == (f1 + halo) * hop
            assert f1 == T or T - f1 >= halo                   # what is left for the next chunk is more than a halo
        assert covered == T

    check()


def test_halo_covers_the_ladder_and_the_exact_pieces():
    mel = ab.MelSpec(sample_rate=16000, hop_length=512, n_fft=2048)
    assert lt.halo_frames(mel, 1024) == lt.ALIGN                                          # n_fft / hop = 4 frames, rounded up
    hcqt = ab.HCQT(sample_rate=22050, hop_length=256, n_bins=360, bins_per_octave=60)
    h = lt.halo_frames(hcqt, 2048)
    # deepest level: n_fft 1024 at 2^7 plus the seven filter half lengths -> 90 174 samples = 353 frames; the exact ladder of h = 0.5
    # reaches 352 frames from the start and ~370 from the end of a clip
    assert h % lt.ALIGN == 0 and 384 <= h <= 512


def test_rejects_combos_and_batches():
    mel = ab.MelSpec(sample_rate=16000, hop_length=512)
    with pytest.raises(TypeError):
        lt.process_long_audio(ab.FeatureCombo([mel]), np.zeros(16000, dtype=np.float32))
    with pytest.raises(ValueError):
        lt.process_long_audio(mel, np.zeros((2, 16000), dtype=np.float32))


class _OracleOps:
    """The three steps of a chunk on the CPU (float64 oracle): what _CudaOps does through the C-ABI."""

    def __init__(self, omod, C, F):
        self.o, self.C, self.F, self.device = omod, C, F, torch.device('cpu')

    def raw(self, chunks):
        out = []
        for chunk in chunks:
            S = np.asarray(self.o.process_audio(np.asarray(chunk, dtype=np.float64)))
            out.append(torch.from_numpy(10.0 * np.log10(np.maximum(1e-10, S)).reshape(self.C, self.F, -1)))
        return out

    def reference(self, block, k0, k1, ref):
        ref.copy_(torch.maximum(ref, block[..., k0:k1].amax(dim=(1, 2)).to(ref.dtype)))

    def finish(self, block, k0, k1, ref, dst=None, t_dst=0):
        v = block[..., k0:k1] - ref.to(block.dtype)[:, None, None]
        v = (torch.clamp(v, min=-80.0) / 80 + 1).to(torch.float32)
        if dst is None:
            return v
        dst[..., t_dst:t_dst + (k1 - k0)] = v
        return dst


def test_one_rank_chunks_equal_the_whole_track_on_the_oracle():
    sr, hop = 16000, 512
    audio = piano_like(sr * 8, sr, seed=6)
    mod = ab.MelSpec(sample_rate=sr, hop_length=hop, n_fft=2048, n_mels=40)
    ops = _OracleOps(om.OMelSpec(sample_rate=sr, hop_length=hop, n_fft=2048, n_mels=40, decibels=False), 1, 40)
    full = lt.process_long_audio(mod, audio, chunk_frames=128, ops=ops).numpy()
    want = om.OMelSpec(sample_rate=sr, hop_length=hop, n_fft=2048, n_mels=40).process_audio(audio.astype(np.float64))
    assert full.shape == want.shape and np.abs(full - want).max() * 80 < 1e-4 and full.max() == 1.0


class _OracleAmplitudeOps(_OracleOps):
    """Same, for the modules the reference converts with amplitude_to_db (amin 1e-5 on the magnitude)."""

    def raw(self, chunks):
        out = []
        for chunk in chunks:
            S = np.asarray(self.o.process_audio(np.asarray(chunk, dtype=np.float64)))
            out.append(torch.from_numpy(20.0 * np.log10(np.maximum(1e-5, S)).reshape(self.C, self.F, -1)))
        return out


@pytest.mark.parametrize('name,kw,chunk', [
    ('VQT', dict(sample_rate=22050, hop_length=512), 256),
    # h = 0.5 is early-downsampled by 4 in ONE resample call (its own, longer filter): the halo must cover that reach, too
    ('HCQT', dict(sample_rate=22050, hop_length=256, n_bins=72, bins_per_octave=12, harmonics=[0.5, 1, 2]), 512),
])
def test_halo_is_sufficient_for_the_ladder_on_the_oracle(name, kw, chunk):
    """The halo geometry checked on the ALGORITHM, independent of the kernels: with the float64 oracle computing the chunks, the
    frames a chunk keeps equal the whole-track frames down to the float32 rounding of the stored result (every decimation
    filter of the ladder is a finite FIR, so beyond the halo a frame cannot know where the clip was cut)."""
    sr = kw['sample_rate']
    audio = piano_like(sr * 20, sr, seed=21)
    audio[: len(audio) // 2] *= 0.1
    m = getattr(ab, name)(**kw)
    ops = _OracleAmplitudeOps(getattr(om, 'O' + name)(decibels=False, **kw), m.get_num_channels(), m.get_feature_size())
    halo = lt.halo_frames(m, chunk)
    got = lt.process_long_audio(m, audio, chunk_frames=chunk, ops=ops).numpy()
    want = np.asarray(getattr(om, 'O' + name)(**kw).process_audio(audio.astype(np.float64)))
    assert len(lt.chunk_plan(len(audio), got.shape[-1], kw['hop_length'], max(chunk, 2 * halo), halo)) >= 3
    assert got.shape == want.shape and np.abs(got - want).max() * 80 < 2e-5
    # and a halo that is too short does show: the check above is not vacuous
    short = lt.process_long_audio(m, audio, chunk_frames=chunk, halo=0, ops=ops).numpy()
    assert np.abs(short - want).max() * 80 > 1e-2


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sr, hop = 16000, 512
    audio = piano_like(sr * 12, sr, seed=5)
    audio[: sr * 6] *= 0.05                     # the loud half lies in the other rank's chunks: the reference must cross ranks
    mod = ab.MelSpec(sample_rate=sr, hop_length=hop, n_fft=2048, n_mels=64)          # host-only plan: shapes and geometry
    ops = _OracleOps(om.OMelSpec(sample_rate=sr, hop_length=hop, n_fft=2048, n_mels=64, decibels=False), 1, 64)
    full = lt.process_long_audio(mod, audio, chunk_frames=128, group=dist.group.WORLD, ops=ops)
    own = lt.process_long_audio(mod, audio, chunk_frames=128, group=dist.group.WORLD, ops=ops, gather=False)
    if rank == 0:
        q.put((full.numpy(), sorted(own)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_chunks_equal_the_whole_track():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, own = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sr, hop = 16000, 512
    audio = piano_like(sr * 12, sr, seed=5)
    audio[: sr * 6] *= 0.05
    want = om.OMelSpec(sample_rate=sr, hop_length=hop, n_fft=2048, n_mels=64).process_audio(audio.astype(np.float64))
    assert full.shape == want.shape == (1, 64, 1 + len(audio) // hop)
    assert own == [0, 2]                                      # three chunks of 128 frames, dealt round-robin
    assert np.abs(full - want).max() * 80 < 1e-4              # float32 storage of a float64 computation
    assert full.max() == 1.0                                  # the track's maximum is the reference: 0 dB -> 1.0


# ---------------------------------------------------------------------------------------------------------------------------
# GPU: the native path
# ---------------------------------------------------------------------------------------------------------------------------

CASES = [
    ('MelSpec', dict(sample_rate=16000, hop_length=512, n_fft=2048), 40, 256),
    ('STFT', dict(sample_rate=16000, hop_length=512, n_fft=2048), 40, 256),
    ('SignalPower', dict(sample_rate=16000, hop_length=512), 40, 256),
    ('VQT', dict(sample_rate=22050, hop_length=512), 60, 640),
    ('CQT', dict(sample_rate=22050, hop_length=512, n_bins=192, bins_per_octave=24), 60, 640),
    ('HCQT', dict(sample_rate=22050, hop_length=256, n_bins=360, bins_per_octave=60), 60, 1024),
]


def _graded(a, b):
    """Largest dB difference over the bins within 60 dB of the channel maximum (the bar of tests/test_gpu_parity.py)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    sel = b > 1.0 - 60.0 / 80.0
    return float(np.abs(a - b)[sel].max() * 80.0)


@pytest.mark.gpu
@pytest.mark.parametrize('name,kw,seconds,chunk', CASES)
def test_chunked_track_equals_whole_track(name, kw, seconds, chunk):
    dev = torch.device('cuda', 0)
    sr = kw['sample_rate']
    audio = piano_like(sr * seconds, sr, seed=11)
    audio[: len(audio) // 2] *= 0.1                            # the maximum sits in the second half of the track
    yd = torch.from_numpy(audio).to(dev)
    for decibels in (True, False):
        m = getattr(ab, name)(decibels=decibels, device=dev, **kw)
        whole = m.process_audio(yd)
        got = lt.process_long_audio(m, yd, chunk_frames=chunk)
        plan = lt.chunk_plan(len(audio), int(whole.shape[-1]), kw['hop_length'], max(chunk, 2 * lt.halo_frames(m, chunk)), lt.halo_frames(m, chunk))
        assert len(plan) >= 3 and got.shape == whole.shape and bool(torch.isfinite(got).all())
        w, g = whole.double().cpu().numpy(), got.double().cpu().numpy()
        if decibels:
            scale = 80.0 if name != 'SignalPower' else 1.0
            sel = w > (1.0 - 60.0 / 80.0 if name != 'SignalPower' else -60.0)
            assert np.abs(g - w)[sel].max() * scale < 1e-3, (name, np.abs(g - w)[sel].max() * scale)
            assert float(got.max()) == float(whole.max())     # same reference: the maximum is exactly 1.0 (0 dB) on both paths
        else:
            assert np.linalg.norm(g - w) / np.linalg.norm(w) < 1e-5, (name, np.linalg.norm(g - w) / np.linalg.norm(w))


@pytest.mark.gpu
def test_chunked_melspec_against_the_oracle():
    dev = torch.device('cuda', 0)
    sr, hop = 16000, 512
    audio = piano_like(sr * 30, sr, seed=12)
    m = ab.MelSpec(sample_rate=sr, hop_length=hop, n_fft=2048, device=dev)
    got = lt.process_long_audio(m, audio, chunk_frames=192).double().cpu().numpy()
    want = om.OMelSpec(sample_rate=sr, hop_length=hop, n_fft=2048).process_audio(audio.astype(np.float64))
    assert got.shape == want.shape and _graded(got, want) < 1e-3


@pytest.mark.gpu
def test_range_entry_points_reject_bad_ranges():
    dev = torch.device('cuda', 0)
    m = ab.MelSpec(sample_rate=16000, hop_length=512, device=dev)
    ops = lt._CudaOps(m)
    block = ops.raw([torch.zeros(16000, device=dev)])[0]
    ref = torch.full((1,), float('-inf'), device=dev)
    with pytest.raises(ValueError):
        ops.reference(block, 0, int(block.shape[-1]) + 1, ref)
    with pytest.raises(ValueError):
        ops.finish(block, 5, 3, ref)


@pytest.mark.gpu
def test_one_rank_default_is_the_whole_track_call():
    dev = torch.device('cuda', 0)
    m = ab.HCQT(sample_rate=22050, hop_length=256, n_bins=360, bins_per_octave=60, device=dev)
    yd = torch.from_numpy(piano_like(22050 * 20, 22050, seed=13)).to(dev)
    whole = m.process_audio(yd)
    assert torch.equal(lt.process_long_audio(m, yd), whole)                     # no chunk size asked for, one rank: one call
    own = lt.process_long_audio(m, yd, gather=False)
    assert list(own) == [0] and own[0][:2] == (0, whole.shape[-1]) and torch.equal(own[0][2], whole.reshape(own[0][2].shape))
