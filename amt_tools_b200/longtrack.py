"""
One long track computed as chunks -- on several GPUs, or chunk after chunk on one ("a single very long
track is chunked on frame boundaries with a halo; the global max then needs one scalar max across chunks").

`process_audio` (features/common.py:168-230 and every override) is a pure function of ONE track: frames only see the
samples under their windows (and, for the CQT family, under the finite decimation filters of the ladder), but the dB
post-processing is referenced to the maximum of the whole track (`ref=np.max`, common.py:199, 224-225).  So a track can be
cut on frame boundaries, every chunk extended by a halo of frames that are computed and thrown away, and the chunks
computed apart -- up to ONE exchange: the per-channel maximum over all chunks (C floats, `all_reduce(MAX)`), the only
collective this path has.  Each chunk is then finished against the track's reference while it moves to its place.

    feats = process_long_audio(module, audio, chunk_frames=16384)  # one GPU, chunk after chunk (a memory budget)
    feats = process_long_audio(module, audio, group=dist.group.WORLD)   # chunks dealt round-robin to the ranks

Results equal `module.process_audio(audio)` up to float32 rounding (the fast-convolution blocks of the ladder and the
sliding-DFT tiles start at the chunk's first sample instead of the track's): tests/test_longtrack.py holds them to the
bars of tests/test_gpu_parity.py.  Deviation: a harmonic of an HVQT / HCQT whose own VQT is a frame or two longer than the
common frame count (hvqt.py:123-128) has its maximum taken over the stored frames only.
"""

import json

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

ALIGN = 64      # chunk boundaries in frames: their sample positions are multiples of every ladder level's hop


def halo_frames(module, chunk_frames):
    """
    Frames next to a cut whose values depend on where the clip was cut.  STFT-like modules: the frames whose windows reach
    the cut.  CQT family: a frame of ladder level l under n_fft_l reaches n_fft_l / 2 * 2^l samples, and every 2:1 step adds
    its filter's half length, D * 2^(l - 1) samples; harmonics the reference early-downsamples in one call switch to their
    exact ladder near the clip edges (alt_th / alt_t0 of amtfeat_clip_describe) -- those frames are halo, too.
    """
    hop = int(module.hop_length)
    desc = module.describe()
    desc = json.loads(desc) if isinstance(desc, str) else desc
    items = desc.get('items') or []
    if not items:
        win = int(getattr(module, 'n_fft', 0) or getattr(module, 'win_length', 0) or hop)
        reach = win
    else:
        D = (int(desc['decim_taps']) - 1) // 2
        reach = max(it['n_fft'] // 2 * (1 << it['level']) + D * ((1 << it['level']) - 1) for it in items)
    halo = -(-reach // hop) + 1
    if items and desc.get('exact_ladders'):
        clip = module.describe_clip((chunk_frames + 4 * ALIGN) * hop)
        clip = json.loads(clip) if isinstance(clip, str) else clip
        T = int(clip['frames_computed'])
        halo = max(halo, max(clip['alt_th']), max(T - t0 for t0 in clip['alt_t0'] if t0 >= 0))
    return -(-halo // ALIGN) * ALIGN


def chunk_plan(num_samples, total_frames, hop, chunk_frames, halo):
    """[(f0, f1, a, b, k0)]: frames [f0, f1) of the track come from samples [a, b), where they are frames k0 ... of the chunk."""
    chunk_frames = max(ALIGN, chunk_frames // ALIGN * ALIGN)
    plan = []
    f0 = 0
    while f0 < total_frames:
        f1 = min(total_frames, f0 + chunk_frames)
        if total_frames - f1 < halo:        # a short remainder would be all halo: the last chunk takes it
            f1 = total_frames
        a = max(0, f0 - halo) * hop
        b = num_samples if f1 == total_frames else min(num_samples, (f1 + halo) * hop)
        plan.append((f0, f1, a, b, f0 - a // hop))
        f0 = f1
    return plan


class _CudaOps:
    """The three native steps of a chunk (amtfeat_process_raw / _range_reference / _range_finish)."""

    def __init__(self, module):
        self.m = module
        self.device = module.device
        self.C = int(module.get_num_channels())
        self.F = int(module.get_feature_size())

    def raw(self, chunks):
        """Raw (C, F, T_c) blocks of a list of chunks: ONE ragged batch, one set of launches."""
        m = self.m
        buf, offsets, lengths = m._pack(list(chunks))
        out, (shapes, sizes, out_offsets, _, _, _, _) = m._launch(buf, offsets, lengths, raw=True)
        return [out[o:o + sz].view(self.C, self.F, int(shape[-1])) for o, sz, shape in zip(out_offsets, sizes, shapes)]

    def reference(self, block, k0, k1, ref):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.amtfeat_range_reference(self.m._dev_plan.handle, block.data_ptr(), int(block.shape[-1]), int(k0), int(k1),
                                                        ref.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))

    def finish(self, block, k0, k1, ref, dst=None, t_dst=0):
        """Frames [k0, k1) of the raw block, finished against `ref`: into a new (C, F, k1 - k0) tensor, or into frames
        [t_dst, ...) of `dst` (the track's (C, F, T) block)."""
        if not 0 <= k0 <= k1 <= int(block.shape[-1]):
            raise ValueError('frame range [%d, %d) outside the block of %d frames' % (k0, k1, int(block.shape[-1])))
        out = dst if dst is not None else torch.empty((self.C, self.F, k1 - k0), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.amtfeat_range_finish(self.m._dev_plan.handle, block.data_ptr(), int(block.shape[-1]), int(k0), int(k1),
                                                     ref.data_ptr(), out.data_ptr(), int(out.shape[-1]), int(t_dst),
                                                     torch.cuda.current_stream(self.device).cuda_stream))
        return out


def process_long_audio(module, audio, chunk_frames=None, halo=None, group=None, gather=True, ops=None):
    """
    Features of ONE track, computed as chunks.  `audio`: 1-D float32 (numpy, or a host / device tensor); every rank of `group`
    passes the same track and uploads only the samples of its own chunks.  `gather=True` returns the whole (C, F, T) block on
    every rank (one broadcast per chunk); `gather=False` returns {chunk index: (f0, f1, tensor)} of the rank's own chunks.
    """
    if not hasattr(module, '_launch') or not hasattr(module, 'device'):
        raise TypeError('process_long_audio takes ONE feature module; call it per module of a FeatureCombo')
    if getattr(audio, 'ndim', 1) != 1:
        raise ValueError('expected mono-channel (1-D) audio, got shape %s' % (tuple(audio.shape),))
    ops = ops or _CudaOps(module)
    n = int(audio.shape[-1])
    hop = int(module.hop_length)
    shape = module._out_shape(n)
    T = int(shape[-1]) if len(shape) else 0
    distributed = group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if distributed else (0, 1)
    if chunk_frames is None and world == 1 and ops.__class__ is _CudaOps:
        # one rank and no chunk size asked for: the track is one call (chunking on one GPU only serves a memory budget)
        full = module.process_audio(audio)
        return full if gather else {0: (0, T, full.reshape(ops.C, ops.F, T) if isinstance(full, torch.Tensor) else full)}
    if chunk_frames is None:
        chunk_frames = max(16 * ALIGN, -(-T // (4 * world)))       # a few chunks per rank: the halo stays a small share
    if halo is None:
        halo = halo_frames(module, chunk_frames)
    chunk_frames = max(chunk_frames, 2 * halo)
    plan = chunk_plan(n, T, hop, chunk_frames, halo)
    ref = torch.full((ops.C,), float('-inf'), dtype=torch.float32, device=ops.device)
    mine = [ci for ci in range(len(plan)) if ci % world == rank]
    blocks = dict(zip(mine, ops.raw([audio[plan[ci][2]:plan[ci][3]] for ci in mine]))) if mine else {}
    for ci, block in blocks.items():
        f0, f1, a, b, k0 = plan[ci]
        if int(block.shape[-1]) < k0 + (f1 - f0):
            raise ValueError('chunk %d yields %d frames, %d needed' % (ci, int(block.shape[-1]), k0 + f1 - f0))
        ops.reference(block, k0, k0 + (f1 - f0), ref)
    if distributed:
        dist.all_reduce(ref, op=dist.ReduceOp.MAX, group=group)      # the one exchange step of the path: C floats
    if not gather:
        done = {}
        for ci, block in blocks.items():
            f0, f1, a, b, k0 = plan[ci]
            done[ci] = (f0, f1, ops.finish(block, k0, k0 + (f1 - f0), ref))
        return done
    # the whole block on every rank: own chunks are finished straight into their place, the others arrive by broadcast
    full = torch.empty((ops.C, ops.F, T), dtype=torch.float32, device=ops.device)
    for ci, (f0, f1, a, b, k0) in enumerate(plan):
        if ci in blocks and not distributed:
            ops.finish(blocks.pop(ci), k0, k0 + (f1 - f0), ref, dst=full, t_dst=f0)
            continue
        part = ops.finish(blocks.pop(ci), k0, k0 + (f1 - f0), ref) if ci in blocks else \
            torch.empty((ops.C, ops.F, f1 - f0), dtype=torch.float32, device=ops.device)
        dist.broadcast(part, src=dist.get_global_rank(group, ci % world), group=group)
        full[..., f0:f1] = part
    full = full.reshape(shape)
    return full.cpu().numpy() if getattr(module, 'output', 'torch') == 'numpy' else full
