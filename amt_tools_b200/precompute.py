"""
Bulk feature precompute + cache writer -- the caller of the hot path named by BASELINE config 5
(SURVEY.md 8f rank 1).  The reference computes one track at a time inside `TranscriptionDataset.__getitem__`
and `np.savez_compressed`s synchronously (datasets/common.py:212-295); here a whole corpus shard is pushed
through the GPU in ragged batches while a thread pool compresses and writes the previous batch.

The on-disk format is the reference's, so existing caches stay loadable and new ones are readable by the
unmodified `calculate_feats` (datasets/common.py:242-250):
    <save_loc>/<Dataset>/<features_name()>/<track>.npz   with keys  fs, hop_length, features
(datasets/common.py:259-265, 482-504; tools/constants.py:48-50).
"""

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import shard

KEY_FS, KEY_HOP, KEY_FEATS = 'fs', 'hop_length', 'features'     # amt_tools/tools/constants.py:48-50


def feats_path(save_loc, dataset_name, data_proc, track):
    """datasets/common.py:482-504 get_feats_dir(track)."""
    return os.path.join(save_loc, dataset_name, data_proc.features_name(), '%s.npz' % track)


def _write(path, fs, hop_length, feats, compressed):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    (np.savez_compressed if compressed else np.savez)(path, **{KEY_FS: fs, KEY_HOP: hop_length, KEY_FEATS: feats})
    return path


def precompute_features(tracks, data_proc, save_loc, dataset_name, rank=0, world_size=1, max_batch_seconds=960.0,
                        overwrite=False, compressed=True, writers=4, keep_on_device=False):
    """
    Compute and cache the features of every track this rank owns.

    tracks : dict  track name -> 1-D float32 audio (np.ndarray or torch.Tensor), already at data_proc's sample rate
    Returns {track: path} (and, with keep_on_device=True, {track: (path, CUDA tensor)} for device-resident consumers).
    Tracks whose cache file exists are skipped unless overwrite=True (datasets/common.py:242: cache hit -> load).
    """
    names = sorted(tracks)
    lengths = [int(tracks[n].shape[-1]) for n in names]
    mine = shard.shard_tracks(lengths, world_size, rank)
    todo = [i for i in mine if overwrite or not os.path.exists(feats_path(save_loc, dataset_name, data_proc, names[i]))]
    budget = int(max_batch_seconds * data_proc.get_sample_rate())
    batches = shard.make_batches(todo, lengths, budget)
    fs, hop = data_proc.get_sample_rate(), data_proc.get_hop_length()
    out, pending = {}, []
    copy_stream = torch.cuda.Stream(data_proc.device)
    with ThreadPoolExecutor(max_workers=max(1, writers)) as pool:
        for batch in batches:
            feats = data_proc.process_audio([tracks[names[i]] for i in batch])        # ragged batch on the compute stream
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(data_proc.device))
            with torch.cuda.stream(copy_stream):                                      # D2H overlaps the next batch's kernels
                copy_stream.wait_event(done)
                host = []
                for f in feats:
                    f.record_stream(copy_stream)
                    h = torch.empty(f.shape, dtype=f.dtype, pin_memory=True)
                    h.copy_(f, non_blocking=True)
                    host.append(h)
                copied = torch.cuda.Event()
                copied.record(copy_stream)
            pending.append((batch, feats if keep_on_device else None, host, copied))
            while len(pending) > 1:                                                    # write batch i-1 while batch i computes
                _flush(pending.pop(0), names, fs, hop, save_loc, dataset_name, data_proc, pool, compressed, out)
        while pending:
            _flush(pending.pop(0), names, fs, hop, save_loc, dataset_name, data_proc, pool, compressed, out)
        for k, v in list(out.items()):
            if isinstance(v, tuple):
                out[k] = (v[0].result(), v[1])
            else:
                out[k] = v.result()
    return out


def _flush(item, names, fs, hop, save_loc, dataset_name, data_proc, pool, compressed, out):
    batch, dev, host, copied = item
    copied.synchronize()
    for j, i in enumerate(batch):
        path = feats_path(save_loc, dataset_name, data_proc, names[i])
        fut = pool.submit(_write, path, fs, hop, host[j].numpy(), compressed)
        out[names[i]] = (fut, dev[j]) if dev is not None else fut


def load_features(path):
    """tools/utils.py:3485-3502 load_dict_npz + the unpacking of datasets/common.py:244-250."""
    d = dict(np.load(path, allow_pickle=True))
    return d[KEY_FEATS], d[KEY_FS].item(), d[KEY_HOP].item()
