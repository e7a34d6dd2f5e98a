"""
Bulk feature precompute + cache writer -- the caller of the hot path named by BASELINE config 5
(SURVEY.md 8f rank 1).  The reference computes one track at a time inside `TranscriptionDataset.__getitem__`
and `np.savez_compressed`s synchronously (datasets/common.py:212-295); here a whole corpus shard is pushed
through the C-ABI pipelined executor (amtfeat_pipeline_*) in ragged batches over reused pinned host buffers while a
thread pool compresses and writes the previous batch.

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


COMPRESS_LEVEL = 1          # deflate level of the compressed cache files
_CHUNK = 8 << 20            # bytes of an array one deflate task takes
_zpool = None


def _deflate_pool():
    global _zpool
    if _zpool is None:
        _zpool = ThreadPoolExecutor(max_workers=max(2, os.cpu_count() or 2), thread_name_prefix='amtfeat-deflate')
    return _zpool


def _deflate_chunk(args):
    import zlib
    chunk, level, last = args
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    return co.compress(chunk) + co.flush(zlib.Z_FINISH if last else zlib.Z_FULL_FLUSH)


def _savez_deflate(f, arrays, level):
    """`np.savez_compressed` rewritten for throughput: the same zip-of-.npy container (any `np.load` reads it, as the reference's
    loader does, datasets/common.py:245), but
      * deflate level 1 instead of zlib's default 6 (float32 features: files ~3 % larger, 1.7x - 3x faster), and
      * every array is deflated in 8 MB pieces on a thread pool (zlib releases the GIL) -- each piece ends with a full flush, so the
        concatenation is one valid deflate stream -- instead of one thread per file: a 178 MB HCQT track no longer waits for a
        single core.
    The zip structures (local headers, central directory) are written by hand because zipfile cannot take pre-deflated data;
    members of 4 GiB or more fall back to zipfile (zip64)."""
    import io
    import struct
    import zipfile
    import zlib
    members = []
    for key, val in arrays.items():
        arr = np.asanyarray(val)
        if arr.ndim and not arr.flags.c_contiguous:       # (np.ascontiguousarray would turn the 0-d fs / hop_length into 1-d arrays)
            arr = np.ascontiguousarray(arr)
        hdr = io.BytesIO()
        np.lib.format.write_array_header_1_0(hdr, np.lib.format.header_data_from_array_1_0(arr))
        members.append((key + '.npy', hdr.getvalue(), arr))
    if any(len(h) + a.nbytes >= (1 << 32) - 1 for _, h, a in members):
        with zipfile.ZipFile(f, 'w', zipfile.ZIP_DEFLATED, allowZip64=True, compresslevel=level) as zf:
            for name, h, a in members:
                with zf.open(name, 'w', force_zip64=True) as fp:
                    fp.write(h)
                    fp.write(a.data)
        return
    central = []
    offset = 0
    for name, h, a in members:
        raw = memoryview(np.atleast_1d(a).reshape(-1).view(np.uint8)) if a.nbytes else memoryview(b'')
        pieces = [bytes(h) + bytes(raw[:max(0, _CHUNK - len(h))])]
        pos = max(0, _CHUNK - len(h))
        while pos < len(raw):
            pieces.append(raw[pos:pos + _CHUNK])
            pos += _CHUNK
        tasks = [(pc, level, i == len(pieces) - 1) for i, pc in enumerate(pieces)]
        comp = list(_deflate_pool().map(_deflate_chunk, tasks)) if len(tasks) > 1 else [_deflate_chunk(tasks[0])]
        crc = 0
        for pc in pieces:
            crc = zlib.crc32(pc, crc)
        csize, usize = sum(len(c) for c in comp), len(h) + a.nbytes
        nm = name.encode()
        f.write(struct.pack('<IHHHHHIIIHH', 0x04034b50, 20, 0, 8, 0, 0x21, crc, csize, usize, len(nm), 0) + nm)
        for c in comp:
            f.write(c)
        central.append(struct.pack('<IHHHHHHIIIHHHHHII', 0x02014b50, 20, 20, 0, 8, 0, 0x21, crc, csize, usize, len(nm), 0, 0, 0, 0, 0, offset) + nm)
        offset += 30 + len(nm) + csize
    cd = b''.join(central)
    f.write(cd)
    f.write(struct.pack('<IHHHHIIH', 0x06054b50, 0, 0, len(central), len(central), len(cd), offset, 0))


def _write(path, fs, hop_length, feats, compressed):
    """One cache file in the reference's format.  Written next to its final name and renamed into place: the loader (like the
    reference's, datasets/common.py:242) takes the existence of the file as a cache hit, so a killed run must not leave a
    truncated one behind."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    tmp = '%s.tmp.%d' % (path, os.getpid())
    try:
        arrays = {KEY_FS: fs, KEY_HOP: hop_length, KEY_FEATS: feats}
        with open(tmp, 'wb') as f:
            if compressed:
                _savez_deflate(f, arrays, COMPRESS_LEVEL)
            else:
                np.savez(f, **arrays)
        os.replace(tmp, path)
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)
    return path


def precompute_features(tracks, data_proc, save_loc, dataset_name, rank=0, world_size=1, max_batch_seconds=960.0,
                        overwrite=False, compressed=True, writers=4, keep_on_device=False, ring=3, stats=None):
    """
    Compute and cache the features of every track this rank owns.

    tracks : dict  track name -> 1-D float32 audio (np.ndarray or CPU torch.Tensor), already at data_proc's sample rate
    Returns {track: path} (and, with keep_on_device=True, {track: (path, CUDA tensor)} for device-resident consumers).
    Tracks whose cache file exists are skipped unless overwrite=True (datasets/common.py:242: cache hit -> load).

    The batches run on the C-ABI pipelined executor (amtfeat_pipeline_*: upload / compute / download streams over `ring` device
    staging slots) with `ring` reused pinned host buffers on either side; `writers` threads compress / write batch i - 1 while
    batch i is on the GPU.  `stats` (dict) receives the seconds spent waiting for the GPU and for the writers.
    """
    if not hasattr(data_proc, '_dev_plan') or not hasattr(data_proc, 'device'):
        raise TypeError('precompute_features needs a single feature module of amt_tools_b200.features (one plan, one device); '
                        'run the modules of a FeatureCombo one by one')
    names = sorted(tracks)
    lengths = [int(tracks[n].shape[-1]) for n in names]
    mine = shard.shard_tracks(lengths, world_size, rank)
    todo = [i for i in mine if overwrite or not os.path.exists(feats_path(save_loc, dataset_name, data_proc, names[i]))]
    budget = int(max_batch_seconds * data_proc.get_sample_rate())
    batches = shard.make_batches(todo, lengths, budget)
    if keep_on_device:
        return _precompute_resident(batches, tracks, names, data_proc, save_loc, dataset_name, compressed, writers)
    if not batches:
        return {}
    import time
    from .pipeline import Pipeline
    fs, hop = data_proc.get_sample_rate(), data_proc.get_hop_length()
    # layout of every batch: clip offsets inside the (4-element aligned) packed audio, output offsets, workspace
    plans = []
    for batch in batches:
        lens = [lengths[i] for i in batch]
        in_off, total_in = [], 0
        for n in lens:
            in_off.append(total_in)
            total_in += (n + 3) // 4 * 4
        shapes, sizes, out_off, _, _, ws_bytes, total_out = data_proc._batch_layout(lens)
        plans.append((batch, lens, in_off, max(total_in, 4), shapes, sizes, out_off, max(total_out, 4), ws_bytes))
    max_in, max_out = max(p[3] for p in plans), max(p[7] for p in plans)
    ring = max(1, min(max(2, int(ring)), len(plans)))        # no more staging slots (pinned and device memory) than batches
    t_setup = time.perf_counter()
    pipe = Pipeline(data_proc.device.index, ring, max_in, max_out, max(p[8] for p in plans))
    t_setup = time.perf_counter() - t_setup
    t_alloc = t_fill = 0.0
    # pinned staging buffers, allocated when a slot is first used: page-locking gigabytes takes about as long as writing them out,
    # so the later slots are locked while the GPU and the writers are already busy with the first batch
    h_in, h_out = [None] * ring, [None] * ring

    def alloc_slot(k):
        return (torch.empty(max_in, dtype=torch.float32, pin_memory=True), torch.empty(max_out, dtype=torch.float32, pin_memory=True))

    alloc_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix='amtfeat-pin')
    alloc_futs = [alloc_pool.submit(alloc_slot, k) for k in range(ring)]     # slot 0 first; the others lock pages behind the first batch
    busy = [[] for _ in range(ring)]          # writer futures still reading a host output slot
    out, pending = {}, []
    t_gpu = t_writers = 0.0

    def flush(item, pool):
        nonlocal t_gpu
        slot, ticket, plan = item
        t = time.perf_counter()
        pipe.wait(ticket)
        t_gpu += time.perf_counter() - t
        batch, _, _, _, shapes, sizes, out_off, _, _ = plan
        for j, i in enumerate(batch):
            feats = h_out[slot][out_off[j]:out_off[j] + sizes[j]].view(shapes[j]).numpy()
            fut = pool.submit(_write, feats_path(save_loc, dataset_name, data_proc, names[i]), fs, hop, feats, compressed)
            busy[slot].append(fut)
            out[names[i]] = fut

    try:
        with ThreadPoolExecutor(max_workers=max(1, writers)) as pool:
            for b, plan in enumerate(plans):
                slot = b % ring
                while len(pending) >= max(1, ring - 1):                          # keep ring - 1 batches in flight
                    flush(pending.pop(0), pool)
                t = time.perf_counter()
                for fut in busy[slot]:                                           # the writers of batch b - ring are done with this slot
                    fut.result()
                t_writers += time.perf_counter() - t
                busy[slot] = []
                batch, lens, in_off, total_in, _, _, out_off, total_out, _ = plan
                if h_in[slot] is None:
                    t = time.perf_counter()
                    h_in[slot], h_out[slot] = alloc_futs[slot].result()
                    t_alloc += time.perf_counter() - t
                t = time.perf_counter()
                for i, o, n in zip(batch, in_off, lens):
                    a = tracks[names[i]]
                    a = a.detach().cpu() if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
                    h_in[slot][o:o + n].copy_(a.to(torch.float32))
                t_fill += time.perf_counter() - t
                ticket = pipe.submit(data_proc, h_in[slot], in_off, lens, out_off, h_out[slot], total_in, total_out)
                pending.append((slot, ticket, plan))
            while pending:
                flush(pending.pop(0), pool)
            t = time.perf_counter()
            for k in list(out):
                out[k] = out[k].result()
            t_writers += time.perf_counter() - t
    finally:
        alloc_pool.shutdown(wait=True)
        pipe.wait(-1)
        pipe.close()
    if stats is not None:
        stats.update(wait_gpu_s=t_gpu, wait_writers_s=t_writers, pipeline_create_s=t_setup, pinned_alloc_s=t_alloc, fill_s=t_fill,
                     batches=len(plans), tracks=len(out))
    return out


def _precompute_resident(batches, tracks, names, data_proc, save_loc, dataset_name, compressed, writers):
    """keep_on_device=True: the features also stay on the GPU for a consumer there, so every batch gets its own device
    tensors (torch allocations on the current stream) instead of the pipeline's reused staging slots."""
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
            pending.append((batch, feats, host, copied))
            while len(pending) > 1:                                                    # write batch i-1 while batch i computes
                _flush(pending.pop(0), names, fs, hop, save_loc, dataset_name, data_proc, pool, compressed, out)
        while pending:
            _flush(pending.pop(0), names, fs, hop, save_loc, dataset_name, data_proc, pool, compressed, out)
        for k, v in list(out.items()):
            out[k] = (v[0].result(), v[1])
    return out


def _flush(item, names, fs, hop, save_loc, dataset_name, data_proc, pool, compressed, out):
    batch, dev, host, copied = item
    copied.synchronize()
    for j, i in enumerate(batch):
        path = feats_path(save_loc, dataset_name, data_proc, names[i])
        fut = pool.submit(_write, path, fs, hop, host[j].numpy(), compressed)
        out[names[i]] = (fut, dev[j])


def load_features(path):
    """tools/utils.py:3485-3502 load_dict_npz + the unpacking of datasets/common.py:244-250."""
    d = dict(np.load(path, allow_pickle=True))
    return d[KEY_FEATS], d[KEY_FS].item(), d[KEY_HOP].item()
