"""
Python handle of the C-ABI pipelined host executor (include/amtfeat.h amtfeat_pipeline_*): upload, compute and download streams
chained with events over `nslots` sets of device staging buffers.  Host buffers are pinned torch tensors owned by the caller.
"""

import ctypes as C

from . import _lib


class Pipeline(object):
    def __init__(self, device_index, nslots, max_audio_elems, max_out_elems, max_workspace_bytes):
        self.handle = C.c_void_p()
        _lib.check(_lib.lib.amtfeat_pipeline_create(int(device_index), int(nslots), int(max_audio_elems), int(max_out_elems),
                                                    int(max_workspace_bytes), C.byref(self.handle)))
        self.nslots = int(nslots)

    def submit(self, module, h_audio, in_offsets, lengths, out_offsets, h_out, audio_elems=None, out_elems=None):
        """Enqueue H2D -> process_audio -> D2H of one ragged batch (pinned float32 tensors h_audio / h_out); returns a ticket."""
        ticket = C.c_int64()
        _lib.check(_lib.lib.amtfeat_pipeline_submit(
            self.handle, module._dev_plan.handle, h_audio.data_ptr(), _lib.i64_array(in_offsets), _lib.i64_array(lengths),
            _lib.i64_array(out_offsets), len(lengths), h_out.data_ptr(),
            int(h_audio.numel() if audio_elems is None else audio_elems), int(h_out.numel() if out_elems is None else out_elems),
            C.byref(ticket)))
        return ticket.value

    def wait(self, ticket=-1):
        _lib.check(_lib.lib.amtfeat_pipeline_wait(self.handle, int(ticket)))

    def close(self):
        lib = getattr(_lib, 'lib', None)
        if lib is not None and self.handle is not None and self.handle.value:
            lib.amtfeat_pipeline_destroy(self.handle)
            self.handle = C.c_void_p()

    __del__ = close
