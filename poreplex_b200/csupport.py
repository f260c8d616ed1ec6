"""GPU twin of the reference's one native module, ``poreplex.csupport`` (src/csupport.c).

``detect_events(signal, window_length1=30, window_length2=120, threshold1=3.0, threshold2=9.0,
peak_height=8.0)`` keeps the reference's signature, defaults, return dtype and error behaviour
(csupport.c:70-124, 156-159): the input is coerced to a C-contiguous float32 array, anything
but a 1-D array is a ``ValueError``, a signal that yields no event raises ``csupport.error``.
The detector itself (scrappie's ``event_detection.c``) runs on the GPU through
``pb2_detect_events``; :func:`detect_events_batch` is the form worth calling -- many signals,
one upload, two kernel launches (count, fill), one download.

No CPU fallback: without the CUDA library or a device these functions raise.
"""
import ctypes as C

import numpy as np

from . import _native as N

__all__ = ['detect_events', 'detect_events_batch', 'error', 'EVENT_DTYPE']

# csupport.c:156-159 (SIZEOF_INT == 4)
EVENT_DTYPE = np.dtype([('start', 'u8'), ('length', 'f4'), ('mean', 'f4'), ('stdv', 'f4'),
                        ('pos', 'i4'), ('state', 'i4')])
assert EVENT_DTYPE.itemsize == 28


class error(Exception):
    """``csupport.error`` (csupport.c:189-195)."""


_contexts = {}


def _context(device):
    ctx = _contexts.get(device)
    if ctx is None:
        lib = N.load()
        handle = C.c_void_p()
        rc = lib.pb2_create(int(device), C.byref(handle))
        if rc != 0 or not handle.value:
            raise N.NativeError('poreplex_b200.csupport: cannot open CUDA device %d (rc=%d); '
                                'there is no CPU fallback' % (device, rc))
        ctx = _contexts[device] = (lib, handle)
    return ctx


def _as_signal(signal):
    a = np.ascontiguousarray(np.asarray(signal).astype(np.float32, copy=False))
    if a.ndim != 1:
        raise ValueError('Expects an 1-dimensional array.')
    return a


def detect_events_batch(signals, window_length1=30, window_length2=120, threshold1=3.0,
                        threshold2=9.0, peak_height=8.0, device=0):
    """``[detect_events(s, ...) for s in signals]`` in one pass over the GPU.  A signal without
    events (an empty one) gives an empty table here instead of raising."""
    import torch
    sigs = [_as_signal(s) for s in signals]
    n = len(sigs)
    if n == 0:
        return []
    lib, handle = _context(device)
    lengths = np.array([s.size for s in sigs], np.int64)
    offsets = np.zeros(n, np.int64)
    np.cumsum(lengths[:-1], out=offsets[1:])
    flat = np.concatenate(sigs) if lengths.sum() else np.zeros(1, np.float32)
    dev = torch.device('cuda', int(device))
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        d_sig = torch.from_numpy(flat).to(dev)
        d_off = torch.from_numpy(offsets).to(dev)
        d_len = torch.from_numpy(lengths).to(dev)
        d_cnt = torch.zeros(n, dtype=torch.int64, device=dev)
        par = N.DetectorParams(int(window_length1), int(window_length2), float(threshold1),
                               float(threshold2), float(peak_height))

        def call(counts, ev_off, records):
            rc = lib.pb2_detect_events(handle, d_sig.data_ptr(), d_off.data_ptr(), d_len.data_ptr(),
                                       n, C.byref(par), counts, ev_off, records, st)
            if rc != 0:
                msg = lib.pb2_last_error(handle)
                raise N.NativeError('pb2_detect_events failed (rc=%d): %s'
                                    % (rc, msg.decode() if msg else ''))

        call(d_cnt.data_ptr(), None, None)
        d_evoff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(d_cnt, 0, out=d_evoff[1:])
        ev_off = d_evoff.cpu().numpy()
        total = int(ev_off[-1])
        d_rec = torch.zeros(max(total, 1) * EVENT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        if total:
            call(None, d_evoff.data_ptr(), d_rec.data_ptr())
        table = d_rec.cpu().numpy()[:total * EVENT_DTYPE.itemsize].view(EVENT_DTYPE)
    return [table[ev_off[i]:ev_off[i + 1]].copy() for i in range(n)]


def detect_events(signal, window_length1=30, window_length2=120, threshold1=3.0, threshold2=9.0,
                  peak_height=8.0):
    """csupport.detect_events (csupport.c:70-124)."""
    ev = detect_events_batch([_as_signal(signal)], window_length1, window_length2, threshold1,
                             threshold2, peak_height)[0]
    if len(ev) == 0:
        raise error('Event detection failed.')
    return ev
