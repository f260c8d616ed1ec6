"""Drop-in replacement for ``poreplex/signal_analyzer.py``.

Same public surface (``process_batch``, ``SignalAnalyzer``, ``SignalAnalysis``,
``SignalAnalysisError``), same result dicts (``NanoporeRead.report``,
signal_loader.py:165-198), same result ORDER (early-terminated reads first,
signal_analyzer.py:84-104,131-132), same status / label vocabulary -- but stages A-D
run as ONE batched GPU pass over all reads of the batch through the C ABI
(``SignalEngine.analyze_host`` -> ``pb2_analyze_host``) instead of per-read numpy /
TensorFlow / pomegranate calls.  ``pipeline.py:204`` can call this ``process_batch``
unchanged.

Switch coverage: ``trim_adapter`` (a no-op in the reference at this commit, SURVEY.md F6 --
reproduced), ``barcoding``, ``measure_polya``, ``filter_unsplit_reads``, and the two dump
switches ``dump_adapter_signals`` / ``dump_basecalls`` (signal_analyzer.py:155-211,450-466:
HDF5 part files under ``outputdir/adapter-dumps`` and ``outputdir/events``).  On-the-fly
albacore raises ``NotImplementedError`` when the analyzer is built, which ``process_batch``
reports as a batch-level failure exactly like any other unhandled exception -- never a
silent CPU fallback.
"""
import os
import sys
import traceback
from io import StringIO

import numpy as np

from . import _native as N
from .engine import get_engine, polya_to_dict
from .fast5_source import Fast5Source
from .params import STATUS_NAMES

__all__ = ['SignalAnalyzer', 'SignalAnalysis', 'process_batch']


class SignalAnalysisError(Exception):
    pass


# This function must be picklable.  (signal_analyzer.py:45-58)
def process_batch(batchid, reads, config):
    try:
        with SignalAnalyzer(config, batchid) as analyzer:
            return analyzer.process(reads)
    except Exception as exc:
        exc_type, exc_obj, exc_tb = sys.exc_info()
        filename = os.path.split(exc_tb.tb_frame.f_code.co_filename)[-1]
        errorf = StringIO()
        traceback.print_exc(file=errorf)
        return (-1, '[{filename}:{lineno}] Unhandled exception {name}: {msg}'.format(
                        filename=filename, lineno=exc_tb.tb_lineno,
                        name=type(exc).__name__, msg=str(exc)), errorf.getvalue())


class NanoporeRead:
    """Bookkeeping half of signal_loader.NanoporeRead (signal_loader.py:112-198); the
    numeric half (load_padded_signal_head / load_signal) lives on the GPU."""

    fast5 = error_message = None
    sequence_length = mean_qscore = num_events = 0
    sequence = scaling_params = label = barcode = polya = None
    barcode_bestguess = barcode_quality = None
    segments = None

    def __init__(self, filename, srcdir, read_id):
        self.fullpath = os.path.join(srcdir, filename)
        self.filename = filename
        self.read_id = read_id
        self.status = 'okay'
        self.stopped = False
        self.load()

    def set_status(self, newstatus, stop=False):
        self.status = newstatus
        self.stopped = self.stopped or stop

    def set_error(self, status, error_message):
        self.status = status
        self.error_message = error_message

    def set_scaling_params(self, params):
        self.scaling_params = params

    def set_label(self, newlabel):
        self.label = newlabel

    def set_barcode(self, newbarcode, guess, quality):
        self.barcode = newbarcode
        self.barcode_bestguess = guess
        self.barcode_quality = quality

    def set_polya_tail(self, polya_info):
        self.polya = polya_info

    def set_adapter_trimming_length(self, newlength):
        if self.sequence is None:
            raise Exception('Sequence is not set.')
        self.sequence = self.sequence[:2] + (newlength,)

    def is_stopped(self):
        return self.stopped

    def close(self):
        if self.fast5 is not None:
            self.fast5.close()

    def load(self):
        try:
            fast5 = Fast5Source(self.fullpath, self.read_id)
        except Exception:
            traceback.print_exc()
            self.set_status('irregular_fast5', stop=True)
            return
        self.fast5 = fast5
        self.sampling_rate = fast5.sampling_rate

    def head_is_too_short(self, length_limit, stride, min_length):
        """The length test of load_padded_signal_head (signal_loader.py:212-222)."""
        sigload_length = min(length_limit, self.fast5.duration)
        sigload_length = sigload_length - sigload_length % stride
        # get_raw_data(end=...) clips to the Signal dataset itself (fast5_file.py:124-125):
        # a dataset shorter than its `duration` attribute is judged by what was actually read
        got = min(sigload_length, self.fast5.signal_length())
        got -= got % stride
        return got < min_length

    def report(self):
        rep = {'filename': self.filename, 'read_id': self.read_id, 'status': self.status}
        if self.fast5 is not None:
            rep.update({
                'channel': self.fast5.channel_number,
                'start_time': round(self.fast5.start_time / self.fast5.sampling_rate, 3),
                'run_id': self.fast5.run_id,
                'sample_id': self.fast5.sample_id,
                'duration': self.fast5.duration,
                'num_events': self.num_events,
                'sequence_length': self.sequence_length,
                'mean_qscore': self.mean_qscore,
            })
        if self.sequence is not None:
            rep['sequence'] = self.sequence
        if self.error_message:
            rep['error_message'] = self.error_message
        if self.label is not None:
            rep['label'] = self.label
        if self.barcode is not None:
            rep['barcode'] = self.barcode
            rep['barcode_guess'] = self.barcode_bestguess
            rep['barcode_score'] = self.barcode_quality
        if self.polya is not None:
            rep['polya'] = self.polya
        return rep

    def load_fast5_events(self, want_events=False):
        if self.fast5 is None:
            raise Exception('Fast5 must be open for getting events.')
        bcall = self.fast5.get_basecall(want_events=want_events)
        if bcall is None:
            raise SignalAnalysisError('not_basecalled')
        self.sequence_length = bcall['sequence_length']
        self.mean_qscore = bcall['mean_qscore']
        self.num_events = bcall['num_events']
        self.sequence = bcall['sequence'], bcall['qstring'], 0
        return bcall['events']


class SignalAnalyzer:

    UNSUPPORTED_SWITCHES = {
        'albacore_onthefly': 'on-the-fly albacore basecalling',
    }

    # signal_analyzer.py:62-68 (model_state becomes 'S<kmersize>' in open_dumps, :156)
    _EVENT_DUMP_FIELD_NAMES = [
        'mean', 'start', 'stdv', 'length', 'model_state',
        'move', 'pos', 'end', 'scaled_mean']
    _EVENT_DUMP_FIELD_DTYPES = [
        '<f4', '<u8', '<f4', '<u8', None,
        '<i4', '<u8', '<u8', '<f8']

    def __init__(self, config, batchid):
        for key, what in self.UNSUPPORTED_SWITCHES.items():
            if config.get(key):
                raise NotImplementedError(
                    'poreplex_b200: {} is not built yet on the CUDA path and there is no '
                    'CPU fallback (config[{!r}])'.format(what, key))
        self.config = config
        self.inputdir = config['inputdir']
        self.outputdir = config.get('outputdir')
        self.batchid = batchid
        self.formatted_batchid = format(batchid, '08d')
        device = int(os.environ.get('POREPLEX_B200_DEVICE', config.get('cuda_device', 0)))
        self.engine = get_engine(config, device)
        self.kmersize = config.get('kmersize', 5)
        self.open_dumps()

    # ---- HDF5 dumps (signal_analyzer.py:155-211) ------------------------------------------
    # The reference appends to one file per worker process ('part-<workerid>.h5', h5py mode
    # 'a').  The writer here produces whole files, so every batch writes its own
    # 'part-<workerid>-<batchid>.h5' with the same internal layout; the inventories
    # (io.py:334-376) glob 'part-*.h5' and walk every batch group of every file, so they see
    # the same objects.
    def open_dumps(self):
        import multiprocessing as mp
        from hashlib import sha1
        self.workerid = sha1(mp.current_process().name.encode()).hexdigest()[:16]
        self.EVENT_DUMP_FIELDS = [
            (n, d if d is not None else 'S{}'.format(self.kmersize))
            for n, d in zip(self._EVENT_DUMP_FIELD_NAMES, self._EVENT_DUMP_FIELD_DTYPES)]
        self.adapter_dump = {} if self.config.get('dump_adapter_signals') else None
        self.adapter_dump_list = []
        self.basecall_dump = {} if self.config.get('dump_basecalls') else None

    def dump_path(self, subdir):
        path = os.path.join(self.outputdir, subdir,
                            'part-{}-{}.h5'.format(self.workerid, self.formatted_batchid))
        os.makedirs(os.path.dirname(path), exist_ok=True)
        return path

    def dump_adapter_signal(self, read_id, adapter_signal, segments, stride):
        """SignalAnalysis.dump_adapter_signal (signal_analyzer.py:450-466)."""
        if len(adapter_signal) > 0 and read_id not in self.adapter_dump:
            self.adapter_dump[read_id] = np.array(adapter_signal, dtype=np.float32)
            self.adapter_dump_list.append((read_id, segments['adapter'][0] * stride,
                                           (segments['adapter'][1] + 1) * stride))

    def write_basecalled_events(self, read_id, events, attrs):
        """signal_analyzer.py:184-196."""
        dataset = np.empty(len(events['start']), dtype=self.EVENT_DUMP_FIELDS)
        for name, _ in self.EVENT_DUMP_FIELDS:
            dataset[name] = events[name]
        if read_id not in self.basecall_dump:
            self.basecall_dump[read_id] = (dataset, attrs)

    def close(self):
        """signal_analyzer.py:198-211."""
        from . import hdf5_write as W
        if self.adapter_dump is not None:
            root = W.Group()
            grp = root.group('adapter').group(self.formatted_batchid)
            for read_id, sig in self.adapter_dump.items():
                grp.dataset(read_id, sig)
            catalog = np.array(self.adapter_dump_list,
                               dtype=[('read_id', 'S36'), ('start', 'i8'), ('end', 'i8')])
            root.group('catalog').group('adapter').dataset(self.formatted_batchid, catalog)
            W.write_file(self.dump_path('adapter-dumps'), root)
            self.adapter_dump = None
        if self.basecall_dump is not None:
            root = W.Group()
            grp = root.group('basecalled_events').group(self.formatted_batchid)
            for read_id, (dataset, attrs) in self.basecall_dump.items():
                grp.dataset(read_id, dataset, attrs=dict(attrs))
            W.write_file(self.dump_path('events'), root)
            self.basecall_dump = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def prefetch_signals(self, reads):
        """Raw signals of the whole batch through the native FAST5 loader (thread pool, one open
        per file) when the files are read without h5py; {index in reads: int16 array}.  Reads it
        cannot serve are simply absent and go through Fast5Source one by one."""
        from . import fast5_source
        if fast5_source._h5py() is not fast5_source._MinimalH5py or not reads:
            return {}
        try:
            from . import fast5_loader
            b = fast5_loader.load_batch(reads, inputdir=self.config['inputdir'],
                                        threads=int(self.config.get('ingest_threads', 0)) or None)
        except Exception:
            return {}
        return {i: b['raw'][b['offsets'][i]:b['offsets'][i] + b['lengths'][i]]
                for i in range(len(reads)) if b['status'][i] == fast5_loader.READ_OK}

    def process(self, reads):
        inputdir = self.config['inputdir']
        eng = self.engine
        results, loaded = [], []

        # STAGE A: open reads, early exits (signal_analyzer.py:84-104)
        nextprocs = []
        prefetched = self.prefetch_signals(reads)
        for ridx, (f5file, read_id) in enumerate(reads):
            if not os.path.exists(os.path.join(inputdir, f5file)):
                results.append({'filename': f5file, 'status': 'disappeared'})
                continue
            try:
                npread = NanoporeRead(f5file, inputdir, read_id)
                if npread.fast5 is None:
                    # the reference dereferences the unopened file here
                    raise AttributeError("'NoneType' object has no attribute 'duration'")
                if npread.head_is_too_short(eng.scaler_length, eng.stride,
                                            eng.scaler_min_length):
                    npread.set_status('scaler_signal_too_short', stop=True)
                if npread.is_stopped():
                    results.append(npread.report())
                else:
                    npread._raw = prefetched.get(ridx)
                    if npread._raw is None:
                        npread._raw = npread.fast5.raw_int16()
                    nextprocs.append(SignalAnalysis(npread, self))
                    loaded.append(npread)
            except Exception as exc:
                results.append(self.pack_unhandled_exception(f5file, read_id, exc,
                                                             sys.exc_info()))

        # STAGES B-D on the GPU: one batched pass for every loaded read
        if loaded:
            raw, offsets, lengths = eng.pack_reads([r._raw for r in loaded])
            batch = (raw, offsets, lengths,
                     np.array([r.fast5.range for r in loaded], np.float64),
                     np.array([r.fast5.digitization for r in loaded], np.float64),
                     np.array([r.fast5.offset for r in loaded], np.float64))
            out = eng.analyze_host(
                *batch,
                barcoding=bool(self.config['barcoding']),
                polya=bool(self.config['measure_polya']),
                # the adapter dump stores the scaled pooled signal itself
                keep_pooled=self.adapter_dump is not None,
                # the chimera filter and the event dump use the scaled event means: exact scale/shift
                exact_scaler=bool(self.config.get('filter_unsplit_reads')) or
                self.basecall_dump is not None)
            pooled_at = eng.pooled_offsets(offsets)
            for i, npread in enumerate(loaded):
                npread._raw = None
                if self.adapter_dump is not None:
                    T = int(lengths[i]) // eng.stride
                    npread._pooled = out['pooled'][pooled_at[i]:pooled_at[i] + T]
                npread._gpu = {k: out[k][i] for k in ('status', 'scale_shift', 'segments',
                                                      'barcode', 'barcode_guess',
                                                      'barcode_score')}
                if 'polya' in out:
                    npread._gpu['polya'] = out['polya'][i]
                st = STATUS_NAMES[int(out['status'][i])]
                if st in ('scaling_qc_fail', 'scaler_signal_too_short'):
                    # fit_scalers (signal_loader.py:104-109); a too-short head is normally
                    # caught in stage A -- if the device still reports one, it stops the read
                    # here rather than letting it continue with no scaling parameters
                    npread.set_status(st, stop=True)
                else:
                    npread.set_scaling_params(np.array(out['scale_shift'][i], dtype=np.float32))

        # STAGE C: per-read bookkeeping in input order (signal_analyzer.py:111-124); the
        # chimera filter runs as one batched GPU call between its two halves
        for phase in (1, 2):
            for siganal in nextprocs:
                try:
                    if not siganal.is_stopped() and not siganal.failed:
                        siganal.process(phase)
                except Exception as exc:
                    f5file = siganal.npread.filename
                    read_id = siganal.npread.read_id
                    error = self.pack_unhandled_exception(f5file, read_id, exc, sys.exc_info())
                    siganal.set_error(error)
                    siganal.failed = True
            if phase == 1 and loaded:
                if self.config['filter_unsplit_reads'] or self.basecall_dump is not None:
                    self.derive_event_tables(nextprocs, batch)
                if self.config['filter_unsplit_reads']:
                    self.detect_unsplit_reads(nextprocs, batch)
        for siganal in nextprocs:
            siganal.clear_cache()

        # STAGE E
        for npread in loaded:
            results.append(npread.report())
        return results

    def derive_event_tables(self, analyses, batch):
        """Event tables of every guppy ``Move`` basecall of the batch in one device call
        (Fast5Reader.construct_events_from_moves / convert_events_guppy, fast5_file.py:183-230,
        and the derived columns of SignalAnalysis.load_events, signal_analyzer.py:311-326);
        `analyses` is aligned with the packed `batch`."""
        eng = self.engine
        idx = [i for i, a in enumerate(analyses)
               if not (a.is_stopped() or a.failed) and a.events is not None
               and a.events.get('guppy_move')]
        if not idx:
            return
        raw, offsets, lengths, rng, dig, off = batch
        sub = (raw, offsets[idx], lengths[idx], rng[idx], dig[idx], off[idx])
        ev = [analyses[i].events for i in idx]
        strides = {int(e['block_stride']) for e in ev}
        if len(strides) != 1:
            raise ValueError('reads of one batch must share block_stride')
        cols = ('mean', 'start', 'p_model_state')
        if self.basecall_dump is not None:
            cols = ('mean', 'stdv', 'start', 'end', 'length', 'pos', 'p_model_state',
                    'model_state', 'scaled_mean')
        tables, err = eng.derive_event_tables_host(
            sub, [e['move'] for e in ev], [int(e['first_sample']) for e in ev], strides.pop(),
            sequences=[e['sequence'] for e in ev], qstrings=[e['qstring'] for e in ev],
            scale_shift=np.array([analyses[i].npread._gpu['scale_shift'] for i in idx], np.float32),
            columns=cols)
        for k, i in enumerate(idx):
            a = analyses[i]
            if err[k]:
                # both conditions are checked when the table is loaded (Fast5Source); the device
                # disagreeing is an internal error of this read, never silent
                a.set_error(self.pack_unhandled_exception(
                    a.npread.filename, a.npread.read_id,
                    Exception('event-table derivation failed on the device (code {})'.format(int(err[k]))),
                    (None, None, None)))
                a.failed = True
                continue
            a.events.update(tables[k])

    def detect_unsplit_reads(self, analyses, batch):
        """Batched SignalAnalysis.detect_unsplit_read (signal_analyzer.py:366-443); `analyses`
        is aligned with the packed `batch` (one entry per loaded read)."""
        eng = self.engine
        live = [not (a.is_stopped() or a.failed) and a.events is not None for a in analyses]
        if not any(live):
            return
        flags = eng.detect_unsplit_host(
            [a.events if ok else None for a, ok in zip(analyses, live)],
            np.array([a.npread.sampling_rate for a in analyses], np.float64),
            np.array([a.npread._gpu['scale_shift'] for a in analyses], np.float32),
            np.array([0 if ok else 10 for ok in live], np.int32),
            np.array([a.npread._gpu['segments'] for a in analyses], np.int32),
            batch=batch)
        for a, f in zip(analyses, flags):
            a.unsplit_flag = int(f)

    def pack_unhandled_exception(self, f5filename, read_id, exc, excinfo):
        exc_type, exc_obj, exc_tb = excinfo
        if exc_tb is None:                                  # not raised: report this frame
            srcfilename, lineno = os.path.basename(__file__), sys._getframe().f_lineno
        else:
            srcfilename = os.path.split(exc_tb.tb_frame.f_code.co_filename)[-1]
            lineno = exc_tb.tb_lineno
        errorf = StringIO()
        traceback.print_exc(file=errorf)
        errmsg = ('[{srcfilename}:{lineno}] ({f5filename}#{read_id}) Unhandled '
                  'exception {name}: {msg}\n{exc}'.format(
            srcfilename=srcfilename, lineno=lineno,
            f5filename=f5filename, read_id=read_id, name=type(exc).__name__, msg=str(exc),
            exc=errorf.getvalue()))
        return {
            'filename': f5filename,
            'read_id': read_id,
            'status': 'unknown_error',
            'error_message': errmsg,
        }


class SignalAnalysis:

    def __init__(self, npread, analyzer):
        self.npread = npread
        self.config = analyzer.config
        self.analyzer = analyzer
        self.failed = False          # an unhandled exception ended this read
        self.events = None
        self.unsplit_flag = 0

    def set_error(self, error):
        self.npread.set_error(error['status'], error['error_message'])

    def is_stopped(self):
        return self.npread.is_stopped()

    def clear_cache(self):
        self.npread.close()

    def process(self, phase=None):
        """Consume the GPU results of this read in the reference's order of checks
        (signal_analyzer.py:230-286).  Phase 1 = everything up to load_events, phase 2 =
        trim / chimera verdict / minimum length; ``phase=None`` runs both."""
        npread = self.npread
        gpu = npread._gpu
        eng = self.analyzer.engine
        try:
            if phase in (None, 1):
                status = STATUS_NAMES[int(gpu['status'])]
                if status == 'unknown_error':
                    raise Exception('Viterbi decoding found no path for this read.')
                segments = self.detect_segments()
                npread.segments = segments
                if 'adapter' not in segments:
                    raise SignalAnalysisError('adapter_not_detected')

                if self.analyzer.adapter_dump is not None:     # signal_analyzer.py:243-244
                    a0, a1 = segments['adapter']
                    self.analyzer.dump_adapter_signal(npread.read_id, npread._pooled[a0:a1 + 1],
                                                      segments, eng.stride)

                if self.config['barcoding'] and int(gpu['barcode_score']) >= 0:
                    bc = int(gpu['barcode'])
                    npread.set_barcode(None if bc < 0 else bc, int(gpu['barcode_guess']),
                                       int(gpu['barcode_score']))

                if self.config['measure_polya']:         # signal_analyzer.py:250-256
                    info = polya_to_dict(gpu['polya'], npread.sampling_rate)
                    if info is not None:
                        npread.set_polya_tail(info)

                self.events = self.load_events()

            if phase in (None, 2):
                if self.analyzer.basecall_dump is not None:    # signal_analyzer.py:259-263
                    self.analyzer.write_basecalled_events(
                        npread.read_id, self.dump_columns(self.events),
                        self.get_dump_attributes(npread.segments, eng.stride))

                if self.config['trim_adapter']:
                    self.trim_adapter(self.events, npread.segments, eng.stride)

                if self.config['filter_unsplit_reads']:
                    if self.unsplit_flag < 0:
                        raise Exception('unsplit-read detection failed on the device '
                                        '(code {})'.format(self.unsplit_flag))
                    if self.unsplit_flag:
                        raise SignalAnalysisError('unsplit_read')

                if npread.sequence is not None:
                    readlength = len(npread.sequence[0]) - npread.sequence[2]
                    if readlength < self.config['minimum_sequence_length']:
                        raise SignalAnalysisError('sequence_too_short')
                npread.set_label('pass')

        except SignalAnalysisError as exc:
            outname = 'artifact' if exc.args[0] in ('unsplit_read',) else 'fail'
            npread.set_status(exc.args[0], stop=True)
            npread.set_label(outname)

    def dump_columns(self, events):
        """The event table with the derived columns of load_events (signal_analyzer.py:311-326).
        Guppy Move tables come complete from the device; tables stored in the file (albacore,
        old guppy) get the three derived columns here, as numpy computes them in the
        reference."""
        if 'scaled_mean' in events and 'pos' in events and 'end' in events:
            return events
        ev = dict(events)
        if 'mean' not in ev:
            raise Exception('event dump of a table without a mean column is not supported')
        ev['scaled_mean'] = np.poly1d(self.npread.scaling_params)(np.asarray(ev['mean']))
        ev['pos'] = np.cumsum(ev['move'])
        duration = np.hstack((np.diff(ev['start']), [1])).astype(np.int64)
        ev['end'] = ev['start'] + duration
        return ev

    def get_dump_attributes(self, segments, stride):
        """signal_analyzer.py:288-309."""
        attrlist = []
        if self.npread.scaling_params is not None:
            sp_scale, sp_shift = self.npread.scaling_params
            attrlist.append(('signal_scale', sp_scale))
            attrlist.append(('signal_shift', sp_shift))
        if 'adapter' in segments:
            attrlist.append(('adapter_begin', np.uint32(segments['adapter'][0] * stride)))
            attrlist.append(('adapter_end', np.uint32((segments['adapter'][1] + 1) * stride)))
        if self.npread.polya is not None:
            polya = self.npread.polya
            if 'polya-tail' in segments:
                attrlist.append(('polya_end_debug',
                                 np.uint32((segments['polya-tail'][1] + 1) * stride)))
            attrlist.append(('polya_begin', np.uint32(polya['begin'])))
            attrlist.append(('polya_end', np.uint32(polya['end'])))
            attrlist.append(('spikes', repr(polya['spikes']).encode()))
        return attrlist

    def detect_segments(self):
        """{state name: (first, last)} from the kernel's baked-order table
        (signal_analyzer.py:355-362)."""
        seg = self.npread._gpu['segments']
        names = self.analyzer.engine.state_names
        return {names[s]: (int(seg[s, 0]), int(seg[s, 1]))
                for s in range(len(names)) if seg[s, 0] >= 0}

    def load_events(self):
        events = self.npread.load_fast5_events(
            want_events=bool(self.config['filter_unsplit_reads']) or
            self.analyzer.basecall_dump is not None)
        if self.npread.scaling_params is None:
            raise Exception('Signal scaling is not available yet.')
        return events

    def trim_adapter(self, events, segments, elspan):
        # signal_analyzer.py:328-331: the guard returns whenever the sequence IS set,
        # and load_events() has always just set it -> no-op, kept as is.
        sequence = self.npread.sequence
        if sequence is not None:
            return
