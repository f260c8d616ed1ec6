"""Drop-in replacement for ``poreplex/signal_analyzer.py``.

Same public surface (``process_batch``, ``SignalAnalyzer``, ``SignalAnalysis``,
``SignalAnalysisError``), same result dicts (``NanoporeRead.report``,
signal_loader.py:165-198), same result ORDER (early-terminated reads first,
signal_analyzer.py:84-104,131-132), same status / label vocabulary -- but stages A-D
run as ONE batched GPU pass over all reads of the batch through the C ABI
(``SignalEngine.analyze_host`` -> ``pb2_analyze_host``) instead of per-read numpy /
TensorFlow / pomegranate calls.  ``pipeline.py:204`` can call this ``process_batch``
unchanged.

Switch coverage this round: ``trim_adapter`` (a no-op in the reference at this commit,
SURVEY.md F6 -- reproduced), ``barcoding``, ``measure_polya`` and
``filter_unsplit_reads``.  The dump switches and on-the-fly albacore raise ``NotImplementedError`` when the
analyzer is built, which ``process_batch`` reports as a batch-level failure exactly like
any other unhandled exception -- never a silent CPU fallback.
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
        'dump_adapter_signals': 'adapter signal dumps',
        'dump_basecalls': 'basecalled event dumps',
        'albacore_onthefly': 'on-the-fly albacore basecalling',
    }

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

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        pass

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
                # the chimera filter decodes the scaled event means: exact scale/shift
                exact_scaler=bool(self.config.get('filter_unsplit_reads')))
            for i, npread in enumerate(loaded):
                npread._raw = None
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
            if phase == 1 and self.config['filter_unsplit_reads'] and loaded:
                self.detect_unsplit_reads(nextprocs, batch)
        for siganal in nextprocs:
            siganal.clear_cache()

        # STAGE E
        for npread in loaded:
            results.append(npread.report())
        return results

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
        srcfilename = os.path.split(exc_tb.tb_frame.f_code.co_filename)[-1]
        errorf = StringIO()
        traceback.print_exc(file=errorf)
        errmsg = ('[{srcfilename}:{lineno}] ({f5filename}#{read_id}) Unhandled '
                  'exception {name}: {msg}\n{exc}'.format(
            srcfilename=srcfilename, lineno=exc_tb.tb_lineno,
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

    def detect_segments(self):
        """{state name: (first, last)} from the kernel's baked-order table
        (signal_analyzer.py:355-362)."""
        seg = self.npread._gpu['segments']
        names = self.analyzer.engine.state_names
        return {names[s]: (int(seg[s, 0]), int(seg[s, 1]))
                for s in range(len(names)) if seg[s, 0] >= 0}

    def load_events(self):
        events = self.npread.load_fast5_events(
            want_events=bool(self.config['filter_unsplit_reads']))
        if self.npread.scaling_params is None:
            raise Exception('Signal scaling is not available yet.')
        return events

    def trim_adapter(self, events, segments, elspan):
        # signal_analyzer.py:328-331: the guard returns whenever the sequence IS set,
        # and load_events() has always just set it -> no-op, kept as is.
        sequence = self.npread.sequence
        if sequence is not None:
            return
