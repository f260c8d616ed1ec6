"""Summary writers fed by the GPU path.

Mirrors of the two consumers of the result dicts that close the loop
``process_batch -> pipeline.py -> sequencing_summary.txt / final table``:

* :class:`SequencingSummaryWriter` -- reference ``poreplex/io.py:120-184``: same constructor,
  ``write_results(list[dict])`` and ``close()``, byte-identical ``sequencing_summary.txt``; plus
  ``write_batch`` which takes the per-read metadata once and the engine's result ARRAYS
  (status / label / barcode / barcode_score / poly(A) records as ``SignalEngine.analyze_host``
  returns them) without building dicts.
* :class:`FinalSummaryTracker` -- reference ``poreplex/io.py:236-332``: ``feed_results`` (dicts)
  or ``feed_counts`` (the ``int64[4][5][11]`` histogram ``k_counts`` produces on the device and
  the 8 GPUs all-reduce), ``print_results(file)`` with the reference's layout.  Fed by dicts the
  table is byte-identical to the reference's; fed by the histogram it differs only in how rows
  with EQUAL counts are ordered (the reference breaks such ties by the order in which the keys
  were first seen, which a histogram does not carry; here: status code order).

Nothing in here touches the GPU; it is host-side formatting of what the kernels counted.
"""
import logging
import os
from collections import OrderedDict
from threading import Lock

from .params import STATUS_NAMES

LABEL_NAMES = ('pass', 'fail', 'artifact')          # k_finalize label codes 0, 1, 2 (3 = none)


def output_name_mapping(config):
    """(label_names, barcode_names, output_layout) as commandline.py:137-159 builds them:
    the keys every writer indexes its streams with -- 'artifact' exists only with the
    chimera filter on, barcode keys only with barcoding on."""
    label_names = {'fail': 'fail', 'pass': 'pass'}
    if config['filter_unsplit_reads']:
        label_names['artifact'] = 'artifact'
    if config['barcoding']:
        barcode_names = {None: 'undetermined'}
        for i in range(config['demultiplexing']['number_of_barcodes']):
            barcode_names[i] = 'BC{n}'.format(n=i + 1)
        layout = {(label, bc): os.path.join(labelname, bcname)
                  for label, labelname in label_names.items()
                  for bc, bcname in barcode_names.items()}
    else:
        barcode_names = {None: '-'}
        layout = {(label, None): labelname for label, labelname in label_names.items()}
    return label_names, barcode_names, layout


class FASTQWriter:
    """io.py:38-71: one gzip stream per (label, barcode) output; a read is written with its
    adapter_length trailing bases removed.  (The reference writes BGZF through pysam, a
    gzip-compatible container: the decompressed bytes are what is compared.)"""

    def __init__(self, output_dir, output_layout):
        import gzip
        self.output_dir = output_dir
        self.output_layout = output_layout
        self.lock = Lock()
        self.streams = {}
        for int_name, name in output_layout.items():
            path = os.path.join(output_dir, 'fastq', name + '.fastq.gz')
            os.makedirs(os.path.dirname(path), exist_ok=True)
            self.streams[int_name] = gzip.open(path, 'wb')

    def close(self):
        for stream in self.streams.values():
            stream.close()

    def write_sequences(self, procresult):
        with self.lock:
            for entry in procresult:
                if entry.get('sequence') is not None:
                    seq, qual, adapter_length = entry['sequence']
                    if adapter_length > 0:
                        seq = seq[:-adapter_length]
                        qual = qual[:-adapter_length]
                    output_name = entry['label'], entry.get('barcode')
                    formatted = '@{}\n{}\n+\n{}\n'.format(entry['read_id'], seq, qual)
                    self.streams[output_name].write(formatted.encode('ascii'))


class SequencingSummaryWriter:
    """io.py:120-184."""

    SUMMARY_OUTPUT_FIELDS = [
        'filename', 'read_id', 'run_id', 'channel', 'start_time',
        'duration', 'num_events', 'sequence_length', 'mean_qscore',
        'sample_id', 'status', 'label',
    ]

    def __init__(self, config, output_dir, label_mapping, barcode_mapping):
        self.file = open(os.path.join(output_dir, 'sequencing_summary.txt'), 'w')
        self.lock = Lock()
        self.label_mapping = label_mapping
        self.output_fields = list(self.SUMMARY_OUTPUT_FIELDS)
        self.barcode_mapping = barcode_mapping if config['barcoding'] else None
        if self.barcode_mapping is not None:
            self.output_fields += ['barcode', 'barcode_score']
        self.polya_enabled = bool(config['measure_polya'])
        if self.polya_enabled:
            self.output_fields.append('polya_dwell')
        self.fast5_output = bool(config['fast5_output'])
        print(*self.output_fields, sep='\t', file=self.file)

    def close(self):
        self.file.close()

    def _filename(self, label_name, barcode, filename):
        if not self.fast5_output:
            return filename
        if self.barcode_mapping is not None:
            return os.path.join('fast5', label_name, self.barcode_mapping[barcode], filename)
        return os.path.join('fast5', label_name, filename)

    def _row(self, entry):
        row = dict(entry)
        row['label'] = self.label_mapping[entry['label']]
        row['filename'] = self._filename(row['label'], entry.get('barcode'), entry['filename'])
        if self.barcode_mapping is not None:
            row['barcode'] = self.barcode_mapping[entry.get('barcode')]
            row['barcode_score'] = entry.get('barcode_score', 0)
        if self.polya_enabled:
            row['polya_dwell'] = (format(entry['polya']['dwell_time'], '.4f')
                                  if 'polya' in entry else '')
        return [row[f] for f in self.output_fields]

    def write_results(self, results):
        with self.lock:
            for entry in results:
                if 'label' in entry:                       # reads stopped before stage C: no row
                    print(*self._row(entry), file=self.file, sep='\t')

    def write_batch(self, meta, out, polya_dwell=None):
        """``meta``: per read a dict with the FAST5-side fields (filename ... sample_id);
        ``out``: the engine's result arrays.  A barcode is reported only when assigned
        (signal_loader.py:190), a label of 3 (none) means the read produced no row."""
        with self.lock:
            for i, m in enumerate(meta):
                lab = int(out['label'][i])
                if lab > 2:
                    continue
                entry = dict(m)
                entry['status'] = STATUS_NAMES[int(out['status'][i])]
                entry['label'] = LABEL_NAMES[lab]
                bc = int(out['barcode'][i])
                if bc >= 0:
                    entry['barcode'] = bc
                    entry['barcode_score'] = int(out['barcode_score'][i])
                if polya_dwell is not None and polya_dwell[i] is not None:
                    entry['polya'] = {'dwell_time': polya_dwell[i]}
                print(*self._row(entry), file=self.file, sep='\t')


class FinalSummaryTracker:
    """io.py:236-332."""

    REPORTING_ORDER = ['pass', 'artifact', 'fail']
    FRIENDLY_LABELS = {
        'pass': 'Successfully processed',
        'fail': 'Processing failed',
        'artifact': 'Possible artifact',
    }
    FRIENDLY_STATUS = {
        'fail': {
            'scaler_signal_too_short': 'Signal is too short',
            'sequence_too_short': 'Sequence is too short',
            'irregular_fast5': 'Invalid FAST5 format',
            'basecall_table_incomplete': 'Basecall table does not match',
            'adapter_not_detected': "3' Adapter could not be located",
            'not_basecalled': 'No albacore basecall data found',
            'scaling_qc_fail': 'Signal scaling QC failed',
            'disappeared': 'File is moved to other location',
            'unknown_error': 'File could not be opened due to unknown error',
        },
        'artifact': {
            'unsplit_read': 'Two or more molecules found within a read',
        },
    }
    LABEL_FORMAT = '{:49s} '
    LABEL_BULLET = ' - '
    MINIMUM_COLUMN_WIDTH = 3

    def __init__(self, label_names, barcode_names):
        self.label_names = label_names
        self.barcode_names = barcode_names
        self.counts = OrderedDict()
        self.barcode_reporting_order = sorted(n for n in barcode_names if n is not None) + [None]

    def _add(self, key, n):
        self.counts[key] = self.counts.get(key, 0) + n

    def feed_results(self, results):
        for entry in results:
            self._add((entry.get('label', 'fail'), entry.get('barcode', None), entry['status']), 1)

    def feed_counts(self, counts):
        """``counts[label 0..3][barcode slot 0..4][status 0..10]`` as pb2_count_results /
        ``SignalEngine.count_results`` return it (label 3 = no label -> 'fail', io.py:276;
        barcode slot 0 = undetermined)."""
        for li in range(4):
            label = LABEL_NAMES[li] if li < 3 else 'fail'
            for slot in range(5):
                for si in range(len(STATUS_NAMES)):
                    n = int(counts[li][slot][si])
                    if n:
                        self._add((label, None if slot == 0 else slot - 1, STATUS_NAMES[si]), n)

    def print_results(self, file):
        if hasattr(file, 'write'):
            def emit(*args):
                print(*args, sep='\t', file=file)
        else:
            logger = logging.getLogger('poreplex')

            def emit(*args):
                logger.error(' '.join(map(str, args)))

        emit('==== Result Summary ====')
        width = max(self.MINIMUM_COLUMN_WIDTH, len(format(max(self.counts.values()), 'd')))
        title_fmt = '{{:{}s}} '.format(width)
        number_fmt = '{{:{}d}} '.format(width)
        if len(self.barcode_names) > 1:
            emit(self.LABEL_FORMAT.format('') +
                 ''.join(title_fmt.format(self.barcode_names[bc]) for bc in self.barcode_reporting_order))

        # rows in label order, larger counts first (stable: ties keep first-seen order), then
        # grouped by (label, status) in order of first appearance
        rows = [(k[0], k[1], k[2], v) for k, v in self.counts.items()]
        rows.sort(key=lambda r: (self.REPORTING_ORDER.index(r[0]), -r[3]))
        groups = OrderedDict()
        for label, barcode, status, n in rows:
            groups.setdefault((label, status), {})[barcode] = n

        current = None
        for (label, status), by_barcode in groups.items():
            title = None
            if current != label:
                current = label
                if label in self.FRIENDLY_STATUS:
                    emit(self.LABEL_FORMAT.format(self.FRIENDLY_LABELS[label]))
                else:
                    title = self.FRIENDLY_LABELS[label]
            if title is None:
                title = self.LABEL_BULLET + self.FRIENDLY_STATUS[label][status]
            emit(self.LABEL_FORMAT.format(title) +
                 ''.join(number_fmt.format(by_barcode.get(bc, 0)) for bc in self.barcode_reporting_order))
        emit('')
