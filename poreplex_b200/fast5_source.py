"""FAST5 access for the drop-in ``process_batch``.

Host-side I/O only: mirrors what ``poreplex/fast5_file.py`` (Fast5Reader) hands to the
signal path -- per-read metadata, the int16 ``Signal`` dataset *untouched* (the
int16 -> pA conversion of ``get_raw_data`` runs on the GPU) and the basecall summary
behind ``NanoporeRead.load_fast5_events`` (signal_loader.py:266-279).  Uses whatever
``h5py`` module is importable (the parity tests install an in-memory one); without
h5py it reads the files through ``poreplex_b200.hdf5_min`` (SURVEY.md section 8f rank 2;
``poreplex_b200.fast5_loader`` is the native batch form for the raw signals).
"""
import numpy as np

__all__ = ['Fast5Source']


class _SharedFile:
    """One open hdf5_min file shared by every Fast5Source of the same path: the drop-in keeps a
    handle per loaded read until the batch ends (as the reference does), and a batch is 10^4-10^5
    reads of a few multi-read files -- one mapping per file, not one per read."""

    _open = {}                                    # abspath -> [Hdf5File, reference count]

    def __init__(self, path):
        import os
        from .hdf5_min import Hdf5File
        self._file = None
        self._key = os.path.abspath(path)
        slot = self._open.get(self._key)
        if slot is None:
            slot = self._open[self._key] = [Hdf5File(path), 0]
        slot[1] += 1
        self._file = slot[0]

    def close(self):
        if self._file is None:
            return
        slot = self._open.get(self._key)
        self._file = None
        if slot is not None:
            slot[1] -= 1
            if slot[1] <= 0:
                del self._open[self._key]
                slot[0].close()

    def __del__(self):                            # sources dropped without close() (reads stopped
        try:                                      # before the GPU stage) release their share too
            self.close()
        except Exception:
            pass

    def __getitem__(self, path):
        return self._file[path]

    def __contains__(self, path):
        return path in self._file

    def __iter__(self):
        return iter(self._file)

    def keys(self):
        return self._file.keys()

    @property
    def attrs(self):
        return self._file.attrs


class _MinimalH5py:
    """``h5py.File`` stand-in over poreplex_b200.hdf5_min (read-only; the classic HDF5 format with
    gzip / shuffle / VBZ chunked datasets that FAST5 files use)."""

    @staticmethod
    def File(path, mode='r'):
        if mode != 'r':
            raise ValueError('hdf5_min is read-only')
        return _SharedFile(path)


def _h5py():
    """h5py when it is importable (the parity tests install an in-memory one), else the built-in
    minimal reader: I/O only, the signal path itself has no CPU fallback."""
    try:
        import h5py
    except ImportError:
        return _MinimalH5py
    return h5py


class Fast5Source:
    """Same node layout rules as Fast5Reader.__init__ / load_metadata
    (fast5_file.py:65-120)."""

    RAWSIGNAL_PREFILTER_SIZE = 5          # fast5_file.py:63

    def __init__(self, path, read_id):
        self.path = path
        self.read_id = read_id
        self.handle = _h5py().File(path, 'r')
        try:
            self.is_multiread = 'UniqueGlobalKey' not in self.handle
            if self.is_multiread:
                self.read_node = 'read_{}/Raw'.format(read_id)
                self.channel_node = 'read_{}/channel_id'.format(read_id)
                self.tracking_node = 'read_{}/tracking_id'.format(read_id)
                self.analyses_node = 'read_{}/Analyses'.format(read_id)
            else:
                first_read_name = next(iter(self.handle['Raw/Reads'].keys()))
                self.read_node = 'Raw/Reads/' + first_read_name
                self.channel_node = 'UniqueGlobalKey/channel_id'
                self.tracking_node = 'UniqueGlobalKey/tracking_id'
                self.analyses_node = 'Analyses'
            self._load_metadata()
        except BaseException:
            self.close()                          # an unreadable read must not pin the file open
            raise

    def close(self):
        if self.handle is not None:
            self.handle.close()
            self.handle = None

    def _load_metadata(self):
        sigattrs = self.handle[self.read_node].attrs
        self.duration = int(sigattrs['duration'])
        self.start_time = int(sigattrs['start_time'])
        file_read_id = sigattrs['read_id'].decode()
        if self.read_id is None:
            self.read_id = file_read_id
        elif file_read_id != self.read_id:
            raise ValueError('Unexpected read {} found in {}'.format(file_read_id, self.path))
        chanattrs = self.handle[self.channel_node].attrs
        self.channel_number = chanattrs['channel_number'].decode()
        self.digitization = float(chanattrs['digitisation'])
        self.offset = float(chanattrs['offset'])
        self.range = float(chanattrs['range'])
        self.sampling_rate = float(chanattrs['sampling_rate'])
        trackattrs = self.handle[self.tracking_node].attrs
        self.run_id = trackattrs['run_id'].decode()
        self.sample_id = trackattrs['sample_id'].decode()

    def signal_length(self):
        """len() of the Signal dataset (what get_raw_data clips to, fast5_file.py:123-125)."""
        return len(self.handle[self.read_node + '/Signal'])

    def raw_int16(self):
        """The whole Signal dataset as int16 (no conversion; fast5_file.py:123-128)."""
        node = self.handle[self.read_node + '/Signal']
        return np.ascontiguousarray(node[0:len(node)], dtype=np.int16)

    # -- basecall summary: Fast5Reader.get_basecall (fast5_file.py:133-164) ----------
    def get_basecall(self, analysis_group='Basecall_1D', want_events=False):
        try:
            analnode = self.handle[self.analyses_node]
        except KeyError:
            return None
        analgroups = [name for name in analnode.keys() if name.startswith(analysis_group)]
        if len(analgroups) < 1:
            return None
        analyses = analnode[max(analgroups)]
        groupno = analyses.name.rsplit('_', 1)[-1]
        segattrs = analnode['Segmentation_{}/Summary/segmentation'.format(groupno)].attrs
        summary = {}
        fastqenc = analyses['BaseCalled_template/Fastq'][()].decode().split('\n')
        summary['sequence'] = fastqenc[1]
        summary['qstring'] = fastqenc[3]
        summaryattrs = analyses['Summary/{}_template'.format(analysis_group.lower())].attrs
        summary['block_stride'] = int(summaryattrs.get('block_stride', 15))
        summary['sequence_length'] = int(summaryattrs['sequence_length'])
        summary['mean_qscore'] = float(summaryattrs['mean_qscore'])
        summary['num_events'] = int(segattrs['num_events_template'])
        summary['first_sample_template'] = int(segattrs['first_sample_template'])
        # the reference always builds the event table here; its failure modes
        # (missing tables, unknown k-mer size, length mismatch) must surface even when
        # no consumer needs the columns
        summary['events'] = self._load_events(analyses, summary, want_events)
        return summary

    def _load_events(self, analyses, summary, want_events):
        if 'BaseCalled_template/Events' in analyses:
            table = analyses['BaseCalled_template/Events'][()]
            names = table.dtype.names or ()
            if len(names) <= 3 and 'move' in names:            # old guppy
                cols = {'move': np.asarray(table['move'])}
                for extra in ('model_state', 'p_model_state'):
                    if extra in names:
                        cols[extra] = np.asarray(table[extra])
                return self._convert_guppy(cols, summary, want_events)
            if len(names) == 14:                               # albacore >= 2.3.0
                return {n: np.asarray(table[n]) for n in names}
            raise Exception('Unsupported event table found.')
        if 'BaseCalled_template/Move' in analyses:
            moves = analyses['BaseCalled_template/Move'][()]
            kmer_size = len(summary['sequence']) - int(moves.sum()) + 1
            if kmer_size not in (5, 1):
                raise Exception('Move table is encoded with an unknown kmer-size.')
            cols = {'move': np.asarray(moves)}
            if want_events:
                # construct_events_from_moves (fast5_file.py:183-207): pos, p_model_state and
                # model_state are derived on the GPU for the whole batch
                # (SignalEngine.derive_event_tables_host); what they are derived from:
                cols['guppy_move'] = True
                cols['sequence'] = summary['sequence']
                cols['qstring'] = summary['qstring']
            return self._convert_guppy(cols, summary, want_events)
        raise Exception("Neither `Events' or `Move' table found in the basecall.")

    def _convert_guppy(self, cols, summary, want_events):
        """convert_events_guppy (fast5_file.py:209-230)."""
        n = len(cols['move'])
        first_sample = summary['first_sample_template']
        block_stride = summary['block_stride']
        last_sample = first_sample + block_stride * n
        if not cols.get('guppy_move'):
            cols['start'] = np.arange(first_sample, last_sample, block_stride)
        node = self.handle[self.read_node + '/Signal']
        end = min(last_sample, len(node))
        nraw = max(end - first_sample, 0)
        padded = nraw + ((block_stride - nraw % block_stride) % block_stride)
        if padded // block_stride != n:
            raise Exception('Numbers of events and raw data strides does not match.')
        if want_events:
            # the `mean` / `stdv` columns (medfilt(5) + per-block statistics of the pA signal,
            # fast5_file.py:217-227) are derived on the GPU from the raw signal
            # (k_event_stats); only their coordinates are recorded here
            cols['first_sample'] = first_sample
            cols['block_stride'] = block_stride
        if not cols.get('guppy_move'):
            cols['length'] = np.full(n, block_stride)
        return cols
