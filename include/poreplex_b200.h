/*
 * poreplex_b200.h -- C ABI of the B200-native Poreplex signal path.
 *
 * One shared library (poreplex_b200/libporeplex_b200.so, sm_100a) replaces the
 * numeric work behind the reference's per-batch entry point
 *
 *     poreplex/signal_analyzer.py:46   process_batch(batchid, reads, config)
 *     poreplex/signal_analyzer.py:82   SignalAnalyzer.process(reads)
 *
 * i.e. the calls that the reference makes into numpy / TensorFlow / pomegranate /
 * its csupport extension for every read.  Each entry point below names the
 * reference interface it stands in for.  Plain pointers and sizes only; no Python,
 * torch or CUDA types appear in the signatures (a stream is passed as void*).
 *
 * Conventions
 *   - every function returns 0 on success, a negative PB2_E* code otherwise;
 *     pb2_last_error() gives a message.  Per-read problems are NOT errors: they are
 *     reported as per-read status codes (PB2_ST_*), exactly as the reference turns
 *     them into result-dict statuses (io.py:245-260).
 *   - "dev" pointers are device pointers on the context's GPU; "host" pointers are
 *     host memory (pinned memory makes the copies asynchronous).
 *   - HMM states are always indexed in pomegranate's baked order (sorted by name).
 */
#ifndef POREPLEX_B200_H
#define POREPLEX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB2_ABI_VERSION 1

#define PB2_MAX_STATES 8
#define PB2_MAX_COMP 4
#define PB2_MAX_EDGES 64
#define PB2_MAX_CLASSES 8
#define PB2_MAX_CALIB 64
#define PB2_WINDOW_MAX 512

/* error codes */
#define PB2_OK 0
#define PB2_EINVAL (-1)
#define PB2_ECUDA (-2)
#define PB2_ENOMEM (-3)
#define PB2_ESTATE (-4)      /* parameters not set yet */
#define PB2_EUNSUPPORTED (-5)

/* per-read status codes: poreplex/io.py:245-260 */
enum {
    PB2_ST_OKAY = 0, PB2_ST_DISAPPEARED = 1, PB2_ST_IRREGULAR_FAST5 = 2,
    PB2_ST_SCALER_SIGNAL_TOO_SHORT = 3, PB2_ST_SCALING_QC_FAIL = 4,
    PB2_ST_ADAPTER_NOT_DETECTED = 5, PB2_ST_NOT_BASECALLED = 6,
    PB2_ST_BASECALL_TABLE_INCOMPLETE = 7, PB2_ST_UNSPLIT_READ = 8,
    PB2_ST_SEQUENCE_TOO_SHORT = 9, PB2_ST_UNKNOWN_ERROR = 10, PB2_N_STATUS = 11
};

/* labels (signal_analyzer.py:281-286); PB2_LABEL_NONE = read stopped before stage C */
enum { PB2_LABEL_PASS = 0, PB2_LABEL_FAIL = 1, PB2_LABEL_ARTIFACT = 2, PB2_LABEL_NONE = 3,
       PB2_N_LABEL = 4 };
#define PB2_N_BARCODE_SLOTS 5   /* undetermined + 4 barcodes (io.py:269-278) */

/* switches: commandline.py:279-291 */
#define PB2_FLAG_BARCODING 1u        /* --barcoding      config['barcoding']            */
#define PB2_FLAG_KEEP_POOLED 2u      /* also return the scaled pooled signal            */
#define PB2_FLAG_POLYA 4u            /* --polya          config['measure_polya']        */
#define PB2_FLAG_EXACT_SCALER 8u     /* scale/shift from the exact f32 scaler kernels (the caller
                                        feeds them to the chimera filter or needs them bit-exact) */

typedef struct pb2_context pb2_context;

/* One Keras LSTM / LSTMCell (gate blocks i|f|c|o).  Host pointers, copied. */
typedef struct {
    int32_t in_dim, units;
    int32_t implementation;          /* Keras `implementation`: 1 or 2 (bias order) */
    const float *kernel;             /* [in_dim][4*units] */
    const float *recurrent;          /* [units][4*units]  */
    const float *bias;               /* [4*units]         */
} pb2_lstm_weights;

/* SignalLoader.load_scaler_model (signal_loader.py:49-75): network + input_defs +
 * output_transform + QC bounds (norm.ppf of scaler_qc_threshold, computed by caller). */
typedef struct {
    pb2_lstm_weights l1, l2;
    const float *dense_kernel;       /* [units][2] */
    const float *dense_bias;         /* [2] */
    int32_t stride;                  /* rough_signal_stride (15)   */
    int32_t length;                  /* input_defs.length (30000)  */
    int32_t min_length;              /* input_defs.min_length (9000) */
    double scale_std, scale_mean, shift_std, shift_mean;
    double qc_scale_lo, qc_scale_hi, qc_shift_lo, qc_shift_hi;
} pb2_scaler_params;

/* load_segmentation_model + bake() (worker_persistence.py:95-121), log space */
typedef struct {
    int32_t n_states;
    int32_t n_comp[PB2_MAX_STATES];
    double mu[PB2_MAX_STATES][PB2_MAX_COMP];
    double log_norm[PB2_MAX_STATES][PB2_MAX_COMP];     /* -log(sigma*SQRT_2_PI) */
    double inv_two_var[PB2_MAX_STATES][PB2_MAX_COMP];  /* 1/(2 sigma^2)        */
    double log_weight[PB2_MAX_STATES][PB2_MAX_COMP];
    double log_start[PB2_MAX_STATES];                  /* -inf: no start edge   */
    int32_t in_begin[PB2_MAX_STATES + 1];              /* CSR by destination    */
    int32_t in_src[PB2_MAX_EDGES];
    double in_logp[PB2_MAX_EDGES];
} pb2_hmm_params;

/* BarcodeDemultiplexer.__init__/load_model (barcoding.py:34-70) + config['demultiplexing'] */
typedef struct {
    pb2_lstm_weights fwd, bwd, l2;
    const float *dense_kernel;       /* [l2.units][n_classes] */
    const float *dense_bias;         /* [n_classes] */
    int32_t n_classes;
    int32_t n_decoy;                 /* number_of_decoy_labels */
    int32_t min_length, max_length;  /* minimum/maximum_dna_length (pooled samples) */
    int32_t trim_length;             /* signal_trim_length (300) */
    float pad_value;                 /* PAD_FILLER (-1000) */
    int32_t n_calibration;
    const double *calibration;       /* pred_score[phred] */
    double score_threshold;          /* calibration[barcoding_quality_filter] */
} pb2_demux_params;

/* PolyASignalAnalyzer.__init__ (polya.py:39-48) over config['polya_dwell'].  Values
 * that the reference compares against float32 Series are given as float (NEP 50). */
typedef struct {
    int32_t stride;                  /* rough_signal_stride                              */
    int32_t refinement_expansion;
    int32_t openend_unit;            /* openend_expansion // stride                      */
    int32_t max_extension;           /* maximum_openend_extension                        */
    int32_t window_length1, window_length2;   /* event_detection                         */
    float threshold1, threshold2, peak_height;
    float cutoff_lo, cutoff_hi;      /* f32(mean -/+ sd * z_cutoff)                      */
    float mean_loc;                  /* f32(polya_mean_dist[0])                          */
    float trigger;                   /* f32(polya_mean_trigger_recalibration * sd)       */
    float half_range;                /* f32(sd * z_cutoff)                               */
    float stdv_max;
    double stdv_lo, stdv_hi;         /* polya_stdv_range                                 */
    int32_t spike_tolerance;
    double spike_weight;
    int32_t recal_max_dist;          /* recalibrate_shifted_signal.*                     */
    float recal_min_length;
    float recal_max_stdv;
} pb2_polya_params;

#define PB2_POLYA_MAX_SPIKES 48
/* NanoporeRead.set_polya_tail payload (polya.py:116-121) */
typedef struct {
    int32_t found;                   /* 0: no poly(A) reported for this read             */
    int32_t n_spikes;                /* may exceed PB2_POLYA_MAX_SPIKES (list truncated) */
    int64_t begin, end;              /* raw-sample coordinates                           */
    int64_t dwell_samples;           /* dwell_time = dwell_samples / sampling_rate       */
    int32_t extensions;              /* open-end extensions used                         */
    int32_t flags;
    float spikes[PB2_POLYA_MAX_SPIKES][4];   /* length, mean[k-1], mean[k], mean[k+1]; NaN = absent */
} pb2_polya_result;

/* config['unsplit_read_detection'] (rna-r941.cfg:17-27); durations in seconds -- the
 * kernel applies int(value * sampling_rate) per read (signal_analyzer.py:375-380). */
typedef struct {
    double window_size, window_step, strict_duration;
    double strict_full_length, strict_dna_length, loosen_full_length, loosen_dna_length;
    double basecount_quality_limit, subread_basecount_limit, subread_baseratio_limit;
} pb2_unsplit_params;

/* Basecalled event tables of a batch (NanoporeRead.load_fast5_events): ragged, read i owns
 * rows [event_offsets[i], event_offsets[i+1]); reads without a table have an empty range.
 * The derived columns of load_events (scaled_mean, pos, end) are computed on the device. */
typedef struct {
    int64_t n_events_total;
    const int64_t *event_offsets;    /* [n_reads + 1]                  */
    const int64_t *start;            /* [total] raw-sample index       */
    const float *mean;               /* [total] event mean, pA; NULL = derive on the device
                                        from the raw signal as convert_events_guppy does
                                        (fast5_file.py:209-230), see first_sample     */
    const int32_t *move;             /* [total]                        */
    const double *p_model_state;     /* [total]                        */
    const double *sampling_rate;     /* [n_reads]                      */
    const int64_t *first_sample;     /* [n_reads] first_sample_template (mean == NULL) */
    int32_t block_stride;            /* samples per event row          (mean == NULL) */
} pb2_event_tables;

/* FASTQ side of a batch of guppy basecalls (Fast5Reader.get_basecall, fast5_file.py:149-151):
 * what construct_events_from_moves (fast5_file.py:183-207) reads besides the Move table. */
typedef struct {
    const uint8_t *sequence;         /* concatenated sequences as basecalled (U or T)          */
    const uint8_t *qstring;          /* concatenated quality strings, same offsets             */
    const int64_t *seq_offsets;      /* [n_reads + 1]                                          */
    const double *qual_table;        /* [256] 1 - 10 ** -((q - 33) / 10) per byte value, built
                                        by the caller with the reference's numpy expression    */
} pb2_basecalls;

/* Event-table columns derived on the device (any pointer may be NULL = not wanted); all
 * arrays are [n_events_total] in the order of pb2_event_tables.event_offsets.
 * fast5_file.py:183-230 + signal_analyzer.py:311-326. */
typedef struct {
    float *mean, *stdv;              /* medfilt(5) block mean / np.std, float32                */
    float *scaled_mean;              /* poly1d(scale, shift)(mean), float32 (needs scale_shift) */
    int64_t *start, *end, *length;   /* arange(first, ., stride); start + diff, last + 1; stride */
    int64_t *pos;                    /* cumsum(move)                                            */
    double *p_model_state;           /* qual[cumsum(move) - 1 + posshift]                      */
    uint8_t *model_state;            /* [.][5] k-mer of the reversed sequence, U -> T          */
    int32_t *error;                  /* [n_reads] 0 ok; 1 unknown k-mer size (fast5_file.py:197);
                                        2 events / raw strides mismatch (fast5_file.py:221)    */
} pb2_event_columns;

/* A batch of reads: ragged int16 DAC samples + per-read calibration
 * (Fast5Reader.get_raw_data, fast5_file.py:122-131).  raw_offsets[i] is the element
 * offset of read i in `raw` and must be a multiple of 8 (16-byte aligned reads). */
typedef struct {
    int64_t n_reads;
    int64_t n_raw_total;             /* number of int16 elements in `raw` */
    int64_t max_raw_length;          /* max(raw_lengths) if known on the host, else 0 */
    const int16_t *raw;
    const int64_t *raw_offsets;      /* [n_reads] */
    const int64_t *raw_lengths;      /* [n_reads] samples per read (= Raw.duration) */
    const double *range;             /* [n_reads] channel_id attrs */
    const double *digitisation;      /* [n_reads] */
    const double *offset;            /* [n_reads] */
    /* Optional compressed form of `raw` for the HOST entry points: one streamvbyte-16 stream per
     * read = the body of an ONT VBZ chunk (HDF5 filter 32020, version 1, 2-byte integers, zigzag
     * deltas) after its optional zstd stage; read i occupies bytes
     * [packed_offsets[i], packed_offsets[i + 1]), offsets multiples of 16.  When `packed` is not
     * NULL pb2_analyze_host uploads these bytes (about 1.13 per sample) instead of `raw`, which
     * may be NULL, and decodes them on the device (pb2_svb16_decode) into the layout
     * raw_offsets / raw_lengths describe.  Ignored by the device-resident entry points. */
    const uint8_t *packed;
    const int64_t *packed_offsets;   /* [n_reads + 1] */
} pb2_batch;

/* Per-read results (any pointer may be NULL = not wanted). */
typedef struct {
    int32_t *status;                 /* [n] PB2_ST_*                                   */
    int32_t *label;                  /* [n] PB2_LABEL_*                                */
    float *scale_shift;              /* [n][2]  NanoporeRead.scaling_params            */
    int32_t *segments;               /* [n][PB2_MAX_STATES][2] first,last (pooled idx) */
    int32_t *barcode;                /* [n] -1 = None          set_barcode(...)        */
    int32_t *barcode_guess;          /* [n] argmax - decoys; INT32_MIN = not classified */
    int32_t *barcode_score;          /* [n] calibrated phred;  -1 = not classified     */
    float *class_probs;              /* [n][PB2_MAX_CLASSES] softmax output            */
    float *pooled;                   /* scaled pooled signal, ragged; read i starts at
                                        element (raw_offsets[i] + stride-1) / stride   */
    int64_t *counts;                 /* [PB2_N_LABEL][PB2_N_BARCODE_SLOTS][PB2_N_STATUS]
                                        FinalSummaryTracker.counts (io.py:269-278)     */
    pb2_polya_result *polya;         /* [n] with PB2_FLAG_POLYA                        */
} pb2_results;

/* ---- life cycle --------------------------------------------------------- */
int pb2_abi_version(void);
/* WorkerPersistenceStorage.init_persistence_objects (worker_persistence.py:60-90) */
int pb2_create(int device, pb2_context **out);
void pb2_destroy(pb2_context *ctx);
const char *pb2_last_error(const pb2_context *ctx);

int pb2_set_scaler(pb2_context *ctx, const pb2_scaler_params *p);
/* scan_limit_pooled = segmentation_scan_limit // stride (signal_analyzer.py:347) */
int pb2_set_segmentation_hmm(pb2_context *ctx, const pb2_hmm_params *p,
                             int32_t scan_limit_pooled, int32_t adapter_state);
int pb2_set_demux(pb2_context *ctx, const pb2_demux_params *p);
/* polya_state: baked index of the 'polya-tail' state (-1 if the model has none) */
int pb2_set_polya(pb2_context *ctx, const pb2_polya_params *p, int32_t polya_state);
/* unsplit_read_detection_model (baked) + its switches; state indices are baked indices of
 * 'adapter', 'leader-high', 'leader-low' in THAT model */
int pb2_set_unsplit(pb2_context *ctx, const pb2_hmm_params *hmm, const pb2_unsplit_params *p,
                    int32_t adapter_state, int32_t leader_high_state, int32_t leader_low_state);

/* ---- whole path --------------------------------------------------------- */
/* SignalAnalyzer.process stages A-D for the numeric outputs (signal_analyzer.py:82-134):
 * load_padded_signal_head -> fit_scalers -> load_signal(pool) -> detect_segments ->
 * [push_barcode_signal -> demuxer.predict] -> counts.  `batch`/`res` hold DEVICE
 * pointers; work is enqueued on `stream` (a cudaStream_t) and not synchronised -- except in the
 * `fast` mode (pb2_set_fast_lstm), where the call waits once on `stream` for the number of reads
 * the guards sent to the exact re-run (it sizes the sub-batch launches from it): the kernels
 * before that point have finished when it returns, the re-run, labels and counts have not. */
int pb2_analyze_device(pb2_context *ctx, const pb2_batch *batch, const pb2_results *res,
                       uint32_t flags, void *stream);
/* Same with HOST buffers: copies in, runs, copies out, synchronises.  Batches of >= 65536 reads
 * and >= 256 Mi samples whose reads lie in ascending, non-overlapping order in `raw` are uploaded
 * in chunks that overlap the kernels (whole batch resident on the device; batches over a quarter
 * of the device memory go through two chunk-sized arenas instead); pinned host buffers make the
 * copies asynchronous.  Results do not depend on the chunking except for the approximate floats
 * of the `fast` mode (pb2_set_fast_lstm). */
int pb2_analyze_host(pb2_context *ctx, const pb2_batch *batch, const pb2_results *res,
                     uint32_t flags);

/* ---- single stages over device buffers (parity tests, partial pipelines) -- */
/* fast5_file.py:122-131 + signal_loader.py:224-225,244-247: int16 -> pA -> mean-pool.
 * pooled[pooled_offset(i) + t], t < raw_lengths[i] / stride (unscaled). */
int pb2_pool_signal(pb2_context *ctx, const pb2_batch *batch, float *pooled, void *stream);
/* signal_loader.py:89-109: scaler network on the left-zero-padded pooled head.
 * z_out (optional) receives the raw network outputs [n][2]. */
int pb2_fit_scalers(pb2_context *ctx, const pb2_batch *batch, const float *pooled,
                    int32_t *status, float *scale_shift, float *z_out, void *stream);
/* signal_loader.py:258-262 + signal_analyzer.py:346-364 */
int pb2_detect_segments(pb2_context *ctx, const pb2_batch *batch, const float *pooled,
                        const float *scale_shift, int32_t *status, int32_t *segments,
                        float *pooled_scaled_out, void *stream);
/* generic Viterbi over dense float rows (pomegranate HiddenMarkovModel.viterbi):
 * x[n][ld], lengths[n] -> path[n][ld] (baked state indices), logp[n].
 * which = 0 segmentation model. */
int pb2_viterbi_paths(pb2_context *ctx, int which, const float *x, const int32_t *lengths,
                      int64_t n, int32_t ld, int32_t *path, double *logp, void *stream);
/* barcoding.py:83-101: windows[n][trim_length], pushed[n] */
int pb2_barcode_windows(pb2_context *ctx, const pb2_batch *batch, const float *pooled,
                        const float *scale_shift, const int32_t *status,
                        const int32_t *segments, float *windows, int32_t *pushed,
                        void *stream);
/* barcoding.py:103-118 on explicit windows[n][trim_length] */
int pb2_demux_predict(pb2_context *ctx, const float *windows, const int32_t *pushed,
                      int64_t n, float *class_probs, int32_t *barcode, int32_t *guess,
                      int32_t *score, void *stream);
/* scaler network on explicit heads[n][length/stride] (keras predict) */
int pb2_scaler_predict(pb2_context *ctx, const float *heads, int64_t n, float *z_out,
                       void *stream);
/* polya.py:50-187 + csupport.detect_events, for reads whose status is okay */
int pb2_measure_polya(pb2_context *ctx, const pb2_batch *batch, const float *scale_shift,
                      const int32_t *status, const int32_t *segments, pb2_polya_result *out,
                      void *stream);
/* csupport.detect_events (src/csupport.c:70-124; scrappie event_detection.c:273-324), the
 * reference's own native entry point, for a batch of float32 signals: signal i is
 * signal[offsets[i] .. offsets[i] + lengths[i]).  All pointers are DEVICE pointers.  Two calls:
 *   records == NULL: event_counts[i] = number of events of signal i (0 for an empty signal,
 *                    where csupport raises);
 *   records != NULL: the events of signal i are written as packed 28-byte records
 *                    {u8 start, f4 length, f4 mean, f4 stdv, i4 pos = -1, i4 state = -1}
 *                    (the numpy dtype of csupport.c:156-159) starting at record
 *                    event_offsets[i] (the exclusive prefix sum of the counts).
 * Window lengths up to 255 samples. */
typedef struct pb2_detector_params {
    int64_t window_length1, window_length2;    /* csupport defaults 30, 120 */
    float threshold1, threshold2, peak_height; /* 3.0, 9.0, 8.0 */
} pb2_detector_params;
int pb2_detect_events(pb2_context *ctx, const float *signal, const int64_t *offsets,
                      const int64_t *lengths, int64_t n_signals, const pb2_detector_params *p,
                      int64_t *event_counts, const int64_t *event_offsets, void *records,
                      void *stream);
/* SignalAnalysis.detect_unsplit_read (signal_analyzer.py:366-443) for reads whose status is
 * okay and that have an event table.  `batch` (the reads' raw signal, same read order) is only
 * needed when events->mean is NULL and may be NULL otherwise.  flag[i]: 1 unsplit, 0 not, <0 internal overflow/no path.
 * max_windows >= ceil((last event end - payload start) / window_step) over the batch. */
int pb2_detect_unsplit(pb2_context *ctx, const pb2_batch *batch,
                       const pb2_event_tables *events, int64_t n_reads,
                       const float *scale_shift, const int32_t *status, const int32_t *segments,
                       int32_t max_windows, int32_t *flag, void *stream);
/* same with HOST pointers everywhere (copies in, runs, copies flag out, synchronises) */
int pb2_detect_unsplit_host(pb2_context *ctx, const pb2_batch *batch,
                            const pb2_event_tables *events, int64_t n_reads,
                            const float *scale_shift, const int32_t *status,
                            const int32_t *segments, int32_t max_windows, int32_t *flag);
/* Fast5Reader.construct_events_from_moves + convert_events_guppy (fast5_file.py:183-230) and the
 * derived columns of SignalAnalysis.load_events (signal_analyzer.py:311-326) for a batch of
 * guppy Move tables.  events: event_offsets, move, first_sample, block_stride are read (mean /
 * start / p_model_state are not).  basecalls may be NULL when pos-dependent columns are not
 * wanted; scale_shift [n][2] may be NULL when scaled_mean is not wanted.  Device pointers. */
int pb2_derive_event_tables(pb2_context *ctx, const pb2_batch *batch,
                            const pb2_event_tables *events, const pb2_basecalls *basecalls,
                            const float *scale_shift, const pb2_event_columns *out, void *stream);
/* same with HOST pointers everywhere (copies in, runs, copies the columns out, synchronises) */
int pb2_derive_event_tables_host(pb2_context *ctx, const pb2_batch *batch,
                                 const pb2_event_tables *events, const pb2_basecalls *basecalls,
                                 const float *scale_shift, const pb2_event_columns *out);
/* Device-side half of the VBZ decoder (see pb2_batch.packed): streamvbyte-16 + zigzag + delta
 * -> int16 samples, one warp per read.  Device pointers; *error (may be NULL) is set to 1 when a
 * stream is shorter than its key block promises. */
int pb2_svb16_decode(pb2_context *ctx, const uint8_t *packed, const int64_t *packed_offsets,
                     const int64_t *raw_offsets, const int64_t *raw_lengths, int64_t n_reads,
                     int16_t *raw, int32_t *error, void *stream);
/* io.py:274-278 */
int pb2_count_results(pb2_context *ctx, const int32_t *status, const int32_t *label,
                      const int32_t *barcode, int64_t n, int64_t *counts, void *stream);

/* number of kernel launches issued through this context so far */
int64_t pb2_kernel_launches(const pb2_context *ctx);

/* Verification switch: run the LSTM kernels with IEEE __fdiv_rn instead of the
 * branch-free Newton division and without left-pad skipping in the demultiplexer
 * (identical outputs, slower; see csrc/pb_math.cuh, kernels_lstm.cu). */
int pb2_set_exact_division(pb2_context *ctx, int on);

/* Tensor-core LSTM path (default on).  The recurrent products of keras predict()
 * (signal_loader.py:96-97, barcoding.py:106-107) run on tcgen05 tensor cores as split-fp16
 * GEMMs; every decision taken from those approximate outputs passes a margin test and the
 * reads that fail it are re-run through the exact f32 kernels, so barcode / guess / score
 * stay those of the exact path.  The error bound the margin test assumes for a window's class
 * logits is  demux_margin_delta + demux_probe_gain * s,  s being the shift of the logits
 * under a deliberately coarse second evaluation of layer 2 (the classifier's second LSTM
 * amplifies perturbations by orders of magnitude for a small fraction of windows, so the
 * sensitivity is measured per window).  Pass 0 to keep a value.  on = 0: exact kernels only.
 * on = 2 ("strict"): scaler, segmentation and barcode windows through the exact kernels for
 * every read -- (scale, shift) and the normalised signal are then the reference's float32
 * values bit for bit -- and only the classifier on the tensor cores (margin test + exact
 * re-run of the windows it flags). */
/* Diagnostics: windows of the last classifier launch that needed the second sensitivity probe
 * (the others were settled by the first probe under the wide screening bound). */
int pb2_probe2_rows(pb2_context *ctx, int64_t *rows);
int pb2_set_fast_lstm(pb2_context *ctx, int on, double demux_margin_delta,
                      double demux_probe_gain);
/* Verification: the tensor-core demultiplexer WITHOUT the exact re-run -- approximate class
 * probabilities and logits ([n][PB2_MAX_CLASSES]), tentative calls, and the margin-test verdict
 * per window (unsafe[i] = 1: this window would be re-run; sensitivity[i] = its measured logit
 * shift).  Device pointers; any may be NULL. */
int pb2_demux_predict_tc(pb2_context *ctx, const float *windows, int64_t n, float *class_probs,
                         float *logits, int32_t *barcode, int32_t *guess, int32_t *score,
                         int32_t *unsafe, float *sensitivity, void *stream);
/* Verification: layer-1 outputs of the exact kernels, every position stepped:
 * out[n][signal_trim_length][2 * units] (forward | backward), n <= 4096.  Device pointers. */
int pb2_debug_demux_l1(pb2_context *ctx, const float *windows, int64_t n, float *out, void *stream);
/* After a call: how many windows the last demultiplexer launch re-ran exactly, and whether a
 * tensor-core kernel hit its barrier time-out (results invalid if non-zero).  Synchronises. */
int pb2_recheck_stats(pb2_context *ctx, int64_t *demux_rechecked, int64_t *tc_timeouts);
/* Audit of the tensor-core path in production: additionally re-run a pseudo-random `fraction`
 * (0..1) of the reads that PASSED every guard through the exact kernels and count those whose
 * status / segments / barcode / guess / score differ from the tensor-core values (the exact
 * values are returned either way).  pb2_audit_stats reads and clears the two counters. */
int pb2_set_audit_fraction(pb2_context *ctx, double fraction);
int pb2_audit_stats(pb2_context *ctx, int64_t *audited, int64_t *mismatched);
/* Of the reads the last pb2_analyze_device call re-ran exactly: how many because the QC verdict,
 * the segmentation, the barcode call was inside its error margin (a read can count twice). */
int pb2_rerun_causes(pb2_context *ctx, int64_t *qc, int64_t *segmentation, int64_t *barcode);

/* ---- measurement ----------------------------------------------------------
 * With profiling on, every kernel launch is bracketed by CUDA events on its own
 * stream.  pb2_profile_read synchronises the device, adds the elapsed times up per
 * kernel and resets the event list.  Kernel ids: 0..pb2_profile_kernel_count()-1. */
int pb2_profile_enable(pb2_context *ctx, int on);
int pb2_profile_kernel_count(void);
const char *pb2_profile_kernel_name(int kernel_id);
int pb2_profile_read(pb2_context *ctx, double *total_ms, int64_t *launches, int n_kernels);
/* The same records as a time line instead of sums (and, like pb2_profile_read, consumed by the
 * call): kernel id, start and end of every launch in milliseconds after the first launch's start,
 * in launch order; at most `capacity` entries, *n_out = entries written.  Shows where a stream
 * sat idle between launches (tools/host_path_profile.py). */
int pb2_profile_timeline(pb2_context *ctx, int32_t *ids, double *start_ms, double *end_ms,
                         int64_t capacity, int64_t *n_out);

#ifdef __cplusplus
}
#endif
#endif /* POREPLEX_B200_H */
