/* poreplex_b200_fast5.h -- C ABI of libpb_fast5.so: FAST5 ingest for the signal path.
 *
 * Host-only (no CUDA): reads single- and multi-read FAST5 (HDF5) files without libhdf5 and packs
 * the int16 Signal datasets of a batch of reads, with their calibration, into the ragged layout
 * pb2_analyze_host takes (include/poreplex_b200.h, pb2_batch: every read starts on a 16-byte
 * boundary).  It replaces, for the signal path, what the reference does per read through h5py:
 *   Fast5Reader.__init__ / load_metadata   poreplex/fast5_file.py:65-120
 *   Fast5Reader.get_raw_data (the dataset read; the int16 -> pA conversion stays on the GPU)
 *                                          poreplex/fast5_file.py:122-131
 *   NanoporeRead.__init__ / open           poreplex/signal_loader.py:117-128,200-210
 * (the step immediately before the accelerated path.)
 *
 * Supported HDF5 subset: superblock version 0 (8-byte offsets), version-1 object headers,
 * symbol-table groups, contiguous / compact / chunked (v1 B-tree) datasets, filters deflate (1;
 * own table-driven inflater, no zlib), shuffle (2), fletcher32 (3) and ONT VBZ (32020; needs libzstd.so.1 at run time).  Anything else
 * is reported as PB2F_EFORMAT, never guessed at.
 */
#ifndef POREPLEX_B200_FAST5_H
#define POREPLEX_B200_FAST5_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB2F_ABI_VERSION 1

/* return codes (negative = error) */
#define PB2F_OK 0
#define PB2F_EINVAL (-1)
#define PB2F_EIO (-2)       /* cannot open / map the file */
#define PB2F_EFORMAT (-3)   /* not HDF5, or a feature outside the supported subset */
#define PB2F_ENOTFOUND (-4) /* no such read / node / attribute */
#define PB2F_ENOSPC (-5)    /* destination too small */

/* per-read status of a batch (the reference's vocabulary, io.py:245-260) */
#define PB2F_READ_OK 0
#define PB2F_READ_DISAPPEARED 1 /* file missing: signal_analyzer.py:90-92 */
#define PB2F_READ_IRREGULAR 2   /* unreadable file / unknown read / bad layout: irregular_fast5 */

typedef struct pb2f_file pb2f_file;
typedef struct pb2f_batch pb2f_batch;

typedef struct {
    int64_t signal_length;                             /* elements of .../Signal */
    int64_t duration, start_time;                      /* Raw attributes (fast5_file.py:101-103) */
    double digitisation, offset, range, sampling_rate; /* channel_id (fast5_file.py:109-114) */
    char read_id[64];                                  /* Raw/read_id */
    char channel_number[16];
    char run_id[64], sample_id[64];                    /* tracking_id (fast5_file.py:116-119) */
} pb2f_read_meta;

int pb2f_abi_version(void);
/* message of the last error raised on the calling thread */
const char *pb2f_last_error(void);

/* ---- one file ------------------------------------------------------------------------- */
int pb2f_open(const char *path, pb2f_file **out);
void pb2f_close(pb2f_file *f);
/* 1 for a multi-read file ('UniqueGlobalKey' absent, fast5_file.py:69), else 0 */
int pb2f_is_multiread(const pb2f_file *f);
/* reads in the file; names are the 'read_<id>' group names without the prefix (multi-read) or
 * the Raw/Reads/<name> group name (single-read), in the file's own (sorted) order */
int64_t pb2f_num_reads(pb2f_file *f);
const char *pb2f_read_name(pb2f_file *f, int64_t index);
/* read_id: NULL = the first read (single-read files) */
int pb2f_read_meta_get(pb2f_file *f, const char *read_id, pb2f_read_meta *out);
/* whole Signal dataset as int16; returns the number of samples or a negative error */
int64_t pb2f_read_signal(pb2f_file *f, const char *read_id, int16_t *dst, int64_t capacity);

/* ---- a batch of (path, read_id) pairs -------------------------------------------------- */
/* Phase 1: open the files (each distinct path once), locate the reads, read their metadata on
 * n_threads workers.  Never fails per read: see status. */
int pb2f_batch_open(const char *const *paths, const char *const *read_ids, int64_t n_reads,
                    int n_threads, pb2f_batch **out);
/* per-read arrays of n_reads elements each (any may be NULL) */
int pb2f_batch_meta(const pb2f_batch *b, int32_t *status, int64_t *signal_length,
                    double *range, double *digitisation, double *offset, double *sampling_rate,
                    int64_t *duration, int64_t *start_time);
int pb2f_batch_meta_full(const pb2f_batch *b, int64_t index, pb2f_read_meta *out);
/* raw_offsets[i] for the packed layout (each read on an 8-sample boundary; reads whose status is
 * not PB2F_READ_OK get length 0); returns the total number of int16 elements needed */
int64_t pb2f_batch_plan(const pb2f_batch *b, int64_t *raw_offsets, int64_t *raw_lengths);
/* Phase 2: decode every readable Signal into raw + raw_offsets[i] on n_threads workers.  A read
 * that fails to decode gets status PB2F_READ_IRREGULAR (visible through pb2f_batch_meta). */
int pb2f_batch_read(pb2f_batch *b, int16_t *raw, int64_t raw_capacity, const int64_t *raw_offsets,
                    int n_threads);
void pb2f_batch_close(pb2f_batch *b);

/* The inflater the deflate filter uses (csrc_host/inflate_fast.h), on its own: one complete zlib
 * stream (RFC 1950) -> dst.  Returns the number of bytes written, PB2F_ENOSPC when dst is too
 * small, PB2F_EFORMAT for truncated / corrupt input or an Adler-32 mismatch. */
int64_t pb2f_inflate(const void *src, int64_t src_len, void *dst, int64_t dst_capacity);

/* The compressed upload form of a packed int16 batch (include/poreplex_b200.h, pb2_batch.packed):
 * one streamvbyte-16 stream of zigzag deltas per read -- the body of an ONT VBZ chunk (filter
 * 32020, version 1, 2-byte integers) without its zstd stage -- each starting on a 16-byte
 * boundary.  pb2f_svb16_plan fills packed_offsets[n_reads + 1] and returns the buffer size;
 * pb2f_svb16_encode writes the streams. */
int64_t pb2f_svb16_plan(const int16_t *raw, const int64_t *raw_offsets, const int64_t *raw_lengths,
                        int64_t n_reads, int n_threads, int64_t *packed_offsets);
int pb2f_svb16_encode(const int16_t *raw, const int64_t *raw_offsets, const int64_t *raw_lengths,
                      int64_t n_reads, int n_threads, const int64_t *packed_offsets, uint8_t *packed);

#ifdef __cplusplus
}
#endif
#endif
