/* pb_oracle.h -- CPU ORACLE interface (test infrastructure, not product code).
 * See pb_oracle.c for the reference citations. */
#ifndef PB_ORACLE_H
#define PB_ORACLE_H
#include <stdint.h>

#define ORC_MAX_STATES 8
#define ORC_MAX_COMP 4
#define ORC_MAX_EDGES 64
#define ORC_MAX_UNITS 64
#define ORC_MAX_GATES 256
#define ORC_MAX_CLASSES 8
#define ORC_MAX_WINDOW 512
#define ORC_MAX_CALIB 64

/* status codes: poreplex/io.py:245-260 order */
enum {
    ORC_OKAY = 0, ORC_DISAPPEARED, ORC_IRREGULAR_FAST5, ORC_SCALER_SIGNAL_TOO_SHORT,
    ORC_SCALING_QC_FAIL, ORC_ADAPTER_NOT_DETECTED, ORC_NOT_BASECALLED,
    ORC_BASECALL_TABLE_INCOMPLETE, ORC_UNSPLIT_READ, ORC_SEQUENCE_TOO_SHORT,
    ORC_UNKNOWN_ERROR
};

#define ORC_FLAG_BARCODING 1

typedef struct {
    int32_t in_dim, units, impl;      /* impl: Keras `implementation` (1 or 2) */
    const float *W;                   /* [in_dim][4*units] */
    const float *U;                   /* [units][4*units]  */
    const float *b;                   /* [4*units]         */
} orc_lstm;

typedef struct {
    orc_lstm l1, l2;
    const float *Wd, *bd;             /* Dense(2): [units][2], [2] */
} orc_scaler;

typedef struct {
    orc_lstm fwd, bwd, l2;
    const float *Wd, *bd;             /* Dense(n_classes) */
    int32_t n_classes;
} orc_demux;

typedef struct {
    int32_t n_states;
    int32_t n_comp[ORC_MAX_STATES];
    double mu[ORC_MAX_STATES][ORC_MAX_COMP];
    double lsp[ORC_MAX_STATES][ORC_MAX_COMP];      /* -log(sigma * SQRT_2_PI)  */
    double inv2s2[ORC_MAX_STATES][ORC_MAX_COMP];   /* 1 / (2 sigma^2)          */
    double logw[ORC_MAX_STATES][ORC_MAX_COMP];     /* log mixture weights      */
    double log_start[ORC_MAX_STATES];              /* -inf: no start edge      */
    int32_t in_begin[ORC_MAX_STATES + 1];          /* CSR over destination     */
    int32_t in_src[ORC_MAX_EDGES];
    double in_logp[ORC_MAX_EDGES];
} orc_hmm;

typedef struct {
    uint64_t start;
    float length, mean, stdv;
    int32_t pos, state;
} orc_event;

typedef struct {
    orc_scaler scaler;
    orc_demux demux;
    orc_hmm seg;
    int32_t stride, scaler_length, scaler_min_length, scan_limit;
    double scale_std, scale_mean, shift_std, shift_mean;
    double qc_scale[2], qc_shift[2];
    int32_t adapter_state;
    int32_t demux_minlen, demux_maxlen, demux_trimlen, n_decoy;
    float demux_pad;
    int32_t n_calib;
    double calib[ORC_MAX_CALIB];
    double score_threshold;
} orc_model;

typedef struct {
    int32_t status;
    float z[2];
    float scale, shift;
    double viterbi_logp;
    int32_t seg[ORC_MAX_STATES][2];
    int32_t pushed;
    int32_t barcode, guess, phred;
    float probs[ORC_MAX_CLASSES];
} orc_result;

float orc_tanhf(float);
float orc_sigmoidf(float);
float orc_expf(float);
double orc_exp_neg(double);
double orc_log_1to2(double);
double orc_pair_lse(double, double);

void orc_dac_to_pa(const int16_t *raw, int64_t n, double gain, double offset, float *out);
void orc_pool_mean(const float *x, int64_t npooled, int stride, float *out);
void orc_scale(const float *x, int64_t n, float scale, float shift, float *out);

void orc_lstm_step(const orc_lstm *L, const float *x, float *h, float *c);
void orc_lstm_seq(const orc_lstm *L, const float *x, int T, int reverse, float *hseq, float *hlast);
void orc_scaler_predict(const orc_scaler *S, const float *head, int T, float z[2]);
void orc_demux_predict(const orc_demux *D, const float *win, int T, float *probs);

void orc_hmm_emissions(const orc_hmm *M, double x, double *e);
double orc_viterbi(const orc_hmm *M, const float *x, int T, int32_t *path);
void orc_segments_from_path(const int32_t *path, int T, int n_states, int32_t *seg);

int orc_barcode_window(const float *adapter, int len, int minlen, int maxlen,
                       int trimlen, float pad, float *out);
void orc_barcode_decide(const float *probs, int n_classes, int n_decoy,
                        const double *calib, int n_calib, double score_threshold,
                        int32_t *barcode, int32_t *guess, int32_t *phred);

int64_t orc_detect_events(const float *x, int64_t n, int64_t w1, int64_t w2,
                          float thr1, float thr2, float peak_height,
                          orc_event *ev, int64_t max_events);

void orc_process_read(const orc_model *M, const int16_t *raw, int64_t n,
                      double gain, double offset, int flags, orc_result *R);
void orc_process_batch(const orc_model *M, const int16_t *raw, const int64_t *offsets,
                       const int64_t *lengths, const double *gain, const double *offset,
                       int64_t N, int flags, int nthreads, orc_result *R);
int orc_max_threads(void);
#endif
