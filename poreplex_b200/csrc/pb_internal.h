// pb_internal.h -- context and kernel-launcher declarations shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/poreplex_b200.h"

namespace pb {

// ---- device-side parameter blocks ------------------------------------------
struct HmmDev {                       // lives in __constant__ memory per kernel launch arg
    int32_t n_states;
    int32_t n_comp[PB2_MAX_STATES];
    double mu[PB2_MAX_STATES][PB2_MAX_COMP];
    double log_norm[PB2_MAX_STATES][PB2_MAX_COMP];
    double inv_two_var[PB2_MAX_STATES][PB2_MAX_COMP];
    double log_weight[PB2_MAX_STATES][PB2_MAX_COMP];
    double log_start[PB2_MAX_STATES];
    int32_t in_begin[PB2_MAX_STATES + 1];
    int32_t in_src[PB2_MAX_EDGES];
    double in_logp[PB2_MAX_EDGES];
};

struct LstmDev {                      // device copies of one layer's weights
    int in_dim = 0, units = 0, impl = 0;
    float *kernel = nullptr;          // [in_dim][4H]  as given (row-major)
    float *recurrent = nullptr;       // [H][4H]
    float *bias = nullptr;            // [4H]
};

struct ScalerDev {
    bool set = false;
    LstmDev l1, l2;
    float *dense_kernel = nullptr, *dense_bias = nullptr;
    int stride = 15, length = 30000, min_length = 9000;
    double scale_std = 0, scale_mean = 0, shift_std = 0, shift_mean = 0;
    double qc_scale_lo = 0, qc_scale_hi = 0, qc_shift_lo = 0, qc_shift_hi = 0;
    // state of both layers after n all-zero input steps: [steps+1][4][units]
    // (h1, c1, h2, c2); lets fit_scalers skip the left zero padding bit-exactly
    float *zero_prefix = nullptr;
    int zero_prefix_steps = 0;
};

struct DemuxDev {
    bool set = false;
    LstmDev fwd, bwd, l2;
    float *dense_kernel = nullptr, *dense_bias = nullptr;
    int n_classes = 0, n_decoy = 0, min_length = 0, max_length = 0, trim_length = 0;
    float pad_value = -1000.f;
    int n_calibration = 0;
    double calibration[PB2_MAX_CALIB];
    double score_threshold = 0;
    // left-pad skipping tables (see kernels_lstm.cu DemuxArgs)
    float *pad_state = nullptr;       // [T+1][2][H1]
    float *pad_prefix = nullptr;      // [T][H2/2][4][2]
    double *calibration_dev = nullptr;   // device copy of `calibration` (tensor-core path)
};

struct Workspace {                    // grow-only device scratch
    void *ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace pb

namespace pb {
enum KernelId { K_POOL = 0, K_SCALER_PREPARE, K_SCALER_LSTM, K_SEGMENT, K_VITERBI_PATHS,
                K_WINDOWS, K_DEMUX_L1, K_DEMUX_L2, K_FINALIZE, K_COUNTS, K_MISC, K_POLYA, K_UNSPLIT_WINDOWS,
                K_UNSPLIT_DECIDE, K_EVENT_MEANS, K_DEMUX_TC_L1, K_DEMUX_TC_L2, K_DEMUX_TC_HEAD,
                K_SCALER_TC_L1, K_SCALER_TC_L2, K_SCALER_TC_HEAD, K_DEMUX_TC_PROBE, K_EVENT_POS, K_SVB_DECODE,
                K_SCALER_TC, K_NUM };
struct ProfEvent { int id; cudaEvent_t a, b; };
}

struct pb2_context {
    int device = 0;
    bool profiling = false;
    bool attr_scaler = false, attr_demux = false;   // max-dynamic-smem attributes set
    bool attr_demux_tc = false, attr_scaler_tc = false, attr_scaler_tc2 = false;
    // tensor-core LSTM path (kernels_lstm_tc.cu): approximate outputs + margin test + exact
    // re-run of the reads whose decisions are not safe.  Off = exact kernels only.
    bool fast_lstm = true;
    bool strict_tc_demux = false;    // exact scaler / segmentation / windows, tensor-core classifier
    // per-window bound on the logit error = delta + probe_gain * (logit shift of the coarse probe)
    double demux_margin_delta = 1e-3;
    double demux_probe_gain = 0.1;
    int demux_probes = 2;                         // coarse probe evaluations of layer 2 (1 or 2)
    // Probe 2 only where probe 1 cannot settle the call: a call that is safe under the bound
    // delta + screen_gain * (shift of probe 1 alone) is final; the rest get the second probe and
    // the two-probe rule.  0 (default) = second probe for every window.  Measured on B200
    // (POREPLEX_B200_SCREEN_GAIN=10, the smallest gain with a comfortable margin over the worst
    // single-probe under-estimate seen): 39 % of the windows still need probe 2 -- the 29
    // calibration edges are close together -- and re-running layer 1 for them costs what the
    // skipped probes save (334.1 vs 334.9 ms per step); the list order also makes the unsafe
    // set depend on warp scheduling.  Kept as a knob, off.
    double demux_screen_gain = 0.0;
    int64_t last_probe2_rows = 0;
    int *probe2_count_dev = nullptr;              // device count of the last launch (diagnostics)
    // assumed bounds on the error of the scaler's two raw outputs (z0 -> scale, z1 -> shift); the
    // shift output has the heavier tail (largest seen on 1 M reads: 4.4e-5 / 2.9e-4)
    double scaler_margin_z0 = 2.5e-4, scaler_margin_z1 = 1.5e-3;
    bool demux_tc_ran = false;
    uint32_t audit_threshold = 0;                 // fraction of guard-passing reads re-run to compare, x 2^32
    int *tc_err = nullptr;                        // device word: a tensor-core kernel timed out
    int64_t last_rerun_cause[3] = {0, 0, 0};      // ... because of QC edge / segmentation / barcode call
    int64_t last_rerun_reads = 0;                 // reads the last whole-path call re-ran exactly
    size_t tc_scratch_bytes = (size_t)14 << 30;   // layer-1 sequence scratch per pass
    bool no_pad_skip = false;      // verification mode: step every padded position
    bool exact_division = false;   // verification mode: IEEE __fdiv_rn in the LSTM kernels
    bool generic_viterbi = false;  // verification mode: never use the topology-specialised k_segment
    std::vector<pb::ProfEvent> prof_events;
    std::vector<cudaEvent_t> prof_pool;
    std::string error;
    int64_t launches = 0;
    int sm_count = 148;
    pb::ScalerDev scaler;
    pb::DemuxDev demux;
    pb::HmmDev seg_hmm;
    bool seg_set = false;
    int scan_limit_pooled = 6666;
    int adapter_state = 0;
    pb2_polya_params polya;
    pb::HmmDev unsplit_hmm;
    pb2_unsplit_params unsplit;
    int unsplit_states[3] = {-1, -1, -1};
    bool unsplit_set = false;
    bool polya_set = false;
    int polya_state = -1;
    // scratch
    pb::Workspace ws_pooled, ws_status, ws_label, ws_scale, ws_seg, ws_win, ws_pushed,
        ws_probs, ws_bc, ws_guess, ws_score, ws_h1, ws_bp, ws_counts, ws_batch, ws_misc,
        ws_heads, ws_flags, ws_slots, ws_polya, ws_unsplit, ws_unsplit_host, ws_tstart, ws_evmean,
        ws_hlast, ws_recheck, ws_win2, ws_read2, ws_tcmisc, ws_fast, ws_sub, ws_slotof, ws_probe2;
    // host staging for pb2_analyze_host
    cudaStream_t host_stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // pipelined host path
    int64_t *counts_host = nullptr;                       // pinned, per-chunk counts
    size_t counts_host_bytes = 0;
    int64_t *stage_host = nullptr;                        // pinned, per-chunk rebased offsets (2 arenas)
    size_t stage_host_bytes = 0;
};

namespace pb {

int fail(pb2_context *ctx, int code, const char *fmt, ...);
int check_cuda(pb2_context *ctx, cudaError_t e, const char *what);
void *ws_get(pb2_context *ctx, Workspace &w, size_t bytes);   // nullptr on failure

#define PB_CUDA(ctx, call)                                               \
    do {                                                                 \
        cudaError_t _e = (call);                                         \
        if (_e != cudaSuccess) return pb::check_cuda(ctx, _e, #call);    \
    } while (0)

#define PB_LAUNCH_CHECK(ctx, name)                                       \
    do {                                                                 \
        (ctx)->launches++;                                               \
        cudaError_t _e = cudaGetLastError();                             \
        if (_e != cudaSuccess) return pb::check_cuda(ctx, _e, name);     \
    } while (0)

void prof_begin(pb2_context *ctx, int id, cudaStream_t st);
void prof_end(pb2_context *ctx, cudaStream_t st);

// launch `...` (a <<<>>> expression) as kernel `id`, timed when profiling is on
#define PB_LAUNCH(ctx, id, name, st, ...)                                \
    do {                                                                 \
        if ((ctx)->profiling) pb::prof_begin(ctx, id, st);               \
        __VA_ARGS__;                                                     \
        if ((ctx)->profiling) pb::prof_end(ctx, st);                     \
        PB_LAUNCH_CHECK(ctx, name);                                      \
    } while (0)

// pooled element offset of a read whose raw data starts at element `raw_off`
__host__ __device__ inline int64_t pooled_offset(int64_t raw_off, int stride) {
    return (raw_off + stride - 1) / stride;
}

// ---- kernel launchers (defined in the kernels_*.cu files) ------------------
int launch_pool(pb2_context *ctx, const pb2_batch &b, int stride, float *pooled,
                cudaStream_t st);
int launch_scaler(pb2_context *ctx, const pb2_batch &b, const float *pooled,
                  int32_t *status, float *scale_shift, float *z_out, cudaStream_t st);
int launch_scaler_heads(pb2_context *ctx, const float *heads, int64_t n, float *z_out,
                        cudaStream_t st);
int build_zero_prefix(pb2_context *ctx);
int build_pad_tables(pb2_context *ctx);
int launch_segment(pb2_context *ctx, const pb2_batch &b, const float *pooled,
                   const float *scale_shift, int32_t *status, int32_t *segments,
                   float *pooled_scaled_out, cudaStream_t st);
int launch_segment3(pb2_context *ctx, const pb2_batch &b, const float *pooled, const float *ss3,
                    int32_t *const status[3], int32_t *const segments[3], cudaStream_t st);
int launch_viterbi_paths(pb2_context *ctx, const HmmDev &hmm, const float *x,
                         const int32_t *lengths, int64_t n, int32_t ld, int32_t *path,
                         double *logp, cudaStream_t st);
int launch_windows(pb2_context *ctx, const pb2_batch &b, const float *pooled,
                   const float *scale_shift, const int32_t *status, const int32_t *segments,
                   float *windows, int32_t *pushed, int *slot_count, int32_t *slot_read,
                   cudaStream_t st);
// slot_count / slot_read: compacted windows (row s belongs to read slot_read[s], *slot_count
// rows are valid); both nullptr = windows[n] aligned with reads, optional `pushed` mask
int launch_demux(pb2_context *ctx, const float *windows, const int32_t *pushed, int64_t n,
                 const int *slot_count, const int32_t *slot_read,
                 float *class_probs, int32_t *barcode, int32_t *guess, int32_t *score,
                 cudaStream_t st);
// exact f32 kernels (kernels_lstm.cu) / tensor-core path with margin test (kernels_lstm_tc.cu)
int launch_demux_exact(pb2_context *ctx, const float *windows, const int32_t *pushed, int64_t n,
                       const int *slot_count, const int32_t *slot_read,
                       float *class_probs, int32_t *barcode, int32_t *guess, int32_t *score,
                       cudaStream_t st);
int launch_scaler_prepare(pb2_context *ctx, const pb2_batch &b, int32_t *status, float *scale_shift,
                          int64_t *xoff, int32_t *nreal, cudaStream_t st);
int launch_scaler_tc(pb2_context *ctx, const pb2_batch &b, const float *pooled, int32_t *status,
                     float *scale_shift, float *ss_vertex, int32_t *read_unsafe, float *z_out,
                     cudaStream_t st);
int launch_compare_corners(pb2_context *ctx, int64_t n, const int32_t *st0, const int32_t *st1,
                           const int32_t *st2, const int32_t *sg0, const int32_t *sg1,
                           const int32_t *sg2, int32_t *read_unsafe, cudaStream_t st);
int debug_demux_l1(pb2_context *ctx, const float *windows, int64_t n, float *out, cudaStream_t st);
int launch_demux_tc(pb2_context *ctx, const float *windows, const int32_t *pushed, int64_t n,
                    const int *slot_count, const int32_t *slot_read,
                    float *class_probs, int32_t *barcode, int32_t *guess, int32_t *score,
                    float *logits_out, int32_t *unsafe_out, float *sens_out, bool recheck,
                    cudaStream_t st, int32_t *read_unsafe = nullptr);
int launch_detect_events(pb2_context *ctx, const float *signal, const int64_t *offsets,
                         const int64_t *lengths, int64_t n, const pb2_detector_params &dp,
                         int64_t *counts, const int64_t *event_offsets, void *records,
                         cudaStream_t st);
int launch_polya(pb2_context *ctx, const pb2_batch &b, const float *scale_shift,
                 const int32_t *status, const int32_t *segments, pb2_polya_result *out,
                 cudaStream_t st);
int launch_svb16_decode(pb2_context *ctx, const uint8_t *packed, const int64_t *packed_offsets,
                        const int64_t *raw_offsets, const int64_t *raw_lengths, int64_t n,
                        int16_t *raw, int32_t *error, cudaStream_t st);
int launch_derive_events(pb2_context *ctx, const pb2_batch &b, const pb2_event_tables &ev,
                         const pb2_basecalls *bc, const float *scale_shift,
                         const pb2_event_columns &out, cudaStream_t st);
int launch_unsplit(pb2_context *ctx, const pb2_batch *batch, const pb2_event_tables &ev,
                   int64_t n, const float *scale_shift, const int32_t *status,
                   const int32_t *segments, int32_t max_windows, int32_t *flag, cudaStream_t st);
int launch_finalize(pb2_context *ctx, int64_t n, uint32_t flags, int32_t *status,
                    int32_t *label, int32_t *barcode, int32_t *guess, int32_t *score,
                    cudaStream_t st, const int32_t *pushed_mask = nullptr);
int launch_counts(pb2_context *ctx, const int32_t *status, const int32_t *label,
                  const int32_t *barcode, int64_t n, int64_t *counts, cudaStream_t st);

}  // namespace pb
