// api.cu -- the extern "C" boundary declared in include/poreplex_b200.h:
// context life cycle, parameter upload, stage entry points and the whole-path
// pipeline (device-resident and host-buffer variants).
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <new>

#include "pb_internal.h"

namespace pb {

int fail(pb2_context *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    return code;
}

int check_cuda(pb2_context *ctx, cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return PB2_OK;
    return fail(ctx, PB2_ECUDA, "CUDA error in %s: %s", what, cudaGetErrorString(e));
}

void *ws_get(pb2_context *ctx, Workspace &w, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (w.bytes >= bytes) return w.ptr;
    if (w.ptr) { cudaFree(w.ptr); w.ptr = nullptr; w.bytes = 0; }
    // grow with some slack so that slightly larger batches do not reallocate
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&w.ptr, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&w.ptr, want);
    }
    if (e != cudaSuccess) {
        w.ptr = nullptr;
        fail(ctx, PB2_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        return nullptr;
    }
    w.bytes = want;
    return w.ptr;
}

static cudaEvent_t prof_event(pb2_context *ctx)
{
    if (!ctx->prof_pool.empty()) {
        cudaEvent_t e = ctx->prof_pool.back();
        ctx->prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void prof_begin(pb2_context *ctx, int id, cudaStream_t st)
{
    ProfEvent pe{id, prof_event(ctx), prof_event(ctx)};
    cudaEventRecord(pe.a, st);
    ctx->prof_events.push_back(pe);
}

void prof_end(pb2_context *ctx, cudaStream_t st)
{
    if (!ctx->prof_events.empty()) cudaEventRecord(ctx->prof_events.back().b, st);
}

static const char *const kKernelNames[K_NUM] = {
    "k_pool", "k_scaler_prepare", "k_scaler_lstm", "k_segment", "k_viterbi_paths",
    "k_windows", "k_demux_l1", "k_demux_l2", "k_finalize", "k_counts", "misc", "k_polya",
    "k_unsplit_windows", "k_unsplit_decide", "k_event_means", "k_lstm_tc_demux_l1",
    "k_lstm_tc_demux_l2", "k_demux_head_tc", "k_lstm_tc_scaler_l1", "k_lstm_tc_scaler_l2",
    "k_scaler_head_tc", "k_lstm_tc_demux_l2_probe", "k_event_pos", "k_svb16_decode",
    "k_lstm_tc_scaler"};

static void ws_free(Workspace &w)
{
    if (w.ptr) cudaFree(w.ptr);
    w.ptr = nullptr;
    w.bytes = 0;
}

static int upload(pb2_context *ctx, const float *host, size_t n, float **dev)
{
    if (*dev) { cudaFree(*dev); *dev = nullptr; }
    if (!host) return fail(ctx, PB2_EINVAL, "null weight pointer");
    PB_CUDA(ctx, cudaMalloc(dev, sizeof(float) * n));
    PB_CUDA(ctx, cudaMemcpy(*dev, host, sizeof(float) * n, cudaMemcpyHostToDevice));
    return PB2_OK;
}

static int upload_lstm(pb2_context *ctx, const pb2_lstm_weights &w, LstmDev &d)
{
    if (w.units <= 0 || w.in_dim <= 0 || (w.implementation != 1 && w.implementation != 2))
        return fail(ctx, PB2_EINVAL, "bad LSTM description");
    d.in_dim = w.in_dim; d.units = w.units; d.impl = w.implementation;
    int rc;
    if ((rc = upload(ctx, w.kernel, (size_t)w.in_dim * 4 * w.units, &d.kernel))) return rc;
    if ((rc = upload(ctx, w.recurrent, (size_t)w.units * 4 * w.units, &d.recurrent))) return rc;
    if ((rc = upload(ctx, w.bias, (size_t)4 * w.units, &d.bias))) return rc;
    return PB2_OK;
}

static void free_lstm(LstmDev &d)
{
    cudaFree(d.kernel); cudaFree(d.recurrent); cudaFree(d.bias);
    d.kernel = d.recurrent = d.bias = nullptr;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace pb

using namespace pb;

extern "C" {

int pb2_abi_version(void) { return PB2_ABI_VERSION; }

int pb2_create(int device, pb2_context **out)
{
    if (!out) return PB2_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return PB2_ECUDA;          // no GPU: the product path has no CPU fallback
    }
    pb2_context *ctx = new (std::nothrow) pb2_context();
    if (!ctx) return PB2_ENOMEM;
    ctx->device = device;
    DeviceGuard g(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->host_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return PB2_ECUDA;
    }
    if (const char *env = getenv("POREPLEX_B200_SCREEN_GAIN")) {       // tuning / measurements
        const double v = atof(env);
        if (v >= 0) ctx->demux_screen_gain = v;
    }
    if (const char *env = getenv("POREPLEX_B200_DEMUX_PROBES")) {      // tuning / measurements
        const int v = atoi(env);
        if (v == 1 || v == 2) ctx->demux_probes = v;
    }
    // sticky time-out word of the tensor-core kernels (read and cleared by pb2_recheck_stats)
    // (+ [2], [3]: reads audited / audited reads whose exact results differed, pb2_audit_stats)
    if (cudaMalloc(&ctx->tc_err, 4 * sizeof(int)) != cudaSuccess ||
        cudaMemset(ctx->tc_err, 0, 4 * sizeof(int)) != cudaSuccess) {
        cudaStreamDestroy(ctx->host_stream);
        delete ctx;
        return PB2_ECUDA;
    }
    *out = ctx;
    return PB2_OK;
}

void pb2_destroy(pb2_context *ctx)
{
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    free_lstm(ctx->scaler.l1); free_lstm(ctx->scaler.l2);
    cudaFree(ctx->scaler.dense_kernel); cudaFree(ctx->scaler.dense_bias);
    cudaFree(ctx->scaler.zero_prefix);
    free_lstm(ctx->demux.fwd); free_lstm(ctx->demux.bwd); free_lstm(ctx->demux.l2);
    cudaFree(ctx->demux.dense_kernel); cudaFree(ctx->demux.dense_bias);
    cudaFree(ctx->demux.pad_state); cudaFree(ctx->demux.pad_prefix);
    cudaFree(ctx->demux.calibration_dev);
    cudaFree(ctx->tc_err);
    if (ctx->counts_host) cudaFreeHost(ctx->counts_host);
    if (ctx->stage_host) cudaFreeHost(ctx->stage_host);
    Workspace *all[] = {&ctx->ws_pooled, &ctx->ws_status, &ctx->ws_label, &ctx->ws_scale,
                        &ctx->ws_seg, &ctx->ws_win, &ctx->ws_pushed, &ctx->ws_probs,
                        &ctx->ws_bc, &ctx->ws_guess, &ctx->ws_score, &ctx->ws_h1, &ctx->ws_bp,
                        &ctx->ws_counts, &ctx->ws_batch, &ctx->ws_misc, &ctx->ws_heads,
                        &ctx->ws_flags, &ctx->ws_slots, &ctx->ws_polya,
                        &ctx->ws_unsplit, &ctx->ws_unsplit_host, &ctx->ws_tstart, &ctx->ws_evmean,
                        &ctx->ws_hlast, &ctx->ws_recheck, &ctx->ws_win2, &ctx->ws_read2,
                        &ctx->ws_tcmisc, &ctx->ws_fast, &ctx->ws_sub,
                        &ctx->ws_slotof, &ctx->ws_probe2};
    for (Workspace *w : all) ws_free(*w);
    for (const ProfEvent &pe : ctx->prof_events) { cudaEventDestroy(pe.a); cudaEventDestroy(pe.b); }
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->host_stream) cudaStreamDestroy(ctx->host_stream);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    delete ctx;
}

const char *pb2_last_error(const pb2_context *ctx)
{
    return ctx ? ctx->error.c_str() : "no context (CUDA device unavailable?)";
}

int64_t pb2_kernel_launches(const pb2_context *ctx) { return ctx ? ctx->launches : 0; }

int pb2_profile_enable(pb2_context *ctx, int on)
{
    if (!ctx) return PB2_EINVAL;
    ctx->profiling = on != 0;
    return PB2_OK;
}

int pb2_set_exact_division(pb2_context *ctx, int on)
{
    if (!ctx) return PB2_EINVAL;
    ctx->exact_division = on != 0;
    ctx->no_pad_skip = on != 0;        // verification mode also steps every padded position
    ctx->generic_viterbi = on != 0;    // ... and decodes with the generic (any-topology) Viterbi step
    return PB2_OK;
}

int pb2_set_fast_lstm(pb2_context *ctx, int on, double demux_margin_delta, double demux_probe_gain)
{
    if (!ctx) return PB2_EINVAL;
    ctx->fast_lstm = on == 1;
    ctx->strict_tc_demux = on == 2;
    if (demux_margin_delta > 0) ctx->demux_margin_delta = demux_margin_delta;
    if (demux_probe_gain > 0) ctx->demux_probe_gain = demux_probe_gain;

    return PB2_OK;
}

int pb2_demux_predict_tc(pb2_context *ctx, const float *windows, int64_t n, float *class_probs,
                         float *logits, int32_t *barcode, int32_t *guess, int32_t *score,
                         int32_t *unsafe, float *sensitivity, void *stream)
{
    if (!ctx) return PB2_EINVAL;
    if (!ctx->demux.set) return fail(ctx, PB2_ESTATE, "demux not set");
    DeviceGuard g(ctx->device);
    return launch_demux_tc(ctx, windows, nullptr, n, nullptr, nullptr, class_probs, barcode, guess,
                           score, logits, unsafe, sensitivity, /*recheck=*/false, (cudaStream_t)stream);
}

int pb2_debug_demux_l1(pb2_context *ctx, const float *windows, int64_t n, float *out, void *stream)
{
    if (!ctx || !windows || !out) return PB2_EINVAL;
    if (!ctx->demux.set) return fail(ctx, PB2_ESTATE, "demux not set");
    DeviceGuard g(ctx->device);
    return debug_demux_l1(ctx, windows, n, out, (cudaStream_t)stream);
}

int pb2_set_audit_fraction(pb2_context *ctx, double fraction)
{
    if (!ctx || !(fraction >= 0.0) || fraction > 1.0) return PB2_EINVAL;
    ctx->audit_threshold = fraction >= 1.0 ? 0xFFFFFFFFu : (uint32_t)(fraction * 4294967296.0);
    return PB2_OK;
}

int pb2_audit_stats(pb2_context *ctx, int64_t *audited, int64_t *mismatched)
{
    if (!ctx) return PB2_EINVAL;
    DeviceGuard g(ctx->device);
    int v[2] = {0, 0};
    PB_CUDA(ctx, cudaDeviceSynchronize());
    PB_CUDA(ctx, cudaMemcpy(v, ctx->tc_err + 2, sizeof(v), cudaMemcpyDeviceToHost));
    PB_CUDA(ctx, cudaMemset(ctx->tc_err + 2, 0, sizeof(v)));
    if (audited) *audited = v[0];
    if (mismatched) *mismatched = v[1];
    return PB2_OK;
}

int pb2_probe2_rows(pb2_context *ctx, int64_t *rows)
{
    if (!ctx || !rows) return PB2_EINVAL;
    DeviceGuard g(ctx->device);
    int v = 0;
    if (ctx->probe2_count_dev) {
        PB_CUDA(ctx, cudaDeviceSynchronize());
        PB_CUDA(ctx, cudaMemcpy(&v, ctx->probe2_count_dev, sizeof(int), cudaMemcpyDeviceToHost));
    }
    *rows = v;
    return PB2_OK;
}

int pb2_rerun_causes(pb2_context *ctx, int64_t *qc, int64_t *segmentation, int64_t *barcode)
{
    if (!ctx) return PB2_EINVAL;
    if (qc) *qc = ctx->last_rerun_cause[0];
    if (segmentation) *segmentation = ctx->last_rerun_cause[1];
    if (barcode) *barcode = ctx->last_rerun_cause[2];
    return PB2_OK;
}

int pb2_recheck_stats(pb2_context *ctx, int64_t *demux_rechecked, int64_t *tc_timeouts)
{
    if (!ctx) return PB2_EINVAL;
    DeviceGuard g(ctx->device);
    int32_t v[2] = {0, 0};
    int terr = 0;
    PB_CUDA(ctx, cudaDeviceSynchronize());
    if (ctx->ws_recheck.ptr && ctx->demux_tc_ran)
        PB_CUDA(ctx, cudaMemcpy(v, ctx->ws_recheck.ptr, sizeof(v), cudaMemcpyDeviceToHost));
    PB_CUDA(ctx, cudaMemcpy(&terr, ctx->tc_err, sizeof(int), cudaMemcpyDeviceToHost));
    PB_CUDA(ctx, cudaMemset(ctx->tc_err, 0, sizeof(int)));
    if (tc_timeouts) *tc_timeouts = terr;
    if (demux_rechecked) *demux_rechecked = ctx->last_rerun_reads > 0 ? ctx->last_rerun_reads : v[1];
    return PB2_OK;
}

int pb2_profile_kernel_count(void) { return K_NUM; }

const char *pb2_profile_kernel_name(int id) { return (id >= 0 && id < K_NUM) ? kKernelNames[id] : ""; }

int pb2_profile_read(pb2_context *ctx, double *total_ms, int64_t *launches, int n_kernels)
{
    if (!ctx || !total_ms || !launches) return PB2_EINVAL;
    DeviceGuard g(ctx->device);
    PB_CUDA(ctx, cudaDeviceSynchronize());
    for (int i = 0; i < n_kernels; i++) { total_ms[i] = 0; launches[i] = 0; }
    for (const ProfEvent &pe : ctx->prof_events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess && pe.id < n_kernels) {
            total_ms[pe.id] += ms;
            launches[pe.id] += 1;
        }
        ctx->prof_pool.push_back(pe.a);
        ctx->prof_pool.push_back(pe.b);
    }
    ctx->prof_events.clear();
    cudaGetLastError();
    return PB2_OK;
}

int pb2_profile_timeline(pb2_context *ctx, int32_t *ids, double *start_ms, double *end_ms,
                         int64_t capacity, int64_t *n_out)
{
    if (!ctx || !ids || !start_ms || !end_ms || !n_out) return PB2_EINVAL;
    DeviceGuard g(ctx->device);
    PB_CUDA(ctx, cudaDeviceSynchronize());
    int64_t k = 0;
    const cudaEvent_t t0 = ctx->prof_events.empty() ? nullptr : ctx->prof_events.front().a;
    for (const ProfEvent &pe : ctx->prof_events) {
        float a = 0.f, b = 0.f;
        if (k < capacity && cudaEventElapsedTime(&a, t0, pe.a) == cudaSuccess &&
            cudaEventElapsedTime(&b, t0, pe.b) == cudaSuccess) {
            ids[k] = pe.id; start_ms[k] = a; end_ms[k] = b;
            k++;
        }
        ctx->prof_pool.push_back(pe.a);
        ctx->prof_pool.push_back(pe.b);
    }
    ctx->prof_events.clear();
    cudaGetLastError();
    *n_out = k;
    return PB2_OK;
}

int pb2_set_scaler(pb2_context *ctx, const pb2_scaler_params *p)
{
    if (!ctx || !p) return PB2_EINVAL;
    DeviceGuard g(ctx->device);
    ScalerDev &S = ctx->scaler;
    S.set = false;
    if (p->stride <= 0 || p->length <= 0 || p->length % p->stride != 0)
        return fail(ctx, PB2_EINVAL, "scaler length must be a positive multiple of stride");
    int rc;
    if ((rc = upload_lstm(ctx, p->l1, S.l1))) return rc;
    if ((rc = upload_lstm(ctx, p->l2, S.l2))) return rc;
    if ((rc = upload(ctx, p->dense_kernel, (size_t)p->l2.units * 2, &S.dense_kernel))) return rc;
    if ((rc = upload(ctx, p->dense_bias, 2, &S.dense_bias))) return rc;
    S.stride = p->stride; S.length = p->length; S.min_length = p->min_length;
    S.scale_std = p->scale_std; S.scale_mean = p->scale_mean;
    S.shift_std = p->shift_std; S.shift_mean = p->shift_mean;
    S.qc_scale_lo = p->qc_scale_lo; S.qc_scale_hi = p->qc_scale_hi;
    S.qc_shift_lo = p->qc_shift_lo; S.qc_shift_hi = p->qc_shift_hi;
    if ((rc = build_zero_prefix(ctx))) return rc;
    S.set = true;
    return PB2_OK;
}

int pb2_set_segmentation_hmm(pb2_context *ctx, const pb2_hmm_params *p,
                             int32_t scan_limit_pooled, int32_t adapter_state)
{
    if (!ctx || !p) return PB2_EINVAL;
    if (p->n_states <= 0 || p->n_states > PB2_MAX_STATES - 1)
        return fail(ctx, PB2_EINVAL, "HMM must have 1..%d states", PB2_MAX_STATES - 1);
    if (adapter_state < 0 || adapter_state >= p->n_states || scan_limit_pooled <= 0)
        return fail(ctx, PB2_EINVAL, "bad adapter state / scan limit");
    for (int s = 0; s < p->n_states; s++)
        if (p->n_comp[s] < 1 || p->n_comp[s] > PB2_MAX_COMP)
            return fail(ctx, PB2_EINVAL, "bad mixture size for state %d", s);
    if (p->in_begin[p->n_states] > PB2_MAX_EDGES)
        return fail(ctx, PB2_EINVAL, "too many HMM edges");
    static_assert(sizeof(HmmDev) == sizeof(pb2_hmm_params), "HmmDev must mirror pb2_hmm_params");
    memcpy(&ctx->seg_hmm, p, sizeof(HmmDev));
    ctx->scan_limit_pooled = scan_limit_pooled;
    ctx->adapter_state = adapter_state;
    ctx->seg_set = true;
    return PB2_OK;
}

int pb2_set_demux(pb2_context *ctx, const pb2_demux_params *p)
{
    if (!ctx || !p) return PB2_EINVAL;
    DeviceGuard g(ctx->device);
    DemuxDev &D = ctx->demux;
    D.set = false;
    if (p->n_classes < 1 || p->n_classes > PB2_MAX_CLASSES || p->n_calibration < 1 ||
        p->n_calibration > PB2_MAX_CALIB || !p->calibration)
        return fail(ctx, PB2_EINVAL, "bad demux class / calibration sizes");
    if (p->trim_length < 1 || p->trim_length > PB2_WINDOW_MAX)
        return fail(ctx, PB2_EINVAL, "signal_trim_length out of range");
    // the count tensor has one slot per barcode (io.py:269-278): more barcode classes than
    // slots would be folded into "undetermined" silently
    if (p->n_decoy < 0 || p->n_classes - p->n_decoy > PB2_N_BARCODE_SLOTS - 1)
        return fail(ctx, PB2_EUNSUPPORTED, "%d barcode classes, the count tensor holds %d",
                    p->n_classes - p->n_decoy, PB2_N_BARCODE_SLOTS - 1);
    int rc;
    if ((rc = upload_lstm(ctx, p->fwd, D.fwd))) return rc;
    if ((rc = upload_lstm(ctx, p->bwd, D.bwd))) return rc;
    if ((rc = upload_lstm(ctx, p->l2, D.l2))) return rc;
    if ((rc = upload(ctx, p->dense_kernel, (size_t)p->l2.units * p->n_classes, &D.dense_kernel))) return rc;
    if ((rc = upload(ctx, p->dense_bias, p->n_classes, &D.dense_bias))) return rc;
    D.n_classes = p->n_classes; D.n_decoy = p->n_decoy;
    D.min_length = p->min_length; D.max_length = p->max_length;
    D.trim_length = p->trim_length; D.pad_value = p->pad_value;
    D.n_calibration = p->n_calibration;
    for (int i = 0; i < PB2_MAX_CALIB; i++)
        D.calibration[i] = i < p->n_calibration ? p->calibration[i] : INFINITY;
    D.score_threshold = p->score_threshold;
    if (!D.calibration_dev) PB_CUDA(ctx, cudaMalloc(&D.calibration_dev, sizeof(double) * PB2_MAX_CALIB));
    PB_CUDA(ctx, cudaMemcpy(D.calibration_dev, D.calibration, sizeof(double) * PB2_MAX_CALIB,
                            cudaMemcpyHostToDevice));
    if ((rc = build_pad_tables(ctx))) return rc;
    D.set = true;
    return PB2_OK;
}

int pb2_set_polya(pb2_context *ctx, const pb2_polya_params *p, int32_t polya_state)
{
    if (!ctx || !p) return PB2_EINVAL;
    if (p->stride <= 0 || p->window_length1 < 2 || p->window_length2 < p->window_length1 ||
        p->window_length2 > 30)
        return fail(ctx, PB2_EUNSUPPORTED, "event-detection windows must satisfy 2 <= w1 <= w2 <= 30");
    ctx->polya = *p;
    ctx->polya_state = polya_state;
    ctx->polya_set = true;
    return PB2_OK;
}

int pb2_set_unsplit(pb2_context *ctx, const pb2_hmm_params *hmm, const pb2_unsplit_params *p,
                    int32_t adapter_state, int32_t leader_high_state, int32_t leader_low_state)
{
    if (!ctx || !hmm || !p) return PB2_EINVAL;
    if (hmm->n_states <= 0 || hmm->n_states > PB2_MAX_STATES - 1)
        return fail(ctx, PB2_EINVAL, "HMM must have 1..%d states", PB2_MAX_STATES - 1);
    memcpy(&ctx->unsplit_hmm, hmm, sizeof(HmmDev));
    ctx->unsplit = *p;
    ctx->unsplit_states[0] = adapter_state;
    ctx->unsplit_states[1] = leader_high_state;
    ctx->unsplit_states[2] = leader_low_state;
    ctx->unsplit_set = true;
    return PB2_OK;
}

// ---- single stages ----------------------------------------------------------
static int check_batch(pb2_context *ctx, const pb2_batch *b)
{
    if (!ctx || !b) return PB2_EINVAL;
    if (b->n_reads < 0) return fail(ctx, PB2_EINVAL, "negative read count");
    if (b->n_reads > 0 && ((!b->raw && !b->packed) || (b->packed && !b->packed_offsets) ||
                           !b->raw_offsets || !b->raw_lengths || !b->range ||
                           !b->digitisation || !b->offset))
        return fail(ctx, PB2_EINVAL, "null batch pointer");
    return PB2_OK;
}

int pb2_pool_signal(pb2_context *ctx, const pb2_batch *batch, float *pooled, void *stream)
{
    int rc = check_batch(ctx, batch);
    if (rc) return rc;
    if (!ctx->scaler.set || !ctx->seg_set) return fail(ctx, PB2_ESTATE, "parameters not set");
    DeviceGuard g(ctx->device);
    return launch_pool(ctx, *batch, ctx->scaler.stride, pooled, (cudaStream_t)stream);
}

int pb2_fit_scalers(pb2_context *ctx, const pb2_batch *batch, const float *pooled,
                    int32_t *status, float *scale_shift, float *z_out, void *stream)
{
    int rc = check_batch(ctx, batch);
    if (rc) return rc;
    if (!ctx->scaler.set) return fail(ctx, PB2_ESTATE, "scaler not set");
    DeviceGuard g(ctx->device);
    return launch_scaler(ctx, *batch, pooled, status, scale_shift, z_out, (cudaStream_t)stream);
}

int pb2_scaler_predict(pb2_context *ctx, const float *heads, int64_t n, float *z_out,
                       void *stream)
{
    if (!ctx) return PB2_EINVAL;
    if (!ctx->scaler.set) return fail(ctx, PB2_ESTATE, "scaler not set");
    DeviceGuard g(ctx->device);
    return launch_scaler_heads(ctx, heads, n, z_out, (cudaStream_t)stream);
}

int pb2_detect_segments(pb2_context *ctx, const pb2_batch *batch, const float *pooled,
                        const float *scale_shift, int32_t *status, int32_t *segments,
                        float *pooled_scaled_out, void *stream)
{
    int rc = check_batch(ctx, batch);
    if (rc) return rc;
    if (!ctx->seg_set || !ctx->scaler.set) return fail(ctx, PB2_ESTATE, "parameters not set");
    DeviceGuard g(ctx->device);
    return launch_segment(ctx, *batch, pooled, scale_shift, status, segments,
                          pooled_scaled_out, (cudaStream_t)stream);
}

int pb2_viterbi_paths(pb2_context *ctx, int which, const float *x, const int32_t *lengths,
                      int64_t n, int32_t ld, int32_t *path, double *logp, void *stream)
{
    if (!ctx) return PB2_EINVAL;
    if (which != 0) return fail(ctx, PB2_EUNSUPPORTED, "only the segmentation model is loaded");
    if (!ctx->seg_set) return fail(ctx, PB2_ESTATE, "segmentation HMM not set");
    DeviceGuard g(ctx->device);
    return launch_viterbi_paths(ctx, ctx->seg_hmm, x, lengths, n, ld, path, logp,
                                (cudaStream_t)stream);
}

int pb2_barcode_windows(pb2_context *ctx, const pb2_batch *batch, const float *pooled,
                        const float *scale_shift, const int32_t *status,
                        const int32_t *segments, float *windows, int32_t *pushed, void *stream)
{
    int rc = check_batch(ctx, batch);
    if (rc) return rc;
    if (!ctx->demux.set || !ctx->scaler.set || !ctx->seg_set)
        return fail(ctx, PB2_ESTATE, "parameters not set");
    DeviceGuard g(ctx->device);
    return launch_windows(ctx, *batch, pooled, scale_shift, status, segments, windows, pushed,
                          nullptr, nullptr, (cudaStream_t)stream);
}

int pb2_demux_predict(pb2_context *ctx, const float *windows, const int32_t *pushed, int64_t n,
                      float *class_probs, int32_t *barcode, int32_t *guess, int32_t *score,
                      void *stream)
{
    if (!ctx) return PB2_EINVAL;
    if (!ctx->demux.set) return fail(ctx, PB2_ESTATE, "demux not set");
    DeviceGuard g(ctx->device);
    return launch_demux(ctx, windows, pushed, n, nullptr, nullptr, class_probs, barcode, guess,
                        score, (cudaStream_t)stream);
}

int pb2_measure_polya(pb2_context *ctx, const pb2_batch *batch, const float *scale_shift,
                      const int32_t *status, const int32_t *segments, pb2_polya_result *out,
                      void *stream)
{
    int rc = check_batch(ctx, batch);
    if (rc) return rc;
    if (!ctx->polya_set || !ctx->seg_set) return fail(ctx, PB2_ESTATE, "poly(A) parameters not set");
    DeviceGuard g(ctx->device);
    return launch_polya(ctx, *batch, scale_shift, status, segments, out, (cudaStream_t)stream);
}

int pb2_detect_unsplit(pb2_context *ctx, const pb2_batch *batch, const pb2_event_tables *events,
                       int64_t n_reads, const float *scale_shift, const int32_t *status,
                       const int32_t *segments, int32_t max_windows, int32_t *flag, void *stream)
{
    if (!ctx || !events || !flag) return PB2_EINVAL;
    if (!ctx->unsplit_set || !ctx->seg_set) return fail(ctx, PB2_ESTATE, "unsplit-read model not set");
    DeviceGuard g(ctx->device);
    return launch_unsplit(ctx, batch, *events, n_reads, scale_shift, status, segments, max_windows,
                          flag, (cudaStream_t)stream);
}

int pb2_detect_unsplit_host(pb2_context *ctx, const pb2_batch *hb, const pb2_event_tables *hev,
                            int64_t n, const float *scale_shift, const int32_t *status,
                            const int32_t *segments, int32_t max_windows, int32_t *flag)
{
    if (!ctx || !hev || !flag) return PB2_EINVAL;
    if (!ctx->unsplit_set || !ctx->seg_set) return fail(ctx, PB2_ESTATE, "unsplit-read model not set");
    if (n <= 0) return PB2_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->host_stream;
    const size_t E = (size_t)hev->n_events_total;
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align(bytes + 16); return o; };
    const size_t o_eo = take(sizeof(int64_t) * (n + 1)), o_st = take(sizeof(int64_t) * E);
    const size_t o_mn = take(sizeof(float) * E), o_mv = take(sizeof(int32_t) * E);
    const size_t o_ps = take(sizeof(double) * E), o_rt = take(sizeof(double) * n);
    const size_t o_ss = take(sizeof(float) * 2 * n), o_stat = take(sizeof(int32_t) * n);
    const size_t o_seg = take(sizeof(int32_t) * 2 * PB2_MAX_STATES * n), o_fl = take(sizeof(int32_t) * n);
    const bool derive = hev->mean == nullptr;
    if (derive && (!hb || hb->n_reads != n || !hev->first_sample))
        return fail(ctx, PB2_EINVAL, "deriving event means needs the batch of the same reads");
    const size_t R = derive ? (size_t)hb->n_raw_total : 0;
    const size_t o_raw = take(sizeof(int16_t) * R), o_ro = take(8 * (size_t)n), o_rl = take(8 * (size_t)n);
    const size_t o_rg = take(8 * (size_t)n), o_dg = take(8 * (size_t)n), o_of = take(8 * (size_t)n);
    const size_t o_fs = take(8 * (size_t)n);
    char *base = (char *)ws_get(ctx, ctx->ws_unsplit_host, off);
    if (!base) return PB2_ENOMEM;
#define PB_H2D(o, src, bytes) PB_CUDA(ctx, cudaMemcpyAsync(base + (o), (src), (bytes), cudaMemcpyHostToDevice, st))
    PB_H2D(o_eo, hev->event_offsets, sizeof(int64_t) * (n + 1));
    if (derive) {
        PB_H2D(o_raw, hb->raw, sizeof(int16_t) * R);
        PB_H2D(o_ro, hb->raw_offsets, 8 * (size_t)n);
        PB_H2D(o_rl, hb->raw_lengths, 8 * (size_t)n);
        PB_H2D(o_rg, hb->range, 8 * (size_t)n);
        PB_H2D(o_dg, hb->digitisation, 8 * (size_t)n);
        PB_H2D(o_of, hb->offset, 8 * (size_t)n);
        PB_H2D(o_fs, hev->first_sample, 8 * (size_t)n);
    }
    if (E) {
        PB_H2D(o_st, hev->start, sizeof(int64_t) * E);
        if (!derive) PB_H2D(o_mn, hev->mean, sizeof(float) * E);
        PB_H2D(o_mv, hev->move, sizeof(int32_t) * E);
        PB_H2D(o_ps, hev->p_model_state, sizeof(double) * E);
    }
    PB_H2D(o_rt, hev->sampling_rate, sizeof(double) * n);
    PB_H2D(o_ss, scale_shift, sizeof(float) * 2 * n);
    PB_H2D(o_stat, status, sizeof(int32_t) * n);
    PB_H2D(o_seg, segments, sizeof(int32_t) * 2 * PB2_MAX_STATES * n);
#undef PB_H2D
    pb2_event_tables dev = *hev;
    dev.event_offsets = (const int64_t *)(base + o_eo);
    dev.start = (const int64_t *)(base + o_st);
    dev.mean = derive ? nullptr : (const float *)(base + o_mn);
    dev.first_sample = (const int64_t *)(base + o_fs);
    pb2_batch db = {};
    if (derive) {
        db = *hb;
        db.raw = (const int16_t *)(base + o_raw);
        db.raw_offsets = (const int64_t *)(base + o_ro);
        db.raw_lengths = (const int64_t *)(base + o_rl);
        db.range = (const double *)(base + o_rg);
        db.digitisation = (const double *)(base + o_dg);
        db.offset = (const double *)(base + o_of);
    }
    dev.move = (const int32_t *)(base + o_mv);
    dev.p_model_state = (const double *)(base + o_ps);
    dev.sampling_rate = (const double *)(base + o_rt);
    int rc = launch_unsplit(ctx, derive ? &db : nullptr, dev, n, (const float *)(base + o_ss), (const int32_t *)(base + o_stat),
                            (const int32_t *)(base + o_seg), max_windows, (int32_t *)(base + o_fl), st);
    if (rc) return rc;
    PB_CUDA(ctx, cudaMemcpyAsync(flag, base + o_fl, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    PB_CUDA(ctx, cudaStreamSynchronize(st));
    return PB2_OK;
}

int pb2_derive_event_tables(pb2_context *ctx, const pb2_batch *batch, const pb2_event_tables *events,
                            const pb2_basecalls *basecalls, const float *scale_shift,
                            const pb2_event_columns *out, void *stream)
{
    int rc = check_batch(ctx, batch);
    if (rc) return rc;
    if (!events || !out) return PB2_EINVAL;
    DeviceGuard g(ctx->device);
    return launch_derive_events(ctx, *batch, *events, basecalls, scale_shift, *out, (cudaStream_t)stream);
}

int pb2_derive_event_tables_host(pb2_context *ctx, const pb2_batch *hb, const pb2_event_tables *hev,
                                 const pb2_basecalls *hbc, const float *scale_shift,
                                 const pb2_event_columns *hout)
{
    int rc = check_batch(ctx, hb);
    if (rc) return rc;
    if (!hev || !hout || !hev->event_offsets || !hev->first_sample) return PB2_EINVAL;
    const int64_t n = hb->n_reads;
    if (n <= 0) return PB2_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->host_stream;
    const size_t E = (size_t)hev->n_events_total;
    const size_t S = hbc ? (size_t)hbc->seq_offsets[n] : 0;
    const size_t R = (size_t)hb->n_raw_total;
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align(bytes + 16); return o; };
    const size_t o_raw = take(2 * R), o_ro = take(8 * n), o_rl = take(8 * n), o_rg = take(8 * n);
    const size_t o_dg = take(8 * n), o_of = take(8 * n), o_eo = take(8 * (n + 1)), o_fs = take(8 * n);
    const size_t o_mv = take(4 * E), o_sq = take(S), o_qs = take(S), o_so = take(8 * (n + 1));
    const size_t o_qt = take(8 * 256), o_ss = take(8 * n);
    const size_t o_mean = take(4 * E), o_stdv = take(4 * E), o_sm = take(4 * E), o_start = take(8 * E);
    const size_t o_end = take(8 * E), o_len = take(8 * E), o_pos = take(8 * E), o_pm = take(8 * E);
    const size_t o_ms = take(5 * E), o_err = take(4 * n);
    char *base = (char *)ws_get(ctx, ctx->ws_unsplit_host, off);
    if (!base) return PB2_ENOMEM;
#define PB_H2D(o, src, bytes) if ((src) && (bytes)) PB_CUDA(ctx, cudaMemcpyAsync(base + (o), (src), (bytes), cudaMemcpyHostToDevice, st))
    PB_H2D(o_raw, hb->raw, 2 * R); PB_H2D(o_ro, hb->raw_offsets, 8 * n); PB_H2D(o_rl, hb->raw_lengths, 8 * n);
    PB_H2D(o_rg, hb->range, 8 * n); PB_H2D(o_dg, hb->digitisation, 8 * n); PB_H2D(o_of, hb->offset, 8 * n);
    PB_H2D(o_eo, hev->event_offsets, 8 * (n + 1)); PB_H2D(o_fs, hev->first_sample, 8 * n);
    PB_H2D(o_mv, hev->move, 4 * E);
    if (hbc) {
        PB_H2D(o_sq, hbc->sequence, S); PB_H2D(o_qs, hbc->qstring, S);
        PB_H2D(o_so, hbc->seq_offsets, 8 * (n + 1)); PB_H2D(o_qt, hbc->qual_table, 8 * 256);
    }
    PB_H2D(o_ss, scale_shift, 8 * n);
#undef PB_H2D
    pb2_batch db = *hb;
    db.raw = (const int16_t *)(base + o_raw); db.raw_offsets = (const int64_t *)(base + o_ro);
    db.raw_lengths = (const int64_t *)(base + o_rl); db.range = (const double *)(base + o_rg);
    db.digitisation = (const double *)(base + o_dg); db.offset = (const double *)(base + o_of);
    pb2_event_tables dev = *hev;
    dev.event_offsets = (const int64_t *)(base + o_eo); dev.first_sample = (const int64_t *)(base + o_fs);
    dev.move = hev->move ? (const int32_t *)(base + o_mv) : nullptr;
    pb2_basecalls dbc = {};
    if (hbc) {
        dbc.sequence = hbc->sequence ? (const uint8_t *)(base + o_sq) : nullptr;
        dbc.qstring = hbc->qstring ? (const uint8_t *)(base + o_qs) : nullptr;
        dbc.seq_offsets = (const int64_t *)(base + o_so);
        dbc.qual_table = hbc->qual_table ? (const double *)(base + o_qt) : nullptr;
    }
    pb2_event_columns d = {};
    d.mean = hout->mean ? (float *)(base + o_mean) : nullptr;
    d.stdv = hout->stdv ? (float *)(base + o_stdv) : nullptr;
    d.scaled_mean = hout->scaled_mean ? (float *)(base + o_sm) : nullptr;
    d.start = hout->start ? (int64_t *)(base + o_start) : nullptr;
    d.end = hout->end ? (int64_t *)(base + o_end) : nullptr;
    d.length = hout->length ? (int64_t *)(base + o_len) : nullptr;
    // pos is needed on the device by the columns that hang off it
    d.pos = (hout->pos || hout->p_model_state || hout->model_state) ? (int64_t *)(base + o_pos) : nullptr;
    d.p_model_state = hout->p_model_state ? (double *)(base + o_pm) : nullptr;
    d.model_state = hout->model_state ? (uint8_t *)(base + o_ms) : nullptr;
    d.error = hout->error ? (int32_t *)(base + o_err) : nullptr;
    if (d.model_state) PB_CUDA(ctx, cudaMemsetAsync(d.model_state, 0, 5 * E + 1, st));
    rc = launch_derive_events(ctx, db, dev, hbc ? &dbc : nullptr, scale_shift ? (const float *)(base + o_ss) : nullptr, d, st);
    if (rc) return rc;
#define PB_D2H(field, bytes) if (hout->field && (bytes)) PB_CUDA(ctx, cudaMemcpyAsync(hout->field, d.field, (bytes), cudaMemcpyDeviceToHost, st))
    PB_D2H(mean, 4 * E); PB_D2H(stdv, 4 * E); PB_D2H(scaled_mean, 4 * E); PB_D2H(start, 8 * E);
    PB_D2H(end, 8 * E); PB_D2H(length, 8 * E); PB_D2H(pos, 8 * E); PB_D2H(p_model_state, 8 * E);
    PB_D2H(model_state, 5 * E); PB_D2H(error, 4 * n);
#undef PB_D2H
    PB_CUDA(ctx, cudaStreamSynchronize(st));
    return PB2_OK;
}

int pb2_detect_events(pb2_context *ctx, const float *signal, const int64_t *offsets,
                      const int64_t *lengths, int64_t n_signals, const pb2_detector_params *p,
                      int64_t *event_counts, const int64_t *event_offsets, void *records,
                      void *stream)
{
    if (!ctx) return PB2_EINVAL;
    if (n_signals < 0 || !p || (n_signals > 0 && (!signal || !offsets || !lengths)))
        return fail(ctx, PB2_EINVAL, "detect_events: bad arguments");
    DeviceGuard g(ctx->device);
    return launch_detect_events(ctx, signal, offsets, lengths, n_signals, *p, event_counts,
                                event_offsets, records, (cudaStream_t)stream);
}

int pb2_svb16_decode(pb2_context *ctx, const uint8_t *packed, const int64_t *packed_offsets,
                     const int64_t *raw_offsets, const int64_t *raw_lengths, int64_t n_reads,
                     int16_t *raw, int32_t *error, void *stream)
{
    if (!ctx) return PB2_EINVAL;
    if (n_reads > 0 && (!packed || !packed_offsets || !raw_offsets || !raw_lengths || !raw))
        return fail(ctx, PB2_EINVAL, "svb16_decode: null pointer");
    DeviceGuard g(ctx->device);
    return launch_svb16_decode(ctx, packed, packed_offsets, raw_offsets, raw_lengths, n_reads, raw,
                               error, (cudaStream_t)stream);
}

int pb2_count_results(pb2_context *ctx, const int32_t *status, const int32_t *label,
                      const int32_t *barcode, int64_t n, int64_t *counts, void *stream)
{
    if (!ctx || !counts) return PB2_EINVAL;
    DeviceGuard g(ctx->device);
    return launch_counts(ctx, status, label, barcode, n, counts, (cudaStream_t)stream);
}

// ---- whole path -------------------------------------------------------------
#define WS_OR(user, ws, type, count)                                                        \
    ((user) ? (user) : (type *)ws_get(ctx, ctx->ws, sizeof(type) * (size_t)(count)))


// ---- whole path, tensor-core variant ---------------------------------------------------
// Approximate scaler and demultiplexer on the tensor cores; every read whose integer outputs
// could differ from the exact kernels' (QC verdict on an edge, segmentation not constant over
// the (scale, shift) uncertainty triangle, barcode call inside its error margin) is collected
// and re-run as a sub-batch through the exact kernels -- scaler, segmentation, window,
// demultiplexer -- and its results replace the tentative ones.  One host synchronisation (the
// size of that sub-batch).
// audit_threshold: a pseudo-random fraction audit_threshold / 2^32 of the reads that passed every
// guard is re-run as well (cause bit 8), purely to compare: see k_scatter_sub_results.
__global__ void k_collect_unsafe(int64_t n, int32_t *__restrict__ unsafe, int *count,
                                 int32_t *__restrict__ list, uint32_t audit_threshold, uint32_t seed)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (!unsafe[r] && audit_threshold) {
        uint32_t h = (uint32_t)r * 2654435761u + seed;          // integer hash of the read index
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        if (h < audit_threshold) unsafe[r] = 8;
    }
    if (!unsafe[r]) return;
    list[atomicAdd(count, 1)] = (int32_t)r;
    if (unsafe[r] & 1) atomicAdd(count + 1, 1);       // per-cause tallies (diagnostics)
    if (unsafe[r] & 2) atomicAdd(count + 2, 1);
    if (unsafe[r] & 4) atomicAdd(count + 3, 1);
}

__global__ void k_gather_sub_batch(int n_sub, const int32_t *__restrict__ list,
                                   const int64_t *__restrict__ ro, const int64_t *__restrict__ rl,
                                   const double *__restrict__ rg, const double *__restrict__ dg,
                                   const double *__restrict__ of, int64_t *ro2, int64_t *rl2,
                                   double *rg2, double *dg2, double *of2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sub) return;
    const int32_t r = list[i];
    ro2[i] = ro[r]; rl2[i] = rl[r]; rg2[i] = rg[r]; dg2[i] = dg[r]; of2[i] = of[r];
}

// A read that was re-run only because of its QC verdict or its segmentation keeps its
// tensor-core barcode call when the exact status and segments equal the tentative ones: the
// window then differs from the one that call was made on by the rounding of (scale, shift)
// only (~1e-7), far inside the margin the call passed.  status_win: status as the window
// kernel should see it (not OKAY = no window).
__global__ void k_sub_needs_demux(int n_sub, const int32_t *__restrict__ list,
                                  const int32_t *__restrict__ unsafe, const int32_t *__restrict__ status,
                                  const int32_t *__restrict__ seg, const int32_t *__restrict__ status2,
                                  const int32_t *__restrict__ seg2, int32_t *__restrict__ status_win,
                                  int32_t *__restrict__ need)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sub) return;
    const int64_t r = list[i];
    bool nd = (unsafe[r] & (4 | 8)) != 0 || status[r] != status2[i];
    for (int k = 0; k < PB2_MAX_STATES * 2; k++)
        nd = nd || seg[r * PB2_MAX_STATES * 2 + k] != seg2[(int64_t)i * PB2_MAX_STATES * 2 + k];
    need[i] = nd ? 1 : 0;
    status_win[i] = nd ? status2[i] : PB2_ST_UNKNOWN_ERROR;
}

__global__ void k_scatter_sub_results(int n_sub, const int32_t *__restrict__ list,
                                      const int32_t *__restrict__ status2, const float *__restrict__ ss2,
                                      const int32_t *__restrict__ seg2, const int32_t *__restrict__ pushed2,
                                      const int32_t *__restrict__ bc2, const int32_t *__restrict__ gs2,
                                      const int32_t *__restrict__ sc2, const float *__restrict__ pr2,
                                      const int32_t *__restrict__ need,
                                      const int32_t *__restrict__ unsafe, int *audit,
                                      int32_t *status, float *ss, int32_t *seg, int32_t *pushed,
                                      int32_t *bc, int32_t *gs, int32_t *sc, float *pr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sub) return;
    const int64_t r = list[i];
    if (unsafe[r] == 8) {
        // audited read: it passed every guard, so the exact results must equal the tentative ones
        bool same = status[r] == status2[i];
        for (int k = 0; k < PB2_MAX_STATES * 2; k++)
            same = same && seg[r * PB2_MAX_STATES * 2 + k] == seg2[(int64_t)i * PB2_MAX_STATES * 2 + k];
        if (pushed)
            same = same && pushed[r] == pushed2[i] &&
                   (!pushed2[i] || (bc[r] == bc2[i] && gs[r] == gs2[i] && sc[r] == sc2[i]));
        atomicAdd(audit, 1);
        if (!same) atomicAdd(audit + 1, 1);
    }
    status[r] = status2[i];
    ss[2 * r] = ss2[2 * i]; ss[2 * r + 1] = ss2[2 * i + 1];
    for (int k = 0; k < PB2_MAX_STATES * 2; k++) seg[r * PB2_MAX_STATES * 2 + k] = seg2[(int64_t)i * PB2_MAX_STATES * 2 + k];
    if (pushed && need[i]) {
        pushed[r] = pushed2[i];
        bc[r] = bc2[i]; gs[r] = gs2[i]; sc[r] = sc2[i];
        if (pr) for (int k = 0; k < PB2_MAX_CLASSES; k++) pr[r * PB2_MAX_CLASSES + k] = pr2[(int64_t)i * PB2_MAX_CLASSES + k];
    }
}

// The fast path in two halves so that a host batch cut into chunks can run the tensor-core half
// chunk by chunk (as its uploads arrive) and resolve the unsafe reads of ALL chunks as one
// sub-batch at the end: the exact kernels have long CTAs, and a sub-batch per chunk ends on a
// partly filled wave every time.
//   analyze_fast_tentative: tensor-core scaler, corner segmentations, windows, tensor-core
//       classifier; tentative results in the result arrays, cause bits in `unsafe`.  No host
//       synchronisation.  `unsafe` must be zero on entry.
//   analyze_fast_resolve:   collect the flagged reads of `batch` (the whole batch), re-run them
//       exactly, scatter, then label + counts.  One host synchronisation.
struct FastArrays {
    int32_t *unsafe;               // [n] cause bits
    int32_t *pushed;               // [n] window accepted (barcoding)
    int32_t *barcode, *guess, *score;
};

static int analyze_fast_tentative(pb2_context *ctx, const pb2_batch *batch, const pb2_results *res,
                                  uint32_t flags, cudaStream_t st, float *pooled, int32_t *status,
                                  float *scale_shift, int32_t *segments, const FastArrays &fa)
{
    int rc;
    const int64_t n = batch->n_reads;
    const bool bcd = (flags & PB2_FLAG_BARCODING) != 0;
    const int T = bcd ? ctx->demux.trim_length : 0;
    const size_t SEG = (size_t)PB2_MAX_STATES * 2;
    // scratch: corner (scale, shift) x3, corner status x2 / segments x2 (corner 0 decodes
    // straight into the result arrays)
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += al(bytes); return o; };
    const size_t o_ssv = take(sizeof(float) * 6 * n), o_st1 = take(4 * (size_t)n), o_st2 = take(4 * (size_t)n);
    const size_t o_sg1 = take(4 * SEG * n), o_sg2 = take(4 * SEG * n);
    char *fs = (char *)ws_get(ctx, ctx->ws_fast, off);
    if (!fs) return PB2_ENOMEM;
    float *ssv = (float *)(fs + o_ssv);
    int32_t *st1 = (int32_t *)(fs + o_st1), *st2 = (int32_t *)(fs + o_st2);
    int32_t *sg1 = (int32_t *)(fs + o_sg1), *sg2 = (int32_t *)(fs + o_sg2);
    int32_t *unsafe = fa.unsafe;

    if ((rc = launch_scaler_tc(ctx, *batch, pooled, status, scale_shift, ssv, unsafe, nullptr, st))) return rc;
    // segmentation at the three corners of the uncertainty triangle
    PB_CUDA(ctx, cudaMemcpyAsync(st1, status, 4 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    PB_CUDA(ctx, cudaMemcpyAsync(st2, status, 4 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    {
        int32_t *const st3[3] = {status, st1, st2}, *const sg3[3] = {segments, sg1, sg2};
        if ((rc = launch_segment3(ctx, *batch, pooled, ssv, st3, sg3, st))) return rc;
    }
    if ((rc = launch_compare_corners(ctx, n, status, st1, st2, segments, sg1, sg2, unsafe, st))) return rc;

    if (bcd) {
        float *windows = (float *)ws_get(ctx, ctx->ws_win, sizeof(float) * (size_t)n * T);
        int32_t *slots = (int32_t *)ws_get(ctx, ctx->ws_slots, sizeof(int32_t) * ((size_t)n + 4));
        if (!windows || !slots) return PB2_ENOMEM;
        int *slot_count = (int *)slots;
        int32_t *slot_read = slots + 4;
        if (res->class_probs)
            PB_CUDA(ctx, cudaMemsetAsync(res->class_probs, 0, sizeof(float) * PB2_MAX_CLASSES * (size_t)n, st));
        if ((rc = launch_windows(ctx, *batch, pooled, scale_shift, status, segments, windows, fa.pushed,
                                 slot_count, slot_read, st))) return rc;
        if ((rc = launch_demux_tc(ctx, windows, nullptr, n, slot_count, slot_read, res->class_probs,
                                  fa.barcode, fa.guess, fa.score, nullptr, nullptr, nullptr,
                                  /*recheck=*/false, st, unsafe))) return rc;
    }
    return PB2_OK;
}

// count: device int[4] (zero on entry), list: device int32[n]
static int analyze_fast_resolve(pb2_context *ctx, const pb2_batch *batch, const pb2_results *res,
                                uint32_t flags, cudaStream_t st, float *pooled, int32_t *status,
                                int32_t *label, float *scale_shift, int32_t *segments,
                                const FastArrays &fa, int *count, int32_t *list)
{
    int rc;
    const int64_t n = batch->n_reads;
    const bool bcd = (flags & PB2_FLAG_BARCODING) != 0;
    const int T = bcd ? ctx->demux.trim_length : 0;
    const size_t SEG = (size_t)PB2_MAX_STATES * 2;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    int32_t *unsafe = fa.unsafe, *pushed = fa.pushed;
    int32_t *barcode = fa.barcode, *guess = fa.guess, *score = fa.score;

    // ---- the unsafe reads, exactly ---------------------------------------------------
    PB_LAUNCH(ctx, K_MISC, "k_collect_unsafe", st,
        k_collect_unsafe<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        n, unsafe, count, list, ctx->audit_threshold, (uint32_t)ctx->launches));
    int tally[4] = {0, 0, 0, 0};
    PB_CUDA(ctx, cudaMemcpyAsync(tally, count, sizeof(tally), cudaMemcpyDeviceToHost, st));
    PB_CUDA(ctx, cudaStreamSynchronize(st));
    const int n_sub = tally[0];
    ctx->last_rerun_reads = n_sub;
    for (int i = 0; i < 3; i++) ctx->last_rerun_cause[i] = tally[i + 1];
    if (n_sub > 0) {
        const size_t m = (size_t)n_sub;
        size_t o2 = 0;
        auto take2 = [&](size_t bytes) { size_t o = o2; o2 += al(bytes); return o; };
        const size_t q_ro = take2(8 * m), q_rl = take2(8 * m), q_rg = take2(8 * m), q_dg = take2(8 * m);
        const size_t q_of = take2(8 * m), q_st = take2(4 * m), q_ss = take2(8 * m), q_sg = take2(4 * SEG * m);
        const size_t q_pu = take2(4 * m), q_sl = take2(4 * (m + 4)), q_bc = take2(4 * m), q_gs = take2(4 * m);
        const size_t q_sc = take2(4 * m), q_pr = take2(4 * PB2_MAX_CLASSES * m);
        const size_t q_win = take2(bcd ? sizeof(float) * m * T : 16);
        const size_t q_nd = take2(4 * m), q_sw = take2(4 * m);
        char *sb = (char *)ws_get(ctx, ctx->ws_sub, o2);
        if (!sb) return PB2_ENOMEM;
        pb2_batch sub = *batch;
        sub.n_reads = n_sub;
        sub.raw_offsets = (const int64_t *)(sb + q_ro);
        sub.raw_lengths = (const int64_t *)(sb + q_rl);
        sub.range = (const double *)(sb + q_rg);
        sub.digitisation = (const double *)(sb + q_dg);
        sub.offset = (const double *)(sb + q_of);
        PB_LAUNCH(ctx, K_MISC, "k_gather_sub_batch", st,
            k_gather_sub_batch<<<(unsigned)((n_sub + 255) / 256), 256, 0, st>>>(
            n_sub, list, batch->raw_offsets, batch->raw_lengths, batch->range, batch->digitisation,
            batch->offset, (int64_t *)(sb + q_ro), (int64_t *)(sb + q_rl), (double *)(sb + q_rg),
            (double *)(sb + q_dg), (double *)(sb + q_of)));
        int32_t *status2 = (int32_t *)(sb + q_st), *seg2 = (int32_t *)(sb + q_sg);
        float *ss2 = (float *)(sb + q_ss);
        int32_t *pushed2 = (int32_t *)(sb + q_pu), *slots2 = (int32_t *)(sb + q_sl);
        int32_t *bc2 = (int32_t *)(sb + q_bc), *gs2 = (int32_t *)(sb + q_gs), *sc2 = (int32_t *)(sb + q_sc);
        float *pr2 = (float *)(sb + q_pr), *win2 = (float *)(sb + q_win);
        if ((rc = launch_scaler(ctx, sub, pooled, status2, ss2, nullptr, st))) return rc;
        if ((rc = launch_segment(ctx, sub, pooled, ss2, status2, seg2, nullptr, st))) return rc;
        int32_t *need = (int32_t *)(sb + q_nd), *status_win = (int32_t *)(sb + q_sw);
        PB_LAUNCH(ctx, K_MISC, "k_sub_needs_demux", st,
            k_sub_needs_demux<<<(unsigned)((n_sub + 255) / 256), 256, 0, st>>>(
            n_sub, list, unsafe, status, segments, status2, seg2, status_win, need));
        if (bcd) {
            PB_CUDA(ctx, cudaMemsetAsync(bc2, 0xFF, 4 * m, st));
            PB_CUDA(ctx, cudaMemsetAsync(gs2, 0xFF, 4 * m, st));
            PB_CUDA(ctx, cudaMemsetAsync(sc2, 0xFF, 4 * m, st));
            PB_CUDA(ctx, cudaMemsetAsync(pr2, 0, 4 * PB2_MAX_CLASSES * m, st));
            if ((rc = launch_windows(ctx, sub, pooled, ss2, status_win, seg2, win2, pushed2,
                                     (int *)slots2, slots2 + 4, st))) return rc;
            if ((rc = launch_demux_exact(ctx, win2, nullptr, n_sub, (int *)slots2, slots2 + 4, pr2, bc2,
                                         gs2, sc2, st))) return rc;
        }
        PB_LAUNCH(ctx, K_MISC, "k_scatter_sub_results", st,
            k_scatter_sub_results<<<(unsigned)((n_sub + 255) / 256), 256, 0, st>>>(
            n_sub, list, status2, ss2, seg2, pushed2, bc2, gs2, sc2, pr2, need, unsafe, ctx->tc_err + 2,
            status, scale_shift, segments,
            bcd ? pushed : nullptr, barcode, guess, score, res->class_probs));
    }
    if ((rc = launch_finalize(ctx, n, flags, status, label, barcode, guess, score, st, pushed))) return rc;
    if (res->counts)
        if ((rc = launch_counts(ctx, status, label, barcode, n, res->counts, st))) return rc;
    return PB2_OK;
}

// barcode / guess / score arrays of a fast-path call: the caller's, or scratch when barcoding
// runs and the caller did not ask for one of them
static int fast_arrays(pb2_context *ctx, const pb2_results *res, uint32_t flags, int64_t n,
                       int32_t *unsafe, FastArrays &fa)
{
    fa.unsafe = unsafe;
    fa.pushed = nullptr;
    fa.barcode = res->barcode; fa.guess = res->barcode_guess; fa.score = res->barcode_score;
    if (flags & PB2_FLAG_BARCODING) {
        fa.pushed = (int32_t *)ws_get(ctx, ctx->ws_pushed, sizeof(int32_t) * (size_t)n);
        fa.barcode = WS_OR(res->barcode, ws_bc, int32_t, n);
        fa.guess = WS_OR(res->barcode_guess, ws_guess, int32_t, n);
        fa.score = WS_OR(res->barcode_score, ws_score, int32_t, n);
        if (!fa.pushed || !fa.barcode || !fa.guess || !fa.score) return PB2_ENOMEM;
    }
    return PB2_OK;
}

static int analyze_device_fast(pb2_context *ctx, const pb2_batch *batch, const pb2_results *res,
                               uint32_t flags, cudaStream_t st, float *pooled, int32_t *status,
                               int32_t *label, float *scale_shift, int32_t *segments)
{
    int rc;
    const int64_t n = batch->n_reads;
    // unsafe flags, [4] counters, the unsafe list
    const size_t o_cnt = ((size_t)4 * n + 255) & ~(size_t)255, o_list = o_cnt + 256;
    char *fl = (char *)ws_get(ctx, ctx->ws_flags, o_list + 4 * (size_t)n);
    if (!fl) return PB2_ENOMEM;
    PB_CUDA(ctx, cudaMemsetAsync(fl, 0, o_list, st));
    FastArrays fa;
    if ((rc = fast_arrays(ctx, res, flags, n, (int32_t *)fl, fa))) return rc;
    if ((rc = analyze_fast_tentative(ctx, batch, res, flags, st, pooled, status, scale_shift, segments, fa)))
        return rc;
    return analyze_fast_resolve(ctx, batch, res, flags, st, pooled, status, label, scale_shift, segments,
                                fa, (int *)(fl + o_cnt), (int32_t *)(fl + o_list));
}

int pb2_analyze_device(pb2_context *ctx, const pb2_batch *batch, const pb2_results *res,
                       uint32_t flags, void *stream)
{
    int rc = check_batch(ctx, batch);
    if (rc) return rc;
    if (!res) return PB2_EINVAL;
    if (!ctx->scaler.set || !ctx->seg_set) return fail(ctx, PB2_ESTATE, "parameters not set");
    if ((flags & PB2_FLAG_BARCODING) && !ctx->demux.set)
        return fail(ctx, PB2_ESTATE, "barcoding requested but demux not set");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = batch->n_reads;
    if (n == 0) {
        if (res->counts)
            PB_CUDA(ctx, cudaMemsetAsync(res->counts, 0, sizeof(int64_t) * PB2_N_LABEL *
                                         PB2_N_BARCODE_SLOTS * PB2_N_STATUS, st));
        return PB2_OK;
    }
    const int stride = ctx->scaler.stride;
    const size_t n_pooled = (size_t)(batch->n_raw_total / stride) + 2;

    float *pooled = (float *)ws_get(ctx, ctx->ws_pooled, sizeof(float) * n_pooled);
    int32_t *status = WS_OR(res->status, ws_status, int32_t, n);
    int32_t *label = WS_OR(res->label, ws_label, int32_t, n);
    float *scale_shift = WS_OR(res->scale_shift, ws_scale, float, 2 * n);
    int32_t *segments = WS_OR(res->segments, ws_seg, int32_t, 2 * PB2_MAX_STATES * n);
    if (!pooled || !status || !label || !scale_shift || !segments) return PB2_ENOMEM;
    float *pooled_out = (flags & PB2_FLAG_KEEP_POOLED) ? res->pooled : nullptr;

    if ((rc = launch_pool(ctx, *batch, stride, pooled, st))) return rc;
    ctx->last_rerun_reads = 0;
    // Tensor-core scaler + demultiplexer with exact re-run of the unsafe reads.  Consumers of
    // the scaled signal itself (poly(A) measurement, the returned pooled signal, the chimera
    // filter through PB2_FLAG_EXACT_SCALER) get the exact scaler.
    if (ctx->fast_lstm && !ctx->exact_division &&
        !(flags & (PB2_FLAG_POLYA | PB2_FLAG_KEEP_POOLED | PB2_FLAG_EXACT_SCALER)))
        return analyze_device_fast(ctx, batch, res, flags, st, pooled, status, label, scale_shift, segments);
    if ((rc = launch_scaler(ctx, *batch, pooled, status, scale_shift, nullptr, st))) return rc;
    if ((rc = launch_segment(ctx, *batch, pooled, scale_shift, status, segments, pooled_out, st)))
        return rc;

    int32_t *barcode = nullptr, *guess = nullptr, *score = nullptr;
    if (flags & PB2_FLAG_BARCODING) {
        const int T = ctx->demux.trim_length;
        float *windows = (float *)ws_get(ctx, ctx->ws_win, sizeof(float) * (size_t)n * T);
        int32_t *pushed = (int32_t *)ws_get(ctx, ctx->ws_pushed, sizeof(int32_t) * (size_t)n);
        barcode = WS_OR(res->barcode, ws_bc, int32_t, n);
        guess = WS_OR(res->barcode_guess, ws_guess, int32_t, n);
        score = WS_OR(res->barcode_score, ws_score, int32_t, n);
        int32_t *slots = (int32_t *)ws_get(ctx, ctx->ws_slots, sizeof(int32_t) * ((size_t)n + 4));
        if (!windows || !pushed || !barcode || !guess || !score || !slots) return PB2_ENOMEM;
        int *slot_count = (int *)slots;           // [0] = number of accepted windows
        int32_t *slot_read = slots + 4;
        if (res->class_probs)
            PB_CUDA(ctx, cudaMemsetAsync(res->class_probs, 0,
                                         sizeof(float) * PB2_MAX_CLASSES * (size_t)n, st));
        if ((rc = launch_windows(ctx, *batch, pooled, scale_shift, status, segments, windows,
                                 pushed, slot_count, slot_read, st))) return rc;
        if (ctx->strict_tc_demux && !ctx->exact_division) {
            // "strict" mode: (scale, shift), the normalised signal, segments and windows above are
            // the exact kernels' for every read; only the classifier runs on the tensor cores,
            // with the margin test and an exact re-run of the windows it flags
            if ((rc = launch_demux_tc(ctx, windows, nullptr, n, slot_count, slot_read, res->class_probs,
                                      barcode, guess, score, nullptr, nullptr, nullptr, /*recheck=*/true,
                                      st, nullptr))) return rc;
        } else if ((rc = launch_demux(ctx, windows, nullptr, n, slot_count, slot_read, res->class_probs,
                                      barcode, guess, score, st))) return rc;
    } else {
        barcode = res->barcode; guess = res->barcode_guess; score = res->barcode_score;
    }
    if (flags & PB2_FLAG_POLYA) {
        if (!ctx->polya_set) return fail(ctx, PB2_ESTATE, "poly(A) requested but parameters not set");
        if (!res->polya) return fail(ctx, PB2_EINVAL, "PB2_FLAG_POLYA needs res->polya");
        if ((rc = launch_polya(ctx, *batch, scale_shift, status, segments, res->polya, st))) return rc;
    }
    if ((rc = launch_finalize(ctx, n, flags, status, label, barcode, guess, score, st))) return rc;
    if (res->counts)
        if ((rc = launch_counts(ctx, status, label, barcode, n, res->counts, st))) return rc;
    return PB2_OK;
}

static int analyze_host_single(pb2_context *ctx, const pb2_batch *hb, const pb2_results *hr, uint32_t flags)
{
    int rc = check_batch(ctx, hb);
    if (rc) return rc;
    if (!hr) return PB2_EINVAL;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->host_stream;
    const int64_t n = hb->n_reads;
    const int stride = ctx->scaler.stride > 0 ? ctx->scaler.stride : 15;
    const size_t n_pooled = (size_t)(hb->n_raw_total / stride) + 2;
    const int n_bins = PB2_N_LABEL * PB2_N_BARCODE_SLOTS * PB2_N_STATUS;

    // one device arena for inputs and outputs of this call
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align(bytes); return o; };
    const size_t o_raw = take(sizeof(int16_t) * (size_t)hb->n_raw_total + 16);
    const bool packed = hb->packed != nullptr;
    const size_t packed_bytes = packed && n > 0 ? (size_t)hb->packed_offsets[n] : 0;
    const size_t o_pk = take(packed ? packed_bytes + 16 : 16), o_po = take(packed ? 8 * (size_t)(n + 1) : 16);
    const size_t o_off = take(sizeof(int64_t) * n), o_len = take(sizeof(int64_t) * n);
    const size_t o_rng = take(sizeof(double) * n), o_dig = take(sizeof(double) * n);
    const size_t o_ofs = take(sizeof(double) * n);
    const size_t o_status = take(sizeof(int32_t) * n), o_label = take(sizeof(int32_t) * n);
    const size_t o_ss = take(sizeof(float) * 2 * n);
    const size_t o_seg = take(sizeof(int32_t) * 2 * PB2_MAX_STATES * n);
    const size_t o_bc = take(sizeof(int32_t) * n), o_gs = take(sizeof(int32_t) * n);
    const size_t o_sc = take(sizeof(int32_t) * n);
    const size_t o_pr = take(sizeof(float) * PB2_MAX_CLASSES * n);
    const size_t o_cnt = take(sizeof(int64_t) * n_bins);
    const bool keep = (flags & PB2_FLAG_KEEP_POOLED) && hr->pooled;
    const size_t o_pool = take(keep ? sizeof(float) * n_pooled : 16);
    const bool want_polya = (flags & PB2_FLAG_POLYA) && hr->polya;
    const size_t o_polya = take(want_polya ? sizeof(pb2_polya_result) * (size_t)n : 16);
    char *base = (char *)ws_get(ctx, ctx->ws_batch, off);
    if (!base) return PB2_ENOMEM;

    pb2_batch db = *hb;
    db.raw = (const int16_t *)(base + o_raw);
    db.raw_offsets = (const int64_t *)(base + o_off);
    db.raw_lengths = (const int64_t *)(base + o_len);
    db.range = (const double *)(base + o_rng);
    db.digitisation = (const double *)(base + o_dig);
    db.offset = (const double *)(base + o_ofs);
    if (db.max_raw_length <= 0)
        for (int64_t i = 0; i < n; i++)
            if (hb->raw_lengths[i] > db.max_raw_length) db.max_raw_length = hb->raw_lengths[i];
    db.packed = nullptr; db.packed_offsets = nullptr;
    if (n > 0) {
        PB_CUDA(ctx, cudaMemcpyAsync((void *)db.raw_offsets, hb->raw_offsets, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
        PB_CUDA(ctx, cudaMemcpyAsync((void *)db.raw_lengths, hb->raw_lengths, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
        if (packed) {
            // compressed upload: the streamvbyte bodies cross the bus, the samples are rebuilt here
            PB_CUDA(ctx, cudaMemcpyAsync(base + o_pk, hb->packed, packed_bytes, cudaMemcpyHostToDevice, st));
            PB_CUDA(ctx, cudaMemcpyAsync(base + o_po, hb->packed_offsets, 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
            PB_CUDA(ctx, cudaMemsetAsync(ctx->tc_err + 1, 0, sizeof(int), st));
            if ((rc = launch_svb16_decode(ctx, (const uint8_t *)(base + o_pk), (const int64_t *)(base + o_po),
                                          db.raw_offsets, db.raw_lengths, n, (int16_t *)db.raw,
                                          ctx->tc_err + 1, st))) return rc;
        } else {
            PB_CUDA(ctx, cudaMemcpyAsync((void *)db.raw, hb->raw, sizeof(int16_t) * (size_t)hb->n_raw_total,
                                         cudaMemcpyHostToDevice, st));
        }
        PB_CUDA(ctx, cudaMemcpyAsync((void *)db.range, hb->range, sizeof(double) * n, cudaMemcpyHostToDevice, st));
        PB_CUDA(ctx, cudaMemcpyAsync((void *)db.digitisation, hb->digitisation, sizeof(double) * n, cudaMemcpyHostToDevice, st));
        PB_CUDA(ctx, cudaMemcpyAsync((void *)db.offset, hb->offset, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    }
    pb2_results dr = {};
    dr.status = (int32_t *)(base + o_status);
    dr.label = (int32_t *)(base + o_label);
    dr.scale_shift = (float *)(base + o_ss);
    dr.segments = (int32_t *)(base + o_seg);
    dr.barcode = (int32_t *)(base + o_bc);
    dr.barcode_guess = (int32_t *)(base + o_gs);
    dr.barcode_score = (int32_t *)(base + o_sc);
    dr.class_probs = (float *)(base + o_pr);
    dr.counts = (int64_t *)(base + o_cnt);
    dr.pooled = keep ? (float *)(base + o_pool) : nullptr;
    dr.polya = want_polya ? (pb2_polya_result *)(base + o_polya) : nullptr;
    if ((flags & PB2_FLAG_POLYA) && !want_polya) flags &= ~PB2_FLAG_POLYA;
    if (keep) PB_CUDA(ctx, cudaMemsetAsync(dr.pooled, 0, sizeof(float) * n_pooled, st));
    if (!(flags & PB2_FLAG_BARCODING) && n > 0) {
        PB_CUDA(ctx, cudaMemsetAsync(dr.barcode, 0xFF, sizeof(int32_t) * n, st));
        PB_CUDA(ctx, cudaMemsetAsync(dr.barcode_guess, 0xFF, sizeof(int32_t) * n, st));
        PB_CUDA(ctx, cudaMemsetAsync(dr.barcode_score, 0xFF, sizeof(int32_t) * n, st));
        PB_CUDA(ctx, cudaMemsetAsync(dr.class_probs, 0, sizeof(float) * PB2_MAX_CLASSES * n, st));
    }
    if ((rc = pb2_analyze_device(ctx, &db, &dr, flags, st))) return rc;

#define PB_D2H(field, bytes)                                                                 \
    if (hr->field && (bytes) > 0)                                                            \
        PB_CUDA(ctx, cudaMemcpyAsync(hr->field, dr.field, (bytes), cudaMemcpyDeviceToHost, st))
    PB_D2H(status, sizeof(int32_t) * n);
    PB_D2H(label, sizeof(int32_t) * n);
    PB_D2H(scale_shift, sizeof(float) * 2 * n);
    PB_D2H(segments, sizeof(int32_t) * 2 * PB2_MAX_STATES * n);
    PB_D2H(barcode, sizeof(int32_t) * n);
    PB_D2H(barcode_guess, sizeof(int32_t) * n);
    PB_D2H(barcode_score, sizeof(int32_t) * n);
    PB_D2H(class_probs, sizeof(float) * PB2_MAX_CLASSES * n);
    PB_D2H(counts, sizeof(int64_t) * n_bins);
    if (keep) PB_D2H(pooled, sizeof(float) * n_pooled);
    if (want_polya) PB_D2H(polya, sizeof(pb2_polya_result) * (size_t)n);
#undef PB_D2H
    int svb_err = 0;
    if (packed) PB_CUDA(ctx, cudaMemcpyAsync(&svb_err, ctx->tc_err + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    PB_CUDA(ctx, cudaStreamSynchronize(st));
    if (svb_err) return fail(ctx, PB2_EINVAL, "packed input: a streamvbyte stream is shorter than its keys promise");
    return PB2_OK;
}


// Host-buffer path for large batches: the batch is cut into chunks of reads and run as a
// three-stage pipeline on three streams -- H2D of chunk c+1 and D2H of chunk c-1 overlap
// the kernels of chunk c (two device arenas, events for hand-over).  With pinned host
// buffers the copies disappear behind the compute.
static int analyze_host_pipelined(pb2_context *ctx, const pb2_batch *hb, const pb2_results *hr,
                                  uint32_t flags, const std::vector<int64_t> &bounds)
{
    DeviceGuard g(ctx->device);
    const int n_bins = PB2_N_LABEL * PB2_N_BARCODE_SLOTS * PB2_N_STATUS;
    const int nchunks = (int)bounds.size() - 1;
    if (!ctx->copy_in) PB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    if (!ctx->copy_out) PB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    cudaStream_t compute = ctx->host_stream;
    const bool packed = hb->packed != nullptr;
    int64_t max_reads = 0, max_span = 0, max_pk = 0;
    for (int c = 0; c < nchunks; c++) {
        const int64_t c0 = bounds[c], c1 = bounds[c + 1];
        const int64_t span = hb->raw_offsets[c1 - 1] + hb->raw_lengths[c1 - 1] - hb->raw_offsets[c0];
        if (c1 - c0 > max_reads) max_reads = c1 - c0;
        if (span > max_span) max_span = span;
        if (packed && hb->packed_offsets[c1] - hb->packed_offsets[c0] > max_pk)
            max_pk = hb->packed_offsets[c1] - hb->packed_offsets[c0];
    }
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t m = (size_t)max_reads;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align(bytes); return o; };
    const size_t o_raw = take(sizeof(int16_t) * (size_t)max_span + 32);
    const size_t o_pk = take(packed ? (size_t)max_pk + 32 : 16), o_po = take(packed ? 8 * (m + 1) : 16);
    const size_t o_off = take(8 * m), o_len = take(8 * m), o_rng = take(8 * m), o_dig = take(8 * m);
    const size_t o_ofs = take(8 * m), o_status = take(4 * m), o_label = take(4 * m);
    const size_t o_ss = take(8 * m), o_seg = take(4 * 2 * PB2_MAX_STATES * m);
    const size_t o_bc = take(4 * m), o_gs = take(4 * m), o_sc = take(4 * m);
    const size_t o_pr = take(4 * PB2_MAX_CLASSES * m), o_cnt = take(8 * n_bins);
    const bool want_polya = (flags & PB2_FLAG_POLYA) && hr->polya;
    const size_t o_polya = take(want_polya ? sizeof(pb2_polya_result) * m : 16);
    if ((flags & PB2_FLAG_POLYA) && !want_polya) flags &= ~PB2_FLAG_POLYA;
    const size_t arena_bytes = off;
    char *base = (char *)ws_get(ctx, ctx->ws_batch, 2 * arena_bytes);
    if (!base) return PB2_ENOMEM;
    // pinned landing area for the per-chunk counts, kept across calls
    const size_t counts_bytes = sizeof(int64_t) * n_bins * (size_t)nchunks;
    if (ctx->counts_host_bytes < counts_bytes) {
        if (ctx->counts_host) cudaFreeHost(ctx->counts_host);
        ctx->counts_host = nullptr; ctx->counts_host_bytes = 0;
        PB_CUDA(ctx, cudaMallocHost(&ctx->counts_host, counts_bytes));
        ctx->counts_host_bytes = counts_bytes;
    }
    int64_t *counts_host = ctx->counts_host;
    // pinned staging for the per-chunk rebased offsets: a copy from pageable memory makes the
    // host wait for everything queued before it on the upload stream (the previous chunk's
    // kernels included, through the arena hand-over event), and while the host waits the next
    // chunk's kernels are not queued -- uploads then stop overlapping compute
    const size_t stage_per_arena = 2 * (m + 2);
    if (ctx->stage_host_bytes < 2 * stage_per_arena * 8) {
        if (ctx->stage_host) cudaFreeHost(ctx->stage_host);
        ctx->stage_host = nullptr; ctx->stage_host_bytes = 0;
        PB_CUDA(ctx, cudaMallocHost(&ctx->stage_host, 2 * stage_per_arena * 8));
        ctx->stage_host_bytes = 2 * stage_per_arena * 8;
    }
    cudaEvent_t ev_h2d[2], ev_comp[2], ev_d2h[2];
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&ev_h2d[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_comp[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_d2h[i], cudaEventDisableTiming);
    }
    int64_t *rebased[2] = {ctx->stage_host, ctx->stage_host + stage_per_arena};
    int64_t *rebased_pk[2] = {rebased[0] + m + 1, rebased[1] + m + 1};
    if (packed) PB_CUDA(ctx, cudaMemsetAsync(ctx->tc_err + 1, 0, sizeof(int), compute));
    int rc = PB2_OK;
    auto run = [&]() -> int {
        // results of chunk c leave on the third stream.  Issued one chunk LATE: the caller's
        // result buffers are usually pageable, which makes these copies block the host, and
        // the next chunk's H2D and kernels must already be queued when that happens.
        auto drain = [&](int c) -> int {
            const int a = c & 1;
            char *A = base + (size_t)a * arena_bytes;
            const int64_t c0 = bounds[c], nc = bounds[c + 1] - c0;
            pb2_results dr = {};
            dr.status = (int32_t *)(A + o_status); dr.label = (int32_t *)(A + o_label);
            dr.scale_shift = (float *)(A + o_ss); dr.segments = (int32_t *)(A + o_seg);
            dr.barcode = (int32_t *)(A + o_bc); dr.barcode_guess = (int32_t *)(A + o_gs);
            dr.barcode_score = (int32_t *)(A + o_sc); dr.class_probs = (float *)(A + o_pr);
            dr.counts = (int64_t *)(A + o_cnt);
            dr.polya = want_polya ? (pb2_polya_result *)(A + o_polya) : nullptr;
            cudaStream_t co = ctx->copy_out;
            PB_CUDA(ctx, cudaStreamWaitEvent(co, ev_comp[a], 0));
#define PB_OUT(field, devp, bytes_per)                                                         \
            if (hr->field) PB_CUDA(ctx, cudaMemcpyAsync((char *)hr->field + (size_t)c0 * (bytes_per), devp, \
                                                        (size_t)nc * (bytes_per), cudaMemcpyDeviceToHost, co))
            PB_OUT(status, dr.status, 4);
            PB_OUT(label, dr.label, 4);
            PB_OUT(scale_shift, dr.scale_shift, 8);
            PB_OUT(segments, dr.segments, 4 * 2 * PB2_MAX_STATES);
            PB_OUT(barcode, dr.barcode, 4);
            PB_OUT(barcode_guess, dr.barcode_guess, 4);
            PB_OUT(barcode_score, dr.barcode_score, 4);
            PB_OUT(class_probs, dr.class_probs, 4 * PB2_MAX_CLASSES);
            if (want_polya) PB_OUT(polya, dr.polya, sizeof(pb2_polya_result));
#undef PB_OUT
            PB_CUDA(ctx, cudaMemcpyAsync(counts_host + (size_t)c * n_bins, dr.counts, 8 * n_bins,
                                         cudaMemcpyDeviceToHost, co));
            PB_CUDA(ctx, cudaEventRecord(ev_d2h[a], co));
            return PB2_OK;
        };
        // inputs of chunk c go up on the first stream.  Issued one chunk EARLY (before the
        // kernels of chunk c - 1 are queued): pb2_analyze_device may block the host while it
        // waits for the size of its exact re-run, and the next upload must already be in flight.
        // The arena's input region is free once the kernels of chunk c - 2 are done.
        std::vector<int64_t> spans((size_t)nchunks), max_lens((size_t)nchunks);
        auto upload = [&](int c) -> int {
            const int a = c & 1;
            char *A = base + (size_t)a * arena_bytes;
            const int64_t c0 = bounds[c], c1 = bounds[c + 1], nc = c1 - c0;
            const int64_t roff0 = hb->raw_offsets[c0];
            const int64_t span = hb->raw_offsets[c1 - 1] + hb->raw_lengths[c1 - 1] - roff0;
            if (c >= 2) {
                // the staging slot of this arena was last read by the upload of chunk c - 2
                PB_CUDA(ctx, cudaEventSynchronize(ev_h2d[a]));
                PB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ev_comp[a], 0));
            }
            int64_t max_len = 0;
            for (int64_t i = 0; i < nc; i++) {
                rebased[a][i] = hb->raw_offsets[c0 + i] - roff0;
                if (hb->raw_lengths[c0 + i] > max_len) max_len = hb->raw_lengths[c0 + i];
            }
            spans[c] = span; max_lens[c] = max_len;
            cudaStream_t ci = ctx->copy_in;
            if (packed) {
                const int64_t p0 = hb->packed_offsets[c0];
                for (int64_t i = 0; i <= nc; i++) rebased_pk[a][i] = hb->packed_offsets[c0 + i] - p0;
                PB_CUDA(ctx, cudaMemcpyAsync(A + o_pk, hb->packed + p0, (size_t)(hb->packed_offsets[c1] - p0), cudaMemcpyHostToDevice, ci));
                PB_CUDA(ctx, cudaMemcpyAsync(A + o_po, rebased_pk[a], 8 * (size_t)(nc + 1), cudaMemcpyHostToDevice, ci));
            } else
            PB_CUDA(ctx, cudaMemcpyAsync(A + o_raw, hb->raw + roff0, sizeof(int16_t) * (size_t)span, cudaMemcpyHostToDevice, ci));
            PB_CUDA(ctx, cudaMemcpyAsync(A + o_off, rebased[a], 8 * (size_t)nc, cudaMemcpyHostToDevice, ci));
            PB_CUDA(ctx, cudaMemcpyAsync(A + o_len, hb->raw_lengths + c0, 8 * (size_t)nc, cudaMemcpyHostToDevice, ci));
            PB_CUDA(ctx, cudaMemcpyAsync(A + o_rng, hb->range + c0, 8 * (size_t)nc, cudaMemcpyHostToDevice, ci));
            PB_CUDA(ctx, cudaMemcpyAsync(A + o_dig, hb->digitisation + c0, 8 * (size_t)nc, cudaMemcpyHostToDevice, ci));
            PB_CUDA(ctx, cudaMemcpyAsync(A + o_ofs, hb->offset + c0, 8 * (size_t)nc, cudaMemcpyHostToDevice, ci));
            PB_CUDA(ctx, cudaEventRecord(ev_h2d[a], ci));
            return PB2_OK;
        };
        { int r0 = upload(0); if (r0) return r0; }
        for (int c = 0; c < nchunks; c++) {
            const int a = c & 1;
            char *A = base + (size_t)a * arena_bytes;
            const int64_t nc = bounds[c + 1] - bounds[c];
            if (c + 1 < nchunks) { int r1 = upload(c + 1); if (r1) return r1; }
            const int64_t span = spans[c], max_len = max_lens[c];

            pb2_batch db = {};
            db.n_reads = nc; db.n_raw_total = span; db.max_raw_length = max_len;
            db.raw = (const int16_t *)(A + o_raw);
            db.raw_offsets = (const int64_t *)(A + o_off);
            db.raw_lengths = (const int64_t *)(A + o_len);
            db.range = (const double *)(A + o_rng);
            db.digitisation = (const double *)(A + o_dig);
            db.offset = (const double *)(A + o_ofs);
            pb2_results dr = {};
            dr.status = (int32_t *)(A + o_status); dr.label = (int32_t *)(A + o_label);
            dr.scale_shift = (float *)(A + o_ss); dr.segments = (int32_t *)(A + o_seg);
            dr.barcode = (int32_t *)(A + o_bc); dr.barcode_guess = (int32_t *)(A + o_gs);
            dr.barcode_score = (int32_t *)(A + o_sc); dr.class_probs = (float *)(A + o_pr);
            dr.counts = (int64_t *)(A + o_cnt);
            dr.polya = want_polya ? (pb2_polya_result *)(A + o_polya) : nullptr;
            // the arena's result region is free once the results of chunk c - 2 have left
            if (c >= 2) PB_CUDA(ctx, cudaStreamWaitEvent(compute, ev_d2h[a], 0));
            PB_CUDA(ctx, cudaStreamWaitEvent(compute, ev_h2d[a], 0));
            if (!(flags & PB2_FLAG_BARCODING)) {
                PB_CUDA(ctx, cudaMemsetAsync(dr.barcode, 0xFF, 4 * (size_t)nc, compute));
                PB_CUDA(ctx, cudaMemsetAsync(dr.barcode_guess, 0xFF, 4 * (size_t)nc, compute));
                PB_CUDA(ctx, cudaMemsetAsync(dr.barcode_score, 0xFF, 4 * (size_t)nc, compute));
                PB_CUDA(ctx, cudaMemsetAsync(dr.class_probs, 0, 4 * PB2_MAX_CLASSES * (size_t)nc, compute));
            }
            if (packed) {
                int rd = launch_svb16_decode(ctx, (const uint8_t *)(A + o_pk), (const int64_t *)(A + o_po),
                                             db.raw_offsets, db.raw_lengths, nc, (int16_t *)(A + o_raw),
                                             ctx->tc_err + 1, compute);
                if (rd) return rd;
            }
            int r2 = pb2_analyze_device(ctx, &db, &dr, flags & ~PB2_FLAG_KEEP_POOLED, compute);
            if (r2) return r2;
            PB_CUDA(ctx, cudaEventRecord(ev_comp[a], compute));

            if (c >= 1) { int r3 = drain(c - 1); if (r3) return r3; }
        }
        { int r3 = drain(nchunks - 1); if (r3) return r3; }
        int svb_err = 0;
        if (packed) PB_CUDA(ctx, cudaMemcpyAsync(&svb_err, ctx->tc_err + 1, sizeof(int), cudaMemcpyDeviceToHost, compute));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_out));
        PB_CUDA(ctx, cudaStreamSynchronize(compute));
        if (svb_err) return fail(ctx, PB2_EINVAL, "packed input: a streamvbyte stream is shorter than its keys promise");
        return PB2_OK;
    };
    rc = run();
    if (rc == PB2_OK && hr->counts) {
        for (int i = 0; i < n_bins; i++) {
            int64_t t = 0;
            for (int c = 0; c < nchunks; c++) t += counts_host[(size_t)c * n_bins + i];
            hr->counts[i] = t;
        }
    }
    if (rc != PB2_OK) cudaDeviceSynchronize();
    for (int i = 0; i < 2; i++) { cudaEventDestroy(ev_h2d[i]); cudaEventDestroy(ev_comp[i]); cudaEventDestroy(ev_d2h[i]); }
    return rc;
}

// Host-buffer path for large batches, whole batch resident on the device (the default).
// The raw buffer is uploaded chunk by chunk into ONE device buffer at its own offsets, so a chunk
// is just a range of reads: no rebasing, no arena hand-over.  The kernels of chunk c start when
// its upload has landed and overlap the uploads of the chunks behind it.  In `fast` mode only
// the tensor-core half runs per chunk; the reads any chunk flagged are resolved by the exact
// kernels as ONE sub-batch at the end (a sub-batch per chunk ends on a partly filled wave of the
// long exact CTAs every time, which cost 20-30 ms per million reads) -- or as two, the first half
// way through when the GPU is found waiting for the bus there -- then the counts run over the
// whole batch and the results go back in one go (124 B per read).
// Chunks are small at both ends and large in the middle (plan_chunks): a small first chunk
// because its upload is the one nothing hides, a small last one because its kernels are what
// nothing hides when the bus is the slower side (several GPUs sharing the host's PCIe uplinks).
static int analyze_host_streamed(pb2_context *ctx, const pb2_batch *hb, const pb2_results *hr,
                                 uint32_t flags, const std::vector<int64_t> &bounds)
{
    DeviceGuard g(ctx->device);
    const int n_bins = PB2_N_LABEL * PB2_N_BARCODE_SLOTS * PB2_N_STATUS;
    const int nchunks = (int)bounds.size() - 1;
    const int64_t n = hb->n_reads;
    if (!ctx->scaler.set || !ctx->seg_set) return fail(ctx, PB2_ESTATE, "parameters not set");
    if ((flags & PB2_FLAG_BARCODING) && !ctx->demux.set)
        return fail(ctx, PB2_ESTATE, "barcoding requested but demux not set");
    if (!ctx->copy_in) PB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    cudaStream_t compute = ctx->host_stream, ci = ctx->copy_in;
    const bool packed = hb->packed != nullptr;
    const size_t packed_bytes = packed ? (size_t)hb->packed_offsets[n] : 0;
    const bool want_polya = (flags & PB2_FLAG_POLYA) && hr->polya;
    if ((flags & PB2_FLAG_POLYA) && !want_polya) flags &= ~PB2_FLAG_POLYA;
    flags &= ~PB2_FLAG_KEEP_POOLED;
    const bool fast = ctx->fast_lstm && !ctx->exact_division &&
                      !(flags & (PB2_FLAG_POLYA | PB2_FLAG_EXACT_SCALER));
    const int stride = ctx->scaler.stride > 0 ? ctx->scaler.stride : 15;

    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t m = (size_t)n;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align(bytes); return o; };
    const size_t o_raw = take(sizeof(int16_t) * (size_t)hb->n_raw_total + 32);
    const size_t o_pk = take(packed ? packed_bytes + 32 : 16), o_po = take(packed ? 8 * (m + 1) : 16);
    const size_t o_off = take(8 * m), o_len = take(8 * m), o_rng = take(8 * m), o_dig = take(8 * m);
    const size_t o_ofs = take(8 * m), o_status = take(4 * m), o_label = take(4 * m);
    const size_t o_ss = take(8 * m), o_seg = take(4 * 2 * PB2_MAX_STATES * m);
    const size_t o_bc = take(4 * m), o_gs = take(4 * m), o_sc = take(4 * m);
    const size_t o_pr = take(4 * PB2_MAX_CLASSES * m), o_cnt = take(8 * n_bins);
    const size_t o_polya = take(want_polya ? sizeof(pb2_polya_result) * m : 16);
    const size_t o_pushed = take(4 * m);
    const size_t o_unsafe = take(4 * m), o_fcnt = take(64), o_list = take(4 * m);
    char *A = (char *)ws_get(ctx, ctx->ws_batch, off);
    float *pooled = (float *)ws_get(ctx, ctx->ws_pooled,
                                    sizeof(float) * ((size_t)(hb->n_raw_total / stride) + 2));
    if (!A || !pooled) return PB2_ENOMEM;

    pb2_batch db = *hb;
    db.raw = (const int16_t *)(A + o_raw);
    db.raw_offsets = (const int64_t *)(A + o_off);
    db.raw_lengths = (const int64_t *)(A + o_len);
    db.range = (const double *)(A + o_rng);
    db.digitisation = (const double *)(A + o_dig);
    db.offset = (const double *)(A + o_ofs);
    db.packed = nullptr; db.packed_offsets = nullptr;
    pb2_results dr = {};
    dr.status = (int32_t *)(A + o_status); dr.label = (int32_t *)(A + o_label);
    dr.scale_shift = (float *)(A + o_ss); dr.segments = (int32_t *)(A + o_seg);
    dr.barcode = (int32_t *)(A + o_bc); dr.barcode_guess = (int32_t *)(A + o_gs);
    dr.barcode_score = (int32_t *)(A + o_sc); dr.class_probs = (float *)(A + o_pr);
    dr.counts = (int64_t *)(A + o_cnt);
    dr.polya = want_polya ? (pb2_polya_result *)(A + o_polya) : nullptr;
    FastArrays fa = {(int32_t *)(A + o_unsafe), (int32_t *)(A + o_pushed), dr.barcode, dr.barcode_guess,
                     dr.barcode_score};

    std::vector<cudaEvent_t> ev((size_t)nchunks + 1, nullptr);
    for (auto &e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    // longest read per chunk (sizes the grids): found chunk by chunk while earlier chunks are in flight
    db.max_raw_length = 0;
    auto chunk_max_len = [&](int c) {
        int64_t ml = 0;
        for (int64_t i = bounds[c]; i < bounds[c + 1]; i++)
            if (hb->raw_lengths[i] > ml) ml = hb->raw_lengths[i];
        if (ml > db.max_raw_length) db.max_raw_length = ml;
        return ml;
    };

    int svb_err = 0;
    auto run = [&]() -> int {
        int rc;
        // per-read metadata of the whole batch (40 bytes per read), then the samples chunk by chunk
        PB_CUDA(ctx, cudaMemcpyAsync(A + o_off, hb->raw_offsets, 8 * m, cudaMemcpyHostToDevice, ci));
        PB_CUDA(ctx, cudaMemcpyAsync(A + o_len, hb->raw_lengths, 8 * m, cudaMemcpyHostToDevice, ci));
        PB_CUDA(ctx, cudaMemcpyAsync(A + o_rng, hb->range, 8 * m, cudaMemcpyHostToDevice, ci));
        PB_CUDA(ctx, cudaMemcpyAsync(A + o_dig, hb->digitisation, 8 * m, cudaMemcpyHostToDevice, ci));
        PB_CUDA(ctx, cudaMemcpyAsync(A + o_ofs, hb->offset, 8 * m, cudaMemcpyHostToDevice, ci));
        if (packed)
            PB_CUDA(ctx, cudaMemcpyAsync(A + o_po, hb->packed_offsets, 8 * (m + 1), cudaMemcpyHostToDevice, ci));
        PB_CUDA(ctx, cudaEventRecord(ev[(size_t)nchunks], ci));
        auto upload = [&](int c) -> int {
            const int64_t c0 = bounds[c], c1 = bounds[c + 1];
            if (packed) {
                const int64_t p0 = hb->packed_offsets[c0], p1 = hb->packed_offsets[c1];
                if (p1 > p0)
                    PB_CUDA(ctx, cudaMemcpyAsync(A + o_pk + p0, hb->packed + p0, (size_t)(p1 - p0),
                                                 cudaMemcpyHostToDevice, ci));
            } else {
                const int64_t r0 = hb->raw_offsets[c0];
                const int64_t span = hb->raw_offsets[c1 - 1] + hb->raw_lengths[c1 - 1] - r0;
                if (span > 0)
                    PB_CUDA(ctx, cudaMemcpyAsync(A + o_raw + 2 * (size_t)r0, hb->raw + r0,
                                                 sizeof(int16_t) * (size_t)span, cudaMemcpyHostToDevice, ci));
            }
            PB_CUDA(ctx, cudaEventRecord(ev[(size_t)c], ci));
            return PB2_OK;
        };
        // uploads are issued two chunks ahead of the kernels: from pageable memory a copy blocks
        // the host, and the kernels of the chunks before it must already be queued by then
        if ((rc = upload(0))) return rc;
        if (nchunks > 1 && (rc = upload(1))) return rc;

        PB_CUDA(ctx, cudaStreamWaitEvent(compute, ev[(size_t)nchunks], 0));
        if (fast) PB_CUDA(ctx, cudaMemsetAsync(A + o_unsafe, 0, o_list - o_unsafe, compute));
        if (packed) PB_CUDA(ctx, cudaMemsetAsync(ctx->tc_err + 1, 0, sizeof(int), compute));
        if (!(flags & PB2_FLAG_BARCODING)) {
            PB_CUDA(ctx, cudaMemsetAsync(dr.barcode, 0xFF, 4 * m, compute));
            PB_CUDA(ctx, cudaMemsetAsync(dr.barcode_guess, 0xFF, 4 * m, compute));
            PB_CUDA(ctx, cudaMemsetAsync(dr.barcode_score, 0xFF, 4 * m, compute));
            PB_CUDA(ctx, cudaMemsetAsync(dr.class_probs, 0, 4 * PB2_MAX_CLASSES * m, compute));
        }
        ctx->last_rerun_reads = 0;
        // exact re-run of the flagged reads of [r0, r1) (label included; counts come at the end)
        int64_t resolved = 0, reruns = 0, causes[3] = {0, 0, 0};
        auto resolve_range = [&](int64_t r0, int64_t r1, int slot) -> int {
            if (r1 <= r0) return PB2_OK;
            pb2_batch rb = db;
            rb.n_reads = r1 - r0;
            rb.raw_offsets = db.raw_offsets + r0; rb.raw_lengths = db.raw_lengths + r0;
            rb.range = db.range + r0; rb.digitisation = db.digitisation + r0; rb.offset = db.offset + r0;
            pb2_results rr = {};
            rr.class_probs = dr.class_probs + (size_t)PB2_MAX_CLASSES * r0;
            FastArrays fr = {fa.unsafe + r0, fa.pushed + r0, dr.barcode + r0, dr.barcode_guess + r0,
                             dr.barcode_score + r0};
            int rc2 = analyze_fast_resolve(ctx, &rb, &rr, flags, compute, pooled, dr.status + r0, dr.label + r0,
                                           dr.scale_shift + 2 * r0, dr.segments + 2 * PB2_MAX_STATES * r0, fr,
                                           (int *)(A + o_fcnt) + 8 * slot, (int32_t *)(A + o_list) + r0);
            reruns += ctx->last_rerun_reads;
            for (int i = 0; i < 3; i++) causes[i] += ctx->last_rerun_cause[i];
            return rc2;
        };
        // When the bus is the slower side (several GPUs behind shared PCIe uplinks) the GPU waits for
        // uploads in the middle of the batch and the single re-run at the end is fully exposed: half
        // way, once the kernels of the chunks so far are done, look whether the next upload has
        // landed; if not, the flagged reads so far are re-run now, inside that wait.
        const char *er_env = getenv("POREPLEX_B200_HOST_EARLY_RESOLVE");       // 0 never, 1 always (tests)
        const int early_mode = er_env ? atoi(er_env) : -1;
        const int mid = (fast && nchunks >= 4 && early_mode != 0) ? nchunks / 2 - 1 : -1;
        for (int c = 0; c < nchunks; c++) {
            const int64_t c0 = bounds[c], nc = bounds[c + 1] - c0;
            pb2_batch cb = db;                       // a range of reads of the resident batch
            cb.n_reads = nc; cb.max_raw_length = chunk_max_len(c);
            cb.raw_offsets = db.raw_offsets + c0; cb.raw_lengths = db.raw_lengths + c0;
            cb.range = db.range + c0; cb.digitisation = db.digitisation + c0; cb.offset = db.offset + c0;
            pb2_results cr = {};
            cr.status = dr.status + c0; cr.label = dr.label + c0;
            cr.scale_shift = dr.scale_shift + 2 * c0;
            cr.segments = dr.segments + 2 * PB2_MAX_STATES * c0;
            cr.barcode = dr.barcode + c0; cr.barcode_guess = dr.barcode_guess + c0;
            cr.barcode_score = dr.barcode_score + c0;
            cr.class_probs = dr.class_probs + (size_t)PB2_MAX_CLASSES * c0;
            cr.polya = want_polya ? dr.polya + c0 : nullptr;
            PB_CUDA(ctx, cudaStreamWaitEvent(compute, ev[(size_t)c], 0));
            if (packed &&
                (rc = launch_svb16_decode(ctx, (const uint8_t *)(A + o_pk), (const int64_t *)(A + o_po) + c0,
                                          cb.raw_offsets, cb.raw_lengths, nc, (int16_t *)(A + o_raw),
                                          ctx->tc_err + 1, compute))) return rc;
            if (fast) {
                FastArrays fc = {fa.unsafe + c0, fa.pushed + c0, cr.barcode, cr.barcode_guess, cr.barcode_score};
                if ((rc = launch_pool(ctx, cb, stride, pooled, compute))) return rc;
                if ((rc = analyze_fast_tentative(ctx, &cb, &cr, flags, compute, pooled, cr.status,
                                                 cr.scale_shift, cr.segments, fc))) return rc;
            } else if ((rc = pb2_analyze_device(ctx, &cb, &cr, flags, compute))) {
                return rc;
            }
            if (c + 2 < nchunks && (rc = upload(c + 2))) return rc;
            if (c == mid) {
                PB_CUDA(ctx, cudaStreamSynchronize(compute));
                const cudaError_t q = cudaEventQuery(ev[(size_t)c + 1]);
                if (q != cudaSuccess && q != cudaErrorNotReady) return check_cuda(ctx, q, "cudaEventQuery");
                if (q == cudaErrorNotReady || early_mode == 1) {
                    if ((rc = resolve_range(0, bounds[c + 1], 0))) return rc;
                    resolved = bounds[c + 1];
                }
            }
        }
        if (fast) {
            if ((rc = resolve_range(resolved, n, 1))) return rc;
            ctx->last_rerun_reads = reruns;
            for (int i = 0; i < 3; i++) ctx->last_rerun_cause[i] = causes[i];
        }
        if ((rc = launch_counts(ctx, dr.status, dr.label, dr.barcode, n, dr.counts, compute))) return rc;
#define PB_D2H(field, bytes)                                                                 \
        if (hr->field && (bytes) > 0)                                                        \
            PB_CUDA(ctx, cudaMemcpyAsync(hr->field, dr.field, (bytes), cudaMemcpyDeviceToHost, compute))
        PB_D2H(status, 4 * m);
        PB_D2H(label, 4 * m);
        PB_D2H(scale_shift, 8 * m);
        PB_D2H(segments, 4 * 2 * PB2_MAX_STATES * m);
        PB_D2H(barcode, 4 * m);
        PB_D2H(barcode_guess, 4 * m);
        PB_D2H(barcode_score, 4 * m);
        PB_D2H(class_probs, 4 * PB2_MAX_CLASSES * m);
        PB_D2H(counts, sizeof(int64_t) * n_bins);
        if (want_polya) PB_D2H(polya, sizeof(pb2_polya_result) * m);
#undef PB_D2H
        if (packed) PB_CUDA(ctx, cudaMemcpyAsync(&svb_err, ctx->tc_err + 1, sizeof(int), cudaMemcpyDeviceToHost, compute));
        PB_CUDA(ctx, cudaStreamSynchronize(compute));
        return PB2_OK;
    };
    int rc = run();
    if (rc != PB2_OK) cudaDeviceSynchronize();
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    if (rc == PB2_OK && svb_err)
        return fail(ctx, PB2_EINVAL, "packed input: a streamvbyte stream is shorter than its keys promise");
    return rc;
}

// chunk boundaries of the streamed host path: K chunks whose sizes rise and fall
// (1 2 4 6 6 4 2 1 for the largest batches), cut on whole waves of the tensor-core kernels
static std::vector<int64_t> plan_chunks(int64_t n, int64_t wave)
{
    static const int w8[] = {1, 2, 4, 6, 6, 4, 2, 1}, w6[] = {1, 2, 4, 4, 2, 1}, w4[] = {1, 2, 2, 1},
                     w3[] = {1, 2, 1}, w2[] = {1, 1};
    const double W = (double)n / (double)wave;
    const int *w = w2; int K = 2, sum = 2;
    if (W >= 39) { w = w8; K = 8; sum = 26; }
    else if (W >= 21) { w = w6; K = 6; sum = 14; }
    else if (W >= 9) { w = w4; K = 4; sum = 6; }
    else if (W >= 6) { w = w3; K = 3; sum = 4; }
    std::vector<int64_t> bounds(1, 0);
    int acc = 0;
    for (int c = 0; c < K; c++) {
        acc += w[c];
        int64_t b = (int64_t)((double)n * acc / sum);
        if (c + 1 < K) b -= b % wave; else b = n;
        if (b > bounds.back()) bounds.push_back(b);
    }
    if (bounds.back() != n) bounds.push_back(n);
    return bounds;
}

int pb2_analyze_host(pb2_context *ctx, const pb2_batch *hb, const pb2_results *hr, uint32_t flags)
{
    int rc = check_batch(ctx, hb);
    if (rc) return rc;
    if (!hr) return PB2_EINVAL;
    // Large batches are cut into chunks of reads whose uploads overlap the kernels.  Default: the
    // streamed path (whole batch resident, analyze_host_streamed).  POREPLEX_B200_HOST_PIPELINE=arena
    // selects the two-arena pipeline (device memory for two chunks only; also taken when the batch
    // would need more than a quarter of the device memory): a small first chunk, then
    // POREPLEX_B200_HOST_CHUNKS (default 4) large ones.
    const int64_t n = hb->n_reads;
    int64_t min_elems = (int64_t)256 << 20;      // below this the copies are not worth hiding
    int64_t min_reads = 65536;
    int64_t nchunks = 4;
    bool uniform = false;
    const char *pmode = getenv("POREPLEX_B200_HOST_PIPELINE");
    bool arena = pmode && !strcmp(pmode, "arena");
    if (const char *env = getenv("POREPLEX_B200_HOST_CHUNKS")) {          // tuning: large chunks after the small first one
        const long long v = atoll(env);
        if (v >= 1 && v <= 64) nchunks = v;
    }
    if (const char *env = getenv("POREPLEX_B200_HOST_CHUNK_ELEMS")) {     // tests / tuning
        const long long v = atoll(env);
        if (v > 0) { min_elems = v; min_reads = 2048; nchunks = (hb->n_raw_total + v - 1) / v; uniform = true; }
    }
    {
        size_t free_b = 0, total_b = 0;
        DeviceGuard g(ctx->device);
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); total_b = 0; }
        if ((size_t)hb->n_raw_total * 2 > total_b / 4) arena = true;
    }
    if (arena) while (hb->n_raw_total / nchunks > ((int64_t)4 << 30)) nchunks *= 2;
    const bool keep = (flags & PB2_FLAG_KEEP_POOLED) && hr->pooled;
    if (n < min_reads || keep || hb->n_raw_total < min_elems || nchunks < 2)
        return analyze_host_single(ctx, hb, hr, flags);
    // both chunked paths upload each chunk as ONE span of the raw buffer: that needs reads laid
    // out in ascending, non-overlapping order inside the buffer.  Any other layout is legal for
    // the ABI and takes the single-arena path.
    for (int64_t i = 0; i < n; i++) {
        const int64_t o = hb->raw_offsets[i], l = hb->raw_lengths[i];
        if (o < 0 || l < 0 || o + l > hb->n_raw_total)
            return fail(ctx, PB2_EINVAL, "read %lld lies outside the raw buffer", (long long)i);
        if (i > 0 && o < hb->raw_offsets[i - 1] + hb->raw_lengths[i - 1])
            return analyze_host_single(ctx, hb, hr, flags);
    }
    std::vector<int64_t> bounds;
    const int64_t wave = (int64_t)ctx->sm_count * 128;
    if (uniform) {
        for (int64_t c = 0; c <= nchunks; c++) {
            int64_t b = (n * c) / nchunks;
            if (c < nchunks) b -= b % 128;       // keep chunk starts tile aligned
            if (bounds.empty() || b > bounds.back()) bounds.push_back(b);
        }
    } else if (!arena) {
        bounds = plan_chunks(n, wave);
    } else {
        // Only the FIRST upload is not hidden behind kernels: a small first chunk, then equal
        // large ones, all whole waves of the tensor-core kernels (one 128-read tile per SM).
        int64_t first = 3 * wave;
        if (first > n / 8) first = (n / 8) - (n / 8) % 128;
        bounds.push_back(0);
        if (first > 0) bounds.push_back(first);
        const int64_t rest = n - bounds.back();
        for (int64_t c = 1; c <= nchunks; c++) {
            int64_t b = bounds[first > 0 ? 1 : 0] + (rest * c) / nchunks;
            if (c < nchunks) { b -= b % wave; b -= b % 128; }
            if (b > bounds.back()) bounds.push_back(b);
        }
    }
    if (bounds.back() != n) bounds.push_back(n);
    if (bounds.size() < 3) return analyze_host_single(ctx, hb, hr, flags);
    if (arena) return analyze_host_pipelined(ctx, hb, hr, flags, bounds);
    return analyze_host_streamed(ctx, hb, hr, flags, bounds);
}

}  // extern "C"
