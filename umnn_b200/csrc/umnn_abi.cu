// extern "C" entry points declared in include/umnn_b200.h: argument validation, dispatch to the
// kernels, error strings.  No torch types; plain pointers and sizes only.
#include <map>
#include <mutex>
#include <tuple>
#include <utility>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "umnn_common.cuh"
#include "tc_common.cuh"
#include "tc_layout.cuh"
#include "tc_kernels.cuh"

namespace umnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    (void)cudaGetLastError();  // clear the sticky-free error state
    return (int)e;
}

namespace {
struct LaunchMemo {
    std::mutex mu;
    std::map<int, int> n_sm;                                              // device -> SM count
    std::map<std::pair<const void*, int>, int> smem;                      // (kernel, device) -> opted-in dynamic smem
    std::map<std::pair<const void*, int>, bool> carveout;
    std::map<std::tuple<const void*, int, int, size_t>, int> occ;
};
LaunchMemo& launch_memo() { static LaunchMemo m; return m; }
}  // namespace

cudaError_t current_device(int* dev, int* n_sm) {
    cudaError_t e = cudaGetDevice(dev);
    if (e != cudaSuccess) return e;
    LaunchMemo& m = launch_memo();
    std::lock_guard<std::mutex> lock(m.mu);
    auto it = m.n_sm.find(*dev);
    if (it == m.n_sm.end()) {
        int n = 0;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, *dev);
        if (e != cudaSuccess) return e;
        it = m.n_sm.emplace(*dev, n).first;
    }
    *n_sm = it->second;
    return cudaSuccess;
}

cudaError_t ensure_dynamic_smem(const void* kern, int dev, int bytes) {
    LaunchMemo& m = launch_memo();
    std::lock_guard<std::mutex> lock(m.mu);
    int& have = m.smem[std::make_pair(kern, dev)];
    if (bytes <= have) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

cudaError_t ensure_max_carveout(const void* kern, int dev) {
    LaunchMemo& m = launch_memo();
    std::lock_guard<std::mutex> lock(m.mu);
    bool& done = m.carveout[std::make_pair(kern, dev)];
    if (done) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) done = true;
    return e;
}

cudaError_t cached_occupancy(int* occ, const void* kern, int dev, int threads, size_t smem) {
    LaunchMemo& m = launch_memo();
    std::lock_guard<std::mutex> lock(m.mu);
    const auto key = std::make_tuple(kern, dev, threads, smem);
    auto it = m.occ.find(key);
    if (it == m.occ.end()) {
        int n = 0;
        const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem);
        if (e != cudaSuccess) return e;
        if (m.occ.size() > 4096) m.occ.clear();
        it = m.occ.emplace(key, n).first;
    }
    *occ = it->second;
    return cudaSuccess;
}

int validate_desc(const umnn_desc* d) {
    if (!d) { set_error("desc is NULL"); return UMNN_ERR_NULL; }
    if (d->abi_version != UMNN_ABI_VERSION) {
        set_error("desc.abi_version=%d, library is %d", d->abi_version, UMNN_ABI_VERSION);
        return UMNN_ERR_ABI;
    }
    if (d->layout != UMNN_LAYOUT_STRIDED_D && d->layout != UMNN_LAYOUT_CONTIG) {
        set_error("desc.layout=%d is not a UMNN_LAYOUT_*", d->layout); return UMNN_ERR_DESC;
    }
    if (d->n_samples < 0 || d->n_dims < 1) {
        set_error("desc.n_samples=%lld / n_dims=%d out of range", (long long)d->n_samples, d->n_dims); return UMNN_ERR_DESC;
    }
    if (d->layout == UMNN_LAYOUT_CONTIG && d->n_dims != 1) {
        set_error("UMNN_LAYOUT_CONTIG requires n_dims == 1 (got %d)", d->n_dims); return UMNN_ERR_DESC;
    }
    if (d->n_ctx < 0 || d->n_ctx + 1 > UMNN_MAX_WIDTH) {
        set_error("desc.n_ctx=%d out of range [0, %d]", d->n_ctx, UMNN_MAX_WIDTH - 1); return UMNN_ERR_DESC;
    }
    if (d->n_layers < 2 || d->n_layers > UMNN_MAX_LAYERS) {
        set_error("desc.n_layers=%d out of range [2, %d]", d->n_layers, UMNN_MAX_LAYERS); return UMNN_ERR_DESC;
    }
    if (d->widths[0] != d->n_ctx + 1) {
        set_error("desc.widths[0]=%d must equal 1 + n_ctx = %d", d->widths[0], d->n_ctx + 1); return UMNN_ERR_DESC;
    }
    if (d->widths[d->n_layers] != 1) {
        set_error("desc.widths[n_layers]=%d must be 1", d->widths[d->n_layers]); return UMNN_ERR_DESC;
    }
    for (int l = 1; l < d->n_layers; ++l)
        if (d->widths[l] < 1 || d->widths[l] > UMNN_MAX_WIDTH) {
            set_error("desc.widths[%d]=%d out of range [1, %d]", l, d->widths[l], UMNN_MAX_WIDTH); return UMNN_ERR_DESC;
        }
    if (d->hidden_act != UMNN_ACT_RELU && d->hidden_act != UMNN_ACT_LEAKY_RELU) {
        set_error("desc.hidden_act=%d is not a UMNN_ACT_*", d->hidden_act); return UMNN_ERR_DESC;
    }
    if (d->out_act != UMNN_OUT_ELU_PLUS_1 && d->out_act != UMNN_OUT_SIGMOID) {
        set_error("desc.out_act=%d is not a UMNN_OUT_*", d->out_act); return UMNN_ERR_DESC;
    }
    if (d->nb_steps < 1 || d->nb_steps > UMNN_MAX_STEPS) {
        set_error("desc.nb_steps=%d out of range [1, %d]", d->nb_steps, UMNN_MAX_STEPS); return UMNN_ERR_DESC;
    }
    if (d->precision != UMNN_PREC_FP32 && d->precision != UMNN_PREC_BF16X3 && d->precision != UMNN_PREC_AUTO &&
        d->precision != UMNN_PREC_FP16X3) {
        set_error("desc.precision=%d is not a UMNN_PREC_*", d->precision); return UMNN_ERR_DESC;
    }
    if ((long long)d->n_samples * d->n_dims > (1LL << 40)) {
        set_error("n_samples * n_dims too large"); return UMNN_ERR_DESC;
    }
    return 0;
}

// Which kernel family serves this descriptor.  UMNN_PREC_AUTO picks the FP16X3 tensor-core kernel (guarded, see
// umnn_cc_forward) whenever the shape fits it (checked with the worst-case 2 extra rows per slot so that packing and
// launching agree), else the FP32 kernel.  An explicit tensor-core precision on an unsupported shape fails.
static int resolve_precision(const umnn_desc* d) {
    if (d->precision == UMNN_PREC_AUTO) return tc_unsupported_reason(d, 2) == nullptr ? UMNN_PREC_FP16X3 : UMNN_PREC_FP32;
    return d->precision;
}

static bool is_tc(int prec) { return prec == UMNN_PREC_BF16X3 || prec == UMNN_PREC_FP16X3; }

static int check_tc(const umnn_desc* d, const char* who) {
    const char* why = tc_unsupported_reason(d, 2);
    if (why) {
        set_error("%s: tensor-core precisions unavailable for this shape (%s)", who, why);
        return UMNN_ERR_UNSUPPORTED;
    }
    return 0;
}

// Packed block of the tensor-core precisions:
//   [ bf16 forward blobs | dgrad blobs (when the tensor-core backward serves the shape) ]              UMNN_PREC_BF16X3
//   [ the same, rounded up to 256 bytes | fp16 forward blobs | (256-byte aligned) FP32 block ]          UMNN_PREC_FP16X3
// Under FP16X3 the dgrad blobs (bf16) serve passes D of the backward and the FP32 block serves the guarded re-run
// of a call whose activations left the fp16 range; the bf16 forward blobs are only filled in when the FP32
// backward cannot hold the shape in shared memory (the backward's re-run then falls back to bf16 operands).
static size_t tc_bf16_block_bytes(const umnn_desc* d) {
    return backward_tc_unsupported_reason(d) ? tc_packed_bytes(d) : backward_tc_packed_bytes(d);
}
static size_t tc_fp16_offset(const umnn_desc* d) { return (tc_bf16_block_bytes(d) + 255) / 256 * 256; }
static size_t tc_fp32_offset(const umnn_desc* d) { return (tc_fp16_offset(d) + tc_packed_bytes(d) + 255) / 256 * 256; }
// the backward's guarded re-run: FP32 kernels when they serve the shape, else the bf16 operand split
static bool bwd_rerun_is_fp32(const umnn_desc* d) { return backward_fp32_unsupported_reason(d) == nullptr; }

constexpr size_t kFlagBytes = 256;   // head of the workspace of a guarded FP16X3 call: one int flag

// Marks of the guarded forward calls (see umnn_cc_forward).  Process-wide counter, values >= 2 (0 = cleared by a
// memset, 1 = the mark of memset-based sequences: the backward and calls issued under CUDA-graph capture).
static std::atomic<uint32_t> g_forward_epoch{2};
static int next_epoch() {
    uint32_t e = g_forward_epoch.fetch_add(1, std::memory_order_relaxed);
    while (e < 2 || e > 0x7fffffffu) {            // wrapped: restart above the reserved values
        uint32_t expected = e + 1;
        g_forward_epoch.compare_exchange_strong(expected, 3);
        e = g_forward_epoch.fetch_add(1, std::memory_order_relaxed);
    }
    return (int)e;
}

}  // namespace umnn

using namespace umnn;

extern "C" {

int umnn_abi_version(void) { return UMNN_ABI_VERSION; }

const char* umnn_last_error(void) { return g_err; }

int umnn_cc_tables(int32_t Q, float* nodes_host, float* weights_host) {
    if (!nodes_host || !weights_host) { set_error("umnn_cc_tables: NULL output"); return UMNN_ERR_NULL; }
    if (Q < 1 || Q > UMNN_MAX_STEPS) { set_error("umnn_cc_tables: nb_steps=%d out of range", Q); return UMNN_ERR_DESC; }
    // w_i = sum_{k even} m_k * c_{k,i},  c_{k,i} = (2/Q) cos(k i pi / Q) with column 0 -> 1/Q (cos replaced by
    // 1/2) and column Q halved; m_0 = 1, m_k = 2/(1-k^2).          ParallelNeuralIntegral.py:19-30
    const double pi = 3.14159265358979323846;
    for (int i = 0; i <= Q; ++i) {
        double acc = 0.0;
        for (int k = 0; k <= Q; k += 2) {
            double c;
            if (i == 0) c = 0.5;
            else c = cos((double)k * (double)i * pi / (double)Q);
            if (i == Q) c = 0.5 * c;
            c = c * 2.0 / (double)Q;
            const double m = (k == 0) ? 1.0 : 2.0 / (1.0 - (double)k * (double)k);
            acc += c * m;
        }
        weights_host[i] = (float)acc;
        nodes_host[i] = (float)cos((double)i * pi / (double)Q);
    }
    return 0;
}

int64_t umnn_param_count(const umnn_desc* d) {
    if (validate_desc(d) != 0) return -1;
    int64_t n = 0;
    for (int l = 0; l < d->n_layers; ++l) n += (int64_t)d->widths[l] * d->widths[l + 1] + d->widths[l + 1];
    return n;
}

size_t umnn_packed_params_bytes(const umnn_desc* d) {
    if (validate_desc(d) != 0) return 0;
    switch (resolve_precision(d)) {
        case UMNN_PREC_FP32: return sizeof(float) * (size_t)make_fp32_layout(d).total_floats;
        case UMNN_PREC_BF16X3:
            if (check_tc(d, "umnn_packed_params_bytes")) return 0;
            return tc_bf16_block_bytes(d);
        case UMNN_PREC_FP16X3:
            if (check_tc(d, "umnn_packed_params_bytes")) return 0;
            return tc_fp32_offset(d) + sizeof(float) * (size_t)make_fp32_layout(d).total_floats;
        default: return 0;
    }
}

uint64_t umnn_packed_layout_id(const umnn_desc* d) {
    if (validate_desc(d) != 0) return 0;
    const int prec = resolve_precision(d);
    // FNV-1a over everything that decides where a kernel looks inside the packed block
    uint64_t hsh = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { for (int i = 0; i < 8; ++i) { hsh ^= (v >> (8 * i)) & 0xffu; hsh *= 1099511628211ull; } };
    mix((uint64_t)prec);
    mix((uint64_t)umnn_packed_params_bytes(d));
    mix((uint64_t)d->layout);
    for (int l = 0; l <= d->n_layers; ++l) mix((uint64_t)d->widths[l]);
    if (is_tc(prec) && tc_unsupported_reason(d, 2) == nullptr) {
        mix((uint64_t)tc_two_segments_public());
        mix((uint64_t)tc_bf16_block_bytes(d));
        mix((uint64_t)tc_fp16_offset(d));
        mix((uint64_t)tc_fp32_offset(d));
        mix((uint64_t)(backward_tc_unsupported_reason(d) == nullptr));
        mix((uint64_t)bwd_rerun_is_fp32(d));
    }
    return hsh ? hsh : 1;
}

int umnn_pack_params(const umnn_desc* d, const float* flat_params, void* params_packed, void* stream) {
    int rc = validate_desc(d);
    if (rc) return rc;
    if (!flat_params || !params_packed) { set_error("umnn_pack_params: NULL pointer"); return UMNN_ERR_NULL; }
    const int prec = resolve_precision(d);
    switch (prec) {
        case UMNN_PREC_FP32:
            return launch_pack_fp32(d, flat_params, (float*)params_packed, (cudaStream_t)stream);
        case UMNN_PREC_BF16X3:
        case UMNN_PREC_FP16X3:
            if ((rc = check_tc(d, "umnn_pack_params")) != 0) return rc;
            if (backward_tc_unsupported_reason(d)) {
                // forward only: FP16X3 never reads the bf16 forward blobs (its re-run is the FP32 kernel)
                rc = prec == UMNN_PREC_BF16X3 ? launch_pack_tc(d, flat_params, params_packed, UMNN_OPF_BF16, (cudaStream_t)stream) : 0;
            } else {
                const bool with_forward = prec == UMNN_PREC_BF16X3 || !bwd_rerun_is_fp32(d);
                rc = launch_pack_backward_tc(d, flat_params, params_packed, with_forward, (cudaStream_t)stream);
            }
            if (rc || prec == UMNN_PREC_BF16X3) return rc;
            rc = launch_pack_tc(d, flat_params, (uint8_t*)params_packed + tc_fp16_offset(d), UMNN_OPF_FP16, (cudaStream_t)stream);
            if (rc) return rc;
            return launch_pack_fp32(d, flat_params, (float*)((uint8_t*)params_packed + tc_fp32_offset(d)), (cudaStream_t)stream);
        default:
            set_error("umnn_pack_params: precision %d is not available for this shape", d->precision);
            return UMNN_ERR_UNSUPPORTED;
    }
}

size_t umnn_workspace_bytes(const umnn_desc* d, int32_t for_backward) {
    if (validate_desc(d) != 0) return 0;
    const int prec = resolve_precision(d);
    if (!for_backward) return prec == UMNN_PREC_FP16X3 ? kFlagBytes : 0;
    if (is_tc(prec)) {
        const char* why = check_tc(d, "umnn_workspace_bytes") ? "forward shape unsupported" : backward_tc_unsupported_reason(d);
        if (why) {
            set_error("umnn_workspace_bytes: tensor-core backward unavailable for this shape (%s)", why);
            return 0;
        }
        const size_t tc = backward_tc_workspace_bytes(d);
        if (prec != UMNN_PREC_FP16X3) return tc;
        // [flag | panels of the tensor-core passes, reused by the guarded FP32 re-run]
        size_t body = tc;
        if (bwd_rerun_is_fp32(d)) {
            const size_t f = backward_fp32_workspace_bytes(d, tc);
            if (f > body) body = f;
        }
        return kFlagBytes + body;
    }
    if (backward_fp32_unsupported_reason(d)) {
        set_error("umnn_workspace_bytes: backward unavailable for this shape (%s)", backward_fp32_unsupported_reason(d));
        return 0;
    }
    return backward_fp32_workspace_bytes(d);
}

int umnn_tc_forward_occupancy(const umnn_desc* d, int32_t extra_rows, int32_t* narrow_shape, int32_t* ctas_per_sm) {
    int rc = validate_desc(d);
    if (rc) return rc;
    int narrow = 0, n = 0;
    rc = tc_forward_occupancy(d, extra_rows, &narrow, ctas_per_sm ? &n : nullptr);
    if (rc) return rc;
    if (narrow_shape) *narrow_shape = narrow;
    if (ctas_per_sm) *ctas_per_sm = n;
    return 0;
}

int umnn_cc_forward(const umnn_desc* d, const float* x0, const float* x, const float* h, const void* params_packed,
                    const float* nodes, const float* weights, float* out_integral, float* out_f_at_x,
                    float* out_f_at_x0, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = validate_desc(d);
    if (rc) return rc;
    if (d->n_samples == 0) return 0;
    if (!x || !params_packed || !nodes || !weights || !out_integral || (d->n_ctx > 0 && !h)) {
        set_error("umnn_cc_forward: required pointer is NULL"); return UMNN_ERR_NULL;
    }
    switch (resolve_precision(d)) {
        case UMNN_PREC_FP32:
            return launch_forward_fp32(d, x0, x, h, (const float*)params_packed, nodes, weights, out_integral,
                                       out_f_at_x, out_f_at_x0, nullptr, 0, (cudaStream_t)stream);
        case UMNN_PREC_BF16X3:
            if ((rc = check_tc(d, "umnn_cc_forward")) != 0) return rc;
            return launch_forward_tc(d, x0, x, h, params_packed, nodes, weights, out_integral, out_f_at_x,
                                     out_f_at_x0, UMNN_OPF_BF16, nullptr, nullptr, 0, (cudaStream_t)stream);
        case UMNN_PREC_FP16X3: {
            // fp16 hi/lo operands carry 22 bits (bf16: ~17) but overflow above 65504.  Guarded: the kernel raises a
            // device flag when an activation overflowed (every downstream value is NaN then), and a second launch --
            // the FP32 FFMA kernel, a no-op while the flag is clear -- recomputes the call, so the rare overflow case
            // gets the parity anchor's arithmetic, not a coarser split.  Without a workspace the fp16 launch runs
            // unguarded (overflow surfaces as NaN).
            if ((rc = check_tc(d, "umnn_cc_forward")) != 0) return rc;
            const uint8_t* fp16_blobs = (const uint8_t*)params_packed + tc_fp16_offset(d);
            int* flag = nullptr;
            int epoch = 1;
            if (workspace && workspace_bytes >= sizeof(int)) {
                flag = (int*)workspace;
                // The flag word is never reset between calls: the fp16 kernel raises it by writing THIS call's mark
                // and the re-run launch runs only if it reads that mark back, so whatever an earlier call (or the
                // allocator) left in the word is ignored -- one stream operation less per call.  A chance match with
                // uninitialised memory (2^-31) costs a redundant FP32 re-run, never a wrong result.  Under CUDA-graph
                // capture the mark would be frozen into the graph, so captured sequences clear the word first.
                cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
                UMNN_CUDA_TRY(cudaStreamIsCapturing((cudaStream_t)stream, &cap));
                if (cap != cudaStreamCaptureStatusNone) UMNN_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), (cudaStream_t)stream));
                else epoch = next_epoch();
            }
            rc = launch_forward_tc(d, x0, x, h, fp16_blobs, nodes, weights, out_integral, out_f_at_x, out_f_at_x0,
                                   UMNN_OPF_FP16, nullptr, flag, epoch, (cudaStream_t)stream);
            if (rc || !flag) return rc;
            return launch_forward_fp32(d, x0, x, h, (const float*)((const uint8_t*)params_packed + tc_fp32_offset(d)), nodes,
                                       weights, out_integral, out_f_at_x, out_f_at_x0, flag, epoch, (cudaStream_t)stream);
        }
        default:
            set_error("umnn_cc_forward: precision %d is not available for this shape", d->precision);
            return UMNN_ERR_UNSUPPORTED;
    }
}

int umnn_cc_backward(const umnn_desc* d, const float* x0, const float* x, const float* h, const void* params_packed,
                     const float* nodes, const float* weights, const float* grad_out, const float* grad_f_at_x,
                     float* d_x0, float* d_x, float* d_h, float* d_params, void* workspace, size_t workspace_bytes,
                     void* stream) {
    int rc = validate_desc(d);
    if (rc) return rc;
    const int prec = resolve_precision(d);
    if (is_tc(prec)) {
        if ((rc = check_tc(d, "umnn_cc_backward")) != 0) return rc;
        if (backward_tc_unsupported_reason(d)) {
            set_error("umnn_cc_backward: tensor-core backward unavailable for this shape (%s); use UMNN_PREC_FP32",
                      backward_tc_unsupported_reason(d));
            return UMNN_ERR_UNSUPPORTED;
        }
    }
    if (d->n_samples == 0) {
        if (d_params) UMNN_CUDA_TRY(cudaMemsetAsync(d_params, 0, sizeof(float) * (size_t)umnn_param_count(d), (cudaStream_t)stream));
        return 0;
    }
    if (!x || !params_packed || !nodes || !weights || !grad_out || (d->n_ctx > 0 && !h)) {
        set_error("umnn_cc_backward: required pointer is NULL"); return UMNN_ERR_NULL;
    }
    if (prec == UMNN_PREC_FP16X3) {
        if (!workspace || workspace_bytes < umnn_workspace_bytes(d, 1)) {
            set_error("umnn_cc_backward: workspace of %zu bytes needed, %zu given", umnn_workspace_bytes(d, 1), workspace_bytes);
            return UMNN_ERR_WORKSPACE;
        }
        uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
        const float* fp32_block = bwd_rerun_is_fp32(d) ? (const float*)((const uint8_t*)params_packed + tc_fp32_offset(d)) : nullptr;
        return launch_backward_tc(d, x0, x, h, params_packed, nodes, weights, grad_out, grad_f_at_x, d_x0, d_x, d_h, d_params,
                                  ws + kFlagBytes, workspace_bytes - kFlagBytes, (const uint8_t*)params_packed + tc_fp16_offset(d),
                                  reinterpret_cast<int*>(ws), fp32_block, (cudaStream_t)stream);
    }
    if (is_tc(prec))
        return launch_backward_tc(d, x0, x, h, params_packed, nodes, weights, grad_out, grad_f_at_x, d_x0, d_x, d_h, d_params,
                                  workspace, workspace_bytes, nullptr, nullptr, nullptr, (cudaStream_t)stream);
    return launch_backward_fp32(d, x0, x, h, (const float*)params_packed, nodes, weights, grad_out, grad_f_at_x, d_x0, d_x,
                                d_h, d_params, workspace, workspace_bytes, (cudaStream_t)stream);
}

int umnn_cc_forward_host(const umnn_desc* d, const float* x0_host, const float* x_host, const float* h_host,
                         const float* flat_params_host, float* out_integral_host, float* out_f_at_x_host,
                         float* out_f_at_x0_host, int32_t device) {
    int rc = validate_desc(d);
    if (rc) return rc;
    if (d->n_samples == 0) return 0;
    if (!x_host || !flat_params_host || !out_integral_host || (d->n_ctx > 0 && !h_host)) {
        set_error("umnn_cc_forward_host: required pointer is NULL"); return UMNN_ERR_NULL;
    }
    UMNN_CUDA_TRY(cudaSetDevice(device));
    const size_t n_slots = (size_t)d->n_samples * d->n_dims;
    const size_t xb = n_slots * sizeof(float);
    const size_t hb = (size_t)d->n_samples * d->n_dims * d->n_ctx * sizeof(float);
    const size_t pb = (size_t)umnn_param_count(d) * sizeof(float);
    const size_t packed_b = umnn_packed_params_bytes(d);
    const size_t ws_b = umnn_workspace_bytes(d, 0);
    const int Q = d->nb_steps;
    std::vector<float> tabs(2 * (Q + 1));
    rc = umnn_cc_tables(Q, tabs.data(), tabs.data() + Q + 1);
    if (rc) return rc;

    // one slab: x0 | x | h | flat | packed | tables | out | fx | fx0   (each 256-byte aligned)
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t o_x0 = 0, o_x = o_x0 + al(xb), o_h = o_x + al(xb), o_p = o_h + al(hb), o_pk = o_p + al(pb),
                 o_t = o_pk + al(packed_b), o_out = o_t + al(tabs.size() * sizeof(float)), o_fx = o_out + al(xb),
                 o_fx0 = o_fx + al(xb), o_ws = o_fx0 + al(xb), total = o_ws + al(ws_b);
    char* slab = nullptr;
    UMNN_CUDA_TRY(cudaMalloc((void**)&slab, total));
    cudaStream_t s = 0;
    auto fail = [&](int code) { cudaFree(slab); return code; };
#define UMNN_HOST_TRY(expr)                                                       \
    do {                                                                          \
        cudaError_t _e = (expr);                                                  \
        if (_e != cudaSuccess) return fail(::umnn::cuda_fail(_e, #expr));         \
    } while (0)
    if (x0_host) UMNN_HOST_TRY(cudaMemcpyAsync(slab + o_x0, x0_host, xb, cudaMemcpyHostToDevice, s));
    UMNN_HOST_TRY(cudaMemcpyAsync(slab + o_x, x_host, xb, cudaMemcpyHostToDevice, s));
    if (hb) UMNN_HOST_TRY(cudaMemcpyAsync(slab + o_h, h_host, hb, cudaMemcpyHostToDevice, s));
    UMNN_HOST_TRY(cudaMemcpyAsync(slab + o_p, flat_params_host, pb, cudaMemcpyHostToDevice, s));
    UMNN_HOST_TRY(cudaMemcpyAsync(slab + o_t, tabs.data(), tabs.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = umnn_pack_params(d, (const float*)(slab + o_p), slab + o_pk, s);
    if (rc) return fail(rc);
    rc = umnn_cc_forward(d, x0_host ? (const float*)(slab + o_x0) : nullptr, (const float*)(slab + o_x),
                         (const float*)(slab + o_h), slab + o_pk, (const float*)(slab + o_t),
                         (const float*)(slab + o_t) + Q + 1, (float*)(slab + o_out),
                         out_f_at_x_host ? (float*)(slab + o_fx) : nullptr,
                         out_f_at_x0_host ? (float*)(slab + o_fx0) : nullptr, ws_b ? slab + o_ws : nullptr, ws_b, s);
    if (rc) return fail(rc);
    UMNN_HOST_TRY(cudaMemcpyAsync(out_integral_host, slab + o_out, xb, cudaMemcpyDeviceToHost, s));
    if (out_f_at_x_host) UMNN_HOST_TRY(cudaMemcpyAsync(out_f_at_x_host, slab + o_fx, xb, cudaMemcpyDeviceToHost, s));
    if (out_f_at_x0_host) UMNN_HOST_TRY(cudaMemcpyAsync(out_f_at_x0_host, slab + o_fx0, xb, cudaMemcpyDeviceToHost, s));
    UMNN_HOST_TRY(cudaStreamSynchronize(s));
#undef UMNN_HOST_TRY
    cudaFree(slab);
    return 0;
}

}  // extern "C"
