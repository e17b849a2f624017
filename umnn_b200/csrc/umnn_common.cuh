// Shared host/device helpers for the umnn_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/umnn_b200.h"

namespace umnn {

constexpr float kLeakySlope = 0.01f;  // nn.LeakyReLU() default, models/UMNN/UMNNMAF.py:250

// ---------------------------------------------------------------------------------------------
// error reporting (thread-local string behind umnn_last_error)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define UMNN_CUDA_TRY(expr)                                           \
    do {                                                              \
        cudaError_t _e = (expr);                                      \
        if (_e != cudaSuccess) return ::umnn::cuda_fail(_e, #expr);   \
    } while (0)

int validate_desc(const umnn_desc* d);

// ---------------------------------------------------------------------------------------------
// Host-side memo of what the CUDA runtime would otherwise be asked on every launch (umnn_abi.cu).  A small call
// (config 1: 5 100 rows, ~15 us on the GPU) spent more host time on cudaFuncSetAttribute, the occupancy query and
// the device attribute lookups than on its two launches.
// ---------------------------------------------------------------------------------------------
// current device and its SM count
cudaError_t current_device(int* dev, int* n_sm);
// opt `kern` into `bytes` of dynamic shared memory on the current device unless it already is (monotonic)
cudaError_t ensure_dynamic_smem(const void* kern, int dev, int bytes);
// prefer the largest shared-memory carveout for `kern` (once per kernel and device)
cudaError_t ensure_max_carveout(const void* kern, int dev);
// cudaOccupancyMaxActiveBlocksPerMultiprocessor, memoised per (kernel, device, threads, smem)
cudaError_t cached_occupancy(int* occ, const void* kern, int dev, int threads, size_t smem);

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// FP32 packed parameter block (UMNN_PREC_FP32)
//   hidden layer l (l < L-1):  Wt[kpad_l][npad_l]  (Wt[k][j] = W_l[j][k], zero padded), bias[npad_l]
//   output layer    (l = L-1): w[kpad]  (zero padded), bias[4] (bias in [0])
//   kpad = round_up(n_in, 16), npad = round_up(n_out, 8); all offsets are multiples of 4 floats.
// ---------------------------------------------------------------------------------------------
constexpr int kFp32KChunk = 16;
constexpr int kFp32UnitsPerThread = 8;

struct Fp32Layout {
    int n_layers;
    int nin[UMNN_MAX_LAYERS], nout[UMNN_MAX_LAYERS], kpad[UMNN_MAX_LAYERS], npad[UMNN_MAX_LAYERS];
    int w_off[UMNN_MAX_LAYERS], b_off[UMNN_MAX_LAYERS];  // in floats
    int src_w_off[UMNN_MAX_LAYERS], src_b_off[UMNN_MAX_LAYERS];  // offsets into the flat vector
    // backward (dgrad) copies of the hidden layers: Wd_l[n16][k8] = W_l[n][k] (untransposed), zero padded
    // to n16 = round_up(n_out, 16) rows and k8 = round_up(n_in, 8) columns
    int d_off[UMNN_MAX_LAYERS], n16[UMNN_MAX_LAYERS], k8[UMNN_MAX_LAYERS];
    int total_floats;
    int max_kpad;   // rows of the activation buffer
    int max_npad;   // widest padded hidden layer
};

inline Fp32Layout make_fp32_layout(const umnn_desc* d) {
    Fp32Layout L{};
    L.n_layers = d->n_layers;
    int off = 0, src = 0;
    L.max_kpad = 0;
    L.max_npad = 8;
    for (int l = 0; l < d->n_layers; ++l) {
        const int nin = d->widths[l], nout = d->widths[l + 1];
        L.nin[l] = nin;
        L.nout[l] = nout;
        L.kpad[l] = round_up(nin, kFp32KChunk);
        if (L.kpad[l] > L.max_kpad) L.max_kpad = L.kpad[l];
        L.src_w_off[l] = src;
        src += nin * nout;
        L.src_b_off[l] = src;
        src += nout;
        if (l < d->n_layers - 1) {
            L.npad[l] = round_up(nout, kFp32UnitsPerThread);
            if (L.npad[l] > L.max_npad) L.max_npad = L.npad[l];
            L.w_off[l] = off;
            off += L.kpad[l] * L.npad[l];
            L.b_off[l] = off;
            off += L.npad[l];
        } else {
            L.npad[l] = 1;
            L.w_off[l] = off;
            off += L.kpad[l];
            L.b_off[l] = off;
            off += 4;
        }
    }
    for (int l = 0; l < d->n_layers - 1; ++l) {
        L.n16[l] = round_up(L.nout[l], kFp32KChunk);
        L.k8[l] = round_up(L.nin[l], kFp32UnitsPerThread);
        L.d_off[l] = off;
        off += L.n16[l] * L.k8[l];
    }
    L.total_floats = off;
    return L;
}

// ---------------------------------------------------------------------------------------------
// activations (device)
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float hidden_act(float v, int kind) {
    // UMNN_ACT_LEAKY_RELU: max(v, 0.01 v) for slope < 1;  UMNN_ACT_RELU: max(v, 0)
    return kind == UMNN_ACT_LEAKY_RELU ? fmaxf(v, v * kLeakySlope) : fmaxf(v, 0.0f);
}

__device__ __forceinline__ float out_act(float v, int kind) {
    if (kind == UMNN_OUT_ELU_PLUS_1) {
        // nn.ELU(alpha=1)(v) + 1.  in fp32: expm1(v) + 1 on the negative branch (quantised at 2^-24,
        // exactly 0 below v ~ -17.3), v + 1 on the positive one.  UMNNMAF.py:11-16, MonotonicNN.py:23-27
        return (v > 0.0f ? v : expm1f(v)) + 1.0f;
    }
    return 1.0f / (1.0f + expf(-v));
}

// Abscissa of node i exactly as the reference evaluates it in fp32 (no FMA contraction):
//   X_i = x0 + ((xT - x0) * (t_i + 1)) / 2            ParallelNeuralIntegral.py:55
__device__ __forceinline__ float node_abscissa(float x0, float span, float t) {
    return __fadd_rn(x0, __fmul_rn(__fmul_rn(span, __fadd_rn(t, 1.0f)), 0.5f));
}

// xT = x0 + Q * ((x - x0) / Q)                         ParallelNeuralIntegral.py:49,102
__device__ __forceinline__ float upper_limit(float x0, float x, int Q) {
    const float q = (float)Q;
    return __fadd_rn(x0, __fmul_rn(q, __fdiv_rn(__fsub_rn(x, x0), q)));
}
#endif

}  // namespace umnn

// launchers implemented in the kernel translation units
namespace umnn {
int launch_pack_fp32(const umnn_desc* d, const float* flat, float* packed, cudaStream_t s);
// run_if != NULL: every kernel of the launch is a no-op unless *run_if == run_epoch (the guarded FP32 re-run of an
// FP16X3 call whose activations left the fp16 range, see umnn_cc_forward / umnn_cc_backward)
int launch_forward_fp32(const umnn_desc* d, const float* x0, const float* x, const float* h, const float* packed,
                        const float* nodes, const float* weights, float* out, float* out_fx, float* out_fx0,
                        const int* run_if, int run_epoch, cudaStream_t s);
// fused FP32 backward (cc_backward_fp32.cu).  budget_bytes > 0: size the chunks to fill a workspace of that many
// bytes (fewer launches; used by the guarded re-run, which borrows the tensor-core backward's workspace) instead of
// the L2-sized default.
size_t backward_fp32_workspace_bytes(const umnn_desc* d, size_t budget_bytes = 0);
const char* backward_fp32_unsupported_reason(const umnn_desc* d);
int launch_backward_fp32(const umnn_desc* d, const float* x0, const float* x, const float* h, const float* packed,
                         const float* nodes, const float* weights, const float* grad_out, const float* grad_fx,
                         float* d_x0, float* d_x, float* d_h, float* d_params, void* workspace, size_t workspace_bytes,
                         cudaStream_t s, const int* run_if = nullptr, size_t budget_bytes = 0);
}  // namespace umnn
