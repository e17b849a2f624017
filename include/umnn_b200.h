/*
 * umnn_b200 -- C ABI of the B200-native Clenshaw-Curtis integration hot path of UMNN.
 *
 * This header is the drop-in boundary.  The reference (AWehenkel/UMNN @ 59118c14) is pure
 * Python/PyTorch and has no FFI of its own; each entry point below names the reference
 * function whose work it replaces, and INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add at those call sites.
 *
 * Conventions
 *   - plain C, no torch types.  All `const float*` / `float*` arguments of the device entry points
 *     are DEVICE pointers owned by the caller; nothing is allocated or freed inside; work is
 *     enqueued on `stream` (a cudaStream_t passed as void*) and is asynchronous with respect to
 *     the host.  The *_host entry points take HOST pointers and are synchronous.
 *   - return value: 0 = ok; negative = argument error (UMNN_ERR_*); positive = cudaError_t.
 *     `umnn_last_error()` returns a thread-local human-readable message for the last failure.
 *   - re-entrant; no global mutable state except the thread-local error string and an atomic call counter.
 *   - all tensors are contiguous float32.  B = n_samples, Dx = n_dims, E = n_ctx, Q = nb_steps
 *     (Q+1 quadrature nodes).
 *
 * Data layouts (desc.layout)
 *   UMNN_LAYOUT_STRIDED_D  x0,x: [B][Dx]   h: [B][E][Dx]   slot (b,d) reads h[b][e][d], e=0..E-1
 *                          (IntegrandNetwork.forward, models/UMNN/UMNNMAF.py:263-284)
 *   UMNN_LAYOUT_CONTIG     x0,x: [B][1]    h: [B][E]       slot b reads h[b][0..E-1]; Dx must be 1
 *                          (IntegrandNN.forward, models/UMNN/MonotonicNN.py:26-27, and
 *                           IntegrandNetwork.independant_forward as used by UMNNMAF.invert,
 *                           models/UMNN/UMNNMAF.py:207-216)
 *
 * Parameters travel as ONE flat float32 vector in nn.Sequential order
 *   W1 (out x in, row-major), b1, W2, b2, ..., W_last (1 x H), b_last
 * i.e. exactly `_flatten(integrand.parameters())` (models/UMNN/ParallelNeuralIntegral.py:6-8).
 */
#ifndef UMNN_B200_H
#define UMNN_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define UMNN_API __attribute__((visibility("default")))
#else
#define UMNN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define UMNN_ABI_VERSION 1
#define UMNN_MAX_LAYERS 8       /* Linear layers, output layer included */
#define UMNN_MAX_WIDTH 256      /* widest hidden layer / widest input (1+E) */
#define UMNN_MAX_STEPS 1024     /* Q */

enum { UMNN_LAYOUT_STRIDED_D = 0, UMNN_LAYOUT_CONTIG = 1 };
enum { UMNN_ACT_RELU = 0, UMNN_ACT_LEAKY_RELU = 1 };           /* hidden; leaky slope 0.01 */
enum { UMNN_OUT_ELU_PLUS_1 = 0, UMNN_OUT_SIGMOID = 1 };        /* UMNNMAF.py:11-19 */
/*
 * Arithmetic of the hidden-to-hidden layers (layer 1, the output layer, activations and the quadrature sum are
 * fp32 on CUDA cores in every mode):
 *   UMNN_PREC_FP32    FFMA kernel, any shape within UMNN_MAX_*.
 *   UMNN_PREC_FP16X3  tcgen05 tensor cores, operands split into fp16 hi + lo (3 MMAs per product, fp32 accumulate,
 *                     22 bits per operand: integrals within 3e-6 of the fp32 reference on every tested network).
 *                     fp16 overflows above 65504: the kernel raises a device flag (head of the workspace) when an
 *                     activation leaves that range and a second launch -- the FP32 kernel, a no-op while the flag is
 *                     clear -- recomputes the call, so the rare overflow case gets the parity anchor's arithmetic and
 *                     the host never has to look.
 *   UMNN_PREC_AUTO    FP16X3 where the tensor-core kernel serves the shape (>= 2 hidden layers, widths <= 254,
 *                     parameters within shared memory), else FP32.  This is what the product uses.
 *   UMNN_PREC_BF16X3  DIAGNOSTIC ONLY -- the same scheme with bf16 hi + lo operands (~17 bits per operand, fp32
 *                     exponent range, no guard needed).  It meets the 1e-5 integral bar on default-initialised
 *                     networks (4e-7..8e-7) but NOT on trained-scale weights (1e-5..2.2e-5 measured): it is never
 *                     selected by AUTO and exists for A/B measurements of the operand formats.
 */
enum { UMNN_PREC_FP32 = 0, UMNN_PREC_BF16X3 = 1, UMNN_PREC_AUTO = 2, UMNN_PREC_FP16X3 = 3 };

enum {
    UMNN_ERR_NULL = -1,        /* required pointer is NULL */
    UMNN_ERR_DESC = -2,        /* descriptor out of range / inconsistent */
    UMNN_ERR_WORKSPACE = -3,   /* workspace too small */
    UMNN_ERR_UNSUPPORTED = -4, /* valid request this build cannot serve (e.g. BF16X3 on that shape) */
    UMNN_ERR_ABI = -5          /* desc.abi_version mismatch */
};

typedef struct umnn_desc {
    int32_t abi_version;                  /* must be UMNN_ABI_VERSION */
    int32_t layout;                       /* UMNN_LAYOUT_* */
    int64_t n_samples;                    /* B */
    int32_t n_dims;                       /* Dx */
    int32_t n_ctx;                        /* E  (widths[0] must be 1+E) */
    int32_t n_layers;                     /* number of Linear layers, 2..UMNN_MAX_LAYERS */
    int32_t widths[UMNN_MAX_LAYERS + 1];  /* 1+E, H1, ..., HL, 1 */
    int32_t hidden_act;                   /* UMNN_ACT_* */
    int32_t out_act;                      /* UMNN_OUT_* */
    int32_t nb_steps;                     /* Q >= 1 */
    int32_t precision;                    /* UMNN_PREC_* */
} umnn_desc;

/* ABI version of the loaded library (compare with UMNN_ABI_VERSION). */
UMNN_API int umnn_abi_version(void);

/* Thread-local message of the last failing call on this thread ("" if none). */
UMNN_API const char* umnn_last_error(void);

/*
 * Clenshaw-Curtis tables on the host: nodes t[i] = cos(i*pi/Q) (i=0 is +1, i=Q is -1) and weights
 * w[i], float64 arithmetic rounded to float32.
 * Replaces compute_cc_weights(nb_steps), models/UMNN/ParallelNeuralIntegral.py:14-34
 * (= NeuralIntegral.py:14-34, UMNNMAF._compute_cc_weights UMNNMAF.py:55-69).
 */
UMNN_API int umnn_cc_tables(int32_t nb_steps, float* nodes_host, float* weights_host);

/* Number of floats in the flat parameter vector described by desc (sum of in*out + out). */
UMNN_API int64_t umnn_param_count(const umnn_desc* desc);

/*
 * Size in bytes of the device-side packed parameter block for desc (depends on desc.precision),
 * and the packing itself (transposition, zero padding, BF16 hi/lo split for UMNN_PREC_BF16X3).
 * `flat_params` is the device vector described above.  Re-pack after every parameter update.
 */
UMNN_API size_t umnn_packed_params_bytes(const umnn_desc* desc);
/*
 * Identifier of the packed block's internal layout for desc (resolved precision, offsets of its parts): two
 * descriptors with the same id can share one packed block; 0 = invalid desc.  The layout depends on nb_steps
 * (through the shared-memory fit that UMNN_PREC_AUTO resolves with), so a cache of packed blocks must key on this.
 */
UMNN_API uint64_t umnn_packed_layout_id(const umnn_desc* desc);
UMNN_API int umnn_pack_params(const umnn_desc* desc, const float* flat_params, void* params_packed, void* stream);

/*
 * Scratch the forward (for_backward = 0) / backward (1) entry points need for desc (0 is possible).  The forward
 * needs 256 bytes for UMNN_PREC_FP16X3 (the overflow flag of the guarded re-run) and nothing otherwise; the FP16X3
 * backward's workspace starts with the same 256-byte flag block.
 */
UMNN_API size_t umnn_workspace_bytes(const umnn_desc* desc, int32_t for_backward);

/*
 * Forward: for every slot (b,d)
 *     xT       = x0 + Q*((x-x0)/Q)
 *     integral = ((sum_i w_i * f(x0 + ((xT-x0)*(t_i+1))/2, h_slot)) * (xT-x0)) / 2
 *     f_at_x   = f(x, h_slot)     (optional; the Jacobian point of UMNNMAF.compute_log_jac)
 *     f_at_x0  = f(x0, h_slot)    (optional; needed by the Leibniz backward)
 * in ONE kernel launch; the Q+1 nodes never leave the SM.
 * Replaces integrate(..., compute_grad=False), models/UMNN/ParallelNeuralIntegral.py:37-65 and
 * models/UMNN/NeuralIntegral.py:37-66, IntegrandNetwork.forward models/UMNN/UMNNMAF.py:263-284,
 * IntegrandNN.forward models/UMNN/MonotonicNN.py:26-27, and the point evaluation of
 * UMNNMAF.compute_log_jac models/UMNN/UMNNMAF.py:136-139.
 *   x0            [B][Dx] or NULL (= zeros)
 *   out_integral  [B][Dx]
 *   out_f_at_x    [B][Dx] or NULL;  out_f_at_x0 [B][Dx] or NULL
 *   nodes, weights  device tables of Q+1 floats (umnn_cc_tables / compute_cc_weights)
 *   workspace     umnn_workspace_bytes(desc, 0) bytes, or NULL (UMNN_PREC_FP16X3 then runs unguarded: an activation
 *                 beyond the fp16 range surfaces as NaN).  Its content on entry does not matter and the same buffer may
 *                 serve every call that is ordered on one stream: the overflow flag is marked with a per-call value
 *                 (a process-wide atomic counter, the library's only mutable global besides the error string)
 *                 instead of being cleared, so a guarded call is two launches and no memset.
 */
UMNN_API int umnn_cc_forward(const umnn_desc* desc, const float* x0, const float* x, const float* h,
                    const void* params_packed, const float* nodes, const float* weights,
                    float* out_integral, float* out_f_at_x, float* out_f_at_x0,
                    void* workspace, size_t workspace_bytes, void* stream);

/*
 * Diagnostic (ctas_per_sm != NULL needs a CUDA device; with NULL only the shape is reported): which shape of the tensor-core forward kernel serves desc with `extra_rows`
 * (0..2) point evaluations per slot -- narrow_shape = 1: two CTAs per SM with 128-column tensor-memory regions
 * (every padded hidden width <= 128), 0: one CTA per SM with 256-column regions -- and how many CTAs of it per SM
 * cudaOccupancyMaxActiveClusters accounts for (informational: for the narrow shape the calculator answers 1 while
 * the hardware co-schedules 2).  UMNN_ERR_UNSUPPORTED if the tensor-core kernel cannot serve
 * the shape.
 */
UMNN_API int umnn_tc_forward_occupancy(const umnn_desc* desc, int32_t extra_rows, int32_t* narrow_shape,
                              int32_t* ctas_per_sm);

/*
 * Backward (Leibniz rule): given grad_out = dL/d(integral) [B][Dx]
 *     d_x      =  f(x , h) * grad_out
 *     d_x0     = -f(x0, h) * grad_out
 *     d_params =  sum_{slots,i} grad_out*(xT-x0)/2 * w_i * df/dparams(X_i, h)
 *     d_h      =  sum_i         grad_out*(xT-x0)/2 * w_i * df/dh     (X_i, h)      (same layout as h)
 * Optional grad_f_at_x [B][Dx] (dL/d f_at_x, e.g. from log|df/dx|) is folded in as one more
 * evaluation point at x: it adds to d_params and d_h, and d_x gets grad_f_at_x * df/dx(x,h).
 * Replaces ParallelNeuralIntegral.backward models/UMNN/ParallelNeuralIntegral.py:110-123,
 * integrate(compute_grad=True) :66-80, computeIntegrand :83-94 (and NeuralIntegral.py:47-64,69-75,90-99).
 *   d_params [P] is OVERWRITTEN (not accumulated); any of d_x0, d_x, d_h, d_params may be NULL.
 * desc.precision selects the path (and must be the precision params_packed was packed with):
 * UMNN_PREC_BF16X3 / FP16X3 / AUTO = three tensor-core passes per chunk of rows (forward re-evaluation with operand
 * emission, dgrad with transposed weights, split-K weight-gradient GEMM; FP16X3 re-evaluates with fp16 operands
 * and, if an activation left the fp16 range, repeats the whole backward with the FP32 kernels -- or with bf16
 * operands for the few shapes the FP32 backward cannot hold in shared memory), UMNN_PREC_FP32 = fused FFMA
 * kernel + FFMA split-K GEMM.  workspace must hold umnn_workspace_bytes(desc, 1) bytes (operand panels of one chunk;
 * the batch is processed in chunks of whole slots).  A shape the tensor-core backward cannot serve returns
 * UMNN_ERR_UNSUPPORTED (retry with UMNN_PREC_FP32).  Deterministic (fixed reduction order).
 */
UMNN_API int umnn_cc_backward(const umnn_desc* desc, const float* x0, const float* x, const float* h,
                     const void* params_packed, const float* nodes, const float* weights,
                     const float* grad_out, const float* grad_f_at_x,
                     float* d_x0, float* d_x, float* d_h, float* d_params,
                     void* workspace, size_t workspace_bytes, void* stream);

/*
 * Host-buffer convenience entry (what a non-PyTorch caller binds): allocates device memory on
 * `device`, copies x0/x/h/flat_params in, builds the tables, packs, runs umnn_cc_forward, copies
 * the results back and frees.  Synchronous.  x0_host and the two f outputs may be NULL.
 */
UMNN_API int umnn_cc_forward_host(const umnn_desc* desc, const float* x0_host, const float* x_host,
                         const float* h_host, const float* flat_params_host, float* out_integral_host,
                         float* out_f_at_x_host, float* out_f_at_x0_host, int32_t device);

/*
 * Sampling direction: one round of the bracket refinement of UMNNMAF.invert for one dimension
 * (models/UMNN/UMNNMAF.py:210,213-231; the integral of :213 comes from umnn_cc_forward with
 * UMNN_LAYOUT_CONTIG over the n_grid * n_samples slots, grid-point major).
 *   integral, x_grid  [n_grid][n_samples]  integral from 0 to x_grid, and the grid it was taken on
 *   grid              [n_grid]             relative positions 0 .. 1 (torch.arange(0, 1 + step/2, step))
 *   offset, target    one float per sample with the given element strides (h[:, j] and z[:, j])
 *   scale             device pointer to exp(scaling[j])
 *   left, right       bracket per sample (element stride bracket_stride), updated in place (:229-230)
 *   x_grid_next       [n_grid][n_samples]  grid of the next round (:210); must not alias x_grid
 *   x_mid             closest grid point per sample (:231), stride x_mid_stride; may be NULL
 * With integral == NULL only x_grid_next is laid over the current bracket (first round).
 * Bit-identical to the reference's torch expressions, including its flat neighbour indexing (:226-227).
 */
UMNN_API int umnn_invert_bracket_step(int64_t n_samples, int32_t n_grid, const float* integral, const float* x_grid,
                             const float* grid, const float* offset, int64_t offset_stride, const float* scale,
                             const float* target, int64_t target_stride, float* left, float* right,
                             int64_t bracket_stride, float* x_grid_next, float* x_mid, int64_t x_mid_stride,
                             void* stream);

/*
 * Sampling direction, one whole dimension per call: everything UMNNMAF.invert does for dimension j after the
 * conditioner pass (models/UMNN/UMNNMAF.py:203-231) -- replicate the dimension's context over the grid, reset the
 * bracket to [init_left, init_right] (the reference starts every dimension at [-50, 50], :194-195), then n_iter rounds
 * of {umnn_cc_forward over the n_grid * n_samples grid slots, umnn_invert_bracket_step} -- enqueued by ONE call
 * (1 + 3 * n_iter launches) instead of ~25 host-side operations per round.
 *   desc        UMNN_LAYOUT_CONTIG descriptor of the grid integrals: n_samples MUST be n_grid * n_samples, n_dims 1
 *   h_cols      [n_samples][E]   context of dimension j (h[:, j::D]); column 0 is the offset z0
 *   grid        [n_grid] relative grid positions; scale: device pointer to exp(scaling[j])
 *   target      z[:, j] with element stride target_stride
 *   x_out       receives the closest grid point of the last round per sample, element stride x_out_stride
 *               (write straight into x[:, j])
 *   workspace   umnn_invert_workspace_bytes(desc, n_samples, n_grid) bytes
 */
UMNN_API size_t umnn_invert_workspace_bytes(const umnn_desc* desc, int64_t n_samples, int32_t n_grid);
UMNN_API int umnn_invert_dimension(const umnn_desc* desc, const void* params_packed, const float* nodes,
                          const float* weights, int64_t n_samples, int32_t n_grid, int32_t n_iter,
                          const float* h_cols, const float* grid, const float* scale, const float* target,
                          int64_t target_stride, float init_left, float init_right, float* x_out,
                          int64_t x_out_stride, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UMNN_B200_H */
