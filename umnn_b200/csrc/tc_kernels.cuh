// Parameter blocks shared by the tensor-core kernels (forward / pass F, pass D, pass W).
#pragma once

#include "tc_bwd_layout.cuh"
#include "tc_layout.cuh"

namespace umnn {

// panels written by pass F (EMIT) for one chunk of rows; index j = hidden layer (0 = network input)
struct TcEmit {
    uint8_t* a[UMNN_MAX_LAYERS];       // A_j panels (bf16 hi [and lo] per block, see tc_bwd_layout.cuh)
    uint32_t* mask[UMNN_MAX_LAYERS];   // [R_pad][8] sign bits per 32-column pair, j = 1..J
    float* v;                          // [R_pad] pre-output-activation
    int width[UMNN_MAX_LAYERS];        // panel widths P_j
    long long row_block;               // padded rows per CTA (tiles_per_cta * 128)
    int parts;                         // 1: hi only, 2: hi + lo
    long long r_pad;                   // padded rows of the chunk's panels (n_cta * row_block)
};

struct TcParams {
    const float* x0;
    const float* x;
    const float* h;
    const float* nodes;
    const float* weights;
    const uint8_t* blobs;  // [2][blob_bytes]
    float* out;
    float* out_fx;
    float* out_fx0;
    long long slot0;       // first slot served by this launch (chunked launches of the backward)
    long long n_slots;     // slots served by this launch
    long long slots_per_cta;
    int tiles_per_cta;
    int D, E, layout, Q, rps, out_act;
    int x_row;             // 1: the first extra row of a slot (node Q+1) is evaluated at x, else at x0
    const int* run_if;     // not NULL: the whole launch is a no-op unless *run_if == epoch (guarded re-run)
    int* raise_flag;       // not NULL (fp16 operands): set to `epoch` when an activation overflowed the fp16 range
    int epoch;             // value that marks THIS call in the flag word (see umnn_cc_forward): no reset needed
    TcLayout L;
    TcSmem S;
    TcEmit emit;
};

int launch_forward_tc_emit(const umnn_desc* d, const float* x0, const float* x, const float* h, const void* packed,
                           const float* nodes, const float* weights, long long slot0, long long n_slots_chunk,
                           long long slots_per_cta, int tiles_per_cta, int n_cta, const TcEmit& emit, int opf,
                           const int* run_if, int* raise_flag, int epoch, cudaStream_t s);
bool tc_two_segments_public();

// tensor-core backward (cc_backward_tc.cu)
const char* backward_tc_unsupported_reason(const umnn_desc* d);
size_t backward_tc_workspace_bytes(const umnn_desc* d);
size_t backward_tc_packed_bytes(const umnn_desc* d);      // bf16 forward blobs + dgrad blobs
// with_forward = false: only the dgrad blobs are written (an FP16X3 block whose re-run is the FP32 kernel never
// reads the bf16 forward blobs)
int launch_pack_backward_tc(const umnn_desc* d, const float* flat, void* packed, bool with_forward, cudaStream_t s);
// fwd_blobs_fp16 != NULL: pass F re-evaluates with fp16 operands and raises *flag on overflow; the guarded re-run is
// the FP32 backward on `rerun_fp32_packed` when that is given, else a second tensor-core sequence with bf16 operands
int launch_backward_tc(const umnn_desc* d, const float* x0, const float* x, const float* h, const void* packed,
                       const float* nodes, const float* weights, const float* grad_out, const float* grad_fx,
                       float* d_x0, float* d_x, float* d_h, float* d_params, void* workspace, size_t workspace_bytes,
                       const void* fwd_blobs_fp16, int* flag, const float* rerun_fp32_packed, cudaStream_t s);

}  // namespace umnn
