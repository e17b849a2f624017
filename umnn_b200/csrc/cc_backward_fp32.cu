// Fused Clenshaw-Curtis backward, FP32 FFMA path.
//
// Replaces ParallelNeuralIntegral.backward (models/UMNN/ParallelNeuralIntegral.py:110-123) with
// integrate(compute_grad=True) (:66-80) and computeIntegrand (:83-94): for every slot
//     d_x  =  f(x ,h) * g   (+ grad_f_at_x * df/dx(x,h) when a cotangent of the Jacobian point is given)
//     d_x0 = -f(x0,h) * g
//     d_h  =  sum over the slot's rows of  c_row * df/dh(X_row, h)
//     d_W  =  sum over all rows of         c_row * df/dW(X_row, h)
// with c_row = g * (xT-x0)/2 * w_i on the Q+1 node rows, grad_f_at_x on the row evaluated at x, 0 at x0.
//
// Nothing of size rows x width is kept between forward and backward: the kernel re-evaluates the network
// per 64-row tile (activations in shared memory), back-propagates through it in place, reduces d_h per
// slot (carried across tiles) and streams the (dz_l, a_{l-1}) operand panels of the weight gradient to an
// L2-sized scratch; a split-K FFMA GEMM turns the panels into dW/db partials, reduced in a fixed order
// (deterministic).  The batch is processed in chunks of whole slots so the scratch stays bounded.
#include "umnn_common.cuh"

namespace umnn {

namespace {

constexpr int kTR = 64;   // rows per tile
constexpr int kRG = 16;   // thread columns: 4 rows each
constexpr int kKC = kFp32KChunk;
constexpr int kUT = kFp32UnitsPerThread;
constexpr long long kChunkRowsTarget = 49152;
constexpr int kMaxSplit = 64;

struct BwdParams {
    const float *x0, *x, *h, *packed, *nodes, *weights, *grad_out, *grad_fx;
    float *d_x0, *d_x, *d_h;
    float* scratch;
    const int* run_if;      // not NULL: no-op unless *run_if != 0 (guarded re-run of an FP16X3 backward)
    long long slot0, n_slots_chunk, slots_per_cta, ld;
    int D, E, layout, Q, rps, n_layers, hidden_act, out_act;
    int nin[UMNN_MAX_LAYERS], nout[UMNN_MAX_LAYERS], kpad[UMNN_MAX_LAYERS], npad[UMNN_MAX_LAYERS];
    int w_off[UMNN_MAX_LAYERS], b_off[UMNN_MAX_LAYERS], d_off[UMNN_MAX_LAYERS], n16[UMNN_MAX_LAYERS], k8[UMNN_MAX_LAYERS];
    int act_off[UMNN_MAX_LAYERS + 1];                       // float offsets of the activation buffers in smem
    long long a_panel[UMNN_MAX_LAYERS], dz_panel[UMNN_MAX_LAYERS];  // float offsets of A_l [nin_l][ld], DZ_l [nout_l][ld]
    int act_floats, max_w;                                  // total activation floats; widest staged weight row
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

__device__ __forceinline__ float act_grad(float a, int kind) {
    // derivative of the hidden activation expressed through its OUTPUT a (a > 0 <=> pre-activation > 0)
    return a > 0.0f ? 1.0f : (kind == UMNN_ACT_LEAKY_RELU ? kLeakySlope : 0.0f);
}

// acc[4][8] = sum_k in[k][rows tx*4..+3] * W[k][units ty*8..+7] over `krows` reduction rows; W is a global
// [krows][ldw] row-major panel streamed through the two-stage shared buffer `wst` (stage stride wst_stride)
__device__ __forceinline__ void gemm_tile(const float* in, const float* Wg, int krows, int ldw, bool active, int tx, int ty,
                                          float* wst, int wst_stride, int tid, int nthr, float (&acc)[4][kUT]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < kUT; ++j) acc[i][j] = 0.0f;
    const int n_chunks = krows / kKC;
    const int chunk_f4 = kKC * ldw / 4;
    for (int i = tid; i < chunk_f4; i += nthr) cp_async16(wst + 4 * i, Wg + 4 * i);
    cp_async_commit();
    for (int c = 0; c < n_chunks; ++c) {
        cp_async_wait_all();
        __syncthreads();
        if (c + 1 < n_chunks) {
            float* dst = wst + ((c + 1) & 1) * wst_stride;
            const float* src = Wg + (size_t)(c + 1) * kKC * ldw;
            for (int i = tid; i < chunk_f4; i += nthr) cp_async16(dst + 4 * i, src + 4 * i);
            cp_async_commit();
        }
        if (active) {
            const float* wc = wst + (c & 1) * wst_stride + ty * kUT;
            const float* ac = in + (size_t)c * kKC * kTR + tx * 4;
#pragma unroll
            for (int kk = 0; kk < kKC; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(ac + kk * kTR);
                const float4 w0 = *reinterpret_cast<const float4*>(wc + kk * ldw);
                const float4 w1 = *reinterpret_cast<const float4*>(wc + kk * ldw + 4);
                const float a[4] = {a0.x, a0.y, a0.z, a0.w};
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < kUT; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
        }
    }
    __syncthreads();  // every thread is done with `in` and the stage buffers
}

__global__ void __launch_bounds__(512) cc_backward_fp32_kernel(const BwdParams p) {
    if (p.run_if != nullptr && *p.run_if == 0) return;
    extern __shared__ __align__(16) float smem[];
    float* act = smem;                                  // activation buffers, act_off[l]
    float* wst = act + p.act_floats;                    // [2][kKC][max_w]
    float* dvrow = wst + 2 * kKC * p.max_w;             // [kTR] cotangent of the pre-output-activation
    float* frow = dvrow + kTR;                          // [kTR] f of the row
    float* carry = frow + kTR;                          // [2][E]
    float* tab_t = carry + 2 * (p.E > 0 ? p.E : 1);
    float* tab_w = tab_t + (p.Q + 1);

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int tx = tid & (kRG - 1), ty = tid >> 4;
    const int L = p.n_layers;
    const int wst_stride = kKC * p.max_w;

    const long long rel_begin = (long long)blockIdx.x * p.slots_per_cta;
    long long rel_end = rel_begin + p.slots_per_cta;
    if (rel_end > p.n_slots_chunk) rel_end = p.n_slots_chunk;
    if (rel_begin >= rel_end) return;
    const long long slot_begin = p.slot0 + rel_begin;
    const long long n_rows = (rel_end - rel_begin) * p.rps;
    const long long rowbase = rel_begin * p.rps;        // first row of this CTA inside the chunk panels

    for (int i = tid; i <= p.Q; i += nthr) { tab_t[i] = p.nodes[i]; tab_w[i] = p.weights[i]; }
    for (int i = tid; i < p.act_floats; i += nthr) act[i] = 0.0f;
    for (int i = tid; i < 2 * p.E; i += nthr) carry[i] = 0.0f;
    __syncthreads();

    int tile = 0;
    for (long long row0 = 0; row0 < n_rows; row0 += kTR, ++tile) {
        // ---- input columns: act[0][k][r] = [x_row, h_slot]
        float* a0buf = act + p.act_off[0];
        for (int r = tid; r < kTR; r += nthr) {
            const long long row = row0 + r;
            if (row < n_rows) {
                const long long ls = row / p.rps;
                const int node = (int)(row - ls * p.rps);
                const long long slot = slot_begin + ls;
                const float lo = p.x0 ? p.x0[slot] : 0.0f;
                const float hi = p.x[slot];
                float xi;
                if (node <= p.Q) xi = node_abscissa(lo, __fsub_rn(upper_limit(lo, hi, p.Q), lo), tab_t[node]);
                else xi = (node == p.Q + 1) ? hi : lo;
                a0buf[r] = xi;
                const float* hp;
                int hs;
                if (p.layout == UMNN_LAYOUT_STRIDED_D) {
                    const long long n = slot / p.D;
                    hp = p.h + n * (long long)p.E * p.D + (slot - n * p.D);
                    hs = p.D;
                } else {
                    hp = p.h + slot * (long long)p.E;
                    hs = 1;
                }
                for (int e = 0; e < p.E; ++e) a0buf[(1 + e) * kTR + r] = __ldg(hp + (long long)e * hs);
            } else {
                for (int k = 0; k <= p.E; ++k) a0buf[k * kTR + r] = 0.0f;
            }
        }
        __syncthreads();

        // ---- forward through the hidden layers, keeping every activation
        for (int l = 0; l < L - 1; ++l) {
            float acc[4][kUT];
            const bool active = ty * kUT < p.npad[l];
            gemm_tile(act + p.act_off[l], p.packed + p.w_off[l], p.kpad[l], p.npad[l], active, tx, ty, wst, wst_stride, tid,
                      nthr, acc);
            if (active) {
                float* ob = act + p.act_off[l + 1];
                const float* bg = p.packed + p.b_off[l] + ty * kUT;
#pragma unroll
                for (int j = 0; j < kUT; ++j) {
                    const float b = __ldg(bg + j);
                    float4 o;
                    o.x = hidden_act(acc[0][j] + b, p.hidden_act);
                    o.y = hidden_act(acc[1][j] + b, p.hidden_act);
                    o.z = hidden_act(acc[2][j] + b, p.hidden_act);
                    o.w = hidden_act(acc[3][j] + b, p.hidden_act);
                    *reinterpret_cast<float4*>(ob + (size_t)(ty * kUT + j) * kTR + tx * 4) = o;
                }
            }
            __syncthreads();
        }

        // ---- output layer, f, cotangent of v
        {
            const int l = L - 1;
            const float* wl = p.packed + p.w_off[l];
            const float bl = __ldg(p.packed + p.b_off[l]);
            const float* ab = act + p.act_off[l];
            for (int r = tid; r < kTR; r += nthr) {
                const long long row = row0 + r;
                float dv = 0.0f, f = 0.0f;
                if (row < n_rows) {
                    float v = 0.0f;
                    for (int k = 0; k < p.nin[l]; ++k) v = fmaf(__ldg(wl + k), ab[k * kTR + r], v);
                    v += bl;
                    f = out_act(v, p.out_act);
                    const long long ls = row / p.rps;
                    const int node = (int)(row - ls * p.rps);
                    const long long slot = slot_begin + ls;
                    float c = 0.0f;
                    if (node <= p.Q) {
                        const float lo = p.x0 ? p.x0[slot] : 0.0f;
                        const float span = __fsub_rn(upper_limit(lo, p.x[slot], p.Q), lo);
                        c = __fmul_rn(__fmul_rn(__fmul_rn(p.grad_out[slot], span), 0.5f), tab_w[node]);   // :70-71
                    } else if (node == p.Q + 1 && p.grad_fx) {
                        c = p.grad_fx[slot];
                    }
                    float dact;
                    if (p.out_act == UMNN_OUT_ELU_PLUS_1) dact = v > 0.0f ? 1.0f : expf(v);
                    else dact = f * (1.0f - f);
                    dv = c * dact;
                }
                dvrow[r] = dv;
                frow[r] = f;
            }
        }
        __syncthreads();

        // ---- stream the layer inputs a_l (l = 0..L-1) and dz_{L-1} = dv to the scratch panels
        for (int l = 0; l < L; ++l) {
            const float* ab = act + p.act_off[l];
            float* pg = p.scratch + p.a_panel[l] + rowbase + row0;
            for (int idx = tid; idx < p.nin[l] * kTR; idx += nthr) {
                const int k = idx / kTR, r = idx - k * kTR;
                if (row0 + r < n_rows) pg[(size_t)k * p.ld + r] = ab[idx];
            }
        }
        for (int r = tid; r < kTR; r += nthr)
            if (row0 + r < n_rows) p.scratch[p.dz_panel[L - 1] + rowbase + row0 + r] = dvrow[r];
        __syncthreads();

        // ---- dz of the last hidden layer, in place over its activations
        {
            const int l = L - 1;
            float* ab = act + p.act_off[l];
            const float* wl = p.packed + p.w_off[l];
            for (int idx = tid; idx < p.nin[l] * kTR; idx += nthr) {
                const int k = idx / kTR, r = idx - k * kTR;
                ab[idx] = dvrow[r] * __ldg(wl + k) * act_grad(ab[idx], p.hidden_act);
            }
        }
        __syncthreads();

        // ---- back through the hidden layers: dz_l sits in act[l+1]; da_l = W_l^T dz_l; dz_{l-1} = da_l * act'(a_l)
        for (int l = L - 2; l >= 0; --l) {
            const float* dzb = act + p.act_off[l + 1];
            float* pg = p.scratch + p.dz_panel[l] + rowbase + row0;
            for (int idx = tid; idx < p.nout[l] * kTR; idx += nthr) {
                const int n = idx / kTR, r = idx - n * kTR;
                if (row0 + r < n_rows) pg[(size_t)n * p.ld + r] = dzb[idx];
            }
            float acc[4][kUT];
            const bool active = ty * kUT < p.k8[l];
            gemm_tile(dzb, p.packed + p.d_off[l], p.n16[l], p.k8[l], active, tx, ty, wst, wst_stride, tid, nthr, acc);
            if (active) {
                float* ob = act + p.act_off[l];
#pragma unroll
                for (int j = 0; j < kUT; ++j) {
                    float* dst = ob + (size_t)(ty * kUT + j) * kTR + tx * 4;
                    float4 o = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
                    if (l > 0) {
                        const float4 a = *reinterpret_cast<const float4*>(dst);
                        o.x *= act_grad(a.x, p.hidden_act);
                        o.y *= act_grad(a.y, p.hidden_act);
                        o.z *= act_grad(a.z, p.hidden_act);
                        o.w *= act_grad(a.w, p.hidden_act);
                    }
                    *reinterpret_cast<float4*>(dst) = o;
                }
            }
            __syncthreads();
        }

        // ---- act[0] now holds d f / d input (already weighted by c_row): context gradient per slot,
        //      Leibniz terms and the Jacobian-point term of d_x
        {
            const float* dab = act + p.act_off[0];
            const long long last_row = (row0 + kTR < n_rows ? row0 + kTR : n_rows) - 1;
            const long long s_first = row0 / p.rps, s_last = last_row / p.rps;
            const int ns = (int)(s_last - s_first) + 1;
            const float* cin = carry + (tile & 1) * p.E;
            float* cout = carry + ((tile + 1) & 1) * p.E;
            for (int idx = tid; idx < ns * p.E; idx += nthr) {
                const int i = idx / p.E, e = idx - i * p.E;
                const long long ls = s_first + i;
                const long long a = ls * p.rps, b = a + p.rps - 1;
                const long long lo = a > row0 ? a : row0;
                const long long hi = b < last_row ? b : last_row;
                float sum = (a < row0) ? cin[e] : 0.0f;
                const float* col = dab + (size_t)(1 + e) * kTR;
                for (long long rr = lo; rr <= hi; ++rr) sum += col[(int)(rr - row0)];
                if (b <= last_row) {
                    if (p.d_h) {
                        const long long slot = slot_begin + ls;
                        if (p.layout == UMNN_LAYOUT_STRIDED_D) {
                            const long long n = slot / p.D;
                            p.d_h[n * (long long)p.E * p.D + (long long)e * p.D + (slot - n * p.D)] = sum;
                        } else {
                            p.d_h[slot * (long long)p.E + e] = sum;
                        }
                    }
                } else {
                    cout[e] = sum;
                }
            }
            for (int r = tid; r < kTR; r += nthr) {
                const long long row = row0 + r;
                if (row >= n_rows) continue;
                const long long ls = row / p.rps;
                const int node = (int)(row - ls * p.rps);
                if (node <= p.Q) continue;
                const long long slot = slot_begin + ls;
                const float g = p.grad_out[slot];
                if (node == p.Q + 1) {
                    if (p.d_x) p.d_x[slot] = frow[r] * g + dab[r];        // :115,:123 (+ Jacobian-point term)
                } else {
                    if (p.d_x0) p.d_x0[slot] = -frow[r] * g;              // :116,:123
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: C[n][k] = sum_r DZ[n][r] * A[k][r]  (k == nin: bias, A == 1), split over row slabs
// ---------------------------------------------------------------------------------------------
constexpr int kWT = 128;  // output tile (n and k)
constexpr int kWK = 16;   // rows per shared-memory step
constexpr int kWP = kWT + 4;

// 256 threads, 8 x 8 outputs each (n: tn*4..+3 and 64+tn*4..+3, k likewise) so every smem read is a
// conflict-free float4; operands are transposed on the way in ([dim][rows] panels -> [r][dim] tiles)
// All Linear layers of a chunk in ONE launch (blockIdx.x walks the output tiles of every layer): as the guarded re-run
// of a tensor-core backward this sequence is a string of no-op launches (~2.7 us each), one per layer was 2/3 of them.
struct WgradLayers {
    int n_layers;
    int tile_base[UMNN_MAX_LAYERS + 1];        // first output tile of layer l in blockIdx.x
    int tiles_k[UMNN_MAX_LAYERS];              // tiles along k (inputs + bias column)
    int nout[UMNN_MAX_LAYERS], nin[UMNN_MAX_LAYERS], w_dst[UMNN_MAX_LAYERS], b_dst[UMNN_MAX_LAYERS];
    long long dz_off[UMNN_MAX_LAYERS], a_off[UMNN_MAX_LAYERS];   // panel offsets in the scratch (floats)
};

__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ scratch, const __grid_constant__ WgradLayers W, long long ld,
                                                    long long rows, long long slab, float* __restrict__ part,
                                                    long long part_stride, const int* __restrict__ run_if) {
    if (run_if != nullptr && *run_if == 0) return;
    __shared__ __align__(16) float As[2][kWK][kWP];   // [stage][r][n]
    __shared__ __align__(16) float Bs[2][kWK][kWP];   // [stage][r][k]
    const int tid = threadIdx.x;
    int layer = 0;
    while (layer + 1 < W.n_layers && (int)blockIdx.x >= W.tile_base[layer + 1]) ++layer;
    const int tile = (int)blockIdx.x - W.tile_base[layer];
    const int nout = W.nout[layer], nin = W.nin[layer], w_dst = W.w_dst[layer], b_dst = W.b_dst[layer];
    const float* __restrict__ dz = scratch + W.dz_off[layer];
    const float* __restrict__ a = scratch + W.a_off[layer];
    const int k0 = (tile % W.tiles_k[layer]) * kWT, n0 = (tile / W.tiles_k[layer]) * kWT;
    const long long r_begin = (long long)blockIdx.z * slab;
    long long r_end = r_begin + slab;
    if (r_end > rows) r_end = rows;
    const int tn = (tid >> 4) * 4, tk = (tid & 15) * 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    // each thread moves 2 float4 (along r) per operand per step: 128 dims x 16 rows = 512 float4
    auto load_tile = [&](long long r0, float4 (&va)[2], float4 (&vb)[2]) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int e = tid + it * 256;
            const int dim = e >> 2;              // 0..127
            const long long r = r0 + (e & 3) * 4;
            va[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            vb[it] = va[it];
            const int n = n0 + dim, k = k0 + dim;
            if (n < nout) {
                const float* src = dz + (size_t)n * ld + r;
                if (r + 3 < r_end) va[it] = *reinterpret_cast<const float4*>(src);
                else {
                    if (r < r_end) va[it].x = src[0];
                    if (r + 1 < r_end) va[it].y = src[1];
                    if (r + 2 < r_end) va[it].z = src[2];
                }
            }
            if (k < nin) {
                const float* src = a + (size_t)k * ld + r;
                if (r + 3 < r_end) vb[it] = *reinterpret_cast<const float4*>(src);
                else {
                    if (r < r_end) vb[it].x = src[0];
                    if (r + 1 < r_end) vb[it].y = src[1];
                    if (r + 2 < r_end) vb[it].z = src[2];
                }
            } else if (k == nin) {   // bias column: A == 1 on valid rows
                vb[it].x = (r < r_end) ? 1.f : 0.f;
                vb[it].y = (r + 1 < r_end) ? 1.f : 0.f;
                vb[it].z = (r + 2 < r_end) ? 1.f : 0.f;
                vb[it].w = (r + 3 < r_end) ? 1.f : 0.f;
            }
        }
    };
    auto store_tile = [&](int stage, const float4 (&va)[2], const float4 (&vb)[2]) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int e = tid + it * 256;
            const int dim = e >> 2, r4 = (e & 3) * 4;
            As[stage][r4 + 0][dim] = va[it].x; As[stage][r4 + 1][dim] = va[it].y;
            As[stage][r4 + 2][dim] = va[it].z; As[stage][r4 + 3][dim] = va[it].w;
            Bs[stage][r4 + 0][dim] = vb[it].x; Bs[stage][r4 + 1][dim] = vb[it].y;
            Bs[stage][r4 + 2][dim] = vb[it].z; Bs[stage][r4 + 3][dim] = vb[it].w;
        }
    };

    float4 va[2], vb[2];
    if (r_begin < r_end) {
        load_tile(r_begin, va, vb);
        store_tile(0, va, vb);
    }
    __syncthreads();
    int stage = 0;
    for (long long r0 = r_begin; r0 < r_end; r0 += kWK, stage ^= 1) {
        const bool more = r0 + kWK < r_end;
        if (more) load_tile(r0 + kWK, va, vb);      // global loads in flight during the FMAs below
#pragma unroll
        for (int r = 0; r < kWK; ++r) {
            const float4 x0 = *reinterpret_cast<const float4*>(&As[stage][r][tn]);
            const float4 x1 = *reinterpret_cast<const float4*>(&As[stage][r][64 + tn]);
            const float4 y0 = *reinterpret_cast<const float4*>(&Bs[stage][r][tk]);
            const float4 y1 = *reinterpret_cast<const float4*>(&Bs[stage][r][64 + tk]);
            const float xa[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            const float ya[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xa[i], ya[j], acc[i][j]);
        }
        if (more) store_tile(stage ^ 1, va, vb);
        __syncthreads();
    }
    float* out = part + (size_t)blockIdx.z * part_stride;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = n0 + (i < 4 ? tn + i : 64 + tn + i - 4);
        if (n >= nout) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + (j < 4 ? tk + j : 64 + tk + j - 4);
            if (k < nin) out[w_dst + (size_t)n * nin + k] = acc[i][j];
            else if (k == nin) out[b_dst + n] = acc[i][j];
        }
    }
}

__global__ void reduce_partials_kernel(const float* __restrict__ part, long long part_stride, int nsplit, long long P,
                                       float* __restrict__ d_params, int accumulate, const int* __restrict__ run_if) {
    if (run_if != nullptr && *run_if == 0) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float s = accumulate ? d_params[i] : 0.0f;
    for (int z = 0; z < nsplit; ++z) s += part[(size_t)z * part_stride + i];
    d_params[i] = s;
}

struct BwdPlan {
    Fp32Layout L;
    long long P;                 // parameter count
    long long floats_per_row;    // sum of (nin_l + nout_l)
    long long chunk_slots, chunk_rows, ld;
    int rps;
    size_t scratch_bytes, part_bytes, total_bytes;
    int act_off[UMNN_MAX_LAYERS + 1];
    int act_floats, max_w;
    size_t smem;
    int threads;
};

bool make_plan(const umnn_desc* d, BwdPlan* B, size_t budget_bytes = 0) {
    B->L = make_fp32_layout(d);
    const Fp32Layout& L = B->L;
    B->P = 0;
    B->floats_per_row = 0;
    for (int l = 0; l < d->n_layers; ++l) {
        B->P += (long long)L.nin[l] * L.nout[l] + L.nout[l];
        B->floats_per_row += L.nin[l] + L.nout[l];
    }
    B->rps = d->nb_steps + 3;
    const long long n_slots = d->n_samples * (long long)d->n_dims;
    // a chunk = n_cta x s whole slots, s chosen so that s*rps fills its 64-row tiles with the least padding
    const int n_cta = 148;
    long long rows_target = kChunkRowsTarget;
    if (budget_bytes > 0) {
        // chunks as large as the given workspace allows (never below the default)
        const size_t fixed = (size_t)kMaxSplit * B->P * sizeof(float) + 1024;
        if (budget_bytes > fixed) {
            const long long fit = (long long)((budget_bytes - fixed) / ((size_t)B->floats_per_row * sizeof(float))) - 8;
            if (fit > rows_target) rows_target = fit;
        }
        const long long all_rows = n_slots * B->rps;
        if (rows_target > all_rows + n_cta * (long long)B->rps) rows_target = all_rows + n_cta * (long long)B->rps;
    }
    const long long per_cta_target = rows_target / n_cta;
    long long s_max = per_cta_target / B->rps;
    if (s_max < 1) s_max = 1;
    long long best_s = s_max;
    double best_waste = 2.0;
    for (long long sc = s_max; sc >= 1 && sc * 2 >= s_max; --sc) {
        const long long rows_c = sc * B->rps;
        const double waste = (double)((rows_c + kTR - 1) / kTR * kTR - rows_c) / (double)rows_c;
        if (waste < best_waste - 1e-9) { best_waste = waste; best_s = sc; }
    }
    long long cs = best_s * n_cta;
    if (cs > n_slots) cs = n_slots > 0 ? n_slots : 1;
    B->chunk_slots = cs;
    B->chunk_rows = cs * B->rps;
    B->ld = (B->chunk_rows + 3) / 4 * 4;
    B->scratch_bytes = (size_t)B->floats_per_row * B->ld * sizeof(float);
    B->part_bytes = (size_t)kMaxSplit * B->P * sizeof(float);
    B->total_bytes = B->scratch_bytes + B->part_bytes + 256;
    int off = 0, max_w = 8;
    for (int l = 0; l < d->n_layers; ++l) {
        B->act_off[l] = off;
        off += L.kpad[l] * kTR;                 // act[l] holds the INPUT of layer l (kpad_l rows)
        if (l < d->n_layers - 1) {
            if (L.npad[l] > max_w) max_w = L.npad[l];
            if (L.k8[l] > max_w) max_w = L.k8[l];
        }
    }
    B->act_floats = off;
    B->max_w = max_w;
    int groups = 1;
    for (int l = 0; l < d->n_layers - 1; ++l) {
        if (L.npad[l] / kUT > groups) groups = L.npad[l] / kUT;
        if (L.k8[l] / kUT > groups) groups = L.k8[l] / kUT;
    }
    B->threads = round_up(kRG * groups, 32);
    B->smem = sizeof(float) * ((size_t)B->act_floats + 2 * kKC * max_w + 2 * kTR + 2 * (d->n_ctx > 0 ? d->n_ctx : 1) +
                               2 * (d->nb_steps + 1));
    return B->smem <= 232448 && B->threads <= 512;
}

}  // namespace

const char* backward_fp32_unsupported_reason(const umnn_desc* d) {
    BwdPlan B;
    if (!make_plan(d, &B)) return "activations of one 64-row tile do not fit in 227 KB of shared memory";
    return nullptr;
}

size_t backward_fp32_workspace_bytes(const umnn_desc* d, size_t budget_bytes) {
    BwdPlan B;
    if (!make_plan(d, &B, budget_bytes)) return 0;
    return B.total_bytes;
}

int launch_backward_fp32(const umnn_desc* d, const float* x0, const float* x, const float* h, const float* packed,
                         const float* nodes, const float* weights, const float* grad_out, const float* grad_fx,
                         float* d_x0, float* d_x, float* d_h, float* d_params, void* workspace, size_t workspace_bytes,
                         cudaStream_t s, const int* run_if, size_t budget_bytes) {
    BwdPlan B;
    if (!make_plan(d, &B, budget_bytes)) {
        set_error("umnn_cc_backward: %s", backward_fp32_unsupported_reason(d));
        return UMNN_ERR_UNSUPPORTED;
    }
    if (!workspace || workspace_bytes < B.total_bytes) {
        set_error("umnn_cc_backward: workspace of %zu bytes needed, %zu given", B.total_bytes, workspace_bytes);
        return UMNN_ERR_WORKSPACE;
    }
    const Fp32Layout& L = B.L;
    const long long n_slots = d->n_samples * (long long)d->n_dims;
    int dev = 0, n_sm = 0;
    UMNN_CUDA_TRY(current_device(&dev, &n_sm));
    UMNN_CUDA_TRY(ensure_dynamic_smem((const void*)cc_backward_fp32_kernel, dev, (int)B.smem));

    float* scratch = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    float* part = scratch + (size_t)B.floats_per_row * B.ld;

    BwdParams p{};
    p.x0 = x0; p.x = x; p.h = h; p.packed = packed; p.nodes = nodes; p.weights = weights;
    p.grad_out = grad_out; p.grad_fx = grad_fx; p.d_x0 = d_x0; p.d_x = d_x; p.d_h = d_h; p.scratch = scratch;
    p.run_if = run_if;
    p.D = d->n_dims; p.E = d->n_ctx; p.layout = d->layout; p.Q = d->nb_steps; p.rps = B.rps;
    p.n_layers = d->n_layers; p.hidden_act = d->hidden_act; p.out_act = d->out_act;
    p.ld = B.ld; p.act_floats = B.act_floats; p.max_w = B.max_w;
    long long off = 0;
    for (int l = 0; l < d->n_layers; ++l) {
        p.nin[l] = L.nin[l]; p.nout[l] = L.nout[l]; p.kpad[l] = L.kpad[l]; p.npad[l] = L.npad[l];
        p.w_off[l] = L.w_off[l]; p.b_off[l] = L.b_off[l]; p.d_off[l] = L.d_off[l]; p.n16[l] = L.n16[l]; p.k8[l] = L.k8[l];
        p.act_off[l] = B.act_off[l];
        p.a_panel[l] = off;  off += (long long)L.nin[l] * B.ld;
        p.dz_panel[l] = off; off += (long long)L.nout[l] * B.ld;
    }

    bool first = true;
    for (long long s0 = 0; s0 < n_slots; s0 += B.chunk_slots) {
        const long long cs = (n_slots - s0 < B.chunk_slots) ? (n_slots - s0) : B.chunk_slots;
        const long long rows = cs * B.rps;
        p.slot0 = s0;
        p.n_slots_chunk = cs;
        long long want = (rows + kTR - 1) / kTR;
        if (want > n_sm) want = n_sm;
        if (want > cs) want = cs;
        if (want < 1) want = 1;
        p.slots_per_cta = (cs + want - 1) / want;
        const long long grid = (cs + p.slots_per_cta - 1) / p.slots_per_cta;
        cc_backward_fp32_kernel<<<(unsigned)grid, B.threads, B.smem, s>>>(p);
        UMNN_CUDA_TRY(cudaGetLastError());
        if (d_params) {
            int nsplit = (int)((rows + 1023) / 1024);
            if (nsplit > kMaxSplit) nsplit = kMaxSplit;
            if (nsplit < 1) nsplit = 1;
            long long slab = (rows + nsplit - 1) / nsplit;
            slab = (slab + kWK - 1) / kWK * kWK;
            nsplit = (int)((rows + slab - 1) / slab);
            WgradLayers W{};
            W.n_layers = d->n_layers;
            int n_tiles = 0;
            for (int l = 0; l < d->n_layers; ++l) {
                W.tile_base[l] = n_tiles;
                W.tiles_k[l] = (L.nin[l] + 1 + kWT - 1) / kWT;
                n_tiles += W.tiles_k[l] * ((L.nout[l] + kWT - 1) / kWT);
                W.nout[l] = L.nout[l]; W.nin[l] = L.nin[l]; W.w_dst[l] = L.src_w_off[l]; W.b_dst[l] = L.src_b_off[l];
                W.dz_off[l] = p.dz_panel[l]; W.a_off[l] = p.a_panel[l];
            }
            W.tile_base[d->n_layers] = n_tiles;
            wgrad_kernel<<<dim3((unsigned)n_tiles, 1, (unsigned)nsplit), 256, 0, s>>>(scratch, W, B.ld, rows, slab, part, B.P, run_if);
            UMNN_CUDA_TRY(cudaGetLastError());
            reduce_partials_kernel<<<(unsigned)((B.P + 255) / 256), 256, 0, s>>>(part, B.P, nsplit, B.P, d_params, first ? 0 : 1, run_if);
            UMNN_CUDA_TRY(cudaGetLastError());
        }
        first = false;
    }
    if (n_slots == 0 && d_params) UMNN_CUDA_TRY(cudaMemsetAsync(d_params, 0, sizeof(float) * B.P, s));
    return 0;
}

}  // namespace umnn
