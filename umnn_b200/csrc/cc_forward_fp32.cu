// Fused Clenshaw-Curtis forward, FP32 FFMA path (UMNN_PREC_FP32) -- the parity anchor and the generic
// kernel for every integrand shape (any depth <= UMNN_MAX_LAYERS, any width <= UMNN_MAX_WIDTH).
//
// What one launch does (replaces integrate(), ParallelNeuralIntegral.py:37-65 + IntegrandNetwork.forward,
// UMNNMAF.py:263-284 + the Jacobian point of UMNNMAF.py:136-139):
//   rows = (slot, node) pairs, RPS = Q+1 (+1 for f(x)) (+1 for f(x0)) rows per slot, flattened.
//   A CTA owns a contiguous range of whole slots and walks its rows in tiles of 128.  Per tile:
//     1. build the input columns act[k][row] = [x_node, h_slot[0..E-1]]  (never materialised in HBM)
//     2. every hidden layer:  act <- act_fn(W_l act + b_l)   register tile 8 rows x 8 units per thread,
//        weights streamed L2 -> SMEM in 16-row chunks with cp.async double buffering,
//        activations stay in SMEM (k-major so row reads are conflict-free float4)
//     3. output layer (H -> 1) + ELU+1, weighted by the CC weight of the row's node
//     4. deterministic segmented sum over the rows of each slot (carried across tiles),
//        final scaling (z * (xT-x0)) / 2, store.
//
// HBM traffic = x0, x, h read once (+ L2-resident weights), 1-3 floats written per slot.
#include "umnn_common.cuh"

namespace umnn {

namespace {

constexpr int kTile = 128;                 // rows per tile
constexpr int kRowGroups = 16;             // 16 thread columns x 8 rows
constexpr int kKC = kFp32KChunk;           // weight rows per cp.async stage
constexpr int kUT = kFp32UnitsPerThread;   // units per thread

struct FwdParams {
    const float* x0;
    const float* x;
    const float* h;
    const float* packed;
    const float* nodes;
    const float* weights;
    float* out;
    float* out_fx;
    float* out_fx0;
    const int* run_if;     // not NULL: no-op unless *run_if == run_epoch (guarded re-run of an FP16X3 call)
    int run_epoch;
    long long n_slots;
    long long slots_per_cta;
    int D, E, layout, Q, rps, n_layers, hidden_act, out_act;
    int max_kpad, max_npad;
    int nin[UMNN_MAX_LAYERS], kpad[UMNN_MAX_LAYERS], nout[UMNN_MAX_LAYERS], npad[UMNN_MAX_LAYERS];
    int w_off[UMNN_MAX_LAYERS], b_off[UMNN_MAX_LAYERS];
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

__device__ __forceinline__ const float* slot_ctx(const FwdParams& p, long long slot, int* stride) {
    if (p.layout == UMNN_LAYOUT_STRIDED_D) {
        const long long n = slot / p.D;
        const int d = (int)(slot - n * p.D);
        *stride = p.D;
        return p.h + n * (long long)p.E * p.D + d;
    }
    *stride = 1;
    return p.h + slot * (long long)p.E;
}

template <int HIDDEN_ACT>
__global__ void __launch_bounds__(512) cc_forward_fp32_kernel(const FwdParams p) {
    if (p.run_if != nullptr && *p.run_if != p.run_epoch) return;
    extern __shared__ __align__(16) float smem[];
    float* act = smem;                                        // [max_kpad][kTile]
    float* wst = act + (size_t)p.max_kpad * kTile;            // [2][kKC][max_npad]
    float* fval = wst + 2 * kKC * p.max_npad;                 // [kTile]
    float* tab_t = fval + kTile;                              // [Q+1]
    float* tab_w = tab_t + (p.Q + 1);                         // [Q+1]

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int tx = tid & (kRowGroups - 1);   // row group: rows tx*4..tx*4+3 and 64+tx*4..+3
    const int ty = tid >> 4;                 // unit group: units ty*8..ty*8+7

    const long long slot_begin = (long long)blockIdx.x * p.slots_per_cta;
    long long slot_end = slot_begin + p.slots_per_cta;
    if (slot_end > p.n_slots) slot_end = p.n_slots;
    if (slot_begin >= slot_end) return;
    const long long n_rows = (slot_end - slot_begin) * p.rps;

    for (int i = tid; i <= p.Q; i += nthr) {
        tab_t[i] = p.nodes[i];
        tab_w[i] = p.weights[i];
    }
    for (int i = tid; i < p.max_kpad * kTile; i += nthr) act[i] = 0.0f;
    float carry = 0.0f;  // partial weighted sum of the slot that straddles a tile boundary (warp 0)
    __syncthreads();

    for (long long row0 = 0; row0 < n_rows; row0 += kTile) {
        // ---- 1. input columns -----------------------------------------------------------------
        for (int r = tid; r < kTile; r += nthr) {
            const long long row = row0 + r;
            if (row < n_rows) {
                const long long ls = row / p.rps;
                const int node = (int)(row - ls * p.rps);
                const long long slot = slot_begin + ls;
                const float lo = p.x0 ? p.x0[slot] : 0.0f;
                const float hi = p.x[slot];
                float xi;
                if (node <= p.Q) {
                    const float xT = upper_limit(lo, hi, p.Q);
                    xi = node_abscissa(lo, __fsub_rn(xT, lo), tab_t[node]);
                } else if (node == p.Q + 1 && p.out_fx) {
                    xi = hi;
                } else {
                    xi = lo;
                }
                act[r] = xi;
                int hs;
                const float* hp = slot_ctx(p, slot, &hs);
                for (int e = 0; e < p.E; ++e) act[(1 + e) * kTile + r] = __ldg(hp + (long long)e * hs);
            } else {
                for (int k = 0; k <= p.E; ++k) act[k * kTile + r] = 0.0f;
            }
        }
        __syncthreads();

        // ---- 2. hidden layers -----------------------------------------------------------------
        for (int l = 0; l < p.n_layers - 1; ++l) {
            const int kpad = p.kpad[l], npad = p.npad[l];
            const float* Wg = p.packed + p.w_off[l];
            const int n_chunks = kpad / kKC;
            const int chunk_f4 = kKC * npad / 4;
            const bool active = ty * kUT < npad;

            float acc[8][kUT];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < kUT; ++j) acc[i][j] = 0.0f;

            for (int i = tid; i < chunk_f4; i += nthr) cp_async16(wst + 4 * i, Wg + 4 * i);
            cp_async_commit();

            for (int c = 0; c < n_chunks; ++c) {
                cp_async_wait_all();
                __syncthreads();
                if (c + 1 < n_chunks) {
                    float* dst = wst + ((c + 1) & 1) * kKC * p.max_npad;
                    const float* src = Wg + (size_t)(c + 1) * kKC * npad;
                    for (int i = tid; i < chunk_f4; i += nthr) cp_async16(dst + 4 * i, src + 4 * i);
                    cp_async_commit();
                }
                if (active) {
                    const float* wc = wst + (c & 1) * kKC * p.max_npad + ty * kUT;
                    const float* ac = act + (size_t)c * kKC * kTile + tx * 4;
#pragma unroll
                    for (int kk = 0; kk < kKC; ++kk) {
                        const float4 a0 = *reinterpret_cast<const float4*>(ac + kk * kTile);
                        const float4 a1 = *reinterpret_cast<const float4*>(ac + kk * kTile + 64);
                        const float4 w0 = *reinterpret_cast<const float4*>(wc + kk * npad);
                        const float4 w1 = *reinterpret_cast<const float4*>(wc + kk * npad + 4);
                        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i)
#pragma unroll
                            for (int j = 0; j < kUT; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
                    }
                }
            }
            __syncthreads();  // every thread is done reading act for this layer
            if (active) {
                const float* bg = p.packed + p.b_off[l] + ty * kUT;
#pragma unroll
                for (int j = 0; j < kUT; ++j) {
                    const float b = __ldg(bg + j);
                    float4 o0, o1;
                    o0.x = hidden_act(acc[0][j] + b, HIDDEN_ACT);
                    o0.y = hidden_act(acc[1][j] + b, HIDDEN_ACT);
                    o0.z = hidden_act(acc[2][j] + b, HIDDEN_ACT);
                    o0.w = hidden_act(acc[3][j] + b, HIDDEN_ACT);
                    o1.x = hidden_act(acc[4][j] + b, HIDDEN_ACT);
                    o1.y = hidden_act(acc[5][j] + b, HIDDEN_ACT);
                    o1.z = hidden_act(acc[6][j] + b, HIDDEN_ACT);
                    o1.w = hidden_act(acc[7][j] + b, HIDDEN_ACT);
                    float* dst = act + (size_t)(ty * kUT + j) * kTile + tx * 4;
                    *reinterpret_cast<float4*>(dst) = o0;
                    *reinterpret_cast<float4*>(dst + 64) = o1;
                }
            }
            __syncthreads();
        }

        // ---- 3. output layer + output activation + CC weight -----------------------------------
        {
            const int l = p.n_layers - 1;
            const int nin = p.nin[l];
            const float* wl = p.packed + p.w_off[l];
            const float bl = __ldg(p.packed + p.b_off[l]);
            for (int r = tid; r < kTile; r += nthr) {
                const long long row = row0 + r;
                if (row >= n_rows) continue;
                float v = 0.0f;
                for (int k = 0; k < nin; ++k) v = fmaf(__ldg(wl + k), act[k * kTile + r], v);
                const float f = out_act(v + bl, p.out_act);
                const long long ls = row / p.rps;
                const int node = (int)(row - ls * p.rps);
                if (node <= p.Q) {
                    fval[r] = f * tab_w[node];
                } else {
                    const long long slot = slot_begin + ls;
                    if (node == p.Q + 1 && p.out_fx) p.out_fx[slot] = f;
                    else p.out_fx0[slot] = f;
                }
            }
        }
        __syncthreads();

        // ---- 4. segmented sum over the node rows of every slot touching this tile (warp 0) -----
        if (tid < 32) {
            const long long last_row = (row0 + kTile < n_rows ? row0 + kTile : n_rows) - 1;
            const long long s_first = row0 / p.rps, s_last = last_row / p.rps;
            for (long long ls = s_first; ls <= s_last; ++ls) {
                const long long a = ls * p.rps;            // first node row of the slot
                const long long b = a + p.Q;               // last node row of the slot
                const long long lo = a > row0 ? a : row0;
                const long long hi = b < last_row ? b : last_row;
                float part = 0.0f;
                for (long long r = lo + tid; r <= hi; r += 32) part += fval[(int)(r - row0)];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                if (lo <= hi) {
                    const float total = (a < row0 ? carry : 0.0f) + part;
                    if (b <= last_row) {
                        if (tid == 0) {
                            const long long slot = slot_begin + ls;
                            const float x0v = p.x0 ? p.x0[slot] : 0.0f;
                            const float span = __fsub_rn(upper_limit(x0v, p.x[slot], p.Q), x0v);
                            p.out[slot] = __fmul_rn(__fmul_rn(total, span), 0.5f);
                        }
                        carry = 0.0f;
                    } else {
                        carry = total;
                    }
                }
            }
        }
        // fval / act of this tile are rewritten only after the next tile's first __syncthreads,
        // and warp 0 reaches that barrier after finishing the sums above.
    }
}

__global__ void pack_fp32_kernel(const float* __restrict__ flat, float* __restrict__ packed, Fp32Layout L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.total_floats) return;
    float v = 0.0f;
    for (int l = 0; l < L.n_layers; ++l) {
        const bool last = (l == L.n_layers - 1);
        const int wsz = last ? L.kpad[l] : L.kpad[l] * L.npad[l];
        const int bsz = last ? 4 : L.npad[l];
        if (i >= L.w_off[l] && i < L.w_off[l] + wsz) {
            const int r = i - L.w_off[l];
            const int k = last ? r : r / L.npad[l];
            const int j = last ? 0 : r % L.npad[l];
            if (k < L.nin[l] && j < L.nout[l]) v = flat[L.src_w_off[l] + j * L.nin[l] + k];
            break;
        }
        if (i >= L.b_off[l] && i < L.b_off[l] + bsz) {
            const int j = i - L.b_off[l];
            if (j < L.nout[l]) v = flat[L.src_b_off[l] + j];
            break;
        }
        if (!last && i >= L.d_off[l] && i < L.d_off[l] + L.n16[l] * L.k8[l]) {
            const int r = i - L.d_off[l];
            const int n = r / L.k8[l], k = r % L.k8[l];
            if (n < L.nout[l] && k < L.nin[l]) v = flat[L.src_w_off[l] + n * L.nin[l] + k];
            break;
        }
    }
    packed[i] = v;
}

}  // namespace

int launch_pack_fp32(const umnn_desc* d, const float* flat, float* packed, cudaStream_t s) {
    const Fp32Layout L = make_fp32_layout(d);
    const int threads = 256;
    const int blocks = (L.total_floats + threads - 1) / threads;
    pack_fp32_kernel<<<blocks, threads, 0, s>>>(flat, packed, L);
    UMNN_CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_forward_fp32(const umnn_desc* d, const float* x0, const float* x, const float* h, const float* packed,
                        const float* nodes, const float* weights, float* out, float* out_fx, float* out_fx0,
                        const int* run_if, int run_epoch, cudaStream_t s) {
    const Fp32Layout L = make_fp32_layout(d);
    FwdParams p{};
    p.run_if = run_if;
    p.run_epoch = run_epoch;
    p.x0 = x0; p.x = x; p.h = h; p.packed = packed; p.nodes = nodes; p.weights = weights;
    p.out = out; p.out_fx = out_fx; p.out_fx0 = out_fx0;
    p.n_slots = d->n_samples * (long long)d->n_dims;
    p.D = d->n_dims; p.E = d->n_ctx; p.layout = d->layout; p.Q = d->nb_steps;
    p.rps = d->nb_steps + 1 + (out_fx ? 1 : 0) + (out_fx0 ? 1 : 0);
    p.n_layers = d->n_layers; p.hidden_act = d->hidden_act; p.out_act = d->out_act;
    p.max_kpad = L.max_kpad; p.max_npad = L.max_npad;
    for (int l = 0; l < d->n_layers; ++l) {
        p.nin[l] = L.nin[l]; p.kpad[l] = L.kpad[l]; p.nout[l] = L.nout[l]; p.npad[l] = L.npad[l];
        p.w_off[l] = L.w_off[l]; p.b_off[l] = L.b_off[l];
    }
    if (p.n_slots == 0) return 0;

    int dev = 0, n_sm = 0;
    UMNN_CUDA_TRY(current_device(&dev, &n_sm));

    const int n_groups = L.max_npad / kUT;
    const int threads = round_up(kRowGroups * n_groups, 32);  // 16 x ceil(maxH/8) rounded to whole warps, <= 512
    const size_t smem = sizeof(float) * ((size_t)L.max_kpad * kTile + 2 * kKC * L.max_npad + kTile + 2 * (d->nb_steps + 1));

    auto kern = d->hidden_act == UMNN_ACT_LEAKY_RELU ? cc_forward_fp32_kernel<UMNN_ACT_LEAKY_RELU>
                                                     : cc_forward_fp32_kernel<UMNN_ACT_RELU>;
    UMNN_CUDA_TRY(ensure_dynamic_smem((const void*)kern, dev, (int)smem));
    int occ = 0;
    UMNN_CUDA_TRY(cached_occupancy(&occ, (const void*)kern, dev, threads, smem));
    if (occ < 1) {
        set_error("cc_forward_fp32: kernel does not fit on an SM (threads=%d smem=%zu)", threads, smem);
        return UMNN_ERR_UNSUPPORTED;
    }
    // persistent-style grid: at most occ CTAs per SM, each owning whole slots; never more CTAs than
    // there are tiles of work.
    const long long total_rows = p.n_slots * p.rps;
    long long want = (total_rows + kTile - 1) / kTile;
    const long long cap = (long long)n_sm * occ;
    if (want > cap) want = cap;
    if (want > p.n_slots) want = p.n_slots;
    if (want < 1) want = 1;
    p.slots_per_cta = (p.n_slots + want - 1) / want;
    const long long grid = (p.n_slots + p.slots_per_cta - 1) / p.slots_per_cta;
    kern<<<(unsigned)grid, threads, smem, s>>>(p);
    UMNN_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace umnn
