// Tensor-core backward of the Clenshaw-Curtis integral (tcgen05), three passes per chunk of rows.
//
// Replaces ParallelNeuralIntegral.backward (models/UMNN/ParallelNeuralIntegral.py:110-123),
// integrate(compute_grad=True) (:66-80) and computeIntegrand (:83-94).
//
//   pass F  cc_forward_tc_kernel<.., EMIT=true> (cc_forward_tc.cu): re-evaluates the network for the chunk and
//           writes the activation panels A_0..A_J (bf16 hi [and lo], UMMA-tiled), their sign masks (pair-major) and v per
//           row; with fp16 operands by default, so that its signs are those of the forward that was differentiated.
//   pass D  cc_dgrad_tc_kernel (here): the same persistent CTA-pair machinery run on the TRANSPOSED weights:
//           dz_J = dv * w_out (.) act'(a_J) is rank-1 on CUDA cores, every further layer is
//           da = dz . W (3 MMAs per K block, A operand in TMEM) followed by an in-place epilogue
//           dz = da (.) act'(a) -> hi/lo; the last MMA layer yields d f / d input, reduced per slot into d_h
//           (and the Jacobian-point term of d_x) by the last column group behind "d0 full / empty" named barriers.
//           Emits the dz panels; the prep warps prefetch the tile's sign masks into L2.
//   pass W  cc_wgrad_tc_kernel (here): dW_j = DZ_j^T A_{j-1} for every layer at once; the panels are ready-made
//           MN-major UMMA tiles moved by bulk-TMA; every CTA pair reduces its slab of rows into TMEM
//           accumulators (all layers resident: 464 of 512 columns at [200]^3) and writes one partial, summed
//           in a fixed order by reduce_partials_tc_kernel.  Column H_{j-1} of A_{j-1} holds 1.0, so the bias
//           gradient is one more column of the same GEMM.
//
// Scratch = panels of one chunk (2.7 KB per row at [200]^3 with hi-only panels, 5.3 KB with hi + lo), bounded by
// kBwdMaxTiles tiles per CTA.
#include "tc_common.cuh"
#include "tc_kernels.cuh"

#include <stdlib.h>
#include <string.h>

namespace umnn {

namespace {

using namespace tc;

// ------------------------------------------------------------------------------------------------------
// pass D
// ------------------------------------------------------------------------------------------------------
constexpr int kEpiWarps = 16;
constexpr int kColGroups = kEpiWarps / 4;
constexpr int kPrepWarps = 3;
constexpr int kMmaWarp = kEpiWarps + kPrepWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kPrepThreads = kPrepWarps * 32;
constexpr uint32_t kColP = 0, kColQ = kTcRegionCols;

constexpr int BAR_READY = 0;                                  // [layer][8]
constexpr int BAR_ACC = BAR_READY + kBwdMaxHidden * 8;        // [layer][2]
constexpr int BAR_PREP_FULL = BAR_ACC + kBwdMaxHidden * 2;
constexpr int BAR_PREP_EMPTY = BAR_PREP_FULL + kTcPrepBufs;
constexpr int BAR_WLOAD = BAR_PREP_EMPTY + kTcPrepBufs;
constexpr int BAR_PEER = BAR_WLOAD + 1;
constexpr int BAR_COUNT = BAR_PEER + 1;
static_assert(BAR_COUNT <= kTcDgradBars, "barrier table too small");

struct TcDgradParams {
    const float *x0, *x, *weights, *grad_out, *grad_fx;
    const uint8_t* blobs;                      // dgrad blobs [2][blob_bytes]
    const float* v;                            // [R_pad]
    const uint32_t* mask[UMNN_MAX_LAYERS];     // j = 1..J
    uint8_t* dz[UMNN_MAX_LAYERS + 1];          // DZ_j panels, j = 1..J+1
    float *d_x0, *d_x, *d_h;
    long long slot0, n_slots, slots_per_cta, row_block, r_pad;
    int tiles_per_cta, D, E, layout, Q, rps, out_act;
    int dz_parts;                              // parts of the dz panels: 1 = hi only, 2 = hi + lo
    int dz_head_parts;                         // parts of DZ_J and DZ_{J+1} (see BwdPanels)
    const int* run_if;                         // not NULL: no-op unless *run_if != 0 (guarded bf16 re-run)
    TcDgradLayout L;
    TcDgradSmem S;
};

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"r"(kEpiThreads) : "memory"); }
// producer / consumer barriers between the warps that park the input gradient in shared memory and the column group
// that reduces it (PTX ISA, bar.arrive + bar.sync): 3 = "d0 full", 4 = "d0 empty"
__device__ __forceinline__ void named_arrive(int id, int n_threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n_threads) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int n_threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory"); }
constexpr int BARID_D0_FULL = 3, BARID_D0_EMPTY = 4;

// the chain needs hi AND lo of dz for its own MMAs, so both exist in registers; the panel keeps `parts` of them
__device__ __forceinline__ void emit16(const PanelRow& R, int col, const uint32_t (&o)[16], int parts) {
    uint8_t* g0 = R.base + panel_granule(R, col);
    uint8_t* g1 = R.base + panel_granule(R, col + 8);
    *reinterpret_cast<uint4*>(g0) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4*>(g1) = make_uint4(o[4], o[5], o[6], o[7]);
    if (parts == 2) {
        *reinterpret_cast<uint4*>(g0 + R.lo_off) = make_uint4(o[8], o[9], o[10], o[11]);
        *reinterpret_cast<uint4*>(g1 + R.lo_off) = make_uint4(o[12], o[13], o[14], o[15]);
    }
}

// v * act'(.) for column c (0..31) of a pair from the recorded sign mask (set bit = negative pre-activation)
template <int HIDDEN_ACT>
__device__ __forceinline__ float times_slope(float v, uint32_t bits, int c) {
    if (bits & (1u << mask_bitpos(c))) v = (HIDDEN_ACT == UMNN_ACT_LEAKY_RELU) ? v * kLeakySlope : 0.0f;
    return v;
}

template <int HIDDEN_ACT>
__global__ void __launch_bounds__(kThreads, 1) cc_dgrad_tc_kernel(const __grid_constant__ TcDgradParams p) {
    if (p.run_if != nullptr && *p.run_if == 0) return;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte aligned base, by pointer arithmetic on the __shared__ array so that the compiler keeps the
    // address space (LDS/STS instead of generic LD/ST on every table and scratch access)
#if defined(UMNN_TC_SMEM_GENERIC) && UMNN_TC_SMEM_GENERIC
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
#else
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
#endif
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform role index (see cc_forward_tc.cu)
    const uint32_t rank = cluster_ctarank();
    const TcDgradLayout& L = p.L;
    const int T = p.tiles_per_cta, J = L.J;

    float* dvrow = reinterpret_cast<float*>(smem + p.S.off_dv);      // [bufs][128]
    float* frow = reinterpret_cast<float*>(smem + p.S.off_f);        // [bufs][128]
    int* nodeid = reinterpret_cast<int*>(smem + p.S.off_node);       // [bufs][128]
    float* d0 = reinterpret_cast<float*>(smem + p.S.off_d0);         // [128][n0pad + 1]
    float* carry = reinterpret_cast<float*>(smem + p.S.off_carry);   // [2][E]
    float* tab_w = reinterpret_cast<float*>(smem + p.S.off_tabw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.S.off_bars);
    uint32_t* holder = reinterpret_cast<uint32_t*>(smem + p.S.off_holder);
    const float* wlast = reinterpret_cast<const float*>(smem + L.off_wlast);
    const int d0_stride = L.n0pad + 1;

    const long long slot_begin = p.slot0 + (long long)blockIdx.x * p.slots_per_cta;
    long long slot_end = slot_begin + p.slots_per_cta;
    if (slot_end > p.slot0 + p.n_slots) slot_end = p.slot0 + p.n_slots;
    const long long n_rows = slot_end > slot_begin ? (slot_end - slot_begin) * p.rps : 0;
    const long long cta_row0 = (long long)blockIdx.x * p.row_block;

    if (tid == 0) {
        for (int m = 0; m < kBwdMaxHidden; ++m) {
            for (int j = 0; j < 8; ++j) mbar_init(&bars[BAR_READY + m * 8 + j], 8);
            for (int s = 0; s < 2; ++s) mbar_init(&bars[BAR_ACC + m * 2 + s], 1);
        }
        for (int b = 0; b < kTcPrepBufs; ++b) {
            mbar_init(&bars[BAR_PREP_FULL + b], kPrepWarps);
            mbar_init(&bars[BAR_PREP_EMPTY + b], kEpiWarps);
        }
        mbar_init(&bars[BAR_WLOAD], 1);
        mbar_init(&bars[BAR_PEER], 2);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc<2>(holder, 512);
    for (int i = tid; i <= p.Q; i += kThreads) tab_w[i] = p.weights[i];
    for (int i = tid; i < 2 * p.E; i += kThreads) carry[i] = 0.0f;
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tbase = *holder;

    if (warp == kMmaWarp) {
        // =========================================================== MMA issuer (same protocol as the forward)
        if (elect_one_sync()) {
            mbar_expect_tx(&bars[BAR_WLOAD], L.blob_bytes);
            const uint8_t* src = p.blobs + (size_t)rank * L.blob_bytes;
            for (uint32_t off = 0; off < L.blob_bytes; off += 32768u) {
                const uint32_t n = (L.blob_bytes - off < 32768u) ? (L.blob_bytes - off) : 32768u;
                bulk_g2s(smem + off, src + off, n, &bars[BAR_WLOAD]);
            }
        }
        __syncwarp();
        mbar_wait(&bars[BAR_WLOAD], 0, 100);
        if (elect_one_sync()) mbar_arrive_cluster(&bars[BAR_PEER], 0);
        __syncwarp();
        // lane-0 broadcasts keep the issuer's descriptors / tensor-memory addresses in uniform registers (cc_forward_tc.cu)
        const uint32_t rank_u = __shfl_sync(0xffffffffu, rank, 0);
        const uint32_t tbase = __shfl_sync(0xffffffffu, *holder, 0);
        if (rank_u == 0) {
            mbar_wait_warp(&bars[BAR_PEER], 0, 101);
            const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
            for (int t = 0; t < T; ++t) {
                const uint32_t par = (uint32_t)(t & 1);
                for (int m = 0; m < J; ++m) {
                    const TcChainLayer& y = L.layer[m];
                    const uint32_t a_base = tbase + ((m & 1) ? kColP : kColQ);
                    const uint32_t col_d = (m & 1) ? kColQ : kColP;
                    const int n_kb = y.kpad / 16;
                    MmaSegment gs[2];
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const uint32_t lbo = (uint32_t)(y.seg_n[s] / 16) * 128u;
                        gs[s].idesc = make_idesc_bf16_f32(256, y.seg_n[s]);
                        gs[s].step = (uint64_t)((2u * lbo) >> 4);
                        gs[s].d_addr = tbase + col_d + (uint32_t)y.seg_begin[s];
                        gs[s].bhi = make_smem_desc(sbase + y.b_off[0][s], lbo, 128);
                        gs[s].blo = make_smem_desc(sbase + y.b_off[1][s], lbo, 128);
                    }
                    issue_mma_layer(n_kb, y.nseg, gs[0], gs[1], a_base, &bars[BAR_READY + m * 8], par, &bars[BAR_ACC + m * 2], 200 + m * 8);
                }
            }
        }
        __syncwarp();
    } else if (warp >= kEpiWarps) {
        // =========================================================== prep warps: cotangent of v, f, dz_{J+1} panel
        const int ptid = tid - kEpiThreads;
        for (int u = 0; u < T; ++u) {
            const int b = u % kTcPrepBufs;
            if (u >= kTcPrepBufs) mbar_wait(&bars[BAR_PREP_EMPTY + b], (uint32_t)((u / kTcPrepBufs - 1) & 1), 120 + b);
            const uint32_t row0 = (uint32_t)u * kTcTile;          // 32-bit row arithmetic: a chunk holds <= 4096 rows per CTA
            // Pull this tile's sign masks (written by pass F, by now evicted to HBM) into L2 ahead of the epilogue warps:
            // their mask loads sat on the critical path with DRAM latency (17 % of this kernel's stall samples).  A
            // (layer, pair) run of 128 rows is 4 lines of 128 bytes.
            for (int j = 1; j <= J; ++j) {
                const int n_lines = ((L.P[j] + 31) / 32) * 4;
                for (int idx = ptid; idx < n_lines; idx += kPrepThreads)
                    prefetch_l2(p.mask[j] + (long long)(idx >> 2) * p.r_pad + cta_row0 + row0 + (idx & 3) * 32);
            }
            for (int r = ptid; r < kTcTile; r += kPrepThreads) {
                const uint32_t row = row0 + r;
                const long long pr = cta_row0 + row;
                float dv = 0.0f, f = 0.0f;
                int node = -1;
                if (row < (uint32_t)n_rows) {
                    const uint32_t ls = row / (uint32_t)p.rps;
                    node = (int)(row - ls * (uint32_t)p.rps);
                    const long long slot = slot_begin + ls;
                    const float v = p.v[pr];
                    f = out_act(v, p.out_act);
                    float c = 0.0f;
                    if (node <= p.Q) {
                        const float lo = p.x0 ? __ldg(p.x0 + slot) : 0.0f;
                        const float span = __fsub_rn(upper_limit(lo, __ldg(p.x + slot), p.Q), lo);
                        c = __fmul_rn(__fmul_rn(__fmul_rn(__ldg(p.grad_out + slot), span), 0.5f), tab_w[node]);
                    } else if (node == p.Q + 1 && p.grad_fx) {
                        c = __ldg(p.grad_fx + slot);
                    }
                    const float dact = (p.out_act == UMNN_OUT_ELU_PLUS_1) ? (v > 0.0f ? 1.0f : expf(v)) : f * (1.0f - f);
                    dv = c * dact;
                }
                dvrow[b * kTcTile + r] = dv;
                frow[b * kTcTile + r] = f;
                nodeid[b * kTcTile + r] = node;
                // dz_{J+1} panel (width 16): column 0 = dv
                uint32_t hi, lo2;
                split_bf16x2(dv, 0.0f, hi, lo2);
                *reinterpret_cast<uint4*>(p.dz[J + 1] + panel_offset(pr, 0, 16, 0, p.dz_head_parts, p.r_pad)) = make_uint4(hi, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(p.dz[J + 1] + panel_offset(pr, 8, 16, 0, p.dz_head_parts, p.r_pad)) = make_uint4(0u, 0u, 0u, 0u);
                if (p.dz_head_parts == 2) {
                    *reinterpret_cast<uint4*>(p.dz[J + 1] + panel_offset(pr, 0, 16, 1, 2, p.r_pad)) = make_uint4(lo2, 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(p.dz[J + 1] + panel_offset(pr, 8, 16, 1, 2, p.r_pad)) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAR_PREP_FULL + b]);
        }
    } else {
        // =========================================================== epilogue warps
        const int q = warp & 3, cg = warp >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const bool even_layers = (J & 1) == 0;
        const int PJ = L.P[J];
        const int pairsJ = (PJ + 31) / 32, pairs0 = L.n0pad / 32;
        const int kRedBarThreads = 128 * (pairs0 + 1);      // the column groups that write d0 (pairs0 <= 2) + the one that reduces it
        const uint32_t col_last = ((J - 1) & 1) ? kColQ : kColP;
        const TcChainLayer& ylast = L.layer[J - 1];

        mbar_wait(&bars[BAR_WLOAD], 0, 130);
        // (sample, dimension) of the CTA's first slot: the one 64-bit division; slots further on use 32-bit arithmetic
        const long long n_begin = p.layout == UMNN_LAYOUT_STRIDED_D ? slot_begin / p.D : 0;
        const uint32_t d_begin = (uint32_t)(slot_begin - n_begin * p.D);

        // rank-1 head of the chain for tile `tile`: dz_J = dv * w_out (.) act'(a_J), 32 columns -> region Q
        auto head_pair = [&](int bu, int tile, int pp) {
            const long long pr = cta_row0 + (long long)tile * kTcTile + r;
            const float dv = dvrow[bu * kTcTile + r];
            const uint32_t bits = p.mask[J][(long long)pp * p.r_pad + pr];
            const int halves = (32 * pp + 16 < PJ) ? 2 : 1;
            const PanelRow prow = panel_row(p.dz[J], pr, PJ, p.dz_head_parts, p.r_pad);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (hf < halves) {
                    const float* wv = wlast + 32 * pp + 16 * hf;
                    uint32_t o[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float z0 = times_slope<HIDDEN_ACT>(dv * wv[2 * i], bits, 16 * hf + 2 * i);
                        const float z1 = times_slope<HIDDEN_ACT>(dv * wv[2 * i + 1], bits, 16 * hf + 2 * i + 1);
                        split_bf16x2(z0, z1, o[i], o[8 + i]);
                    }
                    tmem_st16(tbase + lane_sel + kColQ + 32u * pp + 16u * hf, o);
                    emit16(prow, 32 * pp + 16 * hf, o, p.dz_head_parts);
                }
            }
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&bars[BAR_READY + 0 * 8 + pp], 0);
        };

        mbar_wait(&bars[BAR_PREP_FULL + 0], 0, 131);
        for (int pp = cg; pp < pairsJ; pp += kColGroups) head_pair(0, 0, pp);

        for (int t = 0; t < T; ++t) {
            const uint32_t par = (uint32_t)(t & 1);
            const int b = t % kTcPrepBufs;
            const bool has_next = (t + 1 < T);
            const int bn = (t + 1) % kTcPrepBufs;
            const long long pr = cta_row0 + (long long)t * kTcTile + r;

            // ---- da of layer m -> dz (mask) -> A operand of layer m+1, in place; emit the dz panel
            for (int m = 0; m + 1 < J; ++m) {
                const TcChainLayer& y = L.layer[m];
                const int jout = J - m - 1;                          // hidden layer whose dz this is
                const uint32_t col_d = (m & 1) ? kColQ : kColP;
                const int n_pairs = (y.npad + 31) / 32;
                for (int pp = cg; pp < n_pairs; pp += kColGroups) {
                    const int s = (y.nseg == 2 && 32 * pp >= y.seg_begin[1]) ? 1 : 0;
                    const uint32_t bits = p.mask[jout][(long long)pp * p.r_pad + pr];
                    mbar_wait(&bars[BAR_ACC + m * 2 + s], par, 300 + m * 2 + s);
                    tc_fence_after_sync();
                    const uint32_t taddr = tbase + lane_sel + col_d + 32u * pp;
                    const bool two = 32 * pp + 16 < y.npad;
                    uint32_t v0[16], v1[16], o[16];
                    tmem_ld16(taddr, v0);
                    if (two) tmem_ld16(taddr + 16, v1);
                    tmem_ld_wait();
                    const PanelRow prow = panel_row(p.dz[jout], pr, y.npad, p.dz_parts, p.r_pad);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        split_bf16x2(times_slope<HIDDEN_ACT>(__uint_as_float(v0[2 * i]), bits, 2 * i),
                                     times_slope<HIDDEN_ACT>(__uint_as_float(v0[2 * i + 1]), bits, 2 * i + 1), o[i], o[8 + i]);
                    tmem_st16(taddr, o);
                    emit16(prow, 32 * pp, o, p.dz_parts);
                    if (two) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            split_bf16x2(times_slope<HIDDEN_ACT>(__uint_as_float(v1[2 * i]), bits, 16 + 2 * i),
                                         times_slope<HIDDEN_ACT>(__uint_as_float(v1[2 * i + 1]), bits, 17 + 2 * i), o[i], o[8 + i]);
                        tmem_st16(taddr + 16, o);
                        emit16(prow, 32 * pp + 16, o, p.dz_parts);
                    }
                    tmem_st_wait();
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(&bars[BAR_READY + (m + 1) * 8 + pp], 0);
                }
            }

            // ---- last accumulator = d f / d input (n0pad columns) -> shared memory; even chains also write the
            //      head of the next tile over the pair just consumed
            if (has_next && even_layers) mbar_wait(&bars[BAR_PREP_FULL + bn], (uint32_t)(((t + 1) / kTcPrepBufs) & 1), 132);
            const int n_loop = pairs0 > pairsJ ? pairs0 : pairsJ;
            for (int pp = cg; pp < n_loop; pp += kColGroups) {
                if (pp < pairs0) {
                    const int s = (ylast.nseg == 2 && 32 * pp >= ylast.seg_begin[1]) ? 1 : 0;
                    mbar_wait(&bars[BAR_ACC + (J - 1) * 2 + s], par, 310 + s);
                    tc_fence_after_sync();
                    uint32_t v0[16], v1[16];
                    const uint32_t taddr = tbase + lane_sel + col_last + 32u * pp;
                    tmem_ld16(taddr, v0);
                    tmem_ld16(taddr + 16, v1);
                    tmem_ld_wait();
                    if (t > 0) named_sync(BARID_D0_EMPTY, kRedBarThreads);     // the previous tile's reduction has read d0
                    float* dst = d0 + r * d0_stride + 32 * pp;
#pragma unroll
                    for (int i = 0; i < 16; ++i) { dst[i] = __uint_as_float(v0[i]); dst[16 + i] = __uint_as_float(v1[i]); }
                    __threadfence_block();
                    named_arrive(BARID_D0_FULL, kRedBarThreads);
                } else if (even_layers) {
                    for (int s = 0; s < ylast.nseg; ++s) mbar_wait(&bars[BAR_ACC + (J - 1) * 2 + s], par, 320 + s);
                    tc_fence_after_sync();
                }
                if (even_layers && has_next && pp < pairsJ) head_pair(bn, t + 1, pp);
            }
            if (!even_layers) {
                // odd chains: the last accumulator lives in P, which MMA layer 0 of the next tile overwrites -> the head of
                // the next tile is published only after the warps that read P have done so
                epi_bar_sync();
                if (has_next) {
                    for (int s = 0; s < ylast.nseg; ++s) mbar_wait(&bars[BAR_ACC + (J - 1) * 2 + s], par, 330 + s);
                    tc_fence_after_sync();
                    mbar_wait(&bars[BAR_PREP_FULL + bn], (uint32_t)(((t + 1) / kTcPrepBufs) & 1), 133);
                    for (int pp = cg; pp < pairsJ; pp += kColGroups) head_pair(bn, t + 1, pp);
                }
            }

            // ---- per-slot context gradient (carried across tiles), Leibniz terms, Jacobian-point term of d_x: the LAST
            //      column group alone (it converts half as many pairs per layer as the others), behind the d0 full /
            //      d0 empty barriers.  With every epilogue warp in this reduction and a CTA-wide barrier on either side
            //      of it, the next tile's first accumulators waited for it (14 % of the kernel's stall samples).
            if (cg == kColGroups - 1) {
                named_sync(BARID_D0_FULL, kRedBarThreads);
                const int rtid = tid - (kColGroups - 1) * 128;          // 0..127
                const int row0 = t * kTcTile, n_rows_i = (int)n_rows;
                if (row0 < n_rows_i) {
                    const int last_row = (row0 + kTcTile < n_rows_i ? row0 + kTcTile : n_rows_i) - 1;
                    const int s_first = (int)((uint32_t)row0 / (uint32_t)p.rps), s_last = (int)((uint32_t)last_row / (uint32_t)p.rps);
                    const int ns = s_last - s_first + 1;
                    const float* cin = carry + (t & 1) * p.E;
                    float* cout = carry + ((t + 1) & 1) * p.E;
                    // one (slot, e) sum per thread, four independent partial sums (rows mod 4) added in a fixed order;
                    // neighbouring lanes read neighbouring floats of a d0 row
                    for (int idx = rtid; idx < ns * p.E; idx += 128) {
                        const int i = idx / p.E, e = idx - i * p.E;
                        const int ls = s_first + i;
                        const int a = ls * p.rps, bb = a + p.rps - 1;
                        const int lo = a > row0 ? a : row0;
                        const int hi = bb < last_row ? bb : last_row;
                        const float* col = d0 + 1 + e - row0 * d0_stride;
                        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
                        int rr = lo;
                        for (; rr + 3 <= hi; rr += 4) {
                            s0 += col[rr * d0_stride];
                            s1 += col[(rr + 1) * d0_stride];
                            s2 += col[(rr + 2) * d0_stride];
                            s3 += col[(rr + 3) * d0_stride];
                        }
                        for (; rr <= hi; ++rr) s0 += col[rr * d0_stride];
                        float sum = (s0 + s1) + (s2 + s3);
                        if (a < row0) sum += cin[e];
                        if (bb <= last_row) {
                            if (p.d_h) {
                                if (p.layout == UMNN_LAYOUT_STRIDED_D) {
                                    const uint32_t dd = d_begin + (uint32_t)ls, dn = dd / (uint32_t)p.D;   // 32-bit: see n_begin
                                    p.d_h[(n_begin + dn) * (long long)p.E * p.D + (long long)e * p.D + (dd - dn * (uint32_t)p.D)] = sum;
                                } else {
                                    p.d_h[(slot_begin + ls) * (long long)p.E + e] = sum;
                                }
                            }
                        } else {
                            cout[e] = sum;
                        }
                    }
                    const int node = nodeid[b * kTcTile + r];          // r = this thread's row of the tile, as in every column group
                    if (node > p.Q) {
                        const long long slot = slot_begin + (uint32_t)(row0 + r) / (uint32_t)p.rps;
                        const float g = p.grad_out[slot];
                        if (node == p.Q + 1) {
                            if (p.d_x) p.d_x[slot] = frow[b * kTcTile + r] * g + d0[r * d0_stride];
                        } else if (p.d_x0) {
                            p.d_x0[slot] = -frow[b * kTcTile + r] * g;
                        }
                    }
                }
                __threadfence_block();
                if (has_next) named_arrive(BARID_D0_EMPTY, kRedBarThreads);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAR_PREP_EMPTY + b]);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == kMmaWarp) tmem_dealloc<2>(tbase, 512);
}

// dgrad weights: B'_i[n'][k'] = W_{J-i}[k'][n'], bf16 hi / lo, K-major core matrices split over the CTA pair
__device__ __forceinline__ uint16_t bf16_bits(float v) { return (uint16_t)(pack_bf16x2(v, 0.0f) & 0xFFFFu); }
__device__ __forceinline__ float bf16_val(float v) { return __uint_as_float((uint32_t)bf16_bits(v) << 16); }

__global__ void pack_dgrad_weights_kernel(const float* __restrict__ flat, uint8_t* __restrict__ blobs, TcDgradLayout L) {
    const uint32_t per_rank = L.weights_bytes / 2;
    const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= 2 * per_rank) return;
    const uint32_t rank = gi / per_rank;
    const uint32_t byte_off = (gi - rank * per_rank) * 2;
    int m = 0, part = 0, s = 0;
    bool found = false;
    for (int mm = 0; mm < L.J && !found; ++mm)
        for (int pp = 0; pp < 2 && !found; ++pp)
            for (int ss = 0; ss < L.layer[mm].nseg && !found; ++ss) {
                const uint32_t lo = L.layer[mm].b_off[pp][ss];
                const uint32_t sz = (uint32_t)(L.layer[mm].seg_n[ss] / 2) * L.layer[mm].kpad * 2;
                if (byte_off >= lo && byte_off < lo + sz) { m = mm; part = pp; s = ss; found = true; }
            }
    const TcChainLayer& y = L.layer[m];
    const uint32_t idx = (byte_off - y.b_off[part][s]) / 2;
    const int rows_cta = y.seg_n[s] / 2, n8c = rows_cta / 8;
    const int per_kb = 2 * n8c * 64;
    const int kb = idx / per_kb;
    int rem = idx - kb * per_kb;
    const int k8 = rem / (n8c * 64);
    rem -= k8 * n8c * 64;
    const int n8 = rem / 64;
    const int rr = (rem % 64) / 8, kk = rem % 8;
    const int nprime = y.seg_begin[s] + (int)rank * rows_cta + n8 * 8 + rr;   // input unit of Linear layer j
    const int kprime = kb * 16 + k8 * 8 + kk;                                // output unit of Linear layer j
    const int j = L.J - m;                // hidden layer index; Linear layer (0-based in the flat vector) j - 1
    const int h_out = L.H[j], h_in = L.H[j - 1];
    float hi = 0.0f, lo = 0.0f;
    if (nprime < h_in && kprime < h_out) {
        const float w = flat[L.src_w_off[j - 1] + kprime * h_in + nprime];
        hi = bf16_val(w);
        lo = bf16_val(w - hi);
    }
    reinterpret_cast<uint16_t*>(blobs + (size_t)rank * L.blob_bytes)[byte_off / 2] = bf16_bits(part == 0 ? hi : lo);
}

__global__ void pack_dgrad_consts_kernel(const float* __restrict__ flat, uint8_t* __restrict__ blobs, TcDgradLayout L) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= L.P[L.J]) return;
    const float v = (n < L.H[L.J]) ? flat[L.src_w_off[L.J] + n] : 0.0f;     // output layer weights, no bias
    for (int rank = 0; rank < 2; ++rank) *reinterpret_cast<float*>(blobs + (size_t)rank * L.blob_bytes + L.off_wlast + 4 * n) = v;
}

// ------------------------------------------------------------------------------------------------------
// pass W
// ------------------------------------------------------------------------------------------------------
constexpr int kWMaxStages = 8;
constexpr int kWEpiWarps = 4;
constexpr int kWThreads = (kWEpiWarps + 2) * 32;   // warps 0-3 epilogue, 4 producer, 5 MMA

struct TcWgradParams {
    const uint8_t* panel[2 * UMNN_MAX_LAYERS + 2];      // panel base pointers
    float* part;                 // [n_pairs][P]
    long long P;                 // parameters
    long long n_blocks;          // 16-row blocks in the chunk
    TcWgradPlan W;
    int n_stages;                // ring depth (2..kWMaxStages), as many as fit in shared memory
    const int* run_if;           // not NULL: no-op unless *run_if != 0
    int src_w_off[UMNN_MAX_LAYERS], src_b_off[UMNN_MAX_LAYERS];
};

__global__ void __launch_bounds__(kWThreads, 1) cc_wgrad_tc_kernel(const __grid_constant__ TcWgradParams p) {
    if (p.run_if != nullptr && *p.run_if == 0) return;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte aligned base, by pointer arithmetic on the __shared__ array so that the compiler keeps the
    // address space (LDS/STS instead of generic LD/ST on every table and scratch access)
#if defined(UMNN_TC_SMEM_GENERIC) && UMNN_TC_SMEM_GENERIC
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
#else
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
#endif
    __shared__ __align__(8) uint64_t full[kWMaxStages], empty[kWMaxStages], done, peer_ready[kWMaxStages];
    const int kWStages = p.n_stages;
    __shared__ uint32_t holder;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform role index (see cc_forward_tc.cu)
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const TcWgradPlan& W = p.W;

    // this pair's slab of 16-row blocks
    const long long per = (p.n_blocks + n_pairs - 1) / n_pairs;
    const long long blk_begin = (long long)pair * per;
    long long blk_end = blk_begin + per;
    if (blk_end > p.n_blocks) blk_end = p.n_blocks;
    const long long n_kb = blk_end > blk_begin ? blk_end - blk_begin : 0;

    // zero the stage buffers once (M tiles of panels narrower than this CTA's 128-column range stay zero there)
    for (uint32_t i = tid; i < (W.stage_bytes * kWStages) / 16; i += kWThreads)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    if (tid == 0) {
        for (int s = 0; s < kWStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&peer_ready[s], 2);
        }
        mbar_init(&done, 1);
        fence_mbar_init();
    }
    if (warp == kWEpiWarps + 1) tmem_alloc<2>(&holder, 512);
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tbase = holder;

    const int kbs = W.kb_per_stage;
    // pipeline stages this pair walks; 32-bit counters and a rolling (stage, phase) pair below: a 64-bit division
    // in these loops would be a subroutine call, which costs the MMA issuer its uniform registers
    const int n_kb32 = (int)n_kb;
    const int n_st = (n_kb32 + kbs - 1) / kbs;

    if (warp == kWEpiWarps) {
        // =========================================================== producer: ONE bulk-TMA copy per panel and stage
        // (this CTA's half-panel is contiguous over consecutive blocks); the copies of a stage are spread over the lanes
        __shared__ uint32_t cp_dst[32], cp_blk_bytes[32];
        __shared__ unsigned long long cp_src[32];
        __shared__ uint32_t n_copies_s, blk_bytes_s;
        if (lane == 0) {
            uint32_t n = 0, bytes = 0;
            for (int pn = 0; pn < W.n_panels; ++pn) {
                cp_dst[n] = W.tile_off[pn];
                cp_src[n] = (unsigned long long)(p.panel[pn] + (size_t)rank * panel_half_bytes(p.n_blocks * 16, W.panel_width[pn], W.panel_parts[pn]));
                cp_blk_bytes[n] = W.block_bytes[pn];
                bytes += cp_blk_bytes[n];
                ++n;
            }
            n_copies_s = n;
            blk_bytes_s = bytes;
        }
        __syncwarp();
        const uint32_t n_copies = n_copies_s, blk_bytes = blk_bytes_s;
        int st = 0;
        uint32_t round = 0;                                   // how many times the ring has wrapped
        for (int si = 0; si < n_st; ++si) {
            if (round > 0) mbar_wait(&empty[st], (round - 1) & 1u, 400 + st);
            const int kb0 = si * kbs;
            const uint32_t nb = (uint32_t)((n_kb32 - kb0 < kbs) ? (n_kb32 - kb0) : kbs);     // blocks in this stage (tail: fewer)
            if (lane == 0) mbar_expect_tx(&full[st], nb * blk_bytes);
            __syncwarp();
            uint8_t* sb = smem + (size_t)st * W.stage_bytes;
            const unsigned long long blk = (unsigned long long)(blk_begin + kb0);
            for (uint32_t i = lane; i < n_copies; i += 32)
                bulk_g2s(sb + cp_dst[i], reinterpret_cast<const uint8_t*>(cp_src[i] + blk * cp_blk_bytes[i]), nb * cp_blk_bytes[i], &full[st]);
            if (++st == kWStages) { st = 0; ++round; }
        }
        __syncwarp();
    } else if (warp == kWEpiWarps + 1) {
        // =========================================================== MMA issuer
        // Converged warp, warp-uniform control flow (votes in the waits, election inside the MMA statement): the operand
        // descriptors are rebuilt per layer and stage from kernel parameters with uniform-datapath arithmetic.
        const uint32_t rank_u = __shfl_sync(0xffffffffu, rank, 0);
        const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
        const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
        const int n_layers = W.n_layers;
        int st = 0;
        uint32_t ph = 0;
        for (int si = 0; si < n_st; ++si) {
            mbar_wait_warp(&full[st], ph, 410 + st);
            // both CTAs' tiles of this stage must have landed before the pair-wide MMA reads them
            if (elect_one_sync()) mbar_arrive_cluster(&peer_ready[st], 0);
            __syncwarp();
            if (rank_u == 0) {
                mbar_wait_warp(&peer_ready[st], ph, 420 + st);
                tc_fence_after_sync();
                const int kb0 = si * kbs;
                const int nb = (n_kb32 - kb0 < kbs) ? (n_kb32 - kb0) : kbs;
                const uint32_t stage_addr = sbase + (uint32_t)st * W.stage_bytes;
                for (int l = 0; l < n_layers; ++l) {
                    const TcWgradLayer& y = W.layer[l];
                    const uint32_t idesc = make_idesc_bf16_f32(256, y.n_width) | (1u << 15) | (1u << 16);
                    // tiles hold [block][hi | lo] x [k8][W/16 core matrices]; the M tile is read 128 rows deep although only
                    // W/2 are staged (rows beyond feed accumulator rows nobody reads)
                    const uint32_t lbo_m = (uint32_t)(y.m_width / 16) * 128u, lbo_n = (uint32_t)(y.n_width / 16) * 128u;
                    uint64_t ahi = make_smem_desc(stage_addr + W.tile_off[y.m_panel], lbo_m, 128);
                    uint64_t bhi = make_smem_desc(stage_addr + W.tile_off[y.n_panel], lbo_n, 128);
                    const uint64_t a_lo_off = (uint64_t)((16u * (uint32_t)y.m_width) >> 4), b_lo_off = (uint64_t)((16u * (uint32_t)y.n_width) >> 4);
                    const uint64_t ms = (uint64_t)(W.block_bytes[y.m_panel] >> 4), ns = (uint64_t)(W.block_bytes[y.n_panel] >> 4);
                    // hi*hi always; the cross terms only for operands whose panel carries a lo part
                    const bool m_lo = W.panel_parts[y.m_panel] == 2, n_lo = W.panel_parts[y.n_panel] == 2;
                    const uint32_t d_addr = tb + (uint32_t)y.tmem_col;
                    for (int i = 0; i < nb; ++i) {
                        mma_ss_elect<2>(d_addr, ahi, bhi, idesc, (si > 0 || i > 0) ? 1u : 0u);
                        if (m_lo) mma_ss_elect<2>(d_addr, ahi + a_lo_off, bhi, idesc, 1);
                        if (n_lo) mma_ss_elect<2>(d_addr, ahi, bhi + b_lo_off, idesc, 1);
                        ahi += ms;
                        bhi += ns;
                    }
                }
                // release the stage in BOTH CTAs once these MMAs have read it
                if (elect_one_sync()) mma_commit<2>(&empty[st], 0x3);
                __syncwarp();
            }
            if (++st == kWStages) { st = 0; ph ^= 1u; }
        }
        if (rank == 0) {
            if (elect_one_sync()) mma_commit<2>(&done, 0x3);
            __syncwarp();
        }
    } else {
        // =========================================================== epilogue: TMEM accumulators -> this pair's partial
        mbar_wait(&done, 0, 430);
        tc_fence_after_sync();
        float* out = p.part + (size_t)pair * p.P;
        const int lane_row = warp * 32 + lane;                  // accumulator row (TMEM lane) held by this thread
        const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
        for (int l = 0; l < W.n_layers; ++l) {
            const TcWgradLayer& y = W.layer[l];
            // CTA `rank` staged columns [rank * W/2, (rank + 1) * W/2) of the M panel as its rows 0 .. W/2 - 1
            const int row = (lane_row < y.m_width / 2) ? (int)rank * (y.m_width / 2) + lane_row : (1 << 30);
            for (int c = 0; c < y.n_width; c += 16) {
                uint32_t v[16];
                tmem_ld16(tbase + lane_sel + (uint32_t)y.tmem_col + c, v);
                tmem_ld_wait();
                if (n_kb == 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0u;
                }
                if (!y.swapped) {
                    if (row < y.n_out) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int k = c + i;
                            if (k < y.n_in) out[p.src_w_off[y.lin] + (size_t)row * y.n_in + k] = __uint_as_float(v[i]);
                            else if (k == y.ones_col) out[p.src_b_off[y.lin] + row] = __uint_as_float(v[i]);
                        }
                    }
                } else if (c == 0) {
                    if (row < y.n_in) out[p.src_w_off[y.lin] + row] = __uint_as_float(v[0]);
                    else if (row == y.ones_col) out[p.src_b_off[y.lin]] = __uint_as_float(v[0]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == kWEpiWarps + 1) tmem_dealloc<2>(tbase, 512);
}

__global__ void reduce_partials_tc_kernel(const float* __restrict__ part, long long P, int n_part, float* __restrict__ d_params,
                                          int accumulate, const int* __restrict__ run_if) {
    if (run_if != nullptr && *run_if == 0) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float s = accumulate ? d_params[i] : 0.0f;
    for (int z = 0; z < n_part; ++z) s += part[(size_t)z * P + i];
    d_params[i] = s;
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
// Tiles of 128 rows per CTA and chunk.  Every chunk pays the prologues and tails of its four launches (weight staging,
// tensor-memory allocation, cluster barriers, the last tile's drain), the gaps between them and -- through the size of
// the workspace the guarded FP32 re-run borrows -- a share of that re-run's no-op launches.  Sweep on one box
// (profiles/r2_bwd_chunk_sweep.txt; 32 / 64 / 96 / 128 tiles, balanced chunks): config 4 at 8192 samples 85.6 / 80.8 /
// 79.7 / 78.2 ms, config 3 4.74 / 4.43 / 4.26 / 4.29 ms, config 5 4.21 / 3.95 / 3.83 / 3.71 ms.  Scratch: 2.7 KB per row
// = 1.6 GB at 32 tiles, 4.9 GB at 96 (only calls that large allocate it).  UMNN_B200_BWD_TILES = 8..256 overrides.
int bwd_max_tiles() {
    if (const char* e = getenv("UMNN_B200_BWD_TILES")) {
        const int v = atoi(e);
        if (v >= 8 && v <= 256) return v;
    }
    return kBwdMaxTiles;
}

// UMNN_B200_BWD_PANELS = auto (default) | hi_head | hi | a_hilo | hilo.  auto: hi-only panels -- except the two rank-1
// head panels DZ_J and DZ_{J+1}, which keep their lo part (BwdPanels) -- once the call fills at least one whole 32-tile
// chunk of rows (kBwdHiOnlyRows: the regime where the scratch traffic costs time), hi + lo everywhere below (small
// calls are latency bound, the second part is free there and keeps the weight gradient at fp32 grade).
BwdPanels bwd_panels(long long total_rows) {
    const char* e = getenv("UMNN_B200_BWD_PANELS");
    if (e && strcmp(e, "hilo") == 0) return BwdPanels{2, 2, 2};
    if (e && strcmp(e, "a_hilo") == 0) return BwdPanels{2, 1, 2};
    if (e && strcmp(e, "hi_head") == 0) return BwdPanels{1, 1, 2};
    if (e && strcmp(e, "hi") == 0) return BwdPanels{1, 1, 1};       // every panel hi-only: measurement only (see BwdPanels)
    return total_rows >= kBwdHiOnlyRows ? BwdPanels{1, 1, 2} : BwdPanels{2, 2, 2};
}

struct BwdTcPlan {
    BwdPanels panels;
    TcLayout F;            // forward layout (pass F blobs)
    TcDgradLayout G;
    TcWgradPlan W;
    TcDgradSmem GS;
    long long P;
    int rps, n_cta, tiles, n_pairs_w;
    long long slots_per_cta, chunk_slots, row_block, R_pad;
    size_t panel_bytes[2 * UMNN_MAX_LAYERS + 2];
    size_t off_panel[2 * UMNN_MAX_LAYERS + 2], off_mask[UMNN_MAX_LAYERS], off_v, off_part, off_flag, total_bytes;
    size_t w_smem;
    int w_stages;
};

const char* make_bwd_plan(const umnn_desc* d, BwdTcPlan* B) {
    const bool two = tc_two_segments_public();
    if (!make_tc_layout(d, &B->F, two)) return "needs >= 2 hidden layers of width <= 254";
    if (!make_tc_dgrad_layout(d, &B->G, two)) return "input width (1 + E) above 62 or hidden width above 254";
    B->panels = bwd_panels(d->n_samples * (long long)d->n_dims * (d->nb_steps + 3));
    if (!make_tc_wgrad_plan(B->G, &B->W, B->panels)) return "weight-gradient accumulators exceed 512 TMEM columns";
    B->rps = d->nb_steps + 3;
    const TcSmem FS = make_tc_smem(B->F, B->rps, d->nb_steps);
    if (FS.total > kTcMaxSmem) return "forward weights + per-tile context do not fit in 227 KB of shared memory";
    B->GS = make_tc_dgrad_smem(B->G, d->n_ctx, d->nb_steps);
    if (B->GS.total > kTcMaxSmem) return "transposed weights do not fit in 227 KB of shared memory";
    // pass W pipeline: as many 16-row blocks per stage (<= 4) as still leave >= 3 stages in shared memory
    // (UMNN_B200_WGRAD_KBS = 1..4 pins the count: A/B measurements)
    {
        int forced = 0;
        if (const char* e = getenv("UMNN_B200_WGRAD_KBS")) forced = atoi(e);
        int best = 0;
        for (int kbs = 4; kbs >= 1; --kbs) {
            if (forced >= 1 && forced <= 4 && kbs != forced) continue;
            tc_wgrad_set_stage(&B->W, kbs);
            const int stages = (int)((kTcMaxSmem - 2048) / B->W.stage_bytes);
            if (stages >= 3 || (kbs == 1 && stages >= 2) || (forced == kbs && stages >= 2)) { best = kbs; break; }
        }
        if (!best) return "weight-gradient stages do not fit in 227 KB of shared memory";
        B->w_stages = (int)((kTcMaxSmem - 2048) / B->W.stage_bytes);
        if (B->w_stages > kWMaxStages) B->w_stages = kWMaxStages;
    }
    B->w_smem = (size_t)B->W.stage_bytes * B->w_stages + 1024;
    B->P = 0;
    for (int l = 0; l < d->n_layers; ++l) B->P += (long long)d->widths[l] * d->widths[l + 1] + d->widths[l + 1];

    // chunk geometry: n_cta CTAs x `tiles` tiles of 128 rows, whole slots per CTA with the least padding
    const long long n_slots = d->n_samples * (long long)d->n_dims;
    const long long total_rows = n_slots * B->rps;
    long long n_cta = (total_rows + kTcTile - 1) / kTcTile;
    if (n_cta > 148) n_cta = 148;
    if (n_cta > n_slots) n_cta = n_slots;
    n_cta = (n_cta + 1) / 2 * 2;
    if (n_cta < 2) n_cta = 2;
    B->n_cta = (int)n_cta;
    long long s_max = ((long long)bwd_max_tiles() * kTcTile) / B->rps;  // slots that fit in the largest row block
    if (s_max < 1) return "a slot has more rows than one CTA's chunk (Q too large for the tensor-core backward)";
    const long long need = (n_slots + n_cta - 1) / n_cta;               // slots per CTA if everything went in one chunk
    // Several chunks: BALANCED -- as many chunks as the largest row block demands, all of the same size.  With chunks
    // of the maximum size and a remainder, the last chunk kept only part of the CTAs busy for a whole chunk's time
    // (config 3: 5.25 chunks' worth of rows took 6 chunk times).
    long long s_best = need;
    if (need > s_max) {
        const long long n_chunks = (need + s_max - 1) / s_max;
        s_best = (need + n_chunks - 1) / n_chunks;
    }
    if (s_best < 1) s_best = 1;
    B->slots_per_cta = s_best;
    B->tiles = (int)((s_best * B->rps + kTcTile - 1) / kTcTile);
    B->chunk_slots = s_best * n_cta;
    B->row_block = (long long)B->tiles * kTcTile;
    B->R_pad = B->row_block * n_cta;
    B->n_pairs_w = 74;
    if ((long long)B->n_pairs_w > B->R_pad / 16) B->n_pairs_w = (int)(B->R_pad / 16);
    if (B->n_pairs_w < 1) B->n_pairs_w = 1;

    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    for (int pn = 0; pn < B->W.n_panels; ++pn) {
        B->panel_bytes[pn] = (size_t)B->R_pad * B->W.panel_width[pn] * 2 * B->W.panel_parts[pn];
        B->off_panel[pn] = take(B->panel_bytes[pn]);
    }
    for (int j = 1; j <= B->G.J; ++j) B->off_mask[j] = take((size_t)B->R_pad * 8 * 4);
    B->off_v = take((size_t)B->R_pad * 4);
    B->off_part = take((size_t)B->n_pairs_w * B->P * 4);
    B->off_flag = take(sizeof(int));
    B->total_bytes = off + 256;
    return nullptr;
}

}  // namespace

const char* backward_tc_unsupported_reason(const umnn_desc* d) {
    BwdTcPlan B;
    return make_bwd_plan(d, &B);
}

size_t backward_tc_workspace_bytes(const umnn_desc* d) {
    BwdTcPlan B;
    if (make_bwd_plan(d, &B)) return 0;
    return B.total_bytes;
}

size_t backward_tc_packed_bytes(const umnn_desc* d) {
    BwdTcPlan B;
    if (make_bwd_plan(d, &B)) return 0;
    return 2 * (size_t)B.F.blob_bytes + 2 * (size_t)B.G.blob_bytes;
}

int launch_pack_backward_tc(const umnn_desc* d, const float* flat, void* packed, bool with_forward, cudaStream_t s) {
    BwdTcPlan B;
    const char* why = make_bwd_plan(d, &B);
    if (why) { set_error("BF16X3 backward: %s", why); return UMNN_ERR_UNSUPPORTED; }
    if (with_forward) {
        const int rc = launch_pack_tc(d, flat, packed, UMNN_OPF_BF16, s);   // forward blobs first
        if (rc) return rc;
    }
    uint8_t* g = (uint8_t*)packed + 2 * (size_t)B.F.blob_bytes;
    const uint32_t n_w = B.G.weights_bytes;
    pack_dgrad_weights_kernel<<<(n_w + 255) / 256, 256, 0, s>>>(flat, g, B.G);
    UMNN_CUDA_TRY(cudaGetLastError());
    pack_dgrad_consts_kernel<<<(B.G.P[B.G.J] + 255) / 256, 256, 0, s>>>(flat, g, B.G);
    UMNN_CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_backward_tc(const umnn_desc* d, const float* x0, const float* x, const float* h, const void* packed,
                       const float* nodes, const float* weights, const float* grad_out, const float* grad_fx,
                       float* d_x0, float* d_x, float* d_h, float* d_params, void* workspace, size_t workspace_bytes,
                       const void* fwd_blobs_fp16, int* flag, const float* rerun_fp32_packed, cudaStream_t s) {
    BwdTcPlan B;
    const char* why = make_bwd_plan(d, &B);
    if (why) { set_error("BF16X3 backward: %s", why); return UMNN_ERR_UNSUPPORTED; }
    if (!workspace || workspace_bytes < B.total_bytes) {
        set_error("umnn_cc_backward: workspace of %zu bytes needed, %zu given", B.total_bytes, workspace_bytes);
        return UMNN_ERR_WORKSPACE;
    }
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    const int J = B.G.J;
    const long long n_slots = d->n_samples * (long long)d->n_dims;
    const uint8_t* fwd_blobs = (const uint8_t*)packed;
    const uint8_t* dgrad_blobs = fwd_blobs + 2 * (size_t)B.F.blob_bytes;

    TcEmit emit{};
    for (int j = 0; j <= J; ++j) {
        emit.a[j] = ws + B.off_panel[panel_A(j)];
        emit.width[j] = B.G.P[j];
        if (j >= 1) emit.mask[j] = reinterpret_cast<uint32_t*>(ws + B.off_mask[j]);
    }
    emit.v = reinterpret_cast<float*>(ws + B.off_v);
    emit.row_block = B.row_block;
    emit.parts = B.panels.a_parts;
    emit.r_pad = B.R_pad;

    TcDgradParams g{};
    g.x0 = x0; g.x = x; g.weights = weights; g.grad_out = grad_out; g.grad_fx = grad_fx;
    g.blobs = dgrad_blobs; g.v = emit.v;
    for (int j = 1; j <= J; ++j) g.mask[j] = emit.mask[j];
    for (int j = 1; j <= J + 1; ++j) {
        g.dz[j] = ws + B.off_panel[panel_DZ(j, J)];
    }
    g.d_x0 = d_x0; g.d_x = d_x; g.d_h = d_h;
    g.slots_per_cta = B.slots_per_cta; g.row_block = B.row_block; g.tiles_per_cta = B.tiles;
    g.D = d->n_dims; g.E = d->n_ctx; g.layout = d->layout; g.Q = d->nb_steps; g.rps = B.rps; g.out_act = d->out_act;
    g.L = B.G; g.S = B.GS;
    g.dz_parts = B.panels.dz_parts;
    g.dz_head_parts = B.panels.dz_head_parts;
    g.r_pad = B.R_pad;

    TcWgradParams w{};
    for (int pn = 0; pn < B.W.n_panels; ++pn) w.panel[pn] = ws + B.off_panel[pn];
    w.part = reinterpret_cast<float*>(ws + B.off_part);
    w.P = B.P;
    w.n_blocks = B.R_pad / 16;
    w.W = B.W;
    w.n_stages = B.w_stages;
    for (int l = 0; l < d->n_layers; ++l) { w.src_w_off[l] = B.G.src_w_off[l]; w.src_b_off[l] = B.G.src_b_off[l]; }

    auto dkern = d->hidden_act == UMNN_ACT_LEAKY_RELU ? cc_dgrad_tc_kernel<UMNN_ACT_LEAKY_RELU> : cc_dgrad_tc_kernel<UMNN_ACT_RELU>;
    int dev = 0, n_sm_unused = 0;
    UMNN_CUDA_TRY(current_device(&dev, &n_sm_unused));
    UMNN_CUDA_TRY(ensure_dynamic_smem((const void*)dkern, dev, (int)B.GS.total));
    UMNN_CUDA_TRY(ensure_dynamic_smem((const void*)cc_wgrad_tc_kernel, dev, (int)B.w_smem));

    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;

    // fwd_blobs_fp16 != NULL: pass F re-evaluates the network with fp16 hi/lo operands (22 bits: far fewer
    // LeakyReLU units flip side against the fp32 forward than with bf16's ~17).  Its panels stay bf16 (split from
    // the same fp32 activations): one tcgen05.mma cannot take an fp16 and a bf16 operand (illegal instruction on
    // B200), and the dz panels need bf16's exponent range.  An activation beyond the fp16 range raises the flag;
    // the whole backward is then repeated by a second sequence of launches that are no-ops while the flag is clear:
    // the FP32 backward (rerun_fp32_packed given: the rare case gets the parity anchor's arithmetic) or, for shapes
    // the FP32 backward cannot hold in shared memory, the same passes with bf16 operands.
    if (fwd_blobs_fp16 && !flag) { set_error("umnn_cc_backward: FP16X3 needs the flag workspace"); return UMNN_ERR_WORKSPACE; }
    const int n_attempts = (fwd_blobs_fp16 && !rerun_fp32_packed) ? 2 : 1;
    if (fwd_blobs_fp16) UMNN_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), s));
    for (int attempt = 0; attempt < n_attempts; ++attempt) {
        const bool fp16_act = fwd_blobs_fp16 && attempt == 0;
        const int* run_if = (fwd_blobs_fp16 && attempt == 1) ? flag : nullptr;
        g.run_if = run_if;
        w.run_if = run_if;
        bool first = true;
        for (long long s0 = 0; s0 < n_slots; s0 += B.chunk_slots) {
            const long long cs = (n_slots - s0 < B.chunk_slots) ? (n_slots - s0) : B.chunk_slots;
            // pass F
            int rc = launch_forward_tc_emit(d, x0, x, h, fp16_act ? (const uint8_t*)fwd_blobs_fp16 : fwd_blobs, nodes, weights, s0, cs,
                                            B.slots_per_cta, B.tiles, B.n_cta, emit, fp16_act ? UMNN_OPF_FP16 : UMNN_OPF_BF16,
                                            run_if, fp16_act ? flag : nullptr, 1, s);
            if (rc) return rc;
            // pass D
            g.slot0 = s0;
            g.n_slots = cs;
            {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3((unsigned)B.n_cta);
                cfg.blockDim = dim3(kThreads);
                cfg.dynamicSmemBytes = B.GS.total;
                cfg.stream = s;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                UMNN_CUDA_TRY(cudaLaunchKernelEx(&cfg, dkern, g));
            }
            // pass W
            if (d_params) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3((unsigned)(2 * B.n_pairs_w));
                cfg.blockDim = dim3(kWThreads);
                cfg.dynamicSmemBytes = B.w_smem;
                cfg.stream = s;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                UMNN_CUDA_TRY(cudaLaunchKernelEx(&cfg, cc_wgrad_tc_kernel, w));
                reduce_partials_tc_kernel<<<(unsigned)((B.P + 255) / 256), 256, 0, s>>>(w.part, B.P, B.n_pairs_w, d_params,
                                                                                       first ? 0 : 1, run_if);
                UMNN_CUDA_TRY(cudaGetLastError());
            }
            first = false;
        }
    }
    if (fwd_blobs_fp16 && rerun_fp32_packed)
        return launch_backward_fp32(d, x0, x, h, rerun_fp32_packed, nodes, weights, grad_out, grad_fx, d_x0, d_x, d_h, d_params,
                                    workspace, workspace_bytes, s, flag, workspace_bytes);
    return 0;
}

}  // namespace umnn
