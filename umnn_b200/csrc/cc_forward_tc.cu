// Fused Clenshaw-Curtis forward on tcgen05 for sm_100a: 16-bit hi/lo operand split (template parameter OPF: fp16 for
// UMNN_PREC_FP16X3 -- the product path, guarded by a device flag and an FP32 re-run -- or bf16 for the diagnostic
// UMNN_PREC_BF16X3).  With EMIT != 0 the same kernel is pass F of the tensor-core backward (cc_backward_tc.cu).
//
// One persistent CTA PAIR (cluster of 2, tcgen05 cta_group::2) per two SMs.  Every CTA owns a contiguous
// range of whole slots and walks its (slot, node) rows in tiles of 128; the pair runs in lock-step so
// one elected thread of the leader CTA issues every MMA for both (M = 256).
//
//   weights      all hidden-to-hidden matrices, split into hi + lo, live in SHARED MEMORY for the
//                whole kernel (each CTA holds half of the N rows of every matrix; cta_group::2 reads
//                both halves), staged once per CTA by a bulk-TMA copy.
//   activations  live in TENSOR MEMORY only: region P = columns [0,256), Q = [256,512).  An MMA layer
//                reads its A operand (hi/lo pairs) from one region and accumulates fp32 into the
//                other; the epilogue converts the accumulator IN PLACE into the next layer's A operand
//                (16 fp32 columns -> 8 hi + 8 lo packed columns), chunk by chunk, and the next layer's
//                K-block MMAs start as soon as their chunk is converted.
//   per K block  D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo      (fp32-class products)
//   layer 1      rank-1 on CUDA cores: act(x_node * w1x + c_slot), c_slot = b1 + W1h.h_slot once per slot
//   output layer dot product on CUDA cores while the last accumulator is read, ELU+1, CC weight,
//                deterministic segmented sum per slot, (z*(xT-x0))/2.
//
// Warp roles per CTA (wide shape, 640 threads): warps 0-15 epilogue (TMEM lane quadrant = warp%4; column group
// warp/4 selects which 32-column pairs of K blocks it converts; the LAST column group also finalizes the tile's
// rows, fed by the others through "partials full / empty" named barriers), warps 16-18 "prep" (context gather,
// abscissae, c_slot and -- pass F -- the A_0 panel, running tiles ahead), warp 19 MMA issuer (converged warp,
// uniform datapath, one elected lane of the leader CTA).  Narrow shape: 8 epilogue warps, two CTAs per SM.
#include "tc_common.cuh"
#include "tc_layout.cuh"
#include "tc_bwd_layout.cuh"
#include "tc_kernels.cuh"

#include <stdlib.h>

namespace umnn {

namespace {

using namespace tc;

// Kernel shape.
//   WIDE    one CTA per SM, 16 epilogue warps, TMEM regions of 256 columns (hidden widths <= 254).
//   NARROW  two CTAs per SM (two CTA pairs per SM pair), 8 epilogue warps each, TMEM regions of 128 columns
//           (hidden widths <= 126).  At narrow widths the MMA -> epilogue -> MMA chain of one tile is latency
//           bound (an MMA layer lasts ~0.6 us, an epilogue stage ~1.7 us), so a second resident tile fills the
//           tensor pipe while the first one converts.
template <bool NARROW>
struct Shape {
    static constexpr int kEpiWarps = NARROW ? 8 : 16;        // 2 or 4 per TMEM lane quadrant
    static constexpr int kColGroups = kEpiWarps / 4;         // warp/4 handles 32-column pairs p % kColGroups == warp/4
    static constexpr int kPrepWarps = 3;
    static constexpr int kMmaWarp = kEpiWarps + kPrepWarps;
    static constexpr int kThreads = (kMmaWarp + 1) * 32;     // 640 / 384
    static constexpr int kEpiThreads = kEpiWarps * 32;
    static constexpr int kPrepThreads = kPrepWarps * 32;
    static constexpr uint32_t kRegion = NARROW ? kTcNarrowRegionCols : kTcRegionCols;
    static constexpr uint32_t kColP = 0, kColQ = kRegion;
    static constexpr uint32_t kTmemCols = 2 * kRegion;
    // NARROW: 12 warps x 2 CTAs -> 80 registers per thread.  (Registers are handed out to groups of 4 warps: an
    // 11-warp CTA with 88 registers is budgeted as 12 warps and only ONE of them fits an SM -- measured.)
    static constexpr int kCtasPerSm = NARROW ? 2 : 1;
};

// barrier indices
constexpr int BAR_READY = 0;                                   // [layer][8]  A-operand pair p of MMA layer m is in TMEM
constexpr int BAR_ACC = BAR_READY + kTcMaxMmaLayers * 8;       // [layer][2]  accumulator segment s of layer m complete
constexpr int BAR_PREP_FULL = BAR_ACC + kTcMaxMmaLayers * 2;   // [kTcPrepBufs]
constexpr int BAR_PREP_EMPTY = BAR_PREP_FULL + kTcPrepBufs;    // [kTcPrepBufs]
constexpr int BAR_WLOAD = BAR_PREP_EMPTY + kTcPrepBufs;
constexpr int BAR_PEER = BAR_WLOAD + 1;
constexpr int BAR_COUNT = BAR_PEER + 1;
static_assert(BAR_COUNT <= kTcNumBars, "barrier table too small");

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int N_THREADS>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N_THREADS) : "memory"); }
template <int N_THREADS>
__device__ __forceinline__ void prep_bar_sync() { asm volatile("bar.sync 2, %0;" ::"n"(N_THREADS) : "memory"); }
// producer / consumer barriers of the finalize step (PTX ISA: bar.arrive + bar.sync)
constexpr int BARID_FIN_FULL = 3, BARID_FIN_EMPTY = 4, BARID_FIN_GROUP = 5;
template <int ID, int N_THREADS>
__device__ __forceinline__ void named_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N_THREADS) : "memory"); }
template <int ID, int N_THREADS>
__device__ __forceinline__ void named_arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(N_THREADS) : "memory"); }

template <int HIDDEN_ACT>
__device__ __forceinline__ float hact(float v) {
    return HIDDEN_ACT == UMNN_ACT_LEAKY_RELU ? fmaxf(v, v * kLeakySlope) : fmaxf(v, 0.0f);
}

// two 16-byte granules (16 columns) of a bf16 panel: o[0..7] = hi pairs, o[8..15] = lo pairs (PARTS == 2 only)
template <int PARTS>
__device__ __forceinline__ void emit16(const PanelRow& R, int col, const uint32_t (&o)[16]) {
    uint8_t* g0 = R.base + panel_granule(R, col);
    uint8_t* g1 = R.base + panel_granule(R, col + 8);
    *reinterpret_cast<uint4*>(g0) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4*>(g1) = make_uint4(o[4], o[5], o[6], o[7]);
    if constexpr (PARTS == 2) {
        *reinterpret_cast<uint4*>(g0 + R.lo_off) = make_uint4(o[8], o[9], o[10], o[11]);
        *reinterpret_cast<uint4*>(g1 + R.lo_off) = make_uint4(o[12], o[13], o[14], o[15]);
    }
}

// One pair of activations -> hi / lo operand words in the launch's operand format (o_hi, o_lo) and, for pass F,
// the bf16 words of the panel (p_hi and, with EMIT == 2, p_lo; tcgen05 cannot mix an fp16 with a bf16 operand in one
// MMA, so the panels that pass W multiplies with the bf16 dz panels are bf16 in every mode).  EMIT = 0: no panels,
// 1: hi-only panels, 2: hi + lo panels.  TRACK: keep the largest |activation| converted to fp16 (networks whose
// hidden activation would swallow a NaN: ReLU).
template <int OPF, int EMIT, bool TRACK>
__device__ __forceinline__ void split_pair(float2 a, uint32_t& o_hi, uint32_t& o_lo, uint32_t& p_hi, uint32_t& p_lo, float& amax) {
    split2<OPF>(a, o_hi, o_lo);
    if constexpr (EMIT != 0) {
        if constexpr (OPF == UMNN_OPF_BF16) { p_hi = o_hi; p_lo = o_lo; }
        else if constexpr (EMIT == 2) split2<UMNN_OPF_BF16>(a, p_hi, p_lo);
        else p_hi = pack_bf16x2(a.x, a.y);
    }
    if constexpr (TRACK) amax = fmaxf(fmaxf(amax, fabsf(a.x)), fabsf(a.y));
}

// hidden activation of a pair: the slope multiply is one packed instruction
template <int HIDDEN_ACT>
__device__ __forceinline__ float2 hact2(float2 v) {
    if constexpr (HIDDEN_ACT == UMNN_ACT_LEAKY_RELU) {
        const float2 t = fmul2(v, make_float2(kLeakySlope, kLeakySlope));
        return make_float2(fmaxf(v.x, t.x), fmaxf(v.y, t.y));
    } else {
        return make_float2(fmaxf(v.x, 0.0f), fmaxf(v.y, 0.0f));
    }
}

template <int HIDDEN_ACT, int EMIT, bool NARROW, int OPF>
__global__ void __launch_bounds__(Shape<NARROW>::kThreads, Shape<NARROW>::kCtasPerSm)
cc_forward_tc_kernel(const __grid_constant__ TcParams p) {
    using C = Shape<NARROW>;
    constexpr int kEpiWarps = C::kEpiWarps, kColGroups = C::kColGroups, kPrepWarps = C::kPrepWarps;
    constexpr int kMmaWarp = C::kMmaWarp, kThreads = C::kThreads, kEpiThreads = C::kEpiThreads, kPrepThreads = C::kPrepThreads;
    constexpr uint32_t kColP = C::kColP, kColQ = C::kColQ;
    static_assert(!(EMIT != 0 && NARROW), "pass F runs the wide shape");
    constexpr int kParts = EMIT == 2 ? 2 : 1;   // parts of the activation panels (pass F)
    // fp16 operands overflow above 65504.  With LeakyReLU the resulting inf / -inf pair turns every unit of the next
    // layer -- and from there the row's output -- into NaN, which the finalize step sees for free; ReLU would map
    // that NaN to 0, so ReLU networks track the largest converted value instead.
    constexpr bool kTrack = (OPF == UMNN_OPF_FP16) && (HIDDEN_ACT == UMNN_ACT_RELU);
    // guarded re-run (see launch_forward_tc): nothing to do unless the first attempt raised the flag
    if (p.run_if != nullptr && *p.run_if != p.epoch) return;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte aligned base, by pointer arithmetic on the __shared__ array so that the compiler keeps the
    // address space (LDS/STS instead of generic LD/ST on every table and scratch access)
#if defined(UMNN_TC_SMEM_GENERIC) && UMNN_TC_SMEM_GENERIC
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
#else
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
#endif

    const int tid = threadIdx.x, lane = tid & 31;
    // broadcast from lane 0: the compiler then knows the warp index -- and every role branch on it -- is warp-uniform
    // (uniform registers for descriptors and addresses inside the MMA-issuer branch)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t rank = cluster_ctarank();
    const TcLayout& L = p.L;
    const int T = p.tiles_per_cta;

    float* cvec = reinterpret_cast<float*>(smem + p.S.off_cvec);      // [bufs][max_slots][npad1]
    float* hbuf = reinterpret_cast<float*>(smem + p.S.off_hbuf);      // [max_slots][h_stride]
    float* xnode = reinterpret_cast<float*>(smem + p.S.off_xnode);    // [bufs][128]
    int* lsrel = reinterpret_cast<int*>(smem + p.S.off_lsrel);        // [bufs][128]
    int* nodeid = reinterpret_cast<int*>(smem + p.S.off_node);        // [bufs][128]
    float* part = reinterpret_cast<float*>(smem + p.S.off_part);      // [3][128]
    float* fval = reinterpret_cast<float*>(smem + p.S.off_fval);      // [128]
    float* carry2 = reinterpret_cast<float*>(smem + p.S.off_carry);   // [2]
    float* tab_t = reinterpret_cast<float*>(smem + p.S.off_tabt);
    float* tab_w = reinterpret_cast<float*>(smem + p.S.off_tabw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.S.off_bars);
    uint32_t* holder = reinterpret_cast<uint32_t*>(smem + p.S.off_holder);
    const float* w1x = reinterpret_cast<const float*>(smem + L.off_w1x);
    const float* b1p = reinterpret_cast<const float*>(smem + L.off_b1p);
    const float* w1h = reinterpret_cast<const float*>(smem + L.off_w1h);
    const float* w4 = reinterpret_cast<const float*>(smem + L.off_w4);

    const long long slot_begin = p.slot0 + (long long)blockIdx.x * p.slots_per_cta;
    long long slot_end = slot_begin + p.slots_per_cta;
    if (slot_end > p.slot0 + p.n_slots) slot_end = p.slot0 + p.n_slots;
    const long long n_rows = slot_end > slot_begin ? (slot_end - slot_begin) * p.rps : 0;

    // ---------------------------------------------------------------- setup
    if (tid == 0) {
        for (int m = 0; m < kTcMaxMmaLayers; ++m) {
            for (int j = 0; j < 8; ++j) mbar_init(&bars[BAR_READY + m * 8 + j], 8);     // 4 quadrant warps x 2 CTAs
            for (int s = 0; s < 2; ++s) mbar_init(&bars[BAR_ACC + m * 2 + s], 1);       // tcgen05.commit
        }
        for (int b = 0; b < kTcPrepBufs; ++b) {
            mbar_init(&bars[BAR_PREP_FULL + b], kPrepWarps);
            mbar_init(&bars[BAR_PREP_EMPTY + b], kEpiWarps);
        }
        mbar_init(&bars[BAR_WLOAD], 1);
        mbar_init(&bars[BAR_PEER], 2);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc<2>(holder, C::kTmemCols);
    for (int i = tid; i <= p.Q; i += kThreads) {
        tab_t[i] = p.nodes[i];
        tab_w[i] = p.weights[i];
    }
    // context rows: [h_0 .. h_{E-1} | 1.0 | 0 ...]; the prep warps only ever rewrite the first E entries
    const int Hs = p.S.h_stride;
    for (int i = tid; i < p.S.max_slots * Hs; i += kThreads) hbuf[i] = ((i % Hs) == p.E) ? 1.0f : 0.0f;
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tbase = *holder;
    if (warp == kMmaWarp) {
        // =========================================================== MMA issuer warp
        // The whole warp runs this code converged (uniform control flow keeps descriptors and TMEM
        // addresses in uniform registers); one elected lane issues the tcgen05 instructions.
        if (elect_one_sync()) {
            // stage this CTA's parameter blob with bulk-TMA copies (<= 32 KB each) on one mbarrier
            mbar_expect_tx(&bars[BAR_WLOAD], L.blob_bytes);
            const uint8_t* src = p.blobs + (size_t)rank * L.blob_bytes;
            for (uint32_t off = 0; off < L.blob_bytes; off += 32768u) {
                const uint32_t n = (L.blob_bytes - off < 32768u) ? (L.blob_bytes - off) : 32768u;
                bulk_g2s(smem + off, src + off, n, &bars[BAR_WLOAD]);
            }
        }
        __syncwarp();
        mbar_wait(&bars[BAR_WLOAD], 0, 100);
        if (elect_one_sync()) mbar_arrive_cluster(&bars[BAR_PEER], 0);   // this CTA's half of B is resident
        __syncwarp();
        // lane-0 broadcasts: values the compiler cannot see as warp-uniform on its own (special-register reads through
        // asm, a shared-memory load) become uniform-register material
        const uint32_t rank_u = __shfl_sync(0xffffffffu, rank, 0);
        const uint32_t tbase = __shfl_sync(0xffffffffu, *holder, 0);
        if (NARROW) {
            // two CTA pairs share an SM pair here: the pair-collective allocation must have handed both CTAs of THIS
            // pair the same columns (one tcgen05.mma addresses the tensor memory of both).  Checked by the whole warp
            // with a warp-uniform outcome (a divergent branch here would cost the issuer its uniform registers).
            uint32_t peer_base;
            asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(peer_base) : "r"(map_to_cta(smem_u32(holder), rank ^ 1u)) : "memory");
            if (__any_sync(0xffffffffu, peer_base != tbase)) asm volatile("trap;");
        }
        if (rank_u == 0) {
            mbar_wait_warp(&bars[BAR_PEER], 0, 101);
            const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
            for (int t = 0; t < T; ++t) {
                const uint32_t par = (uint32_t)(t & 1);
                for (int m = 0; m < L.n_mma; ++m) {
                    const TcMmaLayer& y = L.layer[m];
                    const uint32_t a_base = tbase + ((m & 1) ? kColP : kColQ);   // A operand region
                    const uint32_t col_d = (m & 1) ? kColQ : kColP;             // accumulator region
                    const int n_kb = y.kpad / 16;
                    MmaSegment g[2];
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const uint32_t lbo = (uint32_t)(y.seg_n[s] / 16) * 128u;
                        g[s].idesc = make_idesc_f32(256, y.seg_n[s], OPF, OPF);
                        g[s].step = (uint64_t)((2u * lbo) >> 4);                  // descriptor address field is in 16-byte units
                        g[s].d_addr = tbase + col_d + (uint32_t)y.seg_begin[s];
                        g[s].bhi = make_smem_desc(sbase + y.b_off[0][s], lbo, 128);
                        g[s].blo = make_smem_desc(sbase + y.b_off[1][s], lbo, 128);
                    }
                    issue_mma_layer(n_kb, y.nseg, g[0], g[1], a_base, &bars[BAR_READY + m * 8], par, &bars[BAR_ACC + m * 2], 200 + m * 8);
                }
            }
        }
        __syncwarp();
    } else if (warp >= kEpiWarps) {
        // =========================================================== prep warps
        // Per tile: context of the slots it touches, abscissa / slot / node of every row, (pass F) the A_0 panel, and
        // c_slot = b1 + W1h . h_slot.  All row arithmetic is 32-bit (a CTA's rows fit: checked by the launcher); 64-bit
        // divisions are subroutine calls on this ISA and the prep warps are the pacing resource of pass F.
        const int ptid = tid - kEpiThreads;
        const uint32_t n_rows32 = (uint32_t)n_rows, rps32 = (uint32_t)p.rps;
        const long long n_begin = p.layout == UMNN_LAYOUT_STRIDED_D ? slot_begin / p.D : 0;   // the one 64-bit division
        const uint32_t d_begin = (uint32_t)(slot_begin - n_begin * p.D);
        mbar_wait(&bars[BAR_WLOAD], 0, 110);
        for (int u = 0; u < T; ++u) {
            const int b = u % kTcPrepBufs;
            if (u >= kTcPrepBufs) mbar_wait(&bars[BAR_PREP_EMPTY + b], (uint32_t)((u / kTcPrepBufs - 1) & 1), 120 + b);
            const uint32_t row0 = (uint32_t)u * kTcTile;
            const uint32_t ls_first = row0 / rps32;
            const bool live = row0 < n_rows32;
            int ns = 1;
            if (live) {
                const uint32_t last = (row0 + kTcTile < n_rows32 ? row0 + kTcTile : n_rows32) - 1;
                ns = (int)(last / rps32 - ls_first) + 1;
            }
            // context of the slots this tile touches -> shared memory.  STRIDED_D stores h as [B][E][D]: the slots of a
            // tile are consecutive in d, so the slot index runs fastest over the threads -- neighbouring lanes read
            // neighbouring floats of one 32-byte sector (the e-major order fetched a sector per lane).  CONTIG ([N][E])
            // is contiguous in e.
            if (live) {
                if (p.layout == UMNN_LAYOUT_STRIDED_D) {
                    for (int idx = ptid; idx < ns * p.E; idx += kPrepThreads) {
                        const int e = idx / ns, i = idx - e * ns;
                        // slot -> (sample, dimension) from the CTA's first slot with 32-bit arithmetic
                        const uint32_t dd = d_begin + ls_first + (uint32_t)i, dn = dd / (uint32_t)p.D;
                        const float* hp = p.h + (n_begin + dn) * (long long)p.E * p.D + (dd - dn * (uint32_t)p.D);
                        hbuf[i * Hs + e] = __ldg(hp + (long long)e * p.D);
                    }
                } else {
                    const float* hp = p.h + (slot_begin + ls_first) * (long long)p.E;      // ns * E contiguous floats
                    for (int idx = ptid; idx < ns * p.E; idx += kPrepThreads) {
                        const int i = idx / p.E, e = idx - i * p.E;
                        hbuf[i * Hs + e] = __ldg(hp + idx);
                    }
                }
            }
            for (int r = ptid; r < kTcTile; r += kPrepThreads) {
                const uint32_t row = row0 + r;
                float xi = 0.0f;
                int rel = 0, node = -1;
                if (row < n_rows32) {
                    const uint32_t ls = row / rps32;
                    node = (int)(row - ls * rps32);
                    rel = (int)(ls - ls_first);
                    const long long slot = slot_begin + ls;
                    const float lo = p.x0 ? __ldg(p.x0 + slot) : 0.0f;
                    const float hi = __ldg(p.x + slot);
                    if (node <= p.Q) {
                        const float xT = upper_limit(lo, hi, p.Q);
                        xi = node_abscissa(lo, __fsub_rn(xT, lo), tab_t[node]);
                    } else if (node == p.Q + 1 && p.x_row) {
                        xi = hi;
                    } else {
                        xi = lo;
                    }
                }
                xnode[b * kTcTile + r] = xi;
                lsrel[b * kTcTile + r] = rel;
                nodeid[b * kTcTile + r] = node;
            }
            prep_bar_sync<kPrepThreads>();
            if (EMIT) {
                // A_0 = [x_row, h_slot, 1, 0...] as bf16 hi / lo (operand of the first layer's weight gradient); the raw
                // inputs stay bf16 in every mode: their range is the caller's, not the network's.  Columns 1.. of a row
                // are the slot's hbuf row as it stands, so an 8-column granule is 8 shared loads and 4 conversions.
                const int W0 = p.emit.width[0];                       // == Hs
                const int n_gran = W0 / 8;                            // 2, 4, 6 or 8: divides the 96 prep threads
                const int gidx = ptid % n_gran, r_step = kPrepThreads / n_gran;
                for (int r = ptid / n_gran; r < kTcTile; r += r_step) {
                    const long long pr = (long long)blockIdx.x * p.emit.row_block + row0 + r;
                    const bool valid = nodeid[b * kTcTile + r] >= 0;
                    const float* hv = hbuf + lsrel[b * kTcTile + r] * Hs + gidx * 8;
                    float v8[8];
                    v8[0] = gidx == 0 ? xnode[b * kTcTile + r] : hv[-1];
#pragma unroll
                    for (int q2 = 1; q2 < 8; ++q2) v8[q2] = hv[q2 - 1];
                    uint32_t hi4[4], lo4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) split_bf16x2(valid ? v8[2 * i] : 0.0f, valid ? v8[2 * i + 1] : 0.0f, hi4[i], lo4[i]);
                    *reinterpret_cast<uint4*>(p.emit.a[0] + panel_offset(pr, gidx * 8, W0, 0, kParts, p.emit.r_pad)) = make_uint4(hi4[0], hi4[1], hi4[2], hi4[3]);
                    if constexpr (EMIT == 2)
                        *reinterpret_cast<uint4*>(p.emit.a[0] + panel_offset(pr, gidx * 8, W0, 1, kParts, p.emit.r_pad)) = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
                }
            }
            // c_slot = b1' + W1h . h_slot: four columns per thread, context and weights read 16 bytes at a time
            // (1.3 instructions per multiply-add instead of 3).  Summation order per element: e ascending, one fma
            // chain -- the same bits as the scalar loop.
            float* cv = cvec + (size_t)b * p.S.max_slots * L.npad1;
            const int n4g = L.npad1 >> 2;
            for (int idx = ptid; idx < ns * n4g; idx += kPrepThreads) {
                const int i = idx / n4g, n = 4 * (idx - i * n4g);
                float4 acc = *reinterpret_cast<const float4*>(b1p + n);
                if (live) {
                    const float* hv = hbuf + i * Hs;
                    const float* wp = w1h + n;
                    int e = 0;
                    for (; e + 4 <= p.E; e += 4) {
                        const float4 hh = *reinterpret_cast<const float4*>(hv + e);
                        const float4 wa = *reinterpret_cast<const float4*>(wp + (e + 0) * L.npad1);
                        const float4 wb = *reinterpret_cast<const float4*>(wp + (e + 1) * L.npad1);
                        const float4 wc = *reinterpret_cast<const float4*>(wp + (e + 2) * L.npad1);
                        const float4 wd = *reinterpret_cast<const float4*>(wp + (e + 3) * L.npad1);
                        acc.x = fmaf(wa.x, hh.x, acc.x); acc.y = fmaf(wa.y, hh.x, acc.y); acc.z = fmaf(wa.z, hh.x, acc.z); acc.w = fmaf(wa.w, hh.x, acc.w);
                        acc.x = fmaf(wb.x, hh.y, acc.x); acc.y = fmaf(wb.y, hh.y, acc.y); acc.z = fmaf(wb.z, hh.y, acc.z); acc.w = fmaf(wb.w, hh.y, acc.w);
                        acc.x = fmaf(wc.x, hh.z, acc.x); acc.y = fmaf(wc.y, hh.z, acc.y); acc.z = fmaf(wc.z, hh.z, acc.z); acc.w = fmaf(wc.w, hh.z, acc.w);
                        acc.x = fmaf(wd.x, hh.w, acc.x); acc.y = fmaf(wd.y, hh.w, acc.y); acc.z = fmaf(wd.z, hh.w, acc.z); acc.w = fmaf(wd.w, hh.w, acc.w);
                    }
                    for (; e < p.E; ++e) {
                        const float hh = hv[e];
                        const float4 wa = *reinterpret_cast<const float4*>(wp + e * L.npad1);
                        acc.x = fmaf(wa.x, hh, acc.x); acc.y = fmaf(wa.y, hh, acc.y); acc.z = fmaf(wa.z, hh, acc.z); acc.w = fmaf(wa.w, hh, acc.w);
                    }
                }
                *reinterpret_cast<float4*>(cv + i * L.npad1 + n) = acc;
            }
            prep_bar_sync<kPrepThreads>();   // hbuf is rewritten by the next tile
            if (lane == 0) mbar_arrive(&bars[BAR_PREP_FULL + b]);
        }
    } else {
        // =========================================================== epilogue warps
        const int q = warp & 3, cg = warp >> 2;
        const int r = q * 32 + lane;                      // row of the tile owned by this thread
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const int n_mma = L.n_mma;
        const bool even_layers = (n_mma & 1) == 0;        // last accumulator shares region Q with layer 1's output
        const int pairs1 = (L.npad1 + 31) / 32, pairsL = (L.npadL + 31) / 32;
        const uint32_t col_last = ((n_mma - 1) & 1) ? kColQ : kColP;   // accumulator region of the last MMA layer
        const TcMmaLayer& ylast = L.layer[n_mma - 1];
        float amax = 0.0f;     // largest |activation| this thread turned into an fp16 operand

        mbar_wait(&bars[BAR_WLOAD], 0, 130);

        const long long cta_row0 = (long long)blockIdx.x * (EMIT ? p.emit.row_block : 0);
        // one 16-column half of layer 1 for the tile `tile` (prep buffer `bu`) -> A operand of MMA layer 0
        // (region Q); returns the sign bits of the 16 activations
        auto l1_half = [&](int bu, int tile, int c16) -> uint32_t {
            const float xn = xnode[bu * kTcTile + r];
            const float* cv = cvec + ((size_t)bu * p.S.max_slots + lsrel[bu * kTcTile + r]) * L.npad1 + 16 * c16;
            const float* wx = w1x + 16 * c16;
            uint32_t o[16], ob[16], pre[16];
            // 16-byte shared loads (both tables are 64-byte aligned per 16 columns; the lanes of a warp read the same
            // or two different slots, i.e. broadcasts)
            const float4* cv4 = reinterpret_cast<const float4*>(cv);
            const float4* wx4 = reinterpret_cast<const float4*>(wx);
            const float2 xn2 = make_float2(xn, xn);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 c = cv4[j], w = wx4[j];
                const float2 va = ffma2(xn2, make_float2(w.x, w.y), make_float2(c.x, c.y));
                const float2 vb = ffma2(xn2, make_float2(w.z, w.w), make_float2(c.z, c.w));
                split_pair<OPF, EMIT, kTrack>(hact2<HIDDEN_ACT>(va), o[2 * j], o[8 + 2 * j], ob[2 * j], ob[8 + 2 * j], amax);
                split_pair<OPF, EMIT, kTrack>(hact2<HIDDEN_ACT>(vb), o[2 * j + 1], o[9 + 2 * j], ob[2 * j + 1], ob[9 + 2 * j], amax);
                if (EMIT) {
                    pre[4 * j] = __float_as_uint(va.x); pre[4 * j + 1] = __float_as_uint(va.y);
                    pre[4 * j + 2] = __float_as_uint(vb.x); pre[4 * j + 3] = __float_as_uint(vb.y);
                }
            }
            tmem_st16(tbase + lane_sel + kColQ + 16u * c16, o);
            if (EMIT) {
                emit16<kParts>(panel_row(p.emit.a[1], cta_row0 + (long long)tile * kTcTile + r, L.npad1, kParts, p.emit.r_pad), 16 * c16, ob);
                return sign_mask16(pre);
            }
            return 0u;
        };
        // publish pair `pp` of MMA layer `m`'s A operand (both CTAs arrive on the leader's barrier)
        auto publish = [&](int m, int pp) {
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&bars[BAR_READY + m * 8 + pp], 0);
        };
        auto l1_pair = [&](int bu, int tile, int pp) {
            uint32_t bits = l1_half(bu, tile, 2 * pp);
            if (32 * pp + 16 < L.npad1) bits |= l1_half(bu, tile, 2 * pp + 1) << 4;
            if (EMIT) p.emit.mask[1][(long long)pp * p.emit.r_pad + cta_row0 + (long long)tile * kTcTile + r] = bits;
            publish(0, pp);
        };

        // prologue: layer 1 of tile 0
        mbar_wait(&bars[BAR_PREP_FULL + 0], 0, 131);
        for (int pp = cg; pp < pairs1; pp += kColGroups) l1_pair(0, 0, pp);

        for (int t = 0; t < T; ++t) {
            const uint32_t par = (uint32_t)(t & 1);
            const int b = t % kTcPrepBufs;
            const bool has_next = (t + 1 < T);
            const int bn = (t + 1) % kTcPrepBufs;

            // ---- accumulator of MMA layer m -> A operand of layer m+1, in place
            for (int m = 0; m + 1 < n_mma; ++m) {
                const TcMmaLayer& y = L.layer[m];
                const uint32_t col_d = (m & 1) ? kColQ : kColP;
                const int n_pairs = (y.npad + 31) / 32;
                for (int pp = cg; pp < n_pairs; pp += kColGroups) {
                    const int s = (y.nseg == 2 && 32 * pp >= y.seg_begin[1]) ? 1 : 0;
                    mbar_wait(&bars[BAR_ACC + m * 2 + s], par, 300 + m * 2 + s);
                    tc_fence_after_sync();
                    const uint32_t taddr = tbase + lane_sel + col_d + 32u * pp;
                    const bool two = 32 * pp + 16 < y.npad;
                    uint32_t v0[16], v1[16], o[16], ob[16];
                    tmem_ld16(taddr, v0);
                    if (two) tmem_ld16(taddr + 16, v1);
                    tmem_ld_wait();
                    const long long pr = cta_row0 + (long long)t * kTcTile + r;
                    PanelRow prow;
                    if (EMIT) {
                        prow = panel_row(p.emit.a[m + 2], pr, y.npad, kParts, p.emit.r_pad);
                        // signs first: the accumulator registers die as they are converted below
                        p.emit.mask[m + 2][(long long)pp * p.emit.r_pad + pr] = sign_mask16(v0) | (two ? sign_mask16(v1) << 4 : 0u);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        split_pair<OPF, EMIT, kTrack>(hact2<HIDDEN_ACT>(make_float2(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1]))),
                                                      o[i], o[8 + i], ob[i], ob[8 + i], amax);
                    tmem_st16(taddr, o);
                    if (EMIT) emit16<kParts>(prow, 32 * pp, ob);
                    if (two) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            split_pair<OPF, EMIT, kTrack>(hact2<HIDDEN_ACT>(make_float2(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1]))),
                                                          o[i], o[8 + i], ob[i], ob[8 + i], amax);
                        tmem_st16(taddr + 16, o);
                        if (EMIT) emit16<kParts>(prow, 32 * pp + 16, ob);
                    }
                    publish(m + 1, pp);
                }
            }

            // ---- last accumulator: output layer dot product; even layer counts also write layer 1 of the
            //      next tile over the pair just consumed
            float partial = 0.0f;
            if (has_next && even_layers) mbar_wait(&bars[BAR_PREP_FULL + bn], (uint32_t)(((t + 1) / kTcPrepBufs) & 1), 132);
            const int n_loop = pairsL > pairs1 ? pairsL : pairs1;
            for (int pp = cg; pp < n_loop; pp += kColGroups) {
                if (pp < pairsL) {
                    const int s = (ylast.nseg == 2 && 32 * pp >= ylast.seg_begin[1]) ? 1 : 0;
                    mbar_wait(&bars[BAR_ACC + (n_mma - 1) * 2 + s], par, 310 + s);
                    tc_fence_after_sync();
                    const uint32_t taddr = tbase + lane_sel + col_last + 32u * pp;
                    const bool two = 32 * pp + 16 < L.npadL;
                    uint32_t v0[16], v1[16];
                    tmem_ld16(taddr, v0);
                    if (two) tmem_ld16(taddr + 16, v1);
                    tmem_ld_wait();
                    const float4* wv4 = reinterpret_cast<const float4*>(w4 + 32 * pp);   // 128-byte aligned
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 w = wv4[j];
                        const float2 a = hact2<HIDDEN_ACT>(make_float2(__uint_as_float(v0[4 * j]), __uint_as_float(v0[4 * j + 1])));
                        const float2 b = hact2<HIDDEN_ACT>(make_float2(__uint_as_float(v0[4 * j + 2]), __uint_as_float(v0[4 * j + 3])));
                        partial = fmaf(a.x, w.x, partial);
                        partial = fmaf(a.y, w.y, partial);
                        partial = fmaf(b.x, w.z, partial);
                        partial = fmaf(b.y, w.w, partial);
                    }
                    if (two) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 w = wv4[4 + j];
                            const float2 a = hact2<HIDDEN_ACT>(make_float2(__uint_as_float(v1[4 * j]), __uint_as_float(v1[4 * j + 1])));
                            const float2 b = hact2<HIDDEN_ACT>(make_float2(__uint_as_float(v1[4 * j + 2]), __uint_as_float(v1[4 * j + 3])));
                            partial = fmaf(a.x, w.x, partial);
                            partial = fmaf(a.y, w.y, partial);
                            partial = fmaf(b.x, w.z, partial);
                            partial = fmaf(b.y, w.w, partial);
                        }
                    }
                    if (EMIT) {
                        // last hidden activations a_J (operand of the output layer's weight gradient) and their signs
                        const long long pr = cta_row0 + (long long)t * kTcTile + r;
                        const PanelRow prow = panel_row(p.emit.a[n_mma + 1], pr, L.npadL, kParts, p.emit.r_pad);
                        uint32_t o[16];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float e0 = hact<HIDDEN_ACT>(__uint_as_float(v0[2 * i])), e1 = hact<HIDDEN_ACT>(__uint_as_float(v0[2 * i + 1]));
                            if constexpr (EMIT == 2) split_bf16x2(e0, e1, o[i], o[8 + i]);
                            else o[i] = pack_bf16x2(e0, e1);
                        }
                        emit16<kParts>(prow, 32 * pp, o);
                        if (two) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float e0 = hact<HIDDEN_ACT>(__uint_as_float(v1[2 * i])), e1 = hact<HIDDEN_ACT>(__uint_as_float(v1[2 * i + 1]));
                                if constexpr (EMIT == 2) split_bf16x2(e0, e1, o[i], o[8 + i]);
                                else o[i] = pack_bf16x2(e0, e1);
                            }
                            emit16<kParts>(prow, 32 * pp + 16, o);
                        }
                        p.emit.mask[n_mma + 1][(long long)pp * p.emit.r_pad + pr] = sign_mask16(v0) | (two ? sign_mask16(v1) << 4 : 0u);
                    }
                } else if (even_layers) {
                    // columns beyond the last accumulator: free once every MMA of this tile is done
                    for (int s = 0; s < ylast.nseg; ++s) mbar_wait(&bars[BAR_ACC + (n_mma - 1) * 2 + s], par, 320 + s);
                    tc_fence_after_sync();
                }
                if (even_layers && has_next && pp < pairs1) l1_pair(bn, t + 1, pp);
            }

            // ---- finalize the rows of this tile: the LAST column group alone (it converts the fewest pairs per layer),
            //      behind producer / consumer barriers ("partials full" / "partials empty").  The other column groups hand
            //      over their partial dot products and go straight on to the next tile: with two CTA-wide barriers around
            //      the finalize step every warp waited for it, and on narrow networks the first MMA layer of the next
            //      tile is shorter than that wait.  Same order of additions as before: same bits.
            constexpr int kFin = kColGroups - 1;
            if (cg != kFin) {
                if (t > 0) named_sync<BARID_FIN_EMPTY, kEpiThreads>();      // the previous tile's partials have been read
                part[cg * kTcTile + r] = partial;
                __threadfence_block();
                named_arrive<BARID_FIN_FULL, kEpiThreads>();
            }
            if (!even_layers) {
                // odd layer counts: the last accumulator lives in P, which MMA layer 0 of the next tile
                // overwrites -> publish layer 1 of the next tile only after every warp has read P
                epi_bar_sync<kEpiThreads>();
                if (has_next) {
                    for (int s = 0; s < ylast.nseg; ++s) mbar_wait(&bars[BAR_ACC + (n_mma - 1) * 2 + s], par, 330 + s);
                    tc_fence_after_sync();
                    mbar_wait(&bars[BAR_PREP_FULL + bn], (uint32_t)(((t + 1) / kTcPrepBufs) & 1), 133);
                    for (int pp = cg; pp < pairs1; pp += kColGroups) l1_pair(bn, t + 1, pp);
                }
            }
            if (cg == kFin) {
                named_sync<BARID_FIN_FULL, kEpiThreads>();
                const int row0 = t * kTcTile;                     // 32-bit row arithmetic (see the prep warps)
                const int n_rows_i = (int)n_rows;
                {
                    const int node = nodeid[b * kTcTile + r];
                    float vtot = kColGroups > 1 ? part[r] : partial;                      // ((p_0 + p_1) + p_2) + p_3
#pragma unroll
                    for (int g = 1; g < kFin; ++g) vtot += part[g * kTcTile + r];
                    if (kColGroups > 1) vtot += partial;
                    if (EMIT) p.emit.v[cta_row0 + row0 + r] = node >= 0 ? vtot : 0.0f;
                    if (node >= 0) {
                        // fp16 operands, LeakyReLU: an overflowed activation has made this row's output NaN (see kTrack)
                        if (OPF == UMNN_OPF_FP16 && !kTrack && p.raise_flag != nullptr && !(fabsf(vtot) <= 3.0e38f)) *p.raise_flag = p.epoch;
                        const float f = out_act(vtot, p.out_act);
                        if (node <= p.Q) {
                            fval[r] = f * tab_w[node];
                        } else if (!EMIT) {
                            const long long slot = slot_begin + (uint32_t)(row0 + r) / (uint32_t)p.rps;
                            if (node == p.Q + 1 && p.x_row) p.out_fx[slot] = f;
                            else p.out_fx0[slot] = f;
                        }
                    }
                }
                if (!EMIT) {
                    named_sync<BARID_FIN_GROUP, 128>();           // fval of the whole tile is there (this group's 4 warps)
                    if (row0 < n_rows_i) {
                        // Per-slot node sums, one slot per warp of the group (slot s_first + q, + 4, ..).  The slot that
                        // straddles two tiles hands its partial sum over through shared memory (double buffered: the
                        // warp that reads tile t's carry-in may run next to the one that writes its carry-out).
                        const int last_row = (row0 + kTcTile < n_rows_i ? row0 + kTcTile : n_rows_i) - 1;
                        const int s_first = (int)((uint32_t)row0 / (uint32_t)p.rps), s_last = (int)((uint32_t)last_row / (uint32_t)p.rps);
                        for (int ls = s_first + q; ls <= s_last; ls += 4) {
                            const int a = ls * p.rps, bb = a + p.Q;
                            const int lo = a > row0 ? a : row0;
                            const int hi = bb < last_row ? bb : last_row;
                            float sum = 0.0f;
                            for (int rr = lo + lane; rr <= hi; rr += 32) sum += fval[rr - row0];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                            if (lo <= hi) {
                                const float total = (a < row0 ? carry2[t & 1] : 0.0f) + sum;
                                if (bb <= last_row) {
                                    if (lane == 0) {
                                        const long long slot = slot_begin + ls;
                                        const float x0v = p.x0 ? p.x0[slot] : 0.0f;
                                        const float span = __fsub_rn(upper_limit(x0v, p.x[slot], p.Q), x0v);
                                        p.out[slot] = __fmul_rn(__fmul_rn(total, span), 0.5f);
                                    }
                                } else if (lane == 0) {
                                    carry2[(t + 1) & 1] = total;
                                }
                            }
                        }
                    }
                }
                __threadfence_block();
                if (has_next) named_arrive<BARID_FIN_EMPTY, kEpiThreads>();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAR_PREP_EMPTY + b]);
        }
        // ReLU networks: an activation beyond the fp16 range became inf in its operand -> ask for the bf16 re-run
        if (kTrack && p.raise_flag != nullptr && amax > kFp16Max) *p.raise_flag = p.epoch;
    }

    // ---------------------------------------------------------------- teardown
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == kMmaWarp) tmem_dealloc<2>(tbase, C::kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// parameter packing
// ---------------------------------------------------------------------------------------------
// 16-bit operand encodings (opf = UMNN_OPF_BF16 / UMNN_OPF_FP16) and their fp32 values
__device__ __forceinline__ uint16_t op_bits(float v, int opf) {
    return (uint16_t)((opf == UMNN_OPF_BF16 ? pack_bf16x2(v, 0.0f) : pack_f16x2(v, 0.0f)) & 0xFFFFu);
}
__device__ __forceinline__ float op_val(float v, int opf) {
    const uint16_t b = op_bits(v, opf);
    return opf == UMNN_OPF_BF16 ? __uint_as_float((uint32_t)b << 16) : __half2float(__ushort_as_half(b));
}

__global__ void pack_tc_weights_kernel(const float* __restrict__ flat, uint8_t* __restrict__ blobs, TcLayout L, int opf) {
    const uint32_t per_rank = L.weights_bytes / 2;  // bf16 elements
    const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= 2 * per_rank) return;
    const uint32_t rank = gi / per_rank;
    const uint32_t byte_off = (gi - rank * per_rank) * 2;
    // locate (layer, part, segment)
    int m = 0, part = 0, s = 0;
    bool found = false;
    for (int mm = 0; mm < L.n_mma && !found; ++mm)
        for (int pp = 0; pp < 2 && !found; ++pp)
            for (int ss = 0; ss < L.layer[mm].nseg && !found; ++ss) {
                const uint32_t lo = L.layer[mm].b_off[pp][ss];
                const uint32_t sz = (uint32_t)(L.layer[mm].seg_n[ss] / 2) * L.layer[mm].kpad * 2;
                if (byte_off >= lo && byte_off < lo + sz) { m = mm; part = pp; s = ss; found = true; }
            }
    const TcMmaLayer& y = L.layer[m];
    const uint32_t idx = (byte_off - y.b_off[part][s]) / 2;
    const int rows_cta = y.seg_n[s] / 2, n8c = rows_cta / 8;
    const int per_kb = 2 * n8c * 64;
    const int kb = idx / per_kb;
    int rem = idx - kb * per_kb;
    const int k8 = rem / (n8c * 64);
    rem -= k8 * n8c * 64;
    const int n8 = rem / 64;
    const int rr = (rem % 64) / 8, kk = rem % 8;
    const int n = y.seg_begin[s] + (int)rank * rows_cta + n8 * 8 + rr;
    const int k = kb * 16 + k8 * 8 + kk;
    const int lin = m + 1;  // Linear layer index in the flat vector
    float hi = 0.0f, lo = 0.0f;
    if (n < y.h_out) {
        if (k < y.h_in) {
            const float w = flat[L.src_w_off[lin] + n * y.h_in + k];
            hi = op_val(w, opf);
            lo = op_val(w - hi, opf);
        } else if (k == y.h_in || k == y.h_in + 1) {
            const float bv = flat[L.src_b_off[lin] + n];
            const float b_hi = op_val(bv, opf);
            const float b_lo = op_val(bv - b_hi, opf);
            if (k == y.h_in) { hi = b_hi; lo = b_lo; }
            else { hi = op_val(bv - b_hi - b_lo, opf); lo = 0.0f; }
        }
    } else if ((n == y.h_out && k == y.h_in) || (n == y.h_out + 1 && k == y.h_in + 1)) {
        hi = 1.0f;   // bias carriers propagate the constant 1
    }
    reinterpret_cast<uint16_t*>(blobs + (size_t)rank * L.blob_bytes)[byte_off / 2] = op_bits(part == 0 ? hi : lo, opf);
}

__global__ void pack_tc_consts_kernel(const float* __restrict__ flat, uint8_t* __restrict__ blobs, TcLayout L, int n_layers) {
    const uint32_t n_f = (L.blob_bytes - L.weights_bytes) / 4;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_f) return;
    const uint32_t byte_off = L.weights_bytes + 4 * i;
    float v = 0.0f;
    const int nin0 = 1 + L.E;
    if (byte_off >= L.off_w4) {
        const int n = (byte_off - L.off_w4) / 4;
        const int last = n_layers - 1;
        if (n < L.hL) v = flat[L.src_w_off[last] + n];
        else if (n == L.hL) v = flat[L.src_b_off[last]];
    } else if (byte_off >= L.off_w1h) {
        const int idx = (byte_off - L.off_w1h) / 4;
        const int e = idx / L.npad1, n = idx % L.npad1;
        if (n < L.h1 && e < L.E) v = flat[L.src_w_off[0] + n * nin0 + 1 + e];
    } else if (byte_off >= L.off_b1p) {
        const int n = (byte_off - L.off_b1p) / 4;
        if (n < L.h1) v = flat[L.src_b_off[0] + n];
        else if (n == L.h1 || n == L.h1 + 1) v = 1.0f;
    } else {
        const int n = (byte_off - L.off_w1x) / 4;
        if (n < L.h1) v = flat[L.src_w_off[0] + n * nin0];
    }
    for (int rank = 0; rank < 2; ++rank) *reinterpret_cast<float*>(blobs + (size_t)rank * L.blob_bytes + byte_off) = v;
}

bool tc_two_segments() {
    const char* e = getenv("UMNN_B200_TC_SEGMENTS");
    return !(e && e[0] == '1');
}

bool tc_narrow_enabled() {
    const char* e = getenv("UMNN_B200_TC_NARROW");
    return !(e && e[0] == '0');
}

template <int EMIT, bool NARROW, int OPF>
int launch_tc_kernel(int hidden_act, const TcParams& p, int n_cta, cudaStream_t s) {
    auto kern = hidden_act == UMNN_ACT_LEAKY_RELU ? cc_forward_tc_kernel<UMNN_ACT_LEAKY_RELU, EMIT, NARROW, OPF>
                                                  : cc_forward_tc_kernel<UMNN_ACT_RELU, EMIT, NARROW, OPF>;
    int dev = 0, n_sm_unused = 0;
    UMNN_CUDA_TRY(current_device(&dev, &n_sm_unused));
    UMNN_CUDA_TRY(ensure_dynamic_smem((const void*)kern, dev, (int)p.S.total));
    // two CTAs per SM need the full shared-memory carveout; without the hint the driver sizes it for one CTA
    if (NARROW) UMNN_CUDA_TRY(ensure_max_carveout((const void*)kern, dev));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)n_cta);
    cfg.blockDim = dim3(Shape<NARROW>::kThreads);
    cfg.dynamicSmemBytes = p.S.total;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    UMNN_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
    return 0;
}

}  // namespace

const char* tc_unsupported_reason(const umnn_desc* d, int extra_rows) {
    TcLayout L;
    if (!make_tc_layout(d, &L, tc_two_segments())) return "needs >= 2 hidden layers of width <= 254";
    const int rps = d->nb_steps + 1 + extra_rows;
    const TcSmem S = make_tc_smem(L, rps, d->nb_steps);
    if (S.total > kTcMaxSmem) return "weights + per-tile context do not fit in 227 KB of shared memory";
    return nullptr;
}

int launch_pack_tc(const umnn_desc* d, const float* flat, void* packed, int opf, cudaStream_t s) {
    TcLayout L;
    if (!make_tc_layout(d, &L, tc_two_segments())) {
        set_error("BF16X3: shape not supported by the tensor-core kernel");
        return UMNN_ERR_UNSUPPORTED;
    }
    const uint32_t n_w = L.weights_bytes;  // = 2 ranks x weights_bytes/2 elements
    pack_tc_weights_kernel<<<(n_w + 255) / 256, 256, 0, s>>>(flat, (uint8_t*)packed, L, opf);
    UMNN_CUDA_TRY(cudaGetLastError());
    const uint32_t n_f = (L.blob_bytes - L.weights_bytes) / 4;
    pack_tc_consts_kernel<<<(n_f + 255) / 256, 256, 0, s>>>(flat, (uint8_t*)packed, L, d->n_layers);
    UMNN_CUDA_TRY(cudaGetLastError());
    return 0;
}

size_t tc_packed_bytes(const umnn_desc* d) {
    TcLayout L;
    if (!make_tc_layout(d, &L, tc_two_segments())) return 0;
    return 2 * (size_t)L.blob_bytes;
}

int launch_forward_tc(const umnn_desc* d, const float* x0, const float* x, const float* h, const void* packed,
                      const float* nodes, const float* weights, float* out, float* out_fx, float* out_fx0,
                      int opf, const int* run_if, int* raise_flag, int epoch, cudaStream_t s) {
    TcParams p{};
    if (!make_tc_layout(d, &p.L, tc_two_segments())) {
        set_error("tensor-core forward: shape not supported");
        return UMNN_ERR_UNSUPPORTED;
    }
    p.x0 = x0; p.x = x; p.h = h; p.nodes = nodes; p.weights = weights; p.blobs = (const uint8_t*)packed;
    p.out = out; p.out_fx = out_fx; p.out_fx0 = out_fx0;
    p.run_if = run_if; p.raise_flag = raise_flag; p.epoch = epoch;
    p.n_slots = d->n_samples * (long long)d->n_dims;
    p.D = d->n_dims; p.E = d->n_ctx; p.layout = d->layout; p.Q = d->nb_steps; p.out_act = d->out_act;
    p.rps = d->nb_steps + 1 + (out_fx ? 1 : 0) + (out_fx0 ? 1 : 0);
    p.x_row = out_fx ? 1 : 0;
    p.slot0 = 0;
    p.S = make_tc_smem(p.L, p.rps, p.Q);
    if (p.S.total > kTcMaxSmem) {
        set_error("tensor-core forward: needs %u bytes of shared memory (max %zu)", p.S.total, kTcMaxSmem);
        return UMNN_ERR_UNSUPPORTED;
    }
    if (p.n_slots == 0) return 0;
    // narrow shape: two co-resident CTAs per SM, each with half of the tensor memory
    const bool narrow = tc_narrow_enabled() && tc_layout_is_narrow(p.L) && p.S.total <= kTcNarrowMaxSmem;
    int dev = 0, n_sm = 0;
    UMNN_CUDA_TRY(current_device(&dev, &n_sm));
    const long long total_rows = p.n_slots * p.rps;
    long long n_cta = (total_rows + kTcTile - 1) / kTcTile;
    const long long cap = (long long)(n_sm / 2) * 2 * (narrow ? 2 : 1);
    if (n_cta > cap) n_cta = cap;
    if (n_cta > p.n_slots) n_cta = p.n_slots;
    n_cta = (n_cta + 1) / 2 * 2;             // whole CTA pairs
    if (n_cta < 2) n_cta = 2;
    p.slots_per_cta = (p.n_slots + n_cta - 1) / n_cta;
    if (p.slots_per_cta * p.rps >= (1LL << 31) - kTcTile) {       // the kernel's row arithmetic is 32-bit
        set_error("tensor-core forward: %lld rows per CTA exceed the 32-bit row index", p.slots_per_cta * p.rps);
        return UMNN_ERR_UNSUPPORTED;
    }
    p.tiles_per_cta = (int)((p.slots_per_cta * p.rps + kTcTile - 1) / kTcTile);

    if (opf == UMNN_OPF_FP16)
        return narrow ? launch_tc_kernel<0, true, UMNN_OPF_FP16>(d->hidden_act, p, (int)n_cta, s)
                      : launch_tc_kernel<0, false, UMNN_OPF_FP16>(d->hidden_act, p, (int)n_cta, s);
    return narrow ? launch_tc_kernel<0, true, UMNN_OPF_BF16>(d->hidden_act, p, (int)n_cta, s)
                  : launch_tc_kernel<0, false, UMNN_OPF_BF16>(d->hidden_act, p, (int)n_cta, s);
}

// diagnostic behind umnn_tc_forward_occupancy: which shape serves the descriptor and how many CTAs of it one SM holds
int tc_forward_occupancy(const umnn_desc* d, int extra_rows, int* narrow_out, int* ctas_per_sm) {
    TcLayout L;
    if (!make_tc_layout(d, &L, tc_two_segments())) {
        set_error("tensor-core forward: shape not supported");
        return UMNN_ERR_UNSUPPORTED;
    }
    const TcSmem S = make_tc_smem(L, d->nb_steps + 1 + extra_rows, d->nb_steps);
    if (S.total > kTcMaxSmem) {
        set_error("tensor-core forward: needs %u bytes of shared memory (max %zu)", S.total, kTcMaxSmem);
        return UMNN_ERR_UNSUPPORTED;
    }
    const bool narrow = tc_narrow_enabled() && tc_layout_is_narrow(L) && S.total <= kTcNarrowMaxSmem;
    if (narrow_out) *narrow_out = narrow ? 1 : 0;
    if (!ctas_per_sm) return 0;          // shape selection only: host arithmetic, no device needed
    const void* kern = narrow ? (const void*)cc_forward_tc_kernel<UMNN_ACT_LEAKY_RELU, 0, true, UMNN_OPF_FP16>
                              : (const void*)cc_forward_tc_kernel<UMNN_ACT_LEAKY_RELU, 0, false, UMNN_OPF_FP16>;
    int dev = 0, n_sm = 0;
    UMNN_CUDA_TRY(current_device(&dev, &n_sm));
    UMNN_CUDA_TRY(ensure_dynamic_smem(kern, dev, (int)S.total));      // through the memo: it must know what the attribute is
    if (narrow) UMNN_CUDA_TRY(ensure_max_carveout(kern, dev));
    // the kernel is launched as clusters of 2: ask how many clusters the device holds at once
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2048);
    cfg.blockDim = dim3(narrow ? Shape<true>::kThreads : Shape<false>::kThreads);
    cfg.dynamicSmemBytes = S.total;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n_clusters = 0;
    UMNN_CUDA_TRY(cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg));
    *ctas_per_sm = n_sm > 0 ? (2 * n_clusters) / n_sm : 0;
    return 0;
}

// pass F of the tensor-core backward: the forward kernel over one chunk of slots with panel emission
int launch_forward_tc_emit(const umnn_desc* d, const float* x0, const float* x, const float* h, const void* packed,
                           const float* nodes, const float* weights, long long slot0, long long n_slots_chunk,
                           long long slots_per_cta, int tiles_per_cta, int n_cta, const TcEmit& emit, int opf,
                           const int* run_if, int* raise_flag, int epoch, cudaStream_t s) {
    TcParams p{};
    if (!make_tc_layout(d, &p.L, tc_two_segments())) {
        set_error("BF16X3: shape not supported by the tensor-core kernel");
        return UMNN_ERR_UNSUPPORTED;
    }
    p.x0 = x0; p.x = x; p.h = h; p.nodes = nodes; p.weights = weights; p.blobs = (const uint8_t*)packed;
    p.out = nullptr; p.out_fx = nullptr; p.out_fx0 = nullptr;
    p.run_if = run_if; p.raise_flag = raise_flag; p.epoch = epoch;
    p.slot0 = slot0; p.n_slots = n_slots_chunk; p.slots_per_cta = slots_per_cta; p.tiles_per_cta = tiles_per_cta;
    p.D = d->n_dims; p.E = d->n_ctx; p.layout = d->layout; p.Q = d->nb_steps; p.out_act = d->out_act;
    p.rps = d->nb_steps + 3;
    p.x_row = 1;
    p.S = make_tc_smem(p.L, p.rps, p.Q);
    p.emit = emit;
    if (p.S.total > kTcMaxSmem) {
        set_error("BF16X3 backward: needs %u bytes of shared memory (max %zu)", p.S.total, kTcMaxSmem);
        return UMNN_ERR_UNSUPPORTED;
    }
    if (emit.parts == 2)
        return opf == UMNN_OPF_FP16 ? launch_tc_kernel<2, false, UMNN_OPF_FP16>(d->hidden_act, p, n_cta, s)
                                    : launch_tc_kernel<2, false, UMNN_OPF_BF16>(d->hidden_act, p, n_cta, s);
    return opf == UMNN_OPF_FP16 ? launch_tc_kernel<1, false, UMNN_OPF_FP16>(d->hidden_act, p, n_cta, s)
                                : launch_tc_kernel<1, false, UMNN_OPF_BF16>(d->hidden_act, p, n_cta, s);
}

bool tc_two_segments_public() { return tc_two_segments(); }

}  // namespace umnn
