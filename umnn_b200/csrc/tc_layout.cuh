// Packed-parameter / shared-memory layout of the BF16x3 tensor-core path (UMNN_PREC_BF16X3).
// Shared by the pack kernel, the forward launcher and the kernel itself.
//
// The integrand MLP  [1+E] -> H1 -> H2 -> ... -> HL -> 1  is split into
//   layer 1      (1+E -> H1)   CUDA cores, rank-1 form:  pre = x_node * w1x + (b1 + W1h . h_slot)
//   MMA layers   (H_{l} -> H_{l+1}, l = 1..L-1)   tcgen05, BF16 hi/lo split (3 MMAs per K block)
//   output layer (HL -> 1)     CUDA cores while the last accumulator is read out of TMEM
//
// Every hidden width is padded to npad = round_up(H + 2, 16).  The two extra units H and H+1 are
// "bias carriers": they hold the constant 1.0 through the whole network (B row H has a single 1.0
// at column k = H_prev, row H+1 at k = H_prev + 1; LeakyReLU/ReLU map 1 -> 1; the bf16 split of 1.0
// is exact), so a layer's bias rides inside the GEMM: B_hi[n][H_prev] = bf16(b), B_lo[n][H_prev] =
// bf16(b - hi), B_hi[n][H_prev + 1] = bf16(b - hi - lo)  -> 24 bits of bias, no epilogue add.
//
// An MMA layer's N range is cut into 1 or 2 segments (each a multiple of 16) so the epilogue of the
// first segment overlaps the MMAs of the second.  With cta_group::2 each CTA of the pair holds HALF of
// every segment's B rows (rank r: rows [seg_begin + r*seg_n/2, +seg_n/2)).  Per (layer, part, segment)
// the CTA's rows are stored as K-major core matrices  [kb][k8 (2)][n8][8 rows][8 bf16]  so that one
// K=16 instruction reads a contiguous 2 * (seg_n/16) * 128-byte block
// (descriptor: LBO = (seg_n/16)*128 between the two K halves, SBO = 128 between 8-row groups).
//
// Per-CTA blob (copied verbatim to shared memory offset 0 by one bulk-TMA copy):
//   [ MMA layer 0: hi seg0 | hi seg1 | lo seg0 | lo seg1 ][ MMA layer 1 ... ]
//   [ w1x[npad1] | b1p[npad1] | W1h[E][npad1] | w4[npadL] ]   (fp32; b1p carries 1.0 at H1, H1+1;
//                                                              w4[HL] = output bias)
#pragma once

#include "umnn_common.cuh"

namespace umnn {

constexpr int kTcMaxMmaLayers = UMNN_MAX_LAYERS - 2;
constexpr int kTcTile = 128;           // rows per CTA tile (256 per CTA pair)
constexpr int kTcPrepBufs = 3;
constexpr int kTcRegionCols = 256;     // TMEM columns per region (P = [0,256), Q = [256,512))
constexpr int kTcNarrowRegionCols = 128;   // "narrow" kernel shape: two CTAs per SM, regions of 128 columns
constexpr size_t kTcMaxSmem = 232448;  // 227 KB
constexpr size_t kTcNarrowMaxSmem = 115712;  // (228 KB - 2 x 1 KB reserved) / 2 CTAs per SM

struct TcMmaLayer {
    int h_in, h_out;      // true widths
    int kpad, npad;       // padded
    int nseg;
    int seg_begin[2], seg_n[2];
    uint32_t b_off[2][2]; // [part hi/lo][segment] byte offset inside the per-CTA blob
};

struct TcLayout {
    int n_mma;                         // number of MMA layers = n_layers - 2
    TcMmaLayer layer[kTcMaxMmaLayers];
    int E, h1, npad1, hL, npadL;
    uint32_t off_w1x, off_b1p, off_w1h, off_w4;   // byte offsets of the fp32 constants in the blob
    uint32_t weights_bytes;            // bytes of the bf16 section
    uint32_t blob_bytes;               // per-CTA blob (multiple of 16)
    int src_w_off[UMNN_MAX_LAYERS], src_b_off[UMNN_MAX_LAYERS];  // offsets into the flat fp32 vector
};

inline int tc_pad(int h) { return round_up(h + 2, 16); }

// false if the shape cannot be served by the tensor-core kernel at all (independent of smem budget)
inline bool make_tc_layout(const umnn_desc* d, TcLayout* L, bool two_segments = true) {
    if (d->n_layers < 3) return false;
    for (int l = 1; l < d->n_layers; ++l)
        if (tc_pad(d->widths[l]) > kTcRegionCols) return false;
    *L = TcLayout{};
    L->n_mma = d->n_layers - 2;
    L->E = d->n_ctx;
    int src = 0;
    for (int l = 0; l < d->n_layers; ++l) {
        L->src_w_off[l] = src;
        src += d->widths[l] * d->widths[l + 1];
        L->src_b_off[l] = src;
        src += d->widths[l + 1];
    }
    L->h1 = d->widths[1];
    L->npad1 = tc_pad(L->h1);
    L->hL = d->widths[d->n_layers - 1];
    L->npadL = tc_pad(L->hL);
    uint32_t off = 0;
    for (int m = 0; m < L->n_mma; ++m) {
        TcMmaLayer& y = L->layer[m];
        y.h_in = d->widths[m + 1];
        y.h_out = d->widths[m + 2];
        y.kpad = tc_pad(y.h_in);
        y.npad = tc_pad(y.h_out);
        // segments are cut at a multiple of 32 columns (the epilogue works on 32-column pairs of K blocks);
        // a second segment narrower than 32 columns is not worth a separate MMA shape
        const int pairs = (y.npad + 31) / 32;
        const int n0 = 32 * ((pairs + 1) / 2);
        if (two_segments && y.npad - n0 >= 32) {
            y.nseg = 2;
            y.seg_n[0] = n0;
            y.seg_n[1] = y.npad - n0;
            y.seg_begin[0] = 0;
            y.seg_begin[1] = n0;
        } else {
            y.nseg = 1;
            y.seg_n[0] = y.npad;
            y.seg_begin[0] = 0;
            y.seg_n[1] = 0;
            y.seg_begin[1] = y.npad;
        }
        for (int part = 0; part < 2; ++part)
            for (int s = 0; s < y.nseg; ++s) {
                y.b_off[part][s] = off;
                off += (uint32_t)(y.seg_n[s] / 2) * y.kpad * 2;
            }
    }
    L->weights_bytes = off;
    L->off_w1x = off;  off += 4u * L->npad1;
    L->off_b1p = off;  off += 4u * L->npad1;
    L->off_w1h = off;  off += 4u * L->npad1 * L->E;
    L->off_w4 = off;   off += 4u * L->npadL;
    L->blob_bytes = (off + 15u) & ~15u;
    return true;
}

// true if every padded width fits the 128-column regions of the narrow kernel shape
inline bool tc_layout_is_narrow(const TcLayout& L) {
    if (L.npad1 > kTcNarrowRegionCols || L.npadL > kTcNarrowRegionCols) return false;
    for (int m = 0; m < L.n_mma; ++m)
        if (L.layer[m].kpad > kTcNarrowRegionCols || L.layer[m].npad > kTcNarrowRegionCols) return false;
    return true;
}

// dynamic shared memory map of the forward kernel (byte offsets from the 1024-aligned base)
struct TcSmem {
    uint32_t off_cvec, off_hbuf, off_xnode, off_lsrel, off_node, off_part, off_fval, off_carry, off_tabt, off_tabw, off_bars, off_holder;
    uint32_t total;
    int max_slots;
    int h_stride;       // floats per slot row of hbuf: [h_0 .. h_{E-1}, 1.0, 0 ...], = round_up(E + 2, 16) = width of panel A_0
};

constexpr int kTcNumBars = kTcMaxMmaLayers * 10 /*ready[layer][8 pairs] + acc_full[layer][2]*/ + 2 * kTcPrepBufs + 2;

inline TcSmem make_tc_smem(const TcLayout& L, int rps, int Q) {
    TcSmem S{};
    S.max_slots = (kTcTile - 1) / rps + 2;   // slots a 128-row window can touch
    uint32_t off = L.blob_bytes;
    S.off_cvec = off;   off += 4u * kTcPrepBufs * S.max_slots * L.npad1;
    // a slot's context row is stored as the tail of its A_0 row ([x | h_0..h_{E-1}, 1, 0..]): 16-byte aligned for the
    // float4 reads of the c_slot product, and pass F copies it into the A_0 panel without per-element branches
    S.h_stride = round_up(L.E + 2, 16);
    S.off_hbuf = off;   off += 4u * S.max_slots * S.h_stride;
    S.off_xnode = off;  off += 4u * kTcPrepBufs * kTcTile;
    S.off_lsrel = off;  off += 4u * kTcPrepBufs * kTcTile;
    S.off_node = off;   off += 4u * kTcPrepBufs * kTcTile;
    S.off_part = off;   off += 4u * 3 * kTcTile;
    S.off_fval = off;   off += 4u * kTcTile;
    S.off_carry = off;  off += 16;          // partial node sum of the slot that straddles two tiles, double buffered
    S.off_tabt = off;   off += 4u * (Q + 1);
    S.off_tabw = off;   off += 4u * (Q + 1);
    off = (off + 7u) & ~7u;
    S.off_bars = off;   off += 8u * kTcNumBars;
    S.off_holder = off; off += 16;
    S.total = off + 1024;   // slack for the 1024-byte alignment of the base
    return S;
}

// operand format of the MMA operands (weights blob and in-TMEM activations): UMNN_OPF_BF16 / UMNN_OPF_FP16
int launch_pack_tc(const umnn_desc* d, const float* flat, void* packed, int opf, cudaStream_t s);
// run_if (device int, may be NULL): the launch is a no-op unless *run_if == epoch.  raise_flag (device int, may be
// NULL; fp16 operands only): set to `epoch` when an activation overflowed the fp16 range.
int launch_forward_tc(const umnn_desc* d, const float* x0, const float* x, const float* h, const void* packed,
                      const float* nodes, const float* weights, float* out, float* out_fx, float* out_fx0,
                      int opf, const int* run_if, int* raise_flag, int epoch, cudaStream_t s);
int tc_forward_occupancy(const umnn_desc* d, int extra_rows, int* narrow_out, int* ctas_per_sm);
size_t tc_packed_bytes(const umnn_desc* d);
// 0 if the tensor-core kernel can serve desc (with `extra_rows` = 0..2 extra rows per slot), else a reason string
const char* tc_unsupported_reason(const umnn_desc* d, int extra_rows);

}  // namespace umnn
