// Layouts of the tensor-core backward (three passes over a chunk of rows, see cc_backward_tc.cu):
//   pass F  forward re-evaluation (cc_forward_tc_kernel<.., EMIT>) -> activation panels A_0..A_J, sign masks, v
//   pass D  dgrad chain with transposed weights -> dz panels DZ_1..DZ_{J+1}, d_h, d_x, d_x0
//   pass W  split-K weight-gradient GEMMs over the panels (tcgen05 SS, MN-major operands)
//
// Hidden layers j = 1..J (J = n_layers - 1), widths H_j padded to P_j = tc_pad(H_j) exactly as in the
// forward (units H_j and H_j + 1 carry the constant 1.0, so column H_j of A_j is the "ones" column that
// turns the bias gradient into one more column of the weight-gradient GEMM).  A_0 = [x_row, h_slot, 1, 0..]
// has width P_0 = round_up(1 + E + 1, 16).
//
// A panel of width W stores, for R_pad rows in blocks of 16 rows, `parts` bf16 values per element: the rounded value
// ("hi", parts = 1) or hi AND the residual lo (parts = 2).  In pass W every panel is one operand of one cta_group::2
// MMA, whose two CTAs each supply half of the columns, so a panel is two half-panels (one per CTA of the pair), each
//   [16-row block][part (parts)][k8 = (row % 16) / 8][col8][row % 8][col % 8]:
// a CTA fetches its half of SEVERAL consecutive blocks (one pipeline stage of pass W) with ONE contiguous bulk-TMA
// copy and every block of it is directly `parts` MN-major UMMA operand tiles (128-byte core matrices;
// LBO = (W/16)*128 between the K halves, SBO = 128).
//
// Which panels carry a lo part is the backward's HBM-traffic / precision trade (BwdPanels below).  The panels only
// feed the WEIGHT gradient dW = sum over rows of dz^T a: per-row rounding errors of a bf16 operand (2^-9 relative,
// independent from row to row) average out over the millions of rows of the sum, so hi-only panels cost the weight
// gradient ~1e-4..1e-3 normwise while halving the scratch (d_h, d_x and d_x0 never read a panel: the dgrad chain
// keeps its 3-term hi/lo products in tensor memory).
#pragma once

#include "tc_layout.cuh"

namespace umnn {

constexpr int kBwdMaxHidden = UMNN_MAX_LAYERS - 1;          // J <= 7
constexpr int kBwdMaxTiles = 96;                            // tiles of 128 rows per CTA and chunk (default; see bwd_max_tiles)
constexpr long long kBwdHiOnlyRows = 148LL * 32 * 128;      // calls of at least this many rows use hi-only operand panels

// bytes of one half-panel: r_pad rows (a multiple of 16) x W/2 columns x parts x 2 bytes
__host__ __device__ inline size_t panel_half_bytes(long long r_pad, int W, int parts) { return (size_t)r_pad * (size_t)(W * parts); }

__host__ __device__ inline size_t panel_offset(long long pr, int c, int W, int part, int parts, long long r_pad) {
    const int hw = W >> 1;                       // columns per CTA half (multiple of 8)
    const int half = c >= hw ? 1 : 0;
    const int cin = c - half * hw;
    return (size_t)half * panel_half_bytes(r_pad, W, parts) + (size_t)(pr >> 4) * (size_t)(32 * hw * parts) + (size_t)part * (size_t)(32 * hw) +
           (size_t)(((pr >> 3) & 1) * (hw >> 3) + (cin >> 3)) * 128 + (size_t)(pr & 7) * 16 + (size_t)(cin & 7) * 2;
}

// Row-resolved view of a panel for the epilogue threads: everything of panel_offset() that depends on the row
// is folded into `base` once per tile, leaving 16*c (+ a constant in the second column half) per 8-column granule.
struct PanelRow {
    uint8_t* base;
    uint32_t hw;           // columns per CTA half
    uint32_t half_off;     // extra offset of the second column half (half-panel bytes - 16*hw; < 4 GiB by the plan)
    uint32_t lo_off;       // offset of the lo part (parts == 2)
};
__host__ __device__ inline PanelRow panel_row(uint8_t* panel, long long pr, int W, int parts, long long r_pad) {
    PanelRow R;
    const uint32_t hw = (uint32_t)W >> 1;
    R.base = panel + (size_t)(pr >> 4) * (size_t)(32 * hw * parts) + (size_t)(((uint32_t)(pr >> 3) & 1u) * hw * 16u + ((uint32_t)pr & 7u) * 16u);
    R.hw = hw;
    R.half_off = (uint32_t)(panel_half_bytes(r_pad, W, parts) - 16u * hw);
    R.lo_off = 32u * hw;
    return R;
}
// byte offset (from PanelRow::base) of the hi granule holding columns c..c+7 (c a multiple of 8)
__host__ __device__ inline uint32_t panel_granule(const PanelRow& R, int c) {
    return 16u * (uint32_t)c + ((uint32_t)c >= R.hw ? R.half_off : 0u);
}

// Sign masks of a 32-column pair: one 32-bit word per (row, pair).  Column c of the pair (0..31) lives at bit
// 8*(c % 4) + (c % 16) / 4 + 4*(c / 16): the layout three byte-permutes per four fp32 values produce.  A set bit
// means the pre-activation is negative (sign bit), i.e. the hidden activation's derivative is its slope; a clear
// bit means derivative 1 (an exactly-zero pre-activation counts as positive here).
__host__ __device__ constexpr int mask_bitpos(int c) { return 8 * (c & 3) + ((c & 15) >> 2) + 4 * (c >> 4); }

struct TcChainLayer {
    int kpad, npad;
    int nseg, seg_begin[2], seg_n[2];
    uint32_t b_off[2][2];
};

// dgrad chain: layer i maps dz_{J-i} (width kpad = P_{J-i}) to da_{J-i-1} (width npad = P_{J-i-1}, or the
// padded input width for the last one) with B'_i[n'][k'] = W_{J-i}[k'][n'] (no bias carriers)
struct TcDgradLayout {
    int J;
    int H[kBwdMaxHidden + 2];   // H[0] = 1 + E (inputs), H[1..J] hidden widths
    int P[kBwdMaxHidden + 2];   // padded panel widths P[0..J]
    int n0pad;                  // padded width of the last dgrad output (inputs), multiple of 32
    TcChainLayer layer[kBwdMaxHidden];
    uint32_t off_wlast;         // fp32 w_{J+1}[P_J] (zero beyond H_J: no bias)
    uint32_t weights_bytes, blob_bytes;
    int src_w_off[UMNN_MAX_LAYERS], src_b_off[UMNN_MAX_LAYERS];
};

inline void tc_cut_segments(TcChainLayer* y, bool two_segments) {
    const int pairs = (y->npad + 31) / 32;
    const int n0 = 32 * ((pairs + 1) / 2);
    if (two_segments && y->npad - n0 >= 32) {
        y->nseg = 2;
        y->seg_begin[0] = 0;  y->seg_n[0] = n0;
        y->seg_begin[1] = n0; y->seg_n[1] = y->npad - n0;
    } else {
        y->nseg = 1;
        y->seg_begin[0] = 0;  y->seg_n[0] = y->npad;
        y->seg_begin[1] = y->npad; y->seg_n[1] = 0;
    }
}

inline bool make_tc_dgrad_layout(const umnn_desc* d, TcDgradLayout* L, bool two_segments = true) {
    if (d->n_layers < 3) return false;
    *L = TcDgradLayout{};
    L->J = d->n_layers - 1;
    int src = 0;
    for (int l = 0; l < d->n_layers; ++l) {
        L->src_w_off[l] = src;
        src += d->widths[l] * d->widths[l + 1];
        L->src_b_off[l] = src;
        src += d->widths[l + 1];
    }
    L->H[0] = d->widths[0];
    L->P[0] = round_up(d->widths[0] + 1, 16);
    for (int j = 1; j <= L->J; ++j) {
        L->H[j] = d->widths[j];
        L->P[j] = tc_pad(d->widths[j]);
        if (L->P[j] > kTcRegionCols) return false;
    }
    L->n0pad = round_up(d->widths[0], 32);
    if (L->n0pad > 64 || L->P[0] > 64) return false;   // the input-gradient tile is reduced through shared memory
    uint32_t off = 0;
    for (int i = 0; i < L->J; ++i) {
        TcChainLayer& y = L->layer[i];
        y.kpad = L->P[L->J - i];
        y.npad = (i < L->J - 1) ? L->P[L->J - i - 1] : L->n0pad;
        tc_cut_segments(&y, two_segments);
        for (int part = 0; part < 2; ++part)
            for (int s = 0; s < y.nseg; ++s) {
                y.b_off[part][s] = off;
                off += (uint32_t)(y.seg_n[s] / 2) * y.kpad * 2;
            }
    }
    L->weights_bytes = off;
    L->off_wlast = off;
    off += 4u * L->P[L->J];
    L->blob_bytes = (off + 15u) & ~15u;
    return true;
}

// shared-memory map of the dgrad kernel
struct TcDgradSmem {
    uint32_t off_dv, off_f, off_node, off_d0, off_carry, off_tabw, off_bars, off_holder, total;
};
constexpr int kTcDgradBars = kBwdMaxHidden * 10 + 2 * kTcPrepBufs + 2;

inline TcDgradSmem make_tc_dgrad_smem(const TcDgradLayout& L, int E, int Q) {
    TcDgradSmem S{};
    uint32_t off = L.blob_bytes;
    S.off_dv = off;    off += 4u * kTcPrepBufs * kTcTile;
    S.off_f = off;     off += 4u * kTcPrepBufs * kTcTile;
    S.off_node = off;  off += 4u * kTcPrepBufs * kTcTile;
    S.off_d0 = off;    off += 4u * kTcTile * (L.n0pad + 1);
    S.off_carry = off; off += 4u * 2 * (E > 0 ? E : 1);
    S.off_tabw = off;  off += 4u * (Q + 1);
    off = (off + 7u) & ~7u;
    S.off_bars = off;  off += 8u * kTcDgradBars;
    S.off_holder = off; off += 16;
    S.total = off + 1024;
    return S;
}

// ---- pass W: one accumulator per Linear layer, all resident in TMEM at once -------------------------
struct TcWgradLayer {
    int m_width, n_width;       // panel widths of the M operand (D rows, over the CTA pair) and the N operand
    int m_panel, n_panel;       // indices into the panel table
    int tmem_col;               // first accumulator column
    int swapped;                // 1: D rows = input units (last Linear layer), 0: D rows = output units
    int lin;                    // Linear layer index in the flat vector
    int n_out, n_in, ones_col;  // true dims; column (or row when swapped) that carries the bias gradient
};

// parts per element of the activation panels (A_0..A_J) and of the dz panels (DZ_1..DZ_{J+1}): 1 = bf16 hi only,
// 2 = hi + lo.  UMNN_B200_BWD_PANELS = auto | hi_head ({1,1,2}) | hi ({1,1,1}) | a_hilo ({2,1,2}) | hilo ({2,2,2}: round 1's scheme).
// dz_head_parts: parts of DZ_J (the rank-1 head dz_J = dv * w_out (.) act'(a_J)) and of DZ_{J+1} (= dv).  Their entries
// are the SAME number for every row that shares dv -- the Jacobian rows of a likelihood have dv = -1/(B f) with f nearly
// constant at initialisation -- so a hi-only panel rounds them all the same way and the error does not average out over
// the rows: 9e-3 of the last hidden layer's weight gradient on a POWER-shaped flow (profiles/r2_panel_precision.txt).
// These two panels therefore keep their lo part whenever the others drop it.
struct BwdPanels { int a_parts, dz_parts, dz_head_parts; };

struct TcWgradPlan {
    int n_layers;               // = J + 1
    TcWgradLayer layer[UMNN_MAX_LAYERS];
    int n_panels;               // panel table: A_0..A_J then DZ_1..DZ_{J+1}
    int panel_width[2 * UMNN_MAX_LAYERS + 2];
    int panel_parts[2 * UMNN_MAX_LAYERS + 2];
    int kb_per_stage;           // 16-row K blocks per pipeline stage (1..4), see tc_wgrad_set_stage
    uint32_t block_bytes[2 * UMNN_MAX_LAYERS + 2];   // bytes of one 16-row block of a CTA's half-panel (16 * W * parts)
    uint32_t stage_bytes;       // bytes staged per stage and CTA (+ slack for the M-tile over-read)
    uint32_t tile_off[2 * UMNN_MAX_LAYERS + 2];      // smem offset of the panel's blocks ([block][part]...) inside a stage
    int tmem_cols_used;
};

// A pipeline stage of pass W holds `kbs` consecutive 16-row blocks of every panel (one bulk-TMA copy per panel and
// stage).  More blocks per stage = fewer barrier round trips and copy requests per row: with one block per stage the
// MMA issuer's per-stage bookkeeping (~1 us) was the bound, not HBM or the tensor pipe (profiles/r2_bwd_passes.md).
inline void tc_wgrad_set_stage(TcWgradPlan* W, int kbs) {
    W->kb_per_stage = kbs;
    uint32_t off = 0;
    for (int p = 0; p < W->n_panels; ++p) {
        W->tile_off[p] = off;
        off += (uint32_t)kbs * W->block_bytes[p];
    }
    // an M tile is read as 128 rows per K half although only W/2 are staged: the tensor core over-reads up to
    // (128 - 8) * 16 bytes past the last K half of the last block -> keep that much slack inside the stage
    W->stage_bytes = off + 2048;
}

// panel indices
inline int panel_A(int j) { return j; }                             // j = 0..J
inline int panel_DZ(int j, int J) { return J + j; }                 // j = 1..J+1  -> J+1 .. 2J+1

inline bool make_tc_wgrad_plan(const TcDgradLayout& G, TcWgradPlan* W, BwdPanels pp = BwdPanels{1, 1, 2}) {
    *W = TcWgradPlan{};
    const int J = G.J;
    W->n_layers = J + 1;
    W->n_panels = 2 * J + 2;
    for (int j = 0; j <= J; ++j) { W->panel_width[panel_A(j)] = G.P[j]; W->panel_parts[panel_A(j)] = pp.a_parts; }
    for (int j = 1; j <= J; ++j) {
        W->panel_width[panel_DZ(j, J)] = G.P[j];
        W->panel_parts[panel_DZ(j, J)] = (j == J) ? pp.dz_head_parts : pp.dz_parts;
    }
    W->panel_width[panel_DZ(J + 1, J)] = 16;
    W->panel_parts[panel_DZ(J + 1, J)] = pp.dz_head_parts;
    int col = 0;
    for (int j = 1; j <= J + 1; ++j) {
        TcWgradLayer& y = W->layer[j - 1];
        y.lin = j - 1;
        y.n_in = G.H[j - 1];
        y.n_out = (j <= J) ? G.H[j] : 1;
        y.swapped = (j == J + 1);
        if (!y.swapped) {
            y.m_panel = panel_DZ(j, J);  y.m_width = G.P[j];
            y.n_panel = panel_A(j - 1);  y.n_width = G.P[j - 1];
            y.ones_col = G.H[j - 1];            // column of A_{j-1} that holds 1.0 (for A_0: index 1+E)
        } else {
            y.m_panel = panel_A(J);      y.m_width = G.P[J];
            y.n_panel = panel_DZ(J + 1, J); y.n_width = 16;
            y.ones_col = G.H[J];                // ROW of D (unit H_J of A_J holds 1.0)
        }
        if (y.m_width > 256 || (y.n_width % 16) != 0 || (y.m_width % 16) != 0) return false;
        y.tmem_col = col;
        col += y.n_width;
    }
    W->tmem_cols_used = col;
    if (col > 512) return false;
    for (int p = 0; p < W->n_panels; ++p)
        W->block_bytes[p] = 16u * (uint32_t)W->panel_width[p] * (uint32_t)W->panel_parts[p];   // `parts` tiles of half the columns
    tc_wgrad_set_stage(W, 1);
    return true;
}

}  // namespace umnn
