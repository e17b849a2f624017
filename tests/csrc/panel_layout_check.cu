// Host-only consistency check of the backward's operand-panel addressing (tc_bwd_layout.cuh): the row-resolved
// view the epilogue threads use (panel_row + panel_granule [+ lo_off]) must address exactly the bytes of
// panel_offset(), for hi-only (parts = 1) and hi + lo (parts = 2) panels; a CTA's half-panel must be contiguous over
// consecutive 16-row blocks (pass W fetches several blocks of it with one bulk copy), and inside a fetched block the
// bytes must be the MN-major UMMA tile pass W's descriptors describe.
#include <cstdio>
#include <set>
#include "../../umnn_b200/csrc/tc_bwd_layout.cuh"
using namespace umnn;
int main() {
    int bad = 0;
    const int widths[] = {16, 32, 48, 64, 112, 208, 256};
    const long long r_pad = 64;
    for (int parts = 1; parts <= 2; ++parts)
        for (int W : widths) {
            std::set<size_t> seen;
            const size_t half_bytes = panel_half_bytes(r_pad, W, parts);
            const size_t block_bytes = (size_t)16 * W * parts;      // one 16-row block of a half-panel (TcWgradPlan::block_bytes)
            if (half_bytes != (size_t)(r_pad / 16) * block_bytes) { printf("half size parts=%d W=%d\n", parts, W); ++bad; }
            for (long long pr = 0; pr < r_pad; ++pr) {
                PanelRow R = panel_row((uint8_t*)nullptr, pr, W, parts, r_pad);
                for (int c = 0; c < W; c += 8)
                    for (int part = 0; part < parts; ++part) {
                        const size_t a = (size_t)(R.base - (uint8_t*)nullptr) + panel_granule(R, c) + (part ? R.lo_off : 0);
                        const size_t b = panel_offset(pr, c, W, part, parts, r_pad);
                        if (a != b) { if (bad < 10) printf("mismatch parts=%d W=%d pr=%lld c=%d part=%d: %zu vs %zu\n", parts, W, pr, c, part, a, b); ++bad; }
                        for (int k = 0; k < 8; ++k) {
                            const size_t e = panel_offset(pr, c + k, W, part, parts, r_pad);
                            if (e != b + 2 * k) ++bad;
                            if (!seen.insert(e).second) { if (bad < 10) printf("alias parts=%d W=%d\n", parts, W); ++bad; }
                            // the element lives in the half-panel of the CTA that owns its column, in block pr / 16
                            const int rank = (c + k) >= W / 2;
                            const size_t lo = (size_t)rank * half_bytes + (size_t)(pr >> 4) * block_bytes;
                            if (e < lo || e >= lo + block_bytes) { if (bad < 10) printf("block range parts=%d W=%d c=%d\n", parts, W, c + k); ++bad; }
                            // position inside the block: [part][k8][col8][row%8][col%8] with LBO = (W/16)*128, part stride 16*W
                            const int cin = (c + k) - rank * (W / 2);
                            const size_t want = (size_t)part * 16 * W + (size_t)((pr >> 3) & 1) * (W / 16) * 128 + (size_t)(cin / 8) * 128 + (pr & 7) * 16 + (cin & 7) * 2;
                            if (e - lo != want) { if (bad < 10) printf("tile pos parts=%d W=%d\n", parts, W); ++bad; }
                        }
                    }
            }
            if (seen.size() != (size_t)r_pad * W * parts) { printf("coverage parts=%d W=%d: %zu\n", parts, W, seen.size()); ++bad; }
            if (*seen.rbegin() + 2 != (size_t)r_pad * W * parts * 2) { printf("extent parts=%d W=%d\n", parts, W); ++bad; }
        }
    printf(bad ? "FAILED %d\n" : "panel layout ok\n", bad);
    return bad ? 1 : 0;
}
