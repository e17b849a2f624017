// Bracket refinement of the sampling direction (UMNNMAF.invert, models/UMNN/UMNNMAF.py:182-232).
//
// For one dimension j the reference repeats `iter` times: lay 10 grid points over [left, right] per sample
// (:210), integrate the monotone integrand from 0 to every grid point (:213, the fused forward kernel with
// the contiguous-context layout), pick the grid point whose image is closest to the target (:218), and
// shrink the bracket to its neighbours (:224-230).  Everything after the integral is ~15 small torch ops;
// here it is one launch, one thread per sample, that also lays the next round's grid, so a round costs two
// launches.  HBM-bound and tiny (44 floats per sample): no staging, coalesced [G][B] accesses.
//
// Arithmetic is kept identical to the torch expressions (separate multiply and add, no FMA contraction),
// including the reference's flat neighbour indexing: x_flat = x[:, :, j].t() has index b*G + g, the left
// neighbour is x_flat[mid - 1] (index -1 wraps to the last element, grid point 0 reads the previous
// SAMPLE's last grid point) and the right neighbour x_flat[(mid + 1) % (B*G)].
#include "umnn_common.cuh"

namespace umnn {

__global__ void __launch_bounds__(128)
invert_bracket_kernel(long long B, int G, const float* __restrict__ integ, const float* __restrict__ x_cur,
                      const float* __restrict__ grid, const float* __restrict__ offset, long long offset_stride,
                      const float* __restrict__ scale, const float* __restrict__ target, long long target_stride,
                      float* __restrict__ left, float* __restrict__ right, long long bracket_stride,
                      float* __restrict__ x_next, float* __restrict__ x_mid, long long x_mid_stride) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float l, r;
    if (integ == nullptr) {
        // first round: only lay the grid over the initial bracket
        l = left[b * bracket_stride];
        r = right[b * bracket_stride];
    } else {
        const float s = *scale;
        const float off = offset[b * offset_stride];
        const float tgt = target[b * target_stride];
        int pos = 0;
        float z_val = __fmul_rn(s, __fadd_rn(off, integ[b]));
        float best = fabsf(__fsub_rn(z_val, tgt));
        for (int g = 1; g < G; ++g) {
            const float z = __fmul_rn(s, __fadd_rn(off, integ[(long long)g * B + b]));
            const float d = fabsf(__fsub_rn(z, tgt));
            // torch.min(dim): first minimum, NaN wins
            if (d < best || (d != d && best == best)) { best = d; pos = g; z_val = z; }
        }
        const long long n = B * G;
        const long long mid = b * G + pos;
        const long long lo = mid > 0 ? mid - 1 : n - 1;
        const long long hi = (mid + 1) % n;
        const float xm = x_cur[(mid % G) * B + mid / G];
        const float xl = x_cur[(lo % G) * B + lo / G];
        const float xr = x_cur[(hi % G) * B + hi / G];
        const float below = z_val < tgt ? 1.0f : 0.0f;
        const float above = __fsub_rn(1.0f, below);
        l = __fadd_rn(__fmul_rn(below, xm), __fmul_rn(above, xl));
        r = __fadd_rn(__fmul_rn(below, xr), __fmul_rn(above, xm));
        left[b * bracket_stride] = l;
        right[b * bracket_stride] = r;
        if (x_mid) x_mid[b * x_mid_stride] = xm;
    }
    const float width = __fsub_rn(r, l);
    for (int g = 0; g < G; ++g) x_next[(long long)g * B + b] = __fadd_rn(__fmul_rn(grid[g], width), l);
}

}  // namespace umnn

extern "C" {

int umnn_invert_bracket_step(int64_t n_samples, int32_t n_grid, const float* integral, const float* x_grid,
                             const float* grid, const float* offset, int64_t offset_stride, const float* scale,
                             const float* target, int64_t target_stride, float* left, float* right,
                             int64_t bracket_stride, float* x_grid_next, float* x_mid, int64_t x_mid_stride,
                             void* stream) {
    using namespace umnn;
    if (n_samples < 0 || n_grid < 2 || n_grid > 1024) {
        set_error("umnn_invert_bracket_step: n_samples=%lld n_grid=%d out of range", (long long)n_samples, n_grid);
        return UMNN_ERR_DESC;
    }
    if (n_samples == 0) return 0;
    if (!grid || !left || !right || !x_grid_next) {
        set_error("umnn_invert_bracket_step: required pointer is NULL");
        return UMNN_ERR_NULL;
    }
    if (integral && (!x_grid || !offset || !scale || !target)) {
        set_error("umnn_invert_bracket_step: a refinement round needs x_grid, offset, scale and target");
        return UMNN_ERR_NULL;
    }
    if (integral && x_grid == x_grid_next) {
        set_error("umnn_invert_bracket_step: x_grid_next must not alias x_grid (neighbouring samples are read)");
        return UMNN_ERR_DESC;
    }
    const int threads = 128;
    const long long blocks = (n_samples + threads - 1) / threads;
    invert_bracket_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        n_samples, n_grid, integral, x_grid, grid, offset, offset_stride, scale, target, target_stride, left, right,
        bracket_stride, x_grid_next, x_mid, x_mid_stride);
    UMNN_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // extern "C"
