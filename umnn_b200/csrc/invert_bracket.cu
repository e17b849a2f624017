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

// First step of one dimension's refinement: replicate the dimension's context rows over the grid points
// (h_rep[g*B + b][:] = h_cols[b][:], the contiguous-context layout of the G*B integrals), reset the bracket to
// [init_left, init_right] and lay the first grid over it -- what UMNNMAF.invert does with ~8 torch ops (:203-210).
__global__ void __launch_bounds__(256)
invert_prepare_kernel(long long B, int G, int E, const float* __restrict__ h_cols, const float* __restrict__ grid,
                      float init_left, float init_right, float* __restrict__ h_rep, float* __restrict__ left,
                      float* __restrict__ right, float* __restrict__ x_first) {
    const long long n = B * (long long)E;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = h_cols[i];
        for (int g = 0; g < G; ++g) h_rep[(long long)g * n + i] = v;
    }
    const float width = __fsub_rn(init_right, init_left);
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += stride) {
        left[b] = init_left;
        right[b] = init_right;
        for (int g = 0; g < G; ++g) x_first[(long long)g * B + b] = __fadd_rn(__fmul_rn(grid[g], width), init_left);
    }
}

}  // namespace umnn

extern "C" {

size_t umnn_invert_workspace_bytes(const umnn_desc* desc, int64_t n_samples, int32_t n_grid) {
    using namespace umnn;
    if (validate_desc(desc) != 0) return 0;
    if (n_samples < 0 || n_grid < 2 || n_grid > 1024 || desc->layout != UMNN_LAYOUT_CONTIG ||
        desc->n_samples != n_samples * (int64_t)n_grid) {
        set_error("umnn_invert_workspace_bytes: desc must be UMNN_LAYOUT_CONTIG with n_samples == n_samples * n_grid");
        return 0;
    }
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t gb = (size_t)n_samples * n_grid;
    return al(umnn_workspace_bytes(desc, 0)) + al(gb * desc->n_ctx * 4) + 3 * al(gb * 4) + 2 * al((size_t)n_samples * 4) + 256;
}

int umnn_invert_dimension(const umnn_desc* desc, const void* params_packed, const float* nodes, const float* weights,
                          int64_t n_samples, int32_t n_grid, int32_t n_iter, const float* h_cols, const float* grid,
                          const float* scale, const float* target, int64_t target_stride, float init_left,
                          float init_right, float* x_out, int64_t x_out_stride, void* workspace, size_t workspace_bytes,
                          void* stream) {
    using namespace umnn;
    int rc = validate_desc(desc);
    if (rc) return rc;
    const size_t need = umnn_invert_workspace_bytes(desc, n_samples, n_grid);
    if (need == 0) return UMNN_ERR_DESC;
    if (n_iter < 1) { set_error("umnn_invert_dimension: n_iter must be >= 1"); return UMNN_ERR_DESC; }
    if (n_samples == 0) return 0;
    if (!params_packed || !nodes || !weights || !grid || !scale || !target || !x_out || (desc->n_ctx > 0 && !h_cols)) {
        set_error("umnn_invert_dimension: required pointer is NULL");
        return UMNN_ERR_NULL;
    }
    if (!workspace || workspace_bytes < need) {
        set_error("umnn_invert_dimension: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
        return UMNN_ERR_WORKSPACE;
    }
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t gb = (size_t)n_samples * n_grid;
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    const size_t fwd_ws = umnn_workspace_bytes(desc, 0);
    void* flag_ws = fwd_ws ? ws : nullptr;
    ws += al(fwd_ws);
    float* h_rep = reinterpret_cast<float*>(ws);  ws += al(gb * desc->n_ctx * 4);
    float* x_a = reinterpret_cast<float*>(ws);    ws += al(gb * 4);
    float* x_b = reinterpret_cast<float*>(ws);    ws += al(gb * 4);
    float* integ = reinterpret_cast<float*>(ws);  ws += al(gb * 4);
    float* left = reinterpret_cast<float*>(ws);   ws += al((size_t)n_samples * 4);
    float* right = reinterpret_cast<float*>(ws);
    cudaStream_t s = (cudaStream_t)stream;
    const long long work = n_samples * (long long)(desc->n_ctx > 0 ? desc->n_ctx : 1);
    long long blocks = (work + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    invert_prepare_kernel<<<(unsigned)blocks, 256, 0, s>>>(n_samples, n_grid, desc->n_ctx, h_cols, grid, init_left, init_right, h_rep,
                                                           left, right, x_a);
    UMNN_CUDA_TRY(cudaGetLastError());
    // offset = first context value of the dimension (z0 of UMNNMAF.py:203): column 0 of h_cols, stride E
    const float* offset = h_cols;
    const int64_t offset_stride = desc->n_ctx;
    for (int it = 0; it < n_iter; ++it) {
        rc = umnn_cc_forward(desc, nullptr, x_a, h_rep, params_packed, nodes, weights, integ, nullptr, nullptr, flag_ws, fwd_ws, stream);
        if (rc) return rc;
        rc = umnn_invert_bracket_step(n_samples, n_grid, integ, x_a, grid, offset, offset_stride, scale, target, target_stride, left,
                                      right, 1, x_b, x_out, x_out_stride, stream);
        if (rc) return rc;
        float* t = x_a; x_a = x_b; x_b = t;
    }
    return 0;
}

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
