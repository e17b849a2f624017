// sm_100a building blocks for the tensor-core path: mbarrier, cluster, TMEM allocation, tcgen05
// MMA / commit / ld / st wrappers and the UMMA shared-memory / instruction descriptors.
// Raw PTX only (no CUTLASS dependency); field layouts follow the PTX ISA "tcgen05" chapter.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace umnn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------
// cluster
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive_release(); cluster_wait_acquire(); }
// shared::cta address of `p` as seen in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster.  Default (.release.cta) semantics:
// the payload travels through tensor memory and is ordered by tcgen05.wait::st + tcgen05.fence, so no
// cluster-scope memory fence (a full MEMBAR) is paid per arrive.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    const uint32_t remote = map_to_cta(smem_u32(bar), rank);
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint (the value CUTLASS passes): ptxas turns it into TRYWAIT + NANOSLEEP.SYNCS +
// re-probe, i.e. the warp sleeps on the barrier between probes instead of spinning.
constexpr uint32_t kMbarSuspendHintNs = 0x989680u;   // 10 ms, the value CUTLASS passes
#ifndef UMNN_TC_WAIT_HINT
#define UMNN_TC_WAIT_HINT 1
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
#if UMNN_TC_WAIT_HINT
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendHintNs)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
#endif
    return ok != 0;
}
__device__ __forceinline__ void prefetch_l2(const void* gptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(gptr)); }
// non-blocking probe (test_wait never suspends): the MMA issuer uses it to decide what to issue next
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// acquire at cluster scope: needed when the arrivals come from the peer CTA
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendHintNs)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded waiting: a protocol bug (or a faulted peer CTA) traps with a message after UMNN_TC_SPIN_LIMIT nanoseconds
// (default ~4.3 s; the clock is read once per 4096 polls) instead of hanging the GPU.  A legitimate wait lasts at
// most a few tiles (microseconds).  0 = wait forever.
#ifndef UMNN_TC_SPIN_LIMIT
#define UMNN_TC_SPIN_LIMIT (1LL << 32)
#endif
// No printf here: a device-side call anywhere in the kernel makes ptxas give up uniform registers for values that are
// live across it (they are caller-saved through vector registers), and the MMA issuer's descriptors are live across
// every wait.  The trap surfaces as cudaErrorLaunchFailure; `tag` / parity stay in registers for cuda-gdb.
__device__ __forceinline__ void mbar_wait_expired(int tag, uint32_t parity) {
    (void)tag; (void)parity;
    asm volatile("trap;");
}
// UMNN_TC_WAIT_STYLE: 1 (default) = read the clock after every failed poll; 0 = clock every 4096 polls; 2 = poll
// count only (2^27 polls), no clock.  A/B on one B200 (scripts/gpu_visit_r1k.sh, ms per step, config 4 at
// B = 8192 / config 3 / config 5): style 1 24.94 / 1.413 / 1.281, style 0 25.05 / 1.426 / 1.290, style 2 with or
// without the hint 25.67 / 1.487 / 1.358 -- the slower the idle warps poll, the more issue slots and LSU bandwidth
// the working warps get (the hint lets ptxas put a NANOSLEEP.SYNCS between two probes; the clock read adds to it).
#ifndef UMNN_TC_WAIT_STYLE
#define UMNN_TC_WAIT_STYLE 1
#endif
#ifndef UMNN_TC_WAIT_SLEEP_NS
#define UMNN_TC_WAIT_SLEEP_NS 50
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
#if !UMNN_TC_SPIN_LIMIT
    (void)tag;
    while (!mbar_try_wait(bar, parity)) {}
#elif UMNN_TC_WAIT_STYLE == 1
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = global_timer_ns();
    while (!mbar_try_wait(bar, parity))
        if (global_timer_ns() - t0 > (unsigned long long)(UMNN_TC_SPIN_LIMIT)) mbar_wait_expired(tag, parity);
#elif UMNN_TC_WAIT_STYLE == 2
    for (int i = 0; i < (1 << 27); ++i)
        if (mbar_try_wait(bar, parity)) return;
    mbar_wait_expired(tag, parity);
#elif UMNN_TC_WAIT_STYLE == 3
    // explicit back-off of UMNN_TC_WAIT_SLEEP_NS between probes, clock every 1024 probes
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0 = 0;
    for (uint32_t n = 1; !mbar_try_wait(bar, parity); ++n) {
        __nanosleep(UMNN_TC_WAIT_SLEEP_NS);
        if ((n & 0x3FFu) == 0) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > (unsigned long long)(UMNN_TC_SPIN_LIMIT)) mbar_wait_expired(tag, parity);
        }
    }
#else
    unsigned long long t0 = 0;
    for (uint32_t n = 1; !mbar_try_wait(bar, parity); ++n) {
        if ((n & 0xFFFu) == 0) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > (unsigned long long)(UMNN_TC_SPIN_LIMIT)) mbar_wait_expired(tag, parity);
        }
    }
#endif
}
// The same wait for a warp that must stay CONVERGED (the MMA issuer): every lane probes and the loop exit is a warp
// vote, i.e. a warp-uniform condition.  With a per-lane exit condition ptxas treats everything after the loop as
// possibly divergent and keeps descriptors / tensor-memory addresses in vector registers, paying an R2UR per operand
// of every tcgen05.mma (18 per K block, ~50 instructions between two MMA triples -- more than the MMAs take on narrow
// layers); with the vote they live in uniform registers.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int tag = 0) {
    if (__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) return;
#if UMNN_TC_SPIN_LIMIT
    const unsigned long long t0 = global_timer_ns();
    while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity)))
        if (global_timer_ns() - t0 > (unsigned long long)(UMNN_TC_SPIN_LIMIT)) mbar_wait_expired(tag, parity);
#else
    (void)tag;
    while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {}
#endif
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int tag = 0) {
#if UMNN_TC_SPIN_LIMIT
    unsigned long long t0 = 0;
    for (uint32_t n = 1; !mbar_try_wait_cluster(bar, parity); ++n) {
        if ((n & 0xFFFu) == 0) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > (unsigned long long)(UMNN_TC_SPIN_LIMIT)) mbar_wait_expired(tag, parity);
        }
    }
#else
    (void)tag;
    while (!mbar_try_wait_cluster(bar, parity)) {}
#endif
}

// true in exactly one (elected) lane of a fully converged warp
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// generic-proxy writes to shared memory -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// tensor memory
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one full warp; writes the TMEM base address (lane 0, first column) to *holder (shared memory)
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* holder, uint32_t ncols) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// TMEM address = (lane << 16) | column
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) { return base + (lane << 16) + col; }

// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp receives lane (lane_base + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// descriptors
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"): the operand is a grid of
// 8-row x 16-byte core matrices, each stored as 128 contiguous bytes (row r of the core matrix at
// byte 16*r).  lbo = byte distance between the two core matrices that are adjacent in K inside one
// K=16 (bf16) instruction; sbo = byte distance between core matrices adjacent in the M/N direction.
//   bits [0,14)  start address >> 4        bits [16,30) lbo >> 4        bits [32,46) sbo >> 4
//   bits [46,48) = 1 (sm_100 descriptor version)       bits [61,64) = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// Instruction descriptor for kind::f16 with BF16 inputs, FP32 accumulation, both operands K-major.
//   [4,6) D format: 1 = F32     [7,10) A format: 1 = BF16     [10,13) B format: 1 = BF16
//   [15] A major (0 = K)        [16] B major (0 = K)          [17,23) N >> 3        [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Same, with the operand formats as arguments (kind::f16 encodes A and B separately: 0 = F16, 1 = BF16).
#define UMNN_OPF_FP16 0
#define UMNN_OPF_BF16 1
__host__ __device__ constexpr uint32_t make_idesc_f32(int M, int N, int a_fmt, int b_fmt) {
    return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// MMA issue (ONE thread) and completion
// ---------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T
template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T     (A: lane = row, 2 bf16 per 32-bit column, even k in the low half)
template <int CG>
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// The same TS MMA issued from CONVERGED code: every lane executes the statement, the election happens inside it and
// only the elected lane's tcgen05.mma is issued.  Unlike `if (elect_one_sync()) mma_ts(...)`, the operands are used by
// convergent code, so warp-uniform descriptors / addresses can stay in uniform registers instead of being broadcast
// from the elected lane's vector registers before every MMA.
template <int CG>
__device__ __forceinline__ void mma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
}
template <int CG>
__device__ __forceinline__ void mma_ss_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// mbarrier arrive (count 1) once every previously issued tcgen05.mma of this thread has completed.
// CG == 2: the arrive is multicast to the barrier at the same offset in every CTA of `cta_mask`.
template <int CG>
__device__ __forceinline__ void mma_commit(uint64_t* bar, uint16_t cta_mask = 0x3) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(cta_mask)
                     : "memory");
}

// ---------------------------------------------------------------------------------------------
// One MMA layer of the activation chain (forward, pass F, pass D), issued by the converged issuer warp.
//
// The layer's N range is cut into one or two segments with separate accumulators; every K block is the hi/lo triple
// D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo with A in tensor memory.  K block kb needs 32-column pair kb/2 of the A
// operand, published by the epilogue warps on ready[kb/2].  Segment 0 is issued first (its commit starts the epilogue
// that overlaps segment 1).  -DUMNN_TC_ISSUE_WORKCONSERVING=1 lets segment 1 advance on the pairs that are already
// there whenever segment 0 is waiting for one that is not; measured on one box against the strict order it changes
// nothing (forward 18.85 vs 18.91 ms at config 4 / 8192, pass F 258 vs 255 us, pass D 321 vs 318 us,
// profiles/r2_issue_order_ab.txt): the chain MMA -> epilogue -> MMA of a tile is latency bound, the issuer's waits are
// a symptom.  All decisions are warp votes or warp-uniform integers, so descriptors and addresses stay in uniform
// registers (see mbar_wait_warp).
// ---------------------------------------------------------------------------------------------
struct MmaSegment {
    uint32_t idesc;      // instruction descriptor (M = 256, N = segment width)
    uint32_t d_addr;     // accumulator (tensor memory)
    uint64_t bhi, blo;   // shared-memory descriptors of K block 0 (hi / lo part of B)
    uint64_t step;       // descriptor increment per K block
};

__device__ __forceinline__ void mma_triple(const MmaSegment& g, uint32_t a_hi, uint64_t bhi, uint64_t blo, uint32_t accumulate) {
    mma_ts_elect<2>(g.d_addr, a_hi, bhi, g.idesc, accumulate);
    mma_ts_elect<2>(g.d_addr, a_hi + 8, bhi, g.idesc, 1);
    mma_ts_elect<2>(g.d_addr, a_hi, blo, g.idesc, 1);
}

__device__ __forceinline__ void issue_mma_layer(int n_kb, int nseg, const MmaSegment& g0, const MmaSegment& g1, uint32_t a_base,
                                                uint64_t* ready, uint32_t par, uint64_t* acc_bar, int tag) {
    int kb0 = 0, kb1 = 0, known = 0;          // next K block per segment; pairs [0, known) are confirmed published
    uint32_t a0 = a_base, a1 = a_base;
    uint64_t bhi0 = g0.bhi, blo0 = g0.blo, bhi1 = g1.bhi, blo1 = g1.blo;
    while (kb0 < n_kb) {
        const int need = kb0 >> 1;
        if (need >= known) {
            if (__all_sync(0xffffffffu, mbar_test(&ready[need], par))) {
                known = need + 1;
                tc_fence_after_sync();
#if defined(UMNN_TC_ISSUE_WORKCONSERVING) && UMNN_TC_ISSUE_WORKCONSERVING      // A/B switch, off in the product build (see above)
            } else if (nseg == 2 && kb1 < n_kb && (kb1 >> 1) < known) {
                mma_triple(g1, a1, bhi1, blo1, kb1 > 0);
                a1 += 16; bhi1 += g1.step; blo1 += g1.step; ++kb1;
                continue;
#endif
            } else {
                mbar_wait_warp(&ready[need], par, tag + need);
                known = need + 1;
                tc_fence_after_sync();
            }
        }
        mma_triple(g0, a0, bhi0, blo0, kb0 > 0);
        a0 += 16; bhi0 += g0.step; blo0 += g0.step; ++kb0;
    }
    if (elect_one_sync()) mma_commit<2>(&acc_bar[0], 0x3);
    __syncwarp();
    if (nseg == 2) {
        while (kb1 < n_kb) {
            mma_triple(g1, a1, bhi1, blo1, kb1 > 0);         // every pair is confirmed: segment 0 has walked all of them
            a1 += 16; bhi1 += g1.step; blo1 += g1.step; ++kb1;
        }
        if (elect_one_sync()) mma_commit<2>(&acc_bar[1], 0x3);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// bf16 hi/lo split of an fp32 pair:  v = hi + lo + O(2^-17 |v|),  packed as {even in low half}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float even, float odd) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(odd), "f"(even));  // first source -> upper half
    return r;
}
__device__ __forceinline__ void split_bf16x2(float even, float odd, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(even, odd);
    const float he = __uint_as_float(hi << 16);
    const float ho = __uint_as_float(hi & 0xFFFF0000u);
    lo = pack_bf16x2(even - he, odd - ho);
}

// fp16 hi/lo split:  v = hi + lo + O(2^-23 |v|) for 2^-3 <= |v| <= 65504 (absolute error <= 2^-25 below: lo is
// subnormal there); |v| > 65504 overflows to inf (hi) and -inf / NaN (lo) -- see the guard in cc_forward_tc_kernel.
__device__ __forceinline__ uint32_t pack_f16x2(float even, float odd) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(odd), "f"(even));  // first source -> upper half
    return r;
}
constexpr float kFp16Max = 65504.0f;

// ---------------------------------------------------------------------------------------------
// packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2 -- one issue slot for two IEEE fp32 operations, results identical
// to the scalar instructions).  UMNN_TC_F32X2=0 compiles the scalar forms instead (experiment switch).
// ---------------------------------------------------------------------------------------------
#ifndef UMNN_TC_F32X2
#define UMNN_TC_F32X2 1
#endif
__device__ __forceinline__ unsigned long long f2_bits(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 f2_from(unsigned long long b) { return *reinterpret_cast<float2*>(&b); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
#if UMNN_TC_F32X2
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
    return f2_from(d);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
#if UMNN_TC_F32X2
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(d);
#else
    return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
#endif
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
#if UMNN_TC_F32X2
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(d);
#else
    return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y));
#endif
}
// hi/lo split of a pair with the residual subtraction as one packed instruction
template <int OPF>
__device__ __forceinline__ void split2(float2 a, uint32_t& hi, uint32_t& lo) {
    float2 h;
    if constexpr (OPF == UMNN_OPF_BF16) {
        hi = pack_bf16x2(a.x, a.y);
        h = make_float2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xFFFF0000u));
    } else {
        hi = pack_f16x2(a.x, a.y);
        h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    }
    const float2 r = fsub2(a, h);
    if constexpr (OPF == UMNN_OPF_BF16) lo = pack_bf16x2(r.x, r.y);
    else lo = pack_f16x2(r.x, r.y);
}

// sign bits of 16 fp32 bit patterns, element 4k + j -> bit 8j + k (see mask_bitpos in tc_bwd_layout.cuh):
// three byte permutes gather the top bytes of four values, one shift + mask drops them into place
__device__ __forceinline__ uint32_t sign_mask16(const uint32_t (&v)[16]) {
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t t0 = __byte_perm(v[4 * k], v[4 * k + 1], 0x0073);
        const uint32_t t1 = __byte_perm(v[4 * k + 2], v[4 * k + 3], 0x0073);
        const uint32_t g = __byte_perm(t0, t1, 0x5410);
        m |= (g >> (7 - k)) & (0x01010101u << k);
    }
    return m;
}

}  // namespace tc
}  // namespace umnn
