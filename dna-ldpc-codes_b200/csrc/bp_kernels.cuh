// bp_kernels.cuh - hand-written sm_100a kernels of the flooding sum-product decoder.
//
// Data layout in HBM (one "wave" = G groups of 32 frames, frame-interleaved so that a warp = 32 frames of a group):
//   msg   [G][E][32] T     ONE in-place message per edge: holds pr (bit->check, p0/p1) before the check-node pass
//                          and lr (check->bit) after it. Edge order = CSR (row-major, ascending column), so the 72
//                          messages x 32 frames of a check are one contiguous 18 KB chunk.
//   lratio[G][N][32] T     channel likelihood ratios
//   decw  [G][N]  u32      hard decisions, bit f of word = frame f of the group (warp ballot)
//   actw  [G]     u32      frames of the group still iterating
// Every global access of a warp is one fully used 256 B (fp64) / 128 B (fp32) segment.
//
// Reference semantics: Iter_Belief_Propagation dec.cpp:632-694, check() check.cpp:28-47,
// Run_Belief_Propagation_Decoder dec.cpp:583-605.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bp_math.cuh"

namespace dnaldpc {

constexpr int kFG = 32;  // frames per group == warp width

template <typename T> __device__ __forceinline__ T ld_stream(const T *p) { return __ldcs(p); }
template <typename T> __device__ __forceinline__ void st_stream(T *p, T v) { __stcs(p, v); }

// ------------------------------------------------------------------------------------------------
// Check-node (row) pass, dec.cpp:644-662.  One thread = one (check i, frame f); one warp = check i of 32 frames.
//   d_k = 1 - 2/(1+pr_k);  F_0 = 1, F_{k+1} = F_k*d_k;  B_last = 1, B_{k-1} = B_k*d_k;  lr_k = (1+F_k*B_k)/(1-F_k*B_k)
// The reference evaluates d_k twice (identical values) and parks F_k in e->lr; here the d_k live in registers,
// the backward products are check-pointed every 8 edges and re-derived block by block with the same
// multiplications in the same order, so every rounded intermediate is the reference's.
// FIRST: iteration 0 of a frame reads pr_e = lratio[col(e)] (Init_Belief_Propagation, dec.cpp:608-629) instead of msg,
// which removes the E-sized initialisation write.
// ------------------------------------------------------------------------------------------------
// Out-of-line fallback for one (check, frame) whose inputs left the proven operand ranges. Same operations in the
// same order as the reference, with nvcc's full-range divisions; F_k is re-derived for every k (O(deg^2)) because the
// single in-place message array has no room to park it. Reached only with invalid (negative / NaN) likelihood ratios.
template <typename T, bool FIRST>
__device__ __noinline__ void row_slow_path(T *base, const T *lr_lane, const int32_t *cols, int deg) {
    auto pr_at = [&](int k) -> T { return FIRST ? lr_lane[(size_t)cols[k] * kFG] : base[(size_t)k * kFG]; };
    T B = T(1);
    for (int k = deg - 1; k >= 0; k--) {
        T F = T(1);
        for (int m = 0; m < k; m++) F = mul_rn(F, check_factor_slow(pr_at(m)));
        const T dk = check_factor_slow(pr_at(k));
        base[(size_t)k * kFG] = check_to_bit_slow(mul_rn(F, B));
        B = mul_rn(B, dk);
    }
}

// The arithmetic of one (check, frame): d[] holds pr_k on entry; writes lr_k to base[k*32]. One basic block.
template <typename T, int DC, bool EXACT, bool FIRST>
__device__ __forceinline__ void row_compute(T (&d)[DC], int deg, T *base, const T *lr_lane, const int32_t *cols) {
    constexpr int NB = (DC + 7) / 8;
    T ck[NB];
    T B = T(1);
    bool bad = false;
#pragma unroll
    for (int k = DC - 1; k >= 0; k--) {
        T dk = (EXACT || k < deg) ? check_factor(d[k], bad) : T(1);  // padding edges: exact identity in both chains
        d[k] = dk;
        if ((k & 7) == 7 || k == DC - 1) ck[k >> 3] = B;
        B = mul_rn(B, dk);
    }
    if (bad) {  // invalid likelihood ratios (negative / NaN): redo this check with full IEEE divisions, nothing stored yet
        row_slow_path<T, FIRST>(base, lr_lane, cols, deg);
        return;
    }
    T F = T(1);
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int bot = b * 8;
        const int top = (bot + 7 < DC - 1) ? bot + 7 : DC - 1;
        T Bv[8];
        Bv[top - bot] = ck[b];
#pragma unroll
        for (int k = top; k > bot; k--) Bv[k - 1 - bot] = mul_rn(Bv[k - bot], d[k]);
#pragma unroll
        for (int k = bot; k <= top; k++) {
            T t = mul_rn(F, Bv[k - bot]);
            T lr = check_to_bit(t);
            if (EXACT || k < deg) st_stream(base + (size_t)k * kFG, lr);
            F = mul_rn(F, d[k]);
        }
    }
}

// Variant A: one warp per (check, group), loads straight into registers. Used for small or sparsely active waves
// (finished lanes load nothing, so a group with one straggler moves one 32 B sector per edge, not 256 B).
template <typename T, int DC, bool EXACT, bool FIRST, int kRowWarps>
__global__ void __launch_bounds__(kRowWarps * 32)
row_pass_kernel(T *__restrict__ msg, const T *__restrict__ lratio, const uint32_t *__restrict__ actw,
                const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx, int M, int N, int E, int g0, int G) {
    const int lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (item >= (long long)G * M) return;
    const int gl = (int)(item / M), i = (int)(item - (long long)gl * M);
    const int g = g0 + gl;
    if (!((actw[g] >> lane) & 1u)) return;  // finished (or padding) frames keep their messages untouched

    const int e0 = EXACT ? i * DC : row_ptr[i];
    const int deg = EXACT ? DC : (row_ptr[i + 1] - e0);
    T *base = msg + ((size_t)g * E + e0) * kFG + lane;

    T d[DC];
#pragma unroll
    for (int k = 0; k < DC; k++) {
        if (EXACT || k < deg) {
            if (FIRST) d[k] = __ldg(lratio + ((size_t)g * N + __ldg(col_idx + e0 + k)) * kFG + lane);
            else d[k] = ld_stream(base + (size_t)k * kFG);
        }
    }
    row_compute<T, DC, EXACT, FIRST>(d, deg, base, lratio + (size_t)g * N * kFG + lane, col_idx + e0);
}

// ---- TMA (cp.async.bulk) + mbarrier helpers --------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy (SASS: UBLKCP), completion signalled on the mbarrier as transferred bytes
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Variant B (bulk path; opt-in with DNALDPC_ROW_IMPL=tma): persistent warps, each owning one shared-memory tile. While a warp computes check t
// from registers, the TMA engine streams the 18 KB of its next check (DC x 32 frames, contiguous in `msg`) into the
// tile with one cp.async.bulk; the loads never occupy registers or LSU slots and their latency hides behind ~3 000
// cycles of fp64 work. FIRST gathers the DC 256-byte lratio segments with one bulk copy per edge instead.
// Measured on B200 (round 1, 4096-frame wave): 1.95 ms per launch vs 1.75-1.79 ms for variant A, because the check-node
// pass is limited by how many bytes an SM can keep in flight (registers + smem are exhausted by the 72 live factors),
// not by exposed load latency; a plain copy with this access pattern needs ~290 KB in flight per SM to reach 6.8 TB/s
// and gets 5.5 TB/s with the 147 KB that 8 warps can hold. Kept as the measured alternative, not the default.
constexpr int kRowTmaWarps = 4;

template <typename T, int DC, bool EXACT, bool FIRST>
__global__ void __launch_bounds__(kRowTmaWarps * 32, 2)
row_pass_tma_kernel(T *__restrict__ msg, const T *__restrict__ lratio, const uint32_t *__restrict__ actw,
                    const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx, int M, int N, int E, int g0, int G) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T *tile = reinterpret_cast<T *>(smem_raw) + (size_t)warp * DC * kFG;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kRowTmaWarps * DC * kFG * sizeof(T)) + warp;
    const long long total = (long long)G * M;
    const long long stride = (long long)gridDim.x * kRowTmaWarps;
    long long item = (long long)blockIdx.x * kRowTmaWarps + warp;
    if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncwarp();
    const uint64_t pol = l2_evict_first_policy();

    auto issue = [&](long long it) {  // whole warp calls; no-op for groups with no active frame
        const int gl = (int)(it / M), i = (int)(it - (long long)gl * M);
        const int g = g0 + gl;
        if (__ldg(actw + g) == 0) return;
        const int e0 = EXACT ? i * DC : __ldg(row_ptr + i);
        const int deg = EXACT ? DC : (__ldg(row_ptr + i + 1) - e0);
        if (deg == 0) return;
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)(deg * kFG * sizeof(T)));
        if (FIRST) {
            __syncwarp();
            for (int k = lane; k < deg; k += 32)
                tma_load_1d(tile + (size_t)k * kFG, lratio + ((size_t)g * N + __ldg(col_idx + e0 + k)) * kFG,
                            (uint32_t)(kFG * sizeof(T)), bar, pol);
        } else if (lane == 0) {
            tma_load_1d(tile, msg + ((size_t)g * E + e0) * kFG, (uint32_t)(deg * kFG * sizeof(T)), bar, pol);
        }
    };

    if (item < total) issue(item);
    uint32_t parity = 0;
    for (; item < total; item += stride) {
        const int gl = (int)(item / M), i = (int)(item - (long long)gl * M);
        const int g = g0 + gl;
        const uint32_t act = __ldg(actw + g);
        const int e0 = EXACT ? i * DC : __ldg(row_ptr + i);
        const int deg = EXACT ? DC : (__ldg(row_ptr + i + 1) - e0);
        const bool on = (act >> lane) & 1u;
        T d[DC];
        if (act != 0 && deg != 0) {
            mbar_wait(bar, parity);
            parity ^= 1u;
#pragma unroll
            for (int k = 0; k < DC; k++)
                if (EXACT || k < deg) d[k] = tile[k * kFG + lane];
            __syncwarp();                       // every lane has its registers: the tile may be overwritten
            if (lane == 0) fence_proxy_async_smem();
        }
        if (item + stride < total) issue(item + stride);
        if (on && deg != 0) {
            T *base = msg + ((size_t)g * E + e0) * kFG + lane;
            row_compute<T, DC, EXACT, FIRST>(d, deg, base, lratio + (size_t)g * N * kFG + lane, col_idx + e0);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Bit-node (column) pass, dec.cpp:667-693.  One thread = one (bit j, frame f); one warp = bit j of 32 frames.
//   P_0 = lratio_j, P_{k+1} = P_k*lr_k; tot = P_last (NaN -> 1); dblk_j = (tot <= 1);
//   S_last = 1, S_{k-1} = S_k*lr_k;  pr_k = P_k*S_k (NaN -> 1)
// The hard decisions of the 32 frames are packed with one warp ballot.
// ------------------------------------------------------------------------------------------------
constexpr int kColWarps = 8;

template <typename T, int DV, bool EXACT>
__global__ void __launch_bounds__(kColWarps * 32)
col_pass_kernel(T *__restrict__ msg, const T *__restrict__ lratio, uint32_t *__restrict__ decw,
                const uint32_t *__restrict__ actw, T *__restrict__ post, const int32_t *__restrict__ col_ptr,
                const int32_t *__restrict__ col_edge, int N, int E, int g0, int cols_per_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = g0 + blockIdx.y;
    const uint32_t act = actw[g];
    if (act == 0) return;
    const bool on = (act >> lane) & 1u;
    T *gmsg = msg + (size_t)g * E * kFG + lane;
    const int jbeg = (blockIdx.x * kColWarps + warp) * cols_per_warp;
    const int jend = min(jbeg + cols_per_warp, N);
    for (int j = jbeg; j < jend; j++) {
        const int c0 = EXACT ? j * DV : __ldg(col_ptr + j);
        const int deg = EXACT ? DV : (__ldg(col_ptr + j + 1) - c0);
        int eid[DV];
#pragma unroll
        for (int k = 0; k < DV; k++) eid[k] = (EXACT || k < deg) ? __ldg(col_edge + c0 + k) : 0;
        T lr[DV];
#pragma unroll
        for (int k = 0; k < DV; k++) lr[k] = (on && (EXACT || k < deg)) ? ld_stream(gmsg + (size_t)eid[k] * kFG) : T(1);
        T P = on ? __ldg(lratio + ((size_t)g * N + j) * kFG + lane) : T(1);
        T p[DV];
#pragma unroll
        for (int k = 0; k < DV; k++) {
            p[k] = P;
            if (EXACT || k < deg) P = mul_rn(P, lr[k]);
        }
        if (P != P) P = T(1);
        const bool bit = (P <= T(1));
        if (post != nullptr && on) post[((size_t)g * N + j) * kFG + lane] = P;
        T S = T(1);
#pragma unroll
        for (int k = DV - 1; k >= 0; k--) {
            if (EXACT || k < deg) {
                T v = mul_rn(p[k], S);
                if (v != v) v = T(1);
                if (on) st_stream(gmsg + (size_t)eid[k] * kFG, v);
                S = mul_rn(S, lr[k]);
            }
        }
        const uint32_t w = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) {
            uint32_t *dst = decw + (size_t)g * N + j;
            *dst = (act == 0xffffffffu) ? w : ((w & act) | (*dst & ~act));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Syndrome + per-frame loop control (check.cpp:28-47 + dec.cpp:594-599). One CTA per group of 32 frames:
// parity word of check i = XOR of the decision words of its bits (32 frames at once), OR-reduced over checks
// with warp shuffles. Frames whose syndrome is zero, or that reached max_iter, leave the active mask and get
// their iteration count n (the value Run_Belief_Propagation_Decoder returns).
// ------------------------------------------------------------------------------------------------
constexpr int kSynThreads = 256;
constexpr int kSynSplit = 8;  // CTAs per group (row chunks); the last one to arrive applies the loop control

__global__ void __launch_bounds__(kSynThreads)
syndrome_update_kernel(const uint32_t *__restrict__ decw, uint32_t *__restrict__ actw, int32_t *__restrict__ iters,
                       uint8_t *__restrict__ okflag, const int32_t *__restrict__ row_ptr,
                       const int32_t *__restrict__ col_idx, uint32_t *__restrict__ unsatw,
                       unsigned int *__restrict__ arrive, int M, int N, int g0, int n, int max_iter,
                       unsigned int *__restrict__ n_active /* counter of this iteration */) {
    const int g = g0 + blockIdx.y;
    const uint32_t act = actw[g];
    if (act == 0) return;  // uniform over the kSynSplit CTAs of the group: actw only changes in the last arriver
    const uint32_t *dw = decw + (size_t)g * N;
    const int rows_per = (M + kSynSplit - 1) / kSynSplit;
    const int i_end = min(M, (int)(blockIdx.x + 1) * rows_per);
    uint32_t acc = 0;
    for (int i = blockIdx.x * rows_per + threadIdx.x; i < i_end; i += kSynThreads) {
        uint32_t p = 0;
        const int e1 = __ldg(row_ptr + i + 1);
        int e = __ldg(row_ptr + i);
        for (; e + 4 <= e1; e += 4) {  // 4 independent gathers in flight
            const uint32_t a = __ldg(dw + __ldg(col_idx + e)), b = __ldg(dw + __ldg(col_idx + e + 1));
            const uint32_t c = __ldg(dw + __ldg(col_idx + e + 2)), d = __ldg(dw + __ldg(col_idx + e + 3));
            p ^= (a ^ b) ^ (c ^ d);
        }
        for (; e < e1; e++) p ^= __ldg(dw + __ldg(col_idx + e));
        acc |= p;
    }
    acc = __reduce_or_sync(0xffffffffu, acc);
    __shared__ uint32_t s_or[kSynThreads / 32];
    __shared__ unsigned int s_ticket;
    if ((threadIdx.x & 31) == 0) s_or[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t part = 0;
#pragma unroll
        for (int w = 0; w < kSynThreads / 32; w++) part |= s_or[w];
        if (part) atomicOr(unsatw + g, part);
        __threadfence();
        s_ticket = atomicAdd(arrive + g, 1u);
    }
    __syncthreads();
    if (s_ticket != kSynSplit - 1) return;
    if (threadIdx.x < 32) {  // last CTA of the group: every partial OR is visible
        __threadfence();
        const uint32_t unsat = *((volatile uint32_t *)(unsatw + g));
        const uint32_t done_ok = act & ~unsat;
        const uint32_t still = (n >= max_iter) ? 0u : (act & unsat);
        const uint32_t done = act & ~still;
        const int f = threadIdx.x;
        if ((done >> f) & 1u) {
            iters[(size_t)g * kFG + f] = n;
            okflag[(size_t)g * kFG + f] = (uint8_t)((done_ok >> f) & 1u);
        }
        if (f == 0) {
            actw[g] = still;
            unsatw[g] = 0;  // re-armed for the next iteration
            arrive[g] = 0;
            if (still) atomicAdd(n_active, (unsigned)__popc(still));
        }
    }
}

// Syndrome of the final decisions as 0/1 chars [F][M] (the `pchk` buffer of check(), check.cpp:28-47).
__global__ void __launch_bounds__(256)
syndrome_bytes_kernel(const uint32_t *__restrict__ decw, const int32_t *__restrict__ row_ptr,
                      const int32_t *__restrict__ col_idx, int M, int N, int nframes, uint8_t *__restrict__ out) {
    const int g = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const uint32_t *dw = decw + (size_t)g * N;
    uint32_t p = 0;
    const int e1 = __ldg(row_ptr + i + 1);
    for (int e = __ldg(row_ptr + i); e < e1; e++) p ^= __ldg(dw + __ldg(col_idx + e));
    for (int f = 0; f < kFG; f++) {
        const long long fr = (long long)g * kFG + f;
        if (fr < nframes) out[(size_t)fr * M + i] = (uint8_t)((p >> f) & 1u);
    }
}

// ------------------------------------------------------------------------------------------------
// Likelihood setup ("channel options"): frame-major input of any kind -> lratio[G][N][32] + initial hard
// decisions (lratio < 1, Init_Belief_Propagation dec.cpp:626) + loop state. A 32x32 tile is transposed through
// shared memory so that both the frame-major reads and the frame-interleaved writes are coalesced.
// ------------------------------------------------------------------------------------------------
enum InKind { IN_LR_F64 = 0, IN_LLR_F64 = 1, IN_BSC_BITS = 2, IN_AWGN_F32 = 3, IN_AWGN_F64 = 4, IN_VOTE_I8 = 5 };

struct SetupArgs {
    const void *data;
    size_t frame_stride;  // bytes
    double param;         // AWGN: 2/sigma^2 is NOT precomputed: LLR = 2*y/(sigma*sigma) like channel.cpp:32
    const double *table;  // BSC: 2 entries, VOTE: 256 entries (device)
    int nframes;          // valid frames of this wave
};

template <int KIND> __device__ __forceinline__ double load_lr(const SetupArgs &a, long long fr, int j) {
    const char *row = (const char *)a.data + (size_t)fr * a.frame_stride;
    if (KIND == IN_LR_F64) return ((const double *)row)[j];
    if (KIND == IN_LLR_F64) return exp(((const double *)row)[j]);
    if (KIND == IN_BSC_BITS) return a.table[(((const uint32_t *)row)[j >> 5] >> (j & 31)) & 1u];
    if (KIND == IN_AWGN_F32) return exp(2.0 * (double)((const float *)row)[j] / (a.param * a.param));
    if (KIND == IN_AWGN_F64) return exp(2.0 * ((const double *)row)[j] / (a.param * a.param));
    return a.table[(int)((const int8_t *)row)[j] + 128];
}

template <typename T, int KIND>
__global__ void __launch_bounds__(256)
setup_kernel(SetupArgs a, T *__restrict__ lratio, uint32_t *__restrict__ decw, int N) {
    __shared__ double tile[32][33];
    const int g = blockIdx.y, j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {  // r = frame within the group, tx = bit within the tile
        const long long fr = (long long)g * kFG + r;
        const int j = j0 + tx;
        double v = 1.0;                 // padding frames: erased, never active
        if (fr < a.nframes && j < N) v = load_lr<KIND>(a, fr, j);
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {  // r = bit within the tile, tx = frame
        const int j = j0 + r;
        const T v = (T)tile[tx][r];
        const uint32_t w = __ballot_sync(0xffffffffu, v < T(1));
        if (j < N) {
            lratio[((size_t)g * N + j) * kFG + tx] = v;
            if (tx == 0) decw[(size_t)g * N + j] = w;
        }
    }
}

__global__ void init_state_kernel(uint32_t *actw, uint32_t *unsatw, unsigned int *arrive, int32_t *iters, uint8_t *okflag,
                                  int G, int nframes) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < G * kFG) { iters[t] = 0; okflag[t] = 0; }
    if (t < G) {
        unsatw[t] = 0;
        arrive[t] = 0;
        const long long lo = (long long)t * kFG;
        const long long cnt = (long long)nframes - lo;
        actw[t] = cnt >= 32 ? 0xffffffffu : (cnt <= 0 ? 0u : ((1u << cnt) - 1u));
    }
}

// ------------------------------------------------------------------------------------------------
// Result gather: bit-transpose the ballot words back to per-frame outputs.
// ------------------------------------------------------------------------------------------------
// packed bits [F][ceil(N/32)]: a warp takes 32 decision words (bits j0..j0+31 of 32 frames) and transposes 32x32 bits.
__global__ void __launch_bounds__(256)
gather_bits_kernel(const uint32_t *__restrict__ decw, int N, int nframes, int words_per_frame, size_t out_stride_words,
                   uint32_t *__restrict__ out) {
    const int g = blockIdx.y;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= words_per_frame) return;
    const int j = w * 32 + lane;
    const uint32_t word = (j < N) ? decw[(size_t)g * N + j] : 0u;
    uint32_t mine = 0;
#pragma unroll
    for (int f = 0; f < 32; f++) {
        const uint32_t b = __ballot_sync(0xffffffffu, (word >> f) & 1u);
        if (lane == f) mine = b;
    }
    const long long fr = (long long)g * kFG + lane;
    if (fr < nframes) out[(size_t)fr * out_stride_words + w] = mine;
}

// 0/1 chars [F][N] (the reference's `char *dblk`)
__global__ void __launch_bounds__(256)
gather_bytes_kernel(const uint32_t *__restrict__ decw, int N, int nframes, uint8_t *__restrict__ out) {
    const int g = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const uint32_t word = decw[(size_t)g * N + j];
    for (int f = 0; f < kFG; f++) {
        const long long fr = (long long)g * kFG + f;
        if (fr < nframes) out[(size_t)fr * N + j] = (uint8_t)((word >> f) & 1u);
    }
}

// posterior [F][N] double from post[G][N][32] (or from lratio for frames that never iterated)
template <typename T>
__global__ void __launch_bounds__(256)
gather_posterior_kernel(const T *__restrict__ post, const T *__restrict__ lratio, const int32_t *__restrict__ iters,
                        int N, int nframes, double *__restrict__ out) {
    __shared__ double tile[32][33];
    const int g = blockIdx.y, j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const bool iterated = iters[(size_t)g * kFG + tx] > 0;
    for (int r = ty; r < 32; r += 8) {  // r = bit, tx = frame
        const int j = j0 + r;
        double v = 0;
        if (j < N) {
            const size_t idx = ((size_t)g * N + j) * kFG + tx;
            v = iterated ? (double)post[idx] : (double)lratio[idx];
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {  // r = frame, tx = bit
        const long long fr = (long long)g * kFG + r;
        const int j = j0 + tx;
        if (fr < nframes && j < N) out[(size_t)fr * N + j] = tile[tx][r];
    }
}

// ------------------------------------------------------------------------------------------------
// Synthetic BSC input generator (benchmarks): counter RNG keyed by (seed, GLOBAL frame, bit).
// Must stay identical to oracle/bp_oracle.c:orc_rng_u64 and tests/oraclelib.py:rng_u64 (the specification).
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
__host__ __device__ __forceinline__ uint64_t rng_u64(uint64_t seed, uint64_t frame, uint64_t bit, uint64_t stream) {
    uint64_t x = mix64(seed * 0x9E3779B97F4A7C15ULL + frame + 0x632BE59BD9B4E019ULL);
    return mix64(x ^ (bit * 0xD6E8FEB86659FD93ULL + stream * 0xA0761D6478BD642FULL + 0x2545F4914F6CDD1DULL));
}

__global__ void __launch_bounds__(256)
synth_bsc_kernel(const uint32_t *__restrict__ cw_bits, int n_cw, uint64_t seed, long long frame0, long long F, int N,
                 int words_per_frame, uint64_t thr, uint32_t *__restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= F * words_per_frame) return;
    const long long fl = t / words_per_frame;
    const int w = (int)(t - fl * words_per_frame);
    const uint64_t f = (uint64_t)(frame0 + fl);
    uint32_t word = cw_bits ? cw_bits[(size_t)(f % (uint64_t)n_cw) * words_per_frame + w] : 0u;
    const uint64_t fkey = mix64(seed * 0x9E3779B97F4A7C15ULL + f + 0x632BE59BD9B4E019ULL);
    for (int b = 0; b < 32; b++) {
        const int j = w * 32 + b;
        if (j < N) {
            const uint64_t x = mix64(fkey ^ ((uint64_t)j * 0xD6E8FEB86659FD93ULL + 0x2545F4914F6CDD1DULL));
            if ((x >> 11) < thr) word ^= (1u << b);
        }
    }
    out[t] = word;
}

// ------------------------------------------------------------------------------------------------
// Diagnostic: inlined in-range sequences vs nvcc's full-range IEEE division, bit for bit.
// ------------------------------------------------------------------------------------------------
__global__ void math_selftest_kernel(long long n, uint64_t seed, unsigned long long *mismatch) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r0 = rng_u64(seed, (uint64_t)i, 0, 0), r1 = rng_u64(seed, (uint64_t)i, 1, 0);
    unsigned long long bad_count = 0;
    // (1) pr >= 0 over the whole exponent range, plus +0, +inf and exact powers of two
    {
        double pr;
        const int sel = (int)(r1 & 15);
        if (sel == 0) pr = 0.0;
        else if (sel == 1) pr = __longlong_as_double(0x7FF0000000000000LL);
        else if (sel == 2) pr = __longlong_as_double((long long)((r0 >> 53) << 52));  // 2^k
        else if (sel < 10) pr = __longlong_as_double((long long)(r0 & 0x7FEFFFFFFFFFFFFFULL));  // any finite positive
        else {  // moderate ratios, the common regime: 2^[-64, 64)
            const uint64_t ex = 1023 - 64 + ((r1 >> 8) & 127);
            pr = __longlong_as_double((long long)((ex << 52) | (r0 & 0xFFFFFFFFFFFFFULL)));
        }
        bool bad = false;
        const double a = check_factor(pr, bad), b = check_factor_slow(pr);
        if (bad || __double_as_longlong(a) != __double_as_longlong(b)) bad_count++;
    }
    // (2) t in [-1, 1]: uniform, and clustered at +-1 (1 - 2^-k * m) where the quotient saturates
    {
        const uint64_t m = r0 & 0xFFFFFFFFFFFFFULL;
        const int sel = (int)((r1 >> 16) & 7);
        double t;
        if (sel == 0) t = 1.0;
        else if (sel == 1) t = -1.0;
        else if (sel == 2) t = 0.0;
        else if (sel < 5) {
            const uint64_t ex = 1023 - 1 - ((r1 >> 24) & 63);  // |t| in [2^-64, 1)
            t = __longlong_as_double((long long)((ex << 52) | m));
        } else {
            const uint64_t ex = 1023 - 1 - ((r1 >> 24) & 63);
            t = 1.0 - __longlong_as_double((long long)((ex << 52) | m));  // close to 1
            if (t > 1.0) t = 1.0;
        }
        if ((r1 >> 40) & 1) t = -t;
        const double a = check_to_bit(t), b = check_to_bit_slow(t);
        if (__double_as_longlong(a) != __double_as_longlong(b)) bad_count++;
    }
    if (bad_count) atomicAdd(mismatch, bad_count);
}

}  // namespace dnaldpc
