// sw_kernels.cuh - sliding-window belief propagation for spatially-coupled LDPC codes.
//
// Reference: Run_SW_Decoder (dec.cpp:2092-2196) with Init_SW_Decoder (2366-2386), Iter_SW_Decoder (2388-2432),
// Check_Update_SW (2478-2500), Variable_Update_SW (2502-2548), Decision_SW (2630-2645), check_bound (check.cpp:49-72)
// -> mod2sparse_mulvec_bound (mod2sparse.cpp:883-912).
//
// Same slot-interleaved layout as the flooding decoder (bp_kernels.cuh): a warp is one node x the 32 frames of a group.
// Differences that follow from the algorithm:
//   * a bit that has left the window keeps its bit->check message while the checks it still touches keep running, so
//     e->pr and e->lr are both live: TWO message arrays pr[G][E][32], lr[G][E][32], both zero at the start
//     (alloc_entry, mod2sparse.cpp:61-62);
//   * a frame that satisfies the bounded syndrome of position t early waits (its messages untouched) for the other
//     frames it shares warps with. Frames are independent, so each frame's arithmetic - and hence its result - is
//     exactly the reference's. Two schedules: the sw_* kernels (first half of the file) walk a whole wave through
//     the positions together (Engine::sw_lockstep, A/B switch DNALDPC_SW_LOCKSTEP=1); the sw2_* kernels (second half,
//     the default, Engine::sw_groups) give every group of 32 frames its own position, refill a group as soon as it is
//     through, pack the lanes of groups in which few frames still iterate and move 1-4 stragglers into side arrays;
//   * checks of degree <= 8 (the (3,6) protograph codes) run through sw_row_reg_kernel: the row's messages in registers,
//     all loads in flight at once, the in-range division sequences of bp_math.cuh (bit-identical to IEEE division on
//     their ranges; anything else falls back, per lane, to the full-range loop); higher degrees use the generic
//     sw_row_kernel with nvcc's full-range division.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bp_kernels.cuh"

namespace dnaldpc {

// Frame-major host layout -> slot-interleaved lratio; every decision starts "set": the reference hands Run_SW_Decoder
// a dblk buffer filled with 2 (DNA_main.cpp:664-666) and mulvec_bound tests u[j] != 0.
__global__ void __launch_bounds__(256)
sw_load_kernel(const double *__restrict__ in, double *__restrict__ lratio, uint32_t *__restrict__ decw, int N, int nf) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, g = blockIdx.y, j0 = blockIdx.x * 32;
    for (int r = ty; r < 32; r += 8) {  // r = slot, tx = bit
        const int slot = g * kFG + r, j = j0 + tx;
        tile[r][tx] = (slot < nf && j < N) ? in[(size_t)slot * N + j] : 1.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {  // r = bit, tx = slot
        const int j = j0 + r;
        if (j < N) {
            lratio[((size_t)g * N + j) * kFG + tx] = tile[tx][r];
            if (tx == 0) decw[(size_t)g * N + j] = 0xffffffffu;
        }
    }
}

// Start of a window position: every frame of the wave runs again, per-position iteration count n = 0.
__global__ void sw_begin_kernel(uint32_t *__restrict__ runw, int32_t *__restrict__ n_pos, int32_t *__restrict__ sum, int G, int nf,
                                int first) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= G * kFG) return;
    n_pos[slot] = 0;
    if (first) sum[slot] = 0;
    if ((slot & 31) == 0) {
        const int left = nf - slot;
        runw[slot >> 5] = left >= 32 ? 0xffffffffu : (left > 0 ? ((1u << left) - 1u) : 0u);
    }
}

// Init_SW_Decoder: e->pr = lratio[j], e->lr = 1 for every entry of the columns [j0, j1).
__global__ void __launch_bounds__(256)
sw_init_kernel(double *__restrict__ pr, double *__restrict__ lr, const double *__restrict__ lratio,
               const int32_t *__restrict__ col_ptr, const int32_t *__restrict__ col_edge, int N, int E, int j0, int j1) {
    const int lane = threadIdx.x & 31, g = blockIdx.y;
    const int j = j0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= j1) return;
    const double v = lratio[((size_t)g * N + j) * kFG + lane];
    for (int k = __ldg(col_ptr + j); k < __ldg(col_ptr + j + 1); k++) {
        const size_t idx = ((size_t)g * E + __ldg(col_edge + k)) * kFG + lane;
        pr[idx] = v;
        lr[idx] = 1.0;
    }
}

// Check_Update_SW for the checks [c0, c1): the whole row, whatever window its bits are in. The forward products are
// parked in e->lr exactly like the reference does.
__global__ void __launch_bounds__(128)
sw_row_kernel(const double *__restrict__ pr, double *__restrict__ lr, const uint32_t *__restrict__ runw,
              const int32_t *__restrict__ row_ptr, int E, int c0, int c1, int G) {
    const int lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int rows = c1 - c0;
    if (item >= (long long)G * rows) return;
    const int g = (int)(item / rows), i = c0 + (int)(item - (long long)g * rows);
    if (!((runw[g] >> lane) & 1u)) return;
    const int e0 = __ldg(row_ptr + i), e1 = __ldg(row_ptr + i + 1);
    const double *p = pr + (size_t)g * E * kFG + lane;
    double *l = lr + (size_t)g * E * kFG + lane;
    double dl = 1.0;
    for (int e = e0; e < e1; e++) {
        l[(size_t)e * kFG] = dl;
        dl = __dmul_rn(dl, check_factor_slow(p[(size_t)e * kFG]));
    }
    dl = 1.0;
    for (int e = e1 - 1; e >= e0; e--) {
        const double t = __dmul_rn(l[(size_t)e * kFG], dl);
        l[(size_t)e * kFG] = check_to_bit_slow(t);
        dl = __dmul_rn(dl, check_factor_slow(p[(size_t)e * kFG]));
    }
}

// The same update with the row in registers (degree <= DC): forward products F_k and backward products B_k are the
// reference's (same multiplications in the same order: F runs up the row, B down), t = F_k * B_k, lr_k = (1+t)/(1-t).
template <int DC>
__global__ void __launch_bounds__(128)
sw_row_reg_kernel(const double *__restrict__ pr, double *__restrict__ lr, const uint32_t *__restrict__ runw,
                  const int32_t *__restrict__ row_ptr, int E, int c0, int c1, int G) {
    const int lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int rows = c1 - c0;
    if (item >= (long long)G * rows) return;
    const int g = (int)(item / rows), i = c0 + (int)(item - (long long)g * rows);
    if (!((runw[g] >> lane) & 1u)) return;
    const int e0 = __ldg(row_ptr + i), deg = __ldg(row_ptr + i + 1) - e0;
    const double *p = pr + ((size_t)g * E + e0) * kFG + lane;
    double *l = lr + ((size_t)g * E + e0) * kFG + lane;
    double d[DC], Bv[DC];
#pragma unroll
    for (int k = 0; k < DC; k++) d[k] = k < deg ? ld_stream(p + (size_t)k * kFG) : 0.0;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < DC; k++) d[k] = k < deg ? check_factor(d[k], bad) : 1.0;  // padding: exact identity in both chains
    if (bad) {  // operands outside the proven ranges (negative / NaN ratios): the full-range loop of sw_row_kernel for this lane
        double dl = 1.0;
        for (int k = 0; k < deg; k++) {
            l[(size_t)k * kFG] = dl;
            dl = __dmul_rn(dl, check_factor_slow(p[(size_t)k * kFG]));
        }
        dl = 1.0;
        for (int k = deg - 1; k >= 0; k--) {
            const double t = __dmul_rn(l[(size_t)k * kFG], dl);
            l[(size_t)k * kFG] = check_to_bit_slow(t);
            dl = __dmul_rn(dl, check_factor_slow(p[(size_t)k * kFG]));
        }
        return;
    }
    double B = 1.0;
#pragma unroll
    for (int k = DC - 1; k >= 0; k--) {
        Bv[k] = B;
        B = __dmul_rn(B, d[k]);
    }
    double F = 1.0;
#pragma unroll
    for (int k = 0; k < DC; k++) {
        const double t = __dmul_rn(F, Bv[k]);
        if (k < deg) st_stream(l + (size_t)k * kFG, check_to_bit(t));
        F = __dmul_rn(F, d[k]);
    }
}

// Variable_Update_SW + Decision_SW for the bits [v0, v1), restricted to the entries whose check lies in [c0, c1).
// The decision's product lratio * lr * ... (Decision_SW) is the forward product of the update (same factors, same
// order); no NaN guard on it: NaN <= 1 is false -> bit 0, like the reference.
__global__ void __launch_bounds__(256)
sw_col_kernel(double *__restrict__ pr, const double *__restrict__ lr, const double *__restrict__ lratio,
              uint32_t *__restrict__ decw, const uint32_t *__restrict__ runw, const int32_t *__restrict__ col_ptr,
              const int32_t *__restrict__ col_edge, const int32_t *__restrict__ edge_row, int N, int E, int v0, int v1, int c0,
              int c1) {
    const int lane = threadIdx.x & 31, g = blockIdx.y;
    const int j = v0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= v1) return;
    const uint32_t run = runw[g];
    if (run == 0) return;
    const bool on = (run >> lane) & 1u;
    const int k0 = __ldg(col_ptr + j), k1 = __ldg(col_ptr + j + 1);
    double *p = pr + (size_t)g * E * kFG + lane;
    const double *l = lr + (size_t)g * E * kFG + lane;
    double acc = on ? lratio[((size_t)g * N + j) * kFG + lane] : 1.0;
    if (on) {
        for (int k = k0; k < k1; k++) {
            const int e = __ldg(col_edge + k), r = __ldg(edge_row + e);
            if (r < c1 && r >= c0) {
                p[(size_t)e * kFG] = acc;
                acc = __dmul_rn(acc, l[(size_t)e * kFG]);
            }
        }
    }
    const uint32_t w = __ballot_sync(0xffffffffu, acc <= 1.0);
    if (lane == 0) {
        uint32_t *dst = decw + (size_t)g * N + j;
        *dst = (w & run) | (*dst & ~run);
    }
    if (on) {
        double s = 1.0;
        for (int k = k1 - 1; k >= k0; k--) {
            const int e = __ldg(col_edge + k), r = __ldg(edge_row + e);
            if (r < c1 && r >= c0) {
                double v = __dmul_rn(p[(size_t)e * kFG], s);
                if (v != v) v = 1.0;
                p[(size_t)e * kFG] = v;
                s = __dmul_rn(s, l[(size_t)e * kFG]);
            }
        }
    }
}

// check_bound + the loop control of Iter_SW_Decoder (dec.cpp:2418-2428), one CTA per group: parity of the checks
// [c0, cc) over the bits [v0, vc) for the 32 frames at once. A running frame whose bounded syndrome is zero, or whose
// count n has reached max_iter, is finished with this position (n is added to its sum); the others iterate once more.
// `final`: the closing check() of Run_SW_Decoder (dec.cpp:2187-2189) over the whole matrix -> success flag, iteration
// figure floor(sum / L) (dec.cpp:2193-2194) and, on request, the syndrome bytes.
__global__ void __launch_bounds__(256)
sw_syn_kernel(const uint32_t *__restrict__ decw, uint32_t *__restrict__ runw, int32_t *__restrict__ n_pos,
              int32_t *__restrict__ sum, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx, int N, int M,
              int v0, int vc, int c0, int cc, int max_iter, unsigned int *__restrict__ remaining, int final, int L, int nf,
              long long frame0, int32_t *__restrict__ iters_out, uint8_t *__restrict__ ok_out, uint8_t *__restrict__ pchk_out) {
    const int g = blockIdx.x;
    const uint32_t run = runw[g];
    const int left = nf - g * kFG;
    const uint32_t valid = left >= 32 ? 0xffffffffu : (left > 0 ? ((1u << left) - 1u) : 0u);
    if (!final && run == 0) return;
    const uint32_t *dw = decw + (size_t)g * N;
    uint32_t acc = 0;
    for (int i = c0 + threadIdx.x; i < cc; i += blockDim.x) {
        uint32_t p = 0;
        for (int e = __ldg(row_ptr + i); e < __ldg(row_ptr + i + 1); e++) {
            const int c = __ldg(col_idx + e);
            if (c >= v0 && c < vc) p ^= dw[c];
        }
        acc |= p;
        if (final && pchk_out)
            for (uint32_t m = valid; m; m &= m - 1) {
                const int f = __ffs(m) - 1;
                pchk_out[(size_t)(frame0 + g * kFG + f) * M + i] = (uint8_t)((p >> f) & 1u);
            }
    }
    acc = __reduce_or_sync(0xffffffffu, acc);
    __shared__ uint32_t s_or[8];
    if ((threadIdx.x & 31) == 0) s_or[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    uint32_t unsat = 0;
    for (int w = 0; w < 8; w++) unsat |= s_or[w];
    const int f = threadIdx.x, slot = g * kFG + f;
    if (final) {
        if ((valid >> f) & 1u) {
            ok_out[frame0 + slot] = (uint8_t)(((unsat >> f) & 1u) ^ 1u);
            iters_out[frame0 + slot] = sum[slot] / L;
        }
        return;
    }
    const bool running = (run >> f) & 1u;
    bool cont = false;
    if (running) {
        const int n = n_pos[slot];
        if (n == max_iter || !((unsat >> f) & 1u)) sum[slot] += n;
        else { n_pos[slot] = n + 1; cont = true; }
    }
    const uint32_t next = __ballot_sync(0xffffffffu, cont);
    if (f == 0) {
        runw[g] = next;
        if (next) atomicAdd(remaining, (unsigned)__popc(next));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Continuous batching by GROUP (round 2, second half). The lock-step kernels above walk a whole wave through the
// positions together, so every position runs until the slowest of up to 4096 frames is done. Here every group of 32
// frames has its own window position: sw2_syn_kernel moves a group on as soon as ITS slowest frame has left the
// position, a group whose last position is done is harvested (closing check(), outputs) and takes the next 32 pending
// frames from one device counter. A tick is one update of every group at its own position; the host only counts
// finished frames, a few ticks late. Per (check, frame) and (bit, frame) the arithmetic is that of the kernels above.
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t SW_LOAD = 1u, SW_INIT = 2u, SW_FINAL = 4u, SW_ZERO = 8u, SW_WANT = 16u;
// Straggler mode of a group (see sw2_thin_in_kernel): SW_THIN while the group's few running frames iterate in the compact
// side arrays, SW_THIN_IN / SW_THIN_OUT = copy their bit->check messages in / back at the head of the next tick.
constexpr uint32_t SW_THIN = 32u, SW_THIN_IN = 64u, SW_THIN_OUT = 128u;
constexpr int kSwThinLanes = 4;     // frames a group can take into straggler mode = doubles per 32-byte sector
constexpr int kSwThinMinLeft = 3;   // ... if they may still run at least this many updates at the position

struct SwGroups {
    int32_t *pos;        // [G] window position of the group, L = no frames
    uint32_t *run;       // [G] lanes that still iterate at this position
    uint32_t *valid;     // [G] lanes that hold a frame
    uint32_t *flags;     // [G] SW_LOAD / SW_ZERO / SW_INIT: work for the head of the next tick; SW_FINAL: harvest me;
                         //     SW_WANT: empty, waiting for frames that are still on their way to the device
    uint32_t *unsat;     // [G] OR over the checks of the closing syndrome, bit = lane
    int32_t *frame0;     // [G] first frame (row of the staged chunk) of the group's frames
    int32_t *n_pos;      // [G*32] extra updates run at this position (Iter_SW_Decoder's n)
    int32_t *sum;        // [G*32] sum of n over the positions done
    const int32_t *sched;  // [L*8] window ranges per position (engine.cu:sw_schedule)
    unsigned long long *next_frame, *done;  // claim counter, frames harvested
    const unsigned long long *avail;        // frames of the chunk that have arrived in HBM (sw2_publish_kernel)
    int32_t *lists;      // [4][G] groups flagged SW_LOAD / SW_INIT / SW_THIN_IN / SW_THIN_OUT in this tick, then the 4 counts
    // straggler mode (thin_pr == nullptr: switched off)
    double *thin_pr, *thin_lr;  // [G][thin_edges][4]: messages of the window's checks for the group's <= 4 stragglers
    uint32_t *thin_map;         // [G] byte k = slot of straggler k, 0xff = none
    int32_t *thin_e0, *thin_ne; // [G] first edge / number of edges of the window's checks when the group went thin
    int thin_edges;             // edges the side arrays hold per group (most edges of any window's checks)
    unsigned long long *hist;   // diagnostics (DNALDPC_SW_HIST=1), else nullptr: [0..32] group-updates by running frames, [33..65] the same in straggler mode
};

// The copy stream says how many frames of the chunk are resident (ordered behind the copy that brought them).
__global__ void sw2_publish_kernel(unsigned long long *avail, unsigned long long frames) {
    *(volatile unsigned long long *)avail = frames;
    __threadfence();
}

// Work lists of a tick's head: the few groups that take new frames / start a position (the load / init kernels then
// walk these instead of launching a CTA per (group, tile) that finds nothing to do).
__global__ void __launch_bounds__(1024)
sw2_list_kernel(SwGroups s, int G) {
    __shared__ int n[4];
    if (threadIdx.x < 4) n[threadIdx.x] = 0;
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const uint32_t f = s.flags[g];
        if (f & SW_LOAD) s.lists[atomicAdd(&n[0], 1)] = g;
        if (f & SW_INIT) s.lists[G + atomicAdd(&n[1], 1)] = g;
        if (f & SW_THIN_IN) s.lists[2 * G + atomicAdd(&n[2], 1)] = g;
        if (f & SW_THIN_OUT) s.lists[3 * G + atomicAdd(&n[3], 1)] = g;
    }
    __syncthreads();
    if (threadIdx.x < 4) s.lists[4 * G + threadIdx.x] = n[threadIdx.x];
}

// Straggler mode. Once a position's quick frames are done, a group's 1-4 frames that iterate on (typically to max_iter)
// would move a 32-byte sector per 8-byte message, next to the values of the waiting frames. Their messages of the
// window's checks (a contiguous edge range in CSR order; every entry the window's updates touch) are therefore copied
// into side arrays with FOUR slots per edge - one sector - and the updates of the position's remaining iterations run
// there; when the position is done the bit->check messages are copied back (the check->bit ones are rewritten by the
// next update before anything reads them). dir = 0: in (list 2), 1: out (list 3).
__global__ void __launch_bounds__(256)
sw2_thin_copy_kernel(double *__restrict__ pr, SwGroups s, int E, int G, int dir) {
    const int n = s.lists[4 * G + 2 + dir];
    for (int a = blockIdx.y; a < n; a += gridDim.y) {
        const int g = s.lists[(2 + dir) * G + a];
        const uint32_t map = s.thin_map[g];
        const int e0 = s.thin_e0[g], ne = s.thin_ne[g];
        double *main_g = pr + ((size_t)g * E + e0) * kFG;
        double *thin_g = s.thin_pr + (size_t)g * s.thin_edges * kSwThinLanes;
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < ne * kSwThinLanes; idx += gridDim.x * blockDim.x) {
            const int e = idx >> 2, k = idx & 3;
            const uint32_t slot = (map >> (8 * k)) & 0xffu;
            if (slot < 32u) {
                if (dir == 0) thin_g[idx] = main_g[(size_t)e * kFG + slot];
                else main_g[(size_t)e * kFG + slot] = thin_g[idx];
            } else if (dir == 0) thin_g[idx] = 1.0;
        }
    }
}
// Harvest bookkeeping + admission, one warp per group (lane = slot). first != 0: start of a chunk, nothing to harvest.
// A group takes the next 32 frames only once they have arrived (avail); until then it waits with SW_WANT.
__global__ void __launch_bounds__(256)
sw2_claim_kernel(SwGroups s, int G, int L, int F, int first, uint32_t load_flags, int32_t *__restrict__ iters_out,
                 uint8_t *__restrict__ ok_out) {
    const int lane = threadIdx.x & 31, g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= G) return;
    const int slot = g * kFG + lane;
    if (!first) {
        const uint32_t fl = s.flags[g];
        if (fl & SW_FINAL) {
            const uint32_t v = s.valid[g];
            if ((v >> lane) & 1u) {
                const int f = s.frame0[g] + lane;
                ok_out[f] = (uint8_t)(((s.unsat[g] >> lane) & 1u) ^ 1u);
                iters_out[f] = s.sum[slot] / L;  // dec.cpp:2193-2194
            }
            if (lane == 0) atomicAdd(s.done, (unsigned long long)__popc(v));
        } else if (!(fl & SW_WANT)) return;
    }
    const unsigned long long none = ~0ull;
    unsigned long long f0 = none, cur = 0;
    if (lane == 0) {
        cur = *(volatile unsigned long long *)s.next_frame;
        const unsigned long long av = *(volatile const unsigned long long *)s.avail;
        while (cur < av) {
            const unsigned long long old = atomicCAS(s.next_frame, cur, cur + 32ull);
            if (old == cur) { f0 = cur; break; }
            cur = old;
        }
    }
    f0 = __shfl_sync(0xffffffffu, f0, 0);
    s.n_pos[slot] = 0;
    s.sum[slot] = 0;
    __syncwarp();  // every lane has read the old frames' state
    if (lane == 0) {
        if (f0 != none) {
            const int left = F - (int)f0;
            const uint32_t v = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
            s.valid[g] = v; s.run[g] = v; s.pos[g] = 0; s.frame0[g] = (int)f0; s.unsat[g] = 0; s.flags[g] = load_flags;
        } else {  // nothing to take: for good once every frame of the chunk has been claimed
            s.valid[g] = 0; s.run[g] = 0; s.pos[g] = L; s.flags[g] = cur >= (unsigned long long)F ? 0u : SW_WANT;
        }
    }
}

// A group's new frames: frame-major staged rows -> slot-interleaved lratio; every decision starts "set" (see sw_load_kernel).
__global__ void __launch_bounds__(256)
sw2_load_kernel(const double *__restrict__ in, double *__restrict__ lratio, uint32_t *__restrict__ decw, SwGroups s, int N, int G) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, j0 = blockIdx.x * 32;
    const int n = s.lists[4 * G];
    for (int a = blockIdx.y; a < n; a += gridDim.y) {
        const int g = s.lists[a];
        const uint32_t v = s.valid[g];
        const size_t f0 = (size_t)s.frame0[g];
        for (int r = ty; r < 32; r += 8) {  // r = slot, tx = bit
            const int j = j0 + tx;
            tile[r][tx] = (((v >> r) & 1u) && j < N) ? in[(f0 + r) * N + j] : 1.0;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {  // r = bit, tx = slot
            const int j = j0 + r;
            if (j < N) {
                lratio[((size_t)g * N + j) * kFG + tx] = tile[tx][r];
                if (tx == 0) decw[(size_t)g * N + j] = 0xffffffffu;
            }
        }
        __syncthreads();
    }
}

// alloc_entry's e->pr = e->lr = 0 for a group's new frames (only for window descriptions under which an entry can be
// read before Init_SW_Decoder has reached its column; the host decides, engine.cu).
__global__ void __launch_bounds__(256)
sw2_zero_kernel(double *__restrict__ pr, double *__restrict__ lr, SwGroups s, int E, int G) {
    const int n = s.lists[4 * G];
    const size_t n2 = (size_t)E * kFG / 2;  // double2 elements per array
    const double2 z = make_double2(0.0, 0.0);
    for (int a = blockIdx.y; a < n; a += gridDim.y) {
        const int g = s.lists[a];
        double2 *pa = (double2 *)(pr + (size_t)g * E * kFG), *pb = (double2 *)(lr + (size_t)g * E * kFG);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
            pa[i] = z;
            pb[i] = z;
        }
    }
}

// Init_SW_Decoder for the columns that enter the window at the group's position.
__global__ void __launch_bounds__(256)
sw2_init_kernel(double *__restrict__ pr, double *__restrict__ lr, const double *__restrict__ lratio, SwGroups s,
                const int32_t *__restrict__ col_ptr, const int32_t *__restrict__ col_edge, int N, int E, int G) {
    const int lane = threadIdx.x & 31;
    const int n = s.lists[4 * G + 1];
    for (int a = blockIdx.y; a < n; a += gridDim.y) {
        const int g = s.lists[G + a];
        const int32_t *r = s.sched + 8 * s.pos[g];
        const int j = r[6] + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        if (j >= r[7]) continue;
        const double v = lratio[((size_t)g * N + j) * kFG + lane];
        for (int k = __ldg(col_ptr + j); k < __ldg(col_ptr + j + 1); k++) {
            const size_t idx = ((size_t)g * E + __ldg(col_edge + k)) * kFG + lane;
            pr[idx] = v;
            lr[idx] = 1.0;
        }
    }
}

// Check_Update_SW for one (check, slot): p / l point at the row's first bit->check / check->bit message of the slot,
// consecutive entries `stride` doubles apart. DC > 0: rows of degree <= DC in registers (sw_row_reg_kernel's
// arithmetic); DC == 0: the generic loop (sw_row_kernel's).
template <int DC>
__device__ __forceinline__ void sw_row_node(const double *__restrict__ p, double *__restrict__ l, int deg, int stride) {
    bool bad = DC == 0;
    if (DC > 0) {
        constexpr int D = DC > 0 ? DC : 1;
        double d[D], Bv[D];
#pragma unroll
        for (int k = 0; k < D; k++) d[k] = k < deg ? ld_stream(p + (size_t)k * stride) : 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) d[k] = k < deg ? check_factor(d[k], bad) : 1.0;  // padding: exact identity in both chains
        if (!bad) {
            double B = 1.0;
#pragma unroll
            for (int k = D - 1; k >= 0; k--) {
                Bv[k] = B;
                B = __dmul_rn(B, d[k]);
            }
            double F = 1.0;
#pragma unroll
            for (int k = 0; k < D; k++) {
                const double t = __dmul_rn(F, Bv[k]);
                if (k < deg) st_stream(l + (size_t)k * stride, check_to_bit(t));
                F = __dmul_rn(F, d[k]);
            }
        }
    }
    if (bad) {  // generic rows, or operands outside the proven ranges: the full-range loop of sw_row_kernel
        double dl = 1.0;
        for (int k = 0; k < deg; k++) {
            l[(size_t)k * stride] = dl;
            dl = __dmul_rn(dl, check_factor_slow(p[(size_t)k * stride]));
        }
        dl = 1.0;
        for (int k = deg - 1; k >= 0; k--) {
            const double t = __dmul_rn(l[(size_t)k * stride], dl);
            l[(size_t)k * stride] = check_to_bit_slow(t);
            dl = __dmul_rn(dl, check_factor_slow(p[(size_t)k * stride]));
        }
    }
}

// Variable_Update_SW + Decision_SW for one (bit j, slot), restricted to the entries whose check lies in [c0, c1)
// (sw_col_kernel's arithmetic); returns the decision. p / l: the slot's message of edge `ebase`, consecutive edges
// `stride` doubles apart; lrat: the slot's channel ratio of bit j. DV > 0: columns of degree <= DV with every load in
// flight before the first multiplication (col_row = check of the k-th entry of a column, so that the in-window test
// needs no dependent lookup): P_k = lratio * lr_0 .. lr_{k-1} over the in-window entries, pr_k = P_k * S_k with S the
// product from the other end - the values sw_col_kernel leaves in e->pr. `on` = false: no access at all.
template <int DV>
__device__ __forceinline__ bool sw_col_node(double *__restrict__ p, const double *__restrict__ l, int stride, int ebase,
                                            const double *__restrict__ lrat, const int32_t *__restrict__ col_ptr,
                                            const int32_t *__restrict__ col_edge, const int32_t *__restrict__ col_row, int j, int c0,
                                            int c1, bool on) {
    if (!on) return true;
    const int k0 = __ldg(col_ptr + j), k1 = __ldg(col_ptr + j + 1);
    double acc = *lrat;
    if (DV > 0) {
        constexpr int D = DV > 0 ? DV : 1;
        int e[D];
        bool in[D];
        double lv[D], P[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            in[k] = false;
            e[k] = 0;
            if (k0 + k < k1) {
                const int row = __ldg(col_row + k0 + k);
                in[k] = row < c1 && row >= c0;
                if (in[k]) e[k] = __ldg(col_edge + k0 + k) - ebase;
            }
        }
#pragma unroll
        for (int k = 0; k < D; k++) lv[k] = in[k] ? ld_stream(l + (size_t)e[k] * stride) : 1.0;
#pragma unroll
        for (int k = 0; k < D; k++) {
            P[k] = acc;
            if (in[k]) acc = __dmul_rn(acc, lv[k]);
        }
        double S = 1.0;
#pragma unroll
        for (int k = D - 1; k >= 0; k--) {
            if (in[k]) {
                double v = __dmul_rn(P[k], S);
                if (v != v) v = 1.0;
                st_stream(p + (size_t)e[k] * stride, v);
                S = __dmul_rn(S, lv[k]);
            }
        }
        return acc <= 1.0;
    }
    for (int k = k0; k < k1; k++) {
        const int row = __ldg(col_row + k);
        if (row < c1 && row >= c0) {
            const size_t o = (size_t)(__ldg(col_edge + k) - ebase) * stride;
            p[o] = acc;
            acc = __dmul_rn(acc, l[o]);
        }
    }
    double sp = 1.0;
    for (int k = k1 - 1; k >= k0; k--) {
        const int row = __ldg(col_row + k);
        if (row < c1 && row >= c0) {
            const size_t o = (size_t)(__ldg(col_edge + k) - ebase) * stride;
            double v = __dmul_rn(p[o], sp);
            if (v != v) v = 1.0;
            p[o] = v;
            sp = __dmul_rn(sp, l[o]);
        }
    }
    return acc <= 1.0;
}

// Lanes a group gives to one node. A group in which only a few frames still iterate at its position (the stragglers that
// run to max_iter while the others wait) would otherwise spend a whole warp per node on 1-4 useful lanes, and the
// update kernels, bound by warps in flight, would take as long as for a full group: with n <= 4 / 8 / 16 running frames
// a warp takes 8 / 4 / 2 nodes at once, lane = (node, k-th running slot).
__device__ __forceinline__ int sw_lanes_per_node(uint32_t run, int packed) {
    const int n = __popc(run);
    return !packed ? 32 : (n <= 4 ? 4 : (n <= 8 ? 8 : (n <= 16 ? 16 : 32)));
}

__device__ __forceinline__ int sw_nth_set_bit(uint32_t m, int n) {  // position of the n-th (0-based) set bit, -1 if there are fewer
    for (int i = 0; i < n; i++) m &= m - 1;
    return m ? __ffs(m) - 1 : -1;
}

constexpr int kSwNodesPerCta = 64;  // checks / bits a CTA of 8 warps walks through

// Which slot a lane works for and where its messages live: the group's slot-interleaved arrays (stride 32, edge 0 of
// the matrix) or, in straggler mode, the side arrays (stride 4, first edge of the window's checks).
struct SwLane {
    int lpn, slot;     // lanes per node; slot = -1: nothing to do
    int stride, ebase;
    double *pr, *lr;   // the slot's message of edge `ebase`
};
__device__ __forceinline__ SwLane sw_lane_setup(const SwGroups &s, double *pr, double *lr, int g, int E, uint32_t run, int lane, int packed) {
    SwLane L;
    if (s.flags[g] & SW_THIN) {
        const int k = lane & (kSwThinLanes - 1);
        const uint32_t sl = (s.thin_map[g] >> (8 * k)) & 0xffu;
        L.lpn = kSwThinLanes;
        L.slot = (sl < 32u && ((run >> sl) & 1u)) ? (int)sl : -1;
        L.stride = kSwThinLanes;
        L.ebase = s.thin_e0[g];
        const size_t o = (size_t)g * s.thin_edges * kSwThinLanes + k;
        L.pr = s.thin_pr + o;
        L.lr = s.thin_lr + o;
        return L;
    }
    L.lpn = sw_lanes_per_node(run, packed);
    const int sl = L.lpn == 32 ? lane : sw_nth_set_bit(run, lane & (L.lpn - 1));
    L.slot = (sl >= 0 && ((run >> sl) & 1u)) ? sl : -1;
    L.stride = kFG;
    L.ebase = 0;
    const size_t o = (size_t)g * E * kFG + (sl >= 0 ? sl : 0);
    L.pr = pr + o;
    L.lr = lr + o;
    return L;
}

template <int DC>
__global__ void __launch_bounds__(256)
sw2_row_kernel(double *__restrict__ pr, double *__restrict__ lr, SwGroups s, const int32_t *__restrict__ row_ptr, int E,
               int packed) {
    const int g = blockIdx.y;
    const uint32_t run = s.run[g];
    if (run == 0) return;
    const int32_t *r = s.sched + 8 * s.pos[g];
    const int i0 = r[2] + blockIdx.x * kSwNodesPerCta, i1 = min(r[3], i0 + kSwNodesPerCta);
    if (i0 >= i1) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const SwLane L = sw_lane_setup(s, pr, lr, g, E, run, lane, packed);
    if (L.slot < 0) return;
    const int npw = 32 / L.lpn, sub = lane / L.lpn;
    for (int i = i0 + warp * npw + sub; i < i1; i += 8 * npw) {
        const int e0 = __ldg(row_ptr + i), deg = __ldg(row_ptr + i + 1) - e0;
        const size_t o = (size_t)(e0 - L.ebase) * L.stride;
        sw_row_node<DC>(L.pr + o, L.lr + o, deg, L.stride);
    }
}

template <int DV>
__global__ void __launch_bounds__(256)
sw2_col_kernel(double *__restrict__ pr, double *__restrict__ lr, const double *__restrict__ lratio,
               uint32_t *__restrict__ decw, SwGroups s, const int32_t *__restrict__ col_ptr, const int32_t *__restrict__ col_edge,
               const int32_t *__restrict__ col_row, int N, int E, int packed) {
    const int g = blockIdx.y;
    const uint32_t run = s.run[g];
    if (run == 0) return;
    const int32_t *r = s.sched + 8 * s.pos[g];
    const int j0 = r[0] + blockIdx.x * kSwNodesPerCta, j1 = min(r[1], j0 + kSwNodesPerCta);
    if (j0 >= j1) return;
    const int c0 = r[2], c1 = r[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const SwLane L = sw_lane_setup(s, pr, lr, g, E, run, lane, packed);
    const int npw = 32 / L.lpn, sub = lane / L.lpn, k = lane & (L.lpn - 1);
    const double *lrat = lratio + (size_t)g * N * kFG + (L.slot >= 0 ? L.slot : 0);
    for (int jb = j0 + warp * npw; jb < j1; jb += 8 * npw) {  // warp-uniform trip count: the lanes of a node shuffle below
        const int j = jb + sub;
        const bool on = L.slot >= 0 && j < j1;
        const bool bit = sw_col_node<DV>(L.pr, L.lr, L.stride, L.ebase, lrat + (size_t)j * kFG, col_ptr, col_edge, col_row, j, c0, c1, on);
        uint32_t w = (on && bit) ? (1u << L.slot) : 0u;
        for (int off = 1; off < L.lpn; off <<= 1) w |= __shfl_xor_sync(0xffffffffu, w, off);  // OR over the node's lanes
        if (k == 0 && j < j1) {
            uint32_t *dst = decw + (size_t)g * N + j;
            *dst = (w & run) | (*dst & ~run);
        }
    }
}

// check_bound + the loop control of Iter_SW_Decoder (sw_syn_kernel) at the group's own position, and the step to the
// next position once no frame of the group iterates any more. One CTA per group.
__global__ void __launch_bounds__(256)
sw2_syn_kernel(const uint32_t *__restrict__ decw, SwGroups s, const int32_t *__restrict__ row_ptr,
               const int32_t *__restrict__ col_idx, int N, int max_iter, int L) {
    const int g = blockIdx.x;
    const uint32_t run = s.run[g];
    if (run == 0) return;
    const int t = s.pos[g];
    const int32_t *r = s.sched + 8 * t;
    const int v0 = r[0], vc = r[4], c0 = r[2], cc = r[5];
    if (s.hist != nullptr && threadIdx.x == 0) atomicAdd(s.hist + __popc(run) + ((s.flags[g] & SW_THIN) ? 33 : 0), 1ull);
    const uint32_t *dw = decw + (size_t)g * N;
    uint32_t acc = 0;
    for (int i = c0 + threadIdx.x; i < cc; i += blockDim.x) {
        uint32_t p = 0;
        for (int e = __ldg(row_ptr + i); e < __ldg(row_ptr + i + 1); e++) {
            const int c = __ldg(col_idx + e);
            if (c >= v0 && c < vc) p ^= dw[c];
        }
        acc |= p;
    }
    acc = __reduce_or_sync(0xffffffffu, acc);
    __shared__ uint32_t s_or[8];
    if ((threadIdx.x & 31) == 0) s_or[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    uint32_t unsat = 0;
    for (int w = 0; w < 8; w++) unsat |= s_or[w];
    const int f = threadIdx.x, slot = g * kFG + f;
    const bool running = (run >> f) & 1u;
    bool cont = false;
    int n = 0;
    if (running) {
        n = s.n_pos[slot];
        if (n == max_iter || !((unsat >> f) & 1u)) s.sum[slot] += n;
        else { n = n + 1; cont = true; }
    }
    const uint32_t next = __ballot_sync(0xffffffffu, cont);
    const uint32_t thin = s.flags[g] & SW_THIN;
    if (next) {
        if (cont) s.n_pos[slot] = n;
        // all running frames of a group started the position together, so they share n
        const int n_run = __shfl_sync(0xffffffffu, n, __ffs(next) - 1);
        if (f == 0) {
            uint32_t fl = thin;
            if (!thin && s.thin_pr != nullptr && __popc(next) <= kSwThinLanes && n_run >= 2 && max_iter - n_run >= kSwThinMinLeft) {
                uint32_t map = 0;
                for (int k = 0; k < kSwThinLanes; k++) {
                    const int sl = sw_nth_set_bit(next, k);
                    map |= (uint32_t)(sl < 0 ? 0xff : sl) << (8 * k);
                }
                const int e0 = __ldg(row_ptr + c0), e1 = __ldg(row_ptr + r[3]);
                if (e1 - e0 <= s.thin_edges) {
                    s.thin_map[g] = map; s.thin_e0[g] = e0; s.thin_ne[g] = e1 - e0;
                    fl = SW_THIN | SW_THIN_IN;
                }
            }
            s.run[g] = next; s.flags[g] = fl;
        }
    } else {  // the group's slowest frame has left position t
        s.n_pos[slot] = 0;
        if (f == 0) {
            if (t + 1 < L) { s.pos[g] = t + 1; s.run[g] = s.valid[g]; s.flags[g] = SW_INIT | (thin ? SW_THIN_OUT : 0u); }
            else { s.run[g] = 0; s.flags[g] = SW_FINAL; }
        }
    }
}

// The closing check() of Run_SW_Decoder (dec.cpp:2187-2189) for the groups that are through: kSwFinalSplit CTAs per
// group, OR of the syndrome words into unsat[g], the syndrome bytes on request.
constexpr int kSwFinalSplit = 8;
__global__ void __launch_bounds__(256)
sw2_final_syn_kernel(const uint32_t *__restrict__ decw, SwGroups s, const int32_t *__restrict__ row_ptr,
                     const int32_t *__restrict__ col_idx, int N, int M, uint8_t *__restrict__ pchk_out) {
    const int g = blockIdx.y;
    if (!(s.flags[g] & SW_FINAL)) return;
    const uint32_t valid = s.valid[g];
    const size_t f0 = (size_t)s.frame0[g];
    const uint32_t *dw = decw + (size_t)g * N;
    uint32_t acc = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        uint32_t p = 0;
        for (int e = __ldg(row_ptr + i); e < __ldg(row_ptr + i + 1); e++) p ^= dw[__ldg(col_idx + e)];
        acc |= p;
        if (pchk_out)
            for (uint32_t m = valid; m; m &= m - 1) {
                const int f = __ffs(m) - 1;
                pchk_out[(f0 + f) * M + i] = (uint8_t)((p >> f) & 1u);
            }
    }
    acc = __reduce_or_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicOr(s.unsat + g, acc);
}

// Decisions of the groups that are through -> frame-major outputs. A warp takes 32 bits x 32 slots: bit f of its 32
// decision words is frame f's word (32 ballots).
__global__ void __launch_bounds__(256)
sw2_output_kernel(const uint32_t *__restrict__ decw, SwGroups s, int N, int wpf, uint32_t *__restrict__ bits,
                  uint8_t *__restrict__ dblk) {
    const int g = blockIdx.y;
    if (!(s.flags[g] & SW_FINAL)) return;
    const int lane = threadIdx.x & 31, w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= wpf) return;
    const uint32_t valid = s.valid[g];
    const size_t f0 = (size_t)s.frame0[g];
    const int j = w * 32 + lane;
    const uint32_t mine = j < N ? decw[(size_t)g * N + j] : 0u;
    uint32_t word = 0;
#pragma unroll
    for (int f = 0; f < 32; f++) {
        const uint32_t b = __ballot_sync(0xffffffffu, (mine >> f) & 1u);
        if (lane == f) word = b;
    }
    if (!((valid >> lane) & 1u)) return;
    if (bits) bits[(f0 + lane) * wpf + w] = word;
    if (dblk) {
        uint8_t *d = dblk + (f0 + lane) * N + (size_t)w * 32;
        const int nb = min(32, N - w * 32);
        if (nb == 32 && (N & 15) == 0) {  // 16-byte aligned rows: two 16-byte stores
            uint32_t q[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t nib = word >> (4 * i);
                q[i] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
            }
            ((uint4 *)d)[0] = make_uint4(q[0], q[1], q[2], q[3]);
            ((uint4 *)d)[1] = make_uint4(q[4], q[5], q[6], q[7]);
        } else {
            for (int b = 0; b < nb; b++) d[b] = (uint8_t)((word >> b) & 1u);
        }
    }
}

// Decisions of a wave back to frame-major outputs (packed words and / or 0-1 chars).
__global__ void __launch_bounds__(256)
sw_output_kernel(const uint32_t *__restrict__ decw, int N, int nf, long long frame0, int wpf, uint32_t *__restrict__ bits,
                 uint8_t *__restrict__ dblk) {
    const int slot = blockIdx.y, w = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nf || w >= wpf) return;
    const uint32_t *dw = decw + (size_t)(slot >> 5) * N;
    const int f = slot & 31;
    uint32_t word = 0;
    for (int b = 0; b < 32; b++) {
        const int j = w * 32 + b;
        const uint32_t bit = j < N ? ((dw[j] >> f) & 1u) : 0u;
        word |= bit << b;
        if (dblk && j < N) dblk[(size_t)(frame0 + slot) * N + j] = (uint8_t)bit;
    }
    if (bits) bits[(size_t)(frame0 + slot) * wpf + w] = word;
}

}  // namespace dnaldpc
