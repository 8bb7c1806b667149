// engine.cu - per-device wave scheduler around the kernels of bp_kernels.cuh.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bp_kernels.cuh"

namespace dnaldpc {

#define CK(call)                                      \
    do {                                              \
        cudaError_t e_ = (call);                      \
        if (e_ != cudaSuccess) return fail(e_, #call); \
    } while (0)

int Engine::fail(cudaError_t e, const char *what) {
    err_ = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return DNALDPC_ERR_CUDA;
}
int Engine::fail(const std::string &m, int rc) {
    err_ = m;
    return rc;
}

static size_t in_elem_stride(int kind, int N) {
    switch (kind) {
        case DNALDPC_IN_LR_F64: case DNALDPC_IN_LLR_F64: case DNALDPC_IN_AWGN_F64: return (size_t)N * 8;
        case DNALDPC_IN_AWGN_F32: return (size_t)N * 4;
        case DNALDPC_IN_BSC_BITS: return (size_t)((N + 31) / 32) * 4;
        case DNALDPC_IN_VOTE_I8: return (size_t)N;
    }
    return 0;
}

Engine::Engine(const Code &code, int device, int precision, int wave_frames)
    : device_(device), precision_(precision) {
    M_ = code.M; N_ = code.N; E_ = code.E;
    max_row_deg_ = code.max_row_deg; max_col_deg_ = code.max_col_deg;
    reg_rows_ = code.regular_rows; reg_cols_ = code.regular_cols;
    esz_ = precision == DNALDPC_PREC_F32 ? 4 : 8;
    if (wave_frames <= 0) wave_frames = 4096;
    wave_frames_ = (wave_frames + 31) / 32 * 32;
    if (max_row_deg_ > 128 || max_col_deg_ > 16) {
        err_ = "unsupported code: row degree > 128 or column degree > 16";
        return;
    }
    auto chk = [&](cudaError_t e, const char *w) { if (e != cudaSuccess && err_.empty()) fail(e, w); };
    chk(cudaSetDevice(device_), "cudaSetDevice");
    if (!err_.empty()) return;
    auto up = [&](int32_t **dst, const std::vector<int32_t> &v) {
        chk(cudaMalloc((void **)dst, std::max<size_t>(v.size(), 1) * sizeof(int32_t)), "cudaMalloc(H tables)");
        if (err_.empty() && !v.empty())
            chk(cudaMemcpy(*dst, v.data(), v.size() * sizeof(int32_t), cudaMemcpyHostToDevice), "cudaMemcpy(H tables)");
    };
    up(&d_row_ptr_, code.row_ptr);
    up(&d_col_idx_, code.col_idx);
    up(&d_col_ptr_, code.col_ptr);
    up(&d_col_edge_, code.col_edge);
    chk(cudaMalloc((void **)&d_table_, 256 * sizeof(double)), "cudaMalloc(table)");
    chk(cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, device_), "cudaDeviceGetAttribute");
    for (auto &h : ev_) for (auto &e : h) chk(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    chk(cudaEventCreateWithFlags(&fork_ev_, cudaEventDisableTiming), "cudaEventCreate");
    for (auto &e : join_ev_) chk(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    for (auto &e : prof_ev_) chk(cudaEventCreate(&e), "cudaEventCreate");
    chk(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking), "cudaStreamCreate");
    for (auto &st : sub_) chk(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate");
}

Engine::~Engine() {
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    void *ptrs[] = {d_row_ptr_, d_col_idx_, d_col_ptr_, d_col_edge_, d_msg_, d_lratio_, d_post_, d_decw_, d_actw_,
                    d_iters_, d_ok_, d_table_, d_counters_, s_in_, s_bits_, s_dblk_, s_post_, s_pchk_, d_unsatw_, d_arrive_};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (h_counters_) cudaFreeHost(h_counters_);
    for (auto &h : ev_) for (auto &e : h) if (e) cudaEventDestroy(e);
    if (fork_ev_) cudaEventDestroy(fork_ev_);
    for (auto &e : join_ev_) if (e) cudaEventDestroy(e);
    for (auto &e : prof_ev_) if (e) cudaEventDestroy(e);
    if (own_stream_) cudaStreamDestroy(own_stream_);
    for (auto &st : sub_) if (st) cudaStreamDestroy(st);
}

int Engine::ensure_wave(int nf, bool want_post) {
    const int G = (nf + 31) / 32;
    if (G > cap_groups_) {
        void *ptrs[] = {d_msg_, d_lratio_, d_post_, d_decw_, d_actw_, d_iters_, d_ok_, d_unsatw_, d_arrive_};
        for (void *p : ptrs) if (p) cudaFree(p);
        d_msg_ = d_lratio_ = d_post_ = nullptr; d_decw_ = d_actw_ = d_unsatw_ = nullptr; d_iters_ = nullptr; d_ok_ = nullptr;
        d_arrive_ = nullptr;
        cap_groups_ = 0;
        CK(cudaMalloc(&d_msg_, std::max<size_t>((size_t)G * E_ * kFG * esz_, 16)));
        CK(cudaMalloc(&d_lratio_, (size_t)G * N_ * kFG * esz_));
        CK(cudaMalloc((void **)&d_decw_, (size_t)G * N_ * sizeof(uint32_t)));
        CK(cudaMalloc((void **)&d_actw_, (size_t)G * sizeof(uint32_t)));
        CK(cudaMalloc((void **)&d_unsatw_, (size_t)G * sizeof(uint32_t)));
        CK(cudaMalloc((void **)&d_arrive_, (size_t)G * sizeof(unsigned)));
        CK(cudaMalloc((void **)&d_iters_, (size_t)G * kFG * sizeof(int32_t)));
        CK(cudaMalloc((void **)&d_ok_, (size_t)G * kFG));
        cap_groups_ = G;
    }
    if (want_post && !d_post_) CK(cudaMalloc(&d_post_, (size_t)cap_groups_ * N_ * kFG * esz_));
    return DNALDPC_OK;
}

int Engine::ensure_counters(int max_iter) {
    const int need = kHalves * (max_iter + 2);
    if (need > cap_counters_) {
        if (d_counters_) cudaFree(d_counters_);
        if (h_counters_) cudaFreeHost(h_counters_);
        d_counters_ = h_counters_ = nullptr; cap_counters_ = 0;
        CK(cudaMalloc((void **)&d_counters_, (size_t)need * sizeof(unsigned)));
        CK(cudaMallocHost((void **)&h_counters_, (size_t)need * sizeof(unsigned)));
        cap_counters_ = need;
    }
    return DNALDPC_OK;
}

// ---- kernel dispatch -----------------------------------------------------------------------------

template <typename T, int DC, bool EXACT>
static void launch_row_t(bool first, bool bulk, int sm_count, T *msg, const T *lratio, const uint32_t *actw,
                         const int32_t *row_ptr, const int32_t *col_idx, int M, int N, int E, int g0, int G, cudaStream_t st) {
    const long long items = (long long)G * M;
    if (bulk) {  // persistent TMA-prefetch variant: two CTAs of 4 warps per SM, one 18 KB tile per warp
        const size_t smem = (size_t)kRowTmaWarps * DC * kFG * sizeof(T) + kRowTmaWarps * sizeof(uint64_t);
        static bool attr_set[2] = {false, false};
        if (!attr_set[first]) {
            if (first) cudaFuncSetAttribute(row_pass_tma_kernel<T, DC, EXACT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            else cudaFuncSetAttribute(row_pass_tma_kernel<T, DC, EXACT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_set[first] = true;
        }
        const int per_sm = smem * 2 <= 220 * 1024 ? 2 : 1;
        const long long want = (items + kRowTmaWarps - 1) / kRowTmaWarps;
        const unsigned grid = (unsigned)std::min<long long>((long long)sm_count * per_sm, want);
        if (first) row_pass_tma_kernel<T, DC, EXACT, true><<<grid, kRowTmaWarps * 32, smem, st>>>(msg, lratio, actw, row_ptr, col_idx, M, N, E, g0, G);
        else row_pass_tma_kernel<T, DC, EXACT, false><<<grid, kRowTmaWarps * 32, smem, st>>>(msg, lratio, actw, row_ptr, col_idx, M, N, E, g0, G);
        return;
    }
    const unsigned grid = (unsigned)((items + 3) / 4);
    if (first) row_pass_kernel<T, DC, EXACT, true, 4><<<grid, 128, 0, st>>>(msg, lratio, actw, row_ptr, col_idx, M, N, E, g0, G);
    else row_pass_kernel<T, DC, EXACT, false, 4><<<grid, 128, 0, st>>>(msg, lratio, actw, row_ptr, col_idx, M, N, E, g0, G);
}

template <typename T> int Engine::launch_row(bool first, int g0, int G, bool dense, cudaStream_t st) {
    T *msg = (T *)d_msg_;
    const T *lr = (const T *)d_lratio_;
    // bulk (TMA) variant for full, large waves; register-load variant when few frames are left or the wave is small
    static const int force = [] { const char *e = getenv("DNALDPC_ROW_IMPL"); return !e ? 0 : (!strcmp(e, "ldg") ? 1 : (!strcmp(e, "tma") ? 2 : 0)); }();
    const bool bulk = force == 2 && (long long)G * M_ >= 4LL * sm_count_ * 8 && dense;
#define ROW(DC, EX) launch_row_t<T, DC, EX>(first, bulk, sm_count_, msg, lr, d_actw_, d_row_ptr_, d_col_idx_, M_, N_, E_, g0, G, st)
    if (reg_rows_ && max_row_deg_ == 72) ROW(72, true);
    else if (max_row_deg_ <= 8) ROW(8, false);
    else if (max_row_deg_ <= 32) ROW(32, false);
    else if (max_row_deg_ <= 72) ROW(72, false);
    else ROW(128, false);
#undef ROW
    stats.kernel_launches++;
    CK(cudaGetLastError());
    return DNALDPC_OK;
}

template <typename T, int DV, bool EXACT>
static void launch_col_t(T *msg, const T *lratio, uint32_t *decw, const uint32_t *actw, T *post, const int32_t *col_ptr,
                         const int32_t *col_edge, int N, int E, int g0, int G, cudaStream_t st) {
    const int cpw = 8;  // columns per warp
    dim3 grid((unsigned)((N + kColWarps * cpw - 1) / (kColWarps * cpw)), (unsigned)G);
    col_pass_kernel<T, DV, EXACT><<<grid, kColWarps * 32, 0, st>>>(msg, lratio, decw, actw, post, col_ptr, col_edge, N, E, g0, cpw);
}

template <typename T> int Engine::launch_col(int g0, int G, bool want_post, cudaStream_t st) {
    T *msg = (T *)d_msg_;
    const T *lr = (const T *)d_lratio_;
    T *post = want_post ? (T *)d_post_ : nullptr;
#define COL(DV, EX) launch_col_t<T, DV, EX>(msg, lr, d_decw_, d_actw_, post, d_col_ptr_, d_col_edge_, N_, E_, g0, G, st)
    if (reg_cols_ && max_col_deg_ == 8) COL(8, true);
    else if (reg_cols_ && max_col_deg_ == 3) COL(3, true);
    else if (max_col_deg_ <= 4) COL(4, false);
    else if (max_col_deg_ <= 8) COL(8, false);
    else COL(16, false);
#undef COL
    stats.kernel_launches++;
    CK(cudaGetLastError());
    return DNALDPC_OK;
}

template <typename T, int KIND>
static void launch_setup(const SetupArgs &a, T *lratio, uint32_t *decw, int N, int G, cudaStream_t st) {
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)G);
    setup_kernel<T, KIND><<<grid, 256, 0, st>>>(a, lratio, decw, N);
}

// ---- one wave ------------------------------------------------------------------------------------

template <typename T>
int Engine::run_wave(const dnaldpc_input &in, int nf, int max_iter, const dnaldpc_output &out, cudaStream_t st) {
    const int G = (nf + 31) / 32;
    const bool want_post = out.posterior != nullptr;
    int rc = ensure_wave(nf, want_post);
    if (rc) return rc;
    rc = ensure_counters(max_iter);
    if (rc) return rc;

    SetupArgs a;
    a.data = in.data;
    a.frame_stride = in.frame_stride ? in.frame_stride : in_elem_stride(in.kind, N_);
    a.param = in.param;
    a.table = d_table_;
    a.nframes = nf;
    if (in.kind == DNALDPC_IN_BSC_BITS || in.kind == DNALDPC_IN_VOTE_I8) {
        double tab[256];
        int cnt = 256;
        if (in.kind == DNALDPC_IN_BSC_BITS) { dnaldpc_bsc_table(in.param, tab); cnt = 2; }
        else if (in.table) memcpy(tab, in.table, sizeof(tab));
        else dnaldpc_vote_table(in.param, tab);
        // tiny synchronous-safe upload: pageable source is copied to a driver staging buffer before returning
        CK(cudaMemcpyAsync(d_table_, tab, cnt * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    T *lr = (T *)d_lratio_;
    switch (in.kind) {
        case DNALDPC_IN_LR_F64: launch_setup<T, IN_LR_F64>(a, lr, d_decw_, N_, G, st); break;
        case DNALDPC_IN_LLR_F64: launch_setup<T, IN_LLR_F64>(a, lr, d_decw_, N_, G, st); break;
        case DNALDPC_IN_BSC_BITS: launch_setup<T, IN_BSC_BITS>(a, lr, d_decw_, N_, G, st); break;
        case DNALDPC_IN_AWGN_F32: launch_setup<T, IN_AWGN_F32>(a, lr, d_decw_, N_, G, st); break;
        case DNALDPC_IN_AWGN_F64: launch_setup<T, IN_AWGN_F64>(a, lr, d_decw_, N_, G, st); break;
        case DNALDPC_IN_VOTE_I8: launch_setup<T, IN_VOTE_I8>(a, lr, d_decw_, N_, G, st); break;
        default: return fail("unknown input kind", DNALDPC_ERR_ARG);
    }
    init_state_kernel<<<(G * kFG + 255) / 256, 256, 0, st>>>(d_actw_, d_unsatw_, d_arrive_, d_iters_, d_ok_, G, nf);
    stats.kernel_launches += 2;
    CK(cudaGetLastError());
    const int cstride = max_iter + 2;
    CK(cudaMemsetAsync(d_counters_, 0, (size_t)kHalves * cstride * sizeof(unsigned), st));

    // Two halves of the wave iterate independently on their own streams (profiling mode: one half, one stream).
    const int nh = (!profiling && G >= kHalves * kMinGroupsPerHalf) ? kHalves : 1;
    int hg0[kHalves], hgn[kHalves];
    cudaStream_t hs[kHalves];
    bool live[kHalves], dense[kHalves];
    for (int h = 0; h < nh; h++) {
        dense[h] = true;
        hg0[h] = (int)((long long)G * h / nh);
        hgn[h] = (int)((long long)G * (h + 1) / nh) - hg0[h];
        hs[h] = nh == 1 ? st : sub_[h];
        live[h] = true;
    }
    if (nh > 1) {
        CK(cudaEventRecord(fork_ev_, st));
        for (int h = 0; h < nh; h++) CK(cudaStreamWaitEvent(hs[h], fork_ev_, 0));
    }
    // dec.cpp:594-599: for (n = 0;; n++) { c = check(); if (n == max_iter || c == 0) break; iterate; }
    for (int n = 0;; n++) {
        bool any = false;
        for (int h = 0; h < nh; h++) {
            if (!live[h]) continue;
            unsigned *cnt = d_counters_ + (size_t)h * cstride + n;
            syndrome_update_kernel<<<dim3(kSynSplit, (unsigned)hgn[h]), kSynThreads, 0, hs[h]>>>(
                d_decw_, d_actw_, d_iters_, d_ok_, d_row_ptr_, d_col_idx_, d_unsatw_, d_arrive_, M_, N_, hg0[h], n, max_iter, cnt);
            stats.kernel_launches++;
            CK(cudaGetLastError());
            if (n == max_iter) { live[h] = false; continue; }
            CK(cudaMemcpyAsync(h_counters_ + (size_t)h * cstride + n, cnt, sizeof(unsigned), cudaMemcpyDeviceToHost, hs[h]));
            CK(cudaEventRecord(ev_[h][n % (kLag + 1)], hs[h]));
            if (n >= kLag) {  // lagged poll: the host runs at most kLag iterations ahead of the device
                CK(cudaEventSynchronize(ev_[h][(n - kLag) % (kLag + 1)]));
                const unsigned left = h_counters_[(size_t)h * cstride + n - kLag];
                if (left == 0) { live[h] = false; continue; }  // later launches are no-ops
                dense[h] = (long long)left * 2 >= (long long)hgn[h] * kFG;  // at least half of the frames still iterate
            }
            any = true;
        }
        if (!any) break;
        if (profiling) CK(cudaEventRecord(prof_ev_[0], st));
        for (int h = 0; h < nh; h++)
            if (live[h]) { rc = launch_row<T>(n == 0, hg0[h], hgn[h], dense[h], hs[h]); if (rc) return rc; }
        if (profiling) CK(cudaEventRecord(prof_ev_[1], st));
        for (int h = 0; h < nh; h++)
            if (live[h]) { rc = launch_col<T>(hg0[h], hgn[h], want_post, hs[h]); if (rc) return rc; }
        if (profiling) {
            CK(cudaEventRecord(prof_ev_[2], st));
            CK(cudaEventSynchronize(prof_ev_[2]));
            float a_ms = 0, b_ms = 0;
            cudaEventElapsedTime(&a_ms, prof_ev_[0], prof_ev_[1]);
            cudaEventElapsedTime(&b_ms, prof_ev_[1], prof_ev_[2]);
            stats.row_ms += a_ms;
            stats.col_ms += b_ms;
        }
    }
    if (nh > 1) {
        for (int h = 0; h < nh; h++) {
            CK(cudaEventRecord(join_ev_[h], hs[h]));
            CK(cudaStreamWaitEvent(st, join_ev_[h], 0));
        }
    }

    // result gather
    const int wpf = (N_ + 31) / 32;
    if (out.bits) {
        dim3 grid((unsigned)((wpf + 7) / 8), (unsigned)G);
        gather_bits_kernel<<<grid, 256, 0, st>>>(d_decw_, N_, nf, wpf, (size_t)wpf, out.bits);
        stats.kernel_launches++;
    }
    if (out.dblk) {
        dim3 grid((unsigned)((N_ + 255) / 256), (unsigned)G);
        gather_bytes_kernel<<<grid, 256, 0, st>>>(d_decw_, N_, nf, out.dblk);
        stats.kernel_launches++;
    }
    if (out.pchk) {
        dim3 grid((unsigned)((M_ + 255) / 256), (unsigned)G);
        syndrome_bytes_kernel<<<grid, 256, 0, st>>>(d_decw_, d_row_ptr_, d_col_idx_, M_, N_, nf, out.pchk);
        stats.kernel_launches++;
    }
    if (out.posterior) {
        dim3 grid((unsigned)((N_ + 31) / 32), (unsigned)G);
        gather_posterior_kernel<T><<<grid, 256, 0, st>>>((const T *)d_post_, (const T *)d_lratio_, d_iters_, N_, nf, out.posterior);
        stats.kernel_launches++;
    }
    CK(cudaGetLastError());
    if (out.iters) CK(cudaMemcpyAsync(out.iters, d_iters_, (size_t)nf * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (out.is_codeword) CK(cudaMemcpyAsync(out.is_codeword, d_ok_, (size_t)nf, cudaMemcpyDeviceToDevice, st));
    stats.waves++;
    return DNALDPC_OK;
}

static dnaldpc_output slice_out(const dnaldpc_output &o, int64_t f0, int M, int N) {
    const size_t wpf = (size_t)(N + 31) / 32;
    dnaldpc_output s = o;
    if (s.bits) s.bits += (size_t)f0 * wpf;
    if (s.dblk) s.dblk += (size_t)f0 * N;
    if (s.iters) s.iters += f0;
    if (s.is_codeword) s.is_codeword += f0;
    if (s.posterior) s.posterior += (size_t)f0 * N;
    if (s.pchk) s.pchk += (size_t)f0 * M;
    return s;
}

int Engine::decode_device(const dnaldpc_input &in, int64_t F, int max_iter, const dnaldpc_output &out, cudaStream_t stream) {
    if (!err_.empty() && d_row_ptr_ == nullptr) return DNALDPC_ERR_CUDA;
    err_.clear();
    if (F < 0 || max_iter < 0 || (F > 0 && in.data == nullptr)) return fail("bad argument", DNALDPC_ERR_ARG);
    CK(cudaSetDevice(device_));
    stats = dnaldpc_stats{};
    stats.frames = F;
    const size_t stride = in.frame_stride ? in.frame_stride : in_elem_stride(in.kind, N_);
    if (stride == 0) return fail("unknown input kind", DNALDPC_ERR_ARG);
    for (int64_t f0 = 0; f0 < F; f0 += wave_frames_) {
        const int nf = (int)std::min<int64_t>(wave_frames_, F - f0);
        dnaldpc_input w = in;
        w.data = (const char *)in.data + (size_t)f0 * stride;
        w.frame_stride = stride;
        const dnaldpc_output o = slice_out(out, f0, M_, N_);
        int rc = precision_ == DNALDPC_PREC_F32 ? run_wave<float>(w, nf, max_iter, o, stream)
                                                : run_wave<double>(w, nf, max_iter, o, stream);
        if (rc) return rc;
    }
    return DNALDPC_OK;
}

void *Engine::stage(void **buf, size_t *cap, size_t need) {
    if (need > *cap) {
        if (*buf) cudaFree(*buf);
        *buf = nullptr; *cap = 0;
        if (cudaMalloc(buf, need) != cudaSuccess) return nullptr;
        *cap = need;
    }
    return *buf;
}

int Engine::decode_host(const dnaldpc_input &in, int64_t F, int max_iter, const dnaldpc_output &out) {
    if (!err_.empty() && d_row_ptr_ == nullptr) return DNALDPC_ERR_CUDA;
    err_.clear();
    if (F < 0 || max_iter < 0 || (F > 0 && in.data == nullptr)) return fail("bad argument", DNALDPC_ERR_ARG);
    CK(cudaSetDevice(device_));
    cudaStream_t st = own_stream_;
    dnaldpc_stats acc{};
    acc.frames = F;
    const size_t stride = in.frame_stride ? in.frame_stride : in_elem_stride(in.kind, N_);
    const size_t packed = in_elem_stride(in.kind, N_);
    if (packed == 0) return fail("unknown input kind", DNALDPC_ERR_ARG);
    const size_t wpf = (size_t)(N_ + 31) / 32;
    const bool host_exp = in.kind == DNALDPC_IN_LLR_F64 && (in.flags & DNALDPC_FLAG_HOST_EXP);
    for (int64_t f0 = 0; f0 < F; f0 += wave_frames_) {
        const int nf = (int)std::min<int64_t>(wave_frames_, F - f0);
        if (!stage(&s_in_, &c_in_, (size_t)nf * packed)) return fail("out of device memory (input staging)", DNALDPC_ERR_NOMEM);
        const char *src = (const char *)in.data + (size_t)f0 * stride;
        dnaldpc_input w = in;
        if (host_exp) {  // LR = exp(LLR) with the host libm, like LDPC_Encode (DNA_main.cpp:1344)
            h_exp_.resize((size_t)nf * N_);
            for (int f = 0; f < nf; f++) {
                const double *row = (const double *)(src + (size_t)f * stride);
                for (int j = 0; j < N_; j++) h_exp_[(size_t)f * N_ + j] = std::exp(row[j]);
            }
            CK(cudaMemcpyAsync(s_in_, h_exp_.data(), (size_t)nf * packed, cudaMemcpyHostToDevice, st));
            w.kind = DNALDPC_IN_LR_F64;
        } else if (stride == packed) {
            CK(cudaMemcpyAsync(s_in_, src, (size_t)nf * packed, cudaMemcpyHostToDevice, st));
        } else {
            CK(cudaMemcpy2DAsync(s_in_, packed, src, stride, packed, (size_t)nf, cudaMemcpyHostToDevice, st));
        }
        w.data = s_in_;
        w.frame_stride = packed;
        dnaldpc_output o{};
        if (out.bits && !(o.bits = (uint32_t *)stage(&s_bits_, &c_bits_, (size_t)nf * wpf * 4))) return fail("out of device memory", DNALDPC_ERR_NOMEM);
        if (out.dblk && !(o.dblk = (uint8_t *)stage(&s_dblk_, &c_dblk_, (size_t)nf * N_))) return fail("out of device memory", DNALDPC_ERR_NOMEM);
        if (out.posterior && !(o.posterior = (double *)stage(&s_post_, &c_post_, (size_t)nf * N_ * 8))) return fail("out of device memory", DNALDPC_ERR_NOMEM);
        if (out.pchk && !(o.pchk = (uint8_t *)stage(&s_pchk_, &c_pchk_, (size_t)nf * M_))) return fail("out of device memory", DNALDPC_ERR_NOMEM);
        stats = dnaldpc_stats{};
        int rc = precision_ == DNALDPC_PREC_F32 ? run_wave<float>(w, nf, max_iter, o, st) : run_wave<double>(w, nf, max_iter, o, st);
        if (rc) return rc;
        acc.kernel_launches += stats.kernel_launches;
        acc.waves += stats.waves;
        acc.row_ms += stats.row_ms;
        acc.col_ms += stats.col_ms;
        if (out.bits) CK(cudaMemcpyAsync(out.bits + (size_t)f0 * wpf, o.bits, (size_t)nf * wpf * 4, cudaMemcpyDeviceToHost, st));
        if (out.dblk) CK(cudaMemcpyAsync(out.dblk + (size_t)f0 * N_, o.dblk, (size_t)nf * N_, cudaMemcpyDeviceToHost, st));
        if (out.posterior) CK(cudaMemcpyAsync(out.posterior + (size_t)f0 * N_, o.posterior, (size_t)nf * N_ * 8, cudaMemcpyDeviceToHost, st));
        if (out.pchk) CK(cudaMemcpyAsync(out.pchk + (size_t)f0 * M_, o.pchk, (size_t)nf * M_, cudaMemcpyDeviceToHost, st));
        if (out.iters) CK(cudaMemcpyAsync(out.iters + f0, d_iters_, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
        if (out.is_codeword) CK(cudaMemcpyAsync(out.is_codeword + f0, d_ok_, (size_t)nf, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    stats = acc;
    if (out.iters) for (int64_t f = 0; f < F; f++) stats.frame_iters += out.iters[f];
    return DNALDPC_OK;
}

int Engine::synth_bsc(const uint32_t *cw_bits, int n_cw, uint64_t seed, int64_t frame0, int64_t F, double eps,
                      uint32_t *out_bits, cudaStream_t stream) {
    err_.clear();
    if (F < 0 || !out_bits || (cw_bits && n_cw <= 0) || !(eps >= 0.0 && eps <= 1.0)) return fail("bad argument", DNALDPC_ERR_ARG);
    CK(cudaSetDevice(device_));
    const int wpf = (N_ + 31) / 32;
    const long long total = (long long)F * wpf;
    if (total == 0) return DNALDPC_OK;
    const uint64_t thr = (uint64_t)(eps * 9007199254740992.0);
    synth_bsc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(cw_bits, n_cw, seed, frame0, F, N_, wpf, thr, out_bits);
    CK(cudaGetLastError());
    return DNALDPC_OK;
}

int math_selftest(long long n, uint64_t seed, long long *mismatches, std::string &err) {
    unsigned long long *d = nullptr, h = 0;
    cudaError_t e = cudaMalloc((void **)&d, sizeof(h));
    if (e == cudaSuccess) e = cudaMemset(d, 0, sizeof(h));
    if (e == cudaSuccess && n > 0) {
        math_selftest_kernel<<<(unsigned)((n + 255) / 256), 256>>>(n, seed, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (d) cudaFree(d);
    if (e != cudaSuccess) { err = std::string("CUDA error: ") + cudaGetErrorString(e); return DNALDPC_ERR_CUDA; }
    *mismatches = (long long)h;
    return DNALDPC_OK;
}

}  // namespace dnaldpc
