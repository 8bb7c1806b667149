// sources.cu - frame sources: host-buffer batches staged through rings, device-resident batches shared by several
// GPUs, re-decoding rounds over row lists. See sources.h and the frame-queue notes in bp_kernels.cuh (publish_kernel).
#include "sources.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

namespace dnaldpc {

size_t packed_stride(int kind, int N) {
    switch (kind) {
        case DNALDPC_IN_LR_F64: case DNALDPC_IN_LLR_F64: case DNALDPC_IN_AWGN_F64: return (size_t)N * 8;
        case DNALDPC_IN_AWGN_F32: return (size_t)N * 4;
        case DNALDPC_IN_BSC_BITS: return (size_t)((N + 31) / 32) * 4;
        case DNALDPC_IN_VOTE_I8: return (size_t)N;
    }
    return 0;
}

static void add_stats(dnaldpc_stats &a, const dnaldpc_stats &s) {
    a.frames += s.frames; a.frame_iters += s.frame_iters; a.kernel_launches += s.kernel_launches; a.waves += s.waves;
    a.row_ms += s.row_ms; a.col_ms += s.col_ms; a.compactions += s.compactions;
}

static int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- host-buffer batches ---------------------------------------------------------------------------------

namespace {

struct HostBatch {  // one batch of HOST buffers, shared by the engines that decode it
    dnaldpc_input in;
    dnaldpc_output out;
    int64_t F = 0;
    int max_iter = 0, N = 0, M = 0, chunk = 32, exp_threads = 1;
    size_t stride = 0, packed = 0, wpf = 0;
    int64_t n_chunks = 0;
    std::atomic<int64_t> next_chunk{0};
    bool host_exp = false;
};

struct Rec { int64_t q0; int n; int64_t f0; };  // a chunk in an engine's queue: positions [q0, q0+n) = frames [f0, f0+n)

class HostSource : public FrameSource {
  public:
    // in_ring / out_ring: frames the device staging rings hold (multiples of the chunk size). keep = rows are never
    // reused (in_ring == out_ring == capacity): the inputs stay resident for later re-decoding rounds.
    // min_out_ring: how far the output ring may shrink when the device cannot hold it (0 = as given).
    HostSource(Engine &e, HostBatch &b, int64_t in_ring, int64_t out_ring, bool keep, int64_t min_out_ring = 0)
        : e_(e), b_(b), in_ring_(in_ring), out_ring_(out_ring), min_out_ring_(min_out_ring ? min_out_ring : out_ring), keep_(keep) {}
    ~HostSource() override { stop(DNALDPC_ERR_CUDA); }

    int prepare(Session &ss) {
        if (cudaSetDevice(e_.device()) != cudaSuccess) return set_fail(DNALDPC_ERR_CUDA, "cudaSetDevice failed");
        const size_t N = (size_t)b_.N, M = (size_t)b_.M;
        bool ok = false;
        for (;;) {  // a device short of memory gets a shorter output ring (admission then waits for stragglers sooner)
            ok = e_.stage(&e_.s_in_, &e_.c_in_, (size_t)in_ring_ * b_.packed) != nullptr;
            ok = ok && e_.stage(&e_.s_iters_, &e_.c_iters_, (size_t)out_ring_ * 4);
            ok = ok && e_.stage(&e_.s_ok_, &e_.c_ok_, (size_t)out_ring_);
            if (b_.out.bits) ok = ok && e_.stage(&e_.s_bits_, &e_.c_bits_, (size_t)out_ring_ * b_.wpf * 4);
            if (b_.out.dblk) ok = ok && e_.stage(&e_.s_dblk_, &e_.c_dblk_, (size_t)out_ring_ * N);
            if (b_.out.posterior) ok = ok && e_.stage(&e_.s_post_, &e_.c_post_, (size_t)out_ring_ * N * 8);
            if (b_.out.pchk) ok = ok && e_.stage(&e_.s_pchk_, &e_.c_pchk_, (size_t)out_ring_ * M);
            if (ok || out_ring_ <= min_out_ring_) break;
            out_ring_ = std::max<int64_t>(min_out_ring_, round_up(out_ring_ / 2, b_.chunk));
        }
        if (!ok) return set_fail(DNALDPC_ERR_NOMEM, "out of device memory (staging rings of a host batch)");
        if (b_.host_exp) {
            bounce_ = (double *)e_.pinned(2 * (size_t)b_.chunk * N * sizeof(double));
            if (!bounce_) return set_fail(DNALDPC_ERR_NOMEM, "out of pinned host memory (host-side exp)");
            for (auto &ev : bounce_ev_)
                if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return set_fail(DNALDPC_ERR_CUDA, "cudaEventCreate failed");
        }
        ss.in = b_.in;
        ss.in.data = e_.s_in_;
        ss.in.frame_stride = b_.packed;
        if (b_.host_exp) { ss.in.kind = DNALDPC_IN_LR_F64; ss.in.param = 0.0; }
        ss.in.flags &= ~DNALDPC_FLAG_HOST_EXP;
        ss.out = dnaldpc_output{};
        ss.out.iters = (int32_t *)e_.s_iters_;
        ss.out.is_codeword = (uint8_t *)e_.s_ok_;
        if (b_.out.bits) ss.out.bits = (uint32_t *)e_.s_bits_;
        if (b_.out.dblk) ss.out.dblk = (uint8_t *)e_.s_dblk_;
        if (b_.out.posterior) ss.out.posterior = (double *)e_.s_post_;
        if (b_.out.pchk) ss.out.pchk = (uint8_t *)e_.s_pchk_;
        ss.max_iter = b_.max_iter;
        ss.max_frames = keep_ ? in_ring_ : b_.F;
        th_in_ = std::thread([this] { in_loop(); });
        th_out_ = std::thread([this] { out_loop(); });
        return DNALDPC_OK;
    }

    int pump(Engine &, int64_t admitted, int64_t low_water, bool *final) override {
        std::lock_guard<std::mutex> lk(mu_);
        if (rc_) return rc_;
        if (!started_ || admitted != admitted_ || low_water != low_water_) {
            started_ = true; admitted_ = admitted; low_water_ = low_water;
            cv_.notify_all();
        }
        *final = done_pub_;
        return DNALDPC_OK;
    }

    void wait_for_frames() override {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait_for(lk, std::chrono::milliseconds(2), [&] { return abort_ || done_pub_ || recs_.size() > seen_recs_; });
        seen_recs_ = recs_.size();
    }

    // After Engine::run_session returned `engine_rc`: lets the output thread drain, joins, returns the first error.
    int finish(int engine_rc) {
        if (engine_rc) return stop(engine_rc);
        {
            std::lock_guard<std::mutex> lk(mu_);
            finished_ = true;
            cv_.notify_all();
        }
        join();
        return rc_;
    }

    const std::string &error() const { return err_; }
    const std::vector<Rec> &records() const { return recs_; }  // after finish()
    int64_t launches() const { return launches_; }

  private:
    int stop(int rc) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            abort_ = true;
            cv_.notify_all();
        }
        join();
        for (auto &ev : bounce_ev_) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
        return rc;
    }
    void join() {
        if (th_in_.joinable()) th_in_.join();
        if (th_out_.joinable()) th_out_.join();
    }
    int set_fail(int rc, const std::string &m) {
        std::lock_guard<std::mutex> lk(mu_);
        if (!rc_) { rc_ = rc; err_ = m; }
        abort_ = true;
        cv_.notify_all();
        return rc;
    }
    bool cuda_ok(cudaError_t e, const char *what) {
        if (e == cudaSuccess) return true;
        set_fail(DNALDPC_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
        return false;
    }

    // LR = exp(scale * LLR) with the host libm, like LDPC_Encode (DNA_main.cpp:1344); rows split over host threads
    void exp_rows(const char *src, int n, double *dst) {
        const int N = b_.N;
        const double sc = b_.in.param == 0.0 ? 1.0 : b_.in.param;
        auto work = [&](int fa, int fb) {
            for (int f = fa; f < fb; f++) {
                const double *row = (const double *)(src + (size_t)f * b_.stride);
                double *o = dst + (size_t)f * N;
                for (int j = 0; j < N; j++) o[j] = std::exp(sc == 1.0 ? row[j] : sc * row[j]);
            }
        };
        const int nthr = std::max(1, std::min(b_.exp_threads, n / 4 + 1));
        if (nthr == 1) { work(0, n); return; }
        std::vector<std::thread> th;
        for (int t = 1; t < nthr; t++) th.emplace_back(work, (int)((int64_t)n * t / nthr), (int)((int64_t)n * (t + 1) / nthr));
        work(0, (int)((int64_t)n / nthr));
        for (auto &t : th) t.join();
    }

    void in_loop() {
        if (!cuda_ok(cudaSetDevice(e_.device()), "cudaSetDevice")) return;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return started_ || abort_; });
            if (abort_) return;
        }
        cudaStream_t st = e_.io_stream(0);
        // the engine's queue of this session is reset in stream order at ready_event (recorded before its first pump)
        if (!cuda_ok(cudaStreamWaitEvent(st, e_.ready_event(), 0), "cudaStreamWaitEvent")) return;
        int64_t q = 0;
        int slot = 0;
        bool bounce_used[2] = {false, false};
        for (;;) {
            const int64_t need = q + b_.chunk;
            // Nothing left to claim: say so at once. The engine starts compacting its drain tail when the source is
            // final, and room in the rings may be a long time coming (it waits for the frames that run to max_iter).
            auto none_left = [&] { return b_.next_chunk.load(std::memory_order_relaxed) >= b_.n_chunks; };
            if (none_left()) break;
            if (keep_) {
                if (need > in_ring_) break;  // this engine's resident buffers are full: the other engines take the rest
            } else {
                std::unique_lock<std::mutex> lk(mu_);
                while (!(abort_ || none_left() || (need <= admitted_ + in_ring_ && need <= retired_ + out_ring_)))
                    cv_.wait_for(lk, std::chrono::milliseconds(1));  // another engine may take the last chunk meanwhile
            }
            if (none_left()) break;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (abort_) return;
            }
            const int64_t c = b_.next_chunk.fetch_add(1);
            if (c >= b_.n_chunks) break;
            const int64_t f0 = c * b_.chunk;
            const int n = (int)std::min<int64_t>(b_.chunk, b_.F - f0);
            char *dst = (char *)e_.s_in_ + (size_t)(q % in_ring_) * b_.packed;
            const char *src = (const char *)b_.in.data + (size_t)f0 * b_.stride;
            if (b_.host_exp) {
                double *hb = bounce_ + (size_t)slot * b_.chunk * b_.N;
                if (bounce_used[slot] && !cuda_ok(cudaEventSynchronize(bounce_ev_[slot]), "cudaEventSynchronize")) return;
                exp_rows(src, n, hb);
                if (!cuda_ok(cudaMemcpyAsync(dst, hb, (size_t)n * b_.packed, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(H2D)")) return;
                if (!cuda_ok(cudaEventRecord(bounce_ev_[slot], st), "cudaEventRecord")) return;
                bounce_used[slot] = true;
                slot ^= 1;
            } else if (b_.stride == b_.packed) {
                if (!cuda_ok(cudaMemcpyAsync(dst, src, (size_t)n * b_.packed, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(H2D)")) return;
            } else {
                if (!cuda_ok(cudaMemcpy2DAsync(dst, b_.packed, src, b_.stride, b_.packed, (size_t)n, cudaMemcpyHostToDevice, st), "cudaMemcpy2DAsync(H2D)")) return;
            }
            static const bool dma_publish = !(getenv("DNALDPC_PUBLISH_KERNEL") && atoi(getenv("DNALDPC_PUBLISH_KERNEL")) != 0);  // A/B switch
            const int rc = e_.publish(nullptr, (int)(q % in_ring_), (int)(q % out_ring_), n, st, dma_publish);
            if (rc) { set_fail(rc, rc == DNALDPC_ERR_CUDA ? "CUDA error while publishing frames to the engine's queue" : "frame queue overflow"); return; }
            launches_++;
            {
                std::lock_guard<std::mutex> lk(mu_);
                recs_.push_back(Rec{q, n, f0});
                cv_.notify_all();
            }
            q += n;
        }
        std::lock_guard<std::mutex> lk(mu_);
        done_pub_ = true;
        cv_.notify_all();
    }

    void out_loop() {
        if (!cuda_ok(cudaSetDevice(e_.device()), "cudaSetDevice")) return;
        cudaStream_t st = e_.io_stream(1);
        const size_t N = (size_t)b_.N, M = (size_t)b_.M, wpf = b_.wpf;
        const dnaldpc_output &o = b_.out;
        size_t k = 0;
        std::vector<Rec> todo;
        for (;;) {
            todo.clear();
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] {
                    if (abort_) return true;
                    if (k < recs_.size() && low_water_ >= recs_[k].q0 + recs_[k].n) return true;
                    return finished_ && done_pub_ && k == recs_.size();
                });
                if (abort_) return;
                while (k + todo.size() < recs_.size() && low_water_ >= recs_[k + todo.size()].q0 + recs_[k + todo.size()].n)
                    todo.push_back(recs_[k + todo.size()]);
            }
            if (todo.empty()) return;  // finished and everything retired
            for (const Rec &r : todo) {
                // the harvest kernels that wrote these rows completed before the engine reported the low-water mark
                const size_t row = (size_t)(r.q0 % out_ring_), f0 = (size_t)r.f0, n = (size_t)r.n;
                bool ok = true;
                if (o.bits) ok = ok && cuda_ok(cudaMemcpyAsync(o.bits + f0 * wpf, (uint32_t *)e_.s_bits_ + row * wpf, n * wpf * 4, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(D2H)");
                if (o.dblk) ok = ok && cuda_ok(cudaMemcpyAsync(o.dblk + f0 * N, (uint8_t *)e_.s_dblk_ + row * N, n * N, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(D2H)");
                if (o.posterior) ok = ok && cuda_ok(cudaMemcpyAsync(o.posterior + f0 * N, (double *)e_.s_post_ + row * N, n * N * 8, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(D2H)");
                if (o.pchk) ok = ok && cuda_ok(cudaMemcpyAsync(o.pchk + f0 * M, (uint8_t *)e_.s_pchk_ + row * M, n * M, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(D2H)");
                if (o.iters) ok = ok && cuda_ok(cudaMemcpyAsync(o.iters + f0, (int32_t *)e_.s_iters_ + row, n * 4, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(D2H)");
                if (o.is_codeword) ok = ok && cuda_ok(cudaMemcpyAsync(o.is_codeword + f0, (uint8_t *)e_.s_ok_ + row, n, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(D2H)");
                if (!ok) return;
            }
            if (!cuda_ok(cudaStreamSynchronize(st), "cudaStreamSynchronize(D2H)")) return;
            k += todo.size();
            std::lock_guard<std::mutex> lk(mu_);
            retired_ = todo.back().q0 + todo.back().n;
            cv_.notify_all();
        }
    }

    Engine &e_;
    HostBatch &b_;
    const int64_t in_ring_;
    int64_t out_ring_;
    const int64_t min_out_ring_;
    const bool keep_;
    std::mutex mu_;
    std::condition_variable cv_;
    // guarded by mu_
    bool started_ = false, done_pub_ = false, finished_ = false, abort_ = false;
    int64_t admitted_ = 0, low_water_ = 0, retired_ = 0;
    std::vector<Rec> recs_;
    size_t seen_recs_ = 0;
    int rc_ = DNALDPC_OK;
    std::string err_;
    // input thread only
    double *bounce_ = nullptr;
    cudaEvent_t bounce_ev_[2] = {nullptr, nullptr};
    int64_t launches_ = 0;
    std::thread th_in_, th_out_;
};

int setup_host_batch(HostBatch &b, const EngineSet &eng, const dnaldpc_input &in, int64_t F, int max_iter, const dnaldpc_output &out, int nd) {
    b.in = in; b.out = out; b.F = F; b.max_iter = max_iter;
    b.N = eng[0]->N(); b.M = eng[0]->M();
    b.packed = packed_stride(in.kind, b.N);
    b.stride = in.frame_stride ? in.frame_stride : b.packed;
    b.wpf = (size_t)(b.N + 31) / 32;
    b.host_exp = in.kind == DNALDPC_IN_LLR_F64 && (in.flags & DNALDPC_FLAG_HOST_EXP) && !(in.flags & DNALDPC_FLAG_MINSUM);
    // chunk: about 32 MB of input (PCIe-efficient copies, small bounce buffers), at most 2048 frames, and small enough
    // that every engine sees several chunks (dynamic balance); a multiple of 32 frames
    int64_t chunk = std::max<int64_t>(32, std::min<int64_t>(2048, ((int64_t)32 << 20) / (int64_t)b.packed / 32 * 32));
    chunk = std::min<int64_t>(chunk, std::max<int64_t>(32, round_up((F + 4 * nd - 1) / (4 * nd), 32)));
    if (const char *t = getenv("DNALDPC_CHUNK_FRAMES")) chunk = std::max<int64_t>(32, round_up(atoll(t), 32));  // test switch
    b.chunk = (int)chunk;
    b.n_chunks = (F + chunk - 1) / chunk;
    b.exp_threads = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency() / (unsigned)std::max(1, nd)));
    return DNALDPC_OK;
}

size_t out_bytes_per_frame(const dnaldpc_output &out, int N, int M) {
    return 5 + (out.bits ? (size_t)((N + 31) / 32) * 4 : 0) + (out.dblk ? (size_t)N : 0) + (out.posterior ? (size_t)N * 8 : 0) + (out.pchk ? (size_t)M : 0);
}

// Runs one session per engine (engine 0 on the calling thread) and folds the results.
template <typename Run>
int run_engines(int n, Run run) {
    std::vector<int> rcs((size_t)n, DNALDPC_OK);
    std::vector<std::thread> th;
    for (int k = 1; k < n; k++) th.emplace_back([&, k] { rcs[(size_t)k] = run(k); });
    rcs[0] = run(0);
    for (auto &t : th) t.join();
    for (int k = 0; k < n; k++) if (rcs[(size_t)k]) return rcs[(size_t)k];
    return DNALDPC_OK;
}

}  // namespace

int decode_host_batch(EngineSet &eng, const dnaldpc_input &in, int64_t F, int max_iter, const dnaldpc_output &out,
                      dnaldpc_stats &st, std::string &err) {
    st = dnaldpc_stats{};
    if (F == 0) return DNALDPC_OK;
    if (!in.data) { err = "null input buffer"; return DNALDPC_ERR_ARG; }
    HostBatch b;
    setup_host_batch(b, eng, in, F, max_iter, out, (int)eng.size());
    const int nd = (int)std::min<int64_t>((int64_t)eng.size(), b.n_chunks);
    // rings: inputs cover one wave of admissions; outputs are retired in queue order, so the ring has to span the frames
    // admitted while a frame that runs to max_iter sits in its slot (up to ~60 000 in the fast regimes)
    const int64_t wave = round_up(eng[0]->wave_frames(), b.chunk);
    const int64_t in_ring = std::min(round_up(F, b.chunk), std::max<int64_t>(4 * b.chunk, wave));
    const int64_t budget = (int64_t)8 << 30;
    int64_t out_ring = std::min<int64_t>(262144, budget / (int64_t)out_bytes_per_frame(out, b.N, b.M));
    const int64_t min_out_ring = std::min(std::max<int64_t>(4 * b.chunk, 2 * wave), round_up(F, b.chunk));
    out_ring = std::max<int64_t>(round_up(out_ring, b.chunk), min_out_ring);
    out_ring = std::min(out_ring, round_up(F, b.chunk));
    int64_t in_ring_used = in_ring, min_out_used = min_out_ring;
    if (const char *t = getenv("DNALDPC_RING_CHUNKS")) {  // test switch "in,out": ring sizes in chunks, to force wrap-around on small batches
        long a = 0, c = 0;
        if (sscanf(t, "%ld,%ld", &a, &c) == 2 && a >= 1 && c >= 1) {
            in_ring_used = a * b.chunk;
            out_ring = min_out_used = c * b.chunk;
        }
    }
    std::vector<std::unique_ptr<HostSource>> src;
    std::vector<Session> ss((size_t)nd);
    for (int k = 0; k < nd; k++) src.emplace_back(new HostSource(*eng[(size_t)k], b, in_ring_used, out_ring, false, min_out_used));
    std::vector<std::string> errs((size_t)nd);
    const int rc = run_engines(nd, [&](int k) {
        Engine &e = *eng[(size_t)k];
        int r = src[(size_t)k]->prepare(ss[(size_t)k]);
        if (r) { src[(size_t)k]->finish(r); errs[(size_t)k] = src[(size_t)k]->error(); return r; }
        r = e.run_session(ss[(size_t)k], *src[(size_t)k], nullptr);
        if (r) errs[(size_t)k] = e.error();
        const int r2 = src[(size_t)k]->finish(r);
        if (!r && r2) errs[(size_t)k] = src[(size_t)k]->error();
        return r ? r : r2;
    });
    for (int k = 0; k < nd; k++) {
        if (!errs[(size_t)k].empty() && err.empty()) err = errs[(size_t)k];
        add_stats(st, eng[(size_t)k]->stats);
        st.kernel_launches += src[(size_t)k]->launches();
    }
    return rc;
}

// ---- device-resident batches -----------------------------------------------------------------------------

namespace {

struct DeviceBatch {
    int64_t F = 0, n_chunks = 0;
    int chunk = 32;
    std::atomic<int64_t> next_chunk{0};
};

class DeviceSource : public FrameSource {
  public:
    DeviceSource(DeviceBatch &b, cudaStream_t st, int64_t lookahead) : b_(b), st_(st), lookahead_(lookahead) {}
    int pump(Engine &e, int64_t admitted, int64_t, bool *final) override {
        while (!exhausted_ && e.published() - admitted < lookahead_) {
            const int64_t c = b_.next_chunk.fetch_add(1);
            if (c >= b_.n_chunks) { exhausted_ = true; break; }
            const int64_t f0 = c * b_.chunk;
            const int n = (int)std::min<int64_t>(b_.chunk, b_.F - f0);
            const int rc = e.publish(nullptr, (int)f0, (int)f0, n, st_);  // rows = frames of the caller's buffers
            if (rc) return rc;
            launches++;
        }
        *final = exhausted_;
        return DNALDPC_OK;
    }
    int64_t launches = 0;

  private:
    DeviceBatch &b_;
    cudaStream_t st_;
    int64_t lookahead_;
    bool exhausted_ = false;
};

}  // namespace

int decode_device_batch(EngineSet &eng, const dnaldpc_input &in, int64_t F, int max_iter, const dnaldpc_output &out,
                        cudaStream_t stream, dnaldpc_stats &st, std::string &err) {
    st = dnaldpc_stats{};
    if (F == 0) return DNALDPC_OK;
    if (!in.data) { err = "null input buffer"; return DNALDPC_ERR_ARG; }
    if (F > 0x7fffffffLL) { err = "more than 2^31-1 frames in one call"; return DNALDPC_ERR_ARG; }
    const int N = eng[0]->N();
    // Engines other than the first reach the buffers through NVLink peer access; those that cannot are left out.
    std::vector<Engine *> use{eng[0].get()};
    if (eng.size() > 1 && F >= 64) {
        cudaPointerAttributes at{};
        int owner = eng[0]->device();
        if (cudaPointerGetAttributes(&at, in.data) == cudaSuccess && at.type == cudaMemoryTypeDevice) owner = at.device;
        else cudaGetLastError();
        for (size_t k = 1; k < eng.size(); k++) {
            const int dev = eng[k]->device();
            bool reach = dev == owner;
            if (!reach) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, dev, owner) == cudaSuccess && can && cudaSetDevice(dev) == cudaSuccess) {
                    const cudaError_t pe = cudaDeviceEnablePeerAccess(owner, 0);
                    reach = pe == cudaSuccess || pe == cudaErrorPeerAccessAlreadyEnabled;
                }
                cudaGetLastError();
            }
            if (reach) use.push_back(eng[k].get());
        }
        if (eng[0]->device() != owner) {  // the first engine is a peer of the buffers as well
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, eng[0]->device(), owner) == cudaSuccess && can && cudaSetDevice(eng[0]->device()) == cudaSuccess)
                cudaDeviceEnablePeerAccess(owner, 0);
            cudaGetLastError();
        }
    }
    const int nd = (int)use.size();
    DeviceBatch b;
    b.F = F;
    if (nd == 1) b.chunk = (int)F;
    else b.chunk = (int)std::max<int64_t>(32, std::min<int64_t>(1024, round_up(F / (16 * nd), 32)));
    b.n_chunks = (F + b.chunk - 1) / b.chunk;
    if (nd > 1) {  // what the caller queued on `stream` (the inputs) must be complete before a peer reads it
        if (cudaSetDevice(eng[0]->device()) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) {
            err = "CUDA error while synchronising the caller's stream";
            return DNALDPC_ERR_CUDA;
        }
    }
    Session ss;
    ss.in = in;
    ss.in.frame_stride = in.frame_stride ? in.frame_stride : packed_stride(in.kind, N);
    ss.out = out;
    ss.max_iter = max_iter;
    ss.max_frames = F;
    std::vector<std::unique_ptr<DeviceSource>> src;
    for (int k = 0; k < nd; k++)
        src.emplace_back(new DeviceSource(b, k == 0 ? stream : use[(size_t)k]->own_stream(),
                                          nd == 1 ? (int64_t)1 << 62 : (int64_t)use[(size_t)k]->wave_frames()));
    std::vector<std::string> errs((size_t)nd);
    const int rc = run_engines(nd, [&](int k) {
        Engine &e = *use[(size_t)k];
        // a caller that keeps iters / is_codeword NULL gets per-engine scratch (Engine::run); with its own arrays every
        // engine writes the rows of the frames it decoded
        int r = e.run_session(ss, *src[(size_t)k], k == 0 ? stream : e.own_stream());
        if (r) errs[(size_t)k] = e.error();
        else if (k > 0 && cudaStreamSynchronize(e.own_stream()) != cudaSuccess) { r = DNALDPC_ERR_CUDA; errs[(size_t)k] = "CUDA error on a peer engine's stream"; }
        return r;
    });
    for (int k = 0; k < nd; k++) {
        if (!errs[(size_t)k].empty() && err.empty()) err = errs[(size_t)k];
        add_stats(st, use[(size_t)k]->stats);
        st.kernel_launches += src[(size_t)k]->launches;
    }
    if (nd > 1) cudaSetDevice(eng[0]->device());
    return rc;
}

// ---- re-decoding sweep over device-resident inputs ---------------------------------------------------------

namespace {

class ListSource : public FrameSource {  // one re-decoding round: the frames of a device-side row list, outputs compact
  public:
    ListSource(const int32_t *list, int n, cudaStream_t st) : list_(list), n_(n), st_(st) {}
    int pump(Engine &e, int64_t, int64_t, bool *final) override {
        if (!done_) {
            const int rc = e.publish(list_, 0, 0, n_, st_);
            if (rc) return rc;
            done_ = true;
        }
        *final = true;
        return DNALDPC_OK;
    }

  private:
    const int32_t *list_;
    int n_;
    cudaStream_t st_;
    bool done_ = false;
};

}  // namespace

int redecode_sweep_batch(EngineSet &eng, const dnaldpc_input &in, int64_t F, int max_iter, const double *params, int n_params,
                         const dnaldpc_output &out, int32_t *rounds, dnaldpc_stats &st, std::string &err) {
    st = dnaldpc_stats{};
    if (F == 0) return DNALDPC_OK;
    if (!in.data) { err = "null input buffer"; return DNALDPC_ERR_ARG; }
    const int N = eng[0]->N(), M = eng[0]->M();
    const size_t wpf = (size_t)(N + 31) / 32;
    const size_t packed = packed_stride(in.kind, N);
    const size_t stride = in.frame_stride ? in.frame_stride : packed;
    // Inputs and outputs of every frame stay in HBM for the whole sweep: batches too large for that are swept in parts.
    {
        const size_t per_frame = packed + out_bytes_per_frame(out, N, M);
        size_t free_b = 0, total_b = 0;
        if (cudaSetDevice(eng[0]->device()) != cudaSuccess || cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { err = "cudaMemGetInfo failed"; return DNALDPC_ERR_CUDA; }
        const int64_t fit = std::max<int64_t>(4096, (int64_t)((double)free_b * 0.4 / (double)per_frame) * (int64_t)eng.size() / 2);
        if (F > fit) {
            for (int64_t f0 = 0; f0 < F; f0 += fit) {
                const int64_t nf = std::min(fit, F - f0);
                dnaldpc_input pi = in;
                pi.data = (const char *)in.data + (size_t)f0 * stride;
                pi.frame_stride = stride;
                dnaldpc_output po = out;
                if (po.bits) po.bits += (size_t)f0 * wpf;
                if (po.dblk) po.dblk += (size_t)f0 * N;
                if (po.iters) po.iters += f0;
                if (po.is_codeword) po.is_codeword += f0;
                if (po.posterior) po.posterior += (size_t)f0 * N;
                if (po.pchk) po.pchk += (size_t)f0 * M;
                dnaldpc_stats ps{};
                const int rc = redecode_sweep_batch(eng, pi, nf, max_iter, params, n_params, po, rounds ? rounds + f0 : nullptr, ps, err);
                add_stats(st, ps);
                if (rc) return rc;
            }
            return DNALDPC_OK;
        }
    }
    std::vector<uint8_t> ok_host((size_t)F, 0);
    dnaldpc_output o0 = out;
    o0.is_codeword = ok_host.data();
    dnaldpc_input in0 = in;
    in0.param = params[0];
    HostBatch b;
    setup_host_batch(b, eng, in0, F, max_iter, o0, (int)eng.size());
    const int nd = (int)std::min<int64_t>((int64_t)eng.size(), b.n_chunks);
    // capacity of an engine's resident buffers: its fair share and a quarter more (the engines pull chunks dynamically)
    const int64_t cap = nd == 1 ? round_up(F, b.chunk) : round_up((F + nd - 1) / nd * 5 / 4 + b.chunk, b.chunk);
    std::vector<std::unique_ptr<HostSource>> src;
    std::vector<Session> ss((size_t)nd);
    for (int k = 0; k < nd; k++) src.emplace_back(new HostSource(*eng[(size_t)k], b, cap, cap, true));
    std::vector<std::string> errs((size_t)nd);
    std::vector<dnaldpc_stats> est((size_t)nd);
    if (rounds) for (int64_t f = 0; f < F; f++) rounds[f] = 0;
    const int rc = run_engines(nd, [&](int k) {
        Engine &e = *eng[(size_t)k];
        HostSource &s = *src[(size_t)k];
        std::string &er = errs[(size_t)k];
        int r = s.prepare(ss[(size_t)k]);
        if (r) { s.finish(r); er = s.error(); return r; }
        r = e.run_session(ss[(size_t)k], s, nullptr);  // round 0: every frame, uploaded and copied back as it goes
        if (r) er = e.error();
        const int r2 = s.finish(r);
        if (!r && r2) { er = s.error(); r = r2; }
        if (r) return r;
        est[(size_t)k] = e.stats;
        est[(size_t)k].kernel_launches += s.launches();
        // frame of every resident row
        int64_t rows = 0;
        for (const Rec &rec : s.records()) rows = std::max(rows, rec.q0 + rec.n);
        if (rows == 0 || n_params <= 1) return DNALDPC_OK;
        std::vector<int64_t> row_frame((size_t)rows);
        for (const Rec &rec : s.records())
            for (int i = 0; i < rec.n; i++) row_frame[(size_t)(rec.q0 + i)] = rec.f0 + i;
        cudaStream_t stq = e.own_stream();
        auto cu = [&](cudaError_t ce, const char *what) {
            if (ce == cudaSuccess) return false;
            er = std::string("CUDA error: ") + cudaGetErrorString(ce) + " in " + what;
            return true;
        };
        if (cu(cudaSetDevice(e.device()), "cudaSetDevice")) return DNALDPC_ERR_CUDA;
        r = e.ensure_lists(rows);
        if (r) { er = e.error(); return r; }
        const int32_t *prev = nullptr;
        int n_prev = (int)rows;
        std::vector<int32_t> list_h;
        std::vector<uint8_t> t_ok, t_u8;
        std::vector<int32_t> t_it;
        std::vector<uint32_t> t_bits;
        std::vector<double> t_post;
        for (int rd = 1; rd < n_params; rd++) {
            int32_t *list = e.d_list_[rd & 1];
            int32_t *count_dev = list + rows;
            // the syndrome flags of the previous round: by row after round 0, compact (in list order) afterwards
            r = e.failed_rows((const uint8_t *)e.s_ok_, prev, n_prev, list, count_dev, stq);
            if (r) { er = e.error(); return r; }
            int32_t K = 0;
            if (cu(cudaMemcpyAsync(&K, count_dev, sizeof(K), cudaMemcpyDeviceToHost, stq), "cudaMemcpyAsync") ||
                cu(cudaStreamSynchronize(stq), "cudaStreamSynchronize")) return DNALDPC_ERR_CUDA;
            if (K == 0) break;
            Session s2 = ss[(size_t)k];
            s2.in.param = params[rd];
            s2.max_frames = K;
            ListSource ls(list, K, stq);
            r = e.run_session(s2, ls, stq);  // inputs: resident rows list[0..K); outputs: compact rows 0..K
            if (r) { er = e.error(); return r; }
            add_stats(est[(size_t)k], e.stats);
            est[(size_t)k].kernel_launches += 2;
            // a frame keeps the result of its last round: compact read-back, scattered by frame on the host
            const size_t Ks = (size_t)K;
            list_h.resize(Ks); t_ok.resize(Ks); t_it.resize(Ks);
            bool bad = cu(cudaMemcpyAsync(list_h.data(), list, Ks * 4, cudaMemcpyDeviceToHost, stq), "cudaMemcpyAsync") ||
                       cu(cudaMemcpyAsync(t_ok.data(), e.s_ok_, Ks, cudaMemcpyDeviceToHost, stq), "cudaMemcpyAsync") ||
                       cu(cudaMemcpyAsync(t_it.data(), e.s_iters_, Ks * 4, cudaMemcpyDeviceToHost, stq), "cudaMemcpyAsync");
            if (!bad && out.bits) { t_bits.resize(Ks * wpf); bad = cu(cudaMemcpyAsync(t_bits.data(), e.s_bits_, Ks * wpf * 4, cudaMemcpyDeviceToHost, stq), "cudaMemcpyAsync"); }
            if (!bad) bad = cu(cudaStreamSynchronize(stq), "cudaStreamSynchronize");
            if (bad) return DNALDPC_ERR_CUDA;
            for (size_t i = 0; i < Ks; i++) {
                const size_t f = (size_t)row_frame[(size_t)list_h[i]];
                ok_host[f] = t_ok[i];
                if (out.iters) out.iters[f] = t_it[i];
                if (rounds) rounds[f] = rd;
                if (out.bits) memcpy(out.bits + f * wpf, &t_bits[i * wpf], wpf * 4);
            }
            auto scatter = [&](void *dev_src, size_t row_bytes, char *dst) {  // large per-frame outputs, one array at a time
                t_u8.resize(Ks * row_bytes);
                if (cu(cudaMemcpyAsync(t_u8.data(), dev_src, Ks * row_bytes, cudaMemcpyDeviceToHost, stq), "cudaMemcpyAsync") ||
                    cu(cudaStreamSynchronize(stq), "cudaStreamSynchronize")) return true;
                for (size_t i = 0; i < Ks; i++) memcpy(dst + (size_t)row_frame[(size_t)list_h[i]] * row_bytes, &t_u8[i * row_bytes], row_bytes);
                return false;
            };
            if (out.dblk && scatter(e.s_dblk_, (size_t)N, (char *)out.dblk)) return DNALDPC_ERR_CUDA;
            if (out.pchk && scatter(e.s_pchk_, (size_t)M, (char *)out.pchk)) return DNALDPC_ERR_CUDA;
            if (out.posterior && scatter(e.s_post_, (size_t)N * 8, (char *)out.posterior)) return DNALDPC_ERR_CUDA;
            prev = list;
            n_prev = K;
        }
        return DNALDPC_OK;
    });
    for (int k = 0; k < nd; k++) {
        if (!errs[(size_t)k].empty() && err.empty()) err = errs[(size_t)k];
        add_stats(st, est[(size_t)k]);
    }
    if (out.is_codeword) memcpy(out.is_codeword, ok_host.data(), (size_t)F);
    return rc;
}

}  // namespace dnaldpc
