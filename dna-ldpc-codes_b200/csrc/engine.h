// engine.h - per-device decoder engine: owns the device copies of the H edge tables, the slot-interleaved message
// arrays and the host side of the slot scheduler (admit -> syndrome/loop control -> check pass -> bit pass per tick;
// early exit on zero syndrome / max_iter per frame, dec.cpp:594-599). Internal C++ interface behind include/dnaldpc.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/dnaldpc.h"
#include "../host/code.h"

namespace dnaldpc {

struct SchedArrays;

class Engine {
  public:
    Engine(const Code &code, int device, int precision, int wave_frames);
    ~Engine();
    bool ok() const { return err_.empty(); }
    const std::string &error() const { return err_; }
    int device() const { return device_; }

    // DEVICE pointers in `in` / `out`; asynchronous on `stream` except for the lagged progress polls.
    int decode_device(const dnaldpc_input &in, int64_t F, int max_iter, const dnaldpc_output &out, cudaStream_t stream);
    // HOST pointers; blocking.
    int decode_host(const dnaldpc_input &in, int64_t F, int max_iter, const dnaldpc_output &out);
    // Sliding-window BP for spatially-coupled codes (Run_SW_Decoder, dec.cpp:2092-2196); HOST pointers; blocking.
    int decode_window_host(const Code &code, const dnaldpc_window &w, const double *lratio, int64_t F, int max_iter,
                           const dnaldpc_output &out);
    int synth_bsc(const uint32_t *cw_bits, int n_cw, uint64_t seed, int64_t frame0, int64_t F, double eps,
                  uint32_t *out_bits, cudaStream_t stream);

    dnaldpc_stats stats{};
    bool profiling = false;
    int trace_ticks_ = 0;  // > 0: record in-pipeline events for that many ticks of the next batch (no synchronisation)
    int trace_result(double *row_ms, double *col_ms, double *sched_ms);

  private:
    template <typename T> int run(const dnaldpc_input &in_dev, int64_t F, int max_iter, const dnaldpc_output &out_dev, cudaStream_t st);
    template <typename T> int launch_row(int g0, int G, cudaStream_t st);
    template <typename T> int launch_col(int g0, int G, bool want_post, cudaStream_t st);
    template <typename T> int launch_harvest_setup(const dnaldpc_input &in, const dnaldpc_output &out, int g0, int G, cudaStream_t st);
    int launch_syndrome(const dnaldpc_output &out, int G, int max_iter, int consider_new, int fixed, unsigned *counter,
                        unsigned *finished, unsigned *rearm, int clear_fresh, int64_t F, cudaStream_t st);
    int ensure_slots(int groups, bool want_post);
    int ensure_frame_scratch(int64_t F);
    int fail(cudaError_t e, const char *what);
    int fail(const std::string &m, int rc);
    void *stage(void **buf, size_t *cap, size_t need);
    void fill_sched(SchedArrays &s) const;

    std::string err_;
    int device_ = 0, precision_ = 0, wave_frames_ = 4096, sm_count_ = 148;
    int M_ = 0, N_ = 0, E_ = 0, max_row_deg_ = 0, max_col_deg_ = 0;
    bool reg_rows_ = false, reg_cols_ = false;
    bool smem_attr_set_[2] = {false, false}, syn_attr_set_ = false;
    bool steady_ = false;  // last polled tick: all slots busy, no frame admitted
    bool minsum_ = false;  // current batch runs the LLR-domain min-sum rules instead of sum-product
    size_t esz_ = 8;
    // H edge tables (device)
    int32_t *d_row_ptr_ = nullptr, *d_col_idx_ = nullptr, *d_col_ptr_ = nullptr, *d_col_edge_ = nullptr;
    // slot state (device)
    int cap_groups_ = 0;
    void *d_msg_ = nullptr, *d_lratio_ = nullptr, *d_post_ = nullptr;
    uint32_t *d_decw_ = nullptr, *d_masks_ = nullptr;  // masks: 6 words per group (act, done, newf, fresh, harv, unsat)
    unsigned int *d_arrive_ = nullptr;
    int32_t *d_slot_ = nullptr;                          // 4 ints per slot (frame, iter, harv_frame, harv_iter)
    void *d_sw_lr_ = nullptr;                            // sliding-window mode: second message array (check -> bit)
    int sw_cap_groups_ = 0;
    int32_t *d_edge_row_ = nullptr;                      // sliding-window mode: check of every edge
    int32_t *d_mv_ = nullptr;                            // drain-tail compaction: src[S], dst[S], {count, K}
    // compact when busy slots <= 15/16 of the packed region: a move is cheap next to the ticks it shortens (measured
    // thresholds 30 / 50 / 70 / 85 / 92 / 97 %: 16 384 frames eps 0.008 1196 / 1110 / 1084 / 1037 / 1016 / 992 ms per batch)
    static constexpr int kCompactNum = 15, kCompactDen = 16;
    unsigned long long *d_next_ = nullptr;
    int32_t *d_iters_ = nullptr;                         // per-frame scratch when the caller does not want them
    uint8_t *d_ok_ = nullptr;
    int64_t cap_frames_ = 0;
    double *d_table_ = nullptr;                          // 256 doubles (BSC uses the first 2)
    // progress counters: a ring, polled kLag ticks behind the device
    static constexpr int kLag = 2, kRing = 16;
    unsigned int *d_counters_ = nullptr, *h_counters_ = nullptr;
    cudaEvent_t ev_[kRing] = {};
    cudaEvent_t prof_ev_[3] = {};
    static constexpr int kTrace = 64;
    cudaEvent_t trace_ev_[3 * kTrace] = {};
    int traced_ = 0;
    cudaStream_t own_stream_ = nullptr;
    // staging for the host-pointer path (device side)
    void *s_in_ = nullptr, *s_bits_ = nullptr, *s_dblk_ = nullptr, *s_post_ = nullptr, *s_pchk_ = nullptr;
    size_t c_in_ = 0, c_bits_ = 0, c_dblk_ = 0, c_post_ = 0, c_pchk_ = 0;
    std::vector<double> h_exp_;
};

int math_selftest(long long n, uint64_t seed, long long *mismatches, std::string &err);

}  // namespace dnaldpc
