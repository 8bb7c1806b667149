// sources.h - how batches reach the engines: the frame sources behind dnaldpc_decode_batch (host buffers, staged through
// rings while the engines decode), dnaldpc_decode_batch_device (HBM-resident buffers, read by every GPU of the decoder
// over NVLink peer access) and dnaldpc_redecode_sweep_ex (re-decoding rounds over the inputs left in HBM).
// Internal C++ interface between capi.cu and engine.cu.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "engine.h"

namespace dnaldpc {

typedef std::vector<std::unique_ptr<Engine>> EngineSet;

size_t packed_stride(int kind, int N);  // bytes of one tightly packed input frame, 0 = unknown kind

// HOST buffers. Frames are cut into chunks that the engines pull from one shared counter (a GPU that finishes early
// takes more of the batch); per engine an input thread copies chunks into a device ring and publishes them to the
// engine's frame queue, an output thread copies finished chunks back, and the engine decodes meanwhile.
int decode_host_batch(EngineSet &eng, const dnaldpc_input &in, int64_t F, int max_iter, const dnaldpc_output &out,
                      dnaldpc_stats &st, std::string &err);

// DEVICE buffers (resident on one GPU; the other engines read and write them through peer access). Engine 0 runs on
// `stream`. Blocking for the host; complete on `stream` when it returns.
int decode_device_batch(EngineSet &eng, const dnaldpc_input &in, int64_t F, int max_iter, const dnaldpc_output &out,
                        cudaStream_t stream, dnaldpc_stats &st, std::string &err);

// Re-decoding sweep (ex_decoder/decoder.py:594-664): round r decodes with in.param = params[r] the frames whose syndrome
// is still non-zero. HOST buffers; inputs are uploaded once and stay in HBM, later rounds run over device-side row lists.
int redecode_sweep_batch(EngineSet &eng, const dnaldpc_input &in, int64_t F, int max_iter, const double *params, int n_params,
                         const dnaldpc_output &out, int32_t *rounds, dnaldpc_stats &st, std::string &err);

}  // namespace dnaldpc
