// lob_ffi.cc -- XLA typed-FFI handlers that put liblobstep.so under jax.ffi.ffi_call, so that jit / vmap / scan of the
// reference's trainer compose around the CUDA step (gymnax_exchange/jaxrl/MARL/ippo_rnn_JAXMARL.py:571,616,661 call
// vmap(env.reset) / vmap(env.step); gymnax_exchange/jaxen/marl_env.py:764-804 is the body these handlers replace).
//
// Built by jaxmarl_hft_b200/ffi_stub.py:build_ffi() ONLY where `jax.ffi.include_dir()` exists (the XLA FFI headers ship
// with jaxlib; this image has neither, so here the file is covered by tests/test_ffi_tables.py through the offset tables
// alone):
//   g++ -O2 -shared -fPIC -std=c++17 -I$(python -c "import jax; print(jax.ffi.include_dir())") -I<repo>/include \
//       -I/usr/local/cuda/include lob_ffi.cc -L<csrc> -llobstep -Wl,-rpath,<csrc> -o liblob_ffi.so
//
// The handlers are TABLE-DRIVEN: the Python stub passes, for every operand and every result, the byte offset of the
// pointer field it fills inside LobStepBuffers / LobReplayBuffers (taken from the ctypes mirror abi.LobStepBuffers, so
// the two sides cannot drift apart); the handler stores the pointers and calls the C ABI.  It allocates nothing,
// enqueues on XLA's stream, never synchronises and keeps no state (re-entrant: one call per device under pmap).
#include <cstdint>
#include <cstring>

#include <cuda_runtime_api.h>

#include "xla/ffi/api/ffi.h"

#include "lobstep.h"

namespace ffi = xla::ffi;

template <typename Bufs>
static ffi::Error Fill(Bufs* b, ffi::Span<const int32_t> arg_off, ffi::Span<const int32_t> ret_off,
                       ffi::RemainingArgs args, ffi::RemainingRets rets) {
  if (arg_off.size() != args.size() || ret_off.size() != rets.size())
    return ffi::Error::InvalidArgument("operand / offset table size mismatch");
  std::memset(b, 0, sizeof(*b));
  for (size_t i = 0; i < args.size(); ++i) {
    auto buf = args.get<ffi::AnyBuffer>(i);
    if (!buf.has_value()) return ffi::Error::InvalidArgument("operand is not a buffer");
    if (arg_off[i] < 0 || arg_off[i] + sizeof(void*) > sizeof(*b)) return ffi::Error::InvalidArgument("bad operand offset");
    *reinterpret_cast<void**>(reinterpret_cast<char*>(b) + arg_off[i]) = buf->untyped_data();
  }
  // results alias the state operands (input_output_aliases: the step updates them in place) or are the pure outputs
  // (obs, reward, done, info) and the scratch workspace
  for (size_t i = 0; i < rets.size(); ++i) {
    auto buf = rets.get<ffi::AnyBuffer>(i);
    if (!buf.has_value()) return ffi::Error::InvalidArgument("result is not a buffer");
    if (ret_off[i] < 0 || ret_off[i] + sizeof(void*) > sizeof(*b)) return ffi::Error::InvalidArgument("bad result offset");
    *reinterpret_cast<void**>(reinterpret_cast<char*>(b) + ret_off[i]) = (*buf)->untyped_data();
  }
  return ffi::Error::Success();
}

// lob_step / lob_reset: MARLEnv.step marl_env.py:776-804, MARLEnv.reset marl_env.py:764 + reset_env :130-207
static ffi::Error LobStepImpl(cudaStream_t stream, ffi::Span<const uint8_t> cfg_bytes, int64_t batch, bool reset_only,
                              ffi::Span<const int32_t> arg_off, ffi::Span<const int32_t> ret_off,
                              ffi::RemainingArgs args, ffi::RemainingRets rets) {
  LobStepConfig cfg;
  if (cfg_bytes.size() != sizeof(cfg) || lob_sizeof_step_config() != (int64_t)sizeof(cfg) ||
      lob_sizeof_step_buffers() != (int64_t)sizeof(LobStepBuffers) || lob_abi_version() != LOB_ABI_VERSION)
    return ffi::Error::InvalidArgument("LobStepConfig / LobStepBuffers size mismatch (ABI version?)");
  std::memcpy(&cfg, cfg_bytes.data(), sizeof(cfg));
  LobStepBuffers b;
  if (auto e = Fill(&b, arg_off, ret_off, args, rets); e.failure()) return e;
  const int rc = reset_only ? lob_reset_launch(&cfg, &b, batch, stream) : lob_step_launch(&cfg, &b, batch, stream);
  return rc == LOB_OK ? ffi::Error::Success() : ffi::Error::InvalidArgument(lob_last_error());
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(LobStep, LobStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<ffi::Span<const uint8_t>>("cfg")
                                  .Attr<int64_t>("batch")
                                  .Attr<bool>("reset_only")
                                  .Attr<ffi::Span<const int32_t>>("arg_off")
                                  .Attr<ffi::Span<const int32_t>>("ret_off")
                                  .RemainingArgs()
                                  .RemainingRets());

// lob_replay: BaseLOBEnv.step_env base_env.py:189-216 / job.scan_through_entire_array JaxOrderBookArrays.py:736-756
static ffi::Error LobReplayImpl(cudaStream_t stream, ffi::Span<const uint8_t> cfg_bytes, int64_t n_books, int64_t n_msgs,
                                int64_t n_msgs_total, ffi::Span<const int32_t> arg_off, ffi::Span<const int32_t> ret_off,
                                ffi::RemainingArgs args, ffi::RemainingRets rets) {
  LobBookConfig cfg;
  if (cfg_bytes.size() != sizeof(cfg) || lob_sizeof_book_config() != (int64_t)sizeof(cfg))
    return ffi::Error::InvalidArgument("LobBookConfig size mismatch (ABI version?)");
  std::memcpy(&cfg, cfg_bytes.data(), sizeof(cfg));
  LobReplayBuffers b;
  if (auto e = Fill(&b, arg_off, ret_off, args, rets); e.failure()) return e;
  b.n_msgs = (int32_t)n_msgs;
  b.n_msgs_total = n_msgs_total;
  const int rc = lob_replay_launch(&cfg, &b, n_books, stream);
  return rc == LOB_OK ? ffi::Error::Success() : ffi::Error::InvalidArgument(lob_last_error());
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(LobReplay, LobReplayImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<ffi::Span<const uint8_t>>("cfg")
                                  .Attr<int64_t>("n_books")
                                  .Attr<int64_t>("n_msgs")
                                  .Attr<int64_t>("n_msgs_total")
                                  .Attr<ffi::Span<const int32_t>>("arg_off")
                                  .Attr<ffi::Span<const int32_t>>("ret_off")
                                  .RemainingArgs()
                                  .RemainingRets());
