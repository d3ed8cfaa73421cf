// caller_slot.hpp -- one caller slot number per calling thread, shared by every adaptor of this directory.
//
// ffgpu_eval's last argument selects one of the library's caller slots (a stream and staging buffers each,
// include/ffgpu.h).  The reference calls the black box and the parser concurrently from the Reconstructor's pool
// threads (Reconstructor.hpp:2981-2983): each of those threads gets a number of its own in order of first
// appearance, whichever adaptor and whichever bunch type (FFInt, FFIntVec<N>) it calls first -- an inline function
// with a function-local counter is one object in the whole program, a counter inside a member template would be
// one per instantiation.
#pragma once

#include <atomic>
#include <cstdint>

namespace firefly {

inline uint32_t gpu_caller_slot() {
  static std::atomic<uint32_t> next{0};
  thread_local const uint32_t mine = next.fetch_add(1, std::memory_order_relaxed);
  return mine;
}

}  // namespace firefly
