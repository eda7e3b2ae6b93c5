// host_pool.h — host-side helper of the *_host entry points: u8 -> f32 widening of gathered frame stacks on a small persistent
// thread pool. The f32 [b][x][y][slot] tensors the reference's ToMultiDimArray produces (breakout_environment.rs:56-77) are 4x the
// bytes of the u8 frames they are made from; the device gathers u8 in that order, PCIe carries 1/4 of the bytes, and the widening
// (value = u8 as f32, exact) happens here, straight into the caller's buffer.
#pragma once
#include <cstddef>
#include <cstdint>

namespace qlc_host {
int pool_threads();                                                   // workers + the calling thread
void widen_u8_f32(const uint8_t* src, float* dst, size_t n);          // dst[i] = (float)src[i], all pool threads

// Streamed form: the gather kernel stores its u8 output straight into page-locked host memory (over PCIe, while it runs) and
// raises flags[i] = flag_value when piece i has landed (fence.sys before the flag). The pool threads take the pieces in order,
// wait for each one's flag and widen it into the caller's buffer, so the conversion of piece i overlaps the transfer of the
// pieces after it - no copy calls, no events. `still_running(ctx)` is polled by the calling thread while it waits: it returns
// false when the producer can no longer raise flags (the stream has finished or failed), which ends the waiting; pieces whose
// flag is still missing then are counted in the return value (0 = all widened).
struct StreamPiece { const uint8_t* src; float* dst; uint32_t n; };   // n elements; dst == nullptr: nothing to do for this flag
// `group` consecutive pieces (the CTAs of one stack) are handed to a thread at a time.
size_t widen_stream(const StreamPiece* pieces, size_t n_pieces, size_t group, const volatile uint32_t* flags, uint32_t flag_value,
                    bool (*still_running)(void*), void* ctx);
}
