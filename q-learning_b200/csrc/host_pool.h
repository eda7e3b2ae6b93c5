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
}
