// host_emu_shim.h -- TEST INFRASTRUCTURE ONLY.
// Lets g++ compile the kernel bodies of network-aware-bwa_b200/csrc/kernels.cuh as plain
// functions, so the `-m "not gpu"` tests can check the kernels' LOGIC (traversal order,
// stack emulation, index re-layout) against the reference on a box without a GPU.  One
// "thread" runs at a time (blockIdx/threadIdx are globals set by the launcher loop).
// Nothing in the product includes this file: libbwagpu.so is built by nvcc without
// BWAGPU_HOST_EMU and has no CPU path.
#pragma once
#include <stdint.h>
#include <string.h>

#undef BWAGPU_LDG256
#define BWAGPU_LDG256 0

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__

struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }
struct emu_dim3 { unsigned x, y, z; };
static emu_dim3 blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, threadIdx = {0, 0, 0}, gridDim = {1, 1, 1};

template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
template <typename T, typename U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T, typename U> static inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <typename T, typename U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <typename T> static inline T atomicCAS(T *p, T cmp, T val) { T o = *p; if (o == cmp) *p = val; return o; }
static inline void __threadfence() {}
template <typename T> static inline void __stcs(T *p, T v) { *p = v; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
#ifndef BWAGPU_WARP_EMU // warp_emu.cpp runs 32 real lanes and supplies real collectives
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline bool __all_sync(unsigned, bool p) { return p; }
static inline void __syncwarp() {}
#endif
