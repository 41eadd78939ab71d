/* TEST INFRASTRUCTURE ONLY.  Minimal host stand-ins for the CUDA device vocabulary used by the reference's
 * my_utils/standard_rasterize_cuda/standard_rasterize_cuda_kernel.cu:1-233 (the two __global__ kernels and their
 * __device__ helpers) so that the UNMODIFIED kernel text can be compiled by g++ and executed thread by thread on
 * the CPU (oracle/ref_raster_driver.cpp).  Nothing here is a restatement of the reference's algorithm. */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#ifndef __restrict__
#define __restrict__
#endif

struct cpu_dim3 { int x = 0, y = 0, z = 0; };
static thread_local cpu_dim3 blockIdx, blockDim, threadIdx;

using std::ceil;
using std::floor;
using std::max;
using std::min;

static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
/* single-threaded emulation: compare-and-swap is trivially atomic */
static inline int atomicCAS(int* a, int cmp, int val) { int old = *a; if (old == cmp) *a = val; return old; }
static inline unsigned long long atomicCAS(unsigned long long* a, unsigned long long cmp, unsigned long long val) {
    unsigned long long old = *a; if (old == cmp) *a = val; return old;
}
