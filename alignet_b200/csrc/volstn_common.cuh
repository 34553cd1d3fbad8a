// volstn_common.cuh -- shared device/host helpers for the sm_100a kernels.
//
// The arithmetic contract (reference stn3dtorch/generic/
// BilinearSamplerVolumetric.c:507-517, 542-550, 566-573, 730-822) is reproduced with explicit
// round-to-nearest intrinsics so that nvcc can never contract a multiply into a following add:
// outputs and gradGrid are then bit-identical to the reference's CPU build (-ffp-contract=off).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/volstn_b200.h"

namespace volstn {

// ---------------------------------------------------------------------------------------------
// host-side bookkeeping
// ---------------------------------------------------------------------------------------------
extern std::atomic<uint64_t> g_launches;           // abi.cu
void set_last_cuda_error(cudaError_t e);           // abi.cu (thread local)

inline int cuda_fail(cudaError_t e) {
  set_last_cuda_error(e);
  // A failing runtime call also leaves its code in the runtime's per-thread "last error"; unless it is taken out here the
  // NEXT launch's after_launch() would report this call's failure as its own (found by the test that unpins a buffer twice).
  (void)cudaGetLastError();
  return VOLSTN_ERR_CUDA;
}
#define VOLSTN_CUDA_TRY(expr)                                   \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) return ::volstn::cuda_fail(_e);      \
  } while (0)

// call after every <<<>>>: counts the launch and converts a launch error into a status
inline int after_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e);
  return VOLSTN_OK;
}

constexpr int kSMs = 148;  // B200

// ---------------------------------------------------------------------------------------------
// exact (never contracted) arithmetic for float and double
// ---------------------------------------------------------------------------------------------
template <typename real> struct Ex;

template <> struct Ex<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float from_int(int i) { return __int2float_rn(i); }
  // (int)floor(c) with the reference's x86-64 semantics: cvttsd2si returns INT_MIN for NaN and
  // for anything outside [-2^31, 2^31).  cvt.rmi.s32.f32 saturates instead (NaN -> 0,
  // +big -> INT_MAX), so those two cases are patched; below -2^31 it already yields INT_MIN.
  static __device__ __forceinline__ int floor_to_int(float c) {
    int i = __float2int_rd(c);
    return (c < 2147483648.0f) ? i : INT32_MIN;
  }
};

template <> struct Ex<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double from_int(int i) { return __int2double_rn(i); }
  static __device__ __forceinline__ int floor_to_int(double c) {
    int i = __double2int_rd(c);
    return (c < 2147483648.0) ? i : INT32_MIN;
  }
};

// One axis of the coordinate transform (...c:507-509):
//   coord = ((g + 1) * (size - 1)) / 2 ;  i0 = floor(coord) ;  w = 1 - (coord - i0)
// `sm1` is (real)(size - 1); the division by two is exact, so * 0.5 is the same number.
template <typename real>
__device__ __forceinline__ real axis_coord(real g, real sm1) {
  using E = Ex<real>;
  return E::mul(E::mul(E::add(g, real(1)), sm1), real(0.5));
}
template <typename real>
__device__ __forceinline__ void axis_finish(real c, int &i0, real &w) {
  using E = Ex<real>;
  i0 = E::floor_to_int(c);
  w = E::sub(real(1), E::sub(c, E::from_int(i0)));
}
template <typename real>
__device__ __forceinline__ void axis_setup(real g, real sm1, int &i0, real &w) {
  axis_finish<real>(axis_coord<real>(g, sm1), i0, w);
}
// floor for 0 <= c < 2^23 (float) without conversion instructions: fl_down(c + 2^23) = 2^23 + floor(c)
// exactly, its low mantissa bits are floor(c), and subtracting 2^23 again is exact.
template <typename real>
__device__ __forceinline__ void axis_finish_fast(real c, int &i0, real &w) {
  using E = Ex<real>;
  if constexpr (sizeof(real) == 4) {
    const float r = __fadd_rd(c, 8388608.0f);
    i0 = __float_as_int(r) - 0x4B000000;
    w = E::sub(1.0f, E::sub(c, __fsub_rn(r, 8388608.0f)));
  } else {
    axis_finish<real>(c, i0, w);
  }
}

// `0 <= i && i <= n-1` for i0 and i0+1 without signed-overflow traps: i0 may be INT_MIN
// (then both tests are false, as in the reference's int arithmetic).
__device__ __forceinline__ bool in_range(int i, int n) { return (unsigned)i < (unsigned)n; }
__device__ __forceinline__ int wrap_add(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int wrap_mul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }

// corner visiting orders (see oracle/volstn_oracle.c header for the derivation)
//   corner k:  x = x0 + (k&1), y = y0 + ((k>>1)&1), z = z0 + ((k>>2)&1)
__device__ __forceinline__ constexpr int order_xyz(int j) { return j; }                          // 0..7
__device__ __forceinline__ constexpr int order_zxy(int j) { return ((j & 1) << 2) | (j >> 1); }  // 0,4,1,5,2,6,3,7

// ---------------------------------------------------------------------------------------------
// memory access helpers (sm_100a PTX)
// ---------------------------------------------------------------------------------------------
// Streaming 128-bit read that does not allocate in L1: the grid / gradOut streams are touched once,
// L1 is kept for the volume gather.
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
  float4 r;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream_f1(const float *p) {
  float r;
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ld_stream_f2(const float2 *p) {
  float2 r;
  asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
// streaming (evict-first) stores for write-once outputs.  Compiler intrinsics, NOT asm with a "memory"
// clobber: a clobber is a compiler barrier, and the gathers of a thread's second voxel must be free to
// move above the first voxel's store (these outputs are never read back inside a kernel).
__device__ __forceinline__ void st_stream_f4(float4 *p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream_f2(float2 *p, float2 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream_f1(float *p, float v) { __stcs(p, v); }

// fire-and-forget reductions into gradVol (REDG.E.ADD.F32 / .F32x2 / .F32x4 on sm_100a).  volatile (they
// have no outputs) but no "memory" clobber, for the same reason: gradVol is write-only inside a kernel.
__device__ __forceinline__ void red_add(float *p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v));
}
__device__ __forceinline__ void red_add2(float *p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b));
}
__device__ __forceinline__ void red_add4(float *p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d));
}
__device__ __forceinline__ void red_add(double *p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v));
}

}  // namespace volstn
