// One source, two builds: the integer / bit-exact kernels (clean_core.cuh, safebox_core.cuh) are written against these
// few macros so that the same code compiles for the device (nvcc) and, FOR THE UNIT TESTS ONLY, as a sequential host
// emulation (-DMTB_HOST_EMUL, tests/host_emul: one "thread", barriers are no-ops, atomics are plain updates).  The
// emulation is never linked into libmtb200.so; the product has no CPU path.
#pragma once
#include <stdint.h>

#ifdef MTB_HOST_EMUL
#include <math.h>
#include <string.h>
#define MTB_HD inline
#define MTB_SYNC() ((void)0)
#define MTB_TID 0
#define MTB_NTHR 1
template <typename T>
inline T mtb_atomic_add(T* p, T v) { T o = *p; *p = o + v; return o; }
inline int mtb_atomic_or(int* p, int v) { int o = *p; *p = o | v; return o; }
inline int mtb_atomic_min(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
inline int mtb_atomic_max(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
inline int mtb_popc(uint32_t v) { return __builtin_popcount(v); }
inline int mtb_ffs(uint32_t v) { return __builtin_ffs((int)v); }
inline unsigned long long mtb_atomic_max64(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
inline unsigned long long mtb_atomic_min64(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
inline float mtb_fsqrt(float a) { return sqrtf(a); }
inline double mtb_dsqrt(double a) { return sqrt(a); }
inline double mtb_rint(double a) { return nearbyint(a); }   // round half to even, like Python's round()
inline unsigned long long mtb_double_bits(double a) { unsigned long long u; memcpy(&u, &a, 8); return u; }
inline double mtb_dmul(double a, double b) { return a * b; }
inline double mtb_ddiv(double a, double b) { return a / b; }
inline double mtb_dadd(double a, double b) { return a + b; }
inline double mtb_dsub(double a, double b) { return a - b; }
#else
#define MTB_HD __device__ __forceinline__
#define MTB_SYNC() __syncthreads()
#define MTB_TID (static_cast<int>(threadIdx.x))
#define MTB_NTHR (static_cast<int>(blockDim.x))
template <typename T>
__device__ __forceinline__ T mtb_atomic_add(T* p, T v) { return atomicAdd(p, v); }
__device__ __forceinline__ int mtb_atomic_or(int* p, int v) { return atomicOr(p, v); }
__device__ __forceinline__ int mtb_atomic_min(int* p, int v) { return atomicMin(p, v); }
__device__ __forceinline__ int mtb_atomic_max(int* p, int v) { return atomicMax(p, v); }
__device__ __forceinline__ int mtb_popc(uint32_t v) { return __popc(v); }
__device__ __forceinline__ int mtb_ffs(uint32_t v) { return __ffs(static_cast<int>(v)); }
__device__ __forceinline__ unsigned long long mtb_atomic_max64(unsigned long long* p, unsigned long long v) { return atomicMax(p, v); }
__device__ __forceinline__ unsigned long long mtb_atomic_min64(unsigned long long* p, unsigned long long v) { return atomicMin(p, v); }
__device__ __forceinline__ float mtb_fsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double mtb_dsqrt(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ double mtb_rint(double a) { return rint(a); }
__device__ __forceinline__ unsigned long long mtb_double_bits(double a) { return static_cast<unsigned long long>(__double_as_longlong(a)); }
// explicit IEEE ops so nvcc never contracts them into FMAs (the reference's CPU arithmetic has none)
__device__ __forceinline__ double mtb_dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double mtb_ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double mtb_dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double mtb_dsub(double a, double b) { return __dadd_rn(a, -b); }
#endif
