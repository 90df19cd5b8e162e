// clshim.h -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Minimal OpenCL-C 1.1 -> C++17 shim so that the reference's own device code
// (/root/reference/kernel/kernel.cl) can be compiled by g++ and executed on the
// host as the parity oracle.  Nothing in this file is
// derived from the reference: it only supplies the OpenCL built-ins that the
// kernel text uses (vector types, convert_*, dot/length/sign, work-item ids,
// atom_min) with IEEE-754 single precision semantics, evaluation order fixed
// left-to-right so the CUDA kernels can reproduce it bit for bit.
#pragma once
#include <cmath>
#include <cstdint>
#include <algorithm>

#define __kernel
#define __global
#define __local static thread_local
#define CLK_GLOBAL_MEM_FENCE 0

struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct int3   { int x, y, z; };
struct int2   { int x, y; };

// swizzle helper: the build recipe rewrites `v.xyz` into `xyz(v)`
static inline float3 xyz(const float4 &v) { return float3{v.x, v.y, v.z}; }
static inline float3 xyz(const float3 &v) { return v; }

static inline float3 operator+(float3 a, float3 b) { return float3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline float3 operator-(float3 a, float3 b) { return float3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline float3 operator*(float3 a, float3 b) { return float3{a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline float3 operator*(float3 a, float s)  { return float3{a.x * s, a.y * s, a.z * s}; }
static inline float3 operator/(float3 a, float s)  { return float3{a.x / s, a.y / s, a.z / s}; }
static inline float3 &operator+=(float3 &a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }

static inline int3 operator>>(int3 a, int s) { return int3{a.x >> s, a.y >> s, a.z >> s}; }
static inline int3 operator<<(int3 a, int s) { return int3{a.x << s, a.y << s, a.z << s}; }
static inline int3 operator&(int3 a, int s)  { return int3{a.x & s, a.y & s, a.z & s}; }
static inline int3 operator+(int3 a, int s)  { return int3{a.x + s, a.y + s, a.z + s}; }

// C truncation toward zero (x86 cvttss2si; out-of-range -> INT_MIN)
static inline int3   convert_int3(float3 a)   { return int3{(int)a.x, (int)a.y, (int)a.z}; }
static inline float3 convert_float3(int3 a)   { return float3{(float)a.x, (float)a.y, (float)a.z}; }
static inline float  convert_float(int a)     { return (float)a; }
static inline float  convert_float(double a)  { return (float)a; }
static inline int    convert_int(float a)     { return (int)a; }

static inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float length(float3 a)        { return sqrtf(dot(a, a)); }
static inline float sign(float a)           { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }

using std::min;
using std::max;
using std::abs;
using std::fabs;
using std::pow;
using std::log2;

// work-item identity, set by the serial scheduler in ref_driver.cpp
struct ClItem { int gid[2]; int lid[2]; int lsz[2]; };
extern thread_local ClItem cl_item;
static inline int get_global_id(int d)  { return cl_item.gid[d]; }
static inline int get_local_id(int d)   { return cl_item.lid[d]; }
static inline int get_local_size(int d) { return cl_item.lsz[d]; }

// cl_khr_global_int32_extended_atomics: unsigned min, returns the old value
// (a real atomic: the timed CPU baseline runs raycast_proj work-group-parallel, as an OpenCL CPU runtime would; serially it
// is the same read-compare-write)
static inline unsigned int atom_min(unsigned int *p, unsigned int v)
{
    unsigned int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
