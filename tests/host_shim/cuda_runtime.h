// Test infrastructure: just enough of the CUDA device vocabulary for g++ to compile the product's arithmetic headers
// (bayes.js_b200/csrc/amwg_math.cuh, amwg_ld.cuh) for the HOST, so that tests/test_device_math_on_host.py can hold the very text the
// GPU runs against the oracle without a GPU. Nothing in the product includes this file.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline int __double2hiint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)(b & 0xffffffffll); }
static inline double __hiloint2double(int hi, int lo) {
  const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint64_t)(uint32_t)lo;
  double x; std::memcpy(&x, &b, 8); return x;
}
static inline double __longlong_as_double(long long b) { double x; std::memcpy(&x, &b, 8); return x; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * (uint64_t)b) >> 32); }
