// dev_math.cuh -- device equivalents of the reference's rml SIMD math (src/rml/rmlv/rmlv_mvec4.hxx,
// rmlv_soa.hxx, rmlm_soa.hxx) at one-lane granularity.
//
// Everything here is IEEE fp32 round-to-nearest evaluated in the order the reference writes it.
// The translation unit is compiled with -fmad=false so nvcc never contracts a*b+c into FFMA
// (the reference is SSE: separate mulps/addps), with the default -prec-div / -prec-sqrt (IEEE
// division and square root, like divps / sqrtps) and without -ftz.
//
// The two operations that are NOT IEEE on the CPU, rcpps and rsqrtps, are reproduced from
// tables harvested from the host CPU (host_luts.cpp):
//   rcpps(x)   depends on sign, exponent and the top 11 mantissa bits  -> 2048 entries
//   rsqrtps(x) depends on exponent parity and the top 10 mantissa bits -> 2 x 1024 entries
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rsr {

struct ApproxLuts {
	uint32_t rcp[2048];
	uint32_t rsqrt[2048];
	alignas(16) uint16_t rcp16[2048];   // rcp at 16 bits per entry (rcp_entry16 below), what the tile kernel stages in shared memory
};

__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }

// _mm_max_ps / _mm_min_ps: second operand wins on NaN / equality (rmlv_mvec4.hxx:543-544)
__device__ __forceinline__ float sse_max(float a, float b) { return (a > b) ? a : b; }
__device__ __forceinline__ float sse_min(float a, float b) { return (a < b) ? a : b; }

// _mm_cvttps_epi32: truncate; out of range and NaN give the "integer indefinite" 0x80000000
__device__ __forceinline__ int cvtt(float f) {
	return (fabsf(f) < 2147483648.0f) ? __float2int_rz(f) : static_cast<int>(0x80000000u); }

// _mm_cvtepi32_ps
__device__ __forceinline__ float itof(int i) { return __int2float_rn(i); }

// rmlv::fract (rmlv_mvec4.hxx:578-585): a - float(trunc(a))
__device__ __forceinline__ float fract_sse(float a) { return a - itof(cvtt(a)); }

// rcpps, bit exact against the harvested table
__device__ __forceinline__ float rcp_intel(float x, const uint32_t* __restrict__ lut) {
	const uint32_t b = f2u(x);
	const uint32_t sign = b & 0x80000000u;
	const int E = static_cast<int>((b >> 23) & 0xffu);
	const uint32_t m = b & 0x007fffffu;
	if (E == 0) { return u2f(sign | 0x7f800000u); }              // zero and denormals -> +-inf
	if (E == 255) { return m ? u2f(b | 0x00400000u) : u2f(sign); }  // nan -> qnan, inf -> +-0
	const uint32_t r = lut[m >> 12];
	const int re = static_cast<int>(r >> 23) + (127 - E);
	if (re <= 0) { return u2f(sign); }                            // would be denormal -> +-0
	return u2f(sign | (static_cast<uint32_t>(re) << 23) | (r & 0x007fffffu)); }

// rsqrtps, bit exact against the harvested table
__device__ __forceinline__ float rsqrt_intel(float x, const uint32_t* __restrict__ lut) {
	const uint32_t b = f2u(x);
	const uint32_t sign = b & 0x80000000u;
	const int E = static_cast<int>((b >> 23) & 0xffu);
	const uint32_t m = b & 0x007fffffu;
	if (E == 255 && m) { return u2f(b | 0x00400000u); }           // nan
	if (E == 0) { return u2f(sign | 0x7f800000u); }               // +-0 and denormals -> +-inf
	if (sign) { return u2f(0xffc00000u); }                        // negative -> real indefinite
	if (E == 255) { return 0.0f; }                                // +inf -> +0
	const int e = E - 127;
	const int p = e & 1;
	const int k = (e - p) >> 1;
	const uint32_t r = lut[(p << 10) | (m >> 13)];
	return u2f(r - (static_cast<uint32_t>(k) << 23)); }

// ---- packed fp32x2 (sm_100a FMUL2 / FFMA2) ------------------------------------------------------
// Two IEEE fp32 lanes per instruction; each lane rounds exactly like the scalar op, so the SSE
// order of operations is kept.  ptxas contracts mul.rn.f32x2 feeding add.rn.f32x2 into one FFMA2
// even under --fmad false (tools/ubench/f32x2.cu), so sums are written fma(a, ONE, b) and
// differences fma(b, -ONE, a) with ONE read from __constant__ memory, which the compiler cannot
// fold: round(a*1 + b) == round(a + b) bit for bit (0 mismatches over 2^23 random pairs, same file).
__constant__ float2 kOne2 = {1.0f, 1.0f};
__constant__ float2 kNegOne2 = {-1.0f, -1.0f};

struct f2 { unsigned long long v; };

__device__ __forceinline__ f2 mk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 dup2(float a) { return mk2(a, a); }
__device__ __forceinline__ float lo2(f2 a) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return x; }
__device__ __forceinline__ float hi2(f2 a) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return y; }
__device__ __forceinline__ f2 one2() { return mk2(kOne2.x, kOne2.y); }
__device__ __forceinline__ f2 negone2() { return mk2(kNegOne2.x, kNegOne2.y); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, float s) { return mul2(a, dup2(s)); }   // SASS: FMUL2 Rd, Ra.F32x2, Rs.F32
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return fma2(a, one2(), b); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return fma2(b, negone2(), a); }

// rcpps for the common case 1 <= exponent <= 252 (result is a normal number): two table lanes at a
// time; `ok` is cleared when a lane needs the general routine
__device__ __forceinline__ float rcp_fast(float x, const uint32_t* __restrict__ lut, bool& ok) {
	const uint32_t b = f2u(x);
	const uint32_t e = b & 0x7f800000u;
	ok = ok && ((e - 0x00800000u) < 0x7e000000u);
	const uint32_t r = lut[(b >> 12) & 0x7ffu];
	return u2f(((r + 0x3f800000u) - e) | (b & 0x80000000u)); }

// The same table at 16 bits per entry, as the tile kernel keeps it in shared memory (4 KB instead of 8): every
// rcpps result for a mantissa in [1, 2) lies in (0.5, 1] -- exponent field 126 -- and carries at most 16
// significant mantissa bits (Intel: 12), so entry = (bits >> 7) & 0xffff loses nothing.  rcp_table_fits16
// (host, rsrcu.cu) checks that this holds for the harvested table before a context is created.
__device__ __forceinline__ uint32_t rcp_entry16(uint32_t bits) { return (bits >> 7) & 0xffffu; }
__device__ __forceinline__ uint32_t rcp_expand16(uint32_t e16) { return 0x3f000000u | (e16 << 7); }

__device__ __forceinline__ float rcp_fast16(float x, const uint16_t* __restrict__ lut, bool& ok) {
	const uint32_t b = f2u(x);
	const uint32_t e = b & 0x7f800000u;
	ok = ok && ((e - 0x00800000u) < 0x7e000000u);
	const uint32_t r = lut[(b >> 12) & 0x7ffu];
	return u2f((((r << 7) + 0x7e800000u) - e) | (b & 0x80000000u)); }   // (0x3f000000 | r << 7) + 0x3f800000 - e

__device__ __forceinline__ float rcp_intel16(float x, const uint16_t* __restrict__ lut) {
	const uint32_t b = f2u(x);
	const uint32_t sign = b & 0x80000000u;
	const int E = static_cast<int>((b >> 23) & 0xffu);
	const uint32_t m = b & 0x007fffffu;
	if (E == 0) { return u2f(sign | 0x7f800000u); }
	if (E == 255) { return m ? u2f(b | 0x00400000u) : u2f(sign); }
	const uint32_t r = rcp_expand16(lut[m >> 12]);
	const int re = static_cast<int>(r >> 23) + (127 - E);
	if (re <= 0) { return u2f(sign); }
	return u2f(sign | (static_cast<uint32_t>(re) << 23) | (r & 0x007fffffu)); }

// rmlv::oneover (rmlv_mvec4.hxx:630-650): rcpps + one Newton-Raphson step
__device__ __forceinline__ float oneover(float a, const uint32_t* __restrict__ rcpLut) {
	const float r = rcp_intel(a, rcpLut);
	const float muls = a * (r * r);
	return (r + r) - muls; }

// qmat4 * qfloat4 (rmlm_soa.hxx:58-64), column-major m
__device__ __forceinline__ void mat4_mul(const float* __restrict__ m, float x, float y, float z, float w,
                                         float& ox, float& oy, float& oz, float& ow) {
	ox = ((m[0] * x + m[4] * y) + m[8] * z) + m[12] * w;
	oy = ((m[1] * x + m[5] * y) + m[9] * z) + m[13] * w;
	oz = ((m[2] * x + m[6] * y) + m[10] * z) + m[14] * w;
	ow = ((m[3] * x + m[7] * y) + m[11] * z) + m[15] * w; }

// mul_w0(qmat4, qfloat3) (rmlm_soa.hxx:42-47)
__device__ __forceinline__ void mat4_mul_w0(const float* __restrict__ m, float x, float y, float z,
                                            float& ox, float& oy, float& oz) {
	ox = (m[0] * x + m[4] * y) + m[8] * z;
	oy = (m[1] * x + m[5] * y) + m[9] * z;
	oz = (m[2] * x + m[6] * y) + m[10] * z; }

}  // namespace rsr
