// host_luts.cpp -- harvest this CPU's rcpps / rsqrtps approximation tables.
//
// The reference's perspective divide (`pdiv` -> `oneover`, src/rgl/rglv/rglv_math.hxx:19-23,
// src/rml/rmlv/rmlv_mvec4.hxx:630-650) and `normalize` (src/rml/rmlv/rmlv_soa.hxx:243-248) start
// from the SSE approximation instructions, whose results are implementation defined.  To be bit
// exact with the reference running on the same host, the device code looks the seed up in tables
// read out of the host instruction here.  Model (checked exhaustively below, not assumed):
//   rcpps(x)   = f(sign, exponent, mantissa >> 12)           -> 2048 entries
//   rsqrtps(x) = f(exponent parity, mantissa >> 13)          -> 2 x 1024 entries
#include <cstdint>
#include <cstring>
#include <xmmintrin.h>

namespace rsr {

static inline uint32_t bitsOf(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float floatOf(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

static inline uint32_t hostRcp(uint32_t x) { return bitsOf(_mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(floatOf(x))))); }
static inline uint32_t hostRsqrt(uint32_t x) { return bitsOf(_mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(floatOf(x))))); }

void harvest_luts(uint32_t* rcp2048, uint32_t* rsqrt2048) {
	for (uint32_t i = 0; i < 2048; ++i) { rcp2048[i] = hostRcp(0x3f800000u | (i << 12)); }
	for (uint32_t i = 0; i < 1024; ++i) {
		rsqrt2048[i] = hostRsqrt(0x3f800000u | (i << 13));
		rsqrt2048[1024 + i] = hostRsqrt(0x40000000u | (i << 13)); } }

// software model, identical to rcp_intel / rsqrt_intel in dev_math.cuh
static inline uint32_t modelRcp(uint32_t b, const uint32_t* lut) {
	const uint32_t sign = b & 0x80000000u;
	const int E = static_cast<int>((b >> 23) & 0xffu);
	const uint32_t m = b & 0x007fffffu;
	if (E == 0) { return sign | 0x7f800000u; }
	if (E == 255) { return m ? (b | 0x00400000u) : sign; }
	const uint32_t r = lut[m >> 12];
	const int re = static_cast<int>(r >> 23) + (127 - E);
	if (re <= 0) { return sign; }
	return sign | (static_cast<uint32_t>(re) << 23) | (r & 0x007fffffu); }

static inline uint32_t modelRsqrt(uint32_t b, const uint32_t* lut) {
	const uint32_t sign = b & 0x80000000u;
	const int E = static_cast<int>((b >> 23) & 0xffu);
	const uint32_t m = b & 0x007fffffu;
	if (E == 255 && m) { return b | 0x00400000u; }
	if (E == 0) { return sign | 0x7f800000u; }
	if (sign) { return 0xffc00000u; }
	if (E == 255) { return 0u; }
	const int e = E - 127;
	const int p = e & 1;
	const int k = (e - p) >> 1;
	const uint32_t r = lut[(p << 10) | (m >> 13)];
	return r - (static_cast<uint32_t>(k) << 23); }

// returns the number of inputs on which the table model and the instruction disagree.
// All 2^23 mantissas at two exponents, plus every exponent at a stride of mantissas, both signs.
uint64_t verify_luts(const uint32_t* rcp2048, const uint32_t* rsqrt2048) {
	uint64_t bad = 0;
	for (uint32_t m = 0; m < (1u << 23); ++m) {
		const uint32_t a = 0x3f800000u | m, b = 0x40000000u | m;
		bad += (hostRcp(a) != modelRcp(a, rcp2048));
		bad += (hostRcp(b) != modelRcp(b, rcp2048));
		bad += (hostRsqrt(a) != modelRsqrt(a, rsqrt2048));
		bad += (hostRsqrt(b) != modelRsqrt(b, rsqrt2048)); }
	for (uint32_t E = 0; E < 256; ++E) {
		for (uint32_t m = 0; m < (1u << 23); m += 4099) {
			for (uint32_t s = 0; s < 2; ++s) {
				const uint32_t x = (s << 31) | (E << 23) | m;
				const uint32_t hr = hostRcp(x), mr = modelRcp(x, rcp2048);
				const uint32_t hq = hostRsqrt(x), mq = modelRsqrt(x, rsqrt2048);
				const bool nanR = ((hr & 0x7f800000u) == 0x7f800000u) && (hr & 0x7fffffu);
				const bool nanQ = ((hq & 0x7f800000u) == 0x7f800000u) && (hq & 0x7fffffu);
				const bool mnanR = ((mr & 0x7f800000u) == 0x7f800000u) && (mr & 0x7fffffu);
				const bool mnanQ = ((mq & 0x7f800000u) == 0x7f800000u) && (mq & 0x7fffffu);
				bad += (nanR || mnanR) ? (nanR != mnanR) : (hr != mr);
				bad += (nanQ || mnanQ) ? (nanQ != mnanQ) : (hq != mq); } } }
	return bad; }

}  // namespace rsr
