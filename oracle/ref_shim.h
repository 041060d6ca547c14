/* TEST INFRASTRUCTURE -- portability shim, force-included (-include) when the unmodified
 * reference sources under /root/reference are compiled with g++ for the parity oracle.
 * The reference is an MSVC/clang-cl project (copts.bzl:1-38); this header supplies the few
 * MSVC-isms its hot path touches.  It contains no renderer logic.
 */
#pragma once
#include <cfloat>
#include <cmath>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <x86intrin.h>

/* rmlv_math.hxx:11 declares `constexpr auto M_PI`, glibc's <cmath> defines a macro */
#undef M_PI

/* ryg-srgb.h:126,183 */
#define __forceinline inline __attribute__((always_inline))

/* rglv_gpu_impl.hxx:56 */
static inline unsigned char _BitScanForward(unsigned long* idx, unsigned long mask) {
	if (mask == 0) { return 0; }
	*idx = static_cast<unsigned long>(__builtin_ctzl(mask));
	return 1; }

/* rclmt_jobsys.cxx:84,91 / rclma_framepool.cxx:32 */
static inline void* _aligned_malloc(size_t size, size_t align) {
	const size_t padded = (size + align - 1) / align * align;
	return aligned_alloc(align, padded); }
static inline void _aligned_free(void* p) { free(p); }

/* rglr_texture_sampler.cxx:157 calls an overload that does not exist, from a class template
 * (RGBA8888 sampler) that is never instantiated; g++ still wants a declaration. */
namespace rqdq { namespace rmlv {
struct mvec4i;
void load_interleaved_lut(const uint32_t*, mvec4i, mvec4i&);
}}

/* src/viewer/node/gllayer.cxx:144 calls std::cosf (MSVC has it; glibc's <cmath> only declares ::cosf) */
namespace std { using ::cosf; }

/* `$many` and `$particles` seed std::mt19937 from std::random_device (node/many.cxx:56, node/particles.cxx:50).  The
 * scene parity tests render one scene through two libraries (pure reference / drop-in) and need the same scene in
 * both, so those two translation units are compiled with -DRSR_FIXED_SEED: the seed is 1. */
#ifdef RSR_FIXED_SEED
namespace std {
struct rsr_fixed_random_device {
	using result_type = unsigned int;
	result_type operator()() { return 1u; }
	static constexpr result_type min() { return 0u; }
	static constexpr result_type max() { return 0xffffffffu; } };
}
#define random_device rsr_fixed_random_device
#endif
