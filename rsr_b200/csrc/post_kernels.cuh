// post_kernels.cuh -- canvas filters that follow a frame on the device: Kawase blur, the glow
// combine + true-colour conversion, the telemetry span bars.  All HBM-bound streaming kernels: one 128-bit load / store
// per thread and access, rows contiguous across a warp.  Included by rsrcu.cu after tile_kernel.cuh (srgb8 / linear8 /
// kSrgbTab4 live there); the translation unit is compiled with -fmad=false, so a*b+c stays mul + add like the
// reference's SSE code.
#pragma once

namespace rsr {

// ---------------------------------------------------------------------------------------------
// rglr::KawaseBlurFilter (src/rgl/rglr/rglr_kawase.cxx:22-81)
//   dst(x, y) = (0 + box(x-d-1, y-d-1) + box(x+d, y-d-1) + box(x-d-1, y+d) + box(x+d, y+d)) * (1/16)
//   box(x, y) = ((in(x, y) + in(x+1, y)) + in(x, y+1)) + in(x+1, y+1), every coordinate clamped to the canvas
// The reference runs an unclamped fast path in the interior; the taps are the same there, so one clamped path
// reproduces both.  Algorithmic bytes: 16 read + 16 written per pixel; everything else is on-chip reuse.
//
// A thread owns one column and walks `rows` rows down it (4 .. 16: enough threads to fill the GPU on a small canvas, little warm-up on a large one) (lanes = consecutive columns: every load and store of a
// warp is one contiguous 512-byte row segment).  The upper row of each of the four boxes at output row y is the lower
// row of the same box at row y - 1, so it stays in registers: 8 instead of 16 128-bit taps per pixel.  (One thread
// per pixel with all 16 taps was L1-bound: 79 % L1 throughput at 37 % of the HBM roofline, profiles/r2_post_kernels.txt.)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 f4_add(const float4 a, const float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_box(const float4 a, const float4 b, const float4 c, const float4 d) { return f4_add(f4_add(f4_add(a, b), c), d); }

__global__ void __launch_bounds__(128)
kawase_kernel(const float4* __restrict__ src, int srcStride, float4* __restrict__ dst, int dstStride, int width, int height, int d, int rows) {
	const int x = blockIdx.x * 128 + threadIdx.x, y0 = blockIdx.y * rows;
	if (x >= width) { return; }
	const int y1 = min(y0 + rows, height);
	// columns of the left (-d-1) and right (+d) boxes
	const int xl0 = min(max(x - d - 1, 0), width - 1), xl1 = min(max(x - d, 0), width - 1);
	const int xr0 = min(max(x + d, 0), width - 1), xr1 = min(max(x + d + 1, 0), width - 1);
	auto rowOf = [&](int r) { return src + static_cast<size_t>(min(max(r, 0), height - 1)) * srcStride; };
	const float4* rt = rowOf(y0 - d - 1);   // upper row of the two upper boxes
	const float4* rb = rowOf(y0 + d);       // upper row of the two lower boxes
	float4 tl0 = __ldg(rt + xl0), tl1 = __ldg(rt + xl1), tr0 = __ldg(rt + xr0), tr1 = __ldg(rt + xr1);
	float4 bl0 = __ldg(rb + xl0), bl1 = __ldg(rb + xl1), br0 = __ldg(rb + xr0), br1 = __ldg(rb + xr1);
	const float k16 = 1.0f / 16.0f;
	for (int y = y0; y < y1; ++y) {
		rt = rowOf(y - d);
		rb = rowOf(y + d + 1);
		const float4 ntl0 = __ldg(rt + xl0), ntl1 = __ldg(rt + xl1), ntr0 = __ldg(rt + xr0), ntr1 = __ldg(rt + xr1);
		const float4 nbl0 = __ldg(rb + xl0), nbl1 = __ldg(rb + xl1), nbr0 = __ldg(rb + xr0), nbr1 = __ldg(rb + xr1);
		float4 ax = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		ax = f4_add(ax, f4_box(tl0, tl1, ntl0, ntl1));   // (-d-1, -d-1)
		ax = f4_add(ax, f4_box(tr0, tr1, ntr0, ntr1));   // ( d,   -d-1)
		ax = f4_add(ax, f4_box(bl0, bl1, nbl0, nbl1));   // (-d-1,  d)
		ax = f4_add(ax, f4_box(br0, br1, nbr0, nbr1));   // ( d,    d)
		dst[static_cast<size_t>(y) * dstStride + x] = make_float4(ax.x * k16, ax.y * k16, ax.z * k16, ax.w * k16);
		tl0 = ntl0; tl1 = ntl1; tr0 = ntr0; tr1 = ntr1;
		bl0 = nbl0; bl1 = nbl1; br0 = nbr0; br1 = nbr1; } }

// ---------------------------------------------------------------------------------------------
// rglr::Texture::maybe_make_mipmap (src/rgl/rglr/rglr_texture.cxx:33-81): the mip chain of a power-of-two square
// RGBA32F texture, stacked under the base level (level L starts at row dim + dim/2 + ... + dim >> (L-1), its rows keep
// the base stride), each texel ((a + b) + c) + d of the 2x2 texels above it, divided by 4.  A level depends on the
// rounded values of the one before, so a CTA builds a little pyramid: from a B x B block of the source level
// (B = 2^k <= 32) it produces the next k levels, the first from global memory, the others from shared memory.
// A 1024^2 render target is two launches (levels 1-5, then 6-10 by one CTA).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mipmap_kernel(float4* __restrict__ tex, int stride, int dim, int srcLevel, int srcSize, int k) {
	__shared__ float4 pyr[256 + 64 + 16 + 4 + 1];
	const int B = 1 << k;
	const int t = threadIdx.x;
	auto rowStart = [&](int level) { int r = 0; for (int l = 0; l < level; ++l) { r += dim >> l; } return r; };
	const float4* src = tex + static_cast<size_t>(rowStart(srcLevel)) * stride;
	int prevOfs = 0, ofs = 0;
	for (int j = 1; j <= k; ++j) {
		const int n = B >> j;   // this level's edge inside the block
		float4* dst = tex + static_cast<size_t>(rowStart(srcLevel + j)) * stride;
		for (int i = t; i < n * n; i += 256) {
			const int x = i % n, y = i / n;
			float4 a, b, c, d;
			if (j == 1) {
				const float4* p = src + static_cast<size_t>(blockIdx.y * B + 2 * y) * stride + blockIdx.x * B + 2 * x;
				a = p[0]; b = p[1]; c = p[stride]; d = p[stride + 1]; }
			else {
				const float4* p = pyr + prevOfs + (2 * y) * (2 * n) + 2 * x;
				a = p[0]; b = p[1]; c = p[2 * n]; d = p[2 * n + 1]; }
			const float4 sum = f4_add(f4_add(f4_add(a, b), c), d);
			const float4 avg = make_float4(sum.x / 4.0f, sum.y / 4.0f, sum.z / 4.0f, sum.w / 4.0f);
			pyr[ofs + y * n + x] = avg;
			dst[static_cast<size_t>(blockIdx.y * n + y) * stride + blockIdx.x * n + x] = avg; }
		__syncthreads();
		prevOfs = ofs;
		ofs += n * n; } }

// ---------------------------------------------------------------------------------------------
// `$glow`: rglr::Filter<GlowShader, sRGB | LinearColor> (rglr_algorithm.hxx:107-144, node/glow.cxx:24-39)
//   out = (image + blur * 0.7) * 0.5, image = quad-swizzled RGBA32F canvas {r[4], g[4], b[4], a[4]} per 2x2 quad,
//   blur = ONE pixel of the linear canvas per quad, read at (x/2, y/2), its r / g / b broadcast over the quad.
// One thread converts two horizontally adjacent quads (the reference's `sub` loop) and writes two 16-byte rows.
// Algorithmic bytes per 4x2 pixels: 2 x 48 (image r,g,b planes) + 2 x 16 (blur) read, 32 written.
// (Measured and not taken: a warp loading eight quads as 32 contiguous plane vectors, lane = quad * 4 + plane, with the
// channels gathered by shuffle -- L1 throughput 62 -> 39 %, but every fourth lane converts an alpha plane nobody needs
// and the kernel turns issue-bound: 20.4 instead of 15.5 us, profiles/README.md.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
glow_kernel(const float4* __restrict__ image, int imageStrideQuads, const float4* __restrict__ blur, int blurStride,
            uint32_t* __restrict__ dst, int dstStride, int width, int height, int gamma) {
	__shared__ uint32_t srgbTab[104];
	if (threadIdx.x < 104) { srgbTab[threadIdx.x] = kSrgbTab4[threadIdx.x]; }
	__syncthreads();
	const int x = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4, y = (blockIdx.y * 8 + (threadIdx.x >> 5)) * 2;
	if (x >= width || y >= height) { return; }
	uint32_t row0[4], row1[4];
#pragma unroll
	for (int sub = 0; sub < 2; ++sub) {
		const size_t q = static_cast<size_t>(y >> 1) * imageStrideQuads + (x >> 1) + sub;
		const float4 r4 = __ldg(image + q * 4), g4 = __ldg(image + q * 4 + 1), b4 = __ldg(image + q * 4 + 2);
		const float4 bl = __ldg(blur + static_cast<size_t>(y >> 1) * blurStride + (x >> 1) + sub);
		const float br = bl.x * 0.700f, bg = bl.y * 0.700f, bb = bl.z * 0.700f;
		const float r[4] = {r4.x, r4.y, r4.z, r4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
		uint32_t out[4];
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			const float cr = (r[l] + br) * 0.5f, cg = (g[l] + bg) * 0.5f, cb = (b[l] + bb) * 0.5f;
			out[l] = gamma ? ((srgb8(cr, srgbTab) << 16) | (srgb8(cg, srgbTab) << 8) | srgb8(cb, srgbTab))
			               : ((linear8(cr) << 16) | (linear8(cg) << 8) | linear8(cb)); }
		row0[2 * sub] = out[0]; row0[2 * sub + 1] = out[1];
		row1[2 * sub] = out[2]; row1[2 * sub + 1] = out[3]; }
	uint32_t* p0 = dst + static_cast<size_t>(y) * dstStride + x;
	uint32_t* p1 = p0 + dstStride;
	if (((reinterpret_cast<uintptr_t>(dst) | (static_cast<uintptr_t>(dstStride) * 4u)) & 15u) == 0) {
		*reinterpret_cast<uint4*>(p0) = make_uint4(row0[0], row0[1], row0[2], row0[3]);
		*reinterpret_cast<uint4*>(p1) = make_uint4(row1[0], row1[1], row1[2], row1[3]); }
	else {
#pragma unroll
		for (int i = 0; i < 4; ++i) { p0[i] = row0[i]; p1[i] = row1[i]; } } }

// ---------------------------------------------------------------------------------------------
// render_jobsys (src/viewer/jobsys_vis.cxx:26-90): one lane of 8-pixel bars per worker, 2-pixel gap; a span is drawn
// from int(start * scale) to int(end * scale) inclusive, its colour picked from 16 by xor-folded bits of `raw`, its
// brightness falling from 1 by 1 / (right - left) per pixel (a running float subtraction: reproduced as one).
// One CTA per lane walks the lane's spans in order (later spans overwrite earlier ones where the integer ends touch).
// ---------------------------------------------------------------------------------------------
struct DevSpan { double start, end; uint32_t raw; int32_t lane; };   // = RsrSpan

__constant__ uint32_t kTaskColors[16] = {
	0xd7de48u, 0xda7bf5u, 0x4ccbdbu, 0xf67e77u, 0x6dd671u, 0x84a6e6u, 0xee913cu, 0xf084b4u,
	0x50d9a7u, 0x66de3eu, 0xd19bdfu, 0xcfa93cu, 0x9fc34fu, 0xf179d9u, 0xb4e532u, 0xebc630u };

__global__ void __launch_bounds__(256)
spans_kernel(uint32_t* __restrict__ canvas, int stride, int width, int height, int left, int top, float scale,
             const DevSpan* __restrict__ spans, int count) {
	const int thick = 8, gap = 2;
	const int lane = blockIdx.x;
	const int barTop = top + lane * (thick + gap);
	for (int si = 0; si < count; ++si) {
		const DevSpan sp = spans[si];
		if (sp.lane != lane) { continue; }
		const int sl = __double2int_rz(sp.start * static_cast<double>(scale)), sr = __double2int_rz(sp.end * static_cast<double>(scale));   // int(double * float)
		uint32_t bits = sp.raw, ax = 0;
		ax ^= bits & 0xffu; bits >>= 6; ax ^= bits & 0xffu; bits >>= 8; ax ^= bits & 0xffu; bits >>= 10; ax ^= bits & 0xffu;
		const uint32_t color = kTaskColors[ax & 15u];
		const float delta = 1.0f / static_cast<float>(sr - sl);
		for (int bx = sl + static_cast<int>(threadIdx.x); bx <= sr; bx += blockDim.x) {
			float bright = 1.0f;
			for (int k = sl; k < bx; ++k) { bright -= delta; }
			// tc_mul: uint8_t(float(channel) * bright), alpha 0
			const uint32_t r = static_cast<uint32_t>(__float2int_rz(static_cast<float>((color >> 16) & 0xffu) * bright)) & 0xffu;
			const uint32_t g = static_cast<uint32_t>(__float2int_rz(static_cast<float>((color >> 8) & 0xffu) * bright)) & 0xffu;
			const uint32_t b = static_cast<uint32_t>(__float2int_rz(static_cast<float>(color & 0xffu) * bright)) & 0xffu;
			const int px = left + bx;
			if (px >= 0 && px < width) {
				for (int by = 0; by < thick; ++by) {
					const int py = barTop + by;
					if (py >= 0 && py < height) { canvas[static_cast<size_t>(py) * stride + px] = (r << 16) | (g << 8) | b; } } } }
		__syncthreads(); } }

}  // namespace rsr
