// programs.cuh -- the reference's shader programs (src/viewer/shaders.hxx, shaders_envmap.hxx,
// shaders_wireframe.hxx) instantiated as CUDA device functors.
//
// The reference's program concept (`rglv::BaseProgram`, src/rgl/rglv/rglv_gpu_shaders.hxx:21-95)
// is kept: each program has an `id`, an `earlyZ` flag, a vertex stage producing gl_Position plus
// a block of varyings, and a fragment stage working on one 2x2 quad (4 lanes) at a time.
// Varyings are a flat float[NV] per vertex; every reference program interpolates all of them
// with the perspective barycentrics BP (`Interpolants::Interpolate`), which the tile kernel does
// generically.
//
// Lane order inside a quad is the reference's: 0=(x,y) 1=(x+1,y) 2=(x,y+1) 3=(x+1,y+1).
#pragma once
#include "dev_math.cuh"

namespace rsr {

constexpr int kMaxVaryings = 16;

struct TexUnit {
	const float4* texels;   // device pointer
	uint32_t texelCount;    // for clamping out-of-range gathers (reference reads out of bounds)
	int width, height, stride;
	int kind;               // 0 = non-pow2 nearest, 1 = pow2 mip nearest, 2 = pow2 mip bilinear
	int power; };

struct DevState {
	float vm[16], pm[16], nm[16], vpm[16];     // rglv::Matrices (rglv_gpu_shaders.hxx:14-18)
	float uniforms[32];
	float DSx, DSy, DOx, DOy;                  // GPU::DSDO (rglv_gpu.hxx:265-270)
	float clearColor[4];
	float clearDepth;
	int programId;
	int cullingEnabled, cullFace;
	int scissorX0, scissorY0, scissorX1, scissorY1;   // GPU::ScissorRect (rglv_gpu.hxx:272-276)
	int depthTest, depthFunc, depthWrite, colorWrite, blend;
	int color0Type, depthType;
	const float* buffers[16];
	TexUnit tu[2];
	const float* tu3;
	int tu3dim; };

// SoA vertex attributes of one vertex (the reference's VertexInput structs)
struct VertexIn {
	float px, py, pz;      // slots 0-2
	float nx, ny, nz;      // slots 3-5
	float kx, ky, kz;      // slots 6-8
	float u, v;            // slots 9-10
	const float* imat; };  // slot 15: per-instance mat4 (column-major)

// what the pow2 samplers need of texture unit 0, staged per triangle in the tile kernel's shared memory (a state lookup in
// global memory per rasterised quad would sit on the critical path in front of the texel taps)
struct TexDesc {
	const float4* texels;
	uint32_t texelCount;
	int kind, power; };

struct FragIn {
	const DevState* st;
	TexDesc tex0;               // texture unit 0 of the triangle's draw
	const uint32_t* rcpLut;
	const uint32_t* rsqrtLut;
	float fragX[4], fragY[4];   // gl_FragCoord
	float depth[4];             // gl_FragDepth
	float BPx[4], BPy[4], BPz[4];
	float4* stage; };           // the warp's texel staging area when all 32 lanes shade together (direct rasteriser), else nullptr

#ifndef RSR_COOP_STAGE
#define RSR_COOP_STAGE 0
#endif
// texels of per-warp staging for the cooperative sampler (a 16x8-pixel region at one texel per pixel: <= 20 x 10); without
// it the per-warp scratch only holds the queued rasteriser's work items (160 x uint16 = 20 texels' worth)
constexpr int kStageTexels = RSR_COOP_STAGE ? 208 : 20;

// ---- texture units (src/rgl/rglr/rglr_texture_sampler.cxx) -------------------------------------

__device__ __forceinline__ float4 fetch_texel(const TexDesc& tu, int ofs) {
	// the reference would read out of bounds for wild coordinates; stay inside the allocation
	const uint32_t o = static_cast<uint32_t>(ofs);
	return __ldg(tu.texels + (o < tu.texelCount ? o : 0u)); }

// coarse per-quad level of detail (rglr_texture_sampler.cxx:25-42): lanes 1 - 0 only
__device__ __forceinline__ int level_of_detail(float u0, float u1, float v0, float v1) {
	const float dux = u1 - u0;
	const float dvx = v1 - v0;
	const float sqd = dux * dux + dvx * dvx;
	return (static_cast<int>(f2u(sqd)) - (127 << 23)) >> 24; }

// one bilinear blend (rglr_texture_sampler.cxx:262-286): ((t00*w00 + t10*w10) + t01*w01) + t11*w11, two channels per packed pair
__device__ __forceinline__ void blend_taps(const float4 p00, const float4 p10, const float4 p01, const float4 p11,
                                           const float a00, const float a10, const float a01, const float a11,
                                           float& r, float& g, float& b, float& a) {
	const f2 rg = add2(add2(add2(mul2(mk2(p00.x, p00.y), a00), mul2(mk2(p10.x, p10.y), a10)), mul2(mk2(p01.x, p01.y), a01)),
	                   mul2(mk2(p11.x, p11.y), a11));
	const f2 ba = add2(add2(add2(mul2(mk2(p00.z, p00.w), a00), mul2(mk2(p10.z, p10.w), a10)), mul2(mk2(p01.z, p01.w), a01)),
	                   mul2(mk2(p11.z, p11.w), a11));
	r = lo2(rg); g = hi2(rg); b = lo2(ba); a = hi2(ba); }

// Samples texture unit `tu` at the four pixels of a quad.  `mask`: the pixels whose result is used (the reference
// samples all four SSE lanes and discards; here unused pixels are skipped and return zeros).
//
// Two routes for the bilinear sampler.  GATHER: every tap is a 128-bit load from global memory (L1).  STAGED
// (f.stage != nullptr: the direct rasteriser, where the 32 lanes of a warp shade the quads of ONE triangle inside
// a 16x8-pixel region, all of them in this call together): when the quads agree on the level of detail, the taps of
// the whole region fall into a small box of texels -- 17 x 9 at one texel per pixel.  The warp loads that box once
// with coalesced 128-bit row loads into its own shared-memory staging area and takes the 16 taps of every quad
// from there: same texels, same weights, same arithmetic as the gather route (bit-identical), but each texel
// crosses the L1 tag stage once per warp instead of once per tap, and a tap address is two adds instead of
// wrap / row / clamp arithmetic.  Staged columns are split by texel-x parity (even columns first, then odd): the
// lanes of a quarter warp sit two pixels = two texels apart, so their taps then read consecutive 16-byte slots
// (no bank conflicts).  A box that does not fit (minification near the next level, anisotropy) or quads that
// disagree on the level take the gather route.
__device__ __forceinline__ void sample_quad(const FragIn& f, const float (&u)[4], const float (&v)[4],
                                            float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], const uint32_t mask) {
	const TexDesc& tu = f.tex0;
	if (tu.kind == 0) {
		// TextureUnitRGBAF32_NM_ONEMAP_WRAP_NEAREST (rglr_texture_sampler.cxx:289-311)
		const TexUnit& gtu = f.st->tu[0];   // (width / height / stride: only this sampler needs them)
		const float fw = itof(gtu.width), fh = itof(gtu.height);
		const float almostOne = u2f(0x3f7fffffu);
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			const int pxU = cvtt(fract_sse(u[l] + 100.0f) * fw);
			const int pxV = cvtt((almostOne - fract_sse(v[l] + 100.0f)) * fh);
			const float4 t = fetch_texel(tu, pxV * gtu.stride + pxU);
			r[l] = t.x; g[l] = t.y; b[l] = t.z; a[l] = t.w; }
		return; }

	const int POWER = tu.power;
	const float baseDim = itof(1 << POWER);
	int lod = level_of_detail(u[0] * baseDim, u[1] * baseDim, v[0] * baseDim, v[1] * baseDim);
	lod = min(max(lod, 0), POWER);

	if (tu.kind == 1) {
		// ..._P2_MIPMAP_WRAP_NEAREST (rglr_texture_sampler.cxx:59-109)
		const float levelDim = itof(1 << (POWER - lod));
		const int levelBeginRow = static_cast<int>((0xfffffffeu << (POWER - lod)) & ((1u << (POWER + 1)) - 1u));
		const float almostOne = u2f(0x3f7fffffu);
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			const int levelX = cvtt(fract_sse(u[l] + 10.0f) * levelDim);
			const int levelY = cvtt((almostOne - fract_sse(v[l] + 10.0f)) * levelDim);
			const int ofs = ((levelBeginRow + levelY) << POWER) + levelX;
			const float4 t = fetch_texel(tu, ofs);
			r[l] = t.x; g[l] = t.y; b[l] = t.z; a[l] = t.w; }
		return; }

	// ..._P2_MIPMAP_WRAP_LINEAR (rglr_texture_sampler.cxx:166-286).  Coordinates and weights are
	// computed for two pixels at a time (packed pairs); the four taps of one pixel are blended two
	// channels at a time (the 128-bit texel is two register pairs) with the weight broadcast.
	const int levelDimI = 1 << (POWER - lod);
	const float levelDim = itof(levelDimI);
	const uint32_t lastTexel = tu.texelCount - 1u;
	const f2 one = one2();
	const f2 negHalf = dup2(-0.5f);
	int tx0[4], ty0[4];
	f2 w00[2], w10[2], w01[2], w11[2];
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		const f2 levelX = mul2(mk2(u[2 * h], u[2 * h + 1]), levelDim);
		const f2 levelY = mul2(mk2(v[2 * h], v[2 * h + 1]), levelDim);
		const f2 sx = add2(levelX, negHalf);   // levelX - 0.5
		const f2 sy = add2(levelY, negHalf);
		tx0[2 * h] = cvtt(lo2(sx)); tx0[2 * h + 1] = cvtt(hi2(sx));
		ty0[2 * h] = cvtt(lo2(sy)); ty0[2 * h + 1] = cvtt(hi2(sy));
		const f2 fx = add2(sub2(levelX, mk2(itof(tx0[2 * h]), itof(tx0[2 * h + 1]))), negHalf);
		const f2 fy = add2(sub2(levelY, mk2(itof(ty0[2 * h]), itof(ty0[2 * h + 1]))), negHalf);
		const f2 fx1 = sub2(one, fx);
		const f2 fy1 = sub2(one, fy);
		w00[h] = mul2(fx1, fy1);
		w10[h] = mul2(fx, fy1);
		w01[h] = mul2(fx1, fy);
		w11[h] = mul2(fx, fy); }

	if (f.stage != nullptr) {
		// ---- STAGED: every lane of the warp is here ---------------------------------------------------------
		const unsigned full = 0xffffffffu;
		const unsigned lane = threadIdx.x & 31u;
		int mnx = 0x7fffffff, mxx = static_cast<int>(0x80000000u), mny = 0x7fffffff, mxy = static_cast<int>(0x80000000u);
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			if ((mask >> l) & 1u) { mnx = min(mnx, tx0[l]); mxx = max(mxx, tx0[l]); mny = min(mny, ty0[l]); mxy = max(mxy, ty0[l]); } }
		const int lodLo = __reduce_min_sync(full, mask ? lod : 0x7fffffff);
		const int lodHi = __reduce_max_sync(full, mask ? lod : static_cast<int>(0x80000000u));
		const int bx0 = __reduce_min_sync(full, mnx), bx1 = __reduce_max_sync(full, mxx);   // (the taps reach one texel further: + 1 below)
		const int by0 = __reduce_min_sync(full, mny), by1 = __reduce_max_sync(full, mxy);
		const int me = bx0 & ~1;                                                   // first staged column (even)
		const unsigned spanX = static_cast<unsigned>(bx1 - me) + 1u, spanY = static_cast<unsigned>(by1 - by0) + 1u;
		const unsigned half = (spanX >> 1) + 1u;                                   // staged columns per parity
		const unsigned pitch = 2u * half, rows = spanY + 1u;
		if (lodLo == lodHi && spanX < 64u && spanY < 64u && pitch <= 32u && pitch * rows <= static_cast<unsigned>(kStageTexels)) {
			const int shift = POWER - lodLo;
			const uint32_t wrapMask = (1u << shift) - 1u;
			const uint32_t lastRow = ((0xffffffffu << shift) & ((1u << (POWER + 1)) - 1u)) - 1u;
			float4* stage = f.stage;
			if (lane < pitch) {
				const unsigned col = lane < half ? lane : lane - half;
				const uint32_t xw = static_cast<uint32_t>(me + 2 * static_cast<int>(col) + (lane >= half ? 1 : 0)) & wrapMask;
#pragma unroll 4
				for (unsigned rr = 0; rr < rows; ++rr) {
					const uint32_t yw = static_cast<uint32_t>(by0 + static_cast<int>(rr)) & wrapMask;
					const uint32_t ofs = min(((lastRow - yw) << POWER) + xw, lastTexel);
					stage[rr * pitch + lane] = __ldg(tu.texels + ofs); } }
			__syncwarp();
#pragma unroll
			for (int l = 0; l < 4; ++l) {
				if ((mask >> l) & 1u) {
					const int hh = l >> 1, j = l & 1;
					const unsigned k = static_cast<unsigned>(tx0[l] - me) >> 1;
					const bool odd = (tx0[l] & 1) != 0;
					const unsigned c0 = odd ? k + half : k, c1 = odd ? k + 1u : k + half;
					const float4* row0 = stage + static_cast<unsigned>(ty0[l] - by0) * pitch;
					const float4* row1 = row0 + pitch;
					const float4 p00 = row0[c0], p10 = row0[c1], p01 = row1[c0], p11 = row1[c1];
					blend_taps(p00, p10, p01, p11, j ? hi2(w00[hh]) : lo2(w00[hh]), j ? hi2(w10[hh]) : lo2(w10[hh]),
					           j ? hi2(w01[hh]) : lo2(w01[hh]), j ? hi2(w11[hh]) : lo2(w11[hh]), r[l], g[l], b[l], a[l]); }
				else { r[l] = 0.0f; g[l] = 0.0f; b[l] = 0.0f; a[l] = 0.0f; } }
			__syncwarp();   // the staging area is free again
			return; } }

	// ---- GATHER ----------------------------------------------------------------------------------------------
	const int wrapMask = levelDimI - 1;
	const int levelLastRow = static_cast<int>((0xffffffffu << (POWER - lod)) & ((1u << (POWER + 1)) - 1u)) - 1;
#pragma unroll
	for (int l = 0; l < 4; ++l) {
		if (!((mask >> l) & 1u)) { r[l] = 0.0f; g[l] = 0.0f; b[l] = 0.0f; a[l] = 0.0f; continue; }
		const int hh = l >> 1, j = l & 1;
		const int x0 = tx0[l] & wrapMask, x1 = (tx0[l] + 1) & wrapMask;
		const int by0 = levelLastRow - (ty0[l] & wrapMask);
		const int by1 = levelLastRow - ((ty0[l] + 1) & wrapMask);
		// (the reference would read out of bounds for a texture without its mip rows; stay inside)
		const float4 p00 = __ldg(tu.texels + min(static_cast<uint32_t>((by0 << POWER) + x0), lastTexel));
		const float4 p10 = __ldg(tu.texels + min(static_cast<uint32_t>((by0 << POWER) + x1), lastTexel));
		const float4 p01 = __ldg(tu.texels + min(static_cast<uint32_t>((by1 << POWER) + x0), lastTexel));
		const float4 p11 = __ldg(tu.texels + min(static_cast<uint32_t>((by1 << POWER) + x1), lastTexel));
		blend_taps(p00, p10, p01, p11, j ? hi2(w00[hh]) : lo2(w00[hh]), j ? hi2(w10[hh]) : lo2(w10[hh]),
		           j ? hi2(w01[hh]) : lo2(w01[hh]), j ? hi2(w11[hh]) : lo2(w11[hh]), r[l], g[l], b[l], a[l]); } }

// DepthTextureUnit::sample (rglr_texture_sampler.hxx:61-79): nearest, clamp-to-border(-1)
__device__ __forceinline__ float sample_depth(const DevState& st, float cx, float cy) {
	const float dimf = itof(st.tu3dim);
	const bool hit = (cx >= 0.0f) && (cx < 1.0f) && (cy >= 0.0f) && (cy < 1.0f);
	const int px = cvtt(cx * dimf);
	const int py = cvtt((1.0f - cy) * dimf);
	const int ofs = (py * st.tu3dim + px) & (st.tu3dim * st.tu3dim - 1);
	const float c = st.tu3 ? __ldg(st.tu3 + ofs) : 0.0f;
	return hit ? c : -1.0f; }

// ---- programs ---------------------------------------------------------------------------------
// ShadeVertex: writes clip position and NV varyings.
// ShadeFragment: attrs[k][lane]; writes colour (r,g,b,a)[lane]; may clear bits of `mask` (discard).

struct ProgBase {   // rglv::BaseProgram (rglv_gpu_shaders.hxx:21-95): the depth-only program of the shadow-map pass (node/gllayer.cxx:44-45, :162-176)
	static constexpr int id = 0;
	static constexpr bool samples = false;   // fragment stage samples texture unit 0 (the tile kernel stages its TexDesc)
	static constexpr bool earlyZ = true;
	static constexpr int NV = 0;
	__device__ static void ShadeVertex(const DevState& s, const VertexIn& v, float (&pos)[4], float*) {
		mat4_mul(s.vpm, v.px, v.py, v.pz, 1.0f, pos[0], pos[1], pos[2], pos[3]); }
	__device__ static void ShadeFragment(const FragIn&, const float (&)[kMaxVaryings][4],
	                                     float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], uint32_t&) {
#pragma unroll
		for (int l = 0; l < 4; ++l) { r[l] = 1.0f; g[l] = 1.0f; b[l] = 1.0f; a[l] = 1.0f; } } };

struct ProgAmy : ProgBase {   // shaders.hxx:69-161
	static constexpr int id = 4;
	static constexpr bool samples = true;
	static constexpr int NV = 2;
	__device__ static void ShadeVertex(const DevState& s, const VertexIn& v, float (&pos)[4], float* vary) {
		vary[0] = v.u; vary[1] = v.v;
		mat4_mul(s.vpm, v.px, v.py, v.pz, 1.0f, pos[0], pos[1], pos[2], pos[3]); }
	__device__ static void ShadeFragment(const FragIn& f, const float (&at)[kMaxVaryings][4],
	                                     float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], uint32_t& mask) {
		sample_quad(f, at[0], at[1], r, g, b, a, mask); } };

struct ProgAlphaTexture : ProgAmy {   // shaders.hxx:164-229
	static constexpr int id = 65;
	static constexpr bool earlyZ = false;
	__device__ static void ShadeFragment(const FragIn& f, const float (&at)[kMaxVaryings][4],
	                                     float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], uint32_t& mask) {
		sample_quad(f, at[0], at[1], r, g, b, a, mask);
#pragma unroll
		for (int l = 0; l < 4; ++l) { if (!(a[l] > 0.0f)) { mask &= ~(1u << l); } } } };

struct ProgText : ProgBase {   // shaders.hxx:232-318
	static constexpr int id = 26;
	static constexpr bool samples = true;
	static constexpr bool earlyZ = false;
	static constexpr int NV = 5;
	__device__ static void ShadeVertex(const DevState& s, const VertexIn& v, float (&pos)[4], float* vary) {
		vary[0] = v.kx; vary[1] = v.ky; vary[2] = v.kz; vary[3] = v.u; vary[4] = v.v;
		mat4_mul(s.vpm, v.px, v.py, v.pz, 1.0f, pos[0], pos[1], pos[2], pos[3]); }
	__device__ static void ShadeFragment(const FragIn& f, const float (&at)[kMaxVaryings][4],
	                                     float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], uint32_t& mask) {
		sample_quad(f, at[3], at[4], r, g, b, a, mask);
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			r[l] *= at[0][l]; g[l] *= at[1][l]; b[l] *= at[2][l];
			if (!(a[l] > 0.0f)) { mask &= ~(1u << l); } } } };

struct ProgDepth : ProgAmy {   // shaders.hxx:321-419
	static constexpr int id = 5;
	static constexpr bool samples = false;
	__device__ static void ShadeFragment(const FragIn& f, const float (&at)[kMaxVaryings][4],
	                                     float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], uint32_t&) {
		const float zNear = 10.0f, zFar = 1000.0f;
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			float tmp = sample_depth(*f.st, at[0][l], at[1][l]);
			tmp = (2.0f * zNear) / ((zFar + zNear) - tmp * (zFar - zNear));
			r[l] = tmp; g[l] = tmp; b[l] = tmp; a[l] = 1.0f; } } };

struct ProgPattern : ProgBase {   // shaders.hxx:422-479; uniforms: vec4 offset, vec4 dim
	static constexpr int id = 41;
	static constexpr bool samples = true;
	__device__ static void ShadeVertex(const DevState& s, const VertexIn& v, float (&pos)[4], float*) {
		mat4_mul(s.vpm, v.px, v.py, v.pz, 1.0f, pos[0], pos[1], pos[2], pos[3]); }
	__device__ static void ShadeFragment(const FragIn& f, const float (&)[kMaxVaryings][4],
	                                     float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], uint32_t& mask) {
		const float offx = f.st->uniforms[0], offy = f.st->uniforms[1], dimy = f.st->uniforms[5];
		float u[4], v[4];
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			u[l] = f.fragX[l] / dimy + offx;
			v[l] = f.fragY[l] / dimy + offy; }
		sample_quad(f, u, v, r, g, b, a, mask); } };

struct ProgMany : ProgBase {   // shaders.hxx:482-583; uniforms: float magic
	static constexpr int id = 6;
	static constexpr int NV = 2;
	__device__ static void ShadeVertex(const DevState& s, const VertexIn& v, float (&pos)[4], float* vary) {
		float p1[4];
		mat4_mul(v.imat, v.px, v.py, v.pz, 1.0f, p1[0], p1[1], p1[2], p1[3]);
		vary[0] = v.u; vary[1] = v.v;
		mat4_mul(s.vpm, p1[0], p1[1], p1[2], p1[3], pos[0], pos[1], pos[2], pos[3]); }
	__device__ static void ShadeFragment(const FragIn& f, const float (&at)[kMaxVaryings][4],
	                                     float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], uint32_t&) {
		const float magic = f.st->uniforms[0];
#pragma unroll
		for (int l = 0; l < 4; ++l) { r[l] = at[0][l]; g[l] = at[1][l]; b[l] = magic; a[l] = 1.0f; } } };

// lightPos - p with the reference's qfloat4::operator- (rmlv_soa.hxx:163: w uses rhs.z), then
// length / normalize / dot over 4 components exactly as OBJ1/OBJ2 do
__device__ __forceinline__ float obj_point_light(const float (&sp)[4], const float (&sn)[4],
                                                 const uint32_t* rsqrtLut) {
	const float dx = 0.0f - sp[0], dy = 0.0f - sp[1], dz = 0.0f - sp[2], dw = 1.0f - sp[2];
	const float d2 = ((dx * dx + dy * dy) + dz * dz) + dw * dw;
	const float distance = sqrtf(d2);
	const float scale = rsqrt_intel(d2, rsqrtLut);
	const float lx = dx * scale, ly = dy * scale, lz = dz * scale, lw = dw * scale;
	float diffuse = sse_max(((sn[0] * lx + sn[1] * ly) + sn[2] * lz) + sn[3] * lw, 0.1f);
	diffuse = diffuse * (100.0f / distance);
	return diffuse; }

struct ProgOBJ1 : ProgBase {   // shaders.hxx:586-698
	static constexpr int id = 7;
	static constexpr int NV = 3;
	__device__ static void ShadeVertex(const DevState& s, const VertexIn& v, float (&pos)[4], float* vary,
	                                   const uint32_t* rsqrtLut) {
		float mvv[4], mvn[4];
		mat4_mul(s.vm, v.px, v.py, v.pz, 1.0f, mvv[0], mvv[1], mvv[2], mvv[3]);
		mat4_mul(s.vm, v.nx, v.ny, v.nz, 0.0f, mvn[0], mvn[1], mvn[2], mvn[3]);
		const float diffuse = obj_point_light(mvv, mvn, rsqrtLut);
		vary[0] = v.kx * diffuse; vary[1] = v.ky * diffuse; vary[2] = v.kz * diffuse;
		mat4_mul(s.vpm, v.px, v.py, v.pz, 1.0f, pos[0], pos[1], pos[2], pos[3]); }
	__device__ static void ShadeFragment(const FragIn&, const float (&at)[kMaxVaryings][4],
	                                     float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], uint32_t&) {
#pragma unroll
		for (int l = 0; l < 4; ++l) { r[l] = at[0][l]; g[l] = at[1][l]; b[l] = at[2][l]; a[l] = 1.0f; } } };

struct ProgOBJ2 : ProgBase {   // shaders.hxx:701-818; varyings sp(4) sn(4) kd(3)
	static constexpr int id = 8;
	static constexpr int NV = 11;
	__device__ static void ShadeVertex(const DevState& s, const VertexIn& v, float (&pos)[4], float* vary) {
		mat4_mul(s.vm, v.px, v.py, v.pz, 1.0f, vary[0], vary[1], vary[2], vary[3]);
		mat4_mul(s.vm, v.nx, v.ny, v.nz, 0.0f, vary[4], vary[5], vary[6], vary[7]);
		vary[8] = v.kx; vary[9] = v.ky; vary[10] = v.kz;
		mat4_mul(s.vpm, v.px, v.py, v.pz, 1.0f, pos[0], pos[1], pos[2], pos[3]); }
	__device__ static void ShadeFragment(const FragIn& f, const float (&at)[kMaxVaryings][4],
	                                     float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], uint32_t&) {
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			const float sp[4] = { at[0][l], at[1][l], at[2][l], at[3][l] };
			const float sn[4] = { at[4][l], at[5][l], at[6][l], at[7][l] };
			const float diffuse = obj_point_light(sp, sn, f.rsqrtLut);
			r[l] = at[8][l] * diffuse; g[l] = at[9][l] * diffuse; b[l] = at[10][l] * diffuse; a[l] = 1.0f; } } };

struct ProgOBJ2S : ProgBase {   // shaders.hxx:821-978; varyings sp(4) sn(4) kd(3) lp(4)
	static constexpr int id = 9;
	static constexpr int NV = 15;
	// uniforms: mat4 modelToShadow (0..15), vec3 lpos (16..18), vec3 ldir (19..21), float lcos (22)
	// (rmlv::vec3 is 3 packed floats in UniformsSD)
	__device__ static void ShadeVertex(const DevState& s, const VertexIn& v, float (&pos)[4], float* vary) {
		mat4_mul(s.vm, v.px, v.py, v.pz, 1.0f, vary[0], vary[1], vary[2], vary[3]);
		mat4_mul(s.nm, v.nx, v.ny, v.nz, 1.0f, vary[4], vary[5], vary[6], vary[7]);
		vary[8] = v.kx; vary[9] = v.ky; vary[10] = v.kz;
		mat4_mul(s.uniforms, v.px, v.py, v.pz, 1.0f, vary[11], vary[12], vary[13], vary[14]);
		mat4_mul(s.vpm, v.px, v.py, v.pz, 1.0f, pos[0], pos[1], pos[2], pos[3]); }
	__device__ static void ShadeFragment(const FragIn& f, const float (&at)[kMaxVaryings][4],
	                                     float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], uint32_t&);
};

struct ProgEnvmap : ProgBase {   // shaders_envmap.hxx:26-140
	static constexpr int id = 10;
	static constexpr bool samples = true;
	static constexpr int NV = 2;
	__device__ static void ShadeVertex(const DevState& s, const VertexIn& v, float (&pos)[4], float* vary,
	                                   const uint32_t* rsqrtLut) {
		float p[4];
		mat4_mul(s.vm, v.px, v.py, v.pz, 1.0f, p[0], p[1], p[2], p[3]);
		const float es = rsqrt_intel((p[0] * p[0] + p[1] * p[1]) + p[2] * p[2], rsqrtLut);
		const float ex = p[0] * es, ey = p[1] * es, ez = p[2] * es;
		// vn = mix(smoothNormal, faceNormal, 0.0F) = a + t*(b-a)   (rmlv_mvec4.hxx:533-535)
		const float vnx = v.nx + 0.0f * (v.kx - v.nx);
		const float vny = v.ny + 0.0f * (v.ky - v.ny);
		const float vnz = v.nz + 0.0f * (v.kz - v.nz);
		float tx, ty, tz;
		mat4_mul_w0(s.nm, vnx, vny, vnz, tx, ty, tz);
		const float ns = rsqrt_intel((tx * tx + ty * ty) + tz * tz, rsqrtLut);
		const float nx = tx * ns, ny = ty * ns, nz = tz * ns;
		// reflect(i, n) = i - 2.0F * dot(n, i) * n   (rglv_math.hxx:44-45)
		const float k = 2.0f * ((nx * ex + ny * ey) + nz * ez);
		const float rx = ex - k * nx, ry = ey - k * ny, rz = ez - k * nz;
		const float m = 2.0f * sqrtf(((rx * rx) + (ry * ry)) + ((rz + 1.0f) * (rz + 1.0f)));
		vary[0] = rx / m + 0.5f;
		vary[1] = ry / m + 0.5f;
		mat4_mul(s.vpm, v.px, v.py, v.pz, 1.0f, pos[0], pos[1], pos[2], pos[3]); }
	__device__ static void ShadeFragment(const FragIn& f, const float (&at)[kMaxVaryings][4],
	                                     float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], uint32_t& mask) {
		sample_quad(f, at[0], at[1], r, g, b, a, mask);
#pragma unroll
		for (int l = 0; l < 4; ++l) { a[l] = 0.5f; } } };

struct ProgWireframe : ProgBase {   // shaders_wireframe.hxx:29-70 (vertex stage = BaseProgram)
	static constexpr int id = 11;
	__device__ static void ShadeVertex(const DevState& s, const VertexIn& v, float (&pos)[4], float*) {
		mat4_mul(s.vpm, v.px, v.py, v.pz, 1.0f, pos[0], pos[1], pos[2], pos[3]); }
	__device__ static void fwidth4(const float (&a)[4], float (&o)[4]) {
		// rglv_fragment.hxx:11-19: dFdx = yyww - xxzz, dFdy = xyxy - zwzw
		const float dx0 = a[1] - a[0], dx1 = a[3] - a[2];
		const float dy0 = a[0] - a[2], dy1 = a[1] - a[3];
		o[0] = fabsf(dx0) + fabsf(dy0); o[1] = fabsf(dx0) + fabsf(dy1);
		o[2] = fabsf(dx1) + fabsf(dy0); o[3] = fabsf(dx1) + fabsf(dy1); }
	__device__ static float smoothstep0(float bb, float t) {
		// rmlv_mvec4.hxx:767-769 with a = 0
		float x = (t - 0.0f) / (bb - 0.0f);
		x = sse_max(sse_min(x, 1.0f), 0.0f);
		return (x * x) * (3.0f - 2.0f * x); }
	__device__ static void ShadeFragment(const FragIn& f, const float (&)[kMaxVaryings][4],
	                                     float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], uint32_t&) {
		float dX[4], dY[4], dZ[4];
		fwidth4(f.BPx, dX); fwidth4(f.BPy, dY); fwidth4(f.BPz, dZ);
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			const float ax = smoothstep0(dX[l] * 1.5f, f.BPx[l]);
			const float ay = smoothstep0(dY[l] * 1.5f, f.BPy[l]);
			const float az = smoothstep0(dZ[l] * 1.5f, f.BPz[l]);
			const float e = sse_min(ax, sse_min(ay, az));
			// mix(mvec4f, mvec4f, mvec4f) = a + t*(b-a)
			r[l] = (0.1f + e * (0.5f - 0.1f)) * 4.0f;
			g[l] = (0.1f + e * (0.6f - 0.1f)) * 4.0f;
			b[l] = (0.1f + e * (0.7f - 0.1f)) * 4.0f;
			a[l] = 0.0f; } } };

__device__ inline void ProgOBJ2S::ShadeFragment(const FragIn& f, const float (&at)[kMaxVaryings][4],
                                                float (&r)[4], float (&g)[4], float (&b)[4], float (&a)[4], uint32_t&) {
	const float* u = f.st->uniforms;
	const float lposx = u[16], lposy = u[17], lposz = u[18];
	const float ldx = u[19], ldy = u[20], ldz = u[21];
	const float lcos = u[22];
	const float lds = rsqrt_intel((ldx * ldx + ldy * ldy) + ldz * ldz, f.rsqrtLut);
	const float ndx = ldx * lds, ndy = ldy * lds, ndz = ldz * lds;
#pragma unroll
	for (int l = 0; l < 4; ++l) {
		const float dx = lposx - at[0][l], dy = lposy - at[1][l], dz = lposz - at[2][l];
		const float d2 = (dx * dx + dy * dy) + dz * dz;
		const float s = rsqrt_intel(d2, f.rsqrtLut);
		const float stlx = dx * s, stly = dy * s, stlz = dz * s;
		const float distanceToLight = sqrtf(d2);
		float attenuation = 100.0f / distanceToLight;
		const float angle = ((-stlx) * ndx + (-stly) * ndy) + (-stlz) * ndz;
		if (angle < lcos) { attenuation = 0.111f; }
		const float n2 = (at[4][l] * at[4][l] + at[5][l] * at[5][l]) + at[6][l] * at[6][l];
		const float ns = rsqrt_intel(n2, f.rsqrtLut);
		const float nx = at[4][l] * ns, ny = at[5][l] * ns, nz = at[6][l] * ns;
		const float diffuseFactor = sse_max(0.0f, (nx * stlx + ny * stly) + nz * stlz);
		const float lpx = at[11][l] / at[14][l];
		const float lpy = at[12][l] / at[14][l];
		const float lpz = at[13][l] / at[14][l];
		const float tmp = sample_depth(*f.st, lpx * 0.5f + 0.5f, lpy * 0.5f + 0.5f);
		if ((lpz - 0.0005f) > tmp) { attenuation = 0.111f; }
		r[l] = (attenuation * at[8][l]) * diffuseFactor;
		g[l] = (attenuation * at[9][l]) * diffuseFactor;
		b[l] = (attenuation * at[10][l]) * diffuseFactor;
		a[l] = 1.0f; } }

}  // namespace rsr
