// kernels.cuh -- the frame pipeline as sm_100a kernels.
//
//   K1 vertex_kernel      ShadeVertex + frustum flags + pdiv + viewport  (rglv_gpu_impl.hxx:372-385)
//   K2 setup_kernel       per-triangle classify / cull / tile bbox, Sutherland-Hodgman clip
//                                                        (rglv_gpu_impl.hxx:391-494, :678-793)
//   K3 bin_kernel<COUNT>  per-(chunk,tile) counts, submission order preserved
//   K4 scan_*             prefix sums -> per-tile list offsets
//   K5 bin_kernel<FILL>   writes the per-tile triangle lists
//   K6 tile_kernel        one CTA per 32x32 screen tile: clear, edge-function raster with depth
//                         test, interpolation, fragment programs, blend, resolve / sRGB store
//                                (rglv_gpu.cxx:263-432, rglv_gpu_impl.hxx:166-222, :841-998,
//                                 rglv_triangle.hxx:79-304, rglr_algorithm.hxx:31-104)
//
// Work that the reference does on one thread (BinImpl) is spread over the whole GPU; ordering is
// kept by construction: triangle ids increase in submission order (per draw: unclipped triangles
// in (instance, index) order, then that draw's clipped fan triangles) and every per-tile list is
// written in increasing id order.
#pragma once
#include <cstddef>
#include "programs.cuh"

namespace rsr {

constexpr int kTile = 32;                 // device tile edge in pixels (independent of the reference's)
constexpr int kTileThreads = 256;         // one thread per 2x2 quad
constexpr int kBatch = 256;               // triangles set up per pass of the tile kernel
constexpr int kMaxChunkShift = 10;        // a bin row covers at most 1024 triangle ids (chosen per frame)
constexpr int kBinWarps = 4;              // warps per bin CTA
constexpr int kMaxFan = 6;                // a clipped triangle has <= 8 vertices => <= 6 fan triangles
constexpr int kClipVaryF4 = 4;            // float4s of varyings per clipped vertex (>= kMaxVaryings/4)

constexpr uint32_t kReject = 0xffffffffu;
constexpr uint32_t kClipSrc = 0x40000000u;
constexpr uint32_t kBackface = 0x01000000u;
constexpr uint32_t kNoClipRec = 0x00ffffffu;

struct DevDraw {
	int state;
	int prims;              // triangles per instance
	int instances;          // >= 1
	int instanced;          // slot 15 matrices are loaded per instance
	int nverts;             // vertices shaded per instance
	int nvary;              // floats of varyings
	int strideF4;           // float4s per post-transform vertex record: 2 + ceil(nvary/4)
	uint32_t vbaseF4;       // first record, in float4 units
	uint32_t flagBase;      // first vertex flag byte
	const uint16_t* indices;   // nullptr = DrawArrays
	uint32_t idBase;        // (unused; triangles are identified by their global index pjobBase + local)
	uint32_t N;             // prims * instances
	uint32_t vjobBase;      // prefix of instances*nverts
	uint32_t pjobBase;      // prefix of N
	uint32_t clipSegBase;   // first clip segment of this draw in the segment table
	uint32_t batchKey; };   // program id | pipeline flags << 8: draws with equal keys may share a raster batch

struct ClipVertex {
	float4 dev;                       // device x, y, ndc z, 1/w  (rglv_gpu_impl.hxx:763-767)
	float4 vary[kClipVaryF4]; };

struct ClipRec {
	int nverts;
	int backfacing;
	uint32_t fan[kMaxFan];            // packed tile bbox per fan triangle or kReject
	uint32_t draw, key, state, pad;   // owning draw, its batch key and state index
	ClipVertex v[8]; };
static_assert(sizeof(ClipRec) % 16 == 0 && offsetof(ClipRec, v) == 48, "ClipRec layout");

// Everything the tile kernel needs to set up one accepted (unclipped) triangle, written once by K2:
// 28.4 fixed-point vertices (back faces already re-wound), ndc z and 1/w per vertex, where the
// varyings live, and the owning draw.  One 80-byte record = five 128-bit loads, no further gathers.
struct TriRec {
	int X[3], Y[3];
	float z[3], iw[3];
	uint32_t vref[3];
	uint32_t draw;
	uint32_t key, state, pad0, pad1; };
static_assert(sizeof(TriRec) == 80, "TriRec layout");

constexpr uint32_t kFanIdBit = 0x80000000u;   // list entry: fan triangle (clipRec index << 3 | k) instead of a triangle index

struct BinSeg {
	uint32_t draw;
	uint32_t kind;    // 0 = triangles [start, start+len), 1 = clip sources [start, start+len)
	uint32_t start;
	uint32_t len; };

enum FrameCmdType : int { kCmdClear = 1, kCmdStoreTC = 3, kCmdStoreFP = 4, kCmdStoreDepth = 5, kCmdStoreHalfFP = 6 };

struct FrameCmd {       // non-draw commands only; draws are found through the tile lists
	int type;
	int state;      // state in effect
	int arg;        // clear bits / gamma flag
	int beforeDraw; // number of draws recorded before this command
	void* dst;      // device destination for stores
	int dstStride;  // in pixels
	int pad2; };

struct FrameParams {
	int width, height;
	int tilesX, tilesY;
	int refTileW, refTileH;      // the reference tile the edge start point is taken from
	int postTileW, postTileH;    // the reference's actual tile size in pixels (FilterTile's running coordinates)
	float guardFactor;           // CalcGuardBandFactor (rglv_view_frustum.hxx:36-39)
	int ndraws, ncmds;
	uint32_t totalVJobs, totalPJobs;
	int segShift;                // log2(triangles per bin segment), 6..10, chosen per frame
	uint32_t clipCapacity;       // ClipRec capacity
	uint32_t listCapacity; };

struct Counters {
	unsigned int clipAlloc;          // ClipRec bump allocator
	unsigned int overflow;           // bit 0 clip records, bit 1 tile lists
	unsigned long long binned;       // triangles accepted for binning (incl. clip fan triangles)
	unsigned long long clipped;      // triangles sent to the clipper
	unsigned long long entries;      // (triangle, tile) pairs
	unsigned long long fragments; }; // pixels written

__device__ __forceinline__ int find_draw(const DevDraw* __restrict__ draws, int ndraws, uint32_t job, bool vertexJobs) {
	int lo = 0, hi = ndraws - 1;
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		const uint32_t b = vertexJobs ? draws[mid].vjobBase : draws[mid].pjobBase;
		if (b <= job) { lo = mid; } else { hi = mid - 1; } }
	return lo; }

__device__ __forceinline__ uint32_t pack_tiles(int tx0, int ty0, int tx1, int ty1) {
	return static_cast<uint32_t>(tx0) | (static_cast<uint32_t>(ty0) << 6) |
	       (static_cast<uint32_t>(tx1) << 12) | (static_cast<uint32_t>(ty1) << 18); }

// ---------------------------------------------------------------------------------------------
// K1: vertex stage
// ---------------------------------------------------------------------------------------------

template <class P>
__device__ __forceinline__ void shade_vertex(const DevState& s, const VertexIn& vi, float (&pos)[4], float* vary,
                                             const uint32_t* rsqrtLut) {
	if constexpr (P::id == ProgOBJ1::id || P::id == ProgEnvmap::id) { P::ShadeVertex(s, vi, pos, vary, rsqrtLut); }
	else { P::ShadeVertex(s, vi, pos, vary); } }

__device__ __forceinline__ void run_vertex_program(const DevState& s, const VertexIn& vi, float (&pos)[4], float* vary,
                                                   const uint32_t* rsqrtLut) {
	switch (s.programId) {
	case ProgAmy::id:          shade_vertex<ProgAmy>(s, vi, pos, vary, rsqrtLut); break;
	case ProgAlphaTexture::id: shade_vertex<ProgAlphaTexture>(s, vi, pos, vary, rsqrtLut); break;
	case ProgText::id:         shade_vertex<ProgText>(s, vi, pos, vary, rsqrtLut); break;
	case ProgDepth::id:        shade_vertex<ProgDepth>(s, vi, pos, vary, rsqrtLut); break;
	case ProgPattern::id:      shade_vertex<ProgPattern>(s, vi, pos, vary, rsqrtLut); break;
	case ProgMany::id:         shade_vertex<ProgMany>(s, vi, pos, vary, rsqrtLut); break;
	case ProgOBJ1::id:         shade_vertex<ProgOBJ1>(s, vi, pos, vary, rsqrtLut); break;
	case ProgOBJ2::id:         shade_vertex<ProgOBJ2>(s, vi, pos, vary, rsqrtLut); break;
	case ProgOBJ2S::id:        shade_vertex<ProgOBJ2S>(s, vi, pos, vary, rsqrtLut); break;
	case ProgEnvmap::id:       shade_vertex<ProgEnvmap>(s, vi, pos, vary, rsqrtLut); break;
	case ProgWireframe::id:    shade_vertex<ProgWireframe>(s, vi, pos, vary, rsqrtLut); break;
	default: pos[0] = pos[1] = pos[2] = 0.0f; pos[3] = 1.0f; break; } }

__global__ void __launch_bounds__(256)
vertex_kernel(const DevDraw* __restrict__ draws, const DevState* __restrict__ states, FrameParams fp,
              const ApproxLuts* __restrict__ luts, float4* __restrict__ ptvb, uint8_t* __restrict__ vflags) {
	const uint32_t job = blockIdx.x * blockDim.x + threadIdx.x;
	if (job >= fp.totalVJobs) { return; }
	const int di = find_draw(draws, fp.ndraws, job, true);
	const DevDraw& d = draws[di];
	const DevState& s = states[d.state];
	const uint32_t local = job - d.vjobBase;
	const uint32_t iid = local / static_cast<uint32_t>(d.nverts);
	const uint32_t v = local - iid * static_cast<uint32_t>(d.nverts);

	// attribute fetch: coalesced SoA loads; defaults as the reference's Loader::LoadLane
	VertexIn vi;
	const float* const* B = s.buffers;
	if (B[0]) { vi.px = __ldg(B[0] + v); vi.py = __ldg(B[1] + v); vi.pz = __ldg(B[2] + v); }
	else { vi.px = vi.py = vi.pz = 0.0f; }
	if (B[3]) { vi.nx = __ldg(B[3] + v); vi.ny = __ldg(B[4] + v); vi.nz = __ldg(B[5] + v); }
	else { vi.nx = 0.0f; vi.ny = 0.0f; vi.nz = 1.0f; }
	if (B[6]) { vi.kx = __ldg(B[6] + v); vi.ky = __ldg(B[7] + v); vi.kz = __ldg(B[8] + v); }
	else { vi.kx = vi.ky = vi.kz = 1.0f; }
	vi.u = B[9] ? __ldg(B[9] + v) : 0.0f;
	vi.v = B[10] ? __ldg(B[10] + v) : 0.0f;
	vi.imat = (d.instanced && B[15]) ? (B[15] + static_cast<size_t>(iid) * 16) : s.vm;

	float pos[4];
	float vary[kMaxVaryings];
#pragma unroll
	for (int k = 0; k < kMaxVaryings; ++k) { vary[k] = 0.0f; }
	run_vertex_program(s, vi, pos, vary, luts->rsqrt);

	// ViewFrustum::Test, SIMD flavour: <= 0 (rglv_view_frustum.hxx:60-73)
	const float gw = pos[3] * fp.guardFactor;
	uint32_t flags = 0;
	flags |= (gw + pos[0] <= 0.0f) ? 1u : 0u;
	flags |= (gw + pos[1] <= 0.0f) ? 2u : 0u;
	flags |= (pos[3] + pos[2] <= 0.0f) ? 4u : 0u;
	flags |= (gw - pos[0] <= 0.0f) ? 8u : 0u;
	flags |= (gw - pos[1] <= 0.0f) ? 16u : 0u;

	// pdiv (rglv_math.hxx:19-23) and viewport transform
	const float iw = oneover(pos[3], luts->rcp);
	const float nx = pos[0] * iw, ny = pos[1] * iw, nz = pos[2] * iw;
	const float devx = nx * s.DSx + s.DOx;
	const float devy = ny * s.DSy + s.DOy;

	float4* rec = ptvb + d.vbaseF4 + static_cast<size_t>(local) * d.strideF4;
	rec[0] = make_float4(devx, devy, nz, iw);
	rec[1] = make_float4(pos[0], pos[1], pos[2], pos[3]);
	const int nf4 = d.strideF4 - 2;
	for (int k = 0; k < nf4; ++k) {
		rec[2 + k] = make_float4(vary[4 * k], vary[4 * k + 1], vary[4 * k + 2], vary[4 * k + 3]); }
	vflags[d.flagBase + local] = static_cast<uint8_t>(flags); }

// ---------------------------------------------------------------------------------------------
// K2: triangle setup (classify, cull, tile bbox) + clipper
// ---------------------------------------------------------------------------------------------

struct CVert { float c[4]; float vary[kMaxVaryings]; };

// tile range of a device-space triangle; formulas of BinTriangles* (rglv_gpu_impl.hxx:452-465)
// and ForEachCoveredTile (:796-817)
__device__ __forceinline__ bool tile_bbox(const DevState& s, int ix0, int iy0, int ix1, int iy1, int ix2, int iy2,
                                          bool needNonEmpty, uint32_t& packed, const FrameParams& fp) {
	const int vminx = max(min(ix0, min(ix1, ix2)), s.scissorX0);
	const int vminy = max(min(iy0, min(iy1, iy2)), s.scissorY0);
	const int vmaxx = min(max(ix0, max(ix1, ix2)) + 1, s.scissorX1 - 1);
	const int vmaxy = min(max(iy0, max(iy1, iy2)) + 1, s.scissorY1 - 1);
	if (needNonEmpty && !((vmaxx > vminx) && (vmaxy > vminy))) { return false; }
	// C++ integer division truncates toward zero (vmax* may be negative for off-screen fans)
	int tx0 = vminx / kTile, ty0 = vminy / kTile, tx1 = vmaxx / kTile, ty1 = vmaxy / kTile;
	if (tx1 < tx0 || ty1 < ty0) { return false; }
	tx0 = max(tx0, 0); ty0 = max(ty0, 0);
	tx1 = min(tx1, fp.tilesX - 1); ty1 = min(ty1, fp.tilesY - 1);
	if (tx1 < tx0 || ty1 < ty0) { return false; }
	packed = pack_tiles(tx0, ty0, tx1, ty1);
	return true; }

__device__ __forceinline__ float clip_dist(int plane, const float* c) {
	// ViewFrustum::IsInside / Distance operands (rglv_view_frustum.hxx:75-94)
	switch (plane) {
	case 0: return c[3] + c[0];   // Left
	case 1: return c[3] + c[1];   // Bottom
	case 2: return c[3] + c[2];   // Near
	case 3: return c[3] - c[0];   // Right
	default: return c[3] - c[1]; } }  // Top

__device__ __noinline__ uint32_t clip_triangle(const DevDraw& d, uint32_t drawIndex, const DevState& s, const FrameParams& fp,
                                               const float4* __restrict__ r0, const float4* __restrict__ r1,
                                               const float4* __restrict__ r2, const ApproxLuts* __restrict__ luts,
                                               ClipRec* __restrict__ clipRecs, Counters* __restrict__ ctr) {
	CVert bufA[9], bufB[9];
	CVert* A = bufA; CVert* Bv = bufB;
	int na = 3;
	const float4* recs[3] = { r0, r1, r2 };
	const int nvary = d.nvary;
	for (int i = 0; i < 3; ++i) {
		const float4 c = recs[i][1];
		A[i].c[0] = c.x; A[i].c[1] = c.y; A[i].c[2] = c.z; A[i].c[3] = c.w;
		for (int k = 0; k < kMaxVaryings; ++k) { A[i].vary[k] = 0.0f; }
		for (int k = 0; k < d.strideF4 - 2; ++k) {
			const float4 q = recs[i][2 + k];
			A[i].vary[4 * k] = q.x; A[i].vary[4 * k + 1] = q.y; A[i].vary[4 * k + 2] = q.z; A[i].vary[4 * k + 3] = q.w; } }

	// Sutherland-Hodgman against Left, Bottom, Near, Right, Top (rglv_gpu_impl.hxx:725-758)
	for (int plane = 0; plane < 5 && na > 0; ++plane) {
		int nb = 0;
		bool hereIn = clip_dist(plane, A[0].c) >= 0.0f;
		for (int hi = 0; hi < na; ++hi) {
			const int ni = (hi + 1) % na;
			const bool nextIn = clip_dist(plane, A[ni].c) >= 0.0f;
			if (hereIn) { Bv[nb++] = A[hi]; }
			if (hereIn != nextIn) {
				const CVert& from = hereIn ? A[hi] : A[ni];
				const CVert& to = hereIn ? A[ni] : A[hi];
				const float da = clip_dist(plane, from.c);
				const float db = clip_dist(plane, to.c);
				const float t = da / (da - db);
				// mix(a, b, t) = (1 - t)*a + t*b   (rmlv_math.hxx:87-90)
				const float omt = 1.0f - t;
				CVert nv;
				for (int k = 0; k < 4; ++k) { nv.c[k] = omt * from.c[k] + t * to.c[k]; }
				for (int k = 0; k < kMaxVaryings; ++k) {
					nv.vary[k] = (k < nvary) ? (omt * from.vary[k] + t * to.vary[k]) : 0.0f; }
				if (nb < 9) { Bv[nb++] = nv; }
				hereIn = !hereIn; } }
		CVert* tmp = A; A = Bv; Bv = tmp;
		na = nb; }
	if (na < 3) { return kClipSrc | kNoClipRec; }
	if (na > 8) { na = 8; }

	// clip -> device coordinates (rglv_gpu_impl.hxx:763-767)
	for (int i = 0; i < na; ++i) {
		const float iw = oneover(A[i].c[3], luts->rcp);
		const float x = A[i].c[0] * iw, y = A[i].c[1] * iw, z = A[i].c[2] * iw;
		A[i].c[0] = x * s.DSx + s.DOx;
		A[i].c[1] = y * s.DSy + s.DOy;
		A[i].c[2] = z;
		A[i].c[3] = iw; }

	// facing / culling / winding (rglv_gpu_impl.hxx:770-782).  Note the bit tests on cullFace here,
	// versus the equality tests of the unclipped path.
	const float d31x = A[2].c[0] - A[0].c[0], d31y = A[2].c[1] - A[0].c[1];
	const float d21x = A[1].c[0] - A[0].c[0], d21y = A[1].c[1] - A[0].c[1];
	const float area = d31x * d21y - d31y * d21x;
	const bool backfacing = area < 0.0f;
	bool willCull = true;
	if (backfacing) { if (!s.cullingEnabled || ((s.cullFace & 2) == 0)) { willCull = false; } }
	else { if (!s.cullingEnabled || ((s.cullFace & 1) == 0)) { willCull = false; } }
	if (willCull) { return kClipSrc | kNoClipRec; }

	const unsigned int slot = atomicAdd(&ctr->clipAlloc, 1u);
	if (slot >= fp.clipCapacity) { atomicOr(&ctr->overflow, 1u); return kClipSrc | kNoClipRec; }
	ClipRec& rec = clipRecs[slot];
	rec.nverts = na;
	rec.backfacing = backfacing ? 1 : 0;
	rec.draw = drawIndex; rec.key = d.batchKey; rec.state = static_cast<uint32_t>(d.state); rec.pad = 0;
	for (int i = 0; i < na; ++i) {
		const CVert& src = backfacing ? A[na - 1 - i] : A[i];
		rec.v[i].dev = make_float4(src.c[0], src.c[1], src.c[2], src.c[3]);
		for (int k = 0; k < kClipVaryF4; ++k) {
			rec.v[i].vary[k] = make_float4(src.vary[4 * k], src.vary[4 * k + 1], src.vary[4 * k + 2], src.vary[4 * k + 3]); } }
	int nbinned = 0;
	for (int f = 0; f < kMaxFan; ++f) {
		uint32_t packed = kReject;
		if (f + 2 < na) {
			const float4 a = rec.v[0].dev, b = rec.v[f + 1].dev, c = rec.v[f + 2].dev;
			// ivec2{vec2} is a C cast: truncation, same saturation as cvtt for our purposes
			uint32_t p;
			if (tile_bbox(s, cvtt(a.x), cvtt(a.y), cvtt(b.x), cvtt(b.y), cvtt(c.x), cvtt(c.y), false, p, fp)) {
				packed = p | (backfacing ? kBackface : 0u);
				++nbinned; } }
		rec.fan[f] = packed; }
	if (nbinned) { atomicAdd(&ctr->binned, static_cast<unsigned long long>(nbinned)); }
	return kClipSrc | slot; }

__global__ void __launch_bounds__(256)
setup_kernel(const DevDraw* __restrict__ draws, const DevState* __restrict__ states, FrameParams fp,
             const ApproxLuts* __restrict__ luts, const float4* __restrict__ ptvb, const uint8_t* __restrict__ vflags,
             uint32_t* __restrict__ triInfo, TriRec* __restrict__ triRecs, ClipRec* __restrict__ clipRecs,
             unsigned int* __restrict__ segActive, Counters* __restrict__ ctr) {
	const uint32_t job = blockIdx.x * blockDim.x + threadIdx.x;
	if (job >= fp.totalPJobs) { return; }
	const int di = find_draw(draws, fp.ndraws, job, false);
	const DevDraw& d = draws[di];
	const DevState& s = states[d.state];
	const uint32_t local = job - d.pjobBase;
	const uint32_t iid = local / static_cast<uint32_t>(d.prims);
	const uint32_t prim = local - iid * static_cast<uint32_t>(d.prims);

	uint32_t i0, i1, i2;
	if (d.indices) {
		i0 = __ldg(d.indices + 3 * prim); i1 = __ldg(d.indices + 3 * prim + 1); i2 = __ldg(d.indices + 3 * prim + 2); }
	else { i0 = 3 * prim; i1 = i0 + 1; i2 = i0 + 2; }
	const uint32_t nverts = static_cast<uint32_t>(d.nverts);
	uint32_t out = kReject;
	if (i0 < nverts && i1 < nverts && i2 < nverts) {
		const uint32_t vb = iid * nverts;
		const uint32_t cf0 = vflags[d.flagBase + vb + i0];
		const uint32_t cf1 = vflags[d.flagBase + vb + i1];
		const uint32_t cf2 = vflags[d.flagBase + vb + i2];
		const float4* r0 = ptvb + d.vbaseF4 + static_cast<size_t>(vb + i0) * d.strideF4;
		const float4* r1 = ptvb + d.vbaseF4 + static_cast<size_t>(vb + i1) * d.strideF4;
		const float4* r2 = ptvb + d.vbaseF4 + static_cast<size_t>(vb + i2) * d.strideF4;
		const bool pointsOutside = (cf0 | cf1 | cf2) != 0;
		const bool primOutside = (cf0 & cf1 & cf2) != 0;
		if (primOutside) { out = kReject; }
		else if (pointsOutside) {
			atomicAdd(&ctr->clipped, 1ull);
			out = clip_triangle(d, static_cast<uint32_t>(di), s, fp, r0, r1, r2, luts, clipRecs, ctr);
			if (out != (kClipSrc | kNoClipRec)) { atomicAdd(&segActive[d.clipSegBase + (local >> fp.segShift)], 1u); } }
		else {
			const float4 a = __ldg(r0), b = __ldg(r1), c = __ldg(r2);
			// rmlg::Area (rmlg_triangle.hxx:18-27)
			const float d31x = c.x - a.x, d31y = c.y - a.y;
			const float d21x = b.x - a.x, d21y = b.y - a.y;
			const float area = d31x * d21y - d31y * d21x;
			const bool front = area > 0.0f;
			const bool keepBacks = !(s.cullingEnabled && s.cullFace == 2);
			const bool keepFronts = !(s.cullingEnabled && s.cullFace == 1);
			const bool notCulled = front ? keepFronts : keepBacks;
			uint32_t packed;
			if (notCulled && tile_bbox(s, cvtt(a.x), cvtt(a.y), cvtt(b.x), cvtt(b.y), cvtt(c.x), cvtt(c.y), true, packed, fp)) {
				out = packed | (front ? 0u : kBackface);
				atomicAdd(&ctr->binned, 1ull);
				// triangle record for the tile kernel; back faces are drawn with i0 <-> i2 swapped
				// (rglv_gpu_impl.hxx:467-470), fixed point = trunc(16 * dev) (:885-886)
				const float4 v0 = front ? a : c, v2 = front ? c : a;
				const uint32_t j0 = front ? i0 : i2, j2 = front ? i2 : i0;
				const uint32_t vbF4 = d.vbaseF4 + vb * d.strideF4 + 2u;
				TriRec rec;
				rec.X[0] = cvtt(16.0f * v0.x); rec.X[1] = cvtt(16.0f * b.x); rec.X[2] = cvtt(16.0f * v2.x);
				rec.Y[0] = cvtt(16.0f * v0.y); rec.Y[1] = cvtt(16.0f * b.y); rec.Y[2] = cvtt(16.0f * v2.y);
				rec.z[0] = v0.z; rec.z[1] = b.z; rec.z[2] = v2.z;
				rec.iw[0] = v0.w; rec.iw[1] = b.w; rec.iw[2] = v2.w;
				rec.vref[0] = vbF4 + j0 * d.strideF4; rec.vref[1] = vbF4 + i1 * d.strideF4; rec.vref[2] = vbF4 + j2 * d.strideF4;
				rec.draw = static_cast<uint32_t>(di); rec.key = d.batchKey; rec.state = static_cast<uint32_t>(d.state);
				rec.pad0 = rec.pad1 = 0;
				const uint4* src = reinterpret_cast<const uint4*>(&rec);
				uint4* dst = reinterpret_cast<uint4*>(triRecs + job);
#pragma unroll
				for (int q = 0; q < 5; ++q) { dst[q] = src[q]; } } } }
	triInfo[job] = out; }

// ---------------------------------------------------------------------------------------------
// K3/K5: order-preserving binning.  One warp owns one chunk (row).  32 ids at a time; the warp
// repeatedly takes the smallest tile index any lane still has to visit (__reduce_min_sync), all
// lanes that cover that tile get consecutive slots in lane (= submission) order via a ballot.
// ---------------------------------------------------------------------------------------------

template <bool FILL>
__device__ __forceinline__ void bin_items(bool valid, uint32_t info, uint32_t id, int tilesX, int bandY0, int bandY1,
                                          uint32_t* __restrict__ row, uint32_t* __restrict__ lists, uint32_t listCapacity) {
	const unsigned lane = threadIdx.x & 31u;
	const unsigned ltMask = (1u << lane) - 1u;
	// this warp only bins into tile rows [bandY0, bandY1]
	const int tx0 = info & 63, tx1 = (info >> 12) & 63;
	const int ty0 = max(static_cast<int>((info >> 6) & 63), bandY0), ty1 = min(static_cast<int>((info >> 18) & 63), bandY1);
	valid = valid && (ty0 <= ty1);
	info = pack_tiles(tx0, max(ty0, 0), tx1, max(ty1, 0));
	const unsigned validMask = __ballot_sync(0xffffffffu, valid);
	if (validMask == 0) { return; }

	// Two ways to hand out list slots in submission order for these 32 triangles:
	//  (a) tile-serial: repeatedly take the smallest tile index any lane still has to visit; the lanes
	//      covering it get consecutive slots by ballot.  One step per distinct tile: good for small
	//      triangles (a few tiles in total).
	//  (b) tile-parallel: every lane owns one tile of the union bbox and walks the 32 triangles in
	//      order.  32 steps per 32 tiles of the union: good for large, overlapping triangles.
	const unsigned ux0 = __reduce_min_sync(0xffffffffu, valid ? static_cast<unsigned>(tx0) : 63u);
	const unsigned uy0 = __reduce_min_sync(0xffffffffu, valid ? static_cast<unsigned>(ty0) : 63u);
	const unsigned ux1 = __reduce_max_sync(0xffffffffu, valid ? static_cast<unsigned>(tx1) : 0u);
	const unsigned uy1 = __reduce_max_sync(0xffffffffu, valid ? static_cast<unsigned>(ty1) : 0u);
	const unsigned sumTiles = __reduce_add_sync(0xffffffffu, valid ? static_cast<unsigned>((tx1 - tx0 + 1) * (ty1 - ty0 + 1)) : 0u);
	const unsigned uw = ux1 - ux0 + 1u, uh = uy1 - uy0 + 1u, U = uw * uh;
	if (sumTiles >= 48u && U * 2u < sumTiles * 5u) {
		for (unsigned base = 0; base < U; base += 32u) {
			const unsigned u = base + lane;
			const int tx = static_cast<int>(ux0 + u % uw), ty = static_cast<int>(uy0 + u / uw);
			const bool own = u < U;
			const uint32_t T = static_cast<uint32_t>(ty * tilesX + tx);
			const uint32_t start = own ? row[T] : 0u;
			uint32_t cnt = 0;
			unsigned m = validMask;
			while (m) {
				const int j = __ffs(m) - 1;
				m &= m - 1;
				const uint32_t bj = __shfl_sync(0xffffffffu, info, j);
				const uint32_t idj = __shfl_sync(0xffffffffu, id, j);
				const int jx0 = bj & 63, jy0 = (bj >> 6) & 63, jx1 = (bj >> 12) & 63, jy1 = (bj >> 18) & 63;
				if (own && tx >= jx0 && tx <= jx1 && ty >= jy0 && ty <= jy1) {
					if (FILL) { const uint32_t pos = start + cnt; if (pos < listCapacity) { lists[pos] = idj; } }
					++cnt; } }
			if (own && cnt) { row[T] = start + cnt; } }
		__syncwarp();
		return; }

	int cx = tx0, cy = ty0;
	uint32_t cur = valid ? static_cast<uint32_t>(cy * tilesX + cx) : 0xffffffffu;
	while (true) {
		const uint32_t B = __reduce_min_sync(0xffffffffu, cur);
		if (B == 0xffffffffu) { break; }
		const bool mine = (cur == B);
		const unsigned m = __ballot_sync(0xffffffffu, mine);
		const uint32_t base = row[B];
		__syncwarp();
		if (mine) {
			if (FILL) {
				const uint32_t pos = base + __popc(m & ltMask);
				if (pos < listCapacity) { lists[pos] = id; } }
			if ((m & ltMask) == 0) { row[B] = base + __popc(m); }
			++cx;
			if (cx > tx1) { cx = tx0; ++cy; }
			cur = (cy > ty1) ? 0xffffffffu : static_cast<uint32_t>(cy * tilesX + cx); }
		__syncwarp(); } }

template <bool FILL>
__global__ void __launch_bounds__(kBinWarps * 32)
bin_kernel(const DevDraw* __restrict__ draws, const BinSeg* __restrict__ segs, const uint32_t* __restrict__ chunkSegBegin,
           int nchunks, int nbands, int ntiles, int tilesX, const uint32_t* __restrict__ triInfo, const ClipRec* __restrict__ clipRecs,
           const unsigned int* __restrict__ segActive, uint32_t* __restrict__ counts, uint32_t* __restrict__ lists,
           uint32_t listCapacity) {
	extern __shared__ uint32_t smemRows[];
	const int warp = threadIdx.x >> 5;
	const unsigned lane = threadIdx.x & 31u;
	// one warp per (chunk, band of tile rows): a frame with few triangles that each cover hundreds of
	// tiles is split over nbands warps per chunk instead of serialising on one
	const int rowIdx = blockIdx.x * kBinWarps + warp;
	if (rowIdx >= nchunks * nbands) { return; }
	const int chunk = rowIdx / nbands, band = rowIdx - chunk * nbands;
	const int tilesY = ntiles / tilesX;
	const int rowsPerBand = (tilesY + nbands - 1) / nbands;
	const int bandY0 = band * rowsPerBand, bandY1 = min(bandY0 + rowsPerBand, tilesY) - 1;
	uint32_t* row = smemRows + static_cast<size_t>(warp) * ntiles;
	uint32_t* grow = counts + static_cast<size_t>(rowIdx) * ntiles;
	for (int t = lane; t < ntiles; t += 32) { row[t] = FILL ? grow[t] : 0u; }
	__syncwarp();

	for (uint32_t si = chunkSegBegin[chunk]; si < chunkSegBegin[chunk + 1]; ++si) {
		const BinSeg seg = segs[si];
		const DevDraw& d = draws[seg.draw];
		if (seg.kind == 0) {
			for (uint32_t i = 0; i < seg.len; i += 32) {
				const uint32_t li = i + lane;
				const uint32_t id = d.pjobBase + seg.start + li;   // global triangle index
				const uint32_t info = (li < seg.len) ? __ldg(triInfo + id) : kReject;
				const bool valid = (info != kReject) && !(info & kClipSrc);
				bin_items<FILL>(valid, info, id, tilesX, bandY0, bandY1, row, lists, listCapacity); } }
		else {
			if (segActive[si] == 0) { continue; }
			for (uint32_t i = 0; i < seg.len; i += 32) {
				const uint32_t li = i + lane;
				const uint32_t src = seg.start + li;
				const uint32_t info = (li < seg.len) ? __ldg(triInfo + d.pjobBase + src) : kReject;
				const bool isClip = (info != kReject) && (info & kClipSrc) && ((info & kNoClipRec) != kNoClipRec);
				unsigned m = __ballot_sync(0xffffffffu, isClip);
				while (m) {
					const int j = __ffs(m) - 1;
					m &= m - 1;
					const uint32_t recIdx = __shfl_sync(0xffffffffu, info, j) & kNoClipRec;
					const uint32_t fi = (lane < kMaxFan) ? clipRecs[recIdx].fan[lane] : kReject;
					const uint32_t id = kFanIdBit | (recIdx << 3) | lane;
					bin_items<FILL>(fi != kReject, fi, id, tilesX, bandY0, bandY1, row, lists, listCapacity); } } } }
	__syncwarp();
	if (!FILL) { for (int t = lane; t < ntiles; t += 32) { grow[t] = row[t]; } } }

// K4: counts[chunk][tile] -> absolute list offsets, in place.  Three small kernels.
__global__ void scan_group_sums(const uint32_t* __restrict__ counts, int nchunks, int ntiles, int chunksPerGroup,
                                uint32_t* __restrict__ gsum) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int g = blockIdx.y;
	if (t >= ntiles) { return; }
	const int c0 = g * chunksPerGroup, c1 = min(c0 + chunksPerGroup, nchunks);
	uint32_t s = 0;
#pragma unroll 8
	for (int c = c0; c < c1; ++c) { s += counts[static_cast<size_t>(c) * ntiles + t]; }
	gsum[static_cast<size_t>(g) * ntiles + t] = s; }

__global__ void __launch_bounds__(1024)
scan_tiles(uint32_t* __restrict__ gsum, int ngroups, int ntiles, uint32_t* __restrict__ tileBase,
           uint32_t* __restrict__ tileCount, Counters* __restrict__ ctr, uint32_t listCapacity) {
	// single CTA; ntiles <= 4096 => <= 4 tiles per thread
	__shared__ uint32_t warpSums[32];
	__shared__ uint32_t carry;
	const int tid = threadIdx.x;
	if (tid == 0) { carry = 0; }
	__syncthreads();
	for (int t0 = 0; t0 < ntiles; t0 += 1024) {
		const int t = t0 + tid;
		uint32_t total = 0;
		if (t < ntiles) {
			for (int g = 0; g < ngroups; g += 8) {
				uint32_t v[8];
#pragma unroll
				for (int k = 0; k < 8; ++k) { v[k] = (g + k < ngroups) ? gsum[static_cast<size_t>(g + k) * ntiles + t] : 0u; }
#pragma unroll
				for (int k = 0; k < 8; ++k) {
					if (g + k < ngroups) { gsum[static_cast<size_t>(g + k) * ntiles + t] = total; total += v[k]; } } } }
		// block exclusive scan of total
		uint32_t incl = total;
		for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) { incl += n; } }
		if ((tid & 31) == 31) { warpSums[tid >> 5] = incl; }
		__syncthreads();
		if (tid < 32) {
			uint32_t w = warpSums[tid];
			for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, w, o); if (tid >= o) { w += n; } }
			warpSums[tid] = w; }
		__syncthreads();
		const uint32_t before = carry + ((tid >> 5) ? warpSums[(tid >> 5) - 1] : 0u) + (incl - total);
		if (t < ntiles) {
			tileBase[t] = before;
			tileCount[t] = total; }
		__syncthreads();
		if (tid == 1023) { carry = before + total; }
		__syncthreads(); }
	if (tid == 0) {
		ctr->entries = carry;
		if (carry > listCapacity) { atomicOr(&ctr->overflow, 2u); } } }

__global__ void scan_apply(uint32_t* __restrict__ counts, int nchunks, int ntiles, int chunksPerGroup,
                           const uint32_t* __restrict__ gsum, const uint32_t* __restrict__ tileBase) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int g = blockIdx.y;
	if (t >= ntiles) { return; }
	const int c0 = g * chunksPerGroup, c1 = min(c0 + chunksPerGroup, nchunks);
	uint32_t run = gsum[static_cast<size_t>(g) * ntiles + t] + tileBase[t];
	// 8 independent loads in flight, then 8 stores: the chain is the adds, not the memory round trips
	for (int c = c0; c < c1; c += 8) {
		uint32_t v[8];
#pragma unroll
		for (int k = 0; k < 8; ++k) { v[k] = (c + k < c1) ? counts[static_cast<size_t>(c + k) * ntiles + t] : 0u; }
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			if (c + k < c1) { counts[static_cast<size_t>(c + k) * ntiles + t] = run; run += v[k]; } } } }

}  // namespace rsr
