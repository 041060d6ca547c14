// kernels.cuh -- the frame pipeline as sm_100a kernels.
//
//   K1 vertex_kernel      ShadeVertex + frustum flags + pdiv + viewport  (rglv_gpu_impl.hxx:372-385)
//   K2 setup_kernel       per-triangle classify / cull / tile bbox, Sutherland-Hodgman clip
//                                                        (rglv_gpu_impl.hxx:391-494, :678-793)
//                         ... and the per-tile entry counts (fire-and-forget atomics); the last CTA
//                         to finish turns the counts into list offsets (exclusive scan)
//   K5 fill_kernel        appends (order key, triangle) entries to the per-tile lists, unordered
//   K6 tile_kernel        one CTA per 32x32 screen tile: sort its list by order key, clear, edge-function raster with depth
//                         test, interpolation, fragment programs, blend, resolve / sRGB store
//                                (rglv_gpu.cxx:263-432, rglv_gpu_impl.hxx:166-222, :841-998,
//                                 rglv_triangle.hxx:79-304, rglr_algorithm.hxx:31-104)
//
// Work that the reference does on one thread (BinImpl) is spread over the whole GPU.  Submission
// order -- which depth-LESS ties and blending depend on -- is carried by a 32-bit ORDER KEY per list
// entry: per draw, [idBase, idBase+N) are its triangles in (instance, index) order and
// [idBase+N, idBase+7N) the <= 6 fan triangles of each clipped source, so "unclipped first, then
// clipped, per draw" (rglv_gpu_impl.hxx:498-508) is ascending key order.  Lists are appended in any
// order with atomics; each tile CTA sorts its own list by key before rasterising.
#pragma once
#include <cstddef>
#include "programs.cuh"

namespace rsr {

constexpr int kTile = 32;                 // device tile edge in pixels (independent of the reference's)
constexpr int kTileThreads = 256;         // one thread per 2x2 quad
constexpr int kBatch = 256;               // triangles set up per pass of the tile kernel
constexpr int kMaxFan = 6;                // a clipped triangle has <= 8 vertices => <= 6 fan triangles
constexpr int kClipVaryF4 = 4;            // float4s of varyings per clipped vertex (>= kMaxVaryings/4)
constexpr int kMaxGroups = 64;            // list cells per tile
constexpr int kLargeTiles = 32;           // default of FrameParams::largeTiles
constexpr int kTileLargeCap = 128;        // queued items one tile can take (more: the frame is rendered again with a higher threshold)

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
	uint32_t idBase;        // first order key of this draw: 7 x (triangles of all earlier draws)
	uint32_t N;             // prims * instances
	uint32_t vjobBase;      // prefix of instances*nverts
	uint32_t pjobBase;      // prefix of N
	uint32_t cullBits;      // bit 0 culling enabled, bits 1-2 cull face (GL_FRONT 1, GL_BACK 2, both 3), bit 3 scissor enabled
	uint32_t batchKey; };   // program id | pipeline flags << 8: draws with equal keys may share a raster batch

struct ClipVertex {
	float4 dev;                       // device x, y, ndc z, 1/w  (rglv_gpu_impl.hxx:763-767)
	float4 vary[kClipVaryF4]; };

struct ClipRec {
	int nverts;
	int backfacing;
	uint32_t fan[kMaxFan];            // packed tile bbox per fan triangle or kReject
	uint32_t draw, key, state, okeyBase;   // owning draw, its batch key, state index | list group << 16, order key of fan triangle 0
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
constexpr uint32_t kRunStartBit = 0x40000000u;   // list entry: first entry of the ascending run one warp appended to the cell

enum FrameCmdType : int { kCmdClear = 1, kCmdStoreTC = 3, kCmdStoreFP = 4, kCmdStoreDepth = 5, kCmdStoreHalfFP = 6, kCmdStoreQuadsFP = 7 };

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
	uint32_t totalKeys;          // order keys in use: 7 x triangles
	int groups, groupShift;      // list cells per tile: cell g of a tile holds the triangles [g << groupShift, (g+1) << groupShift)
	uint32_t largeCapacity;      // capacity of the large-item queue
	int largeTiles;              // an item covering more tiles than this goes to the queue instead of the lists
	uint32_t clipCapacity;       // ClipRec capacity
	uint32_t listCapacity; };

struct Counters {
	unsigned int clipAlloc;          // ClipRec bump allocator
	unsigned int overflow;           // bit 0 clip records, bit 1 tile lists
	unsigned int ticket;             // CTAs of K2 that have finished counting
	unsigned int nLarge;             // items queued for the cooperative part of K5
	unsigned long long binned;       // triangles accepted for binning (incl. clip fan triangles)
	unsigned long long clipped;      // triangles sent to the clipper
	unsigned long long entries;      // (triangle, tile) pairs
	unsigned long long fragments;    // pixels written
	unsigned int chunksRunMerge;     // list chunks the tile kernel ordered by run merge (mode B) ...
	unsigned int chunksKeyRange; };  // ... and by key ranges + bitonic sort (mode C)

// Draw that owns a vertex / triangle job.  The host tabulates, per block of 256 jobs, the draw of the
// block's first job.  A block that lies inside one draw (the usual case) needs no search; a block that
// spans many small draws (hundreds of two-triangle draws) stages their job bases in shared memory
// with one coalesced trip and searches there -- a per-thread binary search over global memory is a
// chain of up to eight dependent misses.  Every thread of the block must call it.
__device__ __forceinline__ int find_draw(const DevDraw* __restrict__ draws, const uint32_t* __restrict__ blockDraw, uint32_t vblock, uint32_t job,
                                         bool vertexJobs) {
	__shared__ uint32_t sBase[260];
	const int first = static_cast<int>(__ldg(blockDraw + vblock));
	const int last = static_cast<int>(__ldg(blockDraw + vblock + 1));
	if (last <= first) { return first; }
	const int n = min(last - first + 1, 258);   // (every draw has at least one job: a block of 256 jobs spans <= 257 draws)
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		sBase[i] = vertexJobs ? __ldg(&draws[first + i].vjobBase) : __ldg(&draws[first + i].pjobBase); }
	__syncthreads();
	int lo = 0, hi = n - 1;
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (sBase[mid] <= job) { lo = mid; } else { hi = mid - 1; } }
	return first + lo; }

// Programmatic dependent launch (sm_90+): every kernel of the frame is launched with the "programmatic
// stream serialization" attribute.  pdl_launch_dependents lets the next kernel's CTAs become resident
// (and run their prologue) as soon as every CTA of this grid has passed the call; pdl_wait blocks until
// the previous grid has completed and its writes are visible.  Every kernel waits before it reads or
// writes anything another kernel of the stream touches, so completion stays transitive.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

__device__ __forceinline__ uint32_t pack_tiles(int tx0, int ty0, int tx1, int ty1) {
	return static_cast<uint32_t>(tx0) | (static_cast<uint32_t>(ty0) << 6) |
	       (static_cast<uint32_t>(tx1) << 12) | (static_cast<uint32_t>(ty1) << 18); }

// ---------------------------------------------------------------------------------------------
// K0: frame upload.  The SMs pull the frame's staging arena (pinned, mapped host memory: per-frame
// buffers and the state / draw tables) into its device mirror and zero the control block.  A copy
// engine would do the same, but a copy and a memset between two kernels of one stream cost two
// engine hand-overs (tens of microseconds of idle GPU per frame, measured: tools/trace_run.py).
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
upload_kernel(const uint4* __restrict__ hostSrc, uint4* __restrict__ dst, size_t n16, uint4* __restrict__ zero, size_t nzero16) {
	pdl_launch_dependents();
	pdl_wait();   // the previous frame's kernels still read the control block / may read this arena's mirror
	const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
	const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	for (size_t i = tid; i < n16; i += stride) { dst[i] = hostSrc[i]; }
	for (size_t i = tid; i < nzero16; i += stride) { zero[i] = make_uint4(0u, 0u, 0u, 0u); } }

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
	case ProgBase::id:         shade_vertex<ProgBase>(s, vi, pos, vary, rsqrtLut); break;
	default: pos[0] = pos[1] = pos[2] = 0.0f; pos[3] = 1.0f; break; } }

__global__ void __launch_bounds__(256)
vertex_kernel(const DevDraw* __restrict__ draws, const uint32_t* __restrict__ blockDraw, const DevState* __restrict__ states, FrameParams fp,
              const ApproxLuts* __restrict__ luts, float4* __restrict__ ptvb, uint8_t* __restrict__ vflags,
              uint4* __restrict__ zero, uint32_t nzero16) {
	pdl_launch_dependents();
	pdl_wait();
	// a replayed (retained) frame has nothing to upload: its first kernel is this one, and it zeroes the frame's control
	// block (counters, cell counts and cursors) instead of a separate K0 launch
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nzero16; i += gridDim.x * blockDim.x) { zero[i] = make_uint4(0u, 0u, 0u, 0u); }
	// (the grid may be smaller than the number of 256-job blocks: a frame whose front end runs under the previous
	// frame's tile kernel is launched on a capped grid so that it shares the SMs instead of displacing the tile CTAs)
	const uint32_t nblocks = (fp.totalVJobs + 255u) >> 8;
	for (uint32_t vblock = blockIdx.x; vblock < nblocks; vblock += gridDim.x) {
	if (vblock != blockIdx.x) { __syncthreads(); }   // find_draw's shared table
	const uint32_t job = vblock * blockDim.x + threadIdx.x;
	const int di = find_draw(draws, blockDraw, vblock, min(job, fp.totalVJobs - 1u), true);
	if (job >= fp.totalVJobs) { continue; }
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
	vflags[d.flagBase + local] = static_cast<uint8_t>(flags); } }

// ---------------------------------------------------------------------------------------------
// K2: triangle setup (classify, cull, tile bbox) + clipper
// ---------------------------------------------------------------------------------------------

struct CVert { float c[4]; float vary[kMaxVaryings]; };

// tile range of a device-space triangle; formulas of BinTriangles* (rglv_gpu_impl.hxx:452-465)
// and ForEachCoveredTile (:796-817)
__device__ __forceinline__ bool tile_bbox(int sx0, int sy0, int sx1, int sy1, int ix0, int iy0, int ix1, int iy1, int ix2, int iy2,
                                          bool needNonEmpty, uint32_t& packed, const FrameParams& fp) {
	const int vminx = max(min(ix0, min(ix1, ix2)), sx0);
	const int vminy = max(min(iy0, min(iy1, iy2)), sy0);
	const int vmaxx = min(max(ix0, max(ix1, ix2)) + 1, sx1 - 1);
	const int vmaxy = min(max(iy0, max(iy1, iy2)) + 1, sy1 - 1);
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

__device__ __noinline__ uint32_t clip_triangle(const DevDraw& d, uint32_t drawIndex, uint32_t local, const DevState& s, const FrameParams& fp,
                                               const float4* __restrict__ r0, const float4* __restrict__ r1,
                                               const float4* __restrict__ r2, const ApproxLuts* __restrict__ luts,
                                               ClipRec* __restrict__ clipRecs, Counters* __restrict__ ctr, const unsigned int slot) {
	CVert bufA[9], bufB[9];
	CVert* A = bufA; CVert* Bv = bufB;
	int na = 3;
	const float4* recs[3] = { r0, r1, r2 };
	const int nvary = d.nvary;
	const int nvary4 = 4 * (d.strideF4 - 2);   // varyings as stored: whole float4s (the pad lanes stay 0)
	for (int i = 0; i < 3; ++i) {
		const float4 c = recs[i][1];
		A[i].c[0] = c.x; A[i].c[1] = c.y; A[i].c[2] = c.z; A[i].c[3] = c.w;
		for (int k = 0; k < d.strideF4 - 2; ++k) {
			const float4 q = recs[i][2 + k];
			A[i].vary[4 * k] = q.x; A[i].vary[4 * k + 1] = q.y; A[i].vary[4 * k + 2] = q.z; A[i].vary[4 * k + 3] = q.w; } }

	// Sutherland-Hodgman against Left, Bottom, Near, Right, Top (rglv_gpu_impl.hxx:725-758)
	for (int plane = 0; plane < 5 && na > 0; ++plane) {
		// a plane every current vertex is inside of reproduces the polygon unchanged: skip the pass (most
		// triangles cross one plane; only the first nvary4 varyings are live, the rest are never read)
		bool allIn = true;
		for (int i = 0; i < na; ++i) { allIn = allIn && (clip_dist(plane, A[i].c) >= 0.0f); }
		if (allIn) { continue; }
		int nb = 0;
		bool hereIn = clip_dist(plane, A[0].c) >= 0.0f;
		for (int hi = 0; hi < na; ++hi) {
			const int ni = (hi + 1) % na;
			const bool nextIn = clip_dist(plane, A[ni].c) >= 0.0f;
			if (hereIn) {
				for (int k = 0; k < 4; ++k) { Bv[nb].c[k] = A[hi].c[k]; }
				for (int k = 0; k < nvary4; ++k) { Bv[nb].vary[k] = A[hi].vary[k]; }
				++nb; }
			if (hereIn != nextIn) {
				const CVert& from = hereIn ? A[hi] : A[ni];
				const CVert& to = hereIn ? A[ni] : A[hi];
				const float da = clip_dist(plane, from.c);
				const float db = clip_dist(plane, to.c);
				const float t = da / (da - db);
				// mix(a, b, t) = (1 - t)*a + t*b   (rmlv_math.hxx:87-90)
				const float omt = 1.0f - t;
				if (nb < 9) {
					for (int k = 0; k < 4; ++k) { Bv[nb].c[k] = omt * from.c[k] + t * to.c[k]; }
					for (int k = 0; k < nvary4; ++k) {
						Bv[nb].vary[k] = (k < nvary) ? (omt * from.vary[k] + t * to.vary[k]) : 0.0f; }
					++nb; }
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

	// (the record was reserved by the caller, one atomic per warp; a triangle that clips away or is culled leaves it unused)
	if (slot >= fp.clipCapacity) { atomicOr(&ctr->overflow, 1u); return kClipSrc | kNoClipRec; }
	ClipRec& rec = clipRecs[slot];
	rec.nverts = na;
	rec.backfacing = backfacing ? 1 : 0;
	// fan triangles are drawn after every unclipped triangle of their draw: they go to the list cell of the draw's last triangle
	rec.draw = drawIndex; rec.key = d.batchKey;
	rec.state = static_cast<uint32_t>(d.state) | (((d.pjobBase + d.N - 1u) >> fp.groupShift) << 16);
	rec.okeyBase = d.idBase + d.N + local * static_cast<uint32_t>(kMaxFan);
	for (int i = 0; i < na; ++i) {
		const CVert& src = backfacing ? A[na - 1 - i] : A[i];
		rec.v[i].dev = make_float4(src.c[0], src.c[1], src.c[2], src.c[3]);
		for (int k = 0; k < kClipVaryF4; ++k) {
			rec.v[i].vary[k] = (4 * k < nvary4) ? make_float4(src.vary[4 * k], src.vary[4 * k + 1], src.vary[4 * k + 2], src.vary[4 * k + 3])
			                                    : make_float4(0.0f, 0.0f, 0.0f, 0.0f); } }
	int nbinned = 0;
	for (int f = 0; f < kMaxFan; ++f) {
		uint32_t packed = kReject;
		if (f + 2 < na) {
			const float4 a = rec.v[0].dev, b = rec.v[f + 1].dev, c = rec.v[f + 2].dev;
			// ivec2{vec2} is a C cast: truncation, same saturation as cvtt for our purposes
			uint32_t p;
			if (tile_bbox(s.scissorX0, s.scissorY0, s.scissorX1, s.scissorY1, cvtt(a.x), cvtt(a.y), cvtt(b.x), cvtt(b.y), cvtt(c.x), cvtt(c.y), false, p, fp)) {
				packed = p | (backfacing ? kBackface : 0u);
				++nbinned; } }
		rec.fan[f] = packed; }
	if (nbinned) { atomicAdd(&ctr->binned, static_cast<unsigned long long>(nbinned)); }
	return kClipSrc | slot; }

// ---------------------------------------------------------------------------------------------
// Binning: one ITEM = one triangle (or clip fan triangle) with its tile range, order key and code.
// A tile's list is split into `groups` CELLS by triangle index range; cell (tile, g) receives its
// entries through an atomic cursor, in any order (the tile kernel sorts each cell by order key).
// K2 counts (COUNT), its last CTA scans the cell counts into list offsets, K5 writes (FILL).
// Lanes of a warp hold consecutive triangles, which mostly hit the same few tiles: lanes aiming at
// the same cell are found with __match_any_sync and share ONE atomic.
// ---------------------------------------------------------------------------------------------

struct LargeItem { uint32_t packed, okey, code, group; };

struct BinArgs {
	uint32_t* cellCount;         // [tile * groups + g], zeroed per frame; COUNT adds
	uint32_t* cellCursor;        // same shape, zeroed per frame; FILL takes slots
	const uint32_t* tileBase;    // [tile]: list offset of the tile's first cell (+ the total at the end)
	const uint32_t* cellRel;     // [tile * groups + g]: offset of the cell inside its tile's list (groups > 1 only)
	uint2* lists;
	LargeItem* large; };

template <bool FILL>
__device__ __forceinline__ void bin_item(bool valid, uint32_t packed, uint32_t okey, uint32_t code, uint32_t group,
                                         const FrameParams& fp, const BinArgs& B, Counters* __restrict__ ctr) {
	const unsigned lane = threadIdx.x & 31u;
	const unsigned ltMask = (1u << lane) - 1u;
	const int tx0 = packed & 63, ty0 = (packed >> 6) & 63, tx1 = (packed >> 12) & 63, ty1 = (packed >> 18) & 63;
	const int w = tx1 - tx0 + 1, ntile = valid ? w * (ty1 - ty0 + 1) : 0;
	const bool large = ntile > fp.largeTiles;
	if (!FILL && large) {
		// never binned: every tile CTA scans the (short) queue itself -- a triangle that covers more than
		// largeTiles tiles costs far more to rasterise than the one bbox test per tile this adds
		const unsigned slot = atomicAdd(&ctr->nLarge, 1u);
		if (slot < fp.largeCapacity) { B.large[slot] = LargeItem{packed, okey, code, group}; }
		else { atomicOr(&ctr->overflow, 4u); } }
	// One-tile items (dense meshes of small triangles: the long lists): the lanes of a warp that aim at
	// the same cell take consecutive slots in lane (= submission) order with ONE atomic, all cells at
	// once.  A warp thus adds one ascending run to a cell, which lets the tile kernel order long cells
	// run by run instead of entry by entry.
	const bool small = valid && !large;
	const bool single = small && ntile == 1 && fp.groups == 1;   // (frames with several cells per tile: see the walk below)
	if (__any_sync(0xffffffffu, single)) {
		const uint32_t cell = single ? static_cast<uint32_t>((ty0 * fp.tilesX + tx0) * fp.groups) + group : (0xffffff00u | lane);
		// (fan triangles carry keys above every unclipped key of their draw: they must not share a run with the
		// unclipped lanes around them, or the run would not ascend -- the fan bit is part of the match key)
		const unsigned peers = __match_any_sync(0xffffffffu, single ? (cell | (code & kFanIdBit)) : cell);
		const int leader = __ffs(peers) - 1;
		const uint32_t rank = __popc(peers & ltMask);
		if (!FILL) {
			if (single && rank == 0) { atomicAdd(B.cellCount + cell, static_cast<uint32_t>(__popc(peers))); } }
		else {
			uint32_t base = 0;
			if (single && rank == 0) {
				base = atomicAdd(B.cellCursor + cell, static_cast<uint32_t>(__popc(peers))) + __ldg(B.tileBase + ty0 * fp.tilesX + tx0);
				if (fp.groups > 1) { base += __ldg(B.cellRel + cell); } }
			base = __shfl_sync(0xffffffffu, base, leader);
			if (single) {
				const uint32_t pos = base + rank;
				if (pos < fp.listCapacity) { B.lists[pos] = make_uint2(okey, rank ? code : (code | kRunStartBit)); } } } }

	// Frames with millions of triangles (several cells per tile): items of up to 4 tiles are walked tile by
	// tile in increasing tile index -- each step serves the smallest tile index any lane is at, the
	// lanes at that tile take consecutive slots in lane order with one atomic -- so that straddling
	// triangles, too, stay inside the one ascending run per warp and cell.  (One atomic round trip per step: too slow for the
	// medium-sized triangles of ordinary frames, whose short lists do not need runs.)
	const bool walked = small && ntile <= 4 && fp.groups > 1;
	if (fp.groups > 1) {
		if (!FILL) {
			// Counting needs no order: round q adds the q-th tile of every item (row-major, like the walk below); the lanes
			// of a round that aim at one cell share one fire-and-forget atomic.  (The count used to run the ordered walk
			// too: a chain of warp votes per distinct tile of the warp, a quarter of K2's stall samples on c3.)
			int cx = tx0, cy = ty0;
			for (int q = 0; q < 4; ++q) {
				const bool on = walked && q < ntile;
				if (!__any_sync(0xffffffffu, on)) { break; }
				const uint32_t cell = on ? static_cast<uint32_t>(cy * fp.tilesX + cx) * static_cast<uint32_t>(fp.groups) + group : (0xffffff00u | lane);
				const unsigned peers = __match_any_sync(0xffffffffu, cell);
				if (on && (peers & ltMask) == 0u) { atomicAdd(B.cellCount + cell, static_cast<uint32_t>(__popc(peers))); }
				if (++cx > tx1) { cx = tx0; ++cy; } } }
		else {
			// The ordered walk, with its atomics deferred: the loop itself is warp votes only.  Step s (the s-th distinct
			// (tile, group) the warp meets, in increasing tile index) is remembered by lane s -- its cell and how many lanes
			// take part -- and every participating lane notes the step and its rank in it (at most four per item).  After
			// the walk (or after 32 steps) the step lanes issue all cursor atomics at once: one round trip instead of one
			// per step (K5 on c3: 370 warp instructions and about three serial atomic latencies per triangle thread before).
			int cx = tx0, cy = ty0;
			uint32_t cur = walked ? static_cast<uint32_t>(cy * fp.tilesX + cx) : 0xffffffffu;
			const uint32_t mygf = group | (code & kFanIdBit);
			uint32_t stepCell = 0, stepTile = 0, stepCnt = 0;   // of the step this lane remembers
			uint32_t mySteps = 0, myRanks = 0;                  // 4 x 8 bits: step (0 .. 31) and rank of this lane's entries
			int nmine = 0, nsteps = 0;
			auto flush = [&]() {
				uint32_t base = 0;
				if (static_cast<int>(lane) < nsteps) {
					base = atomicAdd(B.cellCursor + stepCell, stepCnt) + __ldg(B.tileBase + stepTile) + __ldg(B.cellRel + stepCell); }
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const uint32_t b = __shfl_sync(0xffffffffu, base, static_cast<int>((mySteps >> (8 * k)) & 31u));
					if (k < nmine) {
						const uint32_t rank = (myRanks >> (8 * k)) & 0xffu;
						const uint32_t pos = b + rank;
						if (pos < fp.listCapacity) { B.lists[pos] = make_uint2(okey, rank ? code : (code | kRunStartBit)); } } }
				mySteps = 0; myRanks = 0; nmine = 0; nsteps = 0; };
			while (true) {
				const uint32_t tile = __reduce_min_sync(0xffffffffu, cur);
				if (tile == 0xffffffffu) { break; }
				const int leader = __ffs(__ballot_sync(0xffffffffu, cur == tile)) - 1;
				const uint32_t g = __shfl_sync(0xffffffffu, mygf, leader);   // (lanes differ in group only among clip fans; fans never share a run with unclipped lanes)
				const bool mine = (cur == tile) && (mygf == g);
				const unsigned m = __ballot_sync(0xffffffffu, mine);
				if (static_cast<int>(lane) == nsteps) {
					stepTile = tile; stepCell = tile * static_cast<uint32_t>(fp.groups) + (g & ~kFanIdBit); stepCnt = static_cast<uint32_t>(__popc(m)); }
				if (mine) {
					mySteps |= static_cast<uint32_t>(nsteps) << (8 * nmine);
					myRanks |= static_cast<uint32_t>(__popc(m & ltMask)) << (8 * nmine);
					++nmine;
					++cx;
					if (cx > tx1) { cx = tx0; ++cy; }
					cur = (cy <= ty1) ? static_cast<uint32_t>(cy * fp.tilesX + cx) : 0xffffffffu; }
				if (++nsteps == 32) { flush(); } }
			if (nsteps) { flush(); } } }

	// Items that cover more tiles make short lists (few of them fit a tile): every lane appends on its
	// own, four independent atomics in flight; each entry is a run of its own.
	if (small && !single && !walked) {
		// (row-major walk over the tile range with running coordinates: a division per tile would cost more than the atomic)
		const uint32_t rowStep = static_cast<uint32_t>((fp.tilesX - w) * fp.groups);
		uint32_t cur = static_cast<uint32_t>((ty0 * fp.tilesX + tx0) * fp.groups) + group;
		int col = 0;
		for (int i0 = 0; i0 < ntile; i0 += 4) {
			uint32_t cell[4], pos[4];
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				cell[q] = cur;
				if (i0 + q + 1 < ntile) {
					cur += static_cast<uint32_t>(fp.groups);
					if (++col == w) { col = 0; cur += rowStep; } } }
			if (!FILL) {
#pragma unroll
				for (int q = 0; q < 4; ++q) { if (i0 + q < ntile) { atomicAdd(B.cellCount + cell[q], 1u); } } }
			else {
#pragma unroll
				for (int q = 0; q < 4; ++q) { pos[q] = (i0 + q < ntile) ? atomicAdd(B.cellCursor + cell[q], 1u) : 0u; }
#pragma unroll
				for (int q = 0; q < 4; ++q) {
					if (i0 + q < ntile) {
						const uint32_t tile = (fp.groups == 1) ? cell[q] : cell[q] / static_cast<uint32_t>(fp.groups);
						uint32_t p = pos[q] + __ldg(B.tileBase + tile);
						if (fp.groups > 1) { p += __ldg(B.cellRel + cell[q]); }
						if (p < fp.listCapacity) { B.lists[p] = make_uint2(okey, code | kRunStartBit); } } } } } } }

// the items of triangle `job`: itself, or the fan triangles of its clip record
template <bool FILL>
__device__ __forceinline__ void bin_triangle(uint32_t job, uint2 info, const FrameParams& fp, const ClipRec* __restrict__ clipRecs,
                                             const BinArgs& B, Counters* __restrict__ ctr) {
	const bool isClip = (info.x != kReject) && (info.x & kClipSrc);
	const bool hasRec = isClip && ((info.x & kNoClipRec) != kNoClipRec);
	const uint32_t recIdx = info.x & kNoClipRec;
	const uint32_t group = hasRec ? (clipRecs[recIdx].state >> 16) : (job >> fp.groupShift);
	const int nslots = __any_sync(0xffffffffu, hasRec) ? kMaxFan : 1;
	for (int slot = 0; slot < nslots; ++slot) {
		uint32_t packed = kReject, okey = 0, code = 0;
		if (hasRec) {
			packed = clipRecs[recIdx].fan[slot];
			okey = clipRecs[recIdx].okeyBase + static_cast<uint32_t>(slot);
			code = kFanIdBit | (recIdx << 3) | static_cast<uint32_t>(slot); }
		else if (!isClip && slot == 0 && info.x != kReject) { packed = info.x; okey = info.y; code = job; }
		bin_item<FILL>(packed != kReject, packed, okey, code, group, fp, B, ctr); } }

// Called by every thread of every CTA at the end of a kernel: the last CTA to get here scans
// count[0..n) (n <= 4096) into base[0..n] (threadFenceReduction pattern: every thread's writes and
// atomics are ordered before its CTA's ticket).
#ifdef RSR_PHASE_PROF
__device__ unsigned long long g_k2Times[8];
__device__ unsigned long long g_k2Block[12];   // globaltimer marks of thread 0 of the middle CTA   // globaltimer ns: [0] first CTA start (min), [1] last CTA's ticket, [2] after scan, [3] after tile order
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define K2_MARK(i) do { if (threadIdx.x == 0) { g_k2Times[i] = gtimer(); } } while (0)
// per-CTA phase durations of thread 0 (ns), maximum over CTAs: g_k2Block[i] = max(t_i - t_{i-1})
#define K2B(i) do { if (threadIdx.x == 0) { const unsigned long long now_ = gtimer(); if (i) { atomicMax(&g_k2Block[i], now_ - k2Last); } k2Last = now_; } } while (0)
#else
#define K2_MARK(i) do {} while (0)
#define K2B(i) do {} while (0)
#endif

__device__ __forceinline__ void tile_scan_last_block(const uint32_t* count, uint32_t* __restrict__ base, uint32_t* __restrict__ order,
                                                     const int n, const uint32_t listCapacity, Counters* __restrict__ ctr) {
	__shared__ bool lastBlock;
	__shared__ uint32_t warpSums[8];
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) { lastBlock = (atomicAdd(&ctr->ticket, 1u) == gridDim.x - 1); }
	__syncthreads();
	if (!lastBlock) { return; }
	K2_MARK(1);
	__threadfence();
	uint32_t v[16];   // 256 threads x 16 = 4096 tiles (the largest target is 2048 px = 64 x 64 tiles)
	uint32_t sum = 0;
#pragma unroll
	for (int q = 0; q < 16; ++q) {
		const int i = static_cast<int>(threadIdx.x) * 16 + q;
		v[q] = (i < n) ? __ldcg(count + i) : 0u;
		sum += v[q]; }
	uint32_t incl = sum;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) { incl += x; } }
	if (lane == 31) { warpSums[warp] = incl; }
	__syncthreads();
	uint32_t before = incl - sum;
	for (int w = 0; w < warp; ++w) { before += warpSums[w]; }
#pragma unroll
	for (int q = 0; q < 16; ++q) {
		const int i = static_cast<int>(threadIdx.x) * 16 + q;
		if (i < n) { base[i] = before; }
		before += v[q]; }
	if (threadIdx.x == 255) {
		base[n] = before;
		ctr->entries = before;
		if (before > listCapacity) { atomicOr(&ctr->overflow, 2u); } }

	K2_MARK(2);
#ifdef RSR_NO_TILE_ORDER
	for (int q = 0; q < 16; ++q) { const int i = static_cast<int>(threadIdx.x) * 16 + q; if (i < n) { order[i] = static_cast<uint32_t>(i); } }
	return;
#endif
	// CTA -> tile order for the tile kernel: eight classes of list length (>= 2048, 1024, 512, 256, 64,
	// 16, 1 entries, empty), longest first, tile index order inside a class (neighbouring tiles share
	// texels and vertex records), so that the heavy tiles start early and the light ones fill the tail
	// -- the reference sorts its tile jobs by list length for the same reason (rglv_gpu.cxx:100-108).
	// Stable counting sort: thread t owns tiles 16t..16t+15, warp k scans class k.  (Rolled loops on
	// purpose: this runs once per frame in one CTA, straight from a cold instruction cache.)
	__shared__ uint16_t perThread[8][256];
	__shared__ uint32_t classBase[8];
	__shared__ uint8_t tileClass[4096];
	for (int k = 0; k < 8; ++k) { perThread[k][threadIdx.x] = 0; }
#pragma unroll 1
	for (int q = 0; q < 16; ++q) {
		const int i = static_cast<int>(threadIdx.x) * 16 + q;
		if (i < n) {
			const int lg = min(32 - __clz(v[q]), 12);                    // 0 for an empty list, 12 for >= 2048
			const int k = static_cast<int>((0x0123445566667ull >> (4 * lg)) & 15ull);   // nibble table, lg = 0 first
			tileClass[i] = static_cast<uint8_t>(k);
			perThread[k][threadIdx.x] += 1; } }
	__syncthreads();
	{
		uint32_t tot = 0;
		for (int j = 0; j < 8; ++j) { tot += perThread[warp][lane * 8 + j]; }
		uint32_t inc = tot;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) { inc += x; } }
		uint32_t run = inc - tot;
		for (int j = 0; j < 8; ++j) { const uint32_t c = perThread[warp][lane * 8 + j]; perThread[warp][lane * 8 + j] = static_cast<uint16_t>(run); run += c; }
		if (lane == 31) { classBase[warp] = inc; } }
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t acc = 0;
		for (int k = 0; k < 8; ++k) { const uint32_t c = classBase[k]; classBase[k] = acc; acc += c; } }
	__syncthreads();
#pragma unroll 1
	for (int q = 0; q < 16; ++q) {
		const int i = static_cast<int>(threadIdx.x) * 16 + q;
		if (i < n) {
			const int k = tileClass[i];
			order[classBase[k] + perThread[k][threadIdx.x]] = static_cast<uint32_t>(i);
			perThread[k][threadIdx.x] += 1; } }
	K2_MARK(3); }

#ifndef RSR_SETUP_CTAS
#define RSR_SETUP_CTAS 5   // 48 registers (32 bytes of spill in the rare clip path): c3 setup 108 -> 101 us; 6 CTAs / 40 registers: 98 us but no faster overall
#endif
__global__ void __launch_bounds__(256, RSR_SETUP_CTAS)
setup_kernel(const DevDraw* __restrict__ draws, const uint32_t* __restrict__ blockDraw, const DevState* __restrict__ states, FrameParams fp,
             const ApproxLuts* __restrict__ luts, const float4* __restrict__ ptvb, const uint8_t* __restrict__ vflags,
             uint2* __restrict__ triInfo, TriRec* __restrict__ triRecs, ClipRec* __restrict__ clipRecs,
             BinArgs B, uint32_t* __restrict__ tileBase, uint32_t* __restrict__ tileOrder, Counters* __restrict__ ctr) {
	pdl_launch_dependents();
	pdl_wait();
#ifdef RSR_PHASE_PROF
	if (threadIdx.x == 0) { atomicMin(&g_k2Times[0], gtimer()); if (blockIdx.x == gridDim.x - 1) { g_k2Times[4] = gtimer(); } }
#endif
#ifdef RSR_PHASE_PROF
	unsigned long long k2Last = 0;
#endif
	const uint32_t nblocks = (fp.totalPJobs + 255u) >> 8;   // (capped grids: see vertex_kernel)
	for (uint32_t vblock = blockIdx.x; vblock < nblocks; vblock += gridDim.x) {
	if (vblock != blockIdx.x) { __syncthreads(); }   // find_draw's shared table
	const uint32_t job = vblock * blockDim.x + threadIdx.x;
	uint2 myInfo = make_uint2(kReject, 0u);
	K2B(0);
	const int di = find_draw(draws, blockDraw, vblock, min(job, fp.totalPJobs - 1u), false);
	bool needClip = false;
	const float4* clipR0 = nullptr; const float4* clipR1 = nullptr; const float4* clipR2 = nullptr;
	if (job < fp.totalPJobs) {
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
		K2B(1);
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
			else if (pointsOutside) { needClip = true; clipR0 = r0; clipR1 = r1; clipR2 = r2; }
			else {
				const float4 a = __ldg(r0), b = __ldg(r1), c = __ldg(r2);
				// rmlg::Area (rmlg_triangle.hxx:18-27)
				const float d31x = c.x - a.x, d31y = c.y - a.y;
				const float d21x = b.x - a.x, d21y = b.y - a.y;
				const float area = d31x * d21y - d31y * d21x;
				const bool front = area > 0.0f;
				// (cull mode and the scissor switch ride in the draw record: no trip to the state snapshot here)
				const bool cullingEnabled = (d.cullBits & 1u) != 0;
				const uint32_t cullFace = (d.cullBits >> 1) & 3u;
				const bool keepBacks = !(cullingEnabled && cullFace == 2);
				const bool keepFronts = !(cullingEnabled && cullFace == 1);
				const bool notCulled = front ? keepFronts : keepBacks;
				uint32_t packed;
				bool binned = false;
				if (notCulled) {
					binned = (d.cullBits & 8u) ? tile_bbox(s.scissorX0, s.scissorY0, s.scissorX1, s.scissorY1, cvtt(a.x), cvtt(a.y), cvtt(b.x), cvtt(b.y), cvtt(c.x), cvtt(c.y), true, packed, fp)
					                           : tile_bbox(0, 0, fp.width, fp.height, cvtt(a.x), cvtt(a.y), cvtt(b.x), cvtt(b.y), cvtt(c.x), cvtt(c.y), true, packed, fp); }
				if (binned) {
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
		K2B(2);
		myInfo = make_uint2(out, d.idBase + local);
		if (!needClip) { triInfo[job] = myInfo; } }
	// Clip sources: the warp reserves their records with ONE atomic (lanes reach the clipper through
	// different paths, so per-lane allocations would be one round trip each), then clips.
	{
		const unsigned lane = threadIdx.x & 31u;
		const unsigned m = __ballot_sync(0xffffffffu, needClip);
		if (m) {
			const int leader = __ffs(m) - 1;
			unsigned base = 0;
			if (static_cast<int>(lane) == leader) {
				base = atomicAdd(&ctr->clipAlloc, static_cast<unsigned>(__popc(m)));
				atomicAdd(&ctr->clipped, static_cast<unsigned long long>(__popc(m))); }
			base = __shfl_sync(0xffffffffu, base, leader);
			if (needClip) {
				const DevDraw& d = draws[di];
				const uint32_t local = job - d.pjobBase;
				myInfo.x = clip_triangle(d, static_cast<uint32_t>(di), local, states[d.state], fp, clipR0, clipR1, clipR2, luts, clipRecs, ctr,
				                         base + __popc(m & ((1u << lane) - 1u)));
				triInfo[job] = myInfo; } } }
	bin_triangle<false>(job, myInfo, fp, clipRecs, B, ctr);
	K2B(3); }
#ifdef RSR_PHASE_PROF
	__threadfence();
	K2B(4);
	__syncthreads();
	K2B(5);
#endif
	if (fp.groups > 1) { return; }   // K3 (cell_scan_kernel) prepares the offsets of multi-cell lists
	tile_scan_last_block(B.cellCount, tileBase, tileOrder, fp.tilesX * fp.tilesY, fp.listCapacity, ctr); }

// ---------------------------------------------------------------------------------------------
// K3 (frames with several cells per tile only): one warp per tile turns the tile's cell counts into
// offsets inside the tile's list and the tile's total; the last CTA scans the totals.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
cell_scan_kernel(FrameParams fp, const uint32_t* __restrict__ cellCount, uint32_t* __restrict__ cellRel,
                 uint32_t* __restrict__ tileTotal, uint32_t* __restrict__ tileBase, uint32_t* __restrict__ tileOrder,
                 Counters* __restrict__ ctr) {
	pdl_launch_dependents();
	pdl_wait();
	const int ntiles = fp.tilesX * fp.tilesY;
	const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);
	const int lane = threadIdx.x & 31;
	if (tile < ntiles) {
		const int G = fp.groups;   // <= kMaxGroups = 64: two cells per lane
		const uint32_t* cnt = cellCount + static_cast<size_t>(tile) * G;
		const uint32_t a = (2 * lane < G) ? __ldcg(cnt + 2 * lane) : 0u;
		const uint32_t b = (2 * lane + 1 < G) ? __ldcg(cnt + 2 * lane + 1) : 0u;
		uint32_t incl = a + b;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) { incl += n; } }
		const uint32_t before = incl - (a + b);
		if (2 * lane < G) { cellRel[static_cast<size_t>(tile) * G + 2 * lane] = before; }
		if (2 * lane + 1 < G) { cellRel[static_cast<size_t>(tile) * G + 2 * lane + 1] = before + a; }
		if (lane == 31) { tileTotal[tile] = incl; } }
	tile_scan_last_block(tileTotal, tileBase, tileOrder, ntiles, fp.listCapacity, ctr); }

// ---------------------------------------------------------------------------------------------
// K5: list fill: one thread per triangle id (bin_item<FILL>); queued large items are skipped.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
fill_kernel(FrameParams fp, const uint2* __restrict__ triInfo, const ClipRec* __restrict__ clipRecs,
            BinArgs B, Counters* __restrict__ ctr) {
	pdl_launch_dependents();
	pdl_wait();
	const uint32_t nblocks = (fp.totalPJobs + 255u) >> 8;   // (capped grids: see vertex_kernel)
	for (uint32_t vblock = blockIdx.x; vblock < nblocks; vblock += gridDim.x) {
		const uint32_t job = vblock * blockDim.x + threadIdx.x;
		uint2 info = make_uint2(kReject, 0u);
		if (job < fp.totalPJobs) { info = triInfo[job]; }
		bin_triangle<true>(job, info, fp, clipRecs, B, ctr); } }

// ---------------------------------------------------------------------------------------------
// Split-frame presentation: completion counters instead of a host-side barrier.  Behind every frame's tile kernel
// a one-thread kernel adds 1 to this rank's counter in the presenting GPU's memory (stream order: the tile kernel,
// and with it its peer stores, has completed; a system fence, then a system-scope atomic over NVLink); a stream
// waits until all ranks' counters have reached a frame number.  (Signalling from the tile kernel's
// last CTA instead costs a system-scope fence per CTA: measured 20 % slower on the whole frame.)
// The wait is a bounded spin (2 s) so that a rank that died cannot hang the GPU.
// ---------------------------------------------------------------------------------------------

__global__ void signal_counter_kernel(unsigned long long* counter) {
	// (launched with programmatic stream serialization: resident early, then waits here until the kernels before it
	// in the stream -- the frame's tile kernel -- have completed and their stores are visible)
	pdl_launch_dependents();
	pdl_wait();
	if (threadIdx.x == 0) {
		__threadfence_system();
		atomicAdd_system(counter, 1ull); } }

__global__ void wait_counter_kernel(const unsigned long long* counters, unsigned int count, unsigned long long value, unsigned int* timedOut) {
	// lane i watches counters[i], i + 32, ...: done when every one of the `count` counters has reached `value`
	pdl_launch_dependents();
	pdl_wait();
	unsigned long long t0, now;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	unsigned int next = threadIdx.x;
	while (true) {
		while (next < count) {
			unsigned long long v;
			asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(counters + next) : "memory");
			if (v < value) { break; }
			next += 32u; }
		if (__all_sync(0xffffffffu, next >= count)) { break; }
		__nanosleep(200);
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
		if (__any_sync(0xffffffffu, now - t0 > 2000000000ull)) { if (timedOut && threadIdx.x == 0) { *timedOut = 1u; } break; } } }

}  // namespace rsr
