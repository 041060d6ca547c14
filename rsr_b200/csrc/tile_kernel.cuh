// tile_kernel.cuh -- K6: one CTA per 32x32 screen tile, one thread per 2x2 quad.
//
// The CTA keeps the tile's colour + depth in shared memory (SoA per quad lane: conflict free) for
// the whole frame: clear -> every draw's triangles in submission order -> resolve/stores.  It
// replaces GPU::DrawImpl (rglv_gpu.cxx:263-432) + GPUTileImpl::DrawTriangles / DrawClipped
// (rglv_gpu_impl.hxx:841-998) + VTriangleRasterizer / TriangleRasterizer (rglv_triangle.hxx)
// + TriangleProgram::Render (rglv_gpu_impl.hxx:166-222) + FilterTile (rglr_algorithm.hxx:31-104).
//
// Per draw the tile's list segment is consumed in batches of 256 triangles:
//   setup   thread t sets up triangle t: fixed-point vertices, int32 edge constants evaluated at
//           the reference's own start point, bbox, 1/area, depth and 1/w per vertex -> smem
//   filter  each warp (a 16x8 pixel region) ballots which of the 256 bboxes touch its region
//   raster  for each surviving triangle, in order, every thread evaluates the three edge
//           functions at its quad's four pixels; covered quads run depth test, perspective
//           interpolation and the fragment program, then update the tile in shared memory.
#pragma once
#include "kernels.cuh"

namespace rsr {

// ryg's float->sRGB8 table (3rdparty/ryg-srgb/ryg-srgb.h:71-85, public domain, Fabian Giesen);
// data only -- the reference resolves through it, so bit-exact output needs the same numbers.
__constant__ uint32_t kSrgbTab4[104] = {
	0x0073000d, 0x007a000d, 0x0080000d, 0x0087000d, 0x008d000d, 0x0094000d, 0x009a000d, 0x00a1000d,
	0x00a7001a, 0x00b4001a, 0x00c1001a, 0x00ce001a, 0x00da001a, 0x00e7001a, 0x00f4001a, 0x0101001a,
	0x010e0033, 0x01280033, 0x01410033, 0x015b0033, 0x01750033, 0x018f0033, 0x01a80033, 0x01c20033,
	0x01dc0067, 0x020f0067, 0x02430067, 0x02760067, 0x02aa0067, 0x02dd0067, 0x03110067, 0x03440067,
	0x037800ce, 0x03df00ce, 0x044600ce, 0x04ad00ce, 0x051400ce, 0x057b00c5, 0x05dd00bc, 0x063b00b5,
	0x06970158, 0x07420142, 0x07e30130, 0x087b0120, 0x090b0112, 0x09940106, 0x0a1700fc, 0x0a9500f2,
	0x0b0f01cb, 0x0bf401ae, 0x0ccb0195, 0x0d950180, 0x0e56016e, 0x0f0d015e, 0x0fbc0150, 0x10630143,
	0x11070264, 0x1238023e, 0x1357021d, 0x14660201, 0x156601e9, 0x165a01d3, 0x174401c0, 0x182401af,
	0x18fe0331, 0x1a9602fe, 0x1c1502d2, 0x1d7e02ad, 0x1ed4028d, 0x201a0270, 0x21520256, 0x227d0240,
	0x239f0443, 0x25c003fe, 0x27bf03c4, 0x29a10392, 0x2b6a0367, 0x2d1d0341, 0x2ebe031f, 0x304d0300,
	0x31d105b0, 0x34a80555, 0x37520507, 0x39d504c5, 0x3c37048b, 0x3e7c0458, 0x40a8042a, 0x42bd0401,
	0x44c20798, 0x488e071e, 0x4c1c06b6, 0x4f76065d, 0x52a50610, 0x55ac05cc, 0x5892058f, 0x5b590559,
	0x5e0c0a23, 0x631c0980, 0x67db08f6, 0x6c55087f, 0x70940818, 0x74a007bd, 0x787d076c, 0x7c330723,
};

// RSR_PHASE_PROF (developer build, tools/phase_probe.py): cycles thread 0 of every CTA spends per phase of the tile kernel
#ifdef RSR_PHASE_PROF
__device__ unsigned long long g_phaseCycles[16];
#define PHASE_BEGIN() long long phaseLast = clock64(); int phaseCur = 0
#define PHASE(k) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&g_phaseCycles[phaseCur], static_cast<unsigned long long>(now_ - phaseLast)); phaseLast = now_; phaseCur = (k); } } while (0)
#else
#define PHASE_BEGIN() do {} while (0)
#define PHASE(k) do {} while (0)
#endif

constexpr int kSortCap = 1024;           // most list entries sorted in one round
#ifndef RSR_RUN_CAP
#define RSR_RUN_CAP 1024
#endif
constexpr int kRunCap = RSR_RUN_CAP;     // most runs of one cell that the run merge handles

struct TileShared {
	float chan[4][4][kTileThreads];      // [r,g,b,depth][quad lane][thread]
	int ec[3][kBatch];                   // edge functions at the tile origin
	int edx[3][kBatch];
	int edy[3][kBatch];
	uint32_t bbox[kBatch];               // tile-local minx | miny<<6 | maxx<<12 | maxy<<18 | clipped<<24
	float scale[kBatch];
	float z[3][kBatch];
	float iw[3][kBatch];
	uint32_t vref[3][kBatch];            // float4 index of the vertex' varyings; bit 31 = clip buffer
	uint16_t state[kBatch];              // DevState index of the triangle's draw
	uint4 tex0[kBatch];                  // texture unit 0 of the triangle's draw: texels pointer, texel count, kind | power << 8 (TexDesc)
	uint32_t key[kBatch];                // DevDraw::batchKey of the triangle's draw (program | pipeline flags << 8)
	uint32_t runStartBits[kBatch / 32], tinyBits[kBatch / 32];   // per batch: entries that start a run of equal keys; tiny triangles
	// per-warp scratch: the queued rasteriser's (triangle, quad) work items awaiting shading (160 x uint16), or the direct
	// rasteriser's texel staging area (sample_quad, programs.cuh)
	float4 warpScratch[kTileThreads / 32][kStageTexels];
	alignas(16) uint16_t rcpLut[2048];   // rcpps table at 16 bits per entry (dev_math.cuh: rcp_entry16), staged once per CTA
	uint32_t sorted[kSortCap];           // triangle codes of the current chunk of the tile list, in submission order
	uint32_t cellOff[kMaxGroups + 1];    // list offsets of this tile's cells (clamped to the list capacity)
	uint32_t lgKey[kTileLargeCap], lgCode[kTileLargeCap];   // queued large items that cover this tile
	uint8_t lgGroup[kTileLargeCap];
	alignas(4) uint16_t lgPerGroup[kMaxGroups];
	int lgCount;
	uint32_t srgbTab[104];               // ryg table: per-pixel indices diverge, constant memory would serialise
	int firstBad;
	int sortCount;
	int headDraw;                        // draw and batch key of the first entry of the batch being formed
	uint32_t headKey;
};
static_assert(offsetof(TileShared, ec) % 8 == 0 && sizeof(int) * 9 * kBatch >= 8 * kSortCap, "sort scratch aliases ec/edx/edy");

constexpr int kInlineCmds = 6;

// what the clear / store commands read from their state snapshot
struct CmdState { float clearColor[4]; float clearDepth; int programId; float uniform0; int color0Type; };

struct TileArgs {
	FrameParams fp;
	// the first non-draw commands of the frame (usually all of them: a clear and a store or two) and the
	// state words they use travel as kernel parameters: no trips to memory for them in any tile CTA
	FrameCmd icmd[kInlineCmds];
	CmdState icmdState[kInlineCmds];
	const FrameCmd* cmds;
	const DevDraw* draws;
	const DevState* states;
	const ApproxLuts* luts;
	const float4* ptvb;
	const TriRec* triRecs;
	const ClipRec* clipRecs;
	const uint2* lists;                  // (order key, triangle code) entries, unordered within a tile
	const uint32_t* tileBase;            // [tile]: list offset of the tile's first cell, + the total at the end
	const uint32_t* cellRel;             // [tile * groups + g]: offset of the cell inside the tile's list (groups > 1)
	const LargeItem* large;              // queued large items (kernels.cuh)
	const uint32_t* tileOrder;           // CTA -> tile, tiles with the longest lists first (nullptr: identity)
	uint32_t* runScratch;                // [tile][2 x kRunCap]: run tables of a long list cell (load_chunk, mode B)
	Counters* ctr; };

__device__ __forceinline__ void prefetch_entry(const TileArgs& A, uint32_t id) {
	const void* p = (id & kFanIdBit) ? static_cast<const void*>(A.clipRecs + ((id & ~kFanIdBit) >> 3)) : static_cast<const void*>(A.triRecs + id);
	asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

__device__ __forceinline__ int cmd_before_draw(const TileArgs& A, int ci) {
	return (ci < kInlineCmds) ? A.icmd[ci].beforeDraw : A.cmds[ci].beforeDraw; }

__device__ __forceinline__ bool top_left(int dy, int dx) { return (dy > 0) || (dy == 0 && dx > 0); }

// Edge setup for one triangle of this tile.  `wide` selects the reference's 4-wide int32 path
// (VTriangleRasterizer::Draw, rglv_triangle.hxx:193-240: all products wrap) or the scalar int64
// path used for clipped triangles (TriangleRasterizer::Draw, :79-131).
__device__ __forceinline__ void setup_edges(TileShared& sh, int slot, bool wide, bool clipped,
                                            int x1, int x2, int x3, int y1, int y2, int y3,
                                            int ox, int oy, int rl, int rt, int rr, int rb) {
	int vminx = max(min(min(x1, x2), x3) >> 4, rl);
	const int vmaxx = min((max(max(x1, x2), x3) + 15) >> 4, rr);
	int vminy = max(min(min(y1, y2), y3) >> 4, rt);
	const int vmaxy = min((max(max(y1, y2), y3) + 15) >> 4, rb);
	vminx &= ~1;
	vminy &= ~1;

	const int lminx = max(vminx, ox) - ox, lmaxx = min(vmaxx, ox + kTile) - ox;
	const int lminy = max(vminy, oy) - oy, lmaxy = min(vmaxy, oy + kTile) - oy;
	if (lminx >= lmaxx || lminy >= lmaxy) { sh.bbox[slot] = 0; return; }

	const uint32_t ux1 = x1, ux2 = x2, ux3 = x3, uy1 = y1, uy2 = y2, uy3 = y3;
	const int dx12 = static_cast<int>(ux1 - ux2), dy12 = static_cast<int>(uy2 - uy1);
	const int dx23 = static_cast<int>(ux2 - ux3), dy23 = static_cast<int>(uy3 - uy2);
	const int dx31 = static_cast<int>(ux3 - ux1), dy31 = static_cast<int>(uy1 - uy3);
	int c1, c2, c3;
	float scale;
	if (wide) {
		const uint32_t sx = (static_cast<uint32_t>(vminx) << 4) + 8u;
		const uint32_t sy = (static_cast<uint32_t>(vminy) << 4) + 8u;
		uint32_t u1 = static_cast<uint32_t>(dy12) * (sx - ux1) + static_cast<uint32_t>(dx12) * (sy - uy1);
		uint32_t u2 = static_cast<uint32_t>(dy23) * (sx - ux2) + static_cast<uint32_t>(dx23) * (sy - uy2);
		uint32_t u3 = static_cast<uint32_t>(dy31) * (sx - ux3) + static_cast<uint32_t>(dx31) * (sy - uy3);
		u1 = u1 + (top_left(dy12, dx12) ? 1u : 0u) - 1u;
		u2 = u2 + (top_left(dy23, dx23) ? 1u : 0u) - 1u;
		u3 = u3 + (top_left(dy31, dx31) ? 1u : 0u) - 1u;
		c1 = static_cast<int>(u1) >> 4;
		c2 = static_cast<int>(u2) >> 4;
		c3 = static_cast<int>(u3) >> 4;
		const int sum = static_cast<int>(static_cast<uint32_t>(c1) + static_cast<uint32_t>(c2) + static_cast<uint32_t>(c3));
		scale = 1.0f / itof(sum); }
	else {
		const long long ldx12 = static_cast<long long>(x1) - x2, ldy12 = static_cast<long long>(y2) - y1;
		const long long ldx23 = static_cast<long long>(x2) - x3, ldy23 = static_cast<long long>(y3) - y2;
		const long long ldx31 = static_cast<long long>(x3) - x1, ldy31 = static_cast<long long>(y1) - y3;
		const long long sx = (static_cast<long long>(vminx) << 4) + 8;
		const long long sy = (static_cast<long long>(vminy) << 4) + 8;
		long long l1 = ldy12 * (sx - x1) + ldx12 * (sy - y1);
		long long l2 = ldy23 * (sx - x2) + ldx23 * (sy - y2);
		long long l3 = ldy31 * (sx - x3) + ldx31 * (sy - y3);
		if (ldy12 > 0 || (ldy12 == 0 && ldx12 > 0)) { l1++; } --l1;
		if (ldy23 > 0 || (ldy23 == 0 && ldx23 > 0)) { l2++; } --l2;
		if (ldy31 > 0 || (ldy31 == 0 && ldx31 > 0)) { l3++; } --l3;
		l1 >>= 4; l2 >>= 4; l3 >>= 4;
		c1 = static_cast<int>(l1); c2 = static_cast<int>(l2); c3 = static_cast<int>(l3);
		scale = 1.0f / __ll2float_rn(l1 + l2 + l3); }

	// move the start point from the reference's (vminx, vminy) to this tile's origin
	const uint32_t mx = static_cast<uint32_t>(ox - vminx), my = static_cast<uint32_t>(oy - vminy);
	sh.ec[0][slot] = static_cast<int>(static_cast<uint32_t>(c1) + mx * static_cast<uint32_t>(dy12) + my * static_cast<uint32_t>(dx12));
	sh.ec[1][slot] = static_cast<int>(static_cast<uint32_t>(c2) + mx * static_cast<uint32_t>(dy23) + my * static_cast<uint32_t>(dx23));
	sh.ec[2][slot] = static_cast<int>(static_cast<uint32_t>(c3) + mx * static_cast<uint32_t>(dy31) + my * static_cast<uint32_t>(dx31));
	sh.edx[0][slot] = dx12; sh.edx[1][slot] = dx23; sh.edx[2][slot] = dx31;
	sh.edy[0][slot] = dy12; sh.edy[1][slot] = dy23; sh.edy[2][slot] = dy31;
	sh.scale[slot] = scale;
	sh.bbox[slot] = static_cast<uint32_t>(lminx) | (static_cast<uint32_t>(lminy) << 6) |
	                (static_cast<uint32_t>(lmaxx) << 12) | (static_cast<uint32_t>(lmaxy) << 18) |
	                (clipped ? (1u << 24) : 0u); }

// The 80-byte triangle record (or, for a clip fan triangle, the owner words of its clip record) as the
// tile kernel carries it in registers between "who owns this entry" and "set it up": ONE trip to
// memory per list entry instead of two.
struct EntryRec { uint4 q0, q1, q2, q3, q4; };

__device__ __forceinline__ EntryRec load_entry(const TileArgs& A, uint32_t id) {
	EntryRec e;
	if (id & kFanIdBit) {
		const ClipRec& rec = A.clipRecs[(id & ~kFanIdBit) >> 3];
		e.q0 = e.q1 = e.q2 = make_uint4(0u, 0u, 0u, 0u);
		e.q3 = make_uint4(0u, 0u, 0u, rec.draw);
		e.q4 = make_uint4(rec.key, rec.state & 0xffffu, 0u, 0u); }
	else {
		const uint4* r = reinterpret_cast<const uint4*>(A.triRecs + id);
		e.q0 = __ldg(r); e.q1 = __ldg(r + 1); e.q2 = __ldg(r + 2); e.q3 = __ldg(r + 3); e.q4 = __ldg(r + 4); }
	return e; }

__device__ __forceinline__ void setup_triangle(TileShared& sh, int slot, uint32_t id, const EntryRec& e, const TileArgs& A,
                                               int ox, int oy, int rl, int rt, int rr, int rb) {
	if (!(id & kFanIdBit)) {
		// GPUTileImpl::DrawTriangles (rglv_gpu_impl.hxx:880-946): everything was prepared by K2
		const uint4 q0 = e.q0, q1 = e.q1, q2 = e.q2, q3 = e.q3;
		sh.z[0][slot] = __uint_as_float(q1.z); sh.z[1][slot] = __uint_as_float(q1.w); sh.z[2][slot] = __uint_as_float(q2.x);
		sh.iw[0][slot] = __uint_as_float(q2.y); sh.iw[1][slot] = __uint_as_float(q2.z); sh.iw[2][slot] = __uint_as_float(q2.w);
		sh.vref[0][slot] = q3.x; sh.vref[1][slot] = q3.y; sh.vref[2][slot] = q3.z;
		setup_edges(sh, slot, true, false, static_cast<int>(q0.x), static_cast<int>(q0.y), static_cast<int>(q0.z),
		            static_cast<int>(q0.w), static_cast<int>(q1.x), static_cast<int>(q1.y), ox, oy, rl, rt, rr, rb); }
	else {
		// GPUTileImpl::DrawClipped (rglv_gpu_impl.hxx:949-998)
		const uint32_t recIdx = (id & ~kFanIdBit) >> 3, k = id & 7u;
		const ClipRec& rec = A.clipRecs[recIdx];
		const uint32_t base = recIdx * static_cast<uint32_t>(sizeof(ClipRec) / 16) + 3u;   // float4 index of v[0]
		const uint32_t vsz = sizeof(ClipVertex) / 16;
		const uint32_t j0 = 0, j1 = k + 1, j2 = k + 2;
		const float4 v0 = rec.v[j0].dev, v1 = rec.v[j1].dev, v2 = rec.v[j2].dev;
		sh.z[0][slot] = v0.z; sh.z[1][slot] = v1.z; sh.z[2][slot] = v2.z;
		sh.iw[0][slot] = v0.w; sh.iw[1][slot] = v1.w; sh.iw[2][slot] = v2.w;
		sh.vref[0][slot] = 0x80000000u | (base + j0 * vsz + 1u);
		sh.vref[1][slot] = 0x80000000u | (base + j1 * vsz + 1u);
		sh.vref[2][slot] = 0x80000000u | (base + j2 * vsz + 1u);
		setup_edges(sh, slot, false, true,
		            cvtt(v0.x * 16.0f), cvtt(v1.x * 16.0f), cvtt(v2.x * 16.0f),
		            cvtt(v0.y * 16.0f), cvtt(v1.y * 16.0f), cvtt(v2.y * 16.0f), ox, oy, rl, rt, rr, rb); } }

// stages texture unit 0 of the slot's draw (TexDesc) next to the triangle record
__device__ __forceinline__ void setup_texture(TileShared& sh, int slot, const TileArgs& A, uint32_t state) {
	const TexUnit& tu = A.states[state].tu[0];
	const uintptr_t p = reinterpret_cast<uintptr_t>(tu.texels);
	sh.tex0[slot] = make_uint4(static_cast<uint32_t>(p), static_cast<uint32_t>(p >> 32), tu.texelCount,
	                           static_cast<uint32_t>(tu.kind & 0xff) | (static_cast<uint32_t>(tu.power) << 8)); }

__device__ __forceinline__ bool depth_pass(int func, float frag, float dest) {
	return func == 0 ? (frag < dest) : (func == 1 ? (frag <= dest) : (frag == dest)); }

// FragmentStateKey bits (rglv_gpu.hxx, keyOf in rsrcu.cu) as seen by the tile kernel
constexpr uint32_t kKeyDepthTest = 2u, kKeyBlend = 16u, kKeyDepthWrite = 32u, kKeyColorWrite = 64u;
// the pipeline every bundled scene uses: depth test LESS, depth + colour writes, no blending
constexpr uint32_t kKeyFastMask = 0x7eu, kKeyFastValue = 0x62u;

// TriangleProgram::Render for one quad (rglv_gpu_impl.hxx:166-222).
// FAST: the pipeline flags are the compile-time combination above; otherwise `flags` (uniform for
// the batch) is decoded at run time.  The four lanes of the reference's SSE registers are two
// packed pairs here: (0,1) and (2,3).
#ifndef RSR_STAGE_MIN_LANES
#define RSR_STAGE_MIN_LANES 12
#endif
#ifndef RSR_COOP_STAGE
#define RSR_COOP_STAGE 0   // 1: the direct rasteriser stages texel boxes in shared memory (sample_quad); measured slower than gathers, see profiles/README.md
#endif
// COOP: all 32 lanes of the warp are in the call (the direct rasteriser: one triangle, every lane its own quad), lanes
// whose quad is not covered with triMask = 0; they skip the arithmetic but take part in the cooperative texel staging.
template <class P, bool FAST, bool COOP>
__device__ __forceinline__ unsigned render_quad(TileShared& sh, const int t, const int i, const TileArgs& A, const uint32_t flags,
                                                const DevState& s, const int (&e1)[4], const int (&e2)[4], const uint32_t triMask,
                                                const int px, const int py, const bool clipped) {
	const bool depthTest = FAST ? true : (flags & kKeyDepthTest) != 0;
	const int depthFunc = FAST ? 0 : static_cast<int>((flags >> 2) & 3u);
	const bool depthWrite = FAST ? true : (flags & kKeyDepthWrite) != 0;
	const bool colorWrite = FAST ? true : (flags & kKeyColorWrite) != 0;
	const bool blend = FAST ? false : (flags & kKeyBlend) != 0;

	const float scale = sh.scale[i];
	const f2 one = one2();
	f2 BSx[2], BSy[2], BSz[2];
	float fragDepth[4];
	const float z0 = sh.z[0][i], z1 = sh.z[1][i], z2 = sh.z[2][i];
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		BSx[h] = mul2(mk2(itof(e2[2 * h]), itof(e2[2 * h + 1])), scale);
		BSz[h] = mul2(mk2(itof(e1[2 * h]), itof(e1[2 * h + 1])), scale);
		BSy[h] = sub2(sub2(one, BSx[h]), BSz[h]);
		const f2 fd = add2(add2(mul2(BSx[h], z0), mul2(BSy[h], z1)), mul2(BSz[h], z2));
		fragDepth[2 * h] = lo2(fd); fragDepth[2 * h + 1] = hi2(fd); }

	uint32_t fragMask = triMask;
	if (P::earlyZ && depthTest) {
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			const float dest = sh.chan[3][l][t];
			const bool pass = FAST ? (fragDepth[l] < dest) : depth_pass(depthFunc, fragDepth[l], dest);
			if (!pass) { fragMask &= ~(1u << l); } } }
	// (COOP: texel staging pays when the triangle fills a good part of the warp's region; a triangle that touches a few
	// quads samples with plain gathers)
	unsigned activeLanes = 0;
	if (COOP) { activeLanes = __ballot_sync(0xffffffffu, fragMask != 0); if (activeLanes == 0) { return 0; } }
	else if (fragMask == 0) { return 0; }
	if (P::earlyZ && depthWrite) {
#pragma unroll
		for (int l = 0; l < 4; ++l) { if (fragMask & (1u << l)) { sh.chan[3][l][t] = fragDepth[l]; } } }

	// perspective-correct barycentrics
	FragIn f;
	f.st = &s;
	if constexpr (P::samples) {
		const uint4 td = sh.tex0[i];
		f.tex0.texels = reinterpret_cast<const float4*>(static_cast<uintptr_t>(td.x) | (static_cast<uintptr_t>(td.y) << 32));
		f.tex0.texelCount = td.z; f.tex0.kind = static_cast<int>(td.w & 0xffu); f.tex0.power = static_cast<int>(td.w >> 8); }
	f.rcpLut = A.luts->rcp;
	f.rsqrtLut = A.luts->rsqrt;
	f.stage = (COOP && __popc(activeLanes) >= RSR_STAGE_MIN_LANES) ? sh.warpScratch[threadIdx.x >> 5] : nullptr;
	const float iw0 = sh.iw[0][i], iw1 = sh.iw[1][i], iw2 = sh.iw[2][i];
	f2 BPx[2], BPy[2], BPz[2], wsum[2], wx[2], wz[2];
	float rcp[4];
	bool ok = true;
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		wx[h] = mul2(BSx[h], iw0);
		wz[h] = mul2(BSz[h], iw2);
		wsum[h] = add2(add2(wx[h], mul2(BSy[h], iw1)), wz[h]);
		rcp[2 * h] = rcp_fast16(lo2(wsum[h]), sh.rcpLut, ok);
		rcp[2 * h + 1] = rcp_fast16(hi2(wsum[h]), sh.rcpLut, ok); }
	if (!ok) {
#pragma unroll
		for (int h = 0; h < 2; ++h) {
			rcp[2 * h] = rcp_intel16(lo2(wsum[h]), sh.rcpLut);
			rcp[2 * h + 1] = rcp_intel16(hi2(wsum[h]), sh.rcpLut); } }
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		// oneover (rmlv_mvec4.hxx:630-650): r = rcpps(a); (r + r) - a * (r * r)
		const f2 r = mk2(rcp[2 * h], rcp[2 * h + 1]);
		const f2 fragW = sub2(add2(r, r), mul2(wsum[h], mul2(r, r)));
		BPx[h] = mul2(wx[h], fragW);
		BPz[h] = mul2(wz[h], fragW);
		BPy[h] = sub2(sub2(one, BPx[h]), BPz[h]);
		f.BPx[2 * h] = lo2(BPx[h]); f.BPx[2 * h + 1] = hi2(BPx[h]);
		f.BPy[2 * h] = lo2(BPy[h]); f.BPy[2 * h + 1] = hi2(BPy[h]);
		f.BPz[2 * h] = lo2(BPz[h]); f.BPz[2 * h + 1] = hi2(BPz[h]); }
#pragma unroll
	for (int l = 0; l < 4; ++l) {
		f.depth[l] = fragDepth[l];
		f.fragX[l] = (itof(px) + 0.5f) + static_cast<float>(l & 1);
		// rglv_triangle.hxx:297 (4-wide) vs :160 (scalar, clipped triangles): the two differ by one row
		f.fragY[l] = clipped ? ((itof(A.fp.height - py) - 0.5f) + ((l & 2) ? 0.0f : 1.0f))
		                     : (((itof(A.fp.height) - 0.5f) - itof(py)) - ((l & 2) ? 1.0f : 0.0f)); }

	// varyings: Interpolants::Interpolate(BS, BP) -- every reference program uses BP
	float at[kMaxVaryings][4];
	if constexpr (P::NV > 0) {
		const uint32_t r0 = sh.vref[0][i], r1 = sh.vref[1][i], r2 = sh.vref[2][i];
		const float4* base = (r0 & 0x80000000u) ? reinterpret_cast<const float4*>(A.clipRecs) : A.ptvb;
		const float4* p0 = base + (r0 & 0x7fffffffu);
		const float4* p1 = base + (r1 & 0x7fffffffu);
		const float4* p2 = base + (r2 & 0x7fffffffu);
#pragma unroll
		for (int k4 = 0; k4 < (P::NV + 3) / 4; ++k4) {
			const float4 a = __ldg(p0 + k4), b = __ldg(p1 + k4), c = __ldg(p2 + k4);
			const float av[4] = { a.x, a.y, a.z, a.w }, bv[4] = { b.x, b.y, b.z, b.w }, cv[4] = { c.x, c.y, c.z, c.w };
#pragma unroll
			for (int c4 = 0; c4 < 4; ++c4) {
				const int k = k4 * 4 + c4;
				if (k < P::NV) {
#pragma unroll
					for (int h = 0; h < 2; ++h) {
						const f2 v = add2(add2(mul2(BPx[h], av[c4]), mul2(BPy[h], bv[c4])), mul2(BPz[h], cv[c4]));
						at[k][2 * h] = lo2(v); at[k][2 * h + 1] = hi2(v); } } } } }

	float cr[4], cg[4], cb[4], ca[4];
	P::ShadeFragment(f, at, cr, cg, cb, ca, fragMask);

	if (!P::earlyZ && depthTest) {
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			const float dest = sh.chan[3][l][t];
			const bool pass = FAST ? (fragDepth[l] < dest) : depth_pass(depthFunc, fragDepth[l], dest);
			if (!pass) { fragMask &= ~(1u << l); } }
		if (fragMask == 0) { return 0; } }
	if (!P::earlyZ && depthWrite) {
#pragma unroll
		for (int l = 0; l < 4; ++l) { if (fragMask & (1u << l)) { sh.chan[3][l][t] = fragDepth[l]; } } }

	if (colorWrite) {
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			if (fragMask & (1u << l)) {
				if (blend) {
					// BlendAlpha (rglv_gpu_impl.hxx:80-84)
					const float alpha = ca[l];
					const float oma = 1.0f - alpha;
					sh.chan[0][l][t] = cr[l] * alpha + sh.chan[0][l][t] * oma;
					sh.chan[1][l][t] = cg[l] * alpha + sh.chan[1][l][t] * oma;
					sh.chan[2][l][t] = cb[l] * alpha + sh.chan[2][l][t] * oma; }
				else {
					sh.chan[0][l][t] = cr[l];
					sh.chan[1][l][t] = cg[l];
					sh.chan[2][l][t] = cb[l]; } } } }
	return __popc(fragMask); }

// Can triangle ti cover any pixel of the region [x0, x1] x [y0, y1] (tile-local pixel coordinates, inclusive; whole quads: x0, y0 even, x1, y1 odd)?  Each
// edge function is evaluated at the region corner where it is largest; if it is negative there it is negative at
// every pixel of the region and the per-quad tests can be skipped.  The edge functions live in wrapping 32-bit
// arithmetic (the reference's 4-wide setup, rglv_triangle.hxx:216-238): the shortcut is taken only when the
// corner value is far enough from INT_MIN that no pixel of the region can have wrapped past it (a pixel's value is
// the corner's minus less than 2^23), so the answer is exactly what the per-pixel tests would give.
__device__ __forceinline__ bool region_may_be_covered(const TileShared& sh, int ti, int x0, int y0, int x1, int y1) {
#pragma unroll
	for (int e = 0; e < 3; ++e) {
		const int dy = sh.edy[e][ti], dx = sh.edx[e][ti];
		const uint32_t cx = static_cast<uint32_t>(dy > 0 ? x1 : x0), cy = static_cast<uint32_t>(dx > 0 ? y1 : y0);
		const int emax = static_cast<int>(static_cast<uint32_t>(sh.ec[e][ti]) + cx * static_cast<uint32_t>(dy) + cy * static_cast<uint32_t>(dx));
		if (emax < 0 && emax > static_cast<int>(0x80000000u) + (1 << 23) && abs(dy) < (1 << 17) && abs(dx) < (1 << 17)) { return false; } }
	return true; }

// edge functions of triangle ti at the four pixels of the quad whose tile-local origin is (lx, ly);
// returns the coverage mask (bit l = lane l inside all three edges); 0 if the quad is outside the
// triangle's bbox (the reference never visits it)
__device__ __forceinline__ uint32_t quad_coverage(const TileShared& sh, int ti, int lx, int ly, int (&e1)[4], int (&e2)[4]) {
	const uint32_t bb = sh.bbox[ti];
	const int minx = bb & 63, miny = (bb >> 6) & 63, maxx = (bb >> 12) & 63, maxy = (bb >> 18) & 63;
	if (lx < minx || lx >= maxx || ly < miny || ly >= maxy) { return 0; }
	const uint32_t ulx = lx, uly = ly;
	const uint32_t dy1 = sh.edy[0][ti], dx1 = sh.edx[0][ti];
	const uint32_t dy2 = sh.edy[1][ti], dx2 = sh.edx[1][ti];
	const uint32_t dy3 = sh.edy[2][ti], dx3 = sh.edx[2][ti];
	const uint32_t a1 = static_cast<uint32_t>(sh.ec[0][ti]) + ulx * dy1 + uly * dx1;
	const uint32_t a2 = static_cast<uint32_t>(sh.ec[1][ti]) + ulx * dy2 + uly * dx2;
	const uint32_t a3 = static_cast<uint32_t>(sh.ec[2][ti]) + ulx * dy3 + uly * dx3;
	const uint32_t q1[4] = { a1, a1 + dy1, a1 + dx1, a1 + dx1 + dy1 };
	const uint32_t q2[4] = { a2, a2 + dy2, a2 + dx2, a2 + dx2 + dy2 };
	const uint32_t q3[4] = { a3, a3 + dy3, a3 + dx3, a3 + dx3 + dy3 };
	uint32_t covered = 0;
#pragma unroll
	for (int l = 0; l < 4; ++l) {
		e1[l] = static_cast<int>(q1[l]);
		e2[l] = static_cast<int>(q2[l]);
		if (static_cast<int>(q1[l] | q2[l] | q3[l]) >= 0) { covered |= (1u << l); } }
	return covered; }

// Rasterises the nb triangles that setup_triangle placed in shared memory, in order.
//
// Each warp owns a 16x8-pixel region (32 quads, quad q "belongs" to lane q).  Coverage and shading
// are decoupled so that small triangles do not leave 30 of 32 lanes idle while one quad is shaded:
//   produce  for each triangle whose bbox touches the region (ballot), every lane tests its own
//            quad; covered quads are appended, in triangle order, to a per-warp queue of
//            (triangle, quad) work items;
//   consume  whenever 32 items are queued (or at the end) lane i shades item i -- any triangle, any
//            quad of the region.  Items aimed at the same quad must retire in queue order
//            (depth LESS ties, blending): __match_any_sync groups them and the group is replayed
//            rank by rank.  With little overdraw inside 32 consecutive items that is one pass.
template <class P, bool FAST>
__device__ __forceinline__ unsigned draw_batch_queued(TileShared& sh, const TileArgs& A, const uint32_t flags, const int base, int nb, int ox, int oy) {
	const int t = threadIdx.x;
	const int warp = t >> 5, lane = t & 31;
	const unsigned ltMask = (1u << lane) - 1u;
	const int rx = (warp & 1) * 16, ry = (warp >> 1) * 8;            // region origin (tile-local)
	const int lx = rx + (lane & 7) * 2, ly = ry + (lane >> 3) * 2;   // own quad origin
	uint16_t* queue = reinterpret_cast<uint16_t*>(sh.warpScratch[warp]);
	unsigned frags = 0;
	int qn = 0;              // queued items
	int k = 0, kbase = 0;    // next triangle group / base of the current one
	unsigned pending = 0;    // triangles of the current group that touch the region, not yet tested
	unsigned smallMask = 0;  // ... of which the part inside the region is at most 2x2 quads
	// bbox of triangle (kbase + lane) clipped to the region, in region quad units (valid if pending bit set)
	int bqx0 = 0, bqy0 = 0, bqx1 = 0, bqy1 = 0;

	while (true) {
		// ---- produce ------------------------------------------------------------------------
		while (qn < 32) {
			if (pending == 0) {
				if (k >= nb) { break; }
				const int i = base + k + lane;
				bool hit = false, small = false;
				if (k + lane < nb) {
					const uint32_t bb = sh.bbox[i];
					const int minx = bb & 63, miny = (bb >> 6) & 63, maxx = (bb >> 12) & 63, maxy = (bb >> 18) & 63;
					hit = (minx < rx + 16) && (maxx > rx) && (miny < ry + 8) && (maxy > ry);
					hit = hit && region_may_be_covered(sh, i, max(minx, rx), max(miny, ry), (min(maxx, rx + 16) - 1) | 1, (min(maxy, ry + 8) - 1) | 1);
					if (hit) {
						bqx0 = (max(minx, rx) - rx) >> 1; bqx1 = (min(maxx, rx + 16) - 1 - rx) >> 1;
						bqy0 = (max(miny, ry) - ry) >> 1; bqy1 = (min(maxy, ry + 8) - 1 - ry) >> 1;
						small = (bqx1 - bqx0 <= 1) && (bqy1 - bqy0 <= 1); } }
				pending = __ballot_sync(0xffffffffu, hit);
				smallMask = __ballot_sync(0xffffffffu, small);
				kbase = base + k;
				k += 32;
				continue; }
			const int j = __ffs(pending) - 1;
			if ((smallMask >> j) & 1u) {
				// run of small triangles: one lane per triangle, each tests its <= 2x2 quads
				const unsigned large = pending & ~smallMask;
				const unsigned run = large ? (pending & ((1u << (__ffs(large) - 1)) - 1u)) : pending;
				pending &= ~run;
				const bool mine = (run >> lane) & 1u;
				const int ti = kbase + lane;
				unsigned cov = 0;   // bit c: candidate quad c = (dy*2 + dx) of the 2x2 block is covered
				if (mine) {
#pragma unroll
					for (int c = 0; c < 4; ++c) {
						const int qx = bqx0 + (c & 1), qy = bqy0 + (c >> 1);
						if (qx <= bqx1 && qy <= bqy1) {
							int e1[4], e2[4];
							if (quad_coverage(sh, ti, rx + qx * 2, ry + qy * 2, e1, e2)) { cov |= 1u << c; } } } }
				const int cnt = __popc(cov);
				int incl = cnt;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) { incl += v; } }
				int pos = qn + incl - cnt;
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					if ((cov >> c) & 1u) {
						const int ql = (bqy0 + (c >> 1)) * 8 + bqx0 + (c & 1);
						queue[pos++] = static_cast<uint16_t>((ti << 5) | ql); } }
				qn += __shfl_sync(0xffffffffu, incl, 31);
				__syncwarp();
				continue; }
			pending &= pending - 1;
			const int ti = kbase + j;
			int e1[4], e2[4];
			const bool covered = quad_coverage(sh, ti, lx, ly, e1, e2) != 0;
			const unsigned cm = __ballot_sync(0xffffffffu, covered);
			if (cm) {
				if (covered) { queue[qn + __popc(cm & ltMask)] = static_cast<uint16_t>((ti << 5) | lane); }
				qn += __popc(cm);
				__syncwarp(); } }
		if (qn == 0) { break; }

		// ---- consume up to 32 items ------------------------------------------------------------
		const int n = min(qn, 32);
		const bool have = lane < n;
		const unsigned item = have ? queue[lane] : 0u;
		const int ti = static_cast<int>(item >> 5), ql = static_cast<int>(item & 31u);
		const unsigned peers = __match_any_sync(0xffffffffu, have ? ql : (32 + lane));
		const int rank = __popc(peers & ltMask);
		const int maxRank = static_cast<int>(__reduce_max_sync(0xffffffffu, have ? static_cast<unsigned>(rank) : 0u));
		const int qx = rx + (ql & 7) * 2, qy = ry + (ql >> 3) * 2;
		for (int r = 0; r <= maxRank; ++r) {
			if (have && rank == r) {
				int e1[4], e2[4];
				const uint32_t covered = quad_coverage(sh, ti, qx, qy, e1, e2);
				frags += render_quad<P, FAST, false>(sh, warp * 32 + ql, ti, A, flags, A.states[sh.state[ti]], e1, e2, covered, ox + qx, oy + qy,
				                        (sh.bbox[ti] >> 24) & 1u); }
			__syncwarp(); }
		// keep the items that did not fit (the queue holds at most 31 + 128)
		const int rest = qn - n;
		for (int base = 0; base < rest; base += 32) {
			unsigned carry = 0;
			if (base + lane < rest) { carry = queue[32 + base + lane]; }
			__syncwarp();
			if (base + lane < rest) { queue[base + lane] = static_cast<uint16_t>(carry); }
			__syncwarp(); }
		qn = rest; }
	return frags; }

// Direct variant: every lane tests and shades its own quad, triangle by triangle.  No queue
// traffic; best when triangles are large (most lanes covered) or the batch is short.
template <class P, bool FAST>
__device__ __forceinline__ unsigned draw_batch_direct(TileShared& sh, const TileArgs& A, const uint32_t flags, const int base, int nb, int ox, int oy) {
	const int t = threadIdx.x;
	const int warp = t >> 5, lane = t & 31;
	const int rx = (warp & 1) * 16, ry = (warp >> 1) * 8;
	const int lx = rx + (lane & 7) * 2, ly = ry + (lane >> 3) * 2;
	unsigned frags = 0;
	for (int k = 0; k < nb; k += 32) {
		const int i = base + k + lane;
		bool hit = false;
		if (k + lane < nb) {
			const uint32_t bb = sh.bbox[i];
			const int minx = bb & 63, miny = (bb >> 6) & 63, maxx = (bb >> 12) & 63, maxy = (bb >> 18) & 63;
			hit = (minx < rx + 16) && (maxx > rx) && (miny < ry + 8) && (maxy > ry);
			hit = hit && region_may_be_covered(sh, i, max(minx, rx), max(miny, ry), (min(maxx, rx + 16) - 1) | 1, (min(maxy, ry + 8) - 1) | 1); }
		unsigned m = __ballot_sync(0xffffffffu, hit);
		while (m) {
			const int j = __ffs(m) - 1;
			m &= m - 1;
			const int ti = base + k + j;
#if RSR_COOP_STAGE
			int e1[4] = {0, 0, 0, 0}, e2[4] = {0, 0, 0, 0};
			const uint32_t covered = quad_coverage(sh, ti, lx, ly, e1, e2);
			if (!__any_sync(0xffffffffu, covered != 0)) { continue; }
			// (all 32 lanes go in, uncovered ones with an empty mask: the texel staging of sample_quad is a warp-wide effort)
			frags += render_quad<P, FAST, true>(sh, t, ti, A, flags, A.states[sh.state[ti]], e1, e2, covered, ox + lx, oy + ly, (sh.bbox[ti] >> 24) & 1u); } }
#else
			int e1[4], e2[4];
			const uint32_t covered = quad_coverage(sh, ti, lx, ly, e1, e2);
			if (covered == 0) { continue; }
			frags += render_quad<P, FAST, false>(sh, t, ti, A, flags, A.states[sh.state[ti]], e1, e2, covered, ox + lx, oy + ly, (sh.bbox[ti] >> 24) & 1u); } }
#endif
	return frags; }

// picks the variant per batch: long batches of small triangles go through the work queue; the
// common pipeline state gets the specialised instantiation
template <class P>
__device__ __forceinline__ unsigned draw_batch(TileShared& sh, const TileArgs& A, uint32_t key, int base, int nb, int ox, int oy, bool queued) {
	const uint32_t flags = key >> 8;
	const bool fast = (flags & kKeyFastMask) == kKeyFastValue;
	if (queued) {
		return fast ? draw_batch_queued<P, true>(sh, A, flags, base, nb, ox, oy) : draw_batch_queued<P, false>(sh, A, flags, base, nb, ox, oy); }
	return fast ? draw_batch_direct<P, true>(sh, A, flags, base, nb, ox, oy) : draw_batch_direct<P, false>(sh, A, flags, base, nb, ox, oy); }

// program dispatch for one batch of set-up triangles in slots [base, base + nb).  PROGS: the programs this instantiation
// of the tile kernel carries (bit = prog_bit(id)); a frame is launched on the smallest instantiation that holds its programs
__host__ __device__ constexpr uint32_t prog_bit(int id) {
	return id == ProgAmy::id ? 1u : id == ProgAlphaTexture::id ? 2u : id == ProgText::id ? 4u : id == ProgDepth::id ? 8u :
	       id == ProgPattern::id ? 16u : id == ProgMany::id ? 32u : id == ProgOBJ1::id ? 64u : id == ProgOBJ2::id ? 128u :
	       id == ProgOBJ2S::id ? 256u : id == ProgEnvmap::id ? 512u : id == ProgWireframe::id ? 1024u : id == ProgBase::id ? 2048u : 0u; }
constexpr uint32_t kAllProgs = 0xfffu;
// programs whose fragment stage samples texture unit 0 (P::samples)
constexpr uint32_t kSamplingProgs = prog_bit(ProgAmy::id) | prog_bit(ProgAlphaTexture::id) | prog_bit(ProgText::id) | prog_bit(ProgPattern::id) | prog_bit(ProgEnvmap::id);

// RSR_PREFETCH_VARYINGS=1: the thread that sets a triangle up also asks for its three vertices' varyings (ptvb) in L1.  The
// rasteriser walks a run triangle by triangle and the varyings of each one are the first thing render_quad waits for -- an
// L2 round trip per triangle and warp on the critical path of the warp that owns the batch's triangles; with the lines
// already on their way (one request per triangle and CTA instead of one per warp) that wait shrinks to an L1 hit.
#ifndef RSR_PREFETCH_VARYINGS
#define RSR_PREFETCH_VARYINGS 1
#endif
// float4s of varyings per vertex of program `id` among the programs of this instantiation (0: none / not carried)
template <uint32_t PROGS>
__device__ __forceinline__ int varying_f4s(uint32_t id) {
	int n = 0;
#define RSR_NV_CASE(P) if constexpr ((PROGS & prog_bit(P::id)) != 0u && P::NV > 0) { if (id == static_cast<uint32_t>(P::id)) { n = (P::NV + 3) / 4; } }
	RSR_NV_CASE(ProgAmy) RSR_NV_CASE(ProgOBJ2) RSR_NV_CASE(ProgMany) RSR_NV_CASE(ProgAlphaTexture) RSR_NV_CASE(ProgText) RSR_NV_CASE(ProgDepth)
	RSR_NV_CASE(ProgPattern) RSR_NV_CASE(ProgOBJ1) RSR_NV_CASE(ProgOBJ2S) RSR_NV_CASE(ProgEnvmap) RSR_NV_CASE(ProgWireframe)
#undef RSR_NV_CASE
	return n; }

template <uint32_t PROGS>
__device__ __forceinline__ void prefetch_varyings(const TileArgs& A, uint32_t id, const EntryRec& e) {
#if RSR_PREFETCH_VARYINGS
	if (id & kFanIdBit) { return; }   // (a clip fan's vertices live in its clip record, which setup_triangle reads anyway)
	const int n = varying_f4s<PROGS>(e.q4.x & 0xffu);
	if (n == 0) { return; }
	// <= 4 float4s of a vertex lie in at most two (three for 64 bytes) 32-byte sectors: the first and the last float4 name them
	const uint32_t r[3] = { e.q3.x, e.q3.y, e.q3.z };
#pragma unroll
	for (int v = 0; v < 3; ++v) {
		const float4* p = A.ptvb + r[v];
		asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
		if (n > 1) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p + (n - 1))); } }
#endif
}

template <uint32_t PROGS, class P>
__device__ __forceinline__ bool draw_batch_if(unsigned& frags, TileShared& sh, const TileArgs& A, uint32_t key0, int base, int nb, int ox, int oy, bool queued) {
	if constexpr ((PROGS & prog_bit(P::id)) != 0u) {
		if ((key0 & 0xffu) == static_cast<uint32_t>(P::id)) { frags = draw_batch<P>(sh, A, key0, base, nb, ox, oy, queued); return true; } }
	return false; }

template <uint32_t PROGS>
__device__ __forceinline__ unsigned draw_batch_any(TileShared& sh, const TileArgs& A, uint32_t key0, int base, int nb, int ox, int oy, bool queued) {
	unsigned frags = 0;
	draw_batch_if<PROGS, ProgAmy>(frags, sh, A, key0, base, nb, ox, oy, queued) ||
	draw_batch_if<PROGS, ProgOBJ2>(frags, sh, A, key0, base, nb, ox, oy, queued) ||
	draw_batch_if<PROGS, ProgMany>(frags, sh, A, key0, base, nb, ox, oy, queued) ||
	draw_batch_if<PROGS, ProgAlphaTexture>(frags, sh, A, key0, base, nb, ox, oy, queued) ||
	draw_batch_if<PROGS, ProgText>(frags, sh, A, key0, base, nb, ox, oy, queued) ||
	draw_batch_if<PROGS, ProgDepth>(frags, sh, A, key0, base, nb, ox, oy, queued) ||
	draw_batch_if<PROGS, ProgPattern>(frags, sh, A, key0, base, nb, ox, oy, queued) ||
	draw_batch_if<PROGS, ProgOBJ1>(frags, sh, A, key0, base, nb, ox, oy, queued) ||
	draw_batch_if<PROGS, ProgOBJ2S>(frags, sh, A, key0, base, nb, ox, oy, queued) ||
	draw_batch_if<PROGS, ProgEnvmap>(frags, sh, A, key0, base, nb, ox, oy, queued) ||
	draw_batch_if<PROGS, ProgWireframe>(frags, sh, A, key0, base, nb, ox, oy, queued) ||
	draw_batch_if<PROGS, ProgBase>(frags, sh, A, key0, base, nb, ox, oy, queued);
	return frags; }

// ---- IQPostProgram::ShadeCanvas (src/viewer/shaders.hxx:56-66) --------------------------------
// pow(x, y) = exp2f4(log2f4(x) * y)  (rmlv_mvec4.hxx:652-654, 3rdparty/sse-pow/sse_pow.h:19-95):
// degree-3 / degree-5 minimax polynomials, Horner form, separate mul and add

__device__ __forceinline__ float sse_exp2(float x) {
	x = sse_min(x, 129.00000f);
	x = sse_max(x, -126.99999f);
	const int ipart = __float2int_rn(x - 0.5f);                    // _mm_cvtps_epi32: round to nearest even
	const float fpart = x - itof(ipart);
	const float expipart = u2f(static_cast<uint32_t>(ipart + 127) << 23);
	float p = 7.8024521e-2f;
	p = p * fpart + 2.2606716e-1f;
	p = p * fpart + 6.9583356e-1f;
	p = p * fpart + 9.9992520e-1f;
	return expipart * p; }

__device__ __forceinline__ float sse_log2(float x) {
	const uint32_t i = f2u(x);
	const float e = itof(static_cast<int>((i & 0x7f800000u) >> 23) - 127);
	const float m = u2f((i & 0x007fffffu) | 0x3f800000u);
	float p = 0.0596515482674574969533f;
	p = p * m + -0.465725644288844778798f;
	p = p * m + 1.48116647521213171641f;
	p = p * m + -2.52074962577807006663f;
	p = p * m + 2.8882704548164776201f;
	p = p * (m - 1.0f);
	return p + e; }

__device__ __forceinline__ float sse_pow(float x, float y) { return sse_exp2(sse_log2(x) * y); }

// rmlv::sin (rmlv_mvec4.hxx:735-738): scale by 1/pi, wrap1 (:687-691), sin1hp (:714-719)
__device__ __forceinline__ float sse_sin(float x) {
	float M = x * static_cast<float>(1.0 / 3.14159265358979323846);
	const int whole = cvtt(M);
	M = M - itof(whole);
	M = u2f(f2u(M) ^ (static_cast<uint32_t>(whole) << 31));
	const float y = M - (M * fabsf(M));
	return y * (3.1f + 3.6f * fabsf(y)); }

__device__ __forceinline__ void post_iq(float& r, float& g, float& b, float qx, float qy) {
	r = sse_pow(r, 0.45f); g = sse_pow(g, 0.5f); b = sse_pow(b, 0.55f);
	const float vig = 0.2f + 0.8f * sse_pow((((16.0f * qx) * qy) * (1.0f - qx)) * (1.0f - qy), 0.2f);
	r = r * vig; g = g * vig; b = b * vig;
	// (1.0 / 255.0) * hash3(q.x + 13.0*q.y), hash3(n) = fract(sin({n, n+1, n+2}) * 43758.5453123F)
	const float n = qx + 13.0f * qy;
	const float k = static_cast<float>(1.0 / 255.0);
	r = r + k * fract_sse(sse_sin(n) * 43758.5453123f);
	g = g + k * fract_sse(sse_sin(n + 1.0f) * 43758.5453123f);
	b = b + k * fract_sse(sse_sin(n + 2.0f) * 43758.5453123f); }

// sRGB::to_tc / LinearColor::to_tc (rglr_canvas_util.hxx:15-62, ryg-srgb.h:183-223)
__device__ __forceinline__ uint32_t srgb8(float f, const uint32_t* __restrict__ srgbTab) {
	const float clampMin = u2f((127u - 13u) << 23);
	const float almostOne = u2f(0x3f7fffffu);
	float c = sse_max(f, clampMin);
	c = sse_min(c, almostOne);
	const uint32_t bits = f2u(c);
	const uint32_t tab = srgbTab[(bits >> 20) - (127u - 13u) * 8u];
	const uint32_t tmul = (bits >> 12) & 0xffu;
	// _mm_madd_epi16(tab, tmul | 0x02000000): lo16*lo16 + hi16*hi16
	const uint32_t prod = (tab & 0xffffu) * tmul + (tab >> 16) * 0x200u;
	return prod >> 16; }

__device__ __forceinline__ uint32_t linear8(float f) {
	float r = sse_min(f, 1.0f);
	r = sse_max(r, 0.0f);
	return static_cast<uint32_t>(cvtt(r * 255.0f)); }

// Brings the next entries of the tile's list into sh.sorted, in submission (order-key) order, and
// returns how many.  The list is a sequence of cells (kernels.cuh); cells follow each other in
// submission order, entries inside a cell arrive in any order -- almost: every warp of K5 appends
// the entries it has for a cell as ONE ascending run, and runs of different warps cover disjoint
// key ranges (clip fans aside).  Three ways to order a cell, cheapest first:
//   A  cells that fit one raster batch are taken whole, several at a time, and rank-sorted inside
//      each cell (keys are unique, so the rank is the position);
//   B  a longer cell is a few hundred runs: find them (descents of the key sequence), rank the runs
//      by first key (K5 flags the first entry of every run), check that they do not interleave, and gather entries run by run -- no
//      per-entry sort, any length;
//   C  if the runs do interleave (or there are too many): key ranges [nextKey, hi] are selected
//      with one pass over the cell per range (halved until it fits) and sorted bitonically.
// All threads of the CTA must call it.

struct ListCursor {
	int g;                 // next cell
	uint32_t taken;        // entries already taken from cell g (modes B and C)
	int mode;              // of cell g: 0 = undecided / B, 2 = C
	int runs;              // B: runs of cell g (0 = not analysed yet)
	uint32_t nextKey;      // C: pending entries of cell g have order keys >= nextKey
	uint32_t span;
	uint32_t* runTab; };   // B: this tile's run tables in global memory: [kRunCap] start of each run, [kRunCap] entries before it, in key order

__device__ __forceinline__ void sort_scratch(TileShared& sh, uint2* scr, const int n, const uint32_t begin, const int g0, const int gEnd) {
	const int t = threadIdx.x;
	if (n <= kTileThreads) {
		// rank sort inside each cell (cells are already in order): entry t finds its cell, then counts
		// the smaller keys of that cell
		if (t < n) {
			int lo = g0, hi = gEnd - 1;
			while (lo < hi) {
				const int mid = (lo + hi + 1) >> 1;
				if (sh.cellOff[mid] - begin <= static_cast<uint32_t>(t)) { lo = mid; } else { hi = mid - 1; } }
			const int cb = (gEnd > g0) ? static_cast<int>(sh.cellOff[lo] - begin) : 0;
			const int ce = (gEnd > g0) ? static_cast<int>(sh.cellOff[lo + 1] - begin) : n;
			const uint2 mine = scr[t];
			int rank = cb;
			for (int j = cb; j < ce; ++j) { rank += (scr[j].y < mine.y) ? 1 : 0; }
			sh.sorted[rank] = mine.x; } }
	else {
		int P = 512;
		while (P < n) { P <<= 1; }
		unsigned long long* keys = reinterpret_cast<unsigned long long*>(scr);
		for (int i = n + t; i < P; i += kTileThreads) { keys[i] = ~0ull; }
		__syncthreads();
		for (int k = 2; k <= P; k <<= 1) {
			for (int j = k >> 1; j > 0; j >>= 1) {
				for (int i = t; i < P; i += kTileThreads) {
					const int x = i ^ j;
					if (x > i) {
						const unsigned long long a = keys[i], b = keys[x];
						if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[x] = a; } } }
				__syncthreads(); } }
		for (int i = t; i < n; i += kTileThreads) { sh.sorted[i] = scr[i].x; } }
	__syncthreads(); }

__device__ __forceinline__ int load_chunk(TileShared& sh, const TileArgs& A, const uint2* __restrict__ lists, const int G, const uint32_t totalKeys,
                                          ListCursor& lc, Counters* __restrict__ ctr) {
	const int t = threadIdx.x;
	const int warp = t >> 5, lane = t & 31;
	uint2* scr = reinterpret_cast<uint2*>(&sh.ec[0][0]);   // (code, key): key is the high word of the 64-bit view
	__syncthreads();   // the previous batch is done with the batch records (scratch) and sh.sorted
	while (lc.g < G && sh.cellOff[lc.g + 1] == sh.cellOff[lc.g] && sh.lgPerGroup[lc.g] == 0) { ++lc.g; }
	if (lc.g >= G) { return 0; }
	const uint32_t begin = sh.cellOff[lc.g];
	const uint32_t listSize = sh.cellOff[lc.g + 1] - begin;
	const uint32_t nlarge0 = sh.lgPerGroup[lc.g];
	const uint32_t size0 = listSize + nlarge0;      // entries of cell g: its part of the list + queued large items
	const uint2* list = lists + begin;

	if (size0 <= static_cast<uint32_t>(kBatch)) {
		// ---- A: whole cells ------------------------------------------------------------------
		const int g0 = lc.g;
		int gEnd = g0 + 1;
		uint32_t nl = nlarge0;
		while (gEnd < G && (sh.cellOff[gEnd + 1] - begin) + nl + sh.lgPerGroup[gEnd] <= static_cast<uint32_t>(kBatch)) { nl += sh.lgPerGroup[gEnd]; ++gEnd; }
		const int nlist = static_cast<int>(sh.cellOff[gEnd] - begin);
		const int n = nlist + static_cast<int>(nl);
		if (t < nlist) {
			const uint2 e = __ldg(list + t);
			scr[t] = make_uint2(e.y & ~kRunStartBit, e.x);
			prefetch_entry(A, e.y & ~kRunStartBit); }   // the record's trip to memory overlaps the sort
		if (nl) {
			if (t == 0) { sh.sortCount = nlist; }
			__syncthreads();
			if (t < sh.lgCount && sh.lgGroup[t] >= g0 && sh.lgGroup[t] < gEnd) { scr[atomicAdd(&sh.sortCount, 1)] = make_uint2(sh.lgCode[t], sh.lgKey[t]); } }
		lc.g = gEnd;
		__syncthreads();
		// (with queued items mixed in, cells are no longer contiguous in the scratch: rank over everything)
		sort_scratch(sh, scr, n, begin, nl ? 0 : g0, nl ? 0 : gEnd);
		return n; }

	if (lc.mode != 2 && nlarge0 == 0) {
		// ---- B: runs -------------------------------------------------------------------------
		if (lc.runs == 0) {
			uint32_t* rStart = reinterpret_cast<uint32_t*>(&sh.ec[0][0]);   // 3 x kRunCap words inside the batch records
			uint32_t* rKey = rStart + kRunCap;
			uint32_t* rOrder = rKey + kRunCap;
			static_assert(offsetof(TileShared, vref) - offsetof(TileShared, ec) >= 12 * kRunCap, "run scratch aliases the batch records");
			uint32_t* warpCnt = reinterpret_cast<uint32_t*>(sh.warpScratch);
			if (t == 0) { sh.sortCount = 0; sh.firstBad = 0; }
			__syncthreads();
			for (uint32_t strip = 0; strip < size0; strip += kTileThreads) {
				const uint32_t i = strip + t;
				uint32_t key = 0;
				bool isStart = false;
				if (i < size0) {
					const uint2 e = __ldg(list + i);
					key = e.x;
					isStart = (e.y & kRunStartBit) != 0; }
				const unsigned m = __ballot_sync(0xffffffffu, isStart);
				if (lane == 0) { warpCnt[warp] = __popc(m); }
				__syncthreads();
				uint32_t idx = static_cast<uint32_t>(sh.sortCount) + __popc(m & ((1u << lane) - 1u));
				uint32_t total = 0;
#pragma unroll
				for (int w = 0; w < kTileThreads / 32; ++w) { const uint32_t c = warpCnt[w]; if (w < warp) { idx += c; } total += c; }
				if (isStart && idx < static_cast<uint32_t>(kRunCap)) { rStart[idx] = i; rKey[idx] = key; }
				__syncthreads();
				if (t == 0) { sh.sortCount += static_cast<int>(total); } }
			__syncthreads();
			const int R = sh.sortCount;
			if (R <= kRunCap) {
				// rank the runs by first key; lengths follow from the next run in memory order
				for (int r = t; r < R; r += kTileThreads) {
					const uint32_t k = rKey[r];
					int rank = 0;
					for (int q = 0; q < R; ++q) { rank += (rKey[q] < k) ? 1 : 0; }
					rOrder[rank] = static_cast<uint32_t>(r); }
				__syncthreads();
				// runs must not interleave: the last key of a run lies below the first key of the next one
				for (int p = t; p + 1 < R; p += kTileThreads) {
					const uint32_t r = rOrder[p], nx = rOrder[p + 1];
					const uint32_t end = (r + 1 < static_cast<uint32_t>(R)) ? rStart[r + 1] : size0;
					if (__ldg(list + end - 1).x > rKey[nx]) { sh.firstBad = 1; } }
				// exclusive prefix of the run lengths in key order (serial per thread over R / 256 runs, then a block scan)
				const int per = (R + kTileThreads - 1) / kTileThreads;
				const int p0 = min(t * per, R), p1 = min(p0 + per, R);
				uint32_t sum = 0;
				for (int p = p0; p < p1; ++p) {
					const uint32_t r = rOrder[p];
					sum += ((r + 1 < static_cast<uint32_t>(R)) ? rStart[r + 1] : size0) - rStart[r]; }
				uint32_t incl = sum;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) { incl += v; } }
				__syncthreads();   // (warpCnt reuse; sh.firstBad complete)
				if (lane == 31) { warpCnt[warp] = incl; }
				__syncthreads();
				uint32_t before = incl - sum;
				for (int w = 0; w < warp; ++w) { before += warpCnt[w]; }
				for (int p = p0; p < p1; ++p) {
					const uint32_t r = rOrder[p];
					lc.runTab[p] = rStart[r];
					lc.runTab[kRunCap + p] = before;
					before += ((r + 1 < static_cast<uint32_t>(R)) ? rStart[r + 1] : size0) - rStart[r]; }
				__syncthreads();
				if (sh.firstBad == 0) { lc.runs = R; } }
			if (lc.runs == 0) { lc.mode = 2; } }   // interleaved (clip fans) or too many runs
		if (lc.runs) {
			const int R = lc.runs;
			if (t == 0) { atomicAdd(&ctr->chunksRunMerge, 1u); }
			const int n = static_cast<int>(min(static_cast<uint32_t>(kSortCap), size0 - lc.taken));
			for (int i = t; i < n; i += kTileThreads) {
				const uint32_t pos = lc.taken + static_cast<uint32_t>(i);
				int lo = 0, hi = R - 1;
				while (lo < hi) {
					const int mid = (lo + hi + 1) >> 1;
					if (lc.runTab[kRunCap + mid] <= pos) { lo = mid; } else { hi = mid - 1; } }
				const uint32_t code = __ldg(list + lc.runTab[lo] + (pos - lc.runTab[kRunCap + lo])).y & ~kRunStartBit;
				sh.sorted[i] = code;
				prefetch_entry(A, code); }
			lc.taken += static_cast<uint32_t>(n);
			if (lc.taken >= size0) { ++lc.g; lc.taken = 0; lc.mode = 0; lc.runs = 0; }
			__syncthreads();
			return n; } }

	// ---- C: key ranges -------------------------------------------------------------------------
	if (t == 0) { atomicAdd(&ctr->chunksKeyRange, 1u); }
	if (lc.taken == 0) {
		lc.nextKey = 0;
		const uint32_t rounds = (size0 + (3 * kSortCap / 4) - 1) / (3 * kSortCap / 4);
		lc.span = max(totalKeys / (rounds * static_cast<uint32_t>(G)), 1u); }
	int n;
	while (true) {
		if (t == 0) { sh.sortCount = 0; }
		__syncthreads();
		const uint32_t room = 0xffffffffu - lc.nextKey;
		const uint32_t hi = lc.nextKey + min(lc.span - 1u, room);
		for (uint32_t i = t; i < listSize; i += kTileThreads) {
			const uint2 e = __ldg(list + i);
			if (e.x >= lc.nextKey && e.x <= hi) {
				const int slot = atomicAdd(&sh.sortCount, 1);
				if (slot < kSortCap) { scr[slot] = make_uint2(e.y & ~kRunStartBit, e.x); } } }
		if (nlarge0 && t < sh.lgCount && sh.lgGroup[t] == lc.g && sh.lgKey[t] >= lc.nextKey && sh.lgKey[t] <= hi) {
			const int slot = atomicAdd(&sh.sortCount, 1);
			if (slot < kSortCap) { scr[slot] = make_uint2(sh.lgCode[t], sh.lgKey[t]); } }
		__syncthreads();
		n = sh.sortCount;
		__syncthreads();
		if (n > kSortCap) { lc.span = max(lc.span >> 1, 1u); continue; }
		if (n == 0 && hi != 0xffffffffu) { lc.nextKey = hi + 1u; lc.span = (lc.span < 0x80000000u) ? lc.span * 2u : 0xffffffffu; continue; }
		if (n < kSortCap / 4) { lc.span = (lc.span < 0x80000000u) ? lc.span * 2u : 0xffffffffu; }
		lc.nextKey = hi + 1u;
		break; }
	lc.taken += static_cast<uint32_t>(n);
	if (lc.taken >= size0 || n == 0) { ++lc.g; lc.taken = 0; lc.mode = 0; lc.runs = 0; }
	sort_scratch(sh, scr, n, begin, 0, 0);
	return n; }

// Runs the clear / store commands that precede draw `di` for this thread's own 2x2 quad (tile colour and depth of a
// quad are only ever touched by its thread, so no CTA-wide synchronisation is involved).
__device__ __forceinline__ void exec_cmds(TileShared& sh, const TileArgs& A, int& ci, const int di, const int t, const int px, const int py, const bool onScreen) {
	// non-draw commands that precede that draw
	while (ci < A.fp.ncmds && cmd_before_draw(A, ci) <= di) {
		FrameCmd cmd;
		CmdState s;
		if (ci < kInlineCmds) { cmd = A.icmd[ci]; s = A.icmdState[ci]; }
		else {
			cmd = A.cmds[ci];
			const DevState& gs = A.states[cmd.state];
			s.clearColor[0] = gs.clearColor[0]; s.clearColor[1] = gs.clearColor[1]; s.clearColor[2] = gs.clearColor[2]; s.clearColor[3] = gs.clearColor[3];
			s.clearDepth = gs.clearDepth; s.programId = gs.programId; s.uniform0 = gs.uniforms[0]; s.color0Type = gs.color0Type; }
		++ci;
		switch (cmd.type) {
		case kCmdClear: {
			// GPU::DrawImpl CMD_CLEAR (rglv_gpu.cxx:311-344)
			const bool clearColor = (cmd.arg & 1) != 0, clearDepth = (cmd.arg & 2) != 0;
#pragma unroll
			for (int l = 0; l < 4; ++l) {
				if (clearColor) { sh.chan[0][l][t] = s.clearColor[0]; sh.chan[1][l][t] = s.clearColor[1]; sh.chan[2][l][t] = s.clearColor[2]; }
				if (clearDepth) { sh.chan[3][l][t] = s.clearDepth; } } }
			break;
		case kCmdStoreTC: {
			// GPUBltImpl::StoreTrueColor -> FilterTile<SHADER, sRGB|LinearColor>
			// (every lane converts its quad, on screen or not: the lanes of a pair exchange halves below)
			uint32_t out[4];
#pragma unroll
			for (int l = 0; l < 4; ++l) {
				float r = sh.chan[0][l][t], g = sh.chan[1][l][t], b = sh.chan[2][l][t];
				if (s.programId == 2) {   // ExposurePostProgram (shaders.hxx:40-53)
					const float ex = s.uniform0;
					r = r * ex; g = g * ex; b = b * ex; }
				else if (s.programId == 3) {
					// FilterTile's running fragment coordinate (rglr_algorithm.hxx:75-96): starts at the
					// reference tile's left/top edge and is advanced by repeated float adds per quad
					const int ptx = (px / A.fp.postTileW) * A.fp.postTileW, pty = (py / A.fp.postTileH) * A.fp.postTileH;
					const float iw = 1.0f / itof(A.fp.width), ih = 1.0f / itof(A.fp.height);
					float fcx = (itof(ptx) + 0.5f) / itof(A.fp.width) + ((l & 1) ? iw : 0.0f);
					float fcy = ((itof(A.fp.height - pty)) - 0.5f) / itof(A.fp.height) - ((l & 2) ? ih : 0.0f);
					const float fcdx = iw * 2.0f, fcdy = -ih * 2.0f;
					for (int kx = ptx; kx < px; kx += 2) { fcx += fcdx; }
					for (int ky = pty; ky < py; ky += 2) { fcy += fcdy; }
					post_iq(r, g, b, fcx, fcy); }
				out[l] = (cmd.arg & 1) ? ((srgb8(r, sh.srgbTab) << 16) | (srgb8(g, sh.srgbTab) << 8) | srgb8(b, sh.srgbTab))
				                       : ((linear8(r) << 16) | (linear8(g) << 8) | linear8(b)); }
			// 128-bit stores: lanes 2k and 2k+1 hold horizontally adjacent quads (4 x 2 pixels); they swap halves so that
			// the even lane writes the four pixels of the upper row and the odd lane those of the lower row -- a tile row is
			// eight 16-byte stores = one full 128-byte line.  Needs 16-byte aligned rows (destination, stride and width
			// multiples of 4 pixels: then both lanes of a pair are on screen together); otherwise 64-bit stores.
			uint32_t* dst = static_cast<uint32_t*>(cmd.dst);
			const bool wide = ((reinterpret_cast<uintptr_t>(dst) | (static_cast<uintptr_t>(cmd.dstStride) * 4u)) & 15u) == 0 && (A.fp.width & 3) == 0;
			if (wide) {
				const bool odd = (threadIdx.x & 1) != 0;
				const uint32_t give0 = odd ? out[0] : out[2], give1 = odd ? out[1] : out[3];
				const uint32_t got0 = __shfl_xor_sync(0xffffffffu, give0, 1), got1 = __shfl_xor_sync(0xffffffffu, give1, 1);
				if (onScreen) {
					if (!odd) { *reinterpret_cast<uint4*>(dst + static_cast<size_t>(py) * cmd.dstStride + px) = make_uint4(out[0], out[1], got0, got1); }
					else { *reinterpret_cast<uint4*>(dst + static_cast<size_t>(py + 1) * cmd.dstStride + (px - 2)) = make_uint4(got0, got1, out[2], out[3]); } } }
			else if (onScreen) {
				*reinterpret_cast<uint2*>(dst + static_cast<size_t>(py) * cmd.dstStride + px) = make_uint2(out[0], out[1]);
				*reinterpret_cast<uint2*>(dst + static_cast<size_t>(py + 1) * cmd.dstStride + px) = make_uint2(out[2], out[3]); } }
			break;
		case kCmdStoreFP: {
			// Copy(QFloat4Canvas -> FloatingPointCanvas) (rglr_algorithm.cxx:247-279): alpha = 0
			if (onScreen) {
				float4* dst = static_cast<float4*>(cmd.dst);
#pragma unroll
				for (int l = 0; l < 4; ++l) {
					dst[static_cast<size_t>(py + (l >> 1)) * cmd.dstStride + px + (l & 1)] =
						make_float4(sh.chan[0][l][t], sh.chan[1][l][t], sh.chan[2][l][t], 0.0f); } } }
			break;
		case kCmdStoreHalfFP: {
			// Downsample(QFloat4Canvas -> FloatingPointCanvas) (rglr_algorithm.cxx:118-141): one output
			// pixel per quad, ((p0 + p1) + p2) + p3 then * 0.25, alpha = 0
			if (onScreen) {
				float4* dst = static_cast<float4*>(cmd.dst);
				float avg[3];
#pragma unroll
				for (int ch = 0; ch < 3; ++ch) {
					avg[ch] = (((sh.chan[ch][0][t] + sh.chan[ch][1][t]) + sh.chan[ch][2][t]) + sh.chan[ch][3][t]) * 0.25f; }
				dst[static_cast<size_t>(py >> 1) * cmd.dstStride + (px >> 1)] = make_float4(avg[0], avg[1], avg[2], 0.0f); } }
			break;
		case kCmdStoreQuadsFP: {
			// Copy(QFloat4Canvas | QFloat3Canvas -> QFloat4Canvas) (rglr_algorithm.cxx:320-368): the quad-
			// swizzled layout itself, 64 bytes per 2x2 quad {r[4], g[4], b[4], a[4]}.  The fourth plane is
			// what the reference's tile holds there: depth (RB_COLOR_DEPTH), 1.0 (RB_RGBF32), or the
			// clear colour's alpha (RB_RGBAF32: no program writes alpha)
			if (onScreen) {
				float4* dst = static_cast<float4*>(cmd.dst) + (static_cast<size_t>(py >> 1) * cmd.dstStride + (px >> 1)) * 4;
#pragma unroll
				for (int ch = 0; ch < 3; ++ch) {
					dst[ch] = make_float4(sh.chan[ch][0][t], sh.chan[ch][1][t], sh.chan[ch][2][t], sh.chan[ch][3][t]); }
				if (s.color0Type == 0) { dst[3] = make_float4(sh.chan[3][0][t], sh.chan[3][1][t], sh.chan[3][2][t], sh.chan[3][3][t]); }
				else {
					const float a = (s.color0Type == 1) ? 1.0f : s.clearColor[3];
					dst[3] = make_float4(a, a, a, a); } } }
			break;
		case kCmdStoreDepth: {
			if (onScreen) {
				float* dst = static_cast<float*>(cmd.dst);
				*reinterpret_cast<float2*>(dst + static_cast<size_t>(py) * cmd.dstStride + px) = make_float2(sh.chan[3][0][t], sh.chan[3][1][t]);
				*reinterpret_cast<float2*>(dst + static_cast<size_t>(py + 1) * cmd.dstStride + px) = make_float2(sh.chan[3][2][t], sh.chan[3][3][t]); } }
			break;
		default: break; } } }

// RSR_WARP_WALK=1 (experiment, off): warp-autonomous list walk, see the comment in tile_kernel and profiles/README.md
#ifndef RSR_WARP_WALK
#define RSR_WARP_WALK 0
#endif
#ifndef RSR_WARP_QUEUE_MIN
#define RSR_WARP_QUEUE_MIN 12
#endif

// one warp rasterises the nb (<= 32) triangles it has set up in its own record slots [base, base + nb)
template <uint32_t PROGS>
__device__ __forceinline__ unsigned raster_slots(TileShared& sh, const TileArgs& A, uint32_t key0, int base, int nb, int ox, int oy) {
	__syncwarp();
	const int lane = threadIdx.x & 31;
	bool tiny = false;
	if (lane < nb) {
		const uint32_t bb = sh.bbox[base + lane];
		tiny = bb != 0 && (((bb >> 12) & 63) - (bb & 63)) <= 6 && (((bb >> 18) & 63) - ((bb >> 6) & 63)) <= 6; }
	const int ntiny = __popc(__ballot_sync(0xffffffffu, tiny));
	const bool queued = nb >= RSR_WARP_QUEUE_MIN && ntiny * 2 > nb;
	const unsigned frags = draw_batch_any<PROGS>(sh, A, key0, base, nb, ox, oy, queued);
	__syncwarp();
	return frags; }

#ifndef RSR_TILE_CTAS
#define RSR_TILE_CTAS 3
#endif
template <uint32_t PROGS, int MIN_CTAS>
__global__ void __launch_bounds__(kTileThreads, MIN_CTAS)
tile_kernel(const __grid_constant__ TileArgs A) {
	extern __shared__ __align__(16) unsigned char tileSmem[];
	TileShared& sh = *reinterpret_cast<TileShared*>(tileSmem);
	const int t = threadIdx.x;
	const int warp = t >> 5, lane = t & 31;
	const int lx = (warp & 1) * 16 + (lane & 7) * 2, ly = (warp >> 1) * 8 + (lane >> 3) * 2;
	PHASE_BEGIN();

	// Everything up to the grid dependency wait touches only data that no kernel of the frame writes (the
	// approximation tables, constants): with programmatic dependent launch this prologue runs while the
	// list fill kernel is still draining.
	reinterpret_cast<uint4*>(sh.rcpLut)[t] = __ldg(reinterpret_cast<const uint4*>(A.luts->rcp16) + t);   // 2048 x 16 bit = 256 x 16 bytes
	if (t < 104) { sh.srgbTab[t] = kSrgbTab4[t]; }
	if (t < kMaxGroups) { sh.lgPerGroup[t] = 0; }
	if (t == 0) { sh.lgCount = 0; }
#pragma unroll
	for (int c = 0; c < 4; ++c) {
#pragma unroll
		for (int l = 0; l < 4; ++l) { sh.chan[c][l][t] = 0.0f; } }
	pdl_launch_dependents();
	pdl_wait();
	PHASE(1);

	const int tile = A.tileOrder ? static_cast<int>(__ldg(A.tileOrder + blockIdx.x)) : static_cast<int>(blockIdx.x);   // longest lists first
	const int tileX = tile % A.fp.tilesX, tileY = tile / A.fp.tilesX;
	const int ox = tileX * kTile, oy = tileY * kTile;
	const int px = ox + lx, py = oy + ly;
	const bool onScreen = (px < A.fp.width) && (py < A.fp.height);

	// the reference tile this device tile lies in: only its corner matters (edge start point)
	const int rl = (ox / A.fp.refTileW) * A.fp.refTileW, rt = (oy / A.fp.refTileH) * A.fp.refTileH;
	const int rr = min(rl + A.fp.refTileW, A.fp.width), rb = min(rt + A.fp.refTileH, A.fp.height);

	// this tile's cells; a frame whose lists overflowed is rendered again by the host: stay inside the buffer meanwhile
	const int G = A.fp.groups;
	for (int g = t; g <= G; g += kTileThreads) {
		uint32_t off = __ldg(A.tileBase + tile + (g == G ? 1 : 0));
		if (g > 0 && g < G) { off += __ldg(A.cellRel + static_cast<size_t>(tile) * G + g); }
		sh.cellOff[g] = min(off, A.fp.listCapacity); }
	const unsigned nLarge = min(__ldcg(&A.ctr->nLarge), A.fp.largeCapacity);   // (issued with the offset loads: one trip for both)
	__syncthreads();
	// queued large items (kernels.cuh) that cover this tile: a short scan instead of one list entry per tile
	for (unsigned i = t; i < nLarge; i += kTileThreads) {
		const LargeItem it = A.large[i];
		const int tx0 = it.packed & 63, ty0 = (it.packed >> 6) & 63, tx1 = (it.packed >> 12) & 63, ty1 = (it.packed >> 18) & 63;
		if (tileX >= tx0 && tileX <= tx1 && tileY >= ty0 && tileY <= ty1) {
			const int slot = atomicAdd(&sh.lgCount, 1);
			if (slot < kTileLargeCap) {
				sh.lgKey[slot] = it.okey; sh.lgCode[slot] = it.code; sh.lgGroup[slot] = static_cast<uint8_t>(it.group);
				atomicAdd(reinterpret_cast<unsigned int*>(sh.lgPerGroup) + (it.group >> 1), (it.group & 1u) ? 0x10000u : 1u); } } }
	__syncthreads();
	if (sh.lgCount > kTileLargeCap) {
		__syncthreads();
		if (t == 0) { atomicOr(&A.ctr->overflow, 8u); sh.lgCount = kTileLargeCap; }
		__syncthreads(); }
	ListCursor lc{0, 0u, 0, 0, 0u, 1u, A.runScratch + static_cast<size_t>(tile) * (2 * kRunCap)};
	int chunkN = 0, chunkPos = 0;
	unsigned frags = 0;

	// Frame walk.  A.cmds holds the non-draw commands (clear / stores) in submission order, each
	// tagged with the number of draws recorded before it; draws are discovered from the tile's own
	// id-sorted list, so a tile only pays for the draws that touch it.  Up to 256 consecutive
	// list entries are rasterised as one batch as long as they share a program + pipeline flags
	// (DevDraw::batchKey) and no clear/store command lies between them -- a scene made of hundreds
	// of tiny draws (one per textured quad) still fills whole batches.
	int ci = 0;
#if RSR_WARP_WALK
	// Warp-autonomous walk.  Only fetching + sorting a chunk of the list is CTA-wide; inside a chunk every warp
	// walks the sorted entries on its own: 32 at a time it reads their records, finds where the current run (same
	// program + pipeline flags, no clear / store in between) ends, keeps the triangles whose bounding box touches
	// its 16x8-pixel region, sets those up in its own 32 record slots and rasterises them when the slots are full
	// or the run ends.  No warp ever waits for another one inside a chunk: a batch whose triangles cluster in one
	// region no longer stalls the other seven warps at a barrier.  All warps take the same decisions (they read
	// the same entries), so run boundaries and the commands between runs need no agreement protocol.
	const unsigned ltMask = (1u << lane) - 1u;
	const int wbase = warp * 32;
	const int rx = (warp & 1) * 16, ry = (warp >> 1) * 8;
	while (true) {
		if (chunkPos == chunkN && lc.g < G) {
			PHASE(2);
			chunkN = load_chunk(sh, A, A.lists, G, A.fp.totalKeys, lc, A.ctr);
			chunkPos = 0; }
		PHASE(3);
		int di = A.fp.ndraws;   // draw owning the next list entry (ndraws = none left)
		uint32_t key0 = 0;
		if (chunkPos < chunkN) {
			const EntryRec head = load_entry(A, sh.sorted[chunkPos]);
			di = static_cast<int>(head.q3.w);
			key0 = head.q4.x; }
		PHASE(5);
		exec_cmds(sh, A, ci, di, t, px, py, onScreen);
		if (di >= A.fp.ndraws) { break; }
		PHASE(6);
		const int bound = (ci < A.fp.ncmds) ? cmd_before_draw(A, ci) : 0x7fffffff;   // draws >= bound come after cmds[ci]
		int pos = chunkPos, nrec = 0;
		bool runOver = false;
		while (!runOver) {
			const int e = pos + lane;
			const bool valid = e < chunkN;
			uint32_t myId = 0;
			EntryRec myRec;
			myRec.q0 = myRec.q1 = myRec.q2 = myRec.q3 = myRec.q4 = make_uint4(0u, 0u, 0u, 0u);
			if (valid) { myId = sh.sorted[e]; myRec = load_entry(A, myId); }
			const bool bad = valid && (static_cast<int>(myRec.q3.w) >= bound || myRec.q4.x != key0);
			const unsigned badMask = __ballot_sync(0xffffffffu, bad);
			const int nvalid = min(32, chunkN - pos);
			const int take = badMask ? (__ffs(badMask) - 1) : nvalid;
			runOver = (badMask != 0u) || (pos + take >= chunkN);
			bool mine = static_cast<int>(lane) < take;
			if (mine && !(myId & kFanIdBit)) {
				// conservative pixel bounding box (setup_edges clamps it further) against this warp's region
				const int X0 = static_cast<int>(myRec.q0.x), X1 = static_cast<int>(myRec.q0.y), X2 = static_cast<int>(myRec.q0.z);
				const int Y0 = static_cast<int>(myRec.q0.w), Y1 = static_cast<int>(myRec.q1.x), Y2 = static_cast<int>(myRec.q1.y);
				const int minx = ((min(min(X0, X1), X2) >> 4) & ~1) - ox, maxx = ((max(max(X0, X1), X2) + 15) >> 4) - ox;
				const int miny = ((min(min(Y0, Y1), Y2) >> 4) & ~1) - oy, maxy = ((max(max(Y0, Y1), Y2) + 15) >> 4) - oy;
				mine = (minx < rx + 16) && (maxx > rx) && (miny < ry + 8) && (maxy > ry); }
			const unsigned hitMask = __ballot_sync(0xffffffffu, mine);
			const int nhit = __popc(hitMask);
			if (nrec + nhit > 32) {
				frags += raster_slots<PROGS>(sh, A, key0, wbase, nrec, ox, oy);
				nrec = 0; }
			if (mine) {
				const int slot = wbase + nrec + __popc(hitMask & ltMask);
				sh.state[slot] = static_cast<uint16_t>(myRec.q4.y);
				if constexpr ((PROGS & kSamplingProgs) != 0u) { setup_texture(sh, slot, A, myRec.q4.y & 0xffffu); }
				setup_triangle(sh, slot, myId, myRec, A, ox, oy, rl, rt, rr, rb); }
			nrec += nhit;
			pos += take; }
		if (nrec) { frags += raster_slots<PROGS>(sh, A, key0, wbase, nrec, ox, oy); }
		PHASE(8);
		chunkPos = pos; }
	PHASE(9);
#else
	while (true) {
		if (chunkPos == chunkN && lc.g < G) {
			PHASE(2);
			chunkN = load_chunk(sh, A, A.lists, G, A.fp.totalKeys, lc, A.ctr);
			chunkPos = 0; }
		PHASE(3);
		// every thread fetches the record of "its" entry of the next (up to) 256: one trip to memory serves
		// both the question "which draw / program does the batch start with" and the triangle setup
		const int avail = min(kBatch, chunkN - chunkPos);
		uint32_t myId = 0;
		EntryRec myRec;
		myRec.q3 = make_uint4(0u, 0u, 0u, 0u); myRec.q4 = make_uint4(0u, 0u, 0u, 0u);
		if (t < avail) {
			myId = sh.sorted[chunkPos + t];
			myRec = load_entry(A, myId); }
		__syncthreads();   // previous batch fully rasterised: its records, sh.headDraw / headKey / firstBad may be overwritten
		PHASE(4);
		if (t == 0) { sh.headDraw = avail ? static_cast<int>(myRec.q3.w) : A.fp.ndraws; sh.firstBad = avail; }
		__syncthreads();
		const int di = sh.headDraw;   // draw owning the next list entry (ndraws = none left)
		PHASE(5);
		exec_cmds(sh, A, ci, di, t, px, py, onScreen);
		if (di >= A.fp.ndraws) { break; }
		PHASE(6);

		// The batch ends at the next clear / store command (or after 256 entries).  Inside it entries of different programs /
		// pipeline states form RUNS (DevDraw::batchKey); all of the batch's triangles are set up in one go -- setup does not
		// depend on the program -- and every warp then walks the runs on its own: no CTA barrier between the runs of a batch.
		const int bound = (ci < A.fp.ncmds) ? cmd_before_draw(A, ci) : 0x7fffffff;   // draws >= bound come after cmds[ci]
		if (t < avail && static_cast<int>(myRec.q3.w) >= bound) { atomicMin(&sh.firstBad, t); }
		if (t < avail) { sh.key[t] = myRec.q4.x; prefetch_varyings<PROGS>(A, myId, myRec); }
		__syncthreads();
		const int nb = sh.firstBad;   // >= 1: entry 0 belongs to draw di
		// start the trip for the records of the batch after this one while this one is rasterised
		if (chunkPos + nb + t < chunkN) { prefetch_entry(A, sh.sorted[chunkPos + nb + t]); }
		bool tiny = false, runStart = false;
		if (t < nb) {
			sh.state[t] = static_cast<uint16_t>(myRec.q4.y);
			if constexpr ((PROGS & kSamplingProgs) != 0u) { setup_texture(sh, t, A, myRec.q4.y & 0xffffu); }
			setup_triangle(sh, t, myId, myRec, A, ox, oy, rl, rt, rr, rb);
			const uint32_t bb = sh.bbox[t];
			tiny = bb != 0 && (((bb >> 12) & 63) - (bb & 63)) <= 6 && (((bb >> 18) & 63) - ((bb >> 6) & 63)) <= 6;
			runStart = (t == 0) || (sh.key[t - 1] != myRec.q4.x); }
		const unsigned startBits = __ballot_sync(0xffffffffu, runStart), tinyBits = __ballot_sync(0xffffffffu, tiny);
		if (lane == 0) { sh.runStartBits[warp] = startBits; sh.tinyBits[warp] = tinyBits; }
		__syncthreads();   // (publishes the setup records and the run masks)
		PHASE(7);
#ifndef RSR_QUEUE_MIN_NB
#define RSR_QUEUE_MIN_NB 96
#endif
		for (int r0 = 0; r0 < nb; ) {
			// the run [r0, r1): r1 = next set bit of the start mask after r0
			int r1 = nb;
			if (r0 + 1 < nb) {
				const int wEnd = (nb + 31) >> 5;   // (bits at or beyond nb are never set)
				int w = (r0 + 1) >> 5;
				unsigned m = sh.runStartBits[w] & (0xffffffffu << ((r0 + 1) & 31));
				while (m == 0u && ++w < wEnd) { m = sh.runStartBits[w]; }
				if (m != 0u) { r1 = w * 32 + __ffs(m) - 1; } }
			const int len = r1 - r0;
			bool queued = false;
			if (len >= RSR_QUEUE_MIN_NB) {
				int ntiny = 0;
				for (int w = r0 >> 5; w <= (r1 - 1) >> 5; ++w) {
					unsigned m = sh.tinyBits[w];
					if (w == (r0 >> 5)) { m &= 0xffffffffu << (r0 & 31); }
					if (w == ((r1 - 1) >> 5) && ((r1 & 31) != 0)) { m &= (1u << (r1 & 31)) - 1u; }
					ntiny += __popc(m); }
				queued = ntiny * 2 > len; }
#ifdef RSR_FORCE_QUEUED   // (experiment: the queued rasteriser for every run of at least this many entries; c2 120 -> 153-188 us)
			queued = len >= RSR_FORCE_QUEUED;
#endif
			frags += draw_batch_any<PROGS>(sh, A, sh.key[r0], r0, len, ox, oy, queued);
			r0 = r1; }
		PHASE(8);
		chunkPos += max(nb, 1); }
	PHASE(9);

#endif
	// fragment statistics: one atomic per CTA
	for (int o = 16; o > 0; o >>= 1) { frags += __shfl_down_sync(0xffffffffu, frags, o); }
	__syncthreads();
	if (t == 0) { sh.sortCount = 0; }
	__syncthreads();
	if (lane == 0 && frags) { atomicAdd(&sh.sortCount, static_cast<int>(frags)); }
	__syncthreads();
	if (t == 0 && sh.sortCount) { atomicAdd(&A.ctr->fragments, static_cast<unsigned long long>(sh.sortCount)); }
	PHASE(10); }

}  // namespace rsr
