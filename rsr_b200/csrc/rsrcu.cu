// rsrcu.cu -- the C ABI of include/rsrcu.h: frame recording on the host, kernel orchestration on
// one CUDA stream per context.  Replaces rglv::GPU::RunImpl / BinImpl / DrawImpl
// (src/rgl/rglv/rglv_gpu.cxx:90-432); the CPU job system (src/rcl/rclmt) has no counterpart here:
// tiles are CTAs of one grid launch, frames are ordered by the stream.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <type_traits>
#include <vector>

#include <cuda_runtime.h>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include "../../include/rsrcu.h"
#include "tile_kernel.cuh"
#include "post_kernels.cuh"
#include "march_kernels.cuh"

namespace rsr {
void harvest_luts(uint32_t* rcp2048, uint32_t* rsqrt2048);
uint64_t verify_luts(const uint32_t* rcp2048, const uint32_t* rsqrt2048);
}

namespace {

using namespace rsr;

thread_local std::string g_lastError;

int fail(int code, const char* fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_lastError = buf;
	return code; }

#define CU(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
	return fail(RSRCU_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)

// growable device allocation
struct DevBuf {
	void* ptr{nullptr};
	size_t cap{0};
	cudaError_t reserve(size_t bytes) {
		if (bytes <= cap) { return cudaSuccess; }
		if (ptr) { cudaFree(ptr); ptr = nullptr; cap = 0; }
		size_t want = std::max(bytes + bytes / 4, static_cast<size_t>(1) << 16);
		cudaError_t e = cudaMalloc(&ptr, want);
		if (e == cudaSuccess) { cap = want; }
		return e; }
	void release() { if (ptr) { cudaFree(ptr); ptr = nullptr; cap = 0; } } };

// pinned host staging mirrored by a device arena: all per-frame ("upload always") data and the
// small frame tables go through it with ONE host->device copy per frame
struct UploadArena {
	uint8_t* host{nullptr};
	size_t hostCap{0};
	DevBuf dev;
	size_t used{0};
	cudaError_t reserveHost(size_t bytes) {
		if (bytes <= hostCap) { return cudaSuccess; }
		size_t want = std::max(bytes + bytes / 2, static_cast<size_t>(1) << 20);
		uint8_t* nh = nullptr;
		cudaError_t e = cudaMallocHost(&nh, want);
		if (e != cudaSuccess) { return e; }
		if (host) { std::memcpy(nh, host, used); cudaFreeHost(host); }
		host = nh; hostCap = want;
		return cudaSuccess; }
	// returns offset
	cudaError_t push(const void* src, size_t bytes, size_t& offset) {
		const size_t at = (used + 255) & ~static_cast<size_t>(255);
		cudaError_t e = reserveHost(at + bytes);
		if (e != cudaSuccess) { return e; }
		if (src) { std::memcpy(host + at, src, bytes); }
		used = at + bytes;
		offset = at;
		return cudaSuccess; }
	void release() { if (host) { cudaFreeHost(host); host = nullptr; hostCap = 0; } dev.release(); } };

struct StaticAlloc { void* dev; size_t bytes; float lo, hi; bool ranged; };   // (lo, hi: value range of a float array, when asked for)

// direct-mapped front of the static cache: a frame binds the same few hundred pointers thousands of times
struct PtrCacheEntry { const void* host; size_t bytes; void* dev; float lo, hi; bool ranged; };
struct FrameCacheEntry { const void* host{nullptr}; size_t bytes{0}; size_t off{0}; uint64_t stamp{0}; };   // RSRCU_UPLOAD_FRAME: staged once per frame
constexpr size_t kPtrCacheSize = 2048;
inline size_t ptrCacheIndex(const void* p) { return static_cast<size_t>((reinterpret_cast<uintptr_t>(p) >> 4) * 0x9E3779B97F4A7C15ull >> 53); }

struct PendingCopy { void* hostDst; const void* devSrc; size_t rowBytes; size_t rows; size_t hostPitch; size_t devPitch; };

// a device pointer that is either absolute (static cache) or an offset into the frame arena
struct DevRef { bool arena{false}; size_t off{0}; const void* abs{nullptr}; bool null{true}; float lo{0.0f}, hi{0.0f}; bool ranged{false}; };

struct HostTex { DevRef ref; uint32_t texelCount{0}; int width{8}, height{8}, stride{8}, filter{0}; };

// one state snapshot: the device record, built once when the snapshot is taken (pointer fields hold
// encoded references until rsrcu_end_frame knows the arena's device address), plus what draw-time
// validation needs
struct HostState {
	DevState ds;
	size_t bufferFloats[16];
	float posLo[3], posHi[3];     // value ranges of the position arrays (static buffers only): the draw's object-space bounding box
	bool posRanged;
	int programId, key;
	bool posNull, tex0Null, hasArenaRef;
	int tex0Width, tex0Height, tex0Stride; };

constexpr uint64_t kArenaRefBit = 1ull << 63;   // encoded reference: offset into the frame arena instead of a device address

struct HostDraw {
	DevDraw d;
	DevRef indices; };

}  // namespace

constexpr int kSlots = 3;

// RSRCU_UPLOAD_FRAME, large arrays: the host memory is page-locked in place (cudaHostRegister) and copied by the copy
// engine straight from where it lies -- no staging memcpy on the submitting thread.  One device buffer per slot of the
// frame ring (a frame's copy may not overwrite what an earlier frame's kernels still read).
struct PinnedInPlace {
	size_t bytes{0};
	uintptr_t regBase{0}; size_t regBytes{0};   // the page-aligned range that was registered
	bool ok{false};                              // false: registration failed (shared pages, not host memory...): staged like small arrays
	bool owned{false};                           // this context registered the range (another context may have done it first)
	DevBuf dev[kSlots];
	uint64_t lastFrame{0}; };
constexpr size_t kPinInPlaceMin = static_cast<size_t>(1) << 20;
constexpr size_t kPinnedMark = ~static_cast<size_t>(0);   // FrameCacheEntry::off of an array that went the in-place way


// everything rsrcu_end_frame derives from a recorded frame before it can launch: kept per context for the last
// frame, and copied into an rsrcu_frame by rsrcu_retain_frame
struct FramePlan {
	bool valid{false};
	FrameParams fp{};
	size_t offStates{0}, offDraws{0}, offCmds{0}, offVBlocks{0}, offPBlocks{0}, arenaBytes{0};
	uint64_t ptvbF4{0}, nvertsTotal{0}, pjobs{0};
	int ncmdInline{0};
	FrameCmd icmd[kInlineCmds]{};
	CmdState icmdState[kInlineCmds]{};
	std::vector<PendingCopy> copies;
	void* tcDev{nullptr}; int tcStride{0};   // the frame's last true-colour store (rsrcu_device_truecolor)
	int storesUsed{0};                       // store targets of the context's pool the frame's commands point into
	uint64_t trianglesSubmitted{0};
	uint64_t inputBytes{0};
	uint32_t progMask{0}; };                 // programs the frame's draws use (prog_bit): selects the tile kernel instantiation

struct rsrcu_frame {
	FramePlan plan;
	void* devArena{nullptr};
	std::vector<DevBuf> stores;              // a retained frame owns its store targets (taken from the context's pool)
	int device{0}; };

// what was launched into a slot of the ring (counters, store targets, arena): enough to launch it again when its
// read-back shows that a device-side buffer overflowed
struct Launched {
	const FramePlan* plan{nullptr};
	const uint8_t* arenaDev{nullptr};
	bool checked{true}; };

struct rsrcu_ctx {
	int device{0};
	cudaStream_t stream{nullptr};
	cudaStream_t copyStream{nullptr};   // device->host copies of finished frames overlap the next frame's kernels
	cudaEvent_t evRendered[kSlots]{};
	// read-back of the latest frame: enqueued behind the NEXT frame's upload (or at a sync) -- both go through the
	// copy engines in submission order, and a device->host copy that is still waiting for its kernels must not
	// hold up the upload the following frame's kernels wait for
	std::vector<PendingCopy> deferredCopies;
	// RSRCU_TRACE=1: timestamps of the first 64 frames (kernels start/end on the render stream, read-back start/end on
	// the copy stream), printed by rsrcu_destroy -- a poor man's timeline for pipelining problems
	bool trace{false};
	bool traceStages{false};   // RSRCU_TRACE=2: events 2 and 3 of a frame are front end done / tile kernel start instead of the read-back
	std::vector<cudaEvent_t> traceEv;   // 4 per frame
	int traceFrames{0};
	int traceFrameOfSlot[kSlots]{};
	int deferredSlot{-1};
	cudaEvent_t evCopied[kSlots]{};
	int outSlot{0};                     // store targets and counters live in a ring of kSlots: up to kSlots - 1 finished frames may still be on their way to the host
	cudaEvent_t evStage[8]{};
	int profiling{0};                   // 0 off, 1 tile kernel + whole frame, 2 every stage
	float stageMs[7]{};

	ApproxLuts hostLuts;
	ApproxLuts* devLuts{nullptr};

	// recording
	bool inFrame{false};
	bool framePending{false};
	int width{0}, height{0}, refTileW{64}, refTileH{64}, postTileW{64}, postTileH{64};
	RsrState curState;                 // copy of the caller's state (direct API); packed streams are read in place
	const RsrState* curStatePtr{nullptr};
	bool haveState{false};
	bool stateDirty{true};
	DevRef curBuffers[16];
	size_t curBufferFloats[16]{};
	HostTex curTus[2];
	DevRef curTu3; int curTu3dim{256};
	std::vector<HostState> states;
	std::vector<HostDraw> draws;
	std::vector<FrameCmd> cmds;
	std::vector<int> cmdDstKind;      // 0 none, 1 tc, 2 fp, 3 depth
	std::vector<PendingCopy> copies;
	uint64_t trianglesSubmitted{0};
	uint64_t inputBytes{0};            // vertex SoA floats bound x vertices referenced + indices + instance matrices

	UploadArena arenas[kSlots];       // ring, indexed like the store targets (outSlot): frame N+1 records while frame N's upload is in flight, and the
	                                  // device mirror of a frame stays intact until three frames later (an overflowed frame can be launched again)
	cudaEvent_t arenaFree[kSlots]{};
	std::unordered_map<const void*, StaticAlloc> staticCache;
	std::vector<PtrCacheEntry> ptrCache = std::vector<PtrCacheEntry>(kPtrCacheSize, PtrCacheEntry{nullptr, 0, nullptr, 0.0f, 0.0f, false});
	std::vector<FrameCacheEntry> frameCache = std::vector<FrameCacheEntry>(kPtrCacheSize);
	std::unordered_map<const void*, PinnedInPlace> pinned;   // RSRCU_UPLOAD_FRAME arrays of >= 1 MiB
	cudaStream_t h2dStream{nullptr};
	cudaEvent_t evH2D{nullptr};
	bool h2dPending{false};            // copies of the frame being recorded that its first kernel must wait for
	bool frameUsedPinned{false};       // (such a frame cannot be retained: its inputs live in buffers later frames overwrite)
	bool pinInPlace{false};            // rsrcu_set_pin_in_place / RSRCU_PIN_IN_PLACE=1: opt-in, see include/rsrcu.h
	uint64_t frameStamp{0};            // bumped by rsrcu_begin_frame: entries of earlier frames are stale without clearing
	float guardFactor{1.0f};
	uint64_t drawsCulled{0};
	uint32_t progMask{0};

	// device work buffers
	// the frame's intermediate buffers; two sets so that, in overlap mode, the front end (K0-K5) of frame N+1 can fill one
	// set while the tile kernel of frame N still reads the other
	struct WorkSet { DevBuf ptvb, vflags, triInfo, triRecs, clipRecs, tileBase, cellRel, tileTotal, tileOrder, lists, largeItems, runScratch; } sets[2];
	cudaStream_t tileStream2{nullptr};  // overlap mode: the tile kernels of odd frames run here, so that the tail of frame N's tile kernel and the head of frame N+1's share the GPU
	cudaEvent_t evGate{nullptr};
	bool tile2Pending{false};           // work on tileStream2 that `stream` has not been ordered behind yet
	cudaStream_t frontStream{nullptr};   // overlap mode: K0-K5 run here (high priority), the tile kernel on `stream`
	cudaEvent_t evFrontDone[2]{}, evTileDone[2]{};
	bool overlap{false};
	uint64_t frameNo{0};
	FramePlan slotPlan[kSlots];          // plan of the frame recorded into each slot (the last one: slotPlan[lastArena])
	Launched launched[kSlots];
	int lastArena{-1};
	uint64_t framesRetried{0};
	unsigned int* waitTimedOut{nullptr};        // device flag of wait_counter_kernel
	DevBuf counters[kSlots];   // Counters | cellCount[] | cellCursor[]; alternate per frame like the store targets (read back while the next frame runs)
	// store targets: one device buffer per store command of a frame (a frame may hold several stores of one kind with
	// draws in between: each keeps its own image), a pool per slot of the ring, reused by the frames that follow
	std::vector<DevBuf> storePool[kSlots];
	int storesUsed{0};
	void* tcDev{nullptr};              // recording: the frame's last true-colour target
	void* shownTcDev{nullptr}; int shownTcStride{0};   // of the frame launched last (rsrcu_device_truecolor)
	uint32_t clipCapacity{1u << 16};
	uint32_t listCapacity{1u << 24};
	uint32_t largeCapacity{1u << 16};
	int largeTiles{kLargeTiles};
	int forceGroupShift{0};            // RSRCU_GROUP_SHIFT (tests): list cells of 1 << shift triangles even in small frames
	int tcStride{0};
	Counters* hostCounters{nullptr};   // pinned, [kSlots]
	RsrStats stats{};
	uint64_t launches{0};
	uint64_t lastH2D{0}, lastD2H{0};
	std::chrono::steady_clock::time_point tBegin{};
	uint64_t recordNs{0}, submitNs{0};
	// RSRCU_HOST_PROF=1: where the submitting thread's time goes (averages printed by rsrcu_destroy)
	// device canvases (rsrcu_canvas_alloc), cross-context ordering (rsrcu_wait_for), post filters, marching cubes
	std::vector<void*> canvases;
	cudaEvent_t evExt{nullptr};          // recorded on a producer context's stream; the next frame's first stream waits for it
	bool extWaitPending{false};
	DevBuf spanBuf, glowOut;
	float stageStartMs[8]{};             // profiling level 2: time of evStage[i] after evStage[0]
	bool stageSpansValid{false};
	DevBuf mcBlocks, mcTotals, mcBase, mcVerts[3];   // rsrcu_march_surface
	int mcTurn{0};
	bool hostProf{false};
	uint64_t hp[6]{};   // layout, tables, plan, reserve, launches, read-back enqueue (ns, summed)
	uint64_t hpFrames{0};
	std::chrono::steady_clock::time_point hpLast{};
};

namespace {

// the tile kernel keeps the rcpps table at 16 bits per entry (dev_math.cuh: rcp_entry16)
bool rcpTableFits16(const uint32_t* rcp2048) {
	for (int i = 0; i < 2048; ++i) { if ((rcp2048[i] >> 23) != 126u || (rcp2048[i] & 0x7fu) != 0u) { return false; } }
	return true; }

int keyOf(const RsrState& s) {
	uint32_t key = 0;
	key |= static_cast<uint32_t>(s.scissor_enabled != 0);
	key |= (s.depth_test_enabled != 0) << 1;
	if (s.depth_test_enabled) { key |= s.depth_func << 2; }
	key |= (s.blending_enabled != 0) << 4;
	key |= (s.depth_write_mask != 0) << 5;
	key |= (s.color_write_mask != 0) << 6;
	key |= s.color0_attachment_type << 7;
	key |= s.depth_attachment_type << 9;
	return static_cast<int>(key); }

// the reference's dispatch tables (src/viewer/shaders.cxx:54-126, shaders_envmap.cxx:18-31,
// shaders_wireframe.cxx:18-23): (program id, FragmentStateKey) pairs that have a tile program
bool drawProgramInstalled(int programId, int key) {
	static const struct { int id; int key; } table[] = {
		{4, 0x6e2}, {4, 0x62}, {4, 0x72},
		{65, 0x6e2}, {65, 0x62}, {65, 0x72},
		{26, 0x72},
		{41, 0x6e2}, {41, 0x62}, {41, 0x72},
		{5, 0x62}, {6, 0x62}, {7, 0x62}, {8, 0x62}, {9, 0x62},
		{11, 0x62},
		{10, 0x22}, {10, 0x5a}, {10, 0x6e2}, {10, 0x62}, {10, 0x72}, {10, 0x50},
		{0, 0x6a2} };   // BaseProgram, depth only: the shadow-map GPU of a `$layer` (src/viewer/node/gllayer.cxx:44-45)
	for (const auto& e : table) { if (e.id == programId && e.key == key) { return true; } }
	return false; }

int programVaryings(int programId) {
	switch (programId) {
	case 4: case 65: case 5: case 6: case 10: return 2;
	case 26: return 5;
	case 7: return 3;
	case 8: return 11;
	case 9: return 15;
	default: return 0; } }

bool bltProgramInstalled(int programId) { return programId == 1 || programId == 2 || programId == 3; }

// ---- host-side matrix preparation, same operation order as the reference ----------------------

// rmlm::operator*(mat4, mat4) (rmlm_mat4.hxx:197-207), column-major
void mat4Mul(const float* lhs, const float* rhs, float* out) {
	for (int row = 0; row < 4; ++row) {
		for (int col = 0; col < 4; ++col) {
			float ax;
			ax = lhs[0 * 4 + row] * rhs[col * 4 + 0];
			ax += lhs[1 * 4 + row] * rhs[col * 4 + 1];
			ax += lhs[2 * 4 + row] * rhs[col * 4 + 2];
			ax += lhs[3 * 4 + row] * rhs[col * 4 + 3];
			out[col * 4 + row] = ax; } } }

// rmlm::inverse (rmlm_mat4.cxx:25-186): cofactor expansion (the gluInvertMatrix scheme).  Each
// output element is a signed sum of six triple products, accumulated left to right; the table
// lists the factors and signs in that order so the float result is identical.
void mat4Inverse(const float* m, float* inv) {
	struct Term { int s, a, b, c; };
	static const struct { int out; Term t[6]; } rows[16] = {
		{0,  {{+1,5,10,15},{-1,5,11,14},{-1,9,6,15},{+1,9,7,14},{+1,13,6,11},{-1,13,7,10}}},
		{4,  {{-1,4,10,15},{+1,4,11,14},{+1,8,6,15},{-1,8,7,14},{-1,12,6,11},{+1,12,7,10}}},
		{8,  {{+1,4,9,15},{-1,4,11,13},{-1,8,5,15},{+1,8,7,13},{+1,12,5,11},{-1,12,7,9}}},
		{12, {{-1,4,9,14},{+1,4,10,13},{+1,8,5,14},{-1,8,6,13},{-1,12,5,10},{+1,12,6,9}}},
		{1,  {{-1,1,10,15},{+1,1,11,14},{+1,9,2,15},{-1,9,3,14},{-1,13,2,11},{+1,13,3,10}}},
		{5,  {{+1,0,10,15},{-1,0,11,14},{-1,8,2,15},{+1,8,3,14},{+1,12,2,11},{-1,12,3,10}}},
		{9,  {{-1,0,9,15},{+1,0,11,13},{+1,8,1,15},{-1,8,3,13},{-1,12,1,11},{+1,12,3,9}}},
		{13, {{+1,0,9,14},{-1,0,10,13},{-1,8,1,14},{+1,8,2,13},{+1,12,1,10},{-1,12,2,9}}},
		{2,  {{+1,1,6,15},{-1,1,7,14},{-1,5,2,15},{+1,5,3,14},{+1,13,2,7},{-1,13,3,6}}},
		{6,  {{-1,0,6,15},{+1,0,7,14},{+1,4,2,15},{-1,4,3,14},{-1,12,2,7},{+1,12,3,6}}},
		{10, {{+1,0,5,15},{-1,0,7,13},{-1,4,1,15},{+1,4,3,13},{+1,12,1,7},{-1,12,3,5}}},
		{14, {{-1,0,5,14},{+1,0,6,13},{+1,4,1,14},{-1,4,2,13},{-1,12,1,6},{+1,12,2,5}}},
		{3,  {{-1,1,6,11},{+1,1,7,10},{+1,5,2,11},{-1,5,3,10},{-1,9,2,7},{+1,9,3,6}}},
		{7,  {{+1,0,6,11},{-1,0,7,10},{-1,4,2,11},{+1,4,3,10},{+1,8,2,7},{-1,8,3,6}}},
		{11, {{-1,0,5,11},{+1,0,7,9},{+1,4,1,11},{-1,4,3,9},{-1,8,1,7},{+1,8,3,5}}},
		{15, {{+1,0,5,10},{-1,0,6,9},{-1,4,1,10},{+1,4,2,9},{+1,8,1,6},{-1,8,2,5}}} };
	for (const auto& r : rows) {
		float acc = 0.0f;
		for (int k = 0; k < 6; ++k) {
			const Term& t = r.t[k];
			volatile float p = m[t.a] * m[t.b];
			volatile float q = p * m[t.c];
			const float term = q;
			if (k == 0) { acc = (t.s > 0) ? term : -term; }
			else { acc = (t.s > 0) ? (acc + term) : (acc - term); } }
		inv[r.out] = acc; }
	const float det = ((m[0] * inv[0] + m[1] * inv[4]) + m[2] * inv[8]) + m[3] * inv[12];
	const float invdet = 1.0f / det;
	for (int i = 0; i < 16; ++i) { inv[i] *= invdet; } }

void mat4Transpose(const float* m, float* out) {
	for (int c = 0; c < 4; ++c) { for (int r = 0; r < 4; ++r) { out[c * 4 + r] = m[r * 4 + c]; } } }

const void* resolve(const rsrcu_ctx* c, const DevRef& r) {
	if (r.null) { return nullptr; }
	if (r.arena) { return static_cast<const uint8_t*>(c->arenas[c->outSlot].dev.ptr) + r.off; }
	return r.abs; }

template <class T>
const T* encodeRef(const DevRef& r, bool& anyArena) {
	if (r.null) { return nullptr; }
	if (r.arena) { anyArena = true; return reinterpret_cast<const T*>(static_cast<uintptr_t>(kArenaRefBit | r.off)); }
	return static_cast<const T*>(r.abs); }

template <class T>
void patchRef(const T*& p, const uint8_t* arenaDev) {
	const uint64_t v = reinterpret_cast<uintptr_t>(p);
	if (v & kArenaRefBit) { p = reinterpret_cast<const T*>(arenaDev + (v & ~kArenaRefBit)); } }

// value range of a float array (bounding boxes of static meshes); false if it holds a NaN
bool floatRange(const float* p, size_t n, float& lo, float& hi) {
	if (n == 0) { return false; }
	float a = p[0], b = p[0];
	bool ok = true;
	for (size_t i = 0; i < n; ++i) { const float v = p[i]; ok = ok && (v == v); a = v < a ? v : a; b = v > b ? v : b; }
	lo = a; hi = b;
	return ok; }

int uploadData(rsrcu_ctx* c, const void* host, size_t bytes, int upload, DevRef& out, bool wantRange = false) {
	if (!host || bytes == 0) { out = DevRef{}; return RSRCU_OK; }
	if (upload == RSRCU_UPLOAD_DEVICE) {   // already on the device (a canvas, marching-cubes output): used in place
		out = DevRef{}; out.null = false; out.arena = false; out.abs = host;
		return RSRCU_OK; }
	if (upload == RSRCU_UPLOAD_STATIC) {
		PtrCacheEntry& pe = c->ptrCache[ptrCacheIndex(host)];
		if (pe.host == host && pe.bytes == bytes && (pe.ranged || !wantRange)) {
			out.null = false; out.arena = false; out.abs = pe.dev; out.lo = pe.lo; out.hi = pe.hi; out.ranged = pe.ranged;
			return RSRCU_OK; }
		auto it = c->staticCache.find(host);
		if (it != c->staticCache.end() && it->second.bytes == bytes) {
			StaticAlloc& sa = it->second;
			if (wantRange && !sa.ranged) { sa.ranged = floatRange(static_cast<const float*>(host), bytes / 4, sa.lo, sa.hi); if (!sa.ranged) { sa.lo = 0.0f; sa.hi = 0.0f; } }
			pe = PtrCacheEntry{host, bytes, sa.dev, sa.lo, sa.hi, sa.ranged || wantRange};   // (a failed range is not tried again: ranged stays false in `out`)
			out.null = false; out.arena = false; out.abs = sa.dev; out.lo = sa.lo; out.hi = sa.hi; out.ranged = sa.ranged;
			pe.ranged = sa.ranged;
			return RSRCU_OK; }
		if (it != c->staticCache.end()) { cudaFree(it->second.dev); c->staticCache.erase(it); }
		void* d = nullptr;
		CU(cudaSetDevice(c->device));
		CU(cudaMalloc(&d, bytes));
		CU(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
		StaticAlloc sa{d, bytes, 0.0f, 0.0f, false};
		if (wantRange) { sa.ranged = floatRange(static_cast<const float*>(host), bytes / 4, sa.lo, sa.hi); if (!sa.ranged) { sa.lo = 0.0f; sa.hi = 0.0f; } }
		c->staticCache[host] = sa;
		pe = PtrCacheEntry{host, bytes, d, sa.lo, sa.hi, sa.ranged};
		out.null = false; out.arena = false; out.abs = d; out.lo = sa.lo; out.hi = sa.hi; out.ranged = sa.ranged;
		return RSRCU_OK; }
	if (upload == RSRCU_UPLOAD_FRAME) {
		// the reference's own contract: GL records the pointer, the renderer reads it at Run -- every bind of one pointer
		// inside a frame sees the same bytes, so they are staged once (a field of 576 quads binds its two textures 576 times)
		FrameCacheEntry& fe = c->frameCache[ptrCacheIndex(host)];
		if (fe.stamp == c->frameStamp && fe.host == host && fe.bytes >= bytes) {
			out.null = false; out.ranged = false;
			if (fe.off == kPinnedMark) { out.arena = false; out.abs = c->pinned[host].dev[c->outSlot].ptr; }
			else { out.arena = true; out.off = fe.off; }
			return RSRCU_OK; }
		if (c->pinInPlace && bytes >= kPinInPlaceMin) {
			PinnedInPlace& pp = c->pinned[host];
			if (pp.bytes != bytes) {   // first sight, or the address now holds something else
				if (pp.owned) { cudaHostUnregister(reinterpret_cast<void*>(pp.regBase)); cudaGetLastError(); }
				pp.bytes = bytes;
				// Only the pages that lie wholly inside the array are locked: its first and last page may be shared with heap
				// neighbours, and a page-locked page poisons every later copy to or from whatever else lives on it (the copy
				// engine refuses sources / destinations that are only partly registered).  The two edge fragments (< 4 KiB
				// each) are copied as pageable memory.
				pp.regBase = (reinterpret_cast<uintptr_t>(host) + 4095) & ~static_cast<uintptr_t>(4095);
				const uintptr_t regEnd = (reinterpret_cast<uintptr_t>(host) + bytes) & ~static_cast<uintptr_t>(4095);
				pp.regBytes = regEnd > pp.regBase ? regEnd - pp.regBase : 0;
				CU(cudaSetDevice(c->device));
				pp.owned = pp.regBytes != 0 && cudaHostRegister(reinterpret_cast<void*>(pp.regBase), pp.regBytes, cudaHostRegisterDefault) == cudaSuccess;
				if (!pp.owned) { cudaGetLastError(); }
				pp.ok = pp.owned; }
			pp.lastFrame = c->frameStamp;
			if (pp.ok) {
				CU(pp.dev[c->outSlot].reserve(bytes));
				// the slot's buffer was last read by the frame three submissions ago: its kernels are done when it has been read back
				CU(cudaStreamWaitEvent(c->h2dStream, c->evCopied[c->outSlot], 0));
				uint8_t* dev = static_cast<uint8_t*>(pp.dev[c->outSlot].ptr);
				const uint8_t* src = static_cast<const uint8_t*>(host);
				const size_t head = pp.regBase - reinterpret_cast<uintptr_t>(host), tailOff = head + pp.regBytes;
				cudaError_t ce = cudaMemcpyAsync(dev + head, src + head, pp.regBytes, cudaMemcpyHostToDevice, c->h2dStream);
				if (ce == cudaSuccess && head) { ce = cudaMemcpyAsync(dev, src, head, cudaMemcpyHostToDevice, c->h2dStream); }
				if (ce == cudaSuccess && tailOff < bytes) { ce = cudaMemcpyAsync(dev + tailOff, src + tailOff, bytes - tailOff, cudaMemcpyHostToDevice, c->h2dStream); }
				if (ce == cudaSuccess) {
					c->h2dPending = true; c->frameUsedPinned = true;
					fe = FrameCacheEntry{host, bytes, kPinnedMark, c->frameStamp};
					out.null = false; out.arena = false; out.abs = pp.dev[c->outSlot].ptr; out.ranged = false;
					return RSRCU_OK; }
				// (another party has page-locked part of the range?) -- stage this array like a small one from now on
				cudaGetLastError();
				if (std::getenv("RSRCU_PIN_DEBUG")) {
					std::fprintf(stderr, "rsrcu: in-place copy of %p (%zu bytes; registered %p + %zu, owned %d) refused: %s; staging it\n", host, bytes,
					             reinterpret_cast<void*>(pp.regBase), pp.regBytes, pp.owned ? 1 : 0, cudaGetErrorString(ce)); }
				if (pp.owned) { cudaHostUnregister(reinterpret_cast<void*>(pp.regBase)); cudaGetLastError(); pp.owned = false; }
				pp.ok = false; } }
		size_t off = 0;
		CU(c->arenas[c->outSlot].push(host, bytes, off));
		fe = FrameCacheEntry{host, bytes, off, c->frameStamp};
		out.null = false; out.arena = true; out.off = off; out.ranged = false;
		return RSRCU_OK; }
	size_t off = 0;
	CU(c->arenas[c->outSlot].push(host, bytes, off));
	out.null = false; out.arena = true; out.off = off; out.ranged = false;
	return RSRCU_OK; }

// Is the box [lo, hi] (object space), transformed by vpm, outside one plane of the reference's guard-band frustum
// (ViewFrustum::Test, rglv_view_frustum.hxx:60-73) with all of its corners?  Then every vertex of the draw carries
// that plane's flag, every triangle is dropped by the "all three share a flag" test (rglv_gpu_impl.hxx:427-433), and
// the draw can be skipped: nothing of it reaches a tile.  The plane values are linear in the object coordinates, so
// the corners bound them; evaluated in double with a margin far above the float rounding of the vertex kernel, so
// the decision never disagrees with the per-vertex flags.
bool boxOutsideFrustum(const float* vpm, const float* lo, const float* hi, float guardFactor) {
	double val[5][8], mag[5][8];
	for (int k = 0; k < 8; ++k) {
		const double x = (k & 1) ? hi[0] : lo[0], y = (k & 2) ? hi[1] : lo[1], z = (k & 4) ? hi[2] : lo[2];
		double c[4], m[4];
		for (int r = 0; r < 4; ++r) {
			c[r] = vpm[r] * x + vpm[4 + r] * y + vpm[8 + r] * z + vpm[12 + r];
			m[r] = std::fabs(vpm[r] * x) + std::fabs(vpm[4 + r] * y) + std::fabs(vpm[8 + r] * z) + std::fabs(static_cast<double>(vpm[12 + r])); }
		const double g = guardFactor;
		val[0][k] = g * c[3] + c[0]; mag[0][k] = g * m[3] + m[0];
		val[1][k] = g * c[3] + c[1]; mag[1][k] = g * m[3] + m[1];
		val[2][k] = c[3] + c[2];     mag[2][k] = m[3] + m[2];
		val[3][k] = g * c[3] - c[0]; mag[3][k] = g * m[3] + m[0];
		val[4][k] = g * c[3] - c[1]; mag[4][k] = g * m[3] + m[1]; }
	for (int p = 0; p < 5; ++p) {
		bool all = true;
		for (int k = 0; k < 8; ++k) { all = all && (val[p][k] <= -(1e-4 * mag[p][k] + 1e-30)); }   // (false for NaN / inf)
		if (all) { return true; } }
	return false; }

// GL::MaybeUpdateState (rglv_gl.cxx:100-106): snapshot on first use after a change.  The device
// record (DevState) is built here, once; rsrcu_end_frame only copies it into the upload arena.
int snapshotState(rsrcu_ctx* c) {
	if (!c->haveState) { return fail(RSRCU_ERR_INVALID, "no state set (rsrcu_set_state) before a command"); }
	if (!c->stateDirty && !c->states.empty()) { return RSRCU_OK; }
	c->states.emplace_back();
	HostState& hs = c->states.back();
	const RsrState& s = *c->curStatePtr;
	DevState& ds = hs.ds;
	const int W = c->width, H = c->height;
	std::memcpy(ds.vm, s.view_matrix, sizeof(ds.vm));
	std::memcpy(ds.pm, s.projection_matrix, sizeof(ds.pm));
	// MakeMatrices (rglv_gpu_impl.hxx:45-51)
	if (s.program_id == 9 || s.program_id == 10) {   // only OBJ2S and Envmap read gl_NormalMatrix
		float inv[16];
		mat4Inverse(s.view_matrix, inv);
		mat4Transpose(inv, ds.nm); }
	mat4Mul(s.projection_matrix, s.view_matrix, ds.vpm);
	if (s.uniforms_valid) { std::memcpy(ds.uniforms, s.uniforms, sizeof(ds.uniforms)); }
	// GPU::DSDO (rglv_gpu.hxx:265-270): integer halves
	const int vw = (s.viewport_size[0] > 0 && s.viewport_size[1] > 0) ? s.viewport_size[0] : W;
	const int vh = (s.viewport_size[0] > 0 && s.viewport_size[1] > 0) ? s.viewport_size[1] : H;
	ds.DSx = static_cast<float>(vw / 2);
	ds.DSy = static_cast<float>(-vh / 2);
	ds.DOx = static_cast<float>(vw / 2 + s.viewport_origin[0]);
	ds.DOy = static_cast<float>(H - (vh / 2 + s.viewport_origin[1]));
	std::memcpy(ds.clearColor, s.clear_color, sizeof(ds.clearColor));
	ds.clearDepth = s.clear_depth;
	ds.programId = s.program_id;
	ds.cullingEnabled = s.culling_enabled; ds.cullFace = s.cull_face;
	if (!s.scissor_enabled) { ds.scissorX0 = 0; ds.scissorY0 = 0; ds.scissorX1 = W; ds.scissorY1 = H; }
	else {
		// gl_offset_and_size_to_irect (rglv_gpu.hxx:258-263)
		const int left = s.scissor_origin[0], right = left + s.scissor_size[0];
		const int bottom = H - s.scissor_origin[1] - 1, top = bottom - s.scissor_size[1];
		ds.scissorX0 = left; ds.scissorY0 = top; ds.scissorX1 = right; ds.scissorY1 = bottom; }
	ds.depthTest = s.depth_test_enabled; ds.depthFunc = s.depth_func; ds.depthWrite = s.depth_write_mask;
	ds.colorWrite = s.color_write_mask; ds.blend = s.blending_enabled;
	ds.color0Type = s.color0_attachment_type; ds.depthType = s.depth_attachment_type;
	bool anyArena = false;
	for (int b = 0; b < 16; ++b) {
		ds.buffers[b] = encodeRef<float>(c->curBuffers[b], anyArena);
		hs.bufferFloats[b] = c->curBufferFloats[b]; }
	for (int u = 0; u < 2; ++u) {
		const HostTex& ht = c->curTus[u];
		TexUnit& tu = ds.tu[u];
		tu.texels = encodeRef<float4>(ht.ref, anyArena);
		tu.texelCount = ht.texelCount;
		tu.width = ht.width; tu.height = ht.height; tu.stride = ht.stride;
		// MakeTextureUnit (rglr_texture_sampler.cxx:314-363)
		int power = 0;
		while ((1 << (power + 1)) <= ht.width) { ++power; }
		const bool isPow2 = ((1 << power) == ht.width) && ht.width == ht.height && ht.stride == ht.width;
		tu.power = power;
		tu.kind = !isPow2 ? 0 : (ht.filter ? 2 : 1); }
	ds.tu3 = encodeRef<float>(c->curTu3, anyArena);
	ds.tu3dim = c->curTu3dim;
	hs.posRanged = true;
	for (int b = 0; b < 3; ++b) { hs.posRanged = hs.posRanged && !c->curBuffers[b].null && c->curBuffers[b].ranged; hs.posLo[b] = c->curBuffers[b].lo; hs.posHi[b] = c->curBuffers[b].hi; }
	hs.programId = s.program_id;
	hs.key = keyOf(s);
	hs.posNull = c->curBuffers[0].null;
	hs.tex0Null = c->curTus[0].ref.null;
	hs.tex0Width = c->curTus[0].width; hs.tex0Height = c->curTus[0].height; hs.tex0Stride = c->curTus[0].stride;
	hs.hasArenaRef = anyArena;
	c->stateDirty = false;
	return RSRCU_OK; }

// every kernel of the frame is launched with programmatic stream serialization (see pdl_wait in kernels.cuh)
// Instantiations of the tile kernel: the general one carries all eleven programs; frames whose draws use a subset get a
// kernel compiled for just those programs (registers are allotted for the programs present, not for the hungriest of all)
using TileKernelFn = void (*)(const TileArgs);
struct TileVariant { uint32_t progs; TileKernelFn fn; const char* name; };
constexpr uint32_t kPAmy = prog_bit(ProgAmy::id), kPOBJ2 = prog_bit(ProgOBJ2::id), kPMany = prog_bit(ProgMany::id), kPBase = prog_bit(ProgBase::id);
#ifndef RSR_TILE_CTAS_AMY
#define RSR_TILE_CTAS_AMY 3   // (4 CTAs / 64 registers: measured slower on c3 and c4, spills)
#endif
#ifndef RSR_TILE_CTAS_LIT
#define RSR_TILE_CTAS_LIT 4   // c2: 121 -> 117 us
#endif
const TileVariant kTileVariants[] = {
	{ kPBase, tile_kernel<kPBase, RSR_TILE_CTAS_LIT>, "base (depth only)" },
	{ kPAmy, tile_kernel<kPAmy, RSR_TILE_CTAS_AMY>, "amy" },
	{ kPOBJ2 | kPMany, tile_kernel<kPOBJ2 | kPMany, RSR_TILE_CTAS_LIT>, "obj2+many" },
	{ kPAmy | kPOBJ2 | kPMany, tile_kernel<kPAmy | kPOBJ2 | kPMany, RSR_TILE_CTAS_LIT>, "amy+obj2+many" },
	{ kAllProgs, tile_kernel<kAllProgs, RSR_TILE_CTAS>, "all" } };

const TileVariant& tileVariantFor(uint32_t progMask) {
	static const int forced = std::getenv("RSRCU_TILE_VARIANT") ? std::atoi(std::getenv("RSRCU_TILE_VARIANT")) : -1;
	constexpr int n = static_cast<int>(sizeof(kTileVariants) / sizeof(kTileVariants[0]));
	if (forced >= 0 && forced < n && (progMask & ~kTileVariants[forced].progs) == 0u) { return kTileVariants[forced]; }
	for (const TileVariant& v : kTileVariants) { if ((progMask & ~v.progs) == 0u) { return v; } }
	return kTileVariants[n - 1]; }

// (off for the front-end kernels of a frame in overlap mode: a dependent kernel that is resident early only to sit in
// griddepcontrol.wait holds registers and thread slots the concurrent tile kernel of the previous frame could use)
thread_local bool g_pdl = true;   // (contexts may be driven from different host threads: rsr_b200.SubmitPool)

template <class... KArgs, class... Args>
cudaError_t launchPdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args&&... args) {
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(grid, 1, 1);
	cfg.blockDim = dim3(block, 1, 1);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = g_pdl ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...); }

// orders the context's stream behind the tile kernels that went to the second tile stream (overlap mode)
int joinTileStreams(rsrcu_ctx* c) {
	if (c->tile2Pending) {
		CU(cudaStreamWaitEvent(c->stream, c->evTileDone[1], 0));
		c->tile2Pending = false; }
	return RSRCU_OK; }

int flushDeferredCopies(rsrcu_ctx* c) {
	if (c->deferredSlot < 0) { return RSRCU_OK; }
	const int slot = c->deferredSlot;
	c->deferredSlot = -1;
	CU(cudaStreamWaitEvent(c->copyStream, c->evRendered[slot], 0));
	const int tf = c->trace ? c->traceFrameOfSlot[slot] : -1;
	if (tf >= 0 && !c->traceStages) { CU(cudaEventRecord(c->traceEv[4 * tf + 2], c->copyStream)); }
	for (const PendingCopy& pc : c->deferredCopies) {
		if (pc.hostPitch == pc.rowBytes && pc.devPitch == pc.rowBytes) {   // contiguous on both sides: one linear copy (a 2-D copy pays per row)
			CU(cudaMemcpyAsync(pc.hostDst, pc.devSrc, pc.rowBytes * pc.rows, cudaMemcpyDeviceToHost, c->copyStream)); }
		else {
			CU(cudaMemcpy2DAsync(pc.hostDst, pc.hostPitch, pc.devSrc, pc.devPitch, pc.rowBytes, pc.rows, cudaMemcpyDeviceToHost, c->copyStream)); } }
	CU(cudaMemcpyAsync(c->hostCounters + slot, c->counters[slot].ptr, sizeof(Counters), cudaMemcpyDeviceToHost, c->copyStream));
	CU(cudaEventRecord(c->evCopied[slot], c->copyStream));
	if (tf >= 0 && !c->traceStages) { CU(cudaEventRecord(c->traceEv[4 * tf + 3], c->copyStream)); }
	return RSRCU_OK; }

// Launches the kernels of a frame whose tables (plan) are final: K0 (upload of `uploadBytes` from `hostArena`, or
// nothing for a retained frame whose tables already live on the device, plus the zeroed control block), K1-K6,
// and the read-back.  `arenaDev` is where the tables live on the device.
int launchFrame(rsrcu_ctx* c, const FramePlan& plan, const uint8_t* arenaDev, const uint8_t* hostArena, size_t uploadBytes, int arenaIdx) {
	// (the capacities are the context's current ones: a frame launched again after an overflow, or a retained frame
	// replayed after one, runs with the grown buffers)
	FrameParams fp = plan.fp;
	fp.largeCapacity = c->largeCapacity; fp.largeTiles = c->largeTiles; fp.clipCapacity = c->clipCapacity; fp.listCapacity = c->listCapacity;
	c->launched[c->outSlot] = Launched{&plan, arenaDev, false};
	c->shownTcDev = plan.tcDev; c->shownTcStride = plan.tcStride;
	const int ntiles = fp.tilesX * fp.tilesY;
	const uint64_t ptvbF4 = plan.ptvbF4, nvertsTotal = plan.nvertsTotal, pjobs = plan.pjobs;
	const size_t offStates = plan.offStates, offDraws = plan.offDraws, offCmds = plan.offCmds, offVBlocks = plan.offVBlocks, offPBlocks = plan.offPBlocks;
	// overlap mode: the front end of this frame runs on its own high-priority stream into work set (frame & 1) and only
	// the tile kernel on the context's stream, so K0-K5 of frame N+1 fill the SMs the tile kernel of frame N leaves idle
	const int si = c->overlap ? static_cast<int>(c->frameNo & 1) : 0;
	++c->frameNo;
	rsrcu_ctx::WorkSet& w = c->sets[si];
	cudaStream_t st = c->overlap ? c->frontStream : c->stream;
	static const bool dualTile = std::getenv("RSRCU_TILE_STREAMS") ? std::atoi(std::getenv("RSRCU_TILE_STREAMS")) == 2 : true;
	cudaStream_t tileStream = (c->overlap && dualTile && si == 1) ? c->tileStream2 : c->stream;
	if (tileStream != c->stream) { c->tile2Pending = true; }
	c->launches = 0;

	// ---- device buffers -------------------------------------------------------------------
	CU(w.ptvb.reserve(std::max<uint64_t>(1, ptvbF4) * 16));
	CU(w.vflags.reserve(std::max<uint64_t>(1, nvertsTotal)));
	CU(w.triInfo.reserve(std::max<uint64_t>(1, pjobs) * sizeof(uint2)));
	CU(w.triRecs.reserve(std::max<uint64_t>(1, pjobs) * sizeof(TriRec)));
	CU(w.clipRecs.reserve(static_cast<size_t>(c->clipCapacity) * sizeof(ClipRec)));
	const size_t ncells = static_cast<size_t>(ntiles) * fp.groups;
	const size_t ctrlBytes = 64 + ncells * 8;   // Counters | cellCount[] | cellCursor[]
	static_assert(sizeof(Counters) <= 64, "control block layout");
	CU(c->counters[c->outSlot].reserve(ctrlBytes));
	CU(w.tileBase.reserve((static_cast<size_t>(ntiles) + 1) * 4));
	CU(w.cellRel.reserve(ncells * 4));
	CU(w.tileTotal.reserve(static_cast<size_t>(ntiles) * 4));
	CU(w.tileOrder.reserve(static_cast<size_t>(ntiles) * 4));
	CU(w.lists.reserve(static_cast<size_t>(c->listCapacity) * sizeof(uint2)));
	CU(w.largeItems.reserve(static_cast<size_t>(c->largeCapacity) * sizeof(LargeItem)));
	CU(w.runScratch.reserve(static_cast<size_t>(ntiles) * 2 * kRunCap * sizeof(uint32_t)));

	const uint8_t* ab = arenaDev;
	const DevState* dStates = reinterpret_cast<const DevState*>(ab + offStates);
	const DevDraw* dDraws = reinterpret_cast<const DevDraw*>(ab + offDraws);
	const FrameCmd* dCmds = reinterpret_cast<const FrameCmd*>(ab + offCmds);
	Counters* dCtr = static_cast<Counters*>(c->counters[c->outSlot].ptr);
	BinArgs bin{};
	bin.cellCount = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(c->counters[c->outSlot].ptr) + 64);
	bin.cellCursor = bin.cellCount + ncells;
	bin.tileBase = static_cast<const uint32_t*>(w.tileBase.ptr);
	bin.cellRel = static_cast<const uint32_t*>(w.cellRel.ptr);
	bin.lists = static_cast<uint2*>(w.lists.ptr);
	bin.large = static_cast<LargeItem*>(w.largeItems.ptr);

	if (c->hostProf) { const auto now_ = std::chrono::steady_clock::now(); c->hp[3] += static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(now_ - c->hpLast).count()); c->hpLast = now_; }
	static const bool frontPdl = std::getenv("RSRCU_FRONT_PDL") ? std::atoi(std::getenv("RSRCU_FRONT_PDL")) != 0 : false;
	// overlap mode: the front-end kernels run on at most this many CTAs (grid-stride over their 256-job blocks), about three
	// per SM, so that they share the SMs with the previous frame's tile kernel instead of displacing its CTAs
	static const unsigned frontCtas = std::getenv("RSRCU_FRONT_CTAS") ? static_cast<unsigned>(std::atoi(std::getenv("RSRCU_FRONT_CTAS"))) : 444u;
	const unsigned gridCap = (c->overlap && frontCtas > 0) ? frontCtas : 0xffffffffu;
	g_pdl = !c->overlap || frontPdl;
	if (c->extWaitPending) { CU(cudaStreamWaitEvent(st, c->evExt, 0)); c->extWaitPending = false; }   // rsrcu_wait_for: a producer context's canvases
	bool waitH2D = false;
	if (c->h2dPending) {   // arrays copied in place (uploadData): the frame's first kernel waits for the copy engine
		CU(cudaEventRecord(c->evH2D, c->h2dStream));
		CU(cudaStreamWaitEvent(st, c->evH2D, 0));
		c->h2dPending = false;
		waitH2D = true; }
	if (c->overlap) { CU(cudaStreamWaitEvent(st, c->evTileDone[si], 0)); }   // the tile kernel of the frame that last used this work set (and this arena mirror)
	if (c->profiling) { CU(cudaEventRecord(c->evStage[0], st)); }
	CU(cudaStreamWaitEvent(st, c->evCopied[c->outSlot], 0));   // the frame that last used this slot (counters, store targets) has been read back
	// (a replayed frame has nothing to upload: the vertex kernel zeroes the control block itself)
	const bool zeroInVertexKernel = uploadBytes == 0 && fp.totalVJobs > 0;
	if (!zeroInVertexKernel) {
		// K0: upload + zeroed control block in one kernel (kernels.cuh)
		const size_t n16 = (uploadBytes + 15) / 16, nz16 = (ctrlBytes + 15) / 16;
		const int blocks = static_cast<int>(std::min<size_t>(148 * 8, (std::max(n16, nz16) + 255) / 256));
		CU(launchPdl(upload_kernel, static_cast<unsigned>(std::max(blocks, 1)), 256u, 0, st, reinterpret_cast<const uint4*>(hostArena),
			reinterpret_cast<uint4*>(const_cast<uint8_t*>(arenaDev)), n16, reinterpret_cast<uint4*>(dCtr), nz16));
		++c->launches; }
	if (arenaIdx >= 0) { CU(cudaEventRecord(c->arenaFree[arenaIdx], st)); }
	if (c->profiling > 1) { CU(cudaEventRecord(c->evStage[1], st)); }
	const int traceIdx = (c->trace && c->traceFrames < 64) ? c->traceFrames++ : -1;
	if (traceIdx >= 0) { c->traceFrameOfSlot[c->outSlot] = traceIdx; CU(cudaEventRecord(c->traceEv[4 * traceIdx], st)); }
	else { c->traceFrameOfSlot[c->outSlot] = -1; }

	if (fp.totalVJobs) {
		CU(launchPdl(vertex_kernel, std::min(gridCap, (fp.totalVJobs + 255) / 256), 256u, 0, st, dDraws, reinterpret_cast<const uint32_t*>(ab + offVBlocks), dStates, fp,
			static_cast<const ApproxLuts*>(c->devLuts), static_cast<float4*>(w.ptvb.ptr), static_cast<uint8_t*>(w.vflags.ptr),
			reinterpret_cast<uint4*>(dCtr), zeroInVertexKernel ? static_cast<uint32_t>((ctrlBytes + 15) / 16) : 0u));
		++c->launches; }
	if (c->profiling > 1) { CU(cudaEventRecord(c->evStage[2], st)); }
	if (fp.totalPJobs) {
		CU(launchPdl(setup_kernel, std::min(gridCap, (fp.totalPJobs + 255) / 256), 256u, 0, st, dDraws, reinterpret_cast<const uint32_t*>(ab + offPBlocks), dStates, fp,
			static_cast<const ApproxLuts*>(c->devLuts), static_cast<const float4*>(w.ptvb.ptr), static_cast<const uint8_t*>(w.vflags.ptr),
			static_cast<uint2*>(w.triInfo.ptr), static_cast<TriRec*>(w.triRecs.ptr), static_cast<ClipRec*>(w.clipRecs.ptr),
			bin, static_cast<uint32_t*>(w.tileBase.ptr), static_cast<uint32_t*>(w.tileOrder.ptr), dCtr));
		++c->launches; }
	if (c->profiling > 1) { CU(cudaEventRecord(c->evStage[3], st)); CU(cudaEventRecord(c->evStage[4], st)); }
	if (fp.totalPJobs && fp.groups > 1) {
		CU(launchPdl(cell_scan_kernel, static_cast<unsigned>((ntiles + 7) / 8), 256u, 0, st, fp, static_cast<const uint32_t*>(bin.cellCount), static_cast<uint32_t*>(w.cellRel.ptr),
			static_cast<uint32_t*>(w.tileTotal.ptr), static_cast<uint32_t*>(w.tileBase.ptr), static_cast<uint32_t*>(w.tileOrder.ptr), dCtr));
		++c->launches; }
	if (c->profiling > 1) { CU(cudaEventRecord(c->evStage[5], st)); }
	if (fp.totalPJobs) {
		CU(launchPdl(fill_kernel, std::min(gridCap, (fp.totalPJobs + 255) / 256), 256u, 0, st, fp, static_cast<const uint2*>(w.triInfo.ptr),
			static_cast<const ClipRec*>(w.clipRecs.ptr), bin, dCtr));
		++c->launches; }
	if (traceIdx >= 0 && c->traceStages) { CU(cudaEventRecord(c->traceEv[4 * traceIdx + 2], st)); }
	if (c->overlap) {
		CU(cudaEventRecord(c->evFrontDone[si], st));
		CU(cudaStreamWaitEvent(tileStream, c->evFrontDone[si], 0)); }
	if (traceIdx >= 0 && c->traceStages) { CU(cudaEventRecord(c->traceEv[4 * traceIdx + 3], tileStream)); }
	if (c->profiling) { CU(cudaEventRecord(c->evStage[6], tileStream)); }

	g_pdl = true;
	TileArgs ta{};
	for (int i = 0; i < plan.ncmdInline; ++i) { ta.icmd[i] = plan.icmd[i]; ta.icmdState[i] = plan.icmdState[i]; }
	ta.fp = fp; ta.cmds = dCmds; ta.draws = dDraws; ta.states = dStates; ta.luts = c->devLuts;
	ta.ptvb = static_cast<const float4*>(w.ptvb.ptr);
	ta.triRecs = static_cast<const TriRec*>(w.triRecs.ptr);
	ta.clipRecs = static_cast<const ClipRec*>(w.clipRecs.ptr);
	ta.lists = static_cast<const uint2*>(w.lists.ptr);
	ta.tileBase = bin.tileBase;
	ta.cellRel = bin.cellRel;
	ta.large = bin.large;
	ta.tileOrder = fp.totalPJobs ? static_cast<const uint32_t*>(w.tileOrder.ptr) : nullptr;
	ta.runScratch = static_cast<uint32_t*>(w.runScratch.ptr);
	ta.ctr = dCtr;
	CU(launchPdl(tileVariantFor(plan.progMask).fn, static_cast<unsigned>(ntiles), static_cast<unsigned>(kTileThreads), sizeof(TileShared), tileStream, ta));
	++c->launches;
	CU(cudaGetLastError());
	if (c->profiling) { CU(cudaEventRecord(c->evStage[7], tileStream)); }

	if (c->overlap) { CU(cudaEventRecord(c->evTileDone[si], tileStream)); }

	if (c->hostProf) { const auto now_ = std::chrono::steady_clock::now(); c->hp[4] += static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(now_ - c->hpLast).count()); c->hpLast = now_; }
	c->lastH2D = uploadBytes;
	c->lastD2H = 0;
	CU(cudaEventRecord(c->evRendered[c->outSlot], tileStream));
	if (traceIdx >= 0) { CU(cudaEventRecord(c->traceEv[4 * traceIdx + 1], tileStream)); }
	for (const PendingCopy& pc : plan.copies) { c->lastD2H += pc.rowBytes * pc.rows; }
	c->deferredCopies = plan.copies;
	c->stats.triangles_submitted = plan.trianglesSubmitted;
	c->deferredSlot = c->outSlot;
	c->framePending = true;
	// the read-back goes to the copy stream right away (it waits for evRendered there): the upload is a kernel
	// (K0), so a queued device->host copy cannot hold up the next frame's upload on a copy engine
	{ const int r = flushDeferredCopies(c); if (r != RSRCU_OK) { return r; } }
	if (c->hostProf) { const auto now_ = std::chrono::steady_clock::now(); c->hp[5] += static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(now_ - c->hpLast).count()); c->hpLast = now_; }
	if (c->hostProf) { ++c->hpFrames; }
	// RSRCU_UPLOAD_FRAME promises the host its arrays back when the frame has been submitted: the in-place copies
	// (started while the frame was being recorded) must have left host memory by then
	if (waitH2D) { CU(cudaEventSynchronize(c->evH2D)); }
	return RSRCU_OK; }


}  // namespace

extern "C" {

const char* rsrcu_last_error(void) { return g_lastError.c_str(); }

// RSRCU_SEGV_TRACE=1 (developer aid): a SIGSEGV prints the native call stack (offsets resolve with addr2line on the same .so)
static void segvTrace(int) {
	void* frames[48];
	const int n = backtrace(frames, 48);
	backtrace_symbols_fd(frames, n, 2);
	_exit(139); }

int rsrcu_create(int device, rsrcu_ctx** out) {
	if (!out) { return fail(RSRCU_ERR_INVALID, "out is null"); }
	if (std::getenv("RSRCU_SEGV_TRACE")) { signal(SIGSEGV, segvTrace); }
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		return fail(RSRCU_ERR_NO_DEVICE, "no CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e)); }
	if (device < 0 || device >= n) { return fail(RSRCU_ERR_INVALID, "device %d out of range (%d devices)", device, n); }
	CU(cudaSetDevice(device));
	auto* c = new rsrcu_ctx();
	c->device = device;
	CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	CU(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
	CU(cudaStreamCreateWithFlags(&c->tileStream2, cudaStreamNonBlocking));
	CU(cudaEventCreateWithFlags(&c->evGate, cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&c->evExt, cudaEventDisableTiming));
	CU(cudaStreamCreateWithFlags(&c->h2dStream, cudaStreamNonBlocking));
	CU(cudaEventCreateWithFlags(&c->evH2D, cudaEventDisableTiming));
	if (const char* v = std::getenv("RSRCU_PIN_IN_PLACE")) { c->pinInPlace = std::atoi(v) != 0; }
	{
		int lo = 0, hi = 0;
		CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
		const char* pr = std::getenv("RSRCU_FRONT_PRIO");   // (experiments: 0 = lowest priority for the front-end stream)
		CU(cudaStreamCreateWithPriority(&c->frontStream, cudaStreamNonBlocking, (pr && std::atoi(pr) == 0) ? lo : hi)); }
	for (auto& ev : c->evFrontDone) { CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); }
	for (auto& ev : c->evTileDone) { CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); }
	for (auto& ev : c->evRendered) { CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); }
	for (auto& ev : c->evCopied) { CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); }
	for (auto& ev : c->evStage) { CU(cudaEventCreate(&ev)); }
	for (auto& ev : c->arenaFree) { CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); }
	rsr::harvest_luts(c->hostLuts.rcp, c->hostLuts.rsqrt);
	static std::once_flag once;
	static uint64_t mismatches = 0;
	std::call_once(once, [&]() { mismatches = rsr::verify_luts(c->hostLuts.rcp, c->hostLuts.rsqrt); });
	if (mismatches != 0) {
		delete c;
		return fail(RSRCU_ERR_UNSUPPORTED, "host rcpps/rsqrtps do not follow the table model (%llu mismatches); "
		            "bit-exact parity with the reference on this CPU is not possible", static_cast<unsigned long long>(mismatches)); }
	if (!rcpTableFits16(c->hostLuts.rcp)) {
		delete c;
		return fail(RSRCU_ERR_UNSUPPORTED, "host rcpps results do not fit the 16-bit table the tile kernel uses (exponent 126, <= 16 mantissa bits)"); }
	c->hostProf = std::getenv("RSRCU_HOST_PROF") != nullptr;
	if (std::getenv("RSRCU_TRACE")) {
		c->trace = true;
		c->traceStages = std::atoi(std::getenv("RSRCU_TRACE")) == 2;
		c->traceEv.resize(4 * 64);
		for (auto& ev : c->traceEv) { CU(cudaEventCreate(&ev)); } }
	if (const char* v = std::getenv("RSRCU_LARGE_TILES")) { const int n = std::atoi(v); if (n > 0) { c->largeTiles = n; } }   // tests
	if (const char* v = std::getenv("RSRCU_GROUP_SHIFT")) { const int n = std::atoi(v); if (n >= 5 && n < 31) { c->forceGroupShift = n; } }
	if (const char* cap = std::getenv("RSRCU_LIST_CAPACITY")) {   // initial tile-list capacity in entries (tests)
		const long v = std::atol(cap);
		if (v > 0) { c->listCapacity = static_cast<uint32_t>(v); } }
	for (const TileVariant& v : kTileVariants) { CU(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(TileShared)))); }
	for (int i = 0; i < 2048; ++i) { c->hostLuts.rcp16[i] = static_cast<uint16_t>((c->hostLuts.rcp[i] >> 7) & 0xffffu); }
	CU(cudaMalloc(&c->devLuts, sizeof(ApproxLuts)));
	CU(cudaMemcpy(c->devLuts, &c->hostLuts, sizeof(ApproxLuts), cudaMemcpyHostToDevice));
	CU(cudaMallocHost(&c->hostCounters, kSlots * sizeof(Counters)));
	std::memset(c->hostCounters, 0, kSlots * sizeof(Counters));
	*out = c;
	return RSRCU_OK; }

int rsrcu_destroy(rsrcu_ctx* c) {
	if (!c) { return RSRCU_OK; }
	cudaSetDevice(c->device);
	flushDeferredCopies(c);
	cudaStreamSynchronize(c->frontStream);
	cudaStreamSynchronize(c->stream);
	cudaStreamSynchronize(c->tileStream2);
	cudaStreamSynchronize(c->copyStream);
	if (c->hostProf && c->hpFrames) {
		const double n = static_cast<double>(c->hpFrames) * 1e3;
		std::fprintf(stderr, "rsrcu host profile over %llu frames (us per frame): layout %.1f, tables %.1f, plan %.1f, reserve %.1f, launches %.1f, read-back enqueue %.1f\n",
		             static_cast<unsigned long long>(c->hpFrames), c->hp[0] / n, c->hp[1] / n, c->hp[2] / n, c->hp[3] / n, c->hp[4] / n, c->hp[5] / n); }
	if (c->trace) {
		for (int f = 0; f + 1 < c->traceFrames; ++f) {
			float t[4] = {0, 0, 0, 0};
			for (int k = 0; k < 4; ++k) { if (cudaEventElapsedTime(&t[k], c->traceEv[0], c->traceEv[4 * f + k]) != cudaSuccess) { t[k] = -1.0f; cudaGetLastError(); } }
			if (c->traceStages) { std::fprintf(stderr, "rsrcu trace frame %2d: front end %8.1f .. %8.1f us   tile kernel %8.1f .. %8.1f us\n", f, t[0] * 1e3f, t[2] * 1e3f, t[3] * 1e3f, t[1] * 1e3f); }
			else { std::fprintf(stderr, "rsrcu trace frame %2d: kernels %8.1f .. %8.1f us   read-back %8.1f .. %8.1f us\n", f, t[0] * 1e3f, t[1] * 1e3f, t[2] * 1e3f, t[3] * 1e3f); } }
		for (auto& ev : c->traceEv) { cudaEventDestroy(ev); } }
	for (auto& kv : c->staticCache) { cudaFree(kv.second.dev); }
	for (void* p : c->canvases) { cudaFree(p); }
	if (c->h2dStream) { cudaStreamSynchronize(c->h2dStream); }
	for (auto& kv : c->pinned) {
		if (kv.second.owned) { cudaHostUnregister(reinterpret_cast<void*>(kv.second.regBase)); cudaGetLastError(); }
		for (auto& b : kv.second.dev) { b.release(); } }
	if (c->evH2D) { cudaEventDestroy(c->evH2D); }
	if (c->h2dStream) { cudaStreamDestroy(c->h2dStream); }
	for (DevBuf* b : { &c->spanBuf, &c->glowOut, &c->mcBlocks, &c->mcTotals, &c->mcBase, &c->mcVerts[0], &c->mcVerts[1], &c->mcVerts[2] }) { b->release(); }
	if (c->evExt) { cudaEventDestroy(c->evExt); }
	for (auto& w : c->sets) {
		for (DevBuf* b : { &w.ptvb, &w.vflags, &w.triInfo, &w.triRecs, &w.clipRecs, &w.tileBase, &w.cellRel, &w.tileTotal, &w.tileOrder, &w.lists, &w.largeItems, &w.runScratch }) { b->release(); } }
	for (auto& b : c->counters) { b.release(); }
	for (auto& pool : c->storePool) { for (auto& b : pool) { b.release(); } }
	for (auto& a : c->arenas) { a.release(); }
	for (auto& ev : c->arenaFree) { cudaEventDestroy(ev); }
	if (c->devLuts) { cudaFree(c->devLuts); }
	if (c->waitTimedOut) { cudaFree(c->waitTimedOut); }
	if (c->hostCounters) { cudaFreeHost(c->hostCounters); }
	for (auto& ev : c->evStage) { cudaEventDestroy(ev); }
	for (auto& ev : c->evRendered) { cudaEventDestroy(ev); }
	for (auto& ev : c->evCopied) { cudaEventDestroy(ev); }
	for (auto& ev : c->evFrontDone) { cudaEventDestroy(ev); }
	for (auto& ev : c->evTileDone) { cudaEventDestroy(ev); }
	cudaStreamDestroy(c->copyStream);
	cudaStreamDestroy(c->frontStream);
	cudaStreamDestroy(c->stream);
	cudaStreamDestroy(c->tileStream2);
	delete c;
	return RSRCU_OK; }

int rsrcu_set_host_luts(rsrcu_ctx* c, const uint32_t* rcp2048, const uint32_t* rsqrt2x1024) {
	if (!c || !rcp2048 || !rsqrt2x1024) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	if (!rcpTableFits16(rcp2048)) { return fail(RSRCU_ERR_UNSUPPORTED, "rcpps table does not fit the 16-bit form the tile kernel uses (exponent 126, <= 16 mantissa bits)"); }
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->frontStream));
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaStreamSynchronize(c->tileStream2));
	std::memcpy(c->hostLuts.rcp, rcp2048, sizeof(c->hostLuts.rcp));
	std::memcpy(c->hostLuts.rsqrt, rsqrt2x1024, sizeof(c->hostLuts.rsqrt));
	for (int i = 0; i < 2048; ++i) { c->hostLuts.rcp16[i] = static_cast<uint16_t>((c->hostLuts.rcp[i] >> 7) & 0xffffu); }
	CU(cudaMemcpy(c->devLuts, &c->hostLuts, sizeof(ApproxLuts), cudaMemcpyHostToDevice));
	return RSRCU_OK; }

int rsrcu_get_host_luts(rsrcu_ctx* c, uint32_t* rcp2048, uint32_t* rsqrt2x1024) {
	if (!c || !rcp2048 || !rsqrt2x1024) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	std::memcpy(rcp2048, c->hostLuts.rcp, sizeof(c->hostLuts.rcp));
	std::memcpy(rsqrt2x1024, c->hostLuts.rsqrt, sizeof(c->hostLuts.rsqrt));
	return RSRCU_OK; }

int rsrcu_release_static(rsrcu_ctx* c) {
	if (!c) { return fail(RSRCU_ERR_INVALID, "null context"); }
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->frontStream));
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaStreamSynchronize(c->tileStream2));
	for (auto& kv : c->staticCache) { cudaFree(kv.second.dev); }
	c->staticCache.clear();
	std::fill(c->ptrCache.begin(), c->ptrCache.end(), PtrCacheEntry{nullptr, 0, nullptr, 0.0f, 0.0f, false});
	return RSRCU_OK; }

int rsrcu_begin_frame(rsrcu_ctx* c, int width, int height, int tileWBlocks, int tileHBlocks) {
	if (!c) { return fail(RSRCU_ERR_INVALID, "null context"); }
	if (width <= 0 || height <= 0 || (width & 1) || (height & 1)) {
		return fail(RSRCU_ERR_INVALID, "target %dx%d: dimensions must be positive and even (2x2 quads)", width, height); }
	if (width > 2048 || height > 2048) {
		return fail(RSRCU_ERR_UNSUPPORTED, "target %dx%d: the reference's guard band ends at 2048 px "
		            "(rglv_view_frustum.hxx:36-39); render larger images as sub-frames", width, height); }
	if (tileWBlocks <= 0 || tileHBlocks <= 0) { return fail(RSRCU_ERR_INVALID, "tile blocks must be positive"); }
	CU(cudaSetDevice(c->device));
	// take the other staging arena; wait only until the frame that last used it has been uploaded
	c->outSlot = (c->outSlot + 1) % kSlots;
	CU(cudaEventSynchronize(c->arenaFree[c->outSlot]));
	c->width = width; c->height = height;
	{
		// CalcGuardBandFactor (rglv_view_frustum.hxx:36-39)
		const int half = std::max(width, height) / 2;
		c->guardFactor = (2048.0f - static_cast<float>(half)) / static_cast<float>(half); }
	const int rw = tileWBlocks * 8, rh = tileHBlocks * 8;
	// the device tile must lie inside one reference tile to reproduce its start point exactly;
	// otherwise use the device tile itself (identical unless an int32 edge product overflows)
	c->refTileW = (rw % kTile == 0) ? rw : kTile;
	c->refTileH = (rh % kTile == 0) ? rh : kTile;
	c->postTileW = rw; c->postTileH = rh;
	c->states.clear(); c->draws.clear(); c->cmds.clear(); c->cmdDstKind.clear(); c->copies.clear();
	c->slotPlan[c->outSlot].valid = false;
	c->launched[c->outSlot] = Launched{};
	c->storesUsed = 0; c->tcDev = nullptr;
	c->arenas[c->outSlot].used = 0;
	++c->frameStamp;
	c->h2dPending = false; c->frameUsedPinned = false;
	if ((c->frameStamp & 127u) == 0 && !c->pinned.empty()) {
		// arrays not bound for a while: let their pages go (the host may have freed them long ago)
		for (auto it = c->pinned.begin(); it != c->pinned.end();) {
			if (c->frameStamp - it->second.lastFrame > 256) {
				CU(cudaDeviceSynchronize());
				if (it->second.owned) { cudaHostUnregister(reinterpret_cast<void*>(it->second.regBase)); cudaGetLastError(); }
				for (auto& b : it->second.dev) { b.release(); }
				it = c->pinned.erase(it); }
			else { ++it; } } }
	c->trianglesSubmitted = 0; c->inputBytes = 0; c->progMask = 0;
	c->haveState = false; c->stateDirty = true;
	for (auto& b : c->curBuffers) { b = DevRef{}; }
	for (auto& f : c->curBufferFloats) { f = 0; }
	c->curTus[0] = HostTex{}; c->curTus[1] = HostTex{};
	c->curTu3 = DevRef{}; c->curTu3dim = 256;
	c->inFrame = true;
	c->tBegin = std::chrono::steady_clock::now();
	return RSRCU_OK; }

int rsrcu_set_state(rsrcu_ctx* c, const RsrState* st) {
	if (!c || !st) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	if (!c->inFrame) { return fail(RSRCU_ERR_INVALID, "rsrcu_set_state outside begin/end frame"); }
	c->curState = *st;
	c->curStatePtr = &c->curState;
	c->haveState = true;
	c->stateDirty = true;
	return RSRCU_OK; }

int rsrcu_bind_buffer(rsrcu_ctx* c, int slot, const float* host, size_t nFloats, int upload) {
	if (!c || slot < 0 || slot >= 16) { return fail(RSRCU_ERR_INVALID, "bad buffer slot %d", slot); }
	if (!c->inFrame) { return fail(RSRCU_ERR_INVALID, "rsrcu_bind_buffer outside begin/end frame"); }
	int r = uploadData(c, host, nFloats * sizeof(float), upload, c->curBuffers[slot], slot < 3);
	if (r != RSRCU_OK) { return r; }
	c->curBufferFloats[slot] = host ? nFloats : 0;
	c->stateDirty = true;
	return RSRCU_OK; }

int rsrcu_bind_texture(rsrcu_ctx* c, int unit, const float* host, int width, int height, int stride, int filter,
                       int rowsInMemory, int upload) {
	if (!c || unit < 0 || unit > 1) { return fail(RSRCU_ERR_INVALID, "bad texture unit %d", unit); }
	if (!c->inFrame) { return fail(RSRCU_ERR_INVALID, "rsrcu_bind_texture outside begin/end frame"); }
	HostTex& t = c->curTus[unit];
	const size_t texels = static_cast<size_t>(stride) * static_cast<size_t>(rowsInMemory);
	int r = uploadData(c, host, texels * 16, upload, t.ref);
	if (r != RSRCU_OK) { return r; }
	t.texelCount = static_cast<uint32_t>(texels);
	t.width = width; t.height = height; t.stride = stride; t.filter = filter;
	c->stateDirty = true;
	return RSRCU_OK; }

int rsrcu_bind_depth_texture(rsrcu_ctx* c, const float* host, int dim, int upload) {
	if (!c) { return fail(RSRCU_ERR_INVALID, "null context"); }
	if (!c->inFrame) { return fail(RSRCU_ERR_INVALID, "rsrcu_bind_depth_texture outside begin/end frame"); }
	int r = uploadData(c, host, static_cast<size_t>(dim) * dim * sizeof(float), upload, c->curTu3);
	if (r != RSRCU_OK) { return r; }
	c->curTu3dim = dim;
	c->stateDirty = true;
	return RSRCU_OK; }

// the device target of the frame's next store command: its own buffer from the slot's pool
static cudaError_t takeStore(rsrcu_ctx* c, size_t bytes, void*& out) {
	auto& pool = c->storePool[c->outSlot];
	if (static_cast<size_t>(c->storesUsed) >= pool.size()) { pool.emplace_back(); }
	DevBuf& b = pool[c->storesUsed];
	const cudaError_t e = b.reserve(bytes);
	if (e == cudaSuccess) { ++c->storesUsed; out = b.ptr; }
	return e; }

static int pushCmd(rsrcu_ctx* c, int type, int arg, void* dst, int stride, int kind) {
	int r = snapshotState(c);
	if (r != RSRCU_OK) { return r; }
	FrameCmd cmd{};
	cmd.type = type; cmd.state = static_cast<int>(c->states.size()) - 1; cmd.arg = arg; cmd.dst = dst; cmd.dstStride = stride;
	cmd.beforeDraw = static_cast<int>(c->draws.size());
	c->cmds.push_back(cmd);
	c->cmdDstKind.push_back(kind);
	return RSRCU_OK; }

int rsrcu_clear(rsrcu_ctx* c, int bits) {
	if (!c || !c->inFrame) { return fail(RSRCU_ERR_INVALID, "rsrcu_clear outside begin/end frame"); }
	if (bits & RSRCU_GL_STENCIL_BUFFER_BIT) { return fail(RSRCU_ERR_UNSUPPORTED, "CLEAR on STENCIL not implemented (rglv_gpu.cxx:313-315)"); }
	if (c->haveState && c->curStatePtr->color0_attachment_type == RSRCU_RB_COLOR_DEPTH &&
	    bits != (RSRCU_GL_COLOR_BUFFER_BIT | RSRCU_GL_DEPTH_BUFFER_BIT)) {
		return fail(RSRCU_ERR_UNSUPPORTED, "must clear color and depth when using RB_COLOR_DEPTH (rglv_gpu.cxx:317-319)"); }
	return pushCmd(c, kCmdClear, bits, nullptr, 0, 0); }

static int recordDraw(rsrcu_ctx* c, int count, const uint16_t* indices, int instanceCount, int upload, bool arrays) {
	if (!c || !c->inFrame) { return fail(RSRCU_ERR_INVALID, "draw outside begin/end frame"); }
	if (count < 0) { return fail(RSRCU_ERR_INVALID, "negative count"); }
	int r = snapshotState(c);
	if (r != RSRCU_OK) { return r; }
	const HostState& hs = c->states.back();
	const int key = hs.key, programId = hs.programId;
	if (!drawProgramInstalled(programId, key)) {
		return fail(RSRCU_ERR_NO_PROGRAM, "no dispatch entry for program %d state key 0x%x (src/viewer/shaders.cxx:54-126)", programId, key); }
	if (programId == 4 || programId == 65 || programId == 26 || programId == 41 || programId == 10) {
		// texture units with pow2 dims outside 4..1024 cannot be made by the reference (exit(1))
		if (hs.tex0Null) { return fail(RSRCU_ERR_INVALID, "program %d samples texture unit 0 but none is bound", programId); }
		int power = 0; while ((1 << (power + 1)) <= hs.tex0Width) { ++power; }
		const bool isPow2 = ((1 << power) == hs.tex0Width) && hs.tex0Width == hs.tex0Height && hs.tex0Stride == hs.tex0Width;
		if (isPow2 && (hs.tex0Width < 4 || hs.tex0Width > 1024)) {
			return fail(RSRCU_ERR_UNSUPPORTED, "can't make TextureUnit for pow2 size %d (rglr_texture_sampler.cxx:333-357)", hs.tex0Width); } }
	const bool instanced = instanceCount > 0;
	const int instances = instanced ? instanceCount : 1;
	if (instances > 65536) { return fail(RSRCU_ERR_INVALID, "instance id must fit uint16 (rglv_gpu_impl.hxx:490)"); }
	const int prims = count / 3;
	if (prims == 0) { return RSRCU_OK; }
	// whole-draw frustum rejection (sub-frames of a large target see most of the scene off screen): every program except
	// Many transforms the position attribute by vpm alone
	if (!instanced && programId != 6 && hs.posRanged && boxOutsideFrustum(hs.ds.vpm, hs.posLo, hs.posHi, c->guardFactor)) {
		c->trianglesSubmitted += static_cast<uint64_t>(prims);
		++c->drawsCulled;
		return RSRCU_OK; }

	HostDraw hd{};
	DevDraw& d = hd.d;
	d.state = static_cast<int>(c->states.size()) - 1;
	d.prims = prims;
	d.instances = instances;
	d.instanced = instanced ? 1 : 0;
	// vertex extent: explicit buffer length (slot 0) or, for arrays, the count
	size_t nverts = 0;
	if (!arrays && !indices) { return fail(RSRCU_ERR_INVALID, "null index pointer"); }
	if (arrays) { nverts = static_cast<size_t>(prims) * 3; }
	else {
		nverts = hs.bufferFloats[0];
		if (hs.posNull) {
			// no position buffer: every vertex is the origin; need max index + 1
			uint16_t mx = 0;
			for (int i = 0; i < prims * 3; ++i) { mx = std::max(mx, indices[i]); }
			nverts = static_cast<size_t>(mx) + 1; } }
	if (!arrays && !hs.posNull && nverts == 0) { return fail(RSRCU_ERR_INVALID, "position buffer has no length"); }
	if (arrays && !hs.posNull && hs.bufferFloats[0] < nverts) {
		return fail(RSRCU_ERR_INVALID, "DrawArrays count %d exceeds bound position buffer (%zu floats)", count, hs.bufferFloats[0]); }
	for (int slot : {1, 2, 3, 4, 5, 6, 7, 8, 9, 10}) {
		if (hs.ds.buffers[slot] != nullptr && hs.bufferFloats[slot] < nverts) {
			return fail(RSRCU_ERR_INVALID, "buffer slot %d shorter (%zu) than the position buffer (%zu)", slot, hs.bufferFloats[slot], nverts); } }
	if (instanced && (programId == 6) && (hs.ds.buffers[15] == nullptr || hs.bufferFloats[15] < static_cast<size_t>(instances) * 16)) {
		return fail(RSRCU_ERR_INVALID, "instanced draw needs %d mat4 in slot 15", instances); }
	d.nverts = static_cast<int>(nverts);
	d.nvary = programVaryings(programId);
	d.strideF4 = 2 + (d.nvary + 3) / 4;
	d.N = static_cast<uint32_t>(prims) * static_cast<uint32_t>(instances);
	d.batchKey = static_cast<uint32_t>(programId & 0xff) | (static_cast<uint32_t>(key) << 8);
	d.cullBits = (hs.ds.cullingEnabled ? 1u : 0u) | ((static_cast<uint32_t>(hs.ds.cullFace) & 3u) << 1) | ((key & 1) ? 8u : 0u);
	if (!arrays) {
		r = uploadData(c, indices, static_cast<size_t>(prims) * 3 * sizeof(uint16_t), upload, hd.indices);
		if (r != RSRCU_OK) { return r; } }
	c->trianglesSubmitted += d.N;
	{
		int attrs = 0;
		for (int slot = 0; slot <= 10; ++slot) { attrs += hs.ds.buffers[slot] != nullptr ? 1 : 0; }
		c->inputBytes += static_cast<uint64_t>(nverts) * 4u * attrs + (arrays ? 0u : static_cast<uint64_t>(prims) * 6u) + (instanced ? static_cast<uint64_t>(instances) * 64u : 0u); }
	c->progMask |= prog_bit(programId);
	c->draws.push_back(hd);
	return RSRCU_OK; }

int rsrcu_draw_elements(rsrcu_ctx* c, int count, const uint16_t* indices, int hint, int instanceCount, int upload) {
	(void)hint;   // DENSE|READ4 only selects between two equivalent CPU bin loops (rglv_gpu_impl.hxx:274-294)
	return recordDraw(c, count, indices, instanceCount, upload, false); }

int rsrcu_draw_arrays(rsrcu_ctx* c, int count, int instanceCount) {
	return recordDraw(c, count, nullptr, instanceCount, RSRCU_UPLOAD_ALWAYS, true); }

int rsrcu_store_color_tc(rsrcu_ctx* c, int gamma, uint32_t* dst, int width, int height, int stridePx) {
	if (!c || !c->inFrame) { return fail(RSRCU_ERR_INVALID, "store outside begin/end frame"); }
	if (width != c->width || height != c->height) { return fail(RSRCU_ERR_INVALID, "store canvas %dx%d != target %dx%d", width, height, c->width, c->height); }
	if (c->haveState && !bltProgramInstalled(c->curStatePtr->program_id)) {
		return fail(RSRCU_ERR_NO_PROGRAM, "no blt dispatch entry for program %d (src/viewer/shaders.cxx:57-67)", c->curStatePtr->program_id); }
	CU(cudaSetDevice(c->device));
	void* dev = nullptr;
	CU(takeStore(c, static_cast<size_t>(width) * height * 4, dev));
	c->tcStride = width; c->tcDev = dev;
	int r = pushCmd(c, kCmdStoreTC, gamma ? 1 : 0, dev, width, 1);
	if (r != RSRCU_OK) { return r; }
	if (dst) {
		c->copies.push_back(PendingCopy{dst, dev, static_cast<size_t>(width) * 4, static_cast<size_t>(height),
		                                static_cast<size_t>(stridePx) * 4, static_cast<size_t>(width) * 4}); }
	return RSRCU_OK; }

int rsrcu_store_color_tc_device(rsrcu_ctx* c, int gamma, void* deviceDst, int width, int height, int stridePx) {
	if (!c || !c->inFrame) { return fail(RSRCU_ERR_INVALID, "store outside begin/end frame"); }
	if (!deviceDst) { return fail(RSRCU_ERR_INVALID, "null device destination"); }
	if (width != c->width || height != c->height) { return fail(RSRCU_ERR_INVALID, "store canvas %dx%d != target %dx%d", width, height, c->width, c->height); }
	if (c->haveState && !bltProgramInstalled(c->curStatePtr->program_id)) {
		return fail(RSRCU_ERR_NO_PROGRAM, "no blt dispatch entry for program %d", c->curStatePtr->program_id); }
	c->tcStride = stridePx; c->tcDev = deviceDst;
	return pushCmd(c, kCmdStoreTC, gamma ? 1 : 0, deviceDst, stridePx, 1); }

int rsrcu_enable_peer_access(rsrcu_ctx* c, int peerDevice) {
	if (!c) { return fail(RSRCU_ERR_INVALID, "null context"); }
	if (peerDevice == c->device) { return RSRCU_OK; }
	CU(cudaSetDevice(c->device));
	int can = 0;
	CU(cudaDeviceCanAccessPeer(&can, c->device, peerDevice));
	if (!can) { return fail(RSRCU_ERR_UNSUPPORTED, "device %d cannot access device %d's memory", c->device, peerDevice); }
	const cudaError_t e = cudaDeviceEnablePeerAccess(peerDevice, 0);
	if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return RSRCU_OK; }
	CU(e);
	return RSRCU_OK; }

int rsrcu_store_color_fp(rsrcu_ctx* c, float* dst, int width, int height, int stridePx, int half) {
	if (!c || !c->inFrame) { return fail(RSRCU_ERR_INVALID, "store outside begin/end frame"); }
	// the half-size canvas holds one pixel per 2x2 quad of the target (GL::StoreColor(dst, downsample), rglv_gl.cxx)
	const int wantW = half ? c->width / 2 : c->width, wantH = half ? c->height / 2 : c->height;
	if (width != wantW || height != wantH) { return fail(RSRCU_ERR_INVALID, "store canvas %dx%d != %s target %dx%d", width, height, half ? "half" : "full", wantW, wantH); }
	CU(cudaSetDevice(c->device));
	void* dev = nullptr;
	CU(takeStore(c, static_cast<size_t>(width) * height * 16, dev));
	int r = pushCmd(c, half ? kCmdStoreHalfFP : kCmdStoreFP, 0, dev, width, 2);
	if (r != RSRCU_OK) { return r; }
	if (dst) {
		c->copies.push_back(PendingCopy{dst, dev, static_cast<size_t>(width) * 16, static_cast<size_t>(height),
		                                static_cast<size_t>(stridePx) * 16, static_cast<size_t>(width) * 16}); }
	return RSRCU_OK; }

int rsrcu_store_color_quads(rsrcu_ctx* c, float* dst, int width, int height, int strideQuads) {
	if (!c || !c->inFrame) { return fail(RSRCU_ERR_INVALID, "store outside begin/end frame"); }
	if (width != c->width || height != c->height) { return fail(RSRCU_ERR_INVALID, "store canvas %dx%d != target %dx%d", width, height, c->width, c->height); }
	if (strideQuads < width / 2) { return fail(RSRCU_ERR_INVALID, "quad canvas stride %d < %d quads per row", strideQuads, width / 2); }
	CU(cudaSetDevice(c->device));
	const size_t rowBytes = static_cast<size_t>(width / 2) * 64, rows = static_cast<size_t>(height / 2);
	void* dev = nullptr;
	CU(takeStore(c, rowBytes * rows, dev));
	int r = pushCmd(c, kCmdStoreQuadsFP, 0, dev, width / 2, 2);
	if (r != RSRCU_OK) { return r; }
	if (dst) {
		c->copies.push_back(PendingCopy{dst, dev, rowBytes, rows, static_cast<size_t>(strideQuads) * 64, rowBytes}); }
	return RSRCU_OK; }

int rsrcu_store_depth(rsrcu_ctx* c, float* dst) {
	if (!c || !c->inFrame) { return fail(RSRCU_ERR_INVALID, "store outside begin/end frame"); }
	CU(cudaSetDevice(c->device));
	void* dev = nullptr;
	CU(takeStore(c, static_cast<size_t>(c->width) * c->height * 4, dev));
	int r = pushCmd(c, kCmdStoreDepth, 0, dev, c->width, 3);
	if (r != RSRCU_OK) { return r; }
	if (dst) {
		c->copies.push_back(PendingCopy{dst, dev, static_cast<size_t>(c->width) * 4, static_cast<size_t>(c->height),
		                                static_cast<size_t>(c->width) * 4, static_cast<size_t>(c->width) * 4}); }
	return RSRCU_OK; }

int rsrcu_end_frame(rsrcu_ctx* c) {
	if (!c || !c->inFrame) { return fail(RSRCU_ERR_INVALID, "rsrcu_end_frame without begin"); }
	CU(cudaSetDevice(c->device));
	c->inFrame = false;
	const auto tSubmit = std::chrono::steady_clock::now();
	c->hpLast = tSubmit;
	c->recordNs = static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(tSubmit - c->tBegin).count());

	const int W = c->width, H = c->height;
	FrameParams fp{};
	fp.width = W; fp.height = H;
	fp.tilesX = (W + kTile - 1) / kTile; fp.tilesY = (H + kTile - 1) / kTile;
	fp.refTileW = c->refTileW; fp.refTileH = c->refTileH;
	fp.postTileW = c->postTileW; fp.postTileH = c->postTileH;
	fp.guardFactor = c->guardFactor;
	fp.ndraws = static_cast<int>(c->draws.size());
	fp.ncmds = static_cast<int>(c->cmds.size());
	const int ntiles = fp.tilesX * fp.tilesY;

	// ---- layout: jobs, order keys, vertex records ---------------------------------------------
	uint64_t vjobs = 0, pjobs = 0, ids = 0, ptvbF4 = 0, nvertsTotal = 0;
	for (size_t di = 0; di < c->draws.size(); ++di) {
		DevDraw& d = c->draws[di].d;
		d.vjobBase = static_cast<uint32_t>(vjobs);
		d.pjobBase = static_cast<uint32_t>(pjobs);
		d.idBase = static_cast<uint32_t>(ids);
		d.vbaseF4 = static_cast<uint32_t>(ptvbF4);
		d.flagBase = static_cast<uint32_t>(nvertsTotal);
		const uint64_t nv = static_cast<uint64_t>(d.nverts) * d.instances;
		vjobs += nv; nvertsTotal += nv; ptvbF4 += nv * d.strideF4;
		pjobs += d.N;
		ids += static_cast<uint64_t>(d.N) * (1 + kMaxFan); }
	if (c->states.size() > 65535) { return fail(RSRCU_ERR_UNSUPPORTED, "more than 65535 state snapshots in one frame"); }
	if (pjobs >= 0x7ffffff0ull || ids >= 0xfffffff0ull || c->clipCapacity >= (1u << 24) || ptvbF4 >= 0x7ffffff0ull || vjobs >= 0xfffffff0ull) {
		return fail(RSRCU_ERR_UNSUPPORTED, "frame too large for 32-bit ids (%llu ids, %llu vertex records)",
		            static_cast<unsigned long long>(ids), static_cast<unsigned long long>(ptvbF4)); }
	fp.totalVJobs = static_cast<uint32_t>(vjobs);
	fp.totalPJobs = static_cast<uint32_t>(pjobs);
	fp.totalKeys = static_cast<uint32_t>(std::max<uint64_t>(ids, 1));
	// list cells per tile: one for small frames (a tile's whole list is sorted at once); frames with millions of
	// triangles split every list by triangle index range so that a cell stays within one raster batch
	fp.groupShift = 31;
	if (c->forceGroupShift) { fp.groupShift = c->forceGroupShift; while ((pjobs >> fp.groupShift) >= static_cast<uint64_t>(kMaxGroups)) { ++fp.groupShift; } }
	else if (pjobs > (1u << 18)) { fp.groupShift = 15; while ((pjobs >> fp.groupShift) >= static_cast<uint64_t>(kMaxGroups)) { ++fp.groupShift; } }
	fp.groups = static_cast<int>(pjobs >> fp.groupShift) + 1;
	fp.largeCapacity = c->largeCapacity;
	fp.largeTiles = c->largeTiles;
	fp.clipCapacity = c->clipCapacity;
	fp.listCapacity = c->listCapacity;

	if (c->hostProf) { const auto now_ = std::chrono::steady_clock::now(); c->hp[0] += static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(now_ - c->hpLast).count()); c->hpLast = now_; }
	// ---- frame tables into the arena (state / draw tables need final device addresses) -----
	size_t offStates = 0, offDraws = 0, offCmds = 0, offVBlocks = 0, offPBlocks = 0;
	const size_t nVBlocks = static_cast<size_t>((vjobs + 255) / 256) + 1, nPBlocks = static_cast<size_t>((pjobs + 255) / 256) + 1;
	CU(c->arenas[c->outSlot].push(nullptr, sizeof(DevState) * std::max<size_t>(1, c->states.size()), offStates));
	CU(c->arenas[c->outSlot].push(nullptr, sizeof(DevDraw) * std::max<size_t>(1, c->draws.size()), offDraws));
	CU(c->arenas[c->outSlot].push(nullptr, sizeof(FrameCmd) * std::max<size_t>(1, c->cmds.size()), offCmds));
	CU(c->arenas[c->outSlot].push(nullptr, sizeof(uint32_t) * nVBlocks, offVBlocks));
	CU(c->arenas[c->outSlot].push(nullptr, sizeof(uint32_t) * nPBlocks, offPBlocks));
	CU(c->arenas[c->outSlot].dev.reserve(c->arenas[c->outSlot].used));

	{
		const uint8_t* arenaDev = static_cast<const uint8_t*>(c->arenas[c->outSlot].dev.ptr);
		DevState* out = reinterpret_cast<DevState*>(c->arenas[c->outSlot].host + offStates);
		for (size_t i = 0; i < c->states.size(); ++i) {
			const HostState& hs = c->states[i];
			std::memcpy(out + i, &hs.ds, sizeof(DevState));
			if (hs.hasArenaRef) {
				for (int b = 0; b < 16; ++b) { patchRef(out[i].buffers[b], arenaDev); }
				patchRef(out[i].tu[0].texels, arenaDev); patchRef(out[i].tu[1].texels, arenaDev);
				patchRef(out[i].tu3, arenaDev); } } }
	for (size_t i = 0; i < c->draws.size(); ++i) {
		DevDraw d = c->draws[i].d;
		d.indices = static_cast<const uint16_t*>(resolve(c, c->draws[i].indices));
		std::memcpy(c->arenas[c->outSlot].host + offDraws + i * sizeof(DevDraw), &d, sizeof(d)); }
	{
		// draw of the first job of every block of 256 vertex / triangle jobs (find_draw, kernels.cuh)
		uint32_t* vb = reinterpret_cast<uint32_t*>(c->arenas[c->outSlot].host + offVBlocks);
		uint32_t* pb = reinterpret_cast<uint32_t*>(c->arenas[c->outSlot].host + offPBlocks);
		size_t dv = 0, dp = 0;
		const size_t nd = c->draws.size();
		for (size_t b = 0; b < nVBlocks; ++b) {
			while (dv + 1 < nd && c->draws[dv + 1].d.vjobBase <= b * 256) { ++dv; }
			vb[b] = static_cast<uint32_t>(dv); }
		for (size_t b = 0; b < nPBlocks; ++b) {
			while (dp + 1 < nd && c->draws[dp + 1].d.pjobBase <= b * 256) { ++dp; }
			pb[b] = static_cast<uint32_t>(dp); } }
	if (!c->cmds.empty()) { std::memcpy(c->arenas[c->outSlot].host + offCmds, c->cmds.data(), sizeof(FrameCmd) * c->cmds.size()); }

	if (c->hostProf) { const auto now_ = std::chrono::steady_clock::now(); c->hp[1] += static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(now_ - c->hpLast).count()); c->hpLast = now_; }
	FramePlan& plan = c->slotPlan[c->outSlot];
	plan.fp = fp;
	plan.offStates = offStates; plan.offDraws = offDraws; plan.offCmds = offCmds; plan.offVBlocks = offVBlocks; plan.offPBlocks = offPBlocks;
	plan.ptvbF4 = ptvbF4; plan.nvertsTotal = nvertsTotal; plan.pjobs = pjobs;
	plan.arenaBytes = c->arenas[c->outSlot].used;
	plan.ncmdInline = 0;
	for (int i = 0; i < kInlineCmds && i < fp.ncmds; ++i) {
		plan.icmd[i] = c->cmds[i];
		const DevState& ds = c->states[c->cmds[i].state].ds;
		CmdState& cs = plan.icmdState[i];
		std::memcpy(cs.clearColor, ds.clearColor, sizeof(cs.clearColor));
		cs.clearDepth = ds.clearDepth; cs.programId = ds.programId; cs.uniform0 = ds.uniforms[0]; cs.color0Type = ds.color0Type;
		++plan.ncmdInline; }
	plan.copies = c->copies;
	plan.tcDev = c->tcDev; plan.tcStride = c->tcStride; plan.storesUsed = c->storesUsed;
	plan.trianglesSubmitted = c->trianglesSubmitted; plan.inputBytes = c->inputBytes; plan.progMask = c->progMask;
	plan.valid = true;
	CU(c->arenas[c->outSlot].dev.reserve(c->arenas[c->outSlot].used));
	if (c->hostProf) { const auto now_ = std::chrono::steady_clock::now(); c->hp[2] += static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(now_ - c->hpLast).count()); c->hpLast = now_; }
	c->lastArena = c->outSlot;
	const int r = launchFrame(c, plan, static_cast<const uint8_t*>(c->arenas[c->outSlot].dev.ptr), c->arenas[c->outSlot].host, c->arenas[c->outSlot].used, c->outSlot);
	c->submitNs = static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - tSubmit).count());
	return r; }

int rsrcu_retain_frame(rsrcu_ctx* c, rsrcu_frame** out) {
	if (!c || !out) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	if (c->inFrame || c->lastArena < 0 || !c->slotPlan[c->lastArena].valid) { return fail(RSRCU_ERR_INVALID, "rsrcu_retain_frame: no submitted frame (call it after rsrcu_end_frame, before the next rsrcu_begin_frame)"); }
	if (c->slotPlan[c->lastArena].fp.ncmds > kInlineCmds) { return fail(RSRCU_ERR_UNSUPPORTED, "a retained frame holds at most %d clear / store commands", kInlineCmds); }
	if (c->frameUsedPinned) { return fail(RSRCU_ERR_UNSUPPORTED, "a frame with RSRCU_UPLOAD_FRAME arrays of 1 MiB or more cannot be retained (they live in buffers later frames overwrite): bind them RSRCU_UPLOAD_STATIC"); }
	CU(cudaSetDevice(c->device));
	const UploadArena& ar = c->arenas[c->lastArena];
	const FramePlan& plan = c->slotPlan[c->lastArena];
	auto* f = new rsrcu_frame();
	f->plan = plan;
	f->device = c->device;
	// the frame's store targets leave the context's pool with it: later frames (of any size) get fresh buffers
	{
		auto& pool = c->storePool[c->lastArena];
		for (int i = 0; i < plan.storesUsed && i < static_cast<int>(pool.size()); ++i) { f->stores.push_back(pool[i]); pool[i] = DevBuf{}; } }
	const size_t bytes = std::max<size_t>(plan.arenaBytes, 16);
	if (cudaMalloc(&f->devArena, bytes) != cudaSuccess) { delete f; return fail(RSRCU_ERR_CUDA, "cudaMalloc of %zu bytes for a retained frame failed", bytes); }
	// the tables hold device addresses; those that point into the context's arena mirror move to the private copy
	std::vector<uint8_t> tmp(ar.host, ar.host + plan.arenaBytes);
	const uintptr_t oldBase = reinterpret_cast<uintptr_t>(ar.dev.ptr), newBase = reinterpret_cast<uintptr_t>(f->devArena);
	auto rebase = [&](auto& ptr) {
		const uintptr_t v = reinterpret_cast<uintptr_t>(ptr);
		if (v >= oldBase && v < oldBase + plan.arenaBytes) { ptr = reinterpret_cast<std::remove_reference_t<decltype(ptr)>>(v - oldBase + newBase); } };
	DevState* st = reinterpret_cast<DevState*>(tmp.data() + plan.offStates);
	for (size_t i = 0; i < c->states.size(); ++i) {
		for (int b = 0; b < 16; ++b) { rebase(st[i].buffers[b]); }
		rebase(st[i].tu[0].texels); rebase(st[i].tu[1].texels); rebase(st[i].tu3); }
	DevDraw* dr = reinterpret_cast<DevDraw*>(tmp.data() + plan.offDraws);
	for (size_t i = 0; i < c->draws.size(); ++i) { rebase(dr[i].indices); }
	if (cudaMemcpy(f->devArena, tmp.data(), plan.arenaBytes, cudaMemcpyHostToDevice) != cudaSuccess) {
		cudaFree(f->devArena); delete f; return fail(RSRCU_ERR_CUDA, "upload of a retained frame failed"); }
	// the slot's launch record now refers to a plan whose store targets belong to the retained frame: settle it here
	c->slotPlan[c->lastArena].valid = false;
	*out = f;
	return RSRCU_OK; }

int rsrcu_replay_frame(rsrcu_ctx* c, rsrcu_frame* f) {
	if (!c || !f) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	if (c->inFrame) { return fail(RSRCU_ERR_INVALID, "rsrcu_replay_frame inside begin/end frame"); }
	if (f->device != c->device) { return fail(RSRCU_ERR_INVALID, "frame was retained on device %d", f->device); }
	CU(cudaSetDevice(c->device));
	const auto t0 = std::chrono::steady_clock::now();
	if (!f->plan.copies.empty()) {
		// the store targets are the recorded ones: the previous frame's read-back (possibly of the same buffer) goes first
		CU(cudaStreamWaitEvent(c->overlap ? c->frontStream : c->stream, c->evCopied[c->outSlot], 0)); }
	c->outSlot = (c->outSlot + 1) % kSlots;   // counters ring; the store targets stay the ones recorded
	c->recordNs = 0;
	const int r = launchFrame(c, f->plan, static_cast<const uint8_t*>(f->devArena), nullptr, 0, -1);
	c->submitNs = static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count());
	return r; }

int rsrcu_release_frame(rsrcu_ctx* c, rsrcu_frame* f) {
	if (!f) { return RSRCU_OK; }
	if (c) {
		cudaSetDevice(c->device);
		cudaStreamSynchronize(c->frontStream);
		cudaStreamSynchronize(c->stream);
		cudaStreamSynchronize(c->tileStream2); }
	if (c) { for (auto& L : c->launched) { if (L.plan == &f->plan) { L = Launched{}; } } }
	if (f->devArena) { cudaFree(f->devArena); }
	for (auto& b : f->stores) { b.release(); }
	delete f;
	return RSRCU_OK; }

// Reads the overflow bits a finished frame left in its counters and grows the buffer concerned.  Returns the
// message for the caller (nullptr: the frame is complete).
static const char* growAfterOverflow(rsrcu_ctx* c, const Counters& k, char* buf, size_t n) {
	if (k.overflow & 2u) {
		const uint64_t need = k.entries + k.entries / 4;
		c->listCapacity = static_cast<uint32_t>(std::min<uint64_t>(std::max<uint64_t>(need, c->listCapacity), 0xfffffff0ull));
		std::snprintf(buf, n, "tile list capacity exceeded (%llu entries)", static_cast<unsigned long long>(k.entries));
		return buf; }
	if (k.overflow & 8u) {
		c->largeTiles = std::min(c->largeTiles * 4, 1 << 12);
		std::snprintf(buf, n, "a tile is covered by more than %d queued large triangles (threshold now %d tiles)", kTileLargeCap, c->largeTiles);
		return buf; }
	if (k.overflow & 4u) {
		c->largeCapacity = std::max(c->largeCapacity * 2, k.nLarge + k.nLarge / 4);
		std::snprintf(buf, n, "large-item queue exceeded (%u items)", k.nLarge);
		return buf; }
	if (k.overflow & 1u) {
		c->clipCapacity = std::max(c->clipCapacity * 2, k.clipAlloc + k.clipAlloc / 4);
		std::snprintf(buf, n, "clip record capacity exceeded (%u needed)", k.clipAlloc);
		return buf; }
	return nullptr; }

// Waits for the read-back of the frame launched into `slot`.  A frame whose tile lists, clip records or large-item
// queue overflowed was rendered truncated: the buffer is grown and the frame is launched again from its tables,
// which are still on the device (arena ring / retained frame), until it fits -- the caller never sees a truncated
// frame.  `force`: launch it again even if it fitted (a frame submitted before it was launched again, and the two
// may share a store destination).
static int settleSlot(rsrcu_ctx* c, int slot, bool force, bool& relaunched) {
	Launched& L = c->launched[slot];
	relaunched = false;
	if (!L.plan) { return RSRCU_OK; }
	for (int attempt = 0; ; ++attempt) {
		CU(cudaEventSynchronize(c->evCopied[slot]));
		char msg[160];
		const char* why = growAfterOverflow(c, c->hostCounters[slot], msg, sizeof(msg));
		if (!why && !force) { L.checked = true; return RSRCU_OK; }
		if (why && (attempt >= 8 || c->clipCapacity >= (1u << 24))) {
			L.checked = true;
			return fail(RSRCU_ERR_OVERFLOW, "%s: still not enough after %d attempts", why, attempt); }
		force = false;
		relaunched = true;
		++c->framesRetried;
		const FramePlan* plan = L.plan;
		const uint8_t* arenaDev = L.arenaDev;
		const int keep = c->outSlot;
		void* const keepTc = c->shownTcDev; const int keepStride = c->shownTcStride;
		c->outSlot = slot;
		const int r = launchFrame(c, *plan, arenaDev, nullptr, 0, -1);
		c->outSlot = keep;
		if (slot != keep) { c->shownTcDev = keepTc; c->shownTcStride = keepStride; }
		if (r != RSRCU_OK) { return r; } } }

int rsrcu_sync(rsrcu_ctx* c) {
	if (!c) { return fail(RSRCU_ERR_INVALID, "null context"); }
	CU(cudaSetDevice(c->device));
	{ const int r = flushDeferredCopies(c); if (r != RSRCU_OK) { return r; } }
	// every frame still in flight, oldest first
	bool again = false;
	for (int k = kSlots - 1; k >= 0; --k) {
		const int slot = (c->outSlot + kSlots - k) % kSlots;
		if (c->launched[slot].plan && !c->launched[slot].checked) {
			bool relaunched = false;
			const int r = settleSlot(c, slot, again, relaunched);
			if (r != RSRCU_OK) { return r; }
			again = again || relaunched; } }
	CU(cudaStreamSynchronize(c->frontStream));
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaStreamSynchronize(c->tileStream2));
	CU(cudaStreamSynchronize(c->copyStream));
	if (!c->framePending) { return RSRCU_OK; }
	c->framePending = false;
	const Counters& k = c->hostCounters[c->outSlot];
	if (c->launched[c->outSlot].plan) { c->stats.triangles_submitted = c->launched[c->outSlot].plan->trianglesSubmitted; c->stats.input_bytes = c->launched[c->outSlot].plan->inputBytes; }
	c->stats.triangles_binned = k.binned;
	c->stats.triangles_clipped = k.clipped;
	c->stats.bin_entries = k.entries;
	c->stats.fragments_shaded = k.fragments;
	c->stats.kernel_launches = c->launches;
	c->stats.h2d_bytes = c->lastH2D;
	c->stats.d2h_bytes = c->lastD2H;
	c->stats.list_chunks_run_merge = k.chunksRunMerge;
	c->stats.list_chunks_key_range = k.chunksKeyRange;
	c->stats.host_record_ns = c->recordNs;
	c->stats.host_submit_ns = c->submitNs;
	c->stats.frames_retried = c->framesRetried;
	c->stats.draws_culled = c->drawsCulled;
	if (c->profiling) {
		for (int i = 0; i < 6; ++i) { c->stageMs[i] = 0.0f; }
		if (c->profiling > 1) { for (int i = 0; i < 5; ++i) { cudaEventElapsedTime(&c->stageMs[i], c->evStage[i + 1], c->evStage[i + 2]); } }
		cudaEventElapsedTime(&c->stageMs[5], c->evStage[6], c->evStage[7]);
		// stage order of the header: vertex, setup, count, scan, fill, tile
		cudaEventElapsedTime(&c->stageMs[6], c->evStage[0], c->evStage[7]);
		c->stageSpansValid = c->profiling > 1;
		if (c->stageSpansValid) {
			for (int i = 0; i < 8; ++i) {
				if (cudaEventElapsedTime(&c->stageStartMs[i], c->evStage[0], c->evStage[i]) != cudaSuccess) { cudaGetLastError(); c->stageSpansValid = false; } } } }
	return RSRCU_OK; }

int rsrcu_run_stream(rsrcu_ctx* c, const void* stream, size_t bytes) {
	if (!c || !stream) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	const uint8_t* p = static_cast<const uint8_t*>(stream);
	const uint8_t* end = p + bytes;
	while (p + 8 <= end) {
		uint32_t op, size;
		std::memcpy(&op, p, 4); std::memcpy(&size, p + 4, 4);
		if (size < 8 || (size & 7) || p + size > end) { return fail(RSRCU_ERR_INVALID, "malformed stream record (op %u size %u)", op, size); }
		const uint8_t* q = p + 8;
		const size_t payload = size - 8;
		auto i32 = [&](int k) { int32_t v; std::memcpy(&v, q + 4 * k, 4); return v; };
		auto u64 = [&](int byteOfs) { uint64_t v; std::memcpy(&v, q + byteOfs, 8); return v; };
		auto need = [&](size_t n) { return payload >= n; };
		int r = RSRCU_OK;
		switch (op) {
		case RSRCU_OP_BEGIN_FRAME: if (!need(16)) { goto bad; } r = rsrcu_begin_frame(c, i32(0), i32(1), i32(2), i32(3)); break;
		case RSRCU_OP_STATE:
			// read in place: the stream outlives the frame it records (rsrcu_end_frame runs inside this call)
			if (!need(sizeof(RsrState)) || !c->inFrame) { goto bad; }
			c->curStatePtr = reinterpret_cast<const RsrState*>(q); c->haveState = true; c->stateDirty = true; break;
		case RSRCU_OP_BIND_BUFFER: if (!need(24)) { goto bad; }
			r = rsrcu_bind_buffer(c, i32(0), reinterpret_cast<const float*>(u64(8)), static_cast<size_t>(u64(16)), i32(1)); break;
		case RSRCU_OP_BIND_TEXTURE: if (!need(40)) { goto bad; }
			r = rsrcu_bind_texture(c, i32(0), reinterpret_cast<const float*>(u64(32)), i32(1), i32(2), i32(3), i32(4), i32(5), i32(6)); break;
		case RSRCU_OP_BIND_DEPTH: if (!need(16)) { goto bad; }
			r = rsrcu_bind_depth_texture(c, reinterpret_cast<const float*>(u64(8)), i32(0), i32(1)); break;
		case RSRCU_OP_CLEAR: if (!need(8)) { goto bad; } r = rsrcu_clear(c, i32(0)); break;
		case RSRCU_OP_DRAW_ELEMENTS: if (!need(24)) { goto bad; }
			r = rsrcu_draw_elements(c, i32(0), reinterpret_cast<const uint16_t*>(u64(16)), i32(1), i32(2), i32(3)); break;
		case RSRCU_OP_DRAW_ARRAYS: if (!need(8)) { goto bad; } r = rsrcu_draw_arrays(c, i32(0), i32(1)); break;
		case RSRCU_OP_STORE_TC: if (!need(24)) { goto bad; }
			r = rsrcu_store_color_tc(c, i32(0), reinterpret_cast<uint32_t*>(u64(16)), i32(1), i32(2), i32(3)); break;
		case RSRCU_OP_STORE_FP: if (!need(24)) { goto bad; }
			r = rsrcu_store_color_fp(c, reinterpret_cast<float*>(u64(16)), i32(1), i32(2), i32(3), i32(0)); break;
		case RSRCU_OP_STORE_QUADS: if (!need(24)) { goto bad; }
			r = rsrcu_store_color_quads(c, reinterpret_cast<float*>(u64(16)), i32(1), i32(2), i32(3)); break;
		case RSRCU_OP_STORE_DEPTH: if (!need(8)) { goto bad; } r = rsrcu_store_depth(c, reinterpret_cast<float*>(u64(0))); break;
		case RSRCU_OP_END_FRAME: r = rsrcu_end_frame(c); break;
		case RSRCU_OP_STORE_TC_DEV: if (!need(24)) { goto bad; }
			r = rsrcu_store_color_tc_device(c, i32(0), reinterpret_cast<void*>(u64(16)), i32(1), i32(2), i32(3)); break;
		case RSRCU_OP_STORE_FP_DEV: if (!need(24)) { goto bad; }
			r = rsrcu_store_color_fp_device(c, reinterpret_cast<void*>(u64(16)), i32(1), i32(2), i32(3), i32(0)); break;
		case RSRCU_OP_STORE_QUADS_DEV: if (!need(24)) { goto bad; }
			r = rsrcu_store_color_quads_device(c, reinterpret_cast<void*>(u64(16)), i32(1), i32(2), i32(3)); break;
		case RSRCU_OP_STORE_DEPTH_DEV: if (!need(8)) { goto bad; } r = rsrcu_store_depth_device(c, reinterpret_cast<void*>(u64(0))); break;
		default: return fail(RSRCU_ERR_INVALID, "unknown stream opcode %u", op); }
		if (r != RSRCU_OK) { return r; }
		p += size;
		continue;
	bad:
		return fail(RSRCU_ERR_INVALID, "stream record op %u too short (%zu bytes)", op, payload); }
	if (c->inFrame && c->haveState && c->curStatePtr != &c->curState) {
		// the frame continues in another call: keep the current state, the stream may go away
		c->curState = *c->curStatePtr;
		c->curStatePtr = &c->curState; }
	return RSRCU_OK; }

int rsrcu_sync_frame(rsrcu_ctx* c, int lag) {
	if (!c || lag < 0 || lag >= kSlots) { return fail(RSRCU_ERR_INVALID, "lag must be 0 .. %d", kSlots - 1); }
	CU(cudaSetDevice(c->device));
	if (lag == 0) { const int r = flushDeferredCopies(c); if (r != RSRCU_OK) { return r; } }
	const int slot = (c->outSlot + kSlots - lag) % kSlots;
	if (!c->launched[slot].plan) { CU(cudaEventSynchronize(c->evCopied[slot])); return RSRCU_OK; }
	bool relaunched = false;
	const int r = settleSlot(c, slot, false, relaunched);
	c->stats.frames_retried = c->framesRetried;
	return r; }

int rsrcu_device_truecolor(rsrcu_ctx* c, void** devPtr, int* stridePx) {
	if (!c || !devPtr || !stridePx) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	*devPtr = c->shownTcDev; *stridePx = c->shownTcStride;
	return RSRCU_OK; }

int rsrcu_stream(rsrcu_ctx* c, void** stream) {
	if (!c || !stream) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	*stream = c->stream;
	return RSRCU_OK; }

int rsrcu_join(rsrcu_ctx* c) {
	if (!c) { return fail(RSRCU_ERR_INVALID, "null context"); }
	CU(cudaSetDevice(c->device));
	return joinTileStreams(c); }

// ---- device canvases, post filters, presentation ----------------------------------------------

int rsrcu_canvas_alloc(rsrcu_ctx* c, size_t bytes, void** devicePtr) {
	if (!c || !devicePtr || bytes == 0) { return fail(RSRCU_ERR_INVALID, "null argument / empty canvas"); }
	CU(cudaSetDevice(c->device));
	void* p = nullptr;
	CU(cudaMalloc(&p, bytes));
	CU(cudaMemset(p, 0, bytes));
	c->canvases.push_back(p);
	*devicePtr = p;
	return RSRCU_OK; }

int rsrcu_canvas_free(rsrcu_ctx* c, void* devicePtr) {
	if (!c) { return fail(RSRCU_ERR_INVALID, "null context"); }
	auto it = std::find(c->canvases.begin(), c->canvases.end(), devicePtr);
	if (it == c->canvases.end()) { return fail(RSRCU_ERR_INVALID, "not a canvas of this context"); }
	CU(cudaSetDevice(c->device));
	CU(cudaDeviceSynchronize());   // (other contexts may still read it)
	CU(cudaFree(devicePtr));
	c->canvases.erase(it);
	return RSRCU_OK; }

int rsrcu_canvas_read(rsrcu_ctx* c, const void* devicePtr, void* hostDst, size_t bytes) {
	if (!c || !devicePtr || !hostDst) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	CU(cudaSetDevice(c->device));
	{ const int r = joinTileStreams(c); if (r != RSRCU_OK) { return r; } }
	CU(cudaMemcpyAsync(hostDst, devicePtr, bytes, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return RSRCU_OK; }

int rsrcu_canvas_write(rsrcu_ctx* c, void* devicePtr, const void* hostSrc, size_t bytes) {
	if (!c || !devicePtr || !hostSrc) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	CU(cudaSetDevice(c->device));
	{ const int r = joinTileStreams(c); if (r != RSRCU_OK) { return r; } }
	CU(cudaMemcpyAsync(devicePtr, hostSrc, bytes, cudaMemcpyHostToDevice, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return RSRCU_OK; }

int rsrcu_store_color_fp_device(rsrcu_ctx* c, void* deviceDst, int width, int height, int stridePx, int half) {
	if (!c || !c->inFrame) { return fail(RSRCU_ERR_INVALID, "store outside begin/end frame"); }
	if (!deviceDst) { return fail(RSRCU_ERR_INVALID, "null device destination"); }
	const int wantW = half ? c->width / 2 : c->width, wantH = half ? c->height / 2 : c->height;
	if (width != wantW || height != wantH) { return fail(RSRCU_ERR_INVALID, "store canvas %dx%d != %s target %dx%d", width, height, half ? "half" : "full", wantW, wantH); }
	if (stridePx < width) { return fail(RSRCU_ERR_INVALID, "canvas stride %d < width %d", stridePx, width); }
	return pushCmd(c, half ? kCmdStoreHalfFP : kCmdStoreFP, 0, deviceDst, stridePx, 2); }

int rsrcu_store_color_quads_device(rsrcu_ctx* c, void* deviceDst, int width, int height, int strideQuads) {
	if (!c || !c->inFrame) { return fail(RSRCU_ERR_INVALID, "store outside begin/end frame"); }
	if (!deviceDst) { return fail(RSRCU_ERR_INVALID, "null device destination"); }
	if (width != c->width || height != c->height) { return fail(RSRCU_ERR_INVALID, "store canvas %dx%d != target %dx%d", width, height, c->width, c->height); }
	if (strideQuads < width / 2) { return fail(RSRCU_ERR_INVALID, "quad canvas stride %d < %d quads per row", strideQuads, width / 2); }
	return pushCmd(c, kCmdStoreQuadsFP, 0, deviceDst, strideQuads, 2); }

int rsrcu_store_depth_device(rsrcu_ctx* c, void* deviceDst) {
	if (!c || !c->inFrame) { return fail(RSRCU_ERR_INVALID, "store outside begin/end frame"); }
	if (!deviceDst) { return fail(RSRCU_ERR_INVALID, "null device destination"); }
	return pushCmd(c, kCmdStoreDepth, 0, deviceDst, c->width, 3); }

int rsrcu_wait_for(rsrcu_ctx* c, rsrcu_ctx* producer) {
	if (!c || !producer) { return fail(RSRCU_ERR_INVALID, "null context"); }
	if (c == producer) { return RSRCU_OK; }
	if (c->device != producer->device) { return fail(RSRCU_ERR_INVALID, "rsrcu_wait_for: contexts on devices %d and %d (use the completion counters across GPUs)", c->device, producer->device); }
	CU(cudaSetDevice(c->device));
	{ const int r = joinTileStreams(producer); if (r != RSRCU_OK) { return r; } }
	if (c->extWaitPending) {   // an earlier producer nobody has waited for yet: fold it into the context's stream first
		CU(cudaStreamWaitEvent(c->stream, c->evExt, 0));
		if (c->overlap) { CU(cudaStreamWaitEvent(c->frontStream, c->evExt, 0)); } }
	CU(cudaEventRecord(c->evExt, producer->stream));
	CU(cudaStreamWaitEvent(c->stream, c->evExt, 0));   // post filters and copies of this context
	c->extWaitPending = true;                             // the next frame's front end (its own stream in overlap mode)
	return RSRCU_OK; }

int rsrcu_kawase_blur(rsrcu_ctx* c, const void* src, int srcStride, void* dst, int dstStride, int width, int height, int dist) {
	if (!c || !src || !dst) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	if (src == dst) { return fail(RSRCU_ERR_INVALID, "rsrcu_kawase_blur: src and dst must differ (the reference ping-pongs two canvases, node/kawase.cxx:96-124)"); }
	if (width <= 0 || height <= 0 || dist < 0 || srcStride < width || dstStride < width) { return fail(RSRCU_ERR_INVALID, "bad canvas geometry"); }
	CU(cudaSetDevice(c->device));
	{ const int r = joinTileStreams(c); if (r != RSRCU_OK) { return r; } }
	// rows per thread: about 300 k threads on a small canvas, at most 16 rows (the first row of a strip costs twice the loads)
	const int rows = static_cast<int>(std::min<long long>(16, std::max<long long>(4, static_cast<long long>(width) * height / 300000)));
	const dim3 grid(static_cast<unsigned>((width + 127) / 128), static_cast<unsigned>((height + rows - 1) / rows));
	kawase_kernel<<<grid, 128, 0, c->stream>>>(static_cast<const float4*>(src), srcStride, static_cast<float4*>(dst), dstStride, width, height, dist, rows);
	CU(cudaGetLastError());
	return RSRCU_OK; }

int rsrcu_make_mipmap(rsrcu_ctx* c, void* texelsDevice, int dim) {
	if (!c || !texelsDevice) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	if (dim < 2 || (dim & (dim - 1)) != 0 || dim > 4096) { return fail(RSRCU_ERR_INVALID, "mipmap: %d is not a power of two in 2 .. 4096 (rglr_texture.cxx:35-52)", dim); }
	CU(cudaSetDevice(c->device));
	{ const int r = joinTileStreams(c); if (r != RSRCU_OK) { return r; } }
	int pow = 0; while ((1 << pow) < dim) { ++pow; }
	int level = 0, size = dim;
	while (level < pow) {
		const int k = std::min(5, pow - level);
		const unsigned blocks = static_cast<unsigned>(size >> k);
		mipmap_kernel<<<dim3(blocks, blocks), 256, 0, c->stream>>>(static_cast<float4*>(texelsDevice), dim, dim, level, size, k);
		CU(cudaGetLastError());
		level += k; size >>= k; }
	return RSRCU_OK; }

int rsrcu_glow(rsrcu_ctx* c, const void* imageQuads, int imageStrideQuads, const void* blur, int blurStride, int gamma,
               uint32_t* dst, int dstIsDevice, int width, int height, int stridePx) {
	if (!c || !imageQuads || !blur || !dst) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	if (width <= 0 || height <= 0 || (width & 3) || (height & 1)) { return fail(RSRCU_ERR_INVALID, "glow: width must be a multiple of 4 and height of 2 (the filter walks 4x2 pixels, rglr_algorithm.hxx:129)"); }
	if (imageStrideQuads < width / 2 || blurStride < width / 2 || stridePx < width) { return fail(RSRCU_ERR_INVALID, "bad canvas stride"); }
	CU(cudaSetDevice(c->device));
	{ const int r = joinTileStreams(c); if (r != RSRCU_OK) { return r; } }
	uint32_t* target = dst;
	int targetStride = stridePx;
	if (!dstIsDevice) {
		CU(c->glowOut.reserve(static_cast<size_t>(width) * height * 4));
		target = static_cast<uint32_t*>(c->glowOut.ptr); targetStride = width; }
	const dim3 grid(static_cast<unsigned>((width / 4 + 31) / 32), static_cast<unsigned>((height / 2 + 7) / 8));
	glow_kernel<<<grid, 256, 0, c->stream>>>(static_cast<const float4*>(imageQuads), imageStrideQuads, static_cast<const float4*>(blur), blurStride,
	                                         target, targetStride, width, height, gamma ? 1 : 0);
	CU(cudaGetLastError());
	c->shownTcDev = target; c->shownTcStride = targetStride;
	if (!dstIsDevice) {
		CU(cudaMemcpy2DAsync(dst, static_cast<size_t>(stridePx) * 4, target, static_cast<size_t>(targetStride) * 4, static_cast<size_t>(width) * 4,
		                     static_cast<size_t>(height), cudaMemcpyDeviceToHost, c->stream));
		CU(cudaStreamSynchronize(c->stream)); }
	return RSRCU_OK; }

int rsrcu_draw_spans(rsrcu_ctx* c, void* truecolor, int stridePx, int width, int height, int left, int top, float xscale,
                     const RsrSpan* spans, int count) {
	if (!c || !truecolor || (count > 0 && !spans)) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	if (count <= 0) { return RSRCU_OK; }
	static_assert(sizeof(RsrSpan) == sizeof(DevSpan), "span layout");
	int lanes = 0;
	for (int i = 0; i < count; ++i) {
		if (spans[i].lane < 0 || spans[i].lane >= 4096) { return fail(RSRCU_ERR_INVALID, "span lane %d out of range", spans[i].lane); }
		lanes = std::max(lanes, spans[i].lane + 1); }
	CU(cudaSetDevice(c->device));
	{ const int r = joinTileStreams(c); if (r != RSRCU_OK) { return r; } }
	CU(cudaStreamSynchronize(c->stream));   // (spanBuf may still be read by an earlier call; overlays are drawn once per presented frame)
	CU(c->spanBuf.reserve(static_cast<size_t>(count) * sizeof(DevSpan)));
	CU(cudaMemcpyAsync(c->spanBuf.ptr, spans, static_cast<size_t>(count) * sizeof(DevSpan), cudaMemcpyHostToDevice, c->stream));
	// const auto scale = xscale * canvas.width();  (jobsys_vis.cxx:30)
	const float scale = xscale * static_cast<float>(width);
	spans_kernel<<<static_cast<unsigned>(lanes), 256, 0, c->stream>>>(static_cast<uint32_t*>(truecolor), stridePx, width, height, left, top, scale,
	                                                                  static_cast<const DevSpan*>(c->spanBuf.ptr), count);
	CU(cudaGetLastError());
	return RSRCU_OK; }

int rsrcu_frame_spans(rsrcu_ctx* c, RsrSpan* out, int capacity, int* count) {
	if (!c || !out || !count) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	*count = 0;
	if (!c->stageSpansValid) { return fail(RSRCU_ERR_INVALID, "no profiled frame: rsrcu_set_profiling(ctx, 2), render, rsrcu_sync"); }
	// evStage: 0 start, 1 after upload, 2 after vertex, 3 after setup, 4 after count, 5 after scan, 6 after fill / tile start, 7 after tile
	static const int from[6] = {0, 1, 2, 4, 5, 6}, to[6] = {1, 2, 3, 5, 6, 7};
	for (int lane = 0; lane < 6 && *count < capacity; ++lane) {
		const double a = static_cast<double>(c->stageStartMs[from[lane]]) * 1e-3, b = static_cast<double>(c->stageStartMs[to[lane]]) * 1e-3;
		if (b <= a) { continue; }
		out[*count] = RsrSpan{a, b, static_cast<uint32_t>(lane) * 0x9e3779b1u, lane};
		++*count; }
	return RSRCU_OK; }

int rsrcu_present(rsrcu_ctx* c, const void* truecolor, int srcStride, void* surface, int surfaceStride, int width, int height) {
	if (!c || !truecolor || !surface) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	if (width <= 0 || height <= 0 || srcStride < width || surfaceStride < width) { return fail(RSRCU_ERR_INVALID, "bad surface geometry"); }
	CU(cudaSetDevice(c->device));
	{ const int r = joinTileStreams(c); if (r != RSRCU_OK) { return r; } }
	CU(cudaMemcpy2DAsync(surface, static_cast<size_t>(surfaceStride) * 4, truecolor, static_cast<size_t>(srcStride) * 4, static_cast<size_t>(width) * 4,
	                     static_cast<size_t>(height), cudaMemcpyDeviceToDevice, c->stream));
	return RSRCU_OK; }

// ---- marching cubes on the device (kernels in march_kernels.cuh) ---------------------------------------

namespace {

// Bourke's polygonise case table, packed (see march_kernels.cuh: kMcTri); printed by tools/gen_mc_table.py
const uint64_t kMcTriHost[256] = {
	0xffffffffffffffffull, 0xfffffffffffff380ull, 0xfffffffffffff910ull, 0xffffffffff189381ull,
	0xfffffffffffffa21ull, 0xffffffffffa21380ull, 0xffffffffff920a29ull, 0xfffffff89a8a2382ull,
	0xfffffffffffff2b3ull, 0xffffffffff0b82b0ull, 0xffffffffffb32091ull, 0xfffffffb89b912b1ull,
	0xffffffffff3ab1a3ull, 0xfffffffab8a801a0ull, 0xfffffff9ab9b3093ull, 0xffffffffffb8aa89ull,
	0xfffffffffffff874ull, 0xffffffffff437034ull, 0xffffffffff748910ull, 0xfffffff137174914ull,
	0xffffffffff748a21ull, 0xfffffffa21403743ull, 0xfffffff748209a29ull, 0xffff4973727929a2ull,
	0xffffffffff2b3748ull, 0xfffffff40242b74bull, 0xfffffffb32748109ull, 0xffff1292b9b49b74ull,
	0xfffffff487ab31a3ull, 0xffff4b7401b41ab1ull, 0xffff30bab9b09874ull, 0xfffffffab99b4b74ull,
	0xfffffffffffff459ull, 0xffffffffff380459ull, 0xffffffffff051450ull, 0xfffffff513538458ull,
	0xffffffffff459a21ull, 0xfffffff594a21803ull, 0xfffffff204245a25ull, 0xffff8434535235a2ull,
	0xffffffffffb32459ull, 0xfffffff594b802b0ull, 0xfffffffb32510450ull, 0xffff584b82852512ull,
	0xfffffff45931ab3aull, 0xffffab81a8180594ull, 0xffff30bab5b05045ull, 0xfffffffb8aa85845ull,
	0xffffffffff975879ull, 0xfffffff375359039ull, 0xfffffff751710870ull, 0xffffffffff753351ull,
	0xfffffff21a759879ull, 0xffff37503505921aull, 0xffff25a758528208ull, 0xfffffff7533525a2ull,
	0xfffffff2b3987597ull, 0xffffb72029279759ull, 0xffff751871810b32ull, 0xfffffff51771b12bull,
	0xffffb3a31a758859ull, 0xf0aba010b7905075ull, 0xf07570805a30b0abull, 0xffffffffff5b75abull,
	0xfffffffffffff56aull, 0xffffffffff6a5380ull, 0xffffffffff6a5109ull, 0xfffffff6a5891381ull,
	0xffffffffff162561ull, 0xfffffff803621561ull, 0xfffffff620609569ull, 0xffff823625285895ull,
	0xffffffffff56ab32ull, 0xfffffff56a02b80bull, 0xfffffff6a5b32910ull, 0xffffb892b92916a5ull,
	0xfffffff315356b36ull, 0xffff6b51505b0b80ull, 0xffff9505606306b3ull, 0xfffffff89bb96956ull,
	0xffffffffff8746a5ull, 0xfffffffa56374034ull, 0xfffffff7486a5091ull, 0xffff49737179156aull,
	0xfffffff874156216ull, 0xffff743403625521ull, 0xffff620560509748ull, 0xf962695923497937ull,
	0xfffffff56a4872b3ull, 0xffffb720242746a5ull, 0xffff6a5b32874910ull, 0xf6a54b7b492b9129ull,
	0xffff6b51535b3748ull, 0xfb404b7b016b5b15ull, 0xf74836b630560950ull, 0xffff9b7974b96956ull,
	0xffffffffffa4694aull, 0xfffffff380a946a4ull, 0xfffffff04606a10aull, 0xffffa16468618138ull,
	0xfffffff462421941ull, 0xffff462942921803ull, 0xffffffffff624420ull, 0xfffffff624428238ull,
	0xfffffff32b46a94aull, 0xffff6a4a94b82280ull, 0xffffa164606102b3ull, 0xf1b8b12184a16146ull,
	0xffff36b319639469ull, 0xf14641916b0181b8ull, 0xfffffff4600636b3ull, 0xffffffffff86b846ull,
	0xfffffffa98a876a7ull, 0xffffa76a907a0370ull, 0xffff0818717a176aull, 0xfffffff37117a76aull,
	0xffff768981861621ull, 0xf937390976192962ull, 0xfffffff206607087ull, 0xffffffffff276237ull,
	0xffff76898a86ab32ull, 0xf7a9a76790b72702ull, 0xfb32a767a1871081ull, 0xffff17616a71b12bull,
	0xf63136b619768698ull, 0xffffffffff76b190ull, 0xffff06b0b3607087ull, 0xfffffffffffff6b7ull,
	0xfffffffffffffb67ull, 0xffffffffff67b803ull, 0xffffffffff67b910ull, 0xfffffff67b138918ull,
	0xffffffffff7b621aull, 0xfffffff7b6803a21ull, 0xfffffff7b69a2092ull, 0xffff89a38a3a27b6ull,
	0xffffffffff726327ull, 0xfffffff026067807ull, 0xfffffff910732672ull, 0xffff678891681261ull,
	0xfffffff73171a67aull, 0xffff801781a7167aull, 0xffff7a69a0a70730ull, 0xfffffff9a88a7a67ull,
	0xffffffffff68b486ull, 0xfffffff640603b63ull, 0xfffffff109648b68ull, 0xffff63b139369649ull,
	0xfffffff1a28b6486ull, 0xffff640b60b03a21ull, 0xffff9a2920b648b4ull, 0xf36463b34923a39aull,
	0xfffffff264248328ull, 0xffffffffff264240ull, 0xffff834642432091ull, 0xfffffff642241491ull,
	0xffff1a6648168318ull, 0xfffffff40660a01aull, 0xf39a9303a6834364ull, 0xffffffffff4a649aull,
	0xffffffffffb67594ull, 0xfffffff67b594380ull, 0xfffffffb67045105ull, 0xffff51345343867bull,
	0xfffffffb6721a459ull, 0xffff594380a217b6ull, 0xffff204a24a45b67ull, 0xf67b25a523453843ull,
	0xfffffff945267327ull, 0xffff786260680459ull, 0xffff045051673263ull, 0xf851584812786826ull,
	0xffff73167161a459ull, 0xf459078701671a61ull, 0xfa737a6a305a4a04ull, 0xffffa84a458a7a67ull,
	0xfffffff98b9b6596ull, 0xffff590650360b63ull, 0xffffb65510b508b0ull, 0xfffffff1355363b6ull,
	0xffff65b8b9b59a21ull, 0xfa21965690b603b0ull, 0xf52025a50865b58bull, 0xffff35a3a25363b6ull,
	0xffff283265825985ull, 0xfffffff260069659ull, 0xf826283865081851ull, 0xffffffffff612651ull,
	0xf698965683a61631ull, 0xffff06505960a01aull, 0xffffffffffa65830ull, 0xfffffffffffff65aull,
	0xffffffffffb57a5bull, 0xfffffff03857ba5bull, 0xfffffff091ba57b5ull, 0xffff1381897ba57aull,
	0xfffffff15717b21bull, 0xffffb27571721380ull, 0xffff7b2209729579ull, 0xf289823295b27257ull,
	0xfffffff573532a52ull, 0xffff52a578258028ull, 0xffff2a37353a5109ull, 0xf25752a278129289ull,
	0xffffffffff573531ull, 0xfffffff571170780ull, 0xfffffff735539309ull, 0xffffffffff795789ull,
	0xfffffff8ba8a5485ull, 0xffff03bba50b5405ull, 0xffff54aba8a48910ull, 0xf41314943b54a4baull,
	0xffff8548b2582152ull, 0xfb151b2b543b0b40ull, 0xf58b8545b2950520ull, 0xffffffffff3b2549ull,
	0xffff483543253a52ull, 0xfffffff0244252a5ull, 0xf910854583a532a3ull, 0xffff2492914252a5ull,
	0xfffffff153358548ull, 0xffffffffff501540ull, 0xffff530509358548ull, 0xfffffffffffff549ull,
	0xfffffffba9b947b4ull, 0xffffba97b9794380ull, 0xffffb470414b1ba1ull, 0xf4bab474a1843413ull,
	0xffff219b294b97b4ull, 0xf3801b2b197b9479ull, 0xfffffff04224b47bull, 0xffff42343824b47bull,
	0xffff947732972a92ull, 0xf70207872a4797a9ull, 0xfa040a1a472a3a73ull, 0xffffffffff4782a1ull,
	0xfffffff317714194ull, 0xffff178180714194ull, 0xffffffffff347304ull, 0xfffffffffffff784ull,
	0xffffffffff8ba8a9ull, 0xfffffffa9bb93903ull, 0xfffffffba88a0a10ull, 0xffffffffffa3ba13ull,
	0xfffffff8b99b1b21ull, 0xffff9b2921b93903ull, 0xffffffffffb08b20ull, 0xfffffffffffffb23ull,
	0xfffffff98aa82832ull, 0xffffffffff2902a9ull, 0xffff8a1810a82832ull, 0xfffffffffffff2a1ull,
	0xffffffffff819831ull, 0xfffffffffffff190ull, 0xfffffffffffff830ull, 0xffffffffffffffffull };

struct McAabb { float ltb[3], rbf[3]; };

// rmlv::mix(a, b, 0.5F) = (1 - t) * a + t * b (rmlv_math.hxx:87-89); this file is compiled without FMA contraction
inline float mixHalf(float a, float b) { const float t = 0.5f; return (1.0f - t) * a + t * b; }

// BlockDivider::Compute (src/viewer/node/mc.cxx:39-88): octants in the reference's order, midpoints by mix()
void mcDivide(const McAabb& b, int limit, std::vector<McAabb>& out) {
	if (limit == 0) { out.push_back(b); return; }
	float mid[3];
	for (int k = 0; k < 3; ++k) { mid[k] = mixHalf(b.ltb[k], b.rbf[k]); }
	for (int o = 0; o < 8; ++o) {
		McAabb s;
		for (int k = 0; k < 3; ++k) {
			const bool hi = (o >> k) & 1;   // bit 0: right half in x, bit 1: lower half in y, bit 2: front half in z
			s.ltb[k] = hi ? mid[k] : b.ltb[k];
			s.rbf[k] = hi ? b.rbf[k] : mid[k]; }
		mcDivide(s, limit - 1, out); } }

// Surface::sample(vec3) (mc.cxx:97-101) on the host: the block-level distance test of ResolveImpl (mc.cxx:241-245) is
// taken here, with this host's libm like the reference
float mcFieldHost(float px, float py, float pz, float T) {
	const float distort = 0.60f * sinf(5.0f * (px + T / 4.0f)) * sinf(2.0f * (py + (T / 1.33f)));
	const float len = sqrtf(px * px + py * py + pz * pz);
	return (len - 3.0f) + (distort * sinf(T / 2.0f) + 1.0f); }

}  // namespace

int rsrcu_march_surface(rsrcu_ctx* c, float timeSeconds, int precision, int forkDepth, float range,
                        const float** outSoa6, RsrMarchBlock* blocks, int blockCapacity, int* blockCount, int* vertexTotal) {
	if (!c || !outSoa6 || !blocks || !blockCount || !vertexTotal) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	if (precision != 8 && precision != 16 && precision != 32 && precision != 64 && precision != 128) {
		return fail(RSRCU_ERR_INVALID, "precision %d must be one of 8, 16, 32, 64, 128 (mc.cxx:344-354)", precision); }
	if (forkDepth < 0 || forkDepth > 4) { return fail(RSRCU_ERR_INVALID, "forkDepth %d exceeds limit 0 .. 4 (mc.cxx:357-360)", forkDepth); }
	if (!(range >= 0.0001f)) { return fail(RSRCU_ERR_INVALID, "range %g is too small (mc.cxx:363-366)", static_cast<double>(range)); }
	const int dim = precision >> forkDepth;
	if (dim < 1 || dim > kMcMaxDim) {
		return fail(RSRCU_ERR_UNSUPPORTED, "%d cells per block edge: 1 .. %d supported (the reference's 64-wide slices hold dim + 1 values plus padding)", dim, kMcMaxDim); }
	CU(cudaSetDevice(c->device));
	{ const int r = joinTileStreams(c); if (r != RSRCU_OK) { return r; } }
	static std::once_flag once;
	static cudaError_t tableErr = cudaSuccess;
	std::call_once(once, [&]() { tableErr = cudaMemcpyToSymbol(kMcTri, kMcTriHost, sizeof(kMcTriHost)); });
	CU(tableErr);

	// Main (mc.cxx:171-193): the root block, its subdivision, one job per block -- here one CTA per block that passes
	// the distance test
	std::vector<McAabb> all;
	mcDivide(McAabb{{-range, range, -range}, {range, -range, range}}, forkDepth, all);
	std::vector<McBlock> active;
	active.reserve(all.size());
	for (const McAabb& b : all) {
		const float delta = (b.rbf[0] - b.ltb[0]) / static_cast<float>(dim);
		const float mx = mixHalf(b.ltb[0], b.rbf[0]), my = mixHalf(b.ltb[1], b.rbf[1]), mz = mixHalf(b.ltb[2], b.rbf[2]);
		const float dx = b.ltb[0] - mx, dy = b.ltb[1] - my, dz = b.ltb[2] - mz;
		const float R = sqrtf(dx * dx + dy * dy + dz * dz);
		const float D = mcFieldHost(mx, my, mz, timeSeconds);
		if (fabsf(D) * 0.5f > R) { continue; }
		active.push_back(McBlock{b.ltb[0], b.ltb[1], b.ltb[2], delta}); }
	*blockCount = 0; *vertexTotal = 0;
	for (int k = 0; k < 6; ++k) { outSoa6[k] = nullptr; }
	const int n = static_cast<int>(active.size());
	if (n == 0) { return RSRCU_OK; }

	// slabs: enough CTAs to fill the GPU when there are few large blocks (a slab pays one extra slice: at least 2 layers)
	const int slabLayers = std::min(dim, std::max(2, static_cast<int>(static_cast<long long>(dim) * n / 592)));
	const int slabs = (dim + slabLayers - 1) / slabLayers;
	const size_t nslab = static_cast<size_t>(n) * slabs;
	CU(c->mcBlocks.reserve(static_cast<size_t>(n) * sizeof(McBlock)));
	CU(c->mcTotals.reserve((nslab + static_cast<size_t>(n)) * 4));   // slab totals | block totals
	CU(c->mcBase.reserve((nslab + 1) * 4));
	uint32_t* dSlabTotals = static_cast<uint32_t*>(c->mcTotals.ptr);
	uint32_t* dBlockTotals = dSlabTotals + nslab;
	CU(cudaMemcpyAsync(c->mcBlocks.ptr, active.data(), static_cast<size_t>(n) * sizeof(McBlock), cudaMemcpyHostToDevice, c->stream));
	McOut out{};
	march_kernel<false><<<static_cast<unsigned>(nslab), 256, 0, c->stream>>>(static_cast<const McBlock*>(c->mcBlocks.ptr), timeSeconds, dim, slabs, slabLayers,
		dSlabTotals, nullptr, out);
	march_scan_kernel<<<1, 256, 0, c->stream>>>(dSlabTotals, static_cast<uint32_t*>(c->mcBase.ptr), dBlockTotals, n, slabs);
	CU(cudaGetLastError());
	std::vector<uint32_t> totals(static_cast<size_t>(n)), slabBase(nslab + 1);
	CU(cudaMemcpyAsync(totals.data(), dBlockTotals, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaMemcpyAsync(slabBase.data(), c->mcBase.ptr, (nslab + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	const uint32_t padded = slabBase[nslab];
	if (padded == 0) { return RSRCU_OK; }
	std::vector<uint32_t> base(static_cast<size_t>(n));
	for (int i = 0; i < n; ++i) { base[static_cast<size_t>(i)] = slabBase[static_cast<size_t>(i) * slabs]; }
	int nonEmpty = 0;
	for (int i = 0; i < n; ++i) { nonEmpty += totals[static_cast<size_t>(i)] ? 1 : 0; }
	if (nonEmpty > blockCapacity) { return fail(RSRCU_ERR_INVALID, "%d non-empty blocks, room for %d", nonEmpty, blockCapacity); }

	// the vertex arrays rotate over three buffers like the node's (mc.cxx:115, :206-208): a frame still in flight keeps
	// reading the arrays of the previous call
	DevBuf& vb = c->mcVerts[c->mcTurn];
	c->mcTurn = (c->mcTurn + 1) % 3;
	CU(vb.reserve(static_cast<size_t>(padded) * 6 * 4));
	CU(cudaMemsetAsync(vb.ptr, 0, static_cast<size_t>(padded) * 6 * 4, c->stream));   // the padding vertices are zeros (VertexArray_F3F3F3::pad)
	for (int k = 0; k < 6; ++k) { out.a[k] = static_cast<float*>(vb.ptr) + static_cast<size_t>(k) * padded; outSoa6[k] = out.a[k]; }
	march_kernel<true><<<static_cast<unsigned>(nslab), 256, 0, c->stream>>>(static_cast<const McBlock*>(c->mcBlocks.ptr), timeSeconds, dim, slabs, slabLayers,
		nullptr, static_cast<const uint32_t*>(c->mcBase.ptr), out);
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(c->stream));   // (frames may run their front end on another stream: the arrays are complete on return)
	int nb = 0;
	for (int i = 0; i < n; ++i) {
		if (totals[static_cast<size_t>(i)]) { blocks[nb++] = RsrMarchBlock{static_cast<int32_t>(base[static_cast<size_t>(i)]), static_cast<int32_t>(totals[static_cast<size_t>(i)])}; } }
	*blockCount = nb;
	*vertexTotal = static_cast<int>(padded);
	return RSRCU_OK; }

int rsrcu_set_pin_in_place(rsrcu_ctx* c, int enabled) {
	if (!c) { return fail(RSRCU_ERR_INVALID, "null context"); }
	if (c->inFrame) { return fail(RSRCU_ERR_INVALID, "rsrcu_set_pin_in_place inside begin/end frame"); }
	CU(cudaSetDevice(c->device));
	c->pinInPlace = enabled != 0;
	if (!c->pinInPlace && !c->pinned.empty()) {
		// everything submitted has to be through with the buffers before the pages are released
		{ const int r = rsrcu_sync(c); if (r != RSRCU_OK) { return r; } }
		CU(cudaStreamSynchronize(c->h2dStream));
		for (auto& kv : c->pinned) {
			if (kv.second.owned) { cudaHostUnregister(reinterpret_cast<void*>(kv.second.regBase)); cudaGetLastError(); }
			for (auto& b : kv.second.dev) { b.release(); } }
		c->pinned.clear(); }
	return RSRCU_OK; }

int rsrcu_get_stats(rsrcu_ctx* c, RsrStats* out) {
	if (!c || !out) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	*out = c->stats;
	return RSRCU_OK; }

#ifdef RSR_PHASE_PROF
/* developer builds only (rsr_b200.build.build_variant("phase", ["RSR_PHASE_PROF"])): accumulated clock64 cycles per tile-kernel phase; clears them */
int rsrcu_debug_phase_cycles(unsigned long long* out16) {
	if (cudaMemcpyFromSymbol(out16, g_phaseCycles, sizeof(unsigned long long) * 16) != cudaSuccess) { return RSRCU_ERR_CUDA; }
	unsigned long long zero[16] = {};
	cudaMemcpyToSymbol(g_phaseCycles, zero, sizeof(zero));
	return RSRCU_OK; }
int rsrcu_debug_k2_times(unsigned long long* out8) {
	if (cudaMemcpyFromSymbol(out8, g_k2Times, sizeof(unsigned long long) * 8) != cudaSuccess) { return RSRCU_ERR_CUDA; }
	cudaMemcpyFromSymbol(out8 + 8, g_k2Block, sizeof(unsigned long long) * 8);
	unsigned long long init[8] = { ~0ull, 0, 0, 0, 0, 0, 0, 0 };
	unsigned long long zero12[12] = {};
	cudaMemcpyToSymbol(g_k2Block, zero12, sizeof(zero12));
	cudaMemcpyToSymbol(g_k2Times, init, sizeof(init));
	return RSRCU_OK; }
#endif

int rsrcu_signal_counter(rsrcu_ctx* c, void* deviceCounter) {
	if (!c || !deviceCounter) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	if (reinterpret_cast<uintptr_t>(deviceCounter) & 7u) { return fail(RSRCU_ERR_INVALID, "completion counter must be 8-byte aligned"); }
	CU(cudaSetDevice(c->device));
	{ const int r = joinTileStreams(c); if (r != RSRCU_OK) { return r; } }
	CU(launchPdl(signal_counter_kernel, 1u, 32u, 0, c->stream, static_cast<unsigned long long*>(deviceCounter)));
	return RSRCU_OK; }

int rsrcu_wait_counters(rsrcu_ctx* c, const void* deviceCounters, int count, uint64_t value) {
	if (!c || !deviceCounters || count <= 0) { return fail(RSRCU_ERR_INVALID, "bad argument"); }
	CU(cudaSetDevice(c->device));
	if (!c->waitTimedOut) {
		CU(cudaMalloc(&c->waitTimedOut, sizeof(unsigned int)));
		CU(cudaMemset(c->waitTimedOut, 0, sizeof(unsigned int))); }
	CU(launchPdl(wait_counter_kernel, 1u, 32u, 0, c->stream, static_cast<const unsigned long long*>(deviceCounters), static_cast<unsigned int>(count),
	             static_cast<unsigned long long>(value), c->waitTimedOut));
	// the tile kernels of later frames must stay behind this wait on whichever stream they run
	CU(cudaEventRecord(c->evGate, c->stream));
	CU(cudaStreamWaitEvent(c->tileStream2, c->evGate, 0));
	return RSRCU_OK; }

int rsrcu_set_overlap(rsrcu_ctx* c, int enabled) {
	if (!c) { return fail(RSRCU_ERR_INVALID, "null context"); }
	if (c->inFrame) { return fail(RSRCU_ERR_INVALID, "rsrcu_set_overlap inside begin/end frame"); }
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->frontStream));
	CU(cudaStreamSynchronize(c->stream));
	CU(cudaStreamSynchronize(c->tileStream2));
	c->overlap = enabled != 0;
	return RSRCU_OK; }

int rsrcu_set_profiling(rsrcu_ctx* c, int enabled) {
	if (!c) { return fail(RSRCU_ERR_INVALID, "null context"); }
	c->profiling = enabled < 0 ? 0 : (enabled > 2 ? 2 : enabled);
	return RSRCU_OK; }

int rsrcu_get_stage_ms(rsrcu_ctx* c, float* ms7) {
	if (!c || !ms7) { return fail(RSRCU_ERR_INVALID, "null argument"); }
	// evStage: 0 start, 1 after upload, 2 after vertex, 3 after setup, 4 after count, 5 after scan, 6 after fill, 7 after tile
	for (int i = 0; i < 7; ++i) { ms7[i] = c->stageMs[i]; }
	return RSRCU_OK; }

}  // extern "C"
