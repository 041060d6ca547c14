// rglv_gpu_cuda.cxx -- the reference-side binding: `rqdq::rglv::GPU::RunImpl` on top of include/rsrcu.h.
//
// This is the file a maintainer of usrlocalben/rsr adds to src/rgl/rglv/ (and links with -lrsrcu) to render on a
// B200: it REPLACES the body of GPU::RunImpl (src/rgl/rglv/rglv_gpu.cxx:90-116).  Everything above that function --
// the node graph, `rglv::GL` recording (rglv_gl.hxx:182-344, rglv_gl.cxx), `GLState`, the packed command stream
// (rglv_packed_stream.hxx), program ids, `rqv::Install` -- is the reference's own, unmodified code; this function
// walks the stream `GL` recorded, exactly like `GPU::BinImpl` does (rglv_gpu.cxx:119-260), and forwards every
// command to the C ABI.  The CUDA context takes the place of the bin + tile jobs; the `Finalize` job and the
// IC double-buffer swap (rglv_gpu.hxx:176-192) stay.
//
// It is compiled for real by oracle/build_ref.sh against the reference tree (into oracle/_ref/librsr_dropin.so:
// the reference's translation units, with RunImpl's body in rglv_gpu.cxx compiled out, plus this file) and
// tests/test_dropin_gpu.py drives the reference's own `rglv::GL` through it and compares with the pure reference.
//
// Two things the reference API leaves implicit have to be found here:
//  * buffer extents -- `GL::UseBuffer` takes raw pointers (rglv_gl.hxx:104).  Like the reference's binner
//    (rglv_gpu_impl.hxx:332-334) the extent of a draw's vertex arrays is the largest index it uses + 1
//    (DrawArrays: the count; slot 15: 16 floats per instance), found by scanning the index buffer;
//  * texture rows -- `TextureState` has no mip flag, but the sampler `MakeTextureUnit` picks for a power-of-two
//    square texture (rglr_texture_sampler.cxx:314-363) reads the mip chain stacked below the base level
//    (rglr_texture.cxx:33-81), so such a texture is 2 x height rows in memory.
// The class has no member for the CUDA context (the reference's headers are not edited): a side table keyed by
// the GPU object holds it.  A maintainer would add `rsrcu_ctx* cuda_` to the class instead.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <mutex>
#include <unordered_map>

#include "src/rcl/rclmt/rclmt_jobsys.hxx"
#include "src/rgl/rglr/rglr_canvas.hxx"
#include "src/rgl/rglv/rglv_gl.hxx"
#include "src/rgl/rglv/rglv_gpu.hxx"
#include "src/rgl/rglv/rglv_gpu_protocol.hxx"

#include "rsrcu.h"

namespace rqdq {
namespace rglv {

namespace {

struct CudaSide {
	rsrcu_ctx* ctx{nullptr};
	bool frameInFlight{false};
	std::unordered_map<uint64_t, int> extentCache; };   // (index pointer, count) -> vertex extent, when indices are static

std::mutex g_sideMutex;
std::unordered_map<const GPU*, CudaSide> g_side;
int g_device = 0;
// what may be assumed about the memory behind the pointers GL recorded: nothing by default (copied every frame);
// a host that knows its meshes / textures are immutable says so (rsr_dropin_set_upload_policy)
// (RSRCU_UPLOAD_FRAME: the reference reads every pointer at Run, so one staging per frame and pointer is exact)
int g_bufferPolicy = RSRCU_UPLOAD_FRAME, g_texturePolicy = RSRCU_UPLOAD_FRAME, g_indexPolicy = RSRCU_UPLOAD_FRAME;

// Canvases that live on the device (rsr_b200/host/post_nodes_cuda.cxx: the `$buffers` / `$kawase` / `$glow` chain).
// A device canvas is an ordinary rglr canvas object around memory from CudaCanvasAlloc; GL::StoreColor(&canvas) is
// recorded as always, and the store commands below turn into the _device variants when the canvas's memory is in
// this table.  `writer` is the context whose stream wrote the canvas last: a filter that reads it runs there.
struct DeviceRange { size_t bytes; rsrcu_ctx* writer; };
std::map<uintptr_t, DeviceRange> g_deviceCanvases;   // by start address (g_sideMutex)
rsrcu_ctx* g_utilityCtx = nullptr;                    // owns the canvas memory (any context of the device may use it)

DeviceRange* FindDeviceRange(const void* p) {
	const auto a = reinterpret_cast<uintptr_t>(p);
	auto it = g_deviceCanvases.upper_bound(a);
	if (it == g_deviceCanvases.begin()) { return nullptr; }
	--it;
	return a < it->first + it->second.bytes ? &it->second : nullptr; }

bool IsDeviceCanvas(const void* p, rsrcu_ctx* writer) {
	std::lock_guard<std::mutex> lock(g_sideMutex);
	DeviceRange* r = FindDeviceRange(p);
	if (r != nullptr) { r->writer = writer; }
	return r != nullptr; }

[[noreturn]] void Die(const char* what) {
	// same policy as a missing dispatch entry in the reference (rglv_gpu.cxx:199-202)
	std::cerr << "rsrcu: " << what << ": " << rsrcu_last_error() << "\n";
	std::exit(1); }

#define RSRCU_DO(call) do { if ((call) != RSRCU_OK) { Die(#call); } } while (0)

CudaSide& SideOf(const GPU* gpu) {
	std::lock_guard<std::mutex> lock(g_sideMutex);
	CudaSide& s = g_side[gpu];
	if (s.ctx == nullptr) {
		if (const char* dev = std::getenv("RSRCU_DEVICE")) { g_device = std::atoi(dev); }
		RSRCU_DO(rsrcu_create(g_device, &s.ctx)); }
	return s; }

RsrState ToRsrState(const GLState& s, const void* uniforms) {
	RsrState o{};
	o.clear_color[0] = s.clearColor.x; o.clear_color[1] = s.clearColor.y; o.clear_color[2] = s.clearColor.z; o.clear_color[3] = s.clearColor.w;
	o.clear_depth = s.clearDepth;
	o.culling_enabled = s.cullingEnabled ? 1 : 0;
	o.cull_face = s.cullFace;
	o.scissor_enabled = s.scissorEnabled ? 1 : 0;
	o.scissor_origin[0] = s.scissorOrigin.x; o.scissor_origin[1] = s.scissorOrigin.y;
	o.scissor_size[0] = s.scissorSize.x;     o.scissor_size[1] = s.scissorSize.y;
	o.viewport_origin[0] = s.viewportOrigin.x; o.viewport_origin[1] = s.viewportOrigin.y;
	if (s.viewportSize.has_value()) { o.viewport_size[0] = s.viewportSize->x; o.viewport_size[1] = s.viewportSize->y; }
	o.blending_enabled = s.blendingEnabled ? 1 : 0;
	o.color_write_mask = s.colorWriteMask ? 1 : 0;
	o.depth_write_mask = s.depthWriteMask ? 1 : 0;
	o.depth_test_enabled = s.depthTestEnabled ? 1 : 0;
	o.depth_func = s.depthFunc;
	o.program_id = s.programId;
	o.color0_attachment_type = s.color0AttachmentType;
	o.depth_attachment_type = s.depthAttachmentType;
	std::memcpy(o.view_matrix, s.viewMatrix.ff.data(), sizeof(o.view_matrix));
	std::memcpy(o.projection_matrix, s.projectionMatrix.ff.data(), sizeof(o.projection_matrix));
	std::memcpy(o.normal_matrix, s.normalMatrix.ff.data(), sizeof(o.normal_matrix));
	if (uniforms != nullptr) {
		o.uniforms_valid = 1;
		std::memcpy(o.uniforms, uniforms, sizeof(float) * UNIFORM_BUFFER_SIZE); }
	return o; }

bool IsPow2Square(const TextureState& tu) {
	return tu.width == tu.height && tu.stride == tu.width && tu.width > 0 && (tu.width & (tu.width - 1)) == 0; }

// textures and the depth texture are part of the state snapshot
void BindTextures(rsrcu_ctx* ctx, const GLState& st) {
	for (int u = 0; u < 2; ++u) {
		const TextureState& tu = st.tus[u];
		if (tu.ptr == nullptr) { continue; }
		const int rows = IsPow2Square(tu) ? 2 * tu.height : tu.height;
		RSRCU_DO(rsrcu_bind_texture(ctx, u, reinterpret_cast<const float*>(tu.ptr), tu.width, tu.height, tu.stride, tu.filter, rows, g_texturePolicy)); }
	if (st.tu3ptr != nullptr) {
		RSRCU_DO(rsrcu_bind_depth_texture(ctx, st.tu3ptr, st.tu3dim, g_texturePolicy)); } }

// The buffer slots a program's Loader dereferences (src/viewer/shaders.hxx, shaders_envmap.hxx, rglv_gpu_shaders.hxx:
// LoadMD / LoadLane).  GLState::buffers keeps whatever an earlier draw bound to the other slots -- arrays that may be
// shorter than this draw or gone already (the `$writer` node's text after a mesh) -- and the reference never reads
// them; neither may the upload.
unsigned SlotsOf(int programId) {
	constexpr unsigned kPos = 0x007u, kNrm = 0x038u, kKd = 0x1c0u, kUv = 0x600u;
	switch (programId) {
	case 4: case 5: case 6: return kPos | kNrm | kUv;          // Amy, Depth, Many (+ slot 15, bound separately)
	case 65: return kPos | kUv;                                  // AlphaTexture
	case 26: return kPos | kKd | kUv;                            // Text
	case 7: case 8: case 9: case 10: return kPos | kNrm | kKd;  // OBJ1, OBJ2, OBJ2S, Envmap
	default: return kPos; } }                                    // BaseProgram, Pattern, Wireframe

// vertex arrays: bound per draw, once the draw's extent is known
void BindBuffers(rsrcu_ctx* ctx, const GLState& st, int nverts, int instances) {
	const unsigned used = SlotsOf(st.programId);
	for (int slot = 0; slot <= 10; ++slot) {
		const float* p = ((used >> slot) & 1u) ? st.buffers[slot] : nullptr;
		RSRCU_DO(rsrcu_bind_buffer(ctx, slot, p, p != nullptr ? static_cast<size_t>(nverts) : 0, g_bufferPolicy)); }
	const float* mats = st.buffers[15];
	RSRCU_DO(rsrcu_bind_buffer(ctx, 15, instances > 0 ? mats : nullptr, (instances > 0 && mats != nullptr) ? static_cast<size_t>(instances) * 16 : 0, g_bufferPolicy)); }

int VertexExtent(CudaSide& side, const uint16_t* indices, int count) {
	const uint64_t key = (reinterpret_cast<uintptr_t>(indices) * 0x9E3779B97F4A7C15ull) ^ static_cast<uint64_t>(count);
	if (g_indexPolicy == RSRCU_UPLOAD_STATIC) {
		if (auto it = side.extentCache.find(key); it != side.extentCache.end()) { return it->second; } }
	uint16_t mx = 0;
	for (int i = 0; i < count; ++i) { mx = indices[i] > mx ? indices[i] : mx; }
	const int extent = static_cast<int>(mx) + 1;
	if (g_indexPolicy == RSRCU_UPLOAD_STATIC) { side.extentCache[key] = extent; }
	return extent; }

// the stream decode of GPU::BinImpl (rglv_gpu.cxx:119-243), one C-ABI call per command
void SubmitFrame(CudaSide& side, GL& gl, rmlv::ivec2 sizeInPixels, rmlv::ivec2 tileInBlocks) {
	rsrcu_ctx* ctx = side.ctx;
	auto& cs = gl.commands_;
	cs.appendByte(CMD_EOF);
	RSRCU_DO(rsrcu_begin_frame(ctx, sizeInPixels.x, sizeInPixels.y, tileInBlocks.x, tileInBlocks.y));
	const GLState* st = nullptr;
	bool done = false;
	while (!done) {
		const auto cmd = cs.consumeByte();
		switch (cmd) {
		case CMD_EOF:
			done = true;
			break;
		case CMD_STATE: {
			st = static_cast<const GLState*>(cs.consumePtr());
			const RsrState rs = ToRsrState(*st, gl.GetUniformBufferAddr(st->uniformsOfs));
			RSRCU_DO(rsrcu_set_state(ctx, &rs));
			BindTextures(ctx, *st); }
			break;
		case CMD_CLEAR:
			RSRCU_DO(rsrcu_clear(ctx, cs.consumeByte()));
			break;
		case CMD_STORE_COLOR_HALF_LINEAR_FP: {
			auto* c = static_cast<rglr::FloatingPointCanvas*>(cs.consumePtr());
			if (IsDeviceCanvas(c->data(), ctx)) { RSRCU_DO(rsrcu_store_color_fp_device(ctx, c->data(), c->width(), c->height(), c->stride(), 1)); }
			else { RSRCU_DO(rsrcu_store_color_fp(ctx, reinterpret_cast<float*>(c->data()), c->width(), c->height(), c->stride(), 1)); } }
			break;
		case CMD_STORE_COLOR_FULL_LINEAR_FP: {
			auto* c = static_cast<rglr::FloatingPointCanvas*>(cs.consumePtr());
			if (IsDeviceCanvas(c->data(), ctx)) { RSRCU_DO(rsrcu_store_color_fp_device(ctx, c->data(), c->width(), c->height(), c->stride(), 0)); }
			else { RSRCU_DO(rsrcu_store_color_fp(ctx, reinterpret_cast<float*>(c->data()), c->width(), c->height(), c->stride(), 0)); } }
			break;
		case CMD_STORE_COLOR_FULL_QUADS_FP: {
			auto* c = static_cast<rglr::QFloat4Canvas*>(cs.consumePtr());
			if (IsDeviceCanvas(c->data(), ctx)) { RSRCU_DO(rsrcu_store_color_quads_device(ctx, c->data(), c->width(), c->height(), c->stride())); }
			else { RSRCU_DO(rsrcu_store_color_quads(ctx, reinterpret_cast<float*>(c->data()), c->width(), c->height(), c->stride())); } }
			break;
		case CMD_STORE_COLOR_FULL_LINEAR_TC: {
			const auto enableGamma = cs.consumeByte();
			auto* c = static_cast<rglr::TrueColorCanvas*>(cs.consumePtr());
			RSRCU_DO(rsrcu_store_color_tc(ctx, enableGamma, reinterpret_cast<uint32_t*>(c->data()), c->width(), c->height(), c->stride())); }
			break;
		case CMD_STORE_DEPTH_FULL_LINEAR_FP:
			RSRCU_DO(rsrcu_store_depth(ctx, static_cast<float*>(cs.consumePtr())));
			break;
		case CMD_DRAW_ARRAYS: {
			const auto count = cs.consumeInt();
			BindBuffers(ctx, *st, count, 0);
			RSRCU_DO(rsrcu_draw_arrays(ctx, count, 0)); }
			break;
		case CMD_DRAW_ARRAYS_INSTANCED: {
			const auto count = cs.consumeInt();
			const auto instanceCnt = cs.consumeInt();
			BindBuffers(ctx, *st, count, instanceCnt);
			RSRCU_DO(rsrcu_draw_arrays(ctx, count, instanceCnt)); }
			break;
		case CMD_DRAW_ELEMENTS: {
			cs.consumeByte();   // 0x14: 16-bit indices, triangles
			const auto hint = cs.consumeByte();
			const auto count = cs.consumeInt();
			const auto* indices = static_cast<const uint16_t*>(cs.consumePtr());
			BindBuffers(ctx, *st, VertexExtent(side, indices, count), 0);
			RSRCU_DO(rsrcu_draw_elements(ctx, count, indices, hint, 0, g_indexPolicy)); }
			break;
		case CMD_DRAW_ELEMENTS_INSTANCED: {
			cs.consumeByte();
			const auto count = cs.consumeInt();
			const auto* indices = static_cast<const uint16_t*>(cs.consumePtr());
			const auto instanceCnt = cs.consumeInt();
			BindBuffers(ctx, *st, VertexExtent(side, indices, count), instanceCnt);
			RSRCU_DO(rsrcu_draw_elements(ctx, count, indices, 0, instanceCnt, g_indexPolicy)); }
			break;
		default:
			std::cerr << "rsrcu: unknown command " << static_cast<int>(cmd) << " in the GL stream\n";
			std::exit(1); } }
	RSRCU_DO(rsrcu_end_frame(ctx)); }

}  // namespace


void GPU::RunImpl(rclmt::jobsys::Job* job) {
	namespace jobsys = rclmt::jobsys;
	auto finalizeJob = Finalize();
	if (job != nullptr) {
		jobsys::move_links(job, finalizeJob); }

	CudaSide& side = SideOf(this);
	SubmitFrame(side, IC(), bufferDimensionsInPixels_, tileDimensionsInBlocks_);
	if (!doubleBuffer) {
		// the frame is complete (store destinations written) when Run's links fire, as in the reference
		RSRCU_DO(rsrcu_sync(side.ctx));
		SwapBuffers(); }
	else {
		// doubleBuffer (rglv_gpu.cxx:16,111-112): the reference bins this frame while it draws the previous one, whose
		// canvases are complete when this Run ends.  Same contract here: this frame is in flight on the GPU, the
		// previous one has landed; Finalize swaps the contexts.
		if (side.frameInFlight) { RSRCU_DO(rsrcu_sync_frame(side.ctx, 1)); }
		side.frameInFlight = true; }
	jobsys::run(finalizeJob); }


// hooks for the host program (not part of the reference's API)
void CudaRelease(const GPU* gpu) {
	std::lock_guard<std::mutex> lock(g_sideMutex);
	if (auto it = g_side.find(gpu); it != g_side.end()) {
		if (it->second.ctx != nullptr) { rsrcu_destroy(it->second.ctx); }
		g_side.erase(it); } }

void CudaFlush(const GPU* gpu) {
	std::lock_guard<std::mutex> lock(g_sideMutex);
	if (auto it = g_side.find(gpu); it != g_side.end() && it->second.ctx != nullptr) {
		RSRCU_DO(rsrcu_sync(it->second.ctx));
		it->second.frameInFlight = false; } }

// ---- device canvases (see g_deviceCanvases) ---------------------------------------------------------------------
void* CudaCanvasAlloc(size_t bytes) {
	std::lock_guard<std::mutex> lock(g_sideMutex);
	if (g_utilityCtx == nullptr) {
		if (const char* dev = std::getenv("RSRCU_DEVICE")) { g_device = std::atoi(dev); }
		RSRCU_DO(rsrcu_create(g_device, &g_utilityCtx)); }
	void* p = nullptr;
	RSRCU_DO(rsrcu_canvas_alloc(g_utilityCtx, bytes, &p));
	g_deviceCanvases[reinterpret_cast<uintptr_t>(p)] = DeviceRange{bytes, nullptr};
	return p; }

void CudaCanvasFree(void* p) {
	std::lock_guard<std::mutex> lock(g_sideMutex);
	if (p != nullptr && g_deviceCanvases.erase(reinterpret_cast<uintptr_t>(p)) != 0) { RSRCU_DO(rsrcu_canvas_free(g_utilityCtx, p)); } }

// the context whose stream wrote the device canvas last (a store command of its frame, or a filter run on it)
rsrcu_ctx* CudaWriterOf(const void* p) {
	std::lock_guard<std::mutex> lock(g_sideMutex);
	DeviceRange* r = FindDeviceRange(p);
	return r != nullptr ? r->writer : nullptr; }

void CudaNoteWriter(const void* p, rsrcu_ctx* ctx) { IsDeviceCanvas(p, ctx); }

void CudaSetUploadPolicy(int buffers, int textures, int indices) {
	g_bufferPolicy = buffers; g_texturePolicy = textures; g_indexPolicy = indices; }


}  // namespace rglv
}  // namespace rqdq
