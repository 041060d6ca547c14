// rglv_cuda.hxx -- C++ host-side mirror of the reference's drawing API over the rsrcu C ABI.
//
// `rglvcu::GL` has the recording surface of `rqdq::rglv::GL` (src/rgl/rglv/rglv_gl.hxx:182-344: same
// method names, argument order and meaning) and `rglvcu::GPU` the surface of `rqdq::rglv::GPU`
// (src/rgl/rglv/rglv_gpu.hxx:152-168: IC(), Reset(), Run()).  GL calls are recorded into the packed
// stream documented in include/rsrcu.h; Run() hands the frame to librsrcu.so with one call.
// Differences a caller sees, all forced by leaving the CPU:
//   * UseBuffer / BindTexture / DrawElements take the element count next to the pointer (the
//     reference scans indices for the extent, rglv_gpu_impl.hxx:332-334) and an upload policy;
//   * matrices are the 16 floats of rmlm::mat4::ff (column-major);
//   * Run() returns an error code instead of a jobsys::Job* -- the CPU job system is replaced by
//     the context's CUDA stream; Sync() waits for the frame and fills the store destinations.
// Header-only; link with -lrsrcu.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rsrcu.h"

namespace rglvcu {

constexpr int GL_TRIANGLES = 0;
constexpr int GL_CULL_FACE = 1, GL_SCISSOR_TEST = 2, GL_BLEND = 3, GL_DEPTH_TEST = 4;
constexpr int GL_FRONT = 1, GL_BACK = 2, GL_FRONT_AND_BACK = 3;
constexpr int GL_UNSIGNED_SHORT = 1;
constexpr int GL_NEAREST_MIPMAP_NEAREST = 0, GL_LINEAR_MIPMAP_NEAREST = 1;
constexpr int GL_COLOR_BUFFER_BIT = 1, GL_DEPTH_BUFFER_BIT = 2, GL_STENCIL_BUFFER_BIT = 4;
constexpr int GL_LESS = 0, GL_LEQUAL = 1, GL_EQUAL = 2;
constexpr int GL_DEPTH_ATTACHMENT = 0, GL_STENCIL_ATTACHMENT = 1, GL_COLOR_ATTACHMENT0 = 2;
constexpr int RB_COLOR_DEPTH = 0, RB_RGBF32 = 1, RB_RGBAF32 = 2, RB_F32 = 3;

struct Error : std::runtime_error {
	int code;
	Error(int c, const char* msg) : std::runtime_error(msg), code(c) {} };

class GL {
public:
	void Enable(int value) { Cap(value, 1); }
	void Disable(int value) { Cap(value, 0); }
	void DepthFunc(int v) { cs_.depth_func = v; dirty_ = true; }
	void DepthWriteMask(bool v) { cs_.depth_write_mask = v; dirty_ = true; }
	void ColorWriteMask(bool v) { cs_.color_write_mask = v; dirty_ = true; }
	void CullFace(int v) { cs_.cull_face = v; dirty_ = true; }
	void Scissor(int x, int y, int w, int h) {
		cs_.scissor_origin[0] = x; cs_.scissor_origin[1] = y; cs_.scissor_size[0] = w; cs_.scissor_size[1] = h; dirty_ = true; }
	void Viewport(int x, int y, int w, int h) {
		cs_.viewport_origin[0] = x; cs_.viewport_origin[1] = y; cs_.viewport_size[0] = w; cs_.viewport_size[1] = h; dirty_ = true; }
	void UseProgram(int v) { cs_.program_id = v; dirty_ = true; }
	void RenderbufferType(int attachment, int type) {
		if (attachment == GL_DEPTH_ATTACHMENT) { cs_.depth_attachment_type = type; }
		else if (attachment == GL_COLOR_ATTACHMENT0) { cs_.color0_attachment_type = type; }
		dirty_ = true; }
	// AllocUniformBuffer + UseUniforms in one step: <= 128 bytes of UniformsSD
	void UseUniforms(const void* data, size_t bytes) {
		if (bytes > sizeof(cs_.uniforms)) { throw Error(RSRCU_ERR_INVALID, "uniform block larger than 128 bytes"); }
		std::memset(cs_.uniforms, 0, sizeof(cs_.uniforms));
		std::memcpy(cs_.uniforms, data, bytes);
		cs_.uniforms_valid = 1; dirty_ = true; }
	void BindTexture(int unit, const float* texels, int width, int height, int stride, int mode,
	                 int rowsInMemory, int upload = RSRCU_UPLOAD_STATIC) {
		const int32_t p[8] = { unit, width, height, stride, mode, rowsInMemory, upload, 0 };
		Emit(RSRCU_OP_BIND_TEXTURE, p, sizeof(p), reinterpret_cast<uint64_t>(texels)); }
	void BindTexture3(const float* depth, int dim, int upload = RSRCU_UPLOAD_ALWAYS) {
		const int32_t p[2] = { dim, upload };
		Emit(RSRCU_OP_BIND_DEPTH, p, sizeof(p), reinterpret_cast<uint64_t>(depth)); }
	void UseBuffer(int idx, const float* ptr, size_t nFloats, int upload = RSRCU_UPLOAD_STATIC) {
		const int32_t p[2] = { idx, upload };
		const uint64_t q[2] = { reinterpret_cast<uint64_t>(ptr), static_cast<uint64_t>(nFloats) };
		Header(RSRCU_OP_BIND_BUFFER, sizeof(p) + sizeof(q));
		Bytes(p, sizeof(p)); Bytes(q, sizeof(q)); }
	void ClearColor(float r, float g, float b) {
		cs_.clear_color[0] = r; cs_.clear_color[1] = g; cs_.clear_color[2] = b; cs_.clear_color[3] = 1.0f; dirty_ = true; }
	void ClearDepth(float v) { cs_.clear_depth = v; dirty_ = true; }
	void ViewMatrix(const float* m16) { std::memcpy(cs_.view_matrix, m16, 64); dirty_ = true; }
	void ProjectionMatrix(const float* m16) { std::memcpy(cs_.projection_matrix, m16, 64); dirty_ = true; }
	void NormalMatrix(const float* m16) { std::memcpy(cs_.normal_matrix, m16, 64); dirty_ = true; }

	void DrawElements(int mode, int count, int type, const uint16_t* indices, uint8_t hint = 0, int upload = RSRCU_UPLOAD_STATIC) {
		(void)mode; (void)type;
		MaybeUpdateState();
		const int32_t p[4] = { count, hint, 0, upload };
		Emit(RSRCU_OP_DRAW_ELEMENTS, p, sizeof(p), reinterpret_cast<uint64_t>(indices)); }
	void DrawArrays(int mode, int start, int count) {
		(void)mode; (void)start;
		MaybeUpdateState();
		const int32_t p[2] = { count, 0 };
		Emit(RSRCU_OP_DRAW_ARRAYS, p, sizeof(p)); }
	void DrawElementsInstanced(int mode, int count, int type, const uint16_t* indices, int instanceCnt, int upload = RSRCU_UPLOAD_STATIC) {
		(void)mode; (void)type;
		MaybeUpdateState();
		const int32_t p[4] = { count, 0, instanceCnt, upload };
		Emit(RSRCU_OP_DRAW_ELEMENTS, p, sizeof(p), reinterpret_cast<uint64_t>(indices)); }
	void DrawArraysInstanced(int mode, int start, int count, int instanceCnt) {
		(void)mode; (void)start;
		MaybeUpdateState();
		const int32_t p[2] = { count, instanceCnt };
		Emit(RSRCU_OP_DRAW_ARRAYS, p, sizeof(p)); }
	void Clear(uint8_t bits) {
		MaybeUpdateState();
		const int32_t p[2] = { bits, 0 };
		Emit(RSRCU_OP_CLEAR, p, sizeof(p)); }
	// StoreColor(TrueColorCanvas*, enableGammaCorrection)
	void StoreColor(uint32_t* dst, int width, int height, int stridePx, bool enableGammaCorrection) {
		MaybeUpdateState();
		const int32_t p[4] = { enableGammaCorrection ? 1 : 0, width, height, stridePx };
		Emit(RSRCU_OP_STORE_TC, p, sizeof(p), reinterpret_cast<uint64_t>(dst)); }
	// StoreColor(FloatingPointCanvas*, downsample)
	void StoreColor(float* dst, int width, int height, int stridePx, bool downsample) {
		MaybeUpdateState();
		const int32_t p[4] = { downsample ? 1 : 0, width, height, stridePx };
		Emit(RSRCU_OP_STORE_FP, p, sizeof(p), reinterpret_cast<uint64_t>(dst)); }
	// StoreColor(QFloat4Canvas*): quad-swizzled RGBA32F, 64 bytes per 2x2 quad, stride in quads
	void StoreColorQuads(float* dst, int width, int height, int strideQuads) {
		MaybeUpdateState();
		const int32_t p[4] = { 0, width, height, strideQuads };
		Emit(RSRCU_OP_STORE_QUADS, p, sizeof(p), reinterpret_cast<uint64_t>(dst)); }
	void StoreDepth(float* dst) {
		MaybeUpdateState();
		Emit(RSRCU_OP_STORE_DEPTH, nullptr, 0, reinterpret_cast<uint64_t>(dst)); }
	// the same stores into DEVICE canvases (GPU::CanvasAlloc): the frame's float / depth results never cross PCIe and
	// can be bound again as texture (BindTexture with RSRCU_UPLOAD_DEVICE), shadow map (BindTexture3) or filtered
	void StoreColorDevice(void* deviceDst, int width, int height, int stridePx, bool downsample) {
		MaybeUpdateState();
		const int32_t p[4] = { downsample ? 1 : 0, width, height, stridePx };
		Emit(RSRCU_OP_STORE_FP_DEV, p, sizeof(p), reinterpret_cast<uint64_t>(deviceDst)); }
	void StoreColorQuadsDevice(void* deviceDst, int width, int height, int strideQuads) {
		MaybeUpdateState();
		const int32_t p[4] = { 0, width, height, strideQuads };
		Emit(RSRCU_OP_STORE_QUADS_DEV, p, sizeof(p), reinterpret_cast<uint64_t>(deviceDst)); }
	void StoreDepthDevice(void* deviceDst) {
		MaybeUpdateState();
		Emit(RSRCU_OP_STORE_DEPTH_DEV, nullptr, 0, reinterpret_cast<uint64_t>(deviceDst)); }
	void Finish() {}
	void Reset() {
		stream_.clear();
		ResetState();
		dirty_ = true; }

	const std::vector<uint8_t>& stream() const { return stream_; }

private:
	friend class GPU;
	void ResetState() {   // GLState::reset, rglv_gl.hxx:152-179
		std::memset(&cs_, 0, sizeof(cs_));
		cs_.clear_color[3] = 1.0f; cs_.clear_depth = 1.0f;
		cs_.cull_face = GL_BACK;
		cs_.color_write_mask = 1; cs_.depth_write_mask = 1; cs_.depth_test_enabled = 1; cs_.depth_func = GL_LESS;
		cs_.color0_attachment_type = RB_COLOR_DEPTH; cs_.depth_attachment_type = RB_COLOR_DEPTH;
		for (int i = 0; i < 4; ++i) { cs_.view_matrix[i * 5] = cs_.projection_matrix[i * 5] = cs_.normal_matrix[i * 5] = 1.0f; } }
	void Cap(int value, int on) {
		if (value == GL_CULL_FACE) { cs_.culling_enabled = on; }
		else if (value == GL_SCISSOR_TEST) { cs_.scissor_enabled = on; }
		else if (value == GL_BLEND) { cs_.blending_enabled = on; }
		else if (value == GL_DEPTH_TEST) { cs_.depth_test_enabled = on; }
		else { throw std::runtime_error("unknown glEnable value"); }
		dirty_ = true; }
	void MaybeUpdateState() {   // rglv_gl.cxx:100-106
		if (!dirty_) { return; }
		Header(RSRCU_OP_STATE, sizeof(cs_));
		Bytes(&cs_, sizeof(cs_));
		Pad();
		dirty_ = false; }
	void Header(uint32_t op, size_t payload) {
		const uint32_t size = static_cast<uint32_t>((8 + payload + 7) & ~static_cast<size_t>(7));
		Bytes(&op, 4); Bytes(&size, 4); }
	void Bytes(const void* p, size_t n) {
		const auto* b = static_cast<const uint8_t*>(p);
		stream_.insert(stream_.end(), b, b + n); }
	void Pad() { while (stream_.size() & 7) { stream_.push_back(0); } }
	void Emit(uint32_t op, const void* ints, size_t intBytes, uint64_t ptr) {
		Header(op, intBytes + 8);
		if (intBytes) { Bytes(ints, intBytes); }
		Bytes(&ptr, 8);
		Pad(); }
	void Emit(uint32_t op, const void* ints, size_t intBytes) {
		Header(op, intBytes);
		Bytes(ints, intBytes);
		Pad(); }

	std::vector<uint8_t> stream_;
	RsrState cs_{};
	bool dirty_{true}; };


class GPU {
public:
	explicit GPU(int device = 0) {
		if (int rc = rsrcu_create(device, &ctx_); rc != RSRCU_OK) { throw Error(rc, rsrcu_last_error()); } }
	~GPU() { rsrcu_destroy(ctx_); }
	GPU(const GPU&) = delete;
	GPU& operator=(const GPU&) = delete;

	auto IC() -> GL& { return ic_; }
	static void Check(int rc) { if (rc != RSRCU_OK) { throw Error(rc, rsrcu_last_error()); } }

	// GPU::Reset (rglv_gpu.cxx:50-56)
	void Reset(int widthPx, int heightPx, int tileBlocksX = 8, int tileBlocksY = 8) {
		ic_.Reset();
		const int32_t p[4] = { widthPx, heightPx, tileBlocksX, tileBlocksY };
		ic_.Emit(RSRCU_OP_BEGIN_FRAME, p, sizeof(p)); }

	// GPU::Run: decode + upload + kernels are enqueued; returns before the frame is finished
	void Run() {
		ic_.Emit(RSRCU_OP_END_FRAME, nullptr, 0);
		if (int rc = rsrcu_run_stream(ctx_, ic_.stream_.data(), ic_.stream_.size()); rc != RSRCU_OK) {
			throw Error(rc, rsrcu_last_error()); } }

	// what waiting on the reference's Finalize job does
	void Sync() { if (int rc = rsrcu_sync(ctx_); rc != RSRCU_OK) { throw Error(rc, rsrcu_last_error()); } }

	// beyond the reference's surface (include/rsrcu.h): pipelining, retained frames, split-frame presentation
	void SyncFrame(int lag) { Check(rsrcu_sync_frame(ctx_, lag)); }
	void SetOverlap(bool on) { Check(rsrcu_set_overlap(ctx_, on ? 1 : 0)); }
	/// orders Stream() behind every frame submitted so far (overlap mode runs tile kernels on two streams)
	void Join() { Check(rsrcu_join(ctx_)); }
	void EnablePeerAccess(int peerDevice) { Check(rsrcu_enable_peer_access(ctx_, peerDevice)); }
	rsrcu_frame* Retain() { rsrcu_frame* f = nullptr; Check(rsrcu_retain_frame(ctx_, &f)); return f; }
	void Replay(rsrcu_frame* f) { Check(rsrcu_replay_frame(ctx_, f)); }
	void Release(rsrcu_frame* f) { Check(rsrcu_release_frame(ctx_, f)); }

	// device canvases and the nodes next to the rasteriser (include/rsrcu.h): `$kawase`, `$glow`,
	// `$renderToTexture`'s mip chain, `$mc`, the shadow pass of a second context, the telemetry bars, presentation
	void* CanvasAlloc(size_t bytes) { void* p = nullptr; Check(rsrcu_canvas_alloc(ctx_, bytes, &p)); return p; }
	void CanvasFree(void* p) { Check(rsrcu_canvas_free(ctx_, p)); }
	void CanvasRead(const void* devicePtr, void* hostDst, size_t bytes) { Check(rsrcu_canvas_read(ctx_, devicePtr, hostDst, bytes)); }
	void CanvasWrite(void* devicePtr, const void* hostSrc, size_t bytes) { Check(rsrcu_canvas_write(ctx_, devicePtr, hostSrc, bytes)); }
	void WaitFor(GPU& producer) { Check(rsrcu_wait_for(ctx_, producer.ctx_)); }
	void KawaseBlur(const void* src, int srcStridePx, void* dst, int dstStridePx, int width, int height, int dist) {
		Check(rsrcu_kawase_blur(ctx_, src, srcStridePx, dst, dstStridePx, width, height, dist)); }
	void Glow(const void* imageQuads, int imageStrideQuads, const void* blur, int blurStridePx, bool sRGB, uint32_t* dst, bool dstIsDevice,
	          int width, int height, int stridePx) {
		Check(rsrcu_glow(ctx_, imageQuads, imageStrideQuads, blur, blurStridePx, sRGB ? 1 : 0, dst, dstIsDevice ? 1 : 0, width, height, stridePx)); }
	void MakeMipmap(void* texelsDevice, int dim) { Check(rsrcu_make_mipmap(ctx_, texelsDevice, dim)); }
	struct Surface { const float* soa[6]; std::vector<RsrMarchBlock> blocks; int vertexTotal; };
	Surface MarchSurface(float timeSeconds, int precision, int forkDepth, float range) {
		Surface s{}; s.blocks.resize(4096);
		int n = 0;
		Check(rsrcu_march_surface(ctx_, timeSeconds, precision, forkDepth, range, s.soa, s.blocks.data(), 4096, &n, &s.vertexTotal));
		s.blocks.resize(static_cast<size_t>(n));
		return s; }
	void DrawSpans(void* truecolorDevice, int stridePx, int width, int height, int left, int top, float xscale, const std::vector<RsrSpan>& spans) {
		Check(rsrcu_draw_spans(ctx_, truecolorDevice, stridePx, width, height, left, top, xscale, spans.data(), static_cast<int>(spans.size()))); }
	std::vector<RsrSpan> FrameSpans() {
		std::vector<RsrSpan> spans(8); int n = 0;
		Check(rsrcu_frame_spans(ctx_, spans.data(), 8, &n));
		spans.resize(static_cast<size_t>(n));
		return spans; }
	void Present(const void* truecolorDevice, int srcStridePx, void* surfaceDevice, int surfaceStridePx, int width, int height) {
		Check(rsrcu_present(ctx_, truecolorDevice, srcStridePx, surfaceDevice, surfaceStridePx, width, height)); }

	rsrcu_ctx* context() { return ctx_; }

private:
	rsrcu_ctx* ctx_{nullptr};
	GL ic_; };

}  // namespace rglvcu
