/* TEST INFRASTRUCTURE -- C entry points around the UNMODIFIED reference renderer.
 *
 * This file is compiled together with the reference's own sources (read in place from
 * /root/reference, never copied into this repo) into oracle/_ref/librsr_ref.so by
 * oracle/build_ref.sh.  It only forwards calls to the reference's public API:
 *
 *   rglv::GL   recording calls            src/rgl/rglv/rglv_gl.hxx:182-344
 *   rglv::GPU  Reset / IC / Run           src/rgl/rglv/rglv_gpu.hxx:152-168
 *   rqv::Install (program dispatch table) src/viewer/shaders.cxx:54-126
 *   jobsys     init / run / wait          src/rcl/rclmt/rclmt_jobsys.hxx:46-88
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load the resulting library.  The product (rsr_b200/) never does.
 */
#include <chrono>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <vector>

#include "src/rcl/rclmt/rclmt_jobsys.hxx"
#include "src/rgl/rglr/rglr_canvas.hxx"
#include "src/rgl/rglr/rglr_texture.hxx"
#include "src/rgl/rglv/rglv_gl.hxx"
#include "src/rgl/rglv/rglv_gpu.hxx"
#include "src/rgl/rglv/rglv_math.hxx"
#include "src/rgl/rglv/rglv_mesh.hxx"
#include "src/rgl/rglv/rglv_mesh_util.hxx"
#include "src/rgl/rglv/rglv_obj.hxx"
#include "src/rgl/rglv/rglv_vao.hxx"
#include "src/rgl/rglv/rglv_triangle.hxx"
#include "src/rml/rmlm/rmlm_mat4.hxx"
#include "src/rml/rmlv/rmlv_mvec4.hxx"
#include "src/viewer/shaders.hxx"
#include "src/rgl/rglr/rglr_algorithm.hxx"
#include "src/rgl/rglr/rglr_canvas_util.hxx"
#include "src/rgl/rglr/rglr_kawase.hxx"
#include "src/rgl/rglv/rglv_marching_cubes.hxx"
#include "src/rgl/rglr/rglr_fragmentcursor.hxx"
#include "src/rgl/rglv/rglv_gpu_impl.hxx"
#include "src/rgl/rglv/rglv_gpu_shaders.hxx"
#include "src/viewer/jobsys_vis.hxx"

using namespace rqdq;
namespace jobsys = rclmt::jobsys;

#ifdef RSR_CUDA_RUNIMPL
/* the drop-in build (librsr_dropin.so): GPU::RunImpl comes from rsr_b200/host/rglv_gpu_cuda.cxx */
namespace rqdq { namespace rglv {
void CudaRelease(const GPU* gpu);
void CudaFlush(const GPU* gpu);
void CudaSetUploadPolicy(int buffers, int textures, int indices);
}}
#endif

namespace {

bool g_jobsysReady = false;
bool g_working = false;

struct RefGPU {
	rglv::GPU gpu;
	std::deque<rglr::TrueColorCanvas> tcCanvases;
	std::deque<rglr::FloatingPointCanvas> fpCanvases;
	std::deque<rglr::QFloat4Canvas> qfCanvases;
	explicit RefGPU(int threads) : gpu(threads, "oracle") {} };

inline rglv::GL& IC(void* h) { return static_cast<RefGPU*>(h)->gpu.IC(); }

rmlm::mat4 ToMat4(const float* m16) {
	std::array<float, 16> a;
	std::memcpy(a.data(), m16, sizeof(float) * 16);
	return rmlm::mat4{a}; }

/* coverage-only fragment processor for the fill-rule known-answer test
 * (same role as TestTargetProgram in rglv_triangle.t.cxx:35-131) */
struct CoverageProgram {
	uint8_t* out;
	int w;
	int x0{0}, y0{0}, x{0}, y{0};
	void Begin(int bx, int by) { x0 = x = bx; y0 = y = by; }
	void CR() { y += 2; x = x0; }
	void Right2() { x += 2; }
	void Render(const rmlv::qfloat2, const rmlv::mvec4i triMask, rglv::BaryCoord) {
		/* lane order: 0=(x,y) 1=(x+1,y) 2=(x,y+1) 3=(x+1,y+1); mask -1 == NOT covered */
		for (int li = 0; li < 4; ++li) {
			if (triMask.si[li] == 0) {
				out[(y + (li >> 1)) * w + x + (li & 1)] = 1; }}} };

}  // namespace

extern "C" {

/* ---- process-wide job system -------------------------------------------------------- */

int ref_init(int threads) {
	if (!g_jobsysReady) {
		jobsys::telemetryEnabled = false;
		jobsys::init(threads);
		g_jobsysReady = true; }
	return jobsys::numThreads; }

void ref_work_start() {
	if (!g_working) { jobsys::work_start(); g_working = true; } }

void ref_work_end() {
	if (g_working) { jobsys::work_end(); g_working = false; } }

/* stop and join the worker threads (call before process exit; std::thread dtors would abort) */
void ref_shutdown() {
	if (!g_jobsysReady) { return; }
	if (g_working) { jobsys::work_end(); g_working = false; }
	jobsys::stop();
	jobsys::join();
	g_jobsysReady = false; }

void ref_set_double_buffer(int enabled) { rglv::doubleBuffer = (enabled != 0); }

/* ---- gpu object --------------------------------------------------------------------- */

void* ref_gpu_create() {
	auto* h = new RefGPU(jobsys::numThreads);
	rqv::Install(h->gpu);
	return h; }

void ref_gpu_destroy(void* h) {
#ifdef RSR_CUDA_RUNIMPL
	rglv::CudaRelease(&static_cast<RefGPU*>(h)->gpu);
#endif
	delete static_cast<RefGPU*>(h); }

/* 1 in librsr_dropin.so (GPU::Run renders through librsrcu.so), 0 in the pure reference */
int ref_is_dropin() {
#ifdef RSR_CUDA_RUNIMPL
	return 1;
#else
	return 0;
#endif
}

#ifdef RSR_CUDA_RUNIMPL
/* drop-in only: wait for the frame still in flight in doubleBuffer mode; upload policy of the binding */
void ref_dropin_flush(void* h) { rglv::CudaFlush(&static_cast<RefGPU*>(h)->gpu); }
void ref_dropin_set_upload_policy(int buffers, int textures, int indices) { rglv::CudaSetUploadPolicy(buffers, textures, indices); }
#endif

void ref_gpu_reset(void* h, int w, int hgt, int tileBlocksX, int tileBlocksY) {
	auto* g = static_cast<RefGPU*>(h);
	g->gpu.Reset(rmlv::ivec2{w, hgt}, rmlv::ivec2{tileBlocksX, tileBlocksY});
	g->tcCanvases.clear();
	g->fpCanvases.clear();
	g->qfCanvases.clear(); }

/* Run one frame to completion.  Caller brackets with ref_work_start/ref_work_end. */
void ref_gpu_run(void* h) {
	auto* g = static_cast<RefGPU*>(h);
	jobsys::reset();
	auto* done = jobsys::make_job(jobsys::noop);
	auto* job = g->gpu.Run();
	jobsys::add_link(job, done);  // RunImpl moves links to its Finalize job (rglv_gpu.cxx:92-94)
	jobsys::run(job);
	jobsys::wait(done); }

/* ---- GL recording (rglv_gl.hxx) ------------------------------------------------------ */

void ref_gl_enable(void* h, int cap) { IC(h).Enable(cap); }
void ref_gl_disable(void* h, int cap) { IC(h).Disable(cap); }
void ref_gl_depth_func(void* h, int v) { IC(h).DepthFunc(v); }
void ref_gl_depth_write_mask(void* h, int v) { IC(h).DepthWriteMask(v != 0); }
void ref_gl_color_write_mask(void* h, int v) { IC(h).ColorWriteMask(v != 0); }
void ref_gl_cull_face(void* h, int v) { IC(h).CullFace(v); }
void ref_gl_scissor(void* h, int x, int y, int w, int hgt) { IC(h).Scissor(x, y, w, hgt); }
void ref_gl_viewport(void* h, int x, int y, int w, int hgt) { IC(h).Viewport(x, y, w, hgt); }
void ref_gl_use_program(void* h, int id) { IC(h).UseProgram(id); }
void ref_gl_renderbuffer_type(void* h, int attachment, int type) { IC(h).RenderbufferType(attachment, type); }
void ref_gl_clear_color(void* h, float r, float g, float b) { IC(h).ClearColor(rmlv::vec3{r, g, b}); }
void ref_gl_clear_depth(void* h, float d) { IC(h).ClearDepth(d); }
void ref_gl_view_matrix(void* h, const float* m) { IC(h).ViewMatrix(ToMat4(m)); }
void ref_gl_projection_matrix(void* h, const float* m) { IC(h).ProjectionMatrix(ToMat4(m)); }
void ref_gl_normal_matrix(void* h, const float* m) { IC(h).NormalMatrix(ToMat4(m)); }
void ref_gl_use_buffer(void* h, int slot, const float* ptr) { IC(h).UseBuffer(slot, ptr); }

void ref_gl_uniforms(void* h, const void* data, int nbytes) {
	auto [id, ptr] = IC(h).AllocUniformBuffer();
	std::memset(ptr, 0, sizeof(float) * rglv::UNIFORM_BUFFER_SIZE);
	std::memcpy(ptr, data, static_cast<size_t>(nbytes));
	IC(h).UseUniforms(id); }

void ref_gl_bind_texture(void* h, int unit, const float* texels, int w, int hgt, int stride, int mode) {
	IC(h).BindTexture(unit, reinterpret_cast<const PixelToaster::FloatingPointPixel*>(texels), w, hgt, stride, mode); }

void ref_gl_bind_texture3(void* h, const float* depth, int dim) { IC(h).BindTexture3(depth, dim); }

void ref_gl_clear(void* h, int bits) { IC(h).Clear(static_cast<uint8_t>(bits)); }

void ref_gl_draw_elements(void* h, int count, const uint16_t* indices, int hint) {
	IC(h).DrawElements(rglv::GL_TRIANGLES, count, rglv::GL_UNSIGNED_SHORT, indices, static_cast<uint8_t>(hint)); }

void ref_gl_draw_arrays(void* h, int count) {
	IC(h).DrawArrays(rglv::GL_TRIANGLES, 0, count); }

void ref_gl_draw_elements_instanced(void* h, int count, const uint16_t* indices, int instanceCnt) {
	IC(h).DrawElementsInstanced(rglv::GL_TRIANGLES, count, rglv::GL_UNSIGNED_SHORT, indices, instanceCnt); }

void ref_gl_draw_arrays_instanced(void* h, int count, int instanceCnt) {
	IC(h).DrawArraysInstanced(rglv::GL_TRIANGLES, 0, count, instanceCnt); }

void ref_gl_store_color_tc(void* h, uint32_t* dst, int w, int hgt, int stride, int gamma) {
	auto* g = static_cast<RefGPU*>(h);
	g->tcCanvases.emplace_back(reinterpret_cast<PixelToaster::TrueColorPixel*>(dst), w, hgt, stride);
	IC(h).StoreColor(&g->tcCanvases.back(), gamma != 0); }

void ref_gl_store_color_fp(void* h, float* dst, int w, int hgt, int stride, int downsample) {
	auto* g = static_cast<RefGPU*>(h);
	g->fpCanvases.emplace_back(reinterpret_cast<PixelToaster::FloatingPointPixel*>(dst), w, hgt, stride);
	IC(h).StoreColor(&g->fpCanvases.back(), downsample != 0); }

/* dst: quad-swizzled canvas, 64-byte qfloat4 {r[4], g[4], b[4], a[4]} per 2x2 quad, 16-byte aligned (streaming stores) */
void ref_gl_store_color_quads(void* h, float* dst, int w, int hgt, int strideQuads) {
	auto* g = static_cast<RefGPU*>(h);
	g->qfCanvases.emplace_back(w, hgt, reinterpret_cast<rmlv::qfloat4*>(dst), strideQuads);
	IC(h).StoreColor(&g->qfCanvases.back()); }

void ref_gl_store_depth(void* h, float* dst) { IC(h).StoreDepth(dst); }

/* ---- helpers used to prepare inputs / pin primitives --------------------------------- */

/* in: dim*dim RGBA32F texels; out: dim*(2*dim) texels with the stacked mip chain
 * (rglr_texture.cxx:33-81) */
void ref_make_mipmap(const float* in, int dim, float* out) {
	rglr::Texture t;
	t.resize(dim, dim);
	std::memcpy(t.buf.data(), in, sizeof(float) * 4 * dim * dim);
	t.maybe_make_mipmap();
	std::memcpy(out, t.buf.data(), sizeof(float) * 4 * dim * dim * 2); }

/* rcpps / rsqrtps / oneover of this host CPU through the reference's own wrappers */
void ref_rcp(const float* in, float* out, int n) {
	for (int i = 0; i < n; ++i) { out[i] = _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps(in[i]))); } }
void ref_rsqrt(const float* in, float* out, int n) {
	for (int i = 0; i < n; ++i) { out[i] = rmlv::rsqrt(rmlv::mvec4f{in[i]}).get_x(); } }
void ref_oneover(const float* in, float* out, int n) {
	for (int i = 0; i < n; ++i) { out[i] = rmlv::oneover(rmlv::mvec4f{in[i]}).get_x(); } }

/* mat4 helpers (rmlm_mat4.hxx:197-216, rmlm_mat4.cxx:25) so tests can pin the host-side
 * matrix preparation */
void ref_mat4_mul(const float* a, const float* b, float* out) {
	auto r = ToMat4(a) * ToMat4(b);
	std::memcpy(out, r.ff.data(), sizeof(float) * 16); }
/* the reference's camera matrices (rglv_math.cxx:66-106), column-major out: used to build the bundled-scene fixtures */
void ref_look_at(const float* eye, const float* center, const float* up, float* out16) {
	const auto m = rglv::LookAt(rmlv::vec3{eye[0], eye[1], eye[2]}, rmlv::vec3{center[0], center[1], center[2]}, rmlv::vec3{up[0], up[1], up[2]});
	std::memcpy(out16, m.ff.data(), sizeof(float) * 16); }

void ref_perspective2(float fovy, float aspect, float znear, float zfar, float* out16) {
	const auto m = rglv::Perspective2(fovy, aspect, znear, zfar);
	std::memcpy(out16, m.ff.data(), sizeof(float) * 16); }

/* The reference's own mesh path: LoadOBJ (rglv_obj.cxx:193-283: parse, triangulate, normals) then
 * MakeArray(mesh, spec, ...) (rglv_mesh_util.cxx:81-136: de-duplicated SoA vertex arrays + uint16 indices, padded to
 * a multiple of four) -- what the viewer's $mesh node binds with UseBuffer(0 / 3 / 6) (src/viewer/node/mesh.cxx:40,
 * 101-104).  Returns the vertex count (padding included) or -1; arrays are [9][maxVerts]-style: a0.x a0.y a0.z a1.x ... */
int ref_load_obj_arrays(const char* path, const char* spec, float* soa9, int maxVerts, uint16_t* idx, int maxIdx, int* nIdx) {
	try {
		const auto mesh = rglv::LoadOBJ(std::pmr::string(path));
		rglv::VertexArray_F3F3F3 buf;
		rcls::vector<uint16_t> indices;
		rglv::MakeArray(mesh, std::string(spec), buf, indices);
		const int nv = buf.size();
		*nIdx = static_cast<int>(indices.size());
		if (nv > maxVerts || *nIdx > maxIdx) { return -1; }
		const rglv::Float3Array* arrs[3] = { &buf.a0, &buf.a1, &buf.a2 };
		for (int a = 0; a < 3; ++a) {
			std::memcpy(soa9 + (a * 3 + 0) * maxVerts, arrs[a]->x.data(), sizeof(float) * nv);
			std::memcpy(soa9 + (a * 3 + 1) * maxVerts, arrs[a]->y.data(), sizeof(float) * nv);
			std::memcpy(soa9 + (a * 3 + 2) * maxVerts, arrs[a]->z.data(), sizeof(float) * nv); }
		std::memcpy(idx, indices.data(), sizeof(uint16_t) * indices.size());
		return nv; }
	catch (...) { return -1; } }

/* The viewer's $perspective camera node (src/viewer/node/perspective.cxx:51-69), restated here because the node
 * classes live in anonymous namespaces of translation units that need the whole scene compiler: direction and right
 * vector from the two angles, up = cross(right, dir), view = LookAt(position, position + dir, up),
 * projection = translate(origin) * Perspective2(fov, aspect, 10, 1000) with origin = 0. */
void ref_perspective_camera(const float* position, float ha, float va, float fov, float aspect, float* view16, float* proj16) {
	const rmlv::vec3 pos{position[0], position[1], position[2]};
	const rmlv::vec3 dir{ cosf(va)*sinf(ha), sinf(va), cosf(va)*cosf(ha) };
	const rmlv::vec3 right{ sinf(ha-3.14F/2.0F), 0.0F, cosf(ha-3.14F/2.0F) };
	const rmlv::vec3 up = cross(right, dir);
	const auto v = rglv::LookAt(pos, pos + dir, up);
	auto p = rglv::Perspective2(fov, aspect, 10, 1000);
	p = rmlm::mat4::translate(0.0F, 0.0F, 0) * p;
	std::memcpy(view16, v.ff.data(), sizeof(float) * 16);
	std::memcpy(proj16, p.ff.data(), sizeof(float) * 16); }

void ref_mat4_inverse(const float* a, float* out) {
	auto r = rmlm::inverse(ToMat4(a));
	std::memcpy(out, r.ff.data(), sizeof(float) * 16); }

/* scalar rasteriser coverage, device-space float vertices (rglv_triangle.hxx:69-167);
 * out = w*h bytes, 1 where a pixel is covered */
void ref_raster_coverage(const float* xy6, int w, int hgt, uint8_t* out) {
	std::memset(out, 0, static_cast<size_t>(w) * hgt);
	CoverageProgram cp{out, w};
	rmlg::irect rect{rmlv::ivec2{0, 0}, rmlv::ivec2{w, hgt}};
	rglv::TriangleRasterizer<false, CoverageProgram> tr(cp, rect, hgt);
	tr.Draw(rmlv::vec4{xy6[0], xy6[1], 0, 1}, rmlv::vec4{xy6[2], xy6[3], 0, 1}, rmlv::vec4{xy6[4], xy6[5], 0, 1}); }

/* 4-wide rasteriser coverage from 28.4 fixed-point vertices (rglv_triangle.hxx:193-304) */
void ref_vraster_coverage(const int* x3, const int* y3, int rx0, int ry0, int rx1, int ry1, int w, int hgt, uint8_t* out) {
	std::memset(out, 0, static_cast<size_t>(w) * hgt);
	CoverageProgram cp{out, w};
	struct LaneProgram : CoverageProgram { void Lane(int) {} };
	LaneProgram lp{{out, w}};
	rmlg::irect rect{rmlv::ivec2{rx0, ry0}, rmlv::ivec2{rx1, ry1}};
	rglv::VTriangleRasterizer<false, LaneProgram> tr(lp, rect, hgt);
	tr.Draw(rmlv::mvec4i{x3[0]}, rmlv::mvec4i{x3[1]}, rmlv::mvec4i{x3[2]},
	        rmlv::mvec4i{y3[0]}, rmlv::mvec4i{y3[1]}, rmlv::mvec4i{y3[2]}, 1); }

/* ---- post filters, shadow-map program, marching cubes, telemetry overlay ------------------------ */

/* rglr::KawaseBlurFilter (src/rgl/rglr/rglr_kawase.cxx:22-81), the reference's own function over caller memory */
void ref_kawase_blur(const float* src, int srcStride, float* dst, int dstStride, int w, int hgt, int dist) {
	rglr::FloatingPointCanvas s(reinterpret_cast<PixelToaster::FloatingPointPixel*>(const_cast<float*>(src)), w, hgt, srcStride);
	rglr::FloatingPointCanvas d(reinterpret_cast<PixelToaster::FloatingPointPixel*>(dst), w, hgt, dstStride);
	rglr::KawaseBlurFilter(s, d, dist, 0, hgt); }

namespace {
/* `$glow`'s canvas shader.  It lives in an unnamed namespace of src/viewer/node/glow.cxx:24-39 and cannot be linked, so
 * its one expression is restated here; the filter that applies it, rglr::Filter<SHADER, CONVERTER>
 * (rglr_algorithm.hxx:107-144), and the converters (rglr_canvas_util.hxx) are the reference's own. */
struct GlowShaderRestated {
	static rmlv::qfloat3 ShadeCanvas(rmlv::qfloat2, rmlv::qfloat3 c1, rmlv::qfloat3 c2) {
		const rmlv::qfloat blurAmt{ 0.700F };
		const rmlv::qfloat brightness{ 0.5F };
		rmlv::qfloat3 out;
		out = (c1 + c2*blurAmt) * brightness;
		return out; }};
}

void ref_glow_filter(const float* quads, int strideQuads, const float* blur, int blurW, int blurH, int blurStride,
                     uint32_t* dst, int w, int hgt, int stride, int gamma) {
	rglr::QFloat4Canvas src0(w, hgt, reinterpret_cast<rmlv::qfloat4*>(const_cast<float*>(quads)), strideQuads);
	rglr::FloatingPointCanvas src1(reinterpret_cast<PixelToaster::FloatingPointPixel*>(const_cast<float*>(blur)), blurW, blurH, blurStride);
	rglr::TrueColorCanvas out(reinterpret_cast<PixelToaster::TrueColorPixel*>(dst), w, hgt, stride);
	rmlg::irect rect{ { 0, 0 }, { w, hgt } };
	if (gamma) { rglr::Filter<GlowShaderRestated, rglr::sRGB>(src0, src1, out, rect); }
	else { rglr::Filter<GlowShaderRestated, rglr::LinearColor>(src0, src1, out, rect); } }

/* the shadow-map GPU of a `$layer` (src/viewer/node/gllayer.cxx:39-45): BaseProgram, depth only */
void ref_gpu_install_shadow_program(void* h) {
	auto& gpu = static_cast<RefGPU*>(h)->gpu;
	gpu.Install(0, 0, rglv::GPUBinImpl<rglv::BaseProgram>::MakeBinProgramPtrs());
	gpu.Install(0, 0x6a2, rglv::GPUTileImpl<rglr::QFloat3FragmentCursor, rglr::QFloatFragmentCursor, rglv::BaseProgram, false, true, rglv::DepthLT, true, false, rglv::BlendOff>::MakeDrawProgramPtrs()); }

/* the reference's marching-cubes tables (rglv_marching_cubes.cxx) */
void ref_mc_tables(int16_t* edgeFlags256, int8_t* tri256x16, uint8_t* edgeConn12x2) {
	for (int i = 0; i < 256; ++i) {
		edgeFlags256[i] = rglv::cube_edge_flags[i];
		for (int k = 0; k < 16; ++k) { tri256x16[i * 16 + k] = static_cast<int8_t>(rglv::tritable[i][k]); } }
	for (int e = 0; e < 12; ++e) { edgeConn12x2[2 * e] = rglv::edge_connection[e][0]; edgeConn12x2[2 * e + 1] = rglv::edge_connection[e][1]; } }

namespace {
/* `$mc`'s field and block walk.  Both live in the unnamed namespace of src/viewer/node/mc.cxx (Surface :95-110,
 * BlockDivider :33-93, Impl::ResolveImpl :230-300) and cannot be linked without the node graph, so they are restated
 * here line by line; what they call -- rglv::march_sdf_vao, the case tables, rmlv's vec / mvec4f arithmetic and its
 * sine approximation -- is the reference's own code.  `sin` / `abs` on floats are the float overloads, as with the
 * reference's native compiler. */
struct McSurface {
	float timeInSeconds_;
	float sample(rmlv::vec3 pos) const {
		float distort = 0.60F * sinf(5.0F*(pos.x + timeInSeconds_ / 4.0F))* sinf(2.0F*(pos.y + (timeInSeconds_ / 1.33F)));
		return (length(pos) - 3.0F) + (distort * sinf(timeInSeconds_ / 2.0F) + 1.0F); }
	rmlv::mvec4f sample(rmlv::qfloat3 pos) const {
		using rmlv::mvec4f;
		auto T = mvec4f{ timeInSeconds_ };
		auto distort = mvec4f{0.60F} * sin(5.0F*(pos.x + T / 4.0F))* sin(2.0F*(pos.y + (T / 1.33F)));
		return (rmlv::length(pos) - 3.0F) + (distort * sin(T / 2.0F) + 1.0F); }};

struct McAABB { rmlv::vec3 leftTopBack, rightBottomFront; };

void McDivide(McAABB b, int limit, std::vector<McAABB>& out) {
	using rmlv::vec3;
	if (limit == 0) { out.emplace_back(b); return; }
	const auto mid = mix(b.leftTopBack, b.rightBottomFront, 0.5F);
	const auto ltb = b.leftTopBack;
	const auto rbf = b.rightBottomFront;
	const McAABB sub[8] = {
		{ vec3{ ltb.x, ltb.y, ltb.z }, vec3{ mid.x, mid.y, mid.z } }, { vec3{ mid.x, ltb.y, ltb.z }, vec3{ rbf.x, mid.y, mid.z } },
		{ vec3{ ltb.x, mid.y, ltb.z }, vec3{ mid.x, rbf.y, mid.z } }, { vec3{ mid.x, mid.y, ltb.z }, vec3{ rbf.x, rbf.y, mid.z } },
		{ vec3{ ltb.x, ltb.y, mid.z }, vec3{ mid.x, mid.y, rbf.z } }, { vec3{ mid.x, ltb.y, mid.z }, vec3{ rbf.x, mid.y, rbf.z } },
		{ vec3{ ltb.x, mid.y, mid.z }, vec3{ mid.x, rbf.y, rbf.z } }, { vec3{ mid.x, mid.y, mid.z }, vec3{ rbf.x, rbf.y, rbf.z } } };
	for (const auto& s : sub) { McDivide(s, limit - 1, out); } }

/* ResolveImpl (mc.cxx:230-300); returns false when the block is skipped by the distance test */
bool McResolve(const McSurface& field_, McAABB block, int dim, rglv::VertexArray_F3F3F3& vao) {
	using rmlv::vec3; using rmlv::mvec4f; using rmlv::qfloat;
	const int stride = 64;
	std::array<std::array<float, stride*stride>, 2> buf;
	int top = 1, bot = 0;
	float sy = block.leftTopBack.y;
	float delta = (block.rightBottomFront.x - block.leftTopBack.x) / float(dim);
	const qfloat vdelta{ delta * 4 };
	const auto mid = mix(block.leftTopBack, block.rightBottomFront, 0.5F);
	const auto R = length(block.leftTopBack - mid);
	float D = field_.sample(mid);
	if (fabsf(D)*0.5F > R) { return false; }
	auto fillSlice = [&]() {
		mvec4f fooZ{ block.leftTopBack.z };
		mvec4f fooY{ sy };
		for (int iz=0; iz<dim+1; iz++, fooZ += delta) {
			mvec4f fooX{ block.leftTopBack.x };
			fooX += mvec4f{ 0, delta, delta*2, delta*3 };
			for (int ix{0}; ix<dim+1; ix+=4, fooX+=vdelta) {
				auto distance = field_.sample({ fooX, fooY, fooZ });
				_mm_storeu_ps(&(buf[bot][iz*stride + ix]), distance.v); }}};
	fillSlice();
	sy -= delta;
	vec3 origin = block.leftTopBack;
	for (int iy=0; iy<dim; iy++, sy -= delta) {
		origin.y -= delta;
		std::swap(top, bot);
		fillSlice();
		origin.z = block.leftTopBack.z;
		for (int iz = 0; iz < dim; iz++, origin.z += delta) {
			origin.x = block.leftTopBack.x;
			for (int ix = 0; ix < dim; ix++, origin.x += delta) {
				rglv::Cell cell;
				cell.value[0] = buf[bot][ iz   *stride + ix];     cell.pos[0] = origin;
				cell.value[1] = buf[bot][ iz   *stride + ix + 1]; cell.pos[1] = vec3{ origin.x + delta, origin.y, origin.z };
				cell.value[2] = buf[top][ iz   *stride + ix + 1]; cell.pos[2] = vec3{ origin.x + delta, origin.y + delta, origin.z };
				cell.value[3] = buf[top][ iz   *stride + ix];     cell.pos[3] = vec3{ origin.x,         origin.y + delta, origin.z };
				cell.value[4] = buf[bot][(iz+1)*stride + ix];     cell.pos[4] = vec3{ origin.x,         origin.y, origin.z + delta };
				cell.value[5] = buf[bot][(iz+1)*stride + ix + 1]; cell.pos[5] = vec3{ origin.x + delta, origin.y, origin.z + delta };
				cell.value[6] = buf[top][(iz+1)*stride + ix + 1]; cell.pos[6] = vec3{ origin.x + delta, origin.y + delta, origin.z + delta };
				cell.value[7] = buf[top][(iz+1)*stride + ix];     cell.pos[7] = vec3{ origin.x,         origin.y + delta, origin.z + delta };
				rglv::march_sdf_vao(vao, delta, cell, field_); }}}
	return true; }
}

/* `$mc`'s Main (mc.cxx:171-193) with the jobs run one after the other in block order.  Output: vertex SoA
 * (x | y | z and nx | ny | nz, `cap` floats each), per non-empty block its first vertex (padded to 4 like the device
 * library) and vertex count.  Returns the padded vertex total, or -1 when `cap` / `blockCap` is too small. */
int ref_march_surface(float t, int precision, int forkDepth, float range, float* pos3, float* nrm3, int cap,
                      int* blockFirst, int* blockVerts, int blockCap, int* nblocks) {
	McSurface field{t};
	std::vector<McAABB> blocks;
	McDivide(McAABB{ rmlv::vec3{-range, range, -range}, rmlv::vec3{ range, -range, range} }, forkDepth, blocks);
	const int subDim = precision >> forkDepth;
	int total = 0, nb = 0;
	for (const auto& b : blocks) {
		rglv::VertexArray_F3F3F3 vao;
		if (!McResolve(field, b, subDim, vao)) { continue; }
		const int n = vao.size();
		if (n == 0) { continue; }
		const int padded = (n + 3) & ~3;
		if (total + padded > cap || nb >= blockCap) { return -1; }
		for (int i = 0; i < n; ++i) {
			pos3[total + i] = vao.a0.x[i]; pos3[cap + total + i] = vao.a0.y[i]; pos3[2 * cap + total + i] = vao.a0.z[i];
			nrm3[total + i] = vao.a1.x[i]; nrm3[cap + total + i] = vao.a1.y[i]; nrm3[2 * cap + total + i] = vao.a1.z[i]; }
		for (int i = n; i < padded; ++i) {
			pos3[total + i] = pos3[cap + total + i] = pos3[2 * cap + total + i] = 0.0F;
			nrm3[total + i] = nrm3[cap + total + i] = nrm3[2 * cap + total + i] = 0.0F; }
		blockFirst[nb] = total; blockVerts[nb] = n; ++nb;
		total += padded; }
	*nblocks = nb;
	return total; }

/* render_jobsys (src/viewer/jobsys_vis.cxx:26-90) over a given list of spans: they are put where the reference's
 * job system keeps its measurements (jobsys::measurements_pt, one vector per worker = lane) */
void ref_render_spans(uint32_t* canvas, int w, int hgt, int stride, int left, int top, float xscale,
                      const double* startEnd, const uint32_t* raw, const int* lane, int count) {
	auto saved = jobsys::measurements_pt;
	int lanes = 0;
	for (int i = 0; i < count; ++i) { lanes = std::max(lanes, lane[i] + 1); }
	jobsys::measurements_pt.assign(lanes, {});
	for (int i = 0; i < count; ++i) {
		jobsys::measurements_pt[lane[i]].push_back(jobsys::JobStat{ startEnd[2 * i], startEnd[2 * i + 1], raw[i] }); }
	rglr::TrueColorCanvas c(reinterpret_cast<PixelToaster::TrueColorPixel*>(canvas), w, hgt, stride);
	rqv::render_jobsys(left, top, xscale, c);
	jobsys::measurements_pt = saved; }

}  // extern "C"
