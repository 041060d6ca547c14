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

}  // extern "C"
