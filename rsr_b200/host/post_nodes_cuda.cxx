// post_nodes_cuda.cxx -- the viewer's `$buffers` -> `$kawase` -> `$glow` chain with its canvases on the device.
//
// What a maintainer of usrlocalben/rsr adds next to rglv_gpu_cuda.cxx so that a glow scene (data/scene/
// instanced-cubes.lua) stops moving float canvases over PCIe: these three node classes REPLACE
// src/viewer/node/buffers.cxx, kawase.cxx and glow.cxx (same JSON names, same inputs / slots / parameters, so scene
// files do not change).  The reference's nodes store the frame as a quad-swizzled RGBA32F canvas + a half-size linear
// one in host memory (41 MB per 1080p frame), blur the half-size one on the job system (rglr::KawaseBlurFilter) and
// combine both into the window's true-colour canvas (rglr::Filter<GlowShader, ...>).  Here the two float canvases are
// device memory (CudaCanvasAlloc): GL::StoreColor(&canvas) is recorded exactly as before and the binding turns it into
// a device store; the blur passes and the combine are rsrcu_kawase_blur / rsrcu_glow on the stream of the context that
// rendered the frame; only the 8-bit frame comes back.  Results are bit-identical to the reference's chain
// (tests/test_scenes_gpu.py renders instanced-cubes.lua through both).
//
// Compiled by oracle/build_ref.sh into oracle/_ref/librsr_dropin.so in place of the three reference translation units.
#include <iostream>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <tuple>

#include "src/rcl/rclmt/rclmt_jobsys.hxx"
#include "src/rgl/rglr/rglr_canvas.hxx"
#include "src/viewer/compile.hxx"
#include "src/viewer/node/base.hxx"
#include "src/viewer/node/i_canvas.hxx"
#include "src/viewer/node/i_gpu.hxx"
#include "src/viewer/node/i_output.hxx"

#include "rsrcu.h"

namespace rqdq {
namespace rglv {
void* CudaCanvasAlloc(size_t bytes);
void CudaCanvasFree(void* p);
rsrcu_ctx* CudaWriterOf(const void* devicePtr);
void CudaNoteWriter(const void* devicePtr, rsrcu_ctx* ctx);
}  // namespace rglv

namespace {

using namespace rqv;
namespace jobsys = rclmt::jobsys;

void Check(int rc, const char* what) {
	if (rc != RSRCU_OK) { throw std::runtime_error(std::string(what) + ": " + rsrcu_last_error()); } }

// device memory behind one canvas object, re-allocated when the target size changes
class DeviceBlock {
	void* ptr_{nullptr};
	size_t bytes_{0};
public:
	DeviceBlock() = default;
	DeviceBlock(const DeviceBlock&) = delete;
	auto operator=(const DeviceBlock&) -> DeviceBlock& = delete;
	~DeviceBlock() { rglv::CudaCanvasFree(ptr_); }
	auto Reserve(size_t bytes) -> void* {
		if (bytes > bytes_) { rglv::CudaCanvasFree(ptr_); ptr_ = rglv::CudaCanvasAlloc(bytes); bytes_ = bytes; }
		return ptr_; } };

template <class NODE>
auto JobFor(NODE* self, void (NODE::*fn)()) -> jobsys::Job* {
	struct Thunk { static void Run(jobsys::Job*, unsigned, std::tuple<NODE*, void (NODE::*)()>* d) { (std::get<0>(*d)->*std::get<1>(*d))(); } };
	return jobsys::make_job(Thunk::Run, std::tuple{self, fn}); }

// ---- `$buffers` (replaces node/buffers.cxx) -----------------------------------------------------------------------
class DeviceBuffers final : public ICanvas {
	DeviceBlock colorMem_, halfMem_;
	std::optional<rglr::QFloat4Canvas> color_;
	std::optional<rglr::FloatingPointCanvas> half_;
	const bool wantColor_, wantHalf_;
	IGPU* gpu_{nullptr};
public:
	DeviceBuffers(std::string_view id, InputList inputs, bool color, bool half) :
		ICanvas(id, std::move(inputs)), wantColor_(color), wantHalf_(half) {}

	auto Connect(std::string_view attr, NodeBase* other, std::string_view slot) -> bool override {
		if (attr != "gpu") { return ICanvas::Connect(attr, other, slot); }
		gpu_ = dynamic_cast<IGPU*>(other);
		if (gpu_ == nullptr) { TYPE_ERROR(IGPU); }
		return gpu_ != nullptr; }
	void DisconnectAll() override { ICanvas::DisconnectAll(); gpu_ = nullptr; }
	void AddDeps() override { AddDep(gpu_); }
	auto IsValid() -> bool override {
		if (gpu_ == nullptr) { std::cerr << "buffers(" << get_id() << ") has no gpu" << std::endl; return false; }
		return ICanvas::IsValid(); }

	void Main() override {
		gpu_->AddLink(AfterAll(JobFor(this, &DeviceBuffers::Record)));
		gpu_->Run(); }

	void Record() {
		auto& ic = gpu_->IC();
		const auto size = gpu_->GetTargetSize();
		if (wantColor_) {
			const int strideQuads = size.x / 2;
			auto* mem = static_cast<rmlv::qfloat4*>(colorMem_.Reserve(static_cast<size_t>(strideQuads) * (size.y / 2) * 64));
			color_.emplace(size.x, size.y, mem, strideQuads);
			ic.StoreColor(&*color_); }
		if (wantHalf_) {
			auto* mem = static_cast<PixelToaster::FloatingPointPixel*>(halfMem_.Reserve(static_cast<size_t>(size.x / 2) * (size.y / 2) * 16));
			half_.emplace(mem, size.x / 2, size.y / 2, size.x / 2);
			ic.StoreColor(&*half_, /*downsample=*/true); }
		ic.Finish();
		auto renderJob = gpu_->Render();
		jobsys::add_link(renderJob, JobFor(this, &DeviceBuffers::Done));
		jobsys::run(renderJob); }

	void Done() { RunLinks(); }

	auto GetCanvas(std::string_view slot) -> std::pair<int, const void*> override {
		if (slot == "color" && color_) { return {ICanvas::CT_FLOAT4_QUADS, &*color_}; }
		if (slot == "half" && half_) { return {ICanvas::CT_FLOAT4_LINEAR, &*half_}; }
		throw std::runtime_error("renderbuffer: GetCanvas with invalid name (the device chain stores color and half)"); } };

struct DeviceBuffersCompiler final : NodeCompiler {
	void Build() override {
		if (!Input("gpu", /*required=*/true)) { return; }
		if (DataBool("depth", false)) { std::cerr << "buffers: depth canvases are not stored by the reference either (node/buffers.cxx:79)\n"; }
		out_ = std::make_shared<DeviceBuffers>(id_, std::move(inputs_), DataBool("color", false), DataBool("half", false)); } };

// ---- `$kawase` (replaces node/kawase.cxx) -------------------------------------------------------------------------
class DeviceKawase final : public ICanvas {
	const int intensity_;
	ICanvas* input_{nullptr};
	std::string inputSlot_;
	DeviceBlock memA_, memB_;
	std::optional<rglr::FloatingPointCanvas> a_, b_;
	const rglr::FloatingPointCanvas* output_{nullptr};
public:
	DeviceKawase(std::string_view id, InputList inputs, int intensity) : ICanvas(id, std::move(inputs)), intensity_(intensity) {}

	auto Connect(std::string_view attr, NodeBase* other, std::string_view slot) -> bool override {
		if (attr != "input") { return ICanvas::Connect(attr, other, slot); }
		input_ = dynamic_cast<ICanvas*>(other);
		if (input_ == nullptr) { TYPE_ERROR(ICanvas); return false; }
		inputSlot_ = slot;
		return true; }
	void DisconnectAll() override { input_ = nullptr; ICanvas::DisconnectAll(); }
	void AddDeps() override { ICanvas::AddDeps(); AddDep(input_); }
	auto IsValid() -> bool override {
		if (input_ == nullptr) { std::cerr << "kawase(" << get_id() << ") has no input" << std::endl; return false; }
		return ICanvas::IsValid(); }

	void Main() override {
		input_->AddLink(AfterAll(JobFor(this, &DeviceKawase::Blur)));
		input_->Run(); }

	void Blur() {
		const auto in = input_->GetCanvas(inputSlot_);
		if (in.first != ICanvas::CT_FLOAT4_LINEAR) { throw std::runtime_error("blur requires a FloatingPointCanvas"); }
		const auto* src = static_cast<const rglr::FloatingPointCanvas*>(in.second);
		output_ = src;
		if (intensity_ > 0) {
			rsrcu_ctx* ctx = rglv::CudaWriterOf(src->cdata());
			if (ctx == nullptr) { throw std::runtime_error("kawase: the input canvas is not on the device"); }
			const int w = src->width(), h = src->height();
			const size_t bytes = static_cast<size_t>(w) * h * 16;
			a_.emplace(static_cast<PixelToaster::FloatingPointPixel*>(memA_.Reserve(bytes)), w, h, w);
			if (intensity_ > 1) { b_.emplace(static_cast<PixelToaster::FloatingPointPixel*>(memB_.Reserve(bytes)), w, h, w); }
			// node/kawase.cxx:96-124: pass d reads what pass d - 1 wrote, two canvases take turns
			const rglr::FloatingPointCanvas* from = src;
			rglr::FloatingPointCanvas* to = &*a_;
			for (int dist = 0; dist < intensity_; ++dist) {
				Check(rsrcu_kawase_blur(ctx, from->cdata(), from->stride(), to->data(), to->stride(), w, h, dist), "rsrcu_kawase_blur");
				rglv::CudaNoteWriter(to->data(), ctx);
				from = to;
				to = (to == &*a_ && b_) ? &*b_ : &*a_; }
			output_ = from; }
		RunLinks(); }

	auto GetCanvas(std::string_view) -> std::pair<int, const void*> override { return {ICanvas::CT_FLOAT4_LINEAR, output_}; } };

struct DeviceKawaseCompiler final : NodeCompiler {
	void Build() override {
		if (!Input("input", /*required=*/true)) { return; }
		out_ = std::make_shared<DeviceKawase>(id_, std::move(inputs_), DataInt("intensity", 1)); } };   // (taskSize sized the CPU jobs)

// ---- `$glow` (replaces node/glow.cxx) -----------------------------------------------------------------------------
class DeviceGlow final : public IOutput {
	const bool sRGB_;
	rglr::TrueColorCanvas* out_{nullptr};
	ICanvas* image_{nullptr};
	ICanvas* blur_{nullptr};
	std::string imageSlot_, blurSlot_;
public:
	DeviceGlow(std::string_view id, InputList inputs, bool sRGB) : IOutput(id, std::move(inputs)), sRGB_(sRGB) {}

	auto Connect(std::string_view attr, NodeBase* other, std::string_view slot) -> bool override {
		ICanvas** which = attr == "image" ? &image_ : attr == "blur" ? &blur_ : nullptr;
		if (which == nullptr) { return IOutput::Connect(attr, other, slot); }
		*which = dynamic_cast<ICanvas*>(other);
		if (*which == nullptr) { TYPE_ERROR(ICanvas); return false; }
		(attr == "image" ? imageSlot_ : blurSlot_) = slot;
		return true; }
	void DisconnectAll() override { IOutput::DisconnectAll(); image_ = nullptr; blur_ = nullptr; }
	void AddDeps() override { IOutput::AddDeps(); AddDep(image_); AddDep(blur_); }
	void Reset() override { IOutput::Reset(); out_ = nullptr; }
	auto IsValid() -> bool override {
		if (image_ == nullptr || blur_ == nullptr) { std::cerr << "glow(" << get_id() << ") needs image and blur" << std::endl; return false; }
		return IOutput::IsValid(); }

	void Main() override {
		auto job = Render();
		image_->AddLink(AfterAll(job));
		blur_->AddLink(AfterAll(job));
		image_->Run();
		blur_->Run(); }

	auto Render() -> jobsys::Job* override { return JobFor(this, &DeviceGlow::Combine); }

	void Combine() {
		const auto img = image_->GetCanvas(imageSlot_);
		const auto blr = blur_->GetCanvas(blurSlot_);
		if (img.first != ICanvas::CT_FLOAT4_QUADS) { throw std::runtime_error("expected image to be CT_FLOAT4_QUADS"); }
		if (blr.first != ICanvas::CT_FLOAT4_LINEAR) { throw std::runtime_error("expected blur to be CT_FLOAT4_LINEAR"); }
		const auto* image = static_cast<const rglr::QFloat4Canvas*>(img.second);
		const auto* blur = static_cast<const rglr::FloatingPointCanvas*>(blr.second);
		rsrcu_ctx* ctx = rglv::CudaWriterOf(blur->cdata());
		if (ctx == nullptr || rglv::CudaWriterOf(image->cdata()) == nullptr) { throw std::runtime_error("glow: the input canvases are not on the device"); }
		// (host destination: written when the call returns, like the reference's filter jobs before RunLinks)
		Check(rsrcu_glow(ctx, image->cdata(), image->stride(), blur->cdata(), blur->stride(), sRGB_ ? 1 : 0,
		                 reinterpret_cast<uint32_t*>(out_->data()), /*dst_is_device=*/0, image->width(), image->height(), out_->stride()), "rsrcu_glow");
		RunLinks(); }

	void SetOutputCanvas(rglr::TrueColorCanvas* canvas) override { out_ = canvas; } };

struct DeviceGlowCompiler final : NodeCompiler {
	void Build() override {
		if (!Input("image", /*required=*/true)) { return; }
		if (!Input("blur", /*required=*/true)) { return; }
		out_ = std::make_shared<DeviceGlow>(id_, std::move(inputs_), DataBool("sRGB", true)); } };

struct Registration { Registration() {
	auto& reg = NodeRegistry::GetInstance();
	reg.Register("$buffers", []() { return std::make_unique<DeviceBuffersCompiler>(); });
	reg.Register("$kawase", []() { return std::make_unique<DeviceKawaseCompiler>(); });
	reg.Register("$glow", []() { return std::make_unique<DeviceGlowCompiler>(); });
}} registration{};

}  // namespace
}  // namespace rqdq
