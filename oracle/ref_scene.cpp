/* TEST INFRASTRUCTURE -- the reference's own scene front-end behind a C entry point.
 *
 * Compiled by oracle/build_ref.sh together with the reference's node graph where it lies under /root/reference
 * (src/viewer/node/*.cxx, src/viewer/compile.cxx, the JSON / mesh / texture / font loaders underneath) into BOTH
 * oracle/_ref/librsr_ref.so (the pure CPU reference) and oracle/_ref/librsr_dropin.so (the same translation units with
 * GPU::RunImpl replaced by the C-ABI binding, rsr_b200/host/rglv_gpu_cuda.cxx).  A bundled data/scene/X.lua file is
 * turned into JSON by the reference's own data/scene/host.lua run by the vendored Lua interpreter (build_ref.sh), and
 * that JSON drives the node graph here exactly as src/viewer/perf.cxx does (Application::Main :173-206,
 * PrepareBuiltInNodes :247-256, ComputeAndRenderFrame :258-280, MaybeRecompile :282-330): the driver below restates
 * those ~60 lines of perf.cxx (its main() cannot be linked into a library and its mesh-store call is stale,
 * `load_dir`); every node, the compiler, the linker and the renderers underneath are the reference's own code.
 * tests/test_scenes_gpu.py renders the bundled scenes through both libraries and compares the frames bit for bit. */
#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "src/rcl/rclma/rclma_framepool.hxx"
#include "src/rcl/rclmt/rclmt_jobsys.hxx"
#include "src/rcl/rclr/rclr_algorithm.hxx"
#include "src/rgl/rglr/rglr_canvas.hxx"
#include "src/rgl/rglr/rglr_texture_store.hxx"
#include "src/rgl/rglv/rglv_camera.hxx"
#include "src/rgl/rglv/rglv_gpu.hxx"
#include "src/rgl/rglv/rglv_mesh_store.hxx"
#include "src/rml/rmlv/rmlv_vec.hxx"
#include "src/viewer/compile.hxx"
#include "src/viewer/node/i_output.hxx"
#include "src/viewer/node/multivalue.hxx"
#include "src/viewer/node/uicamera.hxx"
#include "3rdparty/gason/gason.h"

using namespace rqdq;
namespace jobsys = rclmt::jobsys;
namespace framepool = rclma::framepool;

namespace {

struct Scene {
	rglv::MeshStore meshStore;
	rglr::TextureStore textureStore;
	rglv::HandyCam camera;
	std::vector<char> jsonText;              // gason parses in place
	std::unique_ptr<JsonAllocator> jsonAlloc;
	JsonValue jsonRoot;
	rqv::NodeList nodes;
	std::shared_ptr<rqv::MultiValueNode> globals, sync;
	std::shared_ptr<rqv::UICamera> uiCamera; };

bool g_framepoolReady = false;

}  // namespace

extern "C" {

/* perf.cxx Application::Main :173-199 -- dataDir holds mesh/ and texture/; the process's working directory must be the
 * directory that CONTAINS data/ (`$image` nodes open "data/texture/x.png" relative to it, node/image.cxx).  jobsys must
 * be running (ref_init).  Returns nullptr when the scene does not compile / link. */
void* ref_scene_load(const char* jsonPath, const char* dataDir) {
	if (!g_framepoolReady) { framepool::Init(); g_framepoolReady = true; }
	rglv::doubleBuffer = false;   // a frame's pixels are in the canvas when its root job is done (perf.cxx runs with the one-frame delay)
	auto s = std::make_unique<Scene>();
	const std::string dd(dataDir);
	s->textureStore.LoadDir(dd + "/texture");
	s->meshStore.LoadDir(dd + "/mesh/");
	// PrepareBuiltInNodes (perf.cxx:247-256)
	s->globals = std::make_shared<rqv::MultiValueNode>("globals", rqv::InputList());
	s->sync = std::make_shared<rqv::MultiValueNode>("sync", rqv::InputList());
	s->uiCamera = std::make_shared<rqv::UICamera>("uiCamera", rqv::InputList(), s->camera);
	std::ifstream in(jsonPath, std::ios::binary);
	if (!in) { std::cerr << "ref_scene_load: cannot open " << jsonPath << "\n"; return nullptr; }
	s->jsonText.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
	s->jsonText.push_back('\0');
	s->jsonAlloc = std::make_unique<JsonAllocator>();
	char* endp = nullptr;
	if (jsonParse(s->jsonText.data(), &endp, &s->jsonRoot, *s->jsonAlloc) != JSON_OK) {
		std::cerr << "ref_scene_load: JSON parse error in " << jsonPath << "\n"; return nullptr; }
	// MaybeRecompile (perf.cxx:296-321)
	bool success;
	rqv::NodeList userNodes;
	std::tie(success, userNodes) = rqv::CompileDocument(s->jsonRoot, s->meshStore);
	if (!success) { std::cerr << "ref_scene_load: compile failed\n"; return nullptr; }
	rqv::NodeList all{ s->globals, s->sync, s->uiCamera };
	all.insert(end(all), begin(userNodes), end(userNodes));
	if (!rqv::Link(all)) { std::cerr << "ref_scene_load: link failed\n"; return nullptr; }
	s->nodes = std::move(all);
	return s.release(); }

void ref_scene_free(void* h) {
	auto* s = static_cast<Scene*>(h);
	if (s) { rclr::for_each(s->nodes, [](auto& n) { n->DisconnectAll(); }); delete s; } }

/* one frame at wall-clock time t into out (w x h, 0x00RRGGBB): the globals of perf.cxx:201-205, then
 * ComputeAndRenderFrame (perf.cxx:258-280).  Brackets the frame with work_start / work_end. */
int ref_scene_render(void* h, int w, int hgt, int tileX, int tileY, float t, uint32_t* out) {
	auto* s = static_cast<Scene*>(h);
	rglr::TrueColorCanvas canvas(reinterpret_cast<PixelToaster::TrueColorPixel*>(out), w, hgt);
	s->globals->Upsert("wallclock", t);
	s->globals->Upsert("windowSize", rmlv::vec2(float(w), float(hgt)));
	s->globals->Upsert("tileSize", rmlv::vec2(float(tileX), float(tileY)));
	s->globals->Upsert("windowAspect", w / float(hgt));
	const std::string_view selector{"root"};
	const auto match = rclr::find_if(s->nodes, [=](const auto& node) { return node->get_id() == selector; });
	if (match == end(s->nodes)) { std::cerr << "output node \"root\" not found\n"; return 1; }
	auto* node = dynamic_cast<rqv::IOutput*>(match->get());
	if (node == nullptr) { std::cerr << "node \"root\" is not an OutputNode\n"; return 2; }
	jobsys::work_start();
	jobsys::reset();
	framepool::Reset();
	auto rootJob = jobsys::make_job(jobsys::noop);
	for (auto& n : s->nodes) { n->Reset(); }
	rqv::ComputeIndegreesFrom(node);
	node->set_indegreeWaitCnt(1);
	node->SetOutputCanvas(&canvas);
	node->AddLink(rootJob);
	node->Run();
	jobsys::wait(rootJob);
	jobsys::work_end();
	return 0; }

/* perf.cxx's timed loop (Application::Main :207-232): `frames` frames at t0, t0 + dt, ... between one work_start /
 * work_end, with the reference's default doubleBuffer mode if asked (a frame's pixels then land one Run later, which is
 * what a viewer sees too).  Returns the elapsed seconds. */
double ref_scene_bench(void* h, int w, int hgt, int tileX, int tileY, int frames, float t0, float dt, int doubleBuffer, uint32_t* out) {
	auto* s = static_cast<Scene*>(h);
	rglr::TrueColorCanvas canvas(reinterpret_cast<PixelToaster::TrueColorPixel*>(out), w, hgt);
	s->globals->Upsert("windowSize", rmlv::vec2(float(w), float(hgt)));
	s->globals->Upsert("tileSize", rmlv::vec2(float(tileX), float(tileY)));
	s->globals->Upsert("windowAspect", w / float(hgt));
	const std::string_view selector{"root"};
	const auto match = rclr::find_if(s->nodes, [=](const auto& node) { return node->get_id() == selector; });
	if (match == end(s->nodes)) { return -1.0; }
	auto* node = dynamic_cast<rqv::IOutput*>(match->get());
	if (node == nullptr) { return -1.0; }
	rglv::doubleBuffer = doubleBuffer != 0;
	jobsys::work_start();
	const auto begin = std::chrono::steady_clock::now();
	for (int fn = 0; fn < frames; ++fn) {
		s->globals->Upsert("wallclock", t0 + dt * float(fn));
		jobsys::reset();
		framepool::Reset();
		auto rootJob = jobsys::make_job(jobsys::noop);
		for (auto& n : s->nodes) { n->Reset(); }
		rqv::ComputeIndegreesFrom(node);
		node->set_indegreeWaitCnt(1);
		node->SetOutputCanvas(&canvas);
		node->AddLink(rootJob);
		node->Run();
		jobsys::wait(rootJob); }
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - begin).count();
	jobsys::work_end();
	rglv::doubleBuffer = false;
	return secs; }

}  // extern "C"
