// cpp_demo.cpp -- renders one depth-tested, bilinear-textured quad through the C++ mirror
// (rsr_b200/host/rglv_cuda.hxx) and prints the FNV-1a hash of the 320x180 frame.
// tests/test_cpp_host.py compares the hash with the Python route on the same inputs; a second frame goes through the
// device-resident glow chain (quads + half-size stores, two Kawase passes, glow) and is compared with the reference's filters.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../rsr_b200/host/rglv_cuda.hxx"

int main() {
	using namespace rglvcu;
	const int W = 320, H = 180, D = 16;
	// SoA quad, uv, indices
	alignas(64) static float px[4] = { -1, 1, 1, -1 }, py[4] = { -1, -1, 1, 1 }, pz[4] = { 0, 0.3f, 0, -0.3f };
	alignas(64) static float u[4] = { 0, 1, 1, 0 }, v[4] = { 0, 0, 1, 1 };
	static uint16_t idx[6] = { 0, 1, 2, 0, 2, 3 };
	// texture + stacked mip chain rows (content: simple integer pattern)
	std::vector<float> tex(static_cast<size_t>(D) * 2 * D * 4, 0.0f);
	for (int y = 0; y < D; ++y) for (int x = 0; x < D; ++x) for (int c = 0; c < 4; ++c) {
		tex[(static_cast<size_t>(y) * D + x) * 4 + c] = static_cast<float>((x * 7 + y * 13 + c * 5) % 32) / 32.0f; }
	// column-major matrices: view = translate(0,0,-3); projection = perspective(45 deg, 16:9, 1, 10)
	float view[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, -3, 1 };
	const float f = 1.0f / std::tan(0.39269908f);
	float proj[16] = { f / (16.0f / 9.0f), 0, 0, 0, 0, f, 0, 0, 0, 0, -11.0f / 9.0f, -1, 0, 0, -20.0f / 9.0f, 0 };

	std::vector<uint32_t> out(static_cast<size_t>(W) * H, 0), glow(static_cast<size_t>(W) * H, 0);
	try {
		GPU gpu(0);
		gpu.Reset(W, H, 8, 8);
		auto& gl = gpu.IC();
		gl.ClearColor(0.2f, 0.3f, 0.4f);
		gl.ClearDepth(1.0f);
		gl.Clear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
		gl.UseProgram(4);   // Amy
		gl.ViewMatrix(view);
		gl.ProjectionMatrix(proj);
		gl.UseBuffer(0, px, 4); gl.UseBuffer(1, py, 4); gl.UseBuffer(2, pz, 4);
		gl.UseBuffer(9, u, 4); gl.UseBuffer(10, v, 4);
		gl.BindTexture(0, tex.data(), D, D, D, GL_LINEAR_MIPMAP_NEAREST, 2 * D);
		gl.DrawElements(GL_TRIANGLES, 6, GL_UNSIGNED_SHORT, idx);
		gl.UseProgram(1);   // Default post
		gl.StoreColor(out.data(), W, H, W, true);
		gpu.Run();
		gpu.Sync();
		// the same frame through the glow chain, its float canvases on the device: quads + half-size stores,
		// two Kawase passes, the glow combine into host memory (second hash)
		void* quads = gpu.CanvasAlloc(static_cast<size_t>(W / 2) * (H / 2) * 64);
		void* half = gpu.CanvasAlloc(static_cast<size_t>(W / 2) * (H / 2) * 16);
		void* ping = gpu.CanvasAlloc(static_cast<size_t>(W / 2) * (H / 2) * 16);
		gpu.Reset(W, H, 8, 8);
		gl.ClearColor(0.2f, 0.3f, 0.4f);
		gl.ClearDepth(1.0f);
		gl.Clear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
		gl.UseProgram(4);
		gl.ViewMatrix(view);
		gl.ProjectionMatrix(proj);
		gl.UseBuffer(0, px, 4); gl.UseBuffer(1, py, 4); gl.UseBuffer(2, pz, 4);
		gl.UseBuffer(9, u, 4); gl.UseBuffer(10, v, 4);
		gl.BindTexture(0, tex.data(), D, D, D, GL_LINEAR_MIPMAP_NEAREST, 2 * D);
		gl.DrawElements(GL_TRIANGLES, 6, GL_UNSIGNED_SHORT, idx);
		gl.UseProgram(1);
		gl.StoreColorQuadsDevice(quads, W, H, W / 2);
		gl.StoreColorDevice(half, W / 2, H / 2, W / 2, /*downsample=*/true);
		gpu.Run();
		gpu.KawaseBlur(half, W / 2, ping, W / 2, W / 2, H / 2, 0);
		gpu.KawaseBlur(ping, W / 2, half, W / 2, W / 2, H / 2, 1);
		gpu.Glow(quads, W / 2, half, W / 2, true, glow.data(), false, W, H, W);
		gpu.Sync();
		gpu.CanvasFree(quads); gpu.CanvasFree(half); gpu.CanvasFree(ping); }
	catch (const Error& e) {
		std::printf("error %d: %s\n", e.code, e.what());
		return e.code == RSRCU_ERR_NO_DEVICE ? 3 : 1; }
	uint64_t h = 1469598103934665603ull;
	for (uint32_t p : out) { for (int b = 0; b < 4; ++b) { h ^= (p >> (8 * b)) & 0xff; h *= 1099511628211ull; } }
	std::printf("fnv1a %016llx\n", static_cast<unsigned long long>(h));
	h = 1469598103934665603ull;
	for (uint32_t p : glow) { for (int b = 0; b < 4; ++b) { h ^= (p >> (8 * b)) & 0xff; h *= 1099511628211ull; } }
	std::printf("glow %016llx\n", static_cast<unsigned long long>(h));
	return 0; }
