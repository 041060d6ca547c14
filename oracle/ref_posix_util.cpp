/* TEST INFRASTRUCTURE -- POSIX stand-ins for the string / path helpers the reference's OBJ loader
 * (src/rgl/rglv/rglv_obj.cxx) calls.  Their own translation units (src/rcl/rclt/rclt_util.cxx,
 * src/rcl/rcls/rcls_file.cxx) include <Windows.h> and cannot be compiled here; these are written from
 * the declared interfaces (rclt_util.hxx:26-35, rcls_file.hxx:11-38) with the semantics the loaders rely on:
 * whitespace trimming, "first word / rest" split, directory part of a path, path join.  The parsing itself
 * (vertices, faces, materials, triangulation, normals) and MakeArray stay the reference's compiled code. */
#include "src/rcl/rcls/rcls_file.hxx"
#include "src/rcl/rclt/rclt_util.hxx"
#include "3rdparty/pixeltoaster/PixelToaster.h"

#include <algorithm>
#include <chrono>
#include <cctype>
#include <fstream>
#include <iterator>
#include <dirent.h>
#include <sys/stat.h>

namespace rqdq {
namespace rclt {

namespace {
bool blank(char ch) { return std::isspace(static_cast<unsigned char>(ch)) != 0; }
}

auto TrimView(std::string_view s) -> std::string_view {
	size_t a = 0, b = s.size();
	while (a < b && blank(s[a])) { ++a; }
	while (b > a && blank(s[b - 1])) { --b; }
	return s.substr(a, b - a); }

auto Trim(std::string_view s) -> std::string { return std::string(TrimView(s)); }

auto Split1View(std::string_view s) -> std::pair<std::string_view, std::string_view> {
	size_t i = 0;
	while (i < s.size() && !blank(s[i])) { ++i; }
	const std::string_view head = s.substr(0, i);
	while (i < s.size() && blank(s[i])) { ++i; }
	return { head, s.substr(i) }; }

}  // namespace rclt

namespace rcls {

/* the scene front-end (mesh / texture stores, JSON files): directory listing in name order, whole-file reads */
auto ListDir(std::string_view dir, std::pmr::memory_resource* mem) -> std::pmr::vector<std::pmr::string> {
	std::pmr::vector<std::pmr::string> out(mem);
	const std::string d(dir);
	if (DIR* h = opendir(d.c_str())) {
		while (dirent* e = readdir(h)) {
			const std::string_view name(e->d_name);
			if (name == "." || name == "..") { continue; }
			out.emplace_back(name); }
		closedir(h); }
	std::sort(out.begin(), out.end());   // (readdir order is arbitrary; both libraries must load the stores in one order)
	return out; }

auto GetMTime(const std::string& path) -> int64_t {
	struct stat st{};
	return stat(path.c_str(), &st) == 0 ? static_cast<int64_t>(st.st_mtime) : 0; }

void LoadBytes(const std::string& path, std::vector<char>& buf) {
	std::ifstream in(path, std::ios::binary);
	buf.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>()); }

auto LoadBytes(const std::string& path) -> std::vector<char> { std::vector<char> b; LoadBytes(path, b); return b; }

void JoinPath(std::string_view a, std::string_view b, std::pmr::string& out) {
	out.assign(a);
	if (!out.empty() && out.back() != '/' && out.back() != '\\') { out.push_back('/'); }
	out.append(b); }

auto JoinPath(std::string a, const std::string& b) -> std::string {
	if (!a.empty() && a.back() != '/' && a.back() != '\\') { a.push_back('/'); }
	return a + b; }

auto DirNameView(std::string_view fn) -> std::string_view {
	const size_t cut = fn.find_last_of("/\\");
	return cut == std::string_view::npos ? std::string_view{} : fn.substr(0, cut + 1); }

auto JoinPath(std::string_view a, std::string_view b, std::pmr::memory_resource* mem) -> std::pmr::string {
	std::pmr::string out(a, mem);
	if (!out.empty() && out.back() != '/' && out.back() != '\\') { out.push_back('/'); }
	if (!b.empty() && b.front() == '/') { out.assign(b); } else { out.append(b); }
	return out; }

auto SplitPath(std::pmr::string p) -> std::pair<std::pmr::string, std::pmr::string> {
	const size_t cut = p.find_last_of("/\\");
	if (cut == std::pmr::string::npos) { return { std::pmr::string{}, p }; }
	std::pmr::string head = p.substr(0, cut + 1), tail = p.substr(cut + 1);
	while (head.size() > 1 && (head.back() == '/' || head.back() == '\\')) { head.pop_back(); }
	return { head, tail }; }

}  // namespace rcls
}  // namespace rqdq

/* PixelToaster's platform layer (window, timer) is not built; `$particles` holds a PixelToaster::Timer
 * (node/particles.cxx:49), whose factory is this */
namespace PixelToaster {
namespace {
class ChronoTimer final : public TimerInterface {
	std::chrono::steady_clock::time_point t0_{std::chrono::steady_clock::now()}, last_{t0_};
	static double secs(std::chrono::steady_clock::duration d) { return std::chrono::duration<double>(d).count(); }
public:
	void reset() override { t0_ = last_ = std::chrono::steady_clock::now(); }
	double time() override { return secs(std::chrono::steady_clock::now() - t0_); }
	double delta() override { const auto now = std::chrono::steady_clock::now(); const double d = secs(now - last_); last_ = now; return d; }
	double resolution() override { return 1e-9; }
	void wait(double) override {} };
}
TimerInterface* createTimer() { return new ChronoTimer(); }
}  // namespace PixelToaster
