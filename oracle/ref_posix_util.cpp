/* TEST INFRASTRUCTURE -- POSIX stand-ins for the string / path helpers the reference's OBJ loader
 * (src/rgl/rglv/rglv_obj.cxx) calls.  Their own translation units (src/rcl/rclt/rclt_util.cxx,
 * src/rcl/rcls/rcls_file.cxx) include <Windows.h> and cannot be compiled here; these are written from
 * the declared interfaces (rclt_util.hxx:26-35, rcls_file.hxx:20-38) with the semantics the loader relies on:
 * whitespace trimming, "first word / rest" split, directory part of a path, path join.  The parsing itself
 * (vertices, faces, materials, triangulation, normals) and MakeArray stay the reference's compiled code. */
#include "src/rcl/rcls/rcls_file.hxx"
#include "src/rcl/rclt/rclt_util.hxx"

#include <cctype>

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

auto Split1View(std::string_view s) -> std::pair<std::string_view, std::string_view> {
	size_t i = 0;
	while (i < s.size() && !blank(s[i])) { ++i; }
	const std::string_view head = s.substr(0, i);
	while (i < s.size() && blank(s[i])) { ++i; }
	return { head, s.substr(i) }; }

}  // namespace rclt

namespace rcls {

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
