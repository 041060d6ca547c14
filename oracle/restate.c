/* restate.c -- TEST INFRASTRUCTURE: plain-C, scalar, single-threaded restatement of the
 * reference's frame-rendering hot path.  It is the CHECKER, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may build or call it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (a) the reference's own known-answer test (D3D top-left fill rule, rglv_triangle.t.cxx:205-228),
 *   (b) the unmodified reference compiled into oracle/_ref/librsr_ref.so, on seeded scenes, and
 *   (c) the golden frames under tests/golden/ that the reference rendered (make_golden.py).
 *
 * One deliberate difference of FORM: the reference is tiled and multithreaded and 4-wide SSE; this
 * file is a serial loop "for each triangle, for each reference tile it touches, for each 2x2 quad".
 * Tiles are disjoint and each tile consumes its triangles in submission order, so the result is
 * the same.  Every arithmetic step is the reference's, in the reference's order (see citations).
 *
 * rcpps / rsqrtps are taken from tables (arguments), because they are CPU-model specific:
 * rst_harvest_luts() reads them out of this host's instructions.
 *
 * Scope: programs Amy (4), Many (6), OBJ2 (8); depth LESS/LEQUAL/EQUAL, depth write, colour
 * write, alpha blend; clipping; pow2 mip-mapped nearest/bilinear textures; Default post + sRGB /
 * linear true-colour store.  Build: gcc -O2 -msse2 -ffp-contract=off -shared -fPIC.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>

typedef struct RstDraw {
	float vm[16], pm[16];           /* column-major, rmlm::mat4::ff */
	float uniforms[32];
	int program;                    /* 4 Amy, 6 Many, 8 OBJ2 */
	int culling, cullFace;
	int depthTest, depthFunc, depthWrite, colorWrite, blend;
	int width, height;              /* target */
	int tileW, tileH;               /* reference tile size in pixels */
	int vpx, vpy, vpw, vph;         /* viewport (vpw = 0: whole target) */
	const float* buffers[16];
	const float* tex; int texDim; int texFilter;
	const uint32_t* rcpLut; const uint32_t* rsqrtLut;
} RstDraw;

static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* ---- rcpps / rsqrtps table models --------------------------------------------------------- */

void rst_harvest_luts(uint32_t* rcp2048, uint32_t* rsqrt2048) {
	for (uint32_t i = 0; i < 2048; ++i) {
		rcp2048[i] = f2u(_mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(u2f(0x3f800000u | (i << 12)))))); }
	for (uint32_t i = 0; i < 1024; ++i) {
		rsqrt2048[i] = f2u(_mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(u2f(0x3f800000u | (i << 13))))));
		rsqrt2048[1024 + i] = f2u(_mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(u2f(0x40000000u | (i << 13)))))); } }

float rst_rcp(float x, const uint32_t* lut) {
	uint32_t b = f2u(x), sign = b & 0x80000000u, m = b & 0x7fffffu;
	int E = (int)((b >> 23) & 0xff);
	if (E == 0) return u2f(sign | 0x7f800000u);
	if (E == 255) return m ? u2f(b | 0x400000u) : u2f(sign);
	uint32_t r = lut[m >> 12];
	int re = (int)(r >> 23) + (127 - E);
	if (re <= 0) return u2f(sign);
	return u2f(sign | ((uint32_t)re << 23) | (r & 0x7fffffu)); }

float rst_rsqrt(float x, const uint32_t* lut) {
	uint32_t b = f2u(x), sign = b & 0x80000000u, m = b & 0x7fffffu;
	int E = (int)((b >> 23) & 0xff);
	if (E == 255 && m) return u2f(b | 0x400000u);
	if (E == 0) return u2f(sign | 0x7f800000u);
	if (sign) return u2f(0xffc00000u);
	if (E == 255) return 0.0f;
	int e = E - 127, p = e & 1, k = (e - p) >> 1;
	uint32_t r = lut[(p << 10) | (m >> 13)];
	return u2f(r - ((uint32_t)k << 23)); }

/* rmlv::oneover, rmlv_mvec4.hxx:630-650 */
static float oneover(float a, const uint32_t* lut) {
	float r = rst_rcp(a, lut);
	float muls = a * (r * r);
	return (r + r) - muls; }

/* _mm_cvttps_epi32 */
static int cvtt(float f) { return (fabsf(f) < 2147483648.0f) ? (int)f : (int)0x80000000u; }
static float sse_max(float a, float b) { return a > b ? a : b; }
static float sse_min(float a, float b) { return a < b ? a : b; }

/* qmat4 * qfloat4, rmlm_soa.hxx:58-64 */
static void mat_mul(const float* m, const float* v, float* o) {
	for (int r = 0; r < 4; ++r) { o[r] = ((m[r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2]) + m[12 + r] * v[3]; } }

/* mat4 * mat4, rmlm_mat4.hxx:197-207 */
static void mat_mul44(const float* l, const float* r, float* o) {
	for (int row = 0; row < 4; ++row) for (int col = 0; col < 4; ++col) {
		float ax = l[row] * r[col * 4];
		ax += l[4 + row] * r[col * 4 + 1];
		ax += l[8 + row] * r[col * 4 + 2];
		ax += l[12 + row] * r[col * 4 + 3];
		o[col * 4 + row] = ax; } }

/* ---- vertex stage --------------------------------------------------------------------------- */

#define NVMAX 12
typedef struct { float clip[4]; float vary[NVMAX]; } VOut;

static int nvary_of(int program) { return program == 8 ? 11 : 2; }

/* SHADER::Loader::LoadLane + ShadeVertex: Amy shaders.hxx:101-141, Many :525-566, OBJ2 :737-791 */
static void shade_vertex(const RstDraw* d, const float* vpm, int idx, int iid, VOut* o) {
	const float* const* B = d->buffers;
	float p[4] = { B[0] ? B[0][idx] : 0, B[0] ? B[1][idx] : 0, B[0] ? B[2][idx] : 0, 1.0f };
	float n[4] = { B[3] ? B[3][idx] : 0, B[3] ? B[4][idx] : 0, B[3] ? B[5][idx] : 1, 0.0f };
	float kd[3] = { B[6] ? B[6][idx] : 1, B[6] ? B[7][idx] : 1, B[6] ? B[8][idx] : 1 };
	float uv[2] = { B[9] ? B[9][idx] : 0, B[10] ? B[10][idx] : 0 };
	memset(o->vary, 0, sizeof(o->vary));
	if (d->program == 6) {
		float p1[4];
		mat_mul(B[15] + 16 * iid, p, p1);
		o->vary[0] = uv[0]; o->vary[1] = uv[1];
		mat_mul(vpm, p1, o->clip); }
	else if (d->program == 8) {
		mat_mul(d->vm, p, o->vary);          /* sp */
		mat_mul(d->vm, n, o->vary + 4);      /* sn */
		o->vary[8] = kd[0]; o->vary[9] = kd[1]; o->vary[10] = kd[2];
		mat_mul(vpm, p, o->clip); }
	else {
		o->vary[0] = uv[0]; o->vary[1] = uv[1];
		mat_mul(vpm, p, o->clip); } }

/* ViewFrustum::Test (SIMD flavour), rglv_view_frustum.hxx:60-73 */
static int frustum_flags(const float* c, float factor) {
	float w = c[3] * factor;
	return (w + c[0] <= 0) | ((w + c[1] <= 0) << 1) | ((c[3] + c[2] <= 0) << 2) | ((w - c[0] <= 0) << 3) | ((w - c[1] <= 0) << 4); }

typedef struct { float x, y, z, iw; } Dev;

/* pdiv + viewport, rglv_math.hxx:19-23, rglv_gpu.hxx:265-270 */
static Dev to_device(const RstDraw* d, const float* c) {
	int vw = d->vpw > 0 ? d->vpw : d->width, vh = d->vpw > 0 ? d->vph : d->height;
	float DSx = (float)(vw / 2), DSy = (float)(-vh / 2);
	float DOx = (float)(vw / 2 + d->vpx), DOy = (float)(d->height - (vh / 2 + d->vpy));
	float iw = oneover(c[3], d->rcpLut);
	Dev o;
	o.x = (c[0] * iw) * DSx + DOx;
	o.y = (c[1] * iw) * DSy + DOy;
	o.z = c[2] * iw;
	o.iw = iw;
	return o; }

/* ---- texture units, rglr_texture_sampler.cxx --------------------------------------------------- */

static float fract_sse(float a) { return a - (float)cvtt(a); }

static void sample_quad(const RstDraw* d, const float* u, const float* v, float out[4][4]) {
	int POWER = 0;
	while ((1 << (POWER + 1)) <= d->texDim) ++POWER;
	float baseDim = (float)(1 << POWER);
	/* LevelOfDetail :25-42: lanes 1 - 0 */
	float dux = u[1] * baseDim - u[0] * baseDim, dvx = v[1] * baseDim - v[0] * baseDim;
	float sqd = dux * dux + dvx * dvx;
	int lod = ((int)f2u(sqd) - (127 << 23)) >> 24;
	if (lod < 0) lod = 0;
	if (lod > POWER) lod = POWER;
	const float* T = d->tex;
	if (!d->texFilter) {   /* ..._WRAP_NEAREST :74-109 */
		float levelDim = (float)(1 << (POWER - lod));
		int beginRow = (int)((0xfffffffeu << (POWER - lod)) & ((1u << (POWER + 1)) - 1u));
		for (int l = 0; l < 4; ++l) {
			int lx = cvtt(fract_sse(u[l] + 10.0f) * levelDim);
			int ly = cvtt((u2f(0x3f7fffffu) - fract_sse(v[l] + 10.0f)) * levelDim);
			const float* t = T + 4 * (size_t)(((beginRow + ly) << POWER) + lx);
			for (int c = 0; c < 4; ++c) out[c][l] = t[c]; }
		return; }
	/* ..._WRAP_LINEAR :183-286 */
	int levelDimI = 1 << (POWER - lod), mask = levelDimI - 1;
	int lastRow = (int)((0xffffffffu << (POWER - lod)) & ((1u << (POWER + 1)) - 1u)) - 1;
	float levelDim = (float)levelDimI;
	for (int l = 0; l < 4; ++l) {
		float X = u[l] * levelDim, Y = v[l] * levelDim;
		int tx0 = cvtt(X - 0.5f), ty0 = cvtt(Y - 0.5f), tx1 = tx0 + 1, ty1 = ty0 + 1;
		float fx = (X - (float)tx0) - 0.5f, fy = (Y - (float)ty0) - 0.5f;
		float fx1 = 1.0f - fx, fy1 = 1.0f - fy;
		float w00 = fx1 * fy1, w10 = fx * fy1, w01 = fx1 * fy, w11 = fx * fy;
		tx0 &= mask; ty0 &= mask; tx1 &= mask; ty1 &= mask;
		int by0 = lastRow - ty0, by1 = lastRow - ty1;
		const float* p00 = T + 4 * (size_t)((by0 << POWER) + tx0);
		const float* p10 = T + 4 * (size_t)((by0 << POWER) + tx1);
		const float* p01 = T + 4 * (size_t)((by1 << POWER) + tx0);
		const float* p11 = T + 4 * (size_t)((by1 << POWER) + tx1);
		for (int c = 0; c < 4; ++c) out[c][l] = ((p00[c] * w00 + p10[c] * w10) + p01[c] * w01) + p11[c] * w11; } }

/* ---- fragment stage: TriangleProgram::Render, rglv_gpu_impl.hxx:166-222 ------------------------- */

typedef struct { float z[3], iw[3]; float vary[3][NVMAX]; } TriData;

static void shade_fragment(const RstDraw* d, float at[NVMAX][4], float col[4][4]) {
	if (d->program == 6) {   /* Many: shaders.hxx:583 */
		for (int l = 0; l < 4; ++l) { col[0][l] = at[0][l]; col[1][l] = at[1][l]; col[2][l] = d->uniforms[0]; col[3][l] = 1.0f; } }
	else if (d->program == 8) {   /* OBJ2: shaders.hxx:809-818 with qfloat4::operator- (rmlv_soa.hxx:163) */
		for (int l = 0; l < 4; ++l) {
			float dx = 0.0f - at[0][l], dy = 0.0f - at[1][l], dz = 0.0f - at[2][l], dw = 1.0f - at[2][l];
			float d2 = ((dx * dx + dy * dy) + dz * dz) + dw * dw;
			float distance = sqrtf(d2);
			float s = rst_rsqrt(d2, d->rsqrtLut);
			float lx = dx * s, ly = dy * s, lz = dz * s, lw = dw * s;
			float diffuse = sse_max(((at[4][l] * lx + at[5][l] * ly) + at[6][l] * lz) + at[7][l] * lw, 0.1f);
			diffuse = diffuse * (100.0f / distance);
			col[0][l] = at[8][l] * diffuse; col[1][l] = at[9][l] * diffuse; col[2][l] = at[10][l] * diffuse; col[3][l] = 1.0f; } }
	else { sample_quad(d, at[0], at[1], col); } }

static int depth_pass(int func, float frag, float dest) {
	return func == 0 ? frag < dest : (func == 1 ? frag <= dest : frag == dest); }

/* one 2x2 quad at (x, y); e1/e2: edge values per lane; mask bit l = lane covered */
static void render_quad(const RstDraw* d, const TriData* t, float* fb, int x, int y, const int* e1, const int* e2,
                        unsigned mask, float scale, uint64_t* frags) {
	float BSx[4], BSy[4], BSz[4], depth[4];
	float* px[4];
	for (int l = 0; l < 4; ++l) {
		px[l] = fb + 4 * ((size_t)(y + (l >> 1)) * d->width + x + (l & 1));
		BSx[l] = (float)e2[l] * scale;
		BSz[l] = (float)e1[l] * scale;
		BSy[l] = (1.0f - BSx[l]) - BSz[l];
		depth[l] = (BSx[l] * t->z[0] + BSy[l] * t->z[1]) + BSz[l] * t->z[2]; }
	if (d->depthTest) {   /* all three programs are earlyZ */
		for (int l = 0; l < 4; ++l) if ((mask >> l & 1) && !depth_pass(d->depthFunc, depth[l], px[l][3])) mask &= ~(1u << l);
		if (!mask) return; }
	if (d->depthWrite) for (int l = 0; l < 4; ++l) if (mask >> l & 1) px[l][3] = depth[l];
	float at[NVMAX][4];
	int nv = nvary_of(d->program);
	for (int l = 0; l < 4; ++l) {
		float fw = oneover((BSx[l] * t->iw[0] + BSy[l] * t->iw[1]) + BSz[l] * t->iw[2], d->rcpLut);
		float BPx = (t->iw[0] * BSx[l]) * fw, BPz = (t->iw[2] * BSz[l]) * fw, BPy = (1.0f - BPx) - BPz;
		for (int k = 0; k < nv; ++k) at[k][l] = (BPx * t->vary[0][k] + BPy * t->vary[1][k]) + BPz * t->vary[2][k]; }
	float col[4][4];
	shade_fragment(d, at, col);
	if (d->colorWrite) for (int l = 0; l < 4; ++l) if (mask >> l & 1) {
		if (d->blend) {   /* BlendAlpha :80-84 */
			float a = col[3][l], oma = 1.0f - a;
			for (int c = 0; c < 3; ++c) px[l][c] = col[c][l] * a + px[l][c] * oma; }
		else { for (int c = 0; c < 3; ++c) px[l][c] = col[c][l]; } }
	*frags += (uint64_t)__builtin_popcount(mask); }

/* ---- rasteriser: VTriangleRasterizer::Draw (int32, wraps) rglv_triangle.hxx:193-304 and
 *      TriangleRasterizer::Draw (int64 setup) :79-167, for one rectangle ----------------------- */

typedef void (*QuadFn)(void* ctx, int x, int y, const int* e1, const int* e2, unsigned mask, float scale);

static void raster_rect(int wide, int x1, int x2, int x3, int y1, int y2, int y3, int rl, int rt, int rr, int rb,
                        QuadFn fn, void* ctx) {
	int minx = x1 < x2 ? x1 : x2; if (x3 < minx) minx = x3;
	int maxx = x1 > x2 ? x1 : x2; if (x3 > maxx) maxx = x3;
	int miny = y1 < y2 ? y1 : y2; if (y3 < miny) miny = y3;
	int maxy = y1 > y2 ? y1 : y2; if (y3 > maxy) maxy = y3;
	int vminx = minx >> 4; if (vminx < rl) vminx = rl;
	int vmaxx = (maxx + 15) >> 4; if (vmaxx > rr) vmaxx = rr;
	int vminy = miny >> 4; if (vminy < rt) vminy = rt;
	int vmaxy = (maxy + 15) >> 4; if (vmaxy > rb) vmaxy = rb;
	vminx &= ~1; vminy &= ~1;
	int dx12, dy12, dx23, dy23, dx31, dy31, c1, c2, c3;
	float scale;
	dx12 = (int)((uint32_t)x1 - (uint32_t)x2); dy12 = (int)((uint32_t)y2 - (uint32_t)y1);
	dx23 = (int)((uint32_t)x2 - (uint32_t)x3); dy23 = (int)((uint32_t)y3 - (uint32_t)y2);
	dx31 = (int)((uint32_t)x3 - (uint32_t)x1); dy31 = (int)((uint32_t)y1 - (uint32_t)y3);
	if (wide) {
		uint32_t sx = ((uint32_t)vminx << 4) + 8u, sy = ((uint32_t)vminy << 4) + 8u;
		uint32_t u1 = (uint32_t)dy12 * (sx - (uint32_t)x1) + (uint32_t)dx12 * (sy - (uint32_t)y1);
		uint32_t u2 = (uint32_t)dy23 * (sx - (uint32_t)x2) + (uint32_t)dx23 * (sy - (uint32_t)y2);
		uint32_t u3 = (uint32_t)dy31 * (sx - (uint32_t)x3) + (uint32_t)dx31 * (sy - (uint32_t)y3);
		u1 += (uint32_t)(dy12 > 0 || (dy12 == 0 && dx12 > 0)) - 1u;
		u2 += (uint32_t)(dy23 > 0 || (dy23 == 0 && dx23 > 0)) - 1u;
		u3 += (uint32_t)(dy31 > 0 || (dy31 == 0 && dx31 > 0)) - 1u;
		c1 = (int)u1 >> 4; c2 = (int)u2 >> 4; c3 = (int)u3 >> 4;
		scale = 1.0f / (float)(int)((uint32_t)c1 + (uint32_t)c2 + (uint32_t)c3); }
	else {
		int64_t ldx12 = (int64_t)x1 - x2, ldy12 = (int64_t)y2 - y1, ldx23 = (int64_t)x2 - x3, ldy23 = (int64_t)y3 - y2;
		int64_t ldx31 = (int64_t)x3 - x1, ldy31 = (int64_t)y1 - y3;
		int64_t sx = ((int64_t)vminx << 4) + 8, sy = ((int64_t)vminy << 4) + 8;
		int64_t l1 = ldy12 * (sx - x1) + ldx12 * (sy - y1);
		int64_t l2 = ldy23 * (sx - x2) + ldx23 * (sy - y2);
		int64_t l3 = ldy31 * (sx - x3) + ldx31 * (sy - y3);
		if (ldy12 > 0 || (ldy12 == 0 && ldx12 > 0)) l1++; --l1;
		if (ldy23 > 0 || (ldy23 == 0 && ldx23 > 0)) l2++; --l2;
		if (ldy31 > 0 || (ldy31 == 0 && ldx31 > 0)) l3++; --l3;
		l1 >>= 4; l2 >>= 4; l3 >>= 4;
		c1 = (int)l1; c2 = (int)l2; c3 = (int)l3;
		scale = 1.0f / (float)(l1 + l2 + l3); }
	uint32_t r1 = (uint32_t)c1, r2 = (uint32_t)c2, r3 = (uint32_t)c3;
	for (int y = vminy; y < vmaxy; y += 2, r1 += 2u * (uint32_t)dx12, r2 += 2u * (uint32_t)dx23, r3 += 2u * (uint32_t)dx31) {
		uint32_t q1 = r1, q2 = r2, q3 = r3;
		for (int x = vminx; x < vmaxx; x += 2, q1 += 2u * (uint32_t)dy12, q2 += 2u * (uint32_t)dy23, q3 += 2u * (uint32_t)dy31) {
			int e1[4] = { (int)q1, (int)(q1 + (uint32_t)dy12), (int)(q1 + (uint32_t)dx12), (int)(q1 + (uint32_t)dx12 + (uint32_t)dy12) };
			int e2[4] = { (int)q2, (int)(q2 + (uint32_t)dy23), (int)(q2 + (uint32_t)dx23), (int)(q2 + (uint32_t)dx23 + (uint32_t)dy23) };
			int e3[4] = { (int)q3, (int)(q3 + (uint32_t)dy31), (int)(q3 + (uint32_t)dx31), (int)(q3 + (uint32_t)dx31 + (uint32_t)dy31) };
			unsigned mask = 0;
			for (int l = 0; l < 4; ++l) if ((e1[l] | e2[l] | e3[l]) >= 0) mask |= 1u << l;
			if (mask) fn(ctx, x, y, e1, e2, mask, scale); } } }

typedef struct { uint8_t* out; int w; } CovCtx;
static void cov_quad(void* c, int x, int y, const int* e1, const int* e2, unsigned mask, float scale) {
	(void)e1; (void)e2; (void)scale;
	CovCtx* cc = (CovCtx*)c;
	for (int l = 0; l < 4; ++l) if (mask >> l & 1) cc->out[(y + (l >> 1)) * cc->w + x + (l & 1)] = 1; }

/* coverage of one triangle given 28.4 fixed-point vertices; used for the fill-rule KAT */
void rst_raster_coverage(int wide, const int* x3, const int* y3, int rl, int rt, int rr, int rb, int w, int h, uint8_t* out) {
	memset(out, 0, (size_t)w * h);
	CovCtx c = { out, w };
	raster_rect(wide, x3[0], x3[1], x3[2], y3[0], y3[1], y3[2], rl, rt, rr, rb, cov_quad, &c); }

typedef struct { const RstDraw* d; const TriData* t; float* fb; uint64_t* frags; } DrawCtx;
static void draw_quad(void* c, int x, int y, const int* e1, const int* e2, unsigned mask, float scale) {
	DrawCtx* dc = (DrawCtx*)c;
	render_quad(dc->d, dc->t, dc->fb, x, y, e1, e2, mask, scale, dc->frags); }

static void raster_over_tiles(const RstDraw* d, int wide, const int* X, const int* Y, int tx0, int ty0, int tx1, int ty1,
                              const TriData* t, float* fb, uint64_t* frags) {
	DrawCtx dc = { d, t, fb, frags };
	int tilesX = (d->width + d->tileW - 1) / d->tileW, tilesY = (d->height + d->tileH - 1) / d->tileH;
	if (tx0 < 0) tx0 = 0; if (ty0 < 0) ty0 = 0;
	if (tx1 > tilesX - 1) tx1 = tilesX - 1; if (ty1 > tilesY - 1) ty1 = tilesY - 1;
	for (int ty = ty0; ty <= ty1; ++ty) for (int tx = tx0; tx <= tx1; ++tx) {
		int rl = tx * d->tileW, rt = ty * d->tileH;
		int rr = rl + d->tileW < d->width ? rl + d->tileW : d->width;
		int rb = rt + d->tileH < d->height ? rt + d->tileH : d->height;
		raster_rect(wide, X[0], X[1], X[2], Y[0], Y[1], Y[2], rl, rt, rr, rb, draw_quad, &dc); } }

/* ---- clipper: ClipTriangles, rglv_gpu_impl.hxx:678-793 ------------------------------------------ */

static float plane_dist(int plane, const float* c) {
	switch (plane) { case 0: return c[3] + c[0]; case 1: return c[3] + c[1]; case 2: return c[3] + c[2];
	                 case 3: return c[3] - c[0]; default: return c[3] - c[1]; } }

static void clip_and_draw(const RstDraw* d, const VOut* v0, const VOut* v1, const VOut* v2, float* fb, uint64_t* frags) {
	VOut A[10], B[10];
	int na = 3, nv = nvary_of(d->program);
	A[0] = *v0; A[1] = *v1; A[2] = *v2;
	for (int plane = 0; plane < 5 && na > 0; ++plane) {
		int nb = 0;
		int hereIn = plane_dist(plane, A[0].clip) >= 0;
		for (int hi = 0; hi < na; ++hi) {
			int ni = (hi + 1) % na;
			int nextIn = plane_dist(plane, A[ni].clip) >= 0;
			if (hereIn) B[nb++] = A[hi];
			if (hereIn != nextIn) {
				const VOut* from = hereIn ? &A[hi] : &A[ni];
				const VOut* to = hereIn ? &A[ni] : &A[hi];
				float da = plane_dist(plane, from->clip), db = plane_dist(plane, to->clip);
				float t = da / (da - db), omt = 1.0f - t;      /* Distance :86-94, mix rmlv_math.hxx:87-90 */
				VOut n;
				memset(&n, 0, sizeof(n));
				for (int k = 0; k < 4; ++k) n.clip[k] = omt * from->clip[k] + t * to->clip[k];
				for (int k = 0; k < nv; ++k) n.vary[k] = omt * from->vary[k] + t * to->vary[k];
				if (nb < 10) B[nb++] = n;
				hereIn = !hereIn; } }
		memcpy(A, B, sizeof(VOut) * nb);
		na = nb; }
	if (na < 3) return;
	Dev dev[10];
	for (int i = 0; i < na; ++i) dev[i] = to_device(d, A[i].clip);
	float d31x = dev[2].x - dev[0].x, d31y = dev[2].y - dev[0].y, d21x = dev[1].x - dev[0].x, d21y = dev[1].y - dev[0].y;
	int backfacing = (d31x * d21y - d31y * d21x) < 0;
	int willCull = 1;
	if (backfacing) { if (!d->culling || (d->cullFace & 2) == 0) willCull = 0; }
	else { if (!d->culling || (d->cullFace & 1) == 0) willCull = 0; }
	if (willCull) return;
	if (backfacing) for (int i = 0; i < na / 2; ++i) {
		Dev td = dev[i]; dev[i] = dev[na - 1 - i]; dev[na - 1 - i] = td;
		VOut tv = A[i]; A[i] = A[na - 1 - i]; A[na - 1 - i] = tv; }
	for (int f = 1; f < na - 1; ++f) {
		int ids[3] = { 0, f, f + 1 };
		/* ForEachCoveredTile :796-817 */
		int ix[3], iy[3], X[3], Y[3];
		TriData t;
		for (int k = 0; k < 3; ++k) {
			const Dev* p = &dev[ids[k]];
			ix[k] = cvtt(p->x); iy[k] = cvtt(p->y);
			X[k] = cvtt(p->x * 16.0f); Y[k] = cvtt(p->y * 16.0f);    /* DrawClipped :997-998 */
			t.z[k] = p->z; t.iw[k] = p->iw;
			memcpy(t.vary[k], A[ids[k]].vary, sizeof(t.vary[k])); }
		int mnx = ix[0] < ix[1] ? ix[0] : ix[1]; if (ix[2] < mnx) mnx = ix[2];
		int mxx = ix[0] > ix[1] ? ix[0] : ix[1]; if (ix[2] > mxx) mxx = ix[2];
		int mny = iy[0] < iy[1] ? iy[0] : iy[1]; if (iy[2] < mny) mny = iy[2];
		int mxy = iy[0] > iy[1] ? iy[0] : iy[1]; if (iy[2] > mxy) mxy = iy[2];
		int vminx = mnx > 0 ? mnx : 0, vminy = mny > 0 ? mny : 0;
		int vmaxx = mxx + 1 < d->width - 1 ? mxx + 1 : d->width - 1;
		int vmaxy = mxy + 1 < d->height - 1 ? mxy + 1 : d->height - 1;
		raster_over_tiles(d, 0, X, Y, vminx / d->tileW, vminy / d->tileH, vmaxx / d->tileW, vmaxy / d->tileH, &t, fb, frags); } }

/* ---- one draw: BinTriangles* + DrawTriangles, rglv_gpu_impl.hxx:315-675, :841-947 ----------------
 * fb: height*width*4 floats (r, g, b, depth).  indices == NULL: DrawArrays.  instances == 0: not instanced.
 * returns the number of pixels written */
uint64_t rst_draw(const RstDraw* d, int count, const uint16_t* indices, int instances, float* fb) {
	float vpm[16];
	mat_mul44(d->pm, d->vm, vpm);
	int half = (d->width > d->height ? d->width : d->height) / 2;
	float factor = (2048.0f - (float)half) / (float)half;    /* rglv_view_frustum.hxx:36-39 */
	int prims = count / 3, ninst = instances > 0 ? instances : 1;
	uint64_t frags = 0;
	/* unclipped triangles of every instance first, then the clip queue (:498-508) */
	for (int pass = 0; pass < 2; ++pass)
	for (int iid = 0; iid < ninst; ++iid)
	for (int p = 0; p < prims; ++p) {
		int i0 = indices ? indices[3 * p] : 3 * p, i1 = indices ? indices[3 * p + 1] : 3 * p + 1, i2 = indices ? indices[3 * p + 2] : 3 * p + 2;
		VOut v[3];
		shade_vertex(d, vpm, i0, iid, &v[0]); shade_vertex(d, vpm, i1, iid, &v[1]); shade_vertex(d, vpm, i2, iid, &v[2]);
		int cf0 = frustum_flags(v[0].clip, factor), cf1 = frustum_flags(v[1].clip, factor), cf2 = frustum_flags(v[2].clip, factor);
		if (cf0 & cf1 & cf2) continue;
		if (cf0 | cf1 | cf2) { if (pass == 1) clip_and_draw(d, &v[0], &v[1], &v[2], fb, &frags); continue; }
		if (pass == 1) continue;
		Dev a = to_device(d, v[0].clip), b = to_device(d, v[1].clip), c = to_device(d, v[2].clip);
		float d31x = c.x - a.x, d31y = c.y - a.y, d21x = b.x - a.x, d21y = b.y - a.y;
		int front = (d31x * d21y - d31y * d21x) > 0;
		int keepBacks = !(d->culling && d->cullFace == 2), keepFronts = !(d->culling && d->cullFace == 1);
		if (!(front ? keepFronts : keepBacks)) continue;
		int ix[3] = { cvtt(a.x), cvtt(b.x), cvtt(c.x) }, iy[3] = { cvtt(a.y), cvtt(b.y), cvtt(c.y) };
		int mnx = ix[0] < ix[1] ? ix[0] : ix[1]; if (ix[2] < mnx) mnx = ix[2];
		int mxx = ix[0] > ix[1] ? ix[0] : ix[1]; if (ix[2] > mxx) mxx = ix[2];
		int mny = iy[0] < iy[1] ? iy[0] : iy[1]; if (iy[2] < mny) mny = iy[2];
		int mxy = iy[0] > iy[1] ? iy[0] : iy[1]; if (iy[2] > mxy) mxy = iy[2];
		int vminx = mnx > 0 ? mnx : 0, vminy = mny > 0 ? mny : 0;
		int vmaxx = mxx + 1 < d->width - 1 ? mxx + 1 : d->width - 1;
		int vmaxy = mxy + 1 < d->height - 1 ? mxy + 1 : d->height - 1;
		if (!(vmaxx > vminx && vmaxy > vminy)) continue;
		/* back faces are drawn with i0 <-> i2 swapped (:467-470) */
		const Dev* dv[3] = { front ? &a : &c, &b, front ? &c : &a };
		const VOut* vv[3] = { front ? &v[0] : &v[2], &v[1], front ? &v[2] : &v[0] };
		int X[3], Y[3];
		TriData t;
		for (int k = 0; k < 3; ++k) {
			X[k] = cvtt(16.0f * dv[k]->x); Y[k] = cvtt(16.0f * dv[k]->y);      /* :885-886 */
			t.z[k] = dv[k]->z; t.iw[k] = dv[k]->iw;
			memcpy(t.vary[k], vv[k]->vary, sizeof(t.vary[k])); }
		raster_over_tiles(d, 1, X, Y, vminx / d->tileW, vminy / d->tileH, vmaxx / d->tileW, vmaxy / d->tileH, &t, fb, &frags); }
	return frags; }

/* CMD_CLEAR with RB_COLOR_DEPTH, rglv_gpu.cxx:316-322 */
void rst_clear(float* fb, int w, int h, const float* rgb, float depth) {
	for (size_t i = 0; i < (size_t)w * h; ++i) { fb[4 * i] = rgb[0]; fb[4 * i + 1] = rgb[1]; fb[4 * i + 2] = rgb[2]; fb[4 * i + 3] = depth; } }

/* ryg float->sRGB8 table, 3rdparty/ryg-srgb/ryg-srgb.h:71-85 (public domain) */
static const uint32_t tab4[104] = {
	0x0073000d, 0x007a000d, 0x0080000d, 0x0087000d, 0x008d000d, 0x0094000d, 0x009a000d, 0x00a1000d,
	0x00a7001a, 0x00b4001a, 0x00c1001a, 0x00ce001a, 0x00da001a, 0x00e7001a, 0x00f4001a, 0x0101001a,
	0x010e0033, 0x01280033, 0x01410033, 0x015b0033, 0x01750033, 0x018f0033, 0x01a80033, 0x01c20033,
	0x01dc0067, 0x020f0067, 0x02430067, 0x02760067, 0x02aa0067, 0x02dd0067, 0x03110067, 0x03440067,
	0x037800ce, 0x03df00ce, 0x044600ce, 0x04ad00ce, 0x051400ce, 0x057b00c5, 0x05dd00bc, 0x063b00b5,
	0x06970158, 0x07420142, 0x07e30130, 0x087b0120, 0x090b0112, 0x09940106, 0x0a1700fc, 0x0a9500f2,
	0x0b0f01cb, 0x0bf401ae, 0x0ccb0195, 0x0d950180, 0x0e56016e, 0x0f0d015e, 0x0fbc0150, 0x10630143,
	0x11070264, 0x1238023e, 0x1357021d, 0x14660201, 0x156601e9, 0x165a01d3, 0x174401c0, 0x182401af,
	0x18fe0331, 0x1a9602fe, 0x1c1502d2, 0x1d7e02ad, 0x1ed4028d, 0x201a0270, 0x21520256, 0x227d0240,
	0x239f0443, 0x25c003fe, 0x27bf03c4, 0x29a10392, 0x2b6a0367, 0x2d1d0341, 0x2ebe031f, 0x304d0300,
	0x31d105b0, 0x34a80555, 0x37520507, 0x39d504c5, 0x3c37048b, 0x3e7c0458, 0x40a8042a, 0x42bd0401,
	0x44c20798, 0x488e071e, 0x4c1c06b6, 0x4f76065d, 0x52a50610, 0x55ac05cc, 0x5892058f, 0x5b590559,
	0x5e0c0a23, 0x631c0980, 0x67db08f6, 0x6c55087f, 0x70940818, 0x74a007bd, 0x787d076c, 0x7c330723,
};

static uint32_t srgb8(float f) {   /* float_to_srgb8_var2_SSE2, ryg-srgb.h:183-223 */
	float c = sse_max(f, u2f((127u - 13u) << 23));
	c = sse_min(c, u2f(0x3f7fffffu));
	uint32_t bits = f2u(c), tab = tab4[(bits >> 20) - (127u - 13u) * 8u], t = (bits >> 12) & 0xffu;
	return ((tab & 0xffffu) * t + (tab >> 16) * 0x200u) >> 16; }

static uint32_t linear8(float f) {   /* to_tc_basic, rglr_canvas_util.hxx:24-33 */
	float r = sse_min(f, 1.0f);
	r = sse_max(r, 0.0f);
	return (uint32_t)cvtt(r * 255.0f); }

/* FilterTile<DefaultPostProgram, sRGB|LinearColor>, rglr_algorithm.hxx:69-104 */
void rst_store_tc(const float* fb, int w, int h, int gamma, uint32_t* out, int stride) {
	for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
		const float* p = fb + 4 * ((size_t)y * w + x);
		out[(size_t)y * stride + x] = gamma ? (srgb8(p[0]) << 16) | (srgb8(p[1]) << 8) | srgb8(p[2])
		                                    : (linear8(p[0]) << 16) | (linear8(p[1]) << 8) | linear8(p[2]); } }
