/* rsrcu.h -- C ABI of the B200 (sm_100a) rasteriser that replaces rsr's software `rglv::GPU`.
 *
 * Boundary: the reference records GL calls into `rglv::GL` (src/rgl/rglv/rglv_gl.hxx:182-344) and
 * `rglv::GPU::RunImpl` (src/rgl/rglv/rglv_gpu.cxx:90-116) decodes that stream, bins triangles
 * (`BinImpl` :119-260) and draws tiles (`DrawImpl` :263-432).  A drop-in replaces exactly
 * RunImpl: it walks the same command stream and, per command, calls one function below.
 * INTEGRATION.md shows the ~60-line RunImpl replacement a maintainer would add.
 *
 * Conventions
 *  - plain C, no exceptions cross the boundary; every call returns RSRCU_OK (0) or an error code,
 *    `rsrcu_last_error()` returns a thread-local human-readable message for the last failure.
 *  - one context per `rglv::GPU`; calls on one context are serialised by the caller (the reference
 *    runs RunImpl on a single job thread); contexts are independent and may live on different
 *    devices (split-frame sharding: one context per GPU).
 *  - host pointers are borrowed only for the duration of the call (data is staged/uploaded before
 *    the call returns) unless a call says otherwise (store destinations are written by
 *    `rsrcu_end_frame`/`rsrcu_sync`).
 *  - matrices are 16 floats, column-major, exactly `rmlm::mat4::ff` (src/rml/rmlm/rmlm_mat4.hxx:14).
 *  - there is no CPU fallback: if no CUDA device is present `rsrcu_create` fails.
 */
#ifndef RSRCU_H
#define RSRCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSRCU_OK                 0
#define RSRCU_ERR_NO_DEVICE      1   /* no CUDA device / driver */
#define RSRCU_ERR_CUDA           2   /* a CUDA runtime call failed */
#define RSRCU_ERR_INVALID        3   /* bad argument or call order */
#define RSRCU_ERR_NO_PROGRAM     4   /* (program id, fragment state key) not in the dispatch table;
                                        the reference calls std::exit(1) here (rglv_gpu.cxx:199-202) */
#define RSRCU_ERR_UNSUPPORTED    5   /* state the reference itself cannot render (e.g. >2048 px) */
#define RSRCU_ERR_OVERFLOW       6   /* a device-side buffer still overflows after the library grew it and rendered
                                        the frame again several times (rsrcu_sync / rsrcu_sync_frame retry internally) */

/* constants: same numeric values as src/rgl/rglv/rglv_gl.hxx:20-62 */
#define RSRCU_GL_FRONT 1
#define RSRCU_GL_BACK 2
#define RSRCU_GL_NEAREST_MIPMAP_NEAREST 0
#define RSRCU_GL_LINEAR_MIPMAP_NEAREST 1
#define RSRCU_GL_COLOR_BUFFER_BIT 1
#define RSRCU_GL_DEPTH_BUFFER_BIT 2
#define RSRCU_GL_STENCIL_BUFFER_BIT 4
#define RSRCU_GL_LESS 0
#define RSRCU_GL_LEQUAL 1
#define RSRCU_GL_EQUAL 2
#define RSRCU_RB_COLOR_DEPTH 0
#define RSRCU_RB_RGBF32 1
#define RSRCU_RB_RGBAF32 2
#define RSRCU_RB_F32 3
#define RSRCU_HINT_READ4 1
#define RSRCU_HINT_DENSE 2

/* upload policy for rsrcu_bind_buffer / rsrcu_bind_texture / index data */
#define RSRCU_UPLOAD_ALWAYS 0   /* contents may have changed since the last call: copy again */
#define RSRCU_UPLOAD_STATIC 1   /* (pointer, size) identifies immutable data: copy once, then reuse */
#define RSRCU_UPLOAD_FRAME 3    /* contents stay as they are until rsrcu_end_frame -- the reference's own contract (GL records
                                   pointers, the renderer reads them at Run): staged once per frame however often the
                                   pointer is bound (the drop-in binding's default) */
#define RSRCU_UPLOAD_DEVICE 2   /* the pointer IS device memory on the context's device (a canvas of rsrcu_canvas_alloc,
                                   the output of rsrcu_march_surface, any CUDA allocation): used in place, nothing is
                                   copied.  The caller orders producer and consumer (same context: stream order;
                                   another context: rsrcu_wait_for) */

typedef struct rsrcu_ctx rsrcu_ctx;

/* POD mirror of `rglv::GLState` (src/rgl/rglv/rglv_gl.hxx:80-179) without the raw pointers:
 * vertex buffers, textures and uniforms are attached with the bind calls below and are captured
 * together with this struct by the next draw / clear / store, like `GL::MaybeUpdateState`
 * (src/rgl/rglv/rglv_gl.cxx:100-106) captures `cs_`. */
typedef struct RsrState {
	float clear_color[4];         /* GLState::clearColor */
	float clear_depth;            /* GLState::clearDepth */
	int32_t culling_enabled;      /* GL_CULL_FACE */
	int32_t cull_face;            /* RSRCU_GL_FRONT / RSRCU_GL_BACK (3 = both; see DESIGN.md quirk list) */
	int32_t scissor_enabled;      /* no program of the reference's table is installed with scissor */
	int32_t scissor_origin[2], scissor_size[2];
	int32_t viewport_origin[2];
	int32_t viewport_size[2];     /* {0,0} = unset => target size (std::optional in the reference) */
	int32_t blending_enabled;
	int32_t color_write_mask;
	int32_t depth_write_mask;
	int32_t depth_test_enabled;
	int32_t depth_func;           /* RSRCU_GL_LESS / LEQUAL / EQUAL */
	int32_t program_id;           /* ids of src/viewer/shaders*.hxx: Amy 4, Many 6, OBJ2 8, ... */
	int32_t color0_attachment_type; /* RSRCU_RB_* */
	int32_t depth_attachment_type;
	float view_matrix[16];
	float projection_matrix[16];
	float normal_matrix[16];      /* carried for API parity; like the reference, shading uses
	                                 transpose(inverse(view)) (rglv_gpu_impl.hxx:45-51) */
	uint32_t uniforms_valid;      /* 0 = UseUniforms(-1) */
	float uniforms[32];           /* one UNIFORM_BUFFER_SIZE block (rglv_gl.hxx:62) */
} RsrState;

/* ---- lifetime ------------------------------------------------------------------------------ */

/* Creates a context on CUDA device `device`.  Harvests this host CPU's rcpps / rsqrtps tables
 * (the reference's `oneover` / `normalize` depend on them: rmlv_mvec4.hxx:630-650,
 * rmlv_soa.hxx:243-248), verifies the table model against the instruction, uploads them. */
int rsrcu_create(int device, rsrcu_ctx** out);
int rsrcu_destroy(rsrcu_ctx* ctx);
const char* rsrcu_last_error(void);

/* Overrides the harvested approximation tables (to reproduce a frame rendered by the reference on
 * another CPU, e.g. a committed golden image).  rcp: 2048 entries = bits of rcpps(1 + i/2048);
 * rsqrt: 2 x 1024 entries = bits of rsqrtps(1 + i/1024) then rsqrtps(2 * (1 + i/1024)). */
int rsrcu_set_host_luts(rsrcu_ctx* ctx, const uint32_t* rcp2048, const uint32_t* rsqrt2x1024);
int rsrcu_get_host_luts(rsrcu_ctx* ctx, uint32_t* rcp2048, uint32_t* rsqrt2x1024);

/* Drops every RSRCU_UPLOAD_STATIC allocation (device copies of meshes / textures).  Static data is
 * keyed by host pointer + byte size and must stay alive and unchanged while cached; call this
 * before freeing or rewriting such memory. */
int rsrcu_release_static(rsrcu_ctx* ctx);

/* ---- frame recording: one call per reference stream command ---------------------------------- */

/* GPU::Reset (rglv_gpu.cxx:50-56).  tile_*_blocks are the reference's tile size in 8x8 blocks
 * (default 8x8 => 64x64 px); the device uses its own 32x32 tiles and only needs these to
 * reproduce the reference's int32 edge-function start point.  width,height <= 2048 and even. */
int rsrcu_begin_frame(rsrcu_ctx* ctx, int width, int height, int tile_w_blocks, int tile_h_blocks);

/* CMD_STATE (rglv_gpu.cxx:139-147) */
int rsrcu_set_state(rsrcu_ctx* ctx, const RsrState* state);

/* GLState::buffers[slot] (rglv_gl.hxx:104): SoA float arrays, slots 0-2 position, 3-5 normal,
 * 6-8 colour/kd, 9-10 uv, 15 per-instance mat4 array.  n_floats makes the extent explicit (the
 * reference scans the index buffer instead, rglv_gpu_impl.hxx:332-334).  NULL unbinds. */
int rsrcu_bind_buffer(rsrcu_ctx* ctx, int slot, const float* host_ptr, size_t n_floats, int upload);

/* GLState::tus[unit] (rglv_gl.hxx:64-77,106): RGBA32F texels, `height` rows of the base level;
 * for power-of-two square textures the mip chain is stacked below (2*height rows in memory,
 * rglr_texture.cxx:33-81) and must be included: rows_in_memory says how many rows to upload. */
int rsrcu_bind_texture(rsrcu_ctx* ctx, int unit, const float* host_ptr, int width, int height,
                       int stride, int filter, int rows_in_memory, int upload);

/* GLState::tu3ptr/tu3dim: float32 depth (shadow) map, dim x dim */
int rsrcu_bind_depth_texture(rsrcu_ctx* ctx, const float* host_ptr, int dim, int upload);

/* CMD_CLEAR (rglv_gpu.cxx:148-155, :311-344) */
int rsrcu_clear(rsrcu_ctx* ctx, int bits);

/* CMD_DRAW_ELEMENTS / _INSTANCED (rglv_gpu.cxx:216-243): count = number of indices (3 per
 * triangle), uint16 indices, instance_count = 0 selects the non-instanced path. */
int rsrcu_draw_elements(rsrcu_ctx* ctx, int count, const uint16_t* indices, int hint,
                        int instance_count, int upload);

/* CMD_DRAW_ARRAYS / _INSTANCED (rglv_gpu.cxx:192-215) */
int rsrcu_draw_arrays(rsrcu_ctx* ctx, int count, int instance_count);

/* CMD_STORE_COLOR_FULL_LINEAR_TC (rglv_gpu.cxx:177-185, :384-395): post program of the current
 * state + sRGB/linear conversion to 0x00RRGGBB.  dst = host memory (pinned or pageable), written
 * when the frame completes; dst == NULL keeps the result on the device only
 * (see rsrcu_device_truecolor). */
int rsrcu_store_color_tc(rsrcu_ctx* ctx, int enable_gamma, uint32_t* dst, int width, int height,
                         int stride_px);

/* Same store, but the destination is DEVICE memory owned by the caller (e.g. a torch tensor that is
 * then gathered to the presenting GPU over NCCL): the tile kernel resolves straight into it. */
int rsrcu_store_color_tc_device(rsrcu_ctx* ctx, int enable_gamma, void* device_dst, int width, int height,
                                int stride_px);

/* Split-frame presentation over NVLink: lets this context's kernels store into memory that lives on
 * `peer_device` (cudaDeviceEnablePeerAccess; already-enabled is not an error).  The destination of
 * rsrcu_store_color_tc_device may then be a (CUDA-IPC-opened or same-process) pointer into the presenting
 * GPU's frame buffer: the tile kernel's resolve writes travel as peer stores while it rasterises, so
 * there is no separate gather step. */
int rsrcu_enable_peer_access(rsrcu_ctx* ctx, int peer_device);

/* Split-frame completion without a host-side barrier or a collective: 64-bit counters in device memory (of this or --
 * after rsrcu_enable_peer_access -- of another GPU, 8-byte aligned), one per rank, as stream operations.
 * rsrcu_signal_counter enqueues on this context's stream "add 1 to *device_counter", executed once every frame
 * submitted so far has completed and all its stores, peer stores included, are visible system-wide.
 * rsrcu_wait_counters enqueues a wait until each of the `count` consecutive counters has reached `value`.
 * Each rank signals its own counter behind its units of frame f; a rank waits for "all >= f - 1" before it reuses a
 * buffer of the double-buffered presented frame, the presenter for "all >= f + 1" before it presents frame f.
 * A wait gives up after 2 s (a rank that died must not hang the GPU). */
int rsrcu_signal_counter(rsrcu_ctx* ctx, void* device_counter);
int rsrcu_wait_counters(rsrcu_ctx* ctx, const void* device_counters, int count, uint64_t value);

/* CMD_STORE_COLOR_FULL_LINEAR_FP (half = 0; Copy, rglr_algorithm.cxx:247-279) and
 * CMD_STORE_COLOR_HALF_LINEAR_FP (half = 1; Downsample, rglr_algorithm.cxx:118-141: one pixel per
 * 2x2 quad, ((p0 + p1) + p2) + p3 times 0.25) (rglv_gpu.cxx:156-169, :345-370): RGBA32F pixels,
 * alpha = 0 as in the reference.  width/height describe dst: the target's size, or half of it. */
int rsrcu_store_color_fp(rsrcu_ctx* ctx, float* dst, int width, int height, int stride_px, int half);

/* CMD_STORE_COLOR_FULL_QUADS_FP (rglv_gpu.cxx:170-176, :371-383; Copy, rglr_algorithm.cxx:320-368):
 * dst is a QFloat4Canvas -- 64 bytes per 2x2 quad {r[4], g[4], b[4], a[4]}, lanes = (x,y), (x+1,y),
 * (x,y+1), (x+1,y+1); stride in quads.  The fourth plane is what the reference's tile buffer holds:
 * depth for RB_COLOR_DEPTH, 1.0 for RB_RGBF32, the clear colour's alpha for RB_RGBAF32. */
int rsrcu_store_color_quads(rsrcu_ctx* ctx, float* dst, int width, int height, int stride_quads);

/* CMD_STORE_DEPTH_FULL_LINEAR_FP (rglv_gpu.cxx:186-191, :396-405).  The reference implements it
 * for RB_F32 depth attachments only; here RB_COLOR_DEPTH (depth in alpha) is accepted too. */
int rsrcu_store_depth(rsrcu_ctx* ctx, float* dst);

/* CMD_EOF: uploads the recorded frame, enqueues every kernel and the device->host copies of the
 * store destinations on the context's stream, and returns without waiting. */
int rsrcu_end_frame(rsrcu_ctx* ctx);

/* Waits for every submitted frame.  A frame whose tile lists, clip records or large-item queue did not fit is
 * never delivered truncated: the buffer is grown and the frame is launched again from its device-resident tables
 * (RsrStats::frames_retried counts those launches). */
int rsrcu_sync(rsrcu_ctx* ctx);

/* Retained frames.  The reference records every frame anew; a caller that submits the same recorded
 * frame again and again (a static scene, or the sub-frames of a split frame whose camera did not move)
 * can keep the frame's tables on the device instead: rsrcu_retain_frame, called after rsrcu_end_frame /
 * rsrcu_run_stream and before the next rsrcu_begin_frame, snapshots the last submitted frame -- state and
 * draw tables, per-frame (UPLOAD_ALWAYS) data, clear / store commands (at most 6) with their destinations --
 * into an rsrcu_frame; rsrcu_replay_frame launches its kernels without decoding, table building or upload.
 * Store destinations are the ones recorded: host destinations are written again by every replay (wait
 * with rsrcu_sync / rsrcu_sync_frame before reading them), device-resident results stay in the same buffer. */
typedef struct rsrcu_frame rsrcu_frame;
int rsrcu_retain_frame(rsrcu_ctx* ctx, rsrcu_frame** out);
int rsrcu_replay_frame(rsrcu_ctx* ctx, rsrcu_frame* frame);
int rsrcu_release_frame(rsrcu_ctx* ctx, rsrcu_frame* frame);

/* Frame overlap (off by default).  When on, the front end of a frame (upload, vertex, setup, binning: latency-bound
 * kernels that leave most of the GPU idle) runs on a second, high-priority stream into its own set of intermediate
 * buffers, and only the tile kernel runs on the context's stream: the front end of frame N+1 executes while the tile
 * kernel of frame N is still rasterising -- the counterpart of the reference binning frame N+1 on the main thread
 * while the tile jobs of frame N run (doubleBuffer, rglv_gpu.cxx:16,111-112).  The tile kernels of consecutive frames
 * alternate between the context's stream and a second one (the tail of one and the head of the next share the GPU):
 * call rsrcu_join before enqueueing work of your own on rsrcu_stream() that must come after every submitted frame
 * (rsrcu_sync, rsrcu_sync_frame and rsrcu_signal_counter order themselves). */
int rsrcu_set_overlap(rsrcu_ctx* ctx, int enabled);
/* orders the stream of rsrcu_stream() behind every frame submitted so far (no host wait) */
int rsrcu_join(rsrcu_ctx* ctx);

/* Frames are pipelined (upload arena double-buffered; store targets, counters and their device->host
 * copies in a ring of three; copies run on a second stream): rsrcu_sync_frame(ctx, lag) waits only
 * until the frame `lag` submissions before the most recent one (lag 0 = the latest, at most 2) has
 * landed in its store destinations, so frame N's read-back overlaps the recording and the kernels of
 * the frames after it -- the GPU counterpart of the reference's doubleBuffer mode
 * (rglv_gpu.cxx:16,111-112).  A store destination must not be reused before its frame has been
 * waited for.  Like rsrcu_sync it launches a frame again whose device-side buffers overflowed (the frame's
 * tables stay on the device until three frames later), so the destination always holds the complete frame. */
int rsrcu_sync_frame(rsrcu_ctx* ctx, int lag);

/* ---- packed command stream ---------------------------------------------------------------------
 * The reference hands `GPU::RunImpl` one byte-packed command stream per frame
 * (`FastPackedStream`, src/rgl/rglv/rglv_packed_stream.hxx:14-131; opcodes rglv_gpu_protocol.hxx:7-53).
 * rsrcu_run_stream is the same idea for this ABI: one call decodes a whole recorded frame
 * (begin_frame ... end_frame) without per-command FFI overhead.  Records are 8-byte aligned:
 *   uint32 opcode, uint32 record_bytes (header included), payload
 * payloads (all fields little-endian, pointers as uint64 host addresses, borrowed for the call):
 *   RSRCU_OP_BEGIN_FRAME   int32 width, height, tile_w_blocks, tile_h_blocks
 *   RSRCU_OP_STATE         RsrState
 *   RSRCU_OP_BIND_BUFFER   int32 slot, upload; uint64 ptr; uint64 n_floats
 *   RSRCU_OP_BIND_TEXTURE  int32 unit, width, height, stride, filter, rows_in_memory, upload, pad; uint64 ptr
 *   RSRCU_OP_BIND_DEPTH    int32 dim, upload; uint64 ptr
 *   RSRCU_OP_CLEAR         int32 bits, pad
 *   RSRCU_OP_DRAW_ELEMENTS int32 count, hint, instance_count, upload; uint64 ptr
 *   RSRCU_OP_DRAW_ARRAYS   int32 count, instance_count
 *   RSRCU_OP_STORE_TC      int32 gamma, width, height, stride_px; uint64 ptr
 *   RSRCU_OP_STORE_FP      int32 half, width, height, stride_px; uint64 ptr
 *   RSRCU_OP_STORE_DEPTH   uint64 ptr
 *   RSRCU_OP_END_FRAME     (no payload)
 *   RSRCU_OP_STORE_TC_DEV  int32 gamma, width, height, stride_px; uint64 device ptr
 *   RSRCU_OP_STORE_QUADS   int32 pad, width, height, stride_quads; uint64 ptr
 *   RSRCU_OP_STORE_FP_DEV / _QUADS_DEV / _DEPTH_DEV   same payloads as STORE_FP / STORE_QUADS / STORE_DEPTH, device ptr
 */
#define RSRCU_OP_BEGIN_FRAME 1
#define RSRCU_OP_STATE 2
#define RSRCU_OP_BIND_BUFFER 3
#define RSRCU_OP_BIND_TEXTURE 4
#define RSRCU_OP_BIND_DEPTH 5
#define RSRCU_OP_CLEAR 6
#define RSRCU_OP_DRAW_ELEMENTS 7
#define RSRCU_OP_DRAW_ARRAYS 8
#define RSRCU_OP_STORE_TC 9
#define RSRCU_OP_STORE_FP 10
#define RSRCU_OP_STORE_DEPTH 11
#define RSRCU_OP_END_FRAME 12
#define RSRCU_OP_STORE_TC_DEV 13
#define RSRCU_OP_STORE_QUADS 14
#define RSRCU_OP_STORE_FP_DEV 15
#define RSRCU_OP_STORE_QUADS_DEV 16
#define RSRCU_OP_STORE_DEPTH_DEV 17
int rsrcu_run_stream(rsrcu_ctx* ctx, const void* stream, size_t bytes);

/* ---- device-side access (viewer presents from the device-resolved buffer; bench; sharding) --- */

/* device pointer + pitch (pixels) of the last true-colour store of the current/last frame */
int rsrcu_device_truecolor(rsrcu_ctx* ctx, void** dev_ptr, int* stride_px);
/* the CUDA stream (cudaStream_t) all work of this context is enqueued on */
int rsrcu_stream(rsrcu_ctx* ctx, void** stream);

/* ---- device canvases: render-to-texture, shadow maps and glow chains that never leave the device ----
 * The reference's node graph passes canvases between nodes in host memory: a `$layer` renders its lights' depth maps
 * with a second rglv::GPU and StoreDepth (src/viewer/node/gllayer.cxx:154-181) before the main pass samples them;
 * `$buffers` exposes StoreColor results (quads / half-size linear, node/buffers.cxx), `$kawase` blurs them
 * (node/kawase.cxx:83-129), `$glow` combines image + blur into the true-colour output (node/glow.cxx:146-160),
 * `$rendertotexture` feeds a frame back as a texture.  Here those canvases are device memory: the store commands take
 * a device destination, the bind calls take a device source (RSRCU_UPLOAD_DEVICE), the post filters are kernels on the
 * context's stream, and only the final true-colour frame crosses PCIe. */
int rsrcu_canvas_alloc(rsrcu_ctx* ctx, size_t bytes, void** device_ptr);   /* 256-byte aligned, zero-filled */
int rsrcu_canvas_free(rsrcu_ctx* ctx, void* device_ptr);
/* blocking copies, ordered behind everything submitted to the context so far (tests, fixtures, debugging) */
int rsrcu_canvas_read(rsrcu_ctx* ctx, const void* device_ptr, void* host_dst, size_t bytes);
int rsrcu_canvas_write(rsrcu_ctx* ctx, void* device_ptr, const void* host_src, size_t bytes);

/* rsrcu_store_color_fp / _quads / rsrcu_store_depth with a DEVICE destination (same layouts, same arguments) */
int rsrcu_store_color_fp_device(rsrcu_ctx* ctx, void* device_dst, int width, int height, int stride_px, int half);
int rsrcu_store_color_quads_device(rsrcu_ctx* ctx, void* device_dst, int width, int height, int stride_quads);
int rsrcu_store_depth_device(rsrcu_ctx* ctx, void* device_dst);

/* Orders `ctx` behind `producer` (two contexts on the same device, e.g. the shadow-map context of a `$layer` and the
 * main one): everything submitted to `producer` so far completes before anything submitted to `ctx` from now on
 * starts.  No host wait.  (jobsys::add_link(gpu_.Run(), lightJobs[li]) ... jobsys::wait, gllayer.cxx:177-187) */
int rsrcu_wait_for(rsrcu_ctx* ctx, rsrcu_ctx* producer);

/* rglr::KawaseBlurFilter (src/rgl/rglr/rglr_kawase.cxx:22-81): dst(x, y) = 1/16 of the sum of four 2x2 boxes at the
 * offsets (-d-1,-d-1) (d,-d-1) (-d-1,d) (d,d), coordinates clamped to the canvas; RGBA32F linear canvases, strides in
 * pixels, src != dst.  `$kawase` with intensity N = N calls with dist 0 .. N-1 ping-ponging two canvases. */
int rsrcu_kawase_blur(rsrcu_ctx* ctx, const void* src_device, int src_stride_px, void* dst_device, int dst_stride_px,
                      int width, int height, int dist);

/* rglr::Texture::maybe_make_mipmap (src/rgl/rglr/rglr_texture.cxx:33-81), what `$renderToTexture` does to a power-of-two
 * square target before it is sampled (node/rendertotexture.cxx:91): texels_device holds 2 * dim rows of dim RGBA32F
 * texels, the base level in the first dim rows; the chain is written below it (each level ((a + b) + c) + d of the
 * 2x2 texels above, / 4).  Stream-ordered on the context. */
int rsrcu_make_mipmap(rsrcu_ctx* ctx, void* texels_device, int dim);

/* `$glow` (node/glow.cxx:24-39, :146-160; rglr::Filter<GlowShader, sRGB|LinearColor>, rglr_algorithm.hxx:107-144):
 * out = (image + blur * 0.7) * 0.5 per quad, converted to 0x00RRGGBB.  image: quad-swizzled canvas of the frame
 * (rsrcu_store_color_quads[_device]); blur: linear RGBA32F canvas read at (x/2, y/2) -- one blur pixel per quad, as
 * the reference does.  dst: host memory (written when the call returns) or, with dst_is_device, device memory
 * (stream-ordered; rsrcu_device_truecolor then reports it). */
int rsrcu_glow(rsrcu_ctx* ctx, const void* image_quads_device, int image_stride_quads, const void* blur_device,
               int blur_stride_px, int enable_gamma, uint32_t* dst, int dst_is_device, int width, int height, int stride_px);

/* ---- geometry produced on the device -----------------------------------------------------------------
 * `$mc` (src/viewer/node/mc.cxx:230-300 + rglv::march_sdf_vao, src/rgl/rglv/rglv_marching_cubes.hxx:66-109): marching
 * cubes over the node's signed-distance field (sphere of radius 3 + sine distortion, mc.cxx:95-110) on a
 * precision^3 grid over [-range, range]^3, split into 8^fork_depth blocks like the reference's jobs.  The vertices
 * (position SoA in slots 0-2, normal SoA in slots 3-5, three per triangle, DrawArrays order) are written to device
 * memory owned by the context and never cross PCIe: bind them with RSRCU_UPLOAD_DEVICE and rsrcu_draw_arrays.
 * Blocks are emitted in the reference's block order and cells in its y / z / x order, so the triangle sequence is
 * the one the reference draws (its jobs append to per-block arrays that are drawn in allocation order; here the
 * order is deterministic).  Positions are bit-identical to the reference's; normals go through libm's sinf there and
 * CUDA's sinf here and agree to about 1e-6.  out_soa6: device pointers to x, y, z, nx, ny, nz (vertex_total floats
 * each; valid until the third call from now on this context); blocks: per non-empty block its first vertex (a
 * multiple of 4: the reference pads every block's arrays for its 4-wide loader, the padding is zeros) and its vertex
 * count -- the reference issues one DrawArrays per block (mc.cxx:212-226).  Blocking (the counts come back to the
 * host).  1 <= precision >> fork_depth <= 32. */
typedef struct RsrMarchBlock { int32_t first_vertex, vertex_count; } RsrMarchBlock;
int rsrcu_march_surface(rsrcu_ctx* ctx, float time_seconds, int precision, int fork_depth, float range,
                        const float** out_soa6, RsrMarchBlock* blocks, int block_capacity, int* block_count, int* vertex_total);

/* ---- presentation + telemetry --------------------------------------------------------------------------
 * The viewer draws a per-thread timeline of job spans over the frame (render_jobsys, src/viewer/jobsys_vis.cxx:26-90:
 * one 8-pixel bar per lane, 2-pixel gap, brightness ramp 1 -> 0 along the span, colours picked by hashed bits) and
 * hands the canvas to the window.  Here the lanes are the frame's pipeline stages timed with CUDA events instead of
 * jobsys::measurements_pt, the bars are drawn by a kernel into the device-resolved true-colour buffer, and the
 * presentation step is a device-to-device blit into the presentation surface (the interop-mapped swap-chain image in a
 * viewer; any device allocation here). */
typedef struct RsrSpan { double start, end; uint32_t raw; int32_t lane; } RsrSpan;   /* jobsys::JobStat (rclmt_jobsys.hxx:94-97) + its lane */
int rsrcu_draw_spans(rsrcu_ctx* ctx, void* truecolor_device, int stride_px, int width, int height,
                     int left, int top, float xscale, const RsrSpan* spans, int count);
/* the spans of the context's last profiled frame (rsrcu_set_profiling level 2): one lane per stage, raw = stage index */
int rsrcu_frame_spans(rsrcu_ctx* ctx, RsrSpan* out, int capacity, int* count);
int rsrcu_present(rsrcu_ctx* ctx, const void* truecolor_device, int src_stride_px, void* surface_device,
                  int surface_stride_px, int width, int height);

/* RSRCU_UPLOAD_FRAME arrays of 1 MiB or more copied IN PLACE (opt-in; also RSRCU_PIN_IN_PLACE=1): instead of a staging
 * memcpy on the submitting thread the array's pages are locked where they lie (cudaHostRegister, once per pointer)
 * and the copy engine reads them directly, once per frame; rsrcu_end_frame returns when the copies have left host
 * memory.  ONLY for hosts whose large arrays outlive their last use by the context: memory that is freed while it is
 * still registered poisons later copies to or from whatever is allocated there next (this is why it is not the
 * default: the reference only promises its pointers until a frame's Finalize).  Only the pages wholly inside an array
 * are locked (its first and last page may be shared with heap neighbours, whose own copies a locked page would break);
 * the two edge fragments travel as pageable memory.  A copy the copy engine refuses falls back to staging.  Disabling
 * it waits for everything submitted and releases all pages.  A frame that used it cannot be retained. */
int rsrcu_set_pin_in_place(rsrcu_ctx* ctx, int enabled);

/* counters of the last completed frame (filled by rsrcu_sync) */
typedef struct RsrStats {
	uint64_t triangles_submitted;   /* sum over draws of prims x instances */
	uint64_t triangles_binned;      /* accepted after cull / frustum */
	uint64_t triangles_clipped;     /* sent through the clipper */
	uint64_t bin_entries;           /* (triangle, tile) pairs */
	uint64_t fragments_shaded;      /* pixels that passed coverage and depth and were written */
	uint64_t kernel_launches;       /* kernels launched for the frame */
	uint64_t h2d_bytes;             /* bytes copied host->device for the frame (upload arena) */
	uint64_t d2h_bytes;             /* bytes copied device->host for the frame (store destinations) */
	uint64_t list_chunks_run_merge; /* tile-list chunks ordered by run merge ... */
	uint64_t list_chunks_key_range; /* ... and by key ranges + bitonic sort (long lists) */
	uint64_t host_record_ns;        /* host time spent recording the frame (begin_frame .. end_frame, all calls) */
	uint64_t host_submit_ns;        /* host time spent in rsrcu_end_frame (tables, upload, launches) */
	uint64_t frames_retried;        /* cumulative: frames launched again because a device-side buffer overflowed */
	uint64_t input_bytes;           /* the frame's draw inputs: bound vertex SoA floats x vertices + indices + instance matrices */
	uint64_t draws_culled;          /* cumulative: draws skipped because their bounding box lies outside one guard-band plane */
} RsrStats;
int rsrcu_get_stats(rsrcu_ctx* ctx, RsrStats* out);

/* last frame's device time per stage in milliseconds (CUDA events on the context stream):
 * [0] vertex [1] setup+clip+tile counts [2] unused (0) [3] cell scan (frames with several list cells
 * per tile) [4] list fill [5] tile sort+raster+resolve [6] whole frame incl. upload.
 * level 0: off; 1: events around the tile kernel and the frame only ([5], [6]); 2: every stage (the
 * extra events cost a few microseconds of stream time per frame). */
int rsrcu_set_profiling(rsrcu_ctx* ctx, int level);
int rsrcu_get_stage_ms(rsrcu_ctx* ctx, float* ms7);

#ifdef __cplusplus
}
#endif
#endif /* RSRCU_H */
