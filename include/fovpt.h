/* fovpt.h -- C ABI of libfovpt, the B200-native (sm_100a) replacement for the OptiX-backed
 * hot path of bipul-mohanto/fovPathTracing_optixCodeLatest.
 *
 * Everything the reference delegated to OptiX + its own device programs sits behind these
 * entry points; the C++ `SampleRenderer` shim (host/SimplePathtracer.cpp) and the python
 * binding call nothing else.  Plain pointers and sizes only.  Every function returns 0 on
 * success and non-zero on failure; fovpt_last_error() then holds the message (the shim rethrows
 * it as std::runtime_error, the reference's error convention, sutil/Exception.h:93-194).
 *
 * LaunchParams (248 B), Material (104 B) and Probe (64 B) cross this ABI as byte blobs with the
 * reference's layouts (PT_sv5_/LaunchParams.h:49-91, Material.h:11-70, Probe.cuh:6-21),
 * static_asserted in csrc/fovpt_layout.h.  Pointers inside LaunchParams are DEVICE pointers.
 *
 * reference interface each entry point replaces (file:line under /root/reference/PT_sv5_):
 */
#ifndef FOVPT_H
#define FOVPT_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fovpt_ctx fovpt_ctx;

/* Runtime form of the reference's compile-time literals; fovpt_default_config() yields sv5's. */
typedef struct fovpt_config {
    int32_t max_depth;        /* deviceProgram.cu:515 literal 4 (orig: 8, PT_original/deviceProgram.cu:429) */
    float   tmin, tmax;       /* deviceProgram.cu:41-42 : 0.01f, 1e16f; need 0 <= tmin < tmax          */
    float   exposure;         /* deviceProgram.cu:586 : 2^4                                            */
    int32_t accumulate;       /* 0 = sv5 (lerp commented out :565-581); 1 = vmv2 lerp                   *
                               * (PT_sv4_vmv2/deviceProgram.cu:545-553)                                 */
    int32_t foveated;         /* SimplePathtracer.cpp:22-23 FOV_ON / FOV_OFF                           */
    int32_t inner_radius, outer_radius;           /* SimplePathtracer.cpp:20-21 : 74, 241               */
    int32_t spp_periphery, spp_mid, spp_fovea;    /* SimplePathtracer.cpp:142,170,193 : 8, 16, 32       */
    int32_t spp_uniform;      /* SimplePathtracer.cpp:96 : 4 (FOV_OFF)                                  */
    int32_t seed_all_passes_with_subframe;  /* 0 = sv5 (mid/fovea reseed with subframe 0 and never     *
                               * accumulate, SimplePathtracer.cpp:162,174,196); 1 = every ring uses the *
                               * running subframe_index and accumulates (needed for C2/C3/C5 spp)       */
    int32_t skip_wasted_trace;/* 1 = do not trace the (max_depth+1)-th segment whose radiance the        *
                               * reference computes and discards (deviceProgram.cu:515);     *
                               * images are identical either way                                         */
    int32_t fast_math;        /* 0 = IEEE sin/cos/pow/log in shading; 1 = SFU approximations, the       *
                               * reference's --use_fast_math build (CMakeLists.txt:181)                  */
    int32_t filter;           /* reconstruction of the sub-sampled rings.  0 = the reference's box splat  *
                               * (each launch index fills fillSize^2 pixels, deviceProgram.cu:546-556);   *
                               * 1 = bilinear interpolation between neighbouring launch indices of the    *
                               * same ring -- fills the reference's empty Filter.h (Filter.h:3-7); no      *
                               * reference semantics exist for it (SURVEY 8a R-FILTER, 8f-3)               */
    int32_t features;         /* denoiser feature buffers.  0 = sv5 as shipped (the three stores are commented     *
                               * out, deviceProgram.cu:612-614); 1 = execute them: frame.normal_buffer /            *
                               * albedo_buffer receive the mean first-hit normal / albedo of the pixel's launch      *
                               * index (:509-512, :541-542, :653-665), frame.color_buffer the accumulated colour     *
                               * (LaunchParams.h:55-57); null pointers are skipped                                   */
} fovpt_config;

typedef struct fovpt_stats {
    uint64_t samples;          /* camera paths started (one iteration of deviceProgram.cu:448-539)     */
    uint64_t radiance_rays;    /* closest-hit traces actually issued                                    */
    uint64_t shadow_rays;      /* any-hit traces actually issued                                        */
    uint64_t kernel_launches;  /* libfovpt kernels launched since the last reset                        */
    uint64_t frames;
    uint64_t bvh_nodes, bvh_tris;
    float    bvh_build_ms;
    float    last_frame_ms;    /* CUDA-event time of the last fovpt_render_frame / fovpt_launch         */
    float    last_extend_ms, last_shade_ms, last_shadow_ms, last_raygen_ms, last_resolve_ms;
    uint64_t trace_launches;   /* launches of the traversal kernel since the last reset (roofline: bytes per launch) */
} fovpt_stats;

/* -- context: SampleRenderer::initOptix/createContext, SimplePathtracer.cpp:310-340.  `stream` is a
 *    cudaStream_t to run on, or NULL to create one (the reference creates its own, :339). */
int         fovpt_create(fovpt_ctx** out, int device, void* stream);
int         fovpt_destroy(fovpt_ctx*);
const char* fovpt_last_error(fovpt_ctx*);
void*       fovpt_stream(fovpt_ctx*);
void        fovpt_default_config(fovpt_config*);
int         fovpt_set_config(fovpt_ctx*, const fovpt_config*);
int         fovpt_get_config(fovpt_ctx*, fovpt_config*);

/* -- scene: SampleRenderer::buildAccel + createTextures + buildSBT, SimplePathtracer.cpp:602-746,
 *    748-799, 534-599.  Host pointers.  mesh = one (shape, material) group of loadOBJ; returns the
 *    mesh / texture id (>= 0) or a negative error.  Indices are bounds-checked, vertices must be finite.
 *    fovpt_scene_add_mesh copies the arrays straight into device memory before it returns (no host-side copy is kept; the
 *    caller may free them at once); vtx / idx may be NULL for an empty mesh.  Adding a mesh after a commit invalidates the
 *    committed scene until fovpt_scene_commit is called again; a second commit without changes rebuilds the same scene. */
int fovpt_scene_begin(fovpt_ctx*);
int fovpt_scene_add_texture(fovpt_ctx*, const uint32_t* rgba8, int w, int h);
int fovpt_scene_add_mesh(fovpt_ctx*, const float* vtx, size_t nvtx, const uint32_t* idx, size_t ntri,
                         const float* texcoord_or_null, const void* material104, int texture_id);
/* builds the BVH (on the device: Morton codes, radix sort, PLOC clustering, collapse into compressed wide nodes; on the host:
 * the top of the tree by binned SAH over the last <= 1 M clusters, a few tens of ms on the machine's cores);
 * *traversable_out is what LaunchParams.traversable (offset 160) must carry. */
int fovpt_scene_commit(fovpt_ctx*, uint64_t* traversable_out);

/* -- probe: SampleRenderer::setProbe + CUDAProbeData::createBuffer, SimplePathtracer.cpp:292-308,
 *    Probe.h:102-124.  Host arrays in; probe_out64 receives a Probe blob holding device pointers. */
int fovpt_probe_upload(fovpt_ctx*, int w, int h, const float* rgba, const float* pdfX, const float* cdfX,
                       const float* pdfY, const float* cdfY, const float offset[3], void* probe_out64);

/* -- frame buffers: SampleRenderer::resize, SimplePathtracer.cpp:228-274.  Frees and reallocates
 *    (accumulation is dropped, as in the reference); returns the device pointers to put in
 *    LaunchParams.frame.{accum_buffer,frame_buffer}. */
int fovpt_resize(fovpt_ctx*, int w, int h, void** accum_out, void** frame_out);

/* -- the denoiser buffers SampleRenderer::resize also allocates (SimplePathtracer.cpp:245-260): color / normal / albedo,
 *    float4 per pixel at the size of the last fovpt_resize (which frees them); returns the device pointers to put in
 *    LaunchParams.frame.{color_buffer,normal_buffer,albedo_buffer}.  Written only with fovpt_config.features = 1. */
int fovpt_feature_buffers(fovpt_ctx*, void** color_out, void** normal_out, void** albedo_out);
/* blocking D2H of the three feature buffers (any destination may be NULL) */
int fovpt_download_features(fovpt_ctx*, float* color_rgba, float* normal_rgba, float* albedo_rgba, size_t npix);

/* -- one optixLaunch (SimplePathtracer.cpp:148-157,178-187,200-209): uploads the 248-byte params and
 *    runs raygen for launch indices [0,lw)x[0,lh).  Asynchronous on the context's stream. */
int fovpt_launch(fovpt_ctx*, const void* launch_params248, size_t size, unsigned lw, unsigned lh);

/* -- SampleRenderer::render() (SimplePathtracer.cpp:77-214) as ONE submission: all passes of the
 *    configured schedule share one wavefront; advances frame.subframe_index in the caller's blob
 *    exactly as render() does.  Asynchronous; pair with fovpt_sync (CUDA_SYNC_CHECK, :212). */
int fovpt_render_frame(fovpt_ctx*, void* launch_params248_inout);
/* -- n_subframes successive render() calls of a progressive host loop (main.cpp:402-423 with frame.subframe_index
 *    advancing, PT_original/main.cpp:259,286) as ONE submission: the subframes share one wavefront -- path = (subframe,
 *    launch index, sample), seeds depend only on those (deviceProgram.cu:411) -- so the late, thin bounces of all
 *    subframes fill the machine together, and the per-frame accumulation (fovpt_config.accumulate) is replayed per pixel
 *    in subframe order.  Accumulation and frame buffer are bit-identical to n_subframes calls of fovpt_render_frame;
 *    only the intermediate RGBA8 frames are never materialised.  Batches larger than the path budget are split. */
int fovpt_render_frames(fovpt_ctx*, void* launch_params248_inout, int n_subframes);
int fovpt_sync(fovpt_ctx*);

/* -- SampleRenderer::downloadPixels (SimplePathtracer.cpp:276-280): blocking D2H of npix RGBA8
 *    pixels from the frame buffer most recently rendered into. */
int fovpt_download(fovpt_ctx*, uint32_t* dst, size_t npix);
/* the frame buffer of the last render was borrowed (render(CUDAOutputBuffer&), SimplePathtracer.cpp:216-226) and is about
 * to go away: fovpt_download falls back to the buffers of fovpt_resize instead of reading freed memory. */
int fovpt_forget_frame(fovpt_ctx*);
int fovpt_download_accum(fovpt_ctx*, float* dst_rgba, size_t npix);
/* checkpoint / resume of a progressive render (the accumulation buffer and frame.subframe_index are the only
 * persistent state of the path; the reference never saves them): blocking H2D of npix float4 values into the accumulation
 * buffer of fovpt_resize.  Together with fovpt_download_accum and the caller's own frame.subframe_index a run continues
 * bit-identically where it stopped (SampleRenderer::saveCheckpoint / loadCheckpoint). */
int fovpt_upload_accum(fovpt_ctx*, const float* src_rgba, size_t npix);

/* -- ray queries (optixTrace equivalents, deviceProgram.cu:196-248), for parity tests and tools.
 *    rays: n x {ox,oy,oz,tmin, dx,dy,dz,tmax} host floats (a negative tmin counts as 0).  mode 0: closest hit -> tri = global
 *    triangle id or -1, tuv = t,u,v.  mode 1: any front-facing hit -> tri = 1/0.  modes 2 / 3: as 0 / 1
 *    but tuv = (inner nodes visited, leaves visited, triangles tested) -- BVH-quality introspection. */
int fovpt_trace(fovpt_ctx*, const float* rays, size_t n, int mode, int32_t* tri, float* tuv);
/* texture-filter cross-check (test hook): fetches texture `texture_id` of the committed scene at n (u,v) pairs twice -- hw:
 * through a cudaTextureObject_t built exactly like the reference's (createTextures, SimplePathtracer.cpp:748-799; fetch
 * deviceProgram.cu:655-665), sw: through the software fetch (mode < 0: the product's own; modes 0..4: candidate weight
 * quantisations used to pin the definition).  Both outputs are n x RGBA floats. */
int fovpt_texture_check(fovpt_ctx*, int texture_id, const float* uv, size_t n, int mode, float* sw_rgba, float* hw_rgba);
/* sample-0 primary hit of every launch index of one launch (host outputs, lw*lh entries):
 * tri = global triangle id, -1 miss, -2 outside the ring; tuv = t,u,v; dir = camera direction. */
int fovpt_primary_hits(fovpt_ctx*, const void* launch_params248, unsigned lw, unsigned lh,
                       int32_t* tri, float* tuv, float* dir);

/* -- multi-GPU partition: this context renders only pixels whose tile is owned by `rank`.
 *    owner: tiles_x*tiles_y bytes (row-major, tile (0,0) at pixel (0,0)); NULL clears. */
int fovpt_set_tile_owner(fovpt_ctx*, int tile_size, int tiles_x, int tiles_y, const uint8_t* owner, int rank);

/* -- multi-GPU frame gather over peer memory (SURVEY 8e "one collective per frame"; no reference counterpart).
 *    One process per GPU.  The rank that displays the frame (the owner) allocates the shared RGBA8 frame with
 *    fovpt_peer_frame_create and hands the 128-byte handle to the other ranks (any transport: a broadcast, a pipe);
 *    they map it with fovpt_peer_frame_open (CUDA IPC, NVLink peer access).  Every rank passes the returned pointer
 *    as LaunchParams.frame.frame_buffer: the resolve kernel of fovpt_render_frame then stores the pixels of the
 *    rank's own tiles (fovpt_set_tile_owner) straight into the owner's memory -- the gather is fused into the
 *    kernel that produces the pixels.  Ordering is device-side, no host synchronisation and no collective call:
 *      contributor, per frame:  fovpt_render_frame; fovpt_peer_frame_signal      (arrival flag in the owner's memory)
 *      owner, per frame:        fovpt_render_frame; fovpt_peer_frame_wait;       (its stream waits for every arrival)
 *                               ...read / download the frame...; fovpt_peer_frame_release
 *    and a contributor's next resolve kernel waits for that release before it overwrites the frame.  A rank that
 *    does not show up within FOVPT_PEER_TIMEOUT_MS (environment, default 5000) makes the next fovpt_sync fail
 *    instead of hanging the stream.  All calls are asynchronous on the context stream except create/open/close. */
typedef struct fovpt_peer_handle { unsigned char opaque[128]; } fovpt_peer_handle;
int fovpt_peer_frame_create(fovpt_ctx*, size_t npix, int world, int owner_rank, fovpt_peer_handle* handle_out, uint32_t** frame_out);
int fovpt_peer_frame_open(fovpt_ctx*, const fovpt_peer_handle*, int rank, uint32_t** frame_out);
/* the shared allocation also holds a float4 accumulation buffer (north_star: "gather of the accumulation buffer"): with
 * mirror != 0 (the default while a peer frame is open) every rank's resolve kernel stores the accumulation value of its own
 * pixels into it next to the RGBA8 pixel, so after fovpt_peer_frame_wait the owner holds the complete float image of the
 * N-GPU job (fovpt_download_accum reads it).  *accum_out (may be NULL) receives the device pointer. */
int fovpt_peer_frame_accum(fovpt_ctx*, int mirror, float** accum_out);
int fovpt_peer_frame_signal(fovpt_ctx*);
int fovpt_peer_frame_wait(fovpt_ctx*);
int fovpt_peer_frame_release(fovpt_ctx*);
int fovpt_peer_frame_close(fovpt_ctx*);

/* per-kernel statistics (fovpt_stats.last_extend_ms ... last_resolve_ms) cost one CUDA event record between consecutive
 * kernels (~1 % of a C3 frame); on by default, a host that does not read them switches them off. */
int fovpt_set_kernel_timing(fovpt_ctx*, int on);
int fovpt_get_stats(fovpt_ctx*, fovpt_stats*);
int fovpt_reset_stats(fovpt_ctx*);
/* BVH introspection for validity tests: branching factor (8 or 4) and node size in bytes (96 / 64, layout in
 * csrc/bvh.cuh); copies the quantised nodes / sorted triangle ids to host. */
int fovpt_bvh_layout(fovpt_ctx*, int* wide, int* node_bytes);
int fovpt_bvh_info(fovpt_ctx*, uint64_t* n_nodes, uint64_t* n_tris);
int fovpt_bvh_read(fovpt_ctx*, void* nodes, uint32_t* sorted_tri_ids);

#ifdef __cplusplus
}
#endif
#endif
