// api.cu -- the C ABI of libfovpt (include/fovpt.h): context, scene ingest, probe, frame buffers and
// the frame scheduler that strings the wavefront kernels together.  Each entry point names, in
// fovpt.h, the reference host function it stands in for (PT_sv5_/SimplePathtracer.cpp).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>      // header-only; ranges are no-ops unless a profiler injects the NVTX library

#include "../../include/fovpt.h"
#include "bvh_build.h"
#include "wavefront_launch.h"

using namespace fovpt;

namespace {

// One mesh of the scene being assembled.  Geometry does not stay on the host: fovpt_scene_add_mesh copies it straight from the
// caller's arrays into device arenas (vertices, mesh-local indices, texcoords), the record keeps where it went.
struct MeshRec {
    size_t vbase, tbase, nv, nt;       // first vertex / triangle in the arenas, counts
    MaterialBlob mat; int texture_id; bool has_tc;
};
// Device array that grows by doubling (old contents moved device-to-device); `cap` / `used` in elements.
template <typename T> struct Arena { T* p = nullptr; size_t cap = 0, used = 0; };
struct HostTex { int w, h; std::vector<uint32_t> px; };

struct Mark { cudaEvent_t ev; int kind; };      // start of a timed interval on the context stream; it ends at the next mark

}  // namespace

// flag words behind the shared frame of fovpt_peer_frame_*: arrival[0..63] by rank, then release, then the owner's error word
constexpr int kPeerMaxWorld = 64, kPeerRelease = 64, kPeerError = 65, kPeerFlagWords = 128;

struct fovpt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    fovpt_config cfg;
    int n_sm = 148;
    LaunchCfg L;

    // scene being assembled: mesh records + device arenas (filled by fovpt_scene_add_mesh), textures staged on the host
    std::vector<MeshRec> meshes;
    Arena<float> a_vtx; Arena<uint32_t> a_idx; Arena<float2> a_tc;     // 3 floats / 3 mesh-local indices / 1 texcoord per element group
    std::vector<HostTex> textures;
    // scene (device)
    BuiltBvh bvh;
    SceneDev scene{};
    uint32_t* d_idx = nullptr; uint32_t* d_tri_mesh = nullptr; uint2* d_mesh_table = nullptr;     // derived at commit: global indices, mesh per triangle
    MaterialBlob* d_mat = nullptr; int2* d_mesh_tex = nullptr; int4* d_tex_info = nullptr; uint32_t* d_texels = nullptr;
    bool committed = false;

    // probe
    float4* p_data = nullptr; float *p_pdfX = nullptr, *p_cdfX = nullptr, *p_pdfY = nullptr, *p_cdfY = nullptr;
    uint16_t *p_gy = nullptr, *p_gx = nullptr;       // CDF guide tables (shade.cuh ProbeGuide), valid for p_cdfX / p_cdfY

    // frame buffers owned by the library (fovpt_resize)
    int fb_w = 0, fb_h = 0;
    float4* fb_accum = nullptr; uint32_t* fb_frame = nullptr;
    uint32_t* last_frame_ptr = nullptr; float4* last_accum_ptr = nullptr; size_t last_npix = 0;

    // wavefront storage
    size_t path_cap = 0;
    PathState ps{};
    Queues q{};
    size_t li_cap = 0, needed_cap = 0; float4* li_backplate = nullptr; float4* li_color = nullptr; uint8_t* needed = nullptr;
    size_t li_feat_cap = 0; float4 *li_normal = nullptr, *li_albedo = nullptr;      // fovpt_config.features only
    float4 *fb_color = nullptr, *fb_normal = nullptr, *fb_albedo = nullptr;         // fovpt_feature_buffers
    size_t px_cap = 0; uint32_t* owner_map = nullptr;
    std::vector<unsigned char> owner_key;      // what the cached owner map was computed for
    DevStats* d_stats = nullptr;
    unsigned char* h_pinned = nullptr; size_t h_pinned_sz = 0;

    // partition
    uint8_t* d_tile_owner = nullptr; int tile_size = 0, tiles_x = 0, tiles_y = 0, rank = 0;

    // shared frame of a multi-GPU job (fovpt_peer_frame_*): `base` is the owner's allocation (mapped through CUDA IPC on
    // the contributors) = RGBA8 frame, then kPeerFlagWords u32: arrival[rank] | release | owner-side error, then the float4 accumulation buffer
    struct PeerFrame {
        unsigned char* base = nullptr; uint32_t* flags = nullptr; uint32_t* d_err = nullptr;
        float4* accum = nullptr;            // shared accumulation buffer behind the flags (every rank mirrors its pixels into it)
        bool mirror_accum = true;
        bool open = false, owner = false; int world = 0, owner_rank = 0, rank = 0; uint32_t seq = 0; size_t npix = 0;
        long long timeout_clocks = 0;
    } peer;

    // stats
    uint64_t launches = 0, frames = 0, trace_launches = 0;
    float build_ms = 0.f;
    cudaEvent_t ev_frame0 = nullptr, ev_frame1 = nullptr;
    std::vector<Mark> marks; size_t marks_used = 0;
    float last_ms[6] = {0, 0, 0, 0, 0, 0};     // frame, raygen, extend, shade, shadow, resolve
    bool kernel_timing = true;                  // per-kernel statistics (last_*_ms) through chained CUDA events: fovpt_set_kernel_timing
    bool frame_timed = false;
};

namespace {

enum { K_RAYGEN = 1, K_EXTEND = 2, K_SHADE = 3, K_SHADOW = 4, K_RESOLVE = 5 };

int fail(fovpt_ctx* c, const std::string& m) { if (c) c->err = m; return 1; }
int cuda_fail(fovpt_ctx* c, cudaError_t e, const char* what)
{
    return fail(c, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(c, e__, #call); } while (0)

template <typename T> cudaError_t dmalloc(T** p, size_t n) { return cudaMalloc((void**)p, (n ? n : 1) * sizeof(T)); }

// Per-kernel times come from ONE event between consecutive kernels (an interval runs from its mark to the next mark or to the
// end of the frame, so it includes the launch gap behind the kernel): 22 event records per depth-8 frame.  A pair around every
// launch (54 records) cost 127 us of a 4.7 ms C3 frame and 140 us of a 1.5 ms 1/8-frame share.
void mark(fovpt_ctx* c, int kind)
{
    if (!c->kernel_timing) return;
    if (c->marks_used == c->marks.size()) { Mark m; cudaEventCreate(&m.ev); m.kind = 0; c->marks.push_back(m); }
    Mark& m = c->marks[c->marks_used++];
    m.kind = kind;
    cudaEventRecord(m.ev, c->stream);
}

void free_scene(fovpt_ctx* c)
{
    free_bvh(&c->bvh);
    cudaFree(c->d_idx); cudaFree(c->d_tri_mesh); cudaFree(c->d_mesh_table); cudaFree(c->d_mat);
    cudaFree(c->d_mesh_tex); cudaFree(c->d_tex_info); cudaFree(c->d_texels);
    c->d_idx = nullptr; c->d_tri_mesh = nullptr; c->d_mesh_table = nullptr; c->d_mat = nullptr;
    c->d_mesh_tex = nullptr; c->d_tex_info = nullptr; c->d_texels = nullptr;
    c->committed = false;
}

// what fovpt_scene_add_mesh uploaded (the committed scene keeps using the arenas; only scene_begin / destroy drop them)
void free_arenas(fovpt_ctx* c)
{
    cudaFree(c->a_vtx.p); cudaFree(c->a_idx.p); cudaFree(c->a_tc.p);
    c->a_vtx = Arena<float>{}; c->a_idx = Arena<uint32_t>{}; c->a_tc = Arena<float2>{};
    c->meshes.clear();
}

template <typename T> cudaError_t arena_reserve(Arena<T>& a, size_t need, cudaStream_t st)
{
    if (need <= a.cap) return cudaSuccess;
    const size_t cap = std::max(need, a.cap * 2);
    T* p = nullptr;
    cudaError_t e = cudaMalloc((void**)&p, cap * sizeof(T) + 32);
    if (e != cudaSuccess) return e;
    if (a.used) e = cudaMemcpyAsync(p, a.p, a.used * sizeof(T), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cudaFree(p); return e; }
    cudaFree(a.p);
    a.p = p; a.cap = cap;
    return cudaSuccess;
}

void free_paths(fovpt_ctx* c)
{
    cudaFree(c->ps.direct); cudaFree(c->ps.indirect); cudaFree(c->ps.alpha); cudaFree(c->ps.feat_n); cudaFree(c->ps.feat_a);
    for (RayBuf* b : {&c->q.cur, &c->q.next}) { cudaFree(b->o); cudaFree(b->d); cudaFree(b->thr); cudaFree(b->rng); cudaFree(b->meta); *b = RayBuf{}; }
    cudaFree(c->q.shadow.o); cudaFree(c->q.shadow.d); cudaFree(c->q.shadow.pend); cudaFree(c->q.shadow.pid); c->q.shadow = ShadowBuf{};
    cudaFree(c->q.hit); c->q.hit = nullptr;
    c->ps = PathState{}; c->path_cap = 0;
}

int ensure_paths(fovpt_ctx* c, size_t n)
{
    const bool feat = c->cfg.features != 0;
    if (n <= c->path_cap && (!feat || c->ps.feat_n)) return 0;
    n = std::max(n, c->path_cap);
    free_paths(c);
    CU(dmalloc(&c->ps.direct, n)); CU(dmalloc(&c->ps.indirect, n)); CU(dmalloc(&c->ps.alpha, n));
    if (feat) { CU(dmalloc(&c->ps.feat_n, n)); CU(dmalloc(&c->ps.feat_a, n)); }
    for (RayBuf* b : {&c->q.cur, &c->q.next}) {
        CU(dmalloc(&b->o, n)); CU(dmalloc(&b->d, n)); CU(dmalloc(&b->thr, n)); CU(dmalloc(&b->rng, n)); CU(dmalloc(&b->meta, n));
    }
    CU(dmalloc(&c->q.shadow.o, n)); CU(dmalloc(&c->q.shadow.d, n)); CU(dmalloc(&c->q.shadow.pend, n)); CU(dmalloc(&c->q.shadow.pid, n));
    CU(dmalloc(&c->q.hit, n));
    c->path_cap = n;
    return 0;
}

// n: entries of the per-(layer, launch index) buffers; n_needed: launch indices of one layer (the `needed` bytes)
int ensure_li(fovpt_ctx* c, size_t n, size_t n_needed)
{
    if (n > c->li_cap) {
        cudaFree(c->li_backplate); cudaFree(c->li_color);
        c->li_backplate = nullptr; c->li_color = nullptr; c->li_cap = 0;
        CU(dmalloc(&c->li_backplate, n)); CU(dmalloc(&c->li_color, n));
        c->li_cap = n;
    }
    if (c->cfg.features && n > c->li_feat_cap) {
        cudaFree(c->li_normal); cudaFree(c->li_albedo);
        c->li_normal = nullptr; c->li_albedo = nullptr; c->li_feat_cap = 0;
        CU(dmalloc(&c->li_normal, n)); CU(dmalloc(&c->li_albedo, n));
        c->li_feat_cap = n;
    }
    if (n_needed > c->needed_cap) {
        cudaFree(c->needed); c->needed = nullptr; c->needed_cap = 0;
        CU(dmalloc(&c->needed, n_needed));
        c->needed_cap = n_needed; c->owner_key.clear();
    }
    return 0;
}

void nvtx_push(const char* name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }

int ensure_px(fovpt_ctx* c, size_t n)
{
    if (n <= c->px_cap) return 0;
    cudaFree(c->owner_map);
    CU(dmalloc(&c->owner_map, n));
    c->px_cap = n; c->owner_key.clear();
    return 0;
}

int ensure_pinned(fovpt_ctx* c, size_t n)
{
    if (n <= c->h_pinned_sz) return 0;
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    c->h_pinned = nullptr; c->h_pinned_sz = 0;
    CU(cudaMallocHost((void**)&c->h_pinned, n));
    c->h_pinned_sz = n;
    return 0;
}

// One pass of the frame schedule, host side (SimplePathtracer.cpp:137-209)
struct PassHost { unsigned lw, lh, factor; int fill; float r_inner, r_outer; unsigned spp, off_x, off_y, redraw, subframe, sub_step; };

void fill_frame(fovpt_ctx* c, const LaunchParamsBlob& lp, const std::vector<PassHost>& passes, FrameDev* F, size_t* total_li)
{
    *F = FrameDev{};
    F->w = lp.frame.size_x; F->h = lp.frame.size_y;
    F->cx = lp.frame.c_x; F->cy = lp.frame.c_y;
    memcpy(F->eye, lp.camera.eye, 12); memcpy(F->U, lp.camera.U, 12); memcpy(F->V, lp.camera.V, 12); memcpy(F->W, lp.camera.W, 12);
    F->probe = lp.probe;
    if (c->p_gy && lp.probe.cdfValuesX == c->p_cdfX && lp.probe.cdfValuesY == c->p_cdfY && lp.probe.width <= 65535 && lp.probe.height <= 65535) {
        F->guide.gy = c->p_gy; F->guide.gx = c->p_gx;
    }
    F->accum = lp.frame.accum_buffer; F->frame = lp.frame.frame_buffer;
    if (c->cfg.features) { F->feat_color = lp.frame.color_buffer; F->feat_normal = lp.frame.normal_buffer; F->feat_albedo = lp.frame.albedo_buffer; }
    F->npass = (int)passes.size();
    size_t li = 0;
    for (size_t p = 0; p < passes.size(); ++p) {
        const PassHost& h = passes[p];
        PassDev& d = F->pass[p];
        d.lw = h.lw; d.lh = h.lh; d.factor = h.factor; d.fill = (unsigned)h.fill; d.spp = h.spp; d.redraw = h.redraw; d.subframe = h.subframe; d.sub_step = h.sub_step;
        d.uoff_x = h.off_x; d.uoff_y = h.off_y; d.off_x = (int)h.off_x; d.off_y = (int)h.off_y;
        d.r_inner = h.r_inner; d.r_outer = h.r_outer;
        d.pid_base = 0; d.li_base = (unsigned)li;
        li += (size_t)h.lw * h.lh;
    }
    *total_li = li;
    F->nsub = 1; F->li_total = (unsigned)li;
    F->max_depth = c->cfg.max_depth; F->tmin = c->cfg.tmin; F->tmax = c->cfg.tmax; F->exposure = c->cfg.exposure;
    F->accumulate = c->cfg.accumulate; F->skip_wasted = c->cfg.skip_wasted_trace; F->filter = c->cfg.filter;
    F->tile_owner = c->d_tile_owner; F->tile_size = c->tile_size; F->tiles_x = c->tiles_x; F->tiles_y = c->tiles_y; F->rank = c->rank;
}

constexpr size_t kPathBudget = (size_t)48 << 20;      // at most this many paths in flight

size_t paths_per_layer(const std::vector<PassHost>& passes)
{
    size_t n = 0;
    for (auto& p : passes) n += (size_t)p.lw * p.lh * p.spp;
    return n;
}

// Runs `nsub` accumulated subframes of the passes through ONE wavefront (layer s of pass p is seeded with
// subframe + s * sub_step); asynchronous.  nsub > 1 requires nsub * paths_per_layer <= kPathBudget (fovpt_render_frames
// splits longer batches); a single layer larger than the budget is cut into row chunks as before.
int run_frame(fovpt_ctx* c, const LaunchParamsBlob& lp, const std::vector<PassHost>& passes, int nsub = 1)
{
    if (!c->committed) return fail(c, "scene not committed");
    if (!lp.probe.data) return fail(c, "LaunchParams.probe is empty (call fovpt_probe_upload / setProbe first)");
    if (!lp.frame.accum_buffer || !lp.frame.frame_buffer) return fail(c, "LaunchParams.frame buffers are null (call fovpt_resize first)");
    if (passes.empty() || passes.size() > (size_t)kMaxPasses) return fail(c, "bad pass count");
    for (auto& p : passes) {
        if (p.spp == 0 || p.spp > 4096) return fail(c, "samples_per_launch out of range");
        if ((size_t)p.lw * p.lh >= (1u << 28)) return fail(c, "launch too large");
        if (p.factor == 0 || p.fill <= 0) return fail(c, "factor / fillSize must be positive");
    }
    if (nsub < 1 || (nsub > 1 && (size_t)nsub * paths_per_layer(passes) > kPathBudget)) return fail(c, "subframe batch exceeds the path budget");
    cudaSetDevice(c->device);
    FrameDev F; size_t total_li = 0;
    fill_frame(c, lp, passes, &F, &total_li);
    F.nsub = nsub;
    const size_t npix = (size_t)F.w * F.h;
    // a tile table / shared frame that does not cover the frame would be read / written out of bounds by the kernels
    if (F.tile_owner && ((size_t)F.tiles_x * F.tile_size < (size_t)F.w || (size_t)F.tiles_y * F.tile_size < (size_t)F.h))
        return fail(c, "tile owner table does not cover the frame (fovpt_set_tile_owner after a resize?)");
    if (c->peer.open && (unsigned char*)F.frame == c->peer.base && c->peer.npix < npix)
        return fail(c, "the shared peer frame is smaller than the frame being rendered");
    if (ensure_li(c, total_li * (size_t)nsub, total_li) || ensure_px(c, npix)) return 1;

    nvtx_push("fovpt frame");
    c->marks_used = 0;
    cudaEventRecord(c->ev_frame0, c->stream);

    // owner map: recomputed only when the frame geometry / gaze / partition changes
    std::vector<unsigned char> key(sizeof(int) * 4 + sizeof(PassDev) * kMaxPasses + sizeof(void*) + 5 * sizeof(int));
    {
        unsigned char* k = key.data();
        int head[4] = {F.w, F.h, (int)F.cx, (int)F.cy};
        memcpy(k, head, sizeof(head)); k += sizeof(head);
        PassDev geo[kMaxPasses]; memset(geo, 0, sizeof(geo));
        for (int p = 0; p < F.npass; ++p) { geo[p] = F.pass[p]; geo[p].spp = 0; geo[p].redraw = 0; geo[p].subframe = 0; geo[p].sub_step = 0; }
        memcpy(k, geo, sizeof(geo)); k += sizeof(geo);
        const void* to = F.tile_owner; memcpy(k, &to, sizeof(void*)); k += sizeof(void*);
        int tail[5] = {F.tile_size, F.tiles_x, F.tiles_y, F.rank, F.npass};
        memcpy(k, tail, sizeof(tail));
    }
    if (key != c->owner_key) {
        CU(cudaMemsetAsync(c->needed, 0, total_li, c->stream));
        mark(c, K_RESOLVE);
        launch_owner_map(c->L, F, c->owner_map, c->needed); c->launches++;
        c->owner_key = key;
    }

    // work groups: every group is one trip through the wavefront with at most kPathBudget paths in flight
    struct Group { ChunkTable gen, res; size_t paths; };
    std::vector<Group> groups;
    {
        Group cur{}; size_t used = 0, used_li = 0;
        auto flush = [&]() { if (cur.gen.n) { cur.gen.total = (unsigned)used; cur.res.total = (unsigned)used_li; cur.paths = used; groups.push_back(cur); } cur = Group{}; used = 0; used_li = 0; };
        auto push = [&](int p, unsigned row0, unsigned rows, unsigned ns) {
            const PassDev& P = F.pass[p];
            ChunkDev g{p, row0, rows, 0u, ns, (unsigned)used, (unsigned)used};
            ChunkDev r = g; r.first = (unsigned)used_li;
            cur.res.c[cur.res.n++] = r; cur.gen.c[cur.gen.n++] = g;
            used += (size_t)ns * rows * P.lw * P.spp; used_li += (size_t)ns * rows * P.lw;
        };
        if ((size_t)nsub * paths_per_layer(passes) <= kPathBudget) {
            for (int p = 0; p < F.npass; ++p) push(p, 0u, F.pass[p].lh, (unsigned)nsub);
            flush();
        } else {
            for (int p = 0; p < F.npass; ++p) {
                const PassDev& P = F.pass[p];
                const size_t per_row = (size_t)P.lw * P.spp;
                unsigned row = 0;
                while (row < P.lh) {
                    const size_t room = kPathBudget > used ? kPathBudget - used : 0;
                    unsigned rows = (unsigned)std::min<size_t>(P.lh - row, per_row ? room / per_row : P.lh);
                    if (rows == 0 || cur.gen.n == kMaxChunks) {
                        if (cur.gen.n == 0) rows = 1;      // a single row larger than the budget: take it alone
                        else { flush(); continue; }
                    }
                    push(p, row, rows, 1u);
                    row += rows;
                }
            }
            flush();
        }
    }
    size_t max_paths = 0;
    for (auto& g : groups) max_paths = std::max(max_paths, g.paths);
    if (max_paths >= 0xff000000ull) return fail(c, "path batch exceeds 32-bit ids");
    if (ensure_paths(c, max_paths)) return 1;
    F.path_cap = (unsigned)std::min<size_t>(c->path_cap, 0xffffffffu);

    const bool feat = c->cfg.features != 0;
    PathState ps = c->ps;
    if (!feat) { ps.feat_n = nullptr; ps.feat_a = nullptr; }
    const int iters = c->cfg.max_depth + 1 + (c->scene.any_catcher ? 4 : 0);
    for (auto& g : groups) {
        mark(c, K_RAYGEN);
        CU(cudaMemsetAsync(c->q.counts, 0, 8 * sizeof(uint32_t), c->stream));
        launch_raygen(c->L, F, g.gen, c->needed, ps, c->q, c->li_backplate, c->d_stats); c->launches++;
        // iteration i: k_trace extends the radiance rays of bounce i and resolves the shadow rays that
        // k_shade(i-1) emitted (counter slot alternates 2 / 5), k_shade(i) consumes the hits.
        for (int it = 0; it <= iters; ++it) {
            const int slot_prev = (it & 1) ? 2 : 5, slot_new = (it & 1) ? 5 : 2;
            char range[32]; snprintf(range, sizeof(range), "fovpt bounce %d", it);
            nvtx_push(range);
            mark(c, K_EXTEND);
            launch_trace(c->L, c->scene, ps, c->q, slot_prev, c->cfg.tmin, c->cfg.tmax, c->d_stats); c->launches++; c->trace_launches++;
            if (it == iters) { nvtx_pop(); break; }      // the last launch only drains the final shadow rays
            if (feat && it == 0) { launch_features(c->L, c->scene, ps, c->q, c->cfg.fast_math != 0); c->launches++; }
            mark(c, K_SHADE);
            launch_shade(c->L, F, c->scene, ps, c->q, slot_new, c->cfg.fast_math != 0); c->launches++;
            std::swap(c->q.cur, c->q.next);                // the device-side counters are swapped by k_shade's last CTA
            nvtx_pop();
        }
        mark(c, K_RESOLVE);
        launch_resolve_li(c->L, F, g.res, c->needed, c->ps, c->li_backplate, c->li_color, feat ? c->li_normal : nullptr, feat ? c->li_albedo : nullptr); c->launches++;
    }
    // a contributor to a shared frame stores its pixels into the owner's memory: not before the owner released the last frame
    float4* mirror = nullptr;
    if (c->peer.open && (unsigned char*)F.frame == c->peer.base) {
        if (!c->peer.owner)
            launch_peer_wait(c->L, c->peer.flags + kPeerRelease, 1, -1, c->peer.seq, c->peer.d_err, c->peer.timeout_clocks);
        if (c->peer.mirror_accum) mirror = c->peer.accum;
    }
    launch_resolve_px(c->L, F, c->owner_map, c->li_color, c->needed, mirror, feat ? c->li_normal : nullptr, feat ? c->li_albedo : nullptr, c->cfg.fast_math != 0); c->launches++;
    cudaEventRecord(c->ev_frame1, c->stream);
    nvtx_pop();
    c->frame_timed = true;
    c->frames += (uint64_t)nsub;
    c->last_frame_ptr = F.frame; c->last_accum_ptr = mirror ? mirror : F.accum; c->last_npix = npix;
    CU(cudaGetLastError());
    return 0;
}

void collect_times(fovpt_ctx* c)
{
    if (!c->frame_timed) return;
    cudaEventSynchronize(c->ev_frame1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev_frame0, c->ev_frame1);
    for (int i = 0; i < 6; ++i) c->last_ms[i] = 0.f;
    c->last_ms[0] = ms;
    for (size_t i = 0; i < c->marks_used; ++i) {
        const Mark& m = c->marks[i];
        if (m.kind <= 0 || m.kind > 5) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, m.ev, i + 1 < c->marks_used ? c->marks[i + 1].ev : c->ev_frame1) == cudaSuccess) c->last_ms[m.kind] += t;
    }
    c->frame_timed = false;
}

std::vector<PassHost> schedule(const fovpt_config& cfg, const LaunchParamsBlob& lp)
{
    std::vector<PassHost> v;
    const int w = lp.frame.size_x, h = lp.frame.size_y;
    const unsigned sub = lp.frame.subframe_index;
    const bool ext = cfg.seed_all_passes_with_subframe != 0;
    if (!cfg.foveated) {     // FOV_OFF block, SimplePathtracer.cpp:85-131
        v.push_back({(unsigned)w, (unsigned)h, 1u, 1, 0.f, 1000000000.f, (unsigned)cfg.spp_uniform, 0u, 0u, 0u, ext ? sub : 0u, ext ? 1u : 0u});
        return v;
    }
    const int ri = cfg.inner_radius, ro = cfg.outer_radius;
    const unsigned cx = lp.frame.c_x, cy = lp.frame.c_y;
    // periphery :137-157
    v.push_back({(unsigned)(w / 4), (unsigned)(h / 4), 4u, 4, (float)ro, 1000000000.f, (unsigned)cfg.spp_periphery, 0u, 0u, 0u, sub, 1u});
    // mid ring :160-187 (launch dims are the float r_outer truncated)
    const float r2 = (float)(ro + 2);
    v.push_back({(unsigned)r2, (unsigned)r2, 2u, 2, (float)ri, r2, (unsigned)cfg.spp_mid, cx - (unsigned)(ro + 2), cy - (unsigned)(ro + 2),
                 ext ? 0u : 1u, ext ? sub : 0u, ext ? 1u : 0u});
    // fovea :189-209
    const float r3 = (float)(ri + 1);
    v.push_back({(unsigned)(r3 * 2), (unsigned)(r3 * 2), 1u, 1, 0.f, r3, (unsigned)cfg.spp_fovea, cx - (unsigned)(ri + 1), cy - (unsigned)(ri + 1),
                 ext ? 0u : 1u, ext ? sub : 0u, ext ? 1u : 0u});
    return v;
}

}  // namespace

// ====================================================================== C ABI
extern "C" {

void fovpt_default_config(fovpt_config* k)
{
    memset(k, 0, sizeof(*k));
    k->max_depth = 4; k->tmin = 0.01f; k->tmax = 1e16f; k->exposure = 16.0f; k->accumulate = 0; k->foveated = 1;
    k->inner_radius = 74; k->outer_radius = 241; k->spp_periphery = 8; k->spp_mid = 16; k->spp_fovea = 32; k->spp_uniform = 4;
    k->seed_all_passes_with_subframe = 0; k->skip_wasted_trace = 1; k->fast_math = 0; k->filter = 0; k->features = 0;
}

int fovpt_create(fovpt_ctx** out, int device, void* stream)
{
    if (!out) return 1;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return 2;   // "no CUDA capable devices", SimplePathtracer.cpp:318-321
    fovpt_ctx* c = new fovpt_ctx();
    c->device = device;
    fovpt_default_config(&c->cfg);
    if (cudaSetDevice(device) != cudaSuccess) { delete c; return 3; }
    cudaFree(0);
    if (stream) c->stream = (cudaStream_t)stream;
    else { if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return 4; } c->own_stream = true; }
    if (const char* e = getenv("FOVPT_KERNEL_TIMING")) c->kernel_timing = atoi(e) != 0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    c->n_sm = prop.multiProcessorCount;
    int be = 8, bs = 4;
    query_occupancy(&be, &bs);
    c->L.stream = c->stream;
    c->L.grid_trace = c->n_sm * (be > 0 ? be : 8);
    c->L.grid_shade = c->n_sm * (bs > 0 ? bs : 4);
    c->L.grid_light = c->n_sm * 8;
    if (dmalloc(&c->q.counts, 8) != cudaSuccess || dmalloc(&c->d_stats, 1) != cudaSuccess) { delete c; return 5; }
    cudaMemset(c->q.counts, 0, 32); cudaMemset(c->d_stats, 0, sizeof(DevStats));
    cudaEventCreate(&c->ev_frame0); cudaEventCreate(&c->ev_frame1);
    *out = c;
    return 0;
}

int fovpt_destroy(fovpt_ctx* c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    fovpt_peer_frame_close(c);
    free_scene(c); free_arenas(c); free_paths(c);
    cudaFree(c->p_data); cudaFree(c->p_pdfX); cudaFree(c->p_cdfX); cudaFree(c->p_pdfY); cudaFree(c->p_cdfY); cudaFree(c->p_gy); cudaFree(c->p_gx);
    cudaFree(c->fb_accum); cudaFree(c->fb_frame); cudaFree(c->li_backplate); cudaFree(c->li_color); cudaFree(c->needed);
    cudaFree(c->li_normal); cudaFree(c->li_albedo); cudaFree(c->fb_color); cudaFree(c->fb_normal); cudaFree(c->fb_albedo);
    cudaFree(c->owner_map); cudaFree(c->q.counts); cudaFree(c->d_stats); cudaFree(c->d_tile_owner);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    for (auto& m : c->marks) cudaEventDestroy(m.ev);
    cudaEventDestroy(c->ev_frame0); cudaEventDestroy(c->ev_frame1);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

const char* fovpt_last_error(fovpt_ctx* c) { return c ? c->err.c_str() : "null context"; }
void* fovpt_stream(fovpt_ctx* c) { return c ? (void*)c->stream : nullptr; }

int fovpt_set_config(fovpt_ctx* c, const fovpt_config* k)
{
    if (!c || !k) return 1;
    if (k->max_depth < 0 || k->max_depth > 200) return fail(c, "max_depth out of range");
    if (k->foveated && (k->inner_radius < 0 || k->outer_radius < k->inner_radius)) return fail(c, "bad radii");
    if (k->spp_periphery <= 0 || k->spp_mid <= 0 || k->spp_fovea <= 0 || k->spp_uniform <= 0) return fail(c, "spp must be positive");
    if (k->filter < 0 || k->filter > 1) return fail(c, "filter must be 0 (box) or 1 (bilinear)");
    if (k->features < 0 || k->features > 1) return fail(c, "features must be 0 or 1");
    if (!(k->tmin >= 0.f) || !(k->tmax > k->tmin)) return fail(c, "need 0 <= tmin < tmax");
    c->cfg = *k;
    return 0;
}
int fovpt_set_kernel_timing(fovpt_ctx* c, int on)
{
    if (!c) return 1;
    c->kernel_timing = on != 0;
    return 0;
}
int fovpt_get_config(fovpt_ctx* c, fovpt_config* k) { if (!c || !k) return 1; *k = c->cfg; return 0; }

int fovpt_scene_begin(fovpt_ctx* c)
{
    if (!c) return 1;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->textures.clear();
    free_scene(c); free_arenas(c);
    return 0;
}

int fovpt_scene_add_texture(fovpt_ctx* c, const uint32_t* rgba8, int w, int h)
{
    if (!c || !rgba8 || w <= 0 || h <= 0) { fail(c, "bad texture"); return -1; }
    try {
        HostTex t; t.w = w; t.h = h; t.px.assign(rgba8, rgba8 + (size_t)w * h);
        c->textures.push_back(std::move(t));
    } catch (const std::exception& e) { fail(c, std::string("fovpt_scene_add_texture: ") + e.what()); return -1; }
    return (int)c->textures.size() - 1;
}

int fovpt_scene_add_mesh(fovpt_ctx* c, const float* vtx, size_t nvtx, const uint32_t* idx, size_t ntri,
                         const float* tc, const void* material104, int texture_id)
{
    if (!c || (!vtx && nvtx) || (!idx && ntri) || !material104) { fail(c, "null mesh argument"); return -1; }
    // validation as two branch-free reductions (they vectorise; an early-exit loop over 30 M indices does not)
    uint32_t max_idx = 0;
    for (size_t i = 0; i < 3 * ntri; ++i) max_idx = std::max(max_idx, idx[i]);
    if (ntri && max_idx >= nvtx) { fail(c, "mesh index out of range (see SURVEY A.9: one usemtl per shape)"); return -1; }
    uint32_t not_finite = 0;
    for (size_t i = 0; i < 3 * nvtx; ++i) { uint32_t b; memcpy(&b, vtx + i, 4); not_finite |= (uint32_t)((b & 0x7f800000u) == 0x7f800000u); }
    if (not_finite) { fail(c, "non-finite vertex"); return -1; }
    if (c->a_vtx.used / 3 + nvtx >= (1ull << 32) || c->a_idx.used / 3 + ntri >= (1u << 29)) { fail(c, "too many vertices / triangles"); return -1; }
    // Straight from the caller's arrays into the device arenas: no host-side copy of the geometry exists.  (cudaMemcpy from
    // pageable memory returns when the source has been consumed, so the caller may free its arrays right away.)
    cudaSetDevice(c->device);
    if (c->committed) {                                   // the committed scene reads the arenas (texcoords): growing them invalidates it
        cudaStreamSynchronize(c->stream);
        c->committed = false;                             // frames fail with "scene not committed" until fovpt_scene_commit runs again
    }
    cudaError_t e = arena_reserve(c->a_vtx, c->a_vtx.used + 3 * nvtx, c->stream);
    if (e == cudaSuccess) e = arena_reserve(c->a_idx, c->a_idx.used + 3 * ntri, c->stream);
    if (e == cudaSuccess) e = arena_reserve(c->a_tc, c->a_tc.used + nvtx, c->stream);
    MeshRec m;
    m.vbase = c->a_vtx.used / 3; m.tbase = c->a_idx.used / 3; m.nv = nvtx; m.nt = ntri; m.texture_id = texture_id; m.has_tc = tc != nullptr;
    memcpy(&m.mat, material104, sizeof(MaterialBlob));
    if (e == cudaSuccess && nvtx) e = cudaMemcpy(c->a_vtx.p + c->a_vtx.used, vtx, 3 * nvtx * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && ntri) e = cudaMemcpy(c->a_idx.p + c->a_idx.used, idx, 3 * ntri * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && nvtx) e = tc ? cudaMemcpy(c->a_tc.p + c->a_tc.used, tc, nvtx * sizeof(float2), cudaMemcpyHostToDevice)
                                         : cudaMemset(c->a_tc.p + c->a_tc.used, 0, nvtx * sizeof(float2));
    if (e != cudaSuccess) { cuda_fail(c, e, "fovpt_scene_add_mesh upload"); return -1; }
    c->a_vtx.used += 3 * nvtx; c->a_idx.used += 3 * ntri; c->a_tc.used += nvtx;
    c->meshes.push_back(m);
    return (int)c->meshes.size() - 1;
}

static int scene_commit(fovpt_ctx* c, uint64_t* traversable_out);
int fovpt_scene_commit(fovpt_ctx* c, uint64_t* traversable_out)
{
    if (!c) return 1;
    try { return scene_commit(c, traversable_out); }
    catch (const std::exception& e) { return fail(c, std::string("fovpt_scene_commit: ") + e.what()); }      // no exception crosses the C ABI
}
static int scene_commit(fovpt_ctx* c, uint64_t* traversable_out)
{
    cudaSetDevice(c->device);
    const bool trace = getenv("FOVPT_BUILD_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_last = now();
    auto phase = [&](const char* name) { if (trace) { const double t = now(); fprintf(stderr, "[fovpt commit] %-27s %9.3f ms\n", name, t - t_last); t_last = t; } };
    free_scene(c);                                   // derived data only: the arenas filled by fovpt_scene_add_mesh stay
    const size_t nv = c->a_vtx.used / 3, nt = c->a_idx.used / 3;
    // per-mesh tables (small, host): where each mesh starts, its material and texture
    std::vector<uint2> table; std::vector<MaterialBlob> mats; std::vector<int2> mesh_tex;
    int any_catcher = 0;
    for (const MeshRec& m : c->meshes) {
        table.push_back(make_uint2((uint32_t)m.tbase, (uint32_t)m.vbase));
        mats.push_back(m.mat);
        const bool has_tex = m.texture_id >= 0 && m.texture_id < (int)c->textures.size();
        mesh_tex.push_back(make_int2(has_tex ? m.texture_id : -1, m.has_tc ? 1 : 0));
        if (m.mat.flags & kMaterialShadowCatcher) any_catcher = 1;
    }
    std::vector<int4> tex_info; std::vector<uint32_t> texels;
    for (auto& t : c->textures) {
        tex_info.push_back(make_int4((int)texels.size(), t.w, t.h, 0));
        texels.insert(texels.end(), t.px.begin(), t.px.end());
    }
    phase("mesh tables (host)");
    CU(dmalloc(&c->d_idx, 3 * nt)); CU(dmalloc(&c->d_tri_mesh, nt)); CU(dmalloc(&c->d_mesh_table, table.size()));
    CU(dmalloc(&c->d_mat, mats.size())); CU(dmalloc(&c->d_mesh_tex, mesh_tex.size()));
    CU(dmalloc(&c->d_tex_info, tex_info.size())); CU(dmalloc(&c->d_texels, texels.size()));
    CU(cudaMemcpyAsync(c->d_mesh_table, table.data(), table.size() * sizeof(uint2), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_mat, mats.data(), mats.size() * sizeof(MaterialBlob), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_mesh_tex, mesh_tex.data(), mesh_tex.size() * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_tex_info, tex_info.data(), tex_info.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_texels, texels.data(), texels.size() * 4, cudaMemcpyHostToDevice, c->stream));
    // mesh-local indices -> indices into the vertex soup, and the mesh id of every triangle, on the device
    if (nt) { launch_rebase_indices(c->stream, c->a_idx.p, c->d_mesh_table, (uint32_t)table.size(), (uint32_t)nt, c->d_idx, c->d_tri_mesh); c->launches++; }
    CU(cudaStreamSynchronize(c->stream));            // the host vectors above go out of scope
    (void)nv;
    phase("tables + index rebase");
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, c->stream);
    int leaf_max = 2;        // measured on C3: 2 -> 4.49 ms, 3 -> 4.52 ms, 4 -> 4.56 ms of k_trace per frame
    if (const char* e = getenv("FOVPT_LEAF_MAX")) { const int v = atoi(e); if (v >= 1 && v <= 4) leaf_max = v; }   // tuning knob
    cudaError_t be = build_bvh(c->a_vtx.p, c->d_idx, c->d_tri_mesh, (uint32_t)nt, leaf_max, kBvhWide, c->stream, &c->bvh, &c->launches);
    cudaEventRecord(e1, c->stream);
    if (be != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return cuda_fail(c, be, "build_bvh"); }
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&c->build_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    phase("build_bvh");
    static_assert(3 * kMaxWideDepth + 1 <= (uint32_t)kStackSize, "traversal stack must hold 3 deferred children per level");
    if (c->bvh.max_depth > kMaxWideDepth) return fail(c, "BVH deeper than the traversal stack allows (3 deferred children per level)");

    c->scene.bvh.nodes = c->bvh.nodes; c->scene.bvh.tris = c->bvh.tris;
    c->scene.bvh.n_nodes = c->bvh.n_nodes; c->scene.bvh.n_tris = c->bvh.n_tris;
    c->scene.tri_mesh = c->d_tri_mesh; c->scene.tri_vidx = c->d_idx; c->scene.texcoords = c->a_tc.p;
    c->scene.materials = c->d_mat; c->scene.mesh_tex = c->d_mesh_tex; c->scene.tex_info = c->d_tex_info; c->scene.texels = c->d_texels;
    c->scene.any_catcher = any_catcher;
    c->scene.n_meshes = (uint32_t)mats.size(); c->scene.n_textures = (uint32_t)tex_info.size(); c->scene.n_vertices = (uint32_t)nv;
    c->committed = true;
    if (traversable_out) *traversable_out = (uint64_t)(uintptr_t)c->bvh.nodes | 1ull;
    return 0;
}

int fovpt_probe_upload(fovpt_ctx* c, int w, int h, const float* rgba, const float* pdfX, const float* cdfX,
                       const float* pdfY, const float* cdfY, const float offset[3], void* probe_out64)
{
    if (!c) return 1;
    if (w <= 0 || h <= 0 || !rgba || !pdfX || !cdfX || !pdfY || !cdfY || !probe_out64) return fail(c, "Probe Data is not valid");   // Probe.h:104-105
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->p_data); cudaFree(c->p_pdfX); cudaFree(c->p_cdfX); cudaFree(c->p_pdfY); cudaFree(c->p_cdfY); cudaFree(c->p_gy); cudaFree(c->p_gx);
    c->p_gy = c->p_gx = nullptr;
    const size_t n = (size_t)w * h;
    CU(dmalloc(&c->p_data, n)); CU(dmalloc(&c->p_pdfX, n)); CU(dmalloc(&c->p_cdfX, n)); CU(dmalloc(&c->p_pdfY, (size_t)h)); CU(dmalloc(&c->p_cdfY, (size_t)h));
    CU(cudaMemcpy(c->p_data, rgba, n * 16, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->p_pdfX, pdfX, n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->p_cdfX, cdfX, n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->p_pdfY, pdfY, (size_t)h * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->p_cdfY, cdfY, (size_t)h * 4, cudaMemcpyHostToDevice));
    ProbeBlob pb; memset(&pb, 0, sizeof(pb));
    pb.width = w; pb.height = h; pb.data = c->p_data;
    if (offset) memcpy(pb.offset, offset, 12);
    pb.pdfValuesX = c->p_pdfX; pb.cdfValuesX = c->p_cdfX; pb.pdfValuesY = c->p_pdfY; pb.cdfValuesY = c->p_cdfY;
    memcpy(probe_out64, &pb, sizeof(pb));
    if (w <= 65535 && h <= 65535) {                       // guide tables for the CDF inversion (exactly equivalent search)
        CU(dmalloc(&c->p_gy, (size_t)kGuide + 1)); CU(dmalloc(&c->p_gx, (size_t)h * (kGuide + 1)));
        launch_probe_guides(c->L, pb, c->p_gy, c->p_gx);
        c->launches++;
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
    }
    return 0;
}

int fovpt_resize(fovpt_ctx* c, int w, int h, void** accum_out, void** frame_out)
{
    if (!c) return 1;
    if (w <= 0 || h <= 0) return fail(c, "bad frame size");
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->fb_accum); cudaFree(c->fb_frame); cudaFree(c->fb_color); cudaFree(c->fb_normal); cudaFree(c->fb_albedo);
    c->fb_accum = nullptr; c->fb_frame = nullptr; c->fb_color = c->fb_normal = c->fb_albedo = nullptr;
    const size_t n = (size_t)w * h;
    CU(dmalloc(&c->fb_accum, n)); CU(dmalloc(&c->fb_frame, n));
    CU(cudaMemset(c->fb_accum, 0, n * 16)); CU(cudaMemset(c->fb_frame, 0, n * 4));
    c->fb_w = w; c->fb_h = h;
    c->last_frame_ptr = c->fb_frame; c->last_accum_ptr = c->fb_accum; c->last_npix = n;
    if (accum_out) *accum_out = c->fb_accum;
    if (frame_out) *frame_out = c->fb_frame;
    return 0;
}

int fovpt_feature_buffers(fovpt_ctx* c, void** color_out, void** normal_out, void** albedo_out)
{
    if (!c) return 1;
    if (c->fb_w <= 0 || c->fb_h <= 0) return fail(c, "call fovpt_resize first");
    cudaSetDevice(c->device);
    const size_t n = (size_t)c->fb_w * c->fb_h;
    if (!c->fb_color) {
        CU(dmalloc(&c->fb_color, n)); CU(dmalloc(&c->fb_normal, n)); CU(dmalloc(&c->fb_albedo, n));
        CU(cudaMemset(c->fb_color, 0, n * 16)); CU(cudaMemset(c->fb_normal, 0, n * 16)); CU(cudaMemset(c->fb_albedo, 0, n * 16));
    }
    if (color_out) *color_out = c->fb_color;
    if (normal_out) *normal_out = c->fb_normal;
    if (albedo_out) *albedo_out = c->fb_albedo;
    return 0;
}

int fovpt_download_features(fovpt_ctx* c, float* color_rgba, float* normal_rgba, float* albedo_rgba, size_t npix)
{
    if (!c) return 1;
    if (!c->fb_color || npix > (size_t)c->fb_w * c->fb_h) return fail(c, "no feature buffers of that size (fovpt_feature_buffers)");
    cudaSetDevice(c->device);
    if (color_rgba) CU(cudaMemcpyAsync(color_rgba, c->fb_color, npix * 16, cudaMemcpyDeviceToHost, c->stream));
    if (normal_rgba) CU(cudaMemcpyAsync(normal_rgba, c->fb_normal, npix * 16, cudaMemcpyDeviceToHost, c->stream));
    if (albedo_rgba) CU(cudaMemcpyAsync(albedo_rgba, c->fb_albedo, npix * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int fovpt_launch(fovpt_ctx* c, const void* lp248, size_t size, unsigned lw, unsigned lh)
{
    if (!c) return 1;
    if (!lp248 || size != sizeof(LaunchParamsBlob)) return fail(c, "launch params must be the 248-byte LaunchParams");
    LaunchParamsBlob lp; memcpy(&lp, lp248, sizeof(lp));
    if (lp.frame.size_x <= 0 || lp.frame.size_y <= 0) return fail(c, "frame.size not set");
    if (lp.frame.factor_x != lp.frame.factor_y) return fail(c, "anisotropic factor not supported");
    if (lw == 0 || lh == 0) return 0;
    try {
        std::vector<PassHost> p;
        p.push_back({lw, lh, lp.frame.factor_x, lp.frame.fillSize, lp.frame.r_inner, lp.frame.r_outer, lp.samples_per_launch,
                     lp.frame.offset_x, lp.frame.offset_y, lp.frame.redraw, lp.frame.subframe_index, 0u});
        return run_frame(c, lp, p);
    } catch (const std::exception& e) { return fail(c, std::string("fovpt_launch: ") + e.what()); }
}

static int render_frames(fovpt_ctx* c, void* lp248, int n_subframes);
int fovpt_render_frames(fovpt_ctx* c, void* lp248, int n_subframes)
{
    if (!c) return 1;
    try { return render_frames(c, lp248, n_subframes); }
    catch (const std::exception& e) { return fail(c, std::string("fovpt_render_frames: ") + e.what()); }
}
static int render_frames(fovpt_ctx* c, void* lp248, int n_subframes)
{
    if (!lp248) return fail(c, "null launch params");
    if (n_subframes < 1 || n_subframes > 65536) return fail(c, "n_subframes out of range");
    LaunchParamsBlob* lp = (LaunchParamsBlob*)lp248;
    if (lp->frame.size_x == 0) return 0;                  // SimplePathtracer.cpp:81-82
    std::vector<PassHost> p = schedule(c->cfg, *lp);
    // as many subframes per wavefront as the path budget holds (C3: all 8; C5: one, itself cut into row chunks)
    const size_t per_layer = std::max<size_t>(paths_per_layer(p), 1);
    const int per_batch = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_subframes, kPathBudget / per_layer));
    for (int done = 0; done < n_subframes;) {
        const int nsub = std::min(per_batch, n_subframes - done);
        LaunchParamsBlob tmp; memcpy(&tmp, lp, sizeof(tmp));
        int rc = run_frame(c, tmp, p, nsub);
        if (rc) return rc;
        done += nsub;
        lp->frame.subframe_index += (unsigned)nsub;
        for (auto& ph : p) ph.subframe += (unsigned)nsub * ph.sub_step;
    }
    // what render() leaves behind in the host copy of the params (:137-211)
    const PassHost& last = p.back();
    lp->frame.factor_x = lp->frame.factor_y = last.factor; lp->frame.factor_z = 1;
    lp->frame.fillSize = last.fill; lp->frame.r_inner = last.r_inner; lp->frame.r_outer = last.r_outer;
    lp->samples_per_launch = last.spp; lp->frame.offset_x = last.off_x; lp->frame.offset_y = last.off_y; lp->frame.redraw = last.redraw;
    return 0;
}

int fovpt_render_frame(fovpt_ctx* c, void* lp248) { return fovpt_render_frames(c, lp248, 1); }

int fovpt_sync(fovpt_ctx* c)
{
    if (!c) return 1;
    cudaSetDevice(c->device);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    collect_times(c);
    if (const unsigned site = read_check_failure()) return fail(c, "bounds check " + std::to_string(site) + " failed in a kernel (FOVPT_CHECKED build, sites in csrc/bvh.cuh / wavefront.cu)");
    if (c->peer.open) {                       // a wait on the shared frame that timed out left the number of the missing rank + 1
        uint32_t* where = c->peer.owner ? c->peer.flags + kPeerError : c->peer.d_err;
        uint32_t e = 0;
        CU(cudaMemcpy(&e, where, 4, cudaMemcpyDeviceToHost));
        if (e) {
            CU(cudaMemset(where, 0, 4));
            return fail(c, c->peer.owner ? "peer frame: rank " + std::to_string(e - 1) + " did not deliver its tiles in time"
                                         : std::string("peer frame: the owner did not release the previous frame in time"));
        }
    }
    return 0;
}

int fovpt_download(fovpt_ctx* c, uint32_t* dst, size_t npix)
{
    if (!c || !dst) return 1;
    if (!c->last_frame_ptr || npix > c->last_npix) return fail(c, "no frame buffer of that size");
    cudaSetDevice(c->device);
    cudaPointerAttributes pa;
    const bool dst_pinned = cudaPointerGetAttributes(&pa, dst) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (dst_pinned) {                                     // caller's buffer is page-locked: DMA straight into it
        CU(cudaMemcpyAsync(dst, c->last_frame_ptr, npix * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    } else {                                              // pageable: stage through the context's pinned buffer
        if (ensure_pinned(c, npix * 4)) return 1;
        CU(cudaMemcpyAsync(c->h_pinned, c->last_frame_ptr, npix * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        memcpy(dst, c->h_pinned, npix * 4);
    }
    collect_times(c);
    return 0;
}

int fovpt_forget_frame(fovpt_ctx* c)
{
    if (!c) return 1;
    cudaSetDevice(c->device);
    CU(cudaStreamSynchronize(c->stream));
    c->last_frame_ptr = c->fb_frame; c->last_accum_ptr = c->fb_accum; c->last_npix = (size_t)c->fb_w * c->fb_h;
    return 0;
}

int fovpt_download_accum(fovpt_ctx* c, float* dst, size_t npix)
{
    if (!c || !dst) return 1;
    if (!c->last_accum_ptr || npix > c->last_npix) return fail(c, "no accumulation buffer of that size");
    cudaSetDevice(c->device);
    CU(cudaMemcpyAsync(dst, c->last_accum_ptr, npix * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    collect_times(c);
    return 0;
}

int fovpt_upload_accum(fovpt_ctx* c, const float* src, size_t npix)
{
    if (!c || !src) return 1;
    if (!c->fb_accum || npix != (size_t)c->fb_w * c->fb_h) return fail(c, "fovpt_upload_accum: size differs from the accumulation buffer of fovpt_resize");
    cudaSetDevice(c->device);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(c->fb_accum, src, npix * 16, cudaMemcpyHostToDevice));
    return 0;
}

int fovpt_trace(fovpt_ctx* c, const float* rays, size_t n, int mode, int32_t* tri, float* tuv)
{
    if (!c || !rays || !tri) return 1;
    if (!c->committed) return fail(c, "scene not committed");
    if (mode < 0 || mode > 3) return fail(c, "mode must be 0 (closest), 1 (any), 2/3 (the same with visit counters in tuv)");
    if (n == 0) return 0;
    cudaSetDevice(c->device);
    float* d_r = nullptr; int32_t* d_t = nullptr; float* d_tuv = nullptr;
    CU(dmalloc(&d_r, 8 * n)); CU(dmalloc(&d_t, n)); CU(dmalloc(&d_tuv, 3 * n));
    CU(cudaMemcpyAsync(d_r, rays, 32 * n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(d_tuv, 0, 12 * n, c->stream));
    launch_trace_query(c->L, c->scene, d_r, n, mode, d_t, d_tuv);
    c->launches++;
    CU(cudaMemcpyAsync(tri, d_t, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    if (tuv) CU(cudaMemcpyAsync(tuv, d_tuv, 12 * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    cudaFree(d_r); cudaFree(d_t); cudaFree(d_tuv);
    return 0;
}

int fovpt_texture_check(fovpt_ctx* c, int texture_id, const float* uv, size_t n, int mode, float* sw_rgba, float* hw_rgba)
{
    if (!c || !uv || !sw_rgba || !hw_rgba) return 1;
    if (!c->committed) return fail(c, "scene not committed");
    if (texture_id < 0 || texture_id >= (int)c->textures.size()) return fail(c, "no such texture");
    if (n == 0) return 0;
    cudaSetDevice(c->device);
    const HostTex& t = c->textures[(size_t)texture_id];
    // the reference's texture object, field by field (SimplePathtracer.cpp:748-799)
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
    cudaArray_t arr = nullptr;
    CU(cudaMallocArray(&arr, &cd, (size_t)t.w, (size_t)t.h));
    const size_t pitch = (size_t)t.w * 4;
    CU(cudaMemcpy2DToArray(arr, 0, 0, t.px.data(), pitch, pitch, (size_t)t.h, cudaMemcpyHostToDevice));
    cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td; memset(&td, 0, sizeof(td));
    td.addressMode[0] = cudaAddressModeWrap; td.addressMode[1] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 1;
    td.maxAnisotropy = 1; td.maxMipmapLevelClamp = 99; td.minMipmapLevelClamp = 0; td.mipmapFilterMode = cudaFilterModePoint;
    td.borderColor[0] = 1.0f; td.sRGB = 0;
    cudaTextureObject_t tex = 0;
    CU(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    float2* d_uv = nullptr; float4 *d_sw = nullptr, *d_hw = nullptr;
    CU(dmalloc(&d_uv, n)); CU(dmalloc(&d_sw, n)); CU(dmalloc(&d_hw, n));
    CU(cudaMemcpyAsync(d_uv, uv, n * 8, cudaMemcpyHostToDevice, c->stream));
    int4 info; CU(cudaMemcpy(&info, c->d_tex_info + texture_id, sizeof(int4), cudaMemcpyDeviceToHost));
    launch_texture_probe(c->L, tex, c->d_texels + info.x, t.w, t.h, d_uv, n, mode, d_sw, d_hw);
    c->launches++;
    CU(cudaMemcpyAsync(sw_rgba, d_sw, n * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(hw_rgba, d_hw, n * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    cudaDestroyTextureObject(tex); cudaFreeArray(arr); cudaFree(d_uv); cudaFree(d_sw); cudaFree(d_hw);
    return 0;
}

int fovpt_primary_hits(fovpt_ctx* c, const void* lp248, unsigned lw, unsigned lh, int32_t* tri, float* tuv, float* dir)
{
    if (!c || !lp248 || !tri || !tuv || !dir) return 1;
    if (!c->committed) return fail(c, "scene not committed");
    LaunchParamsBlob lp; memcpy(&lp, lp248, sizeof(lp));
    std::vector<PassHost> p;
    p.push_back({lw, lh, lp.frame.factor_x, lp.frame.fillSize, lp.frame.r_inner, lp.frame.r_outer, lp.samples_per_launch,
                 lp.frame.offset_x, lp.frame.offset_y, lp.frame.redraw, lp.frame.subframe_index, 0u});
    FrameDev F; size_t total_li = 0;
    fill_frame(c, lp, p, &F, &total_li);
    cudaSetDevice(c->device);
    const size_t n = (size_t)lw * lh;
    int32_t* d_t = nullptr; float* d_tuv = nullptr; float* d_dir = nullptr;
    CU(dmalloc(&d_t, n)); CU(dmalloc(&d_tuv, 3 * n)); CU(dmalloc(&d_dir, 3 * n));
    CU(cudaMemsetAsync(d_tuv, 0, 12 * n, c->stream)); CU(cudaMemsetAsync(d_dir, 0, 12 * n, c->stream));
    launch_primary_hits(c->L, F, c->scene, d_t, d_tuv, d_dir);
    c->launches++;
    CU(cudaMemcpyAsync(tri, d_t, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(tuv, d_tuv, 12 * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(dir, d_dir, 12 * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    cudaFree(d_t); cudaFree(d_tuv); cudaFree(d_dir);
    return 0;
}

int fovpt_set_tile_owner(fovpt_ctx* c, int tile_size, int tiles_x, int tiles_y, const uint8_t* owner, int rank)
{
    if (!c) return 1;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_tile_owner); c->d_tile_owner = nullptr;
    c->owner_key.clear();
    if (!owner) { c->tile_size = c->tiles_x = c->tiles_y = 0; c->rank = 0; return 0; }
    if (tile_size <= 0 || tiles_x <= 0 || tiles_y <= 0) return fail(c, "bad tile grid");
    if (rank < 0 || rank > 255) return fail(c, "rank must be 0..255 (the owner table holds one byte per tile)");
    CU(dmalloc(&c->d_tile_owner, (size_t)tiles_x * tiles_y));
    CU(cudaMemcpy(c->d_tile_owner, owner, (size_t)tiles_x * tiles_y, cudaMemcpyHostToDevice));
    c->tile_size = tile_size; c->tiles_x = tiles_x; c->tiles_y = tiles_y; c->rank = rank;
    return 0;
}

int fovpt_get_stats(fovpt_ctx* c, fovpt_stats* s)
{
    if (!c || !s) return 1;
    cudaSetDevice(c->device);
    CU(cudaStreamSynchronize(c->stream));
    collect_times(c);
    DevStats d;
    CU(cudaMemcpy(&d, c->d_stats, sizeof(d), cudaMemcpyDeviceToHost));
    memset(s, 0, sizeof(*s));
    s->samples = d.samples; s->radiance_rays = d.radiance_rays; s->shadow_rays = d.shadow_rays;
#ifdef FOVPT_TRACE_STATS
    {
        const unsigned long long* q = d.dbg;
        fprintf(stderr, "[trace stats] node_exec %llu lanes/exec %.2f (per exec: blocked %.2f, empty while queue open %.2f, empty after queue drained %.2f) | "
                        "leaf_exec %llu lanes/exec %.2f tris/lane %.2f nmax %.2f | fetch rounds %llu\n",
                q[0], q[0] ? (double)q[1] / q[0] : 0., q[0] ? (double)q[10] / q[0] : 0., q[0] ? (double)q[8] / q[0] : 0., q[0] ? (double)q[9] / q[0] : 0.,
                q[2], q[2] ? (double)q[3] / q[2] : 0., q[3] ? (double)q[4] / q[3] : 0., q[2] ? (double)q[5] / q[2] : 0., q[6]);
    }
#endif
    s->kernel_launches = c->launches; s->frames = c->frames; s->trace_launches = c->trace_launches;
    s->bvh_nodes = c->bvh.n_nodes; s->bvh_tris = c->bvh.n_tris; s->bvh_build_ms = c->build_ms;
    s->last_frame_ms = c->last_ms[0]; s->last_raygen_ms = c->last_ms[K_RAYGEN]; s->last_extend_ms = c->last_ms[K_EXTEND];
    s->last_shade_ms = c->last_ms[K_SHADE]; s->last_shadow_ms = c->last_ms[K_SHADOW]; s->last_resolve_ms = c->last_ms[K_RESOLVE];
    return 0;
}

int fovpt_reset_stats(fovpt_ctx* c)
{
    if (!c) return 1;
    cudaSetDevice(c->device);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemset(c->d_stats, 0, sizeof(DevStats)));
    c->launches = 0; c->frames = 0; c->trace_launches = 0;
    return 0;
}

int fovpt_bvh_info(fovpt_ctx* c, uint64_t* n_nodes, uint64_t* n_tris)
{
    if (!c || !c->committed) return 1;
    if (n_nodes) *n_nodes = c->bvh.n_nodes;
    if (n_tris) *n_tris = c->bvh.n_tris;
    return 0;
}

int fovpt_bvh_layout(fovpt_ctx* c, int* wide, int* node_bytes)
{
    if (!c) return 1;
    if (wide) *wide = kBvhWide;
    if (node_bytes) *node_bytes = 64;
    return 0;
}

int fovpt_bvh_read(fovpt_ctx* c, void* nodes64, uint32_t* sorted_tri_ids)
{
    if (!c || !c->committed) return 1;
    cudaSetDevice(c->device);
    if (nodes64 && c->bvh.n_nodes) CU(cudaMemcpy(nodes64, c->bvh.nodes, (size_t)c->bvh.n_nodes * 64, cudaMemcpyDeviceToHost));
    if (sorted_tri_ids && c->bvh.n_tris) CU(cudaMemcpy(sorted_tri_ids, c->bvh.sorted_ids, (size_t)c->bvh.n_tris * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// ---- shared frame over peer memory.  handle = { cudaIpcMemHandle_t (64 B) | npix u64 | world i32 | owner rank i32 | magic u32 }
namespace {
struct PeerHandleBlob { cudaIpcMemHandle_t mem; unsigned long long npix; int world, owner_rank; unsigned magic; unsigned pad_[11]; };
static_assert(sizeof(PeerHandleBlob) == sizeof(fovpt_peer_handle), "fovpt_peer_handle is 128 opaque bytes");
constexpr unsigned kPeerMagic = 0x46505446u;
size_t peer_frame_bytes(size_t npix) { return (npix * 4 + 255) & ~(size_t)255; }
long long peer_timeout_clocks(int device)
{
    double ms = 5000.0;
    if (const char* e = getenv("FOVPT_PEER_TIMEOUT_MS")) ms = atof(e) > 0 ? atof(e) : ms;
    int khz = 1900000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    return (long long)(ms * (double)khz);
}
}  // namespace

int fovpt_peer_frame_create(fovpt_ctx* c, size_t npix, int world, int rank, fovpt_peer_handle* handle_out, uint32_t** frame_out)
{
    if (!c || !handle_out || !frame_out) return 1;
    if (c->peer.open) return fail(c, "a peer frame is already open on this context");
    if (npix == 0 || world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world) return fail(c, "peer frame: bad size, world or rank");
    cudaSetDevice(c->device);
    const size_t fb = peer_frame_bytes(npix);
    unsigned char* base = nullptr;
    CU(cudaMalloc((void**)&base, fb + kPeerFlagWords * 4 + npix * sizeof(float4)));
    CU(cudaMemset(base, 0, fb + kPeerFlagWords * 4 + npix * sizeof(float4)));
    PeerHandleBlob h; memset(&h, 0, sizeof(h));
    cudaError_t e = cudaIpcGetMemHandle(&h.mem, base);
    if (e != cudaSuccess) { cudaFree(base); return cuda_fail(c, e, "cudaIpcGetMemHandle"); }
    h.npix = npix; h.world = world; h.owner_rank = rank; h.magic = kPeerMagic;
    memcpy(handle_out, &h, sizeof(h));
    c->peer = fovpt_ctx::PeerFrame{};
    c->peer.base = base; c->peer.flags = (uint32_t*)(base + fb); c->peer.accum = (float4*)(base + fb + kPeerFlagWords * 4);
    c->peer.open = true; c->peer.owner = true;
    c->peer.world = world; c->peer.owner_rank = rank; c->peer.rank = rank; c->peer.npix = npix;
    c->peer.timeout_clocks = peer_timeout_clocks(c->device);
    *frame_out = (uint32_t*)base;
    return 0;
}

int fovpt_peer_frame_open(fovpt_ctx* c, const fovpt_peer_handle* handle, int rank, uint32_t** frame_out)
{
    if (!c || !handle || !frame_out) return 1;
    if (c->peer.open) return fail(c, "a peer frame is already open on this context");
    PeerHandleBlob h; memcpy(&h, handle, sizeof(h));
    if (h.magic != kPeerMagic || h.world < 1 || h.world > kPeerMaxWorld) return fail(c, "peer frame: not a handle from fovpt_peer_frame_create");
    if (rank < 0 || rank >= h.world || rank == h.owner_rank) return fail(c, "peer frame: rank out of range (or the owner's)");
    cudaSetDevice(c->device);
    void* base = nullptr;
    CU(cudaIpcOpenMemHandle(&base, h.mem, cudaIpcMemLazyEnablePeerAccess));
    uint32_t* d_err = nullptr;
    cudaError_t e = dmalloc(&d_err, 1);
    if (e == cudaSuccess) e = cudaMemset(d_err, 0, 4);
    if (e != cudaSuccess) { cudaIpcCloseMemHandle(base); return cuda_fail(c, e, "peer frame error word"); }
    c->peer = fovpt_ctx::PeerFrame{};
    c->peer.base = (unsigned char*)base; c->peer.flags = (uint32_t*)((unsigned char*)base + peer_frame_bytes(h.npix)); c->peer.d_err = d_err;
    c->peer.accum = (float4*)((unsigned char*)base + peer_frame_bytes(h.npix) + kPeerFlagWords * 4);
    c->peer.open = true; c->peer.owner = false; c->peer.world = h.world; c->peer.owner_rank = h.owner_rank; c->peer.rank = rank;
    c->peer.npix = (size_t)h.npix; c->peer.timeout_clocks = peer_timeout_clocks(c->device);
    *frame_out = (uint32_t*)base;
    return 0;
}

int fovpt_peer_frame_accum(fovpt_ctx* c, int mirror, float** accum_out)
{
    if (!c) return 1;
    if (!c->peer.open) return fail(c, "peer frame: not open");
    c->peer.mirror_accum = mirror != 0;
    if (accum_out) *accum_out = (float*)c->peer.accum;
    return 0;
}

int fovpt_peer_frame_signal(fovpt_ctx* c)
{
    if (!c) return 1;
    if (!c->peer.open || c->peer.owner) return fail(c, "peer frame: signal is the contributor's call");
    cudaSetDevice(c->device);
    launch_peer_signal(c->L, c->peer.flags + c->peer.rank, ++c->peer.seq);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

int fovpt_peer_frame_wait(fovpt_ctx* c)
{
    if (!c) return 1;
    if (!c->peer.open || !c->peer.owner) return fail(c, "peer frame: wait is the owner's call");
    cudaSetDevice(c->device);
    ++c->peer.seq;
    if (c->peer.world > 1) {
        launch_peer_wait(c->L, c->peer.flags, c->peer.world, c->peer.owner_rank, c->peer.seq, c->peer.flags + kPeerError, c->peer.timeout_clocks);
        c->launches++;
    }
    CU(cudaGetLastError());
    return 0;
}

int fovpt_peer_frame_release(fovpt_ctx* c)
{
    if (!c) return 1;
    if (!c->peer.open || !c->peer.owner) return fail(c, "peer frame: release is the owner's call");
    cudaSetDevice(c->device);
    launch_peer_signal(c->L, c->peer.flags + kPeerRelease, c->peer.seq);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

int fovpt_peer_frame_close(fovpt_ctx* c)
{
    if (!c) return 1;
    if (!c->peer.open) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->last_frame_ptr == (uint32_t*)c->peer.base) { c->last_frame_ptr = nullptr; c->last_npix = 0; }
    if (c->last_accum_ptr == c->peer.accum) c->last_accum_ptr = nullptr;
    if (c->peer.owner) cudaFree(c->peer.base);
    else { cudaIpcCloseMemHandle(c->peer.base); cudaFree(c->peer.d_err); }
    c->peer = fovpt_ctx::PeerFrame{};
    cudaGetLastError();
    return 0;
}

}  // extern "C"
