// oracle/bvh.h -- TEST INFRASTRUCTURE (CPU oracle), never linked into the product library.
//
// The part of the path that the reference delegates to OptiX (optixAccelBuild and optixTrace,
// PT_sv5_/SimplePathtracer.cpp:602-746, PT_sv5_/deviceProgram.cu:196-248) has no source in
// the reference.  This header is the oracle's canonical stand-in:
//   * a deterministic binned-SAH BVH2 (16 bins, leaves <= 4 triangles),
//   * closest-hit traversal: min t over tmin < t < tmax, ties -> lowest global triangle id,
//   * any-hit traversal with back-face culling (occlusion semantics of deviceProgram.cu:224-300),
//   * brute-force versions of both, used by tests to show the BVH never changes a result,
//   * visit counters (n_node, n_tri) that define the ALGORITHMIC bytes per ray of SURVEY 8(d).
// The ray/triangle arithmetic itself is include/fovpt_exact.h, shared verbatim with the kernels.
// "Parity unpinned": no OptiX build exists to check these semantics against; see DESIGN.md.
#pragma once
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../include/fovpt_exact.h"

namespace obvh {

struct Hit { int mesh; int prim; uint32_t tri; float t, u, v; };

struct Counters {
    std::atomic<uint64_t> closest_rays{0}, any_rays{0};
    std::atomic<uint64_t> closest_nodes{0}, closest_tris{0};
    std::atomic<uint64_t> any_nodes{0}, any_tris{0};
    void reset() { closest_rays = any_rays = closest_nodes = closest_tris = any_nodes = any_tris = 0; }
};

struct Node {            // 64 bytes, the unit SURVEY 8(d) charges per inner-node visit
    float lo[2][3];
    float hi[2][3];
    int32_t child[2];    // >= 0: inner node index ; < 0: leaf, first = ~child
    int32_t count[2];    // triangles in the leaf (0 for inner)
};

struct Scene {
    std::vector<fx_tri>   tris;       // global triangle soup, mesh-major order
    std::vector<uint32_t> tri_mesh;   // mesh of each global triangle
    std::vector<uint32_t> mesh_first; // first global triangle of each mesh
    std::vector<Node>     nodes;      // nodes[0] is the root (if tris.size() > 1 leaf capacity)
    std::vector<uint32_t> order;      // leaf order -> global triangle id
    float                 pad = 0.f;
    bool                  single_leaf = false;

    void clear() { tris.clear(); tri_mesh.clear(); mesh_first.clear(); nodes.clear(); order.clear(); }

    void add_mesh(const float* vtx, const uint32_t* idx, size_t ntri)
    {
        uint32_t m = (uint32_t)mesh_first.size();
        mesh_first.push_back((uint32_t)tris.size());
        for (size_t i = 0; i < ntri; ++i) {
            const float* a = vtx + 3 * (size_t)idx[3 * i + 0];
            const float* b = vtx + 3 * (size_t)idx[3 * i + 1];
            const float* c = vtx + 3 * (size_t)idx[3 * i + 2];
            fx_vec3 A = {a[0], a[1], a[2]}, B = {b[0], b[1], b[2]}, C = {c[0], c[1], c[2]};
            tris.push_back(fx_make_tri(A, B, C));
            tri_mesh.push_back(m);
        }
    }

    // ---- build -----------------------------------------------------------------------
    struct Box {
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        void grow(const float* p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
        void grow(const Box& b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
        float area() const {
            float d[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
            if (d[0] < 0) return 0.f;
            return 2.f * (d[0] * d[1] + d[1] * d[2] + d[2] * d[0]);
        }
    };
    static Box tri_box(const fx_tri& t)
    {
        Box b;
        float v0[3] = {t.v0.x, t.v0.y, t.v0.z};
        float v1[3] = {t.v1.x, t.v1.y, t.v1.z};
        float v2[3] = {t.v2.x, t.v2.y, t.v2.z};
        b.grow(v0); b.grow(v1); b.grow(v2);
        // an ulp-scale margin on top of the exact vertex bounds (the global pad of commit() is what absorbs the tolerance of the ray/triangle test)
        for (int a = 0; a < 3; ++a) {
            float m = std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a])) * 2.4e-7f;
            b.lo[a] -= m; b.hi[a] += m;
        }
        return b;
    }

    struct BuildRef { Box box; float c[3]; uint32_t id; };

    void commit()
    {
        nodes.clear(); order.clear();
        size_t n = tris.size();
        if (n == 0) return;
        std::vector<BuildRef> refs(n);
        Box all;
        for (size_t i = 0; i < n; ++i) {
            refs[i].box = tri_box(tris[i]);
            refs[i].id = (uint32_t)i;
            for (int a = 0; a < 3; ++a) refs[i].c[a] = 0.5f * (refs[i].box.lo[a] + refs[i].box.hi[a]);
            all.grow(refs[i].box);
        }
        // global conservative padding of every box: the slab test is then allowed ulp-level
        // error without ever rejecting a box whose triangle the exact test accepts.
        float S = 0.f;
        for (int a = 0; a < 3; ++a) S = std::max(S, std::max(std::fabs(all.lo[a]), std::fabs(all.hi[a])));
        pad = S * (1.0f / 65536.0f);
        order.reserve(n);
        if (n <= 4) {
            single_leaf = true;
            for (size_t i = 0; i < n; ++i) order.push_back((uint32_t)i);
            return;
        }
        single_leaf = false;
        nodes.reserve(n);
        nodes.emplace_back();
        build_node(0, refs, 0, n);
    }

    // returns child encoding for range [b,e): leaf if small, else recurses.
    void make_child(int node, int side, std::vector<BuildRef>& refs, size_t b, size_t e)
    {
        Box bx;
        for (size_t i = b; i < e; ++i) bx.grow(refs[i].box);
        for (int a = 0; a < 3; ++a) { nodes[node].lo[side][a] = bx.lo[a] - pad; nodes[node].hi[side][a] = bx.hi[a] + pad; }
        if (e - b <= 4 && !should_split_small(refs, b, e)) {
            // deterministic leaf order: ascending global id
            std::sort(refs.begin() + b, refs.begin() + e, [](const BuildRef& x, const BuildRef& y) { return x.id < y.id; });
            nodes[node].child[side] = ~(int32_t)order.size();
            nodes[node].count[side] = (int32_t)(e - b);
            for (size_t i = b; i < e; ++i) order.push_back(refs[i].id);
        } else {
            int c = (int)nodes.size();
            nodes.emplace_back();
            nodes[node].child[side] = c;
            nodes[node].count[side] = 0;
            build_node(c, refs, b, e);
        }
    }
    static bool should_split_small(std::vector<BuildRef>&, size_t, size_t) { return false; }

    void build_node(int node, std::vector<BuildRef>& refs, size_t b, size_t e)
    {
        const int NB = 16;
        Box cb;
        for (size_t i = b; i < e; ++i) cb.grow(refs[i].c);
        int best_axis = -1, best_bin = -1;
        float best_cost = FLT_MAX;
        for (int a = 0; a < 3; ++a) {
            float ext = cb.hi[a] - cb.lo[a];
            if (!(ext > 0.f)) continue;
            Box bb[NB]; int cnt[NB] = {0};
            float scale = NB / ext;
            for (size_t i = b; i < e; ++i) {
                int k = std::min(NB - 1, std::max(0, (int)((refs[i].c[a] - cb.lo[a]) * scale)));
                bb[k].grow(refs[i].box); cnt[k]++;
            }
            float la[NB], ra[NB]; int lc[NB], rc[NB];
            Box acc; int c = 0;
            for (int k = 0; k < NB; ++k) { acc.grow(bb[k]); c += cnt[k]; la[k] = acc.area(); lc[k] = c; }
            acc = Box(); c = 0;
            for (int k = NB - 1; k >= 0; --k) { acc.grow(bb[k]); c += cnt[k]; ra[k] = acc.area(); rc[k] = c; }
            for (int k = 0; k < NB - 1; ++k) {
                if (lc[k] == 0 || rc[k + 1] == 0) continue;
                float cost = la[k] * lc[k] + ra[k + 1] * rc[k + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = k; }
            }
        }
        size_t mid;
        if (best_axis < 0) {
            mid = (b + e) / 2;    // all centroids coincide: median split by id order
            std::sort(refs.begin() + b, refs.begin() + e, [](const BuildRef& x, const BuildRef& y) { return x.id < y.id; });
        } else {
            float ext = cb.hi[best_axis] - cb.lo[best_axis];
            float scale = NB / ext, lo = cb.lo[best_axis];
            int a = best_axis, kb = best_bin;
            auto it = std::stable_partition(refs.begin() + b, refs.begin() + e, [&](const BuildRef& r) {
                int k = std::min(NB - 1, std::max(0, (int)((r.c[a] - lo) * scale)));
                return k <= kb;
            });
            mid = (size_t)(it - refs.begin());
            if (mid == b || mid == e) mid = (b + e) / 2;
        }
        make_child(node, 0, refs, b, mid);
        make_child(node, 1, refs, mid, e);
    }

    // ---- traversal -------------------------------------------------------------------
    struct RayPrep { float o[3], inv[3]; };

    static inline bool slab(const float* lo, const float* hi, const RayPrep& r, float tmin, float tmax, float* tnear)
    {
        float t0 = tmin, t1 = tmax;
        for (int a = 0; a < 3; ++a) {
            float ta = (lo[a] - r.o[a]) * r.inv[a];
            float tb = (hi[a] - r.o[a]) * r.inv[a];
            float tn = fminf(ta, tb), tf = fmaxf(ta, tb);   // fmin/fmax drop NaN (0*inf)
            t0 = fmaxf(t0, tn); t1 = fminf(t1, tf);
        }
        *tnear = t0;
        return t0 <= t1 * 1.0000012f + 0.0f;
    }

    inline void leaf_closest(uint32_t first, int count, const fx_rayspace& d, float tmin, float tmax,
                             float& best_t, uint32_t& best_id, float& bu, float& bv, uint64_t& ntri) const
    {
        for (int i = 0; i < count; ++i) {
            uint32_t id = order[first + i];
            float t, u, v;
            ++ntri;
            if (fx_ray_tri_wt(d, tris[id], &t, &u, &v) == 0) continue;
            if (!(t > tmin && t < tmax)) continue;
            if (fx_hit_better(t, id, best_t, best_id)) { best_t = t; best_id = id; bu = u; bv = v; }
        }
    }

    bool closest(const fx_vec3& o, const fx_vec3& dir, float tmin, float tmax, Hit* hit, Counters* ctr = nullptr) const
    {
        const fx_rayspace d = fx_rayspace_init(o, dir);                // per-ray constants of the watertight triangle test
        float best_t = INFINITY, bu = 0, bv = 0; uint32_t best_id = 0xffffffffu;
        uint64_t nnode = 0, ntri = 0;
        if (!tris.empty()) {
            if (single_leaf) {
                leaf_closest(0, (int)order.size(), d, tmin, tmax, best_t, best_id, bu, bv, ntri);
            } else {
                RayPrep r = {{o.x, o.y, o.z}, {1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z}};
                int stack[128]; int sp = 0; stack[sp++] = 0;
                while (sp) {
                    const Node& nd = nodes[stack[--sp]];
                    ++nnode;
                    float tn[2]; bool h[2];
                    float lim = fminf(tmax, best_t);
                    h[0] = slab(nd.lo[0], nd.hi[0], r, tmin, lim, &tn[0]);
                    h[1] = slab(nd.lo[1], nd.hi[1], r, tmin, lim, &tn[1]);
                    int first = (h[0] && h[1] && tn[1] < tn[0]) ? 1 : 0;
                    for (int k = 0; k < 2; ++k) {
                        int s = k == 0 ? first : 1 - first;
                        if (!h[s]) continue;
                        if (nd.child[s] < 0) {
                            leaf_closest((uint32_t)~nd.child[s], nd.count[s], d, tmin, tmax, best_t, best_id, bu, bv, ntri);
                        }
                    }
                    // push far inner child first so the near one is popped first
                    for (int k = 1; k >= 0; --k) {
                        int s = k == 0 ? first : 1 - first;
                        if (h[s] && nd.child[s] >= 0) stack[sp++] = nd.child[s];
                    }
                }
            }
        }
        if (ctr) { ctr->closest_rays++; ctr->closest_nodes += nnode; ctr->closest_tris += ntri; }
        if (best_id == 0xffffffffu) return false;
        hit->tri = best_id; hit->mesh = (int)tri_mesh[best_id]; hit->prim = (int)(best_id - mesh_first[hit->mesh]);
        hit->t = best_t; hit->u = bu; hit->v = bv;
        return true;
    }

    // popped nodes are re-tested against the shrunken interval implicitly: the slab test above
    // was done with the limit at push time; a stale entry only costs work, never correctness.

    inline bool leaf_any(uint32_t first, int count, const fx_rayspace& d, float tmin, float tmax, uint64_t& ntri) const
    {
        for (int i = 0; i < count; ++i) {
            uint32_t id = order[first + i];
            float t, u, v;
            ++ntri;
            if (fx_ray_tri_wt(d, tris[id], &t, &u, &v) != 1) continue;     // front faces only
            if (t > tmin && t < tmax) return true;
        }
        return false;
    }

    bool any_front(const fx_vec3& o, const fx_vec3& dir, float tmin, float tmax, Counters* ctr = nullptr) const
    {
        const fx_rayspace d = fx_rayspace_init(o, dir);
        uint64_t nnode = 0, ntri = 0; bool occ = false;
        if (!tris.empty()) {
            if (single_leaf) {
                occ = leaf_any(0, (int)order.size(), d, tmin, tmax, ntri);
            } else {
                RayPrep r = {{o.x, o.y, o.z}, {1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z}};
                int stack[128]; int sp = 0; stack[sp++] = 0;
                while (sp && !occ) {
                    const Node& nd = nodes[stack[--sp]];
                    ++nnode;
                    for (int s = 0; s < 2 && !occ; ++s) {
                        float tn;
                        if (!slab(nd.lo[s], nd.hi[s], r, tmin, tmax, &tn)) continue;
                        if (nd.child[s] < 0) occ = leaf_any((uint32_t)~nd.child[s], nd.count[s], d, tmin, tmax, ntri);
                        else stack[sp++] = nd.child[s];
                    }
                }
            }
        }
        if (ctr) { ctr->any_rays++; ctr->any_nodes += nnode; ctr->any_tris += ntri; }
        return occ;
    }

    // ---- brute force (tests only) ----------------------------------------------------
    bool closest_brute(const fx_vec3& o, const fx_vec3& d, float tmin, float tmax, Hit* hit) const
    {
        float best_t = INFINITY, bu = 0, bv = 0; uint32_t best_id = 0xffffffffu;
        for (uint32_t id = 0; id < tris.size(); ++id) {
            float t, u, v;
            if (fx_ray_tri(o, d, tris[id], &t, &u, &v) == 0) continue;
            if (!(t > tmin && t < tmax)) continue;
            if (fx_hit_better(t, id, best_t, best_id)) { best_t = t; best_id = id; bu = u; bv = v; }
        }
        if (best_id == 0xffffffffu) return false;
        hit->tri = best_id; hit->mesh = (int)tri_mesh[best_id]; hit->prim = (int)(best_id - mesh_first[hit->mesh]);
        hit->t = best_t; hit->u = bu; hit->v = bv;
        return true;
    }
    bool any_front_brute(const fx_vec3& o, const fx_vec3& d, float tmin, float tmax) const
    {
        for (uint32_t id = 0; id < tris.size(); ++id) {
            float t, u, v;
            if (fx_ray_tri(o, d, tris[id], &t, &u, &v) != 1) continue;
            if (t > tmin && t < tmax) return true;
        }
        return false;
    }
};

}  // namespace obvh
