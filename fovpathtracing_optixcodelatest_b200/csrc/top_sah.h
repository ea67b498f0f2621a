// top_sah.h -- the top of the BVH: binned surface-area-heuristic build over the clusters the bottom-up PLOC phase leaves
// (bvh_build.cu).  Plain C++ (no CUDA), so the builder's host logic is unit-tested on CPU (tests/c/top_sah_test.cpp).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <system_error>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace fovpt {
namespace topsah {

struct TopCluster { float lo[3]; int cid; float hi[3]; uint32_t cnt; };     // two 16-byte halves: (lo, id) (hi, triangle count)
struct TopNode { int left, right; float lo[3], hi[3]; uint32_t cnt; };

inline float half_area(const float lo[3], const float hi[3])
{
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}

// Builds the subtree over c[b, e) (reordering that range) into nodes[base, base + (e - b) - 1) and returns its id (first_id +
// index; a single cluster is its own subtree).  A subtree over m clusters has exactly m - 1 inner nodes, so every range of
// the node array belongs to one call: ids do not depend on which thread runs what.
inline int top_sah(TopCluster* c, size_t b, size_t e, int first_id, TopNode* nodes, size_t base, int depth)
{
    if (e - b == 1) return c[b].cid;
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f}, clo[3] = {3e38f, 3e38f, 3e38f}, chi[3] = {-3e38f, -3e38f, -3e38f};
    uint32_t total = 0;
    for (size_t i = b; i < e; ++i) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], c[i].lo[a]); hi[a] = std::max(hi[a], c[i].hi[a]);
            const float m = 0.5f * (c[i].lo[a] + c[i].hi[a]);
            clo[a] = std::min(clo[a], m); chi[a] = std::max(chi[a], m);
        }
        total += c[i].cnt;
    }
    size_t mid = b;
    const size_t m = e - b;
    constexpr size_t kSmall = 16;
    if (m == 2) {
        mid = b + 1;
    } else if (m <= kSmall) {
        // small node (half of all nodes have <= 8 clusters, and setting up 3 x 32 bins costs far more than they are worth there):
        // exact sweep -- every split of the centroid order of every axis
        TopCluster sorted[3][kSmall];
        float bc = 3e38f; size_t bk = 0; int ba = 0;
        for (int a = 0; a < 3; ++a) {
            TopCluster* tmp = sorted[a];
            std::copy(c + b, c + e, tmp);
            for (size_t i = 1; i < m; ++i) {                                // insertion sort by centroid
                const TopCluster x = tmp[i]; const float key = x.lo[a] + x.hi[a];
                size_t j = i;
                for (; j > 0 && tmp[j - 1].lo[a] + tmp[j - 1].hi[a] > key; --j) tmp[j] = tmp[j - 1];
                tmp[j] = x;
            }
            float ra[kSmall]; uint32_t rn[kSmall];
            float l3[3] = {3e38f, 3e38f, 3e38f}, h3[3] = {-3e38f, -3e38f, -3e38f}; uint32_t nn = 0;
            for (size_t k = m; k-- > 0;) {
                for (int d = 0; d < 3; ++d) { l3[d] = std::min(l3[d], tmp[k].lo[d]); h3[d] = std::max(h3[d], tmp[k].hi[d]); }
                nn += tmp[k].cnt; ra[k] = half_area(l3, h3); rn[k] = nn;
            }
            for (int d = 0; d < 3; ++d) { l3[d] = 3e38f; h3[d] = -3e38f; }
            nn = 0;
            for (size_t k = 0; k + 1 < m; ++k) {
                for (int d = 0; d < 3; ++d) { l3[d] = std::min(l3[d], tmp[k].lo[d]); h3[d] = std::max(h3[d], tmp[k].hi[d]); }
                nn += tmp[k].cnt;
                const float cost = half_area(l3, h3) * (float)nn + ra[k + 1] * (float)rn[k + 1];
                const size_t off_centre = k + 1 > m / 2 ? k + 1 - m / 2 : m / 2 - (k + 1), best_off = bk > m / 2 ? bk - m / 2 : m / 2 - bk;
                if (cost < bc || (cost == bc && off_centre < best_off)) { bc = cost; bk = k + 1; ba = a; }
            }
        }
        if (bk) { std::copy(sorted[ba], sorted[ba] + m, c + b); mid = b + bk; }
    }
    constexpr int kBins = 32;
    const int nb = (int)std::min<size_t>(kBins, std::max<size_t>(4, m / 4));      // bins in use: the set-up and the two scans are per node, not per cluster
    struct Bin { float lo[3], hi[3]; uint32_t n; };
    float best_cost = 3e38f; int best_axis = -1, best_bin = 0;
    float sc[3]; bool use[3];
    for (int a = 0; a < 3; ++a) { use[a] = m > kSmall && depth < 48 && chi[a] - clo[a] > 0.f; sc[a] = use[a] ? (float)nb / (chi[a] - clo[a]) : 0.f; }
    if (use[0] || use[1] || use[2]) {                      // deeper than 48 levels (adversarial input): halve the list
        Bin bins[3][kBins];
        for (int a = 0; a < 3; ++a) for (int k = 0; k < nb; ++k) { bins[a][k].n = 0; for (int d = 0; d < 3; ++d) { bins[a][k].lo[d] = 3e38f; bins[a][k].hi[d] = -3e38f; } }
#if defined(__SSE2__)
        // one pass over the clusters fills the bins of all three axes; boxes move as 16-byte halves (lane 3 = id / count bits, ignored)
        alignas(16) float vlo[3][kBins][4], vhi[3][kBins][4];
        for (int a = 0; a < 3; ++a) for (int k = 0; k < nb; ++k) { _mm_store_ps(vlo[a][k], _mm_set1_ps(3e38f)); _mm_store_ps(vhi[a][k], _mm_set1_ps(-3e38f)); }
        for (size_t i = b; i < e; ++i) {
            const __m128 l = _mm_loadu_ps(c[i].lo), h = _mm_loadu_ps(c[i].hi);
            for (int a = 0; a < 3; ++a) {
                if (!use[a]) continue;
                const int k = std::min(nb - 1, std::max(0, (int)((0.5f * (c[i].lo[a] + c[i].hi[a]) - clo[a]) * sc[a])));
                bins[a][k].n += c[i].cnt;
                _mm_store_ps(vlo[a][k], _mm_min_ps(_mm_load_ps(vlo[a][k]), l));
                _mm_store_ps(vhi[a][k], _mm_max_ps(_mm_load_ps(vhi[a][k]), h));
            }
        }
        for (int a = 0; a < 3; ++a) for (int k = 0; k < nb; ++k) for (int d = 0; d < 3; ++d) { bins[a][k].lo[d] = vlo[a][k][d]; bins[a][k].hi[d] = vhi[a][k][d]; }
#else
        for (size_t i = b; i < e; ++i)                      // one pass over the clusters fills the bins of all three axes
            for (int a = 0; a < 3; ++a) {
                if (!use[a]) continue;
                Bin& B = bins[a][std::min(nb - 1, std::max(0, (int)((0.5f * (c[i].lo[a] + c[i].hi[a]) - clo[a]) * sc[a])))];
                B.n += c[i].cnt;
                for (int d = 0; d < 3; ++d) { B.lo[d] = std::min(B.lo[d], c[i].lo[d]); B.hi[d] = std::max(B.hi[d], c[i].hi[d]); }
            }
#endif
        for (int a = 0; a < 3; ++a) {
            if (!use[a]) continue;
            float rl[kBins][3], rh[kBins][3]; uint32_t rn[kBins];           // suffix boxes
            float l3[3] = {3e38f, 3e38f, 3e38f}, h3[3] = {-3e38f, -3e38f, -3e38f}; uint32_t nn = 0;
            for (int k = nb - 1; k >= 0; --k) {
                for (int d = 0; d < 3; ++d) { l3[d] = std::min(l3[d], bins[a][k].lo[d]); h3[d] = std::max(h3[d], bins[a][k].hi[d]); rl[k][d] = l3[d]; rh[k][d] = h3[d]; }
                nn += bins[a][k].n; rn[k] = nn;
            }
            for (int d = 0; d < 3; ++d) { l3[d] = 3e38f; h3[d] = -3e38f; }
            nn = 0;
            for (int k = 0; k + 1 < nb; ++k) {                              // split after bin k
                for (int d = 0; d < 3; ++d) { l3[d] = std::min(l3[d], bins[a][k].lo[d]); h3[d] = std::max(h3[d], bins[a][k].hi[d]); }
                nn += bins[a][k].n;
                if (nn == 0 || rn[k + 1] == 0) continue;
                const float cost = half_area(l3, h3) * (float)nn + half_area(rl[k + 1], rh[k + 1]) * (float)rn[k + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = k; }
            }
        }
    }
    if (best_axis >= 0) {
        const int a = best_axis;
        TopCluster* it = std::partition(c + b, c + e, [&](const TopCluster& x) {
            return std::min(nb - 1, std::max(0, (int)((0.5f * (x.lo[a] + x.hi[a]) - clo[a]) * sc[a]))) <= best_bin; });
        mid = (size_t)(it - c);
    }
    if (mid == b || mid == e) mid = b + (e - b) / 2;                        // coincident centroids: split the list in half
    const size_t ml = mid - b;
    int l, r;
    if (e - b > 16384 && depth < 6) {                                       // big subtrees: left half on a thread of its own
        std::thread t;
        bool spawned = true;
        try { t = std::thread([&]() { l = top_sah(c, b, mid, first_id, nodes, base + 1, depth + 1); }); }
        catch (const std::system_error&) { spawned = false; }              // no thread to be had: do the left half here as well
        r = top_sah(c, mid, e, first_id, nodes, base + ml, depth + 1);
        if (spawned) t.join(); else l = top_sah(c, b, mid, first_id, nodes, base + 1, depth + 1);
    } else {
        l = top_sah(c, b, mid, first_id, nodes, base + 1, depth + 1);
        r = top_sah(c, mid, e, first_id, nodes, base + ml, depth + 1);
    }
    TopNode& nd = nodes[base];
    nd.left = l; nd.right = r; nd.cnt = total;
    for (int d = 0; d < 3; ++d) { nd.lo[d] = lo[d]; nd.hi[d] = hi[d]; }
    return first_id + (int)base;
}

}  // namespace topsah
}  // namespace fovpt
