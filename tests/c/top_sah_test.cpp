// top_sah_test.cpp -- CPU unit test of the BVH builder's host stage (csrc/top_sah.h): the binned-SAH tree over the clusters
// that the device-side PLOC phase leaves.  Checks, for random, clustered, degenerate and tiny inputs: every cluster is a leaf
// exactly once, every node's box and triangle count are the union / sum of its children, ids are dense, the result does not
// depend on thread timing, well separated groups are split first, and coincident boxes give a balanced (shallow) tree.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#include "top_sah.h"

using namespace fovpt::topsah;

static int fails = 0;
#define CHECK(x) do { if (!(x)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #x); ++fails; } } while (0)

struct Walk { size_t leaves = 0; int depth = 0; };

// returns (lo, hi, cnt) of the subtree; leaf ids are -(1 + cluster number)
static void walk(int id, int first_id, const std::vector<TopNode>& nodes, const std::vector<TopCluster>& in, std::vector<int>& seen,
                 float lo[3], float hi[3], uint32_t& cnt, Walk& w, int depth)
{
    w.depth = std::max(w.depth, depth);
    if (id < first_id) {                                   // a cluster
        const int k = -id - 1;
        CHECK(k >= 0 && k < (int)in.size());
        ++seen[(size_t)k]; ++w.leaves;
        std::memcpy(lo, in[(size_t)k].lo, 12); std::memcpy(hi, in[(size_t)k].hi, 12); cnt = in[(size_t)k].cnt;
        return;
    }
    const size_t n = (size_t)(id - first_id);
    CHECK(n < nodes.size());
    float l1[3], h1[3], l2[3], h2[3]; uint32_t c1 = 0, c2 = 0;
    walk(nodes[n].left, first_id, nodes, in, seen, l1, h1, c1, w, depth + 1);
    walk(nodes[n].right, first_id, nodes, in, seen, l2, h2, c2, w, depth + 1);
    for (int d = 0; d < 3; ++d) {
        lo[d] = std::min(l1[d], l2[d]); hi[d] = std::max(h1[d], h2[d]);
        CHECK(nodes[n].lo[d] == lo[d] && nodes[n].hi[d] == hi[d]);
    }
    cnt = c1 + c2;
    CHECK(nodes[n].cnt == cnt);
}

static std::vector<TopNode> build(std::vector<TopCluster> c, int first_id, int* root, double* ms = nullptr)
{
    std::vector<TopNode> nodes(c.size() > 1 ? c.size() - 1 : 0);
    const auto t0 = std::chrono::steady_clock::now();
    *root = top_sah(c.data(), 0, c.size(), first_id, nodes.data(), 0, 0);
    if (ms) *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return nodes;
}

static Walk validate(const std::vector<TopCluster>& in, int first_id, const std::vector<TopNode>& nodes, int root)
{
    std::vector<int> seen(in.size(), 0);
    float lo[3], hi[3]; uint32_t cnt = 0; Walk w;
    walk(root, first_id, nodes, in, seen, lo, hi, cnt, w, 0);
    for (int s : seen) CHECK(s == 1);
    uint32_t total = 0; for (auto& k : in) total += k.cnt;
    CHECK(cnt == total && w.leaves == in.size());
    return w;
}

static std::vector<TopCluster> random_clusters(size_t n, unsigned seed, bool clumped)
{
    std::mt19937 g(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::vector<TopCluster> c(n);
    for (size_t i = 0; i < n; ++i) {
        float ctr[3];
        for (int d = 0; d < 3; ++d) ctr[d] = clumped ? (float)((i * 7 + (size_t)d) % 5) * 100.f + U(g) : U(g) * 1000.f;
        c[i].cid = -(int)i - 1; c[i].cnt = 1 + (uint32_t)(U(g) * 4.f);
        for (int d = 0; d < 3; ++d) { const float r = U(g) * 2.f; c[i].lo[d] = ctr[d] - r; c[i].hi[d] = ctr[d] + r; }
    }
    return c;
}

int main()
{
    const int first_id = 1000000;
    // tiny inputs
    for (size_t n : {(size_t)1, (size_t)2, (size_t)3, (size_t)7}) {
        auto c = random_clusters(n, 1, false); int root;
        auto nodes = build(c, first_id, &root);
        if (n == 1) CHECK(root == c[0].cid && nodes.empty());
        else validate(c, first_id, nodes, root);
    }
    // random and clumped, small enough to stay on one thread and large enough to spawn threads; twice -> identical
    for (size_t n : {(size_t)1000, (size_t)50000, (size_t)300000}) {
        for (int clumped = 0; clumped < 2; ++clumped) {
            auto c = random_clusters(n, 7 + (unsigned)n, clumped != 0); int r1, r2; double ms = 0;
            auto n1 = build(c, first_id, &r1, &ms), n2 = build(c, first_id, &r2);
            const Walk w = validate(c, first_id, n1, r1);
            CHECK(r1 == r2 && n1.size() == n2.size() && std::memcmp(n1.data(), n2.data(), n1.size() * sizeof(TopNode)) == 0);
            CHECK(w.depth <= 64);
            std::printf("n = %7zu %s: depth %2d, %.1f ms\n", n, clumped ? "clumped" : "uniform", w.depth, ms);
        }
    }
    // two well separated groups: the root separates them
    {
        auto c = random_clusters(4000, 3, false);
        for (size_t i = 0; i < c.size(); i += 2) { c[i].lo[1] += 1e5f; c[i].hi[1] += 1e5f; }
        int root; auto nodes = build(c, first_id, &root);
        validate(c, first_id, nodes, root);
        const TopNode& top = nodes[(size_t)(root - first_id)];
        auto side = [&](int id) { return id < first_id ? c[(size_t)(-id - 1)].lo[1] > 5e4f : nodes[(size_t)(id - first_id)].lo[1] > 5e4f; };
        CHECK(side(top.left) != side(top.right));
        auto span = [&](int id) { return id < first_id ? 0.f : nodes[(size_t)(id - first_id)].hi[1] - nodes[(size_t)(id - first_id)].lo[1]; };
        CHECK(span(top.left) < 2000.f && span(top.right) < 2000.f);
    }
    // coincident boxes (duplicated geometry): halved lists, depth ~ log2 n
    {
        auto c = random_clusters(100000, 5, false);
        for (auto& k : c) { for (int d = 0; d < 3; ++d) { k.lo[d] = 1.f; k.hi[d] = 2.f; } }
        int root; auto nodes = build(c, first_id, &root);
        const Walk w = validate(c, first_id, nodes, root);
        CHECK(w.depth <= 20);
    }
    std::printf(fails ? "top_sah: %d check(s) failed\n" : "top_sah ok\n", fails);
    return fails ? 1 : 0;
}
