// CPU unit test of the grid nearest-neighbour search logic (gennbv_b200/csrc/nn_grid.cuh): the same functions the CUDA
// kernels call are compiled for the host, the grid is built serially (count -> exclusive scan -> fill, as chamfer.cu does in
// parallel) and every query's result is compared, bit for bit, with a scan over all points.
// Built and run by tests/test_nn_grid_host.py:  g++ -O2 -ffp-contract=off nn_grid_host_test.cpp && ./a.out
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../gennbv_b200/csrc/nn_grid.cuh"

using namespace gnbv;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double urand() {            // xorshift64*, deterministic across platforms
    rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
    return (double)((rng_state * 2685821657736338717ull) >> 11) / 9007199254740992.0;
}

struct Cloud { std::vector<float> p; int n() const { return (int)p.size() / 3; } };

static Cloud box_surface(int n, float sx, float sy, float sz, float jitter) {
    Cloud c;
    for (int i = 0; i < n; ++i) {
        float u = (float)urand(), v = (float)urand();
        int face = (int)(urand() * 6) % 6;
        float x, y, z;
        switch (face) {
            case 0: x = 0; y = u * sy; z = v * sz; break;
            case 1: x = sx; y = u * sy; z = v * sz; break;
            case 2: y = 0; x = u * sx; z = v * sz; break;
            case 3: y = sy; x = u * sx; z = v * sz; break;
            case 4: z = 0; x = u * sx; y = v * sy; break;
            default: z = sz; x = u * sx; y = v * sy; break;
        }
        c.p.push_back(x - sx / 2 + jitter * (float)(urand() - 0.5));
        c.p.push_back(y - sy / 2 + jitter * (float)(urand() - 0.5));
        c.p.push_back(z + jitter * (float)(urand() - 0.5));
    }
    return c;
}

static Cloud volume(int n, float s, float ox) {
    Cloud c;
    for (int i = 0; i < 3 * n; ++i) c.p.push_back((float)(urand() - 0.5) * s + (i % 3 == 0 ? ox : 0.f));
    return c;
}

struct Grid { NNGridMeta m; std::vector<int> cell_end; std::vector<Float4> pts; };

static Grid build(const Cloud& r, int C) {
    Grid g;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < r.n(); ++i)
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], r.p[3 * i + a]); hi[a] = fmaxf(hi[a], r.p[3 * i + a]); }
    memset(&g.m, 0, sizeof(g.m));
    nn_make_meta(g.m, lo, hi, r.n(), C);
    int ncell = g.m.dims[0] * g.m.dims[1] * g.m.dims[2];
    std::vector<int> cnt(ncell > 0 ? ncell : 1, 0);
    std::vector<int> lin(r.n());
    for (int i = 0; i < r.n(); ++i) {
        int c[3];
        nn_cell_of(g.m, r.p[3 * i], r.p[3 * i + 1], r.p[3 * i + 2], c);
        for (int a = 0; a < 3; ++a)
            if (c[a] < 0 || c[a] >= g.m.dims[a]) { printf("FAIL cell out of range\n"); exit(1); }
        lin[i] = nn_cell_linear(g.m, c);
        cnt[lin[i]]++;
    }
    int run = 0;
    for (int c = 0; c < ncell; ++c) { int t = cnt[c]; cnt[c] = run; run += t; }      // exclusive scan
    g.pts.resize(r.n() > 0 ? r.n() : 1);
    for (int i = 0; i < r.n(); ++i) {                                               // fill: cnt[c] becomes the end offset
        Float4 v; v.x = r.p[3 * i]; v.y = r.p[3 * i + 1]; v.z = r.p[3 * i + 2]; v.w = 0.f;
        g.pts[cnt[lin[i]]++] = v;
    }
    g.cell_end = cnt;
    return g;
}

static long check(const char* name, const Cloud& q, const Cloud& r, int C) {
    Grid g = build(r, C);
    long bad = 0;
    for (int i = 0; i < q.n(); ++i) {
        const float qx = q.p[3 * i], qy = q.p[3 * i + 1], qz = q.p[3 * i + 2];
        float want = FLT_MAX;
        for (int j = 0; j < r.n(); ++j) {
            Float4 v; v.x = r.p[3 * j]; v.y = r.p[3 * j + 1]; v.z = r.p[3 * j + 2]; v.w = 0.f;
            want = fminf(want, nn_sqdist(qx, qy, qz, v));
        }
        const float got = nn_query(g.m, g.cell_end.data(), g.pts.data(), qx, qy, qz);
        if (memcmp(&got, &want, 4) != 0) {
            if (bad < 5) printf("  %s: query %d got %.9g want %.9g\n", name, i, got, want);
            ++bad;
        }
    }
    printf("%-44s C=%3d  nq=%6d nr=%6d dims=%dx%dx%d  mismatches=%ld\n", name, C, q.n(), r.n(), g.m.dims[0], g.m.dims[1], g.m.dims[2], bad);
    return bad;
}

int main() {
    long bad = 0;
    Cloud gt = box_surface(7000, 5.3f, 4.1f, 3.7f, 0.0f), scan = box_surface(10000, 5.3f, 4.1f, 3.7f, 0.02f);
    for (int i = 0; i < scan.n() * 3; ++i) scan.p[i] = nearbyintf(scan.p[i] * 100.f) / 100.f;        // 1 cm lattice, as the eval env
    bad += check("scan(1cm) -> gt surface", scan, gt, 71);
    bad += check("gt surface -> scan(1cm)", gt, scan, 122);
    bad += check("surface, coarse grid", scan, gt, 3);
    bad += check("surface, one cell", scan, gt, 1);
    bad += check("surface, finest grid (escape path)", gt, box_surface(300, 5.3f, 4.1f, 3.7f, 0.f), 160);
    Cloud va = volume(4000, 4.f, 0.f), vb = volume(6000, 4.f, 0.1f);
    bad += check("volume -> volume", va, vb, 18);
    bad += check("volume -> volume, sparse cells", va, vb, 64);
    bad += check("queries far outside the box", volume(2000, 400.f, 30.f), vb, 32);
    bad += check("identical clouds", vb, vb, 24);
    Cloud one; one.p = {0.5f, -1.f, 2.f};
    bad += check("single reference point", va, one, 16);
    Cloud same; for (int i = 0; i < 50; ++i) { same.p.push_back(1.f); same.p.push_back(2.f); same.p.push_back(3.f); }
    bad += check("coincident reference points", va, same, 16);
    Cloud plane; for (int i = 0; i < 5000; ++i) { plane.p.push_back((float)urand() * 3); plane.p.push_back((float)urand() * 2); plane.p.push_back(0.25f); }
    bad += check("planar reference cloud", va, plane, 40);
    Cloud line; for (int i = 0; i < 1000; ++i) { line.p.push_back(0.f); line.p.push_back(0.f); line.p.push_back((float)urand()); }
    bad += check("collinear reference cloud", va, line, 160);
    Cloud empty;
    bad += check("empty reference cloud", va, empty, 8);
    bad += check("empty query cloud", empty, vb, 8);
    printf(bad ? "FAILED: %ld mismatches\n" : "OK (%ld mismatches)\n", bad);
    return bad ? 1 : 0;
}
