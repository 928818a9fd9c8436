// trav_sim.cpp — CPU model of the traversal kernels' SIMT execution (design tool, not product code).
//
// Loads a scene with the product's own host code (scene_host / scene_xml), builds the pair layout and the wide layouts,
// generates primary + diffuse bounce rays in the order the wavefront queues hold them, traverses every ray on the CPU
// with each candidate scheme while recording its step trace, checks that all schemes find the same closest hit, and then
// replays the traces through a lock-step warp model that counts issued warp instructions and active lanes per
// instruction (= ncu's smsp__thread_inst_executed_per_inst_executed).  Used in round 2 to choose between per-thread,
// dynamic-fetch and warp-cooperative traversal before spending GPU minutes.
//
//   g++ -O2 -std=c++17 -fopenmp -Iinclude -Icuda-pt_b200/csrc tools/sim/trav_sim.cpp cuda-pt_b200/csrc/scene_host.cpp \
//       cuda-pt_b200/csrc/scene_xml.cpp cuda-pt_b200/csrc/image_io.cpp cuda-pt_b200/csrc/jpeg_read.cpp \
//       cuda-pt_b200/csrc/hdr_read.cpp cuda-pt_b200/csrc/nvdb_read.cpp -lz -o /tmp/trav_sim
#include "scene_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

using namespace cptb;

struct V3 {
    float x, y, z;
};
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline V3 norm(V3 a) {
    const float l = 1.f / std::sqrt(dot(a, a));
    return a * l;
}
static inline int as_int(float f) {
    int32_t v;
    memcpy(&v, &f, 4);
    return v;
}

struct Ray {
    V3 o, d;
    float tmax;
};
struct Hit {
    float t = 1e7f, u = 0, v = 0;
    int prim = -1;
};

static float half_to_float(uint16_t h) {
    const uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023;
    uint32_t bits;
    if (e == 0) {
        if (m == 0)
            bits = s << 31;
        else {
            int ee = -1;
            uint32_t mm = m;
            do {
                ee++;
                mm <<= 1;
            } while (!(mm & 1024));
            bits = (s << 31) | uint32_t(127 - 15 - ee) << 23 | (mm & 1023) << 13;
        }
    } else if (e == 31)
        bits = (s << 31) | 0x7f800000u | (m << 13);
    else
        bits = (s << 31) | (e + 112) << 23 | (m << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

struct Scene {
    HostScene hs;
    std::vector<F4> tris;
    PairLayout pl;
    WideLayout w4, w8, w2;
    std::vector<int> prim_orig;
};

static bool slab(const F4 &lo, const F4 &hi, V3 inv_d, V3 o_div, float &t_near) {
    const float ax = std::fma(lo.x, inv_d.x, -o_div.x), ay = std::fma(lo.y, inv_d.y, -o_div.y), az = std::fma(lo.z, inv_d.z, -o_div.z);
    const float bx = std::fma(hi.x, inv_d.x, -o_div.x), by = std::fma(hi.y, inv_d.y, -o_div.y), bz = std::fma(hi.z, inv_d.z, -o_div.z);
    const float min_x = ax < bx ? ax : bx, max_x = ax < bx ? bx : ax;
    const float min_y = ay < by ? ay : by, max_y = ay < by ? by : ay;
    const float min_z = az < bz ? az : bz, max_z = az < bz ? bz : az;
    t_near = std::fmax(min_x, std::fmax(min_y, min_z));
    const float t_far = std::fmin(max_x, std::fmin(max_y, max_z));
    return (t_far > t_near) && (t_far > 0);
}

static float hit_tri(const Scene &sc, int idx, V3 o, V3 d, float &su, float &sv, float max_range) {
    const F4 p0 = sc.tris[3 * size_t(idx)], v1 = sc.tris[3 * size_t(idx) + 1], v2 = sc.tris[3 * size_t(idx) + 2];
    const V3 v = o - V3{p0.x, p0.y, p0.z};
    const V3 t1{std::fma(v2.y, -d.z, d.y * v2.z), std::fma(-d.x, v2.z, v2.x * d.z), std::fma(v2.x, -d.y, d.x * v2.y)};
    const V3 t2{std::fma(-d.y, v1.z, v1.y * d.z), std::fma(v1.x, -d.z, d.x * v1.z), std::fma(-d.x, v1.y, v1.x * d.y)};
    const float det = std::fma(-p0.w, d.x, std::fma(v1.x, t1.x, v2.x * t2.x));
    const float inv_det = 1.f / det;
    const float a = dot(t1, v) * inv_det, b = dot(t2, v) * inv_det;
    const float t = (std::fma(p0.w, v.x, std::fma(v1.w, v.y, v2.w * v.z))) * inv_det;
    su = a, sv = b;
    return (a > 0 && b > 0 && a + b < 1 && t > 1e-3f && t < max_range) ? t : 0.f;
}

static int cached_entry(const Scene &sc, V3 inv_d, V3 o_div, float limit, int miss_value, int *iters = nullptr) {
    int node_idx = miss_value;
    const int n = int(sc.hs.cache_nodes.size());
    int it = 0;
    for (int i = 0; i < n;) {
        ++it;
        const U4 c = sc.hs.cache_nodes[size_t(i)];
        const F4 lo{half_to_float(uint16_t(c.x & 0xffff)), half_to_float(uint16_t(c.y & 0xffff)), half_to_float(uint16_t(c.z & 0xffff)), 0};
        const F4 hi{half_to_float(uint16_t(c.x >> 16)), half_to_float(uint16_t(c.y >> 16)), half_to_float(uint16_t(c.z >> 16)), 0};
        float tn;
        bool hit = slab(lo, hi, inv_d, o_div, tn) && tn < limit;
        const int all_offset = int(c.w & 0xff), gmem = int(c.w >> 8);
        const int inc = (!hit) * all_offset + int(hit && all_offset != 1);
        hit = hit && all_offset == 1;
        i = hit ? n : i + inc;
        node_idx = hit ? gmem : node_idx;
    }
    if (iters)
        *iters = it;
    return node_idx;
}

static bool prim_before(const Scene &sc, int a, int b) {
    if (b < 0)
        return false;
    return sc.prim_orig.empty() ? a < b : sc.prim_orig[size_t(a)] < sc.prim_orig[size_t(b)];
}

// ------------------------------------------------------------------ traces
// pair / wide per-thread traversal: one record per loop iteration
struct Step {
    uint8_t kind;  // 0 inner, 1 leaf
    uint8_t cnt;   // leaf: triangles tested; inner(wide per-thread): children tested
    uint8_t pops;  // stack entries examined after this step (0 = descended without popping)
    int32_t ref = 0; // pair traversal: node index (inner) / first primitive (leaf) - for the L1 line model
};
// cooperative traversal: one record per node visit of the lane group
struct CStep {
    uint8_t leaf_lanes; // lanes whose child is a hit leaf (tested in place)
    uint8_t leaf_max;   // largest primitive count among them
    uint16_t leaf_tris; // total triangles tested in this step
    uint8_t pushes;     // entries pushed
    uint8_t pops;       // entries examined after the step
};
struct Trace {
    std::vector<Step> pair, w4t, w8t;
    std::vector<CStep> c4, c8;
    int cache_iters = 0;
};

static Hit trace_pair(const Scene &sc, const Ray &r, bool anyhit, Trace *tr, std::vector<Step> *out) {
    Hit h;
    h.t = r.tmax;
    const V3 inv_d{1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z}, o_div{r.o.x * inv_d.x, r.o.y * inv_d.y, r.o.z * inv_d.z};
    int ci = 0;
    const int start = cached_entry(sc, inv_d, o_div, r.tmax, anyhit ? sc.hs.num_nodes() : 0, &ci);
    if (tr)
        tr->cache_iters = ci;
    if (anyhit && start >= sc.hs.num_nodes())
        return h;
    int stack_ref[128];
    float stack_t[128];
    int sp = 0, cur = sc.pl.root;
    while (true) {
        Step st{0, 0, 0};
        bool descend = false;
        if (cur >= 0) {
            const F4 a0 = sc.pl.wnodes[4 * size_t(cur)], a1 = sc.pl.wnodes[4 * size_t(cur) + 1], b0 = sc.pl.wnodes[4 * size_t(cur) + 2],
                     b1 = sc.pl.wnodes[4 * size_t(cur) + 3];
            float t0, t1;
            const bool h0 = slab(a0, a1, inv_d, o_div, t0) && t0 < h.t && as_int(a1.w) > start;
            const bool h1 = slab(b0, b1, inv_d, o_div, t1) && t1 < h.t && as_int(b1.w) > start;
            const int r0 = as_int(a0.w), r1 = as_int(b0.w);
            st.kind = 0, st.cnt = 2, st.ref = cur;
            if (h0 && h1) {
                const bool swap = !anyhit && t1 < t0;
                stack_ref[sp] = swap ? r0 : r1;
                stack_t[sp] = swap ? t0 : t1;
                ++sp;
                cur = swap ? r1 : r0;
                descend = true;
            } else if (h0 || h1) {
                cur = h0 ? r0 : r1;
                descend = true;
            }
        } else {
            const int packed = ~cur, base = packed & ((1 << 26) - 1), cnt = packed >> 26;
            st.kind = 1, st.cnt = uint8_t(cnt), st.ref = base;
            for (int idx = base; idx < base + cnt; idx++) {
                float u, v;
                const float dist = hit_tri(sc, idx, r.o, r.d, u, v, anyhit ? r.tmax : INFINITY);
                if (anyhit) {
                    if (dist > 1e-3f) {
                        h.prim = idx, h.t = dist;
                        if (out)
                            out->push_back(st);
                        return h;
                    }
                    continue;
                }
                const bool valid = dist > 1e-3f && (dist < h.t || (dist == h.t && prim_before(sc, idx, h.prim)));
                if (valid)
                    h.t = dist, h.u = u, h.v = v, h.prim = idx;
            }
        }
        bool found = descend;
        if (!descend) {
            while (sp > 0) {
                --sp;
                st.pops++;
                if (anyhit || stack_t[sp] < h.t) {
                    cur = stack_ref[sp];
                    found = true;
                    break;
                }
            }
        }
        if (out)
            out->push_back(st);
        if (!found)
            break;
    }
    return h;
}

// per-thread traversal of a wide layout (one lane tests all children; sorted push)
static Hit trace_wide_thread(const Scene &sc, const WideLayout &w, const Ray &r, std::vector<Step> *out) {
    Hit h;
    h.t = r.tmax;
    const V3 inv_d{1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z}, o_div{r.o.x * inv_d.x, r.o.y * inv_d.y, r.o.z * inv_d.z};
    const int start = cached_entry(sc, inv_d, o_div, r.tmax, 0);
    int stack_ref[256];
    float stack_t[256];
    int sp = 0, cur = 0;
    const int W = w.width;
    while (true) {
        Step st{0, 0, 0};
        bool descend = false;
        if (cur >= 0) {
            st.kind = 0, st.cnt = uint8_t(W);
            int refs[8];
            float ts[8];
            int n = 0;
            for (int i = 0; i < W; i++) {
                const F4 lo = w.nodes[(size_t(cur) * W + i) * 2], hi = w.nodes[(size_t(cur) * W + i) * 2 + 1];
                float t;
                if (slab(lo, hi, inv_d, o_div, t) && t < h.t && as_int(hi.w) > start)
                    refs[n] = as_int(lo.w), ts[n] = t, n++;
            }
            // sort descending by t so that the nearest ends on top
            for (int i = 1; i < n; i++)
                for (int j = i; j > 0 && ts[j] > ts[j - 1]; j--)
                    std::swap(ts[j], ts[j - 1]), std::swap(refs[j], refs[j - 1]);
            if (n > 0) {
                for (int i = 0; i < n - 1; i++)
                    stack_ref[sp] = refs[i], stack_t[sp] = ts[i], sp++;
                cur = refs[n - 1];
                descend = true;
            }
        } else {
            const int packed = ~cur, base = packed & ((1 << 26) - 1), cnt = packed >> 26;
            st.kind = 1, st.cnt = uint8_t(cnt);
            for (int idx = base; idx < base + cnt; idx++) {
                float u, v;
                const float dist = hit_tri(sc, idx, r.o, r.d, u, v, INFINITY);
                const bool valid = dist > 1e-3f && (dist < h.t || (dist == h.t && prim_before(sc, idx, h.prim)));
                if (valid)
                    h.t = dist, h.u = u, h.v = v, h.prim = idx;
            }
        }
        bool found = descend;
        if (!descend) {
            while (sp > 0) {
                --sp;
                st.pops++;
                if (stack_t[sp] < h.t) {
                    cur = stack_ref[sp];
                    found = true;
                    break;
                }
            }
        }
        if (out)
            out->push_back(st);
        if (!found)
            break;
    }
    return h;
}

// cooperative traversal: W lanes per ray, lane i owns child i; hit leaves are tested in place; the nearest hit inner child
// is entered, the others are pushed (unsorted) and popped nearest-of-the-top-W first
static Hit trace_coop(const Scene &sc, const WideLayout &w, const Ray &r, bool anyhit, bool pop_nearest, std::vector<CStep> *out,
                      int *max_sp = nullptr) {
    Hit h;
    h.t = r.tmax;
    const V3 inv_d{1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z}, o_div{r.o.x * inv_d.x, r.o.y * inv_d.y, r.o.z * inv_d.z};
    const int start = cached_entry(sc, inv_d, o_div, r.tmax, anyhit ? sc.hs.num_nodes() : 0);
    if (anyhit && start >= sc.hs.num_nodes())
        return h;
    std::vector<int> stack_ref(512);
    std::vector<float> stack_t(512);
    int sp = 0, cur = 0;
    const int W = w.width;
    while (true) {
        CStep st{0, 0, 0, 0, 0};
        int refs[8];
        float ts[8];
        int n = 0;
        for (int i = 0; i < W; i++) {
            const F4 lo = w.nodes[(size_t(cur) * W + i) * 2], hi = w.nodes[(size_t(cur) * W + i) * 2 + 1];
            float t;
            if (!(slab(lo, hi, inv_d, o_div, t) && t < h.t && as_int(hi.w) > start))
                continue;
            const int ref = as_int(lo.w);
            if (ref < 0) {
                const int packed = ~ref, base = packed & ((1 << 26) - 1), cnt = packed >> 26;
                st.leaf_lanes++;
                st.leaf_max = std::max<uint8_t>(st.leaf_max, uint8_t(cnt));
                st.leaf_tris += cnt;
                for (int idx = base; idx < base + cnt; idx++) {
                    float u, v;
                    const float dist = hit_tri(sc, idx, r.o, r.d, u, v, anyhit ? r.tmax : INFINITY);
                    if (anyhit) {
                        if (dist > 1e-3f) {
                            h.prim = idx, h.t = dist;
                            if (out)
                                out->push_back(st);
                            return h;
                        }
                        continue;
                    }
                    const bool valid = dist > 1e-3f && (dist < h.t || (dist == h.t && prim_before(sc, idx, h.prim)));
                    if (valid)
                        h.t = dist, h.u = u, h.v = v, h.prim = idx;
                }
            } else {
                refs[n] = ref, ts[n] = t, n++;
            }
        }
        // inner children still in front of the (possibly updated) closest hit
        int m = 0;
        for (int i = 0; i < n; i++)
            if (anyhit || ts[i] < h.t)
                refs[m] = refs[i], ts[m] = ts[i], m++;
        bool found = false;
        if (m > 0) {
            int best = 0;
            for (int i = 1; i < m; i++)
                if (!anyhit && ts[i] < ts[best])
                    best = i;
            cur = refs[best];
            for (int i = 0; i < m; i++)
                if (i != best)
                    stack_ref[sp] = refs[i], stack_t[sp] = ts[i], sp++;
            st.pushes = uint8_t(m - 1);
            found = true;
        } else {
            while (sp > 0) {
                st.pops++;
                if (pop_nearest && !anyhit) {
                    const int k = std::min(sp, W);
                    int best = -1;
                    for (int i = 0; i < k; i++)
                        if (stack_t[sp - 1 - i] < h.t && (best < 0 || stack_t[sp - 1 - i] < stack_t[sp - 1 - best]))
                            best = i;
                    if (best < 0) {
                        sp -= k;
                        continue;
                    }
                    cur = stack_ref[sp - 1 - best];
                    stack_ref[sp - 1 - best] = stack_ref[sp - 1];
                    stack_t[sp - 1 - best] = stack_t[sp - 1];
                    --sp;
                    found = true;
                    break;
                }
                --sp;
                if (anyhit || stack_t[sp] < h.t) {
                    cur = stack_ref[sp];
                    found = true;
                    break;
                }
            }
        }
        if (max_sp)
            *max_sp = std::max(*max_sp, sp);
        if (out)
            out->push_back(st);
        if (!found)
            break;
    }
    return h;
}

// ------------------------------------------------------------------ warp models
struct Cost { // warp instructions
    double inner = 59, leaf0 = 14, tri = 66, pop = 9, cache = 24, fetch = 30, epi = 0;
    double w_child = 24, w_over = 30;                    // per-thread wide: per child tested + per-node overhead (sort, push)
    double c_slab = 27, c_tri0 = 8, c_tri = 62, c_sel = 16, c_push = 6, c_pop = 10, c_fetch = 34; // cooperative
};
struct Tally {
    double inst = 0, thread_inst = 0;
    void add(double n, double lanes) { inst += n, thread_inst += n * lanes; }
    double lanes() const { return inst > 0 ? thread_inst / inst : 0; }
};

// per-thread traversal, static chunks of 32 (what k_trace<2> does), optionally with refill from a buffer of `buf` rays
static Tally model_thread(const std::vector<const std::vector<Step> *> &rays, const std::vector<int> &cache_iters, const Cost &c, int buf, bool wide,
                          int refill_threshold = 1) {
    Tally T;
    const size_t N = rays.size();
    for (size_t base = 0; base < N; base += size_t(buf)) {
        const size_t end = std::min(N, base + size_t(buf));
        size_t next = base;
        const std::vector<Step> *cur[32];
        size_t pos[32];
        int nact = 0;
        // prologue: the first 32 rays (cached phase per lane)
        auto fetch = [&](int lane) {
            if (next < end) {
                cur[lane] = rays[next], pos[lane] = 0;
                ++next;
                return true;
            }
            cur[lane] = nullptr;
            return false;
        };
        int maxc = 0, got = 0;
        for (int l = 0; l < 32; l++) {
            const size_t idx = next;
            if (fetch(l))
                maxc = std::max(maxc, cache_iters[idx]), got++;
        }
        nact = got;
        T.add(c.fetch, got);
        T.add(c.cache * maxc, got * 0.6); // cached-phase loop, divergent trip counts (rough)
        while (nact > 0) {
            int n_in = 0, n_leaf = 0, max_cnt = 0, sum_cnt = 0, max_pop = 0, sum_pop = 0, n_pop = 0;
            for (int l = 0; l < 32; l++) {
                if (!cur[l])
                    continue;
                const Step &s = (*cur[l])[pos[l]];
                if (s.kind == 0)
                    n_in++;
                else
                    n_leaf++, max_cnt = std::max<int>(max_cnt, s.cnt), sum_cnt += s.cnt;
                if (s.pops)
                    n_pop++, max_pop = std::max<int>(max_pop, s.pops), sum_pop += s.pops;
            }
            if (n_in)
                T.add(wide ? c.w_over : c.inner, n_in);
            if (wide && n_in) {
                // children tests of the wide node
                int W = 0;
                for (int l = 0; l < 32; l++)
                    if (cur[l] && (*cur[l])[pos[l]].kind == 0)
                        W = (*cur[l])[pos[l]].cnt;
                T.add(c.w_child * W, n_in);
            }
            if (n_leaf) {
                T.add(c.leaf0, n_leaf);
                T.inst += c.tri * max_cnt, T.thread_inst += c.tri * sum_cnt;
            }
            if (n_pop)
                T.inst += c.pop * max_pop, T.thread_inst += c.pop * sum_pop;
            // advance
            int idle = 0;
            for (int l = 0; l < 32; l++) {
                if (!cur[l]) {
                    idle++;
                    continue;
                }
                if (++pos[l] >= cur[l]->size())
                    cur[l] = nullptr, nact--, idle++;
            }
            if (buf > 32 && idle >= refill_threshold && next < end) {
                int gotn = 0, mc = 0;
                for (int l = 0; l < 32; l++)
                    if (!cur[l]) {
                        const size_t idx = next;
                        if (fetch(l))
                            gotn++, nact++, mc = std::max(mc, cache_iters[idx]);
                    }
                T.add(c.fetch, gotn);
                T.add(c.cache * mc, gotn * 0.6);
            }
        }
    }
    return T;
}

// ------------------------------------------------------------------ L1 line model of the static per-thread kernel (round 2)
// The per-thread traversal kernels are bound by the L1/TEX pipe: a warp-level load costs one wavefront per distinct 128-byte line its
// active lanes touch.  This replays the pair traces of 32 rays in lock step (the static kernel) and counts, per ray, the wavefronts of
// the node loads (2 x LDG.256 per visit: one 64-byte node = half a line) and of the triangle loads (3 x LDG.128 per triangle) - as a
// function of WHICH rays share a warp.  `order` = the rays in the order they are handed to warps, 32 at a time.
struct LineStats {
    double node_wavefronts = 0, tri_wavefronts = 0, node_lane_loads = 0, tri_lane_loads = 0, iterations = 0, lane_iterations = 0;
    size_t rays = 0;
};
static LineStats model_lines(const std::vector<const std::vector<Step> *> &rays, const std::vector<uint32_t> &order) {
    LineStats L;
    const size_t N = order.size();
    for (size_t base = 0; base < N; base += 32) {
        const size_t n = std::min<size_t>(32, N - base);
        const std::vector<Step> *cur[32];
        size_t pos[32];
        int nact = 0;
        for (size_t l = 0; l < 32; l++) {
            cur[l] = l < n ? rays[order[base + l]] : nullptr, pos[l] = 0;
            if (cur[l] && cur[l]->empty())
                cur[l] = nullptr;
            nact += cur[l] != nullptr;
        }
        L.rays += n;
        while (nact > 0) {
            int lines[32], nl = 0, n_in = 0, max_cnt = 0;
            for (int l = 0; l < 32; l++) {
                if (!cur[l])
                    continue;
                const Step &s = (*cur[l])[pos[l]];
                if (s.kind == 0) {
                    n_in++;
                    const int line = s.ref >> 1; // two 64-byte pair nodes per 128-byte line
                    bool seen = false;
                    for (int k = 0; k < nl; k++)
                        seen |= lines[k] == line;
                    if (!seen)
                        lines[nl++] = line;
                } else
                    max_cnt = std::max<int>(max_cnt, s.cnt);
            }
            L.node_wavefronts += 2.0 * nl, L.node_lane_loads += 2.0 * n_in;
            for (int j = 0; j < max_cnt; j++) // the triangle loop: iteration j, loads k = 0..2 of 16 bytes at (first + j) * 48 + 16 k
                for (int k = 0; k < 3; k++) {
                    nl = 0;
                    int lanes = 0;
                    for (int l = 0; l < 32; l++) {
                        if (!cur[l])
                            continue;
                        const Step &s = (*cur[l])[pos[l]];
                        if (s.kind != 1 || s.cnt <= j)
                            continue;
                        lanes++;
                        const int line = int((size_t(s.ref + j) * 48 + size_t(k) * 16) >> 7);
                        bool seen = false;
                        for (int q = 0; q < nl; q++)
                            seen |= lines[q] == line;
                        if (!seen)
                            lines[nl++] = line;
                    }
                    L.tri_wavefronts += nl, L.tri_lane_loads += lanes;
                }
            L.iterations += 1, L.lane_iterations += nact;
            for (int l = 0; l < 32; l++)
                if (cur[l] && ++pos[l] >= cur[l]->size())
                    cur[l] = nullptr, nact--;
        }
    }
    return L;
}
// sort keys for "which rays share a warp": direction octant, then a 30-bit Morton code of the origin inside the scene's bounds
static uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    auto spread = [](uint32_t v) {
        v &= 1023u;
        v = (v | (v << 16)) & 0x030000FFu;
        v = (v | (v << 8)) & 0x0300F00Fu;
        v = (v | (v << 4)) & 0x030C30C3u;
        v = (v | (v << 2)) & 0x09249249u;
        return v;
    };
    return spread(x) | (spread(y) << 1) | (spread(z) << 2);
}

// cooperative: G lanes per ray, 32/G groups per warp, flat loop, chunk of `buf` rays per warp
static Tally model_coop(const std::vector<const std::vector<CStep> *> &rays, const std::vector<int> &cache_iters, const Cost &c, int G, int buf) {
    Tally T;
    const size_t N = rays.size();
    const int NG = 32 / G;
    for (size_t base = 0; base < N; base += size_t(buf)) {
        const size_t end = std::min(N, base + size_t(buf));
        // prologue: cached phase, one ray per lane, buf/32 rounds
        for (size_t b = base; b < end; b += 32) {
            int mc = 0;
            const int cntl = int(std::min<size_t>(32, end - b));
            for (int l = 0; l < cntl; l++)
                mc = std::max(mc, cache_iters[b + size_t(l)]);
            T.add(c.fetch, cntl);
            T.add(c.cache * mc, cntl * 0.6);
        }
        size_t next = base;
        const std::vector<CStep> *cur[8];
        size_t pos[8];
        int nact = 0;
        for (int g = 0; g < NG; g++) {
            cur[g] = next < end ? rays[next++] : nullptr;
            pos[g] = 0;
            nact += cur[g] != nullptr;
        }
        T.add(c.c_fetch, nact * G);
        while (nact > 0) {
            int leaf_lanes = 0, leaf_max = 0, leaf_tris = 0, n_push = 0, max_pop = 0, sum_pop = 0;
            for (int g = 0; g < NG; g++) {
                if (!cur[g])
                    continue;
                const CStep &s = (*cur[g])[pos[g]];
                leaf_lanes += s.leaf_lanes, leaf_max = std::max<int>(leaf_max, s.leaf_max), leaf_tris += s.leaf_tris;
                n_push += s.pushes > 0;
                if (s.pops)
                    max_pop = std::max<int>(max_pop, s.pops), sum_pop += s.pops;
            }
            T.add(c.c_slab, nact * G);
            if (leaf_lanes) {
                T.add(c.c_tri0, leaf_lanes);
                T.inst += c.c_tri * leaf_max, T.thread_inst += c.c_tri * leaf_tris;
            }
            T.add(c.c_sel, nact * G);
            if (n_push)
                T.add(c.c_push, n_push * G * 0.5);
            if (max_pop)
                T.inst += c.c_pop * max_pop, T.thread_inst += c.c_pop * sum_pop * G;
            int fetched = 0;
            for (int g = 0; g < NG; g++) {
                if (!cur[g])
                    continue;
                if (++pos[g] >= cur[g]->size()) {
                    cur[g] = next < end ? rays[next++] : nullptr;
                    pos[g] = 0;
                    if (cur[g])
                        fetched++;
                    else
                        nact--;
                }
            }
            if (fetched)
                T.add(c.c_fetch, fetched * G);
        }
    }
    return T;
}


// ------------------------------------------------------------------ lock-step model of the planned kernel
// Warp-cooperative traversal with DEFERRED leaf tests: G lanes per ray; a node visit tests G children; hit leaves are not
// tested in place but appended to the group's pending list (<= 2G leaf references); the warp runs a "triangle phase" -
// lane i of a group tests pending leaf i - only when enough lanes would be busy, or when a group cannot go on without
// it (list above G, or traversal finished with leaves pending).  Deferral only delays updates of the closest distance,
// i.e. it weakens culling (more visits), never the result.
struct DefStats {
    double inst = 0, thread_inst = 0, visits = 0, tri_tests = 0, tri_phases = 0, rays = 0, mismatches = 0, l1_wavefronts = 0;
};
static DefStats sim_coop_deferred(const Scene &sc, const WideLayout &w, const std::vector<Ray> &rays, const std::vector<Hit> &ref, const Cost &c,
                                  int chunk, int flush_lanes, int slab_nodes /* nodes served from shared memory */, int min_blocked = 1,
                                  double visit_cost = 0, double tri_cost = 0) {
    DefStats S;
    const int G = w.width, NG = 32 / G;
    struct Group {
        long ray = -1;
        int cur = 0, sp = 0, start = 0, npend = 0;
        bool walking = false; // has a node to visit
        Hit h;
        V3 inv_d, o_div;
        int stack_ref[256];
        float stack_t[256];
        int pend[16];
    };
    std::vector<Group> g;
    g.resize(size_t(NG));
    const size_t N = rays.size();
    for (size_t base = 0; base < N; base += size_t(chunk)) {
        const size_t end = std::min(N, base + size_t(chunk));
        size_t next = base;
        // prologue (full lanes): ray load + cached phase, one ray per lane
        for (size_t b = base; b < end; b += 32) {
            const int cntl = int(std::min<size_t>(32, end - b));
            S.inst += c.fetch + c.cache * 6, S.thread_inst += (c.fetch + c.cache * 6 * 0.7) * cntl;
        }
        auto fetch = [&](Group &q) {
            if (next >= end) {
                q.ray = -1, q.walking = false;
                return false;
            }
            q.ray = long(next++);
            const Ray &r = rays[size_t(q.ray)];
            q.inv_d = V3{1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z}, q.o_div = V3{r.o.x * q.inv_d.x, r.o.y * q.inv_d.y, r.o.z * q.inv_d.z};
            q.start = cached_entry(sc, q.inv_d, q.o_div, r.tmax, 0);
            q.h = Hit();
            q.h.t = r.tmax;
            q.cur = 0, q.sp = 0, q.npend = 0, q.walking = true;
            return true;
        };
        int live = 0;
        for (auto &q : g)
            live += fetch(q);
        S.inst += c.c_fetch, S.thread_inst += c.c_fetch * live * G;
        while (live > 0) {
            // ---- node visit for every walking group
            int walkers = 0;
            for (auto &q : g) {
                if (q.ray < 0 || !q.walking)
                    continue;
                walkers++;
                S.visits++;
                S.l1_wavefronts += q.cur < slab_nodes ? 0.25 * G * 32 / 128.0 : G * 32 / 128.0;
                const Ray &r = rays[size_t(q.ray)];
                int refs[8];
                float ts[8];
                int n = 0;
                for (int i = 0; i < G; i++) {
                    const F4 lo = w.nodes[(size_t(q.cur) * G + i) * 2], hi = w.nodes[(size_t(q.cur) * G + i) * 2 + 1];
                    float t;
                    if (!(slab(lo, hi, q.inv_d, q.o_div, t) && t < q.h.t && as_int(hi.w) > q.start))
                        continue;
                    const int rf = as_int(lo.w);
                    if (rf < 0)
                        q.pend[q.npend++] = rf;
                    else
                        refs[n] = rf, ts[n] = t, n++;
                }
                (void)r;
                bool found = false;
                if (n > 0) {
                    int best = 0;
                    for (int i = 1; i < n; i++)
                        if (ts[i] < ts[best])
                            best = i;
                    q.cur = refs[best];
                    for (int i = 0; i < n; i++)
                        if (i != best)
                            q.stack_ref[q.sp] = refs[i], q.stack_t[q.sp] = ts[i], q.sp++;
                    found = true;
                } else {
                    while (q.sp > 0) { // nearest of the top G
                        const int k = std::min(q.sp, G);
                        int best = -1;
                        for (int i = 0; i < k; i++)
                            if (q.stack_t[q.sp - 1 - i] < q.h.t && (best < 0 || q.stack_t[q.sp - 1 - i] < q.stack_t[q.sp - 1 - best]))
                                best = i;
                        S.inst += 0; // accounted below as one pop block per iteration
                        if (best < 0) {
                            q.sp -= k;
                            continue;
                        }
                        q.cur = q.stack_ref[q.sp - 1 - best];
                        q.stack_ref[q.sp - 1 - best] = q.stack_ref[q.sp - 1], q.stack_t[q.sp - 1 - best] = q.stack_t[q.sp - 1];
                        --q.sp;
                        found = true;
                        break;
                    }
                }
                q.walking = found;
                if (q.npend > G)
                    q.walking = false, q.cur = found ? q.cur : -1; // must flush first (keeps cur for afterwards)
                if (!found)
                    q.cur = -1;
            }
            if (walkers || visit_cost > 0) {
                const double vc = visit_cost > 0 ? visit_cost : c.c_slab + c.c_sel + c.c_push + c.c_pop * 0.6;
                S.inst += vc;
                S.thread_inst += vc * (visit_cost > 0 ? 32 : walkers * G);
            }
            // ---- triangle phase?
            int busy = 0, blocked = 0;
            for (auto &q : g) {
                if (q.ray < 0)
                    continue;
                busy += std::min(q.npend, G);
                blocked += (!q.walking && q.npend > 0);
            }
            if (busy >= flush_lanes || blocked >= min_blocked || (blocked > 0 && walkers - blocked <= 0)) {
                int maxc = 0, act = 0;
                double tsum = 0;
                for (auto &q : g) {
                    if (q.ray < 0 || q.npend == 0)
                        continue;
                    const Ray &r = rays[size_t(q.ray)];
                    const int k = std::min(q.npend, G);
                    for (int i = 0; i < k; i++) {
                        const int packed = ~q.pend[q.npend - 1 - i], pb = packed & ((1 << 26) - 1), cnt = packed >> 26;
                        maxc = std::max(maxc, cnt), act++, tsum += cnt;
                        S.l1_wavefronts += 1.5 * cnt;
                        for (int idx = pb; idx < pb + cnt; idx++) {
                            float u, v;
                            const float dist = hit_tri(sc, idx, r.o, r.d, u, v, INFINITY);
                            const bool valid = dist > 1e-3f && (dist < q.h.t || (dist == q.h.t && prim_before(sc, idx, q.h.prim)));
                            if (valid)
                                q.h.t = dist, q.h.u = u, q.h.v = v, q.h.prim = idx;
                        }
                    }
                    q.npend -= k;
                    if (q.npend <= G && q.cur >= 0)
                        q.walking = true;
                }
                S.tri_phases++, S.tri_tests += tsum;
                if (tri_cost > 0)
                    S.inst += 25 + tri_cost * maxc, S.thread_inst += 25 * act + tri_cost * tsum;
                else
                    S.inst += c.c_tri0 + 12 + c.c_tri * maxc, S.thread_inst += (c.c_tri0 + 12) * act + c.c_tri * tsum;
            }
            // ---- retire / refill
            int fetched = 0;
            for (auto &q : g) {
                if (q.ray < 0 || q.walking || q.npend > 0 || q.cur >= 0)
                    continue;
                const Hit &a = q.h, &b = ref[size_t(q.ray)];
                S.mismatches += !(a.prim == b.prim && a.t == b.t && a.u == b.u && a.v == b.v);
                S.rays++;
                if (fetch(q))
                    fetched++;
                else
                    live--;
            }
            if (fetched)
                S.inst += c.c_fetch, S.thread_inst += c.c_fetch * fetched * G;
        }
        // epilogue (full lanes) is common to all schemes and not counted
    }
    return S;
}


// ------------------------------------------------------------------ lock-step model: per-thread pair traversal with deferred leaves
// One ray per lane (the existing kernels' mapping).  A lane that reaches a leaf does not test it: the leaf goes to the lane's
// pending list (K entries) and the lane pops on.  The warp runs a triangle phase (every lane with a pending leaf tests one)
// when enough lanes have one, or when a lane is blocked (list full, or traversal finished with leaves pending).
// refill > 0: a lane that finished takes the next ray of the warp's buffer of `buf` rays as soon as `refill` lanes are idle.
static DefStats sim_thread_deferred(const Scene &sc, const std::vector<Ray> &rays, const std::vector<Hit> &ref, const Cost &c, int K, int flush_lanes,
                                    int buf, int refill) {
    DefStats S;
    struct Lane {
        long ray = -1;
        int cur = 0, sp = 0, start = 0, npend = 0;
        bool walking = false;
        Hit h;
        V3 inv_d, o_div;
        int stack_ref[128];
        float stack_t[128];
        int pend[8];
    };
    std::vector<Lane> L;
    L.resize(32);
    const size_t N = rays.size();
    for (size_t base = 0; base < N; base += size_t(buf)) {
        const size_t end = std::min(N, base + size_t(buf));
        size_t next = base;
        auto fetch = [&](Lane &q) {
            if (next >= end) {
                q.ray = -1, q.walking = false;
                return false;
            }
            q.ray = long(next++);
            const Ray &r = rays[size_t(q.ray)];
            q.inv_d = V3{1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z}, q.o_div = V3{r.o.x * q.inv_d.x, r.o.y * q.inv_d.y, r.o.z * q.inv_d.z};
            q.start = cached_entry(sc, q.inv_d, q.o_div, r.tmax, 0);
            q.h = Hit();
            q.h.t = r.tmax;
            q.cur = sc.pl.root, q.sp = 0, q.npend = 0, q.walking = true;
            return true;
        };
        int live = 0;
        for (auto &q : L)
            live += fetch(q);
        S.inst += c.fetch + c.cache * 6, S.thread_inst += (c.fetch + c.cache * 6 * 0.7) * live;
        while (live > 0) {
            int n_inner = 0, n_leafpend = 0, max_pop = 0, sum_pop = 0;
            for (auto &q : L) {
                if (q.ray < 0 || !q.walking)
                    continue;
                bool descend = false;
                int pops = 0;
                if (q.cur >= 0) {
                    n_inner++;
                    S.visits++;
                    S.l1_wavefronts += 2;
                    const F4 a0 = sc.pl.wnodes[4 * size_t(q.cur)], a1 = sc.pl.wnodes[4 * size_t(q.cur) + 1], b0 = sc.pl.wnodes[4 * size_t(q.cur) + 2],
                             b1 = sc.pl.wnodes[4 * size_t(q.cur) + 3];
                    float t0, t1;
                    const bool h0 = slab(a0, a1, q.inv_d, q.o_div, t0) && t0 < q.h.t && as_int(a1.w) > q.start;
                    const bool h1 = slab(b0, b1, q.inv_d, q.o_div, t1) && t1 < q.h.t && as_int(b1.w) > q.start;
                    const int r0 = as_int(a0.w), r1 = as_int(b0.w);
                    if (h0 && h1) {
                        const bool swap = t1 < t0;
                        q.stack_ref[q.sp] = swap ? r0 : r1, q.stack_t[q.sp] = swap ? t0 : t1, q.sp++;
                        q.cur = swap ? r1 : r0;
                        descend = true;
                    } else if (h0 || h1) {
                        q.cur = h0 ? r0 : r1;
                        descend = true;
                    }
                } else {
                    n_leafpend++;
                    q.pend[q.npend++] = q.cur;
                }
                if (!descend) {
                    bool found = false;
                    while (q.sp > 0) {
                        --q.sp, pops++;
                        if (q.stack_t[q.sp] < q.h.t) {
                            q.cur = q.stack_ref[q.sp];
                            found = true;
                            break;
                        }
                    }
                    if (!found)
                        q.walking = false, q.cur = 0x7fffffff;
                }
                if (pops)
                    max_pop = std::max(max_pop, pops), sum_pop += pops;
                if (q.npend >= K && q.walking && q.cur < 0)
                    q.walking = false; // next step would need a free pending slot
            }
            if (n_inner)
                S.inst += c.inner, S.thread_inst += c.inner * n_inner;
            if (n_leafpend)
                S.inst += 8, S.thread_inst += 8.0 * n_leafpend;
            if (max_pop)
                S.inst += c.pop * max_pop, S.thread_inst += c.pop * sum_pop;
            S.inst += 10, S.thread_inst += 10.0 * 32; // scheduling votes
            // triangle phase
            int have = 0, blocked = 0;
            for (auto &q : L) {
                if (q.ray < 0)
                    continue;
                have += q.npend > 0;
                blocked += (!q.walking && q.npend > 0);
            }
            if (have >= flush_lanes || blocked > 0) {
                int maxc = 0, act = 0;
                double tsum = 0;
                for (auto &q : L) {
                    if (q.ray < 0 || q.npend == 0)
                        continue;
                    const Ray &r = rays[size_t(q.ray)];
                    const int packed = ~q.pend[--q.npend], pb = packed & ((1 << 26) - 1), cnt = packed >> 26;
                    maxc = std::max(maxc, cnt), act++, tsum += cnt;
                    S.l1_wavefronts += 3.0 * cnt;
                    for (int idx = pb; idx < pb + cnt; idx++) {
                        float u, v;
                        const float dist = hit_tri(sc, idx, r.o, r.d, u, v, INFINITY);
                        const bool valid = dist > 1e-3f && (dist < q.h.t || (dist == q.h.t && prim_before(sc, idx, q.h.prim)));
                        if (valid)
                            q.h.t = dist, q.h.u = u, q.h.v = v, q.h.prim = idx;
                    }
                    if (q.cur != 0x7fffffff)
                        q.walking = true;
                }
                S.tri_phases++, S.tri_tests += tsum;
                S.inst += c.leaf0 + c.tri * maxc, S.thread_inst += c.leaf0 * act + c.tri * tsum;
            }
            // retire / refill
            int idle = 0;
            for (auto &q : L)
                if (q.ray >= 0 && !q.walking && q.npend == 0 && q.cur == 0x7fffffff) {
                    const Hit &a = q.h, &b = ref[size_t(q.ray)];
                    S.mismatches += !(a.prim == b.prim && a.t == b.t && a.u == b.u && a.v == b.v);
                    S.rays++;
                    q.ray = -1;
                    live--;
                }
            for (auto &q : L)
                idle += q.ray < 0;
            if (refill > 0 && idle >= refill && next < end) {
                int got = 0;
                for (auto &q : L)
                    if (q.ray < 0 && fetch(q))
                        got++, live++;
                // refill = epilogue of the finished rays + prologue of the new ones at `got` lanes (prologue/epilogue of the static
                // scheme run at 32 lanes and are not counted there either, except the cached phase)
                S.inst += c.fetch + c.cache * 6, S.thread_inst += (c.fetch + c.cache * 6 * 0.7) * got;
            }
        }
    }
    return S;
}

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: trav_sim scene.{xml,cpts} [x0 y0 w h] [bounces]\n");
        return 1;
    }
    Scene sc;
    const std::string path = argv[1];
    if (path.size() > 4 && path.substr(path.size() - 4) == ".xml")
        load_xml(path, sc.hs);
    else
        load_cpts(path, sc.hs);
    build_triangles(sc.hs, sc.tris);
    build_pair_layout(sc.hs, sc.pl, 2);
    if (sc.pl.refined_leaves > 0) {
        std::vector<F4> t2(sc.tris.size());
        for (size_t i = 0; i < sc.pl.perm.size(); i++)
            for (int k = 0; k < 3; k++)
                t2[3 * i + k] = sc.tris[3 * size_t(sc.pl.perm[i]) + k];
        sc.tris.swap(t2);
        sc.prim_orig.assign(sc.pl.perm.begin(), sc.pl.perm.end());
    }
    build_wide_layout(sc.pl, 4, sc.w4);
    build_wide_layout(sc.pl, 8, sc.w8);
    const int W = sc.hs.width(), H = sc.hs.height();
    int x0 = argc > 5 ? atoi(argv[2]) : 0, y0 = argc > 5 ? atoi(argv[3]) : 0, ww = argc > 5 ? atoi(argv[4]) : W, wh = argc > 5 ? atoi(argv[5]) : H;
    const int bounces = argc > 6 ? atoi(argv[6]) : 3;
    printf("scene %s: %dx%d, %d prims, %d nodes (%zu cached), pair nodes %zu depth %d refined %d\n", path.c_str(), W, H, sc.hs.num_prims(),
           sc.hs.num_nodes(), sc.hs.cache_nodes.size(), sc.pl.wnodes.size() / 4, sc.pl.max_depth, sc.pl.refined_leaves);
    for (const WideLayout *w : {&sc.w4, &sc.w8})
        printf("wide%d: %d nodes (%.1f MB), depth %d, max stack %d, avg children %.2f (inner %d, leaf %d)\n", w->width, w->num_nodes,
               w->nodes.size() * 16.0 / 1e6, w->max_depth, w->max_stack, w->avg_children, w->inner_slots, w->leaf_slots);

    // primary rays, 8x4 tiles in row-major tile order (k_raygen)
    const float *cam = sc.hs.camera;
    auto colv = [&](int i) { return V3{cam[i], cam[3 + i], cam[6 + i]}; };
    auto rot = [&](V3 v) { return V3{cam[0] * v.x + cam[1] * v.y + cam[2] * v.z, cam[3] * v.x + cam[4] * v.y + cam[5] * v.z, cam[6] * v.x + cam[7] * v.y + cam[8] * v.z}; };
    const V3 ct{cam[9], cam[10], cam[11]};
    std::mt19937 rng(1234);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::vector<Ray> rays;
    for (int ty = y0 / 4; ty < (y0 + wh) / 4; ty++)
        for (int tx = x0 / 8; tx < (x0 + ww) / 8; tx++)
            for (int l = 0; l < 32; l++) {
                const int px = tx * 8 + (l & 7), py = ty * 4 + (l >> 3);
                const float xp = U(rng) + px, yp = U(rng) + py;
                const float nx = (xp - cam[CPTS_CAM_HW]) * cam[CPTS_CAM_INV_FOCAL] * cam[CPTS_CAM_SIGN_X], ny = (cam[CPTS_CAM_HH] - yp) * cam[CPTS_CAM_INV_FOCAL];
                V3 dir = norm(rot(V3{nx, ny, 1.f}));
                Ray r{ct, dir, 1e7f};
                if (cam[CPTS_CAM_APERTURE] > 0) {
                    const V3 focus = ct + dir * cam[CPTS_CAM_FOCAL_DIST];
                    const float rr = std::sqrt(U(rng)) * cam[CPTS_CAM_APERTURE], ph = 6.2831853f * U(rng);
                    r.o = ct + colv(0) * (rr * std::cos(ph)) + colv(1) * (rr * std::sin(ph));
                    r.d = norm(focus - r.o);
                }
                rays.push_back(r);
            }
    Cost cost;
    for (int b = 0; b <= bounces; b++) {
        const size_t N = rays.size();
        std::vector<Trace> tr(N);
        std::vector<Hit> hits(N);
        size_t mism4 = 0, mism8 = 0, mismt4 = 0, mismt8 = 0, mism8n = 0, plain8 = 0, plain4 = 0;
        int max_sp4 = 0, max_sp8 = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : mism4, mism8, mismt4, mismt8, mism8n, plain8, plain4) reduction(max : max_sp4, max_sp8)
        for (size_t i = 0; i < N; i++) {
            hits[i] = trace_pair(sc, rays[i], false, &tr[i], &tr[i].pair);
            const Hit a = trace_coop(sc, sc.w4, rays[i], false, true, &tr[i].c4, &max_sp4);
            const Hit c8 = trace_coop(sc, sc.w8, rays[i], false, true, &tr[i].c8, &max_sp8);
            std::vector<CStep> tmp8, tmp4;
            const Hit c8n = trace_coop(sc, sc.w8, rays[i], false, false, &tmp8);
            trace_coop(sc, sc.w4, rays[i], false, false, &tmp4);
            plain8 += tmp8.size(), plain4 += tmp4.size();
            const Hit t4 = trace_wide_thread(sc, sc.w4, rays[i], &tr[i].w4t);
            const Hit t8 = trace_wide_thread(sc, sc.w8, rays[i], &tr[i].w8t);
            auto same = [&](const Hit &x) { return x.prim == hits[i].prim && x.t == hits[i].t && x.u == hits[i].u && x.v == hits[i].v; };
            mism4 += !same(a), mism8 += !same(c8), mismt4 += !same(t4), mismt8 += !same(t8), mism8n += !same(c8n);
        }
        double s_pair = 0, s_pair_in = 0, s_pair_leaf = 0, s_tri = 0, s_c4 = 0, s_c8 = 0, s_t4 = 0, s_t8 = 0, s_c8_tris = 0, s_c8_leafsteps = 0, s_cache = 0;
        size_t nhit = 0;
        std::vector<int> cache_iters(N);
        std::vector<size_t> lens(N);
        for (size_t i = 0; i < N; i++) {
            s_pair += tr[i].pair.size();
            for (const Step &s : tr[i].pair)
                s_pair_in += s.kind == 0, s_pair_leaf += s.kind == 1, s_tri += s.kind == 1 ? s.cnt : 0;
            s_c4 += tr[i].c4.size(), s_c8 += tr[i].c8.size(), s_t4 += tr[i].w4t.size(), s_t8 += tr[i].w8t.size();
            for (const CStep &s : tr[i].c8)
                s_c8_tris += s.leaf_tris, s_c8_leafsteps += s.leaf_lanes > 0;
            nhit += hits[i].prim >= 0;
            cache_iters[i] = tr[i].cache_iters, s_cache += tr[i].cache_iters;
            lens[i] = tr[i].pair.size();
        }
        std::vector<size_t> sorted = lens;
        std::sort(sorted.begin(), sorted.end());
        printf("\n== bounce %d: %zu rays, %.1f%% hit.  mismatches vs pair: coop4 %zu coop8 %zu coop8(no pop-nearest) %zu thread4 %zu thread8 %zu; max sp coop4 %d coop8 %d\n",
               b, N, 100.0 * nhit / N, mism4, mism8, mism8n, mismt4, mismt8, max_sp4, max_sp8);
        printf("   per ray: pair steps %.1f (inner %.1f leaf %.1f, tris %.1f; median %zu p90 %zu p99 %zu max %zu) | cached-phase iters %.1f\n", s_pair / N,
               s_pair_in / N, s_pair_leaf / N, s_tri / N, sorted[N / 2], sorted[N * 9 / 10], sorted[N * 99 / 100], sorted[N - 1], s_cache / N);
        printf("   per ray: plain LIFO pop: coop4 visits %.1f coop8 visits %.1f\n", double(plain4) / N, double(plain8) / N);
        printf("   per ray: coop4 visits %.1f | coop8 visits %.1f (tris %.1f, visits with leaf tests %.1f) | thread4 steps %.1f | thread8 steps %.1f\n", s_c4 / N,
               s_c8 / N, s_c8_tris / N, s_c8_leafsteps / N, s_t4 / N, s_t8 / N);
        std::vector<const std::vector<Step> *> p(N), t4(N), t8(N);
        std::vector<const std::vector<CStep> *> c4(N), c8(N);
        for (size_t i = 0; i < N; i++)
            p[i] = &tr[i].pair, t4[i] = &tr[i].w4t, t8[i] = &tr[i].w8t, c4[i] = &tr[i].c4, c8[i] = &tr[i].c8;
        auto show = [&](const char *name, const Tally &T) { printf("   %-34s warp-inst/ray %8.1f   lanes/inst %5.2f\n", name, T.inst / N, T.lanes()); };
        show("per-thread pair, static 32", model_thread(p, cache_iters, cost, 32, false));
        { // L1 line model: wavefronts per ray of the static kernel for different assignments of rays to warps
            V3 lo{1e30f, 1e30f, 1e30f}, hi{-1e30f, -1e30f, -1e30f};
            for (const Ray &r : rays)
                lo = V3{std::min(lo.x, r.o.x), std::min(lo.y, r.o.y), std::min(lo.z, r.o.z)}, hi = V3{std::max(hi.x, r.o.x), std::max(hi.y, r.o.y), std::max(hi.z, r.o.z)};
            auto key = [&](size_t i, int mode) -> uint64_t {
                const Ray &r = rays[i];
                const uint32_t oct = (r.d.x < 0) | ((r.d.y < 0) << 1) | ((r.d.z < 0) << 2);
                auto q = [&](float v, float a, float b) { return uint32_t(std::min(1023.f, std::max(0.f, (v - a) / std::max(b - a, 1e-20f) * 1023.f))); };
                const uint32_t mo = morton3(q(r.o.x, lo.x, hi.x), q(r.o.y, lo.y, hi.y), q(r.o.z, lo.z, hi.z));
                // finer direction cell: 6 x 8 x 8 cube-map cells
                const float ax = std::fabs(r.d.x), ay = std::fabs(r.d.y), az = std::fabs(r.d.z);
                int face;
                float u, v;
                if (ax >= ay && ax >= az)
                    face = r.d.x < 0, u = r.d.y / ax, v = r.d.z / ax;
                else if (ay >= az)
                    face = 2 + (r.d.y < 0), u = r.d.x / ay, v = r.d.z / ay;
                else
                    face = 4 + (r.d.z < 0), u = r.d.x / az, v = r.d.y / az;
                const uint32_t cell = uint32_t(face) * 64 + uint32_t(std::min(7.f, (u * 0.5f + 0.5f) * 8.f)) * 8 + uint32_t(std::min(7.f, (v * 0.5f + 0.5f) * 8.f));
                if (mode == 1)
                    return (uint64_t(oct) << 32) | mo;       // octant, then origin
                if (mode == 2)
                    return (uint64_t(mo >> 12) << 32) | (uint64_t(cell) << 20) | (mo & 0xfffu); // coarse origin cell (18 bits), direction cell, fine origin
                return (uint64_t(cell) << 32) | mo;         // direction cell, then origin
            };
            auto report = [&](const char *name, const std::vector<uint32_t> &order) {
                const LineStats S = model_lines(p, order);
                printf("   L1 lines, %-44s node wavefronts/ray %6.1f (of %6.1f lane loads)  tri wavefronts/ray %5.1f (of %5.1f)  lanes/iteration %5.2f  warp iterations/ray %5.1f\n", name,
                       S.node_wavefronts / S.rays, S.node_lane_loads / S.rays, S.tri_wavefronts / S.rays, S.tri_lane_loads / S.rays, S.lane_iterations / std::max(1.0, S.iterations),
                       S.iterations / S.rays);
            };
            std::vector<uint32_t> order(N);
            for (size_t i = 0; i < N; i++)
                order[i] = uint32_t(i);
            report("queue order (what k_trace<2> gets)", order);
            for (int window : {256, 1024, 4096, 65536})
                for (int mode : {1, 2, 3}) {
                    std::vector<uint32_t> o2 = order;
                    for (size_t b0 = 0; b0 < N; b0 += size_t(window)) {
                        const size_t e0 = std::min(N, b0 + size_t(window));
                        std::sort(o2.begin() + b0, o2.begin() + e0, [&](uint32_t a, uint32_t c) { return key(a, mode) < key(c, mode); });
                    }
                    char nm[96];
                    snprintf(nm, sizeof nm, "sorted in windows of %d by %s", window, mode == 1 ? "octant | origin" : (mode == 2 ? "origin cell | dir cell" : "dir cell | origin"));
                    report(nm, o2);
                }
        }
        show("per-thread pair, refill buf 128", model_thread(p, cache_iters, cost, 128, false, 4));
        show("per-thread pair, refill buf 512", model_thread(p, cache_iters, cost, 512, false, 4));
        show("per-thread pair, refill buf 4096", model_thread(p, cache_iters, cost, 4096, false, 4));
        show("per-thread wide4, static 32", model_thread(t4, cache_iters, cost, 32, true));
        show("per-thread wide8, static 32", model_thread(t8, cache_iters, cost, 32, true));
        show("per-thread wide4, refill buf 512", model_thread(t4, cache_iters, cost, 512, true, 4));
        show("coop G=4, chunk 32", model_coop(c4, cache_iters, cost, 4, 32));
        show("coop G=4, chunk 64", model_coop(c4, cache_iters, cost, 4, 64));
        show("coop G=8, chunk 32", model_coop(c8, cache_iters, cost, 8, 32));
        show("coop G=8, chunk 64", model_coop(c8, cache_iters, cost, 8, 64));
        show("coop G=8, chunk 256", model_coop(c8, cache_iters, cost, 8, 256));
        for (int bufsz : {64, 128, 256, 1024}) {
            const DefStats D = sim_thread_deferred(sc, rays, hits, cost, 4, 16, bufsz, 8);
            printf("   per-thread deferred K=4 flush>=16 refill 8, chunk %4d: warp-inst/ray %7.1f lanes/inst %5.2f  visits/ray %5.1f tri-phases/ray %4.2f\n", bufsz,
                   D.inst / D.rays, D.thread_inst / D.inst, D.visits / D.rays, D.tri_phases / D.rays);
        }
        for (int K : {2, 4})
            for (int fl : {1, 8, 16, 24})
                for (int rf : {0, 8}) {
                    const DefStats D = sim_thread_deferred(sc, rays, hits, cost, K, fl, rf ? 4096 : 32, rf);
                    printf("   per-thread deferred K=%d flush>=%2d refill %d: warp-inst/ray %7.1f lanes/inst %5.2f  visits/ray %5.1f tri-tests/ray %4.1f "
                           "tri-phases/ray %4.2f L1 wavefronts/ray %5.1f mismatches %g\n",
                           K, fl, rf, D.inst / D.rays, D.thread_inst / D.inst, D.visits / D.rays, D.tri_tests / D.rays, D.tri_phases / D.rays,
                           D.l1_wavefronts / D.rays, D.mismatches);
                }
        for (int mb : {1, 2, 3, 4, 6})
            for (int ch : {32, 64}) {
                const DefStats D = sim_coop_deferred(sc, sc.w4, rays, hits, cost, ch, 99, 0, mb, 107, 65);
                printf("   coop G=4 measured costs, flush when >= %d groups blocked, chunk %d: warp-inst/ray %7.1f lanes/inst %5.2f visits/ray %5.1f "
                       "tri-phases/ray %4.2f mismatches %g\n", mb, ch, D.inst / D.rays, D.thread_inst / D.inst, D.visits / D.rays, D.tri_phases / D.rays, D.mismatches);
            }
        for (int G : {4, 8})
            for (int fl : {1, 12, 20, 28})
                for (int slab_n : {0, 256}) {
                    if (slab_n && fl != 20)
                        continue;
                    const DefStats D = sim_coop_deferred(sc, G == 4 ? sc.w4 : sc.w8, rays, hits, cost, 32, fl, slab_n);
                    printf("   coop G=%d deferred, flush>=%2d lanes, slab %3d: warp-inst/ray %7.1f lanes/inst %5.2f  visits/ray %5.1f tri-tests/ray %4.1f "
                           "tri-phases/ray %4.2f L1 wavefronts/ray %5.1f mismatches %g\n",
                           G, fl, slab_n, D.inst / D.rays, D.thread_inst / D.inst, D.visits / D.rays, D.tri_tests / D.rays, D.tri_phases / D.rays,
                           D.l1_wavefronts / D.rays, D.mismatches);
                }
        if (b == bounces)
            break;
        // next bounce: cosine-distributed direction about the geometric normal facing the incoming ray; survivors keep queue order
        std::vector<Ray> nxt;
        for (size_t i = 0; i < N; i++) {
            if (hits[i].prim < 0)
                continue;
            const F4 p0 = sc.tris[3 * size_t(hits[i].prim)], e1 = sc.tris[3 * size_t(hits[i].prim) + 1], e2 = sc.tris[3 * size_t(hits[i].prim) + 2];
            (void)p0;
            V3 n = norm(cross(V3{e1.x, e1.y, e1.z}, V3{e2.x, e2.y, e2.z}));
            if (dot(n, rays[i].d) > 0)
                n = n * -1.f;
            const V3 pos = rays[i].o + rays[i].d * hits[i].t;
            const float u1 = U(rng), u2 = U(rng);
            const float rr = std::sqrt(u1), ph = 6.2831853f * u2;
            const V3 a = std::fabs(n.x) > 0.5f ? V3{0, 1, 0} : V3{1, 0, 0};
            const V3 tx = norm(cross(a, n)), ty = cross(n, tx);
            const V3 d = norm(tx * (rr * std::cos(ph)) + ty * (rr * std::sin(ph)) + n * std::sqrt(std::max(0.f, 1.f - u1)));
            if (U(rng) < 0.85f) // some absorption so that queues thin out like the real thing
                nxt.push_back(Ray{pos, d, 1e7f});
        }
        rays.swap(nxt);
    }
    return 0;
}
