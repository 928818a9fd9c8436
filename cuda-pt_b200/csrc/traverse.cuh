// traverse.cuh — BVH traversal for the wavefront kernels.
//
// Two interchangeable back-ends produce the reference's hit set:
//
//  * `closest_hit_dfs` / `occluded_dfs`: the reference's own stackless
//    skip-offset DFS over its 32 B linear nodes (reference:
//    src/renderer/tracing_func.cuh:44-99,117-181).  Kept as the strict
//    comparison mode (CPTB200_TRAVERSAL=dfs) and for tiny scenes.
//
//  * `closest_hit_ordered` / `occluded_ordered`: near-child-first traversal with
//    a short per-thread stack over "pair nodes" (both children's boxes in one
//    64 B record, built by bvh_relayout.cpp from the SAME tree, boxes and leaf
//    ranges).  Every box keeps the reference's fp32 bounds and is tested with
//    the reference's slab arithmetic (src/core/aabb.cuh:58-65), so the set of
//    leaves whose primitives get tested is the same; only the visiting order
//    changes (DESIGN.md §"Why ordered traversal is parity-safe").
//
// Both honour the reference's cached near-root phase: up to 255 fp16 nodes in
// shared memory choose the DFS index at which global traversal starts, and
// everything that precedes that index in DFS order is skipped
// (tracing_func.cuh:124-142; SURVEY.md App. A.9 — including the round-to-nearest
// fp16 bounds quirk).
#pragma once
#include "pt_device.cuh"

namespace cptb {

// slab test, same fma/min/max structure as AABB::intersect (aabb.cuh:58-65, vec3.cuh:263-276)
// tie rule of the closest hit: the reference keeps the first hit in its DFS order = the smaller reference primitive index.
// Refined leaves (scene_host.cpp build_pair_layout) re-order primitives inside a leaf; prim_orig maps back.
PT_DEV bool prim_before(const DevScene &sc, int a, int b) {
    if (b < 0)
        return false;
    return sc.prim_orig ? __ldg(sc.prim_orig + a) < __ldg(sc.prim_orig + b) : a < b;
}

PT_DEV bool slab_test(float minx, float miny, float minz, float maxx, float maxy, float maxz, V3 inv_d, V3 o_div,
                      float &t_near) {
    const float ax = fmaf(minx, inv_d.x, -o_div.x), ay = fmaf(miny, inv_d.y, -o_div.y), az = fmaf(minz, inv_d.z, -o_div.z);
    const float bx = fmaf(maxx, inv_d.x, -o_div.x), by = fmaf(maxy, inv_d.y, -o_div.y), bz = fmaf(maxz, inv_d.z, -o_div.z);
    const float min_x = ax < bx ? ax : bx, max_x = ax < bx ? bx : ax;
    const float min_y = ay < by ? ay : by, max_y = ay < by ? by : ay;
    const float min_z = az < bz ? az : bz, max_z = az < bz ? bz : az;
    t_near = fmaxf(min_x, fmaxf(min_y, min_z));
    const float t_far = fminf(max_x, fminf(max_y, max_z));
    return (t_far > t_near) && (t_far > 0);
}

// Same test with min/max instructions instead of compare+select pairs.  Value-identical to slab_test whenever no
// operand is NaN, i.e. for rays whose reciprocal direction stays below 1e28 in all three components (then every
// product and difference above is finite); the two differ only in how a NaN operand is routed, so rays with a (near-)
// zero direction component take the select form (dispatch in closest_hit<2> / unoccluded<2>).  6 FMNMX + 2 FMNMX3 per
// box instead of 3 FSETP + 6 FSEL + 2 FMNMX3.
PT_DEV bool slab_test_finite(float minx, float miny, float minz, float maxx, float maxy, float maxz, V3 inv_d, V3 o_div,
                             float &t_near) {
    const float ax = fmaf(minx, inv_d.x, -o_div.x), ay = fmaf(miny, inv_d.y, -o_div.y), az = fmaf(minz, inv_d.z, -o_div.z);
    const float bx = fmaf(maxx, inv_d.x, -o_div.x), by = fmaf(maxy, inv_d.y, -o_div.y), bz = fmaf(maxz, inv_d.z, -o_div.z);
    t_near = fmaxf(fminf(ax, bx), fmaxf(fminf(ay, by), fminf(az, bz)));
    const float t_far = fminf(fmaxf(ax, bx), fminf(fmaxf(ay, by), fmaxf(az, bz)));
    return (t_far > t_near) && (t_far > 0);
}
template <bool FINITE>
PT_DEV bool slab_any(float minx, float miny, float minz, float maxx, float maxy, float maxz, V3 inv_d, V3 o_div, float &t_near) {
    if constexpr (FINITE)
        return slab_test_finite(minx, miny, minz, maxx, maxy, maxz, inv_d, o_div, t_near);
    else
        return slab_test(minx, miny, minz, maxx, maxy, maxz, inv_d, o_div, t_near);
}

// cached fp16 phase: returns the DFS node index at which global traversal starts, or `miss_value`
// when no cached frontier node is intersected (bvh.cuh:182-251, tracing_func.cuh:53-65,128-140)
PT_DEV int cached_entry(const uint4 *__restrict__ s_cached, int cache_num, V3 inv_d, V3 o_div, float limit, int miss_value) {
    int node_idx = miss_value;
    float t_near = 0;
    for (int i = 0; i < cache_num;) {
        const uint4 c = s_cached[i];
        const float2 xx = __half22float2(*reinterpret_cast<const __half2 *>(&c.x));
        const float2 yy = __half22float2(*reinterpret_cast<const __half2 *>(&c.y));
        const float2 zz = __half22float2(*reinterpret_cast<const __half2 *>(&c.z));
        bool hit = slab_test(xx.x, yy.x, zz.x, xx.y, yy.y, zz.y, inv_d, o_div, t_near) && t_near < limit;
        const int all_offset = int(c.w & 0xFFu), gmem_index = int((c.w >> 8) & 0x00FFFFFFu);
        const int increment = (!hit) * all_offset + int(hit && all_offset != 1);
        hit = hit && all_offset == 1;
        i = hit ? cache_num : (i + increment);
        node_idx = hit ? gmem_index : node_idx;
    }
    return node_idx;
}

struct Hit {
    float t, u, v;
    int prim;     // -1: miss
    int obj_word; // raw per-primitive object word
};

// ---------------------------------------------------------------------------------------------
// strict mode: reference DFS
// ---------------------------------------------------------------------------------------------
PT_DEV Hit closest_hit_dfs(const DevScene &sc, const uint4 *s_cached, V3 o, V3 d, float t_max = kMaxDist) {
    Hit h{t_max, 0.f, 0.f, -1, 0};
    const V3 inv_d = rcp3(d), o_div = o * inv_d;
    int node_idx = cached_entry(s_cached, sc.num_cache, inv_d, o_div, t_max, 0);
    float t_near = 0;
    while (node_idx < sc.num_nodes) {
        const float4 lo = __ldg(sc.nodes + 2 * node_idx), hi = __ldg(sc.nodes + 2 * node_idx + 1);
        const int beg = __float_as_int(lo.w);
        int end = __float_as_int(hi.w);
        const bool hit = slab_test(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, inv_d, o_div, t_near) && t_near < h.t;
        node_idx += (!hit) * (end < 0 ? -end : 1) + int(hit);
        end = hit ? end + beg : 0;
        for (int idx = beg; idx < end; idx++) {
            float u, v;
            const float dist = hit_triangle(sc.tris, idx, o, d, u, v, __int_as_float(0x7f800000));
            const bool valid = dist > kEps && (dist < h.t || (dist == h.t && h.prim >= 0 && prim_before(sc, idx, h.prim)));
            h.t = valid ? dist : h.t;
            h.u = valid ? u : h.u;
            h.v = valid ? v : h.v;
            h.prim = valid ? idx : h.prim;
        }
    }
    if (h.prim >= 0)
        h.obj_word = __ldg(sc.obj_idxs + h.prim);
    return h;
}

// true == unoccluded
PT_DEV bool unoccluded_dfs(const DevScene &sc, const uint4 *s_cached, V3 o, V3 d, float max_dist) {
    const V3 inv_d = rcp3(d), o_div = o * inv_d;
    int node_idx = cached_entry(s_cached, sc.num_cache, inv_d, o_div, max_dist, sc.num_nodes);
    float t_near = 0;
    while (node_idx < sc.num_nodes) {
        const float4 lo = __ldg(sc.nodes + 2 * node_idx), hi = __ldg(sc.nodes + 2 * node_idx + 1);
        const int beg = __float_as_int(lo.w);
        int end = __float_as_int(hi.w);
        const bool hit = slab_test(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, inv_d, o_div, t_near) && t_near < max_dist;
        node_idx += (!hit) * (end < 0 ? -end : 1) + int(hit);
        end = hit ? end + beg : 0;
        for (int idx = beg; idx < end; idx++) {
            float u, v;
            const float dist = hit_triangle(sc.tris, idx, o, d, u, v, max_dist);
            if (dist > kEps)
                return false;
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// ordered mode: pair nodes + per-thread stack
//   pair node (4 x float4): {c0.min, ref0} {c0.max, end0} {c1.min, ref1} {c1.max, end1}
//   ref >= 0 : index of an inner pair node;  ref < 0 : leaf, ~ref = prim_base | (prim_count << 26)
//   end      : one past the last reference-DFS node index inside that child's subtree
// ---------------------------------------------------------------------------------------------
constexpr int kStackSize = 48;
constexpr int kLeafCountShift = 26;
constexpr int kLeafBaseMask = (1 << kLeafCountShift) - 1;

constexpr int kTraversalDone = 0x7fffffff; // "cur" value meaning: stack exhausted

// one pair node = 64 B.  WIDE: two 256-bit loads (LDG.E.256, sm_100) instead of four 128-bit ones - half the
// load instructions and L1 line visits per node for the same bytes
template <bool WIDE>
PT_DEV void load_pair_node(const float4 *__restrict__ wnodes, int cur, float4 &a0, float4 &a1, float4 &b0, float4 &b1) {
    if constexpr (WIDE) {
        const float4 *p = wnodes + 4 * cur;
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(a0.x), "=f"(a0.y), "=f"(a0.z), "=f"(a0.w), "=f"(a1.x), "=f"(a1.y), "=f"(a1.z), "=f"(a1.w)
                     : "l"(p));
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(b0.x), "=f"(b0.y), "=f"(b0.z), "=f"(b0.w), "=f"(b1.x), "=f"(b1.y), "=f"(b1.z), "=f"(b1.w)
                     : "l"(p + 2));
    } else {
        a0 = __ldg(wnodes + 4 * cur), a1 = __ldg(wnodes + 4 * cur + 1), b0 = __ldg(wnodes + 4 * cur + 2),
        b1 = __ldg(wnodes + 4 * cur + 3);
    }
}

// variant with one inner-or-leaf step per loop iteration (lanes alternate freely between the two paths)
// CPTB_STACK64 (default 1): one 8-byte stack entry {ref, bits(t)} per push / pop (STL.64 / LDL.64) instead of two 4-byte arrays -
// half the local-memory instructions of the stack traffic (27 % of k_trace<2>'s L1 wavefronts).  Measured: k_trace 3.31 -> 3.28 ms per
// sample (profiles/r02_stream_experiments.md)
#ifndef CPTB_STACK64
#define CPTB_STACK64 1
#endif
template <bool WIDE, bool FINITE>
PT_DEV Hit closest_hit_ordered(const DevScene &sc, const uint4 *s_cached, V3 o, V3 d, float t_max = kMaxDist) {
    Hit h{t_max, 0.f, 0.f, -1, 0};
    const V3 inv_d = rcp3(d), o_div = o * inv_d;
    const int start = cached_entry(s_cached, sc.num_cache, inv_d, o_div, t_max, 0);
#if CPTB_STACK64
    uint2 stack[kStackSize];
#else
    int stack_ref[kStackSize];
    float stack_t[kStackSize];
#endif
    int sp = 0;
    int cur = sc.wroot;
    const float inf = __int_as_float(0x7f800000);
    while (true) {
        if (cur >= 0) {
            float4 a0, a1, b0, b1;
            load_pair_node<WIDE>(sc.wnodes, cur, a0, a1, b0, b1);
            float t0, t1;
            const bool h0 = slab_any<FINITE>(a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, inv_d, o_div, t0) && t0 < h.t &&
                            __float_as_int(a1.w) > start;
            const bool h1 = slab_any<FINITE>(b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, inv_d, o_div, t1) && t1 < h.t &&
                            __float_as_int(b1.w) > start;
            const int r0 = __float_as_int(a0.w), r1 = __float_as_int(b0.w);
            if (h0 && h1) {
                const bool swap = t1 < t0;
#if CPTB_STACK64
                stack[sp] = make_uint2(uint32_t(swap ? r0 : r1), __float_as_uint(swap ? t0 : t1));
#else
                stack_ref[sp] = swap ? r0 : r1;
                stack_t[sp] = swap ? t0 : t1;
#endif
                ++sp;
                cur = swap ? r1 : r0;
                continue;
            }
            if (h0 || h1) {
                cur = h0 ? r0 : r1;
                continue;
            }
        } else {
            const int packed = ~cur, base = packed & kLeafBaseMask, cnt = packed >> kLeafCountShift;
            for (int idx = base; idx < base + cnt; idx++) {
                float u, v;
                const float dist = hit_triangle(sc.tris, idx, o, d, u, v, inf);
                // equal distances: the reference keeps the first hit in DFS order == the smaller primitive index
                const bool valid = dist > kEps && (dist < h.t || (dist == h.t && prim_before(sc, idx, h.prim)));
                h.t = valid ? dist : h.t;
                h.u = valid ? u : h.u;
                h.v = valid ? v : h.v;
                h.prim = valid ? idx : h.prim;
            }
        }
        bool found = false;
        while (sp > 0) {
            --sp;
#if CPTB_STACK64
            const uint2 e = stack[sp];
            if (__uint_as_float(e.y) < h.t) {
                cur = int(e.x);
                found = true;
                break;
            }
#else
            if (stack_t[sp] < h.t) {
                cur = stack_ref[sp];
                found = true;
                break;
            }
#endif
        }
        if (!found)
            break;
    }
    if (h.prim >= 0)
        h.obj_word = __ldg(sc.obj_idxs + h.prim);
    return h;
}

template <bool WIDE, bool FINITE>
PT_DEV bool unoccluded_ordered(const DevScene &sc, const uint4 *s_cached, V3 o, V3 d, float max_dist) {
    const V3 inv_d = rcp3(d), o_div = o * inv_d;
    const int start = cached_entry(s_cached, sc.num_cache, inv_d, o_div, max_dist, sc.num_nodes);
    if (start >= sc.num_nodes)
        return true;
    int stack_ref[kStackSize];
    int sp = 0;
    int cur = sc.wroot;
    while (true) {
        if (cur >= 0) {
            float4 a0, a1, b0, b1;
            load_pair_node<WIDE>(sc.wnodes, cur, a0, a1, b0, b1);
            float t0, t1;
            const bool h0 = slab_any<FINITE>(a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, inv_d, o_div, t0) && t0 < max_dist &&
                            __float_as_int(a1.w) > start;
            const bool h1 = slab_any<FINITE>(b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, inv_d, o_div, t1) && t1 < max_dist &&
                            __float_as_int(b1.w) > start;
            const int r0 = __float_as_int(a0.w), r1 = __float_as_int(b0.w);
            if (h0 && h1) {
                stack_ref[sp++] = r1;
                cur = r0;
                continue;
            }
            if (h0 || h1) {
                cur = h0 ? r0 : r1;
                continue;
            }
        } else {
            const int packed = ~cur, base = packed & kLeafBaseMask, cnt = packed >> kLeafCountShift;
            for (int idx = base; idx < base + cnt; idx++) {
                float u, v;
                const float dist = hit_triangle(sc.tris, idx, o, d, u, v, max_dist);
                if (dist > kEps)
                    return false;
            }
        }
        if (sp == 0)
            break;
        cur = stack_ref[--sp];
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// "vote" variant: warp-synchronous step scheduler.  Every lane is either at an inner pair node or inside a
// leaf (one triangle per step).  Each iteration the warp ballots which of the two step kinds more lanes are
// waiting for and executes only that one, so a single code path runs per iteration at >= 50 % lane
// utilisation of the lanes still alive, instead of both paths back to back.  Must be called by all 32 lanes
// (lanes without a ray pass active = false).  Same boxes, same tests, same tie rule as the other ordered modes.
// ---------------------------------------------------------------------------------------------
PT_DEV Hit closest_hit_vote(const DevScene &sc, const uint4 *s_cached, V3 o, V3 d, bool active) {
    Hit h{kMaxDist, 0.f, 0.f, -1, 0};
    const V3 inv_d = rcp3(d), o_div = o * inv_d;
    const int start = active ? cached_entry(s_cached, sc.num_cache, inv_d, o_div, kMaxDist, 0) : 0;
    int stack_ref[kStackSize];
    float stack_t[kStackSize];
    int sp = 0;
    int cur = sc.wroot, ti = 0, te = 0;
    const float inf = __int_as_float(0x7f800000);
    while (true) {
        const bool want_inner = active && ti >= te, want_tri = active && ti < te;
        const unsigned mi = __ballot_sync(0xffffffffu, want_inner), mt = __ballot_sync(0xffffffffu, want_tri);
        if ((mi | mt) == 0u)
            break;
        int next = kTraversalDone; // set when this lane has to move on to a new node reference
        bool advance = false;
        if (__popc(mi) >= __popc(mt)) {
            if (want_inner) {
                const float4 a0 = __ldg(sc.wnodes + 4 * cur), a1 = __ldg(sc.wnodes + 4 * cur + 1),
                             b0 = __ldg(sc.wnodes + 4 * cur + 2), b1 = __ldg(sc.wnodes + 4 * cur + 3);
                float t0, t1;
                const bool h0 = slab_test(a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, inv_d, o_div, t0) && t0 < h.t &&
                                __float_as_int(a1.w) > start;
                const bool h1 = slab_test(b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, inv_d, o_div, t1) && t1 < h.t &&
                                __float_as_int(b1.w) > start;
                const int r0 = __float_as_int(a0.w), r1 = __float_as_int(b0.w);
                advance = true;
                if (h0 && h1) {
                    const bool swap = t1 < t0;
                    stack_ref[sp] = swap ? r0 : r1;
                    stack_t[sp] = swap ? t0 : t1;
                    ++sp;
                    next = swap ? r1 : r0;
                } else if (h0 || h1) {
                    next = h0 ? r0 : r1;
                }
            }
        } else if (want_tri) {
            float u, v;
            const float dist = hit_triangle(sc.tris, ti, o, d, u, v, inf);
            const bool valid = dist > kEps && (dist < h.t || (dist == h.t && prim_before(sc, ti, h.prim)));
            h.t = valid ? dist : h.t;
            h.u = valid ? u : h.u;
            h.v = valid ? v : h.v;
            h.prim = valid ? ti : h.prim;
            ++ti;
            advance = ti >= te;
        }
        if (advance) {
            if (next == kTraversalDone) { // pop the next entry that can still beat the closest hit
                while (sp > 0) {
                    --sp;
                    if (stack_t[sp] < h.t) {
                        next = stack_ref[sp];
                        break;
                    }
                }
            }
            if (next == kTraversalDone) {
                active = false;
            } else if (next < 0) {
                const int packed = ~next;
                ti = packed & kLeafBaseMask;
                te = ti + (packed >> kLeafCountShift);
            } else {
                cur = next;
                ti = te = 0;
            }
        }
    }
    if (h.prim >= 0)
        h.obj_word = __ldg(sc.obj_idxs + h.prim);
    return h;
}

PT_DEV bool unoccluded_vote(const DevScene &sc, const uint4 *s_cached, V3 o, V3 d, float max_dist, bool active) {
    const V3 inv_d = rcp3(d), o_div = o * inv_d;
    const int start = active ? cached_entry(s_cached, sc.num_cache, inv_d, o_div, max_dist, sc.num_nodes) : sc.num_nodes;
    bool clear = true;
    active = active && start < sc.num_nodes;
    int stack_ref[kStackSize];
    int sp = 0;
    int cur = sc.wroot, ti = 0, te = 0;
    while (true) {
        const bool want_inner = active && ti >= te, want_tri = active && ti < te;
        const unsigned mi = __ballot_sync(0xffffffffu, want_inner), mt = __ballot_sync(0xffffffffu, want_tri);
        if ((mi | mt) == 0u)
            break;
        int next = kTraversalDone;
        bool advance = false;
        if (__popc(mi) >= __popc(mt)) {
            if (want_inner) {
                const float4 a0 = __ldg(sc.wnodes + 4 * cur), a1 = __ldg(sc.wnodes + 4 * cur + 1),
                             b0 = __ldg(sc.wnodes + 4 * cur + 2), b1 = __ldg(sc.wnodes + 4 * cur + 3);
                float t0, t1;
                const bool h0 = slab_test(a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, inv_d, o_div, t0) && t0 < max_dist &&
                                __float_as_int(a1.w) > start;
                const bool h1 = slab_test(b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, inv_d, o_div, t1) && t1 < max_dist &&
                                __float_as_int(b1.w) > start;
                const int r0 = __float_as_int(a0.w), r1 = __float_as_int(b0.w);
                advance = true;
                if (h0 && h1) {
                    stack_ref[sp++] = r1;
                    next = r0;
                } else if (h0 || h1) {
                    next = h0 ? r0 : r1;
                }
            }
        } else if (want_tri) {
            float u, v;
            const float dist = hit_triangle(sc.tris, ti, o, d, u, v, max_dist);
            if (dist > kEps) {
                clear = false;
                active = false;
            }
            ++ti;
            advance = active && ti >= te;
        }
        if (advance) {
            if (next == kTraversalDone && sp > 0)
                next = stack_ref[--sp];
            if (next == kTraversalDone) {
                active = false;
            } else if (next < 0) {
                const int packed = ~next;
                ti = packed & kLeafBaseMask;
                te = ti + (packed >> kLeafCountShift);
            } else {
                cur = next;
                ti = te = 0;
            }
        }
    }
    return clear;
}

// ---------------------------------------------------------------------------------------------
// per-thread traversal of the 4-ary wide layout (traversal=5; scene_host.cpp build_wide_layout, the records the cooperative
// kernels use): one ray per lane like the ordered mode, but every visit tests the four children of a node - half as many
// dependent node fetches per ray as the pair layout (27 instead of 46 on vader), at about the same number of box tests.
// The four results are sorted with a 5-exchange network, the nearest hit child is entered, the others are pushed far to near.
// Same boxes, leaves, strict culls and tie rule as the ordered pair traversal.
// ---------------------------------------------------------------------------------------------
constexpr int kWideStack = 64; // per-thread stack of traverse_wide4 (every child that is not entered is pushed, leaves too)
PT_DEV void load_wide_child(const float4 *__restrict__ p, float4 &lo, float4 &hi) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                 : "l"(p));
}
#define CPTB_SORT2(ta, ra, tb, rb)                                                                                     \
    {                                                                                                                  \
        const bool sw = tb < ta;                                                                                       \
        const float tt = sw ? tb : ta;                                                                                 \
        const int rr = sw ? rb : ra;                                                                                   \
        tb = sw ? ta : tb, rb = sw ? ra : rb, ta = tt, ra = rr;                                                        \
    }
template <bool FINITE, bool ANYHIT>
PT_DEV Hit traverse_wide4(const DevScene &sc, const float4 *__restrict__ wide, const uint4 *s_cached, V3 o, V3 d, float t_max) {
    Hit h{t_max, 0.f, 0.f, -1, 0};
    const V3 inv_d = rcp3(d), o_div = o * inv_d;
    const int start = cached_entry(s_cached, sc.num_cache, inv_d, o_div, t_max, ANYHIT ? sc.num_nodes : 0);
    if (ANYHIT && start >= sc.num_nodes)
        return h;
    int stack_ref[kWideStack];
    float stack_t[kWideStack];
    int sp = 0;
    int cur = 0; // wide root
    const float inf = __int_as_float(0x7f800000);
    while (true) {
        if (cur >= 0) {
            const float4 *p = wide + cur * 8;
            float4 l0, h0, l1, h1, l2, h2, l3, h3;
            load_wide_child(p, l0, h0), load_wide_child(p + 2, l1, h1), load_wide_child(p + 4, l2, h2), load_wide_child(p + 6, l3, h3);
            float t0, t1, t2, t3;
            const bool b0 = slab_any<FINITE>(l0.x, l0.y, l0.z, h0.x, h0.y, h0.z, inv_d, o_div, t0) && t0 < h.t && __float_as_int(h0.w) > start;
            const bool b1 = slab_any<FINITE>(l1.x, l1.y, l1.z, h1.x, h1.y, h1.z, inv_d, o_div, t1) && t1 < h.t && __float_as_int(h1.w) > start;
            const bool b2 = slab_any<FINITE>(l2.x, l2.y, l2.z, h2.x, h2.y, h2.z, inv_d, o_div, t2) && t2 < h.t && __float_as_int(h2.w) > start;
            const bool b3 = slab_any<FINITE>(l3.x, l3.y, l3.z, h3.x, h3.y, h3.z, inv_d, o_div, t3) && t3 < h.t && __float_as_int(h3.w) > start;
            t0 = b0 ? t0 : inf, t1 = b1 ? t1 : inf, t2 = b2 ? t2 : inf, t3 = b3 ? t3 : inf;
            int r0 = __float_as_int(l0.w), r1 = __float_as_int(l1.w), r2 = __float_as_int(l2.w), r3 = __float_as_int(l3.w);
            const int n = int(b0) + int(b1) + int(b2) + int(b3);
            if (n > 0) {
                // ascending by entry distance (misses carry +inf and sink to the end)
                CPTB_SORT2(t0, r0, t1, r1)
                CPTB_SORT2(t2, r2, t3, r3)
                CPTB_SORT2(t0, r0, t2, r2)
                CPTB_SORT2(t1, r1, t3, r3)
                CPTB_SORT2(t1, r1, t2, r2)
                if (n > 3)
                    stack_ref[sp] = r3, stack_t[sp] = t3, ++sp;
                if (n > 2)
                    stack_ref[sp] = r2, stack_t[sp] = t2, ++sp;
                if (n > 1)
                    stack_ref[sp] = r1, stack_t[sp] = t1, ++sp;
                cur = r0;
                continue;
            }
        } else {
            const int packed = ~cur, base = packed & kLeafBaseMask, cnt = packed >> kLeafCountShift;
            for (int idx = base; idx < base + cnt; idx++) {
                float u, v;
                const float dist = hit_triangle(sc.tris, idx, o, d, u, v, ANYHIT ? t_max : inf);
                if constexpr (ANYHIT) {
                    if (dist > kEps) {
                        h.prim = idx;
                        return h;
                    }
                } else {
                    const bool valid = dist > kEps && (dist < h.t || (dist == h.t && prim_before(sc, idx, h.prim)));
                    h.t = valid ? dist : h.t;
                    h.u = valid ? u : h.u;
                    h.v = valid ? v : h.v;
                    h.prim = valid ? idx : h.prim;
                }
            }
        }
        bool found = false;
        while (sp > 0) {
            --sp;
            if (ANYHIT || stack_t[sp] < h.t) {
                cur = stack_ref[sp];
                found = true;
                break;
            }
        }
        if (!found)
            break;
    }
    return h;
}
#undef CPTB_SORT2
PT_DEV bool ray_is_finite(V3 d) {
    const V3 inv = rcp3(d);
    const float big = 1e28f;
    return fabsf(inv.x) < big && fabsf(inv.y) < big && fabsf(inv.z) < big;
}
PT_DEV Hit closest_hit_wide4(const DevScene &sc, const float4 *wide, const uint4 *s_cached, V3 o, V3 d) {
    return ray_is_finite(d) ? traverse_wide4<true, false>(sc, wide, s_cached, o, d, kMaxDist) : traverse_wide4<false, false>(sc, wide, s_cached, o, d, kMaxDist);
}
PT_DEV bool unoccluded_wide4(const DevScene &sc, const float4 *wide, const uint4 *s_cached, V3 o, V3 d, float max_dist) {
    const Hit h = ray_is_finite(d) ? traverse_wide4<true, true>(sc, wide, s_cached, o, d, max_dist) : traverse_wide4<false, true>(sc, wide, s_cached, o, d, max_dist);
    return h.prim < 0;
}

// MODE 5: per-thread traversal of the 4-ary wide layout (instantiated for k_raygen / k_shadow_apply / k_trace only).
// MODE: 0 ordered (per-thread stack in local memory), 1 reference DFS order, 2 ordered with 256-bit node loads,
// 3 ordered warp-vote scheduler.
// `active` = this lane carries a ray; MODE 3 is warp-synchronous and must be called by all 32 lanes.
template <int MODE> PT_DEV Hit closest_hit(const DevScene &sc, const uint4 *s_cached, V3 o, V3 d, bool active = true) {
    if constexpr (MODE == 3)
        return closest_hit_vote(sc, s_cached, o, d, active);
    Hit miss{kMaxDist, 0.f, 0.f, -1, 0};
    if (!active)
        return miss;
    if constexpr (MODE == 0)
        return closest_hit_ordered<false, false>(sc, s_cached, o, d);
    else if constexpr (MODE == 1)
        return closest_hit_dfs(sc, s_cached, o, d);
    else if constexpr (MODE == 5)
        return closest_hit_wide4(sc, sc.wide, s_cached, o, d);
    else
    {
        const V3 inv = rcp3(d);
        const float big = 1e28f; // |1/d| below this keeps every slab product and difference finite for scene coordinates < 1e8
        if (fabsf(inv.x) < big && fabsf(inv.y) < big && fabsf(inv.z) < big)
            return closest_hit_ordered<true, true>(sc, s_cached, o, d);
        return closest_hit_ordered<true, false>(sc, s_cached, o, d);
    }
}
// closest hit closer than `t_max` (the volumetric path's transmittance march, megakernel_vpt.cu:104-160); per-lane call
template <int MODE> PT_DEV Hit closest_hit_bounded(const DevScene &sc, const uint4 *s_cached, V3 o, V3 d, float t_max) {
    if constexpr (MODE == 1)
        return closest_hit_dfs(sc, s_cached, o, d, t_max);
    else if constexpr (MODE == 0)
        return closest_hit_ordered<false, false>(sc, s_cached, o, d, t_max);
    else
        return closest_hit_ordered<true, false>(sc, s_cached, o, d, t_max);
}
template <int MODE> PT_DEV bool unoccluded(const DevScene &sc, const uint4 *s_cached, V3 o, V3 d, float max_dist, bool active = true) {
    if constexpr (MODE == 3)
        return unoccluded_vote(sc, s_cached, o, d, max_dist, active);
    if (!active)
        return true;
    if constexpr (MODE == 0)
        return unoccluded_ordered<false, false>(sc, s_cached, o, d, max_dist);
    else if constexpr (MODE == 1)
        return unoccluded_dfs(sc, s_cached, o, d, max_dist);
    else if constexpr (MODE == 5)
        return unoccluded_wide4(sc, sc.wide, s_cached, o, d, max_dist);
    else
    {
        const V3 inv = rcp3(d);
        const float big = 1e28f;
        if (fabsf(inv.x) < big && fabsf(inv.y) < big && fabsf(inv.z) < big)
            return unoccluded_ordered<true, true>(sc, s_cached, o, d, max_dist);
        return unoccluded_ordered<true, false>(sc, s_cached, o, d, max_dist);
    }
}

} // namespace cptb
