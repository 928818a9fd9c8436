// traverse_coop.cuh — warp-cooperative traversal of the wide layout (traversal=4), sm_100a.
//
// Why (profiles/r01_ncu_summary.md): the per-thread traversal kernels issue at 6.7 of 32 active lanes
// and keep the L1 data pipe 66 % busy, because every lane fetches its own 64 B node (two wavefronts per lane and visit)
// and because a warp pays for an inner step, a leaf step and a pop in almost every iteration, each for a few lanes.
// Here a ray belongs to a GROUP of G lanes (G = 4 or 8, 32 / G rays in flight per warp):
//
//   * node visit: lane i of the group loads child i of the wide node (one coalesced 256-bit load per lane, G * 32 B =
//     one or two cache lines per visit instead of two wavefronts per lane) and runs the reference's slab test on it
//     (src/core/aabb.cuh:58-65, same fp32 boxes as the reference's nodes; scene_host.cpp build_wide_layout);
//   * the nearest hit inner child is entered (one REDUX.MIN over the group), the others go to the group's stack in
//     shared memory (LIFO, so its depth is bounded by WideLayout::max_stack); a popped entry that is no longer in front
//     of the closest hit is dropped;
//   * hit LEAVES are not tested where they are found: they go to the group's pending list, and the warp runs a
//     "triangle phase" - lane i of every group tests pending leaf i (src/core/primitives.cuh:57-87) - only when enough
//     lanes would be busy or a group cannot go on without it.  Deferral delays updates of the closest distance, i.e.
//     it weakens culling a little (+3 % visits on vader), never the result;
//   * a group that finishes its ray takes the next ray of the warp's 32-ray chunk; chunk prologue (ray fetch, the fp16
//     cached phase of src/renderer/tracing_func.cuh:124-142) and epilogue (interaction, Russian roulette, compaction)
//     run with one ray per lane.
//
// The cached fp16 nodes and the top of the wide tree (the "treelet slab": nodes are numbered breadth-first) are staged
// into shared memory once per CTA with TMA bulk copies (cp.async.bulk + mbarrier; SASS: UBLKCP / SYNCS).
//
// Hit set: identical to the ordered pair traversal (same boxes, same leaves, same strict culls, same tie rule: among
// equal distances the smaller reference primitive index wins) - tools/sim/trav_sim.cpp checks it on the CPU,
// tests/test_gpu.py on the GPU against the reference's frames.
#pragma once
#include "traverse.cuh"

#include <cstdint>

namespace cptb {

struct CoopParams {
    const float4 *wide; // 2 * G float4 per node, node 0 = root
    int slab_nodes;     // nodes [0, slab_nodes) are resident in shared memory (kernels instantiated with SLAB only)
    int stack_cap;      // stack entries per group
};

// shared-memory carve-up of one CTA (bytes; every region 16 B aligned)
struct CoopLayout {
    int cached_off, slab_off, warp_off, warp_bytes, total;
};
__host__ __device__ inline CoopLayout coop_layout(int num_cache, int slab_nodes, int G, int stack_cap, int warps) {
    CoopLayout L;
    L.cached_off = 16; // [0,16): mbarrier
    L.slab_off = L.cached_off + ((num_cache * 16 + 127) / 128) * 128;
    L.warp_off = L.slab_off + slab_nodes * G * 32;
    const int ng = 32 / G;
    // rays 32 x 64 B, hits 32 x 16 B, next-ray counter 16 B, pending lists ng x 2G x 4 B, stacks ng x cap x 8 B
    L.warp_bytes = 2048 + 512 + 16 + ng * 2 * G * 4 + ng * stack_cap * 8;
    L.warp_bytes = (L.warp_bytes + 15) / 16 * 16;
    L.total = L.warp_off + warps * L.warp_bytes;
    return L;
}

struct CoopShared {
    const uint4 *cached;
    const float4 *slab;
    float4 *rays; // 4 per ray: {o.xyz, bits(start)} {d.xyz, t_max (<= 0: no ray)} {1/d.xyz, -} {o/d.xyz, -}
    float4 *hits; // closest: {t, u, v, bits(prim)}; any-hit: x = 1 visible / 0 occluded
    int *next;
    int *pend;
    uint2 *stack;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

// Stage the cached nodes and the treelet slab with TMA bulk copies, then carve the per-warp regions.
// Must be called by every thread of the CTA, once, before anything else touches dynamic shared memory.
__device__ __forceinline__ CoopShared coop_setup(unsigned char *smem, const DevScene &sc, const CoopParams &cp, int G, bool with_slab) {
    const int slab_nodes = with_slab ? cp.slab_nodes : 0;
    const CoopLayout L = coop_layout(sc.num_cache, slab_nodes, G, cp.stack_cap, int(blockDim.x >> 5));
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    const uint32_t bar_a = smem_addr(bar);
    const uint32_t cached_bytes = uint32_t(sc.num_cache) * 16u, slab_bytes = uint32_t(slab_nodes) * uint32_t(G) * 32u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(cached_bytes + slab_bytes) : "memory");
        if (cached_bytes)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(smem + L.cached_off)),
                         "l"(sc.cache_nodes), "r"(cached_bytes), "r"(bar_a)
                         : "memory");
        if (slab_bytes)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(smem + L.slab_off)),
                         "l"(cp.wide), "r"(slab_bytes), "r"(bar_a)
                         : "memory");
    }
    // every thread waits for the bytes to land (phase 0)
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(bar_a), "r"(0u)
                         : "memory");
        }
    }
    CoopShared S;
    S.cached = reinterpret_cast<const uint4 *>(smem + L.cached_off);
    S.slab = reinterpret_cast<const float4 *>(smem + L.slab_off);
    unsigned char *w = smem + L.warp_off + (threadIdx.x >> 5) * L.warp_bytes;
    const int ng = 32 / G;
    S.rays = reinterpret_cast<float4 *>(w);
    S.hits = reinterpret_cast<float4 *>(w + 2048);
    S.next = reinterpret_cast<int *>(w + 2560);
    S.pend = reinterpret_cast<int *>(w + 2576);
    S.stack = reinterpret_cast<uint2 *>(w + 2576 + ng * 2 * G * 4);
    return S;
}

__device__ __forceinline__ void load_child(const float4 *__restrict__ p, float4 &lo, float4 &hi) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                 : "l"(p));
}

// rays the cooperative traversal takes: every component of 1/d stays below 1e28 (see closest_hit<2>)
__device__ __forceinline__ bool coop_ray_ok(V3 inv) {
    const float big = 1e28f;
    return fabsf(inv.x) < big && fabsf(inv.y) < big && fabsf(inv.z) < big;
}

// chunk prologue, one ray per lane: the ray record the lane groups read (t_max <= 0: no ray in this slot)
__device__ __forceinline__ void coop_store_ray(const CoopShared &S, int lane, V3 o, V3 d, V3 inv_d, V3 o_div, int start, float t_max) {
    S.rays[4 * lane] = make_float4(o.x, o.y, o.z, __int_as_float(start));
    S.rays[4 * lane + 1] = make_float4(d.x, d.y, d.z, t_max);
    S.rays[4 * lane + 2] = make_float4(inv_d.x, inv_d.y, inv_d.z, 0.f);
    S.rays[4 * lane + 3] = make_float4(o_div.x, o_div.y, o_div.z, 0.f);
}

// minimum over the lanes of a group, every lane gets it.  Butterfly of full-mask shuffles: the exchange partners stay
// inside the aligned group of G lanes.  (redux.sync would be one instruction, but its result lands in a uniform register:
// with one member mask per group ptxas serialises the groups - MATCH.ANY + one CREDUX pass per distinct mask.)
template <int G> __device__ __forceinline__ unsigned group_min(unsigned v) {
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 2));
    if constexpr (G == 8)
        v = min(v, __shfl_xor_sync(0xffffffffu, v, 4));
    return v;
}

// Traverse the rays of the warp's chunk (S.rays) and leave the results in S.hits.
// Warp-synchronous: all 32 lanes call it together; S.rays / S.hits must be written, *S.next = 0, and __syncwarp()ed before.
// Control flow is warp-uniform wherever lanes talk to each other (ballots, shuffles use the full mask); what differs
// between the groups of a warp is predicated.  One loop iteration = one node visit per group; everything else (triangle
// phase, retiring a ray, taking the next) sits in a "service" block the warp enters only when some group needs it.
template <int G, bool ANYHIT, bool SLAB, int UNROLL = 2> __device__ __forceinline__ void coop_traverse(const DevScene &sc, const CoopParams &cp, const CoopShared &S, int lane) {
    static_assert(G == 4 || G == 8, "lane group of 4 or 8");
    constexpr unsigned GM = (1u << G) - 1u, FULL = 0xffffffffu;
    const int li = lane & (G - 1), gbase = lane & ~(G - 1);
    const unsigned lt = (1u << li) - 1u;
    int *const pend = S.pend + (lane / G) * 2 * G;
    uint2 *const stack = S.stack + (lane / G) * cp.stack_cap;
    const float4 *const gchild = cp.wide + li * 2;
    const float4 *const schild = S.slab + li * 2;
    const float inf = __int_as_float(0x7f800000);

    int rcur = -1, cur = -1, sp = 0, npend = 0, start = 0, bprim = -1;
    V3 inv_d{0, 0, 0}, o_div{0, 0, 0};
    float ht = 0.f, bt = 0.f, bu = 0.f, bv = 0.f;
    int more = 1; // the chunk may still hold rays for this group

    auto take = [&](bool want) { // warp-uniform call; groups with `want` take the next ray of the chunk (or go idle)
        int r = 32;
        if (want && more && li == 0)
            r = atomicAdd(S.next, 1);
        r = __shfl_sync(FULL, r, gbase);
        if (want) {
            rcur = -1, cur = -1, sp = 0, npend = 0;
            if (r >= 32) {
                more = 0;
            } else {
                rcur = r;
                const float4 a = S.rays[4 * r], b = S.rays[4 * r + 1], c = S.rays[4 * r + 2], e = S.rays[4 * r + 3];
                start = __float_as_int(a.w);
                inv_d = xyz(c), o_div = xyz(e);
                ht = bt = b.w;
                bu = bv = 0.f, bprim = -1;
                cur = b.w > 0.f ? 0 : -1; // no ray in this slot: retired by the next service block
                if (ANYHIT && start >= sc.num_nodes) // no cached frontier node in front of the ray: unoccluded (tracing_func.cuh:49)
                    cur = -1;
            }
        }
    };
    take(true);

    while (true) {
        __syncwarp();
        // ---- node visits (groups without a node to visit run along with every test predicated off).  UNROLL visits per look at
        // the service conditions: a group that gets blocked or finishes in between idles for at most UNROLL - 1 visits
#pragma unroll
        for (int rep = 0; rep < UNROLL; rep++) {
            const bool walk = cur >= 0 && npend <= G;
            const int node = cur > 0 ? cur : 0;
            float4 lo, hi;
            if (SLAB && node < cp.slab_nodes)
                lo = schild[node * (2 * G)], hi = schild[node * (2 * G) + 1];
            else
                load_child(gchild + node * (2 * G), lo, hi);
            float t;
            // min/max form of the slab test: value-identical to the reference's compare+select form for rays whose 1/d is finite
            // (traverse.cuh slab_test_finite); the kernels route the other rays through the per-thread traversal in their prologue
            const bool hit = slab_test_finite(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, inv_d, o_div, t) && t < ht && __float_as_int(hi.w) > start && walk;
            const int ref = __float_as_int(lo.w);
            const bool hleaf = hit && ref < 0, hinner = hit && ref >= 0;
            const unsigned lm = (__ballot_sync(FULL, hleaf) >> gbase) & GM;
            const unsigned im = (__ballot_sync(FULL, hinner) >> gbase) & GM;
            if (hleaf)
                pend[npend + __popc(lm & lt)] = ref;
            npend += __popc(lm);
            int w;
            if constexpr (ANYHIT) {
                w = __ffs(im) - 1;
            } else {
                const unsigned key = hinner ? ((__float_as_uint(fmaxf(t, 0.f)) & ~GM) | unsigned(li)) : 0xffffffffu;
                w = int(group_min<G>(key) & GM);
            }
            const int next = __shfl_sync(FULL, ref, gbase + (w & (G - 1)));
            if (walk) {
                if (im) {
                    const unsigned others = im & ~(1u << w);
                    if ((others >> li) & 1u)
                        stack[sp + __popc(others & lt)] = make_uint2(uint32_t(ref), __float_as_uint(t));
                    sp += __popc(others);
                    cur = next;
                } else {
                    cur = -1;
                    while (sp > 0) { // LIFO; every lane of the group reads the same entry (broadcast)
                        const uint2 e = stack[--sp];
                        if (ANYHIT || __uint_as_float(e.y) < ht) {
                            cur = int(e.x);
                            break;
                        }
                    }
                }
            }
            if (rep + 1 < UNROLL)
                __syncwarp();
        }
        // ---- service: a group that cannot go on without a triangle phase (pending list above G, or traversal finished with
        // leaves pending), or that finished its ray
        const bool blocked = npend > 0 && !(cur >= 0 && npend <= G);
        if (!__any_sync(FULL, blocked || (rcur >= 0 && cur < 0)))
            continue;
        __syncwarp();
        if (__any_sync(FULL, blocked)) { // triangle phase: every group tests its pending leaves (up to G of them, lane i leaf i)
            const int nb = npend < G ? npend : G;
            bool occluded = false;
            if (li < nb) {
                const float4 a = S.rays[4 * rcur], b = S.rays[4 * rcur + 1];
                const V3 o = xyz(a), d = xyz(b);
                const int packed = ~pend[npend - 1 - li], base = packed & kLeafBaseMask, cnt = packed >> kLeafCountShift;
                for (int idx = base; idx < base + cnt; idx++) {
                    float u, v;
                    const float dist = hit_triangle(sc.tris, idx, o, d, u, v, ANYHIT ? ht : inf);
                    if constexpr (ANYHIT) {
                        occluded = occluded || dist > kEps;
                    } else {
                        const bool better = dist > kEps && (dist < bt || (dist == bt && prim_before(sc, idx, bprim)));
                        bt = better ? dist : bt;
                        bu = better ? u : bu;
                        bv = better ? v : bv;
                        bprim = better ? idx : bprim;
                    }
                }
            }
            npend -= nb;
            if constexpr (ANYHIT) {
                if ((__ballot_sync(FULL, occluded) >> gbase) & GM) {
                    if (li == 0)
                        S.hits[rcur].x = 0.f;
                    cur = -1, npend = 0, sp = 0;
                }
            } else {
                ht = __uint_as_float(group_min<G>(__float_as_uint(bt)));
            }
        }
        const bool fin = rcur >= 0 && cur < 0 && npend == 0;
        if (__any_sync(FULL, fin)) { // retire finished rays, take the next ray of the chunk
            if constexpr (!ANYHIT) {
                const bool cand = fin && bprim >= 0 && bt == ht;
                const unsigned cm = (__ballot_sync(FULL, cand) >> gbase) & GM;
                bool win = cand;
                if (cand && __popc(cm) > 1) { // equal distances in two lanes (rare): the reference keeps the first hit in its DFS order
                    const int key = sc.prim_orig ? __ldg(sc.prim_orig + bprim) : bprim;
                    win = key == __reduce_min_sync(cm << gbase, key);
                }
                if (win)
                    S.hits[rcur] = make_float4(bt, bu, bv, __int_as_float(bprim));
            }
            take(fin);
            if (__ballot_sync(FULL, rcur >= 0) == 0u)
                break;
        }
    }
    __syncwarp();
}

} // namespace cptb
