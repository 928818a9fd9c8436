// wavefront.cu — the per-bounce wavefront path-tracing hot path as sm_100a kernels.
//
// Pipeline for one batch of S samples over P = W*H pixels (slots = S*P paths in flight):
//
//   k_raygen      camera ray + primary closest hit, payload init, hit paths -> queue[0]
//   for bounce b: k_shade  (NEE shadow ray + MIS + emission + BSDF sample)  over queue[b]
//                 k_trace  (closest hit + interaction + miss/envmap + Russian roulette),
//                          survivors are compacted into queue[b+1] with warp ballots
//   k_splat       NaN scrub, ordered per-sample accumulation, running mean, gamma, Welford variance
//
// This replaces the reference's raygen_primary_hit_shader / fused_ray_bounce_shader /
// fused_closesthit_shader / radiance_splat (src/pt_impl/wavefront_pt.cu:71-314,476-503) and its
// host loop with thrust::sort + lower_bound + D2H count per bounce
// (src/pt_impl/wf_path_tracer.cu:175-230).  Differences that matter on B200:
//   * no host synchronisation inside a batch: queue sizes stay in device memory, every kernel is a
//     persistent grid (k x 148 SMs) whose warps pull 32-path chunks with one atomic per chunk;
//   * compaction is fused into k_trace's epilogue (ballot + popc prefix, one atomic per warp);
//     the image is indexed by path slot, so processing order never changes pixel values;
//   * S samples are in flight at once (180 GB HBM3e) so late bounces still fill 148 SMs;
//   * near-child-first traversal over a re-laid-out copy of the SAME tree (traverse.cuh).
#include "traverse.cuh"
#include "traverse_coop.cuh"
#include "volume.cuh"
#include "wavefront.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <stdexcept>

namespace cptb {

#define CUDA_OK(expr)                                                                                                  \
    do {                                                                                                               \
        cudaError_t err__ = (expr);                                                                                    \
        if (err__ != cudaSuccess) {                                                                                    \
            cudaGetLastError(); /* clear the non-sticky error so that later calls do not re-report it */              \
            char buf__[512];                                                                                           \
            snprintf(buf__, sizeof(buf__), "CUDA error %s at %s:%d (%s)", cudaGetErrorString(err__), __FILE__,         \
                     __LINE__, #expr);                                                                                 \
            throw std::runtime_error(buf__);                                                                           \
        }                                                                                                              \
    } while (0)

constexpr int kBlock = 128;
#ifndef VPT_MIN_BLOCKS
#define VPT_MIN_BLOCKS 8 // register budget of the volumetric megakernel: 8 CTAs/SM (64 registers) measured best of 1/4/6/8
#endif
#ifndef VPTW_SHADE_MIN_BLOCKS
#define VPTW_SHADE_MIN_BLOCKS 6 // volumetric wavefront, traversal-free kernels (k_vpt_pre / k_vpt_post)
#endif
#ifndef VPTW_MARCH_MIN_BLOCKS
#define VPTW_MARCH_MIN_BLOCKS 10 // volumetric wavefront, transmittance march: 48 registers (fog scene: march 6.04 -> 5.60 ms; 8: 6.00, 12: 5.67)
#endif
#ifndef SHADE_MIN_BLOCKS
#define SHADE_MIN_BLOCKS 6 // register budget of the traversal-free shading kernels: 6 CTAs/SM -> <= 85 registers
#endif

struct PathState {
    float4 *ray_o; // o.xyz, hit_t
    float4 *ray_d; // d.xyz, tag bits
    float4 *thp;
    float4 *L;
    float4 *isect; // n.xyz, half2 uv
    uint2 *rng;
    float *pdf;
    uint32_t *lobes; // megakernel-rules mode: diffuse | specular << 8 | transmit << 16 (| volume << 24) bounce counters
    uint2 *vstate;   // volumetric wavefront: x = nested-volume BankStack, y = emitter id | medium-event flag << 8
};

struct BatchParams {
    int W, H, P, S, first_accum, seed_offset;
    int row0, row1; // rows [row0,row1) of the frame are rendered (tile split across GPUs); pixel indices stay global
    int rules; // 0: wavefront termination rules, 1: megakernel ("pt") rules, 2: volumetric ("vpt")
    int max_depth, max_diffuse, max_specular, max_transmit, max_volume; // MaxDepthParams (volumetric kernel)
};

// termination parameters of the megakernel rules (MaxDepthParams, src/core/max_depth.h:26-42)
struct PtRules {
    int enabled, bounce, max_diffuse, max_specular, max_transmit;
};

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_cached_nodes(const DevScene &sc, uint4 *s_cached) {
    for (int i = threadIdx.x; i < sc.num_cache; i += blockDim.x)
        s_cached[i] = sc.cache_nodes[i];
    __syncthreads();
}

// one atomic per warp chunk; returns the chunk index
__device__ __forceinline__ int warp_next_chunk(int *counter, int lane) {
    int v = 0;
    if (lane == 0)
        v = atomicAdd(counter, 1);
    return __shfl_sync(0xffffffffu, v, 0);
}

// warp-aggregated append (ballot + popc prefix); must be called by the full warp
__device__ __forceinline__ void warp_push(uint32_t *queue, int *count, bool pred, uint32_t value, int lane) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0)
        return;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader)
        base = atomicAdd(count, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred)
        queue[base + __popc(mask & ((1u << lane) - 1u))] = value;
}

// warp-aggregated allocation of one output index per predicated lane; must be called by the full warp
__device__ __forceinline__ int warp_alloc(int *count, bool pred, int lane) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0)
        return -1;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader)
        base = atomicAdd(count, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return pred ? base + __popc(mask & ((1u << lane) - 1u)) : -1;
}

__device__ __forceinline__ V4 null_or_env(const DevScene &sc, V3 dir) {
    // c_emitter[envmap_id]->eval_le(&ray.d): envmap_id == 0 selects the null emitter -> (0,0,0,1)
    if (sc.envmap_id > 0)
        return emitter_eval(sc.emitters[sc.envmap_id], dir, make_float4(0, 0, 0, 0), false);
    return mk4(0, 0, 0, 1);
}

// CompactedObjInfo::solid_angle_pdf (object.cuh:51-56): the quotient is formed in double (1e-4 is a double
// literal) and rounded to float by the return type
__device__ __forceinline__ float solid_angle_pdf(float inv_area, V3 normal, V3 incid, float min_depth) {
    return inv_area * min_depth * min_depth / max(fabsf(dot(incid, normal)), 1e-4);
}

// ------------------------------------------------------------------------------------------------
// k_raygen — reference: raygen_primary_hit_shader (wavefront_pt.cu:71-133)
// one warp = one 8x4 pixel tile of one sample
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kBlock)
    k_raygen(DevScene sc, CameraPOD cam, PathState st, BatchParams bp, uint32_t *q_out, int *q_count, int *work) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    const int lane = threadIdx.x & 31;
    const int ty0 = bp.row0 >> 2;
    const int tiles_x = (bp.W + 7) >> 3, tiles_y = ((bp.row1 + 3) >> 2) - ty0, tiles = tiles_x * tiles_y;
    const int total = tiles * bp.S;
    for (int task = warp_next_chunk(work, lane); task < total; task = warp_next_chunk(work, lane)) {
        const int s = task / tiles, tile = task - s * tiles;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int px = (tx << 3) + (lane & 7), py = ((ty + ty0) << 2) + (lane >> 3);
        const bool inside = px < bp.W && py >= bp.row0 && py < bp.row1;
        bool hit_any = false;
        uint32_t slot = 0;
        V3 o{0, 0, 0}, d{0, 0, 1};
        if (inside) {
            const int gidx = px + py * bp.W;
            slot = uint32_t(s) * uint32_t(bp.P) + uint32_t(gidx);
            Rng rng;
            rng.seed(gidx, bp.seed_offset + (bp.first_accum + s) * kSeedScaler);
            camera_ray(cam, px, py, rng, o, d);
            st.thp[slot] = make_float4(1, 1, 1, 1);
            st.pdf[slot] = 1.f;
            st.rng[slot] = make_uint2(rng.d0, rng.d1);
            if (bp.rules)
                st.lobes[slot] = 0u;
            if (bp.rules == 2)
                st.vstate[slot] = make_uint2(BankStack::make(sc.cam_vol).bits, 0u);
        }
        const Hit h = closest_hit<MODE>(sc, s_cached, o, d, inside); // all 32 lanes call (warp-synchronous modes)
        if (inside) {
            uint32_t tag = 0;
            V4 L = mk4(0, 0, 0, 1);
            if (h.prim >= 0) {
                tag = TAG_ACTIVE | TAG_HIT | uint32_t(h.prim);
                st.isect[slot] = make_interaction(sc, h.prim, h.u, h.v);
                hit_any = true;
            } else if (bp.rules) {
                L += mk4(1, 1, 1, 1) * null_or_env(sc, d); // megakernel_pt.cu:181-184: radiance(0,0,0,1) += throughput * Le
            } else {
                L = null_or_env(sc, d);
            }
            st.L[slot] = to_f4(L);
            st.ray_o[slot] = make_float4(o.x, o.y, o.z, h.t);
            st.ray_d[slot] = make_float4(d.x, d.y, d.z, __uint_as_float(tag));
        }
        warp_push(q_out, q_count, hit_any, slot, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// NEE set-up shared by k_shade and k_shadow: object -> (material, emitter), emitter + primitive pick, emitter
// sample, shadow ray (wavefront_pt.cu:237-272).  Consumes 1 + 1 + 2 draws of the path's sampler.
// ------------------------------------------------------------------------------------------------
struct NeeSetup {
    uint32_t emit_id;
    int material_id, emitter_id;
    ObjectPOD ob;
    V3 hit_pos, sd;    // shadow ray origin / unit direction
    float len;         // distance to the sampled point
    float direct_pdf;
    V4 direct_comp;    // sampled emitter radiance
};
__device__ __forceinline__ void nee_setup(const DevScene &sc, V3 o, V3 d, float hit_t, uint32_t tag, float4 it, Rng &rng,
                                          NeeSetup &ns) {
    const int object_id = __ldg(sc.obj_idxs + (tag & TAG_INDEX)) & 0x000fffff; // extract_object_info (tracing_func.cuh:38-41)
    ns.ob = sc.objects[object_id];
    ns.material_id = ns.ob.bsdf_emitter >> 16;
    ns.emitter_id = ((ns.ob.bsdf_emitter & 0xFFFF) ^ 0x8000) - 0x8000;

    // ---- pick an emitter (tracing_func.cuh:210-225)
    int num = sc.num_emitters;
    num -= ns.emitter_id > 0 && num > 1;
    uint32_t emit_id = (rng.next_u32() % uint32_t(num)) + 1;
    emit_id += emit_id >= uint32_t(ns.emitter_id) && ns.emitter_id > 0;
    ns.direct_pdf = 1.f / float(num);
    emit_id = emit_id * uint32_t(ns.emitter_id == 0 || num > 1);
    ns.emit_id = emit_id;
    const EmitterPOD &em = sc.emitters[emit_id];
    // ---- pick a primitive on it (object.cuh:45-49, wavefront_pt.cu:254-259)
    const ObjectPOD eo = sc.objects[em.obj_ref * (em.obj_ref >= 0)];
    ns.direct_pdf *= eo.inv_area;
    int emit_prim = int(rng.next_u32() % uint32_t(eo.prim_num)) + eo.prim_offset;
    emit_prim = __ldg(sc.emitter_prims + emit_prim);

    ns.hit_pos = advance(o, d, hit_t);
    ns.direct_comp = mk4(0, 0, 0, 1);
    ns.sd = emitter_sample(sc, em, ns.hit_pos, xyz(it), ns.direct_comp, ns.direct_pdf, rng.next2d(), emit_prim) - ns.hit_pos;
    ns.len = length(ns.sd);
    ns.sd *= __frcp_rn(ns.len);
}

// ------------------------------------------------------------------------------------------------
// k_shadow — the NEE visibility ray of fused_ray_bounce_shader (wavefront_pt.cu:273-275) as its own
// low-register traversal kernel: re-derives the shadow ray from the path's sampler state (which it does not
// advance) and leaves the answer in bit 31 of the ray tag for k_shade<SPLIT>
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_shadow(DevScene sc, PathState st, const uint32_t *__restrict__ queue,
                                                   const int *__restrict__ q_count, int *work,
                                                   unsigned long long *shadow_counter) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    const int lane = threadIdx.x & 31;
    const int count = *q_count;
    unsigned int local_shadow = 0;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int i = chunk * 32 + lane;
        const bool valid = i < count;
        uint32_t slot = 0, tag = 0;
        bool need_shadow = false;
        V3 so{0, 0, 0}, sd{0, 0, 1};
        float range = 0;
        if (valid) {
            slot = queue[i];
            const float4 ro = st.ray_o[slot], rd = st.ray_d[slot];
            const uint2 rs = st.rng[slot];
            Rng rng{rs.x, rs.y};
            tag = __float_as_uint(rd.w);
            NeeSetup ns;
            nee_setup(sc, xyz(ro), xyz(rd), ro.w, tag, st.isect[slot], rng, ns);
            need_shadow = ns.emit_id != 0;
            so = ns.hit_pos, sd = ns.sd, range = ns.len - kEps;
        }
        local_shadow += need_shadow;
        const bool visible = unoccluded<MODE>(sc, s_cached, so, sd, range, need_shadow);
        if (valid)
            st.ray_d[slot].w = __uint_as_float(visible && need_shadow ? (tag | TAG_NEE_VISIBLE) : (tag & ~TAG_NEE_VISIBLE));
    }
    for (int off = 16; off > 0; off >>= 1)
        local_shadow += __shfl_down_sync(0xffffffffu, local_shadow, off);
    if (lane == 0 && local_shadow)
        atomicAdd(shadow_counter, (unsigned long long)local_shadow);
}

// ------------------------------------------------------------------------------------------------
// Dynamic-fetch pipeline ("pipeline" 2).  Traversal is the whole cost of this path and a warp that walks 32
// rays to completion idles most of its lanes on the longest ray, so traversal runs in two pure kernels
// (k_anyhit, k_intersect) whose warps keep their lanes full: whenever kRefill lanes are idle they pull the
// next rays of the queue.  Everything that is not traversal moves into coherent companion kernels:
//   k_shadow_gen  NEE set-up -> shadow ray {o, range}{d, start} + queue of rays that need a traversal
//   k_anyhit      any-hit traversal, sets TAG_NEE_VISIBLE
//   k_shade<-1>   shading (reads the bit)
//   k_intersect   closest-hit traversal -> hit record {t, u, v, prim}
//   k_resolve     interaction, miss/envmap, Russian roulette, compaction (the epilogue of k_trace)
// Same boxes, same slab / triangle arithmetic, same tie rule as the ordered traversal of traverse.cuh.
// ------------------------------------------------------------------------------------------------
constexpr int kGrab = 128;  // rays a warp reserves per global atomic
// kRefill (refill a warp when at least this many lanes are idle) and kSteps (traversal steps between refill
// checks) are launch parameters: env CPTB200_DYN_REFILL (default 8) / CPTB200_DYN_STEPS (default 4)

constexpr unsigned char kVisOddRay = 2; // stream traversal: shadow ray left to k_shadow_finish (see stream_fetch_stage)
constexpr int kPrimOddRay = -2;          // same for an extension ray: hit record left to k_resolve
struct RayBufs {
    float4 *sh_o; // shadow ray origin, w = range
    float4 *sh_d; // shadow ray direction, w = bits(start DFS index of the cached phase)
    float4 *hit;  // closest hit {t, u, v, bits(prim)}
};

__global__ void __launch_bounds__(kBlock) k_shadow_gen(DevScene sc, PathState st, RayBufs rb, const uint32_t *__restrict__ queue,
                                                       const int *__restrict__ q_count, int *work, uint32_t *sq,
                                                       int *sq_count, unsigned long long *shadow_counter) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    const int lane = threadIdx.x & 31;
    const int count = *q_count;
    unsigned int local_shadow = 0;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int i = chunk * 32 + lane;
        uint32_t slot = 0;
        bool traverse = false;
        if (i < count) {
            slot = queue[i];
            const float4 ro = st.ray_o[slot], rd = st.ray_d[slot];
            const uint2 rs = st.rng[slot];
            Rng rng{rs.x, rs.y};
            uint32_t tag = __float_as_uint(rd.w) & ~TAG_NEE_VISIBLE;
            NeeSetup ns;
            nee_setup(sc, xyz(ro), xyz(rd), ro.w, tag, st.isect[slot], rng, ns);
            if (ns.emit_id != 0) {
                ++local_shadow;
                const float range = ns.len - kEps;
                const V3 inv_d = rcp3(ns.sd), o_div = ns.hit_pos * inv_d;
                const int start = cached_entry(s_cached, sc.num_cache, inv_d, o_div, range, sc.num_nodes);
                if (start >= sc.num_nodes) {
                    tag |= TAG_NEE_VISIBLE; // no cached frontier node on the way: unoccluded (tracing_func.cuh:49)
                } else {
                    traverse = true;
                    rb.sh_o[slot] = make_float4(ns.hit_pos.x, ns.hit_pos.y, ns.hit_pos.z, range);
                    rb.sh_d[slot] = make_float4(ns.sd.x, ns.sd.y, ns.sd.z, __int_as_float(start));
                }
            }
            st.ray_d[slot].w = __uint_as_float(tag);
        }
        warp_push(sq, sq_count, traverse, slot, lane);
    }
    for (int off = 16; off > 0; off >>= 1)
        local_shadow += __shfl_down_sync(0xffffffffu, local_shadow, off);
    if (lane == 0 && local_shadow)
        atomicAdd(shadow_counter, (unsigned long long)local_shadow);
}

// warp-level work distribution for the dynamic-fetch kernels: the warp owns [next, end) of the queue and
// reserves kGrab more entries with one atomic when it runs dry.  Returns how many entries the idle lanes may
// take now (0 when the queue is exhausted) and advances `next`; `first` = index of the first one.
__device__ __forceinline__ int warp_take(int *work, int count, int lane, int n_idle, int &next, int &end, bool &more, int &first) {
    if (next >= end) {
        int b = 0;
        if (lane == 0)
            b = atomicAdd(work, kGrab);
        b = __shfl_sync(0xffffffffu, b, 0);
        next = b;
        end = min(b + kGrab, count);
        if (b >= count) {
            more = false;
            return 0;
        }
    }
    const int take = min(n_idle, end - next);
    first = next;
    next += take;
    return take;
}

constexpr int kSmemTop = 8; // stack entries per thread kept in shared memory ([level][thread]: bank == lane, conflict-free)

// APPLY = false: pipeline 2 (rays from k_shadow_gen carry their cached-phase start index, answer goes to the tag bit);
// APPLY = true: pipeline 4 (rays from k_shade<-2>, cached phase here, an unoccluded ray swaps in the prepared radiance)
template <int MINB, bool APPLY>
__global__ void __launch_bounds__(kBlock, MINB) k_anyhit(DevScene sc, PathState st, RayBufs rb, const uint32_t *__restrict__ sq,
                                                         const int *__restrict__ sq_count, int *work, int kRefill, int kSteps) {
    extern __shared__ uint4 s_cached[];
    if constexpr (APPLY)
        stage_cached_nodes(sc, s_cached);
    int *s_stack = reinterpret_cast<int *>(s_cached + (APPLY ? sc.num_cache : 0)) + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int count = *sq_count;
    int spill[kStackSize];
    int sp = 0, cur = kTraversalDone, start = 0;
    uint32_t slot = 0;
    V3 o{0, 0, 0}, d{0, 0, 1}, inv_d{0, 0, 0}, o_div{0, 0, 0};
    float range = 0;
    bool active = false, more = true;
    int next = 0, end = 0;
    auto push = [&](int ref) {
        if (sp < 2 * kSmemTop)
            s_stack[sp * kBlock] = ref;
        else
            spill[sp - 2 * kSmemTop] = ref;
        ++sp;
    };
    auto pop = [&]() {
        if (sp == 0)
            return kTraversalDone;
        --sp;
        return sp < 2 * kSmemTop ? s_stack[sp * kBlock] : spill[sp - 2 * kSmemTop];
    };
    while (true) {
        const unsigned act = __ballot_sync(0xffffffffu, active);
        const int n_idle = 32 - __popc(act);
        if (more && n_idle >= kRefill) {
            int first = 0;
            const int take = warp_take(work, count, lane, n_idle, next, end, more, first);
            const int rank = __popc(~act & ((1u << lane) - 1u));
            if (!active && rank < take) {
                slot = sq[first + rank];
                const float4 a = rb.sh_o[slot], b = rb.sh_d[slot];
                o = xyz(a), d = xyz(b), range = a.w;
                inv_d = rcp3(d), o_div = o * inv_d;
                if constexpr (APPLY)
                    start = cached_entry(s_cached, sc.num_cache, inv_d, o_div, range, sc.num_nodes);
                else
                    start = __float_as_int(b.w);
                cur = sc.wroot, sp = 0, active = true;
                if (APPLY && start >= sc.num_nodes) { // no cached frontier node on the way: unoccluded (tracing_func.cuh:49)
                    st.L[slot] = rb.hit[slot];
                    active = false;
                }
            }
        } else if (act == 0)
            break;
#pragma unroll 1
        for (int k = 0; k < kSteps; k++) {
            if (active) {
                bool occluded = false;
                if (cur >= 0) {
                    float4 a0, a1, b0, b1;
                    load_pair_node<true>(sc.wnodes, cur, a0, a1, b0, b1);
                    float t0, t1;
                    const bool h0 = slab_test(a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, inv_d, o_div, t0) && t0 < range &&
                                    __float_as_int(a1.w) > start;
                    const bool h1 = slab_test(b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, inv_d, o_div, t1) && t1 < range &&
                                    __float_as_int(b1.w) > start;
                    const int r0 = __float_as_int(a0.w), r1 = __float_as_int(b0.w);
                    if (h0 && h1) {
                        push(r1);
                        cur = r0;
                    } else if (h0 || h1) {
                        cur = h0 ? r0 : r1;
                    } else {
                        cur = pop();
                    }
                } else {
                    const int packed = ~cur, base = packed & kLeafBaseMask, cnt = packed >> kLeafCountShift;
                    for (int idx = base; idx < base + cnt; idx++) {
                        float u, v;
                        if (hit_triangle(sc.tris, idx, o, d, u, v, range) > kEps) {
                            occluded = true;
                            break;
                        }
                    }
                    cur = pop();
                }
                if (occluded || cur == kTraversalDone) {
                    if (!occluded) {
                        if constexpr (APPLY) {
                            st.L[slot] = rb.hit[slot];
                        } else {
                            float *tagp = &st.ray_d[slot].w;
                            *tagp = __uint_as_float(__float_as_uint(*tagp) | TAG_NEE_VISIBLE);
                        }
                    }
                    active = false;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_shade — reference: fused_ray_bounce_shader (wavefront_pt.cu:213-314)
// MODE >= 0: the shadow ray is traced here; MODE == -1 (split pipeline): k_shadow already traced it;
// MODE == -2 (deferred-NEE pipeline): nothing is traced yet - the radiance is written for the occluded case,
// the value for the unoccluded case goes to rb.hit together with the shadow ray, and only paths for which the
// two differ (in any bit) are queued for k_shadow_apply.  NEE samples that cannot change the pixel (emitter
// facing away, BSDF value zero, delta lobes) therefore never cost a traversal, and the result stays exact.
// ------------------------------------------------------------------------------------------------
// DENSE (deferred-NEE pipeline only): the path records themselves are kept compact instead of a queue of slot ids - the
// kernel reads generation `st` at indices [0, count) (entries whose path ended in the last trace are skipped) and writes
// the shaded paths compactly into generation `dn.out`, shadow rays and Lvis at the new index (see k_vptd_post for why)
struct DenseIO {
    PathState out;
    const uint32_t *pix_in;
    uint32_t *pix_out;
    int *n_out;
};
template <int MODE, bool DENSE = false>
__global__ void __launch_bounds__(kBlock, (MODE < 0 ? SHADE_MIN_BLOCKS : 1)) k_shade(DevScene sc, PathState st, const uint32_t *__restrict__ queue,
                                                  const int *__restrict__ q_count, int *work,
                                                  unsigned long long *shadow_counter, int secondary_bounce, PtRules pt,
                                                  RayBufs rb, uint32_t *sq, int *sq_count, DenseIO dn = DenseIO{}) {
    extern __shared__ uint4 s_cached[];
    if constexpr (MODE >= 0)
        stage_cached_nodes(sc, s_cached);
    const int lane = threadIdx.x & 31;
    const int count = *q_count;
    unsigned int local_shadow = 0;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int i = chunk * 32 + lane;
        bool valid = i < count;
        // ---- part 1 (per lane): load the path, pick an emitter + primitive, sample it, set up the shadow ray
        uint32_t slot = 0, tag = 0, emit_id = 0;
        int oslot = 0; // where the path's state goes: its slot, or (DENSE) a fresh compact index of the next generation
        if constexpr (DENSE) {
            slot = uint32_t(i);
            valid = valid && (__float_as_uint(st.ray_d[valid ? i : 0].w) & TAG_ACTIVE) != 0;
            oslot = warp_alloc(dn.n_out, valid, lane);
        }
        PathState &so = DENSE ? dn.out : st;
        V4 thp = mk4(0, 0, 0, 0), rdc = thp, direct_comp = mk4(0, 0, 0, 1);
        V3 o{0, 0, 0}, d{0, 0, 1}, hit_pos{0, 0, 0}, sd{0, 0, 1};
        float4 it = make_float4(0, 0, 1, 0);
        float hit_t = 0, direct_pdf = 1, emit_len_mis = 1;
        Rng rng{0, 0};
        int material_id = 0, emitter_id = 0;
        ObjectPOD ob{0, 0, 1, 0.f};
        if (valid) {
            if constexpr (!DENSE) {
                slot = queue[i];
                oslot = int(slot);
            }
            thp = from_f4(st.thp[slot]);
            rdc = from_f4(st.L[slot]);
            const float4 ro = st.ray_o[slot], rd = st.ray_d[slot];
            const uint2 rs = st.rng[slot];
            rng = Rng{rs.x, rs.y};
            it = st.isect[slot];
            o = xyz(ro), d = xyz(rd);
            hit_t = ro.w;
            tag = __float_as_uint(rd.w);
            NeeSetup ns;
            nee_setup(sc, o, d, hit_t, tag, it, rng, ns);
            emit_id = ns.emit_id, material_id = ns.material_id, emitter_id = ns.emitter_id, ob = ns.ob;
            hit_pos = ns.hit_pos, sd = ns.sd, emit_len_mis = ns.len, direct_pdf = ns.direct_pdf, direct_comp = ns.direct_comp;
            o = hit_pos;
        }
        // ---- part 2 (whole warp): NEE shadow ray (wavefront_pt.cu:273-275)
        const bool need_shadow = valid && emit_id != 0;
        bool visible, queue_shadow = false;
        if constexpr (MODE >= 0) {
            local_shadow += need_shadow;
            visible = unoccluded<MODE>(sc, s_cached, hit_pos, sd, emit_len_mis - kEps, need_shadow);
        } else if constexpr (MODE == -1) {
            visible = (tag & TAG_NEE_VISIBLE) != 0;
            tag &= ~TAG_NEE_VISIBLE;
        } else {
            local_shadow += need_shadow;
            visible = true; // evaluate the unoccluded candidate; k_shadow_apply decides
        }
        const float shadow_range = emit_len_mis - kEps;
        // ---- part 3 (per lane): MIS, emission, BSDF sample, write back
        if (valid) {
            const Material &mat = sc.mats[material_id];
            const EmitterPOD &em = sc.emitters[emit_id];
            float pdf = st.pdf[slot];
            // emission MIS weight against light sampling (wavefront_pt.cu:291-298 == megakernel_pt.cu:96-103, object.cuh:51-56)
            const float emission_weight =
                pdf / (pdf + solid_angle_pdf(ob.inv_area, shading_normal_world(mat, it), d, hit_t) * (emitter_id > 0) *
                                 secondary_bounce * ((tag & TAG_DELTA) == 0));
            // L: radiance if the shadow ray is blocked, Lv: if it is not (or was not needed)
            V4 L, Lv;
            const V4 emis = thp * emitter_eval(sc.emitters[emitter_id], d, it, true);
            if (!pt.enabled) {
                V4 rdc_v = rdc;
                if (need_shadow && visible) { // wavefront_pt.cu:276-289
                    emit_len_mis = direct_pdf + bsdf_pdf(sc, mat, it, sd, d) * float(em.is_sphere != 0);
                    rdc_v += thp * direct_comp * bsdf_eval(sc, mat, it, sd, d) *
                             (float(emit_len_mis > kEps) * __frcp_rn(emit_len_mis < kEps ? 1.f : emit_len_mis));
                }
                // (2) account for emission, and accumulate to payload buffer (wavefront_pt.cu:299-302)
                Lv = emis * emission_weight + rdc_v;
                L = emis * emission_weight + rdc;
            } else {
                // megakernel order (megakernel_pt.cu:105-152): emission first, then the NEE term, then the NaN scrub
                L = rdc;
                L += emis * emission_weight;
                Lv = L;
                if (need_shadow && visible) {
                    emit_len_mis = direct_pdf + bsdf_pdf(sc, mat, it, sd, d) * float(em.is_sphere != 0);
                    Lv += thp * direct_comp * bsdf_eval(sc, mat, it, sd, d) *
                          (float(emit_len_mis > kEps) * __frcp_rn(emit_len_mis < kEps ? 1.f : emit_len_mis));
                }
            }

            // ---- sample the next direction (wavefront_pt.cu:304-312)
            int lobe = LOBE_NONE;
            d = bsdf_sample(sc, mat, d, it, thp, pdf, rng, lobe);
            tag &= ~TAG_DELTA;
            tag |= uint32_t((lobe & LOBE_SPECULAR) > 0) << 30;
            if (pt.enabled) { // megakernel_pt.cu:159-180
                if (bad_rgb(L))
                    L = splat4(0);
                if (bad_rgb(Lv))
                    Lv = splat4(0);
                uint32_t c = st.lobes[slot];
                c += uint32_t((lobe & LOBE_DIFFUSE) > 0) + (uint32_t((lobe & LOBE_SPECULAR) > 0) << 8) +
                     (uint32_t((lobe & LOBE_TRANSMIT) > 0) << 16);
                so.lobes[oslot] = c;
                bool go_on = !(int(c & 0xFF) >= pt.max_diffuse || int((c >> 8) & 0xFF) >= pt.max_specular ||
                               int((c >> 16) & 0xFF) >= pt.max_transmit);
                const float max_value = max3(thp);
                go_on = go_on && !(max_value < kThpEps);
                if (go_on && pt.bounce >= 1 && max_value < kRRThreshold) {
                    if (rng.next1d() > max_value)
                        go_on = false;
                    else
                        thp *= float(1. / max_value);
                }
                if (!go_on)
                    tag &= ~TAG_ACTIVE;
            }
            if constexpr (MODE == -2) {
                const float4 a = to_f4(L), b = to_f4(Lv);
                queue_shadow = need_shadow && (__float_as_uint(a.x) != __float_as_uint(b.x) || __float_as_uint(a.y) != __float_as_uint(b.y) ||
                                               __float_as_uint(a.z) != __float_as_uint(b.z) || __float_as_uint(a.w) != __float_as_uint(b.w));
                if (queue_shadow) {
                    rb.sh_o[oslot] = make_float4(hit_pos.x, hit_pos.y, hit_pos.z, shadow_range);
                    rb.sh_d[oslot] = make_float4(sd.x, sd.y, sd.z, 0.f);
                    rb.hit[oslot] = b;
                }
                so.L[oslot] = a;
            } else {
                so.L[oslot] = to_f4(Lv);
            }

            so.thp[oslot] = to_f4(thp);
            so.rng[oslot] = make_uint2(rng.d0, rng.d1);
            so.ray_o[oslot] = make_float4(o.x, o.y, o.z, hit_t);
            so.ray_d[oslot] = make_float4(d.x, d.y, d.z, __uint_as_float(tag));
            so.pdf[oslot] = pdf;
            if constexpr (DENSE)
                dn.pix_out[oslot] = dn.pix_in[slot];
        }
        if constexpr (MODE == -2)
            warp_push(sq, sq_count, queue_shadow, uint32_t(oslot), lane);
    }
    // per-warp ray statistics (Mrays/s accounting)
    for (int off = 16; off > 0; off >>= 1)
        local_shadow += __shfl_down_sync(0xffffffffu, local_shadow, off);
    if (lane == 0 && local_shadow)
        atomicAdd(shadow_counter, (unsigned long long)local_shadow);
}

// deferred-NEE pipeline: trace the queued shadow rays; an unoccluded ray swaps in the radiance k_shade<-2> prepared
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_shadow_apply(DevScene sc, PathState st, RayBufs rb, const uint32_t *__restrict__ sq,
                                                         const int *__restrict__ sq_count, int *work) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    const int lane = threadIdx.x & 31;
    const int count = *sq_count;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int i = chunk * 32 + lane;
        const bool valid = i < count;
        uint32_t slot = 0;
        V3 o{0, 0, 0}, d{0, 0, 1};
        float range = 0;
        if (valid) {
            slot = sq[i];
            const float4 a = rb.sh_o[slot], b = rb.sh_d[slot];
            o = xyz(a), d = xyz(b), range = a.w;
        }
        const bool visible = unoccluded<MODE>(sc, s_cached, o, d, range, valid);
        if (valid && visible)
            st.L[slot] = rb.hit[slot];
    }
}

// what happens to a path once its extension ray is traced (wavefront_pt.cu:160-206): interaction at a hit +
// Russian roulette, environment / null emitter on a miss.  Returns whether the path goes on to the next bounce.
__device__ __forceinline__ bool resolve_hit(const DevScene &sc, PathState &st, uint32_t slot, uint32_t tag, V3 d, float hit_t,
                                            float hit_u, float hit_v, int prim, int bounce, int pt_rules) {
    bool alive = false;
    V4 thp = from_f4(st.thp[slot]);
    if (prim >= 0) {
        tag |= TAG_HIT | TAG_ACTIVE | uint32_t(prim);
        st.isect[slot] = make_interaction(sc, prim, hit_u, hit_v);
        alive = true;
        const float max_value = max3(thp);
        if (!pt_rules && bounce >= kRRBounce && max_value < kRRThreshold) {
            const uint2 rs = st.rng[slot];
            Rng rng{rs.x, rs.y};
            const bool survive = rng.next1d() < max_value && max_value > kThpEps;
            alive = survive;
            if (!survive)
                tag &= ~TAG_ACTIVE;
            thp *= survive ? 1.f / max_value : 0.f;
            st.thp[slot] = to_f4(thp);
            st.rng[slot] = make_uint2(rng.d0, rng.d1);
        }
    } else {
        V4 L = from_f4(st.L[slot]);
        L += thp * null_or_env(sc, d);
        st.L[slot] = to_f4(L);
        tag &= ~TAG_ACTIVE;
    }
    st.ray_o[slot].w = hit_t;
    st.ray_d[slot].w = __uint_as_float(tag);
    return alive;
}

// closest-hit traversal with dynamic ray fetch (see the pipeline note above k_shadow_gen)
template <int MINB>
__global__ void __launch_bounds__(kBlock, MINB) k_intersect(DevScene sc, PathState st, RayBufs rb, const uint32_t *__restrict__ queue,
                                                            const int *__restrict__ q_count, int *work, int kRefill, int kSteps) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    uint2 *s_stack = reinterpret_cast<uint2 *>(s_cached + sc.num_cache) + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int count = *q_count;
    uint2 spill[kStackSize];
    int sp = 0, cur = kTraversalDone, start = 0;
    uint32_t slot = 0;
    V3 o{0, 0, 0}, d{0, 0, 1}, inv_d{0, 0, 0}, o_div{0, 0, 0};
    float ht = kMaxDist, hu = 0, hv = 0;
    int hprim = -1;
    bool active = false, more = true;
    int next = 0, end = 0;
    const float inf = __int_as_float(0x7f800000);
    while (true) {
        const unsigned act = __ballot_sync(0xffffffffu, active);
        const int n_idle = 32 - __popc(act);
        if (more && n_idle >= kRefill) {
            int first = 0;
            const int take = warp_take(work, count, lane, n_idle, next, end, more, first);
            const int rank = __popc(~act & ((1u << lane) - 1u));
            if (!active && rank < take) {
                slot = queue[first + rank];
                const float4 a = st.ray_o[slot], b = st.ray_d[slot];
                if (__float_as_uint(b.w) & TAG_ACTIVE) { // megakernel rules: k_shade may have ended the path
                    o = xyz(a), d = xyz(b);
                    inv_d = rcp3(d), o_div = o * inv_d;
                    start = cached_entry(s_cached, sc.num_cache, inv_d, o_div, kMaxDist, 0);
                    ht = kMaxDist, hu = hv = 0.f, hprim = -1;
                    cur = sc.wroot, sp = 0, active = true;
                }
            }
        } else if (act == 0)
            break;
#pragma unroll 1
        for (int k = 0; k < kSteps; k++) {
            if (active) {
                bool pop = false;
                if (cur >= 0) {
                    float4 a0, a1, b0, b1;
                    load_pair_node<true>(sc.wnodes, cur, a0, a1, b0, b1);
                    float t0, t1;
                    const bool h0 = slab_test(a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, inv_d, o_div, t0) && t0 < ht &&
                                    __float_as_int(a1.w) > start;
                    const bool h1 = slab_test(b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, inv_d, o_div, t1) && t1 < ht &&
                                    __float_as_int(b1.w) > start;
                    const int r0 = __float_as_int(a0.w), r1 = __float_as_int(b0.w);
                    if (h0 && h1) {
                        const bool swap = t1 < t0;
                        const uint2 e = make_uint2(uint32_t(swap ? r0 : r1), __float_as_uint(swap ? t0 : t1));
                        if (sp < kSmemTop)
                            s_stack[sp * kBlock] = e;
                        else
                            spill[sp - kSmemTop] = e;
                        ++sp;
                        cur = swap ? r1 : r0;
                    } else if (h0 || h1) {
                        cur = h0 ? r0 : r1;
                    } else {
                        pop = true;
                    }
                } else {
                    const int packed = ~cur, base = packed & kLeafBaseMask, cnt = packed >> kLeafCountShift;
                    for (int idx = base; idx < base + cnt; idx++) {
                        float u, v;
                        const float dist = hit_triangle(sc.tris, idx, o, d, u, v, inf);
                        // equal distances: the reference keeps the first hit in DFS order == the smaller primitive index
                        const bool better = dist > kEps && (dist < ht || (dist == ht && prim_before(sc, idx, hprim)));
                        ht = better ? dist : ht;
                        hu = better ? u : hu;
                        hv = better ? v : hv;
                        hprim = better ? idx : hprim;
                    }
                    pop = true;
                }
                if (pop) {
                    cur = kTraversalDone;
                    while (sp > 0) {
                        --sp;
                        const uint2 e = sp < kSmemTop ? s_stack[sp * kBlock] : spill[sp - kSmemTop];
                        if (__uint_as_float(e.y) < ht) {
                            cur = int(e.x);
                            break;
                        }
                    }
                    if (cur == kTraversalDone) {
                        rb.hit[slot] = make_float4(ht, hu, hv, __int_as_float(hprim));
                        active = false;
                    }
                }
            }
        }
    }
}

// the per-path epilogue of k_trace for the dynamic-fetch pipeline: reads k_intersect's hit records in queue order
__global__ void __launch_bounds__(kBlock) k_resolve(DevScene sc, PathState st, RayBufs rb, const uint32_t *__restrict__ queue,
                                                    const int *__restrict__ q_count, int *work, uint32_t *q_next,
                                                    int *q_next_count, int bounce, int pt_rules) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    const int lane = threadIdx.x & 31;
    const int count = *q_count;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int i = chunk * 32 + lane;
        bool alive = false;
        uint32_t slot = 0;
        if (i < count) {
            slot = queue[i];
            const float4 rd = st.ray_d[slot];
            const uint32_t tag = __float_as_uint(rd.w) & 0x50000000u; // reset(): clear hit flag and hit index
            if (tag & TAG_ACTIVE) {
                float4 h = rb.hit[slot];
                if (__float_as_int(h.w) == kPrimOddRay) { // stream traversal: a ray with a zero direction component, left to us
                    const Hit hh = closest_hit_ordered<true, false>(sc, s_cached, xyz(st.ray_o[slot]), xyz(rd));
                    h = make_float4(hh.t, hh.u, hh.v, __int_as_float(hh.prim));
                }
                alive = resolve_hit(sc, st, slot, tag, xyz(rd), h.x, h.y, h.z, __float_as_int(h.w), bounce, pt_rules);
            }
        }
        warp_push(q_next, q_next_count, alive, slot, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// k_trace — reference: fused_closesthit_shader (wavefront_pt.cu:141-207) + the sort/lower_bound
// compaction of wf_path_tracer.cu:197-202
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_trace(DevScene sc, PathState st, const uint32_t *__restrict__ queue,
                                                  const int *__restrict__ q_count, int *work, uint32_t *q_next,
                                                  int *q_next_count, int bounce, int pt_rules) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    const int lane = threadIdx.x & 31;
    const int count = *q_count;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int i = chunk * 32 + lane;
        bool alive = false;
        uint32_t slot = 0, tag = 0;
        bool valid = i < count;
        V3 o{0, 0, 0}, d{0, 0, 1};
        if (valid) {
            slot = queue[i];
            const float4 ro = st.ray_o[slot], rd = st.ray_d[slot];
            o = xyz(ro), d = xyz(rd);
            tag = __float_as_uint(rd.w) & 0x50000000u; // reset(): clear hit flag and hit index (and the split pipeline's NEE bit)
            valid = (tag & TAG_ACTIVE) != 0;           // megakernel rules: k_shade may have ended the path
        }
        const Hit h = closest_hit<MODE>(sc, s_cached, o, d, valid); // all 32 lanes call
        if (valid)
            alive = resolve_hit(sc, st, slot, tag, d, h.t, h.u, h.v, h.prim, bounce, pt_rules);
        warp_push(q_next, q_next_count, alive, slot, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// traversal=6: per-thread ordered traversal of the pair layout with DEFERRED LEAVES and LANE REFILL.
// The lock-step model (tools/sim/trav_sim.cpp) says where the per-thread kernels lose their lanes: in ~97 % of
// its iterations a warp pays for a leaf step (14 + 66 instructions per triangle) on behalf of ~3 lanes, and it idles the lanes
// whose rays finished early.  Here a lane that reaches a leaf does not test it: the leaf goes to the lane's pending list (2
// entries) and the lane pops on; the warp runs a triangle phase (every lane with a pending leaf tests one) only when enough
// lanes have one or a lane cannot go on without it.  Lanes whose ray is done take the next rays of the queue as soon as
// `refill` lanes are idle; the per-ray prologue (fetch, 1/d, fp16 cached phase) and epilogue (interaction, Russian roulette,
// compaction) run then, for those lanes.  Same boxes, same tests, same tie rule as closest_hit_ordered; deferral only delays
// updates of the closest distance (weaker culling, same result).
// ------------------------------------------------------------------------------------------------
#ifndef DQ_MIN_BLOCKS
#define DQ_MIN_BLOCKS 10
#endif
template <bool SHADOW>
__global__ void __launch_bounds__(kBlock, DQ_MIN_BLOCKS)
    k_traverse_dq(DevScene sc, PathState st, RayBufs rb, const uint32_t *__restrict__ queue, const int *__restrict__ q_count, int *work,
                  uint32_t *q_next, int *q_next_count, int bounce, int pt_rules, int kRefill, int kFlush, int kSteps) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    const int lane = threadIdx.x & 31;
    const int count = *q_count;
    int stack_ref[kStackSize];
    float stack_t[kStackSize];
    int sp = 0, cur = kTraversalDone, start = 0, npend = 0, pend0 = 0, pend1 = 0, hprim = -1;
    uint32_t slot = 0, tag = 0;
    V3 o{0, 0, 0}, d{0, 0, 1}, inv_d{0, 0, 0}, o_div{0, 0, 0};
    float ht = kMaxDist, hu = 0.f, hv = 0.f;
    bool active = false, has_result = false, more = true;
    int next = 0, end = 0;
    const float inf = __int_as_float(0x7f800000);
    while (true) {
        unsigned act = __ballot_sync(0xffffffffu, active);
        const int n_idle = 32 - __popc(act);
        if ((more && n_idle >= kRefill) || act == 0u) {
            // ---- epilogue of the rays that finished since the last refill
            if constexpr (!SHADOW) {
                bool alive = false;
                if (has_result)
                    alive = resolve_hit(sc, st, slot, tag, d, ht, hu, hv, hprim, bounce, pt_rules);
                warp_push(q_next, q_next_count, alive, slot, lane);
            } else {
                if (has_result && hprim < 0) // unoccluded: the radiance k_shade<-2> prepared for that case
                    st.L[slot] = rb.hit[slot];
            }
            has_result = false;
            // ---- prologue of the next rays
            if (more) {
                int first = 0;
                const int take = warp_take(work, count, lane, n_idle, next, end, more, first);
                const int rank = __popc(~act & ((1u << lane) - 1u));
                if (!active && rank < take) {
                    slot = queue[first + rank];
                    bool valid = true;
                    if constexpr (SHADOW) {
                        const float4 a = rb.sh_o[slot], b = rb.sh_d[slot];
                        o = xyz(a), d = xyz(b), ht = a.w;
                    } else {
                        const float4 a = st.ray_o[slot], b = st.ray_d[slot];
                        o = xyz(a), d = xyz(b), ht = kMaxDist;
                        tag = __float_as_uint(b.w) & 0x50000000u; // reset(): clear hit flag and hit index
                        valid = (tag & TAG_ACTIVE) != 0;         // megakernel rules: k_shade may have ended the path
                    }
                    if (valid) {
                        inv_d = rcp3(d), o_div = o * inv_d;
                        hu = hv = 0.f, hprim = -1, npend = 0, sp = 0;
                        const float big = 1e28f;
                        if (!(fabsf(inv_d.x) < big && fabsf(inv_d.y) < big && fabsf(inv_d.z) < big)) {
                            // (near-)zero direction component: the compare+select slab form, per thread, right here
                            if constexpr (SHADOW) {
                                hprim = unoccluded_ordered<true, false>(sc, s_cached, o, d, ht) ? -1 : 0;
                            } else {
                                const Hit hh = closest_hit_ordered<true, false>(sc, s_cached, o, d);
                                ht = hh.t, hu = hh.u, hv = hh.v, hprim = hh.prim;
                            }
                            has_result = true;
                        } else {
                            start = cached_entry(s_cached, sc.num_cache, inv_d, o_div, ht, SHADOW ? sc.num_nodes : 0);
                            if (SHADOW && start >= sc.num_nodes) { // no cached frontier node on the way: unoccluded (tracing_func.cuh:49)
                                has_result = true;
                            } else {
                                cur = sc.wroot, active = true;
                            }
                        }
                    }
                }
            }
            act = __ballot_sync(0xffffffffu, active);
            if (act == 0u) {
                if (!more && !__any_sync(0xffffffffu, has_result))
                    break;
                continue; // results of rays that needed no traversal: handled by the next round
            }
        }
#pragma unroll 1
        for (int k = 0; k < kSteps; k++) {
            if (active && cur != kTraversalDone && !(cur < 0 && npend == 2)) {
                bool pop = false;
                if (cur >= 0) {
                    float4 a0, a1, b0, b1;
                    load_pair_node<true>(sc.wnodes, cur, a0, a1, b0, b1);
                    float t0, t1;
                    const bool h0 = slab_test_finite(a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, inv_d, o_div, t0) && t0 < ht && __float_as_int(a1.w) > start;
                    const bool h1 = slab_test_finite(b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, inv_d, o_div, t1) && t1 < ht && __float_as_int(b1.w) > start;
                    const int r0 = __float_as_int(a0.w), r1 = __float_as_int(b0.w);
                    if (h0 && h1) {
                        const bool swap = !SHADOW && t1 < t0;
                        stack_ref[sp] = swap ? r0 : r1;
                        stack_t[sp] = swap ? t0 : t1;
                        ++sp;
                        cur = swap ? r1 : r0;
                    } else if (h0 || h1) {
                        cur = h0 ? r0 : r1;
                    } else {
                        pop = true;
                    }
                } else { // a leaf: remember it, go on with the stack
                    if (npend == 0)
                        pend0 = cur;
                    else
                        pend1 = cur;
                    ++npend;
                    pop = true;
                }
                if (pop) {
                    cur = kTraversalDone;
                    while (sp > 0) {
                        --sp;
                        if (SHADOW || stack_t[sp] < ht) {
                            cur = stack_ref[sp];
                            break;
                        }
                    }
                }
            }
        }
        // ---- triangle phase: when enough lanes hold a pending leaf, or a lane cannot go on without it
        const bool blocked = active && npend > 0 && (cur == kTraversalDone || (cur < 0 && npend == 2));
        const unsigned have = __ballot_sync(0xffffffffu, active && npend > 0);
        if (__popc(have) >= kFlush || __any_sync(0xffffffffu, blocked)) {
            if (active && npend > 0) {
                --npend;
                const int packed = ~(npend == 1 ? pend1 : pend0), base = packed & kLeafBaseMask, cnt = packed >> kLeafCountShift;
                for (int idx = base; idx < base + cnt; idx++) {
                    float u, v;
                    const float dist = hit_triangle(sc.tris, idx, o, d, u, v, SHADOW ? ht : inf);
                    if constexpr (SHADOW) {
                        if (dist > kEps)
                            hprim = idx, cur = kTraversalDone, npend = 0, sp = 0; // occluded: done
                    } else {
                        const bool better = dist > kEps && (dist < ht || (dist == ht && prim_before(sc, idx, hprim)));
                        ht = better ? dist : ht;
                        hu = better ? u : hu;
                        hv = better ? v : hv;
                        hprim = better ? idx : hprim;
                    }
                }
            }
        }
        if (active && cur == kTraversalDone && npend == 0) { // this ray is done; its epilogue runs with the next refill
            active = false;
            has_result = true;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// traversal=7 ("stream"): traversal=6's walk (per-thread ordered traversal, leaves deferred into a 2-entry pending list, warp-wide
// triangle phases, idle lanes refilled) with everything that made traversal=6 wait on memory taken out of the walk:
//  * ray fetch is an asynchronous pipeline per warp.  One stage per pass through the main loop: reserve an aligned 32-ray
//    chunk of the queue (atomic), `cp.async` its slot ids into shared memory, `cp.async`-gather the ray records into the warp's
//    ring in shared memory, and - a pass later, when they have landed - run the per-ray prologue (1/d, the fp16 cached phase of
//    tracing_func.cuh:124-142, rays that need no traversal answered on the spot) with one ray per lane at 32 of 32 lanes and
//    compact the rays that do need a traversal to the ring's tail.  The walk never waits for global memory on behalf of a new
//    ray: a lane whose ray is done takes the next one from shared memory.
//  * the epilogue runs per CHUNK, in queue order, at 32 of 32 lanes: a finished ray stores its hit record and counts its chunk
//    down; when the last ray of a chunk is done the warp runs interaction, Russian roulette and compaction for the chunk's 32
//    queue entries (stream_resolve_chunk = k_trace's epilogue).  The next bounce's queue therefore keeps the slot order the
//    shading kernel's gathers rely on (traversal=6 retired rays one by one and k_shade paid for it).  Shadow rays store one
//    visibility byte; the radiance swap runs in k_shadow_finish.
// Same boxes, same tests, same tie rule as closest_hit_ordered; deferral only delays updates of the closest distance.
// The cold stages are out-of-line functions on purpose: inlined, their registers make ptxas spill the walk's loop-carried state.
// Build knobs (measured one by one, profiles/r02_stream_experiments.md): SQ_REGSTATE (the walking ray's cold state in registers - 56
// per thread at 9 CTAs per SM - instead of shared memory), SQ_SMEM_STACK (stack entries per lane in shared memory), SQ_RING, SQ_MIN_BLOCKS,
// SQ_INLINE_RESOLVE (0: hit records only, k_resolve does the epilogue as its own launch), SQ_SHADOW_DEFER (0: any-hit rays test leaves
// where they find them).
// ------------------------------------------------------------------------------------------------
#ifndef SQ_REGSTATE
#define SQ_REGSTATE 1
#endif
#ifndef SQ_INLINE_RESOLVE
#define SQ_INLINE_RESOLVE 1
#endif
#ifndef SQ_MIN_BLOCKS
#define SQ_MIN_BLOCKS (SQ_REGSTATE ? 9 : 10)
#endif
#ifndef SQ_RING
#define SQ_RING 32
#endif
constexpr int kRing = SQ_RING; // ring entries per warp: a power of two, >= 32
static_assert((kRing & (kRing - 1)) == 0 && kRing >= 32 && kRing <= 1024, "SQ_RING must be a power of two in [32, 1024]");
#ifndef SQ_SMEM_STACK
#define SQ_SMEM_STACK 12
#endif
#ifndef SQ_SHADOW_DEFER
#define SQ_SHADOW_DEFER 1
#endif
constexpr int kSqSmemStack = SQ_SMEM_STACK; // closest-hit stack entries per lane kept in shared memory (the rest in local memory)
constexpr int kChunkTab = 16;                      // chunks a warp may have open (rays walking or waiting in the ring)
constexpr int kStartBits = 27, kStartMask = (1 << kStartBits) - 1; // ring word: cached-phase start index | chunk-table entry << 27
constexpr int kLaneWords = SQ_REGSTATE ? 1 : 14;   // per-lane words kept in shared memory (enum below)
constexpr size_t kLaneBytes = size_t((kLaneWords * 32 * 4 + 15) / 16 * 16);
constexpr size_t kRingBytes = kLaneBytes + size_t(kRing) * 36 + size_t(kChunkTab) * 8;
enum { W_FREG = 0, W_OX, W_OY, W_OZ, W_DX, W_DY, W_DZ, W_ID, W_HU, W_HV, W_PRIM, W_PEND0, W_PEND1, W_CHUNK };

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned s = unsigned(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    const unsigned s = unsigned(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// what the out-of-line stages need, once per CTA in shared memory (passing the kernel's parameter struct by reference would make
// nvcc copy it to local memory and route the walk's own accesses through that copy)
struct StreamCtx {
    DevScene sc;
    PathState st;
    RayBufs rb;
    const uint32_t *queue;
    uint32_t *q_next;
    int *q_next_count;
    unsigned char *vis;
    int count, bounce, pt_rules, pad;
};
constexpr size_t kCtxBytes = (sizeof(StreamCtx) + 15) / 16 * 16;

// warp-uniform state of a warp's fetch pipeline, packed so that it crosses the stage function in two registers
struct StreamFetch {
    uint32_t ht;  // head | tail << 16: ring[head, tail) = rays ready to walk (16-bit wrapping counters, index & (kRing - 1))
    uint32_t ctl; // stage (2 bits) | rays of the chunk in flight << 2 | queue exhausted << 8 | chunk-table entry << 9
    __device__ int head() const { return int(ht & 0xffffu); }
    __device__ int tail() const { return int(ht >> 16); }
    __device__ int avail() const { return (tail() - head()) & 0xffff; }
    __device__ int stage() const { return int(ctl & 3u); }
    __device__ int fcount() const { return int((ctl >> 2) & 63u); }
    __device__ bool more() const { return (ctl & 256u) == 0u; }
    __device__ int entry() const { return int((ctl >> 9) & 15u); }
};

// k_trace's epilogue for the 32 queue entries [fbase, fbase + 32) of one finished chunk (see k_resolve).  Full warp.
__device__ __noinline__ void stream_resolve_chunk(const StreamCtx *ctx, const uint4 *s_cached, int fbase) {
    const int lane = threadIdx.x & 31;
    const int i = fbase + lane;
    bool alive = false;
    uint32_t slot = 0;
    if (i < ctx->count) {
        PathState st = ctx->st;
        slot = ctx->queue[i];
        const float4 rd = st.ray_d[slot];
        const uint32_t tag = __float_as_uint(rd.w) & 0x50000000u; // reset(): clear hit flag and hit index
        if (tag & TAG_ACTIVE) {
            float4 h = ctx->rb.hit[slot];
            if (__float_as_int(h.w) == kPrimOddRay) { // a ray with a zero direction component, left to us by the prologue
                const Hit hh = closest_hit_ordered<true, false>(ctx->sc, s_cached, xyz(st.ray_o[slot]), xyz(rd));
                h = make_float4(hh.t, hh.u, hh.v, __int_as_float(hh.prim));
            }
            alive = resolve_hit(ctx->sc, st, slot, tag, xyz(rd), h.x, h.y, h.z, __float_as_int(h.w), ctx->bounce, ctx->pt_rules);
        }
    }
    warp_push(ctx->q_next, ctx->q_next_count, alive, slot, lane);
}

// One stage of the fetch pipeline.  Must be called by the full warp.  `chunk`: lane 0 holds the chunk number reserved a pass ago.
template <bool SHADOW>
__device__ __noinline__ StreamFetch stream_fetch_stage(StreamFetch f, int chunk, const StreamCtx *ctx, const uint4 *s_cached, unsigned char *wbase) {
    const int lane = threadIdx.x & 31;
    uint32_t *freg = reinterpret_cast<uint32_t *>(wbase) + W_FREG * 32 + lane;
    float4 *ringA = reinterpret_cast<float4 *>(wbase + kLaneBytes); // o.xyz, tmax
    float4 *ringB = ringA + kRing;                                  // d.xyz, bits(start | entry << 27)
    uint32_t *ringS = reinterpret_cast<uint32_t *>(ringB + kRing);   // slot (closest hit) / queue position (shadow)
    int *tab_base = reinterpret_cast<int *>(ringS + kRing);
    volatile int *tab_cnt = tab_base + kChunkTab;
    const int tail = f.tail(), stage = f.stage(), fcount = f.fcount();
    constexpr bool kTable = !SHADOW && SQ_INLINE_RESOLVE;
    if (stage == 3) { // the chunk's ray records have landed at ring[tail .. tail + fcount): prologue, one ray per lane
        cp_async_wait_all();
        __syncwarp();
        const int pos = (tail + lane) & (kRing - 1);
        const int num_cache = ctx->sc.num_cache, num_nodes = ctx->sc.num_nodes;
        bool trav = false;
        float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 1, 0);
        uint32_t rid = 0;
        if (lane < fcount) {
            a = ringA[pos], b = ringB[pos], rid = ringS[pos];
            bool valid = true;
            if constexpr (!SHADOW) {
                valid = (__float_as_uint(b.w) & TAG_ACTIVE) != 0; // megakernel rules: k_shade may have ended the path
                a.w = kMaxDist;
            }
            if (valid) {
                const V3 po = xyz(a), pd = xyz(b);
                const V3 pinv = rcp3(pd), podiv = po * pinv;
                const float big = 1e28f;
                if (!(fabsf(pinv.x) < big && fabsf(pinv.y) < big && fabsf(pinv.z) < big)) {
                    // (near-)zero direction component: needs the compare+select slab form.  Rare; marked here and traced per thread by
                    // the code that reads the result anyway (k_shadow_finish / the chunk epilogue)
                    if constexpr (SHADOW)
                        ctx->vis[rid] = kVisOddRay;
                    else
                        ctx->rb.hit[rid] = make_float4(kMaxDist, 0.f, 0.f, __int_as_float(kPrimOddRay));
                } else {
                    const int st0 = cached_entry(s_cached, num_cache, pinv, podiv, a.w, SHADOW ? num_nodes : 0);
                    if (SHADOW && st0 >= num_nodes) {
                        ctx->vis[rid] = 1; // no cached frontier node on the way: unoccluded (tracing_func.cuh:49)
                    } else {
                        trav = true;
                        b.w = __int_as_float(st0 | (kTable ? f.entry() << kStartBits : 0));
                    }
                }
            }
        }
        __syncwarp(); // every lane has read its record before the compacted ones are written
        const unsigned m = __ballot_sync(0xffffffffu, trav);
        if (trav) {
            const int p = (tail + __popc(m & ((1u << lane) - 1u))) & (kRing - 1);
            ringA[p] = a, ringB[p] = b, ringS[p] = rid;
        }
        if constexpr (kTable) {
            const int e = f.entry();
            if (lane == 0)
                tab_cnt[e] = __popc(m); // 0: nothing of this chunk walks, its epilogue is due right away (the caller looks at the table)
        }
        __syncwarp();
        f.ht = (f.ht & 0xffffu) | (uint32_t((tail + __popc(m)) & 0xffff) << 16);
        f.ctl &= ~255u; // stage 0
        if (kTable && m == 0u)
            f.ctl |= 0x2000u;
    } else if (stage == 2) { // slot ids are here (shared memory): start the gathers of the ray records
        cp_async_wait_all();
        if (lane < fcount) {
            const int pos = (tail + lane) & (kRing - 1);
            const uint32_t slot = *freg;
            const float4 *src_o = SHADOW ? ctx->rb.sh_o : ctx->st.ray_o, *src_d = SHADOW ? ctx->rb.sh_d : ctx->st.ray_d;
            cp_async16(ringA + pos, src_o + slot);
            cp_async16(ringB + pos, src_d + slot);
            if constexpr (!SHADOW)
                ringS[pos] = slot;
        }
        cp_async_commit();
        f.ctl |= 3u;
    } else { // stage 1: the chunk number is here: fetch the slot ids
        const int fbase = __shfl_sync(0xffffffffu, chunk, 0) * 32;
        const int count = ctx->count;
        if (fbase >= count) {
            f.ctl = 256u; // queue exhausted, stage 0
        } else {
            int e = 0;
            if constexpr (kTable) { // an entry of the warp's chunk table; none free: try again next pass
                const unsigned free_mask = __ballot_sync(0xffffffffu, lane < kChunkTab && tab_cnt[lane & (kChunkTab - 1)] < 0);
                if (free_mask == 0u)
                    return f;
                e = __ffs(free_mask) - 1;
                if (lane == 0)
                    tab_base[e] = fbase, tab_cnt[e] = 1 << 20; // open; the prologue sets the number of walking rays
            }
            const int n = min(32, count - fbase);
            if (lane < n) {
                cp_async4(freg, ctx->queue + fbase + lane);
                if constexpr (SHADOW)
                    ringS[(tail + lane) & (kRing - 1)] = uint32_t(fbase + lane);
            }
            cp_async_commit();
            f.ctl = 2u | (uint32_t(n) << 2) | (uint32_t(e) << 9);
        }
    }
    return f;
}

// epilogues of the chunks in `done` (bit e = chunk-table entry e).  Heavy (uses the whole register file): the caller parks its
// walk state in local memory around this call (see k_traverse_stream), so that ptxas keeps that state in registers elsewhere.
__device__ __noinline__ void stream_resolve_done(unsigned done, const StreamCtx *ctx, const uint4 *s_cached, unsigned char *wbase) {
    const int lane = threadIdx.x & 31;
    int *tab_base = reinterpret_cast<int *>(wbase + kLaneBytes + size_t(kRing) * 36);
    volatile int *tab_cnt = tab_base + kChunkTab;
    while (done) {
        const int e = __ffs(done) - 1;
        done &= done - 1u;
        stream_resolve_chunk(ctx, s_cached, tab_base[e]);
        __syncwarp();
        if (lane == 0)
            tab_cnt[e] = -1;
    }
    __syncwarp();
}

template <bool SHADOW>
__global__ void __launch_bounds__(kBlock, SQ_MIN_BLOCKS)
    k_traverse_stream(DevScene sc, PathState st, RayBufs rb, const uint32_t *__restrict__ queue, const int *__restrict__ q_count, int *work,
                      unsigned char *__restrict__ vis, uint32_t *q_next, int *q_next_count, int bounce, int pt_rules, int kRefill, int kFlush, int kSteps) {
    extern __shared__ uint4 s_cached[];
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    StreamCtx *ctx = reinterpret_cast<StreamCtx *>(s_cached + sc.num_cache);
    unsigned char *wbase = reinterpret_cast<unsigned char *>(ctx) + kCtxBytes + size_t(threadIdx.x >> 5) * kRingBytes;
    if (threadIdx.x == 0) {
        ctx->sc = sc, ctx->st = st, ctx->rb = rb, ctx->queue = queue, ctx->q_next = q_next, ctx->q_next_count = q_next_count, ctx->vis = vis;
        ctx->count = *q_count, ctx->bounce = bounce, ctx->pt_rules = pt_rules, ctx->pad = 0;
    }
    const float4 *ringA = reinterpret_cast<const float4 *>(wbase + kLaneBytes);
    const float4 *ringB = ringA + kRing;
    const uint32_t *ringS = reinterpret_cast<const uint32_t *>(ringB + kRing);
    int *tab_base = reinterpret_cast<int *>(const_cast<uint32_t *>(ringS) + kRing);
    volatile int *tab_cnt = tab_base + kChunkTab;
    if (lane < kChunkTab)
        tab_cnt[lane] = -1; // all entries free
    stage_cached_nodes(sc, s_cached); // (ends with __syncthreads: the context and the tables are visible)
    constexpr bool kTable = !SHADOW && SQ_INLINE_RESOLVE;
    // the walking ray's cold state - what the box tests do not need: o, d (triangle tests), id / hu / hv / hprim / chunk (retire),
    // the two pending leaves
#if SQ_REGSTATE
    V3 ro{0, 0, 0}, rd{0, 0, 1};
    float rhu = 0.f, rhv = 0.f;
    int rid = 0, rprim = -1, rpend0 = 0, rpend1 = 0, rchunk = 0;
#else
    // word w of lane l at lsf[w * 32] (conflict-free)
    float *lsf = reinterpret_cast<float *>(wbase) + lane;
    int *lsi = reinterpret_cast<int *>(lsf);
#endif
    // ---- walk state, one ray per lane
    uint2 stack[SHADOW ? 1 : kStackSize - kSqSmemStack]; // closest hit: {ref, bits(t)}, one 8-byte access per push / pop
    // SQ_SMEM_STACK: the first entries of every lane's stack in shared memory ([level][thread]: conflict-free, one or two wavefronts per
    // warp instruction whatever the lanes' depths, and 32 walking lanes per warp no longer thrash L1 with their stacks)
    uint2 *sstack = reinterpret_cast<uint2 *>(reinterpret_cast<unsigned char *>(ctx) + kCtxBytes + size_t(kBlock / 32) * kRingBytes) + threadIdx.x;
    int stack_ref[SHADOW ? kStackSize - kSqSmemStack : 1];
    int *sstack_ref = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(ctx) + kCtxBytes + size_t(kBlock / 32) * kRingBytes) + threadIdx.x;
    int sp = 0, cur = kTraversalDone, start = 0, npend = 0;
    V3 inv_d{0, 0, 0}, o_div{0, 0, 0};
    float ht = kMaxDist;
    bool active = false; // (shadow rays: ht < 0 marks an occluded ray)
    StreamFetch f{0u, 0u};
    int pending = 0; // lane 0: the chunk number reserved for the fetch pipeline's next round
    unsigned done = 0u; // chunk-table entries whose last ray retired in the previous pass
    const float inf = __int_as_float(0x7f800000);
    while (true) {
        // ---------------- epilogues of the chunks that finished in the previous pass (about once per 32 rays)
        if constexpr (kTable) {
            if (done != 0u) {
                // park the walk state in local memory around the call: the callee needs the whole register file, and values that
                // live across it would be spilled by ptxas - with reloads inside the walk loop
                volatile float park[24];
                park[0] = inv_d.x, park[1] = inv_d.y, park[2] = inv_d.z, park[3] = o_div.x, park[4] = o_div.y, park[5] = o_div.z, park[6] = ht;
                park[7] = __int_as_float(sp), park[8] = __int_as_float(cur), park[9] = __int_as_float(start);
                park[10] = __int_as_float(npend | (active ? 4 : 0));
#if SQ_REGSTATE
                park[11] = ro.x, park[12] = ro.y, park[13] = ro.z, park[14] = rd.x, park[15] = rd.y, park[16] = rd.z, park[17] = rhu, park[18] = rhv;
                park[19] = __int_as_float(rid), park[20] = __int_as_float(rprim), park[21] = __int_as_float(rpend0), park[22] = __int_as_float(rpend1);
                park[23] = __int_as_float(rchunk);
#endif
                stream_resolve_done(done, ctx, s_cached, wbase);
                done = 0u;
                inv_d = V3{park[0], park[1], park[2]}, o_div = V3{park[3], park[4], park[5]}, ht = park[6];
                sp = __float_as_int(park[7]), cur = __float_as_int(park[8]), start = __float_as_int(park[9]);
                const int fl = __float_as_int(park[10]);
                npend = fl & 3, active = (fl & 4) != 0;
#if SQ_REGSTATE
                ro = V3{park[11], park[12], park[13]}, rd = V3{park[14], park[15], park[16]}, rhu = park[17], rhv = park[18];
                rid = __float_as_int(park[19]), rprim = __float_as_int(park[20]), rpend0 = __float_as_int(park[21]), rpend1 = __float_as_int(park[22]);
                rchunk = __float_as_int(park[23]);
#endif
            }
        }
        // ---------------- fetch pipeline: one stage per pass
        if (f.stage() != 0) {
            f = stream_fetch_stage<SHADOW>(f, pending, ctx, s_cached, wbase);
            if (kTable && (f.ctl & 0x2000u)) { // a chunk without walking rays: its epilogue runs at the top of the next pass
                f.ctl &= ~0x2000u;
                done |= __ballot_sync(0xffffffffu, lane < kChunkTab && tab_cnt[lane & (kChunkTab - 1)] == 0);
            }
        } else if (f.more() && f.avail() <= kRing - 32) { // room for a chunk: reserve one; the number is used a pass later
            if (lane == 0) // atom.inc, not atom.add: ptxas turns an add into its warp-aggregated form, whose shuffle waits for the result here
                asm volatile("atom.global.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(pending) : "l"(work) : "memory");
            f.ctl |= 1u;
        }
        // ---------------- refill idle lanes from the ring
        unsigned act = __ballot_sync(0xffffffffu, active);
        const int avail = f.avail();
        if (avail > 0) {
            const int n_idle = 32 - __popc(act);
            if (n_idle >= kRefill || act == 0u || (!f.more() && f.stage() == 0 && n_idle > 0)) {
                const int take = min(n_idle, avail);
                const int rank = __popc(~act & lt_mask);
                if (!active && rank < take) {
                    const int p = (f.head() + rank) & (kRing - 1);
                    const float4 a = ringA[p], b = ringB[p];
                    const int sw = __float_as_int(b.w);
#if SQ_REGSTATE
                    rid = int(ringS[p]), ro = xyz(a), rd = xyz(b);
                    rhu = rhv = 0.f, rprim = -1, rchunk = sw >> kStartBits;
#else
                    lsi[W_ID * 32] = int(ringS[p]);
                    lsf[W_OX * 32] = a.x, lsf[W_OY * 32] = a.y, lsf[W_OZ * 32] = a.z;
                    lsf[W_DX * 32] = b.x, lsf[W_DY * 32] = b.y, lsf[W_DZ * 32] = b.z;
                    if constexpr (!SHADOW)
                        lsf[W_HU * 32] = 0.f, lsf[W_HV * 32] = 0.f, lsi[W_PRIM * 32] = -1, lsi[W_CHUNK * 32] = sw >> kStartBits;
#endif
                    ht = a.w, start = sw & kStartMask;
                    inv_d = rcp3(xyz(b)), o_div = xyz(a) * inv_d;
                    npend = 0, sp = 0;
                    cur = sc.wroot, active = true;
                }
                f.ht = (f.ht & 0xffff0000u) | uint32_t((f.head() + take) & 0xffff);
                act = __ballot_sync(0xffffffffu, active);
            }
        }
        if (act == 0u) {
            if (!f.more() && f.stage() == 0 && f.avail() == 0 && done == 0u)
                break;
            continue; // nothing to walk yet: let the fetch pipeline advance
        }
        // ---------------- walk
#pragma unroll 1
        for (int k = 0; k < kSteps; k++) {
            if (active && cur != kTraversalDone && !(cur < 0 && npend == 2)) {
                bool pop = false;
                if (cur >= 0) {
                    float4 a0, a1, b0, b1;
                    load_pair_node<true>(sc.wnodes, cur, a0, a1, b0, b1);
                    float t0, t1;
                    const bool h0 = slab_test_finite(a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, inv_d, o_div, t0) && t0 < ht && __float_as_int(a1.w) > start;
                    const bool h1 = slab_test_finite(b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, inv_d, o_div, t1) && t1 < ht && __float_as_int(b1.w) > start;
                    const int r0 = __float_as_int(a0.w), r1 = __float_as_int(b0.w);
                    if (h0 && h1) {
                        const bool swap = !SHADOW && t1 < t0;
                        if constexpr (SHADOW) {
                            if (kSqSmemStack > 0 && sp < kSqSmemStack)
                                sstack_ref[sp * kBlock] = r1;
                            else
                                stack_ref[sp - kSqSmemStack] = r1;
                        } else
                        {
                            const uint2 e = make_uint2(uint32_t(swap ? r0 : r1), __float_as_uint(swap ? t0 : t1));
                            if (kSqSmemStack > 0 && sp < kSqSmemStack)
                                sstack[sp * kBlock] = e;
                            else
                                stack[sp - kSqSmemStack] = e;
                        }
                        ++sp;
                        cur = swap ? r1 : r0;
                    } else if (h0 || h1) {
                        cur = h0 ? r0 : r1;
                    } else {
                        pop = true;
                    }
                } else if (SHADOW && !SQ_SHADOW_DEFER) { // any-hit, leaves tested where they are found (an occluder ends the ray at once)
#if SQ_REGSTATE
                    const V3 o = ro, d = rd;
#else
                    const V3 o{lsf[W_OX * 32], lsf[W_OY * 32], lsf[W_OZ * 32]}, d{lsf[W_DX * 32], lsf[W_DY * 32], lsf[W_DZ * 32]};
#endif
                    const int packed = ~cur, base = packed & kLeafBaseMask, cnt = packed >> kLeafCountShift;
                    bool occl = false;
                    for (int idx = base; idx < base + cnt; idx++) {
                        float u, v;
                        const float dist = hit_triangle(sc.tris, idx, o, d, u, v, ht);
                        if (dist > kEps)
                            occl = true;
                    }
                    if (occl)
                        ht = -1.f, sp = 0;
                    pop = true;
                } else { // a leaf: remember it, go on with the stack
#if SQ_REGSTATE
                    if (npend == 0)
                        rpend0 = cur;
                    else
                        rpend1 = cur;
#else
                    lsi[(W_PEND0 + npend) * 32] = cur;
#endif
                    ++npend;
                    pop = true;
                }
                if (pop) {
                    cur = kTraversalDone;
                    while (sp > 0) {
                        --sp;
                        if constexpr (SHADOW) {
                            cur = (kSqSmemStack > 0 && sp < kSqSmemStack) ? sstack_ref[sp * kBlock] : stack_ref[sp - kSqSmemStack];
                            break;
                        } else {
                            const uint2 e = (kSqSmemStack > 0 && sp < kSqSmemStack) ? sstack[sp * kBlock] : stack[sp - kSqSmemStack];
                            if (__uint_as_float(e.y) < ht) {
                                cur = int(e.x);
                                break;
                            }
                        }
                    }
                }
            }
        }
        // ---------------- triangle phase: when enough lanes hold a pending leaf, or a lane cannot go on without it
        const bool blocked = active && npend > 0 && (cur == kTraversalDone || (cur < 0 && npend == 2));
        const unsigned have = __ballot_sync(0xffffffffu, active && npend > 0);
        if (__popc(have) >= kFlush || __any_sync(0xffffffffu, blocked)) {
            if (active && npend > 0) {
                --npend;
#if SQ_REGSTATE
                const int packed = ~(npend == 1 ? rpend1 : rpend0), base = packed & kLeafBaseMask, cnt = packed >> kLeafCountShift;
                const V3 o = ro, d = rd;
#else
                const int packed = ~lsi[(W_PEND0 + npend) * 32], base = packed & kLeafBaseMask, cnt = packed >> kLeafCountShift;
                const V3 o{lsf[W_OX * 32], lsf[W_OY * 32], lsf[W_OZ * 32]}, d{lsf[W_DX * 32], lsf[W_DY * 32], lsf[W_DZ * 32]};
#endif
                if constexpr (SHADOW) {
                    bool occl = false;
                    for (int idx = base; idx < base + cnt; idx++) {
                        float u, v;
                        const float dist = hit_triangle(sc.tris, idx, o, d, u, v, ht);
                        if (dist > kEps)
                            occl = true;
                    }
                    if (occl) // occluded: done
                        ht = -1.f, cur = kTraversalDone, npend = 0, sp = 0;
                } else {
#if SQ_REGSTATE
                    float bu = rhu, bv = rhv;
                    int hprim = rprim;
#else
                    float bu = 0.f, bv = 0.f;
                    int hprim = lsi[W_PRIM * 32];
#endif
                    bool changed = false;
                    for (int idx = base; idx < base + cnt; idx++) {
                        float u, v;
                        const float dist = hit_triangle(sc.tris, idx, o, d, u, v, inf);
                        const bool better = dist > kEps && (dist < ht || (dist == ht && prim_before(sc, idx, hprim)));
                        ht = better ? dist : ht;
                        bu = better ? u : bu;
                        bv = better ? v : bv;
                        hprim = better ? idx : hprim;
                        changed |= better;
                    }
#if SQ_REGSTATE
                    rhu = bu, rhv = bv, rprim = hprim;
                    (void)changed;
#else
                    if (changed)
                        lsf[W_HU * 32] = bu, lsf[W_HV * 32] = bv, lsi[W_PRIM * 32] = hprim;
#endif
                }
            }
        }
        // ---------------- retire
        const bool retire = active && cur == kTraversalDone && npend == 0;
        if (retire) {
            active = false;
#if SQ_REGSTATE
            const uint32_t id = uint32_t(rid);
            const float hu = rhu, hv = rhv;
            const int hprim = rprim, chunk_e = rchunk;
#else
            const uint32_t id = uint32_t(lsi[W_ID * 32]);
            const float hu = SHADOW ? 0.f : lsf[W_HU * 32], hv = SHADOW ? 0.f : lsf[W_HV * 32];
            const int hprim = SHADOW ? 0 : lsi[W_PRIM * 32], chunk_e = SHADOW ? 0 : lsi[W_CHUNK * 32];
#endif
            if constexpr (SHADOW) {
                vis[id] = ht < 0.f ? 0 : 1;
            } else {
                rb.hit[id] = make_float4(ht, hu, hv, __int_as_float(hprim));
                if constexpr (kTable)
                    atomicSub(const_cast<int *>(tab_cnt) + chunk_e, 1);
            }
            (void)hu, (void)hv, (void)hprim, (void)chunk_e;
        }
        if constexpr (kTable) {
            if (__any_sync(0xffffffffu, retire)) { // chunks whose last ray just finished: their epilogue runs at the top of the next pass
                __syncwarp();
                done |= __ballot_sync(0xffffffffu, lane < kChunkTab && tab_cnt[lane & (kChunkTab - 1)] == 0);
            }
        }
    }
}

// the radiance swap of k_shadow_apply for the stream traversal: an unoccluded shadow ray takes the value k_shade<-2> prepared.
// Rays the stream kernel marked kVisOddRay (zero direction component) are traced here, per thread, with the select-form slab test.
__global__ void __launch_bounds__(256) k_shadow_finish(DevScene sc, PathState st, RayBufs rb, const uint32_t *__restrict__ sq, const int *__restrict__ sq_count,
                                                       const unsigned char *__restrict__ vis) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    const int count = *sq_count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const unsigned char v = vis[i];
        if (v == 0)
            continue;
        const uint32_t slot = sq[i];
        bool visible = true;
        if (v == kVisOddRay) {
            const float4 a = rb.sh_o[slot], b = rb.sh_d[slot];
            visible = unoccluded_ordered<true, false>(sc, s_cached, xyz(a), xyz(b), a.w);
        }
        if (visible)
            st.L[slot] = rb.hit[slot];
    }
}

// ------------------------------------------------------------------------------------------------
// warp-cooperative forms of the three traversal kernels of the default pipeline (traversal=4, traverse_coop.cuh):
// same prologue / epilogue per path as k_raygen / k_shadow_apply / k_trace, one ray per lane; in between the warp's
// 32 rays are traversed by lane groups over the wide layout.  Dynamic shared memory: coop_layout().
// ------------------------------------------------------------------------------------------------
constexpr int kCoopBlock = 256;

template <int G, bool SLAB>
__global__ void __launch_bounds__(kCoopBlock)
    k_trace_coop(DevScene sc, CoopParams cp, PathState st, const uint32_t *__restrict__ queue, const int *__restrict__ q_count, int *work,
                 uint32_t *q_next, int *q_next_count, int bounce, int pt_rules) {
    extern __shared__ __align__(128) unsigned char coop_smem[];
    const CoopShared S = coop_setup(coop_smem, sc, cp, G, SLAB);
    const int lane = threadIdx.x & 31;
    const int count = *q_count;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int i = chunk * 32 + lane;
        uint32_t slot = 0, tag = 0;
        bool valid = i < count;
        V3 o{0, 0, 0}, d{0, 0, 1};
        int start = 0;
        if (valid) {
            slot = queue[i];
            const float4 ro = st.ray_o[slot], rd = st.ray_d[slot];
            o = xyz(ro), d = xyz(rd);
            tag = __float_as_uint(rd.w) & 0x50000000u; // reset(): clear hit flag and hit index
            valid = (tag & TAG_ACTIVE) != 0;           // megakernel rules: k_shade may have ended the path
        }
        float4 h0 = make_float4(kMaxDist, 0.f, 0.f, __int_as_float(-1));
        bool coop_ray = valid;
        const V3 inv_d = rcp3(d), o_div = o * inv_d;
        if (valid && !coop_ray_ok(inv_d)) { // a (near-)zero direction component: per-thread traversal with the compare+select slab test
            const Hit hh = closest_hit_ordered<true, false>(sc, S.cached, o, d);
            h0 = make_float4(hh.t, hh.u, hh.v, __int_as_float(hh.prim));
            coop_ray = false;
        }
        if (coop_ray)
            start = cached_entry(S.cached, sc.num_cache, inv_d, o_div, kMaxDist, 0);
        coop_store_ray(S, lane, o, d, inv_d, o_div, start, coop_ray ? kMaxDist : -1.f);
        S.hits[lane] = h0;
        if (lane == 0)
            *S.next = 0;
        __syncwarp();
        coop_traverse<G, false, SLAB>(sc, cp, S, lane);
        const float4 h = S.hits[lane];
        bool alive = false;
        if (valid)
            alive = resolve_hit(sc, st, slot, tag, d, h.x, h.y, h.z, __float_as_int(h.w), bounce, pt_rules);
        warp_push(q_next, q_next_count, alive, slot, lane);
    }
}

template <int G, bool SLAB>
__global__ void __launch_bounds__(kCoopBlock)
    k_shadow_apply_coop(DevScene sc, CoopParams cp, PathState st, RayBufs rb, const uint32_t *__restrict__ sq, const int *__restrict__ sq_count, int *work) {
    extern __shared__ __align__(128) unsigned char coop_smem[];
    const CoopShared S = coop_setup(coop_smem, sc, cp, G, SLAB);
    const int lane = threadIdx.x & 31;
    const int count = *sq_count;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int i = chunk * 32 + lane;
        const bool valid = i < count;
        uint32_t slot = 0;
        V3 o{0, 0, 0}, d{0, 0, 1};
        float range = -1.f;
        int start = 0;
        if (valid) {
            slot = sq[i];
            const float4 a = rb.sh_o[slot], b = rb.sh_d[slot];
            o = xyz(a), d = xyz(b), range = a.w;
        }
        float vis0 = 1.f;
        const V3 inv_d = rcp3(d), o_div = o * inv_d;
        if (valid && !coop_ray_ok(inv_d)) { // per-thread traversal with the compare+select slab test
            vis0 = unoccluded_ordered<true, false>(sc, S.cached, o, d, range) ? 1.f : 0.f;
            range = -1.f;
        }
        if (range > 0.f)
            start = cached_entry(S.cached, sc.num_cache, inv_d, o_div, range, sc.num_nodes);
        coop_store_ray(S, lane, o, d, inv_d, o_div, start, range); // a range <= 0 cannot be occluded (no t in (EPSILON, range))
        S.hits[lane] = make_float4(vis0, 0.f, 0.f, 0.f);
        if (lane == 0)
            *S.next = 0;
        __syncwarp();
        coop_traverse<G, true, SLAB>(sc, cp, S, lane);
        if (valid && S.hits[lane].x != 0.f)
            st.L[slot] = rb.hit[slot];
    }
}

template <int G, bool SLAB>
__global__ void __launch_bounds__(kCoopBlock)
    k_raygen_coop(DevScene sc, CoopParams cp, CameraPOD cam, PathState st, BatchParams bp, uint32_t *q_out, int *q_count, int *work) {
    extern __shared__ __align__(128) unsigned char coop_smem[];
    const CoopShared S = coop_setup(coop_smem, sc, cp, G, SLAB);
    const int lane = threadIdx.x & 31;
    const int ty0 = bp.row0 >> 2;
    const int tiles_x = (bp.W + 7) >> 3, tiles_y = ((bp.row1 + 3) >> 2) - ty0, tiles = tiles_x * tiles_y;
    const int total = tiles * bp.S;
    for (int task = warp_next_chunk(work, lane); task < total; task = warp_next_chunk(work, lane)) {
        const int s = task / tiles, tile = task - s * tiles;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int px = (tx << 3) + (lane & 7), py = ((ty + ty0) << 2) + (lane >> 3);
        const bool inside = px < bp.W && py >= bp.row0 && py < bp.row1;
        bool hit_any = false;
        uint32_t slot = 0;
        V3 o{0, 0, 0}, d{0, 0, 1};
        int start = 0;
        if (inside) {
            const int gidx = px + py * bp.W;
            slot = uint32_t(s) * uint32_t(bp.P) + uint32_t(gidx);
            Rng rng;
            rng.seed(gidx, bp.seed_offset + (bp.first_accum + s) * kSeedScaler);
            camera_ray(cam, px, py, rng, o, d);
            st.thp[slot] = make_float4(1, 1, 1, 1);
            st.pdf[slot] = 1.f;
            st.rng[slot] = make_uint2(rng.d0, rng.d1);
            if (bp.rules)
                st.lobes[slot] = 0u;
            if (bp.rules == 2)
                st.vstate[slot] = make_uint2(BankStack::make(sc.cam_vol).bits, 0u);
        }
        float4 h0 = make_float4(kMaxDist, 0.f, 0.f, __int_as_float(-1));
        bool coop_ray = inside;
        const V3 inv_d = rcp3(d), o_div = o * inv_d;
        if (inside && !coop_ray_ok(inv_d)) {
            const Hit hh = closest_hit_ordered<true, false>(sc, S.cached, o, d);
            h0 = make_float4(hh.t, hh.u, hh.v, __int_as_float(hh.prim));
            coop_ray = false;
        }
        if (coop_ray)
            start = cached_entry(S.cached, sc.num_cache, inv_d, o_div, kMaxDist, 0);
        coop_store_ray(S, lane, o, d, inv_d, o_div, start, coop_ray ? kMaxDist : -1.f);
        S.hits[lane] = h0;
        if (lane == 0)
            *S.next = 0;
        __syncwarp();
        coop_traverse<G, false, SLAB>(sc, cp, S, lane);
        if (inside) {
            const float4 h = S.hits[lane];
            const int prim = __float_as_int(h.w);
            uint32_t tag = 0;
            V4 L = mk4(0, 0, 0, 1);
            if (prim >= 0) {
                tag = TAG_ACTIVE | TAG_HIT | uint32_t(prim);
                st.isect[slot] = make_interaction(sc, prim, h.y, h.z);
                hit_any = true;
            } else if (bp.rules) {
                L += mk4(1, 1, 1, 1) * null_or_env(sc, d);
            } else {
                L = null_or_env(sc, d);
            }
            st.L[slot] = to_f4(L);
            st.ray_o[slot] = make_float4(o.x, o.y, o.z, h.x);
            st.ray_d[slot] = make_float4(d.x, d.y, d.z, __uint_as_float(tag));
        }
        warp_push(q_out, q_count, hit_any, slot, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// k_vpt — volumetric path tracer: the reference's render_vpt_kernel (src/pt_impl/megakernel_vpt.cu:203-456) and its
// transmittance march occlusion_transmittance_estimate (:104-201) on this library's traversal, BSDF, emitter and
// medium code.  One path per thread, whole path inside the kernel (the reference's structure; wavefront-ising the
// medium events is the follow-up), S samples in flight, radiance to the path pool's L for k_splat.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__device__ V4 nee_transmittance(const DevScene &sc, const uint4 *s_cached, V3 o, V3 d, BankStack nested, float max_dist, Rng &rng,
                                unsigned int &segments) {
    float total_dist = 0;
    V4 Tr = splat4(1.f);
    while (total_dist < max_dist) {
        const float min_dist0 = max_dist - total_dist;
        const Hit h = closest_hit_bounded<MODE>(sc, s_cached, o, d, min_dist0);
        ++segments;
        const float min_dist = h.prim >= 0 ? h.t : min_dist0;
        const int info = h.prim >= 0 ? h.obj_word : int(0xffffffffu); // INVALID_OBJ
        const bool is_alpha = (info & 0x40000000) == 0x40000000;
        int active_medium = (info >> 20) & 0xff;
        if (active_medium == 0 || !is_alpha) { // opaque surface in the way
            Tr = splat4(0.f);
            break;
        }
        const bool is_in_medium = h.prim >= 0 && dot(interp_normal(sc, max(h.prim, 0), h.u, h.v), d) > 0;
        if (is_in_medium) {
            active_medium = nested.pop();
        } else {
            const int old_active = nested.top();
            nested.push(active_medium);
            active_medium = old_active;
        }
        const MediumPOD &am = sc.media[active_medium];
        // grid media: Medium::transmittance(ray, sp, dist) ignores `dist`; the march's ray keeps hit_t = MAX_DIST
        Tr *= am.type == MEDIUM_GRID ? grid_transmittance(sc.grids[am.grid], o, d, kMaxDist, rng) : medium_transmittance(am, min_dist);
        total_dist += min_dist;
        o = advance(o, d, min_dist);
    }
    return Tr;
}

template <int MODE>
__global__ void __launch_bounds__(kBlock, VPT_MIN_BLOCKS) k_vpt(DevScene sc, CameraPOD cam, PathState st, BatchParams bp, int *work,
                                                 unsigned long long *stats) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    const int lane = threadIdx.x & 31;
    const int ty0 = bp.row0 >> 2;
    const int tiles_x = (bp.W + 7) >> 3, tiles_y = ((bp.row1 + 3) >> 2) - ty0, tiles = tiles_x * tiles_y;
    const int total = tiles * bp.S;
    unsigned int n_closest = 0, n_shadow = 0, n_shaded = 0, n_segments = 0;
    for (int task = warp_next_chunk(work, lane); task < total; task = warp_next_chunk(work, lane)) {
        const int s = task / tiles, tile = task - s * tiles;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int px = (tx << 3) + (lane & 7), py = ((ty + ty0) << 2) + (lane >> 3);
        if (!(px < bp.W && py >= bp.row0 && py < bp.row1))
            continue;
        const int gidx = px + py * bp.W;
        const uint32_t slot = uint32_t(s) * uint32_t(bp.P) + uint32_t(gidx);
        Rng rng;
        rng.seed(gidx, bp.seed_offset + (bp.first_accum + s) * kSeedScaler);
        BankStack nested = BankStack::make(sc.cam_vol);
        V3 o, d;
        camera_ray(cam, px, py, rng, o, d);
        int diff_b = 0, spec_b = 0, trans_b = 0, volm_b = 0;
        V4 throughput = mk4(1, 1, 1), radiance = mk4(0, 0, 0);
        float emission_weight = 1.f;
        bool delta = false;
        for (int b = 0; b < bp.max_depth; b++) {
            // ---- step 1: closest hit
            const Hit h = closest_hit<MODE>(sc, s_cached, o, d, true);
            if (b > 0)
                ++n_closest;
            if (h.prim < 0) { // definitely not inside a volume
                radiance += throughput * null_or_env(sc, d);
                break;
            }
            const int obj_info = h.obj_word;
            const int hit_med_idx = (obj_info >> 20) & 0xff, object_id = obj_info & 0x000fffff;
            const ObjectPOD ob = sc.objects[object_id];
            const int material_id = ob.bsdf_emitter >> 16, emitter_id = ((ob.bsdf_emitter & 0xFFFF) ^ 0x8000) - 0x8000;
            const bool hit_emitter = emitter_id > 0;
            // ---- medium event or surface event
            const MediumPOD &medium = sc.media[nested.top()];
            const MediumSample md =
                medium.type == MEDIUM_GRID ? grid_sample(sc.grids[medium.grid], medium, o, d, h.t, rng) : medium_sample(medium, h.t, rng);
            throughput *= md.local_thp;
            const bool alpha_mask = (obj_info & 0x40000000) == 0x40000000;
            if (alpha_mask && md.flag == 0) {
                // reached a cullable medium boundary: step through it, no NEE, the bounce counter still advances
                o = advance(o, d, md.dist);
                const float4 it = make_interaction(sc, h.prim, h.u, h.v);
                const bool same_hemisphere = dot(d, xyz(it)) > 0;
                if (same_hemisphere) // leaving the medium
                    nested.pop();
                else
                    nested.push(hit_med_idx);
                continue;
            }
            ++n_shaded;
            V4 direct_comp = splat4(0, 1), nee_tr = splat4(0, 1);
            // ---- sample an emitter (the hit emitter is excluded only for surface events)
            float direct_pdf = 1;
            const int no_sample = md.flag > 0 ? 0 : emitter_id;
            int num = sc.num_emitters;
            num -= no_sample > 0 && num > 1;
            uint32_t emit_id = (rng.next_u32() % uint32_t(num)) + 1;
            emit_id += emit_id >= uint32_t(no_sample) && no_sample > 0;
            direct_pdf = 1.f / float(num);
            emit_id = emit_id * uint32_t(no_sample == 0 || num > 1);
            const EmitterPOD &em = sc.emitters[emit_id];
            const ObjectPOD eo = sc.objects[em.obj_ref * (em.obj_ref >= 0)];
            direct_pdf *= eo.inv_area;
            int emit_prim = int(rng.next_u32() % uint32_t(eo.prim_num)) + eo.prim_offset;
            emit_prim = __ldg(sc.emitter_prims + emit_prim);
            const V3 so = advance(o, d, md.dist);
            V3 sd = emitter_sample(sc, em, so, V3{0, 0, 1}, direct_comp, direct_pdf, rng.next2d(), emit_prim) - so;
            float emit_len_mis = length(sd);
            sd *= __frcp_rn(emit_len_mis);
            if (emit_id != 0) {
                ++n_shadow;
                nee_tr = nee_transmittance<MODE>(sc, s_cached, so, sd, nested, emit_len_mis - kEps, rng, n_segments);
            }
            int lobe = LOBE_NONE;
            V4 nee_thp;
            if (md.flag > 0) { // medium event
                // Medium::query_emission: homogeneous media do not emit; grid media look up the blackbody table
                radiance += throughput * (medium.type == MEDIUM_GRID ? grid_emission(sc.grids[medium.grid], sc.blackbody, so, rng) : splat4(0, 1));
                const float phase_pdf = phase_eval(medium, sd, d);
                nee_thp = splat4(phase_pdf, 1.f);
                emit_len_mis = direct_pdf + phase_pdf;
                lobe = LOBE_VOLUME;
                d = medium_scatter(medium, d, throughput, rng, emission_weight);
                delta = false;
            } else {
                const float4 it = make_interaction(sc, h.prim, h.u, h.v);
                const Material &mat = sc.mats[material_id];
                emission_weight = emission_weight / (emission_weight + solid_angle_pdf(ob.inv_area, shading_normal_world(mat, it), d, h.t) *
                                                                           hit_emitter * (b > 0) * (!delta));
                { // (throughput * Le) * w + radiance with the last product fused, as in the reference's SASS
                    const V4 tl = throughput * emitter_eval(sc.emitters[emitter_id], d, it, true);
                    radiance = V4{fmaf(tl.x, emission_weight, radiance.x), fmaf(tl.y, emission_weight, radiance.y),
                                  fmaf(tl.z, emission_weight, radiance.z), fmaf(tl.w, emission_weight, radiance.w)};
                }
                emit_len_mis = direct_pdf + bsdf_pdf(sc, mat, it, sd, d) * float(em.is_sphere != 0);
                nee_thp = bsdf_eval(sc, mat, it, sd, d);
                const bool same_hemisphere = dot(d, xyz(it)) > 0;
                d = bsdf_sample(sc, mat, d, it, throughput, emission_weight, rng, lobe);
                delta = (lobe & LOBE_SPECULAR) > 0;
                // crossing the interface of a bounded medium: incident and scattered direction on opposite sides
                if (((dot(xyz(it), d) > 0) ^ same_hemisphere) == false) {
                    if (same_hemisphere)
                        nested.pop();
                    else
                        nested.push(hit_med_idx);
                }
            }
            radiance += nee_tr * throughput * direct_comp *
                        (nee_thp * float(emit_len_mis > kEps) * __frcp_rn(emit_len_mis < kEps ? 1.f : emit_len_mis));
            o = so;
            if (bad_rgb(radiance))
                radiance = splat4(0);
            diff_b += (lobe & LOBE_DIFFUSE) > 0;
            spec_b += (lobe & LOBE_SPECULAR) > 0;
            trans_b += (lobe & LOBE_TRANSMIT) > 0;
            volm_b += (lobe & LOBE_VOLUME) > 0;
            if (diff_b >= bp.max_diffuse || spec_b >= bp.max_specular || trans_b >= bp.max_transmit || volm_b >= bp.max_volume)
                break;
            const float max_value = max3(throughput);
            if (max_value < kThpEps)
                break;
            if (b >= kRRBounce && max_value < kRRThreshold) {
                if (rng.next1d() > max_value)
                    break;
                throughput *= float(1. / max_value);
            }
        }
        st.L[slot] = to_f4(radiance);
    }
    for (int off = 16; off > 0; off >>= 1) {
        n_closest += __shfl_down_sync(0xffffffffu, n_closest, off);
        n_shadow += __shfl_down_sync(0xffffffffu, n_shadow, off);
        n_shaded += __shfl_down_sync(0xffffffffu, n_shaded, off);
        n_segments += __shfl_down_sync(0xffffffffu, n_segments, off);
    }
    if (lane == 0) {
        atomicAdd(stats + 0, (unsigned long long)n_closest);
        atomicAdd(stats + 1, (unsigned long long)n_shadow);
        atomicAdd(stats + 2, (unsigned long long)n_shaded);
        atomicAdd(stats + 3, (unsigned long long)n_segments);
    }
}


// ------------------------------------------------------------------------------------------------
// Volumetric path, wavefront form: the loop body of render_vpt_kernel (megakernel_vpt.cu:242-452) cut
// at its two traversal points, so that every kernel runs one kind of work over a compacted queue:
//   k_raygen / k_trace      closest hit of bounce b (miss: environment, path ends)                     [traversal]
//   k_vpt_pre               distance sampling in the current medium (delta tracking for grids); cullable boundary
//                           reached -> step through and go straight to the next trace; otherwise pick the emitter and
//                           the point on it, build the shadow ray                                       [no traversal]
//   k_vpt_march             transmittance march along the shadow ray (bounded closest hits + medium
//                           transmittance / ratio tracking)                                             [traversal]
//   k_vpt_post              medium event: emission, phase function; surface event: emitter hit with MIS, BSDF; NEE
//                           contribution, lobe caps, throughput cut-off, Russian roulette                [no traversal]
// Each path draws from its own sampler in the megakernel's order (distance sampling, emitter pick, march, emission
// jitter, scattering, roulette), so the frames are the megakernel's frames bit for bit.
// ------------------------------------------------------------------------------------------------
struct VptBufs {
    float4 *sh_o;  // shadow origin = event position, emit_len_mis (distance to the sampled emitter point)
    float4 *sh_d;  // shadow direction, direct_pdf
    float4 *dcomp; // emitter sample's radiance term (direct_comp)
    float4 *tr;    // result of the march
};

__global__ void __launch_bounds__(kBlock, VPTW_SHADE_MIN_BLOCKS)
    k_vpt_pre(DevScene sc, PathState st, VptBufs vb, const uint32_t *__restrict__ queue, const int *__restrict__ q_count, int *work,
              uint32_t *q_shade, int *q_shade_count, uint32_t *q_march, int *q_march_count, uint32_t *q_next, int *q_next_count) {
    const int lane = threadIdx.x & 31;
    const int count = *q_count;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int i = chunk * 32 + lane;
        bool to_shade = false, to_march = false, to_next = false;
        uint32_t slot = 0;
        if (i < count) {
            slot = queue[i];
            const float4 ro = st.ray_o[slot], rd = st.ray_d[slot];
            V3 o = xyz(ro);
            const V3 d = xyz(rd);
            const float hit_t = ro.w;
            const uint32_t tag = __float_as_uint(rd.w);
            const int prim = int(tag & TAG_INDEX);
            const int obj_info = __ldg(sc.obj_idxs + prim);
            const int hit_med_idx = (obj_info >> 20) & 0xff, object_id = obj_info & 0x000fffff;
            uint2 vs = st.vstate[slot];
            BankStack nested{vs.x};
            const uint2 rs = st.rng[slot];
            Rng rng{rs.x, rs.y};
            V4 throughput = from_f4(st.thp[slot]);
            const MediumPOD &medium = sc.media[nested.top()];
            const MediumSample md =
                medium.type == MEDIUM_GRID ? grid_sample(sc.grids[medium.grid], medium, o, d, hit_t, rng) : medium_sample(medium, hit_t, rng);
            throughput *= md.local_thp;
            st.thp[slot] = to_f4(throughput);
            const bool alpha_mask = (obj_info & 0x40000000) == 0x40000000;
            if (alpha_mask && md.flag == 0) { // cullable boundary reached: step through it, no NEE, next bounce
                o = advance(o, d, md.dist);
                const float4 it = st.isect[slot];
                if (dot(d, xyz(it)) > 0)
                    nested.pop();
                else
                    nested.push(hit_med_idx);
                st.ray_o[slot] = make_float4(o.x, o.y, o.z, hit_t);
                st.vstate[slot] = make_uint2(nested.bits, 0u);
                st.rng[slot] = make_uint2(rng.d0, rng.d1);
                to_next = true;
            } else {
                const ObjectPOD ob = sc.objects[object_id];
                const int emitter_id = ((ob.bsdf_emitter & 0xFFFF) ^ 0x8000) - 0x8000;
                V4 direct_comp = splat4(0, 1);
                float direct_pdf = 1;
                const int no_sample = md.flag > 0 ? 0 : emitter_id; // the hit emitter is excluded only for surface events
                int num = sc.num_emitters;
                num -= no_sample > 0 && num > 1;
                uint32_t emit_id = (rng.next_u32() % uint32_t(num)) + 1;
                emit_id += emit_id >= uint32_t(no_sample) && no_sample > 0;
                direct_pdf = 1.f / float(num);
                emit_id = emit_id * uint32_t(no_sample == 0 || num > 1);
                const EmitterPOD &em = sc.emitters[emit_id];
                const ObjectPOD eo = sc.objects[em.obj_ref * (em.obj_ref >= 0)];
                direct_pdf *= eo.inv_area;
                int emit_prim = int(rng.next_u32() % uint32_t(eo.prim_num)) + eo.prim_offset;
                emit_prim = __ldg(sc.emitter_prims + emit_prim);
                const V3 so = advance(o, d, md.dist);
                V3 sd = emitter_sample(sc, em, so, V3{0, 0, 1}, direct_comp, direct_pdf, rng.next2d(), emit_prim) - so;
                const float emit_len_mis = length(sd);
                sd *= __frcp_rn(emit_len_mis);
                vb.sh_o[slot] = make_float4(so.x, so.y, so.z, emit_len_mis);
                vb.sh_d[slot] = make_float4(sd.x, sd.y, sd.z, direct_pdf);
                vb.dcomp[slot] = to_f4(direct_comp);
                st.vstate[slot] = make_uint2(nested.bits, emit_id | (md.flag << 8));
                st.rng[slot] = make_uint2(rng.d0, rng.d1);
                to_shade = true;
                to_march = emit_id != 0;
            }
        }
        warp_push(q_next, q_next_count, to_next, slot, lane);
        warp_push(q_shade, q_shade_count, to_shade, slot, lane);
        warp_push(q_march, q_march_count, to_march, slot, lane);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kBlock, VPTW_MARCH_MIN_BLOCKS) k_vpt_march(DevScene sc, PathState st, VptBufs vb, const uint32_t *__restrict__ queue,
                                                      const int *__restrict__ q_count, int *work, unsigned long long *stats) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    const int lane = threadIdx.x & 31;
    const int count = *q_count;
    unsigned int n_segments = 0;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int i = chunk * 32 + lane;
        if (i < count) {
            const uint32_t slot = queue[i];
            const float4 a = vb.sh_o[slot], b = vb.sh_d[slot];
            const uint2 rs = st.rng[slot];
            Rng rng{rs.x, rs.y};
            const V4 tr = nee_transmittance<MODE>(sc, s_cached, xyz(a), xyz(b), BankStack{st.vstate[slot].x}, a.w - kEps, rng, n_segments);
            vb.tr[slot] = to_f4(tr);
            st.rng[slot] = make_uint2(rng.d0, rng.d1);
        }
    }
    for (int off = 16; off > 0; off >>= 1)
        n_segments += __shfl_down_sync(0xffffffffu, n_segments, off);
    if (lane == 0 && n_segments)
        atomicAdd(stats + 3, (unsigned long long)n_segments);
}

__global__ void __launch_bounds__(kBlock, VPTW_SHADE_MIN_BLOCKS)
    k_vpt_post(DevScene sc, PathState st, VptBufs vb, BatchParams bp, const uint32_t *__restrict__ queue, const int *__restrict__ q_count,
               int *work, uint32_t *q_next, int *q_next_count, int b) {
    const int lane = threadIdx.x & 31;
    const int count = *q_count;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int i = chunk * 32 + lane;
        bool alive = false;
        uint32_t slot = 0;
        if (i < count) {
            slot = queue[i];
            const float4 ro = st.ray_o[slot], rd = st.ray_d[slot];
            V3 d = xyz(rd);
            const float hit_t = ro.w;
            const uint32_t tag = __float_as_uint(rd.w);
            const int prim = int(tag & TAG_INDEX);
            bool delta = (tag & TAG_DELTA) != 0;
            const int obj_info = __ldg(sc.obj_idxs + prim);
            const int hit_med_idx = (obj_info >> 20) & 0xff, object_id = obj_info & 0x000fffff;
            const ObjectPOD ob = sc.objects[object_id];
            const int material_id = ob.bsdf_emitter >> 16, emitter_id = ((ob.bsdf_emitter & 0xFFFF) ^ 0x8000) - 0x8000;
            const bool hit_emitter = emitter_id > 0;
            const uint2 vs = st.vstate[slot];
            BankStack nested{vs.x};
            const uint32_t emit_id = vs.y & 0xffu;
            const bool medium_event = (vs.y >> 8) & 1u;
            const uint2 rs = st.rng[slot];
            Rng rng{rs.x, rs.y};
            V4 throughput = from_f4(st.thp[slot]), radiance = from_f4(st.L[slot]);
            float emission_weight = st.pdf[slot];
            const float4 so4 = vb.sh_o[slot], sd4 = vb.sh_d[slot];
            const V3 so = xyz(so4), sd = xyz(sd4);
            float emit_len_mis = so4.w;
            const float direct_pdf = sd4.w;
            const V4 direct_comp = from_f4(vb.dcomp[slot]);
            const V4 nee_tr = emit_id != 0 ? from_f4(vb.tr[slot]) : splat4(0, 1);
            const MediumPOD &medium = sc.media[nested.top()];
            const EmitterPOD &em = sc.emitters[emit_id];
            int lobe = LOBE_NONE;
            V4 nee_thp;
            if (medium_event) {
                radiance += throughput * (medium.type == MEDIUM_GRID ? grid_emission(sc.grids[medium.grid], sc.blackbody, so, rng) : splat4(0, 1));
                const float phase_pdf = phase_eval(medium, sd, d);
                nee_thp = splat4(phase_pdf, 1.f);
                emit_len_mis = direct_pdf + phase_pdf;
                lobe = LOBE_VOLUME;
                d = medium_scatter(medium, d, throughput, rng, emission_weight);
                delta = false;
            } else {
                const float4 it = st.isect[slot];
                const Material &mat = sc.mats[material_id];
                emission_weight = emission_weight / (emission_weight + solid_angle_pdf(ob.inv_area, shading_normal_world(mat, it), d, hit_t) *
                                                                           hit_emitter * (b > 0) * (!delta));
                {
                    const V4 tl = throughput * emitter_eval(sc.emitters[emitter_id], d, it, true);
                    radiance = V4{fmaf(tl.x, emission_weight, radiance.x), fmaf(tl.y, emission_weight, radiance.y),
                                  fmaf(tl.z, emission_weight, radiance.z), fmaf(tl.w, emission_weight, radiance.w)};
                }
                emit_len_mis = direct_pdf + bsdf_pdf(sc, mat, it, sd, d) * float(em.is_sphere != 0);
                nee_thp = bsdf_eval(sc, mat, it, sd, d);
                const bool same_hemisphere = dot(d, xyz(it)) > 0;
                d = bsdf_sample(sc, mat, d, it, throughput, emission_weight, rng, lobe);
                delta = (lobe & LOBE_SPECULAR) > 0;
                if (((dot(xyz(it), d) > 0) ^ same_hemisphere) == false) {
                    if (same_hemisphere)
                        nested.pop();
                    else
                        nested.push(hit_med_idx);
                }
            }
            radiance += nee_tr * throughput * direct_comp *
                        (nee_thp * float(emit_len_mis > kEps) * __frcp_rn(emit_len_mis < kEps ? 1.f : emit_len_mis));
            if (bad_rgb(radiance))
                radiance = splat4(0);
            uint32_t c = st.lobes[slot];
            c += uint32_t((lobe & LOBE_DIFFUSE) > 0) | (uint32_t((lobe & LOBE_SPECULAR) > 0) << 8) | (uint32_t((lobe & LOBE_TRANSMIT) > 0) << 16) |
                 (uint32_t((lobe & LOBE_VOLUME) > 0) << 24);
            alive = !(int(c & 0xff) >= bp.max_diffuse || int((c >> 8) & 0xff) >= bp.max_specular || int((c >> 16) & 0xff) >= bp.max_transmit ||
                      int(c >> 24) >= bp.max_volume);
            if (alive) {
                const float max_value = max3(throughput);
                if (max_value < kThpEps) {
                    alive = false;
                } else if (b >= kRRBounce && max_value < kRRThreshold) {
                    if (rng.next1d() > max_value)
                        alive = false;
                    else
                        throughput *= float(1. / max_value);
                }
            }
            st.L[slot] = to_f4(radiance);
            if (alive) {
                st.lobes[slot] = c;
                st.thp[slot] = to_f4(throughput);
                st.pdf[slot] = emission_weight;
                st.rng[slot] = make_uint2(rng.d0, rng.d1);
                st.vstate[slot] = make_uint2(nested.bits, 0u);
                st.ray_o[slot] = make_float4(so.x, so.y, so.z, hit_t);
                st.ray_d[slot] = make_float4(d.x, d.y, d.z, __uint_as_float(TAG_ACTIVE | (delta ? TAG_DELTA : 0u)));
            }
        }
        warp_push(q_next, q_next_count, alive, slot, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// Volumetric wavefront, dense form (default): same cuts as above, but the path RECORDS are kept dense instead of the
// queue of slot ids.  ncu on the slot form (profiles/r01_ncu_volumetric_summary.md): by bounce 5 only ~30 % of the slots
// are alive, every 16 B field read lands in its own sector, DRAM moves 6x the useful bytes and the traversal-free
// kernels wait on it.  Here generation g of the path arrays holds the n alive paths at indices [0, n): k_vptd_post
// writes the survivors of generation g compactly into generation g^1 (warp-aggregated allocation), every kernel
// reads and writes fully coalesced SoA, a path carries its slot (sample x pixel) in `pix` and deposits its final
// radiance into Lout[slot] when it ends.  The image is indexed by slot, each path uses its own sampler: order-free,
// so the frames stay the megakernel's (and the reference's) bit for bit.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t VK_SHADE = 0u, VK_CONTINUE = 1u, VK_DEAD = 2u; // event word bits 16-17 (k_vptd_pre -> k_vptd_post)

template <int MODE>
__global__ void __launch_bounds__(kBlock)
    k_vptd_raygen(DevScene sc, CameraPOD cam, PathState out, uint32_t *pix, float4 *Lout, BatchParams bp, int *n_out, int *work) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    const int lane = threadIdx.x & 31;
    const int ty0 = bp.row0 >> 2;
    const int tiles_x = (bp.W + 7) >> 3, tiles_y = ((bp.row1 + 3) >> 2) - ty0, tiles = tiles_x * tiles_y;
    const int total = tiles * bp.S;
    for (int task = warp_next_chunk(work, lane); task < total; task = warp_next_chunk(work, lane)) {
        const int s = task / tiles, tile = task - s * tiles;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int px = (tx << 3) + (lane & 7), py = ((ty + ty0) << 2) + (lane >> 3);
        const bool inside = px < bp.W && py >= bp.row0 && py < bp.row1;
        uint32_t slot = 0;
        V3 o{0, 0, 0}, d{0, 0, 1};
        Rng rng{0, 0};
        if (inside) {
            const int gidx = px + py * bp.W;
            slot = uint32_t(s) * uint32_t(bp.P) + uint32_t(gidx);
            rng.seed(gidx, bp.seed_offset + (bp.first_accum + s) * kSeedScaler);
            camera_ray(cam, px, py, rng, o, d);
        }
        const Hit h = closest_hit<MODE>(sc, s_cached, o, d, inside);
        const bool alive = inside && h.prim >= 0;
        const int j = warp_alloc(n_out, alive, lane);
        if (alive) {
            out.ray_o[j] = make_float4(o.x, o.y, o.z, h.t);
            out.ray_d[j] = make_float4(d.x, d.y, d.z, __uint_as_float(TAG_ACTIVE | TAG_HIT | uint32_t(h.prim)));
            out.thp[j] = make_float4(1, 1, 1, 1);
            out.L[j] = make_float4(0, 0, 0, 1);
            out.isect[j] = make_interaction(sc, h.prim, h.u, h.v);
            out.rng[j] = make_uint2(rng.d0, rng.d1);
            out.pdf[j] = 1.f;
            if (bp.rules)
                out.lobes[j] = 0u;
            if (bp.rules == 2)
                out.vstate[j] = make_uint2(BankStack::make(sc.cam_vol).bits, 0u);
            pix[j] = slot;
        } else if (inside) {
            V4 L = mk4(0, 0, 0, 1);
            if (bp.rules)
                L += mk4(1, 1, 1, 1) * null_or_env(sc, d); // megakernel_pt.cu:181-184: radiance(0,0,0,1) += throughput * Le
            else
                L = null_or_env(sc, d);
            Lout[slot] = to_f4(L);
        }
    }
}

__global__ void __launch_bounds__(kBlock, VPTW_SHADE_MIN_BLOCKS)
    k_vptd_pre(DevScene sc, PathState st, VptBufs vb, const uint32_t *__restrict__ queue, const int *__restrict__ n_ptr, int *work,
               uint32_t *q_march, int *n_march, unsigned long long *stats) {
    const int lane = threadIdx.x & 31;
    const int count = *n_ptr;
    unsigned int n_shaded = 0;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int k = chunk * 32 + lane;
        // queue: indices of the hits in the order k_vptd_trace's warps finished (nullptr at bounce 0: identity)
        const int i = (queue && k < count) ? int(queue[k]) : k;
        bool to_march = false;
        if (k < count) {
            const float4 ro = st.ray_o[i], rd = st.ray_d[i];
            const uint32_t tag = __float_as_uint(rd.w);
            uint2 vs = st.vstate[i];
            if (!(tag & TAG_ACTIVE)) { // ended at the trace (miss): its radiance is already deposited
                st.vstate[i] = make_uint2(vs.x, VK_DEAD << 16);
            } else {
                V3 o = xyz(ro);
                const V3 d = xyz(rd);
                const float hit_t = ro.w;
                const int prim = int(tag & TAG_INDEX);
                const int obj_info = __ldg(sc.obj_idxs + prim);
                const int hit_med_idx = (obj_info >> 20) & 0xff, object_id = obj_info & 0x000fffff;
                BankStack nested{vs.x};
                const uint2 rs = st.rng[i];
                Rng rng{rs.x, rs.y};
                V4 throughput = from_f4(st.thp[i]);
                const MediumPOD &medium = sc.media[nested.top()];
                const MediumSample md = medium.type == MEDIUM_GRID ? grid_sample(sc.grids[medium.grid], medium, o, d, hit_t, rng)
                                                                   : medium_sample(medium, hit_t, rng);
                throughput *= md.local_thp;
                st.thp[i] = to_f4(throughput);
                const bool alpha_mask = (obj_info & 0x40000000) == 0x40000000;
                if (alpha_mask && md.flag == 0) { // cullable boundary reached: step through it, no NEE, next bounce
                    o = advance(o, d, md.dist);
                    const float4 it = st.isect[i];
                    if (dot(d, xyz(it)) > 0)
                        nested.pop();
                    else
                        nested.push(hit_med_idx);
                    st.ray_o[i] = make_float4(o.x, o.y, o.z, hit_t);
                    st.vstate[i] = make_uint2(nested.bits, VK_CONTINUE << 16);
                } else {
                    ++n_shaded;
                    const ObjectPOD ob = sc.objects[object_id];
                    const int emitter_id = ((ob.bsdf_emitter & 0xFFFF) ^ 0x8000) - 0x8000;
                    V4 direct_comp = splat4(0, 1);
                    float direct_pdf = 1;
                    const int no_sample = md.flag > 0 ? 0 : emitter_id;
                    int num = sc.num_emitters;
                    num -= no_sample > 0 && num > 1;
                    uint32_t emit_id = (rng.next_u32() % uint32_t(num)) + 1;
                    emit_id += emit_id >= uint32_t(no_sample) && no_sample > 0;
                    direct_pdf = 1.f / float(num);
                    emit_id = emit_id * uint32_t(no_sample == 0 || num > 1);
                    const EmitterPOD &em = sc.emitters[emit_id];
                    const ObjectPOD eo = sc.objects[em.obj_ref * (em.obj_ref >= 0)];
                    direct_pdf *= eo.inv_area;
                    int emit_prim = int(rng.next_u32() % uint32_t(eo.prim_num)) + eo.prim_offset;
                    emit_prim = __ldg(sc.emitter_prims + emit_prim);
                    const V3 so = advance(o, d, md.dist);
                    V3 sd = emitter_sample(sc, em, so, V3{0, 0, 1}, direct_comp, direct_pdf, rng.next2d(), emit_prim) - so;
                    const float emit_len_mis = length(sd);
                    sd *= __frcp_rn(emit_len_mis);
                    vb.sh_o[i] = make_float4(so.x, so.y, so.z, emit_len_mis);
                    vb.sh_d[i] = make_float4(sd.x, sd.y, sd.z, direct_pdf);
                    vb.dcomp[i] = to_f4(direct_comp);
                    st.vstate[i] = make_uint2(nested.bits, emit_id | (md.flag << 8) | (VK_SHADE << 16));
                    to_march = emit_id != 0;
                }
                st.rng[i] = make_uint2(rng.d0, rng.d1);
            }
        }
        warp_push(q_march, n_march, to_march, uint32_t(i), lane);
    }
    for (int off = 16; off > 0; off >>= 1)
        n_shaded += __shfl_down_sync(0xffffffffu, n_shaded, off);
    if (lane == 0 && n_shaded)
        atomicAdd(stats + 2, (unsigned long long)n_shaded);
}

__global__ void __launch_bounds__(kBlock, VPTW_SHADE_MIN_BLOCKS)
    k_vptd_post(DevScene sc, PathState st, PathState out, const uint32_t *__restrict__ pix, uint32_t *pix_out, VptBufs vb, float4 *Lout,
                BatchParams bp, const uint32_t *__restrict__ queue, const int *__restrict__ n_ptr, int *n_out, int *n_out_back, int cap,
                int *work, int b) {
    const int lane = threadIdx.x & 31;
    const int count = *n_ptr;
    const bool last = b + 1 >= bp.max_depth;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int k = chunk * 32 + lane;
        const int i = (queue && k < count) ? int(queue[k]) : k;
        bool alive = false, went_on = false; // went_on: stepped through a boundary, direction unchanged
        float4 n_o, n_d, n_thp, n_L, n_isect;
        uint2 n_rng, n_vs;
        float n_pdf = 0;
        uint32_t n_lobes = 0, slot = 0;
        if (k < count) {
            const uint2 vs = st.vstate[i];
            const uint32_t kind = vs.y >> 16;
            if (kind != VK_DEAD) {
                slot = pix[i];
                const float4 ro = st.ray_o[i], rd = st.ray_d[i];
                n_L = st.L[i];
                if (kind == VK_CONTINUE) { // stepped through a cullable boundary in k_vptd_pre: unchanged, on to the next trace
                    alive = went_on = true;
                    n_o = ro, n_d = rd, n_thp = st.thp[i], n_isect = st.isect[i], n_rng = st.rng[i], n_vs = make_uint2(vs.x, 0u);
                    n_pdf = st.pdf[i], n_lobes = st.lobes[i];
                } else {
                    V3 d = xyz(rd);
                    const float hit_t = ro.w;
                    const uint32_t tag = __float_as_uint(rd.w);
                    const int prim = int(tag & TAG_INDEX);
                    bool delta = (tag & TAG_DELTA) != 0;
                    const int obj_info = __ldg(sc.obj_idxs + prim);
                    const int hit_med_idx = (obj_info >> 20) & 0xff, object_id = obj_info & 0x000fffff;
                    const ObjectPOD ob = sc.objects[object_id];
                    const int material_id = ob.bsdf_emitter >> 16, emitter_id = ((ob.bsdf_emitter & 0xFFFF) ^ 0x8000) - 0x8000;
                    const bool hit_emitter = emitter_id > 0;
                    BankStack nested{vs.x};
                    const uint32_t emit_id = vs.y & 0xffu;
                    const bool medium_event = (vs.y >> 8) & 1u;
                    const uint2 rs = st.rng[i];
                    Rng rng{rs.x, rs.y};
                    V4 throughput = from_f4(st.thp[i]), radiance = from_f4(n_L);
                    float emission_weight = st.pdf[i];
                    const float4 so4 = vb.sh_o[i], sd4 = vb.sh_d[i];
                    const V3 so = xyz(so4), sd = xyz(sd4);
                    float emit_len_mis = so4.w;
                    const float direct_pdf = sd4.w;
                    const V4 direct_comp = from_f4(vb.dcomp[i]);
                    const V4 nee_tr = emit_id != 0 ? from_f4(vb.tr[i]) : splat4(0, 1);
                    const MediumPOD &medium = sc.media[nested.top()];
                    const EmitterPOD &em = sc.emitters[emit_id];
                    int lobe = LOBE_NONE;
                    V4 nee_thp;
                    if (medium_event) {
                        radiance += throughput * (medium.type == MEDIUM_GRID ? grid_emission(sc.grids[medium.grid], sc.blackbody, so, rng) : splat4(0, 1));
                        const float phase_pdf = phase_eval(medium, sd, d);
                        nee_thp = splat4(phase_pdf, 1.f);
                        emit_len_mis = direct_pdf + phase_pdf;
                        lobe = LOBE_VOLUME;
                        d = medium_scatter(medium, d, throughput, rng, emission_weight);
                        delta = false;
                        n_isect = make_float4(0, 0, 0, 0);
                    } else {
                        const float4 it = st.isect[i];
                        n_isect = it;
                        const Material &mat = sc.mats[material_id];
                        emission_weight = emission_weight / (emission_weight + solid_angle_pdf(ob.inv_area, shading_normal_world(mat, it), d, hit_t) *
                                                                                   hit_emitter * (b > 0) * (!delta));
                        {
                            const V4 tl = throughput * emitter_eval(sc.emitters[emitter_id], d, it, true);
                            radiance = V4{fmaf(tl.x, emission_weight, radiance.x), fmaf(tl.y, emission_weight, radiance.y),
                                          fmaf(tl.z, emission_weight, radiance.z), fmaf(tl.w, emission_weight, radiance.w)};
                        }
                        emit_len_mis = direct_pdf + bsdf_pdf(sc, mat, it, sd, d) * float(em.is_sphere != 0);
                        nee_thp = bsdf_eval(sc, mat, it, sd, d);
                        const bool same_hemisphere = dot(d, xyz(it)) > 0;
                        d = bsdf_sample(sc, mat, d, it, throughput, emission_weight, rng, lobe);
                        delta = (lobe & LOBE_SPECULAR) > 0;
                        if (((dot(xyz(it), d) > 0) ^ same_hemisphere) == false) {
                            if (same_hemisphere)
                                nested.pop();
                            else
                                nested.push(hit_med_idx);
                        }
                    }
                    radiance += nee_tr * throughput * direct_comp *
                                (nee_thp * float(emit_len_mis > kEps) * __frcp_rn(emit_len_mis < kEps ? 1.f : emit_len_mis));
                    if (bad_rgb(radiance))
                        radiance = splat4(0);
                    uint32_t c = st.lobes[i];
                    c += uint32_t((lobe & LOBE_DIFFUSE) > 0) | (uint32_t((lobe & LOBE_SPECULAR) > 0) << 8) |
                         (uint32_t((lobe & LOBE_TRANSMIT) > 0) << 16) | (uint32_t((lobe & LOBE_VOLUME) > 0) << 24);
                    alive = !(int(c & 0xff) >= bp.max_diffuse || int((c >> 8) & 0xff) >= bp.max_specular ||
                              int((c >> 16) & 0xff) >= bp.max_transmit || int(c >> 24) >= bp.max_volume);
                    if (alive) {
                        const float max_value = max3(throughput);
                        if (max_value < kThpEps) {
                            alive = false;
                        } else if (b >= kRRBounce && max_value < kRRThreshold) {
                            if (rng.next1d() > max_value)
                                alive = false;
                            else
                                throughput *= float(1. / max_value);
                        }
                    }
                    n_L = to_f4(radiance);
                    n_o = make_float4(so.x, so.y, so.z, hit_t);
                    n_d = make_float4(d.x, d.y, d.z, __uint_as_float(TAG_ACTIVE | (delta ? TAG_DELTA : 0u)));
                    n_thp = to_f4(throughput), n_rng = make_uint2(rng.d0, rng.d1), n_vs = make_uint2(nested.bits, 0u);
                    n_pdf = emission_weight, n_lobes = c;
                }
                if (!alive || last) { // the path ends here (or the bounce loop does): final radiance of its slot
                    Lout[slot] = n_L;
                    alive = false;
                }
            }
        }
        // unscattered continuations fill the next generation from the front, scattered survivors from the back: the next
        // trace then runs warps of coherent rays and warps of incoherent rays instead of mixing them (lanes per
        // instruction 5.5 -> 6.8 on the fog scene, which is what the slot-queue form gets from its two producers)
        const int jf = warp_alloc(n_out, alive && went_on, lane), jb = warp_alloc(n_out_back, alive && !went_on, lane);
        const int j = went_on ? jf : cap - 1 - jb;
        if (alive) {
            out.ray_o[j] = n_o, out.ray_d[j] = n_d, out.thp[j] = n_thp, out.L[j] = n_L, out.isect[j] = n_isect;
            out.rng[j] = n_rng, out.vstate[j] = n_vs, out.pdf[j] = n_pdf, out.lobes[j] = n_lobes;
            pix_out[j] = slot;
        }
    }
}

// closest hit of the dense generation, in place: hit -> interaction + hit record; miss -> environment, final radiance
template <int MODE>
__global__ void __launch_bounds__(kBlock, 12) k_vptd_trace(DevScene sc, PathState st, const uint32_t *__restrict__ pix, float4 *Lout,
                                                       const int *__restrict__ n_ptr, const int *__restrict__ n_back_ptr, int cap, int *work,
                                                       uint32_t *q_hit, int *n_hit) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    const int lane = threadIdx.x & 31;
    const int n_front = *n_ptr, count = n_front + *n_back_ptr;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int k = chunk * 32 + lane;
        const bool valid = k < count;
        const int i = k < n_front ? k : cap - 1 - (k - n_front); // front range, then back range (see k_vptd_post)
        V3 o{0, 0, 0}, d{0, 0, 1};
        uint32_t tag = 0;
        if (valid) {
            const float4 ro = st.ray_o[i], rd = st.ray_d[i];
            o = xyz(ro), d = xyz(rd);
            tag = __float_as_uint(rd.w) & (TAG_ACTIVE | TAG_DELTA);
        }
        const Hit h = closest_hit<MODE>(sc, s_cached, o, d, valid);
        if (valid) {
            if (h.prim >= 0) {
                st.isect[i] = make_interaction(sc, h.prim, h.u, h.v);
                st.ray_o[i].w = h.t;
                st.ray_d[i].w = __uint_as_float(tag | TAG_HIT | uint32_t(h.prim));
            } else {
                V4 L = from_f4(st.L[i]);
                L += from_f4(st.thp[i]) * null_or_env(sc, d);
                Lout[pix[i]] = to_f4(L);
                st.ray_d[i].w = __uint_as_float(0u);
            }
        }
        // hits in the order the warps finish: rays of similar traversal length - similar region of the scene - end up in the
        // same warps of the next kernels (measured: traversal over plain index order runs 9 % slower)
        warp_push(q_hit, n_hit, valid && h.prim >= 0, uint32_t(i), lane);
    }
}

// k_trace over a dense generation, in place (surface path, deferred-NEE pipeline): hit -> interaction, Russian roulette;
// the path ends (miss, roulette, or ended by k_shade under the megakernel rules) -> final radiance to Lout[slot]
template <int MODE>
__global__ void __launch_bounds__(kBlock, 12) k_trace_dense(DevScene sc, PathState st, const uint32_t *__restrict__ pix, float4 *Lout,
                                                        const int *__restrict__ n_ptr, int *work, int bounce, int pt_rules) {
    extern __shared__ uint4 s_cached[];
    stage_cached_nodes(sc, s_cached);
    const int lane = threadIdx.x & 31;
    const int count = *n_ptr;
    for (int chunk = warp_next_chunk(work, lane); chunk * 32 < count; chunk = warp_next_chunk(work, lane)) {
        const int i = chunk * 32 + lane;
        bool valid = i < count, ended = false;
        uint32_t tag = 0;
        V3 o{0, 0, 0}, d{0, 0, 1};
        if (valid) {
            const float4 ro = st.ray_o[i], rd = st.ray_d[i];
            o = xyz(ro), d = xyz(rd);
            tag = __float_as_uint(rd.w) & 0x50000000u;
            ended = (tag & TAG_ACTIVE) == 0; // megakernel rules: k_shade ended the path
            valid = !ended;
        }
        const Hit h = closest_hit<MODE>(sc, s_cached, o, d, valid); // all 32 lanes call
        if (valid)
            ended = !resolve_hit(sc, st, uint32_t(i), tag, d, h.t, h.u, h.v, h.prim, bounce, pt_rules);
        if (ended)
            Lout[pix[i]] = st.L[i];
    }
}

// final radiance of the paths still in flight after the last bounce
__global__ void k_deposit(PathState st, const uint32_t *__restrict__ pix, float4 *Lout, const int *__restrict__ n_ptr) {
    const int count = *n_ptr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        Lout[pix[i]] = st.L[i];
}

__global__ void k_fold_vptd_stats(const int *__restrict__ cnt_n, const int *__restrict__ cnt_b, const int *__restrict__ cnt_m, int depth,
                                  unsigned long long *stats) {
    if (threadIdx.x == 0) {
        unsigned long long closest = 0, shadow = 0;
        for (int b = 0; b < depth; b++) {
            closest += b > 0 ? (unsigned long long)cnt_n[b] + (unsigned long long)cnt_b[b] : 0ull;
            shadow += (unsigned long long)cnt_m[b];
        }
        stats[0] += closest, stats[1] += shadow;
    }
}

// ray accounting of the volumetric wavefront: closest-hit rays = paths entering a trace after bounce 0, shadow rays =
// marches, shaded = events that ran NEE
__global__ void k_fold_vpt_stats(const int *__restrict__ cnt_t, const int *__restrict__ cnt_m, const int *__restrict__ cnt_s, int depth,
                                 unsigned long long *stats) {
    if (threadIdx.x == 0) {
        unsigned long long closest = 0, shadow = 0, shaded = 0;
        for (int b = 0; b < depth; b++) {
            closest += b > 0 ? (unsigned long long)cnt_t[b] : 0ull;
            shadow += (unsigned long long)cnt_m[b];
            shaded += (unsigned long long)cnt_s[b];
        }
        stats[0] += closest, stats[1] += shadow, stats[2] += shaded;
    }
}

// ------------------------------------------------------------------------------------------------
// k_splat — reference: radiance_splat (wavefront_pt.cu:476-503), estimate_variance
// (tracing_func.cuh:194-208), Vec4::gamma_corr (vec4.cuh:213-217)
// samples of a batch are added in sample order, so the fp32 sum is the one the reference produces
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_splat(const float4 *__restrict__ Lbuf, float4 *__restrict__ image,
                                               float4 *__restrict__ out, float *__restrict__ var, int P, int S,
                                               int first_accum, int write_output, int gamma, int scrub, int p_begin,
                                               int p_end) {
    for (int p = p_begin + blockIdx.x * blockDim.x + threadIdx.x; p < p_end; p += gridDim.x * blockDim.x) {
        V4 acc = from_f4(image[p]);
        float variance = var ? var[p] : 0.f;
        for (int s = 0; s < S; s++) {
            V4 L = from_f4(Lbuf[size_t(s) * P + p]);
            L = (scrub && bad_rgb(L)) ? mk4(0, 0, 0, 1) : L; // wavefront_pt.cu:487 (the megakernel scrubs per bounce instead)
            if (var) {
                const int accum_cnt = first_accum + s;
                // rounding points pinned to the reference kernel's SASS: only `cur - new_sum * rcp` and
                // `(n-1) * var + prod` are fused there
                const float cur_val = __fmul_rn(__fadd_rn(__fadd_rn(L.x, L.y), L.z), 0.33333333f);
                float old_mean = __fmul_rn(__fadd_rn(__fadd_rn(acc.x, acc.y), acc.z), 0.33333333f);
                const float new_sum = __fadd_rn(cur_val, old_mean), rcp_cnt = 1.f / float(accum_cnt);
                old_mean = accum_cnt > 1 ? __fdividef(old_mean, float(accum_cnt - 1)) : 0.f;
                const float d_new = fmaf(new_sum, -rcp_cnt, cur_val), d_old = __fsub_rn(cur_val, old_mean);
                variance = __fmul_rn(fmaf(float(accum_cnt - 1), variance, __fmul_rn(d_old, d_new)), rcp_cnt);
            }
            acc += L;
        }
        image[p] = to_f4(acc);
        if (var)
            var[p] = variance;
        if (write_output) {
            V4 o = acc;
            o *= 1.f / float(first_accum + S - 1);
            if (gamma)
                o = mk4(powf(fmaxf(0.f, o.x), 1.f / 2.2f), powf(fmaxf(0.f, o.y), 1.f / 2.2f),
                        powf(fmaxf(0.f, o.z), 1.f / 2.2f));
            out[p] = to_f4(o);
        }
    }
}

// folds the per-bounce queue sizes of a finished batch into the cumulative ray statistics
__global__ void k_fold_stats(const int *__restrict__ q_counts, int traced_bounces, int shaded_bounces,
                             unsigned long long *stats, const int *__restrict__ sq_counts, int shadow_queues) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long queued = 0;
        for (int b = 0; b < shadow_queues; b++)
            queued += (unsigned long long)sq_counts[b];
        stats[3] += queued;
        unsigned long long closest = 0, shaded = 0;
        for (int b = 0; b < shaded_bounces; b++)
            shaded += (unsigned long long)q_counts[b];
        for (int b = 0; b < traced_bounces; b++)
            closest += (unsigned long long)q_counts[b];
        stats[0] += closest;
        stats[2] += shaded;
    }
}

// derived material / emitter parameters are computed on the device with the same math functions the
// reference uses inside its device-side constructors (virtual_funcs.cu:85-98, bsdf_plastic.cu:25-36,
// fresnel.cuh:62-75, cuda_utils.cuh:115-120, emitter.cuh:349-353)
// Hamilton product with the rounding points of the reference's create_envmap_source kernel (quaternion.cuh:54-60):
// fma(a, b, +-(c * d)) for the first two terms of a component, then one fma per remaining term
struct Quat {
    float x, y, z, w;
};
__device__ Quat qmul(Quat a, Quat q) {
    return Quat{fmaf(-a.z, q.y, fmaf(a.y, q.z, fmaf(a.w, q.x, __fmul_rn(a.x, q.w)))),
                fmaf(a.z, q.x, fmaf(a.y, q.w, fmaf(a.w, q.y, -__fmul_rn(a.x, q.z)))),
                fmaf(a.z, q.w, fmaf(-a.y, q.x, fmaf(a.w, q.z, __fmul_rn(a.x, q.y)))),
                fmaf(-a.z, q.z, fmaf(-a.y, q.y, fmaf(a.w, q.w, -__fmul_rn(a.x, q.x))))};
}

__device__ float roughness_to_alpha_dev(float roughness) {
    roughness = fmaxf(roughness, 1e-3f);
    float x = logf(roughness);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
__device__ float diffuse_fresnel_dev(float ior) {
    const int sample_cnt = 131072;
    double diff_fresnel = 0.0;
    float fb = fresnel_simple(ior, 0.0f);
    for (int i = 1; i <= sample_cnt; i++) {
        float cos_theta2 = float(i) / sample_cnt;
        float fa = fresnel_simple(ior, min(sqrtf(cos_theta2), 1.0f));
        diff_fresnel += double(fa + fb) * (0.5f / sample_cnt);
        fb = fa;
    }
    return float(diff_fresnel);
}
__global__ void k_derive_tables(Material *mats, int num_mats, EmitterPOD *ems, int num_ems) {
    const int i = threadIdx.x;
    if (i < num_mats) {
        Material &m = mats[i];
        if (m.type == MAT_GGX) {
            m.k_d = make_float4(0, 0, 0, 0);
            m.k_s = make_float4(roughness_to_alpha_dev(m.raw[0]), roughness_to_alpha_dev(m.raw[1]), 1, 1);
        } else if (m.type == MAT_PLASTIC || m.type == MAT_PLASTIC_FWD) {
            m.eta = 1.f / m.raw[0];
            m.precomp_diff_f = m.type == MAT_PLASTIC ? diffuse_fresnel_dev(m.raw[0]) : 0.f;
        }
    }
    if (i < num_ems && ems[i].type == EM_ENVMAP) {
        EmitterPOD &e = ems[i];
        float s, c;
        sincosf(e.azimuth * 0.5f, &s, &c);
        Quat yaw{0 * s, 0 * s, 1 * s, c};
        sincosf(e.zenith * 0.5f, &s, &c);
        Quat pit{1 * s, 0 * s, 0 * s, c};
        Quat r = qmul(yaw, pit);
        e.rot = make_float4(r.x, r.y, r.z, r.w);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    void alloc(size_t count) {
        release();
        n = count;
        if (count)
            CUDA_OK(cudaMalloc(&p, count * sizeof(T)));
    }
    void upload(const std::vector<T> &v) {
        alloc(v.size());
        if (!v.empty())
            CUDA_OK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
    void release() {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

struct Renderer::Impl {
    DevBuf<float4> nodes, norms, arena; // arena = [pair nodes | wide nodes | triangles]
    size_t arena_nodes = 0, arena_wide = 0, l2_window_bytes = 0;
    int wide_width = 4, wide_nodes = 0, wide_max_stack = 0, wide_max_stack_all = 0, wide_depth = 0;
    double wide_fill = 0;
    int coop_blocks_per_sm = 0; // occupancy of the cooperative kernels with the current shared-memory carve-up (smallest of the three)
    int coop_blocks[3] = {0, 0, 0}; // k_trace_coop, k_shadow_apply_coop, k_raygen_coop: CTAs per SM each
    size_t coop_smem = 0;
    DevBuf<uint4> cache_nodes, uvs;
    DevBuf<int> obj_idxs, emitter_prims, prim_orig;
    int refined_leaves = 0;
    DevBuf<ObjectPOD> objects;
    DevBuf<Material> mats;
    DevBuf<MediumPOD> media;
    DevBuf<GridPOD> grids;
    DevBuf<int> grid_bricks;
    DevBuf<float> grid_data;
    int max_volume = 0;
    DevBuf<EmitterPOD> emitters;
    DevScene dev{};
    CameraPOD cam{};
    // path pool
    DevBuf<float4> ray_o, ray_d, thp, L, isect;
    DevBuf<uint2> rng;
    DevBuf<float> pdf;
    DevBuf<uint32_t> lobes;
    int caps[3] = {16, 16, 16};
    DevBuf<uint32_t> queue[2], shadow_queue;
    DevBuf<float4> sh_o, sh_d, hit; // dynamic-fetch pipeline: shadow rays and hit records
    DevBuf<unsigned char> vis;      // stream traversal: one visibility byte per shadow-queue entry
    int stream_blocks[2] = {0, 0};  // stream traversal: resident CTAs per SM of k_traverse_stream<false / true>
    bool auto_batch = false;        // the batch in flight takes part in the automatic traversal choice
    DevBuf<float4> vtr;             // volumetric wavefront: march results
    DevBuf<uint2> vstate;           // volumetric wavefront: nested-volume stack + event word
    DevBuf<uint32_t> queue3;        // volumetric wavefront: fourth queue (hit / trace ping-pong, shade, march)
    struct PathGen {                // volumetric wavefront, dense form: one generation of the compact path arrays
        DevBuf<float4> ray_o, ray_d, thp, L, isect;
        DevBuf<uint2> rng, vstate;
        DevBuf<float> pdf;
        DevBuf<uint32_t> lobes, pix;
        void alloc(size_t n) {
            ray_o.alloc(n), ray_d.alloc(n), thp.alloc(n), L.alloc(n), isect.alloc(n), rng.alloc(n), vstate.alloc(n), pdf.alloc(n), lobes.alloc(n),
                pix.alloc(n);
        }
        PathState view() { return PathState{ray_o.p, ray_d.p, thp.p, L.p, isect.p, rng.p, pdf.p, lobes.p, vstate.p}; }
    } gen[2];
    DevBuf<int> counters;                 // [0..D] queue sizes, [D+1 ..] work counters
    DevBuf<unsigned long long> stats;     // closest, shadow, shaded
    std::vector<cudaTextureObject_t> tex_objs;
    std::vector<cudaArray_t> tex_arrays;
    int capacity_samples = 0;
    int pair_depth = 0, nesting_violations = 0, clipped_refs = 0;
    float max_overshoot = 0.f;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    size_t scene_bytes = 0;
    // optional per-launch timing
    std::vector<cudaEvent_t> kev;
    std::vector<int> kev_kind;     // kind of the launch between kev[i] and kev[i+1]
    std::vector<int> kev_bounce;
    std::vector<int> h_counts;
};

static CameraPOD camera_from_floats(const float *c) {
    CameraPOD cam;
    cam.R.r0 = V3{c[0], c[1], c[2]};
    cam.R.r1 = V3{c[3], c[4], c[5]};
    cam.R.r2 = V3{c[6], c[7], c[8]};
    cam.t = V3{c[9], c[10], c[11]};
    cam.inv_focal = c[CPTS_CAM_INV_FOCAL];
    cam.hw = c[CPTS_CAM_HW];
    cam.hh = c[CPTS_CAM_HH];
    cam.sign_x = c[CPTS_CAM_SIGN_X];
    cam.aperture = c[CPTS_CAM_APERTURE];
    cam.focal_dist = c[CPTS_CAM_FOCAL_DIST];
    return cam;
}

static cudaTextureObject_t make_texture2d(const HostTexture &t, std::vector<cudaArray_t> &arrays) {
    // reference: src/impl/textures.cu — bilinear, wrap, normalised coordinates
    cudaChannelFormatDesc desc =
        t.channels == 4 ? cudaCreateChannelDesc<float4>() : cudaCreateChannelDesc<float2>();
    cudaArray_t arr;
    CUDA_OK(cudaMallocArray(&arr, &desc, t.width, t.height));
    const size_t row = size_t(t.width) * t.channels * sizeof(float);
    CUDA_OK(cudaMemcpy2DToArray(arr, 0, 0, t.texels.data(), row, row, t.height, cudaMemcpyHostToDevice));
    arrays.push_back(arr);
    cudaResourceDesc res{};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 1;
    cudaTextureObject_t obj = 0;
    CUDA_OK(cudaCreateTextureObject(&obj, &res, &td, nullptr));
    return obj;
}

Renderer::Renderer(const HostScene &scene, int device, int seed_offset)
    : device_(device), seed_offset_(seed_offset), impl_(new Impl) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw std::runtime_error("cpt_b200: no CUDA device available (this library has no CPU fallback)");
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    num_sms_ = prop.multiProcessorCount;
    validate_scene(scene);
    w_ = scene.width();
    h_ = scene.height();
    max_depth_ = scene.config[CPTS_CFG_MAX_DEPTH];
    // renderer type of the XML (scene.cu:975-1000): 0 / 8 = megakernel PT -> its termination rules, 1 = wavefront
    rules_ = (scene.config[CPTS_CFG_RENDERER] == 0 || scene.config[CPTS_CFG_RENDERER] == 8) ? 1 : (scene.config[CPTS_CFG_RENDERER] == 6 ? 2 : 0);
    CUDA_OK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreate(&impl_->ev0));
    CUDA_OK(cudaEventCreate(&impl_->ev1));

    Impl &m = *impl_;
    m.caps[0] = scene.config[CPTS_CFG_MAX_DIFFUSE];
    m.caps[1] = scene.config[CPTS_CFG_MAX_SPECULAR];
    m.caps[2] = scene.config[CPTS_CFG_MAX_TRANSMIT];
    // ---- geometry / BVH
    {
        std::vector<float4> tmp(scene.nodes.size());
        memcpy(tmp.data(), scene.nodes.data(), tmp.size() * sizeof(float4));
        m.nodes.upload(tmp);
        std::vector<uint4> ctmp(scene.cache_nodes.size());
        memcpy(ctmp.data(), scene.cache_nodes.data(), ctmp.size() * sizeof(uint4));
        m.cache_nodes.upload(ctmp);
        std::vector<F4> tris, norms;
        std::vector<U4> uvs;
        build_triangles(scene, tris);
        build_normals_uvs(scene, norms, uvs);
        std::vector<float4> t4(tris.size()), n4(norms.size());
        memcpy(t4.data(), tris.data(), t4.size() * sizeof(float4));
        memcpy(n4.data(), norms.data(), n4.size() * sizeof(float4));
        std::vector<uint4> u4(uvs.size());
        memcpy(u4.data(), uvs.data(), u4.size() * sizeof(uint4));
        // reference leaves larger than this get a tree of their own in the pair layout (CPTB200_REFINE_LEAVES, 0 = off); their
        // primitives are re-ordered inside the leaf's range, every per-primitive array follows the permutation
        int refine_above = 2;
        if (const char *e = getenv("CPTB200_REFINE_LEAVES"))
            refine_above = atoi(e);
        PairLayout pl;
        build_pair_layout(scene, pl, refine_above);
        std::vector<int> prim_obj_words(scene.obj_idxs.begin(), scene.obj_idxs.end());
        std::vector<int> inv_perm;
        if (pl.refined_leaves > 0) {
            const std::vector<int32_t> &perm = pl.perm;
            const size_t P = perm.size();
            std::vector<float4> t2(t4.size()), n2(n4.size());
            std::vector<uint4> u2(u4.size());
            inv_perm.resize(P);
            for (size_t i = 0; i < P; i++) {
                const size_t o = size_t(perm[i]);
                for (int k = 0; k < 3; k++)
                    t2[3 * i + k] = t4[3 * o + k], n2[3 * i + k] = n4[3 * o + k];
                u2[i] = u4[o];
                prim_obj_words[i] = scene.obj_idxs[o];
                inv_perm[o] = int(i);
            }
            t4.swap(t2), n4.swap(n2), u4.swap(u2);
            m.prim_orig.upload(std::vector<int>(perm.begin(), perm.end()));
        }
        m.refined_leaves = pl.refined_leaves;
        m.norms.upload(n4);
        m.uvs.upload(u4);
        std::vector<float4> w4(pl.wnodes.size());
        memcpy(w4.data(), pl.wnodes.data(), w4.size() * sizeof(float4));
        // wide layout for the cooperative traversal (traversal=4): the pair tree collapsed to 4 (CPTB200_COOP_WIDTH=8: 8) children
        WideLayout wl;
        if (const char *e = getenv("CPTB200_COOP_WIDTH"))
            m.wide_width = atoi(e) == 8 ? 8 : 4;
        build_wide_layout(pl, m.wide_width, wl);
        std::vector<float4> ww(wl.nodes.size());
        memcpy(ww.data(), wl.nodes.data(), ww.size() * sizeof(float4));
        m.wide_nodes = wl.num_nodes, m.wide_max_stack = wl.max_stack, m.wide_max_stack_all = wl.max_stack_all, m.wide_depth = wl.max_depth, m.wide_fill = wl.avg_children;
        // traversal arena: pair nodes, wide nodes and triangles in one allocation, so that one L2 access-policy window can pin
        // them all (the 4-5 GB path-state pool streams through L2 every bounce and would otherwise keep evicting them)
        if (kTriStride == 4) { // 64-byte triangle records (two 256-bit loads): {v0,a20}{e1,a21}{e2,a22}{0}
            std::vector<float4> t64(t4.size() / 3 * 4, make_float4(0, 0, 0, 0));
            for (size_t i = 0; i < t4.size() / 3; i++)
                for (int k = 0; k < 3; k++)
                    t64[4 * i + k] = t4[3 * i + k];
            t4.swap(t64);
        }
        m.arena.alloc(w4.size() + ww.size() + t4.size());
        CUDA_OK(cudaMemcpy(m.arena.p, w4.data(), w4.size() * sizeof(float4), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(m.arena.p + w4.size(), ww.data(), ww.size() * sizeof(float4), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(m.arena.p + w4.size() + ww.size(), t4.data(), t4.size() * sizeof(float4), cudaMemcpyHostToDevice));
        m.arena_nodes = w4.size();
        m.arena_wide = ww.size();
        m.pair_depth = pl.max_depth;
        m.nesting_violations = pl.nesting_violations;
        m.clipped_refs = pl.clipped_refs;
        m.dev.wroot = pl.root;
        // ordered traversal's preconditions: stack depth, and child boxes inside their parents' boxes.  The
        // reference's SBVH pads clipped child boxes by AABB_EPS (1e-5) beyond the parent; that much is tolerated (a hit
        // would have to lie within the padding of a box edge to be affected), anything larger falls back to the
        // reference's own visiting order
        m.max_overshoot = pl.max_overshoot;
        if (pl.max_depth + 2 > kStackSize || pl.max_overshoot > 2.5e-5f) {
            ordered_ok_ = false;
            traversal_ = TRAVERSAL_DFS;
        }
        m.obj_idxs.upload(prim_obj_words);
        // emitter primitive map is padded like path_tracer.cu:76-87 (zero filled here)
        std::vector<int> ep(scene.emitter_prims.begin(), scene.emitter_prims.end());
        if (!inv_perm.empty())
            for (int &e : ep)
                e = inv_perm[size_t(e)];
        ep.resize(std::max<size_t>(ep.size(), size_t(std::max(scene.objects.front().prim_num, 1))), 0);
        m.emitter_prims.upload(ep);
        std::vector<ObjectPOD> objs(scene.objects.size());
        for (size_t i = 0; i < objs.size(); i++)
            objs[i] = ObjectPOD{scene.objects[i].bsdf_emitter, scene.objects[i].prim_offset, scene.objects[i].prim_num,
                                scene.objects[i].inv_area};
        m.objects.upload(objs);
        m.scene_bytes = (t4.size() + n4.size() + w4.size() + ww.size() + tmp.size()) * sizeof(float4) + u4.size() * sizeof(uint4);
    }
    // ---- textures
    for (const auto &t : scene.textures)
        m.tex_objs.push_back(make_texture2d(t, m.tex_arrays));
    auto tex_of = [&](int idx) -> cudaTextureObject_t {
        return (idx >= 0 && size_t(idx) < m.tex_objs.size()) ? m.tex_objs[idx] : 0;
    };
    // ---- materials / emitters
    {
        std::vector<Material> mats(scene.bsdfs.size());
        for (size_t i = 0; i < mats.size(); i++) {
            const CptsBsdf &b = scene.bsdfs[i];
            Material mm{};
            mm.type = b.type;
            mm.flags = b.flags;
            mm.k_d = make_float4(b.k_d[0], b.k_d[1], b.k_d[2], b.k_d[3]);
            mm.k_s = make_float4(b.k_s[0], b.k_s[1], b.k_s[2], b.k_s[3]);
            mm.k_g = make_float4(b.k_g[0], b.k_g[1], b.k_g[2], b.k_g[3]);
            if (b.type == CPTS_BSDF_GGX_CONDUCTOR) {
                for (int k = 0; k < 3; k++)
                    mm.eta_t[k] = b.extras[k], mm.kk[k] = b.extras[3 + k];
                mm.raw[0] = b.extras[6], mm.raw[1] = b.extras[7];
            } else if (b.type == CPTS_BSDF_PLASTIC || b.type == CPTS_BSDF_PLASTIC_FORWARD) {
                mm.raw[0] = b.extras[0];
                mm.trans_scaler = b.extras[1];
                mm.thickness = b.extras[2];
                mm.penetrable = b.extras[3] != 0.f;
            } else if (b.type == CPTS_BSDF_DISPERSION) {
                mm.k_d = make_float4(b.extras[0], b.extras[1], 0, 1);
            } else if (b.type == CPTS_BSDF_FORWARD) {
                mm.k_d = mm.k_s = mm.k_g = make_float4(0, 0, 0, 1);
            }
            for (int k = 0; k < 5; k++)
                mm.tex[k] = tex_of(b.tex[k]);
            mats[i] = mm;
        }
        m.mats.upload(mats);
        std::vector<EmitterPOD> ems(scene.emitters.size());
        for (size_t i = 0; i < ems.size(); i++) {
            const CptsEmitter &e = scene.emitters[i];
            EmitterPOD ee{};
            ee.type = e.type;
            ee.obj_ref = e.obj_ref;
            ee.is_sphere = e.is_sphere;
            ee.Le = make_float4(e.Le[0], e.Le[1], e.Le[2], e.Le[3]);
            ee.pos = make_float3(e.pos[0], e.pos[1], e.pos[2]);
            ee.cos_val = e.cos_val;
            ee.tex = tex_of(e.tex);
            ee.scale = e.scale;
            ee.azimuth = e.azimuth;
            ee.zenith = e.zenith;
            ee.rot = make_float4(0, 0, 0, 1);
            if (e.is_sphere)
                throw std::runtime_error("cpt_b200: sphere-bound emitters are not supported (TRIANGLE_ONLY build)");
            if (e.type == CPTS_EMIT_ENVMAP && ee.tex == 0)
                throw std::runtime_error("cpt_b200: environment map emitter without a texture");
            ems[i] = ee;
        }
        m.emitters.upload(ems);
        k_derive_tables<<<1, 64, 0, stream_>>>(m.mats.p, int(mats.size()), m.emitters.p, int(ems.size()));
        CUDA_OK(cudaGetLastError());
    }
    m.dev.nodes = m.nodes.p;
    m.dev.cache_nodes = m.cache_nodes.p;
    m.dev.obj_idxs = m.obj_idxs.p;
    m.dev.prim_orig = m.prim_orig.p; // nullptr: primitives are in the reference's order
    m.dev.tris = m.arena.p + m.arena_nodes + m.arena_wide;
    m.dev.norms = m.norms.p;
    m.dev.uvs = m.uvs.p;
    m.dev.objects = m.objects.p;
    m.dev.emitter_prims = m.emitter_prims.p;
    m.dev.mats = m.mats.p;
    m.dev.emitters = m.emitters.p;
    m.dev.wnodes = m.arena.p;
    m.dev.wide = m.arena.p + m.arena_nodes;
    // L2 residency for the traversal arena (env CPTB200_L2_PERSIST=0 turns it off)
    {
        const char *e = getenv("CPTB200_L2_PERSIST");
        const size_t bytes = m.arena.n * sizeof(float4);
        if ((!e || atoi(e) != 0) && bytes > 0 && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
            const size_t want = std::min<size_t>(bytes, size_t(prop.persistingL2CacheMaxSize));
            size_t have = 0;
            cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize);
            if (have < want)
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaStreamAttrValue attr{};
            attr.accessPolicyWindow.base_ptr = m.arena.p;
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, size_t(prop.accessPolicyMaxWindowSize));
            attr.accessPolicyWindow.hitRatio = float(std::min(1.0, double(want) / double(attr.accessPolicyWindow.num_bytes)));
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(stream_, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess)
                m.l2_window_bytes = attr.accessPolicyWindow.num_bytes;
            cudaGetLastError();
        }
    }
    {
        std::vector<MediumPOD> med(std::max<size_t>(scene.media.size(), 1));
        for (size_t i = 0; i < med.size(); i++) {
            MediumPOD mp{};
            mp.type = -1; // vacuum
            if (i > 0 && i < scene.media.size()) {
                const CptsMedium &c = scene.media[i];
                mp.type = c.type, mp.phase_type = c.phase_type, mp.grid = c.grid;
                mp.sigma_a = make_float4(c.sigma_a[0], c.sigma_a[1], c.sigma_a[2], c.sigma_a[3]);
                mp.sigma_s = make_float4(c.sigma_s[0], c.sigma_s[1], c.sigma_s[2], c.sigma_s[3]);
                mp.sigma_t = make_float4(c.sigma_a[0] + c.sigma_s[0], c.sigma_a[1] + c.sigma_s[1], c.sigma_a[2] + c.sigma_s[2],
                                         c.sigma_a[3] + c.sigma_s[3]); // HomogeneousMedium ctor (homogeneous.cuh:37-42)
                mp.phase = make_float4(c.phase[0], c.phase[1], c.phase[2], c.phase[3]);
            }
            med[i] = mp;
        }
        m.media.upload(med);
        m.dev.media = m.media.p;
        m.dev.num_media = int(med.size()) - 1;
        m.dev.cam_vol = scene.config[CPTS_CFG_CAM_VOL];
        m.max_volume = scene.config[CPTS_CFG_MAX_VOLUME];
    }
    m.dev.grids = nullptr;
    m.dev.blackbody = 0;
    if (!scene.grids.empty()) { // grid media: brick tables + pools as two flat device arrays, channels point into them
        m.grid_bricks.upload(scene.grid_bricks);
        m.grid_data.upload(scene.grid_data);
        std::vector<GridPOD> gp(scene.grids.size());
        for (size_t i = 0; i < gp.size(); i++) {
            const CptsGrid &c = scene.grids[i];
            GridPOD g{};
            g.scale = c.scale, g.temp_scale = c.temp_scale, g.em_scale = c.em_scale, g.majorant = c.majorant;
            for (int k = 0; k < 3; k++)
                g.aabb_min[k] = c.aabb_min[k], g.aabb_max[k] = c.aabb_max[k];
            for (int ch = 0; ch < 3; ch++) {
                const CptsGridChannel &cc = c.ch[ch];
                GridChannelPOD &d = g.ch[ch];
                memcpy(d.inv_mat, cc.inv_mat, sizeof(d.inv_mat));
                memcpy(d.vec, cc.vec, sizeof(d.vec));
                memcpy(d.bmin, cc.bmin, sizeof(d.bmin));
                memcpy(d.nb, cc.nb, sizeof(d.nb));
                memcpy(d.bg, cc.background, sizeof(d.bg));
                d.ncomp = cc.ncomp;
                d.bricks = m.grid_bricks.p + cc.brick_offset;
                d.data = m.grid_data.p + cc.data_offset;
            }
            gp[i] = g;
        }
        m.grids.upload(gp);
        m.dev.grids = m.grids.p;
        m.scene_bytes += scene.grid_bricks.size() * 4 + scene.grid_data.size() * 4;
    }
    m.dev.num_nodes = scene.num_nodes();
    m.dev.num_cache = int(scene.cache_nodes.size());
    m.dev.num_prims = scene.num_prims();
    m.dev.num_emitters = int(scene.emitters.size()) - 1;
    m.dev.envmap_id = scene.config[CPTS_CFG_ENVMAP_ID];
    m.dev.cie = m.dev.d65 = 0;
    // 1-D array textures, linear filter, clamp, normalised coordinates (reference: src/core/xyz.cuh:60-84)
    auto make1d = [&](const void *src, size_t count, bool four) {
        cudaChannelFormatDesc desc = four ? cudaCreateChannelDesc<float4>() : cudaCreateChannelDesc<float>();
        cudaArray_t arr;
        CUDA_OK(cudaMallocArray(&arr, &desc, count));
        const size_t bytes = count * (four ? 16 : 4);
        CUDA_OK(cudaMemcpy2DToArray(arr, 0, 0, src, bytes, bytes, 1, cudaMemcpyHostToDevice));
        m.tex_arrays.push_back(arr);
        cudaResourceDesc res{};
        res.resType = cudaResourceTypeArray;
        res.res.array.array = arr;
        cudaTextureDesc td{};
        td.addressMode[0] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 1;
        cudaTextureObject_t obj = 0;
        CUDA_OK(cudaCreateTextureObject(&obj, &res, &td, nullptr));
        m.tex_objs.push_back(obj);
        return obj;
    };
    if (!scene.cie_xyz.empty()) {
        m.dev.cie = make1d(scene.cie_xyz.data(), scene.cie_xyz.size(), true);
        m.dev.d65 = make1d(scene.d65_spd.data(), scene.d65_spd.size(), false);
    }
    if (!scene.grids.empty()) // vol_grid.cu:304-326
        m.dev.blackbody = make1d(scene.blackbody.data(), scene.blackbody.size(), true);
    m.cam = camera_from_floats(scene.camera);

    const size_t P = size_t(w_) * h_;
    CUDA_OK(cudaMalloc(&d_image_, P * sizeof(float4)));
    CUDA_OK(cudaMalloc(&d_out_, P * sizeof(float4)));
    CUDA_OK(cudaMemsetAsync(d_image_, 0, P * sizeof(float4), stream_));
    CUDA_OK(cudaMemsetAsync(d_out_, 0, P * sizeof(float4), stream_));
    m.stats.alloc(4);
    CUDA_OK(cudaMemsetAsync(m.stats.p, 0, 4 * sizeof(unsigned long long), stream_));
    if (const char *e = getenv("CPTB200_BLOCKS_PER_SM"))
        blocks_per_sm_ = std::max(1, atoi(e));
    if (const char *e = getenv("CPTB200_COOP_FLUSH"))
        coop_flush_ = std::min(32, std::max(1, atoi(e)));
    if (const char *e = getenv("CPTB200_COOP_SLAB"))
        coop_slab_ = std::max(0, atoi(e));
    if (const char *e = getenv("CPTB200_COOP_RAYGEN"))
        coop_raygen_ = atoi(e) != 0;
    set_coop(coop_flush_, coop_slab_);
    if (const char *e = getenv("CPTB200_TRAVERSAL"))
        set_traversal(atoi(e));
    if (const char *e = getenv("CPTB200_DYN_REFILL"))
        dyn_refill_ = std::min(32, std::max(1, atoi(e)));
    if (const char *e = getenv("CPTB200_DYN_STEPS"))
        dyn_steps_ = std::max(1, atoi(e));
    if (const char *e = getenv("CPTB200_DYN_MINB"))
        dyn_minb_ = atoi(e);
    if (const char *e = getenv("CPTB200_SQ_REFILL"))
        set_stream(atoi(e), -1, -1, -1);
    if (const char *e = getenv("CPTB200_SQ_FLUSH"))
        set_stream(-1, atoi(e), -1, -1);
    if (const char *e = getenv("CPTB200_SQ_STEPS"))
        set_stream(-1, -1, atoi(e), -1);
    if (const char *e = getenv("CPTB200_SQ_SHADOW"))
        sq_shadow_ = atoi(e) != 0;
    if (const char *e = getenv("CPTB200_SQ_BLOCKS"))
        set_stream(-1, -1, -1, atoi(e));
    if (const char *e = getenv("CPTB200_DQ_FLUSH"))
        dq_flush_ = std::min(32, std::max(1, atoi(e)));
    // default number of samples in flight: ~32M paths (5 GB of path state at 160 B per path; the traversal kernels - the stream
    // traversal above all - want long queues, and the pool is sized for 180 GB of HBM), at least 1, at most 64
    set_samples_in_flight(int(std::min<size_t>(64, std::max<size_t>(1, (size_t(32) << 20) / P))));
    CUDA_OK(cudaStreamSynchronize(stream_));
}

Renderer::~Renderer() {
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    for (auto t : impl_->tex_objs)
        cudaDestroyTextureObject(t);
    for (auto a : impl_->tex_arrays)
        cudaFreeArray(a);
    if (impl_->ev0)
        cudaEventDestroy(impl_->ev0);
    if (impl_->ev1)
        cudaEventDestroy(impl_->ev1);
    for (auto e : impl_->kev)
        cudaEventDestroy(e);
    cudaFree(d_image_);
    cudaFree(d_out_);
    cudaFree(d_var_);
    impl_.reset();
    cudaStreamDestroy(stream_);
}

// shared memory and occupancy of the cooperative kernels for the current knobs; decides whether traversal=4 may be used
void Renderer::set_coop(int flush_lanes, int slab_nodes) {
    Impl &m = *impl_;
    CUDA_OK(cudaSetDevice(device_));
    coop_flush_ = std::min(32, std::max(1, flush_lanes));
    coop_slab_ = std::min(std::max(0, slab_nodes), m.wide_nodes);
    const int G = m.wide_width;
    coop_ok_ = false;
    m.coop_blocks_per_sm = 0;
    int max_optin = 0;
    CUDA_OK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device_));
    if (G == 8)
        coop_slab_ = 0; // the slab variant is instantiated for groups of 4 only
    for (;; coop_slab_ /= 2) { // shrink the slab until the carve-up fits
        const int cap = std::max(8, (m.wide_max_stack + 7) / 8 * 8);
        const CoopLayout L = coop_layout(m.dev.num_cache, coop_slab_, G, cap, kCoopBlock / 32);
        if (L.total <= max_optin) {
            m.coop_smem = size_t(L.total);
            break;
        }
        if (coop_slab_ == 0)
            return; // the stacks alone do not fit: the tree is too deep for the cooperative kernels
    }
    auto prep = [&](auto kernel) {
        CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(m.coop_smem)));
        int nb = 0;
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kCoopBlock, m.coop_smem));
        return nb;
    };
    if (G == 8)
        m.coop_blocks[0] = prep(k_trace_coop<8, false>), m.coop_blocks[1] = prep(k_shadow_apply_coop<8, false>), m.coop_blocks[2] = prep(k_raygen_coop<8, false>);
    else if (coop_slab_ > 0)
        m.coop_blocks[0] = prep(k_trace_coop<4, true>), m.coop_blocks[1] = prep(k_shadow_apply_coop<4, true>), m.coop_blocks[2] = prep(k_raygen_coop<4, true>);
    else
        m.coop_blocks[0] = prep(k_trace_coop<4, false>), m.coop_blocks[1] = prep(k_shadow_apply_coop<4, false>), m.coop_blocks[2] = prep(k_raygen_coop<4, false>);
    const int nb = std::min(std::min(m.coop_blocks[0], m.coop_blocks[1]), m.coop_blocks[2]);
    m.coop_blocks_per_sm = nb;
    // preconditions of the ordered traversal, plus: the stack carve-up fits, and no leaf holds a clipped (SBVH) reference - the
    // reference's result on those depends on the order in which single leaves are tested (PairLayout::clipped_refs)
    coop_ok_ = nb >= 1 && ordered_ok_ && m.clipped_refs == 0 && m.nesting_violations == 0;
    // per-thread traversal of the same layout (traversal=5): 4-ary records, per-thread stack of kStackSize entries
    wide4_ok_ = G == 4 && ordered_ok_ && m.clipped_refs == 0 && m.nesting_violations == 0 && m.wide_max_stack_all + 1 <= kWideStack;
    // stream traversal (traversal=7): the pair layout per thread with deferred leaves - the same restriction on clipped references
    stream_ok_ = ordered_ok_ && m.clipped_refs == 0 && m.nesting_violations == 0 && m.dev.num_nodes < (1 << kStartBits);
}

void Renderer::set_stream(int refill, int flush, int steps, int blocks) {
    if (refill >= 0)
        sq_refill_ = std::min(32, std::max(1, refill));
    if (flush >= 0)
        sq_flush_ = std::min(32, std::max(1, flush));
    if (steps >= 0)
        sq_steps_ = std::min(64, std::max(1, steps));
    if (blocks >= 0)
        sq_blocks_ = std::min(32, blocks);
}

void Renderer::auto_observe(float ms, int samples) {
    // called after every synchronised batch: the first full-size batch warms up, the second is timed with the ordered traversal,
    // the third with the stream traversal for the closest-hit rays, the fourth with shadow rays in the stream kernel too; from
    // the fifth on the fastest of the three runs (see launch_batch)
    if (!impl_->auto_batch || auto_state_ >= kAutoDecided || samples <= 0)
        return;
    if (auto_state_ == 0 || samples > auto_samples_) { // (a larger batch size than seen so far: start over with it)
        auto_samples_ = samples, auto_state_ = 1;
        return;
    }
    if (samples < auto_samples_) // a remainder batch (spp not a multiple of the samples in flight): not comparable, this trial is
        return;                  // repeated with the next full-size batch
    auto_ms_per_sample_[auto_state_ - 1] = ms / float(samples);
    if (++auto_state_ == kAutoDecided) { // a more elaborate mode has to win by 1.5 %
        const float t2 = auto_ms_per_sample_[0], t7 = auto_ms_per_sample_[1], t7s = auto_ms_per_sample_[2];
        auto_choice_ = std::min(t7, t7s) < 0.985f * t2 ? 7 : 2;
        auto_shadow_ = auto_choice_ == 7 && t7s < 0.985f * t7;
    }
}

void Renderer::set_traversal(int t) {
    if (t == -1) {
        auto_traversal_ = true, auto_state_ = 0, auto_choice_ = 2, auto_shadow_ = false;
        traversal_ = ordered_ok_ ? 2 : TRAVERSAL_DFS;
        return;
    }
    if (t < 0 || t > 7)
        throw std::runtime_error("traversal must be -1 (automatic) or 0..7");
    auto_traversal_ = false;
    if (t == 6 && !ordered_ok_)
        t = TRAVERSAL_DFS;
    if (t == 7 && !stream_ok_)
        t = 2;
    if (t == 4 && !coop_ok_)
        t = 2;
    if (t == 5 && !wide4_ok_)
        t = 2;
    if (t != TRAVERSAL_DFS && !ordered_ok_)
        t = TRAVERSAL_DFS;
    traversal_ = t;
}

void Renderer::set_samples_in_flight(int s) {
    if (s < 1)
        throw std::runtime_error("samples_in_flight must be >= 1");
    const size_t P = size_t(w_) * h_;
    if (P * size_t(s) >= (size_t(1) << 32))
        throw std::runtime_error("samples_in_flight * pixels exceeds the 32-bit slot index");
    s_max_ = s;
    auto_state_ = 0;
}

void Renderer::enable_variance(bool on) {
    CUDA_OK(cudaSetDevice(device_));
    if (on && !d_var_) {
        CUDA_OK(cudaMalloc(&d_var_, size_t(w_) * h_ * sizeof(float)));
        CUDA_OK(cudaMemsetAsync(d_var_, 0, size_t(w_) * h_ * sizeof(float), stream_));
    }
    variance_on_ = on;
}

void Renderer::ensure_capacity(int samples) {
    Impl &m = *impl_;
    const bool vpt_wave = rules_ == 2 && pipeline_ != 0;
    if (samples <= m.capacity_samples && int(m.counters.n) >= 8 * (max_depth_ + 2) + 8 && (pipeline_ < 2 || m.hit.n > 0) &&
        (!vpt_wave || (m.vstate.n > 0 && m.hit.n > 0)) && (!((vpt_wave && pipeline_ >= 2) || (rules_ != 2 && pipeline_ == 5)) || m.gen[1].pix.n > 0))
        return;
    CUDA_OK(cudaStreamSynchronize(stream_));
    const size_t slots = size_t(w_) * h_ * size_t(std::max(samples, m.capacity_samples));
    m.ray_o.alloc(slots);
    m.ray_d.alloc(slots);
    m.thp.alloc(slots);
    m.L.alloc(slots);
    m.isect.alloc(slots);
    m.rng.alloc(slots);
    m.pdf.alloc(slots);
    m.lobes.alloc(slots);
    m.queue[0].alloc(slots);
    m.queue[1].alloc(slots);
    if (vpt_wave) {
        m.vtr.alloc(slots);
        m.vstate.alloc(slots);
        m.queue3.alloc(slots);
    }
    if ((vpt_wave && pipeline_ >= 2) || (rules_ != 2 && pipeline_ == 5)) {
        m.gen[0].alloc(slots);
        m.gen[1].alloc(slots);
    }
    if (pipeline_ >= 2 || vpt_wave) {
        m.shadow_queue.alloc(slots);
        m.sh_o.alloc(slots);
        m.sh_d.alloc(slots);
        m.hit.alloc(slots);
        m.vis.alloc(slots);
    }
    m.counters.alloc(size_t(8 * (max_depth_ + 2) + 8)); // D+2 queue sizes, D+2 shadow-queue sizes, then one work counter per launch (<= 5D+1)
    m.capacity_samples = std::max(samples, m.capacity_samples);
}

void Renderer::launch_batch(int first_accum, int samples, bool gamma, bool write_output, bool use_variance) {
    Impl &m = *impl_;
    CUDA_OK(cudaSetDevice(device_));
    if (row_begin() >= row_end())
        throw std::runtime_error("cpt_b200: empty row range (row_begin >= row_end)");
    // automatic traversal choice (set_traversal(-1), the default): this batch runs the mode the tuning state asks for
    struct Restore {
        int &ref, saved;
        ~Restore() { ref = saved; }
    } restore_traversal{traversal_, traversal_};
    struct RestoreFlag {
        bool &ref, saved;
        ~RestoreFlag() { ref = saved; }
    } restore_shadow{sq_shadow_, sq_shadow_};
    m.auto_batch = auto_traversal_ && stream_ok_ && traversal_ == 2 && rules_ != 2 && pipeline_ == 3;
    if (m.auto_batch) {
        traversal_ = (auto_state_ == 2 || auto_state_ == 3) ? 7 : (auto_state_ >= kAutoDecided ? auto_choice_ : 2);
        sq_shadow_ = auto_state_ == 3 || (auto_state_ >= kAutoDecided && auto_shadow_);
    }
    ensure_capacity(samples);
    const int P = w_ * h_, D = max_depth_;
    PathState st{m.ray_o.p, m.ray_d.p, m.thp.p, m.L.p, m.isect.p, m.rng.p, m.pdf.p, m.lobes.p, m.vstate.p};
    BatchParams bp{w_, h_, P, samples, first_accum, seed_offset_, row0_, row1_ > 0 ? std::min(row1_, h_) : h_, rules_,
                   max_depth_, m.caps[0], m.caps[1], m.caps[2], m.max_volume > 0 ? m.max_volume : max_depth_};
    int *q_counts = m.counters.p;          // D + 1 entries
    int *sq_counts = m.counters.p + (D + 2); // D + 1 entries (dynamic-fetch pipeline)
    int *work = m.counters.p + 2 * (D + 2);  // one per launch
    RayBufs rb{m.sh_o.p, m.sh_d.p, m.hit.p};
    CUDA_OK(cudaMemsetAsync(m.counters.p, 0, m.counters.n * sizeof(int), stream_));
    const size_t smem = size_t(m.dev.num_cache) * sizeof(uint4);
    const size_t smem_dyn = smem + size_t(kSmemTop) * kBlock * sizeof(uint2); // dynamic-fetch kernels: + shared-memory stack
    const size_t smem_any = size_t(kSmemTop) * kBlock * sizeof(uint2);
    const int grid = num_sms_ * blocks_per_sm_;
    // traversal=4: cooperative forms of k_raygen / k_shadow_apply / k_trace in the deferred-NEE surface pipeline; every other
    // kernel (fused shading, volumetric path) keeps the ordered per-thread traversal
    const bool coop = traversal_ == 4 && coop_ok_ && rules_ != 2 && this->pipeline_ == 3;
    const int mode = (traversal_ == 4 || traversal_ == 5 || traversal_ == 6 || traversal_ == 7) ? 2 : traversal_;
    const bool dq = traversal_ == 6 && ordered_ok_ && rules_ != 2 && this->pipeline_ == 3;
    const bool wide4 = traversal_ == 5 && wide4_ok_ && rules_ != 2 && this->pipeline_ == 3;
    const bool sq = traversal_ == 7 && stream_ok_ && rules_ != 2 && this->pipeline_ == 3;
    const size_t smem_sq = smem + kCtxBytes + size_t(kBlock / 32) * kRingBytes + size_t(kSqSmemStack) * kBlock * sizeof(uint2);
    if (sq && m.stream_blocks[0] == 0) { // resident CTAs per SM of the two stream kernels with this scene's shared-memory size
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m.stream_blocks[0], k_traverse_stream<false>, kBlock, smem_sq));
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m.stream_blocks[1], k_traverse_stream<true>, kBlock, smem_sq));
        if (m.stream_blocks[0] < 1 || m.stream_blocks[1] < 1)
            throw std::runtime_error("cpt_b200: the stream traversal kernels do not fit on this device");
    }
    const int grid_sq[2] = {num_sms_ * (sq_blocks_ > 0 ? std::min(sq_blocks_, std::max(1, m.stream_blocks[0])) : std::max(1, m.stream_blocks[0])),
                            num_sms_ * (sq_blocks_ > 0 ? std::min(sq_blocks_, std::max(1, m.stream_blocks[1])) : std::max(1, m.stream_blocks[1]))};
    int coop_grid = num_sms_; // set per kernel by LAUNCH_COOP
    const CoopParams cpar{m.arena.p + m.arena_nodes, m.wide_width == 8 ? 0 : coop_slab_, std::max(8, (m.wide_max_stack + 7) / 8 * 8)};
#define LAUNCH_COOP(which, kernel, ...)                                                                                \
    do {                                                                                                               \
        coop_grid = num_sms_ * std::max(1, m.coop_blocks[which]);                                                      \
        if (m.wide_width == 8)                                                                                         \
            kernel<8, false><<<coop_grid, kCoopBlock, m.coop_smem, stream_>>>(__VA_ARGS__);                             \
        else if (coop_slab_ > 0)                                                                                       \
            kernel<4, true><<<coop_grid, kCoopBlock, m.coop_smem, stream_>>>(__VA_ARGS__);                              \
        else                                                                                                           \
            kernel<4, false><<<coop_grid, kCoopBlock, m.coop_smem, stream_>>>(__VA_ARGS__);                             \
    } while (0)
    const int pipeline_ = (traversal_ == 1 || traversal_ == 3 || traversal_ == 4 || traversal_ == 5 || traversal_ == 6 || traversal_ == 7) ? (this->pipeline_ == 2 ? 1 : (this->pipeline_ == 4 ? 3 : this->pipeline_)) : this->pipeline_; // dynamic fetch exists for the ordered traversal only
#define LAUNCH_MODE(kernel, ...)                                                                                       \
    do {                                                                                                               \
        if (mode == 0)                                                                                                 \
            kernel<0><<<grid, kBlock, smem, stream_>>>(__VA_ARGS__);                                                   \
        else if (mode == 1)                                                                                            \
            kernel<1><<<grid, kBlock, smem, stream_>>>(__VA_ARGS__);                                                   \
        else if (mode == 2)                                                                                            \
            kernel<2><<<grid, kBlock, smem, stream_>>>(__VA_ARGS__);                                                   \
        else                                                                                                           \
            kernel<3><<<grid, kBlock, smem, stream_>>>(__VA_ARGS__);                                                   \
    } while (0)
    CUDA_OK(cudaEventRecord(m.ev0, stream_));
    int wk = 0;
    size_t kev_used = 0;
    m.kev_kind.clear();
    m.kev_bounce.clear();
    auto mark = [&](int kind, int bounce) { // event BEFORE a launch of `kind` (-1: closing event)
        if (!kernel_timing_)
            return;
        if (kev_used == m.kev.size()) {
            cudaEvent_t e;
            CUDA_OK(cudaEventCreate(&e));
            m.kev.push_back(e);
        }
        CUDA_OK(cudaEventRecord(m.kev[kev_used++], stream_));
        if (kind >= 0) {
            m.kev_kind.push_back(kind);
            m.kev_bounce.push_back(bounce);
        }
    };
    if (rules_ == 2 && this->pipeline_ >= 2) { // volumetric path, wavefront with dense path records (see k_vptd_post)
        int *cnt_n = m.counters.p, *cnt_m = cnt_n + (D + 2), *cnt_h = cnt_n + 2 * (D + 2), *cnt_b = cnt_n + 3 * (D + 2), *wkv = cnt_n + 4 * (D + 2);
        const int cap = int(m.gen[0].pix.n);
        uint32_t *q_hit = m.queue[0].p;
        VptBufs vb{m.sh_o.p, m.sh_d.p, m.hit.p, m.vtr.p};
        PathState g[2] = {m.gen[0].view(), m.gen[1].view()};
        uint32_t *gp[2] = {m.gen[0].pix.p, m.gen[1].pix.p};
        uint32_t *q_march = m.queue3.p;
        mark(0, 0);
        LAUNCH_MODE(k_vptd_raygen, m.dev, m.cam, g[0], gp[0], m.L.p, bp, cnt_n, wkv + wk);
        ++wk;
        int launches = 1, cur = 0;
        for (int b = 0; b < D; b++) {
            mark(1, b);
            const uint32_t *q_in = b == 0 ? nullptr : q_hit; // bounce 0: the generation k_vptd_raygen wrote is all hits
            const int *n_in = b == 0 ? cnt_n : cnt_h + b;
            k_vptd_pre<<<grid, kBlock, 0, stream_>>>(m.dev, g[cur], vb, q_in, n_in, wkv + wk, q_march, cnt_m + b, m.stats.p);
            ++wk;
            mark(4, b);
            LAUNCH_MODE(k_vpt_march, m.dev, g[cur], vb, q_march, cnt_m + b, wkv + wk, m.stats.p);
            ++wk;
            mark(1, b);
            k_vptd_post<<<grid, kBlock, 0, stream_>>>(m.dev, g[cur], g[cur ^ 1], gp[cur], gp[cur ^ 1], vb, m.L.p, bp, q_in, n_in, cnt_n + b + 1,
                                                      cnt_b + b + 1, cap, wkv + wk, b);
            ++wk;
            launches += 3;
            cur ^= 1;
            if (b + 1 >= D)
                break;
            mark(2, b);
            LAUNCH_MODE(k_vptd_trace, m.dev, g[cur], gp[cur], m.L.p, cnt_n + b + 1, cnt_b + b + 1, cap, wkv + wk, q_hit, cnt_h + b + 1);
            ++wk, ++launches;
        }
        mark(3, 0);
        k_splat<<<num_sms_ * 4, 256, 0, stream_>>>(m.L.p, d_image_, d_out_, use_variance ? d_var_ : nullptr, P, samples, first_accum,
                                                   write_output ? 1 : 0, gamma ? 1 : 0, 0, row0_ * w_, row_end() * w_);
        mark(-1, 0);
        k_fold_vptd_stats<<<1, 32, 0, stream_>>>(cnt_n, cnt_b, cnt_m, D, m.stats.p);
        CUDA_OK(cudaEventRecord(m.ev1, stream_));
        CUDA_OK(cudaGetLastError());
        host_stats_.samples += uint64_t(w_) * uint64_t(row_end() - row_begin()) * samples;
        host_stats_.launches += launches + 2;
        return;
    }
    if (rules_ == 2 && this->pipeline_ != 0) { // volumetric path, wavefront over slot queues (pipeline 1; see k_vpt_pre)
        int *cnt_h = m.counters.p, *cnt_m = cnt_h + (D + 2), *cnt_s = cnt_h + 2 * (D + 2), *cnt_t = cnt_h + 3 * (D + 2);
        int *wkv = cnt_h + 4 * (D + 2);
        VptBufs vb{m.sh_o.p, m.sh_d.p, m.hit.p, m.vtr.p};
        uint32_t *q_hit = m.queue[0].p, *q_trace = m.queue[1].p, *q_shade = m.shadow_queue.p, *q_march = m.queue3.p;
        mark(0, 0);
        LAUNCH_MODE(k_raygen, m.dev, m.cam, st, bp, q_hit, cnt_h, wkv + wk);
        ++wk;
        int launches = 1;
        for (int b = 0; b < D; b++) {
            mark(1, b);
            k_vpt_pre<<<grid, kBlock, 0, stream_>>>(m.dev, st, vb, q_hit, cnt_h + b, wkv + wk, q_shade, cnt_s + b, q_march, cnt_m + b, q_trace,
                                                    cnt_t + b + 1);
            ++wk;
            mark(4, b);
            LAUNCH_MODE(k_vpt_march, m.dev, st, vb, q_march, cnt_m + b, wkv + wk, m.stats.p);
            ++wk;
            mark(1, b);
            k_vpt_post<<<grid, kBlock, 0, stream_>>>(m.dev, st, vb, bp, q_shade, cnt_s + b, wkv + wk, q_trace, cnt_t + b + 1, b);
            ++wk;
            launches += 3;
            if (b + 1 >= D)
                break;
            mark(2, b);
            LAUNCH_MODE(k_trace, m.dev, st, q_trace, cnt_t + b + 1, wkv + wk, q_hit, cnt_h + b + 1, b + 1, 1);
            ++wk, ++launches;
        }
        mark(3, 0);
        k_splat<<<num_sms_ * 4, 256, 0, stream_>>>(m.L.p, d_image_, d_out_, use_variance ? d_var_ : nullptr, P, samples, first_accum,
                                                   write_output ? 1 : 0, gamma ? 1 : 0, 0, row0_ * w_, row_end() * w_);
        mark(-1, 0);
        k_fold_vpt_stats<<<1, 32, 0, stream_>>>(cnt_t, cnt_m, cnt_s, D, m.stats.p);
        CUDA_OK(cudaEventRecord(m.ev1, stream_));
        CUDA_OK(cudaGetLastError());
        host_stats_.samples += uint64_t(w_) * uint64_t(row_end() - row_begin()) * samples;
        host_stats_.launches += launches + 2;
        return;
    }
    if (rules_ == 2) { // volumetric path, megakernel form (pipeline 0): one kernel per batch + the splat
        mark(0, 0);
        LAUNCH_MODE(k_vpt, m.dev, m.cam, st, bp, work, m.stats.p);
        mark(3, 0);
        k_splat<<<num_sms_ * 4, 256, 0, stream_>>>(m.L.p, d_image_, d_out_, use_variance ? d_var_ : nullptr, P, samples, first_accum,
                                                   write_output ? 1 : 0, gamma ? 1 : 0, 0, row0_ * w_, row_end() * w_);
        mark(-1, 0);
        CUDA_OK(cudaEventRecord(m.ev1, stream_));
        CUDA_OK(cudaGetLastError());
        host_stats_.samples += uint64_t(w_) * uint64_t(row_end() - row_begin()) * samples;
        host_stats_.launches += 2;
        return;
    }
    if (this->pipeline_ == 5) { // deferred NEE over dense path records (see DenseIO / k_vptd_post): surface path
        int *ent = m.counters.p, *sqc = ent + (D + 2), *wkd = ent + 2 * (D + 2);
        PathState g[2] = {m.gen[0].view(), m.gen[1].view()};
        uint32_t *gp[2] = {m.gen[0].pix.p, m.gen[1].pix.p};
        unsigned long long *shadow_counter = m.stats.p + 1;
        mark(0, 0);
        LAUNCH_MODE(k_vptd_raygen, m.dev, m.cam, g[0], gp[0], m.L.p, bp, ent, wkd + wk);
        ++wk;
        int launches = 1, traced = 0, shaded = 0, cur = 0;
        for (int b = 0; b < D; b++) {
            const PtRules pt{rules_, b, m.caps[0], m.caps[1], m.caps[2]};
            mark(1, b);
            k_shade<-2, true><<<grid, kBlock, 0, stream_>>>(m.dev, g[cur], nullptr, ent + b, wkd + wk, shadow_counter, b > 0, pt, rb, m.shadow_queue.p,
                                                             sqc + b, DenseIO{g[cur ^ 1], gp[cur], gp[cur ^ 1], ent + b + 1});
            ++wk, ++launches, ++shaded;
            cur ^= 1;
            mark(4, b);
            LAUNCH_MODE(k_shadow_apply, m.dev, g[cur], rb, m.shadow_queue.p, sqc + b, wkd + wk);
            ++wk, ++launches;
            if (b + 1 >= D)
                break;
            mark(2, b);
            LAUNCH_MODE(k_trace_dense, m.dev, g[cur], gp[cur], m.L.p, ent + b + 1, wkd + wk, b, rules_);
            ++wk, ++launches, ++traced;
        }
        mark(3, 0);
        k_deposit<<<num_sms_ * 8, 256, 0, stream_>>>(g[cur], gp[cur], m.L.p, ent + D);
        k_splat<<<num_sms_ * 4, 256, 0, stream_>>>(m.L.p, d_image_, d_out_, use_variance ? d_var_ : nullptr, P, samples, first_accum,
                                                   write_output ? 1 : 0, gamma ? 1 : 0, rules_ == 0 ? 1 : 0, row0_ * w_, row_end() * w_);
        mark(-1, 0);
        k_fold_stats<<<1, 32, 0, stream_>>>(ent + 1, traced, shaded, m.stats.p, sqc, shaded);
        CUDA_OK(cudaEventRecord(m.ev1, stream_));
        CUDA_OK(cudaGetLastError());
        host_stats_.samples += uint64_t(w_) * uint64_t(row_end() - row_begin()) * samples;
        host_stats_.launches += launches + 3;
        return;
    }
    mark(0, 0);
    if (coop && coop_raygen_) // primary rays are coherent: the per-thread kernel runs them at 15 of 32 lanes and is the faster one
        LAUNCH_COOP(2, k_raygen_coop, m.dev, cpar, m.cam, st, bp, m.queue[0].p, q_counts, work + wk);
    else if (wide4 && coop_raygen_)
        k_raygen<5><<<grid, kBlock, smem, stream_>>>(m.dev, m.cam, st, bp, m.queue[0].p, q_counts, work + wk);
    else
        LAUNCH_MODE(k_raygen, m.dev, m.cam, st, bp, m.queue[0].p, q_counts, work + wk);
    ++wk;
    unsigned long long *shadow_counter = m.stats.p + 1;
    int launches = 1, traced = 0, shaded = 0;
    for (int b = 0; b < D; b++) {
        uint32_t *q_cur = m.queue[b & 1].p, *q_nxt = m.queue[(b + 1) & 1].p;
        const PtRules pt{rules_, b, m.caps[0], m.caps[1], m.caps[2]};
        if (pipeline_ == 3 || pipeline_ == 4) {
            mark(1, b);
            k_shade<-2><<<grid, kBlock, 0, stream_>>>(m.dev, st, q_cur, q_counts + b, work + wk, shadow_counter, b > 0, pt, rb,
                                                      m.shadow_queue.p, sq_counts + b);
            ++wk, ++launches;
            mark(4, b);
            if (pipeline_ == 3 && coop)
                LAUNCH_COOP(1, k_shadow_apply_coop, m.dev, cpar, st, rb, m.shadow_queue.p, sq_counts + b, work + wk);
            else if (pipeline_ == 3 && wide4)
                k_shadow_apply<5><<<grid, kBlock, smem, stream_>>>(m.dev, st, rb, m.shadow_queue.p, sq_counts + b, work + wk);
            else if (pipeline_ == 3 && sq && sq_shadow_) {
                k_traverse_stream<true><<<grid_sq[1], kBlock, smem_sq, stream_>>>(m.dev, st, rb, m.shadow_queue.p, sq_counts + b, work + wk, m.vis.p, nullptr, nullptr, b,
                                                                                    rules_, sq_refill_, sq_flush_, sq_steps_);
                k_shadow_finish<<<num_sms_ * 8, 256, smem, stream_>>>(m.dev, st, rb, m.shadow_queue.p, sq_counts + b, m.vis.p);
                ++launches;
            } else if (pipeline_ == 3 && dq)
                k_traverse_dq<true><<<grid, kBlock, smem, stream_>>>(m.dev, st, rb, m.shadow_queue.p, sq_counts + b, work + wk, nullptr, nullptr, b, rules_,
                                                                     dyn_refill_, dq_flush_, dyn_steps_);
            else if (pipeline_ == 3)
                LAUNCH_MODE(k_shadow_apply, m.dev, st, rb, m.shadow_queue.p, sq_counts + b, work + wk);
            else if (dyn_minb_ >= 12)
                k_anyhit<12, true><<<grid, kBlock, smem_dyn, stream_>>>(m.dev, st, rb, m.shadow_queue.p, sq_counts + b, work + wk, dyn_refill_, dyn_steps_);
            else
                k_anyhit<8, true><<<grid, kBlock, smem_dyn, stream_>>>(m.dev, st, rb, m.shadow_queue.p, sq_counts + b, work + wk, dyn_refill_, dyn_steps_);
        } else if (pipeline_ == 2) {
            mark(4, b);
            k_shadow_gen<<<grid, kBlock, smem, stream_>>>(m.dev, st, rb, q_cur, q_counts + b, work + wk, m.shadow_queue.p,
                                                          sq_counts + b, shadow_counter);
            ++wk, ++launches;
            mark(4, b);
            if (dyn_minb_ >= 12)
                k_anyhit<12, false><<<grid, kBlock, smem_any, stream_>>>(m.dev, st, rb, m.shadow_queue.p, sq_counts + b, work + wk, dyn_refill_, dyn_steps_);
            else
                k_anyhit<8, false><<<grid, kBlock, smem_any, stream_>>>(m.dev, st, rb, m.shadow_queue.p, sq_counts + b, work + wk, dyn_refill_, dyn_steps_);
            ++wk, ++launches;
            mark(1, b);
            k_shade<-1><<<grid, kBlock, 0, stream_>>>(m.dev, st, q_cur, q_counts + b, work + wk, shadow_counter, b > 0, pt, rb, nullptr, nullptr);
        } else if (pipeline_ == 1) {
            mark(4, b);
            LAUNCH_MODE(k_shadow, m.dev, st, q_cur, q_counts + b, work + wk, shadow_counter);
            ++wk, ++launches;
            mark(1, b);
            k_shade<-1><<<grid, kBlock, 0, stream_>>>(m.dev, st, q_cur, q_counts + b, work + wk, shadow_counter, b > 0, pt, rb, nullptr, nullptr);
        } else {
            mark(1, b);
            LAUNCH_MODE(k_shade, m.dev, st, q_cur, q_counts + b, work + wk, shadow_counter, b > 0, pt, rb, nullptr, nullptr);
        }
        ++wk;
        ++launches, ++shaded;
        if (b + 1 >= D)
            break;
        mark(2, b);
        if (pipeline_ == 2 || pipeline_ == 4) {
            if (dyn_minb_ >= 12)
                k_intersect<12><<<grid, kBlock, smem_dyn, stream_>>>(m.dev, st, rb, q_cur, q_counts + b, work + wk, dyn_refill_, dyn_steps_);
            else
                k_intersect<8><<<grid, kBlock, smem_dyn, stream_>>>(m.dev, st, rb, q_cur, q_counts + b, work + wk, dyn_refill_, dyn_steps_);
            ++wk, ++launches;
            mark(2, b);
            k_resolve<<<grid, kBlock, smem, stream_>>>(m.dev, st, rb, q_cur, q_counts + b, work + wk, q_nxt, q_counts + b + 1, b, rules_);
        } else if (coop) {
            LAUNCH_COOP(0, k_trace_coop, m.dev, cpar, st, q_cur, q_counts + b, work + wk, q_nxt, q_counts + b + 1, b, rules_);
        } else if (wide4) {
            k_trace<5><<<grid, kBlock, smem, stream_>>>(m.dev, st, q_cur, q_counts + b, work + wk, q_nxt, q_counts + b + 1, b, rules_);
        } else if (sq) {
            k_traverse_stream<false><<<grid_sq[0], kBlock, smem_sq, stream_>>>(m.dev, st, rb, q_cur, q_counts + b, work + wk, nullptr, q_nxt, q_counts + b + 1, b, rules_,
                                                                                 sq_refill_, sq_flush_, sq_steps_);
#if !SQ_INLINE_RESOLVE
            ++wk, ++launches;
            k_resolve<<<grid, kBlock, smem, stream_>>>(m.dev, st, rb, q_cur, q_counts + b, work + wk, q_nxt, q_counts + b + 1, b, rules_);
#endif
        } else if (dq) {
            k_traverse_dq<false><<<grid, kBlock, smem, stream_>>>(m.dev, st, rb, q_cur, q_counts + b, work + wk, q_nxt, q_counts + b + 1, b, rules_,
                                                                  dyn_refill_, dq_flush_, dyn_steps_);
        } else {
            LAUNCH_MODE(k_trace, m.dev, st, q_cur, q_counts + b, work + wk, q_nxt, q_counts + b + 1, b, rules_);
        }
        ++wk;
        ++launches, ++traced;
    }
    mark(3, 0);
    k_splat<<<num_sms_ * 4, 256, 0, stream_>>>(m.L.p, d_image_, d_out_, use_variance ? d_var_ : nullptr, P, samples,
                                               first_accum, write_output ? 1 : 0, gamma ? 1 : 0, rules_ == 0 ? 1 : 0, row0_ * w_,
                                               (row1_ > 0 ? std::min(row1_, h_) : h_) * w_);
    mark(-1, 0);
    k_fold_stats<<<1, 32, 0, stream_>>>(q_counts, traced, shaded, m.stats.p, sq_counts, pipeline_ >= 2 ? shaded : 0);
    launches += 2;
    CUDA_OK(cudaEventRecord(m.ev1, stream_));
    CUDA_OK(cudaGetLastError());
    host_stats_.samples += uint64_t(w_) * uint64_t(row_end() - row_begin()) * samples;
    host_stats_.launches += launches;
}

void Renderer::collect_kernel_times(int samples) {
    // called after the batch's stream sync: fold per-launch event deltas and queue sizes into the totals
    if (!kernel_timing_)
        return;
    Impl &m = *impl_;
    m.h_counts.resize(size_t(max_depth_) + 2);
    CUDA_OK(cudaMemcpy(m.h_counts.data(), m.counters.p, m.h_counts.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < m.kev_kind.size(); i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, m.kev[i], m.kev[i + 1]);
        const int kind = m.kev_kind[i];
        ktimes_.ms[kind] += ms;
        ktimes_.launches[kind] += 1;
        ktimes_.units[kind] += (kind == 0 || kind == 3) ? uint64_t(w_) * h_ * samples : uint64_t(m.h_counts[m.kev_bounce[i]]);
    }
}

int Renderer::queue_counts(int32_t *paths, int32_t *shadow, int n) {
    Impl &m = *impl_;
    CUDA_OK(cudaSetDevice(device_));
    CUDA_OK(cudaStreamSynchronize(stream_));
    const int D = max_depth_, have = std::min(n, D + 1);
    if (rules_ == 2 || int(m.counters.n) < 2 * (D + 2) || have <= 0)
        return 0;
    std::vector<int> h(size_t(2 * (D + 2)));
    CUDA_OK(cudaMemcpy(h.data(), m.counters.p, h.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int b = 0; b < have; b++) {
        if (paths)
            paths[b] = h[size_t(b)];
        if (shadow)
            shadow[b] = h[size_t(D + 2 + b)];
    }
    return have;
}

const float *Renderer::render_raw(bool gamma) {
    launch_batch(accum_cnt_ + 1, 1, gamma, true, variance_on_);
    CUDA_OK(cudaStreamSynchronize(stream_));
    accum_cnt_++; // only a sample that was actually rendered counts
    float ms = 0;
    cudaEventElapsedTime(&ms, impl_->ev0, impl_->ev1);
    host_stats_.device_ms += ms;
    auto_observe(ms, 1);
    collect_kernel_times(1);
    return reinterpret_cast<const float *>(d_out_);
}

const float *Renderer::render_batch(int spp, bool gamma) {
    if (spp <= 0)
        throw std::runtime_error("render_batch: spp must be positive");
    int done = 0;
    while (done < spp) {
        const int s = std::min(s_max_, spp - done);
        const bool last = done + s >= spp;
        launch_batch(accum_cnt_ + 1, s, gamma, last, variance_on_);
        accum_cnt_ += s;
        done += s;
        CUDA_OK(cudaStreamSynchronize(stream_)); // events are reused per batch
        float ms = 0;
        cudaEventElapsedTime(&ms, impl_->ev0, impl_->ev1);
        host_stats_.device_ms += ms;
        auto_observe(ms, s);
        collect_kernel_times(s);
    }
    return reinterpret_cast<const float *>(d_out_);
}

static inline int to_int_gamma(float x) { // reference: core/cuda_utils.cuh:111-113
    return int(powf(std::min(std::max(x, 0.f), 1.f), 1.f / 2.1f) * 255.f + .5f);
}
static inline int to_int_linear(float x) { return int(std::min(std::max(x, 0.f), 1.f) * 255.f); }

void Renderer::render_offline(int num_iter, bool gamma, uint8_t *rgba8) {
    // TracerBase::render semantics (wf_path_tracer.cu:107-173): seeds use the 0-based iteration index,
    // the sum image is exported as 8-bit with the 1/2.1 gamma of host_device.cuh:63-112
    if (num_iter <= 0)
        throw std::runtime_error("render_offline: num_iter must be positive");
    int done = 0;
    while (done < num_iter) {
        const int s = std::min(s_max_, num_iter - done);
        launch_batch(done, s, false, false, false);
        done += s;
        CUDA_OK(cudaStreamSynchronize(stream_));
        float ms = 0;
        cudaEventElapsedTime(&ms, impl_->ev0, impl_->ev1);
        host_stats_.device_ms += ms;
        auto_observe(ms, s);
    }
    std::vector<float> host(size_t(w_) * h_ * 4);
    download_accumulator(host.data());
    const float inv = 1.f / num_iter;
    for (size_t p = 0; p < size_t(w_) * h_; p++) {
        for (int c = 0; c < 3; c++)
            rgba8[4 * p + c] = uint8_t(gamma ? to_int_gamma(host[4 * p + c] * inv) : to_int_linear(host[4 * p + c] * inv));
        rgba8[4 * p + 3] = 255;
    }
}

void Renderer::download_output(float *host) const {
    CUDA_OK(cudaSetDevice(device_));
    CUDA_OK(cudaMemcpyAsync(host, d_out_, size_t(w_) * h_ * sizeof(float4), cudaMemcpyDeviceToHost, stream_));
    CUDA_OK(cudaStreamSynchronize(stream_));
}
void Renderer::download_accumulator(float *host) const {
    CUDA_OK(cudaSetDevice(device_));
    CUDA_OK(cudaMemcpyAsync(host, d_image_, size_t(w_) * h_ * sizeof(float4), cudaMemcpyDeviceToHost, stream_));
    CUDA_OK(cudaStreamSynchronize(stream_));
}
void Renderer::download_variance(float *host) const {
    if (!d_var_)
        throw std::runtime_error("variance buffer not enabled");
    CUDA_OK(cudaSetDevice(device_));
    CUDA_OK(cudaMemcpyAsync(host, d_var_, size_t(w_) * h_ * sizeof(float), cudaMemcpyDeviceToHost, stream_));
    CUDA_OK(cudaStreamSynchronize(stream_));
}

void Renderer::download_state(const char *which, void *host) const {
    const Impl &m = *impl_;
    const size_t P = size_t(w_) * h_;
    const std::string k = which;
    const void *src = nullptr;
    size_t elem = 16;
    if (k == "thp") src = m.thp.p;
    else if (k == "L") src = m.L.p;
    else if (k == "ray_o") src = m.ray_o.p;
    else if (k == "ray_d") src = m.ray_d.p;
    else if (k == "isect") src = m.isect.p;
    else if (k == "rng") src = m.rng.p, elem = 8;
    else if (k == "pdf") src = m.pdf.p, elem = 4;
    if (!src)
        throw std::runtime_error("download_state: unknown or unallocated array '" + k + "'");
    CUDA_OK(cudaSetDevice(device_));
    CUDA_OK(cudaStreamSynchronize(stream_));
    CUDA_OK(cudaMemcpy(host, src, P * elem, cudaMemcpyDeviceToHost));
}

void Renderer::reset() {
    CUDA_OK(cudaSetDevice(device_));
    CUDA_OK(cudaStreamSynchronize(stream_));
    CUDA_OK(cudaMemsetAsync(d_image_, 0, size_t(w_) * h_ * sizeof(float4), stream_));
    if (d_var_)
        CUDA_OK(cudaMemsetAsync(d_var_, 0, size_t(w_) * h_ * sizeof(float), stream_));
    accum_cnt_ = 0;
}

void Renderer::update_camera(const float cam[CPTS_CAM_COUNT]) { impl_->cam = camera_from_floats(cam); }

void Renderer::sync() const {
    CUDA_OK(cudaSetDevice(device_));
    CUDA_OK(cudaStreamSynchronize(stream_));
}

RenderStats Renderer::stats() {
    CUDA_OK(cudaSetDevice(device_));
    CUDA_OK(cudaStreamSynchronize(stream_));
    unsigned long long h[4] = {0, 0, 0, 0};
    CUDA_OK(cudaMemcpy(h, impl_->stats.p, sizeof(h), cudaMemcpyDeviceToHost));
    RenderStats s = host_stats_;
    s.closest_rays = h[0];
    s.shadow_rays = h[1];
    s.shaded = h[2];
    s.shadow_traced = rules_ == 2 ? h[1] : (pipeline_ >= 2 ? h[3] : h[1]);
    return s;
}

std::string Renderer::info() const {
    std::ostringstream os;
    os << "cpt_b200 wavefront renderer: " << w_ << "x" << h_ << ", max_depth " << max_depth_ << ", "
       << impl_->dev.num_prims << " primitives, " << impl_->dev.num_nodes << " BVH nodes (" << impl_->dev.num_cache
       << " cached), traversal=" << (auto_traversal_ ? (auto_state_ >= kAutoDecided ? (auto_choice_ == 7 ? (auto_shadow_ ? "auto -> stream (closest hit + shadow rays)" : "auto -> stream (closest hit)") : "auto -> ordered-256bit") : "auto (undecided)") : "") << (auto_traversal_ ? " / set: " : "") << (traversal_ == 0 ? "ordered" : (traversal_ == 1 ? "dfs" : (traversal_ == 2 ? "ordered-256bit" : (traversal_ == 3 ? "ordered-vote" : (traversal_ == 4 ? "cooperative" : (traversal_ == 5 ? "wide4 per thread" : (traversal_ == 6 ? "deferred leaves + refill" : "stream")))))))
       << ", wide layout " << impl_->wide_width << "-ary " << impl_->wide_nodes << " nodes (fill " << impl_->wide_fill << ", depth " << impl_->wide_depth
       << ", stack " << impl_->wide_max_stack << "; coop " << (coop_ok_ ? "ok" : "unavailable") << ": " << impl_->coop_blocks[0] << "/" << impl_->coop_blocks[1] << "/" << impl_->coop_blocks[2] << " CTAs/SM (trace/shadow/raygen) x " << impl_->coop_smem
       << " B smem, slab " << coop_slab_ << " nodes, flush " << coop_flush_ << " lanes)"
       << ", pair depth " << impl_->pair_depth << ", nesting violations " << impl_->nesting_violations << " (max " << impl_->max_overshoot << "), clipped references " << impl_->clipped_refs
       << ", refined leaves " << impl_->refined_leaves
       << ", L2 window " << (impl_->l2_window_bytes >> 20) << " MB, pipeline " << pipeline_ << ", rules " << (rules_ == 2 ? "volumetric-pt" : (rules_ ? "megakernel-pt" : "wavefront")) << ", samples in flight " << s_max_ << ", device " << device_ << ", counter " << accum_cnt_;
    return os.str();
}

} // namespace cptb
