// pt_device.cuh — device-side scene tables and the scalar path-tracing math
// (sampler, camera, triangle test, BSDF family, emitters) for the sm_100a
// wavefront kernels.
//
// Everything here is a POD / tag-switch restatement of what the reference does
// with device-heap polymorphic objects.  Each function names the reference
// lines whose arithmetic (operand order, explicit fma, approximate vs rounded
// reciprocal, float-vs-double literals) it has to reproduce so that radiance
// buffers match within the 1e-4 relative-L2 bar.
#pragma once
#include "vecmath.cuh"

namespace cptb {

// ---- constants (reference: src/core/constants.cuh:26-40) ----
constexpr float kEps = 1e-3f;       // EPSILON
constexpr float kThpEps = 1e-4f;    // THP_EPS
constexpr float kMaxDist = 1e7f;    // MAX_DIST
constexpr float kEnvDist = 5e3f;    // ENVMAP_DIST
constexpr float kPi = 3.1415926535897f;
constexpr float kInvPi = 1.f / kPi;
constexpr int kSeedScaler = 11467;  // wavefront_pt.cu:25
constexpr int kRRBounce = 2;        // wavefront_pt.cu:26
constexpr float kRRThreshold = 0.1; // wavefront_pt.cu:27

// lobe flags (reference: core/enums.cuh:41-51)
enum : int {
    LOBE_NONE = 0x00,
    LOBE_DIFFUSE = 0x01,
    LOBE_SPECULAR = 0x02,
    LOBE_GLOSSY = 0x04,
    LOBE_REFLECT = 0x10,
    LOBE_TRANSMIT = 0x20,
    LOBE_VOLUME = 0x40 // SCAT_VOLUME (enums.cuh:48): scattering inside a medium
};
enum : int { MAT_LAMBERT = 0, MAT_SPECULAR, MAT_TRANSLUCENT, MAT_PLASTIC, MAT_PLASTIC_FWD, MAT_GGX, MAT_DISPERSION, MAT_FORWARD };
enum : int { EM_NULL = 0, EM_POINT, EM_AREA, EM_AREA_SPOT, EM_ENVMAP };
enum : int { TEX_DIFF = 0, TEX_SPEC, TEX_GLOS, TEX_NORMAL, TEX_ROUGH };

// ray tag bits (reference: core/ray.cuh:37-97)
constexpr uint32_t TAG_ACTIVE = 0x10000000u, TAG_HIT = 0x20000000u, TAG_DELTA = 0x40000000u, TAG_INDEX = 0x0fffffffu;
constexpr uint32_t TAG_NEE_VISIBLE = 0x80000000u; // ours (split pipeline): k_shadow -> k_shade hand-over, never visible outside a bounce

// ---- device scene tables ----
struct Material { // 160 B
    int type, flags, penetrable, pad0;
    float4 k_d, k_s, k_g;
    float eta_t[3], kk[3];                              // conductor Fresnel terms
    float trans_scaler, thickness, eta, precomp_diff_f; // plastic family
    float raw[2];                                       // raw roughness_x/y (GGX) or ior (plastic) before derivation
    cudaTextureObject_t tex[5];
};

struct EmitterPOD { // 96 B
    int type, obj_ref, is_sphere, pad0;
    float4 Le;
    float3 pos;
    float cos_val;
    cudaTextureObject_t tex;
    float scale, azimuth, zenith, pad1;
    float4 rot; // envmap quaternion (x,y,z,w)
};

struct ObjectPOD { // reference CompactedObjInfo, object.cuh:31-65
    int bsdf_emitter, prim_offset, prim_num;
    float inv_area;
};

struct CameraPOD { // reference DeviceCamera, camera_model.cuh:40-48
    M3 R;
    V3 t;
    float inv_focal, hw, hh, sign_x, aperture, focal_dist;
};

struct DevScene {
    // reference-format acceleration structure (consumed as-is)
    const float4 *nodes;       // 2 float4 per DFS node {min.xyz, base}{max.xyz, cnt | -subtree}
    const uint4 *cache_nodes;  // fp16 near-root nodes
    const int *obj_idxs;       // per primitive object word
    const int *prim_orig;      // reference index of each primitive when refined leaves re-ordered them (nullptr: identity)
    const float4 *tris;        // kTriStride float4 per primitive {v0,a20}{e1,a21}{e2,a22}[pad]
    const float4 *norms;       // 3 float4 per primitive
    const uint4 *uvs;          // 3 half2 per primitive (+pad)
    const ObjectPOD *objects;
    const int *emitter_prims;
    const Material *mats;
    const EmitterPOD *emitters;
    // re-laid-out ordered-traversal structure built at renderer construction
    const float4 *wnodes;      // 4 float4 per inner node (see bvh_relayout.cpp)
    const float4 *wide;        // wide layout (scene_host.cpp build_wide_layout): 2 * width float4 per node, node 0 = root
    int num_nodes, num_cache, num_prims, num_emitters, envmap_id, wroot;
    const struct MediumPOD *media; // [0] = vacuum (volumetric path only)
    int num_media, cam_vol;
    const struct GridPOD *grids;     // grid media (volume.cuh)
    cudaTextureObject_t blackbody;   // emission colour table of grid media
    cudaTextureObject_t cie, d65; // spectral tables for the dispersion BSDF
};

// ---- sampler (reference: src/core/sampler.cuh:28-76) ----
struct Rng {
    uint32_t d0, d1;
    PT_DEV void seed(int seed, int offset) {
        const unsigned long long s = (unsigned long long)(int)(seed + offset); // sign-extending, int wrap-around
        d0 = (uint32_t)s ^ 0x12345678u;
        d1 = (uint32_t)(s >> 32) ^ 0x87654321u;
    }
    PT_DEV uint32_t next_u32() {
        uint32_t t = d0;
        t ^= t << 13;
        t ^= t >> 17;
        t ^= t << 5;
        const uint32_t out = t + d1;
        d1 = d0;
        d0 = out;
        return out;
    }
    PT_DEV float next1d() { return next_u32() * 2.3283064e-10f + (2.3283064e-10f / 2.0f); }
    // x is drawn first, then y
    PT_DEV V2 next2d() {
        const float a = next1d();
        const float b = next1d();
        return V2{a, b};
    }
};

// ---- sampling warps (reference: src/core/sampling.cuh:27-67) ----
PT_DEV V3 sample_cos_hemisphere(V2 uv, float &pdf) {
    float cos_theta = sqrtf(uv.x);
    float sin_theta = sqrtf(1. - uv.x); // double literal in the reference
    pdf = cos_theta * kInvPi;
    float sin_phi = 0, cos_phi = 0;
    sincospif(2.f * uv.y, &sin_phi, &cos_phi);
    return V3{cos_phi * sin_theta, sin_phi * sin_theta, cos_theta};
}
PT_DEV V3 sample_sphere_uniform(V2 uv, float &pdf) {
    float cos_theta = 2.f * uv.x - 1.f;
    float sin_theta = sqrtf(1.f - cos_theta * cos_theta);
    pdf = 0.25f * kInvPi;
    float sin_phi = 0, cos_phi = 0;
    sincospif(2.f * uv.y, &sin_phi, &cos_phi);
    return V3{cos_phi * sin_theta, sin_phi * sin_theta, cos_theta};
}
PT_DEV V2 sample_disk(V2 uv) {
    float r = sqrtf(uv.x);
    float s = 0, c = 0;
    sincospif(2.f * uv.y, &s, &c);
    return V2{r * c, r * s};
}

// ---- camera ray (reference: src/core/camera_model.cuh:58-87; orthographic branch is compiled out upstream only by
// convention, focal_distance==0 still selects it) ----
PT_DEV void camera_ray(const CameraPOD &cam, int px, int py, Rng &rng, V3 &o, V3 &d) {
    float x_pos = rng.next1d() + static_cast<float>(px), y_pos = rng.next1d() + static_cast<float>(py);
    float ndc_x = (x_pos - cam.hw) * cam.inv_focal * cam.sign_x;
    float ndc_y = (cam.hh - y_pos) * cam.inv_focal;
    if (cam.focal_dist == 0) {
        o = cam.t + col(cam.R, 1) * ndc_y + col(cam.R, 0) * ndc_x;
        d = normalized(rotate(cam.R, V3{0, 0, 1}));
        return;
    }
    V3 dir = normalized(rotate(cam.R, V3{ndc_x, ndc_y, 1.0f}));
    if (cam.aperture > 0.0f) {
        V3 focus = cam.t + dir * cam.focal_dist;
        V2 lens = sample_disk(rng.next2d()) * cam.aperture;
        V3 off = col(cam.R, 0) * lens.x + col(cam.R, 1) * lens.y;
        o = cam.t + off;
        d = normalized(focus - o);
    } else {
        o = cam.t;
        d = dir;
    }
}

// ---- triangle test on the 48 B precomputed-adjoint layout
// (reference: src/core/primitives.cuh:57-87, layout src/core/aos.cuh:113-164) ----
// CPTB_TRI64 (build knob, default 0): a triangle record padded from 48 to 64 bytes and read with two 256-bit loads instead of three
// 128-bit ones.  Measured on vader 1080p (profiles/r02_stream_experiments.md): traversal kernels unchanged (3.31 -> 3.32 / 1.52 ->
// 1.52 ms per sample), k_shade 0.95 -> 1.03 (its emitter sampling reads the records too, and the L2-pinned arena grows by 5 MB).
#ifndef CPTB_TRI64
#define CPTB_TRI64 0
#endif
constexpr int kTriStride = CPTB_TRI64 ? 4 : 3;
PT_DEV void load_triangle(const float4 *__restrict__ tris, int index, float4 &p0, float4 &v1, float4 &v2) {
#if CPTB_TRI64
    const float4 *p = tris + 4 * index;
    float pad0, pad1, pad2, pad3;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(p0.x), "=f"(p0.y), "=f"(p0.z), "=f"(p0.w), "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w)
                 : "l"(p));
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v2.x), "=f"(v2.y), "=f"(v2.z), "=f"(v2.w), "=f"(pad0), "=f"(pad1), "=f"(pad2), "=f"(pad3)
                 : "l"(p + 2));
#else
    p0 = __ldg(tris + 3 * index), v1 = __ldg(tris + 3 * index + 1), v2 = __ldg(tris + 3 * index + 2);
#endif
}
PT_DEV float hit_triangle(const float4 *__restrict__ tris, int index, V3 o, V3 d, float &su, float &sv,
                          float max_range) {
    float4 p0, v1, v2;
    load_triangle(tris, index, p0, v1, v2);
    V3 v = o - V3{p0.x, p0.y, p0.z};
    V3 t1 = V3{fmaf(v2.y, -d.z, d.y * v2.z), fmaf(-d.x, v2.z, v2.x * d.z), fmaf(v2.x, -d.y, d.x * v2.y)};
    V3 t2 = V3{fmaf(-d.y, v1.z, v1.y * d.z), fmaf(v1.x, -d.z, d.x * v1.z), fmaf(-d.x, v1.y, v1.x * d.y)};
    // det = fma(-a20, d.x, fma(e1.x, t1.x, e2.x*t2.x)): the contraction the reference's SASS ends up with
    // (FFMA R9 = d.x * -a20 + (FFMA e1.x*t1.x + FMUL e2.x*t2.x) before MUFU.RCP in raygen_primary_hit_shader)
    const float det = fmaf(-p0.w, d.x, mad2(v1.x, t1.x, v2.x, t2.x));
    const float inv_det = 1.f / det;
    const float a = dot(t1, v) * inv_det, b = dot(t2, v) * inv_det,
                t = mad3(p0.w, v.x, v1.w, v.y, v2.w, v.z) * inv_det;
    su = a;
    sv = b;
    return t * (a > 0 && b > 0 && a + b < 1 && t > kEps && t < max_range);
}

// shading normal + half-precision uv at a hit (reference: primitives.cuh:109-123, aos.cuh:305-319,
// vec2_half.cuh:182-186).  Returns {n.xyz, bits(half2 uv)}.
// NormalArray::eval (aos.cuh:305-319): interpolated, normalised vertex normal
PT_DEV V3 interp_normal(const DevScene &sc, int index, float u, float v) {
    const float4 a = __ldg(sc.norms + 3 * index), b = __ldg(sc.norms + 3 * index + 1), c = __ldg(sc.norms + 3 * index + 2);
    return normalized(lerp3(xyz(a), 1.f - u - v, xyz(b), u, xyz(c), v));
}
PT_DEV float4 make_interaction(const DevScene &sc, int index, float u, float v) {
    V3 n = interp_normal(sc, index, u, v);
    const uint4 pk = __ldg(sc.uvs + index);
    const __half2 h1 = *reinterpret_cast<const __half2 *>(&pk.x), h2 = *reinterpret_cast<const __half2 *>(&pk.y),
                  h3 = *reinterpret_cast<const __half2 *>(&pk.z);
    const __half w1 = __float2half(1.f - u - v), w2 = __float2half(u), w3 = __float2half(v);
    // (h1 * w1 + h2 * w2) + h3 * w3, every product and sum rounded to half
    const __half rx = (h1.x * w1 + h2.x * w2) + h3.x * w3;
    const __half ry = (h1.y * w1 + h2.y * w2) + h3.y * w3;
    const __half2 uvh = __halves2half2(rx, ry);
    float4 out;
    out.x = n.x, out.y = n.y, out.z = n.z;
    out.w = __uint_as_float(*reinterpret_cast<const uint32_t *>(&uvh));
    return out;
}
PT_DEV float2 uv_of(float4 isect) {
    const uint32_t bits = __float_as_uint(isect.w);
    return __half22float2(*reinterpret_cast<const __half2 *>(&bits));
}

// ---- texture helpers (reference: src/core/textures.cuh:79-117) ----
PT_DEV V3 shading_normal_world(const Material &m, float4 isect) {
    M3 R = frame_from_z(xyz(isect), false);
    float4 pn = make_float4(0, 0, 1, 0);
    if (m.tex[TEX_NORMAL] != 0) {
        const float2 uv = uv_of(isect);
        pn = tex2D<float4>(m.tex[TEX_NORMAL], uv.x, uv.y);
    }
    return rotate_t(R, normalized(V3{pn.x, pn.y, pn.z}));
}
PT_DEV V3 shading_normal_world(const Material &m, float4 isect, M3 &R) {
    R = frame_from_z(xyz(isect), false);
    float4 pn = make_float4(0, 0, 1, 0);
    if (m.tex[TEX_NORMAL] != 0) {
        const float2 uv = uv_of(isect);
        pn = tex2D<float4>(m.tex[TEX_NORMAL], uv.x, uv.y);
    }
    return rotate_t(R, normalized(V3{pn.x, pn.y, pn.z}));
}
PT_DEV V4 tex_or(cudaTextureObject_t t, float4 isect, float4 dflt) {
    if (t == 0)
        return from_f4(dflt);
    const float2 uv = uv_of(isect);
    return from_f4(tex2D<float4>(t, uv.x, uv.y));
}
PT_DEV V2 rough_or(const Material &m, float4 isect, V2 dflt) {
    if (m.tex[TEX_ROUGH] == 0)
        return dflt;
    const float2 uv = uv_of(isect);
    const float2 r = tex2D<float2>(m.tex[TEX_ROUGH], uv.x, uv.y);
    return V2{r.x, r.y};
}

// ---- Fresnel (reference: src/bsdf/fresnel.cuh:28-140) ----
PT_DEV float fresnel_simple(float eta, float cos_i) {
    eta = cos_i < 0.0f ? 1.0f / eta : eta;
    cos_i = cos_i < 0.0f ? -cos_i : cos_i;
    float sin_t2 = eta * eta * (1.0f - cos_i * cos_i), cos_t = 0, result = 1;
    if (sin_t2 < 1.0f) {
        cos_t = sqrtf(fmaxf(1.0f - sin_t2, 0.0f));
        float Rs = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
        float Rp = (eta * cos_t - cos_i) / (eta * cos_t + cos_i);
        result = (Rs * Rs + Rp * Rp) * 0.5f;
    }
    return result;
}
PT_DEV bool total_internal(float dot_n, float ni, float nr) {
    return (1.f - (ni * ni) / (nr * nr) * (1.f - dot_n * dot_n)) < 0.f;
}
PT_DEV float sgnf(float v) { return float(v > 0.f) - float(v < 0.f); }
PT_DEV V3 snell_refract(V3 incid, V3 n, float &cos_r2, float dot_n, float ni, float nr) {
    float ratio = ni / nr;
    cos_r2 = 1.f - (ratio * ratio) * (1. - dot_n * dot_n); // mixed double arithmetic as upstream
    return normalized(ratio * incid - ratio * dot_n * n + sgnf(dot_n) * sqrtf(fabsf(cos_r2)) * n) * float(cos_r2 > 0.f);
}
PT_DEV float fresnel_dielectric(float n_in, float n_out, float cos_inc, float cos_ref) {
    float n1i = n_in * cos_inc, n2i = n_out * cos_inc, n1r = n_in * cos_ref, n2r = n_out * cos_ref;
    float rs = (n1i - n2r) / (n1i + n2r);
    float rp = (n1r - n2i) / (n1r + n2i);
    return 0.5f * (rs * rs + rp * rp);
}
PT_DEV V3 sqrt3(V3 a) { return V3{sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)}; }
PT_DEV V4 fresnel_conductor(const Material &m, float cos_i) {
    const V3 eta_t{m.eta_t[0], m.eta_t[1], m.eta_t[2]}, k{m.kk[0], m.kk[1], m.kk[2]};
    cos_i = fminf(fmaxf(cos_i, -1), 1);
    float cos2 = cos_i * cos_i;
    float sin2 = 1. - cos2;
    V3 eta2 = eta_t * eta_t;
    V3 etak2 = k * k;
    V3 t0 = eta2 - etak2 - sin2;
    V3 a2b2 = sqrt3(t0 * t0 + 4 * eta2 * etak2);
    V3 t1 = a2b2 + cos2;
    V3 a = sqrt3(0.5f * (a2b2 + t0));
    V3 t2 = 2.f * cos_i * a;
    V3 Rs = (t1 - t2) / (t1 + t2);
    V3 t3 = cos2 * a2b2 + sin2 * sin2;
    V3 t4 = t2 * sin2;
    V3 Rp = Rs * (t3 - t4) / (t3 + t4);
    Rp = 0.5f * (Rp + Rs);
    return V4{Rp.x, Rp.y, Rp.z, 1.f};
}

// ---- GGX microfacet helpers (reference: src/impl/bsdf_ggx.cu:25-189) ----
PT_DEV float ggx_e(V3 l, float ax, float ay) {
    float c2 = l.z * l.z, ic2 = c2 == 0 ? 0 : __frcp_rn(c2);
    return (l.x * l.x * ax * ax + l.y * l.y * ay * ay) * ic2;
}
PT_DEV float ggx_lambda(V3 l, float ax, float ay) {
    float e = ggx_e(l, ax, ay);
    return e == 0 ? 0 : (-1.f + sqrtf(1.f + e)) * 0.5f;
}
PT_DEV float ggx_D(V3 l, float ax, float ay) {
    float c2 = l.z * l.z, ic2 = c2 == 0 ? 0 : __frcp_rn(c2);
    float e = (l.x * l.x / (ax * ax) + l.y * l.y / (ay * ay)) * ic2;
    return __frcp_rn(kPi * ax * ay * c2 * c2 * (1 + e) * (1 + e));
}
PT_DEV float ggx_G1(V3 l, float ax, float ay) { return 1.f / (1.f + ggx_lambda(l, ax, ay)); }
PT_DEV float ggx_G(V3 li, V3 lo, float ax, float ay) { return 1.f / (1.f + ggx_lambda(li, ax, ay) + ggx_lambda(lo, ax, ay)); }
PT_DEV void ggx_slopes(float cos_theta, V2 uv, V2 &slopes) {
    if (cos_theta == 1) {
        float r = sqrtf(uv.x / (1 - uv.y));
        float phi = 2.f * uv.y, sp = 0, cp = 0;
        sincospif(phi, &sp, &cp);
        slopes.x = r * cp;
        slopes.y = r * sp;
        return;
    }
    float sin_theta = sqrtf(fmaxf(0, 1.f - cos_theta * cos_theta));
    float tan_theta = sin_theta / cos_theta;
    float a = 1.f / tan_theta;
    float G1 = 2.f / (1 + sqrtf(1.f + 1.f / (a * a)));
    float A = 2.f * uv.x / G1 - 1;
    float tmp = fminf(1.f / (A * A - 1.f), 1e9);
    float B = tan_theta;
    float D = sqrtf(fmaxf(B * B * tmp * tmp - (A * A - B * B) * tmp, 0));
    float s1 = B * tmp - D;
    float s2 = B * tmp + D;
    slopes.x = (A < 0 || s2 > 1.f / tan_theta) ? s1 : s2;
    float sign = uv.y > 0.5f ? 1.f : -1.f;
    uv.y = 2.f * (uv.y - .5f) * sign;
    float z = (uv.y * (uv.y * (uv.y * 0.27385f - 0.73369f) + 0.46341f)) /
              (uv.y * (uv.y * (uv.y * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
    slopes.y = sign * z * sqrtf(1.f + slopes.x * slopes.x);
}
PT_DEV V2 ggx_sincos_phi(V3 v) {
    float s2 = fmaxf(1.f - v.z * v.z, 0), inv = rsqrtf(s2);
    bool zero = s2 == 0;
    return V2{zero ? 1 : fminf(fmaxf(v.x * inv, -1), 1), zero ? 0 : fminf(fmaxf(v.y * inv, -1), 1)};
}
PT_DEV V3 ggx_sample_wh(V3 li, float ax, float ay, V2 uv) {
    V3 ws = normalized(V3{li.x * ax, li.y * ay, li.z});
    V2 slope;
    ggx_slopes(ws.z, uv, slope);
    V2 cs = ggx_sincos_phi(ws);
    float tmp = cs.x * slope.x - cs.y * slope.y;
    slope.y = (cs.y * slope.x + cs.x * slope.y) * ay;
    slope.x = tmp * ax;
    return normalized(V3{-slope.x, -slope.y, 1.});
}
PT_DEV float ggx_pdf(V3 wh, V3 li, float ax, float ay) {
    float D = ggx_D(wh, ax, ay);
    float cos_ratio = fabsf(dot(li, wh)) / fabsf(li.z);
    return D * ggx_G1(li, ax, ay) * cos_ratio;
}

// ---- dispersion colour helpers (reference: src/impl/bsdf_dispersion.cu:26-44, src/core/xyz.cuh:47-57) ----
PT_DEV V4 wavelength_rgb(const DevScene &sc, float wl) {
    float cie_i = wl - 360.f, d65_i = wl - 300.f;
    V4 c = from_f4(tex1D<float4>(sc.cie, cie_i / 471.f));
    float spd = tex1D<float>(sc.d65, d65_i / 531.f);
    c *= spd / 22.2175f;
    float r = 3.240479f * c.x + -1.537150f * c.y + -0.498535f * c.z;
    float g = -0.969256f * c.x + 1.875991f * c.y + 0.041556f * c.z;
    float b = 0.055648f * c.x + -0.204043f * c.y + 1.057311f * c.z;
    V4 rgb = max4(mk4(r, g, b), splat4(0));
    return rgb / mk4(1.4979, 1.13591, 1.13159);
}

// det-refraction core shared by Translucent and Dispersion (reference: src/bsdf/translucent.cuh:43-98)
PT_DEV V4 dielectric_eval(V3 n, V3 out, V3 in, V4 ks, float eta, bool is_radiance) {
    float dn = dot(in, n);
    float ni = dn < 0 ? 1.f : eta;
    float nr = dn < 0 ? eta : 1.f, cos_r2 = 0;
    float eta2 = ni * ni / (nr * nr);
    V3 ret = normalized(advance(in, n, -2.f * dn)), refr = snell_refract(in, n, cos_r2, dn, ni, nr);
    bool rf = dot(out, ret) > 0.99999f, rr = dot(out, refr) > 0.99999f;
    return ks * float(rf | rr) * (rr && is_radiance ? eta2 : 1.f);
}
PT_DEV V3 dielectric_sample(V3 in, V3 n, V4 ks, float eta, V4 &thp, Rng &rng, float &pdf, int &lobe, bool is_radiance) {
    float dn = dot(in, n);
    float ni = dn < 0 ? 1.f : eta;
    float nr = dn < 0 ? eta : 1.f, cos_r2 = 0;
    float eta2 = ni * ni / (nr * nr);
    V3 ret = normalized(advance(in, n, -2.f * dn)), refr = snell_refract(in, n, cos_r2, dn, ni, nr);
    bool total = total_internal(dn, ni, nr);
    nr = fresnel_dielectric(ni, nr, fabsf(dn), sqrtf(fabsf(cos_r2)));
    bool reflect = total || rng.next1d() < nr; // no draw on total internal reflection
    ret = blend(ret, refr, reflect);
    pdf = total ? 1.f : (reflect ? nr : 1.f - nr);
    lobe = LOBE_SPECULAR | (total || reflect ? LOBE_REFLECT : LOBE_TRANSMIT);
    thp *= ks * (is_radiance && !reflect ? eta2 : 1.f);
    return ret;
}

PT_DEV bool dispersion_wavelength(const Material &m, V3 in, V3 out, V3 n, float &wl) {
    float ci = dot(n, in), co = dot(n, out), si = sqrtf(1.f - ci * ci), so = sqrtf(1.f - co * co);
    float eta = si > so ? si / so : so / si;
    wl = sqrtf(m.k_d.y / fmaxf(eta - m.k_d.x, 1e-5f));
    return wl > 360.f && wl < 360.f + 471.f;
}

// ======================= BSDF interface: pdf / eval / sample =======================
// `in` points towards the surface, `out` away from it.  index==material id only matters for textures,
// which live inside Material here.

PT_DEV float bsdf_pdf(const DevScene &sc, const Material &m, float4 it, V3 out, V3 in) {
    switch (m.type) {
    case MAT_LAMBERT: { // lambert.cuh:38-43
        const V3 n = shading_normal_world(m, it);
        return max(dot(n, out), 0.f) * kInvPi;
    }
    case MAT_PLASTIC: { // bsdf_plastic.cu:38-51
        const V3 n = shading_normal_world(m, it);
        float dwo = fabs(dot(out, n)), dwi = fabs(dot(in, n)), Fi = fresnel_simple(m.eta, dwi),
              sp = Fi / (Fi + m.trans_scaler * (1.0f - Fi));
        V3 rd = -reflection(in, n);
        return dot(rd, out) < 1.f - kThpEps ? kInvPi * dwo * (1.f - sp) : sp;
    }
    case MAT_PLASTIC_FWD: { // bsdf_plastic.cu:148-162
        const V3 n = shading_normal_world(m, it);
        float dwi = dot(in, n), Fi = fresnel_simple(m.eta, -dwi), sp = Fi / (Fi + m.trans_scaler * (1.0f - Fi));
        V3 rd = -reflection(in, n);
        float p = 0;
        p = dot(rd, out) < 1.f - kThpEps ? p : sp;
        p = dot(in, out) < 1.f - kThpEps ? p : 1.f - sp;
        return p;
    }
    case MAT_GGX: { // bsdf_ggx.cu:203-218
        M3 R;
        shading_normal_world(m, it, R);
        const V3 li = -rotate(R, in), lo = rotate(R, out), wh = normalized(lo + li);
        const V2 al = rough_or(m, it, V2{m.k_s.x, m.k_s.y});
        float pv = ggx_pdf(wh, li, al.x, al.y);
        bool nsh = (li.z > 0) ^ (lo.z > 0);
        return nsh ? 0 : pv * __frcp_rn(4.f * dot(wh, li));
    }
    case MAT_DISPERSION: { // bsdf_dispersion.cu:46-61
        const V3 n = shading_normal_world(m, it);
        bool ip = dot(n, in) > 0, op = dot(n, out) > 0;
        float p = 0;
        if ((ip ^ op) == false) {
            float wl = 0;
            p = dispersion_wavelength(m, in, out, n, wl);
            float eta = m.k_d.x + m.k_d.y / (wl * wl), ci = dot(in, n), F = fresnel_simple(eta, -ci);
            p *= (1.f - F) / 471.f;
        }
        return p;
    }
    case MAT_FORWARD: return static_cast<float>(dot(in, out) == 1.f); // forward.cuh:38-41
    default: return 0.f; // specular / det-refraction are delta lobes
    }
}

PT_DEV V4 bsdf_eval(const DevScene &sc, const Material &m, float4 it, V3 out, V3 in) {
    switch (m.type) {
    case MAT_LAMBERT: { // lambert.cuh:45-57
        const V3 n = shading_normal_world(m, it);
        float cos_term = dot(n, out);
        float dot_in = dot(n, in);
        float same_side = (dot_in > 0) ^ (cos_term > 0);
        return tex_or(m.tex[TEX_DIFF], it, m.k_d) * max(0.f, cos_term) * kInvPi * same_side;
    }
    case MAT_SPECULAR: { // specular.cuh:43-51
        const V3 n = shading_normal_world(m, it);
        V3 rd = normalized(advance(in, n, -2.f * dot(in, n)));
        return tex_or(m.tex[TEX_SPEC], it, m.k_s) * float(dot(out, rd) > 0.99999f);
    }
    case MAT_TRANSLUCENT: { // translucent.cuh:100-109 (is_radiance=false via the base-class default)
        const V3 n = shading_normal_world(m, it);
        const V4 ks = tex_or(m.tex[TEX_SPEC], it, m.k_s);
        float eta = rough_or(m, it, V2{m.k_d.x, m.k_d.x}).x;
        return dielectric_eval(n, out, in, ks, eta, false);
    }
    case MAT_PLASTIC: { // bsdf_plastic.cu:53-80 (is_radiance default true, unused)
        const V3 n = shading_normal_world(m, it);
        float rwo = dot(out, n), rwi = dot(in, n);
        float dwo = fabsf(rwo), dwi = fabsf(rwi), Fi = fresnel_simple(m.eta, dwi), Fo = fresnel_simple(m.eta, dwo);
        V3 rd = -reflection(in, n);
        V4 kdiff = tex_or(m.tex[TEX_DIFF], it, m.k_d);
        V4 brdf = (kInvPi * (1.0f - Fi) * (1.0f - Fo) * m.eta * m.eta) * dwo * (kdiff / (-kdiff * m.precomp_diff_f + 1.f)) *
                  exp_xyz(tex_or(m.tex[TEX_GLOS], it, m.k_g) * m.thickness * (-1.0f / dwo - 1.0f / dwi));
        brdf += dot(rd, out) < 1.f - kThpEps ? splat4(0) : splat4(Fi, 1) * tex_or(m.tex[TEX_SPEC], it, m.k_s);
        brdf = (m.penetrable > 0 || (rwo > 0) ^ (rwi > 0)) ? brdf : splat4(0, 1);
        return brdf;
    }
    case MAT_PLASTIC_FWD: { // bsdf_plastic.cu:164-187
        const V3 n = shading_normal_world(m, it);
        float dwi = dot(in, n), Fi = fresnel_simple(m.eta, fabsf(dwi));
        V3 rd = -reflection(in, n);
        V4 brdf = dot(in, out) < 1.f - kThpEps
                      ? splat4(0, 1)
                      : (1.0f - Fi) * (1.0f - Fi) * tex_or(m.tex[TEX_DIFF], it, m.k_d) * m.eta * m.eta *
                            exp_xyz(tex_or(m.tex[TEX_GLOS], it, m.k_g) * m.thickness * (-2.f / fabsf(dwi)));
        brdf += dot(rd, out) < 1.f - kThpEps ? brdf : splat4(Fi, 1) * tex_or(m.tex[TEX_SPEC], it, m.k_s);
        return brdf;
    }
    case MAT_GGX: { // bsdf_ggx.cu:220-232 + GGX::eval 170-188
        M3 R;
        const V3 n = shading_normal_world(m, it, R);
        const V2 al = rough_or(m, it, V2{m.k_s.x, m.k_s.y});
        V3 li = -rotate(R, in), lo = rotate(R, out);
        V3 wh = face_forward(normalized(lo + li));
        V4 fres = fresnel_conductor(m, fabsf(dot(lo, wh)));
        float ci = li.z, co = lo.z;
        bool nsh = (ci > 0) ^ (co > 0);
        V4 g = nsh ? splat4(0) : (ggx_D(wh, al.x, al.y) * fres * __frcp_rn(4.f * ci * co) * ggx_G(li, lo, al.x, al.y));
        return tex_or(m.tex[TEX_GLOS], it, m.k_g) * g * fmaxf(0, dot(out, n));
    }
    case MAT_DISPERSION: { // bsdf_dispersion.cu:63-76; called through the BSDF base pointer, so the base-class
                           // default is_radiance=false applies (bsdf.cuh:54-56), not the override's `true`
        const V3 n = shading_normal_world(m, it);
        float wl = 0;
        bool valid = dispersion_wavelength(m, in, out, n, wl);
        const V4 ks = tex_or(m.tex[TEX_SPEC], it, m.k_s);
        float eta = valid ? m.k_d.x + m.k_d.y / (wl * wl) : m.k_d.x;
        V4 r = dielectric_eval(n, out, in, ks, eta, false);
        r *= valid ? wavelength_rgb(sc, wl) : splat4(1);
        return r;
    }
    case MAT_FORWARD: return splat4(float(dot(in, out) == 1.f), 1); // forward.cuh:43-47
    default: return splat4(0);
    }
}

// samples the next direction, updates throughput / pdf / sampled lobe flags
PT_DEV V3 bsdf_sample(const DevScene &sc, const Material &m, V3 in, float4 it, V4 &thp, float &pdf, Rng &rng, int &lobe) {
    switch (m.type) {
    case MAT_LAMBERT: { // lambert.cuh:59-75
        V3 local = sample_cos_hemisphere(rng.next2d(), pdf);
        const V3 n = shading_normal_world(m, it);
        V3 out = to_world(n, local);
        float di = dot(n, in);
        float dout = dot(n, out);
        thp *= tex_or(m.tex[TEX_DIFF], it, m.k_d) * float((di > 0) ^ (dout > 0));
        lobe = m.flags;
        return out;
    }
    case MAT_SPECULAR: { // specular.cuh:53-65
        lobe = m.flags;
        const V3 n = shading_normal_world(m, it);
        float idn = dot(in, n);
        pdf = 1.f;
        thp *= tex_or(m.tex[TEX_SPEC], it, m.k_s);
        return -reflection(in, n, idn);
    }
    case MAT_TRANSLUCENT: { // translucent.cuh:111-121 (is_radiance false)
        const V3 n = shading_normal_world(m, it);
        const V4 ks = tex_or(m.tex[TEX_SPEC], it, m.k_s);
        float eta = rough_or(m, it, V2{m.k_d.x, m.k_d.x}).x;
        return dielectric_sample(in, n, ks, eta, thp, rng, pdf, lobe, false);
    }
    case MAT_PLASTIC: { // bsdf_plastic.cu:82-130
        const V3 n = shading_normal_world(m, it);
        float rdi = dot(n, in), di = fabsf(rdi);
        float Fi = fresnel_simple(m.eta, di), sp = Fi / (Fi + m.trans_scaler * (1.0f - Fi));
        V3 out;
        if (rng.next1d() < sp) {
            out = -reflection(in, n);
            pdf = sp;
            thp *= splat4(Fi / sp, 1) * tex_or(m.tex[TEX_SPEC], it, m.k_s);
            lobe = LOBE_REFLECT | LOBE_SPECULAR;
        } else {
            float dummy = 1;
            V3 local = sample_cos_hemisphere(rng.next2d(), dummy);
            float Fo = fresnel_simple(m.eta, local.z);
            V4 kdiff = tex_or(m.tex[TEX_DIFF], it, m.k_d);
            V4 lthp = ((1.0f - Fi) * (1.0f - Fo) * m.eta * m.eta) * (kdiff / (-kdiff * m.precomp_diff_f + 1.f)) *
                      exp_xyz(tex_or(m.tex[TEX_GLOS], it, m.k_g) * m.thickness * (-1.0f / local.z - 1.0f / di));
            pdf = kInvPi * local.z * (1.0f - sp);
            thp *= __frcp_rn(1.f - sp) * lthp;
            out = to_world(n, local);
            lobe = LOBE_REFLECT | LOBE_DIFFUSE;
        }
        thp = (m.penetrable > 0 || (rdi > 0) ^ (dot(out, n) > 0)) ? thp : splat4(0, 1);
        return out;
    }
    case MAT_PLASTIC_FWD: { // bsdf_plastic.cu:189-228
        const V3 n = shading_normal_world(m, it);
        float di = dot(n, in), Fi = fresnel_simple(m.eta, fabsf(di)), sp = Fi / (Fi + m.trans_scaler * (1.0f - Fi));
        V3 out;
        if (rng.next1d() < sp) {
            out = -reflection(in, n, di);
            pdf = sp;
            thp *= splat4(Fi / sp, 1) * tex_or(m.tex[TEX_SPEC], it, m.k_s);
            lobe = LOBE_REFLECT | LOBE_SPECULAR;
        } else {
            V4 lthp = ((1.0f - Fi) * (1.0f - Fi) * m.eta * m.eta) * tex_or(m.tex[TEX_DIFF], it, m.k_d) *
                      exp_xyz(tex_or(m.tex[TEX_GLOS], it, m.k_g) * m.thickness * (-2.0f / fabsf(di)));
            pdf = 1.0f - sp;
            thp *= __frcp_rn(1.f - sp) * lthp;
            out = in;
            lobe = LOBE_TRANSMIT | LOBE_SPECULAR;
        }
        return out;
    }
    case MAT_GGX: { // bsdf_ggx.cu:234-266
        M3 R;
        const V3 n = shading_normal_world(m, it, R);
        const V2 al = rough_or(m, it, V2{m.k_s.x, m.k_s.y});
        const V3 li = -rotate(R, in), wh = ggx_sample_wh(li, al.x, al.y, rng.next2d());
        float D = ggx_D(wh, al.x, al.y);
        float dim = dot(li, wh);
        pdf = D * ggx_G1(li, al.x, al.y) * fabsf(dim / li.z);
        pdf = (pdf > 0 && dim > 0) ? (pdf * __frcp_rn(4.f * dim)) : 0;
        V3 lref = reflection(li, wh, dim), ref = rotate_t(R, lref);
        float ci = li.z, co = lref.z;
        V4 fres = fresnel_conductor(m, fabsf(dot(lref, wh)));
        if (ci > 0 && co > 0 && pdf > 0) {
            thp *= (1.f / pdf) * tex_or(m.tex[TEX_GLOS], it, m.k_g) * D * ggx_G(li, lref, al.x, al.y) *
                   __frcp_rn(4.f * ci * co) * fres * fmaxf(dot(n, ref), 0);
        }
        lobe = m.flags;
        return ref;
    }
    case MAT_DISPERSION: { // bsdf_dispersion.cu:78-92 (is_radiance=false: base-class default, see bsdf_eval)
        float wl = rng.next1d() * 471.f + 360.f;
        float eta = m.k_d.x + m.k_d.y / (wl * wl);
        const V3 n = shading_normal_world(m, it);
        const V4 ks = tex_or(m.tex[TEX_SPEC], it, m.k_s);
        V3 r = dielectric_sample(in, n, ks, eta, thp, rng, pdf, lobe, false);
        thp *= wavelength_rgb(sc, wl);
        pdf *= 1.f / 471.f;
        return r;
    }
    default: // forward.cuh:49-56
        pdf = 1.f;
        lobe = m.flags;
        return in;
    }
}

// ======================= emitters (reference: src/core/emitter.cuh:32-359, src/impl/emitter.cu:26-73) =======================
// q * (v,0) * conj(q) (quaternion.cuh:54-84).  The rounding points are pinned to the reference's SASS for
// EnvMapEmitter::eval_le / ::sample: every component is a left-to-right fma chain whose first product pair
// is fma(a, b, +-(c * d)); the zero terms of the pure-vector operand are kept (they only matter for -0).
PT_DEV V3 qrotate(float4 q, V3 v) {
    const float X = fmaf(-v.y, q.z, fmaf(v.z, q.y, fmaf(v.x, q.w, __fmul_rn(0.f, q.x))));
    const float Y = fmaf(v.x, q.z, fmaf(0.f, q.y, fmaf(v.y, q.w, -__fmul_rn(v.z, q.x))));
    const float Z = fmaf(0.f, q.z, fmaf(-v.x, q.y, fmaf(v.z, q.w, __fmul_rn(v.y, q.x))));
    const float W = fmaf(-v.z, q.z, fmaf(-v.y, q.y, fmaf(0.f, q.w, -__fmul_rn(v.x, q.x))));
    return V3{fmaf(q.y, Z, fmaf(q.z, -Y, fmaf(q.w, X, -__fmul_rn(q.x, W)))),
              fmaf(q.x, -Z, fmaf(q.w, Y, fmaf(q.z, X, -__fmul_rn(q.y, W)))),
              fmaf(q.w, Z, fmaf(q.x, Y, fmaf(-q.y, X, -__fmul_rn(q.z, W))))};
}
PT_DEV V4 envmap_lookup(const EmitterPOD &e, V3 dir) {
    V3 d = qrotate(e.rot, dir);
    float ty = acosf(d.z) * kInvPi, px = (atan2f(d.y, d.x) + kPi) * kInvPi * 0.5f;
    return from_f4(tex2D<float4>(e.tex, px, ty)) * e.scale;
}

// radiance leaving an emitter towards -inci_dir at a hit (eval_le)
PT_DEV V4 emitter_eval(const EmitterPOD &e, V3 inci, float4 it, bool has_it) {
    switch (e.type) {
    case EM_AREA: { // emitter.cuh:207-212
        V4 le = from_f4(e.Le);
        if (e.tex != 0) {
            const float2 uv = uv_of(it);
            le = from_f4(tex2D<float4>(e.tex, uv.x, uv.y)) * e.Le.w;
        }
        return blend(le, mk4(0, 0, 0, 1), dot(inci, xyz(it)) < 0);
    }
    case EM_AREA_SPOT: { // emitter.cuh:295-301
        V4 le = from_f4(e.Le);
        if (e.tex != 0) {
            const float2 uv = uv_of(it);
            le = from_f4(tex2D<float4>(e.tex, uv.x, uv.y)) * e.Le.w;
        }
        return blend(le, mk4(0, 0, 0, 1), dot(inci, xyz(it)) < -e.cos_val);
    }
    case EM_ENVMAP: return envmap_lookup(e, inci); // emitter.cu:67-73
    default: return mk4(0, 0, 0, 1);               // null emitter and point source
    }
}

// sample a point on the emitter for NEE; returns the sampled position
PT_DEV V3 emitter_sample(const DevScene &sc, const EmitterPOD &e, V3 hit_pos, V3 hit_n, V4 &le, float &pdf, V2 uv, int prim) {
    switch (e.type) {
    case EM_POINT: { // emitter.cuh:108-115, distance_attenuate :32-34
        V3 diff = V3{e.pos.x, e.pos.y, e.pos.z} - hit_pos;
        le = from_f4(e.Le) * min(1.f / max(length2(diff), 1e-5f), 1.f);
        return V3{e.pos.x, e.pos.y, e.pos.z};
    }
    case EM_AREA:
    case EM_AREA_SPOT: { // emitter.cuh:157-183 / 235-262 (triangle-bound; sphere-bound emitters are outside TRIANGLE_ONLY)
        float ssum = uv.x + uv.y;
        if (!(ssum < 1.f))
            uv = V2{-uv.x + 1.f, -uv.y + 1.f};
        const float4 p0 = __ldg(sc.tris + kTriStride * prim), e1 = __ldg(sc.tris + kTriStride * prim + 1), e2 = __ldg(sc.tris + kTriStride * prim + 2);
        V3 sampled = V3{__fadd_rn(mad2(uv.x, e1.x, uv.y, e2.x), p0.x), __fadd_rn(mad2(uv.x, e1.y, uv.y, e2.y), p0.y),
                        __fadd_rn(mad2(uv.x, e1.z, uv.y, e2.z), p0.z)};
        const float4 na = __ldg(sc.norms + 3 * prim), nb = __ldg(sc.norms + 3 * prim + 1), nc = __ldg(sc.norms + 3 * prim + 2);
        V3 normal = normalized(lerp3(xyz(na), 1.f - uv.x - uv.y, xyz(nb), uv.x, xyz(nc), uv.y));
        V3 to_hit = hit_pos - sampled;
        pdf *= length2(to_hit);
        to_hit = normalized(to_hit);
        ssum = dot(normal, to_hit);
        pdf *= float(ssum > 0) / ssum;
        V4 base = from_f4(e.Le);
        if (e.tex != 0) {
            const uint4 pk = __ldg(sc.uvs + prim);
            const __half2 h1 = *reinterpret_cast<const __half2 *>(&pk.x), h2 = *reinterpret_cast<const __half2 *>(&pk.y),
                          h3 = *reinterpret_cast<const __half2 *>(&pk.z);
            const __half w1 = __float2half(1.f - uv.x - uv.y), w2 = __float2half(uv.x), w3 = __float2half(uv.y);
            const float tu = __half2float((h1.x * w1 + h2.x * w2) + h3.x * w3),
                        tv = __half2float((h1.y * w1 + h2.y * w2) + h3.y * w3);
            base = from_f4(tex2D<float4>(e.tex, tu, tv)) * e.Le.w;
        }
        le = base * float(e.type == EM_AREA ? (ssum > 0) : (ssum > e.cos_val));
        return sampled;
    }
    case EM_ENVMAP: { // emitter.cu:26-46
        V3 dir;
        if (length2(hit_n) > kEps)
            dir = to_world(hit_n, sample_cos_hemisphere(uv, pdf));
        else
            dir = to_world(V3{0, 0, 1}, sample_sphere_uniform(uv, pdf));
        V3 pos = hit_pos + kEnvDist * dir;
        le = envmap_lookup(e, dir);
        return pos;
    }
    default: // abstract emitter, emitter.cuh:64-71
        pdf = 1;
        le = splat4(0);
        return V3{0, 0, 0};
    }
}

} // namespace cptb
