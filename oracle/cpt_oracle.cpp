/*
 * cpt_oracle.cpp — CPU restatement of the reference's wavefront path-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (cuda-pt_b200/, include/) links, imports or
 * executes this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker / the timed CPU baseline.
 *
 * What it restates (reference = Enigmatisms/cuda-pt @ 8ee28d0a, paths relative to /root/reference):
 *   sampler            src/core/sampler.cuh:28-76
 *   warps              src/core/sampling.cuh:27-67
 *   camera             src/core/camera_model.cuh:58-87
 *   slab / fp16 nodes  src/core/aabb.cuh:58-65, src/core/bvh.cuh:154-251
 *   traversal          src/renderer/tracing_func.cuh:44-99 (any hit), :117-181 (closest hit)
 *   triangle test      src/core/primitives.cuh:57-87, layout src/core/aos.cuh:134-164
 *   interaction        src/core/primitives.cuh:109-123, src/core/aos.cuh:305-319, src/core/vec2_half.cuh:182-186
 *   frames             src/core/so3.cuh:201-252
 *   BSDFs              src/bsdf/{lambert,specular,translucent,plastic,ggx_conductor,dispersion,forward}.cuh,
 *                      src/impl/bsdf_{ggx,plastic,dispersion}.cu, src/bsdf/fresnel.cuh:28-140
 *   emitters           src/core/emitter.cuh:32-359, src/impl/emitter.cu:26-73 (point, area, area-spot, environment map)
 *   image textures     src/core/textures.cuh:79-117 (lookups), texture-object set-up src/impl/textures.cu:134-165;
 *                      texels are the ones the reference decoded (scene dump); bilinear filter restated with the
 *                      8-bit fractional weights of the texture unit
 *   per-bounce logic   src/pt_impl/wavefront_pt.cu:71-314 (raygen, shade, closest-hit + RR), :476-503 (splat),
 *                      host loop src/pt_impl/wf_path_tracer.cu:175-230, emitter pick tracing_func.cuh:210-225
 *   megakernel rules   src/pt_impl/megakernel_pt.cu:59-186 ("pt": lobe caps, throughput cut-off, RR from bounce 1)
 *   participating media src/pt_impl/megakernel_vpt.cu:59-456, src/volume/homogeneous.cuh:45-88,
 *                      src/volume/henyey_greenstein.cuh:28-108, src/core/medium.cuh:72-112 ; grid media src/volume/grid.cuh:51-92, src/impl/vol_grid.cu:97-211
 *
 * It is a TOLERANCE oracle: the reference runs on the GPU with --use_fast_math (approximate
 * rcp/sqrt/div, flush-to-zero), which cannot be reproduced bit-for-bit on a CPU.
 * It is pinned against the committed reference frames in tests/golden/ref (same scene dump, same seeds):
 * per-pixel sample streams are identical, so images agree to ~1e-4 relative L2 at 64 spp apart from
 * a handful of paths that flip at discrete decisions.  Integer machinery (RNG, keys, emitter picks,
 * half conversion) is exact and has known-answer tests.
 */
#include "cpts_format.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct V3 {
    float x, y, z;
};
struct V4 {
    float x, y, z, w;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float len2(V3 a) { return dot(a, a); }
inline V3 normalize(V3 a) { return a * (1.f / std::sqrt(len2(a))); }
inline V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 operator*(V4 a, V4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline V4 operator*(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline V4 operator/(V4 a, V4 b) { return {a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w}; }
inline V4 operator+(V4 a, float s) { return {a.x + s, a.y + s, a.z + s, a.w + s}; }
inline V4 v4(float v) { return {v, v, v, v}; }
inline V4 v4(float v, float a) { return {v, v, v, a}; }
inline V4 v4(const float *p) { return {p[0], p[1], p[2], p[3]}; }
inline V4 expxyz(V4 a) { return {std::exp(a.x), std::exp(a.y), std::exp(a.z), 1.f}; }

constexpr float EPS = 1e-3f, THP_EPS = 1e-4f, MAX_DIST = 1e7f, PI = 3.1415926535897f, INV_PI = 1.f / PI;
constexpr int SEED_SCALER = 11467, RR_BOUNCE = 2;
constexpr float RR_THRESHOLD = 0.1f;
enum { LOBE_DIFFUSE = 1, LOBE_SPECULAR = 2, LOBE_GLOSSY = 4, LOBE_REFLECT = 0x10, LOBE_TRANSMIT = 0x20, LOBE_VOLUME = 0x40 };

// ---- half helpers -------------------------------------------------------------------------------
float half_to_float(uint16_t h) {
    const uint32_t sign = uint32_t(h & 0x8000u) << 16, exp = (h >> 10) & 0x1F, man = h & 0x3FFu;
    uint32_t out;
    if (exp == 0) {
        if (man == 0)
            out = sign;
        else {
            int e = -1;
            uint32_t m = man;
            do {
                ++e;
                m <<= 1;
            } while ((m & 0x400u) == 0);
            out = sign | uint32_t(127 - 15 - e) << 23 | ((m & 0x3FFu) << 13);
        }
    } else if (exp == 31)
        out = sign | 0x7F800000u | (man << 13);
    else
        out = sign | (exp + 112) << 23 | (man << 13);
    float f;
    memcpy(&f, &out, 4);
    return f;
}
uint16_t float_to_half(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u)
        return uint16_t(sign | 0x7C00u | (x > 0x7F800000u ? 0x0200u : 0u));
    if (x >= 0x477FF000u)
        return uint16_t(sign | 0x7C00u);
    if (x < 0x33000001u)
        return uint16_t(sign);
    const int exp = int(x >> 23) - 127;
    const uint32_t man = (x & 0x7FFFFFu) | 0x800000u;
    const int shift = exp < -14 ? 13 + (-14 - exp) : 13;
    uint32_t hm = man >> shift;
    const uint32_t rem = man & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (hm & 1u)))
        hm++;
    return uint16_t(sign | (exp < -14 ? hm : ((uint32_t(exp + 15) << 10) + (hm - 0x400u))));
}
inline float hround(float f) { return half_to_float(float_to_half(f)); }

// ---- sampler (sampler.cuh:36-72) ------------------------------------------------------------------
struct Rng {
    uint32_t d0, d1;
    void seed(int s, int offset) {
        const unsigned long long v = (unsigned long long)(int)(uint32_t(s) + uint32_t(offset));
        d0 = uint32_t(v) ^ 0x12345678u;
        d1 = uint32_t(v >> 32) ^ 0x87654321u;
    }
    uint32_t u32() {
        uint32_t t = d0;
        t ^= t << 13;
        t ^= t >> 17;
        t ^= t << 5;
        const uint32_t out = t + d1;
        d1 = d0;
        d0 = out;
        return out;
    }
    float f() { return float(u32()) * 2.3283064e-10f + (2.3283064e-10f / 2.0f); }
};

// ---- scene ----------------------------------------------------------------------------------------
struct Scene {
    int32_t cfg[CPTS_CFG_COUNT];
    float cam[CPTS_CAM_COUNT];
    std::vector<float> nodes;       // 8 floats per node
    std::vector<uint32_t> cache;    // 4 words per cached node
    std::vector<int32_t> obj_idxs, emitter_prims;
    std::vector<float> v[3], n[3], uv[3];
    std::vector<CptsObject> objects;
    std::vector<CptsBsdf> bsdfs;
    std::vector<CptsEmitter> emitters;
    std::vector<float> tris; // 12 floats per primitive
    std::vector<float> cie, d65;
    std::vector<CptsMedium> media;     // [0] = vacuum (volumetric path)
    std::vector<CptsGrid> grids;       // grid media: brick tables + pools (cpts_format.h), emission colour table
    std::vector<int32_t> grid_bricks;
    std::vector<float> grid_data, blackbody;
    std::vector<CptsTexDesc> tex_desc; // image textures as the reference decoded them (float texels)
    std::vector<float> tex_data;
    float env_rot[4] = {0, 0, 0, 1};   // environment-map rotation quaternion (emitter.cuh:349-353)
    // derived material parameters (device constructors of the reference)
    std::vector<float> alpha_x, alpha_y, eta, diff_f;
    int num_nodes() const { return int(nodes.size() / 8); }
};

template <typename T> void read_vec(FILE *fp, const CptsSection &s, std::vector<T> &dst) {
    const size_t bytes = size_t(s.elem_size) * s.count;
    dst.resize(bytes / sizeof(T));
    if (bytes && fread(dst.data(), 1, bytes, fp) != bytes)
        throw std::string("truncated cpts section");
}

float fresnel_simple(float eta, float ci);

float roughness_to_alpha(float r) { // cuda_utils.cuh:115-120
    r = std::max(r, 1e-3f);
    const float x = std::log(r);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
float diffuse_fresnel(float ior) { // fresnel.cuh:62-75
    const int n = 131072;
    double acc = 0.0;
    float fb = fresnel_simple(ior, 0.0f);
    for (int i = 1; i <= n; i++) {
        const float c2 = float(i) / n;
        const float fa = fresnel_simple(ior, std::min(std::sqrt(c2), 1.0f));
        acc += double(fa + fb) * (0.5f / n);
        fb = fa;
    }
    return float(acc);
}

Scene *load_scene(const char *path) {
    FILE *fp = fopen(path, "rb");
    if (!fp)
        throw std::string("cannot open ") + path;
    CptsHeader h{};
    if (fread(&h, sizeof(h), 1, fp) != 1 || memcmp(h.magic, CPTS_MAGIC, 8) != 0) {
        fclose(fp);
        throw std::string("bad cpts magic");
    }
    Scene *sc = new Scene();
    for (uint32_t i = 0; i < h.n_sections; i++) {
        CptsSection s{};
        if (fread(&s, sizeof(s), 1, fp) != 1)
            break;
        s.name[sizeof(s.name) - 1] = 0;
        const std::string name = s.name;
        const size_t bytes = size_t(s.elem_size) * s.count;
        std::vector<int32_t> ti;
        std::vector<float> tf;
        if (name == "config") {
            read_vec(fp, s, ti);
            std::copy(ti.begin(), ti.end(), sc->cfg);
        } else if (name == "camera") {
            read_vec(fp, s, tf);
            std::copy(tf.begin(), tf.end(), sc->cam);
        } else if (name == "nodes")
            read_vec(fp, s, sc->nodes);
        else if (name == "cache_nodes")
            read_vec(fp, s, sc->cache);
        else if (name == "obj_idxs")
            read_vec(fp, s, sc->obj_idxs);
        else if (name == "emitter_prims")
            read_vec(fp, s, sc->emitter_prims);
        else if (name == "objects")
            read_vec(fp, s, sc->objects);
        else if (name == "bsdfs")
            read_vec(fp, s, sc->bsdfs);
        else if (name == "emitters")
            read_vec(fp, s, sc->emitters);
        else if (name == "media")
            read_vec(fp, s, sc->media);
        else if (name == "grids")
            read_vec(fp, s, sc->grids);
        else if (name == "grid_bricks")
            read_vec(fp, s, sc->grid_bricks);
        else if (name == "grid_data")
            read_vec(fp, s, sc->grid_data);
        else if (name == "blackbody")
            read_vec(fp, s, sc->blackbody);
        else if (name == "tex_desc")
            read_vec(fp, s, sc->tex_desc);
        else if (name == "tex_data")
            read_vec(fp, s, sc->tex_data);
        else if (name == "cie_xyz")
            read_vec(fp, s, sc->cie);
        else if (name == "d65_spd")
            read_vec(fp, s, sc->d65);
        else if (name.size() == 2 && name[0] == 'v' && name[1] >= '0' && name[1] <= '2')
            read_vec(fp, s, sc->v[name[1] - '0']);
        else if (name.size() == 2 && name[0] == 'n' && name[1] >= '0' && name[1] <= '2')
            read_vec(fp, s, sc->n[name[1] - '0']);
        else if (name.size() == 3 && name[0] == 'u' && name[2] >= '0' && name[2] <= '2')
            read_vec(fp, s, sc->uv[name[2] - '0']);
        else
            fseek(fp, long(bytes), SEEK_CUR);
        fseek(fp, long((16 - bytes % 16) % 16), SEEK_CUR);
    }
    fclose(fp);
    // precomputed triangles (aos.cuh:134-164)
    const size_t P = sc->v[0].size() / 3;
    sc->tris.resize(12 * P);
    for (size_t i = 0; i < P; i++) {
        const float *a = &sc->v[0][3 * i], *b = &sc->v[1][3 * i], *c = &sc->v[2][3 * i];
        const float d1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, d2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
        float *t = &sc->tris[12 * i];
        t[0] = a[0], t[1] = a[1], t[2] = a[2], t[3] = d1[1] * d2[2] - d1[2] * d2[1];
        t[4] = d1[0], t[5] = d1[1], t[6] = d1[2], t[7] = d2[0] * d1[2] - d1[0] * d2[2];
        t[8] = d2[0], t[9] = d2[1], t[10] = d2[2], t[11] = d1[0] * d2[1] - d2[0] * d1[1];
    }
    const size_t B = sc->bsdfs.size();
    sc->alpha_x.assign(B, 0), sc->alpha_y.assign(B, 0), sc->eta.assign(B, 1), sc->diff_f.assign(B, 0);
    for (size_t i = 0; i < B; i++) {
        const CptsBsdf &b = sc->bsdfs[i];
        if (b.type == CPTS_BSDF_GGX_CONDUCTOR) {
            sc->alpha_x[i] = roughness_to_alpha(b.extras[6]);
            sc->alpha_y[i] = roughness_to_alpha(b.extras[7]);
        } else if (b.type == CPTS_BSDF_PLASTIC || b.type == CPTS_BSDF_PLASTIC_FORWARD) {
            sc->eta[i] = 1.f / b.extras[0];
            if (b.type == CPTS_BSDF_PLASTIC)
                sc->diff_f[i] = diffuse_fresnel(b.extras[0]);
        }
    }
    for (const auto &e : sc->emitters)
        if (e.type == CPTS_EMIT_ENVMAP) { // rot = angleAxis(azimuth, z) * angleAxis(zenith, x)  (quaternion.cuh:45-60)
            const float sy = std::sin(e.azimuth * 0.5f), cy = std::cos(e.azimuth * 0.5f), sp = std::sin(e.zenith * 0.5f),
                        cp = std::cos(e.zenith * 0.5f);
            const float a[4] = {0, 0, sy, cy}, q[4] = {sp, 0, 0, cp};
            sc->env_rot[0] = a[3] * q[0] + a[0] * q[3] + a[1] * q[2] - a[2] * q[1];
            sc->env_rot[1] = a[3] * q[1] - a[0] * q[2] + a[1] * q[3] + a[2] * q[0];
            sc->env_rot[2] = a[3] * q[2] + a[0] * q[1] - a[1] * q[0] + a[2] * q[3];
            sc->env_rot[3] = a[3] * q[3] - a[0] * q[0] - a[1] * q[1] - a[2] * q[2];
        }
    return sc;
}

// ---- geometry -------------------------------------------------------------------------------------
inline bool slab(const float *lo, const float *hi, V3 inv_d, V3 o_div, float &t_near) { // aabb.cuh:58-65
    const float ax = lo[0] * inv_d.x - o_div.x, ay = lo[1] * inv_d.y - o_div.y, az = lo[2] * inv_d.z - o_div.z;
    const float bx = hi[0] * inv_d.x - o_div.x, by = hi[1] * inv_d.y - o_div.y, bz = hi[2] * inv_d.z - o_div.z;
    t_near = std::max(std::min(ax, bx), std::max(std::min(ay, by), std::min(az, bz)));
    const float t_far = std::min(std::max(ax, bx), std::min(std::max(ay, by), std::max(az, bz)));
    return t_far > t_near && t_far > 0;
}

int cached_entry(const Scene &sc, V3 inv_d, V3 o_div, float limit, int miss) { // tracing_func.cuh:53-65 / :128-140
    int node = miss;
    const int cn = int(sc.cache.size() / 4);
    for (int i = 0; i < cn;) {
        const uint32_t *c = &sc.cache[4 * i];
        const float lo[3] = {half_to_float(c[0] & 0xFFFF), half_to_float(c[1] & 0xFFFF), half_to_float(c[2] & 0xFFFF)};
        const float hi[3] = {half_to_float(c[0] >> 16), half_to_float(c[1] >> 16), half_to_float(c[2] >> 16)};
        float tn;
        bool hit = slab(lo, hi, inv_d, o_div, tn) && tn < limit;
        const int off = int(c[3] & 0xFF), gidx = int((c[3] >> 8) & 0xFFFFFF);
        const int inc = (!hit) * off + int(hit && off != 1);
        hit = hit && off == 1;
        i = hit ? cn : i + inc;
        node = hit ? gidx : node;
    }
    return node;
}

inline float hit_tri(const Scene &sc, int idx, V3 o, V3 d, float &u, float &v, float max_range) { // primitives.cuh:57-87
    const float *t = &sc.tris[12 * idx];
    const V3 p{o.x - t[0], o.y - t[1], o.z - t[2]};
    const V3 e1{t[4], t[5], t[6]}, e2{t[8], t[9], t[10]};
    const V3 t1{e2.y * -d.z + d.y * e2.z, -d.x * e2.z + e2.x * d.z, e2.x * -d.y + d.x * e2.y};
    const V3 t2{-d.y * e1.z + e1.y * d.z, e1.x * -d.z + d.x * e1.z, -d.x * e1.y + e1.x * d.y};
    const float det = e1.x * t1.x + e2.x * t2.x - t[3] * d.x, inv = 1.f / det;
    const float a = dot(t1, p) * inv, b = dot(t2, p) * inv, tt = (t[3] * p.x + t[7] * p.y + t[11] * p.z) * inv;
    u = a, v = b;
    return tt * float(a > 0 && b > 0 && a + b < 1 && tt > EPS && tt < max_range);
}

struct Hit {
    float t, u, v;
    int prim;
};
Hit closest_hit(const Scene &sc, V3 o, V3 d, float t_max = MAX_DIST) { // tracing_func.cuh:117-181
    Hit h{t_max, 0, 0, -1};
    const V3 inv_d{1.f / d.x, 1.f / d.y, 1.f / d.z}, o_div = o * inv_d;
    int node = cached_entry(sc, inv_d, o_div, t_max, 0);
    const int N = sc.num_nodes();
    while (node < N) {
        const float *nd = &sc.nodes[8 * node];
        int32_t beg, end;
        memcpy(&beg, nd + 3, 4);
        memcpy(&end, nd + 7, 4);
        float tn;
        const bool hit = slab(nd, nd + 4, inv_d, o_div, tn) && tn < h.t;
        node += (!hit) * (end < 0 ? -end : 1) + int(hit);
        end = hit ? end + beg : 0;
        for (int i = beg; i < end; i++) {
            float u, v;
            const float dist = hit_tri(sc, i, o, d, u, v, INFINITY);
            if (dist > EPS && dist < h.t)
                h = Hit{dist, u, v, i};
        }
    }
    return h;
}
bool unoccluded(const Scene &sc, V3 o, V3 d, float max_dist) { // tracing_func.cuh:44-99
    const V3 inv_d{1.f / d.x, 1.f / d.y, 1.f / d.z}, o_div = o * inv_d;
    const int N = sc.num_nodes();
    int node = cached_entry(sc, inv_d, o_div, max_dist, N);
    while (node < N) {
        const float *nd = &sc.nodes[8 * node];
        int32_t beg, end;
        memcpy(&beg, nd + 3, 4);
        memcpy(&end, nd + 7, 4);
        float tn;
        const bool hit = slab(nd, nd + 4, inv_d, o_div, tn) && tn < max_dist;
        node += (!hit) * (end < 0 ? -end : 1) + int(hit);
        end = hit ? end + beg : 0;
        for (int i = beg; i < end; i++) {
            float u, v;
            if (hit_tri(sc, i, o, d, u, v, max_dist) > EPS)
                return false;
        }
    }
    return true;
}

struct Isect {
    V3 n;
    float u, v; // half-rounded uv
};
Isect interaction(const Scene &sc, int prim, float u, float v) { // primitives.cuh:109-123
    const float w = 1.f - u - v;
    auto at = [&](int k) { return V3{sc.n[k][3 * prim], sc.n[k][3 * prim + 1], sc.n[k][3 * prim + 2]}; };
    Isect it;
    it.n = normalize(at(0) * w + at(1) * u + at(2) * v);
    const float hw = hround(w), hu = hround(u), hv = hround(v);
    for (int c = 0; c < 2; c++) {
        const float a = hround(sc.uv[0][2 * prim + c]), b = hround(sc.uv[1][2 * prim + c]), cc = hround(sc.uv[2][2 * prim + c]);
        const float r = hround(hround(hround(a * hw) + hround(b * hu)) + hround(cc * hv));
        (c == 0 ? it.u : it.v) = r;
    }
    return it;
}

// ---- image textures: what tex2D<floatN> returns for the reference's texture objects (textures.cu:134-165:
// normalised coordinates, wrap addressing, bilinear filter with 8-bit fractional weights) ---------------------
void tex2d(const Scene &sc, int index, float u, float v, float out[4]) {
    const CptsTexDesc &t = sc.tex_desc[index];
    const float *base = sc.tex_data.data() + t.offset;
    const int W = t.width, H = t.height, C = t.channels;
    const float x = (u - std::floor(u)) * W - 0.5f, y = (v - std::floor(v)) * H - 0.5f;
    const float fx = std::floor(x), fy = std::floor(y);
    const float a = std::floor((x - fx) * 256.f + 0.5f) / 256.f, b = std::floor((y - fy) * 256.f + 0.5f) / 256.f;
    const int x0 = ((int(fx) % W) + W) % W, x1 = (x0 + 1) % W, y0 = ((int(fy) % H) + H) % H, y1 = (y0 + 1) % H;
    for (int c = 0; c < 4; c++) {
        if (c >= C) {
            out[c] = 0.f;
            continue;
        }
        const float t00 = base[(size_t(y0) * W + x0) * C + c], t10 = base[(size_t(y0) * W + x1) * C + c],
                    t01 = base[(size_t(y1) * W + x0) * C + c], t11 = base[(size_t(y1) * W + x1) * C + c];
        out[c] = (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
    }
}
// PackedHalf2::lerp of a primitive's vertex uvs (vec2_half.cuh:182-186): every product and sum rounded to half
void lerp_uv(const Scene &sc, int prim, float u, float v, float &tu, float &tv) {
    const float hw = hround(1.f - u - v), hu = hround(u), hv = hround(v);
    for (int c = 0; c < 2; c++) {
        const float a = hround(sc.uv[0][2 * prim + c]), b = hround(sc.uv[1][2 * prim + c]), cc = hround(sc.uv[2][2 * prim + c]);
        (c == 0 ? tu : tv) = hround(hround(hround(a * hw) + hround(b * hu)) + hround(cc * hv));
    }
}

// ---- frames (so3.cuh:201-252) -----------------------------------------------------------------------
struct M3 {
    V3 r[3];
};
M3 frame(V3 to, bool l2w) {
    V3 axis{l2w ? -to.y : to.y, l2w ? to.x : -to.x, 0.f};
    M3 R{{{to.z, 0, 0}, {0, to.z, 0}, {0, 0, to.z}}};
    if (std::fabs(to.z) < 1.f - 1e-5f) {
        const V3 sk[3] = {{0, -axis.z, axis.y}, {axis.z, 0, -axis.x}, {-axis.y, axis.x, 0}};
        axis = normalize(axis);
        const float s = 1.f - to.z, a[3] = {axis.x, axis.y, axis.z};
        for (int i = 0; i < 3; i++)
            R.r[i] = R.r[i] + ((axis * a[i]) * s + sk[i]);
    }
    return R;
}
inline V3 rot(const M3 &m, V3 v) { return {dot(m.r[0], v), dot(m.r[1], v), dot(m.r[2], v)}; }
inline V3 rot_t(const M3 &m, V3 v) {
    return {m.r[0].x * v.x + m.r[1].x * v.y + m.r[2].x * v.z, m.r[0].y * v.x + m.r[1].y * v.y + m.r[2].y * v.z,
            m.r[0].z * v.x + m.r[1].z * v.y + m.r[2].z * v.z};
}
inline V3 to_world(V3 n, V3 local) { return rot(frame(n, true), local); }
// eval_normal / eval_normal_reused (textures.cuh:79-103): R_w2l^T applied to the (normalised) normal-map texel,
// (0,0,1) without a normal map
inline V3 shading_normal(V3 n, M3 *R = nullptr, V3 pn = V3{0, 0, 1}) {
    const M3 m = frame(n, false);
    if (R)
        *R = m;
    return rot_t(m, pn);
}

// ---- sampling warps (sampling.cuh:27-67) --------------------------------------------------------------
V3 cos_hemisphere(float ux, float uy, float &pdf) {
    const float ct = std::sqrt(ux), st = std::sqrt(1.f - ux);
    pdf = ct * INV_PI;
    return {std::cos(2.f * uy * PI) * st, std::sin(2.f * uy * PI) * st, ct};
}

// ---- Fresnel (fresnel.cuh:28-140) -----------------------------------------------------------------------
float fresnel_simple(float eta, float ci) {
    eta = ci < 0.f ? 1.f / eta : eta;
    ci = std::fabs(ci);
    const float st2 = eta * eta * (1.f - ci * ci);
    if (st2 >= 1.f)
        return 1.f;
    const float ct = std::sqrt(std::max(1.f - st2, 0.f));
    const float rs = (eta * ci - ct) / (eta * ci + ct), rp = (eta * ct - ci) / (eta * ct + ci);
    return (rs * rs + rp * rp) * 0.5f;
}
inline float sgn(float v) { return float(v > 0.f) - float(v < 0.f); }
V3 snell(V3 in, V3 n, float &cos_r2, float dn, float ni, float nr) {
    const float ratio = ni / nr;
    cos_r2 = float(1.f - (ratio * ratio) * (1. - dn * dn));
    return normalize(ratio * in - (ratio * dn) * n + (sgn(dn) * std::sqrt(std::fabs(cos_r2))) * n) * float(cos_r2 > 0.f);
}
float fresnel_dielectric(float n_in, float n_out, float ci, float cr) {
    const float a = n_in * ci, b = n_out * ci, c = n_in * cr, d = n_out * cr;
    const float rs = (a - d) / (a + d), rp = (c - b) / (c + b);
    return 0.5f * (rs * rs + rp * rp);
}
V4 fresnel_conductor(const CptsBsdf &m, float ci) {
    ci = std::min(std::max(ci, -1.f), 1.f);
    const float c2 = ci * ci, s2 = 1.f - c2;
    float out[3];
    for (int k = 0; k < 3; k++) {
        const float e2 = m.extras[k] * m.extras[k], k2 = m.extras[3 + k] * m.extras[3 + k];
        const float t0 = e2 - k2 - s2, a2b2 = std::sqrt(t0 * t0 + 4 * e2 * k2), t1 = a2b2 + c2;
        const float a = std::sqrt(0.5f * (a2b2 + t0)), t2 = 2.f * ci * a, rs = (t1 - t2) / (t1 + t2);
        const float t3 = c2 * a2b2 + s2 * s2, t4 = t2 * s2, rp = rs * (t3 - t4) / (t3 + t4);
        out[k] = 0.5f * (rp + rs);
    }
    return {out[0], out[1], out[2], 1.f};
}

// ---- GGX (bsdf_ggx.cu:25-189) ------------------------------------------------------------------------------
float ggx_lambda(V3 l, float ax, float ay) {
    const float c2 = l.z * l.z, ic2 = c2 == 0 ? 0 : 1.f / c2;
    const float e = (l.x * l.x * ax * ax + l.y * l.y * ay * ay) * ic2;
    return e == 0 ? 0 : (-1.f + std::sqrt(1.f + e)) * 0.5f;
}
float ggx_D(V3 l, float ax, float ay) {
    const float c2 = l.z * l.z, ic2 = c2 == 0 ? 0 : 1.f / c2;
    const float e = (l.x * l.x / (ax * ax) + l.y * l.y / (ay * ay)) * ic2;
    return 1.f / (PI * ax * ay * c2 * c2 * (1 + e) * (1 + e));
}
float ggx_G1(V3 l, float ax, float ay) { return 1.f / (1.f + ggx_lambda(l, ax, ay)); }
float ggx_G(V3 a, V3 b, float ax, float ay) { return 1.f / (1.f + ggx_lambda(a, ax, ay) + ggx_lambda(b, ax, ay)); }
void ggx_slopes(float ct, float ux, float uy, float &sx, float &sy) {
    if (ct == 1) {
        const float r = std::sqrt(ux / (1 - uy));
        sx = r * std::cos(2.f * uy * PI), sy = r * std::sin(2.f * uy * PI);
        return;
    }
    const float st = std::sqrt(std::max(0.f, 1.f - ct * ct)), tt = st / ct, a = 1.f / tt;
    const float G1 = 2.f / (1 + std::sqrt(1.f + 1.f / (a * a)));
    const float A = 2.f * ux / G1 - 1, tmp = std::min(1.f / (A * A - 1.f), 1e9f), B = tt;
    const float D = std::sqrt(std::max(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.f));
    const float s1 = B * tmp - D, s2 = B * tmp + D;
    sx = (A < 0 || s2 > 1.f / tt) ? s1 : s2;
    const float sign = uy > 0.5f ? 1.f : -1.f;
    uy = 2.f * (uy - .5f) * sign;
    const float z = (uy * (uy * (uy * 0.27385f - 0.73369f) + 0.46341f)) /
                    (uy * (uy * (uy * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
    sy = sign * z * std::sqrt(1.f + sx * sx);
}
V3 ggx_sample_wh(V3 li, float ax, float ay, float ux, float uy) {
    const V3 ws = normalize(V3{li.x * ax, li.y * ay, li.z});
    float sx, sy;
    ggx_slopes(ws.z, ux, uy, sx, sy);
    const float s2 = std::max(1.f - ws.z * ws.z, 0.f), inv = 1.f / std::sqrt(s2);
    const float cphi = s2 == 0 ? 1 : std::min(std::max(ws.x * inv, -1.f), 1.f),
                sphi = s2 == 0 ? 0 : std::min(std::max(ws.y * inv, -1.f), 1.f);
    const float tmp = cphi * sx - sphi * sy;
    sy = (sphi * sx + cphi * sy) * ay;
    sx = tmp * ax;
    return normalize(V3{-sx, -sy, 1.f});
}

// ---- dispersion colour (bsdf_dispersion.cu:26-44, xyz.cuh:47-57) ------------------------------------------
float table_lookup(const std::vector<float> &tab, int stride, int comp, float x01) { // linear filter, clamp
    const int n = int(tab.size()) / stride;
    float p = x01 * n - 0.5f;
    p = std::min(std::max(p, 0.f), float(n - 1));
    const int i0 = int(std::floor(p)), i1 = std::min(i0 + 1, n - 1);
    const float f = p - i0;
    return tab[i0 * stride + comp] * (1 - f) + tab[i1 * stride + comp] * f;
}
V4 wavelength_rgb(const Scene &sc, float wl) {
    const float ci = (wl - 360.f) / 471.f, di = (wl - 300.f) / 531.f;
    const float spd = table_lookup(sc.d65, 1, 0, di) / 22.2175f;
    const float X = table_lookup(sc.cie, 4, 0, ci) * spd, Y = table_lookup(sc.cie, 4, 1, ci) * spd, Z = table_lookup(sc.cie, 4, 2, ci) * spd;
    const float r = 3.240479f * X - 1.537150f * Y - 0.498535f * Z, g = -0.969256f * X + 1.875991f * Y + 0.041556f * Z,
                b = 0.055648f * X - 0.204043f * Y + 1.057311f * Z;
    // Vec4(r,g,b) has w = 1; the scale constructor too
    return {std::max(r, 0.f) / 1.4979f, std::max(g, 0.f) / 1.13591f, std::max(b, 0.f) / 1.13159f, 1.f};
}

// ---- dielectric core (translucent.cuh:43-98); is_radiance is false on this path (base-class default) ------
V4 dielectric_eval(V3 n, V3 out, V3 in, V4 ks, float eta) {
    const float dn = dot(in, n), ni = dn < 0 ? 1.f : eta, nr = dn < 0 ? eta : 1.f;
    float c2 = 0;
    const V3 ret = normalize(in + n * (-2.f * dn)), refr = snell(in, n, c2, dn, ni, nr);
    return ks * float((dot(out, ret) > 0.99999f) | (dot(out, refr) > 0.99999f));
}
V3 dielectric_sample(V3 in, V3 n, V4 ks, float eta, V4 &thp, Rng &rng, float &pdf, int &lobe) {
    const float dn = dot(in, n), ni = dn < 0 ? 1.f : eta;
    float nr = dn < 0 ? eta : 1.f, c2 = 0;
    const V3 ret = normalize(in + n * (-2.f * dn)), refr = snell(in, n, c2, dn, ni, nr);
    const bool total = (1.f - (ni * ni) / (nr * nr) * (1.f - dn * dn)) < 0.f;
    nr = fresnel_dielectric(ni, nr, std::fabs(dn), std::sqrt(std::fabs(c2)));
    const bool reflect = total || rng.f() < nr;
    pdf = total ? 1.f : (reflect ? nr : 1.f - nr);
    lobe = LOBE_SPECULAR | (total || reflect ? LOBE_REFLECT : LOBE_TRANSMIT);
    thp = thp * ks;
    return reflect ? ret : refr;
}
bool dispersion_wl(const CptsBsdf &m, V3 in, V3 out, V3 n, float &wl) { // dispersion.cuh:70-80
    const float ci = dot(n, in), co = dot(n, out), si = std::sqrt(1.f - ci * ci), so = std::sqrt(1.f - co * co);
    const float eta = si > so ? si / so : so / si;
    wl = std::sqrt(m.extras[1] / std::max(eta - m.extras[0], 1e-5f));
    return wl > 360.f && wl < 831.f;
}

// ---- BSDF interface ---------------------------------------------------------------------------------------
struct Mat {
    const CptsBsdf *b; // colours already replaced by this hit's texture lookups
    float ax, ay, eta, diff_f;
    V3 pn{0, 0, 1};    // tangent-space normal of the normal map at this hit
};
// per-hit material: textures.cuh:79-117 lookups folded into a local copy of the BSDF record
Mat resolve_material(const Scene &sc, int mat_id, float u, float v, CptsBsdf &local) {
    local = sc.bsdfs[mat_id];
    Mat m{&local, sc.alpha_x[mat_id], sc.alpha_y[mat_id], sc.eta[mat_id], sc.diff_f[mat_id]};
    float t[4];
    float *dst[3] = {local.k_d, local.k_s, local.k_g};
    for (int k = 0; k < 3; k++)
        if (local.tex[k] >= 0 && !(k == 0 && (local.type == CPTS_BSDF_TRANSLUCENT || local.type == CPTS_BSDF_DISPERSION))) {
            tex2d(sc, local.tex[k], u, v, t);
            memcpy(dst[k], t, sizeof(t));
        }
    if (local.tex[3] >= 0) {
        tex2d(sc, local.tex[3], u, v, t);
        m.pn = normalize(V3{t[0], t[1], t[2]});
    }
    if (local.tex[4] >= 0) { // eval_rough: GGX alpha pair, or the ior of a det-refraction BSDF (translucent.cuh:106)
        tex2d(sc, local.tex[4], u, v, t);
        if (local.type == CPTS_BSDF_GGX_CONDUCTOR)
            m.ax = t[0], m.ay = t[1];
        else if (local.type == CPTS_BSDF_TRANSLUCENT)
            local.k_d[0] = t[0];
    }
    return m;
}
float bsdf_pdf(const Scene &sc, const Mat &m, V3 nrm, V3 out, V3 in) {
    const CptsBsdf &b = *m.b;
    switch (b.type) {
    case CPTS_BSDF_LAMBERTIAN: return std::max(dot(shading_normal(nrm, nullptr, m.pn), out), 0.f) * INV_PI;
    case CPTS_BSDF_PLASTIC: {
        const V3 n = shading_normal(nrm, nullptr, m.pn);
        const float dwo = std::fabs(dot(out, n)), dwi = std::fabs(dot(in, n)), Fi = fresnel_simple(m.eta, dwi),
                    sp = Fi / (Fi + b.extras[1] * (1.f - Fi));
        const V3 rd = in - n * (2.f * dot(in, n));
        return dot(rd, out) < 1.f - THP_EPS ? INV_PI * dwo * (1.f - sp) : sp;
    }
    case CPTS_BSDF_PLASTIC_FORWARD: {
        const V3 n = shading_normal(nrm, nullptr, m.pn);
        const float dwi = dot(in, n), Fi = fresnel_simple(m.eta, -dwi), sp = Fi / (Fi + b.extras[1] * (1.f - Fi));
        const V3 rd = in - n * (2.f * dwi);
        float p = 0;
        p = dot(rd, out) < 1.f - THP_EPS ? p : sp;
        p = dot(in, out) < 1.f - THP_EPS ? p : 1.f - sp;
        return p;
    }
    case CPTS_BSDF_GGX_CONDUCTOR: {
        M3 R;
        shading_normal(nrm, &R, m.pn);
        const V3 li = -rot(R, in), lo = rot(R, out), wh = normalize(lo + li);
        const float pv = ggx_D(wh, m.ax, m.ay) * ggx_G1(li, m.ax, m.ay) * (std::fabs(dot(li, wh)) / std::fabs(li.z));
        return ((li.z > 0) ^ (lo.z > 0)) ? 0 : pv * (1.f / (4.f * dot(wh, li)));
    }
    case CPTS_BSDF_DISPERSION: {
        const V3 n = shading_normal(nrm, nullptr, m.pn);
        float p = 0;
        if (((dot(n, in) > 0) ^ (dot(n, out) > 0)) == false) {
            float wl = 0;
            p = dispersion_wl(b, in, out, n, wl);
            const float eta = b.extras[0] + b.extras[1] / (wl * wl), F = fresnel_simple(eta, -dot(in, n));
            p *= (1.f - F) / 471.f;
        }
        return p;
    }
    case CPTS_BSDF_FORWARD: return float(dot(in, out) == 1.f);
    default: return 0.f;
    }
}
V4 bsdf_eval(const Scene &sc, const Mat &m, V3 nrm, V3 out, V3 in) {
    const CptsBsdf &b = *m.b;
    const V4 kd = v4(b.k_d), ks = v4(b.k_s), kg = v4(b.k_g);
    switch (b.type) {
    case CPTS_BSDF_LAMBERTIAN: {
        const V3 n = shading_normal(nrm, nullptr, m.pn);
        const float ct = dot(n, out), di = dot(n, in);
        return kd * std::max(0.f, ct) * INV_PI * float((di > 0) ^ (ct > 0));
    }
    case CPTS_BSDF_SPECULAR: {
        const V3 n = shading_normal(nrm, nullptr, m.pn), rd = normalize(in + n * (-2.f * dot(in, n)));
        return ks * float(dot(out, rd) > 0.99999f);
    }
    case CPTS_BSDF_TRANSLUCENT: return dielectric_eval(shading_normal(nrm, nullptr, m.pn), out, in, ks, b.k_d[0]);
    case CPTS_BSDF_PLASTIC: {
        const V3 n = shading_normal(nrm, nullptr, m.pn);
        const float rwo = dot(out, n), rwi = dot(in, n), dwo = std::fabs(rwo), dwi = std::fabs(rwi),
                    Fi = fresnel_simple(m.eta, dwi), Fo = fresnel_simple(m.eta, dwo);
        const V3 rd = in - n * (2.f * rwi);
        V4 brdf = (kd / ((kd * -m.diff_f) + 1.f)) * ((INV_PI * (1.f - Fi) * (1.f - Fo) * m.eta * m.eta) * dwo) *
                  expxyz(kg * b.extras[2] * (-1.f / dwo - 1.f / dwi));
        if (!(dot(rd, out) < 1.f - THP_EPS))
            brdf = brdf + v4(Fi, 1) * ks;
        return (b.extras[3] > 0 || ((rwo > 0) ^ (rwi > 0))) ? brdf : v4(0, 1);
    }
    case CPTS_BSDF_PLASTIC_FORWARD: {
        const V3 n = shading_normal(nrm, nullptr, m.pn);
        const float dwi = dot(in, n), Fi = fresnel_simple(m.eta, std::fabs(dwi));
        const V3 rd = in - n * (2.f * dwi);
        V4 brdf = dot(in, out) < 1.f - THP_EPS
                      ? v4(0, 1)
                      : kd * ((1.f - Fi) * (1.f - Fi)) * m.eta * m.eta * expxyz(kg * b.extras[2] * (-2.f / std::fabs(dwi)));
        brdf = brdf + (dot(rd, out) < 1.f - THP_EPS ? brdf : v4(Fi, 1) * ks);
        return brdf;
    }
    case CPTS_BSDF_GGX_CONDUCTOR: {
        M3 R;
        const V3 n = shading_normal(nrm, &R, m.pn), li = -rot(R, in), lo = rot(R, out);
        V3 wh = normalize(lo + li);
        if (!(wh.z > 0))
            wh = -wh;
        const V4 F = fresnel_conductor(b, std::fabs(dot(lo, wh)));
        if ((li.z > 0) ^ (lo.z > 0))
            return v4(0.f) * kg;
        const V4 g = F * ggx_D(wh, m.ax, m.ay) * (1.f / (4.f * li.z * lo.z)) * ggx_G(li, lo, m.ax, m.ay);
        return kg * g * std::max(0.f, dot(out, n));
    }
    case CPTS_BSDF_DISPERSION: {
        const V3 n = shading_normal(nrm, nullptr, m.pn);
        float wl = 0;
        const bool valid = dispersion_wl(b, in, out, n, wl);
        const float eta = valid ? b.extras[0] + b.extras[1] / (wl * wl) : b.extras[0];
        V4 r = dielectric_eval(n, out, in, ks, eta);
        return r * (valid ? wavelength_rgb(sc, wl) : v4(1.f));
    }
    case CPTS_BSDF_FORWARD: return v4(float(dot(in, out) == 1.f), 1);
    default: return v4(0.f);
    }
}
V3 bsdf_sample(const Scene &sc, const Mat &m, V3 in, V3 nrm, V4 &thp, float &pdf, Rng &rng, int &lobe) {
    const CptsBsdf &b = *m.b;
    const V4 kd = v4(b.k_d), ks = v4(b.k_s), kg = v4(b.k_g);
    switch (b.type) {
    case CPTS_BSDF_LAMBERTIAN: {
        const float ux = rng.f(), uy = rng.f();
        const V3 local = cos_hemisphere(ux, uy, pdf), n = shading_normal(nrm, nullptr, m.pn), out = to_world(n, local);
        thp = thp * kd * float((dot(n, in) > 0) ^ (dot(n, out) > 0));
        lobe = b.flags;
        return out;
    }
    case CPTS_BSDF_SPECULAR: {
        const V3 n = shading_normal(nrm, nullptr, m.pn);
        pdf = 1.f;
        lobe = b.flags;
        thp = thp * ks;
        return in - n * (2.f * dot(in, n));
    }
    case CPTS_BSDF_TRANSLUCENT: return dielectric_sample(in, shading_normal(nrm, nullptr, m.pn), ks, b.k_d[0], thp, rng, pdf, lobe);
    case CPTS_BSDF_PLASTIC: {
        const V3 n = shading_normal(nrm, nullptr, m.pn);
        const float rdi = dot(n, in), di = std::fabs(rdi), Fi = fresnel_simple(m.eta, di), sp = Fi / (Fi + b.extras[1] * (1.f - Fi));
        V3 out;
        if (rng.f() < sp) {
            out = in - n * (2.f * dot(in, n));
            pdf = sp;
            thp = thp * (v4(Fi / sp, 1) * ks);
            lobe = LOBE_REFLECT | LOBE_SPECULAR;
        } else {
            float dummy;
            const float ux = rng.f(), uy = rng.f();
            const V3 local = cos_hemisphere(ux, uy, dummy);
            const float Fo = fresnel_simple(m.eta, local.z);
            const V4 lthp = (kd / ((kd * -m.diff_f) + 1.f)) * ((1.f - Fi) * (1.f - Fo) * m.eta * m.eta) *
                            expxyz(kg * b.extras[2] * (-1.f / local.z - 1.f / di));
            pdf = INV_PI * local.z * (1.f - sp);
            thp = thp * (lthp * (1.f / (1.f - sp)));
            out = to_world(n, local);
            lobe = LOBE_REFLECT | LOBE_DIFFUSE;
        }
        if (!(b.extras[3] > 0 || ((rdi > 0) ^ (dot(out, n) > 0))))
            thp = v4(0, 1);
        return out;
    }
    case CPTS_BSDF_PLASTIC_FORWARD: {
        const V3 n = shading_normal(nrm, nullptr, m.pn);
        const float di = dot(n, in), Fi = fresnel_simple(m.eta, std::fabs(di)), sp = Fi / (Fi + b.extras[1] * (1.f - Fi));
        if (rng.f() < sp) {
            pdf = sp;
            thp = thp * (v4(Fi / sp, 1) * ks);
            lobe = LOBE_REFLECT | LOBE_SPECULAR;
            return in - n * (2.f * di);
        }
        const V4 lthp = kd * ((1.f - Fi) * (1.f - Fi) * m.eta * m.eta) * expxyz(kg * b.extras[2] * (-2.f / std::fabs(di)));
        pdf = 1.f - sp;
        thp = thp * (lthp * (1.f / (1.f - sp)));
        lobe = LOBE_TRANSMIT | LOBE_SPECULAR;
        return in;
    }
    case CPTS_BSDF_GGX_CONDUCTOR: {
        M3 R;
        const V3 n = shading_normal(nrm, &R, m.pn), li = -rot(R, in);
        const float ux = rng.f(), uy = rng.f();
        const V3 wh = ggx_sample_wh(li, m.ax, m.ay, ux, uy);
        const float D = ggx_D(wh, m.ax, m.ay), dim = dot(li, wh);
        pdf = D * ggx_G1(li, m.ax, m.ay) * std::fabs(dim / li.z);
        pdf = (pdf > 0 && dim > 0) ? pdf * (1.f / (4.f * dim)) : 0;
        const V3 lref = wh * (2.f * dim) - li, ref = rot_t(R, lref);
        const float ci = li.z, co = lref.z;
        const V4 F = fresnel_conductor(b, std::fabs(dot(lref, wh)));
        if (ci > 0 && co > 0 && pdf > 0)
            thp = thp * (kg * (1.f / pdf) * D * ggx_G(li, lref, m.ax, m.ay) * (1.f / (4.f * ci * co)) * F * std::max(dot(n, ref), 0.f));
        lobe = b.flags;
        return ref;
    }
    case CPTS_BSDF_DISPERSION: {
        const float wl = rng.f() * 471.f + 360.f, eta = b.extras[0] + b.extras[1] / (wl * wl);
        const V3 r = dielectric_sample(in, shading_normal(nrm, nullptr, m.pn), ks, eta, thp, rng, pdf, lobe);
        thp = thp * wavelength_rgb(sc, wl);
        pdf *= 1.f / 471.f;
        return r;
    }
    default:
        pdf = 1.f;
        lobe = b.flags;
        return in;
    }
}

// ---- emitters (emitter.cuh) ---------------------------------------------------------------------------------
V4 envmap_lookup(const Scene &sc, const CptsEmitter &e, V3 dir) { // emitter.cu:36-46,67-73
    const float *q = sc.env_rot;
    auto mul = [](const float a[4], const float b[4], float o[4]) {
        o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
        o[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
        o[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
        o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    };
    const float vq[4] = {dir.x, dir.y, dir.z, 0.f}, conj[4] = {-q[0], -q[1], -q[2], q[3]};
    float qv[4], r[4];
    mul(q, vq, qv);
    mul(qv, conj, r);
    const float ty = std::acos(std::min(std::max(r[2], -1.f), 1.f)) * INV_PI, px = (std::atan2(r[1], r[0]) + PI) * INV_PI * 0.5f;
    float t[4];
    tex2d(sc, e.tex, px, ty, t);
    return V4{t[0], t[1], t[2], t[3]} * e.scale;
}
V4 emitter_le(const Scene &sc, const CptsEmitter &e, float tu, float tv) { // AreaSource::scaled_Le (emitter.cuh:154-156)
    if (e.tex < 0)
        return v4(e.Le);
    float t[4];
    tex2d(sc, e.tex, tu, tv, t);
    return V4{t[0], t[1], t[2], t[3]} * e.Le[3];
}
V4 emitter_eval(const Scene &sc, const CptsEmitter &e, V3 inci, V3 n, float tu, float tv) {
    if (e.type == CPTS_EMIT_AREA)
        return dot(inci, n) < 0 ? emitter_le(sc, e, tu, tv) : V4{0, 0, 0, 1};
    if (e.type == CPTS_EMIT_AREA_SPOT)
        return dot(inci, n) < -e.cos_val ? emitter_le(sc, e, tu, tv) : V4{0, 0, 0, 1};
    if (e.type == CPTS_EMIT_ENVMAP)
        return envmap_lookup(sc, e, inci);
    return {0, 0, 0, 1};
}
V4 null_or_env(const Scene &sc, V3 dir) { // c_emitter[envmap_id]->eval_le(&ray.d)
    const int id = sc.cfg[CPTS_CFG_ENVMAP_ID];
    return id > 0 ? envmap_lookup(sc, sc.emitters[id], dir) : V4{0, 0, 0, 1};
}
V3 emitter_sample(const Scene &sc, const CptsEmitter &e, V3 hit, V3 hit_n, V4 &le, float &pdf, float ux, float uy, int prim) {
    if (e.type == CPTS_EMIT_POINT) {
        const V3 p{e.pos[0], e.pos[1], e.pos[2]};
        le = v4(e.Le) * std::min(1.f / std::max(len2(p - hit), 1e-5f), 1.f);
        return p;
    }
    if (e.type == CPTS_EMIT_AREA || e.type == CPTS_EMIT_AREA_SPOT) {
        if (!(ux + uy < 1.f))
            ux = 1.f - ux, uy = 1.f - uy;
        const float *t = &sc.tris[12 * prim];
        const V3 p = V3{t[4], t[5], t[6]} * ux + V3{t[8], t[9], t[10]} * uy + V3{t[0], t[1], t[2]};
        auto at = [&](int k) { return V3{sc.n[k][3 * prim], sc.n[k][3 * prim + 1], sc.n[k][3 * prim + 2]}; };
        const V3 n = normalize(at(0) * (1.f - ux - uy) + at(1) * ux + at(2) * uy);
        V3 th = hit - p;
        pdf *= len2(th);
        th = normalize(th);
        const float c = dot(n, th);
        pdf *= float(c > 0) / c;
        float tu = 0, tv = 0;
        if (e.tex >= 0)
            lerp_uv(sc, prim, ux, uy, tu, tv);
        le = emitter_le(sc, e, tu, tv) * float(e.type == CPTS_EMIT_AREA ? c > 0 : c > e.cos_val);
        return p;
    }
    if (e.type == CPTS_EMIT_ENVMAP) { // emitter.cu:26-46; surface hits always have a unit normal
        const V3 dir = to_world(hit_n, cos_hemisphere(ux, uy, pdf));
        le = envmap_lookup(sc, e, dir);
        return hit + dir * 5e3f;
    }
    pdf = 1;
    le = v4(0.f);
    return {0, 0, 0};
}

void camera_ray(const Scene &sc, int px, int py, Rng &rng, V3 &o, V3 &d) {
    // camera (camera_model.cuh:58-87)
    const float *c = sc.cam;
    const float xp = rng.f() + float(px), yp = rng.f() + float(py);
    const float nx = (xp - c[CPTS_CAM_HW]) * c[CPTS_CAM_INV_FOCAL] * c[CPTS_CAM_SIGN_X], ny = (c[CPTS_CAM_HH] - yp) * c[CPTS_CAM_INV_FOCAL];
    const M3 R{{{c[0], c[1], c[2]}, {c[3], c[4], c[5]}, {c[6], c[7], c[8]}}};
    const V3 T{c[9], c[10], c[11]};
    d = normalize(rot(R, V3{nx, ny, 1.f}));
    if (c[CPTS_CAM_FOCAL_DIST] == 0) {
        o = T + V3{c[1], c[4], c[7]} * ny + V3{c[0], c[3], c[6]} * nx;
        d = normalize(rot(R, V3{0, 0, 1}));
    } else if (c[CPTS_CAM_APERTURE] > 0) {
        const V3 focus = T + d * c[CPTS_CAM_FOCAL_DIST];
        const float ux = rng.f(), uy = rng.f(), r = std::sqrt(ux);
        const float lx = r * std::cos(2.f * uy * PI) * c[CPTS_CAM_APERTURE], ly = r * std::sin(2.f * uy * PI) * c[CPTS_CAM_APERTURE];
        o = T + (V3{c[0], c[3], c[6]} * lx + V3{c[1], c[4], c[7]} * ly);
        d = normalize(focus - o);
    } else
        o = T;

}

// ---- one path (wavefront_pt.cu:71-314; loop structure of wf_path_tracer.cu:175-230) ----------------------------
struct Counts {
    uint64_t closest = 0, shadow = 0, shaded = 0;
};
V4 trace_path(const Scene &sc, int px, int py, int accum, int seed_offset, int max_depth, Counts &cnt) {
    const int W = sc.cfg[CPTS_CFG_WIDTH];
    Rng rng;
    rng.seed(px + py * W, int(uint32_t(seed_offset) + uint32_t(accum) * uint32_t(SEED_SCALER)));
    V3 o, d;
    camera_ray(sc, px, py, rng, o, d);

    V4 thp{1, 1, 1, 1}, L{0, 0, 0, 1};
    float pdf = 1.f;
    bool delta = false;
    Hit h = closest_hit(sc, o, d);
    if (h.prim < 0)
        return null_or_env(sc, d); // null emitter: (0,0,0,1)
    Isect it = interaction(sc, h.prim, h.u, h.v);
    const int num_emitters = int(sc.emitters.size()) - 1;
    for (int bounce = 0; bounce < max_depth; bounce++) {
        // ---- shade (fused_ray_bounce_shader)
        cnt.shaded++;
        const int obj_id = sc.obj_idxs[h.prim] & 0x000fffff; // extract_object_info (tracing_func.cuh:38-41)
        const CptsObject &ob = sc.objects[obj_id];
        const int mat_id = ob.bsdf_emitter >> 16, emitter_id = ((ob.bsdf_emitter & 0xFFFF) ^ 0x8000) - 0x8000;
        CptsBsdf local_bsdf;
        const Mat m = resolve_material(sc, mat_id, it.u, it.v, local_bsdf);
        int num = num_emitters;
        num -= emitter_id > 0 && num > 1;
        uint32_t eid = rng.u32() % uint32_t(num) + 1;
        eid += eid >= uint32_t(emitter_id) && emitter_id > 0;
        float direct_pdf = 1.f / float(num);
        eid *= uint32_t(emitter_id == 0 || num > 1);
        const CptsEmitter &em = sc.emitters[eid];
        const CptsObject &eo = sc.objects[em.obj_ref * (em.obj_ref >= 0)];
        direct_pdf *= eo.inv_area;
        const uint32_t pick = rng.u32() % uint32_t(eo.prim_num) + uint32_t(eo.prim_offset);
        const int eprim = pick < sc.emitter_prims.size() ? sc.emitter_prims[pick] : 0;
        const V3 hit_pos = o + d * h.t;
        o = hit_pos;
        V4 direct{0, 0, 0, 1};
        const float ux = rng.f(), uy = rng.f();
        V3 sd = emitter_sample(sc, em, hit_pos, it.n, direct, direct_pdf, ux, uy, eprim) - hit_pos;
        const float dist = std::sqrt(len2(sd));
        sd = sd * (1.f / dist);
        V4 rdc = L;
        if (eid != 0) {
            cnt.shadow++;
            if (unoccluded(sc, hit_pos, sd, dist - EPS)) {
                const float mis = direct_pdf + bsdf_pdf(sc, m, it.n, sd, d) * float(em.is_sphere != 0);
                rdc = rdc + thp * direct * bsdf_eval(sc, m, it.n, sd, d) * (mis > EPS ? 1.f / mis : 0.f);
            }
        }
        const float sa = ob.inv_area * h.t * h.t / std::max(std::fabs(dot(d, shading_normal(it.n, nullptr, m.pn))), 1e-4f);
        const float w = pdf / (pdf + sa * float(emitter_id > 0) * float(bounce > 0) * float(!delta));
        L = (thp * emitter_eval(sc, sc.emitters[emitter_id], d, it.n, it.u, it.v)) * w + rdc;
        int lobe = 0;
        d = bsdf_sample(sc, m, d, it.n, thp, pdf, rng, lobe);
        delta = (lobe & LOBE_SPECULAR) != 0;
        if (bounce + 1 >= max_depth)
            break;
        // ---- trace + Russian roulette (fused_closesthit_shader)
        cnt.closest++;
        h = closest_hit(sc, o, d);
        if (h.prim < 0) {
            L = L + thp * null_or_env(sc, d);
            break;
        }
        it = interaction(sc, h.prim, h.u, h.v);
        const float mx = std::max(thp.x, std::max(thp.y, thp.z));
        if (bounce >= RR_BOUNCE && mx < RR_THRESHOLD) {
            const bool survive = rng.f() < mx && mx > THP_EPS;
            thp = thp * (survive ? 1.f / mx : 0.f);
            if (!survive)
                break;
        }
    }
    return L;
}

// ---- participating media (src/volume/homogeneous.cuh:45-88, src/core/medium.cuh:72-112, src/volume/henyey_greenstein.cuh,
// nested-volume stack of megakernel_vpt.cu:59-94) ------------------------------------------------------------------
struct MediumSampleO {
    V4 local_thp;
    float dist;
    int flag;
};
MediumSampleO medium_sample(const CptsMedium *m, float hit_t, Rng &rng) { // m == nullptr: vacuum
    if (!m)
        return {v4(1.f), hit_t, 0};
    const float st[3] = {m->sigma_a[0] + m->sigma_s[0], m->sigma_a[1] + m->sigma_s[1], m->sigma_a[2] + m->sigma_s[2]};
    // signed `discrete1D() % 3`: negative draws index the members in front of sigma_t
    const int channel = int(rng.u32()) % 3;
    const float st_c = channel >= 0 ? st[channel] : (channel == -1 ? m->sigma_s[3] : m->sigma_s[2]);
    float dist = -std::log(1.f - rng.f()) / st_c;
    const bool is_medium = dist < hit_t;
    dist = std::min(dist, hit_t);
    const float eyx = std::exp(-(st[1] - st[0]) * dist), ezx = std::exp(-(st[2] - st[0]) * dist), ezy = std::exp(-(st[2] - st[1]) * dist);
    const float exy = 1.f / eyx, exz = 1.f / ezx, eyz = 1.f / ezy;
    MediumSampleO out;
    out.local_thp = is_medium ? V4{3.f / (st[0] + st[1] * eyx + st[2] * ezx), 3.f / (st[0] * exy + st[1] + st[2] * ezy),
                                   3.f / (st[0] * exz + st[1] * eyz + st[2]), 1.f} * v4(m->sigma_s)
                              : V4{3.f / (1.f + eyx + ezx), 3.f / (exy + 1.f + ezy), 3.f / (exz + eyz + 1.f), 1.f};
    out.dist = dist, out.flag = int(is_medium);
    return out;
}
V4 medium_transmittance(const CptsMedium *m, float dist) {
    if (!m)
        return v4(1.f);
    return {std::exp(-(m->sigma_a[0] + m->sigma_s[0]) * dist), std::exp(-(m->sigma_a[1] + m->sigma_s[1]) * dist),
            std::exp(-(m->sigma_a[2] + m->sigma_s[2]) * dist), 1.f};
}
float hg_phase(float c, float g) {
    const float g2 = g * g, denom = 1.f + g2 - 2.f * g * c;
    return INV_PI * 0.25f * (1.f - g2) / denom / std::sqrt(denom);
}
V3 hg_sample(Rng &rng, float g) {
    const float ux = rng.f(), uy = rng.f(), g2 = g * g;
    const float sq = (1.f - g2) / (1.f + g - 2.f * g * ux), ct = (1.f + g2 - sq * sq) / (2.f * g), st = std::sqrt(std::max(0.f, 1.f - ct * ct));
    return {std::cos(2.f * uy * PI) * st, std::sin(2.f * uy * PI) * st, ct};
}
float phase_eval(const CptsMedium &m, V3 a, V3 b) {
    const float c = dot(a, b);
    switch (m.phase_type) {
    case CPTS_PHASE_SGGX: // same as isotropic upstream (sggx.cuh:30-45)
    case CPTS_PHASE_ISOTROPIC: return INV_PI * 0.25f;
    case CPTS_PHASE_RAYLEIGH: return 3.f / 16.f * INV_PI * (1.f + c); // rayleigh.cuh:35-37
    case CPTS_PHASE_HG: return hg_phase(c, m.phase[0]);
    case CPTS_PHASE_DUO_HG: return hg_phase(c, m.phase[0]) * m.phase[2] + hg_phase(c, m.phase[1]) * (1.f - m.phase[2]);
    default: return 0.f;
    }
}
V3 medium_scatter(const CptsMedium &m, V3 raydir, V4 &thp, Rng &rng, float &pdf) {
    V3 local{0, 0, 1};
    float weight = 1.f;
    if (m.phase_type == CPTS_PHASE_ISOTROPIC || m.phase_type == CPTS_PHASE_SGGX) {
        const float ux = rng.f(), uy = rng.f(), ct = 2.f * ux - 1.f, st = std::sqrt(1.f - ct * ct);
        local = {std::cos(2.f * uy * PI) * st, std::sin(2.f * uy * PI) * st, ct};
    } else if (m.phase_type == CPTS_PHASE_RAYLEIGH) { // rayleigh.cuh:39-48, weight 1
        const float ux = rng.f(), uy = rng.f(), u = ux * 2.f - 1.f;
        const float ct = std::cbrt(2.f * u + std::sqrt(4.f * u * u + 1.f)), st = std::sqrt(std::max(0.f, 1.f - ct * ct));
        local = {std::cos(2.f * uy * PI) * st, std::sin(2.f * uy * PI) * st, ct};
    } else if (m.phase_type == CPTS_PHASE_HG) {
        local = hg_sample(rng, m.phase[0]);
    } else if (m.phase_type == CPTS_PHASE_DUO_HG) {
        const V3 d1 = hg_sample(rng, m.phase[0]), d2 = hg_sample(rng, m.phase[1]);
        const float p1 = hg_phase(d1.z, m.phase[0]), p2 = hg_phase(d2.z, m.phase[1]), w = m.phase[2];
        const bool first = rng.f() < w;
        weight = (first ? p1 : p2) / (w * p1 + (1.f - w) * p2);
        local = first ? d1 : d2;
    }
    thp = thp * weight;
    pdf = weight;
    return to_world(raydir, local);
}

// ---- grid media (src/volume/grid.cuh:51-92 nearest-voxel lookups, src/impl/vol_grid.cu:97-211 delta / ratio tracking,
// blackbody emission).  The NanoVDB grids arrive as brick grids (include/cpts_format.h): same values, one table read.
long long grid_voxel(const Scene &sc, const CptsGridChannel &c, V3 pos, float offset) {
    const float x = pos.x - c.vec[0], y = pos.y - c.vec[1], z = pos.z - c.vec[2];
    float f[3];
    for (int r = 0; r < 3; r++) // NanoVDB math::matMult(float*): fmaf(x, m0, fmaf(y, m1, z * m2))
        f[r] = std::fmaf(x, c.inv_mat[3 * r], std::fmaf(y, c.inv_mat[3 * r + 1], z * c.inv_mat[3 * r + 2])) + offset;
    const int i = int(std::round(f[0])) - c.bmin[0], j = int(std::round(f[1])) - c.bmin[1], k = int(std::round(f[2])) - c.bmin[2];
    if (unsigned(i) >= unsigned(c.nb[0] * 8) || unsigned(j) >= unsigned(c.nb[1] * 8) || unsigned(k) >= unsigned(c.nb[2] * 8))
        return -1;
    const int b = sc.grid_bricks[c.brick_offset + size_t((i >> 3) * c.nb[1] + (j >> 3)) * c.nb[2] + (k >> 3)];
    if (b < 0)
        return -1;
    return (long long)c.data_offset + ((long long)b * 512 + (((i & 7) << 6) | ((j & 7) << 3) | (k & 7))) * c.ncomp;
}
float grid_scalar(const Scene &sc, const CptsGridChannel &c, V3 pos, float offset) {
    const long long a = grid_voxel(sc, c, pos, offset);
    return a < 0 ? c.background[0] : sc.grid_data[a];
}
bool grid_span(const CptsGrid &g, V3 o, V3 d, float &t_near, float &t_far) { // AABB::intersect (aabb.cuh:78-85)
    const float od[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    t_near = -INFINITY, t_far = INFINITY;
    float lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        const float inv = 1.f / dd[k], od_inv = od[k] * inv;
        const float a = std::fmaf(g.aabb_min[k], inv, -od_inv), b = std::fmaf(g.aabb_max[k], inv, -od_inv);
        lo[k] = a < b ? a : b, hi[k] = a < b ? b : a;
    }
    t_near = std::fmax(lo[0], std::fmax(lo[1], lo[2]));
    t_far = std::fmin(hi[0], std::fmin(hi[1], hi[2]));
    return t_far > t_near && t_far > 0;
}
MediumSampleO grid_sample(const Scene &sc, const CptsMedium &m, V3 o, V3 d, float hit_t, Rng &rng) {
    const CptsGrid &g = sc.grids[m.grid];
    MediumSampleO ds{v4(1.f), hit_t, 0};
    float t_near = 0, t_far = 0;
    if (!grid_span(g, o, d, t_near, t_far))
        return ds;
    t_near = std::max(0.f, t_near), t_far = std::min(t_far, hit_t);
    const float inv_maj = 1.f / (g.scale * g.majorant);
    float t = t_near - std::log(1.f - rng.f()) * inv_maj;
    while (t < t_far) {
        const V3 pos = o + d * t;
        const float offset = rng.f() - 0.5f;
        const float dens = grid_scalar(sc, g.ch[CPTS_GRID_DENSITY], pos, offset) * g.scale;
        if (rng.f() < dens * inv_maj) { // real collision: throughput = albedo
            ds.dist = t, ds.flag = 1;
            const CptsGridChannel &ac = g.ch[CPTS_GRID_ALBEDO];
            if (ac.ncomp == 0) {
                ds.local_thp = {m.sigma_a[0], m.sigma_a[1], m.sigma_a[2], 1.f};
            } else {
                const long long a = grid_voxel(sc, ac, pos, offset);
                ds.local_thp = a < 0 ? V4{ac.background[0], ac.background[1], ac.background[2], 1.f}
                                     : V4{sc.grid_data[a], sc.grid_data[a + 1], sc.grid_data[a + 2], 1.f};
            }
            break;
        }
        t -= std::log(1.f - rng.f()) * inv_maj;
    }
    return ds;
}
// ratio tracking with roulette below 0.1; `far_cap` = the ray's hit_t (MAX_DIST in the NEE march)
V4 grid_transmittance(const Scene &sc, const CptsMedium &m, V3 o, V3 d, float far_cap, Rng &rng) {
    const CptsGrid &g = sc.grids[m.grid];
    float Tr = 1.f, t_near = 0, t_far = 0;
    if (grid_span(g, o, d, t_near, t_far)) {
        t_near = std::max(0.f, t_near), t_far = std::min(t_far, far_cap);
        const float inv_maj = 1.f / (g.scale * g.majorant);
        float t = t_near - std::log(1.f - rng.f()) * inv_maj;
        while (t < t_far) {
            const V3 pos = o + d * t;
            const float offset = rng.f() - 0.5f;
            const float dens = grid_scalar(sc, g.ch[CPTS_GRID_DENSITY], pos, offset) * g.scale;
            Tr *= std::max(0.f, 1.f - dens * inv_maj);
            if (Tr < 0.1f) {
                if (rng.f() >= Tr)
                    return {0, 0, 0, 1};
                Tr = 1.f;
            }
            t -= std::log(1.f - rng.f()) * inv_maj;
        }
    }
    return {Tr, Tr, Tr, 1.f};
}
V4 grid_emission(const Scene &sc, const CptsMedium &m, V3 pos, Rng &rng) { // query_emission: the draw always happens
    const CptsGrid &g = sc.grids[m.grid];
    const float offset = rng.f() - 0.5f;
    const float temp = g.ch[CPTS_GRID_TEMPERATURE].ncomp ? grid_scalar(sc, g.ch[CPTS_GRID_TEMPERATURE], pos, offset) : 0.f;
    const float x = temp * g.temp_scale;
    return V4{table_lookup(sc.blackbody, 4, 0, x), table_lookup(sc.blackbody, 4, 1, x), table_lookup(sc.blackbody, 4, 2, x),
              table_lookup(sc.blackbody, 4, 3, x)} * g.em_scale;
}
struct BankStackO {
    int depth = 0, v[3] = {0, 0, 0};
    int top() const { return depth > 0 ? v[depth - 1] : 0; }
    void push(int x) {
        if (depth < 3 && (x & 0xff) != 0xff)
            v[depth++] = x & 0xff;
    }
    int pop() { return depth > 0 ? v[--depth] : 0; }
};
const CptsMedium *medium_of(const Scene &sc, int idx) { return (idx > 0 && size_t(idx) < sc.media.size()) ? &sc.media[idx] : nullptr; }
V3 interp_normal(const Scene &sc, int prim, float u, float v) {
    auto at = [&](int k) { return V3{sc.n[k][3 * prim], sc.n[k][3 * prim + 1], sc.n[k][3 * prim + 2]}; };
    return normalize(at(0) * (1.f - u - v) + at(1) * u + at(2) * v);
}
// occlusion_transmittance_estimate (megakernel_vpt.cu:104-201)
V4 nee_transmittance(const Scene &sc, V3 o, V3 d, BankStackO nested, float max_dist, Rng &rng) {
    float total = 0;
    V4 Tr = v4(1.f);
    while (total < max_dist) {
        const float bound = max_dist - total;
        const Hit h = closest_hit(sc, o, d, bound);
        const float seg = h.prim >= 0 ? h.t : bound;
        const int info = h.prim >= 0 ? sc.obj_idxs[h.prim] : -1;
        int active = (info >> 20) & 0xff;
        if (active == 0 || (info & 0x40000000) == 0)
            return v4(0.f);
        if (h.prim >= 0 && dot(interp_normal(sc, h.prim, h.u, h.v), d) > 0) {
            active = nested.pop();
        } else {
            const int old = nested.top();
            nested.push(active);
            active = old;
        }
        const CptsMedium *am = medium_of(sc, active);
        Tr = Tr * (am && am->type == CPTS_MEDIUM_GRID ? grid_transmittance(sc, *am, o, d, MAX_DIST, rng) : medium_transmittance(am, seg));
        total += seg;
        o = o + d * seg;
    }
    return Tr;
}

// ---- one path under the megakernel rules: rules 1 = megakernel_pt.cu:59-186, rules 2 = megakernel_vpt.cu:203-456 -------------
V4 trace_path_megakernel(const Scene &sc, int px, int py, int accum, int seed_offset, int max_depth, int rules, Counts &cnt) {
    const int W = sc.cfg[CPTS_CFG_WIDTH];
    Rng rng;
    rng.seed(px + py * W, int(uint32_t(seed_offset) + uint32_t(accum) * uint32_t(SEED_SCALER)));
    V3 o, d;
    camera_ray(sc, px, py, rng, o, d);
    const int max_diffuse = sc.cfg[CPTS_CFG_MAX_DIFFUSE], max_specular = sc.cfg[CPTS_CFG_MAX_SPECULAR], max_transmit = sc.cfg[CPTS_CFG_MAX_TRANSMIT];
    const int max_volume = sc.cfg[CPTS_CFG_MAX_VOLUME] > 0 ? sc.cfg[CPTS_CFG_MAX_VOLUME] : max_depth;
    const int rr_bounce = rules == 2 ? 2 : 1; // megakernel_vpt.cu:28 / megakernel_pt.cu:27
    const int num_emitters = int(sc.emitters.size()) - 1;
    BankStackO nested;
    if (rules == 2 && sc.cfg[CPTS_CFG_CAM_VOL] > 0)
        nested.push(sc.cfg[CPTS_CFG_CAM_VOL]);
    V4 thp{1, 1, 1, 1}, radiance{0, 0, 0, 1};
    float ew = 1.f;
    bool delta = false;
    int diff_b = 0, spec_b = 0, trans_b = 0, volm_b = 0;
    for (int b = 0; b < max_depth; b++) {
        const Hit h = closest_hit(sc, o, d);
        if (b > 0)
            cnt.closest++;
        if (h.prim < 0) {
            radiance = radiance + thp * null_or_env(sc, d);
            break;
        }
        const int info = sc.obj_idxs[h.prim], object_id = info & 0xfffff, hit_med = (info >> 20) & 0xff;
        const CptsObject &ob = sc.objects[object_id];
        const int mat_id = ob.bsdf_emitter >> 16, emitter_id = ((ob.bsdf_emitter & 0xFFFF) ^ 0x8000) - 0x8000;
        const CptsMedium *medium = rules == 2 ? medium_of(sc, nested.top()) : nullptr;
        MediumSampleO md{v4(1.f), h.t, 0};
        if (rules == 2) {
            md = medium && medium->type == CPTS_MEDIUM_GRID ? grid_sample(sc, *medium, o, d, h.t, rng) : medium_sample(medium, h.t, rng);
            thp = thp * md.local_thp;
            if ((info & 0x40000000) && md.flag == 0) { // cullable boundary: step through, no shading
                o = o + d * md.dist;
                if (dot(d, interaction(sc, h.prim, h.u, h.v).n) > 0)
                    nested.pop();
                else
                    nested.push(hit_med);
                continue;
            }
        }
        cnt.shaded++;
        const Isect it = interaction(sc, h.prim, h.u, h.v);
        CptsBsdf local_bsdf;
        const Mat m = resolve_material(sc, mat_id, it.u, it.v, local_bsdf);
        const V3 so = o + d * md.dist;
        if (rules == 1) { // emission before the emitter is sampled (megakernel_pt.cu:96-110)
            const float sa = ob.inv_area * h.t * h.t / std::max(std::fabs(dot(d, shading_normal(it.n, nullptr, m.pn))), 1e-4f);
            ew = ew / (ew + sa * float(emitter_id > 0) * float(b > 0) * float(!delta));
            radiance = radiance + (thp * emitter_eval(sc, sc.emitters[emitter_id], d, it.n, it.u, it.v)) * ew;
        }
        const int no_sample = (rules == 2 && md.flag > 0) ? 0 : emitter_id;
        int num = num_emitters;
        num -= no_sample > 0 && num > 1;
        uint32_t eid = rng.u32() % uint32_t(num) + 1;
        eid += eid >= uint32_t(no_sample) && no_sample > 0;
        float direct_pdf = 1.f / float(num);
        eid *= uint32_t(no_sample == 0 || num > 1);
        const CptsEmitter &em = sc.emitters[eid];
        const CptsObject &eo = sc.objects[em.obj_ref * (em.obj_ref >= 0)];
        direct_pdf *= eo.inv_area;
        const uint32_t pick = rng.u32() % uint32_t(eo.prim_num) + uint32_t(eo.prim_offset);
        const int eprim = pick < sc.emitter_prims.size() ? sc.emitter_prims[pick] : 0;
        V4 direct{0, 0, 0, 1};
        const float ux = rng.f(), uy = rng.f();
        V3 sd = emitter_sample(sc, em, so, rules == 2 ? V3{0, 0, 1} : it.n, direct, direct_pdf, ux, uy, eprim) - so;
        const float dist = std::sqrt(len2(sd));
        sd = sd * (1.f / dist);
        int lobe = 0;
        if (rules == 1) {
            if (eid != 0) {
                cnt.shadow++;
                if (unoccluded(sc, so, sd, dist - EPS)) {
                    const float mis = direct_pdf + bsdf_pdf(sc, m, it.n, sd, d) * float(em.is_sphere != 0);
                    radiance = radiance + thp * direct * bsdf_eval(sc, m, it.n, sd, d) * (mis > EPS ? 1.f / mis : 0.f);
                }
            }
            o = so;
            d = bsdf_sample(sc, m, d, it.n, thp, ew, rng, lobe);
            delta = (lobe & LOBE_SPECULAR) != 0;
        } else {
            V4 nee_tr{0, 0, 0, 1}, nee_thp;
            if (eid != 0) {
                cnt.shadow++;
                nee_tr = nee_transmittance(sc, so, sd, nested, dist - EPS, rng);
            }
            float mis;
            if (md.flag > 0) { // medium event
                radiance = radiance + thp * (medium->type == CPTS_MEDIUM_GRID ? grid_emission(sc, *medium, so, rng) : V4{0, 0, 0, 1});
                const float pp = phase_eval(*medium, sd, d);
                nee_thp = v4(pp, 1.f);
                mis = direct_pdf + pp;
                lobe = LOBE_VOLUME;
                d = medium_scatter(*medium, d, thp, rng, ew);
                delta = false;
            } else {
                const float sa = ob.inv_area * h.t * h.t / std::max(std::fabs(dot(d, shading_normal(it.n, nullptr, m.pn))), 1e-4f);
                ew = ew / (ew + sa * float(emitter_id > 0) * float(b > 0) * float(!delta));
                radiance = radiance + (thp * emitter_eval(sc, sc.emitters[emitter_id], d, it.n, it.u, it.v)) * ew;
                mis = direct_pdf + bsdf_pdf(sc, m, it.n, sd, d) * float(em.is_sphere != 0);
                nee_thp = bsdf_eval(sc, m, it.n, sd, d);
                const bool same = dot(d, it.n) > 0;
                d = bsdf_sample(sc, m, d, it.n, thp, ew, rng, lobe);
                delta = (lobe & LOBE_SPECULAR) != 0;
                if (((dot(it.n, d) > 0) != same) == false) {
                    if (same)
                        nested.pop();
                    else
                        nested.push(hit_med);
                }
            }
            radiance = radiance + nee_tr * thp * direct * (nee_thp * (mis > EPS ? 1.f / mis : 0.f));
            o = so;
        }
        if (std::isnan(radiance.x) || std::isnan(radiance.y) || std::isnan(radiance.z) || std::isinf(radiance.x) || std::isinf(radiance.y) ||
            std::isinf(radiance.z))
            radiance = v4(0.f);
        diff_b += (lobe & LOBE_DIFFUSE) != 0, spec_b += (lobe & LOBE_SPECULAR) != 0, trans_b += (lobe & LOBE_TRANSMIT) != 0,
            volm_b += (lobe & LOBE_VOLUME) != 0;
        if (diff_b >= max_diffuse || spec_b >= max_specular || trans_b >= max_transmit || (rules == 2 && volm_b >= max_volume))
            break;
        const float mx = std::max(thp.x, std::max(thp.y, thp.z));
        if (mx < THP_EPS)
            break;
        if (b >= rr_bounce && mx < RR_THRESHOLD) {
            if (rng.f() > mx)
                break;
            thp = thp * float(1. / mx);
        }
    }
    return radiance;
}

} // namespace

// ---------------------------------------------------------------------------------------------------
extern "C" {

static thread_local std::string g_err;
const char *oracle_last_error() { return g_err.c_str(); }

void *oracle_load(const char *path) {
    try {
        return load_scene(path);
    } catch (const std::string &e) {
        g_err = e;
        return nullptr;
    }
}
void oracle_free(void *h) { delete static_cast<Scene *>(h); }
void oracle_config(void *h, int32_t *cfg32) { memcpy(cfg32, static_cast<Scene *>(h)->cfg, sizeof(int32_t) * CPTS_CFG_COUNT); }

/* Adds `spp` samples (accum counters first_accum .. first_accum+spp-1, as render_raw numbers them) to the
 * running SUM image `sum_rgba` (H*W*4 floats) in sample order with the splat's NaN/Inf scrub
 * (wavefront_pt.cu:486-487).  max_depth <= 0 uses the scene's.  stats3 (optional): closest, shadow, shaded.
 * Rows [row0,row1) only (row1 <= 0: all) so that a timing run can take a bounded sample. */
int oracle_render(void *hs, int first_accum, int spp, int seed_offset, int max_depth, int row0, int row1,
                  float *sum_rgba, uint64_t *stats3) {
    Scene &sc = *static_cast<Scene *>(hs);
    const int W = sc.cfg[CPTS_CFG_WIDTH], H = sc.cfg[CPTS_CFG_HEIGHT];
    if (max_depth <= 0)
        max_depth = sc.cfg[CPTS_CFG_MAX_DEPTH];
    if (row1 <= 0 || row1 > H)
        row1 = H;
    uint64_t c0 = 0, c1 = 0, c2 = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : c0, c1, c2)
    for (int y = row0; y < row1; y++) {
        Counts cnt;
        for (int x = 0; x < W; x++) {
            float *dst = sum_rgba + 4 * (size_t(y) * W + x);
            for (int s = 0; s < spp; s++) {
                V4 L = trace_path(sc, x, y, first_accum + s, seed_offset, max_depth, cnt);
                if (std::isnan(L.x) || std::isnan(L.y) || std::isnan(L.z) || std::isinf(L.x) || std::isinf(L.y) || std::isinf(L.z))
                    L = V4{0, 0, 0, 1};
                dst[0] += L.x, dst[1] += L.y, dst[2] += L.z, dst[3] += L.w;
            }
        }
        c0 += cnt.closest, c1 += cnt.shadow, c2 += cnt.shaded;
    }
    if (stats3)
        stats3[0] = c0, stats3[1] = c1, stats3[2] = c2;
    return 0;
}

/* Same, under the termination rules of the reference's megakernels: rules 1 = "pt" (megakernel_pt.cu), rules 2 = "vpt" with
 * participating media (megakernel_vpt.cu).  No NaN scrub at the splat (these kernels scrub per bounce). */
int oracle_render_megakernel(void *hs, int rules, int first_accum, int spp, int seed_offset, int max_depth, int row0, int row1,
                             float *sum_rgba, uint64_t *stats3) {
    Scene &sc = *static_cast<Scene *>(hs);
    const int W = sc.cfg[CPTS_CFG_WIDTH], H = sc.cfg[CPTS_CFG_HEIGHT];
    if (rules != 1 && rules != 2)
        return 1;
    if (max_depth <= 0)
        max_depth = sc.cfg[CPTS_CFG_MAX_DEPTH];
    if (row1 <= 0 || row1 > H)
        row1 = H;
    uint64_t c0 = 0, c1 = 0, c2 = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : c0, c1, c2)
    for (int y = row0; y < row1; y++) {
        Counts cnt;
        for (int x = 0; x < W; x++) {
            float *dst = sum_rgba + 4 * (size_t(y) * W + x);
            for (int s = 0; s < spp; s++) {
                const V4 L = trace_path_megakernel(sc, x, y, first_accum + s, seed_offset, max_depth, rules, cnt);
                dst[0] += L.x, dst[1] += L.y, dst[2] += L.z, dst[3] += L.w;
            }
        }
        c0 += cnt.closest, c1 += cnt.shadow, c2 += cnt.shaded;
    }
    if (stats3)
        stats3[0] = c0, stats3[1] = c1, stats3[2] = c2;
    return 0;
}

/* known-answer helpers */
void oracle_rng(int seed, int offset, int n, uint32_t *out) {
    Rng r;
    r.seed(seed, offset);
    for (int i = 0; i < n; i++)
        out[i] = r.u32();
}
uint16_t oracle_float_to_half(float f) { return float_to_half(f); }
float oracle_half_to_float(uint16_t h) { return half_to_float(h); }
/* closest hit along (o,d): returns primitive index or -1; out4 = {t,u,v,0} */
int oracle_closest_hit(void *hs, const float *o, const float *d, float *out4) {
    const Hit h = closest_hit(*static_cast<Scene *>(hs), V3{o[0], o[1], o[2]}, V3{d[0], d[1], d[2]});
    out4[0] = h.t, out4[1] = h.u, out4[2] = h.v, out4[3] = 0;
    return h.prim;
}
int oracle_unoccluded(void *hs, const float *o, const float *d, float max_dist) {
    return unoccluded(*static_cast<Scene *>(hs), V3{o[0], o[1], o[2]}, V3{d[0], d[1], d[2]}, max_dist);
}
/* brute force over every primitive (no BVH): the traversal's own oracle */
int oracle_closest_hit_bruteforce(void *hs, const float *o, const float *d, float *out4) {
    const Scene &sc = *static_cast<Scene *>(hs);
    Hit h{MAX_DIST, 0, 0, -1};
    const int P = int(sc.tris.size() / 12);
    for (int i = 0; i < P; i++) {
        float u, v;
        const float t = hit_tri(sc, i, V3{o[0], o[1], o[2]}, V3{d[0], d[1], d[2]}, u, v, INFINITY);
        if (t > EPS && t < h.t)
            h = Hit{t, u, v, i};
    }
    out4[0] = h.t, out4[1] = h.u, out4[2] = h.v, out4[3] = 0;
    return h.prim;
}
}
