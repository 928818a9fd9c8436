// volume.cuh — participating media for the volumetric path: homogeneous media with isotropic,
// Henyey-Greenstein and two-lobe HG phase functions, and the 3-deep nested-volume stack.
//
// Reference: src/volume/homogeneous.cuh:28-88 (distance sampling, transmittance), src/core/medium.cuh:72-112 (Medium
// base = vacuum: surface event, transmittance 1), src/volume/henyey_greenstein.cuh:28-108 (phase functions),
// src/pt_impl/megakernel_vpt.cu:59-94 (BankStack).  POD tables instead of device-side virtual objects.
#pragma once
#include "pt_device.cuh"

namespace cptb {

struct MediumPOD { // 80 B
    int type, phase_type; // type < 0: vacuum (media[0]); 0: homogeneous; 1: grid
    int grid, pad1;       // grid media: index into DevScene::grids
    float4 sigma_a, sigma_s, sigma_t;
    float4 phase; // hg: g ; duo-hg: g1, g2, weight
};

struct MediumSample {
    V4 local_thp;
    float dist;
    uint32_t flag; // 1: medium (scattering) event, 0: surface event
};

// ---- grid media: brick grids in place of the reference's NanoVDB trees (include/cpts_format.h "grids") ----
// One lookup = one brick-table read + one voxel read (two dependent loads), where the reference's accessor walks
// root -> upper -> lower -> leaf for every jittered sample.
struct GridChannelPOD {
    float inv_mat[9], vec[3];
    int bmin[3], nb[3];
    float bg[3];
    int ncomp;
    const int *bricks;
    const float *data;
};
struct GridPOD {
    float scale, temp_scale, em_scale, majorant;
    float aabb_min[3], aabb_max[3];
    GridChannelPOD ch[3]; // density, albedo, temperature
};
enum : int { MEDIUM_HOMOGENEOUS = 0, MEDIUM_GRID = 1 };

enum : int { PHASE_NULL = 0, PHASE_ISOTROPIC = 1, PHASE_HG = 2, PHASE_DUO_HG = 3, PHASE_RAYLEIGH = 4, PHASE_SGGX = 5 };

// ---- nested volumes: one 4-byte bank, x = depth, y/z/w = medium ids (megakernel_vpt.cu:59-94) ----
struct BankStack {
    uint32_t bits; // byte 0 = x (depth), bytes 1..3 = y, z, w
    PT_DEV static BankStack make(int val) {
        BankStack s{0u};
        if (val > 0)
            s.bits = 1u | (uint32_t(uint8_t(val)) << 8);
        return s;
    }
    PT_DEV int depth() const { return int(bits & 0xffu); }
    PT_DEV int top() const {
        const int d = depth();
        return d > 0 ? int((bits >> (8 * d)) & 0xffu) : 0;
    }
    PT_DEV void push(int val) {
        const int d = depth();
        const uint32_t v = uint32_t(val) & 0xffu;
        if (d < 3 && v != 0xffu) {
            const int nd = d + 1;
            bits = (bits & ~(0xffu << (8 * nd)) & ~0xffu) | (v << (8 * nd)) | uint32_t(nd);
        }
    }
    PT_DEV int pop() {
        int res = 0;
        const int d = depth();
        if (d > 0) {
            res = int((bits >> (8 * d)) & 0xffu);
            bits = (bits & ~0xffu) | uint32_t(d - 1);
        }
        return res;
    }
};

// ---- distance sampling (homogeneous.cuh:45-82; vacuum: medium.cuh:84-87, no sampler draws) ----
PT_DEV MediumSample medium_sample(const MediumPOD &m, float hit_t, Rng &rng) {
    MediumSample msp;
    if (m.type < 0) {
        msp.local_thp = splat4(1.f);
        msp.dist = hit_t;
        msp.flag = 0;
        return msp;
    }
    // TinySampler::discrete1D returns a signed int (sampler.cuh:49-58), so `% 3` yields -2..2 and `sigma_t[channel]`
    // reads the members laid out before sigma_t for negative draws: [-1] = sigma_s.w, [-2] = sigma_s.z
    // (homogeneous.cuh:31-34,48-49).  Reproduced: it changes the sampled distance by sigma_a / sigma_t.
    const int channel = int(rng.next_u32()) % 3;
    const float st_c = channel == 0 ? m.sigma_t.x
                                    : (channel == 1 ? m.sigma_t.y : (channel == 2 ? m.sigma_t.z : (channel == -1 ? m.sigma_s.w : m.sigma_s.z)));
    msp.dist = -logf(1.f - rng.next1d()) / st_c;
    const bool is_medium = msp.dist < hit_t;
    msp.dist = fminf(msp.dist, hit_t);
    const float exp_y_x = expf(-(m.sigma_t.y - m.sigma_t.x) * msp.dist), exp_z_x = expf(-(m.sigma_t.z - m.sigma_t.x) * msp.dist),
                exp_z_y = expf(-(m.sigma_t.z - m.sigma_t.y) * msp.dist), exp_x_y = __frcp_rn(exp_y_x), exp_x_z = __frcp_rn(exp_z_x),
                exp_y_z = __frcp_rn(exp_z_y);
    msp.local_thp = is_medium ? mk4(3.f / (m.sigma_t.x + m.sigma_t.y * exp_y_x + m.sigma_t.z * exp_z_x),
                                    3.f / (m.sigma_t.x * exp_x_y + m.sigma_t.y + m.sigma_t.z * exp_z_y),
                                    3.f / (m.sigma_t.x * exp_x_z + m.sigma_t.y * exp_y_z + m.sigma_t.z)) *
                                    from_f4(m.sigma_s)
                              : mk4(3.f / (1.f + exp_y_x + exp_z_x), 3.f / (exp_x_y + 1.f + exp_z_y), 3.f / (exp_x_z + exp_y_z + 1.f));
    msp.flag = uint32_t(is_medium);
    return msp;
}

// transmittance over `dist` (homogeneous.cuh:84-87; vacuum: 1)
PT_DEV V4 medium_transmittance(const MediumPOD &m, float dist) {
    if (m.type < 0)
        return splat4(1.f);
    return mk4(expf(-m.sigma_t.x * dist), expf(-m.sigma_t.y * dist), expf(-m.sigma_t.z * dist), 1.f);
}

// ---- grid media (src/volume/grid.cuh:51-92, src/impl/vol_grid.cu:97-211) ----
// nearest-voxel address of a jittered world position: index = inv_mat * (pos - vec) (NanoVDB math::matMult: two fmaf
// and one rounded product per row), + offset, roundf.  Returns the float offset into the channel's pool or -1.
PT_DEV long long grid_voxel(const GridChannelPOD &c, V3 pos, float offset) {
    const float x = pos.x - c.vec[0], y = pos.y - c.vec[1], z = pos.z - c.vec[2];
    const float fi = __fadd_rn(fmaf(x, c.inv_mat[0], fmaf(y, c.inv_mat[1], __fmul_rn(z, c.inv_mat[2]))), offset);
    const float fj = __fadd_rn(fmaf(x, c.inv_mat[3], fmaf(y, c.inv_mat[4], __fmul_rn(z, c.inv_mat[5]))), offset);
    const float fk = __fadd_rn(fmaf(x, c.inv_mat[6], fmaf(y, c.inv_mat[7], __fmul_rn(z, c.inv_mat[8]))), offset);
    const int i = int(roundf(fi)) - c.bmin[0], j = int(roundf(fj)) - c.bmin[1], k = int(roundf(fk)) - c.bmin[2];
    if (unsigned(i) >= unsigned(c.nb[0] << 3) || unsigned(j) >= unsigned(c.nb[1] << 3) || unsigned(k) >= unsigned(c.nb[2] << 3))
        return -1;
    const int b = __ldg(c.bricks + ((i >> 3) * c.nb[1] + (j >> 3)) * c.nb[2] + (k >> 3));
    if (b < 0)
        return -1;
    return ((long long)b * 512 + (((i & 7) << 6) | ((j & 7) << 3) | (k & 7))) * c.ncomp;
}
PT_DEV float grid_scalar(const GridChannelPOD &c, V3 pos, float offset) {
    const long long a = grid_voxel(c, pos, offset);
    return a < 0 ? c.bg[0] : __ldg(c.data + a);
}
PT_DEV V4 grid_albedo(const GridPOD &g, const MediumPOD &m, V3 pos, float offset) {
    const GridChannelPOD &c = g.ch[1];
    if (c.ncomp == 0) // constant albedo
        return mk4(m.sigma_a.x, m.sigma_a.y, m.sigma_a.z);
    const long long a = grid_voxel(c, pos, offset);
    return a < 0 ? mk4(c.bg[0], c.bg[1], c.bg[2]) : mk4(__ldg(c.data + a), __ldg(c.data + a + 1), __ldg(c.data + a + 2));
}
// AABB::intersect(ray, t_near, t_far) (src/core/aabb.cuh:78-85)
PT_DEV bool grid_span(const GridPOD &g, V3 o, V3 d, float &t_near, float &t_far) {
    const float ix = __frcp_rn(d.x), iy = __frcp_rn(d.y), iz = __frcp_rn(d.z);
    const float ox = __fmul_rn(o.x, ix), oy = __fmul_rn(o.y, iy), oz = __fmul_rn(o.z, iz);
    const float ax = fmaf(g.aabb_min[0], ix, -ox), ay = fmaf(g.aabb_min[1], iy, -oy), az = fmaf(g.aabb_min[2], iz, -oz);
    const float bx = fmaf(g.aabb_max[0], ix, -ox), by = fmaf(g.aabb_max[1], iy, -oy), bz = fmaf(g.aabb_max[2], iz, -oz);
    const float min_x = ax < bx ? ax : bx, max_x = ax < bx ? bx : ax;
    const float min_y = ay < by ? ay : by, max_y = ay < by ? by : ay;
    const float min_z = az < bz ? az : bz, max_z = az < bz ? bz : az;
    t_near = fmaxf(min_x, fmaxf(min_y, min_z));
    t_far = fminf(max_x, fminf(max_y, max_z));
    return t_far > t_near && t_far > 0;
}
// GridVolumeMedium::sample -> delta tracking (vol_grid.cu:104-150)
PT_DEV MediumSample grid_sample(const GridPOD &g, const MediumPOD &m, V3 o, V3 d, float hit_t, Rng &rng) {
    MediumSample ds;
    ds.local_thp = splat4(1.f);
    ds.dist = hit_t;
    ds.flag = 0;
    float t_near = 0, t_far = 0;
    if (grid_span(g, o, d, t_near, t_far)) {
        t_near = fmaxf(0.f, t_near);
        t_far = fminf(t_far, hit_t);
        const float inv_maj = 1.f / (g.scale * g.majorant);
        float t = t_near - logf(1.f - rng.next1d()) * inv_maj;
        while (t < t_far) {
            const V3 pos{fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z)};
            const float offset = rng.next1d() - 0.5f;
            const float dens = grid_scalar(g.ch[0], pos, offset) * g.scale;
            if (rng.next1d() < dens * inv_maj) { // real collision
                ds.dist = t;
                ds.local_thp = grid_albedo(g, m, pos, offset);
                ds.flag = 1;
                break;
            }
            t -= logf(1.f - rng.next1d()) * inv_maj;
        }
    }
    return ds;
}
// GridVolumeMedium::transmittance -> ratio tracking with roulette below 0.1 (vol_grid.cu:117-127,179-199).  `far_cap`
// is the ray's hit_t: the NEE march of the reference hands over a ray whose hit_t is still MAX_DIST, so the estimate
// runs over the whole span of the grid's box in front of the origin (reproduced by the caller passing kMaxDist).
PT_DEV V4 grid_transmittance(const GridPOD &g, V3 o, V3 d, float far_cap, Rng &rng) {
    float Tr = 1.f, t_near = 0, t_far = 0;
    if (grid_span(g, o, d, t_near, t_far)) {
        t_near = fmaxf(0.f, t_near);
        t_far = fminf(t_far, far_cap);
        const float inv_maj = 1.f / (g.scale * g.majorant);
        float t = t_near - logf(1.f - rng.next1d()) * inv_maj;
        bool dead = false;
        while (t < t_far) {
            const V3 pos{fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z)};
            const float offset = rng.next1d() - 0.5f;
            const float dens = grid_scalar(g.ch[0], pos, offset) * g.scale;
            Tr *= fmaxf(0.f, 1.f - dens * inv_maj);
            if (Tr < 0.1f) {
                if (rng.next1d() >= Tr) {
                    dead = true;
                    break;
                }
                Tr = 1.f;
            }
            t -= logf(1.f - rng.next1d()) * inv_maj;
        }
        if (dead)
            Tr = 0.f;
    }
    return mk4(Tr, Tr, Tr, 1.f);
}
// GridVolumeMedium::query_emission (vol_grid.cu:97-102): the jitter draw happens whether or not a temperature grid exists
PT_DEV V4 grid_emission(const GridPOD &g, cudaTextureObject_t blackbody, V3 pos, Rng &rng) {
    const float offset = rng.next1d() - 0.5f;
    const float temp = g.ch[2].ncomp ? grid_scalar(g.ch[2], pos, offset) : 0.f;
    return from_f4(tex1D<float4>(blackbody, temp * g.temp_scale)) * g.em_scale;
}

// ---- phase functions (henyey_greenstein.cuh) ----
// g^2 is a stored member of the reference's phase objects (rounded once in the constructor): keep it a rounded
// product here, or `1 + g*g` contracts into one fma
PT_DEV float sq_rn(float g) { return __fmul_rn(g, g); }
PT_DEV float hg_phase(float cos_theta, float g, float g2) {
    const float denom = 1.f + g2 - 2.f * g * cos_theta;
    return kInvPi * 0.25f * (1.f - g2) / denom * rsqrtf(denom);
}
PT_DEV V3 hg_sample(Rng &rng, float g, float g2) {
    const V2 uv = rng.next2d();
    const float sqr_term = (1.f - g2) / (1.f + g - 2.f * g * uv.x);
    const float cos_theta = (1.f + g2 - sqr_term * sqr_term) / (2.f * g);
    const float sin_theta = sqrtf(fmaxf(0, 1 - cos_theta * cos_theta));
    float sin_phi = 0, cos_phi = 0;
    sincospif(2.f * uv.y, &sin_phi, &cos_phi);
    return V3{cos_phi * sin_theta, sin_phi * sin_theta, cos_theta};
}
PT_DEV float phase_eval(const MediumPOD &m, V3 indir, V3 outdir) {
    switch (m.phase_type) {
    case PHASE_SGGX: // "works the same as IsotropicPhase" upstream (sggx.cuh:30-45)
    case PHASE_ISOTROPIC: return kInvPi * 0.25f;
    case PHASE_RAYLEIGH: return 3.f / 16.f * kInvPi * (1.f + dot(indir, outdir)); // rayleigh.cuh:35-37
    case PHASE_HG: return hg_phase(dot(indir, outdir), m.phase.x, sq_rn(m.phase.x));
    case PHASE_DUO_HG: {
        const float c = dot(indir, outdir);
        // ph1 * w + ph2 * (1 - w): the reference's SASS fuses the first product, rounds the second
        return fmaf(hg_phase(c, m.phase.x, sq_rn(m.phase.x)), m.phase.z, __fmul_rn(hg_phase(c, m.phase.y, sq_rn(m.phase.y)), 1.f - m.phase.z));
    }
    default: return 0.f; // PhaseFunction base (phase.cuh): vacuum never scatters
    }
}
// Medium::scatter (medium.cuh:95-101): samples the phase function, scales the throughput by its weight, returns the
// new world-space direction; `pdf` receives the weight
PT_DEV V3 medium_scatter(const MediumPOD &m, V3 raydir, V4 &throughput, Rng &rng, float &pdf) {
    V3 local{0, 0, 1};
    float weight = 1.f;
    switch (m.phase_type) {
    case PHASE_SGGX:
    case PHASE_ISOTROPIC: {
        float dummy = 0;
        local = sample_sphere_uniform(rng.next2d(), dummy);
        break;
    }
    case PHASE_RAYLEIGH: { // analytic inversion, weight 1 (rayleigh.cuh:39-48)
        const V2 uv = rng.next2d();
        const float u_rs = uv.x * 2.f - 1.f;
        const float cos_theta = cbrtf(2.f * u_rs + sqrtf(4.f * u_rs * u_rs + 1.f));
        const float sin_theta = sqrtf(fmaxf(0, 1 - cos_theta * cos_theta));
        float sin_phi = 0, cos_phi = 0;
        sincospif(2.f * uv.y, &sin_phi, &cos_phi);
        local = V3{cos_phi * sin_theta, sin_phi * sin_theta, cos_theta};
        break;
    }
    case PHASE_HG: local = hg_sample(rng, m.phase.x, sq_rn(m.phase.x)); break;
    case PHASE_DUO_HG: { // MixedHGPhaseFunction::sample: both lobes are sampled, one is kept
        const float g1 = m.phase.x, g2 = m.phase.y, w = m.phase.z;
        const V3 dir1 = hg_sample(rng, g1, sq_rn(g1)), dir2 = hg_sample(rng, g2, sq_rn(g2));
        const float pdf1 = hg_phase(dot(V3{0, 0, 1}, dir1), g1, sq_rn(g1)), pdf2 = hg_phase(dot(V3{0, 0, 1}, dir2), g2, sq_rn(g2));
        const bool use_first = rng.next1d() < w;
        float mis_w = use_first ? pdf1 : pdf2;
        mis_w /= fmaf(w, pdf1, __fmul_rn(1.f - w, pdf2));
        local = use_first ? dir1 : dir2;
        weight = mis_w;
        break;
    }
    default: break;
    }
    throughput *= weight;
    pdf = weight;
    return to_world(raydir, local);
}

} // namespace cptb
