// vecmath.cuh — small POD vector types for the sm_100a path-tracing kernels.
//
// Parity note: the reference is built with
// --use_fast_math, so its results depend on the exact sequence of fp32
// operations (which products are fused, which divisions are approximate).
// The operators below are deliberately written so that, once inlined, every
// scalar expression has the same operand order and the same explicit fmaf /
// __frcp_rn placement as the reference's vector algebra
// (reference: src/core/vec3.cuh:66-310, src/core/vec4.cuh:60-271,
//  src/core/vec2.cuh).  Do not "simplify" them.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define PT_DEV __device__ __forceinline__

namespace cptb {

struct V2 {
    float x, y;
};
struct V3 {
    float x, y, z;
};
struct V4 {
    float x, y, z, w;
};

PT_DEV V2 mk2(float x, float y) { return V2{x, y}; }
PT_DEV V3 mk3(float x, float y, float z) { return V3{x, y, z}; }
PT_DEV V3 mk3(float v) { return V3{v, v, v}; }
PT_DEV V4 mk4(float x, float y, float z, float w = 1.f) { return V4{x, y, z, w}; }
PT_DEV V4 splat4(float v) { return V4{v, v, v, v}; }            // reference Vec4(float)
PT_DEV V4 splat4(float v, float a) { return V4{v, v, v, a}; }   // reference Vec4(float, alpha)
PT_DEV V4 from_f4(float4 f) { return V4{f.x, f.y, f.z, f.w}; }
PT_DEV float4 to_f4(V4 v) { return make_float4(v.x, v.y, v.z, v.w); }
PT_DEV V3 xyz(float4 f) { return V3{f.x, f.y, f.z}; }
PT_DEV V3 xyz(V4 f) { return V3{f.x, f.y, f.z}; }

// Explicitly rounded multiply-add shapes.  nvcc decides per expression which product of `a*b + c*d` it
// fuses; the reference's build (verified in its PTX: intersect_triangle, NormalAoSArray::eval,
// AreaSource::sample, SO3::transposed_rotate) fuses the FIRST product and chains the later ones, so these
// helpers pin that order instead of leaving it to the optimiser.
PT_DEV float mad2(float a, float b, float c, float d) { return fmaf(a, b, __fmul_rn(c, d)); }               // a*b + c*d
PT_DEV float mad3(float a, float b, float c, float d, float e, float f) { return fmaf(e, f, mad2(a, b, c, d)); } // + e*f

// ---------------------------------------------------------------- V2
PT_DEV V2 operator*(V2 a, float b) { return V2{a.x * b, a.y * b}; }

// ---------------------------------------------------------------- V3
PT_DEV V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
PT_DEV V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
PT_DEV V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
PT_DEV V3 operator*(V3 a, float b) { return V3{a.x * b, a.y * b, a.z * b}; }
PT_DEV V3 operator*(float b, V3 a) { return V3{a.x * b, a.y * b, a.z * b}; }
PT_DEV V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
PT_DEV V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
PT_DEV V3 operator-(V3 a, float b) { return V3{a.x - b, a.y - b, a.z - b}; }
PT_DEV V3 operator+(V3 a, float b) { return V3{a.x + b, a.y + b, a.z + b}; }
PT_DEV V3 &operator+=(V3 &a, V3 b) {
    a.x += b.x, a.y += b.y, a.z += b.z;
    return a;
}
PT_DEV V3 &operator*=(V3 &a, float b) {
    a.x *= b, a.y *= b, a.z *= b;
    return a;
}

PT_DEV float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
PT_DEV V3 cross(V3 a, V3 b) {
    return V3{fmaf(a.y, b.z, -a.z * b.y), fmaf(a.z, b.x, -a.x * b.z), fmaf(a.x, b.y, -a.y * b.x)};
}
PT_DEV float length2(V3 a) { return fmaf(a.x, a.x, fmaf(a.y, a.y, a.z * a.z)); }
PT_DEV float length(V3 a) { return norm3df(a.x, a.y, a.z); }
PT_DEV V3 normalized(V3 a) { return a * rnorm3df(a.x, a.y, a.z); }
// o + d * t with one fma per lane
PT_DEV V3 advance(V3 o, V3 d, float t) { return V3{fmaf(d.x, t, o.x), fmaf(d.y, t, o.y), fmaf(d.z, t, o.z)}; }
PT_DEV V3 rcp3(V3 a) { return V3{__frcp_rn(a.x), __frcp_rn(a.y), __frcp_rn(a.z)}; }
PT_DEV V3 face_forward(V3 a) { return a.z > 0 ? a : V3{-a.x, -a.y, -a.z}; }
// 2 (r.n) n - r
PT_DEV V3 reflection(V3 ray, V3 n, float dot_prod) { return 2.f * dot_prod * n - ray; }
PT_DEV V3 reflection(V3 ray, V3 n) { return reflection(ray, n, dot(ray, n)); }

// ---------------------------------------------------------------- V4
PT_DEV V4 operator+(V4 a, V4 b) { return V4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
PT_DEV V4 operator-(V4 a, V4 b) { return V4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
PT_DEV V4 operator-(V4 a) { return V4{-a.x, -a.y, -a.z, -a.w}; }
PT_DEV V4 operator+(V4 a, float b) { return V4{a.x + b, a.y + b, a.z + b, a.w + b}; }
PT_DEV V4 operator*(V4 a, float b) { return V4{a.x * b, a.y * b, a.z * b, a.w * b}; }
PT_DEV V4 operator*(float b, V4 a) { return V4{a.x * b, a.y * b, a.z * b, a.w * b}; }
PT_DEV V4 operator*(V4 a, V4 b) { return V4{a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
PT_DEV V4 operator/(V4 a, V4 b) { return V4{a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w}; }
PT_DEV V4 &operator+=(V4 &a, V4 b) {
    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    return a;
}
PT_DEV V4 &operator*=(V4 &a, V4 b) {
    a.x *= b.x, a.y *= b.y, a.z *= b.z, a.w *= b.w;
    return a;
}
PT_DEV V4 &operator*=(V4 &a, float b) {
    a.x *= b, a.y *= b, a.z *= b, a.w *= b;
    return a;
}
PT_DEV float max3(V4 a) { return max(a.x, max(a.y, a.z)); }
PT_DEV bool bad_rgb(V4 a) {
    return isnan(a.x) || isnan(a.y) || isnan(a.z) || isinf(a.x) || isinf(a.y) || isinf(a.z);
}
PT_DEV V4 exp_xyz(V4 a) { return V4{expf(a.x), expf(a.y), expf(a.z), 1.f}; }
PT_DEV V4 max4(V4 a, V4 b) { return V4{max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w)}; }

// arithmetic blend used all over the reference (core/cuda_utils.cuh:104-106):
// a * p + b * (1 - p) with p in {0,1}
PT_DEV V4 blend(V4 a, V4 b, bool p) { return a * float(p) + b * float(1 - int(p)); }
PT_DEV V3 blend(V3 a, V3 b, bool p) { return a * float(p) + b * float(1 - int(p)); }

// ---------------------------------------------------------------- 3x3
struct M3 {
    V3 r0, r1, r2;
};
PT_DEV V3 rotate(const M3 &m, V3 v) { return V3{dot(m.r0, v), dot(m.r1, v), dot(m.r2, v)}; }
PT_DEV V3 rotate_t(const M3 &m, V3 v) {
    return V3{mad3(m.r0.x, v.x, m.r1.x, v.y, m.r2.x, v.z), mad3(m.r0.y, v.x, m.r1.y, v.y, m.r2.y, v.z),
              mad3(m.r0.z, v.x, m.r1.z, v.y, m.r2.z, v.z)};
}
// a*s + b*t + c*u per component, first product fused (barycentric interpolation of vertex attributes)
PT_DEV V3 lerp3(V3 a, float s, V3 b, float t, V3 c, float u) {
    return V3{mad3(a.x, s, b.x, t, c.x, u), mad3(a.y, s, b.y, t, c.y, u), mad3(a.z, s, b.z, t, c.z, u)};
}
PT_DEV V3 col(const M3 &m, int i) {
    return i == 0 ? V3{m.r0.x, m.r1.x, m.r2.x} : (i == 1 ? V3{m.r0.y, m.r1.y, m.r2.y} : V3{m.r0.z, m.r1.z, m.r2.z});
}

// Rotation taking (0,0,1) to `to` (l2w) or `to` to (0,0,1) (w2l): Rodrigues form
// R = diag(to.z) + (1 - to.z) a a^T + [a_unnormalised]_x with a = (-+to.y, +-to.x, 0), skipped when |to.z| ~ 1
// (reference: src/core/so3.cuh:228-239, 201-217).  Written out per entry with the roundings the
// reference's SASS performs: a.z is exactly 0, so the third row/column reduce to +-to.x, +-to.y; the
// diagonal is to.z + round(a_i^2 * s); the xy entries are round(a_x a_y * s).
PT_DEV M3 frame_from_z(V3 to, bool l2w) {
    M3 R{V3{to.z, 0, 0}, V3{0, to.z, 0}, V3{0, 0, to.z}};
    if (fabsf(to.z) < 1.f - 1e-5f) {
        const V3 a = normalized(V3{l2w ? -to.y : to.y, l2w ? to.x : -to.x, 0.f});
        const float s = 1.f - to.z;
        const float oxx = __fmul_rn(a.x, a.x), oyy = __fmul_rn(a.y, a.y), oxy = __fmul_rn(a.y, a.x);
        const float xy = __fmul_rn(oxy, s);
        const float sx = l2w ? to.x : -to.x, sy = l2w ? to.y : -to.y;
        R.r0 = V3{__fadd_rn(to.z, __fmul_rn(oxx, s)), xy, sx};
        R.r1 = V3{xy, __fadd_rn(to.z, __fmul_rn(oyy, s)), sy};
        R.r2 = V3{-sx, -sy, to.z};
    }
    return R;
}
// local (z-up) direction -> world frame whose z axis is `n`
PT_DEV V3 to_world(V3 n, V3 local) { return rotate(frame_from_z(n, true), local); }

} // namespace cptb
