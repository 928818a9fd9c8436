/*
 * cpts_format.h — the flat scene container (".cpts") shared by the C-ABI
 * library, the CPU oracle and the reference harness.
 *
 * A .cpts file is the POD image of what the reference's `Scene` object hands
 * to a renderer constructor (reference: src/core/scene.cuh:49-107 public
 * fields, consumed by src/renderer/tracer_base.cuh:62-69 and
 * src/pt_impl/path_tracer.cu:49-99): linear BVH nodes, cached fp16 nodes,
 * primitive arrays, primitive->object words, compacted object table, emitter
 * primitive map, plus POD restatements of the device-heap BSDF / Emitter
 * objects (reference: src/core/virtual_funcs.cuh:28-118) and the camera
 * (src/core/camera_model.cuh:40-48).
 *
 * Layout (little endian):
 *   CptsHeader
 *   n_sections x { CptsSection, payload (count*elem_size bytes, padded to 16) }
 *
 * Plain C so that C, C++, CUDA and ctypes/numpy readers agree.
 */
#ifndef CPTS_FORMAT_H
#define CPTS_FORMAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPTS_MAGIC "CPTSCN01"

typedef struct CptsHeader {
    char magic[8];
    uint32_t n_sections;
    uint32_t reserved;
} CptsHeader;

typedef struct CptsSection {
    char name[24];      /* zero padded */
    uint32_t elem_size; /* bytes per element */
    uint32_t reserved;
    uint64_t count;     /* number of elements */
} CptsSection;

/* indices into the "config" section (int32[32]) */
enum {
    CPTS_CFG_WIDTH = 0,
    CPTS_CFG_HEIGHT = 1,
    CPTS_CFG_SPP = 2,
    CPTS_CFG_MAX_DEPTH = 3,
    CPTS_CFG_MAX_DIFFUSE = 4,
    CPTS_CFG_MAX_SPECULAR = 5,
    CPTS_CFG_MAX_TRANSMIT = 6,
    CPTS_CFG_GAMMA = 7,
    CPTS_CFG_RENDERER = 8, /* reference RendererType: 0 = pt megakernel, 1 = wfpt */
    CPTS_CFG_NUM_BSDFS = 9,
    CPTS_CFG_NUM_EMITTERS = 10, /* real emitters; table holds num+1 (index 0 = null emitter) */
    CPTS_CFG_ENVMAP_ID = 11,
    CPTS_CFG_NUM_OBJECTS = 12,
    CPTS_CFG_NUM_PRIMS = 13,
    CPTS_CFG_NUM_NODES = 14,
    CPTS_CFG_NUM_CACHE = 15,
    CPTS_CFG_CACHE_LEVEL = 16,
    CPTS_CFG_MAX_NODE_NUM = 17,
    CPTS_CFG_USE_SBVH = 18,
    CPTS_CFG_MAX_VOLUME = 19, /* MaxDepthParams::max_volume (volumetric path only) */
    CPTS_CFG_NUM_MEDIA = 20,  /* real media; the "media" section holds num+1 (index 0 = vacuum) */
    CPTS_CFG_CAM_VOL = 21,    /* medium the camera sits in (Scene::cam_vol_id) */
    CPTS_CFG_COUNT = 32
};

/* "camera" section: float[18], reference DeviceCamera (camera_model.cuh:40-48) */
enum {
    CPTS_CAM_R = 0,          /* 9 floats, row major */
    CPTS_CAM_T = 9,          /* 3 floats */
    CPTS_CAM_INV_FOCAL = 12,
    CPTS_CAM_HW = 13,
    CPTS_CAM_HH = 14,
    CPTS_CAM_SIGN_X = 15,
    CPTS_CAM_APERTURE = 16,
    CPTS_CAM_FOCAL_DIST = 17,
    CPTS_CAM_COUNT = 18
};

/* BSDF types, same numbering as the reference's BSDFType (core/enums.cuh:55-65) */
enum {
    CPTS_BSDF_LAMBERTIAN = 0,
    CPTS_BSDF_SPECULAR = 1,
    CPTS_BSDF_TRANSLUCENT = 2,
    CPTS_BSDF_PLASTIC = 3,
    CPTS_BSDF_PLASTIC_FORWARD = 4,
    CPTS_BSDF_GGX_CONDUCTOR = 5,
    CPTS_BSDF_DISPERSION = 6,
    CPTS_BSDF_FORWARD = 7
};

/* "bsdfs" section, 128 B per entry. `extras` meaning by type:
 *   GGX conductor : [0..2] eta_t, [3..5] k, [6] roughness_x, [7] roughness_y (raw, pre roughness_to_alpha)
 *   plastic(-fwd) : [0] ior, [1] trans_scaler, [2] thickness, [3] penetrable (0/1)
 *   dispersion    : [0] cauchy A, [1] cauchy B
 * tex[] = {diffuse, specular, glossy, normal, roughness} indices into the
 * "tex_desc" section or -1. */
typedef struct CptsBsdf {
    int32_t type;
    int32_t flags;
    float k_d[4];
    float k_s[4];
    float k_g[4];
    float extras[8];
    int32_t tex[5];
    int32_t pad[5];
} CptsBsdf;

enum {
    CPTS_EMIT_NULL = 0,
    CPTS_EMIT_POINT = 1,
    CPTS_EMIT_AREA = 2,
    CPTS_EMIT_AREA_SPOT = 3,
    CPTS_EMIT_ENVMAP = 4
};

/* "emitters" section, 64 B per entry (entry 0 is the null emitter). */
typedef struct CptsEmitter {
    int32_t type;
    int32_t obj_ref;   /* bound object index, -1 if none */
    int32_t is_sphere;
    int32_t tex;       /* index into "tex_desc" or -1 */
    float Le[4];
    float pos[3];
    float cos_val;
    float scale;
    float azimuth;     /* radians */
    float zenith;      /* radians */
    int32_t pad;
} CptsEmitter;

/* "objects" section, 16 B per entry: reference CompactedObjInfo (object.cuh:31-65) */
typedef struct CptsObject {
    int32_t bsdf_emitter;
    int32_t prim_offset;
    int32_t prim_num;
    float inv_area;
} CptsObject;

/* "media" section, 64 B per entry (optional; volumetric path).  Entry 0 is the vacuum.  The per-primitive object word
 * ("obj_idxs") carries the medium an object bounds in bits 20-27 and the "cullable" (alpha) flag in bit 30
 * (src/impl/scene.cu:903-932). */
enum { CPTS_MEDIUM_HOMOGENEOUS = 0, CPTS_MEDIUM_GRID = 1 };
enum { CPTS_PHASE_NULL = 0, CPTS_PHASE_ISOTROPIC = 1, CPTS_PHASE_HG = 2, CPTS_PHASE_DUO_HG = 3, CPTS_PHASE_RAYLEIGH = 4, CPTS_PHASE_SGGX = 5 };
typedef struct CptsMedium {
    int32_t type;       /* CPTS_MEDIUM_* */
    int32_t phase_type; /* CPTS_PHASE_* */
    float sigma_a[4];   /* already multiplied by the density scale (vol_homo.cu:26-32) */
    float sigma_s[4];
    float phase[4];     /* hg: g; duo-hg: g1, g2, weight */
    int32_t grid;       /* grid media: index into the "grids" section; 0 otherwise */
    int32_t pad;
} CptsMedium;

/* Grid media (reference: GridVolumeMedium, src/volume/grid.cuh:34-121, src/impl/vol_grid.cu:82-211).  A grid medium keeps
 * its constant albedo in CptsMedium::sigma_a and (emission scale, temperature scale) in CptsMedium::sigma_s[0..1]
 * (MediumInfo::MediumParams, core/dynamic_bsdf.cuh:55-60); everything else is one CptsGrid in the "grids" section.
 *
 * A grid has up to three channels (density: float, albedo: 3 floats, temperature: float), each the image of one NanoVDB
 * grid under nearest-voxel lookup (`getValue(Coord)`), stored as a two-level brick grid instead of the NanoVDB tree:
 *   "grid_bricks"  int32   one entry per 8x8x8 brick of the channel's index bounding box, x-major
 *                          ((bx * nb[1] + by) * nb[2] + bz), value = brick number in the channel's pool or -1 (every
 *                          voxel of the brick holds the background value)
 *   "grid_data"    float   brick pools: brick b of a channel starts at data_offset + b * 512 * ncomp; inside a brick a
 *                          voxel sits at (((i & 7) << 6) | ((j & 7) << 3) | (k & 7)) * ncomp  (NanoVDB leaf order)
 * Lookups outside the index bounding box return the background.  `inv_mat` / `vec` are the float copies of the grid's
 * affine map (NanoVDB Map::mInvMatF / mVecF): index = inv_mat * (world - vec), evaluated as in math::matMult. */
typedef struct CptsGridChannel {
    float inv_mat[9];
    float vec[3];
    int32_t bmin[3];     /* index bounding box minimum (a multiple of 8 per axis) */
    int32_t nb[3];       /* bricks per axis */
    float background[3];
    int32_t ncomp;       /* 0: channel absent, 1 or 3 */
    uint64_t brick_offset; /* into "grid_bricks" (entries) */
    uint64_t data_offset;  /* into "grid_data" (floats) */
} CptsGridChannel;       /* 104 B */

enum { CPTS_GRID_DENSITY = 0, CPTS_GRID_ALBEDO = 1, CPTS_GRID_TEMPERATURE = 2 };
typedef struct CptsGrid {
    int32_t medium;       /* index of the medium that owns this grid */
    int32_t pad;
    float scale;          /* density scale */
    float temp_scale;
    float em_scale;
    float majorant;       /* maximum of the density grid (Tree::extrema), unscaled */
    float avg_density;    /* mean of the positive density voxels (residual tracking; unused by the served estimators) */
    float pad1;
    float aabb_min[3];    /* world bounding box of the density grid, as floats (GridVolumeMedium::grid_aabb) */
    float aabb_max[3];
    CptsGridChannel ch[3];
} CptsGrid;               /* 368 B */

/* "blackbody" section: float[4] x N, the emission colour table indexed by normalised temperature (the reference reads
 * scene/data/blackbody.bin into a linear-filtered, clamped, normalised 1D texture: vol_grid.cu:304-326). */

/* "tex_desc" section: one per texture; texels live in "tex_data" (float) at `offset` floats. */
typedef struct CptsTexDesc {
    int32_t width;
    int32_t height;
    int32_t channels; /* 4 or 2 */
    int32_t pad;
    uint64_t offset;
} CptsTexDesc;

#ifdef __cplusplus
}
#endif
#endif /* CPTS_FORMAT_H */
