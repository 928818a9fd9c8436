// scene_host.h — host-side flat scene (what the reference's `Scene` hands to a
// renderer constructor; reference: src/core/scene.cuh:49-107) and the helpers
// that turn it into device-ready arrays.  Pure C++ (no CUDA calls) so that the
// loader, the .cpts reader/writer and the BVH re-layout can be unit-tested on
// a CPU-only box.
#pragma once
#include "cpts_format.h"

#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace cptb {

struct F3 {
    float x, y, z;
};
struct F2 {
    float x, y;
};
struct F4 {
    float x, y, z, w;
};
struct U4 {
    uint32_t x, y, z, w;
};

struct HostTexture {
    int width = 0, height = 0, channels = 4;
    std::vector<float> texels;
};

struct HostScene {
    int32_t config[CPTS_CFG_COUNT] = {0};
    float camera[CPTS_CAM_COUNT] = {0};
    std::vector<F4> nodes;       // 2 per node
    std::vector<U4> cache_nodes; // fp16 near-root nodes
    std::vector<int32_t> obj_idxs;
    std::array<std::vector<F3>, 3> verts, norms;
    std::array<std::vector<F2>, 3> uvs;
    std::vector<CptsObject> objects;
    std::vector<int32_t> emitter_prims;
    std::vector<CptsBsdf> bsdfs;
    std::vector<CptsEmitter> emitters; // [0] = null emitter
    std::vector<HostTexture> textures;
    std::vector<CptsMedium> media; // [0] = vacuum; empty when the scene has no participating media
    // grid media (cpts_format.h "grids"): brick tables + brick pools of every channel, and the emission colour table
    std::vector<CptsGrid> grids;
    std::vector<int32_t> grid_bricks;
    std::vector<float> grid_data;
    std::vector<F4> blackbody;
    std::vector<F4> cie_xyz;    // 471 samples, 360..830 nm (only present when a dispersion BSDF needs it)
    std::vector<float> d65_spd; // 531 samples, 300..830 nm

    int width() const { return config[CPTS_CFG_WIDTH]; }
    int height() const { return config[CPTS_CFG_HEIGHT]; }
    int num_nodes() const { return int(nodes.size() / 2); }
    int num_prims() const { return int(verts[0].size()); }
};

// ---- .cpts container (include/cpts_format.h) ----
void load_cpts(const std::string &path, HostScene &out); // throws std::runtime_error
// the same container handed over in memory (reference-side adapter: integration/b200_wavefront.cuh); the image is only read
void load_cpts_memory(const void *image, size_t bytes, HostScene &out);
void save_cpts(const HostScene &scene, const std::string &path);

// ---- XML scene loader (scene format v1.2; reference behaviour: src/impl/scene.cu:937-1192) ----
void load_xml(const std::string &path, HostScene &out); // throws std::runtime_error

// consistency checks shared by both loaders; throws on malformed scenes
void validate_scene(const HostScene &scene);

// ---- device-ready derived arrays ----
// 48 B precomputed-adjoint triangles {v0,a20}{e1,a21}{e2,a22} (reference: src/core/aos.cuh:134-164)
void build_triangles(const HostScene &scene, std::vector<F4> &tris);
// 3 x {n.xyz,0} per primitive and packed half uvs (reference: aos.cuh:248-262, vec2_half.cuh:160-170)
void build_normals_uvs(const HostScene &scene, std::vector<F4> &norms, std::vector<U4> &uvs);

// Pair-node re-layout for ordered traversal (see traverse.cuh).  Returns the root reference.
struct PairLayout {
    std::vector<F4> wnodes; // 4 per inner node
    int root = 0;
    int max_depth = 0;
    int nesting_violations = 0; // child boxes that are not contained in their parent's box
    float max_overshoot = 0.f;  // by how much at most (the reference's SBVH pads clipped child boxes by AABB_EPS = 1e-5)
    int oversize_leaves = 0;
    // leaf refinement: reference leaves with more than `refine_above` primitives get a small tree of their own inside the
    // pair layout; their primitives are re-ordered INSIDE the leaf's range, perm[new index] = reference index (empty =
    // identity).  The reference tree, its leaf ranges and the cached nodes are untouched.
    std::vector<int32_t> perm;
    int refined_leaves = 0;
    // per pair node: both child boxes lie inside the node's own box (false where the reference's SBVH pads a clipped child
    // beyond its parent).  The wide collapse only merges a node into its parent's record where this holds - elsewhere the
    // node's own box test decides, as in the reference, whether its children are looked at at all.
    std::vector<uint8_t> children_nest;
    // primitive references that stick out of their leaf's box (SBVH: a triangle clipped to the node it was split into).  The
    // reference finds hits of such a triangle only through leaves whose box the ray enters while no closer hit is known, so
    // its result depends on the order in which leaves are tested; traversal modes that test leaves in batches (traversal=4)
    // are not used on such scenes.
    int clipped_refs = 0;
};
void build_pair_layout(const HostScene &scene, PairLayout &out, int refine_above = 0);

// Wide re-layout for the warp-cooperative traversal (traverse_coop.cuh): the pair tree collapsed to `width` (4 or 8)
// children per node.  One node = `width` child records of 32 B, {min.xyz, ref}{max.xyz, end} - the SAME fp32 boxes, leaf
// references and DFS end indices as the pair layout, so that lane i of a ray's lane group loads child i with one 256-bit
// load and tests it with the reference's slab arithmetic.  Collapsing only removes box tests of inner nodes whose
// children's boxes are tested instead (children nest inside their parents), the set of leaves a ray can reach is the one
// of the pair tree.  Nodes are numbered breadth-first, so that the first `n` records are the top of the tree: they form
// the "treelet slab" the traversal kernels stage in shared memory with one bulk copy (TMA) per CTA.
struct WideLayout {
    int width = 8;
    std::vector<F4> nodes; // 2 * width per node
    int num_nodes = 0;
    int max_depth = 0;     // wide levels
    int max_stack = 0;     // worst-case number of deferred inner children along one root-to-leaf path
    int max_stack_all = 0; // the same counting leaf children too (per-thread traversal that pushes every child it does not enter)
    double avg_children = 0; // filled slots per node
    int leaf_slots = 0, inner_slots = 0;
};
void build_wide_layout(const PairLayout &pairs, int width, WideLayout &out);

} // namespace cptb
