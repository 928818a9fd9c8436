// scene_host.cpp — .cpts container IO, scene validation and the derived
// device-ready arrays (precomputed triangles, packed normals/uvs, pair-node
// BVH re-layout).  Host-only C++.
#include "scene_host.h"

#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>

namespace cptb {

namespace {

struct FileCloser {
    FILE *fp;
    ~FileCloser() {
        if (fp)
            fclose(fp);
    }
};

template <typename T> void read_into(FILE *fp, const CptsSection &s, std::vector<T> &dst, const char *what) {
    if (s.elem_size != sizeof(T))
        throw std::runtime_error(std::string("cpts: section '") + what + "' has unexpected element size");
    dst.resize(s.count);
    if (s.count && fread(dst.data(), sizeof(T), s.count, fp) != s.count)
        throw std::runtime_error(std::string("cpts: truncated section '") + what + "'");
}

void skip_pad(FILE *fp, uint64_t bytes) {
    const uint64_t pad = (16 - bytes % 16) % 16;
    if (pad)
        fseek(fp, long(pad), SEEK_CUR);
}

struct SectionWriter {
    FILE *fp;
    uint32_t n = 0;
    void put(const char *name, const void *data, uint32_t elem, uint64_t count) {
        CptsSection s{};
        strncpy(s.name, name, sizeof(s.name) - 1);
        s.elem_size = elem;
        s.count = count;
        fwrite(&s, sizeof(s), 1, fp);
        const uint64_t bytes = uint64_t(elem) * count;
        if (bytes)
            fwrite(data, 1, bytes, fp);
        static const char zeros[16] = {0};
        const uint64_t pad = (16 - bytes % 16) % 16;
        if (pad)
            fwrite(zeros, 1, pad, fp);
        ++n;
    }
};

// IEEE binary32 -> binary16, round to nearest even (what __float2half does on the device,
// reference: src/core/vec2_half.cuh:44-46 via scene.cu:1213-1220)
uint16_t float_to_half_rn(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u) // inf / nan
        return uint16_t(sign | 0x7C00u | (x > 0x7F800000u ? 0x0200u | ((x >> 13) & 0x3FFu) : 0u));
    if (x >= 0x477FF000u) // rounds to >= 65520 -> inf
        return uint16_t(sign | 0x7C00u);
    if (x < 0x33000001u) // rounds to zero (<= 2^-25)
        return uint16_t(sign);
    int exp = int(x >> 23) - 127;
    uint32_t man = (x & 0x7FFFFFu) | 0x800000u;
    int shift;
    uint32_t half_exp;
    if (exp < -14) { // subnormal half
        shift = 13 + (-14 - exp);
        half_exp = 0;
    } else {
        shift = 13;
        half_exp = uint32_t(exp + 15) << 10;
    }
    uint32_t half_man = man >> shift;
    const uint32_t rem = man & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half_man & 1u)))
        half_man++;
    // for normals half_man still carries the implicit bit (0x400): adding folds a mantissa carry into the exponent
    uint32_t out = (exp < -14) ? half_man : (half_exp + (half_man - 0x400u));
    return uint16_t(sign | out);
}

} // namespace

static void load_cpts_stream(FILE *fp, const std::string &path, HostScene &out);

void load_cpts(const std::string &path, HostScene &out) {
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp)
        throw std::runtime_error("cpts: cannot open '" + path + "'");
    load_cpts_stream(fp, path, out);
}

void load_cpts_memory(const void *image, size_t bytes, HostScene &out) {
    if (!image || bytes < sizeof(CptsHeader))
        throw std::runtime_error("cpts: empty in-memory container");
    FILE *fp = fmemopen(const_cast<void *>(image), bytes, "rb");
    if (!fp)
        throw std::runtime_error("cpts: cannot open the in-memory container");
    load_cpts_stream(fp, "<memory>", out);
}

static void load_cpts_stream(FILE *fp, const std::string &path, HostScene &out) {
    FileCloser closer{fp};
    CptsHeader h{};
    if (fread(&h, sizeof(h), 1, fp) != 1 || memcmp(h.magic, CPTS_MAGIC, 8) != 0)
        throw std::runtime_error("cpts: bad magic in '" + path + "'");
    out = HostScene();
    std::vector<CptsTexDesc> tex_desc;
    std::vector<float> tex_data;
    for (uint32_t i = 0; i < h.n_sections; i++) {
        CptsSection s{};
        if (fread(&s, sizeof(s), 1, fp) != 1)
            throw std::runtime_error("cpts: truncated section table");
        s.name[sizeof(s.name) - 1] = 0;
        const std::string name = s.name;
        const uint64_t bytes = uint64_t(s.elem_size) * s.count;
        auto vec3_of = [&](const char *prefix, std::array<std::vector<F3>, 3> &dst) {
            for (int k = 0; k < 3; k++)
                if (name == std::string(prefix) + char('0' + k)) {
                    read_into(fp, s, dst[k], s.name);
                    return true;
                }
            return false;
        };
        if (name == "config") {
            std::vector<int32_t> v;
            read_into(fp, s, v, "config");
            if (v.size() != CPTS_CFG_COUNT)
                throw std::runtime_error("cpts: config section has wrong length");
            std::copy(v.begin(), v.end(), out.config);
        } else if (name == "camera") {
            std::vector<float> v;
            read_into(fp, s, v, "camera");
            if (v.size() != CPTS_CAM_COUNT)
                throw std::runtime_error("cpts: camera section has wrong length");
            std::copy(v.begin(), v.end(), out.camera);
        } else if (name == "nodes")
            read_into(fp, s, out.nodes, "nodes");
        else if (name == "cache_nodes")
            read_into(fp, s, out.cache_nodes, "cache_nodes");
        else if (name == "obj_idxs")
            read_into(fp, s, out.obj_idxs, "obj_idxs");
        else if (vec3_of("v", out.verts) || vec3_of("n", out.norms)) {
        } else if (name == "uv0")
            read_into(fp, s, out.uvs[0], "uv0");
        else if (name == "uv1")
            read_into(fp, s, out.uvs[1], "uv1");
        else if (name == "uv2")
            read_into(fp, s, out.uvs[2], "uv2");
        else if (name == "objects")
            read_into(fp, s, out.objects, "objects");
        else if (name == "emitter_prims")
            read_into(fp, s, out.emitter_prims, "emitter_prims");
        else if (name == "bsdfs")
            read_into(fp, s, out.bsdfs, "bsdfs");
        else if (name == "emitters")
            read_into(fp, s, out.emitters, "emitters");
        else if (name == "cie_xyz")
            read_into(fp, s, out.cie_xyz, "cie_xyz");
        else if (name == "d65_spd")
            read_into(fp, s, out.d65_spd, "d65_spd");
        else if (name == "media")
            read_into(fp, s, out.media, "media");
        else if (name == "grids")
            read_into(fp, s, out.grids, "grids");
        else if (name == "grid_bricks")
            read_into(fp, s, out.grid_bricks, "grid_bricks");
        else if (name == "grid_data")
            read_into(fp, s, out.grid_data, "grid_data");
        else if (name == "blackbody")
            read_into(fp, s, out.blackbody, "blackbody");
        else if (name == "tex_desc")
            read_into(fp, s, tex_desc, "tex_desc");
        else if (name == "tex_data")
            read_into(fp, s, tex_data, "tex_data");
        else
            fseek(fp, long(bytes), SEEK_CUR); // unknown section: forward compatible
        skip_pad(fp, bytes);
    }
    for (const auto &d : tex_desc) {
        HostTexture t;
        t.width = d.width, t.height = d.height, t.channels = d.channels;
        // dimensions first: a negative width / height / channel count must not reach the size arithmetic below
        if (d.width <= 0 || d.height <= 0 || d.width > 65536 || d.height > 65536 || (d.channels != 2 && d.channels != 4))
            throw std::runtime_error("cpts: texture with invalid dimensions (width / height in [1,65536], 2 or 4 channels)");
        const uint64_t n = uint64_t(d.width) * uint64_t(d.height) * uint64_t(d.channels);
        if (uint64_t(d.offset) > tex_data.size() || n > tex_data.size() - uint64_t(d.offset))
            throw std::runtime_error("cpts: texture data out of range");
        t.texels.assign(tex_data.begin() + d.offset, tex_data.begin() + d.offset + n);
        out.textures.push_back(std::move(t));
    }
    validate_scene(out);
}

void save_cpts(const HostScene &sc, const std::string &path) {
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp)
        throw std::runtime_error("cpts: cannot create '" + path + "'");
    FileCloser closer{fp};
    CptsHeader h{};
    memcpy(h.magic, CPTS_MAGIC, 8);
    fwrite(&h, sizeof(h), 1, fp);
    SectionWriter w{fp};
    w.put("config", sc.config, 4, CPTS_CFG_COUNT);
    w.put("camera", sc.camera, 4, CPTS_CAM_COUNT);
    w.put("nodes", sc.nodes.data(), 16, sc.nodes.size());
    w.put("cache_nodes", sc.cache_nodes.data(), 16, sc.cache_nodes.size());
    w.put("obj_idxs", sc.obj_idxs.data(), 4, sc.obj_idxs.size());
    const char *vn[3] = {"v0", "v1", "v2"}, *nn[3] = {"n0", "n1", "n2"}, *tn[3] = {"uv0", "uv1", "uv2"};
    for (int k = 0; k < 3; k++) {
        w.put(vn[k], sc.verts[k].data(), 12, sc.verts[k].size());
        w.put(nn[k], sc.norms[k].data(), 12, sc.norms[k].size());
        w.put(tn[k], sc.uvs[k].data(), 8, sc.uvs[k].size());
    }
    w.put("objects", sc.objects.data(), 16, sc.objects.size());
    w.put("emitter_prims", sc.emitter_prims.data(), 4, sc.emitter_prims.size());
    w.put("bsdfs", sc.bsdfs.data(), sizeof(CptsBsdf), sc.bsdfs.size());
    w.put("emitters", sc.emitters.data(), sizeof(CptsEmitter), sc.emitters.size());
    if (!sc.cie_xyz.empty()) {
        w.put("cie_xyz", sc.cie_xyz.data(), 16, sc.cie_xyz.size());
        w.put("d65_spd", sc.d65_spd.data(), 4, sc.d65_spd.size());
    }
    if (!sc.media.empty())
        w.put("media", sc.media.data(), sizeof(CptsMedium), sc.media.size());
    if (!sc.grids.empty()) {
        w.put("grids", sc.grids.data(), sizeof(CptsGrid), sc.grids.size());
        w.put("grid_bricks", sc.grid_bricks.data(), 4, sc.grid_bricks.size());
        w.put("grid_data", sc.grid_data.data(), 4, sc.grid_data.size());
        w.put("blackbody", sc.blackbody.data(), 16, sc.blackbody.size());
    }
    if (!sc.textures.empty()) {
        std::vector<CptsTexDesc> desc;
        std::vector<float> data;
        for (const auto &t : sc.textures) {
            CptsTexDesc d{};
            d.width = t.width, d.height = t.height, d.channels = t.channels;
            d.offset = data.size();
            data.insert(data.end(), t.texels.begin(), t.texels.end());
            desc.push_back(d);
        }
        w.put("tex_desc", desc.data(), sizeof(CptsTexDesc), desc.size());
        w.put("tex_data", data.data(), 4, data.size());
    }
    h.n_sections = w.n;
    fseek(fp, 0, SEEK_SET);
    fwrite(&h, sizeof(h), 1, fp);
}

void validate_scene(const HostScene &sc) {
    auto fail = [](const std::string &m) { throw std::runtime_error("scene: " + m); };
    const size_t P = sc.verts[0].size();
    if (sc.width() <= 0 || sc.height() <= 0)
        fail("film size must be positive");
    if (uint64_t(sc.width()) * sc.height() > (1u << 24))
        fail("film larger than the 24-bit pixel index limit of the reference (wavefront_pt.cu:49)");
    if (P == 0)
        fail("no primitives");
    for (int k = 0; k < 3; k++)
        if (sc.verts[k].size() != P || sc.norms[k].size() != P || sc.uvs[k].size() != P)
            fail("vertex / normal / uv arrays differ in length");
    if (sc.obj_idxs.size() != P)
        fail("obj_idxs length differs from primitive count");
    if (sc.nodes.empty() || sc.nodes.size() % 2)
        fail("BVH required (path_tracer.cu:73-75)");
    if (sc.cache_nodes.size() > 255)
        fail("more than 255 cached nodes");
    for (size_t i = 0; i < sc.cache_nodes.size(); i++) {
        // {half2 x, half2 y, half2 z, gmem_index << 8 | skip}: the cached walk advances by `skip` on a miss (bvh.cuh:182-251), so a
        // zero skip would spin forever on the device, and the frontier index it hands over must be a node
        const uint32_t w = sc.cache_nodes[i].w, skip = w & 0xffu, gmem = w >> 8;
        if (skip == 0)
            fail("cached BVH node with a zero skip offset");
        if (gmem >= sc.nodes.size() / 2)
            fail("cached BVH node points past the node array");
    }
    for (const auto &t : sc.textures)
        if (t.width <= 0 || t.height <= 0 || (t.channels != 2 && t.channels != 4) ||
            t.texels.size() != size_t(t.width) * size_t(t.height) * size_t(t.channels))
            fail("texture with invalid dimensions or texel count");
    auto tex_ok = [&](int32_t idx) { return idx >= -1 && idx < int32_t(sc.textures.size()); };
    for (const auto &b : sc.bsdfs)
        for (int k = 0; k < 5; k++)
            if (!tex_ok(b.tex[k]))
                fail("BSDF references an unknown texture");
    for (const auto &e : sc.emitters) {
        if (!tex_ok(e.tex))
            fail("emitter references an unknown texture");
        if (e.obj_ref < -1)
            fail("emitter with an invalid object reference (must be -1 or an object index)");
    }
    if (sc.config[CPTS_CFG_ENVMAP_ID] < 0 || size_t(sc.config[CPTS_CFG_ENVMAP_ID]) >= std::max<size_t>(sc.emitters.size(), 1))
        fail("environment-map emitter index out of range");
    if (sc.objects.empty())
        fail("no objects");
    if (sc.bsdfs.empty() || sc.bsdfs.size() > 48)
        fail("BSDF count must be in [1,48] (scene.cu:28-29)");
    if (sc.emitters.empty() || sc.emitters.size() > 9)
        fail("emitter table must hold the null emitter plus at most 8 emitters");
    for (const auto &b : sc.bsdfs)
        if (b.type == CPTS_BSDF_DISPERSION && (sc.cie_xyz.size() != 471 || sc.d65_spd.size() != 531))
            fail("dispersion BSDF needs the CIE XYZ (471) and D65 (531) spectral tables (sections cie_xyz / d65_spd); "
                 "the XML loader does not carry them");
    for (int32_t w : sc.obj_idxs) {
        if (w < 0)
            fail("sphere primitives are not supported (reference default build is TRIANGLE_ONLY, defines.cuh:26)");
        if (size_t(w & 0x000fffff) >= sc.objects.size())
            fail("primitive references an unknown object");
        const int medium = (w >> 20) & 0xff; // bits 20-27: medium bound by the object, bit 30: cullable (scene.cu:903-932)
        if (medium != 0 && size_t(medium) >= sc.media.size())
            fail("primitive references an unknown medium");
    }
    if (sc.media.size() > 256)
        fail("at most 255 media (8-bit medium index)");
    if (sc.config[CPTS_CFG_CAM_VOL] < 0 || (sc.config[CPTS_CFG_CAM_VOL] > 0 && size_t(sc.config[CPTS_CFG_CAM_VOL]) >= sc.media.size()))
        fail("camera medium index out of range");
    for (size_t i = 1; i < sc.media.size(); i++) {
        if (sc.media[i].type == CPTS_MEDIUM_GRID) {
            if (sc.media[i].grid < 0 || size_t(sc.media[i].grid) >= sc.grids.size() || sc.grids[sc.media[i].grid].medium != int(i))
                fail("grid medium without a matching entry in the 'grids' section");
        } else if (sc.media[i].type != CPTS_MEDIUM_HOMOGENEOUS)
            fail("unknown medium type");
        if (sc.media[i].phase_type < CPTS_PHASE_ISOTROPIC || sc.media[i].phase_type > CPTS_PHASE_SGGX)
            fail("unknown phase function type");
    }
    for (const auto &g : sc.grids) { // every brick table and every brick pool must lie inside its section
        if (g.ch[CPTS_GRID_DENSITY].ncomp != 1)
            fail("grid medium without a density channel");
        if (!(g.majorant > 0.f) || !(g.scale > 0.f))
            fail("grid medium with a non-positive majorant or density scale");
        for (int c = 0; c < 3; c++) {
            const CptsGridChannel &ch = g.ch[c];
            if (ch.ncomp == 0)
                continue;
            if (ch.ncomp != (c == CPTS_GRID_ALBEDO ? 3 : 1))
                fail("grid channel with an unexpected component count");
            if (ch.nb[0] <= 0 || ch.nb[1] <= 0 || ch.nb[2] <= 0 || ch.nb[0] > 4096 || ch.nb[1] > 4096 || ch.nb[2] > 4096)
                fail("grid channel with an invalid brick count");
            const uint64_t nbricks = uint64_t(ch.nb[0]) * ch.nb[1] * ch.nb[2];
            if (ch.brick_offset + nbricks > sc.grid_bricks.size())
                fail("grid brick table outside 'grid_bricks'");
            for (uint64_t b = 0; b < nbricks; b++) {
                const int32_t id = sc.grid_bricks[ch.brick_offset + b];
                if (id < -1 || (id >= 0 && ch.data_offset + (uint64_t(id) + 1) * 512 * ch.ncomp > sc.grid_data.size()))
                    fail("grid brick outside 'grid_data'");
            }
        }
    }
    if (!sc.grids.empty() && sc.blackbody.empty())
        fail("grid media need the 'blackbody' emission table (the reference loads it for every scene, vol_grid.cu:304-326)");
    for (const auto &o : sc.objects) {
        const int bsdf = o.bsdf_emitter >> 16;
        int em = o.bsdf_emitter & 0xFFFF;
        em = (em ^ 0x8000) - 0x8000;
        if (bsdf < 0 || size_t(bsdf) >= sc.bsdfs.size())
            fail("object references an unknown BSDF");
        if (em < 0 || size_t(em) >= sc.emitters.size())
            fail("object references an unknown emitter");
        if (em > 0 && (o.prim_num <= 0 || size_t(o.prim_offset) + o.prim_num > sc.emitter_prims.size()))
            fail("emitter object primitive range outside emitter_prims");
    }
    for (int32_t p : sc.emitter_prims)
        if (p < 0 || size_t(p) >= P)
            fail("emitter_prims entry out of range");
    for (const auto &e : sc.emitters)
        if (e.obj_ref >= int(sc.objects.size()))
            fail("emitter bound to an unknown object");
    // leaf ranges
    const int N = sc.num_nodes();
    for (int k = 0; k < N; k++) {
        int32_t base, cnt;
        memcpy(&base, &sc.nodes[2 * k].w, 4);
        memcpy(&cnt, &sc.nodes[2 * k + 1].w, 4);
        if (cnt > 0 && (base < 0 || size_t(base) + cnt > P))
            fail("BVH leaf primitive range out of bounds");
        if (cnt > 1024)
            fail("BVH leaf with more than 1024 primitives (the reference's builders stop at 32)");
        if (cnt < 0 && k - cnt > N)
            fail("BVH skip offset runs past the node array");
        if (cnt == 0)
            fail("BVH node with zero primitive count");
    }
}

void build_triangles(const HostScene &sc, std::vector<F4> &tris) {
    const size_t P = sc.verts[0].size();
    tris.resize(3 * P);
    for (size_t i = 0; i < P; i++) {
        const F3 a = sc.verts[0][i], b = sc.verts[1][i], c = sc.verts[2][i];
        // plain fp32 products and differences, no contraction (host code of the reference is built without FMA)
        const volatile float d1x = b.x - a.x, d1y = b.y - a.y, d1z = b.z - a.z;
        const volatile float d2x = c.x - a.x, d2y = c.y - a.y, d2z = c.z - a.z;
        const volatile float p0 = d1y * d2z, p1 = d1z * d2y, p2 = d2x * d1z, p3 = d1x * d2z, p4 = d1x * d2y, p5 = d2x * d1y;
        const float a20 = p0 - p1, a21 = p2 - p3, a22 = p4 - p5;
        tris[3 * i + 0] = F4{a.x, a.y, a.z, a20};
        tris[3 * i + 1] = F4{d1x, d1y, d1z, a21};
        tris[3 * i + 2] = F4{d2x, d2y, d2z, a22};
    }
}

void build_normals_uvs(const HostScene &sc, std::vector<F4> &norms, std::vector<U4> &uvs) {
    const size_t P = sc.verts[0].size();
    norms.resize(3 * P);
    uvs.resize(P);
    for (size_t i = 0; i < P; i++) {
        for (int k = 0; k < 3; k++) {
            const F3 n = sc.norms[k][i];
            norms[3 * i + k] = F4{n.x, n.y, n.z, 0.f};
        }
        auto pack = [](F2 uv) { return uint32_t(float_to_half_rn(uv.x)) | (uint32_t(float_to_half_rn(uv.y)) << 16); };
        uvs[i] = U4{pack(sc.uvs[0][i]), pack(sc.uvs[1][i]), pack(sc.uvs[2][i]), 0u};
    }
}

// ------------------------------------------------------------------------------------------------
// pair-node layout
// ------------------------------------------------------------------------------------------------
void build_pair_layout(const HostScene &sc, PairLayout &out, int refine_above) {
    const int N = sc.num_nodes();
    out = PairLayout();
    if (refine_above > 0) {
        out.perm.resize(size_t(sc.num_prims()));
        for (size_t i = 0; i < out.perm.size(); i++)
            out.perm[i] = int32_t(i);
    }
    auto node_cnt = [&](int k) {
        int32_t c;
        memcpy(&c, &sc.nodes[2 * k + 1].w, 4);
        return c;
    };
    auto node_base = [&](int k) {
        int32_t b;
        memcpy(&b, &sc.nodes[2 * k].w, 4);
        return b;
    };
    auto subtree = [&](int k) { return node_cnt(k) < 0 ? -node_cnt(k) : 1; };
    auto as_float = [](int32_t v) {
        float f;
        memcpy(&f, &v, 4);
        return f;
    };
    constexpr int kMaxLeaf = (1 << (31 - 26)) - 1; // 5-bit count
    constexpr int kBaseMask = (1 << 26) - 1;

    // emits a child record {min,ref}{max,end}
    struct Child {
        F4 lo, hi;
    };
    auto new_node = [&]() {
        const int idx = int(out.wnodes.size() / 4);
        out.wnodes.resize(out.wnodes.size() + 4, F4{0, 0, 0, 0});
        out.children_nest.push_back(1);
        return idx;
    };
    const Child empty_child{F4{1e30f, 1e30f, 1e30f, as_float(~0)}, F4{-1e30f, -1e30f, -1e30f, as_float(0)}};

    // leaf child (possibly a chain of pair nodes when the primitive count exceeds the 5-bit field;
    // every link repeats the leaf's own box so the set of tested primitives is unchanged)
    std::function<Child(int, int, int)> leaf_child = [&](int k, int base, int cnt) -> Child {
        const F4 lo = sc.nodes[2 * k], hi = sc.nodes[2 * k + 1];
        Child c;
        if (cnt <= kMaxLeaf) {
            if (base > kBaseMask)
                throw std::runtime_error("pair layout: primitive index exceeds 26 bits");
            const int32_t ref = ~int32_t(base | (cnt << 26));
            c.lo = F4{lo.x, lo.y, lo.z, as_float(ref)};
            c.hi = F4{hi.x, hi.y, hi.z, as_float(k + 1)};
            return c;
        }
        out.oversize_leaves++;
        const int idx = new_node();
        const Child a = leaf_child(k, base, kMaxLeaf), b = leaf_child(k, base + kMaxLeaf, cnt - kMaxLeaf);
        out.wnodes[4 * idx + 0] = a.lo, out.wnodes[4 * idx + 1] = a.hi;
        out.wnodes[4 * idx + 2] = b.lo, out.wnodes[4 * idx + 3] = b.hi;
        c.lo = F4{lo.x, lo.y, lo.z, as_float(idx)};
        c.hi = F4{hi.x, hi.y, hi.z, as_float(k + 1)};
        return c;
    };

    // ---- leaf refinement.  A reference leaf of many primitives (the reference builds leaves of up to 16 / 32) costs one
    // triangle test per primitive for every ray that enters its box.  Its primitives get a median-split tree of their own:
    // boxes = triangle bounds padded like the reference pads them (AABB_EPS); a ray that hits a triangle passes through that
    // triangle's bounds, so the closest hit is unchanged and only misses are skipped.  Primitives move inside the leaf's own
    // range only; ties keep resolving by the reference's index (perm).
    struct Tri {
        float lo[3], hi[3], ctr[3];
        int32_t orig;
        bool inside;
    };
    bool any_outside = false;
    auto tri_of = [&](int32_t prim, const F4 &blo, const F4 &bhi) {
        Tri t;
        t.orig = prim;
        const float leaf_lo[3] = {blo.x, blo.y, blo.z}, leaf_hi[3] = {bhi.x, bhi.y, bhi.z};
        bool inside = true;
        for (int a = 0; a < 3; a++) {
            const float p0 = (&sc.verts[0][prim].x)[a], p1 = (&sc.verts[1][prim].x)[a], p2 = (&sc.verts[2][prim].x)[a];
            t.lo[a] = std::min(p0, std::min(p1, p2)) - 1e-5f, t.hi[a] = std::max(p0, std::max(p1, p2)) + 1e-5f;
            inside = inside && t.lo[a] >= leaf_lo[a] && t.hi[a] <= leaf_hi[a];
        }
        // A reference that sticks out of its leaf's box (SBVH: a triangle clipped to the node it was split into) cannot be
        // given bounds of its own: the reference tests every primitive of a leaf it enters, wherever the hit lies - also
        // BEFORE the ray enters the leaf's box - and with its start-at-the-cached-frontier quirk such a hit may be the only
        // way a triangle is found at all (cornell C0).  Distance culling would drop those hits; such references stay in a
        // plain group that is tested on every entry of the leaf (refined_leaf_child), only the others are refined.
        if (!inside)
            any_outside = true;
        t.inside = inside;
        for (int a = 0; a < 3; a++)
            t.ctr[a] = 0.5f * (t.lo[a] + t.hi[a]);
        return t;
    };
    std::function<Child(int, std::vector<Tri> &, int, int, int, int)> refine = [&](int k, std::vector<Tri> &tris, int lo, int hi, int base,
                                                                                    int depth) -> Child {
        // child covering tris[lo, hi) = primitives [base + lo, base + hi) after the permutation
        float blo[3] = {1e30f, 1e30f, 1e30f}, bhi[3] = {-1e30f, -1e30f, -1e30f};
        for (int i = lo; i < hi; i++)
            for (int a = 0; a < 3; a++)
                blo[a] = std::min(blo[a], tris[i].lo[a]), bhi[a] = std::max(bhi[a], tris[i].hi[a]);
        out.max_depth = std::max(out.max_depth, depth);
        Child c;
        const int cnt = hi - lo;
        if (cnt <= 2) {
            const int32_t ref = ~int32_t((base + lo) | (cnt << 26));
            c.lo = F4{blo[0], blo[1], blo[2], as_float(ref)};
            c.hi = F4{bhi[0], bhi[1], bhi[2], as_float(k + 1)};
            return c;
        }
        // split: exact SAH over the centroid-sorted order of each axis (a leaf holds at most a few dozen primitives);
        // falls back to the median of the widest axis when no split beats keeping the range together
        int mid = lo + cnt / 2;
        {
            float best_cost = 1e30f;
            int best_axis = -1, best_pos = -1;
            std::vector<float> right_area(size_t(cnt) + 1);
            for (int a = 0; a < 3; a++) {
                std::sort(tris.begin() + lo, tris.begin() + hi,
                          [a](const Tri &x, const Tri &y) { return x.ctr[a] < y.ctr[a] || (x.ctr[a] == y.ctr[a] && x.orig < y.orig); });
                float rl[3] = {1e30f, 1e30f, 1e30f}, rh[3] = {-1e30f, -1e30f, -1e30f};
                auto area_of = [](const float *l, const float *h) {
                    const float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
                    return 2.f * (dx * dy + dy * dz + dx * dz);
                };
                for (int i = cnt - 1; i > 0; i--) {
                    for (int q = 0; q < 3; q++)
                        rl[q] = std::min(rl[q], tris[lo + i].lo[q]), rh[q] = std::max(rh[q], tris[lo + i].hi[q]);
                    right_area[size_t(i)] = area_of(rl, rh);
                }
                float ll[3] = {1e30f, 1e30f, 1e30f}, lh[3] = {-1e30f, -1e30f, -1e30f};
                for (int i = 1; i < cnt; i++) {
                    for (int q = 0; q < 3; q++)
                        ll[q] = std::min(ll[q], tris[lo + i - 1].lo[q]), lh[q] = std::max(lh[q], tris[lo + i - 1].hi[q]);
                    const float cost = area_of(ll, lh) * float(i) + right_area[size_t(i)] * float(cnt - i);
                    if (cost < best_cost)
                        best_cost = cost, best_axis = a, best_pos = i;
                }
            }
            if (best_axis >= 0) {
                const int a = best_axis;
                if (a != 2)
                    std::sort(tris.begin() + lo, tris.begin() + hi,
                              [a](const Tri &x, const Tri &y) { return x.ctr[a] < y.ctr[a] || (x.ctr[a] == y.ctr[a] && x.orig < y.orig); });
                mid = lo + best_pos;
            }
        }
        const int idx = new_node();
        const Child a = refine(k, tris, lo, mid, base, depth + 1), b = refine(k, tris, mid, hi, base, depth + 1);
        out.wnodes[4 * idx + 0] = a.lo, out.wnodes[4 * idx + 1] = a.hi;
        out.wnodes[4 * idx + 2] = b.lo, out.wnodes[4 * idx + 3] = b.hi;
        c.lo = F4{blo[0], blo[1], blo[2], as_float(idx)};
        c.hi = F4{bhi[0], bhi[1], bhi[2], as_float(k + 1)};
        return c;
    };
    auto refined_leaf_child = [&](int k, int base, int cnt, int depth) -> Child {
        if (base + cnt - 1 > kBaseMask)
            throw std::runtime_error("pair layout: primitive index exceeds 26 bits");
        const F4 lo = sc.nodes[2 * k], hi = sc.nodes[2 * k + 1];
        std::vector<Tri> tris;
        tris.reserve(size_t(cnt));
        any_outside = false;
        for (int i = 0; i < cnt; i++)
            tris.push_back(tri_of(base + i, lo, hi));
        // references that stick out of the leaf's box first (reference order kept among them): they form a plain group
        // with the LEAF's box, which the traversal visits first whenever it enters the leaf (same entry distance as the
        // leaf, never farther than the refined part) - every one of them is tested exactly when the reference tests them.
        const int n_out = int(std::stable_partition(tris.begin(), tris.end(), [](const Tri &t) { return !t.inside; }) - tris.begin());
        const int n_in = cnt - n_out;
        if (n_in <= 2 || n_out > kMaxLeaf)
            return leaf_child(k, base, cnt);
        Child c;
        if (n_out == 0) {
            c = refine(k, tris, 0, cnt, base, depth);
        } else {
            const int idx = new_node();
            const int32_t ref = ~int32_t(base | (n_out << 26));
            const Child in = refine(k, tris, n_out, cnt, base, depth + 1);
            out.wnodes[4 * idx + 0] = F4{lo.x, lo.y, lo.z, as_float(ref)};
            out.wnodes[4 * idx + 1] = F4{hi.x, hi.y, hi.z, as_float(k + 1)};
            out.wnodes[4 * idx + 2] = in.lo, out.wnodes[4 * idx + 3] = in.hi;
            c.lo = F4{lo.x, lo.y, lo.z, as_float(idx)};
            c.hi = F4{hi.x, hi.y, hi.z, as_float(k + 1)};
            out.max_depth = std::max(out.max_depth, depth + 1);
        }
        for (int i = 0; i < cnt; i++)
            out.perm[size_t(base) + i] = tris[i].orig;
        // the child keeps the LEAF's own box: the reference enters the leaf on that test, so do we
        c.lo.x = lo.x, c.lo.y = lo.y, c.lo.z = lo.z;
        c.hi.x = hi.x, c.hi.y = hi.y, c.hi.z = hi.z;
        out.refined_leaves++;
        return c;
    };

    auto contained = [&](int child, int parent) {
        const F4 cl = sc.nodes[2 * child], ch = sc.nodes[2 * child + 1], pl = sc.nodes[2 * parent], ph = sc.nodes[2 * parent + 1];
        const float over = std::max(std::max(std::max(pl.x - cl.x, pl.y - cl.y), pl.z - cl.z),
                                    std::max(std::max(ch.x - ph.x, ch.y - ph.y), ch.z - ph.z));
        if (over > 0.f)
            out.max_overshoot = std::max(out.max_overshoot, over);
        return !(over > 0.f);
    };

    for (int k = 0; k < N; k++) { // clipped references (see PairLayout::clipped_refs)
        if (node_cnt(k) <= 0)
            continue;
        const F4 lo = sc.nodes[2 * k], hi = sc.nodes[2 * k + 1];
        for (int i = node_base(k); i < node_base(k) + node_cnt(k); i++)
            out.clipped_refs += !tri_of(i, lo, hi).inside;
    }
    // explicit stack: (reference node k, slot to fill = pair node index * 2 + child id, depth)
    struct Item {
        int k, slot, depth;
    };
    std::vector<Item> todo;
    const int root_pair = new_node();
    out.root = root_pair;
    out.wnodes[4 * root_pair + 2] = empty_child.lo;
    out.wnodes[4 * root_pair + 3] = empty_child.hi;
    todo.push_back(Item{0, root_pair * 2, 1});
    while (!todo.empty()) {
        const Item it = todo.back();
        todo.pop_back();
        out.max_depth = std::max(out.max_depth, it.depth);
        const int pair_idx = it.slot >> 1, which = it.slot & 1;
        Child c;
        if (node_cnt(it.k) > 0) {
            if (refine_above > 0 && node_cnt(it.k) > refine_above)
                c = refined_leaf_child(it.k, node_base(it.k), node_cnt(it.k), it.depth);
            else
                c = leaf_child(it.k, node_base(it.k), node_cnt(it.k));
        } else {
            const int idx = new_node();
            const F4 lo = sc.nodes[2 * it.k], hi = sc.nodes[2 * it.k + 1];
            c.lo = F4{lo.x, lo.y, lo.z, as_float(idx)};
            c.hi = F4{hi.x, hi.y, hi.z, as_float(it.k + subtree(it.k))};
            const int left = it.k + 1, right = left + subtree(left);
            if (right >= N || right >= it.k + subtree(it.k))
                throw std::runtime_error("pair layout: malformed skip offsets");
            if (!contained(left, it.k))
                out.nesting_violations++, out.children_nest[size_t(idx)] = 0;
            if (!contained(right, it.k))
                out.nesting_violations++, out.children_nest[size_t(idx)] = 0;
            // push right first so the left subtree is laid out next (DFS order => parent/child locality)
            todo.push_back(Item{right, idx * 2 + 1, it.depth + 1});
            todo.push_back(Item{left, idx * 2 + 0, it.depth + 1});
        }
        out.wnodes[4 * pair_idx + 2 * which + 0] = c.lo;
        out.wnodes[4 * pair_idx + 2 * which + 1] = c.hi;
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// wide layout: collapse of the pair tree (see scene_host.h)
// ---------------------------------------------------------------------------------------------------------------------
void build_wide_layout(const PairLayout &pl, int width, WideLayout &out) {
    if (width != 2 && width != 4 && width != 8)
        throw std::runtime_error("wide layout: width must be 2, 4 or 8");
    out = WideLayout();
    out.width = width;
    auto as_int = [](float f) {
        int32_t v;
        memcpy(&v, &f, 4);
        return v;
    };
    auto as_float = [](int32_t v) {
        float f;
        memcpy(&f, &v, 4);
        return f;
    };
    struct Child {
        F4 lo, hi;
    };
    auto area = [](const Child &c) {
        const float dx = c.hi.x - c.lo.x, dy = c.hi.y - c.lo.y, dz = c.hi.z - c.lo.z;
        return (dx < 0 || dy < 0 || dz < 0) ? -1.f : 2.f * (dx * dy + dy * dz + dx * dz);
    };
    auto is_empty = [&](const Child &c) { return as_int(c.lo.w) == ~0 && as_int(c.hi.w) == 0; };
    auto pair_children = [&](int pair_idx, Child &a, Child &b) {
        a = Child{pl.wnodes[4 * size_t(pair_idx) + 0], pl.wnodes[4 * size_t(pair_idx) + 1]};
        b = Child{pl.wnodes[4 * size_t(pair_idx) + 2], pl.wnodes[4 * size_t(pair_idx) + 3]};
    };
    const Child empty{F4{1e30f, 1e30f, 1e30f, as_float(~0)}, F4{-1e30f, -1e30f, -1e30f, as_float(0)}};

    // breadth-first construction: queue of (pair node to collapse, wide node index, depth, stack need of the path so far)
    struct Item {
        int pair_idx, wide_idx, depth, stack, stack_all;
    };
    std::vector<Item> queue;
    auto new_node = [&]() {
        const int idx = out.num_nodes++;
        out.nodes.resize(out.nodes.size() + 2 * size_t(width), F4{0, 0, 0, 0});
        return idx;
    };
    queue.push_back(Item{pl.root, new_node(), 1, 0, 0});
    size_t filled = 0;
    for (size_t qi = 0; qi < queue.size(); qi++) {
        const Item it = queue[qi];
        out.max_depth = std::max(out.max_depth, it.depth);
        std::vector<Child> kids;
        {
            Child a, b;
            pair_children(it.pair_idx, a, b);
            if (!is_empty(a))
                kids.push_back(a);
            if (!is_empty(b))
                kids.push_back(b);
        }
        // greedy collapse (Wald et al. 2008): open the inner child with the largest surface area until the node is full
        while (int(kids.size()) < width) {
            int best = -1;
            float best_area = -1.f;
            for (int i = 0; i < int(kids.size()); i++) {
                const int ref = as_int(kids[size_t(i)].lo.w);
                // a node whose children stick out of its box keeps its own box test (PairLayout::children_nest)
                if (ref >= 0 && (pl.children_nest.empty() || pl.children_nest[size_t(ref)]) && area(kids[size_t(i)]) > best_area)
                    best = i, best_area = area(kids[size_t(i)]);
            }
            if (best < 0)
                break;
            Child a, b;
            pair_children(as_int(kids[size_t(best)].lo.w), a, b);
            kids.erase(kids.begin() + best);
            if (!is_empty(a))
                kids.push_back(a);
            if (!is_empty(b))
                kids.push_back(b);
        }
        int inner = 0;
        for (const Child &c : kids)
            inner += as_int(c.lo.w) >= 0;
        const int need = it.stack + std::max(inner - 1, 0);
        out.max_stack = std::max(out.max_stack, need);
        // a per-thread traversal that also pushes leaf children: all children but the one it enters
        const int need_all = it.stack_all + std::max(int(kids.size()) - 1, 0);
        out.max_stack_all = std::max(out.max_stack_all, need_all);
        filled += kids.size();
        for (int i = 0; i < width; i++) {
            Child c = i < int(kids.size()) ? kids[size_t(i)] : empty;
            const int ref = as_int(c.lo.w);
            if (i < int(kids.size())) {
                if (ref >= 0) {
                    const int w = new_node();
                    queue.push_back(Item{ref, w, it.depth + 1, need, need_all});
                    c.lo.w = as_float(w);
                    out.inner_slots++;
                } else {
                    out.leaf_slots++;
                }
            }
            out.nodes[(size_t(it.wide_idx) * width + size_t(i)) * 2 + 0] = c.lo;
            out.nodes[(size_t(it.wide_idx) * width + size_t(i)) * 2 + 1] = c.hi;
        }
    }
    out.avg_children = out.num_nodes ? double(filled) / out.num_nodes : 0.0;
}

} // namespace cptb
