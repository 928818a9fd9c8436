/*
 * cpts_flatten.cuh — REFERENCE-SIDE integration code (compiled inside the reference's build, never into libcpt_b200.so).
 *
 * Flattens the reference's host-side `Scene` (src/core/scene.cuh:49-107, as built by its own XML loader, BVH / SBVH builders
 * and texture loader) into the flat .cpts container of include/cpts_format.h - the form in which a scene crosses the C ABI
 * (cpt_b200_scene_load_memory / cpt_b200_scene_load).  Used by
 *     integration/b200_wavefront.cuh   the TracerBase adapter: flattens into a memory stream, no file involved
 *     oracle/ref_harness/ref_dump.cu   `ref_dump dump`: the same bytes into a file (golden scenes of the tests)
 * It only reads the reference's public members; texels and spectral tables are read back from the reference's own device
 * textures (whatever its stb_image decoded is what the kernels sample).
 */
#pragma once
#include <cstdio>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include <cmath>
#include <fstream>

#include "core/scene.cuh"
#include "core/xyz.cuh"
#include "cpts_format.h"

#include <nanovdb/io/IO.h>

extern CPT_GPU_CONST Emitter *c_emitter[9];
extern CPT_GPU_CONST BSDF *c_material[48];

// reads the spectral tables back out of the reference's own device textures at texel centres
// (linear filtering at a texel centre returns the texel): CIE 1931 XYZ 360-830 nm and D65 300-830 nm,
// uploaded by ColorSpaceXYZ::init() (src/impl/xyz.cu:808-830).  Run-time extraction, not a source copy.
__global__ void read_spectral_tables(float4 *cie, float *d65) {
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i < ColorSpaceXYZ::CIE_samples)
        cie[i] = tex1D<float4>(XYZ.CIE, (float(i) + 0.5f) / float(ColorSpaceXYZ::CIE_samples));
    if (i < ColorSpaceXYZ::D65_samples)
        d65[i] = tex1D<float>(XYZ.D65, (float(i) + 0.5f) / float(ColorSpaceXYZ::D65_samples));
}

namespace {

struct Writer {
    FILE *fp = nullptr;
    uint32_t n = 0;
    bool in_memory = false; // open_memstream(): the caller patches the section count into the finished buffer
    explicit Writer(const std::string &path) : Writer(fopen(path.c_str(), "wb")) {}
    // any seekable stream: a file, or an open_memstream() buffer (the adapter's in-memory hand-over)
    explicit Writer(FILE *stream) : fp(stream) {
        if (!fp)
            throw std::runtime_error("cpts_flatten: cannot open the output stream");
        CptsHeader h{};
        memcpy(h.magic, CPTS_MAGIC, 8);
        fwrite(&h, sizeof(h), 1, fp);
    }
    void section(const char *name, const void *data, uint32_t elem, uint64_t count) {
        CptsSection s{};
        strncpy(s.name, name, sizeof(s.name) - 1);
        s.elem_size = elem;
        s.count = count;
        fwrite(&s, sizeof(s), 1, fp);
        size_t bytes = size_t(elem) * count;
        if (bytes)
            fwrite(data, 1, bytes, fp);
        static const char zeros[16] = {0};
        size_t pad = (16 - bytes % 16) % 16;
        if (pad)
            fwrite(zeros, 1, pad, fp);
        ++n;
    }
    void close() {
        CptsHeader h{};
        memcpy(h.magic, CPTS_MAGIC, 8);
        h.n_sections = n;
        if (!in_memory) { // a memory stream cannot be patched through the FILE: its length follows the last write position
            fseek(fp, 0, SEEK_SET);
            fwrite(&h, sizeof(h), 1, fp);
        }
        fclose(fp);
    }
};

// colour attribute grammar of the scene format ("#RRGGBB" | "r, g, b" | "v")
void parse_color(const char *txt, float out[3]) {
    std::string v = txt ? txt : "0";
    if (!v.empty() && v[0] == '#') {
        unsigned c = 0;
        sscanf(v.c_str() + 1, "%x", &c);
        out[0] = float((c >> 16) & 0xFF) / 255.0f;
        out[1] = float((c >> 8) & 0xFF) / 255.0f;
        out[2] = float(c & 0xFF) / 255.0f;
        return;
    }
    for (auto &ch : v)
        if (ch == ',')
            ch = ' ';
    std::stringstream ss(v);
    std::vector<float> vals;
    float f;
    while (ss >> f)
        vals.push_back(f);
    if (vals.size() >= 3) {
        out[0] = vals[0], out[1] = vals[1], out[2] = vals[2];
    } else {
        out[0] = out[1] = out[2] = vals.empty() ? 0.f : vals[0];
    }
}

const tinyxml2::XMLElement *child_named(const tinyxml2::XMLElement *parent, const char *tag,
                                        std::initializer_list<const char *> names) {
    for (auto e = parent->FirstChildElement(tag); e; e = e->NextSiblingElement(tag)) {
        const char *n = e->Attribute("name");
        if (!n)
            continue;
        for (auto want : names)
            if (strcmp(n, want) == 0)
                return e;
    }
    return nullptr;
}

int lobe_flags_for(int type) {
    switch (type) {
    case BSDFType::Lambertian: return BSDF_DIFFUSE | BSDF_REFLECT;
    case BSDFType::Specular: return BSDF_SPECULAR | BSDF_REFLECT;
    case BSDFType::Translucent: return BSDF_SPECULAR | BSDF_TRANSMIT;
    case BSDFType::Plastic: return BSDF_SPECULAR | BSDF_DIFFUSE | BSDF_REFLECT;
    case BSDFType::PlasticForward: return BSDF_SPECULAR | BSDF_TRANSMIT | BSDF_REFLECT;
    case BSDFType::GGXConductor: return BSDF_GLOSSY | BSDF_REFLECT;
    case BSDFType::Dispersion: return BSDF_DIFFUSE | BSDF_TRANSMIT;
    default: return BSDF_NONE;
    }
}

// reads the texels of one of the reference's pitched 2-D textures (src/impl/textures.cu:134-165) back to the
// host through the resource descriptor of the texture object; returns the index in `descs`
int grab_texture(cudaTextureObject_t obj, std::vector<CptsTexDesc> &descs, std::vector<float> &data) {
    cudaResourceDesc rd{};
    CUDA_CHECK_RETURN(cudaGetTextureObjectResourceDesc(&rd, obj));
    if (rd.resType != cudaResourceTypePitch2D) {
        fprintf(stderr, "unexpected texture resource type %d\n", int(rd.resType));
        exit(3);
    }
    const int ch = (rd.res.pitch2D.desc.y ? 1 : 0) + (rd.res.pitch2D.desc.z ? 1 : 0) + (rd.res.pitch2D.desc.w ? 1 : 0) + 1;
    CptsTexDesc d{};
    d.width = int(rd.res.pitch2D.width), d.height = int(rd.res.pitch2D.height), d.channels = ch;
    d.offset = data.size();
    const size_t row = size_t(d.width) * ch * sizeof(float);
    data.resize(data.size() + size_t(d.width) * d.height * ch);
    CUDA_CHECK_RETURN(cudaMemcpy2D(data.data() + d.offset, row, rd.res.pitch2D.devPtr, rd.res.pitch2D.pitchInBytes, row,
                                   d.height, cudaMemcpyDeviceToHost));
    descs.push_back(d);
    return int(descs.size()) - 1;
}

struct Writer;
void dump_grids(const Scene &scene, const tinyxml2::XMLElement *root, const std::string &xml_path, std::vector<CptsMedium> &media,
                Writer &w);

void dump_scene(const Scene &scene, const std::string &xml_path, Writer &w) {

    int32_t cfg[CPTS_CFG_COUNT] = {0};
    cfg[CPTS_CFG_WIDTH] = scene.config.width;
    cfg[CPTS_CFG_HEIGHT] = scene.config.height;
    cfg[CPTS_CFG_SPP] = scene.config.spp;
    cfg[CPTS_CFG_MAX_DEPTH] = scene.config.md.max_depth;
    cfg[CPTS_CFG_MAX_DIFFUSE] = scene.config.md.max_diffuse;
    cfg[CPTS_CFG_MAX_SPECULAR] = scene.config.md.max_specular;
    cfg[CPTS_CFG_MAX_TRANSMIT] = scene.config.md.max_tranmit;
    cfg[CPTS_CFG_GAMMA] = scene.config.gamma_correction;
    cfg[CPTS_CFG_RENDERER] = int(scene.rdr_type);
    cfg[CPTS_CFG_NUM_BSDFS] = scene.num_bsdfs;
    cfg[CPTS_CFG_NUM_EMITTERS] = scene.num_emitters;
    cfg[CPTS_CFG_ENVMAP_ID] = scene.envmap_id;
    cfg[CPTS_CFG_NUM_OBJECTS] = int(scene.objects.size());
    cfg[CPTS_CFG_NUM_PRIMS] = scene.num_prims;
    cfg[CPTS_CFG_NUM_NODES] = int(scene.nodes.size() >> 1);
    cfg[CPTS_CFG_NUM_CACHE] = int(scene.cache_nodes.size());
    cfg[CPTS_CFG_CACHE_LEVEL] = scene.config.bvh.cache_level;
    cfg[CPTS_CFG_MAX_NODE_NUM] = scene.config.bvh.max_node_num;
    cfg[CPTS_CFG_USE_SBVH] = scene.config.bvh.use_sbvh;
    cfg[CPTS_CFG_MAX_VOLUME] = scene.config.md.max_volume;
    cfg[CPTS_CFG_NUM_MEDIA] = scene.num_medium;
    cfg[CPTS_CFG_CAM_VOL] = scene.cam_vol_id;
    w.section("config", cfg, 4, CPTS_CFG_COUNT);

    static_assert(sizeof(DeviceCamera) == CPTS_CAM_COUNT * sizeof(float), "camera POD size");
    w.section("camera", scene.cam, 4, CPTS_CAM_COUNT);

    w.section("nodes", scene.nodes.data(), 16, scene.nodes.size());
    static_assert(sizeof(CompactNode) == 16, "compact node size");
    w.section("cache_nodes", scene.cache_nodes.data(), 16, scene.cache_nodes.size());
    w.section("obj_idxs", scene.obj_idxs.data(), 4, scene.obj_idxs.size());

    const char *vnames[3] = {"v0", "v1", "v2"}, *nnames[3] = {"n0", "n1", "n2"},
               *tnames[3] = {"uv0", "uv1", "uv2"};
    for (int k = 0; k < 3; k++) {
        static_assert(sizeof(Vec3) == 12 && sizeof(Vec2) == 8, "vec sizes");
        w.section(vnames[k], scene.verts_list[k].data(), 12, scene.verts_list[k].size());
        w.section(nnames[k], scene.norms_list[k].data(), 12, scene.norms_list[k].size());
        w.section(tnames[k], scene.uvs_list[k].data(), 8, scene.uvs_list[k].size());
    }

    std::vector<CompactedObjInfo> objs;
    for (const auto &o : scene.objects)
        objs.push_back(o.export_gpu());
    static_assert(sizeof(CompactedObjInfo) == sizeof(CptsObject), "object POD size");
    w.section("objects", objs.data(), 16, objs.size());
    w.section("emitter_prims", scene.emitter_prims.data(), 4, scene.emitter_prims.size());

    // ---- BSDF / emitter POD tables: host-side info + a re-read of the XML for
    // the fields the reference does not keep on the host.
    tinyxml2::XMLDocument doc;
    doc.LoadFile(xml_path.c_str());
    const tinyxml2::XMLElement *root = doc.FirstChildElement("scene");

    // texture object tables of the reference: 5 x num_bsdfs, diffuse | specular | glossy | normal | roughness
    // (src/core/textures.cuh:129-139,152-166)
    std::vector<CptsTexDesc> tex_descs;
    std::vector<float> tex_data;
    std::vector<cudaTextureObject_t> tex_table(size_t(scene.num_bsdfs) * 5, 0);
    if (scene.num_bsdfs > 0)
        CUDA_CHECK_RETURN(cudaMemcpy(tex_table.data(), scene.textures.diff_tex,
                                     tex_table.size() * sizeof(cudaTextureObject_t), cudaMemcpyDeviceToHost));
    size_t num_bsdf_tex4 = 0;
    for (size_t i = 0; i < size_t(scene.num_bsdfs) * 4; i++)
        num_bsdf_tex4 += tex_table[i] != 0;

    std::vector<CptsBsdf> bsdfs(scene.num_bsdfs);
    const tinyxml2::XMLElement *be = root->FirstChildElement("brdf");
    for (int i = 0; i < scene.num_bsdfs; i++, be = be->NextSiblingElement("brdf")) {
        const BSDFInfo &info = scene.bsdf_infos[i];
        CptsBsdf b{};
        b.type = info.type;
        b.flags = lobe_flags_for(info.type);
        for (int k = 0; k < 5; k++) {
            const cudaTextureObject_t t = tex_table[size_t(k) * scene.num_bsdfs + i];
            b.tex[k] = t ? grab_texture(t, tex_descs, tex_data) : -1;
        }
        for (int k = 0; k < 4; k++) {
            b.k_d[k] = info.bsdf.k_d[k];
            b.k_s[k] = info.bsdf.k_s[k];
            b.k_g[k] = info.bsdf.k_g[k];
        }
        if (info.type == BSDFType::GGXConductor) {
            const Vec3 eta = METAL_ETA_TS[info.bsdf.mtype], kk = METAL_KS[info.bsdf.mtype];
            for (int k = 0; k < 3; k++) {
                b.extras[k] = eta[k];
                b.extras[3 + k] = kk[k];
            }
            b.extras[6] = info.bsdf.roughness_x();
            b.extras[7] = info.bsdf.roughness_y();
        } else if (info.type == BSDFType::Plastic || info.type == BSDFType::PlasticForward) {
            // the loader resets sigma_a (k_g) to (0,0,0,1) before creating the
            // device object (scene.cu:293), BSDFInfo still holds the XML value
            b.k_g[0] = b.k_g[1] = b.k_g[2] = 0.f, b.k_g[3] = 1.f;
            b.extras[0] = info.bsdf.ior();
            b.extras[1] = info.bsdf.trans_scaler();
            b.extras[2] = info.bsdf.thickness();
            bool pen = false;
            if (auto e = child_named(be, "bool", {"penetrable"}))
                e->QueryBoolAttribute("value", &pen);
            b.extras[3] = pen ? 1.f : 0.f;
        } else if (info.type == BSDFType::Dispersion) {
            const Vec2 p = DISPERSION_PARAMS[info.bsdf.mtype];
            b.extras[0] = p.x();
            b.extras[1] = p.y();
        }
        bsdfs[i] = b;
    }
    w.section("bsdfs", bsdfs.data(), sizeof(CptsBsdf), bsdfs.size());

    std::vector<CptsEmitter> ems(scene.num_emitters + 1);
    memset(ems.data(), 0, ems.size() * sizeof(CptsEmitter));
    ems[0].type = CPTS_EMIT_NULL;
    ems[0].obj_ref = -1;
    ems[0].tex = -1;
    ems[0].Le[3] = 1.f; // Emitter(Vec4(0,0,0)) => w = 1
    size_t next_emitter_tex = num_bsdf_tex4; // host_tex_4d = [BSDF float4 textures..., emitter textures...] (scene.cu:204-232,406-409)
    const tinyxml2::XMLElement *ee = root->FirstChildElement("emitter");
    for (int i = 1; i <= scene.num_emitters; i++, ee = ee->NextSiblingElement("emitter")) {
        CptsEmitter e{};
        e.tex = -1;
        e.obj_ref = -1;
        std::string type = ee->Attribute("type");
        float emission[3] = {0, 0, 0}, scaler[3] = {1, 1, 1};
        if (auto c = child_named(ee, "rgb", {"emission"}))
            parse_color(c->Attribute("value"), emission);
        if (auto c = child_named(ee, "rgb", {"scaler"}))
            parse_color(c->Attribute("value"), scaler);
        for (int k = 0; k < 3; k++)
            e.Le[k] = emission[k] * scaler[k];
        e.Le[3] = 1.f * scaler[0];
        if (const tinyxml2::XMLElement *r = ee->FirstChildElement("ref")) // only the first <ref> is looked at (scene.cu:395-397)
            if (r->Attribute("type") && !strcmp(r->Attribute("type"), "texture")) {
                if (next_emitter_tex >= scene.host_tex_4d.size()) {
                    fprintf(stderr, "emitter texture bookkeeping mismatch\n");
                    exit(3);
                }
                e.tex = grab_texture(scene.host_tex_4d[next_emitter_tex++].object(), tex_descs, tex_data);
            }
        for (size_t o = 0; o < scene.objects.size(); o++)
            if (scene.objects[o].emitter_id == i)
                e.obj_ref = int(o); // last bound shape wins, as in the loader's map
        const tinyxml2::XMLElement *bind = child_named(ee, "string", {"bind_type"});
        e.is_sphere = bind && strcmp(bind->Attribute("value"), "sphere") == 0;
        if (type == "point") {
            e.type = CPTS_EMIT_POINT;
            e.obj_ref = -1;
            if (auto p = child_named(ee, "point", {"center", "pos"})) {
                p->QueryFloatAttribute("x", &e.pos[0]);
                p->QueryFloatAttribute("y", &e.pos[1]);
                p->QueryFloatAttribute("z", &e.pos[2]);
            }
        } else if (type == "area") {
            e.type = CPTS_EMIT_AREA;
        } else if (type == "area-spot") {
            e.type = CPTS_EMIT_AREA_SPOT;
            float ang = 0;
            e.cos_val = 0.99f;
            // the loader only looks at the FIRST <float> child
            const tinyxml2::XMLElement *f = ee->FirstChildElement("float");
            if (f && f->Attribute("name") &&
                (!strcmp(f->Attribute("name"), "half-angle") || !strcmp(f->Attribute("name"), "angle"))) {
                f->QueryFloatAttribute("value", &ang);
                e.cos_val = cosf(ang * DEG2RAD);
            }
        } else if (type == "envmap") {
            e.type = CPTS_EMIT_ENVMAP;
            e.obj_ref = -1;
            e.Le[0] = e.Le[1] = e.Le[2] = 0, e.Le[3] = 1;
            e.scale = 1.f;
            float az = 0, ze = 0;
            if (auto c = child_named(ee, "float", {"scaler"}))
                c->QueryFloatAttribute("value", &e.scale);
            if (auto c = child_named(ee, "float", {"azimuth"}))
                c->QueryFloatAttribute("value", &az);
            if (auto c = child_named(ee, "float", {"zenith"}))
                c->QueryFloatAttribute("value", &ze);
            e.azimuth = az * DEG2RAD;
            e.zenith = ze * DEG2RAD;
        }
        ems[i] = e;
    }
    w.section("emitters", ems.data(), sizeof(CptsEmitter), ems.size());

    if (scene.num_medium > 0) { // host-side records of the media (scene.cu:775-900); entry 0 = vacuum
        std::vector<CptsMedium> media(size_t(scene.num_medium) + 1);
        memset(media.data(), 0, media.size() * sizeof(CptsMedium));
        for (int i = 1; i <= scene.num_medium && size_t(i) < scene.medium_infos.size(); i++) {
            const MediumInfo &mi = scene.medium_infos[i];
            CptsMedium m{};
            m.type = int(mi.mtype);
            m.phase_type = int(mi.ptype);
            for (int k = 0; k < 4; k++) {
                m.sigma_a[k] = mi.med_param.sigma_a[k] * (mi.mtype == MediumType::Homogeneous ? mi.med_param.scale : 1.f);
                m.sigma_s[k] = mi.med_param.sigma_s[k] * (mi.mtype == MediumType::Homogeneous ? mi.med_param.scale : 1.f);
                m.phase[k] = mi.med_param.phase[k];
            }
            media[i] = m;
        }
        dump_grids(scene, root, xml_path, media, w);
        w.section("media", media.data(), sizeof(CptsMedium), media.size());
    }
    if (!tex_descs.empty()) {
        w.section("tex_desc", tex_descs.data(), sizeof(CptsTexDesc), tex_descs.size());
        w.section("tex_data", tex_data.data(), 4, tex_data.size());
    }

    bool has_dispersion = false;
    for (const auto &b : bsdfs)
        has_dispersion |= b.type == CPTS_BSDF_DISPERSION;
    if (has_dispersion) {
        float4 *d_cie;
        float *d_d65;
        CUDA_CHECK_RETURN(cudaMalloc(&d_cie, sizeof(float4) * ColorSpaceXYZ::CIE_samples));
        CUDA_CHECK_RETURN(cudaMalloc(&d_d65, sizeof(float) * ColorSpaceXYZ::D65_samples));
        read_spectral_tables<<<5, 128>>>(d_cie, d_d65);
        std::vector<float4> cie(ColorSpaceXYZ::CIE_samples);
        std::vector<float> d65(ColorSpaceXYZ::D65_samples);
        CUDA_CHECK_RETURN(cudaMemcpy(cie.data(), d_cie, sizeof(float4) * cie.size(), cudaMemcpyDeviceToHost));
        CUDA_CHECK_RETURN(cudaMemcpy(d65.data(), d_d65, sizeof(float) * d65.size(), cudaMemcpyDeviceToHost));
        w.section("cie_xyz", cie.data(), 16, cie.size());
        w.section("d65_spd", d65.data(), 4, d65.size());
        cudaFree(d_cie);
        cudaFree(d_d65);
    }
    w.close();
}


// ---- grid media: the image of a NanoVDB grid under nearest-voxel lookups, as the brick grid of cpts_format.h.
// Values are read through the grid's own host accessor (what GridVolumeMedium::sample_density does on the device,
// src/volume/grid.cuh:51-58), never by walking the tree ourselves.
template <typename GridT, int NC>
void append_channel(const GridT *grid, CptsGridChannel &ch, std::vector<int32_t> &bricks, std::vector<float> &data) {
    const auto &map = grid->map();
    for (int k = 0; k < 9; k++)
        ch.inv_mat[k] = map.mInvMatF[k];
    for (int k = 0; k < 3; k++)
        ch.vec[k] = map.mVecF[k];
    const auto bbox = grid->indexBBox();
    const auto bg = grid->tree().background();
    for (int k = 0; k < 3; k++) {
        if constexpr (NC == 1)
            ch.background[k] = k == 0 ? float(bg) : 0.f;
        else
            ch.background[k] = float(bg[k]);
        ch.bmin[k] = bbox.min()[k] & ~7;
        ch.nb[k] = ((bbox.max()[k] - ch.bmin[k]) >> 3) + 1;
    }
    ch.ncomp = NC;
    ch.brick_offset = bricks.size();
    ch.data_offset = data.size();
    auto acc = grid->getAccessor();
    int32_t next = 0;
    std::vector<float> tmp(512 * NC);
    for (int bx = 0; bx < ch.nb[0]; bx++)
        for (int by = 0; by < ch.nb[1]; by++)
            for (int bz = 0; bz < ch.nb[2]; bz++) {
                bool any = false;
                for (int v = 0; v < 512; v++) {
                    const nanovdb::Coord c(ch.bmin[0] + bx * 8 + (v >> 6), ch.bmin[1] + by * 8 + ((v >> 3) & 7), ch.bmin[2] + bz * 8 + (v & 7));
                    const auto val = acc.getValue(c);
                    if constexpr (NC == 1) {
                        tmp[v] = float(val);
                        any |= val != bg;
                    } else {
                        for (int k = 0; k < 3; k++)
                            tmp[v * 3 + k] = float(val[k]);
                        any |= val != bg;
                    }
                }
                if (any) {
                    bricks.push_back(next++);
                    data.insert(data.end(), tmp.begin(), tmp.end());
                } else {
                    bricks.push_back(-1);
                }
            }
}

void dump_grids(const Scene &scene, const tinyxml2::XMLElement *root, const std::string &xml_path, std::vector<CptsMedium> &media,
                Writer &w) {
    std::string folder = xml_path.substr(0, xml_path.find_last_of("/\\") + 1); // scene.cu: folder_prefix
    std::vector<CptsGrid> grids;
    std::vector<int32_t> bricks;
    std::vector<float> data;
    int idx = 1;
    for (auto me = root->FirstChildElement("medium"); me; me = me->NextSiblingElement("medium"), idx++) {
        if (strcmp(me->Attribute("type"), "grid") != 0 || size_t(idx) >= media.size())
            continue;
        std::string den, alb, ems;
        for (auto e = me->FirstChildElement("string"); e; e = e->NextSiblingElement("string")) {
            const char *n = e->Attribute("name");
            if (strcmp(n, "density") == 0)
                den = folder + e->Attribute("value");
            else if (strcmp(n, "albedo") == 0)
                alb = folder + e->Attribute("value");
            else if (strcmp(n, "emission") == 0)
                ems = folder + e->Attribute("value");
        }
        const MediumInfo &mi = scene.medium_infos[idx];
        CptsGrid g{};
        g.medium = idx;
        g.scale = mi.med_param.scale;
        g.em_scale = mi.med_param.emission_scale();
        g.temp_scale = mi.med_param.temperature_scale();
        auto hd = nanovdb::io::readGrid(den);
        const nanovdb::FloatGrid *dg = hd.grid<float>();
        float lo = 0, hi = 0;
        dg->tree().extrema(lo, hi);
        g.majorant = hi;
        const auto wb = dg->worldBBox();
        for (int k = 0; k < 3; k++) {
            g.aabb_min[k] = float(wb.min()[k]); // GridVolumeMedium ctor: Vec3(bbox.min()[0], ...) (vol_grid.cu:90-94)
            g.aabb_max[k] = float(wb.max()[k]);
        }
        append_channel<nanovdb::FloatGrid, 1>(dg, g.ch[CPTS_GRID_DENSITY], bricks, data);
        { // mean of the positive voxels (compute_volume_sum, vol_grid.cu:45-80; only the unused residual estimator reads it)
            double sum = 0;
            size_t cnt = 0;
            const CptsGridChannel &c = g.ch[CPTS_GRID_DENSITY];
            for (size_t i = c.data_offset; i < data.size(); i++)
                if (data[i] > 0)
                    sum += data[i], cnt++;
            g.avg_density = cnt ? float(sum / double(cnt)) : 0.f;
        }
        if (!alb.empty()) {
            auto ha = nanovdb::io::readGrid(alb);
            append_channel<nanovdb::Vec3fGrid, 3>(ha.grid<nanovdb::Vec3f>(), g.ch[CPTS_GRID_ALBEDO], bricks, data);
        }
        if (!ems.empty()) {
            auto he = nanovdb::io::readGrid(ems);
            append_channel<nanovdb::FloatGrid, 1>(he.grid<float>(), g.ch[CPTS_GRID_TEMPERATURE], bricks, data);
        }
        media[idx].grid = int(grids.size());
        grids.push_back(g);
    }
    if (grids.empty())
        return;
    w.section("grids", grids.data(), sizeof(CptsGrid), grids.size());
    w.section("grid_bricks", bricks.data(), 4, bricks.size());
    w.section("grid_data", data.data(), 4, data.size());
    // the emission colour table, read the way GridVolumeManager::load_black_body_data does (vol_grid.cu:304-326)
    std::ifstream file(folder + "../data/blackbody.bin", std::ios::binary);
    std::vector<float> bb;
    float p[3];
    while (file.read(reinterpret_cast<char *>(p), sizeof(p))) {
        bb.insert(bb.end(), {p[0], p[1], p[2], 1.f});
    }
    w.section("blackbody", bb.data(), 16, bb.size() / 4);
}

// `mkvdb <prefix> <n>`: synthetic NanoVDB assets (the reference's .nvdb files are not distributed): a smoke puff of n
// voxels in height over the tall box of the cornell scene, written with the vendored NanoVDB tools.
void dump_scene(const Scene &scene, const std::string &xml_path, const std::string &out_path) {
    Writer w(out_path);
    dump_scene(scene, xml_path, w);
}

// the .cpts image of `scene` in memory (malloc'ed by open_memstream; caller frees *data)
void flatten_scene_to_memory(const Scene &scene, const std::string &xml_path, char **data, size_t *bytes) {
    FILE *ms = open_memstream(data, bytes);
    Writer w(ms);
    w.in_memory = true;
    dump_scene(scene, xml_path, w); // closes the stream: *data / *bytes are final
    CptsHeader h{};
    memcpy(h.magic, CPTS_MAGIC, 8);
    h.n_sections = w.n;
    memcpy(*data, &h, sizeof(h));
}

} // namespace
