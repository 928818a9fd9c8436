/*
 * ref_dump — TEST INFRASTRUCTURE, not product code.
 *
 * A thin driver around the UNMODIFIED reference (Enigmatisms/cuda-pt).  It is
 * compiled by oracle/build_ref.sh against the sources where they lie under
 * /root/reference and linked into oracle/_ref/ref_dump.  Nothing of the
 * reference is copied into this repository: this file only calls the
 * reference's public C++ interface:
 *     Scene(std::string)                         src/core/scene.cuh:49-107
 *     WavefrontPathTracer / PathTracer<...>      src/renderer/wf_path_tracer.cuh, path_tracer.cuh
 *     TracerBase::render_raw / update_camera ... src/renderer/tracer_base.cuh:86-160
 * exactly the way app/pt_renderer.cu:26-117 and
 * src/pyrender/python_render.cu:71-148 do.
 *
 * Sub-commands
 *   dump   <xml> <out.cpts>
 *          load the XML with the reference loader and write the flat scene
 *          (include/cpts_format.h) that our renderer and the CPU oracle read.
 *   render <xml> <out_prefix> <mode: wfpt|pt|vpt> <seed_offset> <frame,frame,...>
 *          run render_raw() and save the float RGBA output buffer after the
 *          listed (1-based) frame counts as <out_prefix>_fNNNN.bin, and the
 *          scene as <out_prefix>.cpts (same in-process Scene => same BVH /
 *          emitter primitive order, SURVEY.md §7 "hard parts").
 *   mkvdb  <prefix> [n]
 *          write synthetic <prefix>_{density,albedo,temperature}.nvdb with the vendored NanoVDB tools (the reference's
 *          own volume assets are not distributed); no scene is loaded.
 *   vdbdump <file.nvdb> <out.cpts>
 *          brick-grid image of one float grid read voxel by voxel through NanoVDB's host accessor (no scene, no GPU):
 *          the reference answer for tests of our own .nvdb reader.
 *   imgdump <image> <channels> <out.bin>
 *          the pixels the reference's texture loader gets for an image file (its stb_image instance, no scene, no GPU).
 *   bench  <xml> <mode> <warmup> <frames> [seed_offset]
 *          time `frames` render_raw() calls with CUDA events; prints one JSON line.
 */
#include <cstdio>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "core/scene.cuh"
#include "core/xyz.cuh"
// Debug probe only (`state` sub-command): read the reference's per-path payload pool, which the class keeps
// private.  The member layout is unchanged; nothing in the reference is modified or re-implemented.
#define private public
#include "renderer/wf_path_tracer.cuh"
#undef private
#include "renderer/volume_pt.cuh"

#include "cpts_format.h"

#include <nanovdb/io/IO.h>
#include <nanovdb/tools/CreateNanoGrid.h>
#include <nanovdb/tools/GridBuilder.h>
#include <cmath>
#include <fstream>

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
    explicit Writer(const std::string &path) {
        fp = fopen(path.c_str(), "wb");
        if (!fp) {
            fprintf(stderr, "cannot open %s\n", path.c_str());
            exit(2);
        }
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
        fseek(fp, 0, SEEK_SET);
        fwrite(&h, sizeof(h), 1, fp);
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

void dump_scene(const Scene &scene, const std::string &xml_path, const std::string &out_path) {
    Writer w(out_path);

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
int make_vdb(const std::string &prefix, int n) {
    using namespace nanovdb;
    const double height = 3.2, vs = height / n;
    const int nx = int(std::ceil(1.9 / vs)), ny = n, nz = int(std::ceil(1.9 / vs));
    const Vec3d origin(3.685 - 0.5 * nx * vs, 0.06, 3.515 - 0.5 * nz * vs);
    const CoordBBox box(Coord(0), Coord(nx - 1, ny - 1, nz - 1));
    auto shape = [=](const Coord &c, float &r2) {
        const float x = (float(c[0]) + 0.5f) / float(nx) * 2.f - 1.f, y = (float(c[1]) + 0.5f) / float(ny) * 2.f - 1.f,
                    z = (float(c[2]) + 0.5f) / float(nz) * 2.f - 1.f;
        r2 = x * x + y * y + z * z;
        return Vec3f(x, y, z);
    };
    auto density = [=](const Coord &c) -> float {
        float r2;
        const Vec3f p = shape(c, r2);
        if (r2 >= 1.f)
            return 0.f;
        const float wisp = 0.55f + 0.45f * std::sin(9.f * p[0] + 2.f * std::sin(5.f * p[1])) * std::cos(7.f * p[2] + 3.f * p[1]);
        const float v = (1.f - r2) * wisp * (0.35f + 0.65f * (0.5f - 0.5f * p[1]));
        return v > 0.02f ? v : 0.f;
    };
    auto albedo = [=](const Coord &c) -> Vec3f {
        float r2;
        const Vec3f p = shape(c, r2);
        if (r2 >= 1.f)
            return Vec3f(0.f);
        return Vec3f(0.55f + 0.4f * std::sin(4.f * p[1]), 0.6f + 0.3f * std::cos(5.f * p[0]), 0.7f + 0.25f * std::sin(3.f * p[2] + 1.f));
    };
    auto temperature = [=](const Coord &c) -> float {
        float r2;
        const Vec3f p = shape(c, r2);
        const float q = p[0] * p[0] * 6.f + (p[1] + 0.55f) * (p[1] + 0.55f) * 3.f + p[2] * p[2] * 6.f;
        return q < 1.f ? (1.f - q) * 0.9f : 0.f;
    };
    tools::build::Grid<float> gd(0.f, "density", GridClass::FogVolume);
    gd.setTransform(vs, origin);
    gd(density, box);
    io::writeGrid(prefix + "_density.nvdb", tools::createNanoGrid(gd));
    tools::build::Grid<Vec3f> ga(Vec3f(0.f), "albedo", GridClass::Unknown);
    ga.setTransform(vs, origin);
    ga(albedo, box);
    io::writeGrid(prefix + "_albedo.nvdb", tools::createNanoGrid(ga));
    tools::build::Grid<float> gt(0.f, "temperature", GridClass::FogVolume);
    gt.setTransform(vs, origin);
    gt(temperature, box);
    io::writeGrid(prefix + "_temperature.nvdb", tools::createNanoGrid(gt));
    printf("[ref_dump] wrote %s_{density,albedo,temperature}.nvdb (%d x %d x %d voxels of %.5f)\n", prefix.c_str(), nx, ny, nz, vs);
    // a grid with constant regions: full 8^3 leaves of one value are pruned into tile values by the NanoVDB builder, so this
    // file exercises the tile paths of a reader (lower-level tiles) next to ordinary leaves; negative index coordinates too
    auto blocks = [=](const Coord &c) -> float {
        const bool core = c[0] >= -16 && c[0] < 24 && c[1] >= 0 && c[1] < 32 && c[2] >= -8 && c[2] < 16;
        if (core)
            return 0.75f;
        const bool shell = c[0] >= -19 && c[0] < 27 && c[1] >= -3 && c[1] < 35 && c[2] >= -11 && c[2] < 19;
        return shell ? 0.1f + 0.01f * float((c[0] * 7 + c[1] * 3 + c[2] * 5) & 15) : 0.f;
    };
    tools::build::Grid<float> gb(0.f, "blocks", GridClass::FogVolume);
    gb.setTransform(0.05, Vec3d(1.0, 0.5, 2.0));
    gb(blocks, CoordBBox(Coord(-24, -8, -16), Coord(31, 39, 23)));
    io::writeGrid(prefix + "_blocks.nvdb", tools::createNanoGrid(gb));
    auto blocks3 = [=](const Coord &c) -> Vec3f {
        const float v = blocks(c);
        return v == 0.f ? Vec3f(0.f) : (v == 0.75f ? Vec3f(0.75f, 0.5f, 0.25f) : Vec3f(v, 1.f - v, 0.5f * v));
    };
    tools::build::Grid<Vec3f> gb3(Vec3f(0.f), "blocks3", GridClass::Unknown);
    gb3.setTransform(0.05, Vec3d(1.0, 0.5, 2.0));
    gb3(blocks3, CoordBBox(Coord(-24, -8, -16), Coord(31, 39, 23)));
    io::writeGrid(prefix + "_blocks3.nvdb", tools::createNanoGrid(gb3));
    return 0;
}

// `imgdump <image> <channels> <out.bin>`: what the reference's texture loader gets from its image library for this file
// (stbi_load as instantiated in src/impl/textures.cu:27-29, called like textures.cu:59,100): int32 width, height, channels
// in file, channels requested, then the 8-bit pixels.  No scene, no GPU - the reference answer for our image decoders.
extern "C" unsigned char *stbi_load(char const *filename, int *x, int *y, int *channels_in_file, int desired_channels);
extern "C" void stbi_image_free(void *retval_from_stbi_load);
int dump_image(const std::string &path, int req, const std::string &out_path) {
    int w = 0, h = 0, n = 0;
    unsigned char *px = stbi_load(path.c_str(), &w, &h, &n, req);
    if (!px) {
        fprintf(stderr, "imgdump: cannot decode %s\n", path.c_str());
        return 1;
    }
    FILE *fp = fopen(out_path.c_str(), "wb");
    const int32_t hdr[4] = {w, h, n, req};
    fwrite(hdr, 4, 4, fp);
    fwrite(px, 1, size_t(w) * h * (req ? req : n), fp);
    fclose(fp);
    stbi_image_free(px);
    return 0;
}

// `vdbdump <file.nvdb> <out.cpts>`: the brick-grid image of one float grid, every voxel read through NanoVDB's own host
// accessor (append_channel) - the reference answer for readers of the .nvdb container.  No scene, no GPU.
int dump_vdb(const std::string &path, const std::string &out_path) {
    auto h = nanovdb::io::readGrid(path);
    CptsGrid rec{};
    std::vector<int32_t> bricks;
    std::vector<float> data;
    if (const nanovdb::Vec3fGrid *g3 = h.grid<nanovdb::Vec3f>()) { // Vec3f grid: the albedo channel
        append_channel<nanovdb::Vec3fGrid, 3>(g3, rec.ch[CPTS_GRID_ALBEDO], bricks, data);
        Writer w3(out_path);
        w3.section("grids", &rec, sizeof(CptsGrid), 1);
        w3.section("grid_bricks", bricks.data(), 4, bricks.size());
        w3.section("grid_data", data.data(), 4, data.size());
        w3.close();
        printf("[ref_dump] wrote %s (%zu bricks, %zu values, tiles in tree: %u lower, %u upper)\n", out_path.c_str(), bricks.size(), data.size(),
               unsigned(h.gridMetaData()->activeTileCount(1)), unsigned(h.gridMetaData()->activeTileCount(2)));
        return 0;
    }
    const nanovdb::FloatGrid *g = h.grid<float>();
    if (!g) {
        fprintf(stderr, "vdbdump: %s holds neither a float nor a Vec3f grid\n", path.c_str());
        return 1;
    }
    append_channel<nanovdb::FloatGrid, 1>(g, rec.ch[CPTS_GRID_DENSITY], bricks, data);
    float lo = 0, hi = 0;
    g->tree().extrema(lo, hi);
    rec.majorant = hi;
    const auto wb = g->worldBBox();
    for (int k = 0; k < 3; k++)
        rec.aabb_min[k] = float(wb.min()[k]), rec.aabb_max[k] = float(wb.max()[k]);
    Writer w(out_path);
    w.section("grids", &rec, sizeof(CptsGrid), 1);
    w.section("grid_bricks", bricks.data(), 4, bricks.size());
    w.section("grid_data", data.data(), 4, data.size());
    w.close();
    printf("[ref_dump] wrote %s (%zu bricks, %zu values, tiles in tree: %u lower, %u upper)\n", out_path.c_str(), bricks.size(), data.size(),
           g->tree().nodeCount(1) ? unsigned(h.gridMetaData()->activeTileCount(1)) : 0u, unsigned(h.gridMetaData()->activeTileCount(2)));
    return 0;
}

std::unique_ptr<TracerBase> make_renderer(const Scene &scene, const std::string &mode) {
    CUDA_CHECK_RETURN(cudaMemcpyToSymbol(c_material, scene.bsdfs, scene.num_bsdfs * sizeof(BSDF *)));
    CUDA_CHECK_RETURN(
        cudaMemcpyToSymbol(c_emitter, scene.emitters, (scene.num_emitters + 1) * sizeof(Emitter *)));
    std::unique_ptr<TracerBase> r;
    if (mode == "wfpt")
        r = std::make_unique<WavefrontPathTracer>(scene);
    else if (mode == "vpt")
        r = std::make_unique<VolumePathTracer>(scene);
    else
        r = std::make_unique<PathTracer<SingleTileScheduler>>(scene, false);
    return r;
}

std::vector<int> parse_int_list(const char *txt) {
    std::vector<int> out;
    std::stringstream ss(txt);
    std::string tok;
    while (std::getline(ss, tok, ','))
        if (!tok.empty())
            out.push_back(atoi(tok.c_str()));
    return out;
}

} // namespace

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: ref_dump dump|render|bench ... (see header comment)\n");
        return 1;
    }
    std::string cmd = argv[1], xml = argv[2];
    if (cmd == "mkvdb")
        return make_vdb(argv[2], argc > 3 ? atoi(argv[3]) : 64);
    if (cmd == "vdbdump")
        return argc > 3 ? dump_vdb(argv[2], argv[3]) : 1;
    if (cmd == "imgdump")
        return argc > 4 ? dump_image(argv[2], atoi(argv[3]), argv[4]) : 1;
    CUDA_CHECK_RETURN(cudaFree(nullptr));
    ColorSpaceXYZ xyz_host;
    Scene scene(xml);
    xyz_host.init(); // dispersion tables, as pyrender does (python_render.cu:79-82)

    if (cmd == "dump") {
        if (argc < 4)
            return 1;
        dump_scene(scene, xml, argv[3]);
        printf("[ref_dump] wrote %s\n", argv[3]);
        return 0;
    }
    if (cmd == "render") {
        if (argc < 7)
            return 1;
        std::string prefix = argv[3], mode = argv[4];
        int seed_offset = atoi(argv[5]);
        std::vector<int> frames = parse_int_list(argv[6]);
        dump_scene(scene, xml, prefix + ".cpts");
        auto rdr = make_renderer(scene, mode);
        rdr->initialize_var_buffer();
        rdr->set_seed_offset(seed_offset);
        rdr->update_camera(scene.cam);
        rdr->initialize_output_buffer();
        const int W = rdr->width(), H = rdr->height();
        std::vector<float> host(size_t(W) * H * 4), hvar(size_t(W) * H);
        int last = 0;
        for (int f : frames)
            last = std::max(last, f);
        size_t next = 0;
        for (int i = 1; i <= last; i++) {
            const float *dev = rdr->render_raw(scene.config.md, false);
            if (next < frames.size() && frames[next] == i) {
                CUDA_CHECK_RETURN(
                    cudaMemcpy(host.data(), dev, host.size() * sizeof(float), cudaMemcpyDeviceToHost));
                char name[512];
                snprintf(name, sizeof(name), "%s_%s_f%04d.bin", prefix.c_str(), mode.c_str(), i);
                FILE *fp = fopen(name, "wb");
                fwrite(host.data(), sizeof(float), host.size(), fp);
                fclose(fp);
                CUDA_CHECK_RETURN(cudaMemcpy(hvar.data(), rdr->get_variance_buffer(),
                                             hvar.size() * sizeof(float), cudaMemcpyDeviceToHost));
                snprintf(name, sizeof(name), "%s_%s_var_f%04d.bin", prefix.c_str(), mode.c_str(), i);
                fp = fopen(name, "wb");
                fwrite(hvar.data(), sizeof(float), hvar.size(), fp);
                fclose(fp);
                printf("[ref_dump] frame %d saved (%dx%d)\n", i, W, H);
                ++next;
            }
        }
        return 0;
    }
    if (cmd == "bench") {
        if (argc < 6)
            return 1;
        std::string mode = argv[3];
        int warm = atoi(argv[4]), frames = atoi(argv[5]);
        int seed_offset = argc > 6 ? atoi(argv[6]) : 0;
        auto rdr = make_renderer(scene, mode);
        rdr->set_seed_offset(seed_offset);
        rdr->update_camera(scene.cam);
        rdr->initialize_output_buffer();
        for (int i = 0; i < warm; i++)
            rdr->render_raw(scene.config.md, false);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        CUDA_CHECK_RETURN(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < frames; i++)
            rdr->render_raw(scene.config.md, false);
        cudaEventRecord(e1);
        CUDA_CHECK_RETURN(cudaDeviceSynchronize());
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("{\"ref_bench\": true, \"mode\": \"%s\", \"width\": %d, \"height\": %d, "
               "\"frames\": %d, \"warmup\": %d, \"total_ms\": %.4f, \"ms_per_frame\": %.5f, "
               "\"max_depth\": %d, \"num_prims\": %d}\n",
               mode.c_str(), rdr->width(), rdr->height(), frames, warm, ms, ms / frames,
               scene.config.md.max_depth, scene.num_prims);
        return 0;
    }
    if (cmd == "state") {
        // state <xml> <out_prefix> <seed_offset> <max_depth> : one wavefront sample with the depth overridden,
        // then the raw payload arrays (per pixel: thp, L, ray_o, ray_d, interaction, sampler, pdf)
        if (argc < 6)
            return 1;
        std::string prefix = argv[3];
        int seed_offset = atoi(argv[4]), depth = atoi(argv[5]);
        CUDA_CHECK_RETURN(cudaMemcpyToSymbol(c_material, scene.bsdfs, scene.num_bsdfs * sizeof(BSDF *)));
        CUDA_CHECK_RETURN(
            cudaMemcpyToSymbol(c_emitter, scene.emitters, (scene.num_emitters + 1) * sizeof(Emitter *)));
        WavefrontPathTracer rdr(scene);
        rdr.set_seed_offset(seed_offset);
        rdr.update_camera(scene.cam);
        rdr.initialize_output_buffer();
        MaxDepthParams md = scene.config.md;
        md.max_depth = depth;
        rdr.render_raw(md, false);
        const size_t P = size_t(rdr.width()) * rdr.height();
        struct Item {
            const char *name;
            const void *ptr;
            size_t elem;
        } items[] = {{"thp", rdr.payload_buffer.thps, 16},      {"L", rdr.payload_buffer.Ls, 16},
                     {"ray_o", rdr.payload_buffer.ray_os, 16},  {"ray_d", rdr.payload_buffer.ray_ds, 16},
                     {"isect", rdr.payload_buffer.interactions, 16}, {"rng", rdr.payload_buffer.samplers, 8},
                     {"pdf", rdr.payload_buffer.pdfs, 4}};
        for (auto &it : items) {
            std::vector<char> host(P * it.elem);
            CUDA_CHECK_RETURN(cudaMemcpy(host.data(), it.ptr, host.size(), cudaMemcpyDeviceToHost));
            FILE *fp = fopen((prefix + "_d" + std::to_string(depth) + "_" + it.name + ".bin").c_str(), "wb");
            fwrite(host.data(), 1, host.size(), fp);
            fclose(fp);
        }
        printf("[ref_dump] payload state written for depth %d\n", depth);
        return 0;
    }
    fprintf(stderr, "unknown sub-command %s\n", cmd.c_str());
    return 1;
}
