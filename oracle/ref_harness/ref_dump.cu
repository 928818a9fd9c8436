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

#include "../../integration/cpts_flatten.cuh"

namespace {

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
