/*
 * b200_wavefront.cuh — REFERENCE-SIDE adapter: `B200WavefrontTracer : TracerBase`.
 *
 * The one class a maintainer of Enigmatisms/cuda-pt adds to route its path tracers to libcpt_b200.so
 * (INTEGRATION.md).  It is compiled inside the reference's build (integration/build_reference_side.sh does that for
 * `pyrender` and `pt` without touching the reference checkout) and only talks to the library through the C ABI of
 * include/cpt_b200.h.  Replaces, behind the reference's own interface (src/renderer/tracer_base.cuh:29-160):
 *     WavefrontPathTracer            src/renderer/wf_path_tracer.cuh, src/pt_impl/wf_path_tracer.cu:36-230
 *     PathTracer<...> (megakernel)   src/pt_impl/path_tracer.cu:49-172          (rules = 1)
 *     VolumePathTracer               src/pt_impl/volume_pt.cu                   (rules = 2)
 *
 * The `Scene` the reference's own loader built (its XML parser, SBVH builder, stb_image textures, NanoVDB grids) crosses
 * the ABI in memory: integration/cpts_flatten.cuh writes the .cpts container into an open_memstream() buffer,
 * cpt_b200_scene_load_memory() reads it.  No file, no environment variable, no global state.
 *
 * TracerBase's non-virtual setters (update_camera, set_seed_offset, initialize_var_buffer, reset_out_buffer) write base-class
 * members; the adapter forwards their effect at the next render call: the camera is read back from the device copy
 * TracerBase::update_camera wrote (72 B), the seed offset and the variance switch from the members.
 */
#pragma once
#include "cpt_b200.h"
#include "cpts_flatten.cuh"
#include "renderer/tracer_base.cuh"

#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

class B200WavefrontTracer : public TracerBase {
    cpt_b200_scene *flat = nullptr;
    cpt_b200_renderer *r = nullptr;
    DeviceCamera last_cam;
    bool have_cam = false, var_on = false;
    int last_seed = 0, last_reset_cnt = 0;

    static void check(int code) {
        if (code != 0)
            throw std::runtime_error(cpt_b200_last_error()); // errors come back as codes, never exit()
    }

    // effects of the base class's non-virtual setters since the last call
    void sync_base_state(const MaxDepthParams &md) {
        DeviceCamera host_cam;
        CUDA_CHECK_RETURN(cudaMemcpy(&host_cam, camera, sizeof(DeviceCamera), cudaMemcpyDeviceToHost));
        if (!have_cam || memcmp(&host_cam, &last_cam, sizeof(DeviceCamera)) != 0) {
            static_assert(sizeof(DeviceCamera) == CPTS_CAM_COUNT * sizeof(float), "DeviceCamera is the 18-float POD of cpts_format.h");
            check(cpt_b200_update_camera(r, reinterpret_cast<const float *>(&host_cam)));
            last_cam = host_cam, have_cam = true;
        }
        if (seed_offset != last_seed) {
            check(cpt_b200_set_seed_offset(r, seed_offset));
            last_seed = seed_offset;
        }
        if (var_buffer && !var_on) { // initialize_var_buffer() was called (tracer_base.cuh:108-131)
            check(cpt_b200_enable_variance(r, 1));
            var_on = true;
        }
        if (accum_cnt == 0 && cpt_b200_counter(r) != 0) // reset_out_buffer() zeroed the base class's counter
            check(cpt_b200_reset(r));
        check(cpt_b200_set_option(r, "max_depth", md.max_depth));
    }

  public:
    /* rules: -1 = what the scene's <renderer type> asks for (pt / ptd -> megakernel rules, wfpt -> wavefront, vpt -> volumetric) */
    B200WavefrontTracer(const Scene &scene, const std::string &xml_path, int device = 0, int rules = -1) : TracerBase(scene) {
        char *image = nullptr;
        size_t bytes = 0;
        flatten_scene_to_memory(scene, xml_path, &image, &bytes); // Scene -> .cpts container in memory
        const int rc = cpt_b200_scene_load_memory(image, bytes, &flat);
        free(image);
        check(rc);
        check(cpt_b200_create(flat, device, 0, &r));
        if (rules >= 0)
            check(cpt_b200_set_option(r, "rules", rules));
        CUDA_CHECK_RETURN(cudaMalloc(&camera, sizeof(DeviceCamera))); // target of TracerBase::update_camera
        CUDA_CHECK_RETURN(cudaMemcpy(camera, scene.cam, sizeof(DeviceCamera), cudaMemcpyHostToDevice));
    }

    ~B200WavefrontTracer() override {
        cpt_b200_destroy(r);
        cpt_b200_scene_free(flat);
        cudaFree(camera);
        if (output_buffer)
            cudaFree(output_buffer);
        if (var_buffer)
            cudaFree(var_buffer);
    }

    // one more sample; returns the renderer-owned device pointer of the running mean (wf_path_tracer.cu:175-230)
    CPT_CPU const float *render_raw(const MaxDepthParams &md, bool gamma_corr = false) override {
        sync_base_state(md);
        const float *dev = nullptr;
        check(cpt_b200_render_raw(r, gamma_corr ? 1 : 0, &dev));
        accum_cnt = cpt_b200_counter(r);
        return dev;
    }

    // offline render: `num_iter` samples, 8-bit RGBA as DeviceImage::export_cpu produces it (wf_path_tracer.cu:107-173)
    CPT_CPU std::vector<uint8_t> render(const MaxDepthParams &md, int num_iter = 64, bool gamma_correction = true) override {
        sync_base_state(md);
        std::vector<uint8_t> rgba(size_t(w) * size_t(h) * 4);
        check(cpt_b200_render_offline(r, num_iter, gamma_correction ? 1 : 0, rgba.data()));
        return rgba;
    }

    CPT_CPU const float *get_variance_buffer() const override {
        const float *v = nullptr;
        if (var_on && cpt_b200_variance(r, &v) == 0)
            return v;
        return nullptr;
    }
    // render_online (GL PBO path of the interactive viewer) is not served: TracerBase's default throws "Not implemented."
};
