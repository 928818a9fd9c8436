// capi.cu — extern "C" boundary (include/cpt_b200.h).  Exceptions never cross it.
#include "cpt_b200.h"
#include "wavefront.cuh"

#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>

struct cpt_b200_scene {
    cptb::HostScene host;
};
struct cpt_b200_renderer {
    cptb::Renderer *impl;
};

namespace {
thread_local std::string g_error;

template <typename F> int guarded(F &&fn) {
    try {
        fn();
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return 1;
    } catch (...) {
        g_error = "unknown error";
        return 2;
    }
}
#define REQUIRE(ptr)                                                                                                   \
    if (!(ptr)) {                                                                                                      \
        g_error = "null argument: " #ptr;                                                                              \
        return 3;                                                                                                      \
    }
} // namespace

extern "C" {

const char *cpt_b200_last_error(void) { return g_error.c_str(); }
const char *cpt_b200_version(void) { return "cpt_b200 0.1 (sm_100a)"; }

int cpt_b200_scene_load_xml(const char *xml_path, cpt_b200_scene **out) {
    REQUIRE(xml_path);
    REQUIRE(out);
    *out = nullptr;
    return guarded([&] {
        auto s = new cpt_b200_scene();
        try {
            cptb::load_xml(xml_path, s->host);
        } catch (...) {
            delete s;
            throw;
        }
        *out = s;
    });
}

int cpt_b200_scene_load_memory(const void *cpts_image, uint64_t bytes, cpt_b200_scene **out) {
    REQUIRE(cpts_image);
    REQUIRE(out);
    *out = nullptr;
    return guarded([&] {
        auto s = new cpt_b200_scene();
        try {
            cptb::load_cpts_memory(cpts_image, size_t(bytes), s->host);
        } catch (...) {
            delete s;
            throw;
        }
        *out = s;
    });
}

int cpt_b200_scene_load(const char *cpts_path, cpt_b200_scene **out) {
    REQUIRE(cpts_path);
    REQUIRE(out);
    *out = nullptr;
    return guarded([&] {
        auto s = new cpt_b200_scene();
        try {
            cptb::load_cpts(cpts_path, s->host);
        } catch (...) {
            delete s;
            throw;
        }
        *out = s;
    });
}

int cpt_b200_scene_save(const cpt_b200_scene *scene, const char *cpts_path) {
    REQUIRE(scene);
    REQUIRE(cpts_path);
    return guarded([&] { cptb::save_cpts(scene->host, cpts_path); });
}

int cpt_b200_scene_config(const cpt_b200_scene *scene, int32_t *config32) {
    REQUIRE(scene);
    REQUIRE(config32);
    memcpy(config32, scene->host.config, sizeof(scene->host.config));
    return 0;
}

int cpt_b200_scene_set_film(cpt_b200_scene *scene, int width, int height) {
    REQUIRE(scene);
    return guarded([&] {
        if (width <= 0 || height <= 0)
            throw std::runtime_error("film size must be positive");
        // DeviceCamera keeps half extents and 1/focal in pixels (camera_model.cu:51-55): rescale with the width
        float *c = scene->host.camera;
        const float old_w = c[CPTS_CAM_HW] * 2.f;
        c[CPTS_CAM_INV_FOCAL] *= old_w / float(width);
        c[CPTS_CAM_HW] = width * 0.5f;
        c[CPTS_CAM_HH] = height * 0.5f;
        scene->host.config[CPTS_CFG_WIDTH] = width;
        scene->host.config[CPTS_CFG_HEIGHT] = height;
    });
}

int cpt_b200_scene_camera(const cpt_b200_scene *scene, float *camera18) {
    REQUIRE(scene);
    REQUIRE(camera18);
    memcpy(camera18, scene->host.camera, sizeof(scene->host.camera));
    return 0;
}

void cpt_b200_scene_free(cpt_b200_scene *scene) { delete scene; }

int cpt_b200_create(const cpt_b200_scene *scene, int device, int seed_offset, cpt_b200_renderer **out) {
    REQUIRE(scene);
    REQUIRE(out);
    *out = nullptr;
    return guarded([&] {
        auto impl = new cptb::Renderer(scene->host, device, seed_offset);
        *out = new cpt_b200_renderer{impl};
    });
}

void cpt_b200_destroy(cpt_b200_renderer *r) {
    if (r) {
        delete r->impl;
        delete r;
    }
}

int cpt_b200_render_raw(cpt_b200_renderer *r, int gamma, const float **dev_rgba) {
    REQUIRE(r);
    return guarded([&] {
        const float *p = r->impl->render_raw(gamma != 0);
        if (dev_rgba)
            *dev_rgba = p;
    });
}

int cpt_b200_render_batch(cpt_b200_renderer *r, int spp, int gamma, const float **dev_rgba) {
    REQUIRE(r);
    return guarded([&] {
        const float *p = r->impl->render_batch(spp, gamma != 0);
        if (dev_rgba)
            *dev_rgba = p;
    });
}

int cpt_b200_render_to_host(cpt_b200_renderer *r, int spp, int gamma, float *host_rgba) {
    REQUIRE(r);
    REQUIRE(host_rgba);
    return guarded([&] {
        r->impl->render_batch(spp, gamma != 0);
        r->impl->download_output(host_rgba);
    });
}

int cpt_b200_render_offline(cpt_b200_renderer *r, int num_iter, int gamma, uint8_t *host_rgba8) {
    REQUIRE(r);
    REQUIRE(host_rgba8);
    return guarded([&] { r->impl->render_offline(num_iter, gamma != 0, host_rgba8); });
}

int cpt_b200_enable_variance(cpt_b200_renderer *r, int on) {
    REQUIRE(r);
    return guarded([&] { r->impl->enable_variance(on != 0); });
}

int cpt_b200_variance(cpt_b200_renderer *r, const float **dev_var) {
    REQUIRE(r);
    REQUIRE(dev_var);
    *dev_var = r->impl->variance_buffer();
    return 0;
}

int cpt_b200_download_variance(cpt_b200_renderer *r, float *host_var) {
    REQUIRE(r);
    REQUIRE(host_var);
    return guarded([&] { r->impl->download_variance(host_var); });
}

int cpt_b200_output(cpt_b200_renderer *r, const float **dev_rgba) {
    REQUIRE(r);
    REQUIRE(dev_rgba);
    *dev_rgba = r->impl->output_buffer();
    return 0;
}

int cpt_b200_accumulator(cpt_b200_renderer *r, float **dev_rgba_sum) {
    REQUIRE(r);
    REQUIRE(dev_rgba_sum);
    *dev_rgba_sum = r->impl->accumulator();
    return 0;
}

int cpt_b200_download_output(cpt_b200_renderer *r, float *host_rgba) {
    REQUIRE(r);
    REQUIRE(host_rgba);
    return guarded([&] { r->impl->download_output(host_rgba); });
}

int cpt_b200_debug_state(cpt_b200_renderer *r, const char *which, void *host) {
    REQUIRE(r);
    REQUIRE(which);
    REQUIRE(host);
    return guarded([&] { r->impl->download_state(which, host); });
}

int cpt_b200_counter(const cpt_b200_renderer *r) { return r ? r->impl->counter() : -1; }

int cpt_b200_set_counter(cpt_b200_renderer *r, int accum_cnt) {
    REQUIRE(r);
    r->impl->set_accum_count(accum_cnt);
    return 0;
}

int cpt_b200_reset(cpt_b200_renderer *r) {
    REQUIRE(r);
    return guarded([&] { r->impl->reset(); });
}

int cpt_b200_set_seed_offset(cpt_b200_renderer *r, int seed_offset) {
    REQUIRE(r);
    r->impl->set_seed_offset(seed_offset);
    return 0;
}

int cpt_b200_update_camera(cpt_b200_renderer *r, const float *camera18) {
    REQUIRE(r);
    REQUIRE(camera18);
    r->impl->update_camera(camera18);
    return 0;
}

int cpt_b200_width(const cpt_b200_renderer *r) { return r ? r->impl->width() : -1; }
int cpt_b200_height(const cpt_b200_renderer *r) { return r ? r->impl->height() : -1; }

int cpt_b200_set_option(cpt_b200_renderer *r, const char *key, int64_t value) {
    REQUIRE(r);
    REQUIRE(key);
    return guarded([&] {
        const std::string k = key;
        if (k == "samples_in_flight")
            r->impl->set_samples_in_flight(int(value));
        else if (k == "traversal") {
            if (value < -1 || value > 7)
                throw std::runtime_error("traversal must be -1 (automatic: 2 or 7, whichever is faster on the scene), 0 (ordered), 1 (reference dfs order), 2 (ordered, 256-bit node loads), 3 (warp vote), 4 "
                                         "(warp-cooperative, wide layout), 5 (per thread, 4-ary wide layout), 6 (per thread, deferred leaves + lane refill) "
                                         "or 7 (stream: 6 with an asynchronous ray-fetch pipeline and the epilogue in its own kernel)");
            r->impl->set_traversal(int(value)); // modes the scene does not allow are replaced (see Renderer::set_traversal)
        } else if (k == "sq_refill") {
            r->impl->set_stream(int(value), -1, -1, -1);
        } else if (k == "sq_flush") {
            r->impl->set_stream(-1, int(value), -1, -1);
        } else if (k == "sq_steps") {
            r->impl->set_stream(-1, -1, int(value), -1);
        } else if (k == "sq_blocks") {
            r->impl->set_stream(-1, -1, -1, int(value));
        } else if (k == "sq_shadow") {
            r->impl->set_stream_shadow(value != 0);
        } else if (k == "coop_flush") {
            r->impl->set_coop(int(value), int(cpt_b200_get_option(r, "coop_slab")));
        } else if (k == "coop_slab") {
            r->impl->set_coop(int(cpt_b200_get_option(r, "coop_flush")), int(value));
        } else if (k == "rules") {
            if (value < 0 || value > 2)
                throw std::runtime_error("rules must be 0 (wavefront), 1 (megakernel pt) or 2 (volumetric pt)");
            r->impl->set_rules(int(value));
        } else if (k == "split_shadow") {
            r->impl->set_pipeline(value != 0 ? 1 : 0);
        } else if (k == "pipeline") {
            if (value < 0 || value > 6)
                throw std::runtime_error("pipeline must be 0 (fused), 1 (split shadow kernel), 2 (dynamic-fetch traversal kernels), 3 (deferred NEE), "
                                         "4 (deferred NEE + dynamic-fetch traversal kernels) or 5 (deferred NEE over dense path records)");
            r->impl->set_pipeline(int(value));
        } else if (k == "row_begin" || k == "row_end") {
            const int h = r->impl->height();
            if (value < 0 || value > h)
                throw std::runtime_error("row_begin / row_end must lie in [0, height]");
            if (k == "row_begin")
                r->impl->set_rows(int(value), r->impl->row_end());
            else
                r->impl->set_rows(r->impl->row_begin(), int(value));
        } else if (k == "kernel_timing") {
            r->impl->set_kernel_timing(value != 0);
        } else if (k == "max_depth") {
            if (value < 0 || value > 64)
                throw std::runtime_error("max_depth must be in [0,64]");
            r->impl->set_max_depth(int(value));
        } else
            throw std::runtime_error("unknown option '" + k + "'");
    });
}

int64_t cpt_b200_get_option(const cpt_b200_renderer *r, const char *key) {
    if (!r || !key)
        return -1;
    const std::string k = key;
    if (k == "samples_in_flight")
        return r->impl->samples_in_flight();
    if (k == "max_depth")
        return r->impl->max_depth();
    if (k == "rules")
        return r->impl->rules();
    if (k == "split_shadow")
        return r->impl->pipeline() != 0 ? 1 : 0;
    if (k == "pipeline")
        return r->impl->pipeline();
    if (k == "traversal")
        return r->impl->traversal();
    if (k == "traversal_auto")
        return r->impl->traversal_is_auto() ? 1 : 0;
    if (k == "sq_refill")
        return r->impl->stream_knob(0);
    if (k == "sq_flush")
        return r->impl->stream_knob(1);
    if (k == "sq_steps")
        return r->impl->stream_knob(2);
    if (k == "sq_blocks")
        return r->impl->stream_knob(3);
    if (k == "sq_shadow")
        return r->impl->stream_shadow_effective() ? 1 : 0;
    if (k == "coop_flush")
        return r->impl->coop_flush();
    if (k == "coop_slab")
        return r->impl->coop_slab();
    if (k == "row_begin")
        return r->impl->row_begin();
    if (k == "row_end")
        return r->impl->row_end();
    return -1;
}

int cpt_b200_stats(cpt_b200_renderer *r, uint64_t *stats8) {
    REQUIRE(r);
    REQUIRE(stats8);
    return guarded([&] {
        const cptb::RenderStats s = r->impl->stats();
        stats8[0] = s.samples;
        stats8[1] = s.closest_rays;
        stats8[2] = s.shadow_rays;
        stats8[3] = s.shaded;
        stats8[4] = s.launches;
        stats8[5] = uint64_t(s.device_ms * 1000.0);
        stats8[6] = s.shadow_traced;
        stats8[7] = 0;
    });
}

int cpt_b200_kernel_times(cpt_b200_renderer *r, double *ms5, uint64_t *launches5, uint64_t *units5) {
    REQUIRE(r);
    REQUIRE(ms5);
    const cptb::KernelTimes t = r->impl->kernel_times();
    for (int i = 0; i < 5; i++) {
        ms5[i] = t.ms[i];
        if (launches5)
            launches5[i] = t.launches[i];
        if (units5)
            units5[i] = t.units[i];
    }
    return 0;
}

int cpt_b200_queue_counts(cpt_b200_renderer *r, int32_t *paths, int32_t *shadow, int n) {
    if (!r || n < 0)
        return -1;
    int written = 0;
    const int rc = guarded([&] { written = r->impl->queue_counts(paths, shadow, n); });
    return rc ? -1 : written;
}

int cpt_b200_info(const cpt_b200_renderer *r, char *buf, int buf_len) {
    REQUIRE(r);
    REQUIRE(buf);
    const std::string s = r->impl->info();
    if (buf_len > 0) {
        strncpy(buf, s.c_str(), size_t(buf_len - 1));
        buf[buf_len - 1] = 0;
    }
    return 0;
}

uint64_t cpt_b200_stream(const cpt_b200_renderer *r) { return r ? uint64_t(uintptr_t(r->impl->stream())) : 0; }

int cpt_b200_sync(cpt_b200_renderer *r) {
    REQUIRE(r);
    return guarded([&] { r->impl->sync(); });
}

} // extern "C"
