/*
 * cpt_b200.h — C ABI of the B200-native wavefront path tracer (libcpt_b200.so).
 *
 * This is the drop-in boundary for the per-bounce hot path of Enigmatisms/cuda-pt
 * (SURVEY.md §8b).  Every entry point names the reference interface it replaces;
 * INTEGRATION.md shows the ~60-line `B200WavefrontTracer : TracerBase` adapter and the
 * two-line switch in app/pt_renderer.cu / src/pyrender/python_render.cu that route
 * `RendererType::WavefrontPT` here.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a
 * non-zero code on failure (never exit()s, unlike CUDA_CHECK_RETURN in the reference,
 * src/core/cuda_utils.cuh:86-96); cpt_b200_last_error() returns the message of the last
 * failure on the calling thread.  One handle = one CUDA device + one CUDA stream;
 * handles are not thread safe.  There is NO CPU fallback: creating a renderer without
 * a CUDA device fails.
 */
#ifndef CPT_B200_H
#define CPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cpt_b200_scene cpt_b200_scene;       /* host-side flat scene */
typedef struct cpt_b200_renderer cpt_b200_renderer; /* device-side wavefront renderer */

const char *cpt_b200_last_error(void);
const char *cpt_b200_version(void);

/* ---- scene: replaces `Scene::Scene(std::string path)` (src/impl/scene.cu:937-1192) as the
 * producer of what a renderer constructor consumes (src/core/scene.cuh:49-107). ---- */

/* Mitsuba-like XML v1.2 scene file (same grammar as the reference loader). */
int cpt_b200_scene_load_xml(const char *xml_path, cpt_b200_scene **out);
/* Flat container written by cpt_b200_scene_save() or by the reference-side adapter
 * (include/cpts_format.h): the exact arrays of a reference `Scene`. */
int cpt_b200_scene_load(const char *cpts_path, cpt_b200_scene **out);
/* The same container handed over IN MEMORY (no file): what the reference-side adapter integration/b200_wavefront.cuh does
 * with the `Scene` the reference's own loader built (src/core/scene.cuh:49-107 -> integration/cpts_flatten.cuh).  The image
 * is only read during the call and may be freed afterwards. */
int cpt_b200_scene_load_memory(const void *cpts_image, uint64_t bytes, cpt_b200_scene **out);
int cpt_b200_scene_save(const cpt_b200_scene *scene, const char *cpts_path);
/* config[CPTS_CFG_COUNT] as in cpts_format.h (width, height, spp, max_depth, gamma, renderer ...) */
int cpt_b200_scene_config(const cpt_b200_scene *scene, int32_t *config32);
int cpt_b200_scene_set_film(cpt_b200_scene *scene, int width, int height);
/* camera[18]: reference DeviceCamera POD (src/core/camera_model.cuh:40-48) */
int cpt_b200_scene_camera(const cpt_b200_scene *scene, float *camera18);
void cpt_b200_scene_free(cpt_b200_scene *scene);

/* ---- renderer: replaces WavefrontPathTracer (src/renderer/wf_path_tracer.cuh:30-72,
 * src/pt_impl/wf_path_tracer.cu:36-48) and the TracerBase/PathTracer buffer ownership
 * (src/renderer/tracer_base.cuh:62-69, src/pt_impl/path_tracer.cu:49-99). ---- */
int cpt_b200_create(const cpt_b200_scene *scene, int device, int seed_offset, cpt_b200_renderer **out);
void cpt_b200_destroy(cpt_b200_renderer *r);

/* TracerBase::render_raw (wf_path_tracer.cu:175-230): one more sample; *dev_rgba receives the
 * DEVICE pointer of the running-mean H x W x 4 fp32 buffer (owned by the renderer, valid until
 * the next call).  The call returns after the device finished, like the reference. */
int cpt_b200_render_raw(cpt_b200_renderer *r, int gamma, const float **dev_rgba);
/* `spp` more samples with several samples in flight per launch group; same result as calling
 * cpt_b200_render_raw `spp` times (the per-pixel fp32 sum is formed in sample order). */
int cpt_b200_render_batch(cpt_b200_renderer *r, int spp, int gamma, const float **dev_rgba);
/* Host-buffer form used by the end-to-end benchmark and by non-CUDA callers: renders `spp`
 * samples and copies the running mean into host_rgba (H*W*4 floats; pinned memory recommended). */
int cpt_b200_render_to_host(cpt_b200_renderer *r, int spp, int gamma, float *host_rgba);
/* TracerBase::render (wf_path_tracer.cu:107-173): 0-based seeds, sum image, 8-bit RGBA export
 * with the reference's 1/2.1 display gamma (src/core/host_device.cuh:63-112). */
int cpt_b200_render_offline(cpt_b200_renderer *r, int num_iter, int gamma, uint8_t *host_rgba8);

/* TracerBase::get_variance_buffer / PythonRenderer::variance (python_render.cu:62-69) */
int cpt_b200_enable_variance(cpt_b200_renderer *r, int on);
int cpt_b200_variance(cpt_b200_renderer *r, const float **dev_var);
int cpt_b200_download_variance(cpt_b200_renderer *r, float *host_var);
/* running-mean output / sum accumulator (DeviceImage, src/core/host_device.cuh:30-62) */
int cpt_b200_output(cpt_b200_renderer *r, const float **dev_rgba);
int cpt_b200_accumulator(cpt_b200_renderer *r, float **dev_rgba_sum);
int cpt_b200_download_output(cpt_b200_renderer *r, float *host_rgba);

/* parity probe: copies the first W*H entries of a per-path state array of the last batch to the host.
 * which = "thp" | "L" | "ray_o" | "ray_d" | "isect" (16 B each) | "rng" (8 B) | "pdf" (4 B); same layout as
 * the reference's PayLoadBufferSoA (src/renderer/wavefront_pt.cuh:74-213). */
int cpt_b200_debug_state(cpt_b200_renderer *r, const char *which, void *host);

/* TracerBase::cnt / reset_out_buffer / set_seed_offset / update_camera (tracer_base.cuh:135-160) */
int cpt_b200_counter(const cpt_b200_renderer *r);
int cpt_b200_set_counter(cpt_b200_renderer *r, int accum_cnt); /* after an external reduce of the accumulator */
int cpt_b200_reset(cpt_b200_renderer *r);
int cpt_b200_set_seed_offset(cpt_b200_renderer *r, int seed_offset);
int cpt_b200_update_camera(cpt_b200_renderer *r, const float *camera18);
int cpt_b200_width(const cpt_b200_renderer *r);
int cpt_b200_height(const cpt_b200_renderer *r);

/* options: "rules" (0 wavefront termination rules, wavefront_pt.cu:186-195; 1 megakernel "pt" rules,
 * megakernel_pt.cu:159-180; 2 volumetric path tracer with participating media, megakernel_vpt.cu:203-456 - default
 * taken from the scene's <renderer type>), "samples_in_flight" (>=1),
 * "traversal" (-1 automatic, the default: 2 or 7 (closest-hit rays only, or shadow rays too), whichever renders the first batches of
 * the scene fastest - frames are bit-identical either way; 0 / 2 / 3 ordered pair-node traversal: 128-bit loads / 256-bit loads /
 * warp-vote loop; 1 reference DFS order; 4 warp-cooperative traversal of the 4-ary wide layout; 5 per-thread traversal of that
 * layout; 6 per-thread with deferred leaves and lane refill; 7 stream traversal: 6 with an asynchronous ray-fetch pipeline per warp,
 * the walk state in registers, the top of the stack in shared memory and the epilogue per finished 32-ray chunk - a mode whose
 * preconditions the scene does not meet is replaced, read the effective one back with cpt_b200_get_option; "traversal_auto" reads
 * whether the choice is automatic),
 * "sq_refill" / "sq_flush" / "sq_steps" / "sq_blocks" / "sq_shadow" (traversal 7: idle lanes that trigger a refill, lanes with a
 * pending leaf that start a triangle phase, walk steps per pass, CTAs per SM (0: what fits), shadow rays through the stream kernel),
 * "coop_slab" (traversal 4: top nodes of the wide tree staged in shared memory by a TMA bulk copy, default 0),
 * "max_depth" (MaxDepthParams::max_depth, src/core/max_depth.h:26-42), "kernel_timing" (0/1: CUDA events
 * around every launch, read back with cpt_b200_kernel_times), "pipeline" (0: the NEE visibility ray is traced
 * inside the shading kernel; 1: in its own traversal kernel before shading; 2: dynamic-fetch traversal
 * kernels; 3: deferred NEE - shade first, then trace only the shadow rays that can change the radiance; 4: 3 with the
 * dynamic-fetch kernels; 5: 3 over dense path records - the shading kernel compacts the paths themselves instead of a
 * queue of slot ids.  Volumetric scenes (rules 2): 0 = megakernel form, 1 = wavefront over slot queues, >= 2 =
 * wavefront with dense path records),
 * "split_shadow" (legacy alias: 0 -> pipeline 0, 1 -> pipeline 1), "row_begin" / "row_end" (tile split across GPUs:
 * only rows [row_begin,row_end) are traced and accumulated, pixel indices and therefore sample streams stay those of
 * the full frame; set row_end first when moving the window down) */
int cpt_b200_set_option(cpt_b200_renderer *r, const char *key, int64_t value);
int64_t cpt_b200_get_option(const cpt_b200_renderer *r, const char *key);

/* stats[8]: pixel-samples, closest-hit rays after the primary one, shadow rays (one per shaded segment whose
 * sampled emitter is not the null emitter, wavefront_pt.cu:273), shaded segments, kernel launches, device
 * microseconds, shadow rays that needed a BVH traversal (<= shadow rays), 0 (cumulative since creation) */
int cpt_b200_stats(cpt_b200_renderer *r, uint64_t *stats8);
/* cumulative per-kernel device time when "kernel_timing" is on; 5 entries: 0 raygen+primary hit, 1 shade
 * (NEE + BSDF), 2 trace (closest hit + RR + compaction), 3 splat, 4 shadow (NEE visibility kernel of the
 * split pipeline).  units = pixel-samples (0,3) or path segments (1,2,4) those launches processed. */
int cpt_b200_kernel_times(cpt_b200_renderer *r, double *ms5, uint64_t *launches5, uint64_t *units5);
/* queue sizes of the LAST batch, per bounce b = 0 .. n-1: paths[b] = path segments shaded at bounce b (= segments k_trace
 * extended at bounce b), shadow[b] = shadow rays k_shadow_apply traversed at bounce b.  The unit counts of single launches:
 * what an ncu capture of "the (b+1)-th launch of k_trace" processed (profiles/).  Replaces the reference's per-bounce
 * `num_valid_ray` read-back (src/pt_impl/wf_path_tracer.cu:197-202).  Returns the number of entries written (<= n). */
int cpt_b200_queue_counts(cpt_b200_renderer *r, int32_t *paths, int32_t *shadow, int n);
/* PythonRenderer::info (python_render.cuh) — copies a one-line description */
int cpt_b200_info(const cpt_b200_renderer *r, char *buf, int buf_len);
/* the CUDA stream all work of this handle is issued on (as a cudaStream_t value) */
uint64_t cpt_b200_stream(const cpt_b200_renderer *r);
int cpt_b200_sync(cpt_b200_renderer *r);

#ifdef __cplusplus
}
#endif
#endif /* CPT_B200_H */
