// wavefront.cuh — host-side renderer object behind the C-ABI (include/cpt_b200.h).
//
// Replaces, for the wavefront path, the reference's
//   WavefrontPathTracer (src/renderer/wf_path_tracer.cuh, src/pt_impl/wf_path_tracer.cu:36-230)
//   + PathTracer / TracerBase device-buffer ownership (src/pt_impl/path_tracer.cu:49-112,
//     src/renderer/tracer_base.cuh:29-160).
#pragma once
#include "scene_host.h"

#include <cuda_runtime.h>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace cptb {

enum Traversal : int { TRAVERSAL_ORDERED = 0, TRAVERSAL_DFS = 1 };

struct RenderStats {
    uint64_t samples = 0;      // pixel-samples (primary rays)
    uint64_t closest_rays = 0; // extension (closest-hit) traversals after the primary one
    uint64_t shadow_rays = 0;  // NEE shadow rays of the algorithm (one per shaded segment whose sampled emitter is not the null emitter)
    uint64_t shadow_traced = 0; // of those, how many needed a BVH traversal (pipelines 2/3 skip rays that cannot matter)
    uint64_t shaded = 0;       // path segments shaded
    uint64_t launches = 0;     // kernels launched
    double device_ms = 0;      // device time spent in render batches (CUDA events)
};

// per-kernel device time (CUDA events around every launch, only when "kernel_timing" is enabled)
struct KernelTimes {
    double ms[5] = {0, 0, 0, 0, 0};         // raygen, shade, trace, splat, shadow (split pipeline only)
    uint64_t launches[5] = {0, 0, 0, 0, 0};
    uint64_t units[5] = {0, 0, 0, 0, 0};    // pixel-samples, shaded segments, traced segments, pixel-samples, shaded segments
};

class Renderer {
  public:
    Renderer(const HostScene &scene, int device, int seed_offset);
    ~Renderer();
    Renderer(const Renderer &) = delete;
    Renderer &operator=(const Renderer &) = delete;

    // one sample per call, like TracerBase::render_raw: returns the device pointer of the running-mean buffer
    const float *render_raw(bool gamma);
    // `spp` more samples with several samples in flight; output buffer refreshed at the end
    const float *render_batch(int spp, bool gamma);
    // TracerBase::render semantics (0-based seeds, sum image, 8-bit export): fills rgba8 (w*h*4 bytes)
    void render_offline(int num_iter, bool gamma, uint8_t *rgba8);

    void download_output(float *host_rgba) const; // synchronous D2H of the running-mean buffer
    void download_variance(float *host) const;
    void download_accumulator(float *host_rgba) const;
    // debugging / parity probes: the first W*H entries (sample 0 of the last batch) of a path-state array
    void download_state(const char *which, void *host) const;

    const float *output_buffer() const { return reinterpret_cast<const float *>(d_out_); }
    float *accumulator() { return reinterpret_cast<float *>(d_image_); }
    const float *variance_buffer() const { return d_var_; }

    void reset();
    void set_seed_offset(int v) { seed_offset_ = v; }
    void update_camera(const float cam[CPTS_CAM_COUNT]);
    void set_max_depth(int d) { max_depth_ = d, auto_state_ = 0; }
    void set_samples_in_flight(int s);
    // 0 / 2 / 3: ordered pair traversal (128-bit / 256-bit loads / warp vote), 1: the reference's own DFS order, 4: warp-
    // cooperative traversal of the wide layout (traverse_coop.cuh).  A mode whose preconditions the scene does not meet
    // (ordered: stack depth and box nesting, checked at construction; cooperative: shared-memory stack) is replaced by the
    // best mode that is allowed - the constructor's fallback cannot be overridden from outside.
    // -1: automatic (the default): the ordered per-thread traversal (2) or, where the scene allows it, the stream traversal (7) for
    // the closest-hit rays or for closest-hit and shadow rays - whichever renders a batch fastest on this scene; decided from the
    // device times of the first batches (frames are bit-identical either way).  Any explicit mode turns the automatic choice off.
    void set_traversal(int t);
    int traversal() const { return (auto_traversal_ && auto_state_ >= kAutoDecided) ? auto_choice_ : traversal_; }
    bool stream_shadow_effective() const { return (auto_traversal_ && auto_state_ >= kAutoDecided) ? auto_shadow_ : sq_shadow_; }
    bool traversal_is_auto() const { return auto_traversal_; }
    // cooperative traversal knobs: triangle-phase threshold in lanes, treelet slab size in wide nodes
    void set_coop(int flush_lanes, int slab_nodes);
    // stream traversal knobs (traversal 7); a negative value keeps the current setting
    void set_stream(int refill, int flush, int steps, int blocks);
    void set_stream_shadow(bool on) { sq_shadow_ = on; }
    bool stream_shadow() const { return sq_shadow_; }
    int stream_knob(int which) const { return which == 0 ? sq_refill_ : (which == 1 ? sq_flush_ : (which == 2 ? sq_steps_ : sq_blocks_)); }
    int coop_flush() const { return coop_flush_; }
    int coop_slab() const { return coop_slab_; }
    void enable_variance(bool on);
    void set_accum_count(int c) { accum_cnt_ = c; }
    void set_kernel_timing(bool on) { kernel_timing_ = on; }
    void set_rules(int r) { rules_ = r, auto_state_ = 0; }
    // tile split: only rows [row0,row1) of the frame are rendered and accumulated (row1 <= 0: to the bottom)
    void set_rows(int row0, int row1) { row0_ = row0, row1_ = row1; }
    int row_begin() const { return row0_; }
    int row_end() const { return row1_ > 0 ? (row1_ < h_ ? row1_ : h_) : h_; }
    // 0: NEE ray traced inside the shading kernel; 1: own shadow kernel; 2: dynamic-fetch traversal kernels;
    // 3: deferred NEE (shade first, trace only the shadow rays that can change the radiance)
    void set_pipeline(int p) { pipeline_ = p, auto_state_ = 0; }
    int pipeline() const { return pipeline_; }
    int rules() const { return rules_; }
    KernelTimes kernel_times() const { return ktimes_; }
    // queue sizes of the last batch (surface pipelines): shaded / extended segments and traversed shadow rays per bounce
    int queue_counts(int32_t *paths, int32_t *shadow, int n);

    int width() const { return w_; }
    int height() const { return h_; }
    int counter() const { return accum_cnt_; }
    int device() const { return device_; }
    int max_depth() const { return max_depth_; }
    int samples_in_flight() const { return s_max_; }
    cudaStream_t stream() const { return stream_; }
    RenderStats stats();
    std::string info() const;
    void sync() const;

  private:
    struct Impl;
    void launch_batch(int first_accum, int samples, bool gamma, bool write_output, bool use_variance);
    void ensure_capacity(int samples);

    int device_ = 0, w_ = 0, h_ = 0, seed_offset_ = 0, accum_cnt_ = 0, max_depth_ = 0;
    int s_max_ = 1, traversal_ = 2 /* ordered pair-node traversal with 256-bit node loads */, num_sms_ = 148, blocks_per_sm_ = 12, rules_ = 0;
    bool variance_on_ = false, kernel_timing_ = false;
    int pipeline_ = 3, row0_ = 0, row1_ = 0;
    bool ordered_ok_ = true, coop_ok_ = false, coop_raygen_ = false, wide4_ok_ = false, stream_ok_ = false;
    int coop_flush_ = 1 /* unused since the triangle phase runs on demand */, coop_slab_ = 0;
    int dyn_refill_ = 8, dyn_steps_ = 4, dyn_minb_ = 8; // tuning knobs of pipeline 2 (env CPTB200_DYN_*)
    int dq_flush_ = 16;                                  // traversal 6: lanes with a pending leaf that start a triangle phase
    // automatic traversal choice: 0 warm-up batch, 1 time traversal 2, 2 time traversal 7, 3 time traversal 7 with the shadow rays
    // in the stream kernel too, 4 decided
    static constexpr int kAutoDecided = 4;
    bool auto_traversal_ = true, auto_shadow_ = false;
    int auto_state_ = 0, auto_samples_ = 0, auto_choice_ = 2;
    float auto_ms_per_sample_[3] = {0.f, 0.f, 0.f};
    void auto_observe(float ms, int samples);
    bool sq_shadow_ = false; // traversal 7: shadow rays through the stream kernel too (default: the ordered per-thread any-hit kernel)
    int sq_refill_ = 8, sq_flush_ = 16, sq_steps_ = 4, sq_blocks_ = 0; // traversal 7 (stream): refill / triangle-phase thresholds in lanes, walk steps per pass, CTAs per SM (0: what fits)
    KernelTimes ktimes_;
    void collect_kernel_times(int samples);
    cudaStream_t stream_ = nullptr;
    float4 *d_image_ = nullptr, *d_out_ = nullptr;
    float *d_var_ = nullptr;
    std::unique_ptr<Impl> impl_;
    RenderStats host_stats_;
};

} // namespace cptb
