#!/usr/bin/env python
"""Headline benchmark: Mrays/s of the wavefront path-tracing hot path at 1920x1080 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the unmodified reference on the same box

A "step" is one pass of the hot path over one batch of paths: `--spp-per-step` (default 8) samples of the full
1080p frame of `scenes/xml/vader_wfpt_1080p.xml` (BASELINE config 1: vader geometry, 320k triangles, 2 area
lights, thin lens, 16 bounces; the textures / environment map of the upstream scene are not shipped and are
removed).  rays = primary + closest-hit + shadow traversals, counted on the device.

One JSON line is printed by rank 0.  Keys beyond the base contract:
  roofline     dominant kernel, algorithmic bytes (SURVEY §8d figures) / CUDA-event time vs the measured HBM peak
  stages       the same for every kernel of the pipeline
  cpu_baseline the CPU oracle (a port: the reference has no CPU renderer) timed on a bounded sample (N=1 only)
  e2e          the same metric through the C-ABI call with HOST buffers (camera H2D + frame D2H every step)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# algorithmic bytes per unit of work (the reference's state layout; ours is the same 92 B pool)
BYTES_PER_UNIT = {"raygen": 100, "shade": 172, "trace": 140, "splat": 64}
METRIC = "Mrays/s (primary + closest-hit + shadow traversals), wavefront PT, 1920x1080"


WORKLOADS = {"c1": "vader_wfpt_1080p.xml",       # BASELINE config 1 (default): vader geometry, no image assets
             "c2": "vader_tex_wfpt_1080p.xml",    # texture-heavy stand-in: + 2K glossy/roughness/normal/albedo maps + envmap
             "vpt": "medium_cbox_vpt_1080p.xml",  # participating media (homogeneous), volumetric path tracer
             "grid": "grid_cbox_vpt_1080p.xml",   # BASELINE config 4 stand-in: grid (NanoVDB) medium, 304x512x304 voxels
             "fog": "vader_fog_vpt_1080p.xml",    # volumetric path on the 320k-triangle vader geometry inside a thin fog
             "whiskey": "../ref_xml/whiskey.xml"}  # the reference's own wavefront demo scene, unmodified (1024 x 1024, 16 bounces)


def ensure_volumes():
    """The grid workload's .nvdb files (~0.5 GB) are synthetic and too large to ship: written on first use with the
    vendored NanoVDB tools of the reference harness (`ref_dump mkvdb`, no GPU involved).  Both arms read the same files."""
    vol = os.path.join(ROOT, "scenes", "volume")
    done = os.path.join(vol, "puff512.done")
    if os.path.exists(done):
        return
    if int(os.environ.get("RANK", "0")) != 0:  # rank 0 writes them
        t0 = time.time()
        while not os.path.exists(done):
            if time.time() - t0 > 300:
                raise SystemExit("bench.py --scene grid: timed out waiting for rank 0 to write the volumes")
            time.sleep(0.5)
        return
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
    if not os.path.exists(tool):
        raise SystemExit("bench.py --scene grid: scenes/volume/puff512_*.nvdb missing and oracle/_ref/ref_dump (which writes them) not built")
    os.makedirs(vol, exist_ok=True)
    subprocess.run([tool, "mkvdb", os.path.join(vol, "puff512"), "512"], check=True, stdout=subprocess.DEVNULL)
    open(done, "w").write("ok\n")


def pick_scene(arg):
    if arg == "grid":
        ensure_volumes()
    if arg in WORKLOADS:
        return os.path.normpath(os.path.join(ROOT, "scenes", "xml", WORKLOADS[arg])), os.path.basename(WORKLOADS[arg])
    if arg:
        return arg, os.path.basename(arg)
    vader = os.path.join(ROOT, "scenes", "xml", "vader_wfpt_1080p.xml")
    if os.path.exists(os.path.join(ROOT, "scenes", "meshes", "vader", "vader.obj")):
        return vader, "vader_wfpt_1080p.xml"
    # meshes absent (fresh checkout without the asset copy): the committed cornell scene re-targeted to 1080p
    return os.path.join(ROOT, "tests", "golden", "ref", "cbox_wfpt_256.cpts"), "cbox_wfpt_256.cpts@1920x1080"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def rays_table():
    p = os.path.join(ROOT, "profiles", "workload_rays.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def cpu_oracle_rate(scene_cpts, width, height, budget_s=12.0, rules=0):
    """Mrays/s of the CPU oracle (OpenMP, all host cores) on a bounded sample: 1 sample of a centred row band."""
    import oracle
    sc = oracle.OracleScene(scene_cpts)
    mid = height // 2
    t0 = time.perf_counter()
    _, st = sc.render(1, rows=(mid - 4, mid + 4), rules=rules)
    dt = max(time.perf_counter() - t0, 1e-3)
    rows = int(min(height, max(16, 8 * budget_s / dt)))
    rows -= rows % 2
    r0 = max(0, mid - rows // 2)
    t0 = time.perf_counter()
    _, st = sc.render(1, rows=(r0, r0 + rows), rules=rules)
    dt = time.perf_counter() - t0
    return {"value": st["rays"] / dt / 1e6, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"1 spp x rows [{r0},{r0 + rows}) of {width}x{height} ({st['samples']} paths, {st['rays']} rays, {dt:.1f} s), "
                      "oracle/cpt_oracle.cpp with OpenMP; the reference has no CPU renderer",
            "msamples_per_s": st["samples"] / dt / 1e6, "rays_per_sample": st["rays"] / st["samples"]}


def run_reference(args, scene, workload):
    """--impl reference: the UNMODIFIED reference (its GPU kernels driven by its own host loop: thrust sort +
    per-bounce D2H count) through oracle/_ref/ref_dump on this box; CPU-oracle port if that binary is unavailable."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S = args.spp_per_step
    ref_dump = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
    base = {"metric": (METRIC.replace("wavefront PT", "volumetric PT") if "vpt" in workload else METRIC).replace("1920x1080", "1024x1024" if "whiskey" in workload else "1920x1080"), "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "impl": "reference",
            "data": ("cornell box geometry bundled with the reference + a synthetic 304x512x304 NanoVDB smoke volume (density, albedo, temperature grids) + two homogeneous media" if "grid" in workload else
                     "the reference's whiskey scene, unmodified" if "whiskey" in workload else
                     "vader geometry bundled with the reference inside a homogeneous fog bounded by a cullable box" if "fog" in workload else
                     "cornell box geometry bundled with the reference + three homogeneous media" if "vpt" in workload else
                     "vader geometry bundled with the reference; textures/envmap not shipped upstream, removed" if "tex" not in workload else
                     "vader geometry + seeded placeholder 2K textures and environment map (tools/make_assets.py)"),
            "config": {"workload": workload, "width": 1920, "height": 1080, "spp_per_step": S}}
    tbl = rays_table().get(workload)
    if scene.endswith(".xml") and os.path.exists(ref_dump):
        try:
            mode = "vpt" if "vpt" in workload else "wfpt"
            out = subprocess.run([ref_dump, "bench", scene, mode, str(args.warmup * S), str(args.steps * S)],
                                 capture_output=True, text=True, timeout=3000, cwd=ROOT)
            line = [l for l in out.stdout.splitlines() if l.startswith("{\"ref_bench\"")][-1]
            rb = json.loads(line)
            rps = tbl.get("reference_scene_rays_per_sample", tbl["rays_per_sample"]) if tbl else None
            ms_step = rb["ms_per_frame"] * S
            msamples = rb["width"] * rb["height"] / rb["ms_per_frame"] / 1e3
            value = msamples * rps if rps else None
            base.update({"value": value, "ms_per_step": ms_step, "msamples_per_s": msamples,
                         "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": 1, "kind": "reference",
                                          "sample": f"{args.steps * S} x {'VolumePathTracer' if mode == 'vpt' else 'WavefrontPathTracer'}::render_raw (1 spp each) on 1 GPU of this box, "
                                                    "reference kernels + its host loop (one host thread); rays/sample of the "
                                                    "reference-built scene (its SBVH) from profiles/workload_rays.json, counts parity-checked"},
                         "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                         "reference_kind": "reference GPU kernels (sm_100a rebuild) + reference host loop"})
            print(json.dumps(base))
            return
        except Exception as e:  # fall through to the CPU port
            base["reference_error"] = str(e)[:200]
    import cuda_pt_b200 as cpt
    sc = cpt.Scene(scene)
    if not scene.endswith(".xml"):
        sc.set_film(1920, 1080)
    tmp = os.path.join(tempfile.gettempdir(), "bench_scene_ref.cpts")
    sc.save(tmp)
    cb = cpu_oracle_rate(tmp, 1920, 1080, budget_s=20.0, rules={6: 2, 0: 1, 8: 1}.get(sc.config["renderer"], 0))
    base.update({"value": cb["value"], "ms_per_step": None, "cpu_baseline": cb,
                 "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default=None)
    ap.add_argument("--spp-per-step", type=int, default=16)
    ap.add_argument("--traversal", type=int, default=2)
    ap.add_argument("--pipeline", type=int, default=3)
    ap.add_argument("--split", default="samples", choices=["samples", "tiles"],
                    help="multi-GPU partition: samples (weak scaling, default) or tiles = bands of rows (strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    scene, workload = pick_scene(args.scene)
    if args.impl == "reference":
        run_reference(args, scene, workload)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import cuda_pt_b200 as cpt

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    sc = cpt.Scene(scene)
    if not scene.endswith(".xml"):
        sc.set_film(1920, 1080)
    cfg = sc.config
    W, H, S, K, Wm = cfg["width"], cfg["height"], args.spp_per_step, args.steps, args.warmup
    r = cpt.WavefrontRenderer(sc, local_rank, 0)
    r.set_option("samples_in_flight", S)
    r.set_option("traversal", args.traversal)
    r.set_option("pipeline", args.pipeline)
    r.set_option("kernel_timing", 1)
    if args.split == "tiles" and world > 1:
        # strong scaling: the frame is cut into bands of rows, every rank traces all S samples of its band
        from cuda_pt_b200.distributed import row_range
        lo, hi = row_range(H, world, rank)
        r.set_option("row_end", hi)
        r.set_option("row_begin", lo)
    else:
        # weak scaling: every rank renders its own K*S sample numbers of the same seed stream
        r.set_counter(rank * (K + Wm) * S * 2)
    ext = torch.cuda.ExternalStream(r.stream(), device=dev)
    acc = cpt.device_tensor(r.accumulator_ptr(), (H, W, 4), local_rank, r)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step():
        r.render_batch(S, False)
        if world > 1:  # one NCCL reduce of the HDR accumulator per sample batch (ddp_render.py:70-81)
            total = acc.clone()
            dist.all_reduce(total)
            torch.cuda.current_stream(dev).synchronize()

    for _ in range(Wm):
        step()
    barrier()
    st0, kt0 = r.stats(), r.kernel_times()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    t_wall = time.perf_counter()
    for _ in range(K):
        step()
    e1.record(ext)
    barrier()
    t_wall = time.perf_counter() - t_wall
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    st1, kt1 = r.stats(), r.kernel_times()
    rays = st1["rays"] - st0["rays"]
    tvec = torch.tensor([ms, float(rays), float(st1["samples"] - st0["samples"])], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tvec.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tvec.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, rays, samples = float(mx[0]), float(sm[1]), float(sm[2])
    else:
        samples = float(tvec[2])
    value = rays / (ms * 1e-3) / 1e6

    # ---- end-to-end through the C ABI with host buffers: camera in (H2D), frame out (D2H) every step
    pinned = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    cam = sc.camera
    for _ in range(2):
        r.update_camera(cam)
        r.render_to_host(S, False, out=pinned)
    barrier()
    s0 = r.stats()
    t0 = time.perf_counter()
    for _ in range(K):
        r.update_camera(cam)
        r.render_to_host(S, False, out=pinned)
    barrier()
    t_e2e = time.perf_counter() - t0
    e2e_rays = float(r.stats()["rays"] - s0["rays"])
    tv = torch.tensor([t_e2e, e2e_rays], dtype=torch.float64, device=dev)
    if world > 1:
        a = tv.clone()
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = tv.clone()
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        t_e2e, e2e_rays = float(a[0]), float(b[1])
    e2e = {"value": e2e_rays / t_e2e / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 72, "d2h_bytes_per_step": W * H * 16,
           "ms_per_step": t_e2e / K * 1e3, "timing": "host wall clock around K x (cpt_b200_update_camera + cpt_b200_render_to_host into pinned memory)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline (rank 0's kernels)
    peak, peak_src = measured_peaks()
    stages = {}
    total_kernel_ms = sum(kt1[j]["ms"] - kt0[j]["ms"] for j in kt1)
    for k in ("raygen", "shade", "trace", "splat"):
        dms = kt1[k]["ms"] - kt0[k]["ms"]
        units = kt1[k]["units"] - kt0[k]["units"]
        n = kt1[k]["launches"] - kt0[k]["launches"]
        if k == "shade":  # SURVEY 8d's "shade" unit (NEE + BSDF, 172 B/seg) covers our shading kernel + shadow kernel(s)
            dms += kt1["shadow"]["ms"] - kt0["shadow"]["ms"]
            n += kt1["shadow"]["launches"] - kt0["shadow"]["launches"]
        if n == 0 or dms <= 0:
            continue
        bpu = BYTES_PER_UNIT[k]
        if k == "splat":  # one launch folds S samples per pixel: S x 16 B radiance + 48 B accumulator/output per pixel
            units, bpu = n * W * H, 16 * S + 48
        gbs = units * bpu / (dms * 1e-3) / 1e9
        stages[k] = {"ms_per_step": dms / K, "launches_per_step": n / K, "units_per_step": units / K,
                     "bytes_per_unit": bpu, "achieved_gbs": gbs, "frac": gbs / peak,
                     "avg_launch_ms": dms / n, "share_of_step": dms / max(total_kernel_ms, 1e-9)}
    sh_ms = kt1["shadow"]["ms"] - kt0["shadow"]["ms"]
    sh_traced = st1["shadow_traced"] - st0["shadow_traced"]
    if sh_ms > 0:
        stages["shade"]["of_which_shadow_kernels_ms_per_step"] = sh_ms / K
        stages["shade"]["shadow_rays_traced_per_step"] = sh_traced / K
    if "vpt" in workload and args.pipeline == 0 and "raygen" in stages:  # megakernel form: the whole path runs in k_vpt, timed in the raygen slot;
        stages["vpt"] = stages.pop("raygen")       # per path it reads the camera sample and writes 16 B of radiance,
        v = stages["vpt"]                          # everything else stays in registers
        v["bytes_per_unit"] = 16
        v["achieved_gbs"] = v["units_per_step"] * 16 / (v["ms_per_step"] * 1e-3) / 1e9
        v["frac"] = v["achieved_gbs"] / peak
    dom = max(stages, key=lambda k: stages[k]["ms_per_step"])
    prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = None
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    vol = "vpt" in workload
    roofline = {"bound": "hbm", "kernel": {"vpt": "k_vpt", "raygen": "k_vptd_raygen" if vol else "k_raygen",
                                            "shade": "k_vptd_pre+k_vptd_post+k_vpt_march" if vol else "k_shade+k_shadow_apply",
                                            "trace": "k_vptd_trace" if vol else "k_trace", "splat": "k_splat"}[dom],
                "achieved": stages[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": stages[dom]["frac"],
                "traffic": traffic, "peak_source": peak_src,
                "note": "algorithmic bytes = SURVEY.md 8d per-unit figure x units per launch; BVH/triangle reads are L2-served "
                        "(scene fits the 126 MB L2) and traversal is latency/issue bound, so the HBM fraction is small by nature"}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        tmp = os.path.join(tempfile.gettempdir(), "bench_scene.cpts")
        sc.save(tmp)
        cpu_baseline = cpu_oracle_rate(tmp, W, H, rules=int(r.get_option("rules")))

    out = {"metric": (METRIC.replace("wavefront PT", "volumetric PT") if "vpt" in workload else METRIC).replace("1920x1080", f"{W}x{H}"), "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": Wm,
           "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if (args.split == "tiles" and world > 1) else "weak", "vs_baseline": None, "dtype": "f32",
           "data": ("cornell box geometry bundled with the reference + a synthetic 304x512x304 NanoVDB smoke volume (density, albedo, temperature "
                    "grids; `ref_dump mkvdb`) read by our own .nvdb reader + two homogeneous media; fixed seeds"
                    if "grid" in workload else
                    "the reference's whiskey scene, unmodified (scene/xml/whiskey.xml + its meshes); fixed seeds" if "whiskey" in workload else
                    "vader geometry bundled with the reference inside a homogeneous fog bounded by a cullable box (scenes/assets/fogbox); fixed seeds"
                    if "fog" in workload else
                    "cornell box geometry bundled with the reference + three homogeneous media (the reference's medium-cbox scene); fixed seeds"
                    if "vpt" in workload else
                    "vader geometry bundled with the reference + seeded placeholder 2K textures and environment map (tools/make_assets.py); fixed seeds"
                    if "tex" in workload else
                    "vader geometry bundled with the reference (textures/envmap not shipped upstream, removed); fixed seeds"),
           "config": {"workload": workload, "width": W, "height": H, "spp_per_step": S, "max_depth": cfg["max_depth"],
                      "triangles": cfg["num_prims"], "parallelism": (f"{'row-band tile' if args.split == 'tiles' else 'sample'}-split x{world}, 1 NCCL all-reduce of the fp32 accumulator per step"
                                      if world > 1 else "1 GPU"),
                      "l2_note": f"path-state pool {S * W * H * 92 / 1e6:.0f} MB per step >> 126 MB L2 (no flush needed)",
                      "traversal": {0: "ordered pair-node", 1: "reference DFS order", 2: "ordered pair-node, 256-bit node loads", 3: "ordered, warp-vote"}[args.traversal],
                      "pipeline": ({0: "volumetric megakernel", 1: "volumetric wavefront over slot queues"}.get(args.pipeline, "volumetric wavefront, dense path records")
                                   if "vpt" in workload else
                                   {0: "fused", 1: "split shadow kernel", 2: "dynamic-fetch traversal", 3: "deferred NEE", 4: "deferred NEE + dynamic fetch",
                                    5: "deferred NEE, dense path records"}[args.pipeline]),
                      "ray_accounting": "rays = primary + closest-hit + NEE shadow rays of the algorithm (SURVEY.md 8d, one shadow ray per shaded "
                                        "segment whose sampled emitter is not the null emitter); shadow rays that provably cannot change the "
                                        "radiance are not traversed (shadow_traced_frac)"},
           "msamples_per_s": samples / (ms * 1e-3) / 1e6, "rays_per_sample": rays / samples,
           "shadow_traced_frac": (st1["shadow_traced"] - st0["shadow_traced"]) / max(1, st1["shadow_rays"] - st0["shadow_rays"]),
           "gpu_launches": int((st1["launches"] - st0["launches"])), "wall_s": t_wall,
           "clocks": clk, "e2e": e2e, "roofline": roofline, "stages": stages, "cpu_baseline": cpu_baseline}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
