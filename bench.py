#!/usr/bin/env python
"""Headline benchmark: Mrays/s of the wavefront path-tracing hot path at 1920x1080 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the unmodified reference on the same box

A "step" is one pass of the hot path over one batch of paths: `--spp-per-step` (default 64) samples of the full
1080p frame of `scenes/xml/vader_wfpt_1080p.xml` (BASELINE config 1: vader geometry, 320k triangles, 2 area
lights, thin lens, 16 bounces; the textures / environment map of the upstream scene are not shipped and are
removed).  rays = primary + closest-hit + shadow traversals, counted on the device.

The library chooses its traversal kernels by measurement during the first four batches of a renderer (warm-up, ordered
per-thread kernels, stream kernel for closest-hit rays, stream kernels for both; frames are bit-identical either way), so `--warmup 4`
keeps the trials out of the timed region; with `--warmup 3` the last trial is the first timed step.  `config.traversal` names the choice.

One JSON line is printed by rank 0.  Keys beyond the base contract:
  roofline     dominant kernel, algorithmic bytes (SURVEY §8d figures) / CUDA-event time vs the measured HBM peak
  stages       the same for every kernel of the pipeline
  cpu_baseline the CPU oracle (a port: the reference has no CPU renderer) timed on a bounded sample (N=1 only)
  e2e          the same metric through the C-ABI call with HOST buffers: camera H2D + frame D2H every step; at N>1 the frame is
               the REDUCED one (the per-step NCCL reduce to rank 0 is inside the timed region, the D2H happens on rank 0)
  traced_mrays the rays that were actually traversed (deferred NEE skips shadow rays that cannot change the radiance; `value`
               counts the algorithm's rays, like the reference arm)
  dropin       the same step on the scene the REFERENCE's loader builds for this XML (its SBVH; the drop-in case of
               INTEGRATION.md) - N=1 only, when oracle/_ref/ref_dump is on the box
  megakernel_rules  the same step under the reference megakernel's termination rules (rules=1), the like-for-like partner
               of the reference arm's `megakernel_pt` block (the reference's README calls its megakernel the faster renderer)

Reference arm at N>1: every rank drives the unmodified reference on its own GPU (ddp_render.py semantics: full frame per
rank, seed_offset = rank * 4201), value = all ranks' samples / slowest rank's time.
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
             "c3": "dispersion_cbox_wfpt_1080p.xml",  # BASELINE config 3 stand-in: cornell box with Diamond / Sapphire dispersion boxes, 7 bounces
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


def physical_gpu(local_index):
    """nvidia-smi index of CUDA device `local_index` (CUDA_VISIBLE_DEVICES may renumber the devices)."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "").strip()
    if vis:
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if local_index < len(ids):
            return ids[local_index]
    return str(local_index)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, physical_gpu(index)

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


def python_serial_rate(module_dir, scene, frames, gpu):
    """BASELINE.md B5, the "Python serial path for plumbing": one process looping pyrender.PythonRenderer(xml, 0, 0).render()
    (one sample + a deep copy into a fresh torch tensor per call, scripts/serial_render.py:181-192) on one GPU.  `module_dir`
    holds the nanobind module: oracle/_ref/pyrender_ref (the unmodified reference) or oracle/_ref/b200_side (the same module
    with its path-tracer cases routed to libcpt_b200.so through integration/b200_wavefront.cuh)."""
    if not os.path.exists(os.path.join(module_dir, "pyrender.so")) or not scene.endswith(".xml"):
        return None
    code = ("import sys, time, json\nsys.path.insert(0, %r)\nimport torch, pyrender\nr = pyrender.PythonRenderer(%r, 0, 0)\n"
            "for _ in range(4):\n    img = r.render()\ntorch.cuda.synchronize()\nt0 = time.perf_counter()\n"
            "for _ in range(%d):\n    img = r.render()\ntorch.cuda.synchronize()\ndt = time.perf_counter() - t0\n"
            "print(json.dumps({'py_serial': True, 'ms_per_frame': dt / %d * 1e3, 'pixels': int(img.shape[0] * img.shape[1])}))\nr.release()\n"
            % (module_dir, scene, frames, frames))
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = str(gpu)
    try:
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1800, cwd=ROOT, env=env)
        line = [l for l in out.stdout.splitlines() if l.startswith('{"py_serial"')][-1]
        rec = json.loads(line)
        return {"ms_per_frame": rec["ms_per_frame"], "msamples_per_s": rec["pixels"] / rec["ms_per_frame"] / 1e3, "frames": frames,
                "what": "wall clock of a Python loop of PythonRenderer.render() calls (1 spp + deep copy to a fresh CUDA tensor each), module " + os.path.relpath(module_dir, ROOT)}
    except Exception as e:
        return {"error": (str(e) + " " + (out.stderr[-200:] if 'out' in dir() else ""))[:300]}


def _ref_bench(ref_dump, scene, mode, warm, frames, seed_offset, gpu):
    """one `ref_dump bench` process = the reference's own render_raw loop on one GPU; returns its JSON record"""
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = str(gpu)
    out = subprocess.run([ref_dump, "bench", scene, mode, str(warm), str(frames), str(seed_offset)], capture_output=True, text=True, timeout=3000,
                         cwd=ROOT, env=env)
    line = [l for l in out.stdout.splitlines() if l.startswith("{\"ref_bench\"")]
    if not line:
        raise RuntimeError("ref_dump bench produced no record: " + (out.stderr or out.stdout)[-300:])
    return json.loads(line[-1])


def run_reference(args, scene, workload):
    """--impl reference: the UNMODIFIED reference (its GPU kernels driven by its own host loop: thrust sort + per-bounce D2H
    count) through oracle/_ref/ref_dump.  N>1: every rank runs it on its own GPU with seed_offset = rank * 4201 (the reference's
    multi-GPU scheme, scripts/ddp_render.py:141-143: full frame per rank, different seeds); value = all ranks' samples / the
    slowest rank's device time.  The reference's reduce (one all_reduce of the frame every 128 frames, ddp_render.py:161-166) is
    not inside this number - it would add < 0.2 ms per 128 frames.  CPU-oracle port if the binary is unavailable (rank 0 only)."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    S = args.spp_per_step
    ref_dump = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
    base = {"metric": (METRIC.replace("wavefront PT", "volumetric PT") if "vpt" in workload else METRIC).replace("1920x1080", "1024x1024" if "whiskey" in workload else "1920x1080"), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "impl": "reference",
            "data": ("cornell box geometry bundled with the reference + a synthetic 304x512x304 NanoVDB smoke volume (density, albedo, temperature grids) + two homogeneous media" if "grid" in workload else
                     "the reference's whiskey scene, unmodified" if "whiskey" in workload else
                     "vader geometry bundled with the reference inside a homogeneous fog bounded by a cullable box" if "fog" in workload else
                     "cornell box geometry bundled with the reference + three homogeneous media" if "vpt" in workload else
                     "cornell box bundled with the reference, boxes switched to the spectral dispersion BSDF (Diamond / Sapphire)" if "dispersion" in workload else
                     "vader geometry bundled with the reference; textures/envmap not shipped upstream, removed" if "tex" not in workload else
                     "vader geometry + seeded placeholder 2K textures and environment map (tools/make_assets.py)"),
            "config": {"workload": workload, "width": 1920, "height": 1080, "spp_per_step": S}}
    tbl = rays_table().get(workload)
    have_ref = scene.endswith(".xml") and os.path.exists(ref_dump)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")  # host-side exchange of the per-rank records only
    if have_ref:
        err = None
        rec = {}
        try:
            mode = "vpt" if "vpt" in workload else "wfpt"
            gpu = physical_gpu(local_rank)
            rec["main"] = _ref_bench(ref_dump, scene, mode, args.warmup * S, args.steps * S, rank * 4201, gpu)
            if mode == "wfpt" and not args.no_extras:  # the reference's megakernel on the same scene (its README: the faster of its two)
                rec["pt"] = _ref_bench(ref_dump, scene, "pt", args.warmup * S, max(S, args.steps * S // 2), rank * 4201, gpu)
        except Exception as e:
            err = str(e)[:300]
        allrec = [(rec, err)]
        if world > 1:
            allrec = [None] * world
            dist.all_gather_object(allrec, (rec, err))
            dist.destroy_process_group()
        if rank != 0:
            return
        errs = [e for _, e in allrec if e]
        if not errs:
            mains = [r["main"] for r, _ in allrec]
            rps = tbl.get("reference_scene_rays_per_sample", tbl["rays_per_sample"]) if tbl else None
            if rps is None:
                # no committed ray count for this workload: count them on the scene the reference's loader builds (`ref_dump dump`) with our
                # renderer (8 spp; the per-bounce path counts on reference-built scenes are parity-checked).  Only the conversion factor
                # Msamples/s -> Mrays/s comes from here, never a time
                try:
                    import cuda_pt_b200 as cpt
                    tmpc = os.path.join(tempfile.gettempdir(), "bench_ref_count.cpts")
                    subprocess.run([ref_dump, "dump", scene, tmpc], check=True, capture_output=True, timeout=900, cwd=ROOT)
                    rc = cpt.WavefrontRenderer(tmpc, local_rank, 0)
                    rc.render_batch(8, False)
                    stc = rc.stats()
                    rps = stc["rays"] / stc["samples"]
                    rc.close()
                    base["rays_per_sample_source"] = "counted now on the reference-built scene (8 spp)"
                except Exception as e:
                    base["rays_per_sample_error"] = str(e)[:200]
            px = mains[0]["width"] * mains[0]["height"]
            t_max = max(m["total_ms"] for m in mains)
            samples = sum(px * m["frames"] for m in mains)
            msamples = samples / t_max / 1e3
            value = msamples * rps if rps else None
            base.update({"value": value, "ms_per_step": t_max / args.steps, "msamples_per_s": msamples,
                         "per_rank_ms_per_frame": [round(m["ms_per_frame"], 3) for m in mains],
                         "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": world, "kind": "reference",
                                          "sample": f"{args.steps * S} x {'VolumePathTracer' if mode == 'vpt' else 'WavefrontPathTracer'}::render_raw (1 spp each) on each of {world} GPU(s) of this box "
                                                    "(seed_offset = rank * 4201), reference kernels + its host loop (one host thread per GPU); Mrays/s = measured "
                                                    "Msamples/s x rays/sample of the reference-built scene (its SBVH) from profiles/workload_rays.json, i.e. the ratio "
                                                    "between the two arms is the Msamples/s ratio"},
                         "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                         "reference_kind": "reference GPU kernels (sm_100a rebuild) + reference host loop, one process per GPU"})
            pts = [r.get("pt") for r, _ in allrec]
            if all(pts):
                t_pt = max(m["total_ms"] for m in pts)
                ms_pt = sum(px * m["frames"] for m in pts) / t_pt / 1e3
                base["megakernel_pt"] = {"msamples_per_s": ms_pt, "ms_per_frame": [round(m["ms_per_frame"], 3) for m in pts],
                                          "note": "PathTracer<SingleTileScheduler>::render_raw on the same scene (what `pt` runs for type=\"pt\" XMLs): "
                                                  "Russian roulette from bounce 1 and per-lobe caps, so fewer rays per sample than the wavefront renderer; "
                                                  "compare Msamples/s with the `megakernel_rules` block of our arm"}
                base["best_msamples_per_s"] = max(msamples, ms_pt)
            if not args.no_extras:
                ps = python_serial_rate(os.path.join(ROOT, "oracle", "_ref", "pyrender_ref"), scene, max(16, args.steps * S // 4), physical_gpu(0))
                if ps:
                    base["python_serial"] = ps
            print(json.dumps(base))
            return
        base["reference_error"] = "; ".join(errs)[:300]
    elif world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if rank != 0:
        return
    import cuda_pt_b200 as cpt
    sc = cpt.Scene(scene)
    if not scene.endswith(".xml"):
        sc.set_film(1920, 1080)
    tmp = os.path.join(tempfile.gettempdir(), "bench_scene_ref.cpts")
    sc.save(tmp)
    cb = cpu_oracle_rate(tmp, 1920, 1080, budget_s=20.0, rules={6: 2, 0: 1, 8: 1}.get(sc.config["renderer"], 0))
    base.update({"n_gpus": 0, "value": cb["value"], "ms_per_step": None, "cpu_baseline": cb,
                 "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)  # x 64 samples per step = the 1024 spp of BASELINE config 1
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default=None)
    ap.add_argument("--spp-per-step", type=int, default=64)
    ap.add_argument("--traversal", type=int, default=None, help="traversal mode (default: the library's automatic choice between the ordered per-thread "
                    "traversal, 2, and the stream traversal, 7, made from the device times of the warm-up steps; modes the scene does not allow are replaced)")
    ap.add_argument("--pipeline", type=int, default=None)
    ap.add_argument("--no-extras", action="store_true", help="skip the dropin / megakernel-rules legs (and the reference arm's megakernel run)")
    ap.add_argument("--no-balance", action="store_true", help="tile split with static row bands")
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
    if args.traversal is not None:
        r.set_option("traversal", args.traversal)
    if args.pipeline is not None:
        r.set_option("pipeline", args.pipeline)
    pipeline = int(r.get_option("pipeline"))
    r.set_option("kernel_timing", 1)
    from cuda_pt_b200.distributed import DistributedRenderer
    tiles = args.split == "tiles" and world > 1
    # weak scaling (default): a step = S samples PER RANK (rank r renders the r-th block of S of every step's world*S samples of the
    # same seed stream); strong scaling (--split tiles): a step = S samples of the frame, cut into row bands that follow the
    # measured batch times.  Either way ONE reduce of the fp32 accumulator to rank 0 per step, overlapped with the next step.
    dr = DistributedRenderer(r, split="tiles" if tiles else "samples", dst=0, balance=not args.no_balance)
    step_spp = S if tiles else S * world
    ext = torch.cuda.ExternalStream(r.stream(), device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(Wm):
        dr.step(step_spp)
    dr.finish()
    barrier()
    st0, kt0 = r.stats(), r.kernel_times()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    t_wall = time.perf_counter()
    for _ in range(K):
        dr.step(step_spp)
    if world > 1:
        dr.finish()  # current stream now waits for the last reduce: the timed region ends when the reduced frame exists on rank 0
        e1.record(torch.cuda.current_stream(dev))
    else:
        e1.record(ext)
    barrier()
    t_wall = time.perf_counter() - t_wall
    ms = e0.elapsed_time(e1)
    rows_final = list(dr.rows) if tiles else None
    traversal, traversal_auto = int(r.get_option("traversal")), int(r.get_option("traversal_auto")) == 1  # (automatic: decided during the warm-up steps)
    stream_shadow = int(r.get_option("sq_shadow")) == 1
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

    # ---- end-to-end through the C ABI with host buffers: camera in (H2D) on every rank, frame out (D2H) every step.
    # N>1: the frame that leaves the device is the REDUCED one - the per-step reduce to rank 0 and its D2H (rank 0) are timed
    pinned = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    cam = sc.camera
    if world > 1:
        dr.host_out = pinned

    def e2e_step():
        r.update_camera(cam)
        if world > 1:
            dr.step(step_spp)
        else:
            r.render_to_host(S, False, out=pinned)

    for _ in range(2):
        e2e_step()
    dr.finish()
    barrier()
    s0 = r.stats()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    dr.finish()
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
    e2e = {"value": e2e_rays / t_e2e / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 72 * world, "d2h_bytes_per_step": W * H * 16,
           "ms_per_step": t_e2e / K * 1e3,
           "timing": ("host wall clock (max over ranks) around K x (cpt_b200_update_camera + render of the step + NCCL reduce of the fp32 accumulator to "
                      "rank 0 + D2H of the reduced frame into pinned memory on rank 0), drained at the end" if world > 1 else
                      "host wall clock around K x (cpt_b200_update_camera + cpt_b200_render_to_host into pinned memory)")}

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
    if "vpt" in workload and pipeline == 0 and "raygen" in stages:  # megakernel form: the whole path runs in k_vpt, timed in the raygen slot;
        stages["vpt"] = stages.pop("raygen")       # per path it reads the camera sample and writes 16 B of radiance,
        v = stages["vpt"]                          # everything else stays in registers
        v["bytes_per_unit"] = 16
        v["achieved_gbs"] = v["units_per_step"] * 16 / (v["ms_per_step"] * 1e-3) / 1e9
        v["frac"] = v["achieved_gbs"] / peak
    dom = max(stages, key=lambda k: stages[k]["ms_per_step"])
    prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic, l1pipe, traffic_units = None, None, None
    if os.path.exists(prof):
        try:
            ent = json.load(open(prof)).get("trace_stream" if (dom == "trace" and traversal == 7) else dom, {})
            traffic, l1pipe, traffic_units = ent.get("dram_bytes_per_launch"), ent.get("l1tex_data_pipe_pct"), ent.get("units_in_this_launch")
        except Exception:
            traffic = None
    vol = "vpt" in workload
    roofline = {"bound": "hbm", "kernel": {"vpt": "k_vpt", "raygen": "k_vptd_raygen" if vol else "k_raygen",
                                            "shade": "k_vptd_pre+k_vptd_post+k_vpt_march" if vol else "k_shade+k_shadow_apply",
                                            "trace": "k_vptd_trace" if vol else ("k_traverse_stream<closest hit>" if traversal == 7 else "k_trace"), "splat": "k_splat"}[dom],
                "achieved": stages[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": stages[dom]["frac"],
                "traffic": traffic, "traffic_units": traffic_units, "peak_source": peak_src,
                "l1tex_data_pipe_pct_of_peak": l1pipe,
                "note": "algorithmic bytes = SURVEY.md 8d per-unit figure x units per launch; `traffic` = dram bytes of ONE captured launch with "
                        "`traffic_units` units (profiles/ncu_traffic.json).  BVH/triangle reads are L2-served (the arena is pinned in the 126 MB L2) "
                        "and every step of a ray is a dependent fetch: the traversal kernels are bound by the L1/TEX pipe (`l1tex_data_pipe_pct_of_peak`, "
                        "ncu), not by HBM, so the HBM fraction is small by nature (DESIGN.md 4c)"}

    # ---- extra legs (N=1): the reference megakernel's termination rules; the scene as the reference's loader builds it
    extras = {}
    if world == 1 and not args.no_extras and "vpt" not in workload:
        def leg(rr, steps):
            rr.set_option("samples_in_flight", S)
            rr.render_batch(4 * S, False)  # warm-up; the automatic traversal choice needs four batches
            a = rr.stats()
            rr.render_batch(steps * S, False)
            rr.sync()
            b = rr.stats()
            dev_s = (b["device_us"] - a["device_us"]) / 1e6
            return {"value": (b["rays"] - a["rays"]) / dev_s / 1e6, "unit": "Mrays/s", "msamples_per_s": (b["samples"] - a["samples"]) / dev_s / 1e6,
                    "ms_per_step": dev_s / steps * 1e3, "rays_per_sample": (b["rays"] - a["rays"]) / max(1, b["samples"] - a["samples"]), "steps": steps}
        try:
            r.set_option("rules", 1)
            extras["megakernel_rules"] = dict(leg(r, max(2, K // 2)), note="rules=1: lobe caps, throughput cut-off, Russian roulette from bounce 1 "
                                              "(megakernel_pt.cu:162-180), same kernels; compare msamples_per_s with the reference arm's megakernel_pt block")
            r.set_option("rules", 0 if cfg["renderer"] == 1 else 1)
        except Exception as e:
            extras["megakernel_rules"] = {"error": str(e)[:200]}
        ref_dump = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
        if scene.endswith(".xml") and os.path.exists(ref_dump):
            try:
                tmpc = os.path.join(tempfile.gettempdir(), "bench_dropin.cpts")
                subprocess.run([ref_dump, "dump", scene, tmpc], check=True, capture_output=True, timeout=900, cwd=ROOT)
                rd = cpt.WavefrontRenderer(tmpc, local_rank, 0)
                if args.traversal is not None:
                    rd.set_option("traversal", args.traversal)
                extras["dropin"] = dict(leg(rd, max(2, K // 2)), traversal=int(rd.get_option("traversal")),
                                        note="the same step on the scene the reference's own loader builds for this XML (SBVH, leaves of up to 16, dumped by "
                                             "`ref_dump dump`): what a host application that keeps the reference's loader gets (INTEGRATION.md)")
                rd.close()
            except Exception as e:
                extras["dropin"] = {"error": str(e)[:200]}

    if world == 1 and not args.no_extras:
        r.sync()
        ps = python_serial_rate(os.path.join(ROOT, "oracle", "_ref", "b200_side"), scene, max(16, K * S // 4), physical_gpu(local_rank))
        if ps:
            extras["python_serial"] = ps

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        tmp = os.path.join(tempfile.gettempdir(), "bench_scene.cpts")
        sc.save(tmp)
        cpu_baseline = cpu_oracle_rate(tmp, W, H, rules=int(r.get_option("rules")))

    out = {"metric": (METRIC.replace("wavefront PT", "volumetric PT") if "vpt" in workload else METRIC).replace("1920x1080", f"{W}x{H}"), "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": Wm,
           "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if tiles else "weak", "vs_baseline": None, "dtype": "f32",
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
                    "cornell box bundled with the reference, boxes switched to the spectral dispersion BSDF (Diamond / Sapphire; BASELINE config 3 stand-in); fixed seeds"
                    if "dispersion" in workload else
                    "vader geometry bundled with the reference (textures/envmap not shipped upstream, removed); fixed seeds"),
           "config": {"workload": workload, "width": W, "height": H, "spp_per_step": S, "max_depth": cfg["max_depth"],
                      "triangles": cfg["num_prims"], "parallelism": (f"{'row-band tile' if tiles else 'sample'}-split x{world}, 1 NCCL reduce of the fp32 accumulator to rank 0 per step, "
                                                                     "issued from a snapshot on a side stream (overlaps the next step; drained inside the timed region)"
                                      if world > 1 else "1 GPU"),
                      "samples_per_step": step_spp, "row_bands_at_the_end": rows_final,
                      "steady_state_note": f"{K} steps x {step_spp} spp = {K * step_spp} spp timed (a steady-state rate; BASELINE's 1024 spp is the same loop run longer)",
                      "l2_note": f"path-state pool {S * W * H * 92 / 1e6:.0f} MB per step >> 126 MB L2 (no flush needed)",
                      "traversal": {0: "ordered pair-node", 1: "reference DFS order", 2: "ordered pair-node, 256-bit node loads", 3: "ordered, warp-vote",
                                    4: "warp-cooperative, wide layout", 5: "per thread, 4-ary wide layout", 6: "per thread, deferred leaves + refill",
                                    7: "stream (deferred leaves, async ray fetch, epilogue per chunk)" + (", shadow rays too" if stream_shadow else "")}.get(traversal, str(traversal))
                                   + (" [chosen automatically from the warm-up steps]" if traversal_auto else ""),
                      "pipeline": ({0: "volumetric megakernel", 1: "volumetric wavefront over slot queues"}.get(pipeline, "volumetric wavefront, dense path records")
                                   if "vpt" in workload else
                                   {0: "fused", 1: "split shadow kernel", 2: "dynamic-fetch traversal", 3: "deferred NEE", 4: "deferred NEE + dynamic fetch",
                                    5: "deferred NEE, dense path records"}.get(pipeline, str(pipeline))),
                      "ray_accounting": "rays = primary + closest-hit + NEE shadow rays of the algorithm (SURVEY.md 8d, one shadow ray per shaded "
                                        "segment whose sampled emitter is not the null emitter); shadow rays that provably cannot change the "
                                        "radiance are not traversed (shadow_traced_frac)"},
           "msamples_per_s": samples / (ms * 1e-3) / 1e6, "rays_per_sample": rays / samples,
           "shadow_traced_frac": (st1["shadow_traced"] - st0["shadow_traced"]) / max(1, st1["shadow_rays"] - st0["shadow_rays"]),
           "traced_mrays": value * (1.0 - ((st1["shadow_rays"] - st0["shadow_rays"]) - (st1["shadow_traced"] - st0["shadow_traced"])) / max(1.0, float(st1["rays"] - st0["rays"]))) if world == 1 else None,
           "gpu_launches": int((st1["launches"] - st0["launches"])), "wall_s": t_wall,
           "clocks": clk, "e2e": e2e, "roofline": roofline, "stages": stages, "cpu_baseline": cpu_baseline}
    out.update(extras)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
