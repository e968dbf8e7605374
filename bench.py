#!/usr/bin/env python
"""bench.py -- Mrays/s (primary + secondary) of the path-tracing hot path on N B200s.

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                   (the reference shader source / CPU oracle on the host cores, same config)

Workload (BASELINE.json): N = 1 is configs[2] "C3" -- the ~1M-triangle synthetic scene
(ezrt_b200.scenes.s_1m: 999,692 triangles), 1920x1080, Disney BRDF + Sobol (mode disney_sobol_p5),
2 bounces; the north_star target (>= 1 Gray/s) is quoted on this scene.  One step = one pass of the
hot path over one batch: `--spp-per-step` consecutive display() calls (default 16; the default 16
steps make up C3's 256 spp), accumulating into the same framebuffer with frameCounter advancing.
N > 1 is weak scaling: every GPU owns 1920x1080 pixels' worth of 16x16 tiles of an image that
grows with N (N=4 is C5's 3840x2160), and every step ends with the single NCCL framebuffer gather.

value  : total rays of all ranks / max-over-ranks CUDA-event time of the K timed steps, scene and
         framebuffer resident in HBM (ezrt_render_device on the current stream).
e2e    : the same steps through the host-buffer C ABI (ezrt_render): every step uploads lastFrame
         from pinned host memory and reads the new framebuffer back.
roofline / cpu_baseline: see DESIGN.md "Measurement".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per k_extend_accel launch on the C3 workload (16-frame batch), mean of the three launch
# kinds (camera / bounce-1 / bounce-2), from the committed ncu capture profiles/ncu_extend_r1_summary.md
NCU_DRAM_BYTES_PER_EXTEND_LAUNCH = 1841e6

METRIC = "Mrays/s (primary+secondary)"
UNIT = "Mrays/s"
WORKLOADS = {
    # name: (scene builder name, width, height, mode, max_bounce)
    "c3": ("s_1m", 1920, 1080, 2, 2),   # configs[2]: 1M tris, 1080p, Disney + Sobol (mode disney_sobol_p5)
    "c2": ("s_bunny", 1024, 1024, 0, 2),  # configs[1]: bunny-class 5k tris, 1024^2, diffuse-only (mode diffuse_p3)
    "c4": ("s_1m", 1920, 1080, 3, 2),   # configs[3]: C3 + HDR env-map importance sampling + MIS (mode disney_is_mis_p5)
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ezrt", choices=["ezrt", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--spp-per-step", type=int, default=16)
    ap.add_argument("--frames-per-batch", type=int, default=0)
    ap.add_argument("--traverse", default="accel", choices=["accel", "pruned", "reference"])
    ap.add_argument("--pipeline", default="wavefront", choices=["wavefront", "megakernel"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", default="1920x1080x1", help="oracle sample WxHxSPP used for cpu_baseline and B_ray")
    return ap.parse_args()


def weak_image(width, height, n):
    """Image of N GPUs: pixels per GPU fixed (1: WxH, 2: 2WxH, 4: 2Wx2H, 8: 4Wx2H)."""
    a = b = 1
    k = n
    while k > 1:
        if a <= b:
            a *= 2
        else:
            b *= 2
        k //= 2
    if a * b != n:
        a, b = n, 1
    return width * a, height * b


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 50 ms during the timed region (a default run times ~0.4 s)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name):
    from ezrt_b200 import scenes
    builder, w, h, mode, bounces = WORKLOADS[name]
    t0 = time.time()
    tris, nodes, eye, cam = getattr(scenes, builder)()
    hdr = cache = None
    if mode == 3:  # procedural 2k environment map (the reference's chinese_garden_2k.hdr is not on the GPU box)
        from ezrt_b200 import api
        hdr = scenes.synth_hdr(2048, 1024)
        cache = api.hdr_cache(hdr)
    return dict(tris=tris, nodes=nodes, eye=eye, cam=cam, width=w, height=h, mode=mode, max_bounce=bounces, scene=builder,
                hdr=hdr, cache=cache, build_s=time.time() - t0)


def oracle_sample(wl, sample, traverse, threads=0):
    """Time the CPU oracle on a bounded sample of the workload; returns (Mrays/s, counters, seconds)."""
    from ezrt_b200 import api
    from tests import oracle_binding as oracle  # the CPU baseline leg: the one place bench.py may execute oracle/
    w, h, spp = [int(x) for x in sample.lower().split("x")]
    cfg = api.RenderConfig(width=w, height=h, spp=spp, max_bounce=wl["max_bounce"], mode=wl["mode"], eye=tuple(wl["eye"]),
                           camera_rotate=tuple(wl["cam"]), env_color=(0.35, 0.45, 0.6), traverse=traverse)
    t0 = time.perf_counter()
    img, c = oracle.render(wl["tris"], wl["nodes"], cfg, hdr=wl.get("hdr"), hdr_cache=wl.get("cache"), threads=threads)
    dt = time.perf_counter() - t0
    c["image"] = img
    return c["rays"] / dt / 1e6, c, dt


def reference_shader_sample(wl, sample, check_against=None):
    """Time the REFERENCE'S OWN SHADER SOURCE (oracle/_ref/libezrt_refshader.so: P3/P4/P5 fshader.fsh transpiled to C++
    in the authoring container, oracle/ref_shader/) on the same bounded sample, all host threads.  Returns seconds, or
    None when the library is not there.  A scene without an environment map gets a 1x1 map of the constant colour
    (GL_NEAREST), which is what env_color means to the shader."""
    import numpy as np
    from ezrt_b200 import api
    from tests import refshader_binding as refshader  # CPU baseline leg only, like oracle_binding
    if not refshader.available():
        return None
    w, h, spp = [int(x) for x in sample.lower().split("x")]
    cfg = api.RenderConfig(width=w, height=h, spp=spp, max_bounce=wl["max_bounce"], mode=wl["mode"], eye=tuple(wl["eye"]),
                           camera_rotate=tuple(wl["cam"]), env_color=(0.35, 0.45, 0.6), traverse=1)
    hdr, cache, linear = wl.get("hdr"), wl.get("cache"), True
    if hdr is None:
        hdr, cache, linear = np.array([[[0.35, 0.45, 0.6]]], np.float32), None, False
    t0 = time.perf_counter()
    img = refshader.render(wl["tris"], wl["nodes"], cfg, hdr, cache, hdr_linear=linear)
    dt = time.perf_counter() - t0
    if check_against is not None:  # the port and the reference shader must agree bit for bit
        assert img.tobytes() == check_against.tobytes(), "oracle port and transpiled reference shader disagree"
    return dt


def b_ray(c):
    """Algorithmic bytes per ray on the reference layout: 48 N_node + 72 N_tri + 72 H + 24."""
    return (48.0 * c["n_node"] + 72.0 * c["n_tri"] + 72.0 * c["hits"]) / c["rays"] + 24.0


def run_reference(args, rank, world):
    """--impl reference: the reference's own implementation of the path on the host cores, all threads: the transpiled
    reference shaders (oracle/_ref, kind "reference") when that library travelled here, else the oracle port."""
    if rank != 0:
        return
    wl = build_workload(args.workload)
    cores = os.cpu_count() or 1
    sample = args.cpu_sample
    _, c, _ = oracle_sample(wl, sample, 1)  # ray count of the sample (the shader library does not count; same paths, same image)
    kind = "reference" if reference_shader_sample(wl, "64x36x1") is not None else "port"
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            t_start = time.perf_counter()
        if kind == "reference":
            reference_shader_sample(wl, sample, check_against=c["image"] if i == 0 else None)
        else:
            oracle_sample(wl, sample, 1)
    total_s = time.perf_counter() - t_start
    value = c["rays"] * args.steps / total_s / 1e6
    what = ("reference shader source (P3|P4|P5 fshader.fsh of the mode, transpiled to C++, oracle/_ref)" if kind == "reference" else "oracle port")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_s / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "scene": wl["scene"], "triangles": int(wl["tris"].shape[0]), "mode": wl["mode"],
                   "max_bounce": wl["max_bounce"], "step": what + " renders sample " + sample + " (literal hitBVH traversal)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample + " (WxHxspp) of the workload per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Everything but the one JSON line goes to stderr: libraries (NCCL prints its version banner on stdout, the
    reference shaders' host code chats) must not share the stream the driver parses."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    args = parse_args()
    quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from ezrt_b200 import api
    from ezrt_b200 import dist as ezdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = build_workload(args.workload)
    W, H = weak_image(wl["width"], wl["height"], world)
    t0 = time.perf_counter()
    scene = api.Scene(wl["tris"], wl["nodes"], wl.get("hdr"), wl.get("cache"), device=local_rank)
    upload_ms = 1e3 * (time.perf_counter() - t0)
    C = 3
    n_local = api.partition_pixels(W, H, rank, world)
    traverse = {"accel": api.TRAVERSE_ACCEL, "pruned": api.TRAVERSE_PRUNED, "reference": api.TRAVERSE_REFERENCE}[args.traverse]
    pipeline = api.PIPELINE_WAVEFRONT if args.pipeline == "wavefront" else api.PIPELINE_MEGAKERNEL

    def cfg_for(step, profile=0):
        return api.RenderConfig(width=W, height=H, spp=args.spp_per_step, first_frame=step * args.spp_per_step, max_bounce=wl["max_bounce"],
                                mode=wl["mode"], eye=tuple(wl["eye"]), camera_rotate=tuple(wl["cam"]), env_color=(0.35, 0.45, 0.6),
                                traverse=traverse, pipeline=pipeline, part_rank=rank, part_count=world,
                                frames_per_batch=args.frames_per_batch, profile=profile)

    stream = torch.cuda.current_stream()
    d_fb = torch.zeros(max(1, n_local) * C, dtype=torch.float32, device="cuda")

    gatherer = ezdist.FramebufferGather(W, H, C, rank, world, d_fb.device) if world > 1 else None

    def device_step(step, profile=0):
        scene.render_device(cfg_for(step, profile), d_fb, stream)
        if world > 1:
            return gatherer(d_fb)  # the single NCCL collective of the step
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: device-resident inputs -----------------
    for s in range(args.warmup):
        device_step(s)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays = launches = 0
    ktimes = {"extend": [0.0, 0], "shade": [0.0, 0], "shadow": [0.0, 0], "other": [0.0, 0]}
    ev0.record(stream)
    for s in range(args.steps):
        device_step(args.warmup + s, profile=1)
        # counters are read after the loop would need per-step storage; reading them synchronises this step only
        c = scene.counters()
        rays += c.rays
        launches += c.kernel_launches
        for k, (ms, n) in scene.kernel_times().items():
            ktimes[k][0] += ms
            ktimes[k][1] += n
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms, float(rays), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, rays, launches = float(tmax[0]), float(tsum[1]), float(tsum[2])
    value = rays / (ms * 1e-3) / 1e6

    # ---------------- e2e: host buffers through ezrt_render -----------------
    e2e = None
    if not args.no_e2e:
        host_fb = torch.zeros(max(1, n_local) * C, dtype=torch.float32).pin_memory()
        host_np = host_fb.numpy()
        full_host = torch.zeros(W * H * C, dtype=torch.float32).pin_memory() if (world > 1 and rank == 0) else None
        e_rays = 0

        def e2e_step(step):
            if world == 1:
                scene.render(cfg_for(step), framebuffer=host_np)  # H2D lastFrame (step > 0), kernels, D2H, sync
                return scene.counters().rays
            d_fb.copy_(host_fb, non_blocking=True)                  # H2D lastFrame part
            full = device_step(step)                                # kernels + the single NCCL gather
            host_fb.copy_(d_fb, non_blocking=True)                  # D2H this rank's part
            if full is not None:
                full_host.copy_(full.reshape(-1), non_blocking=True)  # D2H the assembled image on rank 0
            torch.cuda.synchronize()
            return scene.counters().rays

        for s in range(args.warmup):
            e2e_step(s)
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            e_rays += e2e_step(args.warmup + s)
        barrier()
        e_ms = 1e3 * (time.perf_counter() - t0)
        te = torch.tensor([e_ms, float(e_rays)], dtype=torch.float64, device="cuda")
        if world > 1:
            tm = te.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            tsu = te.clone(); dist.all_reduce(tsu, op=dist.ReduceOp.SUM)
            e_ms, e_rays = float(tm[0]), float(tsu[1])
        part_bytes = n_local * C * 4
        e2e = {"value": e_rays / (e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": part_bytes,
               "d2h_bytes_per_step": part_bytes + (W * H * C * 4 if world > 1 else 0), "ms_per_step": e_ms / max(1, args.steps)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline + cpu baseline (rank 0) -----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    cpu_baseline = None
    bray_ref = bray_pruned = None
    means = {}
    if not args.no_cpu_baseline:
        rate, c_ref, dt = oracle_sample(wl, args.cpu_sample, 1)
        bray_ref = b_ray(c_ref)
        _, c_pr, _ = oracle_sample(wl, args.cpu_sample, 0)
        bray_pruned = b_ray(c_pr)
        means = {"n_node": c_ref["n_node"] / c_ref["rays"], "n_tri": c_ref["n_tri"] / c_ref["rays"], "hit_frac": c_ref["hits"] / c_ref["rays"],
                 "n_node_pruned": c_pr["n_node"] / c_pr["rays"], "n_tri_pruned": c_pr["n_tri"] / c_pr["rays"]}
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": "oracle render of %s (WxHxspp) of the workload, reference traversal, %.1f s" % (args.cpu_sample, dt)}
        dt_ref = reference_shader_sample(wl, args.cpu_sample, check_against=c_ref["image"])
        if dt_ref is not None:  # the reference's own shader source compiled for the CPU travelled here: report that one
            cpu_baseline = {"value": c_ref["rays"] / dt_ref / 1e6, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
                            "sample": "reference shader source (fshader.fsh transpiled to C++, oracle/_ref) renders %s (WxHxspp) of the workload, "
                                      "%.1f s; image bit-identical to the oracle port's, which runs it at %.2f %s" % (args.cpu_sample, dt_ref, rate, UNIT)}
    ext_ms, ext_n = ktimes["extend"]
    roofline = None
    if bray_ref and ext_ms > 0:
        rays_rank0 = rays / world  # rank 0's own launches were timed; rays are evenly spread by the tile interleave
        achieved = rays_rank0 * bray_ref / (ext_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_extend_accel" if args.traverse == "accel" else "k_extend", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": NCU_DRAM_BYTES_PER_EXTEND_LAUNCH, "traffic_unit": "bytes/launch (dram read+write, ncu --set full, profiles/ncu_extend_r1_summary.md)",
                    "algorithmic_bytes_per_launch": rays_rank0 * bray_ref / max(1, ext_n), "peak_source": peak_kind, "bytes_per_ray": bray_ref, "bytes_per_ray_pruned_policy": bray_pruned,
                    "traffic_gbs": NCU_DRAM_BYTES_PER_EXTEND_LAUNCH / (ext_ms * 1e-3 / max(1, ext_n)) / 1e9,
                    "traffic_frac_of_peak": NCU_DRAM_BYTES_PER_EXTEND_LAUNCH / (ext_ms * 1e-3 / max(1, ext_n)) / 1e9 / hbm_peak,
                    "ray_means": means, "extend_ms_per_launch": ext_ms / max(1, ext_n), "extend_launches": ext_n,
                    "extend_share_of_step": ext_ms / ms,
                    "note": "achieved/frac = ALGORITHMIC demand bytes of the reference layout and traversal policy (SURVEY 8d: 48 N_node + 72 N_tri + 72 H + 24 per ray, oracle counters) with no cross-ray reuse; the kernel walks its own acceleration tree out of L2/L1, so frac >> 1 is expected. traffic_* = measured DRAM bytes (ncu): the kernel is bound by the L1 data pipe and issue slots, not by HBM (measured_limiter).",
                    "measured_limiter": {"source": "profiles/ncu_extend_r1_summary.md (ncu --set full, three launches of one batch)",
                                         "l1tex_data_pipe_lsu_wavefronts_pct": [71.2, 82.5, 84.7], "issue_active_pct": [72.0, 63.4, 59.8],
                                         "dram_throughput_pct": [3.5, 3.8, 4.7], "launches": ["camera rays", "bounce 1", "bounce 2"]}}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + (" (BASELINE configs[2]: 1M-tri scene, 1920x1080, Disney+Sobol)" if args.workload == "c3" else ""),
                   "scene": wl["scene"], "triangles": int(wl["tris"].shape[0]), "bvh_nodes": int(wl["nodes"].shape[0]),
                   "image": [W, H], "pixels_per_gpu": n_local, "spp_per_step": args.spp_per_step, "mode": wl["mode"], "max_bounce": wl["max_bounce"],
                   "traverse": args.traverse, "pipeline": args.pipeline, "parallelism": "tiles%d" % world,
                   "l2": "inputs larger than L2: scene 124 MB + wavefront state > 600 MB per step vs 126 MB L2",
                   "scene_build_s": round(wl["build_s"], 2), "scene_upload_ms": round(upload_ms, 1)},
        "rays_per_step": rays / max(1, args.steps), "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "kernel_ms": {k: v[0] for k, v in ktimes.items()}, "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
