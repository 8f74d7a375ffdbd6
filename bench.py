#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 backend for yuki's per-pixel rendering hot path.

Metric (BASELINE.json): Msamples/s (and Mrays/s) of the path-traced Cornell box, config[1]:
1024x1024, Path integrator with Russian roulette (max_depth 8), 1024 spp stratified 32x32, on N B200s.
A "step" is one complete render of that film. At N > 1 the reference's spiral tile list is interleaved over the
ranks (tile i -> rank i mod N, no data-path collective) and the film is summed to rank 0 with one NCCL reduce.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference      # the CPU restatement of yuki's renderer on the host cores

`value` is measured with the scene and the film resident in HBM (CUDA events on the renderer's stream, max over
ranks). `e2e` is the same metric through the C-ABI call with host buffers: scene upload, tile/job upload and the
film read-back are all inside its timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries the one JSON line and nothing else: whatever libraries write to file descriptor 1 while the bench runs
# (NCCL's version banner under NCCL_DEBUG=VERSION, NCCL_DEBUG=INFO output) is sent to stderr; emit() writes the line to
# the real stdout.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
_REAL_STDOUT = None  # main() moves file descriptor 1 aside


def emit(line: dict):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(json.dumps(line), flush=True)
    else:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

from yuki_b200 import desc as D  # noqa: E402

WORKLOAD = "cornell-box 1024x1024, Path max_depth 8 (Russian roulette after 3 bounces), 1024 spp stratified 32x32"
RES = (1024, 1024)
SPP_NX = SPP_NY = 32
MAX_DEPTH = 8
CPU_SAMPLE_SPP = 32  # bounded CPU sample: sample indices 0..31 of every pixel (accumulate-mode tiles), ~10 s on 15 host threads
SCENE = "cornell"
# The device-resident leg (`value`, `roofline`) runs on one pipe so that the CUDA-event time of a kernel launch is the kernel's
# own; the end-to-end leg (`e2e`) makes the call a user makes, with the library's default of two overlapping pipes.
VALUE_PIPES = 1


def select_workload(name, spp_side):
    """The default (and the driver's) workload is BASELINE.json configs[1]. `--workload c5` switches to configs[4] (the
    10 M-triangle scene at 3840x2160, Path 8, 64x64 = 4096 spp) for scaling studies; `--spp-side` shortens either."""
    global WORKLOAD, RES, SPP_NX, SPP_NY, SCENE, CPU_SAMPLE_SPP, VALUE_PIPES
    if name == "c5":
        VALUE_PIPES = 0  # scaling studies: the library's default (two pipes) on every leg
        SCENE, RES, SPP_NX, SPP_NY = "terrain", (3840, 2160), 64, 64
        CPU_SAMPLE_SPP = 1  # one sample index of every pixel: 8.3 M samples
        WORKLOAD = "10M-triangle terrain + material objects 3840x2160, Path max_depth 8, 4096 spp stratified 64x64"
    if spp_side:
        SPP_NX = SPP_NY = spp_side
        CPU_SAMPLE_SPP = min(CPU_SAMPLE_SPP, spp_side * spp_side)
        WORKLOAD = WORKLOAD.split(",")[0] + f", Path max_depth 8, {spp_side * spp_side} spp stratified {spp_side}x{spp_side} (--spp-side override)"


def workload(xf):
    from yuki_b200 import scenes
    scene, cam = scenes.terrain_room(xf) if SCENE == "terrain" else scenes.cornell(xf, light="rect", tall_box="glass")
    film = D.FilmSettings(RES, 16)
    sampler = D.SamplerType.stratified(SPP_NX, SPP_NY, jitter=True)
    integ = D.IntegratorType.path(MAX_DEPTH)
    return scene, cam, film, sampler, integ


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture of this command (profiles/<round>/traffic.json, written by scripts/ncu_traffic.py). None when absent."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for rnd in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        f = os.path.join(pdir, rnd, "traffic.json")
        if os.path.exists(f):
            with open(f) as fh:
                best = json.load(fh)
    return best


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t.start()

    def stop(self):
        self._stop.set()
        self._t.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = []
        for name, col in (("hw_slowdown", 2), ("hw_thermal_slowdown", 3), ("sw_thermal_slowdown", 4), ("sw_power_cap", 5)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def cpu_sample_tiles(film, first_sample=0, n=CPU_SAMPLE_SPP):
    """Accumulate-mode tile list covering sample indices [first, first+n) of every pixel (render_manager.rs:135-143)."""
    from oracle import oracle as O
    base = O.film_tiles(film)
    out = []
    for s in range(first_sample, first_sample + n):
        t = base.copy()
        t["sample"] = s
        out.append(t)
    return np.concatenate(out)


def time_cpu(threads=0, first_sample=0):
    """Times the CPU restatement of yuki's renderer (oracle) on a bounded sample of the workload."""
    from oracle import oracle as O
    scene, cam, film, sampler, integ = workload(O.transforms)
    osc = O.OracleScene(scene)
    acc = D.FilmSettings(film.res, film.tile_dim, accumulate=True)
    tiles = cpu_sample_tiles(film, first_sample)
    _, _, st = osc.render(cam, acc, sampler, integ, tiles=tiles, threads=threads)
    return st


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path. The Rust renderer cannot be built in
    this image (no cargo/rustc), so this is the line-by-line C++ restatement (oracle/, kind "port") with the
    reference's threading model: hardware_concurrency()-1 workers popping 16x16 spiral tiles from one queue."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.warmup > 0:
        time_cpu(first_sample=0)  # one untimed pass warms caches/page tables; repeating it W times would only burn minutes
    t_total, samples, rays, shadow, threads = 0.0, 0, 0, 0, 0
    for k in range(args.steps):
        st = time_cpu(first_sample=(k * CPU_SAMPLE_SPP) % max(1, SPP_NX * SPP_NY - CPU_SAMPLE_SPP))
        t_total += st.seconds
        samples += st.samples
        rays += st.ray_count
        shadow += st.shadow_rays
        threads = st.threads
    v = samples / t_total / 1e6
    sample = f"{CPU_SAMPLE_SPP} of the {SPP_NX * SPP_NY} samples of every pixel per step ({samples // args.steps} samples/step), same scene/sampler/integrator"
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "each step is a bounded sample of the workload; throughput is per sample"},
        "mrays_per_s": rays / t_total / 1e6, "mrays_per_s_total": (rays + shadow) / t_total / 1e6,
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def large_scene_leg(api, xf, ctx, peak):
    """Supplementary measurement (N == 1 only, not the contract's `value`): the geometry of BASELINE.json configs[4] — the
    10 M-triangle terrain at 3840x2160, Path max_depth 8 — at 4 spp on one pipe, so that the line also carries the
    closest-hit kernel's roofline on a scene whose BVH (554 MB of records + 480 MB of triangles) cannot live in the caches.
    Same accounting as `roofline`: (32 B x node visits + 36 B x shape tests) / CUDA-event time of the k_trace_closest launches."""
    from yuki_b200 import scenes
    t0 = time.perf_counter()
    scene, cam = scenes.terrain_room(xf)
    dev = api.Scene(ctx, scene)
    t_build = time.perf_counter() - t0
    rn = api.Renderer(ctx)
    film = D.FilmSettings((3840, 2160), 16)
    sampler, integ = D.SamplerType.stratified(2, 2, jitter=True), D.IntegratorType.path(MAX_DEPTH)
    rn.render(dev, cam, film, sampler, integ, pipes=1)  # warm-up
    steps, ms, cms, nodes, tris, rays, shadow, samples, launches = 2, 0.0, 0.0, 0, 0, 0, 0, 0, 0
    for _ in range(steps):
        st = rn.render(dev, cam, film, sampler, integ, pipes=1).stats
        ms += st.device_ms; cms += st.trace_closest_ms; nodes += st.closest_nodes; tris += st.closest_tris
        rays += st.ray_count; shadow += st.shadow_rays; samples += st.samples; launches += st.trace_closest_launches
    n_tris, n_nodes = dev.host.n_tris, dev.host.n_nodes
    dev.close()
    alg = 32 * nodes + 36 * tris
    achieved = alg / (max(cms, 1e-9) / 1e3) / 1e9
    return {"workload": "configs[4] geometry: 10M-triangle terrain + material objects 3840x2160, Path max_depth 8, 4 spp stratified 2x2, one pipe",
            "triangles": n_tris, "bvh_nodes": n_nodes, "host_build_and_upload_s": t_build, "steps": steps,
            "msamples_per_s": samples / (ms / 1e3) / 1e6, "mrays_per_s": rays / (ms / 1e3) / 1e6,
            "mrays_per_s_total": (rays + shadow) / (ms / 1e3) / 1e6,
            "roofline": {"bound": "hbm", "kernel": "k_trace_closest", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes_per_launch": alg / max(launches, 1), "avg_launch_ms": cms / max(launches, 1), "launches": launches,
                         "share_of_step": cms / ms, "nodes_per_ray": nodes / max(rays, 1), "shape_tests_per_ray": tris / max(rays, 1)}}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from yuki_b200 import api, capi, transforms as xf

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the backend has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    scene, cam, film, sampler, integ = workload(xf)
    ctx = api.Context(local)
    stream = torch.cuda.ExternalStream(capi.lib().yk_context_stream(ctx._h), device=torch.device("cuda", local))
    host_scene = api.HostScene(scene)
    dev = api.Scene(ctx, scene, host=host_scene)
    rn = api.Renderer(ctx)
    all_tiles = api.film_tiles(film)
    my_tiles = np.ascontiguousarray(all_tiles[rank::world])  # spiral order interleaved over ranks (render_manager.rs:206-210 TODO)
    n_pix = film.res[0] * film.res[1]
    spp = sampler.samples_per_pixel()
    total_samples = n_pix * spp

    with torch.cuda.stream(stream):
        d_film = torch.zeros(n_pix * 3, dtype=torch.float32, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

        def step():
            flush.zero_()
            d_film.zero_()
            r = rn.render(dev, cam, film, sampler, integ, tiles=my_tiles, device_film_ptr=d_film.data_ptr(), pipes=VALUE_PIPES)
            if world > 1:
                dist.reduce(d_film, dst=0, op=dist.ReduceOp.SUM)  # tiles are disjoint: sum == gather, bit-exact
            return r.stats

        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        agg = {"ray_count": 0, "shadow_rays": 0, "closest_nodes": 0, "closest_tris": 0, "trace_closest_ms": 0.0, "trace_any_ms": 0.0,
               "shade_ms": 0.0, "kernel_launches": 0, "trace_closest_launches": 0, "any_nodes": 0, "any_tris": 0}
        for _ in range(args.steps):
            st = step()
            for k in agg:
                agg[k] += getattr(st, k)
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clk = clocks.stop() if rank == 0 else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        counts = torch.tensor([agg["ray_count"], agg["shadow_rays"], agg["kernel_launches"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        ms_total = float(ms.item())
        value = total_samples * args.steps / (ms_total / 1e3) / 1e6

        # ---- e2e: public API with host buffers; scene upload + job upload + film read-back inside the timed region ----
        film_host = np.zeros((film.res[1], film.res[0], 3), np.float32)
        flat = host_scene.flat
        scene_bytes = flat.n_nodes * 32 + flat.n_tris * (36 + 4 + 4 + 4 + 1)
        h2d = scene_bytes + len(my_tiles) * (12 + 8) + 8 + 2 * 64   # scene arrays, tile list + tile-area prefix sums, camera matrices
        d2h = n_pix * 12
        e2e_steps = max(1, min(args.steps, 2))

        film_pinned = torch.empty(n_pix * 3, dtype=torch.float32).pin_memory() if world > 1 and rank == 0 else None

        def e2e_step():
            d2 = api.Scene(ctx, scene, host=host_scene)       # yk_scene_create: host arrays -> HBM
            if world == 1:
                rn.render(d2, cam, film, sampler, integ, tiles=my_tiles, film_out=film_host)   # jobs H2D, film D2H
            else:  # every rank renders its tiles, the film is assembled on rank 0 over NVLink and read back there
                d_film.zero_()
                rn.render(d2, cam, film, sampler, integ, tiles=my_tiles, device_film_ptr=d_film.data_ptr())
                dist.reduce(d_film, dst=0, op=dist.ReduceOp.SUM)
                if rank == 0:
                    film_pinned.copy_(d_film)
            d2.close()

        e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_value = total_samples * e2e_steps / float(e2e_s.item()) / 1e6

    large = None
    if rank == 0 and world == 1 and not args.no_large_scene and SCENE == "cornell":
        large = large_scene_leg(api, xf, ctx, peaks()[0])

    if rank == 0:
        peak, peak_src = peaks()
        traffic = ncu_traffic()
        alg_bytes = 32 * agg["closest_nodes"] + 36 * agg["closest_tris"]
        launches = max(agg["trace_closest_launches"], 1)
        t_closest = max(agg["trace_closest_ms"], 1e-9) / 1e3
        achieved = alg_bytes / t_closest / 1e9
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "tile_dim": 16, "partition": f"spiral tiles interleaved over {world} rank(s)",
                       "l2": "256 MB L2 flush between steps; wavefront state (several GB per batch) exceeds the 126 MB L2" + ("; the 37-node scene is cache-resident by nature" if SCENE == "cornell" else ""),
                       "wavefront": "up to 16 Mi paths per batch, queue lengths on the device (no host sync per bounce)",
                       "pipes": ("value / roofline: one pipe, so that a kernel's event-bracketed time is its own; e2e: the library default, "
                                 "two pipes overlapping one batch's shading with the other's traversal") if VALUE_PIPES == 1
                       else "library default (two pipes) on every leg"},
            "mrays_per_s": float(counts[0].item()) / (ms_total / 1e3) / 1e6,
            "mrays_per_s_total": float((counts[0] + counts[1]).item()) / (ms_total / 1e3) / 1e6,
            "gpu_launches": int(counts[2].item()),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "includes": "yk_scene_create + yk_render with host film (pinned staging inside the library)" if world == 1 else
                    "per rank yk_scene_create + yk_render of its tiles, NCCL sum-reduce of the film to rank 0, read-back to pinned host memory there"},
            "roofline": {"bound": "hbm", "kernel": "k_trace_closest", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_source": (traffic or {}).get("source"),
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / launches, "avg_launch_ms": 1e3 * t_closest / launches,
                         "launches": launches, "share_of_step": agg["trace_closest_ms"] / ms_total,
                         "note": "rank 0; 32 B per node visit + 36 B per triangle test (SURVEY.md §8d); the 37-node scene is L1/L2-resident, so achieved can exceed the HBM peak"},
            "stage_ms_per_step": {k: v / args.steps for k, v in (("trace_closest", agg["trace_closest_ms"]), ("trace_any", agg["trace_any_ms"]),
                                                                 ("shade", agg["shade_ms"])) if v > 0},  # any / shade only with YK_STAGE_TIMING=2
        }
        # CPU baseline (rank 0, N == 1 only): bounded sample of the same workload on the host cores.
        if world == 1 and not args.no_cpu_baseline:
            st = time_cpu()
            sample = f"{CPU_SAMPLE_SPP} of the {spp} samples of every pixel ({st.samples} samples), same scene/sampler/integrator"
            line["cpu_baseline"] = {"value": st.samples / st.seconds / 1e6, "unit": "Msamples/s", "cores": st.threads, "kind": "port",
                                    "sample": sample, "seconds": st.seconds}
        if large is not None:
            line["large_scene"] = large
        emit(line)
    dev.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-large-scene", action="store_true", help="skip the supplementary 10 M-triangle roofline leg (N == 1 only)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"], help="c2 = BASELINE.json configs[1] (default, the contract's line); c5 = configs[4]")
    ap.add_argument("--spp-side", type=int, default=0, help="override the stratified grid side (spp = side^2); the line's config says so")
    args = ap.parse_args()
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    select_workload(args.workload, args.spp_side)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
