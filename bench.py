#!/usr/bin/env python
"""bench.py — clips/s of the clip-transform hot path on N B200s (one process per GPU).

A "step" is one pass of the fused transform over one batch of synthetic dashcam
clips.  Default workload = BASELINE.json configs[1]: train-time augmentation,
32 clips x 16 frames x 720x1280 uint8 -> [32,3,16,224,224] bf16 per GPU
(weak scaling: every rank transforms its own 32-clip shard, no collective on
the data path).  See DESIGN.md "Measurement".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mode custom|train|val] [--impl reference]
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (clips per GPU, T, H, W, crop_size)
    "cfg2": (32, 16, 720, 1280, 224),
    "cfg3": (32, 32, 720, 1280, 320),
    "tiny": (4, 4, 180, 320, 112),
    "cfg2s": (8, 16, 720, 1280, 224),   # a trainer-sized batch (cfg5 uses 8 clips per GPU): band-count tuning
    # BASELINE configs[3]: one 40 s x 30 fps 720p video, val chain, every frame transformed once,
    # 16-frame windows at stride 8 as strided views (149 windows); "clips" = windows
    "cfg4": (1, 1200, 720, 1280, 224),
}
# the live train call site nexar_videos.py:2003-2010
KW = {
    "custom": dict(mode="train", enable_custom_augmentation=True, brightness_range=(0.9, 1.1),
                   contrast_range=(0.9, 1.1), saturation_range=(0.9, 1.1), rotation_range=(-5, 5)),
    "train": dict(mode="train"),      # nexar_videos.py:936
    "val": dict(mode="val"),          # nexar_videos.py:947, nexar_inference.py:218
}
METRIC = "augmented clips/sec and achieved HBM GB/s (% of roofline) at 1/2/4/8 B200 vs host CPU"


def algorithmic_bytes_per_clip(t, h, w, cs, out_bytes):
    """uint8 source read once + output written once."""
    return t * h * w * 3 + 3 * t * cs * cs * out_bytes


def ncu_traffic(mode):
    """dram bytes per launch of the resize kernel from the committed `ncu --set full` capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(mode)
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def count(self, t0, t1):
        return sum(1 for ts, _ in list(self.rows) if t0 - 0.05 <= ts <= t1 + 0.15)

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, cmax = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            mx = max(mx, cmax)
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        if not sm:
            sm = [float(l.split(",")[0]) for _, l in self.rows[-3:] if l and l.split(",")[0].strip().replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_setup(mode, cs):
    from oracle import np_oracle as O
    kw = KW[mode]
    aug = O.AugConfig(rotation_range=(-5, 5)) if kw.get("enable_custom_augmentation") else O.AugConfig()
    return O.TransformConfig(mode=kw["mode"], crop_size=cs,
                             enable_custom_augmentation=bool(kw.get("enable_custom_augmentation")), aug=aug)


def cpu_port_clips_per_s(mode, t, h, w, cs, n_clips, threads):
    """Time the CPU port of the reference transform (oracle/torch_port.py) on ``n_clips`` clips."""
    from oracle import torch_port as P
    from vision_collision_detection_b200.synth import make_clip_np
    torch.set_num_threads(threads)
    cfg = oracle_setup(mode, cs)
    clip = torch.from_numpy(make_clip_np(t, h, w, 0, "dashcam")).permute(3, 0, 1, 2)
    random.seed(1234)
    P.clip_transform(clip[:, :2], cfg, random)          # warm-up (thread pool, allocator)
    t0 = time.perf_counter()
    for _ in range(n_clips):
        P.clip_transform(clip, cfg, random)
    dt = time.perf_counter() - t0
    return n_clips / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (the oracle port) on the host cores, rank 0 only."""
    if rank != 0:
        return
    b, t, h, w, cs = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    sample = max(1, args.ref_clips)
    cps_runs = []
    per_clip = None
    for _ in range(max(1, args.warmup)):
        cps1, _ = cpu_port_clips_per_s(args.mode, t, h, w, cs, 1, cores)
        per_clip = 1.0 / cps1
    # bounded sample: the whole --steps run stays within about two minutes whatever K the driver passes
    sample = max(1, min(sample, int(120.0 / (max(1, args.steps) * per_clip))))
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        cps, _ = cpu_port_clips_per_s(args.mode, t, h, w, cs, sample, cores)
        cps_runs.append(cps)
    wall = time.perf_counter() - t_all0
    value = float(np.mean(cps_runs))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}:{args.mode} {t}x{h}x{w} u8 -> {cs}x{cs} f32 on host CPU"},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} clip(s) of {t}x{h}x{w} per step, torch threads={cores}"},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_cfg4(args, rank, world, dev, out_dtype):
    """Inference windowing (BASELINE configs[3]): per-frame-once val transform of a 1200-frame 720p video."""
    import torch.distributed as dist
    from vision_collision_detection_b200.inference import SlidingWindowTransform, sliding_window_starts
    from vision_collision_detection_b200.synth import make_clip_torch
    _, n, h, w, cs = WORKLOADS["cfg4"]
    video = torch.cat([make_clip_torch(100, h, w, seed=rank * 100 + i, kind="dashcam", device=dev) for i in range(n // 100)])
    sw = SlidingWindowTransform(window=16, stride=8, out_dtype=out_dtype)
    for _ in range(max(3, args.warmup)):
        view = sw.windows(video)
    torch.cuda.synchronize()
    steps = min(args.steps, 50)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        view = sw.windows(video)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        peaks, kind = measured_peaks()
        k = len(sliding_window_starts(n, 16, 8))
        bytes_alg = n * h * w * 3 + n * 3 * cs * cs * (2 if out_dtype == torch.bfloat16 else 4)
        ach = bytes_alg / (ms * 1e-3) / 1e9
        print(json.dumps({
            "metric": METRIC, "value": world * k / (ms * 1e-3), "unit": "windows/s", "n_gpus": world, "steps": steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cfg4: {n} frames {h}x{w} u8 (40 s x 30 fps) -> val chain {cs}x{cs} {args.out_dtype}, every frame once, "
                                   f"{k} windows of 16 at stride 8 as strided views", "windows_shape": list(view.shape)},
            "ms_per_video": ms,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                         "peak_kind": kind, "algorithmic_bytes_per_launch": bytes_alg, "traffic": None},
            "gpu_launches": 2 * steps}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="custom", choices=sorted(KW))
    ap.add_argument("--out-dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--ref-clips", type=int, default=8, help="clips per step for --impl reference")
    ap.add_argument("--cpu-clips", type=int, default=128, help="clips for the cpu_baseline leg (about 10-20 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch.distributed as dist
    from vision_collision_detection_b200 import create_video_transforms
    from vision_collision_detection_b200.engine import get_engine
    from vision_collision_detection_b200.host_pipeline import HostClipPipeline
    from vision_collision_detection_b200.synth import make_clip_torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    b, t, h, w, cs = WORKLOADS[args.workload]
    out_dtype = torch.bfloat16 if args.out_dtype == "bf16" else torch.float32
    if args.workload == "cfg4":
        return run_cfg4(args, rank, world, dev, out_dtype)
    tf = create_video_transforms(**KW[args.mode], crop_size=cs, out_dtype=out_dtype)
    eng = get_engine(dev)
    if os.environ.get("NEXAR_FAST_BANDS"):
        from vision_collision_detection_b200 import _lib as _l
        _l.lib().nexar_set_fast_bands(int(os.environ["NEXAR_FAST_BANDS"]))

    # synthetic device-resident shard (1.4 GB for cfg2: larger than the 126 MB L2, so every step streams from HBM)
    clips = torch.stack([make_clip_torch(t, h, w, seed=rank * 1000 + i, kind="dashcam", device=dev) for i in range(b)])
    out = torch.empty((b, 3, t, cs, cs), dtype=out_dtype, device=dev)
    random.seed(1234 + rank)
    param_sets = [tf.sample_params(b, h, w) for _ in range(4)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        tf.forward_batch(clips, params=param_sets[i % len(param_sets)], out=out)

    # ---- device-resident timing ("value") ------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None   # started before the warm-up: nvidia-smi needs ~0.2 s to stream
    # warm-up: the W steps asked for, and at least 10 steps / 50 ms so that the clocks have ramped and the one-time
    # costs (plan tables, workspace, pinned parameter ring) are behind us even when W is tiny; "warmup" reports the count
    t_w = time.time()
    n_warm = 0
    while n_warm < max(args.warmup, 10) or (time.time() - t_w < 0.05 and n_warm < 1000):
        step(n_warm)
        n_warm += 1
        if n_warm % 8 == 0:
            torch.cuda.synchronize()
    args.warmup = n_warm
    barrier()
    from vision_collision_detection_b200 import _lib
    _lib.lib().nexar_profile_begin(args.steps + 8)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    t_wall1 = time.time()
    k1_ms = _lib.profile_end()
    launches = eng.last_launches * args.steps
    ms_total = e0.elapsed_time(e1)
    if sampler is not None and sampler.proc is not None and sampler.count(t_wall0, t_wall1) < 3:
        # a short timed region (small --steps) can fall between two 100 ms samples: keep the SAME load running, untimed,
        # until the sampler has seen it (the clocks line then says how many samples came from the timed region itself)
        in_region = sampler.count(t_wall0, t_wall1)
        t_probe = time.time()
        while time.time() - t_probe < 0.8 and sampler.count(t_wall0, time.time()) < 3:
            for i in range(16):
                step(i)
            torch.cuda.synchronize()
        clocks = sampler.stop(t_wall0, time.time())
        if clocks is not None:
            clocks["samples_in_timed_region"] = in_region
    else:
        clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    tms = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms.item()) / args.steps
    value = world * b / (ms_step * 1e-3)

    # ---- end to end: pinned host clips -> H2D -> transform -> D2H (through the public host API) ----
    e2e = None
    if not args.no_e2e:
        pipe = HostClipPipeline(tf, n_clips=b, frames=t, height=h, width=w, device=dev)
        host_in = pipe.pinned_input()
        host_in.copy_(clips.cpu())
        e2e_steps = max(3, min(args.steps, 8))
        host_in2 = pipe.pinned_input()          # two input batches alternate: one is being copied while the next is "decoded"
        host_in2.copy_(host_in)
        ins = [host_in, host_in2]
        for i in range(2):
            pipe.run(ins[i & 1], params=param_sets[i % len(param_sets)])
        barrier()
        ts0 = time.perf_counter()
        e0.record()
        prev = None
        for i in range(e2e_steps):                # steady state: batch i+1 is submitted before batch i is collected
            ticket = pipe.submit(ins[i & 1], params=param_sets[i % len(param_sets)])
            if prev is not None:
                host_out = pipe.wait(prev)
            prev = ticket
        host_out = pipe.wait(prev)
        e1.record()
        barrier()
        wall = time.perf_counter() - ts0
        ems = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": world * b / (float(ems.item()) / e2e_steps * 1e-3), "unit": "clips/s",
               "h2d_bytes_per_step": int(host_in.numel()), "d2h_bytes_per_step": int(host_out.numel() * host_out.element_size()),
               "steps": e2e_steps}
        del pipe, host_in, host_in2, ins

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        per_clip = algorithmic_bytes_per_clip(t, h, w, cs, out.element_size())
        # dominant kernel: the resize pass (K1).  Its algorithmic bytes: source read once + what it writes
        # (final output when no clip is augmented, else the fp32 RGBX intermediate of the content box).
        k1_ms_avg = (sum(k1_ms) / len(k1_ms)) if k1_ms else None
        k1_bytes = b * per_clip
        roof = None
        if k1_ms_avg:
            ach = k1_bytes / (k1_ms_avg * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "resize (K1)", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "frac_of_8TBs_nominal": ach / 8000.0, "peak_kind": peak_kind,
                    "kernel_ms": k1_ms_avg, "kernel_share_of_step": k1_ms_avg / ms_step,
                    "algorithmic_bytes_per_launch": k1_bytes, "traffic": ncu_traffic(args.mode) if (args.workload == "cfg2" and args.out_dtype == "bf16") else None,
                    "step_achieved_GBs": b * per_clip / (ms_step * 1e-3) / 1e9}
        cpu = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cps, dt = cpu_port_clips_per_s(args.mode, t, h, w, cs, args.cpu_clips, cores)
            cpu = {"value": cps, "unit": "clips/s", "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_clips} clips of {t}x{h}x{w} ({dt:.1f} s), oracle/torch_port.py, torch threads={cores}"}
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}:{args.mode} {b} clips/GPU x {t}x{h}x{w} u8 -> {cs}x{cs} {args.out_dtype}",
                       "kwargs": "nexar_videos.py:2003-2010" if args.mode == "custom" else args.mode,
                       "l2": "input 1.4 GB/step > 126 MB L2 (no flush needed)", "sharding": f"{b} clips per GPU, no collective"},
            "output_GBs": world * b * 3 * t * cs * cs * out.element_size() / (ms_step * 1e-3) / 1e9,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
