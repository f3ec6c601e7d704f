#!/usr/bin/env python
"""Benchmark of the per-frame hot path (YOLOv8n detect -> DeepSORT track) on B200.

    python bench.py --gpus N --steps K --warmup W            # this build, one rank per GPU
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference path

A step is one pass of the hot path over one batch: one 1080p frame of each of the 64 streams
a GPU owns (BASELINE.json configs[1]; N GPUs own 64*N streams - configs[2] at N = 8 - with no
collective on the data path: weak scaling).  One JSON line is printed by rank 0.
  value     tracked frames/s, whole job, frames already resident in HBM, device-timed (CUDA
            events), max over ranks
  e2e       same metric through the public API with HOST frames: every step copies its frames
            from pinned host memory and reads the track table back
  roofline  the convolution kernel (tcgen05 implicit GEMM): algorithmic FLOPs / event-timed
            kernel time vs the measured bf16 peak
  cpu_baseline  the oracle (CPU restatement of the reference path) on the box's host cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STREAMS_PER_GPU = 64
FRAME_HW = (1080, 1920)
TARGET_DETS = 16.0
RING = 12
METRIC = "tracked_frames_per_sec"
WORKLOAD = "YOLOv8n+DeepSORT-ReID, 64 synthetic 1080p streams per GPU, ~16 tracked detections/frame"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    except Exception:
        return dict(hbm=6650.0, bf16=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline_run(yolo_blob, reid_blob, bias, frames_by_stream, steps, warmup=0):
    """Oracle pipeline (one instance per stream, as one reference process per stream) over
    `steps` time steps of the given host frames.  Returns (frames/s, seconds, cores)."""
    import torch
    from oracle.pipeline import Pipeline
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    pipes = [Pipeline(yolo_blob, reid_blob, yolo_bias=bias) for _ in frames_by_stream]
    n = 0
    t0 = None
    for t in range(warmup + steps):
        if t == warmup:
            t0 = time.perf_counter()
        for s, fr in enumerate(frames_by_stream):
            pipes[s].step(fr[t % len(fr)])
            n += t >= warmup
    dt = time.perf_counter() - t0
    return n / dt, dt, cores


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path (the reference's own runtime,
    TensorRT + its ONNX files, is not available)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch  # noqa: F401
    from ai_camera_b200 import synth
    yolo, reid = synth.make_blobs(synth.blob_dir())
    bias = synth.shifted_class_bias(yolo)
    n_sample = 2  # streams sampled per step (a bounded sample of the 64-stream batch)
    video = synth.SynthVideo(n_sample, FRAME_HW, n_frames=6, device="cpu")
    frames = [[video.ring[t, s].numpy() for t in range(video.n_frames)] for s in range(n_sample)]
    fps, dt, cores = cpu_baseline_run(yolo, reid, bias, frames, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "%d of %d streams per step" % (n_sample, STREAMS_PER_GPU),
                   "runtime": "oracle port: PyTorch-CPU fp32 nets + numpy/scipy tracker"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d streams x %d steps of 1080p frames" % (n_sample, args.steps)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_RESULT = None  # the real stdout; everything else written to file descriptor 1 (NCCL's version banner, library chatter) goes to stderr


def emit(line):
    _RESULT.write(json.dumps(line) + "\n")
    _RESULT.flush()


def main():
    global _RESULT
    sys.stdout.flush()
    _RESULT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)  # stdout carries exactly one JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="aicam", choices=["aicam", "reference"])
    ap.add_argument("--streams", type=int, default=STREAMS_PER_GPU)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    from ai_camera_b200 import _lib, sharding, synth
    from ai_camera_b200.pipeline import TrackingPipeline
    lib = _lib.load()
    S = args.streams
    blob_dir = synth.blob_dir()
    if rank == 0:
        synth.make_blobs(blob_dir)
    if world > 1:
        dist.barrier()
    yolo, reid = synth.make_blobs(blob_dir)
    first_stream, _ = sharding.stream_partition(S * world, world, rank)  # weak scaling: S streams per GPU
    video = synth.SynthVideo(S, FRAME_HW, n_frames=RING, device=dev, first_stream=first_stream)
    pipe = TrackingPipeline(yolo, reid, S, dev, max_tracks=128, max_crops=S * 40)
    delta = synth.DEFAULT_LOGIT_SHIFT
    bias = synth.shifted_class_bias(yolo, delta)
    synth.apply_class_bias(pipe.detector.engine, bias)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident run ----------------------------------------------------------------
    step_no = [0]

    def one_step():
        out = pipe.step(video.frames(step_no[0]))
        step_no[0] += 1
        return out

    pipe.tracker.count_stats = True  # device-side totals of crops / reported tracks (two tiny torch adds per step)
    for _ in range(args.warmup):
        one_step()
    counted0 = lib.aicam_launch_count()
    one_step()
    launches_counted = lib.aicam_launch_count() - counted0  # kernels of ONE step, counted by the library itself
    torch.cuda.synchronize(dev)
    # CUDA graphs: one graph per ring position (the frame pointer is baked into the graph)
    graphs = None
    if not args.no_graph:
        try:
            graphs = []
            stream = torch.cuda.Stream(dev)
            with torch.cuda.stream(stream):
                for k in range(RING):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=stream):
                        pipe.step(video.ring[k])
                    graphs.append(g)
            torch.cuda.synchronize(dev)
        except Exception as e:  # capture unsupported -> eager launches
            sys.stderr.write("cuda graph capture failed (%s); running eager\n" % e)
            graphs = None
            torch.cuda.synchronize(dev)

    def timed_step():
        if graphs is None:
            return one_step()
        graphs[video.index(step_no[0])].replay()
        step_no[0] += 1

    for _ in range(3):
        timed_step()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.aicam_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    pipe.tracker.crop_total.zero_()
    pipe.tracker.track_total.zero_()
    sync_all()
    ev[0].record()
    for k in range(args.steps):
        timed_step()
        ev[k + 1].record()
    torch.cuda.synchronize(dev)
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    launches_eager = launches_counted
    gpu_launches = (lib.aicam_launch_count() - launches0) if graphs is None else launches_eager * args.steps
    crops_per_step = float(pipe.tracker.crop_total.item()) / args.steps      # mean over the timed steps
    tracks_out = float(pipe.tracker.track_total.item()) / args.steps
    overflow = int(pipe.tracker.overflow().any())

    # ---- end to end: host frames in, track tables out, every step -------------------------
    host_ring = [video.ring[k].cpu().pin_memory() for k in range(min(4, RING))]
    dev_in = [torch.empty_like(video.ring[0]) for _ in range(2)]
    T = pipe.tracker.T
    host_out = [torch.empty((S, T, 6), dtype=torch.int32).pin_memory(),
                torch.empty((S, T), dtype=torch.float32).pin_memory(), torch.empty(S, dtype=torch.int32).pin_memory()]
    copy_stream = torch.cuda.Stream(dev)
    e2e_steps = max(4, min(args.steps, 20))

    def e2e_run(n):
        # double buffering: the H2D copy of step k+1 overlaps the compute of step k
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        with torch.cuda.stream(copy_stream):
            dev_in[0].copy_(host_ring[0], non_blocking=True)
            ready[0].record()
        for k in range(n):
            cur, nxt = k & 1, (k + 1) & 1
            if k + 1 < n:
                with torch.cuda.stream(copy_stream):
                    if k >= 1:
                        copy_stream.wait_event(done[nxt])
                    dev_in[nxt].copy_(host_ring[(k + 1) % len(host_ring)], non_blocking=True)
                    ready[nxt].record()
            torch.cuda.current_stream(dev).wait_event(ready[cur])
            ot, oc, on = pipe.step(dev_in[cur])
            host_out[0].copy_(ot, non_blocking=True)
            host_out[1].copy_(oc, non_blocking=True)
            host_out[2].copy_(on, non_blocking=True)
            done[cur].record()
        torch.cuda.synchronize(dev)

    e2e_run(3)
    sync_all()
    # timed on the device: the start event precedes the first copy's stream (the copy stream waits on it), the end
    # event follows the last device-to-host copy on the compute stream
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    copy_stream.wait_event(ee0)
    e2e_run(e2e_steps)
    ee1.record()
    torch.cuda.synchronize(dev)
    e2e_s = ee0.elapsed_time(ee1) * 1e-3
    sync_all()
    h2d = S * FRAME_HW[0] * FRAME_HW[1] * 3
    d2h = sum(t.numel() * t.element_size() for t in host_out)

    # ---- roofline of the convolution kernel (separate, event-instrumented pass) -------------
    lib.aicam_profile_enable(1)
    prof_steps = 6
    prof_crops = 0.0
    for _ in range(prof_steps):
        one_step()
        prof_crops += float(pipe.tracker.crop_count[0].item())  # this step's crops (host sync: profiling pass only)
    prof_crops /= prof_steps
    ms = _lib.C.c_double()
    nl = _lib.C.c_uint64()
    _lib.check(lib.aicam_profile_conv(_lib.C.byref(ms), _lib.C.byref(nl)))
    lib.aicam_profile_enable(0)
    flops_step = S * pipe.detector.engine.flops_per_item() + prof_crops * pipe.tracker.reid.flops_per_item()
    # DRAM bytes of the same kernels for one step, from the committed ncu pass (profiles/r1_step_traffic.json)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_step_traffic_v8.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    conv_ms_step = ms.value / prof_steps
    peaks = measured_peaks()
    achieved_tflops = flops_step / (conv_ms_step * 1e-3) / 1e12 if conv_ms_step > 0 else 0.0
    launches_step = max(1, nl.value // prof_steps)
    # per launch, like `achieved`: DRAM bytes of the convolution launches of one step / their number
    traffic_per_launch = (traffic["dram_bytes"] / traffic["launches"]) if traffic else None

    # ---- gather (the only collective: final stats) -------------------------------------------
    allst = np.asarray(sharding.gather_stats([total_ms, e2e_s, crops_per_step, float(tracks_out), float(overflow)], dev))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    max_ms = float(allst[:, 0].max())
    max_e2e = float(allst[:, 1].max())
    frames_total = world * S * args.steps
    value = frames_total / (max_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "streams_per_gpu": S, "frame": "1080x1920x3 u8",
                   "crops_per_step": float(allst[:, 2].mean()), "tracks_reported_per_step": float(allst[:, 3].mean()),
                   "detections_per_frame": float(allst[:, 2].mean()) / S,
                   "tracker_overflow": bool(allst[:, 4].any()), "l2": "inputs larger than L2 (398 MB of frames per step)",
                   "cuda_graph": graphs is not None, "detector_logit_shift": delta,
                   "weights": "seeded synthetic (no checkpoints offline)"},
        "p50_latency_ms": statistics.median(per_step),
        "clocks": clocks,
        "e2e": {"value": world * S * e2e_steps / max_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "h2d_gbs": h2d * e2e_steps / max_e2e / 1e9,
                "note": "raw 1080p frames from pinned host memory, copy double-buffered against compute; "
                        "bounded by the PCIe host-to-device rate once compute is faster than the copy"},
        "gpu_launches": int(gpu_launches),
        "roofline": {"kernel": "tcgen05 convolution kernels: conv_win_kernel + conv_pair_kernel + reid_stem_pool_kernel "
                               "(all %d launches of a step)" % (nl.value // prof_steps),
                     "bound": "tensor", "achieved": achieved_tflops, "peak": peaks["bf16"], "unit": "TFLOP/s",
                     "frac": achieved_tflops / peaks["bf16"], "traffic": traffic_per_launch, "traffic_unit": "bytes/launch",
                     "traffic_detail": traffic, "peak_source": peaks["source"],
                     "launches_per_step": int(launches_step), "avg_launch_us": 1e3 * conv_ms_step / launches_step,
                     "flops_per_launch": flops_step / launches_step,
                     "kernel_ms_per_step": conv_ms_step, "flops_per_step": flops_step,
                     "crops_per_profiled_step": prof_crops, "profiled_steps": prof_steps},
    }
    if world == 1 and not args.no_cpu_baseline:
        n_sample, cpu_steps = 8, 16  # ~10 s of host work: a bounded sample of the 64-stream workload
        frames = [[video.ring[t, s].cpu().numpy() for t in range(min(RING, 8))] for s in range(n_sample)]
        fps, dt, cores = cpu_baseline_run(yolo, reid, bias, frames, cpu_steps, warmup=1)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "%d streams x %d steps of the same 1080p frames (%.1f s)" % (n_sample, cpu_steps, dt)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
