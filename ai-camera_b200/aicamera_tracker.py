"""The frame loop of the reference (``/root/reference/src/aicamera_tracker.py``) on the B200 path.

``python -m ai_camera_b200.aicamera_tracker --input clip.mp4 ...`` takes the reference's flags
(:20-67: ``--input --webcam_id --output_dir --output_filename --show_display --no_save
--yolo_engine --reid_engine --conf_thresh --device``) and runs its loop (:169-240): read a frame,
``YOLODetector.detect``, ``DeepSORT.update``, account the time the reference's way
(frames / sum of detect + update wall time, :175,199-207,254-258).

Differences, all in the direction of the hardware:
  * one frame upload per frame: ``detect`` leaves the frame in HBM and ``update`` is handed that
    device tensor (the reference passes ``frame_bgr.copy()`` through a second H2D inside its ReID
    preprocessing);
  * ``--streams N`` (new) runs N sources through ONE batched ``TrackingPipeline`` step per time
    step (the multi-camera form of the loop: per-stream order is kept, the N frames of a time
    step travel together; N = 1 keeps the reference-shaped facades);
  * the annotated video the reference writes (:211-236) is drawn ON THE DEVICE (visualization.py: box outlines,
    labels, status panel on a device copy of the frame the detector already uploaded) and only then copied back for
    ``cv2.VideoWriter`` (this image has no NVENC: the encode stays on the host); next to it the track tuples of every
    frame go to ``<stem>.tracks.jsonl``.  Batched mode writes videos only with ``--save_video`` (one per stream).
"""
import argparse
import json
import time
from pathlib import Path
from typing import Iterable, List, Optional

import numpy as np
import torch

from . import config


def parse_arguments(argv=None) -> argparse.Namespace:
    parser = argparse.ArgumentParser(description="AICamera: Real-time Object Detection & Tracking (B200 path)")
    parser.add_argument("--input", type=str, default=None, action="append",
                        help="Path to input video file (repeat for several streams). If None, tries to use webcam.")
    parser.add_argument("--webcam_id", type=int, default=0, help="Webcam ID to use if --input is not specified.")
    parser.add_argument("--output_dir", type=str, default="outputs", help="Directory to save the output.")
    parser.add_argument("--output_filename", type=str, default=None,
                        help="Name of the output file. If None, generated from input name or timestamp.")
    parser.add_argument("--show_display", action="store_true", help="Accepted for compatibility; no window is opened.")
    parser.add_argument("--no_save", action="store_true", help="Do not save the output video / track tables.")
    parser.add_argument("--save_video", action="store_true", help="Batched mode: also write one annotated video per stream.")
    parser.add_argument("--yolo_engine", type=str, default=str(config.YOLO_ENGINE_PATH),
                        help="Path to the YOLO weight blob (.aicw; replaces the TensorRT engine file).")
    parser.add_argument("--reid_engine", type=str, default=str(config.REID_ENGINE_PATH),
                        help="Path to the ReID (DeepSORT) weight blob (.aicw).")
    parser.add_argument("--conf_thresh", type=float, default=config.YOLO_CONF_THRESHOLD,
                        help="Confidence threshold for YOLO detections.")
    parser.add_argument("--device", type=str, default="cuda:0", help="CUDA device; there is no CPU path.")
    parser.add_argument("--streams", type=int, default=0,
                        help="Batched mode: number of streams per step (inputs are cycled to fill it). 0 = one "
                             "reference-shaped single-stream loop per input.")
    parser.add_argument("--max_frames", type=int, default=0, help="Stop after this many frames per stream (0 = all).")
    args = parser.parse_args(argv)
    if args.device == "cpu":
        parser.error("this build has no CPU path (the reference's TensorRT engines do not run on CPU either)")
    return args


class LoopStats:
    """Timing as the reference accounts it (:175,199-207): per frame, wall time of detect + update."""

    def __init__(self):
        self.frame_ms: List[float] = []

    def add(self, seconds: float):
        self.frame_ms.append(1e3 * seconds)

    @property
    def frames(self):
        return len(self.frame_ms)

    @property
    def total_s(self):
        return sum(self.frame_ms) * 1e-3

    def fps(self):
        return self.frames / self.total_s if self.total_s > 0 else 0.0

    def percentile(self, q):
        return float(np.percentile(self.frame_ms, q)) if self.frame_ms else 0.0

    def summary(self):
        return {"frames": self.frames, "total_s": self.total_s, "avg_fps": self.fps(),
                "p50_ms": self.percentile(50), "p99_ms": self.percentile(99)}


def run_single_stream(frames: Iterable[np.ndarray], detector, tracker, on_frame=None, share_upload=True, stats_out=None) -> LoopStats:
    """The reference's while-loop body (:169-207) for one stream.  ``on_frame(idx, frame, dets, tracks)``
    receives what the loop would draw.  share_upload: hand ``update`` the frame ``detect`` already
    uploaded (one H2D per frame instead of two)."""
    stats = LoopStats()
    if stats_out is not None:
        stats_out.append(stats)  # (callbacks read the running fps from it)
    for idx, frame_bgr in enumerate(frames):
        t0 = time.time()
        det_bboxes, det_scores, det_class_ids, _ = detector.detect(frame_bgr)
        frame_arg = detector.device_frame if share_upload and detector.device_frame is not None else frame_bgr.copy()
        tracked = tracker.update(det_bboxes, det_scores, det_class_ids, frame_arg)
        stats.add(time.time() - t0)
        if on_frame is not None:
            on_frame(idx, frame_bgr, (det_bboxes, det_scores, det_class_ids), tracked)
    return stats


def run_batched(sources: List[Iterable[np.ndarray]], pipeline, on_step=None, max_steps=0, stats_out=None) -> LoopStats:
    """N streams, one ``TrackingPipeline.step`` per time step.  Frames are read on the host (cv2), staged in two
    pinned buffers and uploaded on a copy stream while the previous step computes; per-stream frame order is kept.
    Stops when the first source ends.  Time per step = wall time from "frames of the step are on the host" to
    "track tables are on the host" (the detect + update span of the reference loop)."""
    dev = pipeline.device
    its = [iter(s) for s in sources]
    S = pipeline.n_streams
    assert len(its) == S
    stats = LoopStats()
    if stats_out is not None:
        stats_out.append(stats)
    host = dev_buf = None
    T = pipeline.tracker.T
    out_host = [torch.empty((S, T, 6), dtype=torch.int32).pin_memory(), torch.empty((S, T), dtype=torch.float32).pin_memory(),
                torch.empty(S, dtype=torch.int32).pin_memory()]
    step = 0
    while not max_steps or step < max_steps:
        batch = []
        for it in its:
            f = next(it, None)
            if f is None:
                return stats
            batch.append(f)
        if host is None:
            shape = (S,) + tuple(batch[0].shape)
            host = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
            dev_buf = [torch.empty(shape, dtype=torch.uint8, device=dev) for _ in range(2)]
        h = host[step & 1]
        for s, f in enumerate(batch):
            h[s].numpy()[...] = f
        t0 = time.time()
        d = dev_buf[step & 1]
        d.copy_(h, non_blocking=True)
        ot, oc, on = pipeline.step(d)
        out_host[0].copy_(ot, non_blocking=True)
        out_host[1].copy_(oc, non_blocking=True)
        out_host[2].copy_(on, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        stats.add(time.time() - t0)
        if on_step is not None:
            on_step(step, batch, out_host)
        step += 1
    return stats


def video_frames(path_or_id, max_frames=0):
    import cv2
    cap = cv2.VideoCapture(path_or_id)
    if not cap.isOpened():
        raise RuntimeError("Could not open video source (%s)." % path_or_id)
    n = 0
    try:
        while True:
            ok, frame = cap.read()
            if not ok:
                return
            yield frame
            n += 1
            if max_frames and n >= max_frames:
                return
    finally:
        cap.release()


class AnnotatedWriter:
    """The reference's visualisation + save step (:211-236) with the drawing on the device: ``write(frame_dev, tracks,
    lines)`` draws the tracks and the status panel on a device COPY of the frame (the reference draws on
    ``frame_bgr.copy()``), copies it back and hands it to cv2.VideoWriter (mp4v, as the reference :150-157)."""

    def __init__(self, path, device, fps=config.DEFAULT_OUTPUT_FPS):
        from .visualization import Overlay
        self.path, self.fps, self.device = str(path), fps, device
        self.overlay = Overlay(device)
        self.writer = None

    def write(self, frame_dev, tracked, info_lines):
        import cv2
        from .visualization import Overlay
        vis = frame_dev.reshape(frame_dev.shape[-3:]).clone()
        H, W = int(vis.shape[0]), int(vis.shape[1])
        lines = list(info_lines)
        self.overlay.draw(vis, [self.overlay.track_items(tracked, (H, W))],
                          panels=[(lambda img: Overlay._draw_info_panel(img, lines)) if lines else None])
        if self.writer is None:
            self.writer = cv2.VideoWriter(self.path, cv2.VideoWriter_fourcc(*"mp4v"), self.fps, (W, H))
            if not self.writer.isOpened():
                raise RuntimeError("Could not open video writer for %s" % self.path)
        self.writer.write(vis.cpu().numpy())

    def close(self):
        if self.writer is not None:
            self.writer.release()
            self.writer = None


def tracks_to_rows(table_host, s):
    ot, oc, on = table_host
    return [tuple(int(v) for v in ot[s, k, :5]) + (config.CLASSES[int(ot[s, k, 5])], float(oc[s, k])) for k in range(int(on[s]))]


def main(argv=None):
    args = parse_arguments(argv)
    device = torch.device(args.device)
    inputs = args.input or []
    for p in inputs:
        if not Path(p).exists():
            print(f"Error: Input video file not found: {p}")
            return 1
    names = [Path(p).stem for p in inputs] or [f"webcam_{args.webcam_id}"]
    sources_spec = inputs or [args.webcam_id]
    writer = None
    if not args.no_save:
        out_dir = Path(args.output_dir)
        out_dir.mkdir(parents=True, exist_ok=True)
        stem = Path(args.output_filename).stem if args.output_filename else "%s_tracked_%s" % (names[0], time.strftime("%Y%m%d-%H%M%S"))
        writer = open(out_dir / (stem + ".tracks.jsonl"), "w")
        print(f"Track tables will be saved to: {writer.name}")
    videos = []
    source_name = Path(inputs[0]).name if inputs else f"webcam_{args.webcam_id}"
    try:
        if args.streams > 0:
            from .pipeline import TrackingPipeline
            print("Initializing the batched pipeline (%d streams)..." % args.streams)
            pipe = TrackingPipeline(args.yolo_engine, args.reid_engine, args.streams, device,
                                    conf_threshold=args.conf_thresh)
            srcs = [video_frames(sources_spec[s % len(sources_spec)], args.max_frames) for s in range(args.streams)]

            if writer and args.save_video:
                videos = [AnnotatedWriter(out_dir / ("%s_s%02d.mp4" % (stem, s)), device) for s in range(args.streams)]
            dev_frames = {}
            pipe_step = pipe.step

            def step_and_keep(frames_dev, *a, **k):  # (the step's device frames, for the overlay)
                dev_frames["f"] = frames_dev
                return pipe_step(frames_dev, *a, **k)
            pipe.step = step_and_keep
            stats_ref = []

            def on_step(step, batch, table):
                if writer:
                    tabs = [t.numpy() for t in table]
                    fps = stats_ref[0].fps() * args.streams if stats_ref else 0.0
                    for s in range(args.streams):
                        rows = tracks_to_rows(tabs, s)
                        writer.write(json.dumps({"frame": step, "stream": s, "tracks": rows}) + "\n")
                        if videos:
                            videos[s].write(dev_frames["f"][s], rows, ["AICamera: YOLOv8 + DeepSORT", f"Input: {source_name}", "FPS: %.2f" % fps])
                if (step + 1) % 100 == 0:
                    print(f"Processed {step + 1} steps.")
            stats = run_batched(srcs, pipe, on_step, stats_out=stats_ref)
            frames = stats.frames * args.streams
        else:
            from .deepsort_tracker import DeepSORT
            from .yolo_detector import YOLODetector
            print("Initializing YOLOv8 Detector...")
            det = YOLODetector(engine_path=args.yolo_engine, conf_threshold=args.conf_thresh, device=device)
            print("Initializing DeepSORT Tracker...")
            trk = DeepSORT(reid_model_path=args.reid_engine, device=device)

            if writer:
                videos = [AnnotatedWriter(out_dir / (stem + ".mp4"), device)]
                print(f"Output video will be saved to: {videos[0].path}")
            loop_stats = []

            def on_frame(idx, frame, dets, tracks):
                if writer:
                    writer.write(json.dumps({"frame": idx, "stream": 0, "tracks": tracks}) + "\n")
                    fps = loop_stats[0].fps() if loop_stats else 0.0  # frames / sum of detect + update time, as the reference (:204-207)
                    videos[0].write(det.device_frame, tracks, ["AICamera: YOLOv8 + DeepSORT", f"Input: {source_name}", "FPS: %.2f" % fps])
                if (idx + 1) % 100 == 0:
                    print(f"Processed {idx + 1} frames.")
            stats = run_single_stream(video_frames(sources_spec[0], args.max_frames), det, trk, on_frame, stats_out=loop_stats)
            frames = stats.frames
    finally:
        if writer:
            writer.close()
        for v in videos:
            v.close()
    print("\n--- Processing Summary ---")
    print(f"Total frames processed: {frames}")
    print(f"Total time: {stats.total_s:.2f} seconds")
    print(f"Average FPS: {frames / stats.total_s if stats.total_s > 0 else 0:.2f}")
    print("Latency per %s: p50 %.2f ms, p99 %.2f ms" % ("step" if args.streams > 0 else "frame", stats.percentile(50),
                                                        stats.percentile(99)))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
