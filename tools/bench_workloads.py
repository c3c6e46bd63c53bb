"""Frames/s of the other BASELINE.json workloads on ONE B200 (bench.py is the contract line for configs[1]):
  mot : configs[2] — ConvNeXt-L MOT detector + embedding path at 1536x2048 (mode="whole", NMS, embedding sampling), then
        (a) the QDTrack association of the reference's MOT evaluator on the model's own detections and
        (b) ByteTrack association on 100 synthetic objects per frame (random weights give few detections of their own).
  vos : configs[3] — ConvNeXt-L + CondInst mask head at 800x1280, n objects propagated from the first frame.
Eager launches (no CUDA graph: the association step returns to the host every frame), wall clock around synchronised
steps, synthetic video, seeded weights.  usage: bench_workloads.py mot|vos [frames]"""
import json, os, sys, time, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unicorn_b200 import _lib
from unicorn_b200.engine import UnicornEngine
from unicorn_b200.synthetic import make_video, make_detections
from unicorn_b200.weights import make_state_dict

what = sys.argv[1] if len(sys.argv) > 1 else "mot"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = "cuda"


def timed(fn, n, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    l0, t0 = _lib.LAUNCHES, time.perf_counter()
    for i in range(n):
        fn(warm + i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return n / dt, 1e3 * dt / n, (_lib.LAUNCHES - l0) // n


if what == "mot":
    from unicorn_b200.mot import UnicornMOTTracker
    from unicorn_b200.tracker.byte_tracker import BYTETracker
    H, W = 1536, 2048
    cfg = "unicorn_track_large_mot_challenge"
    eng = UnicornEngine(make_state_dict(cfg, 0), cfg)
    frames, _ = make_video(4, H, W, seed=0, n_obj=6)
    frames = [f[None].to(dev) for f in frames]
    args = types.SimpleNamespace(track_thresh=0.5, track_buffer=30, match_thresh=0.8, mot20=False)
    dets100 = make_detections(n_frames=n + 8, n_obj=100, seed=3, W=float(W), H=float(H))

    def pipelined(trk, extra=None):
        """submit(t+1); collect(t): host association of frame t overlaps the device work of frame t+1"""
        trk.submit(frames[0])
        def step(i):
            trk.submit(frames[(i + 1) % 4])
            trk.collect()
            if extra:
                extra(i)
        r = timed(step, n, warm=5)
        trk.collect()
        return r

    trk = UnicornMOTTracker(eng, (H, W))
    fps, ms, launches = timed(lambda i: trk.step_tensor(frames[i % 4]), n)
    print(json.dumps({"workload": "configs[2] MOT 1536x2048 ConvNeXt-L: detector + embedding + QDTrack association, sequential eager",
                      "frames_per_s": round(fps, 2), "ms_per_frame": round(ms, 2), "kernels_per_frame": launches, "n_gpus": 1}))
    trk = UnicornMOTTracker(eng, (H, W), use_graph=True)
    fps, ms, launches = pipelined(trk)
    print(json.dumps({"workload": "configs[2] same, device half as CUDA graphs, association of frame t overlapped with frame t+1",
                      "frames_per_s": round(fps, 2), "ms_per_frame": round(ms, 2), "kernels_per_frame": launches, "n_gpus": 1}))
    # ByteTrack arm: detector (no embedding branch) + BYTETracker.update.  Seeded random weights give only a handful of
    # detections, so a second tracker is fed 100 synthetic objects per frame inside the same loop: the measured rate
    # includes the host cost of a 100-object association while the device runs the next frame.
    bt100 = BYTETracker(args, device=dev)
    trk = UnicornMOTTracker(eng, (H, W), assoc="byte", tracker=BYTETracker(args, device=dev), use_graph=True)
    fps, ms, launches = pipelined(trk, extra=lambda i: bt100.update(dets100[i][0].numpy(), (H, W), (H, W)))
    print(json.dumps({"workload": "configs[2] MOT 1536x2048 ConvNeXt-L detector (CUDA graphs) + ByteTrack association of 100 synthetic "
                                  "objects per frame, pipelined", "frames_per_s": round(fps, 2), "ms_per_frame": round(ms, 2),
                      "kernels_per_frame": launches, "n_gpus": 1}))
    bt = BYTETracker(args, device=dev)
    fps2, ms2, l2 = timed(lambda i: bt.update(dets100[i][0].numpy(), (H, W), (H, W)), n)
    print(json.dumps({"workload": "configs[2] ByteTrack update alone, 100 synthetic objects per frame (host Kalman + LAP, IoU on the GPU)",
                      "frames_per_s": round(fps2, 1), "ms_per_frame": round(ms2, 3), "kernels_per_frame": l2}))
else:
    from unicorn_b200.vos import UnicornVOSTrack
    H, W = 800, 1280
    cfg = "unicorn_track_large_mask"
    eng = UnicornEngine(make_state_dict(cfg, 0), cfg)
    for n_obj in (1, 3):
        frames, boxes = make_video(4, H, W, seed=1, n_obj=n_obj)
        frames = [f[None].to(dev) for f in frames]
        trk = UnicornVOSTrack(eng, (H, W))
        trk.initialize_tensor(frames[0], {o + 1: boxes[0, o].tolist() for o in range(n_obj)})
        fps, ms, launches = timed(lambda i: trk.track_tensor(frames[1 + i % 3]), n)
        print(json.dumps({"workload": f"configs[3] VOS 800x1280 ConvNeXt-L + CondInst mask head, {n_obj} object(s) (eager)",
                          "frames_per_s": round(fps, 2), "ms_per_frame": round(ms, 2), "kernels_per_frame": launches, "n_gpus": 1}))
