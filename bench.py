#!/usr/bin/env python
"""bench.py — matched frame-pairs/sec of the pose-estimation hot path (BASELINE.json metric).

    python bench.py                                           # config 3 on one GPU (the default)
    python bench.py --config {1,2,3,4,5} --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1     # CPU reference arm (fp32 torch + cv2)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N ...                 # one rank per GPU, pairs sharded

BASELINE.json configs:
  1  single 640x480 frame vs 512x512 map tile (the reference's CPU-runnable case; 1 pair per step)
  2  1280x720 frame vs 1024x1024 orthophoto tile, 1 pair per call (what PoseNode._pose does per message,
     pose_node.py:191-205): throughput at batch 1 = 1 / latency
  3  batch of 64 such pairs per step, K = 1024 (DEFAULT: the largest single-GPU configuration)
  4  synthetic flyover stream, 8 candidate rasters per frame, device tile-feature cache, frames dealt round-robin
     over the GPUs and, on a GPU, over --inflight contexts with one host thread each (metric: frames/s localised;
     a step = one frame per GPU)
  5  512 pairs per step in total (512 / N per GPU: STRONG scaling), K = 2048, 2000 RANSAC hypotheses

A step = one pass of the hot path (extract x2 + match + PnP/RANSAC + WGS84 tail) over one batch.  `value` =
matched pairs/s (status OK: >= 15 matches and PnP success, pose_node.py:63,299-307) with inputs resident in HBM;
`e2e` = the same through `PoseEstimator.estimate_batch` with pinned HOST buffers, H2D/D2H inside the timed region.

Arithmetic.  The headline (`value`, `e2e`, `roofline`) is the fp32-FAITHFUL mode (`Config(precision=1)`: split-bf16
tensor-core operands, three MMAs per product, fp32 accumulation, fp32 heads and matcher) because the reference's
tensors are fp32 (pose_node.py:254-287); the bf16 fast mode is measured in the same run and reported under
`bf16_fast`.  `accuracy_vs_fp32` compares BOTH modes with the plain fp32 CPU network on the same pairs.
Pairs are independent, so N GPUs run N disjoint shards with no data-path collective; the only collective is the
NCCL broadcast of the weight blob at start-up.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "matched_frame_pairs_per_sec"
UNIT = "pairs/s"
# algorithmic MACs of the dense stack per INPUT pixel; sum = 84 804
MAC_PER_PX = {"1a": 576, "1b": 36864, "2a": 9216, "2b": 9216, "3a": 4608, "3b": 9216, "4a": 2304, "4b": 2304,
              "Pa": 4608, "Pb": 260, "Da": 4608, "Db": 1024}
FLOP_PER_PIXEL = 2.0 * sum(MAC_PER_PX.values())

CONFIGS = {
    1: dict(frame_hw=(480, 640), tile=512, batch=1, keypoints=1024, iters=2048, scaling="weak",
            workload="config 1: single 640x480 frame vs 512x512 map tile, full extract+match+PnP+WGS84 tail"),
    2: dict(frame_hw=(720, 1280), tile=1024, batch=1, keypoints=1024, iters=2048, scaling="weak",
            workload="config 2: 1280x720 frame vs 1024x1024 orthophoto tile, one pair per call, full extract+match+PnP+WGS84 tail"),
    3: dict(frame_hw=(720, 1280), tile=1024, batch=64, keypoints=1024, iters=2048, scaling="weak",
            workload="config 3: batch of 64 pairs (1280x720 frame vs 1024x1024 orthophoto tile), full extract+match+PnP+WGS84 tail"),
    4: dict(frame_hw=(720, 1280), tile=1024, batch=8, keypoints=1024, iters=2048, scaling="strong",
            workload="config 4: synthetic flyover stream, 8 candidate rasters per frame, device tile-feature cache, frames sharded over the GPUs"),
    5: dict(frame_hw=(720, 1280), tile=1024, batch=512, keypoints=2048, iters=2000, scaling="strong",
            workload="config 5: 512-pair batch per step split over the GPUs, 2048-keypoint cap, 2000 RANSAC hypotheses"),
}
MAX_CHUNK = 64   # pairs per library call (workspace: ~0.7 GB per pair in the fp32-faithful mode)


def make_pairs(n, first_seed, frame_hw, tile, ground=None):
    from gisnav_b200 import synth

    ground = synth.ground_texture(4096, 0) if ground is None else ground
    return [synth.make_pair(ground, first_seed + i, frame_hw, tile) for i in range(n)]


def stack(pairs):
    return (np.stack([p.frame for p in pairs]), np.stack([p.tile for p in pairs]), np.stack([p.dem for p in pairs]),
            np.stack([p.k for p in pairs]), np.stack([p.affine for p in pairs]))


# ---- CPU reference path: fp32 torch-CPU SuperPoint-style stack + fp32 matcher head + cv2.solvePnPRansac --------------
def cpu_reference_pair(pair, params, iters, k_cap, quantize=False, want_sets=False):
    """One pair through the reference-style CPU path in fp32 (`quantize=False`: the reference keeps fp32 tensors,
    pose_node.py:254-287).  Returns (ok, camera centre[, keypoint sets, match set])."""
    from oracle import cv2_ref, matcher_ref, nms_ref, sample_ref, superpoint_ref, tail_ref

    feats = []
    for img in (pair.frame, pair.tile):
        s, d = superpoint_ref.forward_dense(img, params, quantize=quantize)
        xy, _ = nms_ref.select_keypoints(s, max_keypoints=k_cap)
        feats.append((xy, sample_ref.sample_descriptors(d, xy, img.shape)))
    _, idx = matcher_ref.match(feats[0][1], feats[1][1], params, 0.5, quantize=bool(quantize) and quantize != "x3")
    sets = None
    if want_sets:
        sets = ([set(map(tuple, f[0].astype(int).tolist())) for f in feats],
                {(tuple(feats[0][0][i].astype(int)), tuple(feats[1][0][j].astype(int))) for i, j in idx.tolist()})
    if len(idx) < 15:
        return (False, None, sets) if want_sets else (False, None)
    r, t, ok, _ = cv2_ref.compute_pose(pair.k, feats[0][0][idx[:, 0]], feats[1][0][idx[:, 1]], pair.dem,
                                       iterations=iters, with_extras=True)
    if not ok:
        return (False, None, sets) if want_sets else (False, None)
    tail = tail_ref.pose_tail(r, t, pair.affine, pair.tile.shape)
    out = (tail is not None, (-r.T @ t).ravel())
    return out + (sets,) if want_sets else out


def cpu_oracle_centre(pair, params, k_cap, iters, seed=0):
    """fp32 CPU network + the library's own deterministic RANSAC restated in C (oracle/pnp_ref.c): the pose the B200
    path must reproduce when its keypoint and match sets equal the fp32 network's."""
    from oracle import matcher_ref, nms_ref, pnp_ref, sample_ref, superpoint_ref

    feats = []
    for img in (pair.frame, pair.tile):
        s, d = superpoint_ref.forward_dense(img, params, quantize=False)
        xy, _ = nms_ref.select_keypoints(s, max_keypoints=k_cap)
        feats.append((xy, sample_ref.sample_descriptors(d, xy, img.shape)))
    _, idx = matcher_ref.match(feats[0][1], feats[1][1], params, 0.5, quantize=False)
    kp_sets = [set(map(tuple, f[0].astype(int).tolist())) for f in feats]
    m_set = {(tuple(feats[0][0][i].astype(int)), tuple(feats[1][0][j].astype(int))) for i, j in idx.tolist()}
    centre = None
    if len(idx) >= 15:
        obj = pnp_ref.points3d(feats[1][0][idx[:, 1]], pair.dem)
        ref = pnp_ref.solve_pnp_ransac(obj, feats[0][0][idx[:, 0]], pair.k, iters=iters, seed=seed)
        if ref["status"] == 0:
            centre = (-ref["r"].T @ ref["t"]).ravel()
    return centre, kp_sets, m_set


def sift_bf_context_pair(pair):
    """The reference's ACTUAL extractor and its VO matcher, for context: cv2.SIFT_create() uncapped
    on both images (pose_node.py:107,122,230) + BFMatcher.knnMatch(k=2) + ratio 0.7 (twist_node.py:248-267) +
    cv2.solvePnPRansac as compute_pose calls it (_shared.py:109-117).  LightGlue (kornia) is not installed."""
    import cv2

    from oracle import cv2_ref

    sift = cv2.SIFT_create()
    kq, dq = sift.detectAndCompute(pair.frame, None)
    kr, dr = sift.detectAndCompute(pair.tile, None)
    if dq is None or dr is None or len(kr) < 2:
        return False, 0, 0
    good = [m for m, n in cv2.BFMatcher().knnMatch(dq, dr, k=2) if m.distance < 0.7 * n.distance]
    if len(good) < 15:
        return False, len(kq) + len(kr), len(good)
    pq = np.float32([kq[m.queryIdx].pt for m in good]); pr = np.float32([kr[m.trainIdx].pt for m in good])
    pr[:, 0] = np.clip(pr[:, 0], 0, pair.tile.shape[1] - 1); pr[:, 1] = np.clip(pr[:, 1], 0, pair.tile.shape[0] - 1)
    _, _, ok, _ = cv2_ref.compute_pose(pair.k, pq, pr, pair.dem, iterations=10, with_extras=True)
    return bool(ok), len(kq) + len(kr), len(good)


def cpu_threads():
    import cv2
    import torch

    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    cv2.setNumThreads(n)
    return n


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


CPU_PATH = ("fp32 torch-CPU SuperPoint-style stack x2 + fp32 dual-softmax head + cv2.solvePnPRansac(iterationsCount=10, "
            "_shared.py:115) + numpy tail (oracle/, kind = port)")


def workload_config(args, cfg):
    return {
        "workload": cfg["workload"], "config": args.config, "frame_hw": list(cfg["frame_hw"]), "tile": cfg["tile"],
        "pairs_per_step_total": cfg["batch"] if cfg["scaling"] == "strong" else None,
        "pairs_per_step_per_gpu": None if cfg["scaling"] == "strong" else cfg["batch"],
        "max_keypoints": cfg["keypoints"],
        "ransac": f"B200 arm: {cfg['iters']} P3P hypotheses + LM refit; reference arm: cv2.solvePnPRansac iterationsCount=10 as the reference passes",
        "matcher_layers": getattr(args, "matcher_layers", 0),
        "l2_policy": "inputs cycle over distinct pre-staged batches; per-step activation traffic (> 1 GB) exceeds the 126 MB L2",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from gisnav_b200 import weights as W

    cfg = CONFIGS[args.config]
    cores = cpu_threads()
    params = W.unpack(W.load())
    per_step = max(1, args.ref_pairs)
    pairs = make_pairs(per_step * (args.steps + args.warmup), 0, cfg["frame_hw"], cfg["tile"])
    it = iter(pairs)
    for _ in range(args.warmup):
        for _ in range(per_step):
            cpu_reference_pair(next(it), params, args.ref_ransac_iters, cfg["keypoints"])
    ok = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(per_step):
            ok += bool(cpu_reference_pair(next(it), params, args.ref_ransac_iters, cfg["keypoints"])[0])
    dt = time.perf_counter() - t0
    value = ok / dt
    sample = (f"{per_step} pair(s)/step x {args.steps} steps of the same synthetic {cfg['frame_hw'][1]}x{cfg['frame_hw'][0]} vs "
              f"{cfg['tile']}x{cfg['tile']} workload (bounded sample of the config's batch)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "cpu_model": cpu_model(), "path": CPU_PATH, "cv2_ransac_iterations": args.ref_ransac_iters},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "matched_fraction": ok / float(per_step * args.steps),
    }
    print(json.dumps(line), flush=True)


# ---- clocks sampler ---------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML every 5 ms; falls back to
    `nvidia-smi -lms 20` if pynvml is unavailable)."""

    def __init__(self, device_index):
        self.idx = device_index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.mode = None

    def _nvml_loop(self):
        import pynvml as nv

        h = nv.nvmlDeviceGetHandleByIndex(self.idx)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
            time.sleep(0.005)

    def _smi_loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20", "-i",
                                 str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        for line in proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.sm.append(float(f[0])); self.max_mhz = float(f[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)
            if self._stop.is_set():
                break
        proc.terminate()

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES remapping: NVML indexes physical GPUs
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    self.idx = int(vis.split(",")[self.idx])
                except (ValueError, IndexError):
                    pass
            self.mode, target = "nvml", self._nvml_loop
        except Exception:
            self.mode, target = "nvidia-smi", self._smi_loop
        self.thread = threading.Thread(target=target, daemon=True)
        self.thread.start()

    def stop(self):
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=5)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.mode}


# ---- per-kernel roofline ------------------------------------------------------------------------------
def kernel_rooflines(prof, steps, px_step, imgs_step, pairs_step, k_cap, iters, precise, peaks):
    """One row per kernel name of the library's event profile: algorithmic work per step, the
    roof that governs it and the achieved fraction.  Tensor peak = MEASURED_PEAKS.json bf16_tflops_sustained
    (/3 for the split-bf16 layers: three MMAs per algorithmic product); HBM peak = hbm_gbs."""
    tensor = float(peaks.get("bf16_tflops_sustained") or 1400.0)
    hbm = float(peaks.get("hbm_gbs") or 6500.0)
    cells = px_step / 64.0
    conv = lambda keys, div: ("tensor", 2.0 * sum(MAC_PER_PX[k] for k in keys) * px_step / 1e12, "TFLOP", tensor / div)  # noqa: E731
    table = {
        "conv_tc:1a+1b": conv(("1a", "1b"), 1), "conv_tc:1b": conv(("1b",), 1), "conv_tc:2a": conv(("2a",), 1), "conv_tc:2b": conv(("2b",), 1),
        "conv_tc:3a": conv(("3a",), 1), "conv_tc:3b": conv(("3b",), 1), "conv_tc:4a": conv(("4a",), 1), "conv_tc:4b": conv(("4b",), 1),
        "conv_tc:Pa": conv(("Pa",), 1), "conv_tc:Da": conv(("Da",), 1),
        "conv_x3:1b": conv(("1b",), 3), "conv_x3:2a": conv(("2a",), 3), "conv_x3:2b": conv(("2b",), 3), "conv_x3:3a": conv(("3a",), 3),
        "conv_x3:3b": conv(("3b",), 3), "conv_x3:4a": conv(("4a",), 3), "conv_x3:4b": conv(("4b",), 3), "conv_x3:Pa": conv(("Pa",), 3),
        "conv_x3:Da": conv(("Da",), 3),
        # HBM-bound byte work: bytes = compulsory reads + writes per step
        "conv1a_kernel": ("hbm", px_step * (1 + 128) / 1e9, "GB", hbm),
        "conv1a_x3_kernel": ("hbm", px_step * (1 + 256) / 1e9, "GB", hbm),
        "score_head_tc": ("hbm", cells * (512 + 256) / 1e9, "GB", hbm),
        "score_head_f32": ("fp32", 2.0 * MAC_PER_PX["Pb"] * px_step / 1e12, "TFLOP", None),
        "softmax_d2s_kernel": ("hbm", cells * (260 + 256) / 1e9, "GB", hbm),
        "nms_r4_kernel": ("hbm", px_step * 4 / 1e9, "GB", hbm),
        "nms_sparse_kernel": ("hbm", px_step * 4 / 1e9, "GB", hbm),
        "nms_compact_kernel": ("hbm", px_step * 4 / 1e9, "GB", hbm),
        "desc_head_tc": ("tensor", 2.0 * imgs_step * k_cap * 4 * 65536 / 1e12, "TFLOP", tensor),
        "desc_head_f32": ("fp32", 2.0 * imgs_step * k_cap * 4 * 65536 / 1e12, "TFLOP", None),
        "project_tc": ("tensor", 2.0 * imgs_step * k_cap * 65536 / 1e12, "TFLOP", tensor),
        "project_f32": ("fp32", 2.0 * imgs_step * k_cap * 65536 / 1e12, "TFLOP", None),
        "hypothesis_kernel": ("latency", None, None, None), "score_kernel": ("fp32", pairs_step * iters * k_cap * 24 / 1e12, "TFLOP", None),
    }
    # the matcher: ONE S = m_a m_b^T per pair is the algorithmic work, whatever the number of passes
    match_flop = 2.0 * pairs_step * k_cap * k_cap * 256 / 1e12
    rows = []
    match_ms = {True: 0.0, False: 0.0}
    k2_ms = 0.0
    for name, (ms, launches) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        ms_step = ms / steps
        row = {"kernel": name, "ms_per_step": ms_step, "launches_per_step": launches / steps}
        if name.startswith("nms_") or name == "topk_kernel":
            k2_ms += ms_step
        if name.startswith("match_rows") or name.startswith("match_pair") or name == "match_collse_kernel":
            is_tc = name.startswith("match_rows_tc") or name.startswith("match_pair") or name == "match_collse_kernel"
            match_ms[is_tc] += ms_step
            row.update({"bound": "tensor" if is_tc else "fp32", "note": "see `matcher S = m_a m_b^T (all passes)`"})
        elif name in table:
            bound, work, unit, peak = table[name]
            row["bound"] = bound
            if work is not None and ms_step > 0:
                ach = work / (ms_step * 1e-3)
                row.update({"work_per_step": work, "work_unit": unit, "achieved": ach, "achieved_unit": unit + "/s"})
                if peak:
                    row.update({"peak": peak, "frac": ach / peak})
        else:
            row["bound"] = "latency"
        rows.append(row)
    if k2_ms > 0:
        work = px_step * 4 / 1e9
        rows.append({"kernel": "K2 NMS + top-K (all kernels)", "ms_per_step": k2_ms, "bound": "hbm", "work_per_step": work, "work_unit": "GB",
                     "achieved": work / (k2_ms * 1e-3), "achieved_unit": "GB/s", "peak": hbm, "frac": work / (k2_ms * 1e-3) / hbm,
                     "note": "algorithmic work = ONE read of the score map; the chain is a histogram, a compaction pass (the only kernel that "
                             "streams the map), three passes over the listed pixels against the L2-resident map and a per-image top-K: "
                             "launch- and latency-bound, not bandwidth-bound"})
    for is_tc, ms_step in match_ms.items():
        if ms_step > 0:
            ach = match_flop / (ms_step * 1e-3)
            row = {"kernel": "matcher S = m_a m_b^T (all passes)", "ms_per_step": ms_step, "bound": "tensor" if is_tc else "fp32",
                   "work_per_step": match_flop, "work_unit": "TFLOP", "achieved": ach, "achieved_unit": "TFLOP/s",
                   "note": "algorithmic work = ONE S per pair; the kernels compute S and S^T in each of two passes and are bound by the "
                           "one-row-per-thread TMEM epilogue (exp / argmax per score), not by the tensor pipe"}
            if is_tc:
                pk = tensor / (3.0 if precise else 1.0)
                row.update({"peak": pk, "frac": ach / pk})
            rows.append(row)
    return rows


# ---- B200 arm: batch configs (1, 2, 3, 5) -------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import gisnav_b200
    from gisnav_b200 import sharding, weights as W

    cfg = CONFIGS[args.config]
    rank, world, local = sharding.env_rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # the only collective: broadcast the weight blob from rank 0 over NVLink
    blob = W.load() if rank == 0 else None
    wt = sharding.broadcast_weights(blob, W.BLOB_BYTES, device=dev)
    torch.cuda.synchronize()
    if args.config == 4:
        return run_stream(args, cfg, rank, world, local, dev, wt)

    strong = cfg["scaling"] == "strong"
    if strong:
        lo, hi = sharding.shard_range(cfg["batch"], rank, world)
        my_pairs = hi - lo
    else:
        my_pairs = args.batch or cfg["batch"]
    chunk = min(MAX_CHUNK, my_pairs)
    hq, wq = cfg["frame_hw"]
    tile = cfg["tile"]
    k_cap, iters = args.keypoints or cfg["keypoints"], args.ransac_iters or cfg["iters"]

    def make_ctx(precision):
        c = gisnav_b200.Config(max_batch=chunk, max_keypoints=k_cap, ransac_iters=iters, max_image_h=max(hq, tile),
                               max_image_w=max(wq, tile), precision=precision)
        ctx = gisnav_b200.Context(c, device=local, weights_device_ptr=wt.data_ptr(), weights_nbytes=W.BLOB_BYTES)
        if args.matcher_layers > 0:
            # the reference matcher's transformer layers (LightGlue n_layers=9, pose_node.py:109-121).  No LightGlue
            # checkpoint is reachable offline: residual-zero init runs the full arithmetic, matches unchanged.
            ctx.set_matcher_layers(W.pack_layers(W.layers_random_init(args.matcher_layers, seed=0, residual_zero=True), args.matcher_layers))
        return ctx

    modes = [("fp32_faithful", 1), ("bf16_fast", 0)] if args.precision == "both" else [(args.precision, 1 if args.precision == "fp32_faithful" else 0)]
    n_sets = 3 if my_pairs <= 64 else 2  # distinct batches cycled through (inputs differ every step)
    from gisnav_b200 import synth

    ground = synth.ground_texture(4096, 0)
    first = (rank * n_sets * my_pairs) if not strong else 0
    sets = []
    for s in range(n_sets):
        seed0 = first + s * my_pairs if not strong else s * cfg["batch"] + lo
        pairs = make_pairs(my_pairs, seed0, cfg["frame_hw"], tile, ground)
        host = stack(pairs)
        dev_t = tuple(torch.from_numpy(a).to(dev) for a in host)
        pinned = tuple(torch.from_numpy(a).pin_memory() for a in host)
        sets.append((pairs, dev_t, pinned))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_mode(name, precision, sampler):
        ctx = make_ctx(precision)
        pe = gisnav_b200.PoseEstimator(ctx)
        stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)

        def timed(run_step, steps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            results = []
            for i in range(steps):
                results.append(run_step(i))
            e1.record(stream)
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item()), results

        def step_device(i):
            _, d, _ = sets[i % n_sets]
            out = []
            for c0 in range(0, my_pairs, chunk):
                sl = slice(c0, min(c0 + chunk, my_pairs))
                out += pe.estimate_batch_device(d[0][sl], d[1][sl], d[2][sl], d[3][sl], d[4][sl])
            return out

        def step_host(i):
            _, _, p = sets[i % n_sets]
            out = []
            for c0 in range(0, my_pairs, chunk):
                sl = slice(c0, min(c0 + chunk, my_pairs))
                out += pe.estimate_batch(p[0][sl].numpy(), p[1][sl].numpy(), p[2][sl].numpy(), p[3][sl].numpy(), p[4][sl].numpy())
            return out

        for i in range(args.warmup):
            step_device(i)
        # short steps (configs 1 / 2: ~1 ms): keep warming until the clocks have had half a second under load, untimed
        t_warm, extra = time.time(), 0
        while time.time() - t_warm < 0.5 and extra < 2000:
            step_device(args.warmup + extra)
            extra += 1
        if sampler is not None:
            sampler.start()
        # headline: no per-kernel event bracketing inside the timed region
        launches0 = ctx.launch_count
        ms, results = timed(step_device, args.steps)
        launches = ctx.launch_count - launches0
        clocks = sampler.stop() if sampler is not None else None
        for i in range(min(args.warmup, 2)):
            step_host(i)
        ms_e2e, res_e2e = timed(step_host, args.steps)
        # separate, shorter pass with every launch bracketed by an event pair (roofline table)
        prof_steps = max(1, min(args.steps, 5))
        ctx.profile(True)
        ctx.profile_read()
        ms_prof, _ = timed(step_device, prof_steps)
        prof = ctx.profile_read()
        ctx.profile(False)
        # batch-1 latency on the same context (what PoseNode._pose pays per message)
        _, d, _ = sets[0]
        for _ in range(3):
            pe.estimate_batch_device(d[0][:1], d[1][:1], d[2][:1], d[3][:1], d[4][:1])
        n_lat = 20
        ms_lat, _ = timed(lambda i: pe.estimate_batch_device(d[0][i % my_pairs:i % my_pairs + 1], d[1][i % my_pairs:i % my_pairs + 1],
                                                            d[2][i % my_pairs:i % my_pairs + 1], d[3][i % my_pairs:i % my_pairs + 1],
                                                            d[4][i % my_pairs:i % my_pairs + 1]), n_lat)
        matched = sum(1 for step in results for r in step if r.ok)
        matched_e2e = sum(1 for step in res_e2e for r in step if r.ok)
        err_gt = []
        for i, step in enumerate(results[:n_sets]):
            for r, p in zip(step, sets[i % n_sets][0]):
                if r.ok:
                    err_gt.append(np.linalg.norm(r.camera_center - (-p.r_gt.T @ p.t_gt).ravel()))
        # accuracy vs the fp32 CPU network on the first `acc_pairs` pairs of set 0 (rank 0, N = 1)
        acc_rows = None
        if rank == 0 and world == 1 and args.acc_pairs > 0 and oracle_rows:
            n_acc = min(len(oracle_rows), my_pairs, chunk)
            _, d0, _ = sets[0]
            res = pe.estimate_batch_device(d0[0][:n_acc], d0[1][:n_acc], d0[2][:n_acc], d0[3][:n_acc], d0[4][:n_acc])
            acc_rows = []
            for j in range(n_acc):
                centre, kp_sets, m_set = oracle_rows[j]
                kq, kr = pe.slot_keypoints(j), pe.slot_keypoints(chunk + j)
                idx = pe.pair_matches(j)
                g_kp = [set(map(tuple, kq.astype(int).tolist())), set(map(tuple, kr.astype(int).tolist()))]
                g_m = {(tuple(kq[a].astype(int)), tuple(kr[b].astype(int))) for a, b in idx.tolist()}
                acc_rows.append({
                    "kp_differ": sum(len(a ^ b) // 2 for a, b in zip(kp_sets, g_kp)),
                    "kp_total": sum(len(a) for a in kp_sets), "matches_fp32": len(m_set), "matches_gpu": len(g_m),
                    "matches_common": len(m_set & g_m),
                    "centre_diff_px": float(np.linalg.norm(res[j].camera_center - centre)) if (centre is not None and res[j].ok) else None})
        out = dict(name=name, precision=precision, ms=ms, ms_e2e=ms_e2e, ms_prof=ms_prof, prof_steps=prof_steps, prof=prof, launches=launches,
                   matched=matched, matched_e2e=matched_e2e, err_gt=err_gt, clocks=clocks, acc_rows=acc_rows,
                   latency_ms=ms_lat / n_lat, extra_warmup_steps=extra)
        ctx.close()
        return out

    # CPU legs on rank 0 at N=1, bounded samples, before the GPU timing
    cpu_baseline = None
    oracle_rows = []
    if rank == 0 and world == 1 and (args.cpu_pairs > 0 or args.acc_pairs > 0):
        cores = cpu_threads()
        params = W.unpack(W.load())
        sample_pairs = sets[0][0]
        if args.cpu_pairs > 0:
            n_cpu = min(args.cpu_pairs, len(sample_pairs))
            cpu_reference_pair(sample_pairs[0], params, args.ref_ransac_iters, k_cap)  # warm-up
            t0 = time.perf_counter()
            ok = sum(bool(cpu_reference_pair(p, params, args.ref_ransac_iters, k_cap)[0]) for p in sample_pairs[:n_cpu])
            dt = time.perf_counter() - t0
            cpu_baseline = {"value": ok / dt, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": cpu_model(), "dtype": "f32",
                            "sample": f"{n_cpu} pairs of the same workload, {dt:.1f} s", "path": CPU_PATH}
            n_ctx = min(args.sift_pairs, len(sample_pairs))
            if n_ctx > 0:
                t0 = time.perf_counter()
                rows = [sift_bf_context_pair(p) for p in sample_pairs[:n_ctx]]
                dt = time.perf_counter() - t0
                cpu_baseline["reference_extractor_context"] = {
                    "path": "cv2.SIFT_create() uncapped x2 + cv2.BFMatcher.knnMatch(k=2) + ratio 0.7 + cv2.solvePnPRansac (pose_node.py:107,122,230; "
                            "twist_node.py:248-267; LightGlue not installed)",
                    "value": sum(r[0] for r in rows) / dt, "unit": UNIT, "pairs_per_sec_processed": n_ctx / dt,
                    "sample": f"{n_ctx} pairs, {dt:.1f} s", "mean_keypoints_per_pair": float(np.mean([r[1] for r in rows])),
                    "mean_ratio_matches": float(np.mean([r[2] for r in rows]))}
        for p in sample_pairs[: min(args.acc_pairs, len(sample_pairs), chunk)]:
            oracle_rows.append(cpu_oracle_centre(p, params, k_cap, iters))

    outs = []
    for i, (name, precision) in enumerate(modes):
        outs.append(run_mode(name, precision, ClockSampler(local) if (rank == 0 and i == 0) else None))

    head = outs[0]
    cnt = torch.tensor([[o["matched"], o["matched_e2e"], o["launches"]] for o in outs], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(my_pairs * args.steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    cnt = cnt.tolist()
    total_all = float(tot.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        px_step = my_pairs * (hq * wq + tile * tile)
        peak = float(peaks.get("bf16_tflops_sustained") or 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s (of fallback)"

        def mode_report(o, c):
            tables = kernel_rooflines(o["prof"], o["prof_steps"], px_step, 2 * my_pairs, my_pairs, k_cap, iters, o["precision"] == 1, peaks)
            conv_ms = sum(r["ms_per_step"] for r in tables if r["kernel"].startswith("conv") or r["kernel"] in
                          ("score_head_tc", "desc_head_tc", "score_head_f32", "desc_head_f32", "softmax_d2s_kernel", "desc_combine_kernel", "desc_cells_kernel"))
            div = 3.0 if o["precision"] == 1 else 1.0
            stack_ach = FLOP_PER_PIXEL * px_step / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else None
            acc = None
            if o["acc_rows"]:
                d = np.array([r["centre_diff_px"] for r in o["acc_rows"] if r["centre_diff_px"] is not None])
                acc = {"pairs": len(o["acc_rows"]), "comparator": "plain fp32 CPU network + fp32 matcher head + the library's RANSAC restated in C (oracle/)",
                       "pose_rmse_px_vs_fp32_oracle": float(np.sqrt(np.mean(d ** 2))) if len(d) else None,
                       "pose_median_px_vs_fp32_oracle": float(np.median(d)) if len(d) else None,
                       "pose_max_px_vs_fp32_oracle": float(d.max()) if len(d) else None,
                       "frac_pairs_within_1e-3_px": float(np.mean(d <= 1e-3)) if len(d) else None,
                       "frac_pairs_within_1e-2_px": float(np.mean(d <= 1e-2)) if len(d) else None,
                       "keypoint_set_overlap": 1.0 - sum(r["kp_differ"] for r in o["acc_rows"]) / max(1, sum(r["kp_total"] for r in o["acc_rows"])),
                       "keypoints_differing_per_pair": float(np.mean([r["kp_differ"] for r in o["acc_rows"]])),
                       "match_set_overlap": sum(r["matches_common"] for r in o["acc_rows"]) / max(1, sum(r["matches_fp32"] for r in o["acc_rows"]))}
            return {
                "dtype": "bf16x3+f32 (split-bf16 tensor-core operands, 3 MMAs per product, fp32 accumulate; fp32 heads and matcher)" if o["precision"] == 1
                         else "bf16 (bf16 operands and stored activations, fp32 accumulate)",
                "value": c[0] / (o["ms"] * 1e-3), "unit": UNIT, "ms_per_step": o["ms"] / args.steps,
                "e2e_value": c[1] / (o["ms_e2e"] * 1e-3), "e2e_ms_per_step": o["ms_e2e"] / args.steps,
                "pairs_per_sec_processed": total_all / (o["ms"] * 1e-3), "matched_fraction": c[0] / total_all if total_all else None,
                "latency_ms_one_pair_per_call": o["latency_ms"],
                "pose_rmse_px_vs_ground_truth": float(np.sqrt(np.mean(np.square(o["err_gt"])))) if o["err_gt"] else None,
                "accuracy_vs_fp32": acc,
                "roofline_stack": {"kernel": "K1 dense stack (all conv / head launches)", "achieved": stack_ach, "unit": "TFLOP/s (algorithmic)",
                                   "peak": peak / div, "frac": stack_ach / (peak / div) if stack_ach else None, "ms_per_step": conv_ms,
                                   "share_of_step": conv_ms / (o["ms_prof"] / o["prof_steps"])},
                "event_profile_ms_per_step": o["ms_prof"] / o["prof_steps"],
                "roofline_per_kernel": tables,
            }

        reports = {o["name"]: mode_report(o, c) for o, c in zip(outs, cnt)}
        hrep = reports[head["name"]]
        # dominant kernel of the headline mode
        dom_name = "conv_x3:1b" if head["precision"] == 1 else "conv_tc:1a+1b"
        dom = next((r for r in hrep["roofline_per_kernel"] if r["kernel"] == dom_name), None) or \
            max((r for r in hrep["roofline_per_kernel"] if r.get("frac")), key=lambda r: r["ms_per_step"])
        launches_dom = dom.get("launches_per_step", 2) or 2
        traffic = tensor_pipe = None
        try:  # dram bytes per launch from the committed `ncu --set full` capture of this kernel, scaled to this batch
            cap = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_dominant.json")))
            row = cap.get(dom["kernel"])
            if row:
                traffic = (row["dram_read_bytes"] + row["dram_write_bytes"]) * (px_step / launches_dom) / row["pixels_per_launch"]
                tensor_pipe = row.get("tensor_pipe_active_pct")
        except (OSError, ValueError, KeyError):
            pass
        in_b = 256 if head["precision"] == 1 else 1
        out_b = 64 if head["precision"] == 1 else 32
        roofline = {
            "bound": "tensor", "kernel": f"{dom['kernel']} (" + ("conv1b + 2x2 pool, split-bf16: conv_tc_halo_kernel<64,2,3,X3>" if head["precision"] == 1
                                                               else "conv1_fused_kernel: im2col -> tcgen05 conv1a -> TMEM -> tcgen05 conv1b -> pool") + ")",
            "achieved": dom["achieved"], "peak": dom["peak"], "unit": "TFLOP/s", "frac": dom["frac"],
            "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read+write, profiles/r02_ncu_dominant.json)",
            "tensor_pipe_active_pct_ncu": tensor_pipe,   # sm__pipe_tensor_cycles_active from the same capture: the clock-independent utilisation
            "algorithmic_flop_per_launch": dom["work_per_step"] * 1e12 / launches_dom,
            "algorithmic_bytes_per_launch": px_step / launches_dom * (in_b + out_b),
            "peak_source": peak_src + (" / 3: three MMAs per algorithmic product in the fp32-faithful mode" if head["precision"] == 1 else ""),
            "launches_per_step": launches_dom, "kernel_ms_per_step": dom["ms_per_step"],
            "share_of_step": dom["ms_per_step"] / hrep["event_profile_ms_per_step"],
        }
        frame_b = hq * wq
        h2d = my_pairs * (frame_b + 2 * tile * tile + 9 * 8 + 12 * 8)
        d2h = my_pairs * 200
        line = {
            "metric": METRIC, "value": hrep["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": hrep["ms_per_step"], "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "bf16x3+f32" if head["precision"] == 1 else "bf16",
            "data": "synthetic (procedural texture, trained-from-scratch weights)",
            "config": dict(workload_config(args, cfg), extra_warmup_steps=head["extra_warmup_steps"],
                           small_batch_streams="batches of <= 2 pairs run the frame and raster chains on two streams"),
            "e2e": {"value": hrep["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": hrep["e2e_ms_per_step"]},
            "gpu_launches": int(cnt[0][2]),
            "clocks": head["clocks"], "roofline": roofline, "cpu_baseline": cpu_baseline,
            "matched_fraction": hrep["matched_fraction"], "pairs_per_sec_processed": hrep["pairs_per_sec_processed"],
            "latency_ms_one_pair_per_call": hrep["latency_ms_one_pair_per_call"],
            "pose_rmse_px_vs_ground_truth": hrep["pose_rmse_px_vs_ground_truth"],
            "accuracy_vs_fp32": hrep["accuracy_vs_fp32"],
            "roofline_stack": hrep["roofline_stack"], "roofline_per_kernel": hrep["roofline_per_kernel"],
            "headline_mode": head["name"], "pairs_per_step_per_gpu": my_pairs, "pairs_per_library_call": chunk,
            "matcher_layers": args.matcher_layers,
        }
        for name, rep in reports.items():
            if name != head["name"]:
                line[name] = rep
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---- B200 arm: config 4, the flyover stream ------------------------------------------------------------------------
def run_stream(args, cfg, rank, world, local, dev, wt):
    """Config 4: frames of a synthetic flyover dealt round-robin over the GPUs (`sharding.frame_owner`); every frame is
    matched against 8 candidate rasters (3x3 neighbourhood of the nearest tile minus the farthest), raster features
    cached on the device by tile id.  A step = one frame per GPU through `PoseEstimator.estimate_candidates` with HOST
    buffers (this is the e2e path; the stream has no device-resident variant: the value IS end to end)."""
    import cv2
    import torch
    import torch.distributed as dist

    import gisnav_b200
    from gisnav_b200 import sharding, synth, weights as W

    TILE, GRID = cfg["tile"], 512
    h, w = cfg["frame_hw"]
    ground = synth.ground_texture(4096, 0)
    f = 0.32 * w
    k = np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1.0]])
    height = 0.6 * TILE * f / w
    n_grid = (4096 - TILE) // GRID + 1
    n_frames = (args.steps + args.warmup) * world
    rng = np.random.default_rng(4)
    start, leg = np.array([1100.0, 1300.0]), np.array([3.0 * np.cos(0.35), 3.0 * np.sin(0.35)])
    yaw0 = np.radians(12.0)
    mine = []
    for i in range(n_frames):
        c = start + leg * i
        angles = (np.radians(rng.uniform(-3, 3)), np.radians(rng.uniform(-3, 3)), yaw0 + np.radians(rng.uniform(-2, 2)))
        noise_seed = int(rng.integers(1 << 30))
        if sharding.frame_owner(i, world) != rank:
            continue
        r = synth.rot_xyz(*angles)
        t = -r @ np.array([[c[0]], [c[1]], [-height]])
        hmat = k @ np.column_stack((r[:, 0], r[:, 1], t[:, 0]))
        img = cv2.warpPerspective(ground, hmat, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
        img = np.clip(img.astype(np.float32) + np.random.default_rng(noise_seed).standard_normal((h, w)).astype(np.float32) * 2.0, 0, 255).astype(np.uint8)
        gx = int(np.clip(round((c[0] - TILE / 2) / GRID), 1, n_grid - 2)); gy = int(np.clip(round((c[1] - TILE / 2) / GRID), 1, n_grid - 2))
        nb = [(gx + dx, gy + dy) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
        nb.sort(key=lambda g: (g[0] * GRID + TILE / 2 - c[0]) ** 2 + (g[1] * GRID + TILE / 2 - c[1]) ** 2)
        mine.append((img, nb[:8], c))
    precision = 0 if args.precision == "bf16_fast" else 1
    # frames are independent: `inflight` contexts per GPU (own stream, workspace and tile cache), one host thread each, take
    # the frames of this rank in turn — host work and the short kernels of one frame overlap with the other frame's
    inflight = max(1, args.inflight)
    fs = gisnav_b200.FrameStream(inflight, gisnav_b200.Config(max_batch=8, max_image_h=1024, max_image_w=1280, tile_cache=48, precision=precision,
                                                              max_keypoints=args.keypoints or cfg["keypoints"],
                                                              ransac_iters=args.ransac_iters or cfg["iters"]),
                                 device=local, weights_device_ptr=wt.data_ptr(), weights_nbytes=W.BLOB_BYTES)
    ctxs = fs.contexts
    streams = [torch.cuda.ExternalStream(c.stream_ptr, device=dev) for c in ctxs]

    def run(i, pe):
        img, nb, c = mine[i]
        # views into the mosaic: estimate_candidates copies only the rasters the device cache does not hold
        tiles = [ground[gy * GRID: gy * GRID + TILE, gx * GRID: gx * GRID + TILE] for gx, gy in nb]
        ids = np.array([gy * n_grid + gx for gx, gy in nb], np.int64)
        affs = np.stack([synth.tile_affine(gx * GRID, gy * GRID) for gx, gy in nb])
        best, res, hits = pe.estimate_candidates(img, tiles, ids, None, k, affs)
        err = None
        if best is not None:
            gx, gy = nb[best]
            cc = res[best].camera_center
            err = float(np.hypot(cc[0] + gx * GRID - c[0], cc[1] + gy * GRID - c[1]))
        return best, hits, err

    def run_range(lo, hi):
        """frames [lo, hi) of this rank through the in-flight contexts (gisnav_b200/stream.py); results in frame order"""
        return fs.map_frames([(lambda pe, i=i: run(i, pe)) for i in range(lo, hi)])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fs.for_each_context(lambda pe: [run(i, pe) for i in range(args.warmup)])   # every context sees the warm-up frames (and caches their rasters)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = fs.launch_count
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in ctxs]
    for (e0, _), st in zip(ev, streams):
        e0.record(st)
    t0 = time.perf_counter()
    rows = run_range(args.warmup, args.warmup + args.steps)
    for (_, e1), st in zip(ev, streams):
        e1.record(st)
    barrier()
    wall = time.perf_counter() - t0
    # the stream is host-driven (tile gathering, cache bookkeeping): take the larger of device and host time, max over ranks
    ms = torch.tensor([max(max(e0.elapsed_time(e1) for e0, e1 in ev), wall * 1e3)], dtype=torch.float64, device=dev)
    cnt = torch.tensor([sum(r[0] is not None for r in rows), sum(r[1] for r in rows), len(rows), fs.launch_count - l0,
                        sum(r[2] ** 2 for r in rows if r[2] is not None), sum(r[2] is not None for r in rows)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        loc, hits, frames, launches, sq, n_err = cnt.tolist()
        secs = float(ms.item()) * 1e-3
        line = {
            "metric": "localised_frames_per_sec (8 candidate pairs per frame)", "value": loc / secs, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16" if precision == 0 else "bf16x3+f32", "data": "synthetic flyover (procedural texture, trained-from-scratch weights)",
            "config": dict(workload_config(args, cfg), frames_in_flight_per_gpu=inflight),
            "e2e": {"value": loc / secs, "unit": "frames/s", "h2d_bytes_per_step": h * w + 8 * (1 - hits / (8.0 * frames)) * TILE * TILE,
                    "d2h_bytes_per_step": 8 * 200, "note": "host buffers in, host results out: value is the end-to-end number"},
            "candidate_pairs_per_sec": 8 * frames / secs, "cache_hit_rate": hits / (8.0 * frames), "frames": int(frames),
            "position_rmse_px_vs_ground_truth": float(np.sqrt(sq / n_err)) if n_err else None,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    fs.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS), help="BASELINE.json config (default 3: 64 pairs per step, one GPU)")
    ap.add_argument("--precision", default="both", choices=["both", "fp32_faithful", "bf16_fast"],
                    help="both (default): headline = fp32-faithful mode, the bf16 fast mode measured in the same run under `bf16_fast`")
    ap.add_argument("--batch", type=int, default=0, help="override pairs per step per GPU (weak-scaling configs)")
    ap.add_argument("--inflight", type=int, default=4, help="config 4: frames in flight per GPU (contexts + host threads)")
    ap.add_argument("--keypoints", type=int, default=0, help="override the config's keypoint cap")
    ap.add_argument("--ransac-iters", type=int, default=0, help="override the config's hypothesis count")
    ap.add_argument("--matcher-layers", type=int, default=0,
                    help="LightGlue transformer layers in front of the assignment head (0 = head only, the north_star "
                         "matcher; 9 = the reference's LightGlueMatcher depth, untrained residual-zero weights)")
    ap.add_argument("--cpu-pairs", type=int, default=8, help="pairs timed on the host cores for cpu_baseline (N=1 only)")
    ap.add_argument("--sift-pairs", type=int, default=4, help="pairs for the SIFT + BF context timing (N=1 only)")
    ap.add_argument("--acc-pairs", type=int, default=32, help="pairs compared with the fp32 CPU network (N=1 only)")
    ap.add_argument("--ref-pairs", type=int, default=2, help="pairs per step for --impl reference")
    ap.add_argument("--ref-ransac-iters", type=int, default=10,
                    help="iterationsCount of the CPU reference's cv2.solvePnPRansac (the reference passes 10, _shared.py:115)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
