#!/usr/bin/env python
"""bench.py — matched frame-pairs/sec of the pose-estimation hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 10 --warmup 3            # B200 arm
    python bench.py --impl reference --steps 3 --warmup 1     # CPU reference arm (torch + cv2)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N ...                 # one rank per GPU, pairs sharded

Workload (config 2 of BASELINE.json): synthetic 1280x720 query frame vs 1024x1024 orthophoto
raster, full extract + match + PnP/RANSAC + WGS84 tail; a step = one pass over a batch of
`--batch` pairs per GPU.  `value` = matched pairs/s (status OK: >= 15 matches and PnP success,
pose_node.py:63,299-307) with inputs resident in HBM; `e2e` = the same through the public API
with pinned HOST buffers, H2D/D2H inside the timed region.  Pairs are independent, so N GPUs run N
disjoint shards with no data-path collective (weak scaling); the only collective is the NCCL
broadcast of the weight blob at start-up.
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

FRAME_HW = (720, 1280)
TILE = 1024
METRIC = "matched_frame_pairs_per_sec"
UNIT = "pairs/s"
# algorithmic FLOPs of the dense stack per input pixel: 84 804 MAC
FLOP_PER_PIXEL = 169608.0


def make_pairs(n, first_seed=0):
    from gisnav_b200 import synth

    ground = synth.ground_texture(4096, 0)
    return [synth.make_pair(ground, first_seed + i, FRAME_HW, TILE) for i in range(n)]


def stack(pairs):
    return (np.stack([p.frame for p in pairs]), np.stack([p.tile for p in pairs]), np.stack([p.dem for p in pairs]),
            np.stack([p.k for p in pairs]), np.stack([p.affine for p in pairs]))


# ---- CPU reference path (the oracle: torch-CPU SuperPoint stack + matcher head + cv2.solvePnPRansac) ----
def cpu_reference_pair(pair, params, iters, k_cap):
    """One pair through the reference-style CPU path; returns (ok, camera centre)."""
    from oracle import cv2_ref, matcher_ref, nms_ref, sample_ref, superpoint_ref, tail_ref

    feats = []
    for img in (pair.frame, pair.tile):
        s, d = superpoint_ref.forward_dense(img, params)
        xy, _ = nms_ref.select_keypoints(s, max_keypoints=k_cap)
        feats.append((xy, sample_ref.sample_descriptors(d, xy, img.shape)))
    _, idx = matcher_ref.match(feats[0][1], feats[1][1], params, 0.5)
    if len(idx) < 15:
        return False, None
    r, t, ok, _ = cv2_ref.compute_pose(pair.k, feats[0][0][idx[:, 0]], feats[1][0][idx[:, 1]], pair.dem,
                                       iterations=iters, with_extras=True)
    if not ok:
        return False, None
    tail = tail_ref.pose_tail(r, t, pair.affine, pair.tile.shape)
    return tail is not None, (-r.T @ t).ravel()


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


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from gisnav_b200 import weights as W

    cores = cpu_threads()
    params = W.unpack(W.load())
    per_step = max(1, args.ref_pairs)
    pairs = make_pairs(per_step * (args.steps + args.warmup))
    it = iter(pairs)
    for _ in range(args.warmup):
        for _ in range(per_step):
            cpu_reference_pair(next(it), params, args.ref_ransac_iters, args.keypoints)
    ok = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(per_step):
            ok += bool(cpu_reference_pair(next(it), params, args.ref_ransac_iters, args.keypoints)[0])
    dt = time.perf_counter() - t0
    value = ok / dt
    sample = f"{per_step} pair(s)/step x {args.steps} steps of the same synthetic 1280x720 vs 1024x1024 workload"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, per_step),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "cpu_model": cpu_model(),
                         "path": "torch-CPU SuperPoint-style stack x2 + dual-softmax head + cv2.solvePnPRansac + numpy tail"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "matched_fraction": ok / float(per_step * args.steps),
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch):
    return {
        "workload": "config 2: 1280x720 frame vs 1024x1024 orthophoto tile, full extract+match+PnP+WGS84 tail",
        "pairs_per_step_per_gpu": batch, "max_keypoints": args.keypoints, "ransac_iters": args.ransac_iters,
        "matcher_layers": getattr(args, "matcher_layers", 0),
        "l2_policy": "inputs cycle over distinct pre-staged batches; per-step activation traffic (>1 GB) exceeds the 126 MB L2",
    }


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


# ---- B200 arm ---------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import gisnav_b200
    from gisnav_b200 import sharding, weights as W

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
    cfg = gisnav_b200.Config(max_batch=args.batch, max_keypoints=args.keypoints, ransac_iters=args.ransac_iters,
                             max_image_h=1024, max_image_w=1280)
    ctx = gisnav_b200.Context(cfg, device=local, weights_device_ptr=wt.data_ptr(), weights_nbytes=W.BLOB_BYTES)
    pe = gisnav_b200.PoseEstimator(ctx)
    if args.matcher_layers > 0:
        # the reference matcher's transformer layers (LightGlue n_layers=9, pose_node.py:109-121).  No LightGlue
        # checkpoint is reachable offline, so the layers are untrained: residual-zero init (every block's last
        # linear layer is 0) runs the full arithmetic and leaves the trained head's matches unchanged.
        ctx.set_matcher_layers(W.pack_layers(W.layers_random_init(args.matcher_layers, seed=0, residual_zero=True),
                                             args.matcher_layers))

    n_sets = 3  # distinct batches cycled through (inputs differ every step)
    lo, _ = sharding.shard_range(world * n_sets * args.batch, rank, world)
    sets = []
    for s in range(n_sets):
        pairs = make_pairs(args.batch, first_seed=lo + s * args.batch)
        host = stack(pairs)
        dev_t = tuple(torch.from_numpy(a).to(dev) for a in host)
        pinned = tuple(torch.from_numpy(a).pin_memory() for a in host)
        sets.append((pairs, dev_t, pinned))
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
        return pe.estimate_batch_device(d[0], d[1], d[2], d[3], d[4])

    def step_host(i):
        _, _, p = sets[i % n_sets]
        return pe.estimate_batch(p[0].numpy(), p[1].numpy(), p[2].numpy(), p[3].numpy(), p[4].numpy())

    # CPU baseline on rank 0 at N=1, bounded sample, before the GPU timing
    cpu_baseline = None
    oracle_centres = {}
    if rank == 0 and world == 1 and args.cpu_pairs > 0:
        cores = cpu_threads()
        params = W.unpack(W.load())
        sample_pairs = sets[0][0][: args.cpu_pairs]
        cpu_reference_pair(sample_pairs[0], params, args.ref_ransac_iters, args.keypoints)  # warm-up
        t0 = time.perf_counter()
        ok = 0
        for j, p in enumerate(sample_pairs):
            good, c = cpu_reference_pair(p, params, args.ref_ransac_iters, args.keypoints)
            ok += bool(good)
            if good:
                oracle_centres[j] = c
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": ok / dt, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": cpu_model(),
                        "sample": f"{len(sample_pairs)} pairs of the same workload, {dt:.1f} s",
                        "path": "torch-CPU SuperPoint-style stack x2 + dual-softmax head + cv2.solvePnPRansac + numpy tail"}

    for i in range(args.warmup):
        step_device(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.profile(True)
    ctx.profile_read()
    launches0 = ctx.launch_count
    ms, results = timed(step_device, args.steps)
    launches = ctx.launch_count - launches0
    prof = ctx.profile_read()
    ctx.profile(False)
    clocks = sampler.stop() if rank == 0 else None

    matched = sum(1 for step in results for r in step if r.ok)
    total = args.steps * args.batch
    # accuracy vs ground truth (and vs the CPU oracle on the sampled pairs)
    err_gt = []
    for i, step in enumerate(results[:n_sets]):
        pairs = sets[i % n_sets][0]
        for r, p in zip(step, pairs):
            if r.ok:
                err_gt.append(np.linalg.norm(r.camera_center - (-p.r_gt.T @ p.t_gt).ravel()))
    err_or = [np.linalg.norm(results[0][j].camera_center - c) for j, c in oracle_centres.items() if results[0][j].ok]

    for i in range(min(args.warmup, 2)):
        step_host(i)
    ms_e2e, res_e2e = timed(step_host, args.steps)
    matched_e2e = sum(1 for step in res_e2e for r in step if r.ok)

    cnt = torch.tensor([matched, total, launches, matched_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    matched_all, total_all, launches_all, matched_e2e_all = (float(x) for x in cnt.tolist())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        # dominant kernel = the fused conv1a+conv1b+pool launch (K1's full-resolution block): 37 440 MAC per
        # input pixel (576 + 36 864); two launches per step (frames, rasters)
        px_step = args.batch * (FRAME_HW[0] * FRAME_HW[1] + TILE * TILE)
        dom = "conv_tc:1a+1b" if "conv_tc:1a+1b" in prof else max((k for k in prof if k.startswith("conv")), key=lambda k: prof[k][0])
        dom_ms, dom_launches = prof[dom]
        dom_flop_step = 2.0 * 37440.0 * px_step if dom == "conv_tc:1a+1b" else None
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        achieved = dom_flop_step * args.steps / (dom_ms * 1e-3) / 1e12 if dom_flop_step else None
        traffic = None
        try:  # dram bytes per launch from the committed `ncu --set full` capture (batch 16), scaled to this batch
            cap = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_step_batch16.json")))
            rows = [r for r in cap if r["kernel"].startswith("conv1_fused_kernel")]
            if rows and dom == "conv_tc:1a+1b":
                traffic = sum(r["dram_read_mb"] + r["dram_write_mb"] for r in rows) / len(rows) * 1e6 * args.batch / 16.0
        except (OSError, ValueError, KeyError):
            pass
        roofline = {
            "bound": "tensor", "kernel": f"{dom} (conv1_fused_kernel: im2col -> tcgen05 conv1a -> TMEM -> tcgen05 conv1b -> pool)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
            "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read+write, profiles/r01_ncu_full_step_batch16.json)",
            "algorithmic_flop_per_launch": dom_flop_step / 2.0 if dom_flop_step else None,
            "algorithmic_bytes_per_launch": px_step / 2.0 * (1 + 32) if dom_flop_step else None,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s (of fallback)",
            "launches_timed": dom_launches, "kernel_ms_total": dom_ms, "share_of_step": dom_ms / ms if ms > 0 else None,
        }
        # the whole dense stack (all conv*/head launches) against the same peak, for context
        conv_names = [k for k in prof if k.startswith("conv") or k in ("score_head_tc", "desc_head_tc")]
        conv_ms = sum(prof[k][0] for k in conv_names)
        stack_flop = FLOP_PER_PIXEL * px_step
        roofline_stack = {"kernel": "K1 dense stack (all conv*/head launches)", "achieved": stack_flop * args.steps / (conv_ms * 1e-3) / 1e12,
                          "peak": peak, "unit": "TFLOP/s", "frac": stack_flop * args.steps / (conv_ms * 1e-3) / 1e12 / peak,
                          "kernel_ms_total": conv_ms, "share_of_step": conv_ms / ms if ms > 0 else None}
        frame_b = FRAME_HW[0] * FRAME_HW[1]
        h2d = args.batch * (frame_b + 2 * TILE * TILE + 9 * 8 + 12 * 8)
        d2h = args.batch * 200
        line = {
            "metric": METRIC, "value": matched_all / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic (procedural texture, trained-from-scratch weights)",
            "config": workload_config(args, args.batch),
            "e2e": {"value": matched_e2e_all / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches_all),
            "clocks": clocks, "roofline": roofline, "roofline_stack": roofline_stack, "cpu_baseline": cpu_baseline,
            "matched_fraction": matched_all / total_all if total_all else None,
            "pairs_per_sec_processed": total_all / (ms * 1e-3),
            "pose_rmse_px_vs_ground_truth": float(np.sqrt(np.mean(np.square(err_gt)))) if err_gt else None,
            "pose_rmse_px_vs_cpu_oracle": float(np.sqrt(np.mean(np.square(err_or)))) if err_or else None,
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
            "impl_conv": cfg.conv_impl, "impl_match": cfg.match_impl, "matcher_layers": args.matcher_layers,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="pairs per step per GPU")
    ap.add_argument("--keypoints", type=int, default=1024)
    ap.add_argument("--ransac-iters", type=int, default=2048)
    ap.add_argument("--matcher-layers", type=int, default=0,
                    help="LightGlue transformer layers in front of the assignment head (0 = head only, the north_star "
                         "matcher; 9 = the reference's LightGlueMatcher depth, untrained residual-zero weights)")
    ap.add_argument("--cpu-pairs", type=int, default=4, help="pairs timed on the host cores for cpu_baseline (N=1 only)")
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
