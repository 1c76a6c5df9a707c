#!/usr/bin/env python
"""Device timings of the SURVEY.md §8(f) rows built on top of the hot path (not part of bench.py's headline line):

* StereoNode rotate + centre-crop (`gnb_rotate_crop`): reference geometry of BASELINE config 2 — a 1469 x 1469 BGR
  orthoimage + DEM (GISNode requests ceil(hypot(1280, 720)), gis_node.py:361-384) rotated to a 45-degree bucket and
  cropped to 720 x 1280 — device tensors in and out, CUDA-event time of the kernel, algorithmic bytes / time against
  the measured HBM peak, next to cv2 (cvtColor + warpAffine + slice) on the host cores.
* LightGlue transformer layers: per-kernel CUDA-event times of one 16-pair batch step with 9 layers.

    python tools/bench_next_rows.py > profiles/r01_next_rows.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def main():
    import cv2
    import torch

    import gisnav_b200
    from gisnav_b200 import synth, weights as W
    from gisnav_b200.stereo import StereoAligner

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = peaks.get("hbm_gbs") or 6650.0
    dev = torch.device("cuda", 0)
    side, crop = 1469, (720, 1280)
    ground = synth.ground_texture(2048, seed=5, n_shapes=1500)
    ortho = np.ascontiguousarray(np.stack([ground[100:100 + side, 200:200 + side]] * 3, axis=-1))
    dem = synth.smooth_dem(2048, seed=3)[:side, :side].copy()
    ctx = gisnav_b200.Context(gisnav_b200.Config(max_batch=16, max_image_h=1024, max_image_w=1280))
    sa = StereoAligner(ctx)
    o_d, d_d = torch.from_numpy(ortho).to(dev), torch.from_numpy(dem).to(dev)
    for _ in range(3):
        sa.align_device(o_d, d_d, 45, crop)
    ctx.profile(True)
    ctx.profile_read()
    reps = 50
    for i in range(reps):
        sa.align_device(o_d, d_d, 45 * (i % 8), crop)
    prof = ctx.profile_read()
    ctx.profile(False)
    ms, n = prof["rotate_crop_kernel"]
    us = 1000.0 * ms / n
    alg_bytes = crop[0] * crop[1] * (3 + 1 + 2)   # BGR + DEM read over the crop footprint, 2 planes written
    cv2.setNumThreads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    for i in range(10):
        stack = np.dstack((cv2.cvtColor(ortho, cv2.COLOR_BGR2GRAY), dem))
        m = cv2.getRotationMatrix2D((side // 2, side // 2), 45 * (i % 8), 1.0)
        rot = cv2.warpAffine(stack, m, (side, side))
        _ = rot[(side // 2 - crop[0] // 2):(side // 2 - crop[0] // 2) + crop[0], (side // 2 - crop[1] // 2):(side // 2 - crop[1] // 2) + crop[1]]
    cpu_us = 1e6 * (time.perf_counter() - t0) / 10
    print(json.dumps({"row": "StereoNode rotate+crop (stereo_node.py:239,292-335)", "kernel": "rotate_crop_kernel<3>",
                      "workload": "1469x1469 BGR + DEM -> 720x1280, device resident", "us_per_launch": us, "launches": n,
                      "algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (us * 1e-6) / 1e9, "hbm_peak_gbs": hbm,
                      "frac_of_measured_hbm": alg_bytes / (us * 1e-6) / 1e9 / hbm,
                      "note": "5.5 MB per call: launch-latency bound, the reference runs it once per 45-degree yaw change",
                      "cpu_cv2_us": cpu_us, "cpu_cores": os.cpu_count()}))

    # transformer layers, one batch step
    pe = gisnav_b200.PoseEstimator(ctx)
    g4 = synth.ground_texture(4096, 0)
    pairs = [synth.make_pair(g4, i, (720, 1280), 1024) for i in range(16)]
    arrs = (np.stack([p.frame for p in pairs]), np.stack([p.tile for p in pairs]), np.stack([p.dem for p in pairs]),
            np.stack([p.k for p in pairs]), np.stack([p.affine for p in pairs]))
    dt = tuple(torch.from_numpy(a).to(dev) for a in arrs)
    n_layers = 9
    ctx.set_matcher_layers(W.pack_layers(W.layers_random_init(n_layers, seed=0, residual_zero=True), n_layers))
    for _ in range(3):
        res = pe.estimate_batch_device(*dt)
    ctx.profile(True)
    ctx.profile_read()
    steps = 10
    for _ in range(steps):
        res = pe.estimate_batch_device(*dt)
    prof = ctx.profile_read()
    ctx.profile(False)
    tokens = 2 * 16 * 1024
    flop = {"lg_linear<qkv>": 2.0 * tokens * 256 * 768, "lg_linear<out>": 2.0 * tokens * 256 * 256,
            "lg_linear<fc1>": 2.0 * tokens * 512 * 512, "lg_linear<fc2>": 2.0 * tokens * 512 * 256,
            "lg_attn<self>": 4.0 * 32 * 4 * 1024 * 1024 * 64, "lg_attn<cross>": 4.0 * 32 * 4 * 1024 * 1024 * 64}
    tmem_bytes = 2.0 * 32 * 4 * 1024 * 1024 * 4   # attention: S (fp32) is read from TMEM in both passes
    out = {"row": "LightGlue transformer layers (pose_node.py:109-121)", "layers": n_layers, "pairs_per_step": 16,
           "keypoints": 1024, "matched": sum(r.ok for r in res), "kernels": {}}
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    total = 0.0
    for name, (ms, n) in sorted(prof.items()):
        if not name.startswith("lg_"):
            continue
        us = 1000.0 * ms / n
        total += ms / steps
        rec = {"us_per_launch": us, "launches_per_step": n // steps, "ms_per_step": ms / steps}
        if name in flop:
            rec["algorithmic_tflops"] = flop[name] / (us * 1e-6) / 1e12
            rec["frac_of_measured_bf16_sustained"] = rec["algorithmic_tflops"] / peak_tf
        if name.startswith("lg_attn"):
            rec["tmem_read_bytes_per_clk_per_sm"] = tmem_bytes / (us * 1e-6) / 148 / 1.965e9
            rec["bound"] = "TMEM read (64 B/clk/SM LDTM rate): S is read twice"
        out["kernels"][name] = rec
    out["transformer_ms_per_step"] = total
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
