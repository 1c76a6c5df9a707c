#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

* ``pnp_cv2_*.npz`` — inputs and outputs of the REFERENCE's own PnP call
  (``cv2.solvePnPRansac`` + ``cv2.Rodrigues`` with the arguments of
  ros/gisnav/gisnav/core/_shared.py:95-119), executed in this container (cv2 4.13.0) through
  oracle/cv2_ref.py.  These pin oracle/pnp_ref.c and the CUDA solver.
* ``stages_small.npz`` — per-stage outputs of the CPU oracle on a small seeded image pair with
  seeded random weights (``gisnav_b200.weights.random_init(0)``): score map, keypoints,
  descriptors, matches.  The reference has no fixtures for these stages, so these
  pin the oracle against drift; the oracle itself is cross-checked against the independent
  ``transformers`` restatements in tests/test_oracle_cpu.py.

* ``stereo_cv2.npz`` — inputs and outputs of the reference's rotate + centre-crop call sequence
  (stereo_node.py:239,306-335) executed here; pins oracle/stereo_ref.py and csrc/warp.cu.

    python tools/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from gisnav_b200 import synth, weights as W  # noqa: E402
from oracle import cv2_ref, matcher_ref, nms_ref, sample_ref, superpoint_ref, tail_ref  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")


def pnp_fixtures():
    for seed in range(5):
        # seeds 0-3: outliers at least 30 px from the consensus => the inlier set does not depend on which
        # minimal models a RANSAC happens to draw; seed 4 keeps borderline outliers (ambiguous set).
        c = synth.synth_correspondences(seed, n_points=300, outlier_frac=0.2 + 0.05 * seed, noise_px=0.5,
                                        tile_size=256, frame_hw=(240, 320), relief=True,
                                        outlier_min_px=30.0 if seed < 4 else 0.0)
        out = {}
        for iters in (10, 2000):  # 10 is what the reference passes (_shared.py:115)
            r, t, ok, inl = cv2_ref.compute_pose(c["k"], c["mkp_qry"], c["mkp_ref"], c["dem"], iterations=iters,
                                                 with_extras=True)
            mask = np.zeros(len(c["mkp_ref"]), np.uint8)
            if inl is not None:
                mask[inl.ravel()] = 1
            out[f"r_{iters}"], out[f"t_{iters}"], out[f"ok_{iters}"], out[f"mask_{iters}"] = r, t, np.array(ok), mask
        tail = tail_ref.pose_tail(out["r_2000"], out["t_2000"], c["affine"], c["dem"].shape)
        np.savez_compressed(os.path.join(OUT, f"pnp_cv2_{seed}.npz"), mkp_ref=c["mkp_ref"], mkp_qry=c["mkp_qry"],
                            dem=c["dem"], k=c["k"], affine=c["affine"], r_gt=c["r_gt"], t_gt=c["t_gt"],
                            tail_ecef=tail[0], tail_quat=tail[1], tail_lla=tail[2], cv2_version=np.array(cv2.__version__), **out)


def stage_fixtures():
    params = W.unpack(W.pack(W.random_init(0)))
    g = synth.ground_texture(512, seed=3, n_shapes=400)
    a = np.ascontiguousarray(g[100:196, 200:328])  # 96 x 128
    b = np.ascontiguousarray(g[104:200, 206:334])  # shifted view
    res = {}
    for name, img in (("a", a), ("b", b)):
        score, dense = superpoint_ref.forward_dense(img, params)
        xy, sc = nms_ref.select_keypoints(score, max_keypoints=64, threshold=0.005)
        desc = sample_ref.sample_descriptors(dense, xy, img.shape)
        res.update({f"img_{name}": img, f"score_{name}": score, f"dense_{name}": dense.astype(np.float32),
                    f"xy_{name}": xy, f"kpscore_{name}": sc, f"desc_{name}": desc})
    for thr, tag in ((0.0, "t0"), (0.002, "t002")):  # random weights: use low thresholds to get a non-empty set
        ms, idx = matcher_ref.match(res["desc_a"], res["desc_b"], params, threshold=thr)
        res[f"match_scores_{tag}"], res[f"match_idx_{tag}"] = ms, idx
    res["assign"] = matcher_ref.assignment_scores(res["desc_a"], res["desc_b"], params)
    np.savez_compressed(os.path.join(OUT, "stages_small.npz"), **res)


def stereo_fixtures():
    """Outputs of the REFERENCE's own rotate+crop call sequence (cv2.cvtColor + cv2.getRotationMatrix2D +
    cv2.warpAffine + slice, stereo_node.py:239,306-335) on a small seeded orthoimage/DEM, executed here."""
    from oracle import stereo_ref

    rng = np.random.default_rng(7)
    g = synth.ground_texture(256, seed=9, n_shapes=120)
    ortho = np.stack([np.roll(g, 3 * c, axis=c % 2)[:149, :149] for c in range(3)], axis=-1).astype(np.uint8)  # BGR
    dem = rng.integers(0, 40, (149, 149), dtype=np.uint8)
    stack = stereo_ref.cv2_orthoimage_stack(ortho, dem)
    res = {"ortho_bgr": ortho, "dem": dem, "gray": stack[:, :, 0], "cv2_version": np.array(cv2.__version__)}
    for ang in (0, 45, 90, 135, 180, 225, 270, 315):
        cropped, inv = stereo_ref.cv2_rotate_and_crop_center(stack, ang, (72, 104))
        res[f"crop_{ang}"], res[f"inv_{ang}"] = np.ascontiguousarray(cropped), inv
    np.savez_compressed(os.path.join(OUT, "stereo_cv2.npz"), **res)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--only-stereo" not in sys.argv:
        pnp_fixtures()
        stage_fixtures()
    stereo_fixtures()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
