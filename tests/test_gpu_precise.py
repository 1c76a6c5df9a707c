"""GPU parity of the fp32-faithful mode (``Config(precision=1)``, ``-m gpu``): split-bf16 tcgen05 convs
(three MMAs per product into one fp32 TMEM accumulator), fp32 conv1a / 1x1 heads / matcher head.

Two oracles are used and told apart:
* ``superpoint_ref(..., quantize="x3")`` + ``matcher_ref(..., quantize=False)`` restate the mode's own contract
  (same split points): GPU-vs-oracle differences are fp32 summation order only -> tight per-layer bars;
* the PLAIN fp32 network (``quantize=False`` everywhere) is what the reference computes
  (ros/gisnav/gisnav/core/pose_node.py:254-287 keeps fp32 tensors end to end): the end-to-end test states how
  close the mode gets to it (keypoint sets, match sets, camera centre).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from gisnav_b200 import Config, Context, KeypointMatcher, PoseEstimator, synth, weights as W
from gisnav_b200.context import ptr

pytestmark = pytest.mark.gpu

LAYERS = ("conv1a", "pool1", "conv2a", "pool2", "conv3a", "pool3", "conv4a", "conv4b", "convPa", "convDa")


def _split_sum(a):
    """value as the library stores it: hi + lo with hi = bf16(v), lo = bf16(v - hi)."""
    import torch

    from oracle.superpoint_ref import split_hi_lo

    hi, lo = split_hi_lo(torch.from_numpy(np.ascontiguousarray(a)))
    return (hi + lo).numpy()


def _record(key, value):
    path = os.path.join(ROOT, "gpurun_out", "x3_errors.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    try:
        data = json.load(open(path))
    except (OSError, ValueError):
        data = {}
    data[key] = value
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)


def _check_layers(ctx, img, idx, params, report, tag, tol_max, tol_mean):
    from oracle import superpoint_ref

    h, w = img.shape
    ref = superpoint_ref.forward_layers(img, params, quantize="x3")
    for name in LAYERS:
        want = _split_sum(ref[name])
        got = np.empty(want.shape, np.float32)
        ctx.check(ctx._lib.gnb_layer_activation_at(ctx.handle, name.encode(), idx, ptr(got), got.size))
        scale = float(np.abs(want).max()) + 1e-12
        e_max, e_mean = float(np.abs(got - want).max()) / scale, float(np.mean(np.abs(got - want))) / scale
        report[f"{tag}.{name}"] = {"max_rel": e_max, "mean_rel": e_mean, "frac_bit_equal": float(np.mean(got == want))}
        assert np.isfinite(got).all(), name
        assert e_max <= tol_max, (name, e_max)
        assert e_mean <= tol_mean, (name, e_mean)
    s_ref, d_ref = superpoint_ref.forward_dense(img, params, quantize="x3")
    score = np.empty((h, w), np.float32)
    ctx.check(ctx._lib.gnb_layer_activation_at(ctx.handle, b"score", idx, ptr(score), score.size))
    dense = np.empty((h // 8, w // 8, 256), np.float32)
    ctx.check(ctx._lib.gnb_layer_activation_at(ctx.handle, b"dense", idx, ptr(dense), dense.size))
    report[f"{tag}.score"] = {"max_rel": float(np.abs(score - s_ref).max() / s_ref.max()), "mean_rel": float(np.mean(np.abs(score - s_ref)) / s_ref.max())}
    report[f"{tag}.dense"] = {"max_abs": float(np.abs(dense - d_ref).max()), "mean_abs": float(np.mean(np.abs(dense - d_ref)))}
    return score, s_ref, d_ref


@pytest.mark.parametrize("hw", [(96, 128), (40, 48), (120, 200), (64, 72)])
def test_x3_layers_match_oracle_small(rand_blob, rand_params, hw):
    h, w = hw
    img = np.ascontiguousarray(synth.ground_texture(512, seed=11, n_shapes=300)[40:40 + h, 60:60 + w])
    ctx = Context(Config(max_batch=2, max_image_h=256, max_image_w=320, precision=1), weights=rand_blob)
    imgs = np.ascontiguousarray(np.stack([img[::-1].copy(), img]))
    ctx.check(ctx._lib.gnb_dense_batch(ctx.handle, ptr(imgs), 2, h, w, 1, 0))
    report = {}
    _check_layers(ctx, img, 1, rand_params, report, "img1", tol_max=2e-4, tol_mean=2e-5)
    assert report["img1.score"]["max_rel"] <= 1e-3 and report["img1.dense"]["max_abs"] <= 1e-4
    _record(f"small.{h}x{w}", report)
    ctx.close()


@pytest.mark.parametrize("weights", ["random", "trained"])
@pytest.mark.parametrize("hw", [(720, 1280), (1024, 1024)])
def test_x3_full_size_parity(weights, hw):
    from oracle import nms_ref, sample_ref

    h, w = hw
    if weights == "trained" and not os.path.exists(W.DEFAULT_WEIGHTS_PATH):
        pytest.skip("trained weights not present")
    blob = W.load() if weights == "trained" else W.pack(W.random_init(0))
    params = W.unpack(blob)
    g = synth.ground_texture(2048, seed=41, n_shapes=1500)
    imgs = np.ascontiguousarray(np.stack([g[y:y + h, x:x + w] for y, x in ((11, 23), (300, 512), (777, 64))]))
    k = 1024
    ctx = Context(Config(max_batch=3, max_image_h=1024, max_image_w=1280, max_keypoints=k, precision=1), weights=blob)
    ctx.check(ctx._lib.gnb_dense_batch(ctx.handle, ptr(imgs), 3, h, w, 1, 1))
    report = {}
    score, s_ref, d_ref = _check_layers(ctx, imgs[2], 2, params, report, "img2", tol_max=2e-4, tol_mean=2e-5)
    assert report["img2.score"]["max_rel"] <= 1e-3 and report["img2.dense"]["max_abs"] <= 1e-4
    xy = np.empty((k, 2), np.float32); sc = np.empty((k,), np.float32); desc = np.empty((k, 256), np.float32)
    n = C.c_int(0)
    ctx.check(ctx._lib.gnb_slot_keypoints(ctx.handle, 2, ptr(xy), ptr(sc), ptr(desc), k, C.byref(n)))
    xy_ref, sc_ref = nms_ref.select_keypoints(score, max_keypoints=k)     # K2 on the GPU's own map: bit-exact
    np.testing.assert_array_equal(xy[: n.value], xy_ref)
    np.testing.assert_array_equal(sc[: n.value], sc_ref)
    d_want = sample_ref.sample_descriptors(d_ref, xy[: n.value], (h, w))
    report["img2.desc_on_demand"] = {"max_abs": float(np.abs(desc[: n.value] - d_want).max())}
    assert report["img2.desc_on_demand"]["max_abs"] <= 1e-4
    kp_o, _ = nms_ref.select_keypoints(s_ref, max_keypoints=k)
    a = set(map(tuple, xy[: n.value].astype(int).tolist())); b = set(map(tuple, kp_o.astype(int).tolist()))
    report["img2.keypoints_common_with_x3_oracle"] = len(a & b) / max(1, len(b))
    assert len(a & b) >= len(b) - 3
    _record(f"full.{weights}.{h}x{w}", report)
    ctx.close()


@pytest.mark.parametrize("impl", [pytest.param(0, id="tcgen05_x3"), pytest.param(1, id="simt_fp32")])
@pytest.mark.parametrize("shape", [(64, 64, 40), (300, 257, 150), (1, 50, 1), (1024, 1000, 700), (129, 512, 100)])
def test_x3_matcher_head(rand_blob, rand_params, shape, impl):
    """Matcher head of the fp32-faithful mode: fp32 projection, then S on tcgen05 from split-bf16 operands
    (match_impl = 0: oracle quantize="x3") or entirely in fp32 on the CUDA cores (match_impl = 1: the plain fp32 head)."""
    from oracle import matcher_ref

    n, m, shared = shape
    rng = np.random.default_rng(n * 1000 + m)
    a = rng.standard_normal((n, 256)).astype(np.float32)
    b = rng.standard_normal((m, 256)).astype(np.float32)
    perm = rng.permutation(m)[:shared]
    b[perm] = a[:shared] + 0.05 * rng.standard_normal((shared, 256)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    ctx = Context(Config(max_batch=2, max_image_h=64, max_image_w=64, max_keypoints=1024, match_threshold=0.01, precision=1, match_impl=impl),
                  weights=rand_blob)
    sc, idx = KeypointMatcher(ctx).match_arrays(a, b)
    sc_ref, idx_ref = matcher_ref.match(a, b, rand_params, threshold=0.01, quantize="x3" if impl == 0 else False)
    np.testing.assert_array_equal(idx, idx_ref)
    np.testing.assert_allclose(sc, sc_ref, rtol=1e-4)
    sc32, idx32 = matcher_ref.match(a, b, rand_params, threshold=0.01, quantize=False)
    assert len(idx32) == len(idx) and np.array_equal(idx32, idx)          # and the same matches as the plain fp32 head
    np.testing.assert_allclose(sc, sc32, rtol=2e-4)
    ctx.close()


def test_x3_end_to_end_vs_the_fp32_network():
    """What the mode is for: against the PLAIN fp32 pipeline (fp32 convs, fp32 matcher head) on config-2-shaped
    pairs the keypoint sets and the match sets are identical on most pairs, and then so is the pose (keypoints are
    integer pixels: the pose depends on the data only through those sets).  A single flipped decision re-draws RANSAC
    and moves the camera centre by ~0.1 px; bars from tools/precision_study.py (16 pairs, CPU emulation)."""
    from oracle import matcher_ref, nms_ref, pnp_ref, sample_ref, superpoint_ref

    if not os.path.exists(W.DEFAULT_WEIGHTS_PATH):
        pytest.skip("trained weights not present")
    blob = W.load()
    params = W.unpack(blob)
    ground = synth.ground_texture(4096, 0)
    n_pairs = 4
    pairs = [synth.make_pair(ground, s, (720, 1280), 1024) for s in range(n_pairs)]
    ctx = Context(Config(max_batch=n_pairs, max_image_h=1024, max_image_w=1280, max_keypoints=1024, precision=1), weights=blob)
    pe = PoseEstimator(ctx)
    res = pe.estimate_batch(np.stack([p.frame for p in pairs]), np.stack([p.tile for p in pairs]), np.stack([p.dem for p in pairs]),
                            np.stack([p.k for p in pairs]), np.stack([p.affine for p in pairs]))
    rows = []
    for i, (p, r) in enumerate(zip(pairs, res)):
        feats = []
        for img in (p.frame, p.tile):
            s, d = superpoint_ref.forward_dense(img, params, quantize=False)
            xy, _ = nms_ref.select_keypoints(s, max_keypoints=1024)
            feats.append((xy, sample_ref.sample_descriptors(d, xy, img.shape)))
        _, idx = matcher_ref.match(feats[0][1], feats[1][1], params, 0.5, quantize=False)
        obj = pnp_ref.points3d(feats[1][0][idx[:, 1]], p.dem)
        ref = pnp_ref.solve_pnp_ransac(obj, feats[0][0][idx[:, 0]], p.k, iters=ctx.config.ransac_iters)
        c_ref = (-ref["r"].T @ ref["t"]).ravel()
        k = 1024
        kp_same = []
        for slot, f in ((i, feats[0]), (n_pairs + i, feats[1])):
            xy = np.empty((k, 2), np.float32); nn = C.c_int(0)
            ctx.check(ctx._lib.gnb_slot_keypoints(ctx.handle, slot, ptr(xy), None, None, k, C.byref(nn)))
            kp_same.append(len(set(map(tuple, xy[: nn.value].astype(int).tolist())) ^ set(map(tuple, f[0].astype(int).tolist()))) // 2)
        rows.append({"pair": i, "kp_differ": kp_same, "matches_gpu": r.n_matches, "matches_fp32": int(len(idx)),
                     "centre_diff_px": float(np.linalg.norm(r.camera_center - c_ref)), "ok": bool(r.ok)})
    _record("end_to_end_vs_fp32", rows)
    d = np.array([r["centre_diff_px"] for r in rows])
    assert all(r["ok"] for r in rows)
    assert np.median(d) <= 1e-2, rows
    assert d.max() <= 0.5, rows
    assert sum(sum(r["kp_differ"]) for r in rows) <= 2 * n_pairs, rows
    ctx.close()


def test_x3_composes_with_layers_and_candidate_cache():
    """fp32-faithful mode through the other entry points: residual-zero transformer layers leave the poses unchanged,
    and the candidate search (tile-feature cache holding fp32 + split projected descriptors) reproduces the batch path
    on a cold and on a fully cached call."""
    if not os.path.exists(W.DEFAULT_WEIGHTS_PATH):
        pytest.skip("trained weights not present")
    blob = W.load()
    ground = synth.ground_texture(2048, seed=24, n_shapes=3000)
    pairs = [synth.make_pair(ground, s, frame_hw=(240, 320), tile_size=256) for s in (1, 2)]
    ctx = Context(Config(max_batch=4, max_image_h=256, max_image_w=320, max_keypoints=512, precision=1), weights=blob)
    pe = PoseEstimator(ctx)
    args = (np.stack([p.frame for p in pairs]), np.stack([p.tile for p in pairs]), None, np.stack([p.k for p in pairs]),
            np.stack([p.affine for p in pairs]))
    base = pe.estimate_batch(*args)
    assert all(r.ok for r in base)
    decoys = [synth.make_pair(ground, s, frame_hw=(240, 320), tile_size=256).tile for s in (11, 12)]
    tiles = np.stack([decoys[0], pairs[0].tile, decoys[1]])
    ids = np.array([100, 101, 102])
    affs = np.stack([pairs[0].affine] * 3)
    for expect_hits in (0, 3):
        best, res, hits = pe.estimate_candidates(pairs[0].frame, tiles, ids, None, pairs[0].k, affs)
        assert hits == expect_hits and best == 1
        assert (res[1].n_matches, res[1].n_inliers) == (base[0].n_matches, base[0].n_inliers)
        np.testing.assert_array_equal(res[1].r, base[0].r)
        np.testing.assert_array_equal(res[1].ecef, base[0].ecef)
    ctx.set_matcher_layers(W.pack_layers(W.layers_random_init(2, seed=1, residual_zero=True), 2))
    with_layers = pe.estimate_batch(*args)
    for a, b in zip(base, with_layers):
        assert b.ok and a.n_matches == b.n_matches
        np.testing.assert_array_equal(a.r, b.r)
    ctx.close()
