"""GPU parity at BASELINE.json's full sizes (``-m gpu``): the kernels that are benchmarked are compared with
the CPU oracle where they are benchmarked — 1280x720 frames and 1024x1024 rasters, batches of three images
(so the persistent CTAs loop over dozens of tiles each: TMEM double-buffer parity, the patch ring wrap of
``conv1_fused_kernel`` and the multi-image tile indexing all run), image index 2 checked, seeded-random AND
trained weights, every layer of the stack, the score map, K2 bit-exact on a real full-size score map, the
on-demand descriptor head, and K4 bit-exact at N = M = 2048.

The measured per-layer errors are written to ``gpurun_out/k1_fullsize_errors.json`` (copied to ``profiles/``);
each tolerance below is about twice the largest value measured on a B200 (profiles/r02_k1_fullsize_errors.json).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from gisnav_b200 import Config, Context, KeypointMatcher, synth, weights as W
from gisnav_b200.context import ptr

pytestmark = pytest.mark.gpu

LAYERS = ("conv1a", "pool1", "conv2a", "pool2", "conv3a", "pool3", "conv4a", "conv4b", "convPa", "convDa")
# Bars per layer against the oracle's activation AS STORED (rounded to bf16), from the B200 measurements in
# profiles/r02_k1_fullsize_errors.json (4 cases x 2 images):
#   max  |gpu - oracle| / max|oracle| : 2 bf16 ulps of the top binade (measured: at most 1 ulp = 0.0078 everywhere)
#   mean |gpu - oracle| / max|oracle| : ~2x the largest measured value
#   fraction of bit-identical values  : a little below the smallest measured value; conv1a is bit-exact, the
#     fused conv1a+conv1b+pool kernel differs from the oracle in < 0.1 % of its values (one ulp each), and the
#     neighbouring-bf16 noise compounds with depth
TOL_MAX = 0.016
TOL_MEAN = {"conv1a": 1e-9, "pool1": 2e-6, "conv2a": 4e-6, "pool2": 2e-5, "conv3a": 6e-5, "pool3": 1.5e-4,
            "conv4a": 3e-4, "conv4b": 5e-4, "convPa": 6e-4, "convDa": 6e-4}
MIN_EQUAL = {"conv1a": 1.0, "pool1": 0.998, "conv2a": 0.99, "pool2": 0.97, "conv3a": 0.92, "pool3": 0.85,
             "conv4a": 0.75, "conv4b": 0.65, "convPa": 0.65, "convDa": 0.65}


def _bf16(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float32).numpy()


def _images(h, w, n=3):
    g = synth.ground_texture(2048, seed=41, n_shapes=1500)
    offs = [(11, 23), (300, 512), (777, 64)]
    return np.ascontiguousarray(np.stack([g[y:y + h, x:x + w] for y, x in offs[:n]]))


def _record(key, value):
    path = os.path.join(ROOT, "gpurun_out", "k1_fullsize_errors.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    try:
        data = json.load(open(path))
    except (OSError, ValueError):
        data = {}
    data[key] = value
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)


def _blob(kind):
    if kind == "trained":
        if not os.path.exists(W.DEFAULT_WEIGHTS_PATH):
            pytest.skip("trained weights not present")
        return W.load()
    return W.pack(W.random_init(0))


@pytest.mark.parametrize("weights", ["random", "trained"])
@pytest.mark.parametrize("hw", [(720, 1280), (1024, 1024)])
def test_k1_full_size_per_layer_parity(weights, hw):
    from oracle import nms_ref, sample_ref, superpoint_ref

    h, w = hw
    blob = _blob(weights)
    params = W.unpack(blob)
    imgs = _images(h, w)
    k = 1024
    ctx = Context(Config(max_batch=3, max_image_h=1024, max_image_w=1280, max_keypoints=k), weights=blob)
    ctx.check(ctx._lib.gnb_dense_batch(ctx.handle, ptr(imgs), 3, h, w, 1, 1))
    report = {}
    for idx in (2, 0):
        ref = superpoint_ref.forward_layers(imgs[idx], params)
        s_ref, d_ref = superpoint_ref.forward_dense(imgs[idx], params)
        for name in LAYERS:
            want = _bf16(ref[name])     # the library stores activations as bf16: compare with the oracle's value as stored
            got = np.empty(want.shape, np.float32)
            ctx.check(ctx._lib.gnb_layer_activation_at(ctx.handle, name.encode(), idx, ptr(got), got.size))
            scale = float(np.abs(want).max()) + 1e-6
            e_max, e_mean = float(np.abs(got - want).max()) / scale, float(np.mean(np.abs(got - want))) / scale
            frac_exact = float(np.mean(got == want))
            report[f"img{idx}.{name}"] = {"max_rel": e_max, "mean_rel": e_mean, "frac_bit_equal": frac_exact}
            assert e_max <= TOL_MAX, (name, idx, e_max)
            assert e_mean <= TOL_MEAN[name], (name, idx, e_mean)
            assert frac_exact >= MIN_EQUAL[name], (name, idx, frac_exact)
        score = np.empty((h, w), np.float32)
        ctx.check(ctx._lib.gnb_layer_activation_at(ctx.handle, b"score", idx, ptr(score), score.size))
        dense = np.empty((h // 8, w // 8, 256), np.float32)
        ctx.check(ctx._lib.gnb_layer_activation_at(ctx.handle, b"dense", idx, ptr(dense), dense.size))
        smax = float(s_ref.max())
        report[f"img{idx}.score"] = {"max_rel": float(np.abs(score - s_ref).max()) / smax,
                                     "mean_rel": float(np.mean(np.abs(score - s_ref))) / smax}
        report[f"img{idx}.dense"] = {"max_abs": float(np.abs(dense - d_ref).max()), "mean_abs": float(np.mean(np.abs(dense - d_ref)))}
        # measured: score max 0.0077 / mean 0.00037 of the map maximum; unit-norm descriptors max 0.0017 / mean 0.0002
        assert report[f"img{idx}.score"]["max_rel"] <= 0.016 and report[f"img{idx}.score"]["mean_rel"] <= 0.0008
        assert report[f"img{idx}.dense"]["max_abs"] <= 0.004 and report[f"img{idx}.dense"]["mean_abs"] <= 0.0004
        # K2 at full size on the GPU's own score map of batch image idx: bit-exact vs the oracle's NMS / top-K
        xy = np.empty((k, 2), np.float32)
        sc = np.empty((k,), np.float32)
        desc = np.empty((k, 256), np.float32)
        n = C.c_int(0)
        ctx.check(ctx._lib.gnb_slot_keypoints(ctx.handle, idx, ptr(xy), ptr(sc), ptr(desc), k, C.byref(n)))
        xy_ref, sc_ref = nms_ref.select_keypoints(score, max_keypoints=k)
        assert n.value == len(xy_ref)
        np.testing.assert_array_equal(xy[: n.value], xy_ref)
        np.testing.assert_array_equal(sc[: n.value], sc_ref)
        # K3 on demand (convDb only at the cells the keypoints touch) vs the oracle's dense map at the SAME keypoints
        d_want = sample_ref.sample_descriptors(d_ref, xy[: n.value], (h, w))
        report[f"img{idx}.desc_on_demand"] = {"max_abs": float(np.abs(desc[: n.value] - d_want).max()),
                                              "mean_abs": float(np.mean(np.abs(desc[: n.value] - d_want)))}
        assert report[f"img{idx}.desc_on_demand"]["max_abs"] <= 0.0025     # measured 0.0011 / 0.00011
        assert report[f"img{idx}.desc_on_demand"]["mean_abs"] <= 0.00025
        kp_ref, _ = nms_ref.select_keypoints(s_ref, max_keypoints=k)
        a = set(map(tuple, xy[: n.value].astype(int).tolist())); b = set(map(tuple, kp_ref.astype(int).tolist()))
        report[f"img{idx}.keypoints_common_with_oracle"] = len(a & b) / max(1, len(b))
    _record(f"{weights}.{h}x{w}", report)
    ctx.close()


@pytest.mark.parametrize("case", ["oracle_1024", "random_ties_1024x1280"])
def test_k2_full_size_bit_exact(case):
    from oracle import nms_ref, superpoint_ref

    blob = _blob("trained")
    if case == "oracle_1024":
        score, _ = superpoint_ref.forward_dense(_images(1024, 1024, 1)[0], W.unpack(blob))
    else:
        rng = np.random.default_rng(9)
        score = (np.round(rng.random((1024, 1280)) * 64) / 256).astype(np.float32)   # 65 distinct values: heavy ties
    h, w = score.shape
    for k in (1024, 2048):
        ctx = Context(Config(max_batch=1, max_image_h=1024, max_image_w=1280, max_keypoints=k), weights=blob)
        xy = np.empty((k, 2), np.float32)
        sc = np.empty((k,), np.float32)
        n = C.c_int(0)
        score = np.ascontiguousarray(score)
        ctx.check(ctx._lib.gnb_select_keypoints(ctx.handle, ptr(score), h, w, ptr(xy), ptr(sc), k, C.byref(n)))
        xy_ref, sc_ref = nms_ref.select_keypoints(score, max_keypoints=k)
        assert n.value == len(xy_ref) == k
        np.testing.assert_array_equal(xy[: n.value], xy_ref)
        np.testing.assert_array_equal(sc[: n.value], sc_ref)
        ctx.close()


@pytest.mark.parametrize("impl", [pytest.param(0, id="tcgen05"), pytest.param(1, id="simt")])
def test_k4_bit_exact_at_2048(impl):
    from oracle import matcher_ref

    rng = np.random.default_rng(2048)
    n = m = 2048
    a = rng.standard_normal((n, 256)).astype(np.float32)
    b = rng.standard_normal((m, 256)).astype(np.float32)
    perm = rng.permutation(m)[:1500]
    b[perm] = a[:1500] + 0.05 * rng.standard_normal((1500, 256)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    p = W.random_init(0)
    ctx = Context(Config(max_batch=1, max_image_h=64, max_image_w=64, max_keypoints=2048, match_threshold=0.01, match_impl=impl,
                         conv_impl=1), weights=W.pack(p))
    sc, idx = KeypointMatcher(ctx).match_arrays(a, b)
    sc_ref, idx_ref = matcher_ref.match(a, b, W.unpack(W.pack(p)), threshold=0.01)
    np.testing.assert_array_equal(idx, idx_ref)
    np.testing.assert_allclose(sc, sc_ref, rtol=2e-4)
    assert len(idx) >= 1000
    ctx.close()
