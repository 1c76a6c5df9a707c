#!/usr/bin/env python
"""CPU study (oracle only): how far are the bf16 and the split-bf16 ("x3") conv contracts from the plain
fp32 network (fp32 matcher head included), end to end?  "bf16" = bf16 conv operands + bf16 matcher GEMM (the fast
mode), "x3" = split-bf16 convs + fp32 heads + fp32 matcher (precision = 1).  For N config-2 pairs: keypoint-set overlap, match-set overlap and camera-centre
difference against the fp32 oracle.  Writes one JSON line; used to set the tolerances of the fp32-faithful
mode.

    python tools/precision_study.py --pairs 4 [--hw 720 1280 --tile 1024]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_pair(pair, params, mode, k_cap, iters, match_quant=True):
    from oracle import matcher_ref, nms_ref, pnp_ref, sample_ref, superpoint_ref

    feats = []
    for img in (pair.frame, pair.tile):
        s, d = superpoint_ref.forward_dense(img, params, quantize=mode)
        xy, _ = nms_ref.select_keypoints(s, max_keypoints=k_cap)
        feats.append((xy, sample_ref.sample_descriptors(d, xy, img.shape)))
    _, idx = matcher_ref.match(feats[0][1], feats[1][1], params, 0.5, quantize=match_quant)
    out = {"kp": [set(map(tuple, f[0].astype(int).tolist())) for f in feats]}
    out["matches"] = {(tuple(feats[0][0][i].astype(int)), tuple(feats[1][0][j].astype(int))) for i, j in idx.tolist()}
    out["centre"] = None
    if len(idx) >= 15:
        obj = pnp_ref.points3d(feats[1][0][idx[:, 1]], pair.dem)
        ref = pnp_ref.solve_pnp_ransac(obj, feats[0][0][idx[:, 0]], pair.k, iters=iters)
        if ref["status"] == 0:
            out["centre"] = (-ref["r"].T @ ref["t"]).ravel()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4)
    ap.add_argument("--hw", type=int, nargs=2, default=(720, 1280))
    ap.add_argument("--tile", type=int, default=1024)
    ap.add_argument("--keypoints", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=2048)
    ap.add_argument("--first-seed", type=int, default=0)
    args = ap.parse_args()
    import torch

    from gisnav_b200 import synth, weights as W

    torch.set_num_threads(os.cpu_count() or 1)
    params = W.unpack(W.load())
    ground = synth.ground_texture(4096, 0)
    rows = []
    t0 = time.time()
    for s in range(args.first_seed, args.first_seed + args.pairs):
        pair = synth.make_pair(ground, s, tuple(args.hw), args.tile)
        ref = run_pair(pair, params, False, args.keypoints, args.iters, match_quant=False)   # fp32 everywhere
        row = {"seed": s}
        for name, mode, mq in (("bf16", True, True), ("x3", "x3", False), ("x3_tc_matcher", "x3", "x3")):
            got = run_pair(pair, params, mode, args.keypoints, args.iters, match_quant=mq)
            kp_diff = sum(len(a ^ b) // 2 for a, b in zip(ref["kp"], got["kp"]))
            m_common = len(ref["matches"] & got["matches"])
            d = None
            if ref["centre"] is not None and got["centre"] is not None:
                d = float(np.linalg.norm(ref["centre"] - got["centre"]))
            row[name] = {"kp_differ": kp_diff, "matches_ref": len(ref["matches"]), "matches_got": len(got["matches"]),
                         "matches_common": m_common, "centre_diff_px": d}
        rows.append(row)
        print(json.dumps(row), flush=True)
    summary = {"pairs": args.pairs, "seconds": time.time() - t0}
    for name in ("bf16", "x3", "x3_tc_matcher"):
        ds = [r[name]["centre_diff_px"] for r in rows if r[name]["centre_diff_px"] is not None]
        summary[name] = {"median_px": float(np.median(ds)), "max_px": float(np.max(ds)), "rmse_px": float(np.sqrt(np.mean(np.square(ds)))),
                         "frac_le_1e-3": float(np.mean(np.array(ds) <= 1e-3)), "frac_le_1e-2": float(np.mean(np.array(ds) <= 1e-2)),
                         "kp_differ_mean": float(np.mean([r[name]["kp_differ"] for r in rows]))}
    print(json.dumps({"summary": summary}), flush=True)


if __name__ == "__main__":
    main()
