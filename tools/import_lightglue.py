#!/usr/bin/env python
"""Import a LightGlue checkpoint (the matcher the reference runs: ``LightGlueMatcher("sift", {"n_layers": 9, ...})``,
ros/gisnav/gisnav/core/pose_node.py:109-121) into this library's blobs.

    python tools/import_lightglue.py superpoint_lightglue.pth layers.gnbl [--head-into weights.bin --head-out weights_lg.bin]

* transformer layers -> a GNBL layer blob (``gisnav_b200.weights.pack_layers``; load with ``ctx.set_matcher_layers``);
* assignment head of the LAST layer (``log_assignment.{L-1}.final_proj / matchability``) -> the ``match.*`` tensors of a
  GNBW weight blob (``--head-into`` an existing blob, written to ``--head-out``).

State-dict layout per the published LightGlue code (cvg/LightGlue ``lightglue.py``; kornia 0.7.2 vendors the same module
and loads the same files — not verifiable offline):

    posenc.Wr.weight [32, M]                       M = 2 for 256-d SuperPoint-style features (this library's extractor);
                                                   the "sift" variant has M = 4 (scale, orientation) and input_proj 128 -> 256:
                                                   refused, its features are not what this extractor produces
    transformers.i.self_attn.Wqkv.{weight [768,256], bias [768]}   output index = head * 192 + dim * 3 + {0: q, 1: k, 2: v}
    transformers.i.self_attn.out_proj, .ffn.0 (Linear 512 -> 512), .ffn.1 (LayerNorm 512), .ffn.3 (Linear 512 -> 256)
    transformers.i.cross_attn.to_qk (ONE projection used for the queries of one side and the keys of the other),
                              .to_v, .to_out, .ffn.0 / .ffn.1 / .ffn.3
    log_assignment.i.final_proj.{weight [256,256], bias}, log_assignment.i.matchability.{weight [1,256], bias [1]}
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import Dict, Mapping, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _np(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t), np.float32)


def count_layers(sd: Mapping[str, object]) -> int:
    idx = {int(k.split(".")[1]) for k in sd if k.startswith("transformers.")}
    if not idx or sorted(idx) != list(range(len(idx))):
        raise ValueError("state dict has no contiguous transformers.<i>.* entries")
    return len(idx)


def split_wqkv(w: np.ndarray, b: np.ndarray, heads: int = 4) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
    """De-interleave the fused self-attention projection: row h * 3 hd + d * 3 + c  ->  (c, h * hd + d)."""
    dim = w.shape[1]
    hd = dim // heads
    w4 = w.reshape(heads, hd, 3, dim)
    b4 = b.reshape(heads, hd, 3)
    ws = {n: np.ascontiguousarray(w4[:, :, c, :].reshape(dim, dim)) for c, n in enumerate("qkv")}
    bs = {n: np.ascontiguousarray(b4[:, :, c].reshape(dim)) for c, n in enumerate("qkv")}
    return ws, bs


def convert_layers(sd: Mapping[str, object]) -> Tuple[Dict[str, np.ndarray], int]:
    """LightGlue state dict -> (``lg.*`` parameter dict in this library's naming, number of layers)."""
    from gisnav_b200 import weights as W

    if any(k.startswith("input_proj.") for k in sd):
        raise ValueError("checkpoint has an input projection (e.g. the 128-d 'sift' variant): this path carries 256-d descriptors")
    n_layers = count_layers(sd)
    pos = _np(sd["posenc.Wr.weight"])
    if pos.shape != (W.DESC_DIM // W.LG_HEADS // 2, 2):
        raise ValueError(f"posenc.Wr.weight has shape {pos.shape}: only the keypoint-only encoder (M = 2) is supported")
    out: Dict[str, np.ndarray] = {"lg.pos.weight": pos}
    for i in range(n_layers):
        s, c = f"transformers.{i}.self_attn", f"transformers.{i}.cross_attn"
        ws, bs = split_wqkv(_np(sd[s + ".Wqkv.weight"]), _np(sd[s + ".Wqkv.bias"]), W.LG_HEADS)
        for n in "qkv":
            out[f"lg.{i}.self.{n}.weight"], out[f"lg.{i}.self.{n}.bias"] = ws[n], bs[n]
        out[f"lg.{i}.self.o.weight"], out[f"lg.{i}.self.o.bias"] = _np(sd[s + ".out_proj.weight"]), _np(sd[s + ".out_proj.bias"])
        # cross attention: queries and keys share to_qk
        for n in "qk":
            out[f"lg.{i}.cross.{n}.weight"], out[f"lg.{i}.cross.{n}.bias"] = _np(sd[c + ".to_qk.weight"]), _np(sd[c + ".to_qk.bias"])
        out[f"lg.{i}.cross.v.weight"], out[f"lg.{i}.cross.v.bias"] = _np(sd[c + ".to_v.weight"]), _np(sd[c + ".to_v.bias"])
        out[f"lg.{i}.cross.o.weight"], out[f"lg.{i}.cross.o.bias"] = _np(sd[c + ".to_out.weight"]), _np(sd[c + ".to_out.bias"])
        for blk, src in (("self", s), ("cross", c)):
            out[f"lg.{i}.{blk}.fc1.weight"], out[f"lg.{i}.{blk}.fc1.bias"] = _np(sd[src + ".ffn.0.weight"]), _np(sd[src + ".ffn.0.bias"])
            out[f"lg.{i}.{blk}.ln.weight"], out[f"lg.{i}.{blk}.ln.bias"] = _np(sd[src + ".ffn.1.weight"]), _np(sd[src + ".ffn.1.bias"])
            out[f"lg.{i}.{blk}.fc2.weight"], out[f"lg.{i}.{blk}.fc2.bias"] = _np(sd[src + ".ffn.3.weight"]), _np(sd[src + ".ffn.3.bias"])
    for name, shape in W.layer_tensors(n_layers).items():
        if tuple(out[name].shape) != tuple(shape):
            raise ValueError(f"{name}: expected {shape}, checkpoint gives {out[name].shape}")
    return out, n_layers


def convert_head(sd: Mapping[str, object], n_layers: int) -> Dict[str, np.ndarray]:
    """Assignment head of the last layer -> the ``match.*`` tensors of the GNBW blob."""
    p = f"log_assignment.{n_layers - 1}"
    return {"match.proj.weight": _np(sd[p + ".final_proj.weight"]), "match.proj.bias": _np(sd[p + ".final_proj.bias"]),
            "match.m.weight": _np(sd[p + ".matchability.weight"]).reshape(-1), "match.m.bias": _np(sd[p + ".matchability.bias"]).reshape(1)}


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("checkpoint")
    ap.add_argument("layers_out")
    ap.add_argument("--head-into", help="existing GNBW weight blob whose match.* tensors are replaced by the checkpoint's head")
    ap.add_argument("--head-out")
    args = ap.parse_args()
    import torch

    from gisnav_b200 import weights as W

    sd = torch.load(args.checkpoint, map_location="cpu")
    sd = sd.get("state_dict", sd) if isinstance(sd, dict) else sd
    sd = {k[len("matcher."):] if k.startswith("matcher.") else k: v for k, v in sd.items()}
    lp, n_layers = convert_layers(sd)
    with open(args.layers_out, "wb") as f:
        f.write(W.pack_layers(lp, n_layers))
    print(f"{n_layers} layers -> {args.layers_out}")
    if args.head_into:
        params = W.unpack(W.load(args.head_into))
        params.update(convert_head(sd, n_layers))
        with open(args.head_out or args.head_into, "wb") as f:
            f.write(W.pack(params))
        print(f"assignment head of layer {n_layers - 1} -> {args.head_out or args.head_into}")


if __name__ == "__main__":
    main()
