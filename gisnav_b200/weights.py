"""Weight blob for the B200 pose-estimation path (detector/descriptor net + matcher head).

The reference extractor is ``cv2.SIFT_create()`` (ros/gisnav/gisnav/core/pose_node.py:107,122) and
its matcher is kornia's LightGlue (pose_node.py:94-121); neither has weights in the reference
tree.  BASELINE.json's north_star replaces them with a SuperPoint-style conv stack plus the
LightGlue *assignment head* (SURVEY.md §0.1, §8(c)); this module fixes the on-disk layout of those
parameters so the CUDA library, the oracle and the training tool agree.

Blob layout (little endian): 16-byte header ``b"GNBW"``, u32 version, u32 n_floats, u32 reserved,
then ``n_floats`` float32 values: for every entry of :data:`TENSORS`, in order, the tensor in
C order.  Conv weights are ``[Cout, Cin, kh, kw]`` (PyTorch order), biases ``[Cout]``.
"""
from __future__ import annotations

import os
import struct
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

MAGIC = b"GNBW"
VERSION = 1
DESC_DIM = 256

# name, Cin, Cout, kernel
CONV_LAYERS = (
    ("conv1a", 1, 64, 3),
    ("conv1b", 64, 64, 3),
    ("conv2a", 64, 64, 3),
    ("conv2b", 64, 64, 3),
    ("conv3a", 64, 128, 3),
    ("conv3b", 128, 128, 3),
    ("conv4a", 128, 128, 3),
    ("conv4b", 128, 128, 3),
    ("convPa", 128, 256, 3),
    ("convPb", 256, 65, 1),
    ("convDa", 128, 256, 3),
    ("convDb", 256, 256, 1),
)


def _tensor_table() -> "OrderedDict[str, Tuple[int, ...]]":
    t: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for name, cin, cout, k in CONV_LAYERS:
        t[name + ".weight"] = (cout, cin, k, k)
        t[name + ".bias"] = (cout,)
    # LightGlue assignment head: final projection + matchability
    t["match.proj.weight"] = (DESC_DIM, DESC_DIM)
    t["match.proj.bias"] = (DESC_DIM,)
    t["match.m.weight"] = (DESC_DIM,)
    t["match.m.bias"] = (1,)
    return t


TENSORS = _tensor_table()
N_FLOATS = int(sum(int(np.prod(s)) for s in TENSORS.values()))
HEADER_BYTES = 16
BLOB_BYTES = HEADER_BYTES + 4 * N_FLOATS

DEFAULT_WEIGHTS_PATH = os.path.join(os.path.dirname(__file__), "weights", "gnb_superpoint_v1.bin")


def pack(params: Dict[str, np.ndarray]) -> bytes:
    """Serialise a ``name -> array`` dict into the blob layout."""
    chunks = [MAGIC + struct.pack("<III", VERSION, N_FLOATS, 0)]
    for name, shape in TENSORS.items():
        a = np.ascontiguousarray(np.asarray(params[name], dtype=np.float32))
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
        chunks.append(a.astype("<f4").tobytes())
    blob = b"".join(chunks)
    assert len(blob) == BLOB_BYTES
    return blob


def unpack(blob: bytes) -> Dict[str, np.ndarray]:
    """Parse a blob back into a ``name -> float32 array`` dict."""
    if len(blob) != BLOB_BYTES or blob[:4] != MAGIC:
        raise ValueError("not a GNBW weight blob of the expected size")
    version, n, _ = struct.unpack("<III", blob[4:16])
    if version != VERSION or n != N_FLOATS:
        raise ValueError(f"weight blob version/size mismatch: v{version}, {n} floats")
    flat = np.frombuffer(blob, dtype="<f4", offset=HEADER_BYTES)
    out: Dict[str, np.ndarray] = {}
    off = 0
    for name, shape in TENSORS.items():
        cnt = int(np.prod(shape))
        out[name] = flat[off : off + cnt].reshape(shape).copy()
        off += cnt
    return out


def random_init(seed: int = 0, match_temperature: float = 20.0) -> Dict[str, np.ndarray]:
    """Seeded He-normal init (no checkpoint can be downloaded here).

    The matcher head starts as a scaled identity so that ``S = T * <a, b>`` for L2-normalised
    descriptors, with matchability logits large enough that ``logsigmoid`` is ~0.
    """
    rng = np.random.default_rng(seed)
    p: Dict[str, np.ndarray] = {}
    for name, cin, cout, k in CONV_LAYERS:
        fan_in = cin * k * k
        p[name + ".weight"] = (rng.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        p[name + ".bias"] = (rng.standard_normal((cout,)) * 0.01).astype(np.float32)
    # S = (Wa/d^.25).(Wb/d^.25) = T <a,b>  =>  W = sqrt(T) * d^.25 * I
    scale = np.sqrt(match_temperature) * DESC_DIM ** 0.25
    p["match.proj.weight"] = (np.eye(DESC_DIM) * scale).astype(np.float32)
    p["match.proj.bias"] = np.zeros((DESC_DIM,), np.float32)
    p["match.m.weight"] = np.zeros((DESC_DIM,), np.float32)
    p["match.m.bias"] = np.full((1,), 12.0, np.float32)
    return p


def load(path: str | None = None) -> bytes:
    """Read a blob from disk (default: the trained weights shipped in the package)."""
    with open(path or DEFAULT_WEIGHTS_PATH, "rb") as f:
        blob = f.read()
    unpack(blob)  # validates
    return blob


# ---- LightGlue transformer layers --------------------------------------------
# The reference matcher runs nine self+cross attention layers in front of the assignment head
# (LightGlueMatcher(..., n_layers=9), ros/gisnav/gisnav/core/pose_node.py:109-121).  Their parameters
# travel in a second, optional blob: 16-byte header b"GNBL", u32 version, u32 n_layers, u32 n_floats,
# then float32 tensors in the order of `layer_tensors(n_layers)` (Linear weights are [out, in]).
LAYER_MAGIC = b"GNBL"
LAYER_VERSION = 1
LG_HEADS = 4
LG_HIDDEN = 2 * DESC_DIM


def layer_tensors(n_layers: int) -> "OrderedDict[str, Tuple[int, ...]]":
    t: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    t["lg.pos.weight"] = (DESC_DIM // LG_HEADS // 2, 2)   # rotary angle projector, no bias
    for i in range(n_layers):
        for blk in ("self", "cross"):
            p = f"lg.{i}.{blk}"
            for proj in ("q", "k", "v", "o"):
                t[f"{p}.{proj}.weight"] = (DESC_DIM, DESC_DIM)
                t[f"{p}.{proj}.bias"] = (DESC_DIM,)
            t[f"{p}.fc1.weight"] = (LG_HIDDEN, LG_HIDDEN)
            t[f"{p}.fc1.bias"] = (LG_HIDDEN,)
            t[f"{p}.ln.weight"] = (LG_HIDDEN,)
            t[f"{p}.ln.bias"] = (LG_HIDDEN,)
            t[f"{p}.fc2.weight"] = (DESC_DIM, LG_HIDDEN)
            t[f"{p}.fc2.bias"] = (DESC_DIM,)
    return t


def layer_floats(n_layers: int) -> int:
    return int(sum(int(np.prod(s)) for s in layer_tensors(n_layers).values()))


def pack_layers(params: Dict[str, np.ndarray], n_layers: int) -> bytes:
    table = layer_tensors(n_layers)
    chunks = [LAYER_MAGIC + struct.pack("<III", LAYER_VERSION, n_layers, layer_floats(n_layers))]
    for name, shape in table.items():
        a = np.ascontiguousarray(np.asarray(params[name], dtype=np.float32))
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
        chunks.append(a.astype("<f4").tobytes())
    return b"".join(chunks)


def unpack_layers(blob: bytes) -> Tuple[Dict[str, np.ndarray], int]:
    if len(blob) < HEADER_BYTES or blob[:4] != LAYER_MAGIC:
        raise ValueError("not a GNBL layer blob")
    version, n_layers, n = struct.unpack("<III", blob[4:16])
    if version != LAYER_VERSION or n != layer_floats(n_layers) or len(blob) != HEADER_BYTES + 4 * n:
        raise ValueError("layer blob version/size mismatch")
    flat = np.frombuffer(blob, dtype="<f4", offset=HEADER_BYTES)
    out: Dict[str, np.ndarray] = {}
    off = 0
    for name, shape in layer_tensors(n_layers).items():
        cnt = int(np.prod(shape))
        out[name] = flat[off: off + cnt].reshape(shape).copy()
        off += cnt
    return out, n_layers


def layers_random_init(n_layers: int, seed: int = 0, residual_zero: bool = False) -> Dict[str, np.ndarray]:
    """Seeded init of the transformer layers (no LightGlue checkpoint is reachable offline).

    ``residual_zero=True`` zeroes every block's last linear layer (the usual zero-init of a residual
    branch): each block then adds exactly 0 to the residual stream, so untrained layers run their full
    arithmetic yet leave the descriptors — and therefore the matches of the trained head — unchanged.
    """
    rng = np.random.default_rng(seed)
    p: Dict[str, np.ndarray] = {}
    for name, shape in layer_tensors(n_layers).items():
        if name == "lg.pos.weight":
            p[name] = (rng.standard_normal(shape) * 4.0).astype(np.float32)
        elif name.endswith("ln.weight"):
            p[name] = (1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif name.endswith(".bias"):
            p[name] = (0.05 * rng.standard_normal(shape)).astype(np.float32)
        else:
            p[name] = (rng.standard_normal(shape) / np.sqrt(shape[1])).astype(np.float32)
        if residual_zero and ".fc2." in name:
            p[name] = np.zeros(shape, np.float32)
    return p
