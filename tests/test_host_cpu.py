"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute calls without a GPU), struct layouts, record format, sharding logic and a
world_size-2 gloo run of the multi-process plumbing."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import gisnav_b200
from gisnav_b200 import _lib, keypoint_record as kr, sharding, weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gisnav_b200.h")).read()
    declared = set(re.findall(r"\b(gnb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"gnb_ctx"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None


def test_default_config_and_struct_layout():
    lib = _lib.load()
    c = _lib.GnbConfig()
    assert lib.gnb_default_config(C.byref(c)) == 0
    assert (c.max_keypoints, c.nms_radius, c.border, c.min_matches) == (1024, 4, 4, 15)
    assert abs(c.match_threshold - 0.5) < 1e-9 and abs(c.keypoint_threshold - 0.005) < 1e-9 and c.reproj_px == 8.0
    assert C.sizeof(_lib.GnbConfig) == 17 * 4
    assert C.sizeof(_lib.GnbPoseResult) == 6 * 4 + 22 * 8
    d = gisnav_b200.Config()
    assert d.max_keypoints == c.max_keypoints and d.ransac_iters == c.ransac_iters


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.GnbError) as e:
        gisnav_b200.Context(weights=W.pack(W.random_init(0)))
    assert e.value.code == _lib.GNB_E_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gisnav_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle/" not in src or f.endswith((".cu", ".cuh", ".py")), f  # comments may cite it


def test_weight_blob_roundtrip():
    p = W.random_init(1)
    blob = W.pack(p)
    assert len(blob) == W.BLOB_BYTES and W.N_FLOATS == 1366914
    q = W.unpack(blob)
    for k in p:
        np.testing.assert_array_equal(p[k], q[k])
    with pytest.raises(ValueError):
        W.unpack(blob[:-4])
    # SuperPoint's published parameter count
    assert sum(int(np.prod(s)) for n, s in W.TENSORS.items() if not n.startswith("match.")) == 1300865


def test_keypoint_record_layout():
    # byte-compatible with KEYPOINT_DTYPE, ros/gisnav/gisnav/core/_shared.py:26-35
    assert kr.KEYPOINT_DTYPE.itemsize == 532 and kr.KEYPOINT_DTYPE.fields["descriptor"][1] == 20
    assert kr.KEYPOINT_DTYPE_256.itemsize == 1044
    rng = np.random.default_rng(0)
    xy = rng.random((7, 2)).astype(np.float32) * 100
    desc = rng.random((7, 256)).astype(np.float32)
    buf = kr.encode(xy, desc)
    assert len(buf) == 7 * 1044
    out = kr.decode(buf, 256)
    np.testing.assert_array_equal(out["xy"], xy)
    np.testing.assert_array_equal(out["descriptor"], desc)
    assert np.all(out["size"] == 1) and np.all(out["angle"] == 0)
    assert kr.decode(b"", 256)["xy"].shape == (0, 2)


def test_shard_ranges_cover_exactly_once():
    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen.extend(range(lo, hi))
            assert seen == list(range(n))
            sizes = [sharding.shard_range(n, r, world) for r in range(world)]
            assert max(h - l for l, h in sizes) - min(h - l for l, h in sizes) <= 1
    assert [sharding.frame_owner(i, 4) for i in range(6)] == [0, 1, 2, 3, 0, 1]
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


_GLOO_WORKER = r'''
import os, sys, hashlib
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from gisnav_b200 import sharding, weights as W
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
blob = W.pack(W.random_init(0)) if rank == 0 else None
t = sharding.broadcast_weights(blob, W.BLOB_BYTES)
digest = hashlib.sha256(t.numpy().tobytes()).hexdigest()
lo, hi = sharding.shard_range(9, rank, 2)
tot = sharding.gather_counts(np.array([hi - lo, float(sum(range(lo, hi)))]))
print("DIGEST", rank, digest)
if rank == 0:
    print("RESULT", int(tot[0]), int(tot[1]))
dist.barrier(); dist.destroy_process_group()
'''


def test_two_process_gloo_broadcast_and_sharding(tmp_path):
    import hashlib
    import socket

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    want = hashlib.sha256(W.pack(W.random_init(0))).hexdigest()
    for r in range(2):  # every byte of the blob arrived on the non-source rank too
        dig = [l for l in outs[r][0].splitlines() if l.startswith("DIGEST")][0].split()
        assert dig[1] == str(r) and dig[2] == want
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT")][0].split()
    assert (int(line[1]), int(line[2])) == (9, 36)  # 9 pairs processed exactly once across 2 ranks


def test_crs_string_ingest_matches_reference_parser():
    from gisnav_b200 import crs, synth
    from oracle import tail_ref

    a = synth.tile_affine(1234.0, 567.0)
    s_ref = tail_ref.affine_to_proj(a)  # restatement of _transformations.py:274-298
    np.testing.assert_array_equal(crs.proj_to_affine(s_ref), tail_ref.proj_to_affine(s_ref))
    np.testing.assert_array_equal(crs.proj_to_affine(crs.affine_to_proj(a)), a)
    np.testing.assert_array_equal(tail_ref.proj_to_affine(crs.affine_to_proj(a)), a)
    with pytest.raises(ValueError):
        crs.proj_to_affine("+proj=affine +xoff=1")


def test_stereo_host_helpers_match_oracle():
    """Host-side bookkeeping of StereoAligner (yaw bucket, CRS composition) against the oracle restatement of
    stereo_node.py:136-168,208-216; the pixel kernel itself is covered by the GPU tests."""
    from gisnav_b200 import crs, stereo, synth
    from oracle import stereo_ref

    for yaw, roll in ((0, 0), (22, 0), (23, 0), (67.9, 0), (200, 30), (359.9, 0), (-10, 0), (44, 2), (350, 15)):
        assert stereo.map_rotation(yaw, roll) == stereo_ref.map_rotation(yaw, roll)
    a = synth.tile_affine(321.0, 654.0)
    _, inv = stereo_ref.rotate_and_crop_center(np.zeros((149, 149), np.uint8), 135, (72, 104))
    got = stereo.world_to_reference_affine(inv, a)
    np.testing.assert_array_equal(got, stereo_ref.world_to_reference_affine(inv, a))
    # the composed matrix survives the proj-string round trip the message carries (stereo_node.py:257-260)
    np.testing.assert_array_equal(crs.proj_to_affine(crs.affine_to_proj(got)), got)


def test_layer_blob_roundtrip_and_validation():
    """GNBL layer blob (transformer layers in front of the matcher head): layout, round trip, rejection of bad blobs."""
    from gisnav_b200 import weights as W

    n_layers = 2
    table = W.layer_tensors(n_layers)
    assert list(table)[0] == "lg.pos.weight" and table["lg.pos.weight"] == (32, 2)
    assert table["lg.1.cross.fc1.weight"] == (512, 512) and table["lg.0.self.fc2.weight"] == (256, 512)
    per_block = 4 * (256 * 256 + 256) + 512 * 512 + 3 * 512 + 256 * 512 + 256
    assert W.layer_floats(n_layers) == 64 + 2 * n_layers * per_block      # what csrc/lightglue.cu expects
    p = W.layers_random_init(n_layers, seed=5)
    blob = W.pack_layers(p, n_layers)
    assert blob[:4] == b"GNBL" and len(blob) == 16 + 4 * W.layer_floats(n_layers)
    q, n = W.unpack_layers(blob)
    assert n == n_layers and all(np.array_equal(p[k], q[k]) for k in p)
    z = W.layers_random_init(n_layers, seed=5, residual_zero=True)
    assert all(not z[k].any() for k in z if ".fc2." in k) and z["lg.0.self.fc1.weight"].any()
    for bad in (blob[:-4], b"GNBW" + blob[4:], blob[:4] + (99).to_bytes(4, "little") + blob[8:]):
        with pytest.raises(ValueError):
            W.unpack_layers(bad)
    p["lg.0.self.q.weight"] = np.zeros((255, 256), np.float32)
    with pytest.raises(ValueError):
        W.pack_layers(p, n_layers)


def test_stereo_aligner_rotation_cache_semantics():
    """StereoAligner.pnp_image re-warps the raster only when the 45-degree yaw bucket changes, like StereoNode
    (stereo_node.py:218-227,262-265).  The pixel kernel is stubbed out: this is host logic only."""
    from gisnav_b200 import crs, stereo, synth
    from oracle import stereo_ref

    sa = object.__new__(stereo.StereoAligner)          # no Context: align() is replaced below
    sa._previous_map_rotation, sa._cached = None, None
    calls = []

    def fake_align(ortho, dem, angle, shape):
        calls.append(angle)
        _, inv = stereo_ref.rotate_and_crop_center(np.zeros(ortho.shape[:2], np.uint8), angle, shape)
        return np.full(shape, angle % 256, np.uint8), np.zeros(shape, np.uint8), inv

    sa.align = fake_align
    ortho, dem = np.zeros((149, 149, 3), np.uint8), np.zeros((149, 149), np.uint8)
    proj = crs.affine_to_proj(synth.tile_affine(10.0, 20.0))
    out = [sa.pnp_image((72, 104), ortho, dem, proj, yaw) for yaw in (3.0, 20.0, 24.0, 40.0, 80.0, 75.0, 359.0)]
    # buckets: 0, 0, 45, 45, 90, 90, 0  ->  warps at the first frame and at every bucket change
    assert calls == [0, 45, 90, 0]
    assert out[0] is out[1] and out[2] is out[3] and out[4] is out[5]
    for (ref, _, proj_str), angle in zip((out[0], out[2], out[4]), (0, 45, 90)):
        assert ref[0, 0] == angle
        _, inv = stereo_ref.rotate_and_crop_center(np.zeros((149, 149), np.uint8), angle, (72, 104))
        want = stereo_ref.world_to_reference_affine(inv, crs.proj_to_affine(proj))
        np.testing.assert_array_equal(crs.proj_to_affine(proj_str), want)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the B200 arm): one JSON line with the contract keys."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-pairs", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "matched_frame_pairs_per_sec" and line["unit"] == "pairs/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["config"]["workload"].startswith("config 3") and line["config"]["config"] == 3
    assert line["dtype"] == "f32"        # the reference's tensors are fp32 (pose_node.py:254-287): no bf16 emulation on this arm
    assert "iterationsCount=10" in line["config"]["ransac"] and line["cpu_baseline"]["cv2_ransac_iterations"] == 10
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and line["matched_fraction"] == 1.0


def test_bench_roofline_table_and_config_contract():
    """bench.py host logic without a GPU: every BASELINE config is defined with the keys the JSON line needs, and the
    per-kernel roofline table maps event-profile rows to their algorithmic work and governing roof."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert sorted(bench.CONFIGS) == [1, 2, 3, 4, 5]
    assert bench.CONFIGS[3]["batch"] == 64 and bench.CONFIGS[5]["batch"] == 512 and bench.CONFIGS[5]["keypoints"] == 2048
    assert bench.CONFIGS[5]["scaling"] == "strong" and bench.CONFIGS[3]["scaling"] == "weak" and bench.CONFIGS[2]["batch"] == 1
    assert abs(bench.FLOP_PER_PIXEL - 169608.0) < 1e-6      # 84 804 MAC per input pixel
    px = 64 * (720 * 1280 + 1024 * 1024)
    prof = {"conv_x3:1b": (100.0, 10), "nms_sparse_kernel": (5.0, 10), "match_pair_x3<0>": (1.0, 5), "match_pair_x3<1>": (1.5, 5),
            "match_collse_kernel": (0.05, 5), "topk_kernel": (0.5, 10), "some_future_kernel": (0.1, 5)}
    rows = bench.kernel_rooflines(prof, 5, px, 128, 64, 1024, 2048, True, {"bf16_tflops_sustained": 1200.0, "hbm_gbs": 6000.0})
    by = {r["kernel"]: r for r in rows}
    c1b = by["conv_x3:1b"]
    assert c1b["bound"] == "tensor" and abs(c1b["peak"] - 400.0) < 1e-9                      # three MMAs per product: peak / 3
    assert abs(c1b["work_per_step"] - 2 * 36864 * px / 1e12) < 1e-9 and abs(c1b["frac"] - c1b["achieved"] / 400.0) < 1e-12
    assert by["nms_sparse_kernel"]["bound"] == "hbm" and abs(by["nms_sparse_kernel"]["work_per_step"] - 4 * px / 1e9) < 1e-9
    m = by["matcher S = m_a m_b^T (all passes)"]
    assert abs(m["ms_per_step"] - (1.0 + 1.5 + 0.05) / 5) < 1e-9 and abs(m["work_per_step"] - 2 * 64 * 1024 * 1024 * 256 / 1e12) < 1e-12
    assert by["topk_kernel"]["bound"] == "latency" and by["some_future_kernel"]["bound"] == "latency"
    cfg = bench.workload_config(type("A", (), {"config": 5, "matcher_layers": 0})(), bench.CONFIGS[5])
    assert cfg["pairs_per_step_total"] == 512 and cfg["pairs_per_step_per_gpu"] is None and "iterationsCount=10" in cfg["ransac"]
