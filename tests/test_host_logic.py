"""CPU tests of the host-side pieces: parameter parsing (reference YAMLs), config marshalling,
the synthetic generator, stream sharding and the 2-rank gloo packet gather."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REF, ROOT, needs_reference
from kimera_vio_b200 import dist as kd
from kimera_vio_b200.params import CameraParams, FrontendParams
from kimera_vio_b200.rig import StereoRigSetup
from kimera_vio_b200.synth import SynthStream


@needs_reference
def test_euroc_yaml_equals_inlined_defaults():
    p = FrontendParams.from_yaml(os.path.join(REF, "params/Euroc/FrontendParams.yaml"))
    q = FrontendParams.euroc()
    for k, v in vars(q).items():
        w = getattr(p, k)
        if isinstance(v, np.ndarray):
            assert np.array_equal(v, w), k
        else:
            assert v == w, (k, v, w)
    l = CameraParams.from_yaml(os.path.join(REF, "params/Euroc/LeftCameraParams.yaml"))
    assert l.intrinsics == CameraParams.euroc_left().intrinsics
    assert np.allclose(l.T_BS, CameraParams.euroc_left().T_BS)
    assert l.distortion == CameraParams.euroc_left().distortion


@needs_reference
@pytest.mark.parametrize("rig", ["Euroc", "uHumans2", "D455", "KinectAzure", "EurocMono", "RealSenseIR", "Kitti"])
def test_all_shipped_rigs_parse(rig):
    path = os.path.join(REF, "params", rig, "FrontendParams.yaml")
    if not os.path.exists(path):
        pytest.skip("rig not shipped")
    p = FrontendParams.from_yaml(path)
    assert p.feature_detector_type == 3 and p.non_max_suppression_type == 6


def test_config_marshalling_roundtrip():
    from kimera_vio_b200 import lib as kl
    if not os.path.exists(kl.LIB_PATH):
        from kimera_vio_b200 import build
        build.build()
    p = FrontendParams.euroc()
    c = kl.make_config(p, 752, 480, batch=32, sobel_cpu_tail_start=736)
    assert (c.width, c.height, c.batch) == (752, 480, 32)
    assert c.klt_win_size == 24 and c.klt_max_level == 4 and abs(c.klt_eps - 0.1) < 1e-15
    assert c.nr_horizontal_bins == 7 and c.nr_vertical_bins == 5 and sum(c.binning_mask[:35]) == 35
    assert c.min_intra_keyframe_time_ns == 200_000_000 and c.ransac_randomize == 0
    assert c.sobel_cpu_tail_start == 736


def test_synth_stream_deterministic_and_textured():
    rig = StereoRigSetup(CameraParams.euroc_left(), CameraParams.euroc_right())
    assert abs(rig.baseline - 0.110078) < 1e-4
    a = SynthStream(CameraParams.euroc_left(), CameraParams.euroc_right(), rig.R1, seed=7)
    b = SynthStream(CameraParams.euroc_left(), CameraParams.euroc_right(), rig.R1, seed=7)
    fa, fb = a.frame(3), b.frame(3)
    assert np.array_equal(fa.left, fb.left) and np.array_equal(fa.right, fb.right)
    assert fa.left.shape == (480, 752) and fa.left.std() > 10
    R = a.kf_rotation(0, 3)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
    assert np.allclose(a.kf_rotation(2, 2), np.eye(3))


def test_stream_shard_partitions():
    for n, w in ((32, 1), (32, 2), (32, 8), (33, 8), (5, 8)):
        seen = []
        for r in range(w):
            b, e = kd.stream_shard(n, w, r)
            seen += list(range(b, e))
        assert seen == list(range(n))


def _gloo_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = kd.stream_shard(6, world, rank)
    local = torch.full((3, 16), rank + 1, dtype=torch.uint8)      # 3 streams x 16-byte "packets"
    local[:, 0] = torch.arange(b, e, dtype=torch.uint8)
    out = kd.gather_packets(local, dst=0)
    ms = kd.max_over_ranks(10.0 + rank)
    if rank == 0:
        q.put((torch.cat(out)[:, 0].tolist(), [int(t[0, 1]) for t in out], ms))
    dist.destroy_process_group()


def test_gloo_world2_gather_packets():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ids, tags, ms = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ids == [0, 1, 2, 3, 4, 5]          # stream-major order: rank 0's block, then rank 1's
    assert tags == [1, 2]
    assert ms == 11.0                          # max over ranks


def _dt_host_lib():
    """Builds tests/native/dt_host.cpp (the mesh kernel's quad-edge code, delaunay.cuh, compiled for the host)."""
    import ctypes as C
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libdt_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "native", "dt_host.cpp")])
    return C.CDLL(so)


def test_delaunay_restatement_matches_subdiv2d():
    """delaunay.cuh (what mesh_kernel executes) against cv2.Subdiv2D through oracle/mesher.py: identical triangle
    LISTS -- order and first vertex included -- on random sub-pixel points, integer corners (collinear and
    cocircular sets), a coarse grid with duplicates, and points partly outside the image."""
    import ctypes as C
    from oracle.mesher import create_mesh_2d_impl
    lib = _dt_host_lib()
    rng = np.random.default_rng(1)
    for trial in range(120):
        w, h = (752, 480) if trial % 2 else (1280, 720)
        n = int(rng.integers(1, 600))
        kind = trial % 4
        if kind == 0:
            pts = rng.uniform(0, [w, h], (n, 2))
        elif kind == 1:
            pts = np.floor(rng.uniform(0, [w, h], (n, 2)))
        elif kind == 2:
            pts = np.floor(rng.uniform(0, [w / 20, h / 20], (n, 2))) * 20
        else:
            pts = rng.uniform(-5, [w + 5, h + 5], (n, 2))
        pts = np.ascontiguousarray(pts, np.float32)
        tri = np.zeros((4 * n + 16, 6), np.float32)
        nq = C.c_int()
        m = lib.dt_host_mesh(w, h, pts.ctypes.data_as(C.c_void_p), n, tri.ctypes.data_as(C.c_void_p), len(tri), C.byref(nq))
        assert m >= 0
        ref = create_mesh_2d_impl((w, h), [tuple(p) for p in pts])
        assert tri[:m].shape == ref.shape and np.array_equal(tri[:m], ref), (trial, kind, n)
        assert nq.value <= 3 * n + 16
