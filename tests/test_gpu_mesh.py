"""Row f1: Mesher::createMesh2dImpl / createMesh2dStereo (reference src/mesh/Mesher.cpp:1712-1817, :1849-1886)
on the GPU against the oracle (oracle/mesher.py = cv2.Subdiv2D): identical triangle lists, stage level
(kvfe_mesh_2d) and inside the frame-level step (cfg.mesh_2d: the mesh of every keyframe joins the packet)."""
import numpy as np
import pytest

import helpers as H
from kimera_vio_b200 import lib as kl
from oracle import mesher as om

pytestmark = pytest.mark.gpu


def test_mesh_2d_stage_matches_subdiv2d():
    p, rig, ctx = H.euroc_setup(batch=1)
    rng = np.random.default_rng(7)
    g, lefts, _ = H.golden()
    det = ctx.detect(lefts[0], [], [], p.max_features_per_frame)            # real sub-pixel corners
    cases = [("detected", det), ("single", det[:1]), ("two", det[:2]),
             ("integer", np.floor(rng.uniform(0, [rig.W, rig.H], (500, 2))).astype(np.float32)),
             ("grid_dups", (np.floor(rng.uniform(0, [rig.W / 20, rig.H / 20], (400, 2))) * 20).astype(np.float32)),
             ("partly_outside", rng.uniform(-5, [rig.W + 5, rig.H + 5], (300, 2)).astype(np.float32))]
    for name, pts in cases:
        tri = ctx.mesh_2d(pts)
        ref = om.create_mesh_2d_impl((rig.W, rig.H), [tuple(q) for q in pts])
        H.diag("mesh_stage", case=name, n=len(pts), n_tri=len(tri), n_ref=len(ref))
        assert tri.shape == ref.shape and np.array_equal(tri, ref), name
    ctx.close()


def test_mesh_in_the_packet():
    """cfg.mesh_2d: every keyframe packet carries createMesh2dStereo's triangles of that frame."""
    N = 9
    p, rig, ctx0 = H.euroc_setup(batch=1)
    ctx0.close()
    cfg = kl.make_config(p, rig.W, rig.H, batch=2, sobel_cpu_tail_start=H.sobel_cpu_tail_start(rig.W), mesh_2d=True)
    ctx = kl.Context(cfg, rig.to_c())
    plain = kl.Context(kl.make_config(p, rig.W, rig.H, batch=2, sobel_cpu_tail_start=H.sobel_cpu_tail_start(rig.W)), rig.to_c())
    seqs = [H.synth_frames(N, seed=20240 + 1000 * b) for b in range(2)]
    lkf, n_kf = [0, 0], 0
    for k in range(N):
        fr = [seqs[b][1][k] for b in range(2)]
        Rs = np.array([seqs[b][0].kf_rotation(lkf[b], k) for b in range(2)])
        pks = ctx.step([f.left for f in fr], [f.right for f in fr], [f.timestamp for f in fr], Rs)
        pk0 = plain.step([f.left for f in fr], [f.right for f in fr], [f.timestamp for f in fr], Rs)
        for b, pk in enumerate(pks):
            assert pk["n"] == pk0[b]["n"] and np.array_equal(pk["kp_x"], pk0[b]["kp_x"])      # the mesh changes nothing else
            if pk["is_keyframe"]:
                lkf[b] = k
                n_kf += 1
                kps = list(zip(pk["kp_x"], pk["kp_y"]))
                ref, _ = om.create_mesh_2d_stereo((rig.W, rig.H), pk["landmark"], pk["right_status"], kps)
                H.diag("mesh_packet", frame=k, stream=b, n_tri=int(pk["n_mesh_triangles"]), n_ref=len(ref))
                assert pk["mesh_tri"].shape == ref.shape and np.array_equal(pk["mesh_tri"], ref)
                assert len(ref) > 100
            else:
                assert pk["n_mesh_triangles"] == 0
    ctx.close()
    plain.close()
    assert n_kf >= 4
