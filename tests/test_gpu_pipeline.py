"""kvfe_pipeline_* (native dispatcher threads, zero-copy image fetch, SM-published outputs): byte-identical
packets and rectified images versus the blocking kvfe_frontend_step on the same inputs, for pageable
(staged), pinned (zero-copy) and device-resident images, for both rotation input modes, with frames
queued ahead and several streams completing out of order."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from kimera_vio_b200 import lib as kl

pytestmark = pytest.mark.gpu


def mat3(a, b):
    """3x3 product in the device's operation order (common.cuh matmul3 / dot3s: a0*b0 + (a1*b1 + a2*b2))."""
    a, b = np.asarray(a, np.float64).reshape(3, 3), np.asarray(b, np.float64).reshape(3, 3)
    c = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            c[i, j] = float(a[i, 0]) * float(b[0, j]) + (float(a[i, 1]) * float(b[1, j]) + float(a[i, 2]) * float(b[2, j]))
    return c


def reference_run(p, rig, tail, frames, rot_of):
    """Blocking kvfe_frontend_step over one stream; returns raw packets, parsed packets, rectified pairs."""
    ctx = kl.Context(kl.make_config(p, rig.W, rig.H, batch=1, sobel_cpu_tail_start=tail), rig.to_c())
    out, lkf = [], 0
    for k, f in enumerate(frames):
        R = rot_of(k, lkf)
        pk = ctx.step([f.left], [f.right], [f.timestamp], np.array([R]), want_rectified=True)[0]
        pk["R_used"] = R
        out.append(pk)
        if pk["is_keyframe"]:
            lkf = k
    ctx.close()
    return out


def same_packet(a, b):
    for name, _, _ in kl.PACKET_FIELDS:
        x, y = a[name], b[name]
        if x.shape != y.shape or not np.array_equal(x, y, equal_nan=True):
            return False, name
    for k in ("n", "is_keyframe", "mono_status", "stereo_status", "n_smart", "nr_tracked", "mode", "frame_id", "timestamp",
              "nr_mono_inliers", "nr_stereo_inliers", "median_disparity"):
        if a[k] != b[k]:
            return False, k
    for k in ("lkf_T_k_mono", "lkf_T_k_stereo", "info_stereo"):
        if not np.array_equal(a[k], b[k]):
            return False, k
    return True, ""


@pytest.mark.parametrize("memory", ["pageable", "pinned", "device"])
def test_pipeline_matches_blocking_step(memory):
    import torch
    N, S = 10, 3
    p, rig, ctx0 = H.euroc_setup(batch=1)
    ctx0.close()
    tail = H.sobel_cpu_tail_start(rig.W)
    seqs = [H.synth_frames(N, seed=20240 + 1000 * s) for s in range(S)]
    refs = [reference_run(p, rig, tail, fr, lambda k, l, st=st: st.kf_rotation(l, k)) for st, fr in seqs]
    cfg = kl.make_config(p, rig.W, rig.H, batch=1, sobel_cpu_tail_start=tail)
    pipe = kl.Pipeline(cfg, rig.to_c(), n_streams=S, n_workers=2, queue_depth=2, output_slots=3, want_rectified=True,
                       rotation_mode=0, checksum_outputs=True)
    keep = []
    def buf(img):
        if memory == "pageable":
            a = np.ascontiguousarray(img); keep.append(a); return a.ctypes.data
        t = torch.from_numpy(np.ascontiguousarray(img))
        t = t.pin_memory() if memory == "pinned" else t.cuda()
        keep.append(t)
        return t.data_ptr()
    # rotation mode 0: frame k of a stream can only be pushed once frame k-1's keyframe decision is known
    nxt, lkf, done = [0] * S, [0] * S, 0
    for s in range(S):
        f = seqs[s][1][0]
        assert pipe.push(s, buf(f.left), buf(f.right), rig.W, f.timestamp, seqs[s][0].kf_rotation(0, 0), tag=0)
    bad = []
    while done < S * N:
        outs = pipe.pop(timeout_ms=20000)
        assert outs, "pipeline stalled"
        for o in outs:
            s, k = o.stream, int(o.tag)
            d = pipe.parse(o)
            ok, why = same_packet(d, refs[s][k])
            if ok and d["is_keyframe"]:
                ok = np.array_equal(d["left_rect"], refs[s][k]["left_rect"]) and np.array_equal(d["right_rect"], refs[s][k]["right_rect"])
                why = "rectified images"
            if not ok:
                bad.append((s, k, why))
            if d["is_keyframe"]:
                lkf[s] = k
            done += 1
            if k + 1 < N:
                f = seqs[s][1][k + 1]
                assert pipe.push(s, buf(f.left), buf(f.right), rig.W, f.timestamp, seqs[s][0].kf_rotation(lkf[s], k + 1), tag=k + 1)
        pipe.release(outs)
    st = pipe.stats()
    H.diag("pipeline_parity", memory=memory, bad=bad, **st)
    pipe.close()
    assert not bad, bad
    assert st["graph_launches"] in (S * N, 2 * S * N)      # one graph per frame, or two with the split step graphs (default)
    assert (st["staged_copies"] > 0) == (memory == "pageable")


@pytest.mark.parametrize("prefetch,split", [(0, 0), (1, 0), (0, -1)])
def test_pipeline_relative_rotations_queue_ahead(prefetch, split):
    """rotation_mode 1: the front-end accumulates the frame-to-frame rotations itself, so every frame of every
    stream is pushed up front; packets equal the blocking step fed with the same accumulated product.
    prefetch=1: the next frame's images are pulled into the staging slot by the side branch of the step graph."""
    import torch
    N, S = 12, 4
    p, rig, ctx0 = H.euroc_setup(batch=1)
    ctx0.close()
    tail = H.sobel_cpu_tail_start(rig.W)
    seqs = [H.synth_frames(N, seed=20240 + 1000 * (s % 2)) for s in range(S)]
    rel = [[np.eye(3)] + [st.kf_rotation(k - 1, k) for k in range(1, N)] for st, _ in seqs]
    refs = []
    for s, (st, fr) in enumerate(seqs):
        acc = {"R": np.eye(3), "last": -1}
        def rot_of(k, lkf, s=s, acc=acc):
            # device semantics: acc = I after a keyframe, else the previous lkf_R_km1; lkf_R_k = acc * km1_R_k
            base = np.eye(3) if lkf == k - 1 or k == 0 else acc["R"]
            R = mat3(base, rel[s][k])
            acc["R"] = R
            return R
        refs.append(reference_run(p, rig, tail, fr, rot_of))
    cfg = kl.make_config(p, rig.W, rig.H, batch=1, sobel_cpu_tail_start=tail)
    pipe = kl.Pipeline(cfg, rig.to_c(), n_streams=S, n_workers=2, queue_depth=N, output_slots=4, want_rectified=True,
                       rotation_mode=1, checksum_outputs=True, prefetch=prefetch, split_graphs=split)
    keep = []
    for k in range(N):
        for s in range(S):
            f = seqs[s][1][k]
            l, r = torch.from_numpy(f.left).pin_memory(), torch.from_numpy(f.right).pin_memory()
            keep += [l, r]
            assert pipe.push(s, l.data_ptr(), r.data_ptr(), rig.W, f.timestamp, rel[s][k], tag=k)
    done, bad, sums = 0, [], {}
    while done < S * N:
        outs = pipe.pop(timeout_ms=20000)
        assert outs, "pipeline stalled"
        for o in outs:
            d = pipe.parse(o)
            ok, why = same_packet(d, refs[o.stream][int(o.tag)])
            if not ok:
                bad.append((o.stream, int(o.tag), why))
            sums[(o.stream, int(o.tag))] = int(o.checksum)
            done += 1
        pipe.release(outs)
    pipe.close()
    assert not bad, bad
    # streams s and s+2 replay the same sequence: identical outputs, hence identical checksums
    assert all(sums[(s, k)] == sums[(s + 2, k)] for s in range(2) for k in range(N))
