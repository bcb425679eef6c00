#!/usr/bin/env python
"""bench.py -- stereo front-end frame-pairs/s on synthetic 752x480 Euroc-shaped input.

  python bench.py --gpus N --steps K --warmup W          (torchrun for N > 1, one rank per GPU)
  python bench.py --impl reference ...                   (the reference's OpenCV CPU path: oracle/)

A "step" = one pass of the whole hot path over one batch of `--batch` (default 32) stereo
frame-pairs, one per independent camera stream (BASELINE.json configs[1]).  `value` is measured
with the batch already resident in HBM (kvfe_frontend_step_dev); `e2e` goes through the
reference-facing C-ABI call with HOST buffers (kvfe_frontend_step: H2D of both images of every
pair, D2H of every output packet inside the timed region).  Streams shard across ranks with no
data-path collective (weak scaling: the per-GPU batch is fixed).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from kimera_vio_b200.params import CameraParams, FrontendParams  # noqa: E402
from kimera_vio_b200.rig import StereoRigSetup  # noqa: E402
from kimera_vio_b200.synth import SynthStream  # noqa: E402

W, H, N_FEATS = 752, 480, 300
POOL_STREAMS = 4              # distinct synthetic streams; batch slot b replays stream b % POOL_STREAMS
DT_NS = 50_000_000


def frame_pool(n_frames: int, rig: StereoRigSetup):
    """POOL_STREAMS synthetic sequences of n_frames pairs, cached under /tmp (generation is numpy)."""
    cache = "/tmp/kvfe_bench_pool_%dx%d_%d_%d.npz" % (W, H, POOL_STREAMS, n_frames)
    if os.path.exists(cache):
        try:
            z = np.load(cache)
            return z["left"], z["right"], z["rot"]
        except Exception:
            pass                                   # unreadable cache: regenerate (deterministic)
    left = np.zeros((POOL_STREAMS, n_frames, H, W), np.uint8)
    right = np.zeros_like(left)
    rot = np.zeros((POOL_STREAMS, n_frames, n_frames, 3, 3))     # rot[s, lkf, k]
    for s in range(POOL_STREAMS):
        st = SynthStream(CameraParams.euroc_left(), CameraParams.euroc_right(), rig.R1, seed=20240 + 1000 * s)
        for k in range(n_frames):
            f = st.frame(k)
            left[s, k], right[s, k] = f.left, f.right
        for a in range(n_frames):
            for k in range(n_frames):
                rot[s, a, k] = st.kf_rotation(a, k)
    try:
        # several ranks may build the cache at once: write privately, publish atomically
        tmp = "%s.%d.tmp" % (cache, os.getpid())
        with open(tmp, "wb") as fh:
            np.savez(fh, left=left, right=right, rot=rot)
        os.replace(tmp, cache)
    except Exception:
        pass
    return left, right, rot


def slot_timestamp(b: int, k: int) -> int:
    # staggers the keyframe cadence across batch slots: slot b's first gap is (1 + b % 4) periods
    return 1403715273262142976 + (k + (b % 4 if k >= 1 else 0)) * DT_NS


class ClockSampler:
    def __init__(self, gpu_index: int):
        self.rows, self.stop = [], False
        self.idx = gpu_index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# the oracle front-end as the timed CPU baseline (cv2 == the reference's OpenCV code path)
# ------------------------------------------------------------------------------------------------
def _oracle_worker(args):
    stream_id, n_warm, n_timed, threads = args
    import cv2
    from oracle import frontend as ofe
    from oracle.rig import StereoRig
    cv2.setNumThreads(threads)
    rig = StereoRig(CameraParams.euroc_left(), CameraParams.euroc_right())
    setup = StereoRigSetup(CameraParams.euroc_left(), CameraParams.euroc_right())
    left, right, rot = frame_pool(n_warm + n_timed, setup)
    s = stream_id % POOL_STREAMS
    fe = ofe.StereoFrontend(FrontendParams.euroc(), rig)
    lkf = 0
    t0 = None
    n_kf = 0
    for k in range(n_warm + n_timed):
        if k == n_warm:
            t0 = time.perf_counter()
        sf = ofe.StereoFrame.make(k, slot_timestamp(stream_id, k), left[s, k], right[s, k], rig)
        o = fe.spin(sf, rot[s, lkf, k])
        if o.is_keyframe:
            lkf = k
            n_kf += k >= n_warm
    return time.perf_counter() - t0, n_timed, n_kf


def cpu_baseline_single(n_warm: int, n_timed: int):
    dt, n, n_kf = _oracle_worker((0, n_warm, n_timed, 1))
    return {"value": n / dt, "unit": "frame-pairs/s", "cores": 1, "kind": "port",
            "sample": "oracle (cv2 4.13 + numpy RANSAC), 1 stream, %d timed pairs after %d warm-up, %d keyframes, "
                      "cv2.setNumThreads(1)" % (n, n_warm, n_kf)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the oracle: same OpenCV
    calls), one single-threaded process per host core, one camera stream each."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = len(os.sched_getaffinity(0))
    nproc = max(1, min(ncores, 32))     # 32 / 64 / 128 processes were tried on the 128-thread box: 32 is the fastest
    setup = StereoRigSetup(CameraParams.euroc_left(), CameraParams.euroc_right())
    frame_pool(args.warmup + args.steps, setup)          # build the cache once
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(nproc) as pool:
        res = pool.map(_oracle_worker, [(i, args.warmup, args.steps, 1) for i in range(nproc)])
    wall = max(r[0] for r in res)
    total = sum(r[1] for r in res)
    value = total / wall
    line = {
        "impl": "reference", "metric": "stereo front-end frame-pairs/sec @ 752x480", "value": value,
        "unit": "frame-pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32/f32/f64 (OpenCV CPU)", "data": "synthetic",
        "config": {"workload": "Euroc stereo 752x480, %d feats (BASELINE.json configs[1] workload), reference CPU "
                               "front-end: %d independent streams on %d host processes (1 thread each), "
                               "oracle = the reference's OpenCV calls through cv2" % (N_FEATS, nproc, nproc)},
        "cpu_baseline": {"value": value, "unit": "frame-pairs/s", "cores": nproc, "kind": "port",
                         "sample": "%d streams x %d timed pairs, one single-threaded process per core" % (nproc, args.steps)},
        "e2e": {"value": value, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cores_available": ncores, "wall_s_incl_setup": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from kimera_vio_b200 import lib as kl

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, K, Wm, NC = args.batch, args.steps, args.warmup, args.contexts
    assert B % NC == 0
    Bc = B // NC                                   # streams per context (sub-batch in flight on its own CUDA stream)
    n_frames = Wm + K
    rig = StereoRigSetup(CameraParams.euroc_left(), CameraParams.euroc_right())
    left, right, rot = frame_pool(n_frames, rig)
    p = FrontendParams.euroc()
    ctxs = [kl.Context(kl.make_config(p, W, H, batch=Bc), rig.to_c()) for _ in range(NC)]
    streams = [torch.cuda.ExternalStream(kl.load().kvfe_cuda_stream(c.h)) for c in ctxs]
    pkb = ctxs[0].packet_bytes

    def slot(c, i):                                # global batch slot of context c, local stream i
        return c * Bc + i

    # IMU rotations need the last-keyframe index per slot, which depends on the device-side keyframe
    # decisions: run the sequence once (untimed, host path) to learn the keyframe schedule.
    ts_all = np.array([[slot_timestamp(rank * B + b, k) for b in range(B)] for k in range(n_frames)], np.int64)
    lkf = np.zeros(B, np.int64)
    R_all = np.zeros((n_frames, B, 9))
    kf_sched = np.zeros((n_frames, B), bool)
    pk_buf = np.empty(Bc * pkb, np.uint8)
    n_kp = []
    for k in range(n_frames):
        for b in range(B):
            R_all[k, b] = rot[b % POOL_STREAMS, lkf[b], k].reshape(9)
        for c, ctx in enumerate(ctxs):
            lp = (C.c_void_p * Bc)(*[left[slot(c, i) % POOL_STREAMS, k].ctypes.data for i in range(Bc)])
            rp = (C.c_void_p * Bc)(*[right[slot(c, i) % POOL_STREAMS, k].ctypes.data for i in range(Bc)])
            tsk = np.ascontiguousarray(ts_all[k, c * Bc:(c + 1) * Bc])
            Rk = np.ascontiguousarray(R_all[k, c * Bc:(c + 1) * Bc])
            rc = ctx.step_raw(lp, rp, W, tsk, Rk, pk_buf)
            assert rc == 0, kl.load().kvfe_last_error(ctx.h)
            for i, pk in enumerate(ctx.parse_packets(pk_buf)):
                if pk["is_keyframe"]:
                    lkf[slot(c, i)] = k
                    kf_sched[k, slot(c, i)] = True
                if k == n_frames - 1:
                    n_kp.append(pk["n"])
    n_kp_mean = float(np.mean(n_kp))
    ts_c = [[np.ascontiguousarray(ts_all[k, c * Bc:(c + 1) * Bc]) for k in range(n_frames)] for c in range(NC)]
    R_c = [[np.ascontiguousarray(R_all[k, c * Bc:(c + 1) * Bc]) for k in range(n_frames)] for c in range(NC)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident measurement (value) ----------------
    dL = torch.empty((NC, n_frames, Bc, H, W), dtype=torch.uint8, device="cuda")
    dR = torch.empty_like(dL)
    for c in range(NC):
        for k in range(n_frames):
            for i in range(Bc):
                dL[c, k, i].copy_(torch.from_numpy(left[slot(c, i) % POOL_STREAMS, k]))
                dR[c, k, i].copy_(torch.from_numpy(right[slot(c, i) % POOL_STREAMS, k]))
    torch.cuda.synchronize()

    # one host call per step enqueues all sub-batches (kvfe_frontend_step_dev_multi)
    lib = kl.load()
    hctx = (C.c_void_p * NC)(*[c.h for c in ctxs])
    multi_args = []
    for k in range(n_frames):
        multi_args.append(((C.c_void_p * NC)(*[dL[c, k].data_ptr() for c in range(NC)]),
                           (C.c_void_p * NC)(*[dR[c, k].data_ptr() for c in range(NC)]),
                           (C.c_void_p * NC)(*[ts_c[c][k].ctypes.data for c in range(NC)]),
                           (C.c_void_p * NC)(*[R_c[c][k].ctypes.data for c in range(NC)])))

    dev_args = [[(ctxs[c].h, C.c_void_p(dL[c, k].data_ptr()), C.c_void_p(dR[c, k].data_ptr()), C.c_size_t(W),
                  C.c_void_p(ts_c[c][k].ctypes.data), C.c_void_p(R_c[c][k].ctypes.data), None)
                 for k in range(n_frames)] for c in range(NC)]
    dev_api = os.environ.get("KVFE_BENCH_DEV_API", "step_dev")    # "submit_dev": measured slower (48 k vs 57 k)

    def run_dev(k0, k1):
        if dev_api == "step_dev":       # one host call per step enqueues every context; unbounded queue depth
            for k in range(k0, k1):
                a = multi_args[k]
                rc = lib.kvfe_frontend_step_dev_multi(hctx, NC, a[0], a[1], C.c_size_t(W), a[2], a[3])
                assert rc == 0
            return
        # kvfe_frontend_submit_dev: images resident in HBM, no copy-engine operation; up to two steps in
        # flight per context, the host only waits for the step before last of a context
        infl = [0] * NC
        for k in range(k0, k1):
            for c in range(NC):
                if infl[c] == 2:
                    rc = lib.kvfe_frontend_wait(ctxs[c].h)
                    assert rc == 0
                    infl[c] -= 1
                rc = lib.kvfe_frontend_submit_dev(*dev_args[c][k])
                assert rc == 0, lib.kvfe_last_error(ctxs[c].h)
                infl[c] += 1
        dev_pending.append(infl)

    dev_pending = []

    def drain_dev():
        for infl in dev_pending:
            for c in range(NC):
                for _ in range(infl[c]):
                    rc = lib.kvfe_frontend_wait(ctxs[c].h)
                    assert rc == 0
        dev_pending.clear()

    sampler = ClockSampler(local)
    for ctx in ctxs:
        ctx.reset()
    run_dev(0, Wm)
    barrier()
    drain_dev()
    launches0 = sum(c.launches for c in ctxs)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in ctxs]
    with sampler:
        ev0.record(streams[0])
        for st in streams[1:]:
            st.wait_event(ev0)
        t_enq = time.perf_counter()
        run_dev(Wm, n_frames)
        t_enq = time.perf_counter() - t_enq
        for e, st in zip(ev1, streams):
            e.record(st)
        barrier()
    drain_dev()
    dev_ms = max(ev0.elapsed_time(e) for e in ev1)
    launches = sum(c.launches for c in ctxs) - launches0
    clocks = sampler.summary()

    # ---------------- dominant kernel (LK) launch duration, live CUDA events ----------------
    # one context holding the whole batch; the step is run stage by stage with events on the
    # library stream (kvfe_frontend_step_dev_timed); inputs: the same fresh device buffers.
    lk_ms, stage_ms = None, None
    if rank == 0:
        big = kl.Context(kl.make_config(p, W, H, batch=B), rig.to_c())
        tsB = [np.ascontiguousarray(ts_all[k]) for k in range(n_frames)]
        RB = [np.ascontiguousarray(R_all[k]) for k in range(n_frames)]
        dLb = dL.permute(1, 0, 2, 3, 4).reshape(n_frames, B, H, W).contiguous()
        dRb = dR.permute(1, 0, 2, 3, 4).reshape(n_frames, B, H, W).contiguous()
        acc = []
        for k in range(n_frames):
            ms = big.step_dev_timed(dLb[k].data_ptr(), dRb[k].data_ptr(), W, tsB[k], RB[k])
            if k >= Wm:
                acc.append(ms)
        stage_ms = np.mean(acc, axis=0)
        lk_ms = float(stage_ms[8])
        big.close()
        del dLb, dRb

    # ---------------- end-to-end measurement through host buffers (e2e) ----------------
    # host frames: for every frame k the images of all sub-batches are contiguous (group uploads)
    pinL = torch.empty((n_frames, NC, Bc, H, W), dtype=torch.uint8).pin_memory()
    pinR = torch.empty_like(pinL).pin_memory()
    for c in range(NC):
        for k in range(n_frames):
            for i in range(Bc):
                pinL[k, c, i].copy_(torch.from_numpy(left[slot(c, i) % POOL_STREAMS, k]))
                pinR[k, c, i].copy_(torch.from_numpy(right[slot(c, i) % POOL_STREAMS, k]))
    # two pinned packet buffers per context (a step's buffer is owned by the library until wait returns)
    pk_pin = [[torch.empty(Bc * pkb, dtype=torch.uint8).pin_memory() for _ in range(2)] for _ in range(NC)]
    sub_args = [[(ctxs[c].h,
                  (C.c_void_p * Bc)(*[pinL[k, c, i].data_ptr() for i in range(Bc)]),
                  (C.c_void_p * Bc)(*[pinR[k, c, i].data_ptr() for i in range(Bc)]),
                  C.c_size_t(W), C.c_void_p(ts_c[c][k].ctypes.data), C.c_void_p(R_c[c][k].ctypes.data),
                  None) for k in range(n_frames)] for c in range(NC)]     # packets: read in place (packets_view)
    hs = [c.h for c in ctxs]

    # staged uploads: the frames of G consecutive sub-batches travel in one H2D copy per camera, issued one
    # step ahead (kvfe_upload_frames); G = 1 falls back to per-context image copies (kvfe_frontend_submit)
    G = max(1, min(args.upload_group, NC))
    ugroups = [list(range(g, min(g + G, NC))) for g in range(0, NC, G)]
    HT = max(1, min(args.host_threads, len(ugroups)))
    uploads, up_args, subu_args = [], [], {}
    if G > 1:
        for cs in ugroups:
            up = C.c_void_p()
            rc = lib.kvfe_upload_create((C.c_void_p * len(cs))(*[ctxs[c].h.value for c in cs]), C.c_int(len(cs)), C.byref(up))
            assert rc == 0, lib.kvfe_last_error(hs[cs[0]])
            uploads.append(up)
            up_args.append([(up, C.c_void_p(pinL[k, cs[0]].data_ptr()), C.c_void_p(pinR[k, cs[0]].data_ptr()), C.c_size_t(W))
                            for k in range(n_frames)])
            for m, c in enumerate(cs):
                subu_args[c] = [(hs[c], up, C.c_int(m), C.c_void_p(ts_c[c][k].ctypes.data),
                                 C.c_void_p(R_c[c][k].ctypes.data), None) for k in range(n_frames)]

    host_prof = [0.0, 0.0, 0.0, 0]          # [unused, unused, seconds inside submit, number of polls] (diagnostic)

    def run_host_thread(gis, k0, k1):
        # one dispatcher thread serving its contexts in COMPLETION order: poll (kvfe_frontend_ready), collect the
        # finished step (kvfe_frontend_wait, packets read in place), submit the context's next frame at once.  A
        # context never has more than one step in flight -- the IMU rotation of frame k depends on frame k-1's
        # keyframe decision -- but contexts advance independently (a keyframe step takes ~3x a tracking step), and
        # with staged uploads (G > 1) the frames of a group travel one step ahead in one H2D copy per camera.
        cs_all = [c for gi in gis for c in ugroups[gi]]
        nxt = {c: k0 for c in cs_all}
        busy = {c: 0 for c in cs_all}
        depth = 1
        up_next = {gi: k0 for gi in gis}
        remaining = len(cs_all) * (k1 - k0)
        ready, wait, submit, submit_u, upload = (lib.kvfe_frontend_ready, lib.kvfe_frontend_wait, lib.kvfe_frontend_submit,
                                                 lib.kvfe_frontend_submit_uploaded, lib.kvfe_upload_frames)
        prof = host_prof
        while remaining:
            for gi in gis:
                cs = ugroups[gi]
                if G > 1:
                    # keep the upload ring one step ahead of the slowest member of the group
                    lo = min(nxt[c] for c in cs)
                    while up_next[gi] < k1 and up_next[gi] <= lo + 1:
                        rc = upload(*up_args[gi][up_next[gi]])
                        assert rc == 0, lib.kvfe_last_error(hs[cs[0]])
                        up_next[gi] += 1
                for c in cs:
                    if busy[c]:
                        prof[3] += 1
                        if ready(hs[c]) == 1:
                            rc = wait(hs[c])
                            assert rc == 0
                            busy[c] -= 1
                            remaining -= 1
                    k = nxt[c]
                    if busy[c] < depth and k < k1 and (G == 1 or k < up_next[gi]):
                        t_a = time.perf_counter()
                        rc = submit_u(*subu_args[c][k]) if G > 1 else submit(*sub_args[c][k])
                        prof[2] += time.perf_counter() - t_a
                        assert rc == 0, lib.kvfe_last_error(hs[c])
                        nxt[c] = k + 1
                        busy[c] += 1

    def run_host(k0, k1):
        parts = [list(range(len(ugroups)))[t::HT] for t in range(HT)]
        if HT == 1:
            return run_host_thread(parts[0], k0, k1)
        th = [threading.Thread(target=run_host_thread, args=(g, k0, k1)) for g in parts]
        for t in th:
            t.start()
        for t in th:
            t.join()

    for ctx in ctxs:
        ctx.reset()
    run_host(0, Wm)
    barrier()
    host_prof[:] = [0.0, 0.0, 0.0, 0]
    t0 = time.perf_counter()
    run_host(Wm, n_frames)
    barrier()
    e2e_s = time.perf_counter() - t0
    print("[e2e host profile] total %.3f s, %d polls, %.3f s inside submit calls" % (e2e_s, host_prof[3], host_prof[2]), file=sys.stderr)
    for up in uploads:
        lib.kvfe_upload_destroy(up)

    # ---------------- host link rate (what bounds e2e) and single-stream latency ----------------
    link, latency = None, None
    if rank == 0:
        # pinned -> device copy of one step's input volume on one stream, CUDA events
        nb = 2 * B * W * H                                   # one step's H2D volume, contiguous
        hbuf, dbuf = pinL.reshape(-1)[:nb], dL.reshape(-1)[:nb]
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dbuf.copy_(hbuf, non_blocking=True)
        torch.cuda.synchronize()
        c0.record()
        for _ in range(8):
            dbuf.copy_(hbuf, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        link = {"h2d_gbs_measured": 8 * hbuf.numel() / (c0.elapsed_time(c1) * 1e-3) / 1e9,
                "copy_bytes": int(hbuf.numel())}
        # one stream alone on the GPU (batch 1, one context): per-frame latency of the host-buffer call
        one = kl.Context(kl.make_config(p, W, H, batch=1), rig.to_c())
        pk1 = np.empty(pkb, np.uint8)
        lat, lat_kf = [], []
        for k in range(n_frames):
            lp = (C.c_void_p * 1)(pinL[k, 0, 0].data_ptr())
            rp = (C.c_void_p * 1)(pinR[k, 0, 0].data_ptr())
            t1 = time.perf_counter()
            rc = one.step_raw(lp, rp, W, ts_c[0][k][:1], R_c[0][k][:1], pk1)
            dt = (time.perf_counter() - t1) * 1e3
            assert rc == 0
            if k >= Wm:
                (lat_kf if one.parse_packets(pk1)[0]["is_keyframe"] else lat).append(dt)
        one.close()
        allv = np.array(lat + lat_kf)
        latency = {"what": "kvfe_frontend_step, host buffers, batch 1, stream alone on the GPU [ms]",
                   "p50": float(np.percentile(allv, 50)), "p99": float(np.percentile(allv, 99)),
                   "non_keyframe_p50": float(np.median(lat)) if lat else None,
                   "keyframe_p50": float(np.median(lat_kf)) if lat_kf else None, "frames": int(allv.size)}

    times = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(times[0]), float(times[1])

    if rank == 0:
        pairs = world * B * K
        value = pairs / (dev_ms * 1e-3)
        e2e_value = pairs / (e2e_ms * 1e-3)
        rho = float(kf_sched[Wm:].mean())
        b_alg = (1 - rho) * W * H + rho * 4 * W * H + 112 * n_kp_mean      # SURVEY 8(d), measured rho
        peaks, which = measured_peaks()
        # LK kernel: compulsory traffic per launch = previous + current pyramids of every stream
        # (levels 0..4, 1.33 * W * H bytes each); DESIGN.md section 3
        pyr_bytes = sum(((W + (1 << l) - 1) >> l) * ((H + (1 << l) - 1) >> l) for l in range(5))
        lk_alg_bytes = 2 * pyr_bytes * B
        achieved = lk_alg_bytes / (lk_ms * 1e-3) / 1e9
        step_achieved = b_alg * (B * K / (dev_ms * 1e-3)) / 1e9          # whole step, per GPU
        line = {
            "metric": "stereo front-end frame-pairs/sec @ 752x480", "value": value, "unit": "frame-pairs/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": dev_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 images, f32 LK+response, f64 geometry",
            "data": "synthetic",
            "config": {"workload": "Euroc stereo 752x480, %d feats, 1xB200 batch=%d frame-pairs per step "
                                   "(BASELINE.json configs[1]); %d independent streams per GPU" % (N_FEATS, B, B),
                       "batch_per_gpu": B, "sub_batches_in_flight": NC, "e2e_dispatcher_threads": HT, "e2e_upload_group": G, "keyframe_ratio": rho,
                       "mean_keypoints": n_kp_mean,
                       "timing": "CUDA events on the library streams, max over ranks; every step reads fresh "
                                 "device-resident inputs (%d MB per rank > L2), no L2 flush" %
                                 (2 * n_frames * B * H * W // 2 ** 20)},
            "e2e": {"value": e2e_value, "unit": "frame-pairs/s", "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": int(2 * B * W * H + B * 80),
                    "d2h_bytes_per_step": int(B * pkb)},
            "gpu_launches": int(launches),
            "host_enqueue_ms_per_step": 1e3 * t_enq / K,
            "host_link": link, "latency_ms": latency,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "lk_kernel_col<24> (pyramidal LK, dominant: 70% of the step's warp instructions, 27% of its serialised kernel time)",
                         "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": 30850048, "peak_source": which + " (burst copy)",
                         "algorithmic_bytes_per_launch": lk_alg_bytes, "launch_ms": lk_ms,
                         "traffic_source": "profiles/r01_ncu_final.txt: dram__bytes_read.sum + dram__bytes_write.sum, one launch, batch 32",
                         "whole_step": {"algorithmic_bytes_per_frame_pair": b_alg, "achieved": step_achieved,
                                        "frac": step_achieved / peaks["hbm_gbs"]},
                         "stage_ms": [float(v) for v in stage_ms],
                         "note": "latency/issue-bound: serial float chains imposed by bit-exactness, see DESIGN.md"},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single(4, 24)
        print(json.dumps(line))
    for ctx in ctxs:
        ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--upload-group", type=int, default=8, help="sub-batches that share one H2D copy per camera (e2e loop)")
    ap.add_argument("--host-threads", type=int, default=1, help="dispatcher threads of the end-to-end (host buffer) loop")
    ap.add_argument("--contexts", type=int, default=32, help="sub-batches in flight on separate CUDA streams")
    ap.add_argument("--impl", default="kvfe", choices=["kvfe", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
