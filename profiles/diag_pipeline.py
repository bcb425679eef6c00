#!/usr/bin/env python
"""Diagnostic (not a bench): device-resident throughput of kvfe_pipeline under variations of its knobs.
  python profiles/diag_pipeline.py [--passes 160] [--config c2]
Prints one line per variant: ms per pass (one pass = `batch` frame-pairs)."""
import argparse
import os
import sys
import time

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--passes", type=int, default=160)
    ap.add_argument("--config", default="c2")
    ap.add_argument("--variants", default="")
    args = ap.parse_args()
    import torch
    from kimera_vio_b200 import lib as kl
    cfg = B.CONFIGS[args.config]
    W, H, nb = cfg["W"], cfg["H"], cfg["batch"]
    PS, NF = cfg["pool_streams"], cfg["pool_frames"]
    left, right, fwd, bwd = B.frame_pool(args.config)
    lcam, rcam, rig = B.config_rig(cfg)
    p = B.config_params(cfg)
    tail = B.sobel_cpu_tail_start(W)
    dL, dR = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    img = W * H
    n_pass = args.passes
    frame_of = np.array([B.pass_frame(t, NF) for t in range(n_pass)], np.int64)
    rot_tab = np.stack([np.stack([B.pass_rotation(fwd, bwd, s, t, NF).reshape(9) for t in range(n_pass)]) for s in range(PS)])

    pL, pR = torch.from_numpy(left).pin_memory(), torch.from_numpy(right).pin_memory()

    def run_variant(name, streams, batch, workers, in_flight, rect, chk, host=False, prefetch=0):
        kcfg = kl.make_config(p, W, H, batch=batch, sobel_cpu_tail_start=tail)
        pipe = kl.Pipeline(kcfg, rig.to_c(), n_streams=streams, n_workers=workers, queue_depth=n_pass + 8, output_slots=4,
                           want_rectified=rect, rotation_mode=1, checksum_outputs=chk, max_in_flight=in_flight, prefetch=prefetch)
        lib, ph = pipe.lib, pipe.h
        OUTS = (kl.PipelineOutput * 1024)()
        n = n_pass * streams
        sidx = np.tile(np.arange(streams, dtype=np.int32), n_pass)
        tt = np.repeat(np.arange(n_pass, dtype=np.int64), streams)
        ss = sidx.astype(np.int64) % PS
        off = (ss * NF + frame_of[tt]) * img
        lp = ((pL if host else dL).data_ptr() + off).astype(np.uint64)
        rp = ((pR if host else dR).data_ptr() + off).astype(np.uint64)
        ts = (B.T0_NS + (tt + np.where(tt >= 1, np.tile(np.arange(streams) % 4, n_pass), 0)) * B.DT_NS).astype(np.int64)
        Rm = np.ascontiguousarray(rot_tab[ss, tt])
        tags = tt.astype(np.uint64)
        warm = 32 * streams
        for (a, b) in ((0, warm), (warm, n)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            acc = lib.kvfe_pipeline_push_many(ph, b - a, sidx[a:].ctypes.data, lp[a:].ctypes.data, rp[a:].ctypes.data, W,
                                              ts[a:].ctypes.data, Rm[a:].ctypes.data, tags[a:].ctypes.data)
            assert acc == b - a
            done = 0
            while done < b - a:
                m = lib.kvfe_pipeline_pop(ph, OUTS, 1024, 5000)
                assert m > 0
                lib.kvfe_pipeline_release(ph, OUTS, m)
                done += m
            dt = time.perf_counter() - t0
        st = pipe.stats()
        frames = (n - warm) * batch
        print("%-44s %8.3f ms/pass(32 fp)  %9.0f fp/s   host launch %.3f ms/pass" %
              (name, 1e3 * dt / (frames / 32.0), frames / dt, 1e3 * st["launch_seconds"] / (n / streams) * 32.0 / (streams * batch)), flush=True)
        pipe.close()

    V = [
        ("32x1 w4 if2 rect chk (bench)", 32, 1, 4, 2, True, True),
        ("32x1 w4 if2 rect nochk", 32, 1, 4, 2, True, False),
        ("32x1 w4 if2 norect nochk", 32, 1, 4, 2, False, False),
        ("32x1 w1 if2 norect nochk", 32, 1, 1, 2, False, False),
        ("32x1 w8 if2 norect nochk", 32, 1, 8, 2, False, False),
        ("32x1 w4 if1 norect nochk", 32, 1, 4, 1, False, False),
        ("16x1 w4 if2 norect nochk", 16, 1, 4, 2, False, False),
        ("8x1 w4 if2 norect nochk", 8, 1, 4, 2, False, False),
        ("HOST 32x1 w4 if2 rect chk (bench e2e)", 32, 1, 4, 2, True, True, True),
        ("HOST 32x1 w4 if2 rect nochk", 32, 1, 4, 2, True, False, True),
        ("HOST 32x1 w4 if2 norect nochk", 32, 1, 4, 2, False, False, True),
        ("HOST 32x1 w8 if2 rect chk", 32, 1, 8, 2, True, True, True),
        ("HOST 32x1 w4 if2 rect chk PREFETCH", 32, 1, 4, 2, True, True, True, 1),
        ("32x1 w8 if2 rect chk", 32, 1, 8, 2, True, True),
        ("32x1 w2 if2 rect chk", 32, 1, 2, 2, True, True),
        ("HOST 32x1 w8 if2 rect chk", 32, 1, 8, 2, True, True, True),
        ("HOST 32x1 w2 if2 rect chk", 32, 1, 2, 2, True, True, True),
        ("32x1 w4 if2 rect chk PREFETCH", 32, 1, 4, 2, True, True, False, 1),
    ]
    want = set(args.variants.split(",")) if args.variants else None
    for i, v in enumerate(V):
        if want is None or str(i) in want:
            run_variant(*v)


if __name__ == "__main__":
    main()
