#!/usr/bin/env python
"""The optional NCCL gather of the keypoint packets to rank 0: device-resident batch steps (one
context per rank, BASELINE configs[1] batch 32) with and without the gather, one rank per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/bench_gather.py
Prints one JSON line on rank 0: ms per step without / with the gather, bytes gathered per step, and a check that
rank 0 received every rank's packets (header fields of the last step)."""
import json
import os
import sys

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from kimera_vio_b200 import lib as kl
    from kimera_vio_b200.dist import PacketGather
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    cfg = bench.CONFIGS["c2"]
    left, right, fwd, bwd = bench.frame_pool("c2")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, PS, NF, B = cfg["W"], cfg["H"], cfg["pool_streams"], cfg["pool_frames"], cfg["batch"]
    lcam, rcam, rig = bench.config_rig(cfg)
    p = bench.config_params(cfg)
    ctx = kl.Context(kl.make_config(p, W, H, batch=B, sobel_cpu_tail_start=bench.sobel_cpu_tail_start(W)), rig.to_c())
    dL, dR = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    idx = torch.tensor([b % PS for b in range(B)], device="cuda")
    N_STEPS = 60
    frames = [(dL[idx, bench.pass_frame(k, NF)].contiguous(), dR[idx, bench.pass_frame(k, NF)].contiguous()) for k in range(NF)]
    res = {}
    for mode in ("plain", "gather"):
        ctx.reset()
        g = PacketGather(ctx) if mode == "gather" else None
        acc = [np.eye(3) for _ in range(B)]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream = torch.cuda.ExternalStream(int(ctx.lib.kvfe_cuda_stream(ctx.h)))
        for k in range(N_STEPS):
            if k == 12:
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                ev0.record(stream)
            bl, br = frames[bench.pass_frame(k, NF)]
            ts = np.array([bench.slot_timestamp(rank * B + b, k) for b in range(B)], np.int64)
            R = np.stack([bench.mat3(acc[b], bench.pass_rotation(fwd, bwd, b % PS, k, NF)).reshape(9) for b in range(B)])
            for b in range(B):
                acc[b] = R[b].reshape(3, 3)
            ctx.step_dev(bl.data_ptr(), br.data_ptr(), W, ts, np.ascontiguousarray(R))
            if g is not None:
                out = g.gather()
            if k % 4 == 3:            # a keyframe every 4th frame resets the accumulated rotation (approximation for timing only)
                acc = [np.eye(3) for _ in range(B)]
        ev1.record(stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / (N_STEPS - 12)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = float(t[0])
        if g is not None:
            if rank == 0:
                hdr = [np.frombuffer(o[:64].cpu().numpy().tobytes(), np.int32)[:4].tolist() for o in out]
                res["rank0_received_headers"] = hdr
            res["bytes_per_rank_per_step"] = int(ctx.B * ctx.packet_bytes)
            g.close()
    if rank == 0:
        res.update(n_gpus=world, batch_per_gpu=B, what="ms per batch-32 step, max over ranks; gather = dist.gather (NCCL) of the fixed-size packets "
                   "to rank 0 on the context's stream, packets assembled in the send buffer by finalize_kernel")
        res["gather_share"] = (res["gather"] - res["plain"]) / res["gather"]
        print(json.dumps(res))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
