"""Small driver for ncu: N device-resident batch steps (one context, batch 32, BASELINE configs[1] by default).
  KVFE_NO_GRAPH=1 KVFE_STEPS=10 ncu ... python profiles/profile_step.py [config]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from kimera_vio_b200 import lib as kl

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = bench.CONFIGS[name]
W, H, PS, NF = cfg["W"], cfg["H"], cfg["pool_streams"], cfg["pool_frames"]
B = int(os.environ.get("KVFE_BATCH", str(cfg["batch"])))
N = int(os.environ.get("KVFE_STEPS", "10"))
left, right, fwd, bwd = bench.frame_pool(name)
lcam, rcam, rig = bench.config_rig(cfg)
p = bench.config_params(cfg)
ctx = kl.Context(kl.make_config(p, W, H, batch=B, sobel_cpu_tail_start=bench.sobel_cpu_tail_start(W)), rig.to_c())
dL, dR = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
idx = torch.tensor([b % PS for b in range(B)], device="cuda")
acc = [np.eye(3) for _ in range(B)]
for k in range(N):
    f = bench.pass_frame(k, NF)
    bl, br = dL[idx, f].contiguous(), dR[idx, f].contiguous()
    ts = np.array([bench.slot_timestamp(b, k) for b in range(B)], np.int64)
    R = np.zeros((B, 9))
    for b in range(B):
        acc[b] = bench.mat3(acc[b], bench.pass_rotation(fwd, bwd, b % PS, k, NF))
        R[b] = acc[b].reshape(9)
    ctx.step_dev(bl.data_ptr(), br.data_ptr(), W, ts, np.ascontiguousarray(R))
    pk = ctx.read_packets()
    for b in range(B):
        if pk[b]["is_keyframe"]:
            acc[b] = np.eye(3)
print("keyframes in the last step:", sum(int(q["is_keyframe"]) for q in pk), "of", B)
