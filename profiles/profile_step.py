"""Small driver for ncu: N steps of the device-resident batch step (batch 32, Euroc 752x480)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from kimera_vio_b200 import lib as kl
from kimera_vio_b200.params import CameraParams, FrontendParams
from kimera_vio_b200.rig import StereoRigSetup

B = int(os.environ.get("KVFE_BATCH", "32"))
N = int(os.environ.get("KVFE_STEPS", "10"))
rig = StereoRigSetup(CameraParams.euroc_left(), CameraParams.euroc_right())
left, right, rot = bench.frame_pool(20, rig)
ctx = kl.Context(kl.make_config(FrontendParams.euroc(), bench.W, bench.H, batch=B), rig.to_c())
dL = torch.empty((N, B, bench.H, bench.W), dtype=torch.uint8, device="cuda")
dR = torch.empty_like(dL)
for k in range(N):
    for b in range(B):
        dL[k, b].copy_(torch.from_numpy(left[b % bench.POOL_STREAMS, k]))
        dR[k, b].copy_(torch.from_numpy(right[b % bench.POOL_STREAMS, k]))
torch.cuda.synchronize()
lkf = np.zeros(B, np.int64)
for k in range(N):
    ts = np.array([bench.slot_timestamp(b, k) for b in range(B)], np.int64)
    R = np.stack([rot[b % bench.POOL_STREAMS, lkf[b], k].reshape(9) for b in range(B)])
    ctx.step_dev(dL[k].data_ptr(), dR[k].data_ptr(), bench.W, ts, R)
    pk = ctx.read_packets()
    for b in range(B):
        if pk[b]["is_keyframe"]:
            lkf[b] = k
print("modes of last step:", [p["mode"] for p in pk])
