import sys, dataclasses; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cv2
import helpers as H
from kimera_vio_b200.params import FrontendParams, CameraParams
from oracle import frontend as ofe
p, rig, ctx = H.euroc_setup(batch=1)
s, fr = H.synth_frames(2, seed=31337)
img = fr[0].left
det = ofe.FeatureDetector(p)
f = ofe.Frame(0,0,img,CameraParams.euroc_left())
p2 = dataclasses.replace(p, enable_subpixel_corner_refinement=False)
p2, rig2, ctx2 = H.euroc_setup(batch=1, params=p2)
det2 = ofe.FeatureDetector(p2)
e2 = det2.detect_corners(f, 300); g2 = ctx2.detect(img, [], [], 300)
print('binned n', len(e2), len(g2), 'equal', np.array_equal(e2,g2))
e3 = det.detect_corners(f,300); g3=ctx.detect(img,[],[],300)
d=np.abs(e3-g3).max(axis=1)
for i in np.argsort(-d)[:5]:
    print('corner', i, 'err', d[i], 'int', e2[i], 'ref', e3[i], 'gpu', g3[i])
i=int(np.argmax(d))
# replay cornerSubPix iteration by iteration with cv2 (max_iter = 1..N) for the worst corner
c0=np.array(e2[i],np.float32).reshape(1,1,2)
for it in range(1, p.subpix_max_iters+1):
    c=c0.copy()
    cv2.cornerSubPix(img, c, (p.subpix_win_size,)*2, (p.subpix_zero_zone,)*2, (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_COUNT, it, p.subpix_epsilon))
    print(' cv2 iters<=%d ->'%it, c.ravel())
print('params', p.subpix_win_size, p.subpix_zero_zone, p.subpix_max_iters, p.subpix_epsilon)
