import sys, dataclasses; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cv2
import helpers as H
from kimera_vio_b200.params import FrontendParams, CameraParams
from oracle import frontend as ofe
p = dataclasses.replace(FrontendParams.euroc(), min_distance=8)
p, rig, ctx = H.euroc_setup(batch=1, params=p)
g, lefts, _ = H.golden()
img = lefts[2]
det = ofe.FeatureDetector(p)
fr = ofe.Frame(0,0,img,CameraParams.euroc_left())
raw = det.raw_feature_detection(img, det.build_mask(fr))
e = np.array([k.pt for k in raw], np.float32)
gq, resp = ctx.detect_raw(img)
print('raw n', len(e), len(gq), 'equal', np.array_equal(e, gq))
if len(e)==len(gq):
    bad = np.nonzero((e!=gq).any(axis=1))[0]
    print('first diffs', bad[:10], e[bad[:5]], gq[bad[:5]])
else:
    se=set(map(tuple,e)); sg=set(map(tuple,gq)); print('only ref', list(se-sg)[:5], 'only gpu', list(sg-se)[:5])
p2 = dataclasses.replace(p, enable_subpixel_corner_refinement=False)
p2, rig2, ctx2 = H.euroc_setup(batch=1, params=p2)
det2 = ofe.FeatureDetector(p2)
e2 = det2.detect_corners(fr, 300); g2 = ctx2.detect(img, [], [], 300)
print('binned n', len(e2), len(g2), 'equal', np.array_equal(e2,g2))
e3 = det.detect_corners(fr,300); g3=ctx.detect(img,[],[],300)
d=np.abs(e3-g3).max(axis=1); i=np.argmax(d); print('subpix worst', i, d[i], e2[i], e3[i], g3[i])
