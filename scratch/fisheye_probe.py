import numpy as np, cv2, sys
sys.path.insert(0, '.')
from kimera_vio_b200.params import CameraParams
import yaml, math
def load(path):
    txt = open(path).read().replace('%YAML:1.0', '')
    return yaml.safe_load(txt)
L = load('/root/reference/params/RealSenseIR/LeftCameraParams.yaml'); R = load('/root/reference/params/RealSenseIR/RightCameraParams.yaml')
def cam(y):
    i = y['intrinsics']; K = np.array([[i[0],0,i[2]],[0,i[1],i[3]],[0,0,1.0]]); D = np.array(y['distortion_coefficients'], np.float64)
    T = np.array(y['T_BS']['data'], np.float64).reshape(4,4); return K, D, T
KL, DL, TL = cam(L); KR, DR, TR = cam(R)
W, H = L['resolution']
c = np.linalg.inv(np.linalg.inv(TL) @ TR)
Rm, T = c[:3,:3].copy(), c[:3,3].copy()
R1,R2,P1,P2,Q = cv2.fisheye.stereoRectify(KL, DL, KR, DR, (W,H), Rm, T, flags=cv2.CALIB_ZERO_DISPARITY)
print(P1, P2, 1.0/Q[3,2])
mx, my = cv2.fisheye.initUndistortRectifyMap(KL, DL, R1, P1, (W,H), cv2.CV_32FC1)
def inv3(S):
    S = S.reshape(9)
    d = S[0]*(S[4]*S[8]-S[5]*S[7]) - S[1]*(S[3]*S[8]-S[5]*S[6]) + S[2]*(S[3]*S[7]-S[4]*S[6])
    d = 1.0/d
    t = np.empty(9)
    t[0]=(S[4]*S[8]-S[5]*S[7])*d; t[1]=(S[2]*S[7]-S[1]*S[8])*d; t[2]=(S[1]*S[5]-S[2]*S[4])*d
    t[3]=(S[5]*S[6]-S[3]*S[8])*d; t[4]=(S[0]*S[8]-S[2]*S[6])*d; t[5]=(S[2]*S[3]-S[0]*S[5])*d
    t[6]=(S[3]*S[7]-S[4]*S[6])*d; t[7]=(S[1]*S[6]-S[0]*S[7])*d; t[8]=(S[0]*S[4]-S[1]*S[3])*d
    return t
def mymap(K, D, Rr, P, variant):
    PP = P[:3,:3]
    RP = np.empty((3,3))
    for i in range(3):
        for j in range(3):
            if variant & 1: RP[i,j] = (PP[i,0]*Rr[0,j] + PP[i,1]*Rr[1,j]) + PP[i,2]*Rr[2,j]
            else: RP[i,j] = PP[i,0]*Rr[0,j] + (PP[i,1]*Rr[1,j] + PP[i,2]*Rr[2,j])
    iR = inv3(RP) if not (variant & 2) else np.linalg.inv(RP).reshape(9)
    f0,f1,c0,c1 = K[0,0],K[1,1],K[0,2],K[1,2]
    ox = np.empty((H,W), np.float32); oy = np.empty((H,W), np.float32)
    ii = np.arange(H, dtype=np.float64)
    _x = ii*iR[1] + iR[2]; _y = ii*iR[4] + iR[5]; _w = ii*iR[7] + iR[8]
    for j in range(W):
        x = _x/_w; y = _y/_w
        r = np.sqrt(x*x + y*y)
        th = np.arctan(r)
        t2 = th*th; t4 = t2*t2; t6 = t4*t2; t8 = t4*t4
        thd = th*(1 + D[0]*t2 + D[1]*t4 + D[2]*t6 + D[3]*t8)
        sc = np.where(r == 0, 1.0, thd/np.where(r==0,1.0,r))
        u = f0*x*sc + c0; v = f1*y*sc + c1
        ox[:,j] = u.astype(np.float32); oy[:,j] = v.astype(np.float32)
        _x = _x + iR[0]; _y = _y + iR[3]; _w = _w + iR[6]
    return ox, oy
for variant in range(4):
    ox, oy = mymap(KL, DL, R1, P1, variant)
    print(variant, (ox.view(np.int32) != mx.view(np.int32)).sum(), (oy.view(np.int32) != my.view(np.int32)).sum(), np.abs(ox-mx).max())

def my_undist(pts, K, D, Rr, P):
    f0,f1,c0,c1 = K[0,0],K[1,1],K[0,2],K[1,2]
    RR = np.eye(3) if Rr is None else Rr.copy()
    if P is not None:
        PP = P[:3,:3]; RP = np.empty((3,3))
        for i in range(3):
            for j in range(3):
                RP[i,j] = (PP[i,0]*RR[0,j] + PP[i,1]*RR[1,j]) + PP[i,2]*RR[2,j]
        RR = RP
    out = np.empty_like(pts)
    for i,(u,v) in enumerate(pts.astype(np.float64)):
        pw0 = (u - c0)/f0; pw1 = (v - c1)/f1
        thd = math.sqrt(pw0*pw0 + pw1*pw1)
        thd = min(max(-math.pi/2., thd), math.pi/2.)
        conv = False; th = thd; scale = 0.0
        if abs(thd) > 1e-8:
            for j in range(10):
                t2 = th*th; t4 = t2*t2; t6 = t4*t2; t8 = t6*t2
                a = D[0]*t2; b = D[1]*t4; c_ = D[2]*t6; d = D[3]*t8
                fix = (th*(1 + a + b + c_ + d) - thd) / (1 + 3*a + 5*b + 7*c_ + 9*d)
                th = th - fix
                if abs(fix) < 1e-8:
                    conv = True; break
            scale = math.tan(th)/thd
        else:
            conv = True
        flipped = (thd < 0 and th > 0) or (thd > 0 and th < 0)
        if conv and not flipped:
            pu0 = pw0*scale; pu1 = pw1*scale
            pr = [(RR[r,0]*pu0 + RR[r,1]*pu1) + RR[r,2]*1.0 for r in range(3)]
            out[i] = (np.float32(pr[0]/pr[2]), np.float32(pr[1]/pr[2]))
        else:
            out[i] = (-1000000.0, -1000000.0)
    return out
rng = np.random.default_rng(0)
pts = np.stack([rng.uniform(-50, W+50, 4000), rng.uniform(-50, H+50, 4000)], 1).astype(np.float32)
pts[0] = (KL[0,2], KL[1,2]); pts[1] = (np.float32(KL[0,2]), np.float32(KL[1,2]))
for (Rr, P, name) in ((R1, P1, "R,P"), (R1, None, "R"), (None, None, "-"), (None, P1, "P")):
    ref = cv2.fisheye.undistortPoints(pts.reshape(-1,1,2), KL, DL, R=Rr, P=P).reshape(-1,2)
    mine = my_undist(pts, KL, DL, Rr, P)
    neq = (ref.view(np.int32) != mine.view(np.int32)).any(1)
    print(name, neq.sum(), np.abs(ref-mine).max(), (ref[:,0] == -1000000.0).sum())
    if neq.sum(): print(pts[neq][:5], ref[neq][:5], mine[neq][:5])
pts = np.stack([rng.uniform(-3000, 3000, 4000), rng.uniform(-3000, 3000, 4000)], 1).astype(np.float32)
for Dt in (DL, np.array([-0.3, 0.2, -0.5, 0.1]), np.array([0.9, -2.0, 3.0, -1.0])):
    ref = cv2.fisheye.undistortPoints(pts.reshape(-1,1,2), KL, Dt, R=R1, P=P1).reshape(-1,2)
    mine = my_undist(pts, KL, Dt, R1, P1)
    neq = (ref.view(np.int32) != mine.view(np.int32)).any(1)
    print("far", neq.sum(), (ref[:,0] == -1000000.0).sum(), np.isnan(ref).sum())
    if neq.sum(): print(pts[neq][:5], ref[neq][:5], mine[neq][:5])
