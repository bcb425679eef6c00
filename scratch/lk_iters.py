import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cv2
import helpers as H
s, fr = H.synth_frames(3, seed=20240)
I0, I1 = fr[1].left, fr[2].left
pts = cv2.goodFeaturesToTrack(I0, 300, 0.001, 20).reshape(-1,2)
W=24; L=4; eps=0.1
def pyr(I):
    out=[I]
    for _ in range(L): out.append(cv2.pyrDown(out[-1]))
    return out
P0,P1=pyr(I0),pyr(I1)
iters=[[] for _ in range(L+1)]
for p in pts:
    g=np.zeros(2)
    for l in range(L,-1,-1):
        A=P0[l].astype(np.float32); B=P1[l].astype(np.float32)
        c=p/(1<<l)
        Ip=cv2.getRectSubPix(A,(W,W),(float(c[0]),float(c[1])))
        gx=cv2.Scharr(A,cv2.CV_32F,1,0)/32; gy=cv2.Scharr(A,cv2.CV_32F,0,1)/32
        Ix=cv2.getRectSubPix(gx,(W,W),(float(c[0]),float(c[1]))); Iy=cv2.getRectSubPix(gy,(W,W),(float(c[0]),float(c[1])))
        G=np.array([[ (Ix*Ix).sum(),(Ix*Iy).sum()],[(Ix*Iy).sum(),(Iy*Iy).sum()]])
        if np.linalg.det(G)<1e-6: iters[l].append(0); g=g*2; continue
        Gi=np.linalg.inv(G); q=c+g; prev=None; n=0
        for j in range(30):
            n+=1
            Jp=cv2.getRectSubPix(B,(W,W),(float(q[0]),float(q[1])))
            d=Jp-Ip; b=np.array([(d*Ix).sum(),(d*Iy).sum()]); dl=-Gi@b
            q=q+dl
            if dl@dl<=eps*eps: break
            if prev is not None and abs(dl[0]+prev[0])<0.01 and abs(dl[1]+prev[1])<0.01: break
            prev=dl
        iters[l].append(n); g=(q-c)*(2 if l>0 else 1)
for l in range(L,-1,-1):
    a=np.array(iters[l]); m4=a[:len(a)//4*4].reshape(-1,4).max(1)
    print("level",l,"mean iters %.2f"%a.mean(),"max",a.max(),"mean of max-of-4 %.2f"%m4.mean(), np.bincount(a)[:12])
