import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cv2, math
import helpers as H
from kimera_vio_b200.params import FrontendParams, CameraParams
from oracle import frontend as ofe
g,lefts,rights=H.golden()
img=lefts[0]
p=FrontendParams.euroc()
det=ofe.FeatureDetector(p)
fr=ofe.Frame(0,0,img,CameraParams.euroc_left())
kps=det.raw_feature_detection(img,det.build_mask(fr))
kps=det.suppress_non_max(kps,300,752,480)
c0=np.array([k.pt for k in kps],np.float32)
ref=c0.reshape(-1,1,2).copy()
cv2.cornerSubPix(img,ref,(10,10),(-1,-1),(cv2.TERM_CRITERIA_EPS+cv2.TERM_CRITERIA_COUNT,40,0.001))
ref=ref.reshape(-1,2)
win=10; ww=21
mask=np.zeros((ww,ww),np.float32)
for i in range(ww):
    y=np.float32(i-win)/np.float32(win); vy=np.float32(math.exp(-float(y*y))) 
    for j in range(ww):
        x=np.float32(j-win)/np.float32(win)
        mask[i,j]=np.float32(np.exp(np.float32(-y*y))*np.exp(np.float32(-x*x)))
def subpix(pt):
    cT=np.array(pt,np.float32); cI=cT.copy(); it=0; eps2=0.001**2
    while True:
        buf=cv2.getRectSubPix(img,(23,23),(float(cI[0]),float(cI[1])),patchType=cv2.CV_32F)
        sp=buf
        tgx=(sp[1:-1,2:]-sp[1:-1,:-2]).astype(np.float64); tgy=(sp[2:,1:-1]-sp[:-2,1:-1]).astype(np.float64)
        m=mask.astype(np.float64)
        gxx=tgx*tgx*m; gxy=tgx*tgy*m; gyy=tgy*tgy*m
        px,py=np.meshgrid(np.arange(ww)-win,np.arange(ww)-win)
        a=gxx.sum(); b=gxy.sum(); c=gyy.sum(); bb1=(gxx*px+gxy*py).sum(); bb2=(gxy*px+gyy*py).sum()
        det_=a*c-b*b
        if abs(det_)<=2.2e-16**2: break
        sc=1.0/det_
        n=np.array([np.float32(cI[0]+c*sc*bb1-b*sc*bb2), np.float32(cI[1]-b*sc*bb1+a*sc*bb2)],np.float32)
        err=float((n[0]-cI[0])*(n[0]-cI[0])+(n[1]-cI[1])*(n[1]-cI[1]))
        cI=n
        if cI[0]<0 or cI[0]>=752 or cI[1]<0 or cI[1]>=480: break
        it+=1
        if not (it<40 and err>eps2): break
    if abs(cI[0]-cT[0])>win or abs(cI[1]-cT[1])>win: cI=cT
    return cI,it
bad=0
for i,pt in enumerate(c0):
    r,it=subpix(pt)
    d=np.abs(r-ref[i]).max()
    near = pt[0]<12 or pt[0]>752-13 or pt[1]<12 or pt[1]>480-13
    if d>1e-4 or near:
        print(i,pt,ref[i],r,'err',d,'iters',it,'near_border',near)
print('---- kernel patch emulation vs cv2.getRectSubPix along the iterates')
f32=np.float32
I=img.astype(np.float32); Hh,Ww=img.shape
def kernel_patch(cxf,cyf,pw=23,ph=23):
    cx=f32(f32(cxf)-f32((pw-1)*0.5)); cy=f32(f32(cyf)-f32((ph-1)*0.5))
    ipx=int(np.floor(cx)); ipy=int(np.floor(cy))
    out=np.zeros((ph,pw),f32)
    if 0<=ipx and ipx+pw<Ww and 0<=ipy and ipy+ph<Hh:
        a=f32(cx-f32(ipx)); b=f32(cy-f32(ipy)); a=max(a,f32(0.0001))
        a12=f32(a*f32(f32(1)-b)); a22=f32(a*b); b1=f32(f32(1)-b); b2=b
        s=(1.0-float(a))/float(a)
        for r in range(ph):
            p0=I[ipy+r,ipx:ipx+pw+1]; p1=I[ipy+r+1,ipx:ipx+pw+1]
            for j in range(pw):
                t=f32(f32(a12*p0[j+1])+f32(a22*p1[j+1]))
                if j==0: prev=f32(f32(f32(1)-a)*f32(f32(b1*p0[0])+f32(b2*p1[0])))
                else:
                    tp=f32(f32(a12*p0[j])+f32(a22*p1[j])); prev=f32(float(tp)*s)
                out[r,j]=f32(prev+t)
        return out,'int'
    return None,'border'
worst=[]
for i,pt in enumerate(c0):
    cT=np.array(pt,np.float32); cI=cT.copy(); it=0
    while True:
        ref_p=cv2.getRectSubPix(img,(23,23),(float(cI[0]),float(cI[1])),patchType=cv2.CV_32F)
        kp,kind=kernel_patch(cI[0],cI[1])
        if kp is not None:
            d=np.abs(kp-ref_p).max()
            if d>0: worst.append((d,i,it,tuple(cI),kind))
        sp=ref_p
        tgx=(sp[1:-1,2:]-sp[1:-1,:-2]).astype(np.float64); tgy=(sp[2:,1:-1]-sp[:-2,1:-1]).astype(np.float64)
        m=mask.astype(np.float64)
        gxx=tgx*tgx*m; gxy=tgx*tgy*m; gyy=tgy*tgy*m
        px,py=np.meshgrid(np.arange(ww)-win,np.arange(ww)-win)
        a=gxx.sum(); b=gxy.sum(); c=gyy.sum(); bb1=(gxx*px+gxy*py).sum(); bb2=(gxy*px+gyy*py).sum()
        det_=a*c-b*b
        if abs(det_)<=2.2e-16**2: break
        sc=1.0/det_
        n=np.array([np.float32(cI[0]+c*sc*bb1-b*sc*bb2), np.float32(cI[1]-b*sc*bb1+a*sc*bb2)],np.float32)
        err=float((n[0]-cI[0])*(n[0]-cI[0])+(n[1]-cI[1])*(n[1]-cI[1]))
        cI=n
        if cI[0]<0 or cI[0]>=752 or cI[1]<0 or cI[1]>=480: break
        it+=1
        if not (it<40 and err>eps2_ if False else (it<40 and err>1e-6)): break
print(len(worst), sorted(worst,reverse=True)[:8])
