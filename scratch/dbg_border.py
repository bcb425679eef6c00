import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cv2, math
import helpers as H
f32=np.float32
g,lefts,_=H.golden(); img=lefts[2]; Hh,Ww=img.shape; I=img.astype(f32)
def kern_patch(cxf,cyf,pw=23,ph=23):
    cx=f32(f32(cxf)-f32((pw-1)*0.5)); cy=f32(f32(cyf)-f32((ph-1)*0.5))
    ipx=int(np.floor(cx)); ipy=int(np.floor(cy))
    a=f32(cx-f32(ipx)); b=f32(cy-f32(ipy))
    a11=f32((f32(1)-a)*(f32(1)-b)); a12=f32(a*(f32(1)-b)); a21=f32((f32(1)-a)*b); a22=f32(a*b); b1=f32(f32(1)-b); b2=b
    rx = 0 if ipx>=0 else min(-ipx,pw)
    rw = pw if ipx < Ww-pw else max(Ww-ipx-1,0)
    ry = 0 if ipy>=0 else -ipy
    rh = ph if ipy < Hh-ph else max(Hh-ipy-1,0)
    out=np.zeros((ph,pw),f32)
    for r in range(ph):
        y0=min(max(ipy+r,0),Hh-1); y1=y0 if (r<ry or r>=rh) else min(max(ipy+r+1,0),Hh-1)
        for j in range(pw):
            if j<rx: xc=min(max(ipx+rx,0),Ww-1); v=f32(f32(I[y0,xc]*b1)+f32(I[y1,xc]*b2))
            elif j>=rw: xc=min(max(ipx+rw,0),Ww-1); v=f32(f32(I[y0,xc]*b1)+f32(I[y1,xc]*b2))
            else:
                x0=min(max(ipx+j,0),Ww-1); x1=min(max(ipx+j+1,0),Ww-1)
                v=f32(f32(f32(I[y0,x0]*a11)+f32(I[y0,x1]*a12))+f32(f32(I[y1,x0]*a21)+f32(I[y1,x1]*a22)))
            out[r,j]=v
    return out
win=10; ww=21
mask=np.zeros((ww,ww),f32)
for i in range(ww):
    y=f32(i-win)/f32(win)
    for j in range(ww):
        x=f32(j-win)/f32(win); mask[i,j]=f32(np.exp(f32(-y*y))*np.exp(f32(-x*x)))
cI=np.array([749.,408.],f32); it=0
while True:
    ref=cv2.getRectSubPix(img,(23,23),(float(cI[0]),float(cI[1])),patchType=cv2.CV_32F)
    kp=kern_patch(cI[0],cI[1])
    d=np.abs(ref-kp); bad=np.argwhere(ref!=kp)
    print('iter',it,'center',cI,'patch mism',len(bad),'max',d.max(), bad[:4].tolist())
    sp=ref
    tgx=(sp[1:-1,2:]-sp[1:-1,:-2]).astype(np.float64); tgy=(sp[2:,1:-1]-sp[:-2,1:-1]).astype(np.float64)
    m=mask.astype(np.float64); gxx=tgx*tgx*m; gxy=tgx*tgy*m; gyy=tgy*tgy*m
    px,py=np.meshgrid(np.arange(ww)-win,np.arange(ww)-win)
    a=gxx.sum(); b=gxy.sum(); c=gyy.sum(); bb1=(gxx*px+gxy*py).sum(); bb2=(gxy*px+gyy*py).sum()
    det=a*c-b*b; sc=1.0/det
    n=np.array([f32(cI[0]+c*sc*bb1-b*sc*bb2), f32(cI[1]-b*sc*bb1+a*sc*bb2)],f32)
    err=float((n[0]-cI[0])**2+(n[1]-cI[1])**2); cI=n
    if cI[0]<0 or cI[0]>=752 or cI[1]<0 or cI[1]>=480: print('out of image',cI); break
    it+=1
    if not (it<40 and err>1e-6): break
print('final',cI)
c=np.array([[[749.,408.]]],f32); cv2.cornerSubPix(img,c,(10,10),(-1,-1),(cv2.TERM_CRITERIA_EPS+cv2.TERM_CRITERIA_COUNT,40,0.001)); print('cv2',c)
