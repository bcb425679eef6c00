import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cv2
import helpers as H
f32=np.float32
s, fr = H.synth_frames(1, seed=31337)
img = fr[0].left; I=img.astype(f32); Hh,Ww=img.shape
def clampi(v,lo,hi): return max(lo,min(hi,v))
def pair_clamped(cxf,cyf,pw=23,ph=23):
    cx=f32(f32(cxf)-f32((pw-1)*0.5)); cy=f32(f32(cyf)-f32((ph-1)*0.5))
    ipx=int(np.floor(cx)); ipy=int(np.floor(cy))
    a=f32(cx-f32(ipx)); b=f32(cy-f32(ipy))
    a11=f32((f32(1)-a)*(f32(1)-b)); a12=f32(a*(f32(1)-b)); a21=f32((f32(1)-a)*b); a22=f32(a*b)
    out=np.zeros((ph,pw),f32)
    for i in range(ph):
        y0=clampi(ipy+i,0,Hh-1); y1=clampi(ipy+i+1,0,Hh-1)
        for j in range(pw):
            x0=clampi(ipx+j,0,Ww-1); x1=clampi(ipx+j+1,0,Ww-1)
            out[i,j]=f32(f32(I[y0,x0]*a11)+f32(I[y0,x1]*a12))+f32(f32(I[y1,x0]*a21)+f32(I[y1,x1]*a22))
    return out
def ocv_border(cxf,cyf,pw=23,ph=23):
    # restatement of getRectSubPix_Cn_ border branch (imgproc/src/samplers.cpp)
    cx=f32(f32(cxf)-f32((pw-1)*0.5)); cy=f32(f32(cyf)-f32((ph-1)*0.5))
    ipx=int(np.floor(cx)); ipy=int(np.floor(cy))
    a=f32(cx-f32(ipx)); b=f32(cy-f32(ipy))
    a11=f32((f32(1)-a)*(f32(1)-b)); a12=f32(a*(f32(1)-b)); a21=f32((f32(1)-a)*b); a22=f32(a*b)
    b1=f32(f32(1)-b); b2=b
    # adjustRect
    rx=0; ry=0; rw=pw; rh=ph
    sx=ipx; sy=ipy   # src origin (may be moved)
    if ipx>=0: pass
    else: rx=-ipx; rx=min(rx,pw); sx=ipx  # columns j<rx use src[r.x]
    if ipx+pw<Ww: rw=pw
    else:
        rw=Ww-ipx-1
        if rw<0: rw=0
    if ipy>=0: pass
    else: ry=-ipy
    if ipy+ph<Hh: rh=ph
    else: rh=Hh-ipy-1
    out=np.zeros((ph,pw),f32)
    def px(y,x): return I[clampi(y,0,Hh-1), clampi(x,0,Ww-1)]
    for i in range(ph):
        # row pointers: src row = clamp(ipy+i), src2 = next row unless outside
        y=ipy+i
        yy0=clampi(y,0,Hh-1); yy1=clampi(y+1,0,Hh-1)
        for j in range(pw):
            x=ipx+j
            if j<rx:
                out[i,j]=f32(f32(px(yy0,ipx+rx)*b1)+f32(px(yy1,ipx+rx)*b2))
            elif j<rw:
                out[i,j]=f32(f32(f32(f32(I[yy0,x]*a11)+f32(I[yy0,x+1]*a12))+f32(I[yy1,x]*a21))+f32(I[yy1,x+1]*a22))
            else:
                out[i,j]=f32(f32(px(yy0,ipx+rw)*b1)+f32(px(yy1,ipx+rw)*b2))
    return out
tests=[(743.2506,5.202382),(743.0,1.0),(5.3,200.7),(748.6,300.2),(300.4,3.3),(300.4,476.8),(3.2,3.7),(749.1,477.2),(743.25,240.5)]
for (cx,cy) in tests:
    ref=cv2.getRectSubPix(img,(23,23),(cx,cy),patchType=cv2.CV_32F)
    A=pair_clamped(cx,cy); B=ocv_border(cx,cy)
    print((cx,cy),'pair_clamped mism',int((A!=ref).sum()),'ocv_border mism',int((B!=ref).sum()), 'maxdiff A %.3g'%np.abs(A-ref).max())
    if (A!=ref).any():
        ii,jj=np.nonzero(A!=ref); print('   mism cols', sorted(set(jj.tolist()))[:30], 'rows', sorted(set(ii.tolist()))[:30])
