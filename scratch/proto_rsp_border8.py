import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cv2
import helpers as H
f32=np.float32
s, fr = H.synth_frames(1, seed=31337)
img = fr[0].left; I=img.astype(f32); Hh,Ww=img.shape
def model(cxf,cyf,pw=23,ph=23,quirk_form='pb'):
    cx=f32(f32(cxf)-f32((pw-1)*0.5)); cy=f32(f32(cyf)-f32((ph-1)*0.5))
    ipx=int(np.floor(cx)); ipy=int(np.floor(cy))
    a=f32(cx-f32(ipx)); b=f32(cy-f32(ipy))
    a11=f32((f32(1)-a)*(f32(1)-b)); a12=f32(a*(f32(1)-b)); a21=f32((f32(1)-a)*b); a22=f32(a*b)
    a1=f32(f32(1)-a); b1=f32(f32(1)-b)
    rx=min(max(-ipx,0),pw); rw=pw if ipx+pw<Ww else max(Ww-ipx-1,0)
    ry=max(-ipy,0); rh=ph if ipy+ph<Hh else max(Hh-ipy-1,0)
    out=np.zeros((ph,pw),f32)
    for i in range(ph):
        outside = i<ry or i>=rh
        y0 = 0 if i<ry else (Hh-1 if i>=rh else ipy+i)
        y1 = y0 if outside else y0+1
        for j in range(pw):
            if j<rx or j>=rw:
                xc = 0 if j<rx else Ww-1
                if i<ry and j>=rw: xc = Ww-2                       # observed quirk
                p0,p1=I[y0,xc],I[y1,xc]
                out[i,j]=f32(f32(p0*b1)+f32(p1*b))
            else:
                x=ipx+j
                if outside: out[i,j]=f32(np.float64(I[y0,x+1])*np.float64(a)+np.float64(f32(I[y0,x]*a1)))
                else: out[i,j]=f32(f32(I[y0,x]*a11)+f32(I[y0,x+1]*a12))+f32(f32(I[y1,x]*a21)+f32(I[y1,x+1]*a22))
    return out
rng=np.random.default_rng(3)
tests=[(743.2506,5.202382),(743.0,1.0),(300.4,3.3),(300.4,476.8),(3.2,3.7),(749.1,477.2),(5.3,200.7),(748.6,300.2)]
for k in range(200):
    side=k%8
    cx=rng.uniform(0,12) if side in (0,4,5) else (rng.uniform(740,751) if side in (1,6,7) else rng.uniform(20,730))
    cy=rng.uniform(0,12) if side in (2,4,6) else (rng.uniform(468,479) if side in (3,5,7) else rng.uniform(20,460))
    tests.append((float(f32(cx)),float(f32(cy))))
tot=0
for (cx,cy) in tests:
    ref=cv2.getRectSubPix(img,(23,23),(cx,cy),patchType=cv2.CV_32F)
    A=model(cx,cy); m=int((A!=ref).sum()); tot+=m
    if m:
        ii,jj=np.nonzero(A!=ref); print((cx,cy),m,'rows',sorted(set(ii.tolist())),'cols',sorted(set(jj.tolist())),'maxdiff',np.abs(A-ref).max())
print('total mismatches over',len(tests),'patches:',tot)

# alternative for outside columns: fma(p1, b, p0*b1)
def model2(cx,cy):
    A=model(cx,cy)
    return A
