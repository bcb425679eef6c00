import numpy as np, cv2, glob
f32=np.float32
img=cv2.imread(sorted(glob.glob('/root/reference/tests/data/MicroEurocDataset/mav0/cam0/data/*.png'))[10],0)
H,W=img.shape
def model(img,cxf,cyf,pw,ph,variant=0):
    cx=f32(cxf)-f32((pw-1)*0.5); cy=f32(cyf)-f32((ph-1)*0.5)
    ipx=int(np.floor(cx)); ipy=int(np.floor(cy))
    a=f32(cx-f32(ipx)); b=f32(cy-f32(ipy))
    a11=f32((f32(1)-a)*(f32(1)-b)); a12=f32(a*(f32(1)-b)); a21=f32((f32(1)-a)*b); a22=f32(a*b)
    b1=f32(1)-b; b2=b
    rx = 0 if ipx>=0 else min(-ipx,pw)
    rw = pw if ipx < W-pw else max(W-ipx-1,0)
    ry = 0 if ipy>=0 else -ipy
    rh = ph if ipy < H-ph else max(H-ipy-1,0)
    out=np.zeros((ph,pw),f32)
    I=img.astype(f32)
    for r in range(ph):
        y0=min(max(ipy+r,0),H-1)
        y1=y0 if (r<ry or r>=rh) else min(max(ipy+r+1,0),H-1)
        for j in range(pw):
            if j<rx:
                xc=min(max(ipx+rx,0),W-1); v=f32(f32(I[y0,xc]*b1)+f32(I[y1,xc]*b2))
            elif j>=rw:
                xc=min(max(ipx+rw,0),W-1); v=f32(f32(I[y0,xc]*b1)+f32(I[y1,xc]*b2))
            else:
                x0=min(max(ipx+j,0),W-1); x1=min(max(ipx+j+1,0),W-1)
                v=f32(f32(f32(f32(I[y0,x0]*a11)+f32(I[y0,x1]*a12))+f32(I[y1,x0]*a21))+f32(I[y1,x1]*a22))
            out[r,j]=v
    return out
rng=np.random.default_rng(0)
for (cx,cy) in [(5.3,100.2),(745.7,200.4),(300.6,4.4),(300.3,474.6),(3.2,3.7),(748.1,476.9),(11.0,240.0),(11.2,240.0),(10.9,240)]:
    ref=cv2.getRectSubPix(img,(23,23),(cx,cy),patchType=cv2.CV_32F)
    m=model(img,cx,cy,23,23)
    d=np.abs(ref-m)
    print((cx,cy),'mism',(ref!=m).sum(),'max',d.max(), np.argwhere(ref!=m)[:4].tolist())
print('--- variants')
import itertools
def variants(I00,I01,I10,I11,a11,a12,a21,a22):
    d=np.float64
    out={}
    out['seq']=f32(f32(f32(f32(I00*a11)+f32(I01*a12))+f32(I10*a21))+f32(I11*a22))
    out['pair']=f32(f32(f32(I00*a11)+f32(I01*a12))+f32(f32(I10*a21)+f32(I11*a22)))
    out['dbl']=f32(d(I00)*d(a11)+d(I01)*d(a12)+d(I10)*d(a21)+d(I11)*d(a22))
    # fma chain: fma(I11,a22,fma(I10,a21,fma(I01,a12,I00*a11)))
    t=f32(I00*a11); t=f32(d(I01)*d(a12)+d(t)); t=f32(d(I10)*d(a21)+d(t)); t=f32(d(I11)*d(a22)+d(t)); out['fma_seq']=t
    # rows first: (I00*a11 + I10*a21) + (I01*a12 + I11*a22)
    out['cols']=f32(f32(f32(I00*a11)+f32(I10*a21))+f32(f32(I01*a12)+f32(I11*a22)))
    return out
I=img.astype(f32)
for (cxf,cyf) in [(5.3,100.2),(745.7,200.4),(300.3,474.6)]:
    pw=ph=23
    ref=cv2.getRectSubPix(img,(23,23),(cxf,cyf),patchType=cv2.CV_32F)
    cx=f32(cxf)-f32(11.0); cy=f32(cyf)-f32(11.0)
    ipx=int(np.floor(cx)); ipy=int(np.floor(cy))
    a=f32(cx-f32(ipx)); b=f32(cy-f32(ipy))
    a11=f32((f32(1)-a)*(f32(1)-b)); a12=f32(a*(f32(1)-b)); a21=f32((f32(1)-a)*b); a22=f32(a*b)
    rx = 0 if ipx>=0 else min(-ipx,pw)
    rw = pw if ipx < W-pw else max(W-ipx-1,0)
    ry = 0 if ipy>=0 else -ipy
    rh = ph if ipy < H-ph else max(H-ipy-1,0)
    cnt={}
    tot=0
    for r in range(ph):
        y0=min(max(ipy+r,0),H-1); y1=y0 if (r<ry or r>=rh) else min(max(ipy+r+1,0),H-1)
        for j in range(rx,rw):
            x0=min(max(ipx+j,0),W-1); x1=min(max(ipx+j+1,0),W-1)
            v=variants(I[y0,x0],I[y0,x1],I[y1,x0],I[y1,x1],a11,a12,a21,a22)
            tot+=1
            for k,val in v.items():
                cnt[k]=cnt.get(k,0)+(val==ref[r,j])
    print((cxf,cyf),tot,cnt)
