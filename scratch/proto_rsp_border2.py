import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cv2, itertools
import helpers as H
f32=np.float32
s, fr = H.synth_frames(1, seed=31337)
img = fr[0].left; I=img.astype(f32); Hh,Ww=img.shape
def clampi(v,lo,hi): return max(lo,min(hi,v))
def gen(cxf,cyf,rowmap,form,pw=23,ph=23):
    cx=f32(f32(cxf)-f32((pw-1)*0.5)); cy=f32(f32(cyf)-f32((ph-1)*0.5))
    ipx=int(np.floor(cx)); ipy=int(np.floor(cy))
    a=f32(cx-f32(ipx)); b=f32(cy-f32(ipy))
    a11=f32((f32(1)-a)*(f32(1)-b)); a12=f32(a*(f32(1)-b)); a21=f32((f32(1)-a)*b); a22=f32(a*b)
    out=np.zeros((ph,pw),f32)
    for i in range(ph):
        y0,y1=rowmap(ipy+i)
        for j in range(pw):
            x0=clampi(ipx+j,0,Ww-1); x1=clampi(ipx+j+1,0,Ww-1)
            p00,p01,p10,p11=I[y0,x0],I[y0,x1],I[y1,x0],I[y1,x1]
            if form=='pair': v=f32(f32(p00*a11)+f32(p01*a12))+f32(f32(p10*a21)+f32(p11*a22))
            elif form=='seq': v=f32(f32(f32(f32(p00*a11)+f32(p01*a12))+f32(p10*a21))+f32(p11*a22))
            elif form=='vert': # vertical first: (p00*b1+p10*b2)*(1-a) + (p01*b1+p11*b2)*a
                b1=f32(f32(1)-b); c0=f32(f32(p00*b1)+f32(p10*b)); c1=f32(f32(p01*b1)+f32(p11*b)); v=f32(f32(c0*f32(f32(1)-a))+f32(c1*a))
            elif form=='pairT': v=f32(f32(p00*a11)+f32(p10*a21))+f32(f32(p01*a12)+f32(p11*a22))
            out[i,j]=v
    return out
rm_clamp=lambda y:(clampi(y,0,Hh-1),clampi(y+1,0,Hh-1))
tests=[(300.4,3.3),(300.4,476.8),(300.0,2.0),(300.7,5.5)]
for (cx,cy) in tests:
    ref=cv2.getRectSubPix(img,(23,23),(cx,cy),patchType=cv2.CV_32F)
    res={}
    for form in ['pair','seq','vert','pairT']:
        A=gen(cx,cy,rm_clamp,form); res[form]=int((A!=ref).sum())
    print((cx,cy),res)
# inspect one mismatching element in detail
cx,cy=300.4,3.3
ref=cv2.getRectSubPix(img,(23,23),(cx,cy),patchType=cv2.CV_32F)
A=gen(cx,cy,rm_clamp,'pair')
ii,jj=np.nonzero(A!=ref)
i,j=ii[0],jj[0]
cxx=f32(f32(cx)-f32(11)); ipx=int(np.floor(cxx)); a=f32(cxx-f32(ipx)); cyy=f32(f32(cy)-f32(11)); ipy=int(np.floor(cyy)); b=f32(cyy-f32(ipy))
print('elem',i,j,'ref %.9g mine %.9g'%(ref[i,j],A[i,j]),'a',a,'b',b,'pix',I[0,ipx+j],I[0,ipx+j+1])
p0,p1=I[0,ipx+j],I[0,ipx+j+1]
cands={'(1-a)p0+a p1':f32(f32(p0*f32(f32(1)-a))+f32(p1*a)),'p0+a(p1-p0)':f32(p0+f32(a*f32(p1-p0))),
 'double':f32(float(p0)*(1-float(a))+float(p1)*float(a))}
print(cands)
