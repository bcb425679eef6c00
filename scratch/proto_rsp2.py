import numpy as np, cv2, glob
f32=np.float32
img=cv2.imread(sorted(glob.glob('/root/reference/tests/data/MicroEurocDataset/mav0/cam0/data/*.png'))[10],0)
I=img.astype(f32); H,W=img.shape
def variants(cxf,cyf,pw=23,ph=23):
    cx=f32(f32(cxf)-f32(11)); cy=f32(f32(cyf)-f32(11))
    ipx=int(np.floor(cx)); ipy=int(np.floor(cy))
    a=f32(cx-f32(ipx)); b=f32(cy-f32(ipy))
    outs={}
    P00=I[ipy:ipy+ph,ipx:ipx+pw]; P01=I[ipy:ipy+ph,ipx+1:ipx+pw+1]; P10=I[ipy+1:ipy+ph+1,ipx:ipx+pw]; P11=I[ipy+1:ipy+ph+1,ipx+1:ipx+pw+1]
    a11=f32((f32(1)-a)*(f32(1)-b)); a12=f32(a*(f32(1)-b)); a21=f32((f32(1)-a)*b); a22=f32(a*b)
    outs['pair']=(P00*a11+P01*a12)+(P10*a21+P11*a22)
    outs['seq']=((P00*a11+P01*a12)+P10*a21)+P11*a22
    # recurrence
    ac=max(a,f32(0.0001)); a12r=f32(ac*(f32(1)-b)); a22r=f32(ac*b); b1=f32(f32(1)-b); b2=b; s=(1.0-float(ac))/float(ac)
    T=(a12r*I[ipy:ipy+ph,ipx:ipx+pw+1]+a22r*I[ipy+1:ipy+ph+1,ipx:ipx+pw+1]).astype(f32)   # t for columns 0..pw
    prev0=(f32(f32(1)-ac)*(b1*I[ipy:ipy+ph,ipx]+b2*I[ipy+1:ipy+ph+1,ipx])).astype(f32)
    rec=np.zeros((ph,pw),f32)
    rec[:,0]=prev0+T[:,1]
    rec[:,1:]=(T[:,1:pw].astype(np.float64)*s).astype(f32)+T[:,2:pw+1]
    outs['rec']=rec
    return outs
rng=np.random.default_rng(1)
cnt={}
N=0
for k in range(300):
    cx=rng.uniform(30,700); cy=rng.uniform(30,440)
    if k%3==0: cx=float(int(cx)); cy=float(int(cy))
    if k%3==1: cx=float(int(cx))
    ref=cv2.getRectSubPix(img,(23,23),(cx,cy),patchType=cv2.CV_32F)
    for n,v in variants(cx,cy).items():
        cnt[n]=cnt.get(n,0)+int(np.array_equal(v.astype(f32),ref))
    N+=1
print(N,cnt)
