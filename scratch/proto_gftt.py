import numpy as np, cv2, glob, time
f32=np.float32
def refl(i,n):
    i=np.where(i<0,-i,i); i=np.where(i>=n,2*(n-1)-i,i); return i
def sobel_model(img, tail=True):
    H,W=img.shape
    I=img.astype(f32)
    s=f32(1/3060.0); s2=f32(2)*s
    xs=np.arange(W); xm=refl(xs-1,W); xp=refl(xs+1,W)
    ys=np.arange(H); ym=refl(ys-1,H); yp=refl(ys+1,H)
    r=I[:,xp]-I[:,xm]                     # exact
    # Dx = fma(r(y-1)+r(y+1), s, (2s)*r(y))
    a=(r[ym]+r[yp]).astype(np.float64)*np.float64(s)+ (s2*r).astype(np.float64)   # fma emulation: exact product in f64 + add, round once
    Dx=a.astype(f32)
    # t = fma(I(x+1), s, fma(I(x), 2s, s*I(x-1)))
    t0=(s*I[:,xm])                         # f32 product rounded
    t1=(I.astype(np.float64)*np.float64(s2)+t0.astype(np.float64)).astype(f32)
    t=(I[:,xp].astype(np.float64)*np.float64(s)+t1.astype(np.float64)).astype(f32)
    if tail:
        x0=16*((W-1)//16)
        tt=(s*I[:,xm]+s2*I)+s*I[:,xp]
        t[:,x0:]=tt[:,x0:]
    Dy=t[yp]-t[ym]
    return Dx,Dy
def mineig_model(img):
    H,W=img.shape
    Dx,Dy=sobel_model(img)
    xx=Dx*Dx; xy=Dx*Dy; yy=Dy*Dy
    xs=np.arange(W); xm=refl(xs-1,W); xp=refl(xs+1,W)
    out=[]
    for S in (xx,xy,yy):
        S64=S.astype(np.float64)
        R=(S64[:,xm]+S64)+S64[:,xp]
        # sequential column sum
        SUM=np.zeros(W); SUM=SUM+R[1]; SUM=SUM+R[0]
        D=np.empty((H,W),f32)
        for y in range(H):
            yp_=y+1 if y+1<H else H-2
            ym_=y-1 if y-1>=0 else 1
            s0=SUM+R[yp_]; D[y]=s0.astype(f32); SUM=s0-R[ym_]
        out.append(D)
    a=out[0]*f32(0.5); b=out[1]; c=out[2]*f32(0.5)
    return (a+c)-np.sqrt((a-c)*(a-c)+b*b)
files=sorted(glob.glob('/root/reference/tests/data/MicroEurocDataset/mav0/cam0/data/*.png'))[:3]+['/root/reference/tests/data/ForStereoFrame/left_fisheye_img_0.png']
for f in files:
    img=cv2.imread(f,0)
    ref=cv2.cornerMinEigenVal(img,3,ksize=3)
    dx=cv2.Sobel(img,cv2.CV_32F,1,0,ksize=3,scale=1/3060.0); dy=cv2.Sobel(img,cv2.CV_32F,0,1,ksize=3,scale=1/3060.0)
    Dx,Dy=sobel_model(img)
    m=mineig_model(img)
    print(img.shape,'dx mism',(Dx!=dx).sum(),'dy mism',(Dy!=dy).sum(),'eig mism',(m!=ref).sum(), 'max abs', np.abs(m-ref).max())
