import sys; sys.path.insert(0,'/root/repo')
import numpy as np, cv2
from kimera_vio_b200.params import CameraParams
from oracle.rig import StereoRig
def inv3(S):
    d=S[0,0]*(S[1,1]*S[2,2]-S[1,2]*S[2,1])-S[0,1]*(S[1,0]*S[2,2]-S[1,2]*S[2,0])+S[0,2]*(S[1,0]*S[2,1]-S[1,1]*S[2,0])
    d=1.0/d
    t=np.empty(9)
    t[0]=(S[1,1]*S[2,2]-S[1,2]*S[2,1])*d; t[1]=(S[0,2]*S[2,1]-S[0,1]*S[2,2])*d; t[2]=(S[0,1]*S[1,2]-S[0,2]*S[1,1])*d
    t[3]=(S[1,2]*S[2,0]-S[1,0]*S[2,2])*d; t[4]=(S[0,0]*S[2,2]-S[0,2]*S[2,0])*d; t[5]=(S[0,2]*S[1,0]-S[0,0]*S[1,2])*d
    t[6]=(S[1,0]*S[2,1]-S[1,1]*S[2,0])*d; t[7]=(S[0,1]*S[2,0]-S[0,0]*S[2,1])*d; t[8]=(S[0,0]*S[1,1]-S[0,1]*S[1,0])*d
    return t.reshape(3,3)
def mm(A,B,order):
    C=np.empty((3,3))
    for i in range(3):
        for j in range(3):
            p=[A[i,k]*B[k,j] for k in range(3)]
            C[i,j]= (p[0]+p[1])+p[2] if order==0 else p[0]+(p[1]+p[2])
    return C
def maps(K,D,R,P,W,H,order):
    iR=inv3(mm(P[:,:3],R,order))
    fx,fy,cx,cy=K[0,0],K[1,1],K[0,2],K[1,2]
    k1,k2,p1,p2=D.ravel()[:4]; k3=0.0
    u,v=np.meshgrid(np.arange(W,dtype=np.float64),np.arange(H,dtype=np.float64))
    X=(iR[0,0]*u+iR[0,1]*v)+iR[0,2]; Y=(iR[1,0]*u+iR[1,1]*v)+iR[1,2]; Wd=(iR[2,0]*u+iR[2,1]*v)+iR[2,2]
    results={}
    for name,(X_,Y_,W_) in {'direct':(X,Y,Wd), 'rowfirst':(v*iR[0,1]+iR[0,2]+u*iR[0,0], v*iR[1,1]+iR[1,2]+u*iR[1,0], v*iR[2,1]+iR[2,2]+u*iR[2,0])}.items():
        w=1.0/W_; x=X_*w; y=Y_*w
        x2=x*x; y2=y*y; r2=x2+y2; _2xy=2*x*y
        kr=(1+((k3*r2+k2)*r2+k1)*r2)/(1.0)
        xd=(x*kr+p1*_2xy+p2*(r2+2*x2)); yd=(y*kr+p1*(r2+2*y2)+p2*_2xy)
        mx=(fx*xd+cx).astype(np.float32); my=(fy*yd+cy).astype(np.float32)
        results[name]=(mx,my)
    return results
for rigname,(l,r) in {'euroc':(CameraParams.euroc_left(),CameraParams.euroc_right())}.items():
    rig=StereoRig(l,r)
    for cam,R,P,mx,my in ((l,rig.R1,rig.P1,rig.map_lx,rig.map_ly),(r,rig.R2,rig.P2,rig.map_rx,rig.map_ry)):
        for order in (0,1):
            res=maps(cam.K,cam.D,R,P,rig.W,rig.H,order)
            for name,(ax,ay) in res.items():
                print(rigname,order,name,(ax!=mx).sum(),(ay!=my).sum())
