import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cv2, itertools
import helpers as H
f32=np.float32
g,lefts,rights=H.golden()
A,B=lefts[0],lefts[1]
win=24
def refl(i,n):
    i=np.where(i<0,-i,i); return np.where(i>=n,2*(n-1)-i,i)
def scharr(I):
    Hh,Ww=I.shape; P=np.pad(I.astype(np.int32),1,mode='reflect')
    dx=3*(P[:-2,2:]-P[:-2,:-2])+10*(P[1:-1,2:]-P[1:-1,:-2])+3*(P[2:,2:]-P[2:,:-2])
    dy=3*(P[2:,:-2]-P[:-2,:-2])+10*(P[2:,1:-1]-P[:-2,1:-1])+3*(P[2:,2:]-P[:-2,2:])
    return dx,dy
DX,DY=scharr(A)
def patch(img,ix,iy,n,zero_out=False):
    Hh,Ww=img.shape
    ys=iy+np.arange(n); xs=ix+np.arange(n)
    if zero_out:
        out=np.zeros((n,n),np.int64)
        oky=(ys>=0)&(ys<Hh); okx=(xs>=0)&(xs<Ww)
        yy=np.clip(ys,0,Hh-1); xx=np.clip(xs,0,Ww-1)
        out=img[np.ix_(yy,xx)].astype(np.int64)*np.outer(oky,okx)
        return out
    return img[np.ix_(refl(ys,Hh),refl(xs,Ww))].astype(np.int64)
def cvround(x): return int(np.rint(f32(x)))
def weights(a,b):
    iw00=cvround(f32(f32(f32(1)-a)*f32(f32(1)-b))*f32(16384)); iw01=cvround(f32(a*f32(f32(1)-b))*f32(16384)); iw10=cvround(f32(f32(f32(1)-a)*b)*f32(16384))
    return iw00,iw01,iw10,16384-iw00-iw01-iw10
def descale(x,n): return (x+(1<<(n-1)))>>n
def one_iter(pt,acc):
    half=f32(11.5)
    px=f32(pt[0])-half; py=f32(pt[1])-half
    ipx=int(np.floor(px)); ipy=int(np.floor(py))
    a=f32(px-f32(ipx)); b=f32(py-f32(ipy))
    w=weights(a,b)
    Ip=patch(A,ipx,ipy,win+1); Dxp=patch(DX,ipx,ipy,win+1,True); Dyp=patch(DY,ipx,ipy,win+1,True)
    def bil(P,sh): return descale(P[:-1,:-1]*w[0]+P[:-1,1:]*w[1]+P[1:,:-1]*w[2]+P[1:,1:]*w[3],sh)
    I=bil(Ip,9); Ix=bil(Dxp,14); Iy=bil(Dyp,14)
    A11,A12,A22=acc(Ix*Ix),acc(Ix*Iy),acc(Iy*Iy)
    sc=f32(1.0/(1<<20))
    A11=f32(A11*sc);A12=f32(A12*sc);A22=f32(A22*sc)
    D=f32(f32(A11*A22)-f32(A12*A12))
    D=f32(f32(1)/D)
    qx,qy=px,py   # init = prev
    iqx=int(np.floor(qx)); iqy=int(np.floor(qy))
    a=f32(qx-f32(iqx)); b=f32(qy-f32(iqy)); w=weights(a,b)
    Jp=patch(B,iqx,iqy,win+1)
    Jv=descale(Jp[:-1,:-1]*w[0]+Jp[:-1,1:]*w[1]+Jp[1:,:-1]*w[2]+Jp[1:,1:]*w[3],9)
    diff=Jv-I
    b1=f32(acc(diff*Ix)*sc); b2=f32(acc(diff*Iy)*sc)
    dx=f32(f32(f32(A12*b2)-f32(A22*b1))*D); dy=f32(f32(f32(A12*b1)-f32(A11*b2))*D)
    return f32(f32(qx+dx)+half), f32(f32(qy+dy)+half)
def acc_exact(M): return f32(int(M.sum()))
def acc_seq(M):
    s=f32(0)
    for v in M.reshape(-1): s=f32(s+f32(v))
    return s
def make_lane_acc(nl, pairdot=False, final='seq'):
    def acc(M):
        lanes=[f32(0)]*nl
        flat=M.reshape(-1)
        if pairdot:
            # int32 pair sums then float, lanes over pairs
            pairs=flat.reshape(-1,2).sum(axis=1)
            for i,v in enumerate(pairs): 
                l=i%nl; lanes[l]=f32(lanes[l]+f32(int(v)))
        else:
            for i,v in enumerate(flat):
                l=i%nl; lanes[l]=f32(lanes[l]+f32(int(v)))
        if final=='seq':
            s=f32(0)
            for l in lanes: s=f32(s+l)
            return s
        else:
            return f32(f32(lanes[0]+lanes[1])+f32(lanes[2]+lanes[3])) if nl==4 else None
    return acc
pts=cv2.goodFeaturesToTrack(A,60,0.001,20).reshape(-1,2)
crit=(cv2.TERM_CRITERIA_COUNT,1,0.0)
ref,st,_=cv2.calcOpticalFlowPyrLK(A,B,pts.reshape(-1,1,2),None,winSize=(24,24),maxLevel=0,criteria=crit)
ref=ref.reshape(-1,2)
hyps={'exact':acc_exact,'seq':acc_seq,'lane4':make_lane_acc(4),'lane4_pair':make_lane_acc(4,True),'lane4_tree':make_lane_acc(4,False,'tree'),'lane4_pair_tree':make_lane_acc(4,True,'tree'),'lane8':make_lane_acc(8)}
for name,acc in hyps.items():
    ok=0
    for i,p in enumerate(pts):
        if not st[i]: continue
        r=one_iter(p,acc)
        ok+= (r[0]==ref[i,0] and r[1]==ref[i,1])
    print(name,ok,'/',int(st.sum()))
print('--- bigger experiment')
def acc_seq(M): return np.cumsum(M.reshape(-1).astype(f32),dtype=f32)[-1]
def lane(nl,final):
    def acc(M):
        L=np.cumsum(M.reshape(-1,nl).astype(f32),axis=0,dtype=f32)[-1]
        if final=='seq':
            s=f32(0)
            for l in L: s=f32(s+l)
            return s
        if final=='tree' and nl==4: return f32(f32(L[0]+L[1])+f32(L[2]+L[3]))
        if final=='tree' and nl==8: return f32(f32(f32(L[0]+L[1])+f32(L[2]+L[3]))+f32(f32(L[4]+L[5])+f32(L[6]+L[7])))
        if final=='hadd' and nl==4: return f32(f32(L[0]+L[2])+f32(L[1]+L[3]))
    return acc
def rowlane(nl):
    # per-row sequential sums then add rows sequentially
    def acc(M):
        r=np.cumsum(M.astype(f32),axis=1,dtype=f32)[:,-1]
        return np.cumsum(r,dtype=f32)[-1]
    return acc
hyps={'exact':acc_exact,'seq':acc_seq,'lane4_seq':lane(4,'seq'),'lane4_tree':lane(4,'tree'),'lane4_hadd':lane(4,'hadd'),'lane8_seq':lane(8,'seq'),'lane8_tree':lane(8,'tree'),'rows':rowlane(0)}
rng=np.random.default_rng(0)
tot={k:0 for k in hyps}; n=0
for (A_,B_) in ((lefts[0],lefts[1]),(lefts[1],lefts[3]),(lefts[2],lefts[4])):
    A=A_;B=B_; DX,DY=scharr(A)
    pts=np.concatenate([cv2.goodFeaturesToTrack(A,250,0.001,10).reshape(-1,2), np.stack([rng.uniform(30,720,150),rng.uniform(30,450,150)],1).astype(f32)]).astype(f32)
    ref,st,_=cv2.calcOpticalFlowPyrLK(A,B,pts.reshape(-1,1,2),None,winSize=(24,24),maxLevel=0,criteria=crit)
    ref=ref.reshape(-1,2)
    for i,p in enumerate(pts):
        if not st[i]: continue
        n+=1
        for name,acc in hyps.items():
            try:
                r=one_iter(p,acc)
            except Exception as e:
                continue
            tot[name]+= (r[0]==ref[i,0] and r[1]==ref[i,1])
print(n,tot)
print('--- mixed hypotheses (A sums / b sums)')
def one_iter2(pt,accA,accB):
    half=f32(11.5)
    px=f32(pt[0])-half; py=f32(pt[1])-half
    ipx=int(np.floor(px)); ipy=int(np.floor(py))
    a=f32(px-f32(ipx)); b=f32(py-f32(ipy))
    w=weights(a,b)
    Ip=patch(A,ipx,ipy,win+1); Dxp=patch(DX,ipx,ipy,win+1,True); Dyp=patch(DY,ipx,ipy,win+1,True)
    def bil(P,sh): return descale(P[:-1,:-1]*w[0]+P[:-1,1:]*w[1]+P[1:,:-1]*w[2]+P[1:,1:]*w[3],sh)
    I=bil(Ip,9); Ix=bil(Dxp,14); Iy=bil(Dyp,14)
    A11,A12,A22=accA(Ix*Ix),accA(Ix*Iy),accA(Iy*Iy)
    sc=f32(1.0/(1<<20))
    A11=f32(A11*sc);A12=f32(A12*sc);A22=f32(A22*sc)
    D=f32(f32(A11*A22)-f32(A12*A12)); D=f32(f32(1)/D)
    qx,qy=px,py
    iqx=int(np.floor(qx)); iqy=int(np.floor(qy))
    a=f32(qx-f32(iqx)); b=f32(qy-f32(iqy)); w=weights(a,b)
    Jp=patch(B,iqx,iqy,win+1)
    Jv=descale(Jp[:-1,:-1]*w[0]+Jp[:-1,1:]*w[1]+Jp[1:,:-1]*w[2]+Jp[1:,1:]*w[3],9)
    diff=Jv-I
    b1=f32(accB(diff*Ix)*sc); b2=f32(accB(diff*Iy)*sc)
    dx=f32(f32(f32(A12*b2)-f32(A22*b1))*D); dy=f32(f32(f32(A12*b1)-f32(A11*b2))*D)
    return f32(f32(qx+dx)+half), f32(f32(qy+dy)+half)
def lane_pair(nl):
    def acc(M):
        pairs=M.reshape(-1,2).sum(axis=1)
        L=np.cumsum(pairs.reshape(-1,nl).astype(f32),axis=0,dtype=f32)[-1]
        s=f32(0)
        for l in L: s=f32(s+l)
        return s
    return acc
cands={'exact':acc_exact,'seq':acc_seq,'lane4':lane(4,'seq'),'lane8':lane(8,'seq'),'lane4pair':lane_pair(4),'lane2pair':lane_pair(2)}
tot={}
n=0
for (A_,B_) in ((lefts[0],lefts[1]),(lefts[1],lefts[3])):
    A=A_;B=B_; DX,DY=scharr(A)
    pts=np.concatenate([cv2.goodFeaturesToTrack(A,250,0.001,10).reshape(-1,2), np.stack([rng.uniform(30,720,150),rng.uniform(30,450,150)],1).astype(f32)]).astype(f32)
    ref,st,_=cv2.calcOpticalFlowPyrLK(A,B,pts.reshape(-1,1,2),None,winSize=(24,24),maxLevel=0,criteria=crit)
    ref=ref.reshape(-1,2)
    for i,p in enumerate(pts):
        if not st[i]: continue
        n+=1
        for na,fa in cands.items():
            for nb,fb in cands.items():
                if na!='lane4' and nb!='lane4': continue
                r=one_iter2(p,fa,fb)
                tot[(na,nb)]=tot.get((na,nb),0)+(r[0]==ref[i,0] and r[1]==ref[i,1])
print(n,tot)
print('--- b-sum discrimination with large residuals')
tot={}; n=0
cands_b={'exact':acc_exact,'seq':acc_seq,'lane4':lane(4,'seq'),'lane4_tree':lane(4,'tree'),'lane4_hadd':lane(4,'hadd'),'lane8':lane(8,'seq'),'lane4pair':lane_pair(4),'lane2pair':lane_pair(2)}
synthA=lefts[0]; 
for (A_,B_) in ((lefts[0],255-lefts[4]),(lefts[0],np.roll(lefts[2],7,axis=1)),(lefts[1],rights[1])):
    A=A_;B=np.ascontiguousarray(B_); DX,DY=scharr(A)
    pts=np.concatenate([cv2.goodFeaturesToTrack(A,300,0.001,10).reshape(-1,2)]).astype(f32)
    ref,st,_=cv2.calcOpticalFlowPyrLK(A,B,pts.reshape(-1,1,2),None,winSize=(24,24),maxLevel=0,criteria=crit)
    ref=ref.reshape(-1,2)
    for i,p in enumerate(pts):
        if not st[i]: continue
        n+=1
        for nb,fb in cands_b.items():
            r=one_iter2(p,lane(4,'seq'),fb)
            tot[nb]=tot.get(nb,0)+(r[0]==ref[i,0] and r[1]==ref[i,1])
print(n,tot)
print('--- pair8 hypothesis for b sums')
def acc_pair8(M):
    # groups of 8 consecutive pixels in row-major order; 4 accumulators over pairs (k, k+4)
    Gp=M.reshape(-1,8)
    P=(Gp[:,:4]+Gp[:,4:]).astype(np.int64)          # int32 pair sums, columns: pairs (0,4),(1,5),(2,6),(3,7)
    L=np.cumsum(P.astype(f32),axis=0,dtype=f32)[-1] # A,B,C,D
    return f32(f32(L[0]+L[2])+f32(L[1]+L[3]))
tot={'pair8':0,'exact':0}; n=0
for (A_,B_) in ((lefts[0],255-lefts[4]),(lefts[0],np.roll(lefts[2],7,axis=1)),(lefts[1],rights[1]),(lefts[0],lefts[1])):
    A=A_;B=np.ascontiguousarray(B_); DX,DY=scharr(A)
    pts=np.concatenate([cv2.goodFeaturesToTrack(A,300,0.001,10).reshape(-1,2)]).astype(f32)
    ref,st,_=cv2.calcOpticalFlowPyrLK(A,B,pts.reshape(-1,1,2),None,winSize=(24,24),maxLevel=0,criteria=crit)
    ref=ref.reshape(-1,2)
    for i,p in enumerate(pts):
        if not st[i]: continue
        n+=1
        for nb,fb in (('pair8',acc_pair8),('exact',acc_exact)):
            r=one_iter2(p,lane(4,'seq'),fb)
            tot[nb]+=(r[0]==ref[i,0] and r[1]==ref[i,1])
print(n,tot)
print('--- failing cases analysis')
fails=[]
for (A_,B_) in ((lefts[0],255-lefts[4]),(lefts[0],np.roll(lefts[2],7,axis=1)),(lefts[1],rights[1])):
    A=A_;B=np.ascontiguousarray(B_); DX,DY=scharr(A)
    pts=np.concatenate([cv2.goodFeaturesToTrack(A,300,0.001,10).reshape(-1,2)]).astype(f32)
    ref,st,_=cv2.calcOpticalFlowPyrLK(A,B,pts.reshape(-1,1,2),None,winSize=(24,24),maxLevel=0,criteria=crit)
    ref=ref.reshape(-1,2)
    for i,p in enumerate(pts):
        if not st[i]: continue
        res={}
        for na,fa in (('l4seq',lane(4,'seq')),('l4tree',lane(4,'tree')),('l4hadd',lane(4,'hadd'))):
            for nb,fb in (('pair8',acc_pair8),('exact',acc_exact),('lane4',lane(4,'seq'))):
                r=one_iter2(p,fa,fb)
                res[(na,nb)]=(r[0]==ref[i,0] and r[1]==ref[i,1])
        if not all(res.values()):
            r=one_iter2(p,lane(4,'seq'),acc_pair8)
            fails.append((tuple(p),tuple(ref[i]),r,{k:v for k,v in res.items() if v}))
for f in fails: print(f)
print('--- final check l4hadd + pair8')
tot=0; n=0
for (A_,B_) in ((lefts[0],255-lefts[4]),(lefts[0],np.roll(lefts[2],7,axis=1)),(lefts[1],rights[1]),(lefts[0],lefts[1]),(lefts[2],lefts[3]),(rights[0],rights[2])):
    A=A_;B=np.ascontiguousarray(B_); DX,DY=scharr(A)
    pts=np.concatenate([cv2.goodFeaturesToTrack(A,400,0.0005,8).reshape(-1,2)]).astype(f32)
    ref,st,_=cv2.calcOpticalFlowPyrLK(A,B,pts.reshape(-1,1,2),None,winSize=(24,24),maxLevel=0,criteria=crit)
    ref=ref.reshape(-1,2)
    for i,p in enumerate(pts):
        if not st[i]: continue
        n+=1
        r=one_iter2(p,lane(4,'hadd'),acc_pair8)
        tot+=(r[0]==ref[i,0] and r[1]==ref[i,1])
print(n,tot)
