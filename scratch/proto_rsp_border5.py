import numpy as np, cv2
Hh,Ww=480,752
np.set_printoptions(linewidth=250, precision=3, suppress=True)
X8=((np.arange(Ww)[None,:]-500)%256+np.zeros((Hh,1))).astype(np.uint8)      # x-500 near right edge (232..251), wraps elsewhere
X8L=((np.arange(Ww)[None,:])%256+np.zeros((Hh,1))).astype(np.uint8)
Y8=((np.arange(Hh)[:,None])%256+np.zeros((1,Ww))).astype(np.uint8)
Y8B=((np.arange(Hh)[:,None]-300)%256+np.zeros((1,Ww))).astype(np.uint8)
def show(name,cx,cy,ximg,xoff,yimg,yoff,rows,cols):
    rx=cv2.getRectSubPix(ximg,(23,23),(cx,cy),patchType=cv2.CV_32F)+xoff; ry=cv2.getRectSubPix(yimg,(23,23),(cx,cy),patchType=cv2.CV_32F)+yoff
    print(name,"center",(cx,cy),"nominal x0 = %.2f, y0 = %.2f"%(cx-11,cy-11))
    for i in rows:
        print("  row %2d (nominal y %.2f): eff y %s | eff x at cols %s: %s"%(i, cy-11+i, ry[i,[0,11,22]], cols, rx[i,cols]))
show("top-right",743.25,1.5,X8,500,Y8,0,[0,8,9,10,11,12],[16,17,18,19,20,22])
show("bottom-right",743.25,477.5,X8,500,Y8B,300,[10,11,12,13,14,22],[16,17,18,19,20,22])
show("top-left",8.25,1.5,X8L,0,Y8,0,[0,8,9,10,11,12],[0,1,2,3,4,5])
show("bottom-left",8.25,477.5,X8L,0,Y8B,300,[10,11,12,13,14,22],[0,1,2,3,4,5])
