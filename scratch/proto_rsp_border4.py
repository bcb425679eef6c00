import numpy as np, cv2
Hh,Ww=480,752
np.set_printoptions(linewidth=250)
X=(np.zeros((Hh,1))+np.arange(Ww)[None,:]).astype(np.float32)   # value = x (float image)
Y=(np.arange(Hh)[:,None]+np.zeros((1,Ww))).astype(np.float32)
def dec(cx,cy):
    # use float32 source images to decode exact source coordinates (8u path may differ, check after)
    rx=cv2.getRectSubPix(X,(23,23),(cx,cy)); ry=cv2.getRectSubPix(Y,(23,23),(cx,cy))
    return rx,ry
X8=(np.arange(Ww)[None,:]%256+np.zeros((Hh,1))).astype(np.uint8); Y8=((np.arange(Hh)[:,None])%256+np.zeros((1,Ww))).astype(np.uint8)
for name,(cx,cy) in {'top-right':(743.0,1.0),'bottom-right':(743.0,478.0),'top-left':(8.0,1.0),'bottom-left':(8.0,478.0)}.items():
    rx=cv2.getRectSubPix(X8,(23,23),(cx,cy),patchType=cv2.CV_32F); ry=cv2.getRectSubPix(Y8,(23,23),(cx,cy),patchType=cv2.CV_32F)
    print(name,"center",(cx,cy))
    print(" x-source (mod 256) rows 0,9,10,11,12,22:")
    for i in [0,9,10,11,12,21,22]: print("  row",i, rx[i].astype(int))
    print(" y-source col 0 :", ry[:,0].astype(int)); print(" y-source col 22:", ry[:,22].astype(int))
