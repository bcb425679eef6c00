import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cv2
import helpers as H
s, fr = H.synth_frames(1, seed=31337)
img = fr[0].left; Hh,Ww=img.shape
np.set_printoptions(linewidth=200)
ref=cv2.getRectSubPix(img,(23,23),(743.0,1.0),patchType=cv2.CV_32F)
print("ref rows 0..12, cols 15..22 (x=747..754), y=-10..2")
print(ref[0:13,15:23].astype(int))
print("image rows 0..3, cols 745..751")
print(img[0:4,745:752])
# integer patch for a synthetic ramp image to decode the mapping
ramp=(np.arange(Hh)[:,None]*0+np.arange(Ww)[None,:]%251).astype(np.uint8)
r2=cv2.getRectSubPix(ramp,(23,23),(743.0,1.0),patchType=cv2.CV_32F)
print("ramp-x ref row0 cols 10..22:", r2[0,10:23].astype(int), " expected x%251:", [(733+j) for j in range(10,23)])
rampy=((np.arange(Hh)[:,None]*7)%251+np.zeros((1,Ww))).astype(np.uint8)
r3=cv2.getRectSubPix(rampy,(23,23),(743.0,1.0),patchType=cv2.CV_32F)
print("ramp-y ref col0 rows 0..22:", r3[:,0].astype(int))
print("ramp-y ref col22 rows 0..22:", r3[:,22].astype(int))
print("expected rows (7*y)%251 for y=clamp(-10..12):", [(7*max(0,y))%251 for y in range(-10,13)])
