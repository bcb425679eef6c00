// subpix.cuh -- cv::cornerSubPix / cv::getRectSubPix device code shared by the detector
// (FeatureDetector.cpp:283-296) and the optional stereo refinement (StereoMatcher.cpp:404-413).
#pragma once
#include "common.cuh"

#define SUBPIX_MAX_WIN 12   // window half-size supported (Euroc: 10 -> 21x21 window, 23x23 patch)
#define SUBPIX_PATCH ((2 * SUBPIX_MAX_WIN + 3) * (2 * SUBPIX_MAX_WIN + 3))

// cv::getRectSubPix(u8 -> f32), win (pw x ph), centre (cxf, cyf); one warp cooperatively.
static __device__ void get_rect_subpix(const unsigned char* __restrict__ img, int pitch, int W, int H, float cxf,
                                float cyf, int pw, int ph, float* __restrict__ buf, int lane) {
  float cx = cxf - (pw - 1) * 0.5f, cy = cyf - (ph - 1) * 0.5f;
  int ipx = cv_floor(cx), ipy = cv_floor(cy);
  // cv2 4.13, pinned against cv2.getRectSubPix on 208 border patches + 300 interior ones (all bit-exact):
  //  * rows and columns of the window that have both taps inside: (P00*a11 + P01*a12) + (P10*a21 + P11*a22);
  //  * window columns left of / at-or-right-of the last image column: 2-tap vertical blend of the edge column
  //    -- except that rows ABOVE the image take column W-2 on the right side (a quirk of the library's
  //    border path, reproduced);
  //  * window rows above the image / at-or-below the last image row: fma(P01, a, P00 * (1 - a)) of the
  //    edge row.
  {
    float a = cx - ipx, bq = cy - ipy;
    float a11 = (1.f - a) * (1.f - bq), a12 = a * (1.f - bq), a21 = (1.f - a) * bq, a22 = a * bq;
    float b1 = 1.f - bq, b2 = bq, a1 = 1.f - a;
    int rx, rw, ry, rh;
    if (ipx >= 0) rx = 0; else { rx = -ipx; if (rx > pw) rx = pw; }
    if (ipx < W - pw) rw = pw; else { rw = W - ipx - 1; if (rw < 0) rw = 0; }
    if (ipy >= 0) ry = 0; else ry = -ipy;
    if (ipy < H - ph) rh = ph; else { rh = H - ipy - 1; if (rh < 0) rh = 0; }
    for (int i = lane; i < pw * ph; i += 32) {
      int r = i / pw, j = i - r * pw;
      const bool row_out = r < ry || r >= rh;
      int y0 = clampi(ipy + r, 0, H - 1);
      int y1 = row_out ? y0 : clampi(ipy + r + 1, 0, H - 1);
      const unsigned char* p0 = img + (size_t)y0 * pitch;
      const unsigned char* p1 = img + (size_t)y1 * pitch;
      float v;
      if (j < rx) { int xc = clampi(ipx + rx, 0, W - 1); v = p0[xc] * b1 + p1[xc] * b2; }
      else if (j >= rw) {
        int xc = clampi(ipx + rw, 0, W - 1);
        if (r < ry) xc = max(W - 2, 0);
        v = p0[xc] * b1 + p1[xc] * b2;
      } else {
        int x0 = clampi(ipx + j, 0, W - 1), x1 = clampi(ipx + j + 1, 0, W - 1);
        if (row_out) v = fmaf((float)p0[x1], a, p0[x0] * a1);
        else v = (p0[x0] * a11 + p0[x1] * a12) + (p1[x0] * a21 + p1[x1] * a22);
      }
      buf[i] = v;
    }
  }
}

// Refines (x, y) in place; returns through pointers. One warp.
static __device__ void corner_subpix_warp(const unsigned char* __restrict__ img, int pitch, int W, int H, int win,
                                   int max_iters, double eps2, const float* __restrict__ gmask,
                                   float* __restrict__ buf, float* px, float* py, int lane) {
  const int ww = 2 * win + 1, pw = ww + 2;
  const float cTx = *px, cTy = *py;
  float cIx = cTx, cIy = cTy;
  int iter = 0;
  double err = 0;
  do {
    get_rect_subpix(img, pitch, W, H, cIx, cIy, pw, pw, buf, lane);
    __syncwarp();
    double a = 0, bsum = 0, c = 0, bb1 = 0, bb2 = 0;
    for (int k = lane; k < ww * ww; k += 32) {
      int i = k / ww, j = k - i * ww;
      const float* sp = buf + (i + 1) * pw + (j + 1);
      double m = gmask[k];
      double tgx = sp[1] - sp[-1];
      double tgy = sp[pw] - sp[-pw];
      double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
      double pxx = j - win, pyy = i - win;
      a += gxx; bsum += gxy; c += gyy;
      bb1 += gxx * pxx + gxy * pyy;
      bb2 += gxy * pxx + gyy * pyy;
    }
    a = warp_sum_d(a); bsum = warp_sum_d(bsum); c = warp_sum_d(c);
    bb1 = warp_sum_d(bb1); bb2 = warp_sum_d(bb2);
    __syncwarp();
    double det = a * c - bsum * bsum;
    if (fabs(det) <= 2.220446049250313e-16 * 2.220446049250313e-16) break;
    double scale = 1.0 / det;
    float nx = (float)(cIx + c * scale * bb1 - bsum * scale * bb2);
    float ny = (float)(cIy - bsum * scale * bb1 + a * scale * bb2);
    err = (nx - cIx) * (nx - cIx) + (ny - cIy) * (ny - cIy);
    // cv2 4.13: a step that leaves the image is discarded (the previous estimate is kept) -- pinned
    // against cv2.cornerSubPix on a corner 3 px from the right border (tests: min_distance_8 variant)
    if (nx < 0 || nx >= W || ny < 0 || ny >= H) break;
    cIx = nx; cIy = ny;
  } while (++iter < max_iters && err > eps2);
  if (fabsf(cIx - cTx) > win || fabsf(cIy - cTy) > win) { cIx = cTx; cIy = cTy; }
  *px = cIx; *py = cIy;
}

