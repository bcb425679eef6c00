// rgbd.cu -- row f2 (RGB-D, partial): the two functions that are specific to the RGB-D front-end
// (reference src/frontend/RgbdVisionImuFrontend.cpp:183-209, :313-366), at the stage level:
//   * DepthFrame::getDetectionMask (src/frontend/DepthFrame.cpp:75-98): cv::inRange of the raw depth image against
//     [min_depth, max_depth] / depth_to_meters -> the detection mask FeatureDetector::featureDetection takes
//     (kvfe_detect_masked);
//   * RgbdFrame::fillStereoFrame (src/frontend/RgbdFrame.cpp:52-115): the "hallucinated" right frame -- per left
//     keypoint the depth at the truncated raw pixel (DepthFrame::getDepthAtPoint, DepthFrame.cpp:39-73), the virtual
//     disparity fx * virtual_baseline / depth, the right rectified keypoint, keypoints_depth_, keypoints_3d_ =
//     versor * depth / versor.z, and RgbdCamera::distortKeypoints (RgbdCamera.cpp:81-85) for right_frame_.keypoints_.
// Everything else the RGB-D front-end does per frame (tracking, Camera::undistortKeypoints, 2-point / 5-point and
// 1-point / 3-point outlier rejection on the fake stereo camera, detection) goes through entry points that already
// exist; the frame-level RGB-D step graph and PnP (Tracker.cpp:1064-1288) are not built.
// Float arithmetic follows the reference expression by expression (float depth, double fx_b, float disparity).
#include "common.cuh"

// depth_type: 0 = CV_16UC1, 1 = CV_32FC1
__global__ void __launch_bounds__(256) depth_mask_kernel(const unsigned char* __restrict__ depth, size_t pitch_bytes, int depth_type,
                                                         int W, int H, float lo, float hi, unsigned int lo16, unsigned int hi16,
                                                         unsigned char* __restrict__ mask, size_t mask_pitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const unsigned char* row = depth + (size_t)y * pitch_bytes;
  bool in;
  if (depth_type == 1) {
    const float v = reinterpret_cast<const float*>(row)[x];
    in = (v >= lo) && (v <= hi);                       // NaN fails both, as in cv::inRange
  } else {
    const unsigned int v = reinterpret_cast<const unsigned short*>(row)[x];
    in = (v >= lo16) && (v <= hi16);
  }
  mask[(size_t)y * mask_pitch + x] = in ? 255 : 0;
}

// DepthFrame::getDepthAtPoint: static_cast<int> truncation of the raw keypoint, NaN outside the image or below min_depth
__device__ __forceinline__ float depth_at_point(const unsigned char* depth, size_t pitch_bytes, int depth_type, int W, int H,
                                                float px, float py, float depth_to_meters, float min_depth) {
  const float nan = __int_as_float(0x7fc00000);
  const int x = (int)px, y = (int)py;
  if (x < 0 || x >= W || y < 0 || y >= H) return nan;
  const unsigned char* row = depth + (size_t)y * pitch_bytes;
  float d = (depth_type == 1) ? reinterpret_cast<const float*>(row)[x] : (float)reinterpret_cast<const unsigned short*>(row)[x];
  d *= depth_to_meters;
  if (d < min_depth) return nan;
  return d;
}

__global__ void __launch_bounds__(128) rgbd_fill_kernel(DevCfg dc, const CamModel* __restrict__ cams, const unsigned char* __restrict__ depth,
                                                        size_t pitch_bytes, int depth_type, float depth_to_meters, float min_depth,
                                                        double fx_b, const float* __restrict__ kp_x, const float* __restrict__ kp_y,
                                                        const int* __restrict__ left_status, const float* __restrict__ left_x,
                                                        const float* __restrict__ left_y, const double* __restrict__ versors, int n,
                                                        int* __restrict__ right_status, float* __restrict__ right_x,
                                                        float* __restrict__ right_y, double* __restrict__ depth_out,
                                                        double* __restrict__ p3d, float* __restrict__ right_kp_x,
                                                        float* __restrict__ right_kp_y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int rs = left_status[i];
  float rx = 0.f, ry = 0.f;
  double dout = 0.0, X = 0.0, Y = 0.0, Z = 0.0;
  if (rs == KVFE_KP_VALID) {
    const float kd = depth_at_point(depth, pitch_bytes, depth_type, dc.W, dc.H, kp_x[i], kp_y[i], depth_to_meters, min_depth);
    rs = KVFE_KP_NO_DEPTH;
    if (isfinite(kd)) {
      const float disparity = (float)(fx_b / (double)kd);
      const float uR = left_x[i] - disparity;
      if (!(uR < 0.0f)) {
        rs = KVFE_KP_VALID;
        rx = uR; ry = left_y[i];
        dout = (double)kd;
        const double vz = versors[3 * i + 2];
        X = versors[3 * i] * dout / vz; Y = versors[3 * i + 1] * dout / vz; Z = vz * dout / vz;
      }
    }
  }
  right_status[i] = rs; right_x[i] = rx; right_y[i] = ry;
  depth_out[i] = dout;
  p3d[3 * i] = X; p3d[3 * i + 1] = Y; p3d[3 * i + 2] = Z;
  // RgbdCamera::distortKeypoints -> UndistorterRectifier::distortUnrectifyKeypoints (UndistorterRectifier.cpp:213-228)
  float dx = 0.f, dy = 0.f;
  if (rs == KVFE_KP_VALID) {
    const int xx = clampi((int)roundf(rx), 0, dc.W - 1), yy = clampi((int)roundf(ry), 0, dc.H - 1);
    rect_map_at(cams[0], xx, yy, &dx, &dy);
  }
  right_kp_x[i] = dx; right_kp_y[i] = dy;
}

int launch_depth_mask(const DevCfg& dc, const unsigned char* depth, size_t pitch_bytes, int depth_type, float lo, float hi,
                      unsigned int lo16, unsigned int hi16, unsigned char* mask, size_t mask_pitch, cudaStream_t s) {
  dim3 grid((dc.W + 255) / 256, dc.H);
  depth_mask_kernel<<<grid, 256, 0, s>>>(depth, pitch_bytes, depth_type, dc.W, dc.H, lo, hi, lo16, hi16, mask, mask_pitch);
  return 1;
}

int launch_rgbd_fill(const DevCfg& dc, const CamModel* d_cam, const unsigned char* depth, size_t pitch_bytes, int depth_type,
                     float depth_to_meters, float min_depth, double fx_b, const float* kp_x, const float* kp_y, const int* left_status,
                     const float* left_x, const float* left_y, const double* versors, int n, int* right_status, float* right_x,
                     float* right_y, double* depth_out, double* p3d, float* right_kp_x, float* right_kp_y, cudaStream_t s) {
  rgbd_fill_kernel<<<(n + 127) / 128, 128, 0, s>>>(dc, d_cam, depth, pitch_bytes, depth_type, depth_to_meters, min_depth, fx_b, kp_x, kp_y,
                                                    left_status, left_x, left_y, versors, n, right_status, right_x, right_y, depth_out,
                                                    p3d, right_kp_x, right_kp_y);
  return 1;
}
