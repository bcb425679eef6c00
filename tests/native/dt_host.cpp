// Test harness (CPU): drives kimera_vio_b200/csrc/delaunay.cuh -- the code the mesh kernel runs -- compiled
// for the host, so that the quad-edge restatement can be checked against cv2.Subdiv2D without a GPU.
// Built on the fly by tests/test_host_logic.py (g++); not part of libkvfe.so.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "../../kimera_vio_b200/csrc/delaunay.cuh"

extern "C" int dt_host_mesh(int w, int h, const float* xy, int n, float* tri, int max_tri, int* n_qedges) {
  const int max_q = 3 * n + 16, max_v = n + 8;
  std::vector<int> next(4 * max_q), ept(4 * max_q), vfirst(max_v);
  std::vector<float> vx(max_v), vy(max_v);
  std::vector<unsigned char> mask(4 * max_q);
  KvfeSubdiv s;
  s.next = next.data(); s.ept = ept.data(); s.vx = vx.data(); s.vy = vy.data(); s.vfirst = vfirst.data();
  s.max_q = max_q; s.max_v = max_v;
  kvfe_dt::init(s, w, h);
  for (int i = 0; i < n; ++i) {
    const float x = xy[2 * i], y = xy[2 * i + 1];
    // Mesher.cpp:1733-1749: rect.contains(kp) && kp.x >= 0 && kp.y >= 0
    if (!(kvfe_dt::rect_contains(s, x, y) && x >= 0.f && y >= 0.f)) continue;
    kvfe_dt::insert(s, x, y);
  }
  if (n_qedges) *n_qedges = s.nq;
  if (s.error) return -s.error;
  return kvfe_dt::triangle_list(s, mask.data(), tri, max_tri);
}

// debugging aid: the quad-edge arrays after inserting the first n points
extern "C" int dt_host_state(int w, int h, const float* xy, int n, int* next_out, int* ept_out, int cap_q, int* recent) {
  const int max_q = 3 * n + 16, max_v = n + 8;
  std::vector<int> next(4 * max_q), ept(4 * max_q), vfirst(max_v);
  std::vector<float> vx(max_v), vy(max_v);
  KvfeSubdiv s;
  s.next = next.data(); s.ept = ept.data(); s.vx = vx.data(); s.vy = vy.data(); s.vfirst = vfirst.data();
  s.max_q = max_q; s.max_v = max_v;
  kvfe_dt::init(s, w, h);
  for (int i = 0; i < n; ++i) kvfe_dt::insert(s, xy[2 * i], xy[2 * i + 1]);
  for (int i = 0; i < 4 * s.nq && i < 4 * cap_q; ++i) { next_out[i] = next[i]; ept_out[i] = ept[i]; }
  *recent = s.recent;
  return s.nq;
}

// debugging aid: locate() of point n after inserting the first n points, from a chosen start edge (<= 0: recentEdge)
extern "C" int dt_host_locate(int w, int h, const float* xy, int n, int start_edge, int* edge_out, int* recent_before) {
  const int max_q = 3 * n + 16, max_v = n + 8;
  std::vector<int> next(4 * max_q), ept(4 * max_q), vfirst(max_v);
  std::vector<float> vx(max_v), vy(max_v);
  KvfeSubdiv s;
  s.next = next.data(); s.ept = ept.data(); s.vx = vx.data(); s.vy = vy.data(); s.vfirst = vfirst.data();
  s.max_q = max_q; s.max_v = max_v;
  kvfe_dt::init(s, w, h);
  for (int i = 0; i < n; ++i) kvfe_dt::insert(s, xy[2 * i], xy[2 * i + 1]);
  *recent_before = s.recent;
  if (start_edge > 0) s.recent = start_edge;
  int e = 0, v = 0;
  int loc = kvfe_dt::locate(s, xy[2 * n], xy[2 * n + 1], &e, &v);
  *edge_out = e;
  return loc;
}
