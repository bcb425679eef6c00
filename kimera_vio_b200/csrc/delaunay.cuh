// delaunay.cuh -- row f1: the 2-D Delaunay mesh of Mesher::createMesh2dImpl (reference
// src/mesh/Mesher.cpp:1712-1817: cv::Subdiv2D subdiv(rect); subdiv.insert(keypoints);
// subdiv.getTriangleList(); keep the triangles whose three vertices lie inside the image).
//
// The triangle LIST (vertex order inside a triangle and order of the triangles) is part of the
// reference's output -- Mesher hashes vertices in that order (:1795-1806) -- and it is a function of
// cv::Subdiv2D's quad-edge bookkeeping (edge indices, free-list reuse, the walk of locate() from the most
// recent edge).  The structure is therefore restated operation by operation: incremental insertion
// (Guibas-Stibbe), the same predicates in the same precision (float coordinates, double products), the
// same edge numbering.  One thread builds one mesh; all state lives in caller-provided arrays (shared
// memory on the device).  Compiles for host and device: the CPU tests drive the very same code against
// cv2.Subdiv2D (tests/test_host_logic.py) -- plain C++, no CUDA needed.
#pragma once

#ifdef __CUDACC__
#define KVFE_HD __host__ __device__
#else
#define KVFE_HD
#endif

struct KvfeSubdiv {
  int* next;      // [4 * max_q]  QuadEdge::next
  int* ept;       // [4 * max_q]  QuadEdge::pt
  float* vx;      // [max_v]
  float* vy;
  int* vfirst;    // [max_v]      Vertex::firstEdge (doubles as the free list link)
  int nq, nv, max_q, max_v;
  int free_q, free_p, recent;
  float tlx, tly, brx, bry;
  int error;
};

namespace kvfe_dt {

// Subdiv2D edge-walk codes
enum { NEXT_AROUND_ORG = 0x00, NEXT_AROUND_DST = 0x22, PREV_AROUND_ORG = 0x11, PREV_AROUND_DST = 0x33,
       NEXT_AROUND_LEFT = 0x13, NEXT_AROUND_RIGHT = 0x31, PREV_AROUND_LEFT = 0x20, PREV_AROUND_RIGHT = 0x02 };
enum { PTLOC_ERROR = -2, PTLOC_OUTSIDE_RECT = -1, PTLOC_INSIDE = 0, PTLOC_VERTEX = 1, PTLOC_ON_EDGE = 2 };

KVFE_HD inline int sym(int e) { return e ^ 2; }
KVFE_HD inline int rot(int e, int r) { return (e & ~3) + ((e + r) & 3); }
KVFE_HD inline int get_edge(const KvfeSubdiv& s, int e, int t) {
  e = s.next[(e & ~3) + ((e + t) & 3)];
  return (e & ~3) + ((e + (t >> 4)) & 3);
}
KVFE_HD inline int edge_org(const KvfeSubdiv& s, int e) { return s.ept[e]; }
KVFE_HD inline int edge_dst(const KvfeSubdiv& s, int e) { return s.ept[(e & ~3) + ((e + 2) & 3)]; }

KVFE_HD inline double tri_area(float ax, float ay, float bx, float by, float cx, float cy) {
  return ((double)bx - ax) * ((double)cy - ay) - ((double)by - ay) * ((double)cx - ax);
}
KVFE_HD inline int right_of(const KvfeSubdiv& s, float px, float py, int e) {
  const int o = edge_org(s, e), d = edge_dst(s, e);
  const double cw = tri_area(px, py, s.vx[d], s.vy[d], s.vx[o], s.vy[o]);
  return (cw > 0) - (cw < 0);
}
KVFE_HD inline int in_circle3(float px, float py, float ax, float ay, float bx, float by, float cx, float cy) {
  const double eps = 1.1920928955078125e-07 * 0.125;
  double val = ((double)ax * ax + (double)ay * ay) * tri_area(bx, by, cx, cy, px, py);
  val -= ((double)bx * bx + (double)by * by) * tri_area(ax, ay, cx, cy, px, py);
  val += ((double)cx * cx + (double)cy * cy) * tri_area(ax, ay, bx, by, px, py);
  val -= ((double)px * px + (double)py * py) * tri_area(ax, ay, bx, by, cx, cy);
  return val > eps ? 1 : val < -eps ? -1 : 0;
}

KVFE_HD inline void init_qedge(KvfeSubdiv& s, int q, int e) {
  s.next[4 * q] = e; s.next[4 * q + 1] = e + 3; s.next[4 * q + 2] = e + 2; s.next[4 * q + 3] = e + 1;
  s.ept[4 * q] = s.ept[4 * q + 1] = s.ept[4 * q + 2] = s.ept[4 * q + 3] = 0;
}
KVFE_HD inline int new_edge(KvfeSubdiv& s) {
  if (s.free_q <= 0) {
    if (s.nq >= s.max_q) { s.error = 1; return 4; }
    for (int k = 0; k < 4; ++k) { s.next[4 * s.nq + k] = 0; s.ept[4 * s.nq + k] = 0; }   // qedges.push_back(QuadEdge())
    s.free_q = s.nq++;
  }
  const int e = s.free_q * 4;
  s.free_q = s.next[e + 1];
  init_qedge(s, e >> 2, e);
  return e;
}
KVFE_HD inline int new_point(KvfeSubdiv& s, float x, float y) {
  if (s.free_p == 0) {
    if (s.nv >= s.max_v) { s.error = 2; return 1; }
    s.vfirst[s.nv] = 0;
    s.free_p = s.nv++;
  }
  const int v = s.free_p;
  s.free_p = s.vfirst[v];
  s.vx[v] = x; s.vy[v] = y; s.vfirst[v] = 0;
  return v;
}
KVFE_HD inline void splice(KvfeSubdiv& s, int a, int b) {
  int& a_next = s.next[a];
  int& b_next = s.next[b];
  const int a_rot = rot(a_next, 1), b_rot = rot(b_next, 1);
  int& a_rot_next = s.next[a_rot];
  int& b_rot_next = s.next[b_rot];
  int t = a_next; a_next = b_next; b_next = t;
  t = a_rot_next; a_rot_next = b_rot_next; b_rot_next = t;
}
KVFE_HD inline void set_edge_points(KvfeSubdiv& s, int e, int org, int dst) {
  s.ept[e] = org;
  s.ept[(e & ~3) + ((e + 2) & 3)] = dst;
  s.vfirst[org] = e;
  s.vfirst[dst] = e ^ 2;
}
KVFE_HD inline void delete_edge(KvfeSubdiv& s, int e) {
  splice(s, e, get_edge(s, e, PREV_AROUND_ORG));
  const int se = sym(e);
  splice(s, se, get_edge(s, se, PREV_AROUND_ORG));
  const int q = e >> 2;
  s.next[4 * q] = 0;
  s.next[4 * q + 1] = s.free_q;
  s.free_q = q;
}
KVFE_HD inline int connect_edges(KvfeSubdiv& s, int a, int b) {
  const int e = new_edge(s);
  splice(s, e, get_edge(s, a, NEXT_AROUND_LEFT));
  splice(s, sym(e), b);
  set_edge_points(s, e, edge_dst(s, a), edge_org(s, b));
  return e;
}
KVFE_HD inline void swap_edges(KvfeSubdiv& s, int e) {
  const int se = sym(e);
  const int a = get_edge(s, e, PREV_AROUND_ORG);
  const int b = get_edge(s, se, PREV_AROUND_ORG);
  splice(s, e, a);
  splice(s, se, b);
  set_edge_points(s, e, edge_dst(s, a), edge_dst(s, b));
  splice(s, e, get_edge(s, a, NEXT_AROUND_LEFT));
  splice(s, se, get_edge(s, b, NEXT_AROUND_LEFT));
}

// Subdiv2D::initDelaunay(Rect(0, 0, w, h)).  `big_factor`: the bounding triangle's vertices sit at
// big_factor * max(w, h); OpenCV 4.13 (the oracle of this repo) uses 6, releases up to 4.5.x used 3 -- the
// triangulation near the image border and the edge numbering depend on it, so it is a parameter.
KVFE_HD inline void init(KvfeSubdiv& s, int w, int h, float big_factor = 6.f) {
  const float big = big_factor * (float)(w > h ? w : h);
  s.nq = 0; s.nv = 0; s.free_q = 0; s.free_p = 0; s.recent = 0; s.error = 0;
  s.tlx = 0.f; s.tly = 0.f; s.brx = (float)w; s.bry = (float)h;
  s.vx[0] = 0.f; s.vy[0] = 0.f; s.vfirst[0] = 0; s.nv = 1;          // vtx.push_back(Vertex())
  for (int k = 0; k < 4; ++k) { s.next[k] = 0; s.ept[k] = 0; }                      // qedges.push_back(QuadEdge())
  s.nq = 1;
  const int pA = new_point(s, big, 0.f), pB = new_point(s, 0.f, big), pC = new_point(s, -big, -big);
  const int eAB = new_edge(s), eBC = new_edge(s), eCA = new_edge(s);
  set_edge_points(s, eAB, pA, pB);
  set_edge_points(s, eBC, pB, pC);
  set_edge_points(s, eCA, pC, pA);
  splice(s, eAB, sym(eCA));
  splice(s, eBC, sym(eAB));
  splice(s, eCA, sym(eBC));
  s.recent = eAB;
}

KVFE_HD inline int locate(KvfeSubdiv& s, float px, float py, int* out_edge, int* out_vertex) {
  int vertex = 0;
  const int max_edges = s.nq * 4;
  if (px < s.tlx || py < s.tly || px >= s.brx || py >= s.bry) { *out_edge = 0; *out_vertex = 0; return PTLOC_OUTSIDE_RECT; }
  int edge = s.recent;
  int location = PTLOC_ERROR;
  int right_of_curr = right_of(s, px, py, edge);
  if (right_of_curr > 0) { edge = sym(edge); right_of_curr = -right_of_curr; }
  for (int i = 0; i < max_edges; ++i) {
    const int onext = s.next[edge];
    const int dprev = get_edge(s, edge, PREV_AROUND_DST);
    const int r_onext = right_of(s, px, py, onext);
    const int r_dprev = right_of(s, px, py, dprev);
    if (r_dprev > 0) {
      if (r_onext > 0 || (r_onext == 0 && right_of_curr == 0)) { location = PTLOC_INSIDE; break; }
      right_of_curr = r_onext; edge = onext;
    } else {
      if (r_onext > 0) {
        if (r_dprev == 0 && right_of_curr == 0) { location = PTLOC_INSIDE; break; }
        right_of_curr = r_dprev; edge = dprev;
      } else if (right_of_curr == 0) {
        const int d = edge_dst(s, onext);
        if (right_of(s, s.vx[d], s.vy[d], edge) >= 0) edge = sym(edge);
        else { right_of_curr = r_onext; edge = onext; }
      } else {
        right_of_curr = r_onext; edge = onext;
      }
    }
  }
  s.recent = edge;
  if (location == PTLOC_INSIDE) {
    const int o = edge_org(s, edge), d = edge_dst(s, edge);
    const float ox = s.vx[o], oy = s.vy[o], dx = s.vx[d], dy = s.vy[d];
    double t1 = fabs((double)(px - ox)); t1 += fabs((double)(py - oy));
    double t2 = fabs((double)(px - dx)); t2 += fabs((double)(py - dy));
    double t3 = fabs((double)(ox - dx)); t3 += fabs((double)(oy - dy));
    const double FE = 1.1920928955078125e-07;
    if (t1 < FE) { location = PTLOC_VERTEX; vertex = o; edge = 0; }
    else if (t2 < FE) { location = PTLOC_VERTEX; vertex = d; edge = 0; }
    else if ((t1 < t3 || t2 < t3) && fabs(tri_area(px, py, ox, oy, dx, dy)) < FE) { location = PTLOC_ON_EDGE; vertex = 0; }
  }
  if (location == PTLOC_ERROR) { edge = 0; vertex = 0; }
  *out_edge = edge; *out_vertex = vertex;
  return location;
}

// Subdiv2D::insert(Point2f); returns the vertex index, < 0 when the point was rejected
KVFE_HD inline int insert(KvfeSubdiv& s, float px, float py) {
  int curr_point = 0, curr_edge = 0;
  const int location = locate(s, px, py, &curr_edge, &curr_point);
  if (location == PTLOC_ERROR || location == PTLOC_OUTSIDE_RECT) { s.error = 3; return -1; }
  if (location == PTLOC_VERTEX) return curr_point;
  if (location == PTLOC_ON_EDGE) {
    const int deleted = curr_edge;
    s.recent = curr_edge = get_edge(s, curr_edge, PREV_AROUND_ORG);
    delete_edge(s, deleted);
  }
  curr_point = new_point(s, px, py);
  int base_edge = new_edge(s);
  const int first_point = edge_org(s, curr_edge);
  set_edge_points(s, base_edge, first_point, curr_point);
  splice(s, base_edge, curr_edge);
  do {
    base_edge = connect_edges(s, curr_edge, sym(base_edge));
    curr_edge = get_edge(s, base_edge, PREV_AROUND_ORG);
  } while (edge_dst(s, curr_edge) != first_point && !s.error);
  curr_edge = get_edge(s, base_edge, PREV_AROUND_ORG);
  const int max_edges = s.nq * 4;
  for (int i = 0; i < max_edges; ++i) {
    const int temp_edge = get_edge(s, curr_edge, PREV_AROUND_ORG);
    const int temp_dst = edge_dst(s, temp_edge);
    const int curr_org = edge_org(s, curr_edge), curr_dst = edge_dst(s, curr_edge);
    if (right_of(s, s.vx[temp_dst], s.vy[temp_dst], curr_edge) > 0 &&
        in_circle3(s.vx[curr_org], s.vy[curr_org], s.vx[temp_dst], s.vy[temp_dst], s.vx[curr_dst], s.vy[curr_dst],
                   s.vx[curr_point], s.vy[curr_point]) < 0) {
      swap_edges(s, curr_edge);
      curr_edge = get_edge(s, curr_edge, PREV_AROUND_ORG);
    } else if (curr_org == first_point) {
      break;
    } else {
      curr_edge = get_edge(s, get_edge(s, curr_edge, NEXT_AROUND_ORG), PREV_AROUND_LEFT);
    }
  }
  return curr_point;
}

KVFE_HD inline bool rect_contains(const KvfeSubdiv& s, float x, float y) {
  return s.tlx <= x && x < s.brx && s.tly <= y && y < s.bry;      // cv::Rect2f::contains
}

// Subdiv2D::getTriangleList + the Mesher's "good triangle" filter (both keep a triangle iff its three
// vertices are inside the image rectangle).  `mask`: 4 * nq bytes of scratch.  Returns the count;
// tri receives 6 floats per triangle (x0 y0 x1 y1 x2 y2), at most max_tri triangles are written.
KVFE_HD inline int triangle_list(const KvfeSubdiv& s, unsigned char* mask, float* tri, int max_tri) {
  const int total = s.nq * 4;
  for (int i = 0; i < total; ++i) mask[i] = 0;
  int n = 0;
  for (int i = 4; i < total; i += 2) {
    if (mask[i]) continue;
    const int ea = i;
    const int a = edge_org(s, ea);
    if (!rect_contains(s, s.vx[a], s.vy[a])) continue;
    const int eb = get_edge(s, ea, NEXT_AROUND_LEFT);
    const int b = edge_org(s, eb);
    if (!rect_contains(s, s.vx[b], s.vy[b])) continue;
    const int ec = get_edge(s, eb, NEXT_AROUND_LEFT);
    const int c = edge_org(s, ec);
    if (!rect_contains(s, s.vx[c], s.vy[c])) continue;
    mask[ea] = 1; mask[eb] = 1; mask[ec] = 1;
    if (n < max_tri) {
      float* t = tri + 6 * n;
      t[0] = s.vx[a]; t[1] = s.vy[a]; t[2] = s.vx[b]; t[3] = s.vy[b]; t[4] = s.vx[c]; t[5] = s.vy[c];
    }
    ++n;
  }
  return n;
}

}  // namespace kvfe_dt
