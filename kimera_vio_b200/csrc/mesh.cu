// mesh.cu -- row f1: the Mesher's 2-D Delaunay mesh right downstream of the front-end packet.
//   Mesher::createMesh2dStereo  reference src/mesh/Mesher.cpp:1849-1886  (keypoints with a VALID right match
//                               and a live landmark)
//   Mesher::createMesh2dImpl    reference src/mesh/Mesher.cpp:1712-1817  (cv::Subdiv2D incremental Delaunay,
//                               getTriangleList, triangles with all vertices inside the image)
// cv::Subdiv2D's triangle list (order of the triangles, first vertex of each) is a function of its quad-edge
// bookkeeping, so the structure is restated operation by operation (delaunay.cuh) and one thread walks it --
// incremental insertion is sequential by construction.  Parallelism comes from the batch: one warp per camera
// stream, the quad-edge arrays of a stream in shared memory (global scratch when they do not fit), lanes
// cooperate on the ordered selection of the input keypoints and on the copy of the result.
#include "common.cuh"
#include "delaunay.cuh"

struct MeshWs {          // carve-up of one stream's workspace (ints): see mesh_ws_ints()
  int* next; int* ept; int* vfirst; float* vx; float* vy; float* in_x; float* in_y; unsigned char* mask;
};
__host__ __device__ inline size_t mesh_ws_ints(int cap) {
  const size_t q = 3 * (size_t)cap + 16, v = (size_t)cap + 8;
  return 8 * q + 3 * v + 2 * (size_t)cap + q + 8;      // next, ept | vfirst, vx, vy | in_x, in_y | mask (4q bytes)
}
__device__ inline MeshWs mesh_carve(int* base, int cap) {
  const size_t q = 3 * (size_t)cap + 16, v = (size_t)cap + 8;
  MeshWs w;
  w.next = base; w.ept = base + 4 * q; w.vfirst = base + 8 * q;
  w.vx = reinterpret_cast<float*>(base + 8 * q + v); w.vy = w.vx + v;
  w.in_x = w.vy + v; w.in_y = w.in_x + cap;
  w.mask = reinterpret_cast<unsigned char*>(w.in_y + cap);
  return w;
}

// lane 0 builds the triangulation of in_x/in_y[0..n) and writes the triangle list; returns the count
__device__ int mesh_build(const DevCfg& dc, MeshWs& w, int n, float* tri, int max_tri) {
  int n_tri = 0;
  if ((threadIdx.x & 31) == 0) {
    KvfeSubdiv s;
    s.next = w.next; s.ept = w.ept; s.vx = w.vx; s.vy = w.vy; s.vfirst = w.vfirst;
    s.max_q = 3 * dc.cap + 16; s.max_v = dc.cap + 8;
    kvfe_dt::init(s, dc.W, dc.H, dc.subdiv_factor);
    for (int i = 0; i < n && !s.error; ++i) {
      const float x = w.in_x[i], y = w.in_y[i];
      // Mesher.cpp:1733-1749: rect.contains(kp) && kp.x >= 0 && kp.y >= 0
      if (kvfe_dt::rect_contains(s, x, y) && x >= 0.f && y >= 0.f) kvfe_dt::insert(s, x, y);
    }
    n_tri = s.error ? 0 : kvfe_dt::triangle_list(s, w.mask, tri, max_tri);
  }
  return __shfl_sync(KVFE_FULL_MASK, n_tri, 0);
}

// frame level: grid B, one warp per stream; keyframes of the nominal spin and the bootstrap frame
__global__ void __launch_bounds__(32) mesh_kernel(DevCfg dc, DevBuf db, int use_smem) {
  extern __shared__ __align__(16) int mesh_smem[];
  const int b = blockIdx.x, lane = threadIdx.x;
  const StreamState& st = db.st[b];
  if (!(st.mode == 0 || st.mode == 2)) return;
  unsigned char* pk = db.packets + (size_t)b * db.packet_bytes;
  kvfe_packet_header* h = reinterpret_cast<kvfe_packet_header*>(pk);
  float* tri = reinterpret_cast<float*>(pk + db.pk_off[20]);
  MeshWs w = mesh_carve(use_smem ? mesh_smem : db.mesh_ws + (size_t)b * mesh_ws_ints(dc.cap), dc.cap);
  // ordered selection: right keypoint VALID and landmark != -1 (createMesh2dStereo)
  const int fk = b * 3 + st.slot_km1;          // finalize has already rotated the slots: km1 is this frame
  const int n = db.fr.n[fk];
  int m = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const size_t k = (size_t)fk * dc.cap + i;
    const bool keep = i < n && db.fr.rstat[k] == KVFE_KP_VALID && db.fr.lmk[k] != -1;
    const unsigned bal = __ballot_sync(KVFE_FULL_MASK, keep);
    if (keep) {
      const int pos = m + __popc(bal & ((1u << lane) - 1));
      w.in_x[pos] = db.fr.kx[k]; w.in_y[pos] = db.fr.ky[k];
    }
    m += __popc(bal);
  }
  __syncwarp();
  const int n_tri = mesh_build(dc, w, m, tri, 2 * dc.cap);
  if (lane == 0) h->n_mesh_triangles = n_tri < 2 * dc.cap ? n_tri : 2 * dc.cap;
}

// stage level (kvfe_mesh_2d): one mesh of n given keypoints
__global__ void __launch_bounds__(32) mesh_raw_kernel(DevCfg dc, DevBuf db, const float* __restrict__ x, const float* __restrict__ y,
                                                      int n, float* tri, int max_tri, int* n_tri_out, int use_smem) {
  extern __shared__ __align__(16) int mesh_smem[];
  MeshWs w = mesh_carve(use_smem ? mesh_smem : db.mesh_ws, dc.cap);
  for (int i = threadIdx.x; i < n; i += 32) { w.in_x[i] = x[i]; w.in_y[i] = y[i]; }
  __syncwarp();
  const int n_tri = mesh_build(dc, w, n, tri, max_tri);
  if (threadIdx.x == 0) *n_tri_out = n_tri;
}

static const size_t MESH_SMEM_LIMIT = 200 * 1024;
static size_t mesh_smem_bytes(const DevCfg& dc) { return mesh_ws_ints(dc.cap) * sizeof(int); }
bool mesh_fits_smem(const DevCfg& dc) { return mesh_smem_bytes(dc) <= MESH_SMEM_LIMIT; }
size_t mesh_global_ws_ints(const DevCfg& dc) { return mesh_ws_ints(dc.cap); }

int launch_mesh_init(const DevCfg& dc) {
  if (!mesh_fits_smem(dc)) return 0;
  cudaFuncSetAttribute(mesh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mesh_smem_bytes(dc));
  cudaFuncSetAttribute(mesh_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mesh_smem_bytes(dc));
  return 0;
}
int launch_mesh(const DevCfg& dc, const DevBuf& db, cudaStream_t s) {
  const bool sm = mesh_fits_smem(dc);
  mesh_kernel<<<dc.B, 32, sm ? mesh_smem_bytes(dc) : 0, s>>>(dc, db, sm ? 1 : 0);
  return 1;
}
int launch_mesh_raw(const DevCfg& dc, const DevBuf& db, const float* x, const float* y, int n, float* tri, int max_tri,
                    int* n_tri, cudaStream_t s) {
  const bool sm = mesh_fits_smem(dc);
  mesh_raw_kernel<<<1, 32, sm ? mesh_smem_bytes(dc) : 0, s>>>(dc, db, x, y, n, tri, max_tri, n_tri, sm ? 1 : 0);
  return 1;
}
