// context.cu -- life cycle, parameters and device mirrors (C ABI: include/mitgcm_b200.h).
#include <cstdlib>
#include <cstring>
#include "context.h"

namespace mg {

static Ctx g_ctx;
Ctx &ctx() { return g_ctx; }

bool fail(int code, const std::string &msg) {
  g_ctx.lastError = code;
  g_ctx.lastErrorString = msg;
  fprintf(stderr, "mitgcm_b200: error %d: %s\n", code, msg.c_str());
  return false;
}

size_t field_elems(const Geom &g, int id) {
  if (id >= 0 && id < MG_N2D) return g.n2;
  if (id >= 100 && id < MG_N3D_END) return g.n3;
  if (id >= 200 && id < MG_N3DP_END) return g.slab * (size_t)(g.Nr + 1) * g.nTiles;
  if (id >= 300 && id < MG_NJ_END) return (size_t)g.PY * g.nTiles;
  if (id >= 400 && id < MG_NK_END) return (size_t)g.Nr + 1;
  return 0;
}

// DO_FIELDS_BLOCKING_EXCHANGES / SOLVE_FOR_PRESSURE exchange these (do_fields_blocking_exchanges.F:54-66,
// solve_for_pressure.F:316); theta2 / salt2 because CYCLE_TRACER swaps the mirrors.
static bool is_exchanged_field(int id) {
  switch (id) {
    case MG_UVEL: case MG_VVEL: case MG_WVEL: case MG_THETA: case MG_THETA2: case MG_SALT: case MG_SALT2:
    case MG_CG2D_X: case MG_ETAN: case MG_ETAH: case MG_AC2D:
      return true;
    default:
      return false;
  }
}

double *field(int id, bool create) {
  Ctx &c = g_ctx;
  auto it = c.fields.find(id);
  if (it != c.fields.end()) return it->second;
  if (!create) return nullptr;
  size_t n = field_elems(c.g, id);
  if (n == 0) { fail(2, "unknown field id " + std::to_string(id)); return nullptr; }
  double *p = nullptr;
  // fields whose halos the neighbouring ranks fill live in the peer arena (zero-filled at creation)
  if (c.arena && is_exchanged_field(id)) p = static_cast<double *>(arena_alloc(n * sizeof(double)));
  if (!p) {
    if (cudaMalloc(&p, n * sizeof(double)) != cudaSuccess) { fail(3, "cudaMalloc failed for field"); return nullptr; }
    cudaMemsetAsync(p, 0, n * sizeof(double), c.stream);
  }
  c.fields[id] = p;
  return p;
}

void *arena_alloc(size_t bytes) {
  Ctx &c = g_ctx;
  if (!c.arena) return nullptr;
  size_t off = (c.arenaUsed + 255) & ~(size_t)255;
  if (off + bytes > c.arenaBytes) return nullptr;
  c.arenaUsed = off + bytes;
  return c.arena + off;
}

bool is_device_ptr(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

double *to_device(const double *p, size_t n, int slot, bool upload) {
  Ctx &c = g_ctx;
  if (is_device_ptr(p)) return const_cast<double *>(p);
  auto &s = c.stage[slot];
  if (s.second < n) {
    if (s.first) cudaFree(s.first);
    s.first = nullptr;
    if (cudaMalloc(&s.first, n * sizeof(double)) != cudaSuccess) { fail(3, "cudaMalloc failed for staging"); return nullptr; }
    s.second = n;
  }
  if (upload && cudaMemcpyAsync(s.first, p, n * sizeof(double), cudaMemcpyHostToDevice, c.stream) != cudaSuccess) {
    fail(4, "H2D copy failed");
    return nullptr;
  }
  return s.first;
}

bool from_device(double *dst, const double *dev, size_t n) {
  if (dst == dev) return true;
  cudaMemcpyKind kind = is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  MG_CUDA(cudaMemcpyAsync(dst, dev, n * sizeof(double), kind, g_ctx.stream));
  return true;
}

// Halo "push" tables for a periodic nSx x nSy tiling on one process (EXCH1 topology,
// eesupp/src/exch_rx_send_put_{x,y}.template with COMM_PUT faces): the west edge (1,j) of a
// tile is mirrored in the east halo (sNx+1,j) of its west neighbour, and so on.
static bool build_push_tables() {
  Ctx &c = g_ctx;
  const Geom &g = c.g;
  const int per = 2 * g.sNy + 2 * g.sNx;
  std::vector<int> t((size_t)per * g.nTiles);
  auto idx = [&](int i, int j, int tile) {   // Fortran (i,j) -> flat index in a tile2d array
    return (int)((size_t)(i + g.OLx - 1) + (size_t)g.PX * (size_t)(j + g.OLy - 1) + g.slab * (size_t)tile);
  };
  for (int bj = 0; bj < g.nSy; bj++)
    for (int bi = 0; bi < g.nSx; bi++) {
      int tile = bi + g.nSx * bj;
      int tw = (bi + g.nSx - 1) % g.nSx + g.nSx * bj, te = (bi + 1) % g.nSx + g.nSx * bj;
      int ts = bi + g.nSx * ((bj + g.nSy - 1) % g.nSy), tn = bi + g.nSx * ((bj + 1) % g.nSy);
      int *p = &t[(size_t)per * tile];
      for (int j = 1; j <= g.sNy; j++) {
        p[j - 1] = idx(g.sNx + 1, j, tw);
        p[g.sNy + j - 1] = idx(0, j, te);
      }
      for (int i = 1; i <= g.sNx; i++) {
        p[2 * g.sNy + i - 1] = idx(i, g.sNy + 1, ts);
        p[2 * g.sNy + g.sNx + i - 1] = idx(i, 0, tn);
      }
    }
  MG_CUDA(cudaMalloc(&c.pushTab, t.size() * sizeof(int)));
  MG_CUDA(cudaMemcpy(c.pushTab, t.data(), t.size() * sizeof(int), cudaMemcpyHostToDevice));
  return true;
}

void cg2d_free_workspace();

}  // namespace mg

using namespace mg;

extern "C" {

void mitgcm_b200_init_(const int *dims, const int *device, int *ierr) {
  Ctx &c = ctx();
  if (c.ready) mitgcm_b200_finalize_();
  c.lastError = 0;
  Geom &g = c.g;
  g.sNx = dims[0]; g.sNy = dims[1]; g.OLx = dims[2]; g.OLy = dims[3];
  g.nSx = dims[4]; g.nSy = dims[5]; g.Nr = dims[6];
  g.nPx = dims[7]; g.nPy = dims[8]; g.myPx = dims[9]; g.myPy = dims[10];
  *ierr = 1;
  if (g.sNx < 1 || g.sNy < 1 || g.OLx < 1 || g.OLy < 1 || g.nSx < 1 || g.nSy < 1 || g.Nr < 1) {
    fail(10, "bad dims");
    return;
  }
  g.PX = g.sNx + 2 * g.OLx; g.PY = g.sNy + 2 * g.OLy; g.nTiles = g.nSx * g.nSy;
  g.slab = (size_t)g.PX * g.PY; g.n2 = g.slab * g.nTiles; g.n3 = g.n2 * g.Nr;
  // the width-1 push tables (cg2d.cu, cg3d.cu, exch2.cu) keep the flat halo index in 28 bits next to the peer slot
  if (g.n2 >= (size_t)1 << 28) { fail(10, "tile2d array exceeds 2^28 elements (push-table index range)"); return; }
  if (g.nPx < 1 || g.nPy < 1 || g.myPx < 0 || g.myPx >= g.nPx || g.myPy < 0 || g.myPy >= g.nPy) { fail(10, "bad process grid"); return; }
  int dev = *device;
  if (dev < 0) {
    const char *lr = getenv("LOCAL_RANK");
    dev = lr ? atoi(lr) : 0;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fail(11, "no CUDA device: libmitgcm_b200 has no CPU fallback");
    return;
  }
  if (dev >= ndev) {      // never wrap: two ranks on one GPU cannot co-run their spin-waiting cooperative kernels
    fail(11, "device ordinal " + std::to_string(dev) + " out of range (" + std::to_string(ndev) + " visible)");
    return;
  }
  if (cudaSetDevice(dev) != cudaSuccess) { fail(11, "cudaSetDevice failed"); return; }
  c.device = dev;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, c.device);
  c.numSMs = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) != cudaSuccess) { fail(12, "stream"); return; }
  memset(&c.p, 0, sizeof(c.p));
  // defaults of model/src/set_defaults.F for the parameters on the path
  c.p.d[MP_RKSIGN] = -1.0; c.p.d[MP_FREESURFFAC] = 1.0; c.p.d[MP_IMPLICSURFPRESS] = 1.0;
  c.p.d[MP_IMPLICDIV2DFLOW] = 1.0; c.p.d[MP_SIDEDRAGFACTOR] = 2.0;
  c.p.d[MP_AFFACMOM] = c.p.d[MP_VFFACMOM] = c.p.d[MP_CFFACMOM] = c.p.d[MP_MTFACMOM] = 1.0;
  c.p.d[MP_ABEPS] = 0.01; c.p.d[MP_RECIP_RSPHERE] = 1.0 / 6370.0e3; c.p.d[MP_CG2DPCOFFDFAC] = 0.51;
  c.p.i[MI_CG2DPRECONDFREQ - 100] = 1;
  c.p.i[MI_CG2DNORMALISERHS - 100] = 1; c.p.i[MI_CG2DMAXITERS - 100] = 150;
  c.p.i[MI_MOMADVECTION - 100] = 1; c.p.i[MI_MOMVISCOSITY - 100] = 1;
  c.p.i[MI_SELECTBOTDRAGQUADR - 100] = -1; c.p.i[MI_MOMFORCING - 100] = 1;
  c.p.i[MI_MOMDISSIP_IN_AB - 100] = 1; c.p.i[MI_TEMPADVSCHEME - 100] = 2;
  c.p.i[MI_TEMPVERTADVSCHEME - 100] = 2;
  c.p.i[MI_USECORIOLIS - 100] = 1; c.p.i[MI_SELECTVORTSCHEME - 100] = 1;
  c.p.i[MI_MULTIDIMADVECTION - 100] = 1;
  c.p.i[MI_SALTADVSCHEME - 100] = 2; c.p.i[MI_SALTVERTADVSCHEME - 100] = 2;
  if (!build_push_tables()) return;
  c.nRanks = g.nPx * g.nPy;
  c.myRank = g.myPx + g.nPx * g.myPy;
  c.e2Owner.clear();
  c.e2Seq = 0;
  for (int r = 0; r < 8; r++) c.arenaDelta[r] = 0;
  c.attrDyn = c.attrThermo = c.attrVi = c.attrDynTma = false; c.attrDynTmaUV = 0;
  if (c.nRanks > 1) {
    // peer arena: header (flags) + 7 tile3d + 4 tile2d exchanged fields + the CG2D workspace block (9 tile2d + mailboxes)
    // + the two strip buffers of halo.cu: (2 OLx sNy + 4 OLx OLy) x (4 Nr + 4) doubles each
    c.arenaBytes = 65536 + (7 * g.n3 + 4 * g.n2 + 9 * g.n2) * sizeof(double) + 32 * 256 +
                   2 * (size_t)(2L * g.OLx * g.sNy + 4L * g.OLx * g.OLy) * (size_t)(4L * g.Nr + 4) * sizeof(double);
    if (cudaMalloc(&c.arena, c.arenaBytes) != cudaSuccess) { c.arena = nullptr; fail(3, "cudaMalloc failed for the peer arena"); return; }
    if (cudaMemset(c.arena, 0, c.arenaBytes) != cudaSuccess) { fail(3, "peer arena memset"); return; }
    c.arenaUsed = 4096;      // header: exchange flags (halo.cu)
    c.peerArena[c.myRank] = c.arena;
  }
  c.ready = true;
  *ierr = 0;
}

void mitgcm_b200_finalize_(void) {
  Ctx &c = ctx();
  if (!c.ready) return;
  cudaStreamSynchronize(c.stream);
  for (auto &f : c.fields)
    if (!in_arena(f.second)) cudaFree(f.second);
  c.fields.clear();
  for (auto &s : c.stage) cudaFree(s.second.first);
  c.stage.clear();
  if (c.pushTab) cudaFree(c.pushTab);
  c.pushTab = nullptr;
  if (c.e2List) cudaFree(c.e2List);
  c.e2List = nullptr;
  c.csCorners.clear(); c.csFace.clear(); c.csEdges.clear();
  cg3d_free_workspace();
  c.e2Count = 0;
  for (int w = 0; w < 2; w++) {
    if (c.e2UvList[w]) cudaFree(c.e2UvList[w]);
    c.e2UvList[w] = nullptr;
    c.e2UvCount[w] = 0;
  }
  for (void *p : c.pinned) cudaHostUnregister(p);
  c.pinned.clear();
  cg2d_free_workspace();
  col_geom_free();
  halo_free();               // unmaps the peers' arenas
  if (c.arena) cudaFree(c.arena);
  c.arena = nullptr;
  c.arenaBytes = c.arenaUsed = 0;
  for (auto &pa : c.peerArena) pa = nullptr;
  c.nRanks = 1; c.myRank = 0;
  for (auto &e : c.ev) if (e) { cudaEventDestroy(e); e = nullptr; }
  for (auto &e : c.pev) if (e) { cudaEventDestroy(e); e = nullptr; }
  c.launches = 0;
  cudaStreamDestroy(c.stream);
  c.stream = nullptr;
  c.ready = false;
}

int mitgcm_b200_last_error_(void) { return ctx().lastError; }
const char *mitgcm_b200_last_error_string(void) { return ctx().lastErrorString.c_str(); }

void mitgcm_b200_set_param_d_(const int *id, const double *val, int *ierr) {
  if (*id < 0 || *id >= MP_ND) { fail(20, "bad double parameter id"); *ierr = 1; return; }
  ctx().p.d[*id] = *val;
  *ierr = 0;
}
void mitgcm_b200_set_param_i_(const int *id, const int *val, int *ierr) {
  if (*id < 100 || *id >= MI_NI_END) { fail(20, "bad int parameter id"); *ierr = 1; return; }
  ctx().p.i[*id - 100] = *val;
  *ierr = 0;
}

void mitgcm_b200_set_field_(const int *id, const double *host, int *ierr) {
  Ctx &c = ctx();
  *ierr = 1;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  double *d = field(*id);
  if (!d) return;
  size_t n = field_elems(c.g, *id);
  cudaMemcpyKind kind = is_device_ptr(host) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (cudaMemcpyAsync(d, host, n * sizeof(double), kind, c.stream) != cudaSuccess) { fail(4, "set_field copy"); return; }
  if (cudaStreamSynchronize(c.stream) != cudaSuccess) { fail(4, "set_field sync"); return; }
  col_geom_touch(*id);
  *ierr = 0;
}

void mitgcm_b200_get_field_(const int *id, double *host, int *ierr) {
  Ctx &c = ctx();
  *ierr = 1;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  double *d = field(*id);
  if (!d) return;
  size_t n = field_elems(c.g, *id);
  cudaMemcpyKind kind = is_device_ptr(host) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (cudaMemcpyAsync(host, d, n * sizeof(double), kind, c.stream) != cudaSuccess) { fail(4, "get_field copy"); return; }
  if (cudaStreamSynchronize(c.stream) != cudaSuccess) { fail(4, "get_field sync"); return; }
  *ierr = 0;
}

__global__ void fill_kernel(double *p, size_t n, double v) {
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) p[t] = v;
}

void mitgcm_b200_fill_field_(const int *id, const double *value, int *ierr) {
  Ctx &c = ctx();
  *ierr = 1;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  double *d = field(*id);
  if (!d) return;
  c.launches++;
  fill_kernel<<<c.numSMs * 8, 256, 0, c.stream>>>(d, field_elems(c.g, *id), *value);
  if (cudaStreamSynchronize(c.stream) != cudaSuccess) { fail(6, "fill_field"); return; }
  col_geom_touch(*id);
  *ierr = 0;
}

void mitgcm_b200_pin_host_(double *host, const long long *nDoubles, int *ierr) {
  Ctx &c = ctx();
  *ierr = 1;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  for (void *p : c.pinned)
    if (p == host) { *ierr = 0; return; }
  if (cudaHostRegister(host, (size_t)*nDoubles * sizeof(double), cudaHostRegisterDefault) != cudaSuccess) {
    cudaGetLastError();
    fail(7, "pin_host: cudaHostRegister failed (the array stays pageable)");
    return;
  }
  c.pinned.push_back(host);
  *ierr = 0;
}

double *mitgcm_b200_field_ptr(int id) {
  if (!ctx().ready) return nullptr;
  col_geom_touch(id);      // the caller may write through the address
  return field(id);
}

void mitgcm_b200_event_record_(const int *slot) {
  Ctx &c = ctx();
  if (!c.ready || *slot < 0 || *slot >= 16) return;
  if (!c.ev[*slot]) cudaEventCreate(&c.ev[*slot]);
  cudaEventRecord(c.ev[*slot], c.stream);
}
void mitgcm_b200_event_elapsed_ms_(const int *a, const int *b, double *ms) {
  Ctx &c = ctx();
  *ms = -1.0;
  if (!c.ready || !c.ev[*a] || !c.ev[*b]) return;
  cudaEventSynchronize(c.ev[*b]);
  float f = 0.f;
  if (cudaEventElapsedTime(&f, c.ev[*a], c.ev[*b]) == cudaSuccess) *ms = f;
}
long long mitgcm_b200_launch_count_(void) { return ctx().launches; }
void mitgcm_b200_step_timings_(double *ms7) {
  // elapsed time between the phase marks recorded by the last step (MI_PROFILE != 0); in a
  // multi-rank step the caller's NCCL exchange of cg2d_x falls into slot 4 and slot 6 is 0
  Ctx &c = ctx();
  for (int i = 0; i < 7; i++) {
    ms7[i] = 0.0;
    if (!c.pev[i] || !c.pev[i + 1]) continue;
    if (cudaEventSynchronize(c.pev[i + 1]) != cudaSuccess) continue;
    float f = 0.f;
    if (cudaEventElapsedTime(&f, c.pev[i], c.pev[i + 1]) == cudaSuccess && f >= 0.f) ms7[i] = f;
  }
  cudaGetLastError();
}

void mitgcm_b200_sync_(void) {
  if (ctx().ready) cudaStreamSynchronize(ctx().stream);
}

}  // extern "C"
