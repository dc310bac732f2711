// halo.cu -- EXCH_XY_RL / EXCH_XYZ_RL across GPUs as peer-memory pushes (eesupp/src/exch1_rx.template:170-201,
// do_fields_blocking_exchanges.F:54-66), and the CUDA IPC wiring of the peer arenas.
//
// One process per GPU, one tile per rank, ranks on a periodic nPx x nPy grid inside one NVSwitch domain.  Every rank
// keeps the fields its neighbours write into -- the exchanged state, the CG2D workspace block -- in ONE allocation
// (the peer arena, context.cu), exports it as a CUDA IPC handle and maps the arenas of its peers.  An exchange is
// then three small launches on the library stream, with no host synchronisation and no staging buffers:
//   ready : tell the 8 neighbours "my kernels that read these halos are done", wait for the same from them
//   push  : store my edge strips AND corner blocks straight into the halo cells of the 8 neighbours over NVLink
//           (the reference fills corners by doing X before Y; the diagonal neighbour's interior block is the same
//           data, so one phase suffices), system-scope fence, the last CTA raises the "done" flags at the neighbours
//   wait  : spin until all 8 neighbours have raised mine
//   unpack: (only with remote west / east / diagonal neighbours) column strips and corner blocks are 16-byte
//           pieces 16 KB apart in the halo -- stored directly over NVLink they cost 1.5 ms per exchange at 2048^2 x
//           150 levels -- so those travel packed into the receiver's strip buffer (contiguous remote stores) and
//           are scattered into the halo locally; south / north rows are contiguous and are stored directly
// Several fields travel in one exchange (one set of flags).  A direction whose neighbour is this rank (nPx or
// nPy = 1) is the local periodic wrap through the same code path.  An exchange can run on a side stream
// (channel 1: own flags) while the main stream computes: theta's halo travels while DYNAMICS runs.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include "context.h"

namespace mg {

constexpr int HALO_MAXF = 8;
#ifndef HALO_SPIN_LIMIT
#define HALO_SPIN_LIMIT (1LL << 31)
#endif

// arena header layout (bytes): per channel c (0 main, 1 side) at 1024*c:
//   +0   ready[8]  (u64)   +64  done[8] (u64)   +128 last-CTA counter (u32)
struct HaloWs {
  int nbrRank[8] = {};
  long long peerDelta[8] = {};        // byte offset from my arena to the mapping of the neighbour's arena, per direction
  unsigned long long seq[2] = {0, 0};
  cudaStream_t side = nullptr;
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  bool sidePending = false;
  int *errHost = nullptr, *errDev = nullptr;   // mapped pinned word: a timed-out spin reports without a sync
  bool connected = false;
  bool gridMode = true;       // one tile per rank on the periodic nPx x nPy grid; false: several tiles per rank (exch2 tile graph)
  void *ipcBase[8] = {};      // what cudaIpcOpenMemHandle returned per rank (the peer's whole allocation)
  double *strip[2] = {nullptr, nullptr};   // per channel: receive buffer of the packed W / E / corner pieces (in the arena)
  long stripLevels = 0;                    // levels a strip buffer holds
};

struct HaloArgs {
  double *f[HALO_MAXF];
  int nz[HALO_MAXF];
  int nf, totalNz;
  int sNx, sNy, OLx, OLy, PX;
  size_t slab;
  long long peerDelta[8];
  unsigned long long seq;
  unsigned long long *ready, *done;   // my flags (this channel)
  unsigned int *counter;
  int *err;
  double *strip;                      // my strip buffer (this channel); the peer's is at the same arena offset
};

// direction d: 0 W, 1 E, 2 S, 3 N, 4 SW, 5 SE, 6 NW, 7 NE
__host__ __device__ inline int dir_dx(int d) { return d == 0 || d == 4 || d == 6 ? -1 : (d == 1 || d == 5 || d == 7 ? 1 : 0); }
__host__ __device__ inline int dir_dy(int d) { return d == 2 || d == 4 || d == 5 ? -1 : (d == 3 || d == 6 || d == 7 ? 1 : 0); }
__host__ __device__ inline int dir_opp(int d) { return d < 4 ? (d ^ 1) : 11 - d; }

__global__ void halo_ready_kernel(HaloArgs a) {
  const int d = threadIdx.x;
  if (d >= 8) return;
  // my earlier kernels (stream order) no longer read these halos: tell the neighbour that will overwrite side d ...
  unsigned long long *peer = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(a.ready) + a.peerDelta[d]);
  *reinterpret_cast<volatile unsigned long long *>(peer + dir_opp(d)) = a.seq;
  // ... and wait until the neighbour on side d says the same about the halo I am going to write there
  long long spins = 0;
  while (*reinterpret_cast<volatile unsigned long long *>(a.ready + d) < a.seq)
    if (++spins > HALO_SPIN_LIMIT) { *a.err = 81; break; }
}

// Strip buffer layout for an exchange of L levels: [side W | side E | SW | SE | NW | NE], side = the receiver's halo
// side; inside a side [level][row][column].
__device__ __forceinline__ long strip_base(const HaloArgs &a, int side, long n0, long n4) {
  return side < 2 ? side * n0 * a.totalNz : 2 * n0 * a.totalNz + (long)(side - 4) * n4 * a.totalNz;
}

__global__ void __launch_bounds__(256) halo_push_kernel(HaloArgs a) {
  const int w0 = a.OLx, h0 = a.sNy, w2 = a.sNx, h2 = a.OLy;
  // cells per level: W | E (OLx x sNy each), S | N (sNx x OLy each), 4 corners (OLx x OLy each)
  const long n0 = (long)w0 * h0, n2 = (long)w2 * h2, n4 = (long)a.OLx * a.OLy;
  const long perLevel = 2 * n0 + 2 * n2 + 4 * n4;
  const long total = perLevel * a.totalNz;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long levG = t / perLevel;
    long lev = levG, c = t - levG * perLevel;
    int fi = 0;
    while (lev >= a.nz[fi]) { lev -= a.nz[fi]; fi++; }
    int d, w;
    if (c < 2 * n0) { d = c >= n0; c -= d * n0; w = w0; }
    else if ((c -= 2 * n0) < 2 * n2) { d = 2 + (c >= n2); c -= (d - 2) * n2; w = w2; }
    else { c -= 2 * n2; d = 4 + (int)(c / n4); c -= (d - 4) * n4; w = a.OLx; }
    const int bI = (int)(c % w), bJ = (int)(c / w);
    const int dx = dir_dx(d), dy = dir_dy(d);
    // source: my interior strip next to edge d (Fortran indices); destination: the same cells shifted by one tile
    const int i = (dx > 0 ? a.sNx - a.OLx + 1 : 1) + bI, j = (dy > 0 ? a.sNy - a.OLy + 1 : 1) + bJ;
    const size_t src = (size_t)(i + a.OLx - 1) + (size_t)a.PX * (size_t)(j + a.OLy - 1) + a.slab * (size_t)lev;
    const double v = a.f[fi][src];
    if (dx != 0 && a.peerDelta[d] != 0) {      // remote column strip / corner: packed into the receiver's strip buffer
      const int side = dir_opp(d);
      double *sb = reinterpret_cast<double *>(reinterpret_cast<char *>(a.strip) + a.peerDelta[d]);
      sb[strip_base(a, side, n0, n4) + levG * (side < 2 ? n0 : n4) + c] = v;
    } else {
      const int iD = i - dx * a.sNx, jD = j - dy * a.sNy;
      const size_t dst = (size_t)(iD + a.OLx - 1) + (size_t)a.PX * (size_t)(jD + a.OLy - 1) + a.slab * (size_t)lev;
      double *fd = reinterpret_cast<double *>(reinterpret_cast<char *>(a.f[fi]) + a.peerDelta[d]);
      fd[dst] = v;
    }
  }
  // publish: every thread's stores are ordered before this CTA's arrival; the last CTA to arrive raises the flags
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    unsigned int n = atomicAdd(a.counter, 1u);
    last = n == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < 8) {
      const int d = threadIdx.x;
      unsigned long long *peer = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(a.done) + a.peerDelta[d]);
      *reinterpret_cast<volatile unsigned long long *>(peer + dir_opp(d)) = a.seq;
    }
    if (threadIdx.x == 0) *a.counter = 0;
  }
}

__global__ void halo_wait_kernel(HaloArgs a) {
  const int d = threadIdx.x;
  if (d < 8) {
    long long spins = 0;
    while (*reinterpret_cast<volatile unsigned long long *>(a.done + d) < a.seq)
      if (++spins > HALO_SPIN_LIMIT) { *a.err = 82; break; }
  }
  __threadfence_system();
}

// strip buffer -> my west / east / corner halos (sides whose neighbour is another rank)
__global__ void __launch_bounds__(256) halo_unpack_kernel(HaloArgs a) {
  const long n0 = (long)a.OLx * a.sNy, n4 = (long)a.OLx * a.OLy;
  const long perSideSet = 2 * n0 + 4 * n4;
  const long total = perSideSet * a.totalNz;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    // t enumerates the buffer in its own order: side, level, cell
    int side;
    long r = t, cells;
    if (r < 2 * n0 * a.totalNz) { side = (int)(r / (n0 * a.totalNz)); r -= side * n0 * a.totalNz; cells = n0; }
    else { r -= 2 * n0 * a.totalNz; side = 4 + (int)(r / (n4 * a.totalNz)); r -= (long)(side - 4) * n4 * a.totalNz; cells = n4; }
    if (a.peerDelta[side] == 0) continue;      // that neighbour is this rank: it wrote the halo directly
    const long levG = r / cells, c = r - levG * cells;
    long lev = levG;
    int fi = 0;
    while (lev >= a.nz[fi]) { lev -= a.nz[fi]; fi++; }
    const int bI = (int)(c % a.OLx), bJ = (int)(c / a.OLx);
    const int dx = dir_dx(side), dy = dir_dy(side);
    const int iD = (dx < 0 ? 1 - a.OLx : a.sNx + 1) + bI;
    const int jD = dy == 0 ? 1 + bJ : (dy < 0 ? 1 - a.OLy : a.sNy + 1) + bJ;
    a.f[fi][(size_t)(iD + a.OLx - 1) + (size_t)a.PX * (size_t)(jD + a.OLy - 1) + a.slab * (size_t)lev] = a.strip[t];
  }
}

bool halo_connected() { return ctx().halo && ctx().halo->connected; }

void halo_free() {
  Ctx &c = ctx();
  HaloWs *h = c.halo;
  if (!h) return;
  if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
  if (h->evFork) cudaEventDestroy(h->evFork);
  if (h->evJoin) cudaEventDestroy(h->evJoin);
  for (int r = 0; r < 8; r++)
    if (h->ipcBase[r]) cudaIpcCloseMemHandle(h->ipcBase[r]);
  if (h->errHost) cudaFreeHost(h->errHost);
  delete h;
  c.halo = nullptr;
}

int *halo_error_word() { return ctx().halo ? ctx().halo->errDev : nullptr; }

bool halo_check_error() {
  HaloWs *h = ctx().halo;
  if (h && h->errHost && *reinterpret_cast<volatile int *>(h->errHost)) {
    int e = *h->errHost;
    *h->errHost = 0;
    return fail(e, "halo / exch2 exchange: timed out waiting for another rank (re-connect to recover)");
  }
  return true;
}

// Offset of p inside its cudaMalloc allocation: the IPC handle names the whole allocation, and small arenas are
// sub-allocated by the runtime.  cuMemGetAddressRange through the runtime's driver entry point (no -lcuda).
static bool alloc_offset(const void *p, unsigned long long *off) {
  typedef int (*PFN)(unsigned long long *, size_t *, unsigned long long);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
    return fail(72, "cudaGetDriverEntryPoint(cuMemGetAddressRange) failed");
  unsigned long long base = 0;
  size_t size = 0;
  if (reinterpret_cast<PFN>(fn)(&base, &size, (unsigned long long)(uintptr_t)p) != 0) return fail(72, "cuMemGetAddressRange failed");
  *off = (unsigned long long)(uintptr_t)p - base;
  return true;
}

static bool comm_handle(unsigned char *handle72) {
  Ctx &c = ctx();
  if (!c.ready) return fail(30, "mitgcm_b200_init_ not called");
  if (!c.arena) return fail(70, "comm_handle: single-rank context (nPx * nPy = 1) has no peer arena");
  cudaIpcMemHandle_t h;
  MG_CUDA(cudaIpcGetMemHandle(&h, c.arena));
  static_assert(sizeof(h) == 64, "IPC handle size");
  unsigned long long off = 0;
  if (!alloc_offset(c.arena, &off)) return false;
  memcpy(handle72, &h, 64);
  memcpy(handle72 + 64, &off, 8);
  return true;
}

bool cg2d_comm_wire();     // cg2d.cu

static bool comm_connect(int nRanks, int myRank, const unsigned char *handles) {
  Ctx &c = ctx();
  if (!c.ready) return fail(30, "mitgcm_b200_init_ not called");
  const Geom &g = c.g;
  if (!c.arena) return fail(70, "comm_connect: single-rank context has no peer arena");
  if (nRanks != c.nRanks || nRanks > 8) return fail(70, "comm_connect: nRanks must equal nPx*nPy (<= 8)");
  if (myRank != c.myRank) return fail(70, "comm_connect: rank must be myPx + nPx*myPy");
  if (g.OLx > g.sNx || g.OLy > g.sNy) return fail(70, "comm_connect: overlap wider than the tile");
  MG_CUDA(cudaStreamSynchronize(c.stream));
  halo_free();
  HaloWs *h = new HaloWs();
  c.halo = h;
  // several tiles per rank: the ranks are not a periodic grid of tiles; every exchange then follows the exch2 tile
  // graph (exch2.cu reads the peers' arenas), which mitgcm_b200_set_exch2_topology_ must supply
  h->gridMode = g.nTiles == 1;
  for (int r = 0; r < nRanks; r++) {
    if (r == myRank) { c.peerArena[r] = c.arena; continue; }
    cudaIpcMemHandle_t ih;
    unsigned long long off = 0;
    memcpy(&ih, handles + 72 * (size_t)r, 64);
    memcpy(&off, handles + 72 * (size_t)r + 64, 8);
    void *base = nullptr;
    MG_CUDA(cudaIpcOpenMemHandle(&base, ih, cudaIpcMemLazyEnablePeerAccess));
    h->ipcBase[r] = base;
    c.peerArena[r] = static_cast<char *>(base) + off;
  }
  auto rk = [&](int px, int py) { return ((px % g.nPx) + g.nPx) % g.nPx + g.nPx * (((py % g.nPy) + g.nPy) % g.nPy); };
  for (int d = 0; d < 8; d++) {
    h->nbrRank[d] = rk(g.myPx + dir_dx(d), g.myPy + dir_dy(d));
    h->peerDelta[d] = c.peerArena[h->nbrRank[d]] - c.arena;
  }
  for (int r = 0; r < 8; r++) c.arenaDelta[r] = r < nRanks ? c.peerArena[r] - c.arena : 0;
  c.e2Seq = 0;
  MG_CUDA(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  MG_CUDA(cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming));
  MG_CUDA(cudaEventCreateWithFlags(&h->evJoin, cudaEventDisableTiming));
  MG_CUDA(cudaHostAlloc(&h->errHost, sizeof(int), cudaHostAllocMapped));
  *h->errHost = 0;
  MG_CUDA(cudaHostGetDevicePointer(&h->errDev, h->errHost, 0));
  MG_CUDA(cudaMemset(c.arena, 0, 4096));      // flags start over (the caller barriers before the first exchange)
  h->stripLevels = 4L * g.Nr + 4;
  const size_t stripBytes = (size_t)(2L * g.OLx * g.sNy + 4L * g.OLx * g.OLy) * (size_t)h->stripLevels * sizeof(double);
  for (int ch = 0; ch < 2 && h->gridMode; ch++) {
    h->strip[ch] = static_cast<double *>(arena_alloc(stripBytes));
    if (!h->strip[ch]) return fail(3, "comm_connect: peer arena exhausted (strip buffers)");
  }
  if (!cg2d_comm_wire()) return false;
  h->connected = true;
  return true;
}

bool halo_exchange(const int *ids, int n, bool sideStream) {
  Ctx &c = ctx();
  HaloWs *h = c.halo;
  if (!h || !h->connected) return fail(70, "halo_exchange: peers not connected (mitgcm_b200_comm_connect_)");
  if (!h->gridMode) return fail(70, "halo_exchange: several tiles per rank -- exchanges follow the exch2 tile graph (mitgcm_b200_set_exch2_topology_)");
  if (!halo_check_error()) return false;
  if (n < 1 || n > HALO_MAXF) return fail(70, "halo_exchange: 1..8 fields per exchange");
  const Geom &g = c.g;
  const int ch = sideStream ? 1 : 0;
  HaloArgs a{};
  a.nf = n;
  a.totalNz = 0;
  for (int k = 0; k < n; k++) {
    const int id = ids[k];
    double *f = field(id);
    if (!f) return false;
    if (!in_arena(f)) return fail(70, "halo_exchange: field " + std::to_string(id) + " is not in the peer arena");
    a.f[k] = f;
    a.nz[k] = id < MG_N2D ? 1 : g.Nr;
    a.totalNz += a.nz[k];
  }
  a.sNx = g.sNx; a.sNy = g.sNy; a.OLx = g.OLx; a.OLy = g.OLy; a.PX = g.PX; a.slab = g.slab;
  for (int d = 0; d < 8; d++) a.peerDelta[d] = h->peerDelta[d];
  a.seq = ++h->seq[ch];
  char *hdr = c.arena + 1024 * ch;
  a.ready = reinterpret_cast<unsigned long long *>(hdr);
  a.done = reinterpret_cast<unsigned long long *>(hdr + 64);
  a.counter = reinterpret_cast<unsigned int *>(hdr + 128);
  a.err = h->errDev;
  a.strip = h->strip[ch];
  if (a.totalNz > h->stripLevels) return fail(70, "halo_exchange: more levels than the strip buffers hold (4 Nr + 4)");
  bool remoteX = false;
  for (int d = 0; d < 8; d++) remoteX = remoteX || (dir_dx(d) != 0 && h->peerDelta[d] != 0);
  cudaStream_t st = c.stream;
  if (sideStream) {
    MG_CUDA(cudaEventRecord(h->evFork, c.stream));
    MG_CUDA(cudaStreamWaitEvent(h->side, h->evFork, 0));
    st = h->side;
  }
  const long perLevel = 2L * g.OLx * g.sNy + 2L * g.sNx * g.OLy + 4L * g.OLx * g.OLy;
  const long total = perLevel * a.totalNz;
  int blocks = (int)std::min<long>((total + 255) / 256, (long)c.numSMs * 4);
  if (blocks < 1) blocks = 1;
  c.launches += 3;
  halo_ready_kernel<<<1, 32, 0, st>>>(a);
  halo_push_kernel<<<blocks, 256, 0, st>>>(a);
  halo_wait_kernel<<<1, 32, 0, st>>>(a);
  if (remoteX) {
    const long nu = (2L * g.OLx * g.sNy + 4L * g.OLx * g.OLy) * a.totalNz;
    c.launches++;
    halo_unpack_kernel<<<(int)std::min<long>((nu + 255) / 256, (long)c.numSMs * 4), 256, 0, st>>>(a);
  }
  MG_CUDA(cudaGetLastError());
  if (sideStream) {
    MG_CUDA(cudaEventRecord(h->evJoin, h->side));
    h->sidePending = true;
  }
  return true;
}

bool halo_join() {
  Ctx &c = ctx();
  HaloWs *h = c.halo;
  if (!h || !h->sidePending) return true;
  MG_CUDA(cudaStreamWaitEvent(c.stream, h->evJoin, 0));
  h->sidePending = false;
  return true;
}

}  // namespace mg

using namespace mg;

extern "C" {

void mitgcm_b200_comm_handle_(unsigned char *handle72, int *ierr) {
  ctx().lastError = 0;
  *ierr = comm_handle(handle72) ? 0 : 1;
}

void mitgcm_b200_comm_connect_(const int *nRanks, const int *myRank, const unsigned char *handles72, int *ierr) {
  ctx().lastError = 0;
  *ierr = comm_connect(*nRanks, *myRank, handles72) ? 0 : 1;
}

void mitgcm_b200_comm_disconnect_(void) {
  Ctx &c = ctx();
  if (!c.ready) return;
  cudaStreamSynchronize(c.stream);
  halo_free();
  for (int r = 0; r < 8; r++)
    if (r != c.myRank) c.peerArena[r] = nullptr;
}

void mitgcm_b200_halo_exchange_(const int *nFields, const int *ids, int *ierr) {
  ctx().lastError = 0;
  *ierr = (halo_exchange(ids, *nFields, false)) ? 0 : 1;
}

}  // extern "C"
