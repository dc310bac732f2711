// context.h -- process-wide state of libmitgcm_b200: tile geometry, device mirrors
// of the COMMON-block arrays, run-time parameters, scratch buffers and the stream.
// One process <-> one GPU, like one MPI rank of the reference (eesupp/src/ini_procs.F).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdio>
#include <map>
#include <string>
#include <vector>
#include "../../include/mitgcm_b200.h"

namespace mg {

struct Geom {
  int sNx, sNy, OLx, OLy, nSx, nSy, Nr;
  int nPx, nPy, myPx, myPy;
  int PX, PY, nTiles;
  size_t slab;   // PX*PY
  size_t n2;     // slab*nTiles
  size_t n3;     // slab*Nr*nTiles
};

enum FieldKind { K2D, K3D, K3DP, KJ, KK };

struct Params {
  double d[MP_ND];
  int i[MI_NI_END - 100];
  double D(int id) const { return d[id]; }
  int I(int id) const { return i[id - 100]; }
};

struct Ctx {
  bool ready = false;
  Geom g{};
  int device = 0;
  cudaStream_t stream = nullptr;
  Params p{};
  std::map<int, double *> fields;
  int lastError = 0;
  std::string lastErrorString;
  // staging buffers for host-pointer arguments, keyed by a small slot number
  std::map<int, std::pair<double *, size_t>> stage;
  // halo push tables for width-1 exchanges (cg2d): index of the halo cell that mirrors
  // each edge point, per tile: [W(sNy) | E(sNy) | S(sNx) | N(sNx)]
  int *pushTab = nullptr;
  // pkg/exch2 tile graph (exch2.cu): gather list (dst, src) pairs of the full-width scalar exchange
  int *e2List = nullptr;
  int e2Count = 0;
  int *e2UvList[2] = {nullptr, nullptr};   // vector-pair exchange, [withSigns]: 4 ints per entry
  int e2UvCount[2] = {0, 0};
  // cubed sphere: facet corners each local tile owns (1 SW, 2 SE, 4 NE, 8 NW) and its facet number
  std::vector<int> csCorners, csFace, csEdges;    // csEdges: 1 N | 2 S | 4 E | 8 W facet edges the tile touches
  // exch2 tile graph across ranks: W2_tileProc - 1 per tile id (empty: one process holds every tile); the gather
  // lists then carry the owner's rank next to the source index and read it through the peer arenas
  std::vector<int> e2Owner;
  unsigned long long e2Seq = 0;     // rank barriers of the distributed exchange so far (same on every rank)
  // Peer arena (multi-rank runs only): ONE allocation per rank that holds everything the neighbouring GPUs
  // write into -- the exchanged state fields (halo pushes), the CG2D workspace block (edge pushes, mailboxes) and
  // the exchange flags -- so one CUDA IPC mapping per peer serves all of it (halo.cu).
  char *arena = nullptr;
  size_t arenaBytes = 0, arenaUsed = 0;
  char *peerArena[8] = {};          // every rank's arena as mapped here (own arena for my rank)
  long long arenaDelta[8] = {};     // peerArena[r] - arena: a mirror in the arena sits at the same offset on every rank
  int nRanks = 1, myRank = 0;
  struct HaloWs *halo = nullptr;
  // function attributes (dynamic shared memory opt-in) are per device: set once per init
  bool attrDyn = false, attrThermo = false, attrVi = false, attrDynTma = false; int attrDynTmaUV = 0;
  // column geometry (colgeom.cu): the nine 3-D open-fraction / mask arrays compressed to (kLow, hLow, 1/hLow) per
  // column and point type when they have the z-level form; state 0 = not checked since the arrays last changed,
  // 1 = valid, -1 = the arrays do not have that form (general kernels)
  int *cgK[3] = {nullptr, nullptr, nullptr};            // C, W, S: deepest wet level (0 = land)
  double *cgH[3] = {nullptr, nullptr, nullptr}, *cgR[3] = {nullptr, nullptr, nullptr};   // hFac and recip_hFac at that level
  int cgState = 0, cgFails = 0;
  int *cgFlag = nullptr;
  std::vector<void *> pinned;      // host arrays page-locked by mitgcm_b200_pin_host_
  // cg2d workspace
  struct Cg2dWs *cg2d = nullptr;
  int numSMs = 0;
  long long launches = 0;
  cudaEvent_t ev[16] = {};
  cudaEvent_t pev[8] = {};
  double stepMs[7] = {};
};

Ctx &ctx();
bool col_geom_ready();      // colgeom.cu: (re)builds the column geometry when needed; true = valid
void col_geom_free();
void col_geom_touch(int id);  // a geometry mirror may have changed
bool fail(int code, const std::string &msg);          // records error, returns false
#define MG_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      mg::fail(1000 + (int)e_, std::string(#call) + ": " + cudaGetErrorString(e_)); \
      return false;                                                                \
    }                                                                              \
  } while (0)

size_t field_elems(const Geom &g, int id);             // 0 if id is unknown
double *field(int id, bool create = true);             // device mirror (zero-filled on creation)
bool is_device_ptr(const void *p);
// returns a device pointer for `p`: p itself if it is a device pointer, else a staging
// buffer (slot) of n doubles, filled from p when `upload`.
double *to_device(const double *p, size_t n, int slot, bool upload);
bool from_device(double *hostOrDev, const double *dev, size_t n);  // no-op when same pointer
bool exch2_active();                                   // a pkg/exch2 topology has been set
bool exch2_field(double *f, int nz);                   // EXCH2_3D_RX as one gather
bool exch2_uv_field(double *u, double *v, int nz, bool withSigns);   // EXCH2_UV_3D_RX as one gather
bool exch_field(double *f, int nz);
void *arena_alloc(size_t bytes);                       // from the peer arena (zero-filled); nullptr when there is none / full
inline bool in_arena(const void *p) {
  const Ctx &c = ctx();
  return c.arena && (const char *)p >= c.arena && (const char *)p < c.arena + c.arenaBytes;
}
bool halo_connected();                                 // halo.cu: peers mapped, exchanges go through peer pushes
bool halo_exchange(const int *ids, int n, bool sideStream = false);   // EXCH_XY(Z)_RL of several mirrors across ranks
bool halo_join();                                      // main stream waits for a side-stream exchange
void halo_free();
bool halo_check_error();
int *halo_error_word();                                // device view of the mapped word a timed-out spin writes (nullptr: not connected)
bool cg2d_comm_rank_slots();                           // cg2d.cu: push-table peer slots name ranks (exch2 tile graph across ranks)
void cg3d_free_workspace();                            // cg3d.cu                    // EXCH_XY(Z)_RL on a mirror (step.cu)

}  // namespace mg
