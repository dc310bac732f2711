// exch2.cu -- pkg/exch2 (tile-graph halo exchange, e.g. the cubed sphere) for scalar fields on
// one GPU.  The reference walks, per tile and per neighbour entry, an index range of the halo
// and copies it through an affine index map into buffers, twice (EXCH2_RX1_CUBE with
// EXCH_IGNORE_CORNERS, then with EXCH_UPDATE_CORNERS: pkg/exch2/exch2_3d_rx.template:60-80,
// exch2_rx1_cube.template:95-262, exch2_get_scal_bounds.F:44-126, exch2_put_rx1.template:160-175).
// Here the topology tables of COMMON /W2_EXCH2_TOPO_I/ and /W2_EXCH2_HALO_SPEC/
// (W2_EXCH2_TOPOLOGY.h:63-122) are compiled ONCE into a gather list: for every halo cell the cell
// whose pre-exchange value it holds after both passes (the two passes are composed cell by cell,
// later neighbour entries overriding earlier ones exactly as the sequential GETs do).  An exchange
// of any field is then one launch of a coalesced gather over (list entry, level), and the width-1
// exchange inside CG2D (EXCH2_S3D_RX, exch2_s3d_rx.template:45-62) becomes the solver's push table.
//
// Across GPUs (the reference: exch2_send_rx{1,2}.template / exch2_recv_rx{1,2}.template, MPI messages between
// W2_tileProc(source tile) and W2_tileProc(target tile)): every rank holds nSx*nSy tiles of the graph
// (W2_myTileList) and compiles the SAME global provenance map; its gather list keeps the entries whose destination
// is one of its tiles and tags every source with the rank that owns it (rank << 28 | index in that rank's arrays).
// A source on another rank is read straight out of that rank's peer arena over NVLink (the exchanged mirrors sit
// at the same arena offset on every rank, context.cu), between two rank barriers on the stream: "every rank's
// interior is final" before the gather, "every rank has finished reading" after it.  Sources are interior cells
// and destinations halo cells (checked), so the gathers of different ranks never conflict.  The CG2D push table
// gets the same rank tags; the solver already stores edge values through them (cg2d.cu: pdst).
#include <algorithm>
#include <vector>
#include "context.h"

namespace mg {

struct E2Tables {
  int nT = 0, maxN = 0;
  std::vector<int> nN, nid, opp, ndir, pij, oi, oj, iLo, iHi, jLo, jHi, bx, by;
  std::vector<int> local, owner;         // position of the tile in its owner's (bi,bj) order; owning rank (0-based)
  std::vector<int> isN, isS, isE, isW;   // exch2_isNedge .. exch2_isWedge
  int N(int t) const { return nN[t]; }
  int at(const std::vector<int> &a, int n, int t) const { return a[n + maxN * t]; }
  int P(int k, int n, int t) const { return pij[k + 4 * (n + maxN * t)]; }
};

static void target_range(const E2Tables &T, int n, int t, int eW, bool corners, int r[4]) {
  int i0 = T.at(T.iLo, n, t), i1 = T.at(T.iHi, n, t), j0 = T.at(T.jLo, n, t), j1 = T.at(T.jHi, n, t);
  const int grow = corners ? eW - 1 : -1;
  // the four tests are sequential on the updated values, as in exch2_get_scal_bounds.F:50-124
  if (i0 == i1 && i0 == 0) { i0 = 1 - eW; j0 -= grow; j1 += grow; }       // west edge overlap
  if (i0 == i1 && i0 > 1) { i1 = i1 + eW - 1; j0 -= grow; j1 += grow; }   // east
  if (j0 == j1 && j0 == 0) { j0 = 1 - eW; i0 -= grow; i1 += grow; }       // south
  if (j0 == j1 && j0 > 1) { j1 = j1 + eW - 1; i0 -= grow; i1 += grow; }   // north
  r[0] = i0; r[1] = i1; r[2] = j0; r[3] = j1;
}

// provenance of every cell of a (nT, sNy+2*OL, sNx+2*OL) array after the exchange
static bool compile_gather(const E2Tables &T, int sNx, int sNy, int OL, int eW, bool twoPass,
                           std::vector<int> &prov) {
  const int PX = sNx + 2 * OL, PY = sNy + 2 * OL;
  auto flat = [&](int t, int i, int j) { return (t * PY + (j + OL - 1)) * PX + (i + OL - 1); };
  prov.resize((size_t)T.nT * PX * PY);
  for (size_t q = 0; q < prov.size(); q++) prov[q] = (int)q;
  for (int pass = 0; pass < (twoPass ? 2 : 1); pass++) {
    std::vector<int> nw(prov);
    for (int t = 0; t < T.nT; t++)
      for (int n = 0; n < T.N(t); n++) {
        const int s = T.at(T.nid, n, t) - 1, m = T.at(T.opp, n, t) - 1;
        int r[4];
        target_range(T, n, t, eW, pass == 1, r);
        for (int j = r[2]; j <= r[3]; j++)
          for (int i = r[0]; i <= r[1]; i++) {
            const int ic = i + T.bx[t], jc = j + T.by[t];
            const int si = T.P(0, m, s) * ic + T.P(1, m, s) * jc + T.at(T.oi, m, s) - T.bx[s];
            const int sj = T.P(2, m, s) * ic + T.P(3, m, s) * jc + T.at(T.oj, m, s) - T.by[s];
            if (si < 1 - OL || si > sNx + OL || sj < 1 - OL || sj > sNy + OL || i < 1 - OL || i > sNx + OL ||
                j < 1 - OL || j > sNy + OL)
              return fail(71, "exch2: index map leaves the tile array (inconsistent topology tables)");
            nw[flat(t, i, j)] = prov[flat(s, si, sj)];
          }
      }
    prov.swap(nw);
  }
  return true;
}

struct PeerDeltas { long long d[8]; };      // byte offset from my arena to rank r's, as mapped here (0 for my rank)

// value of element `idx` of the mirror `f` as rank `r` holds it
__device__ __forceinline__ double peer_load(const double *f, const PeerDeltas &pd, int r, size_t idx) {
  const long long d = pd.d[r];
  const double *p = reinterpret_cast<const double *>(reinterpret_cast<const char *>(f) + d) + idx;
  return d ? __ldcv(p) : *p;      // peer memory: never from a stale cache line
}

__global__ void gather_kernel(double *f, const int2 *lst, int n, int nz, size_t slab, PeerDeltas pd) {
  const size_t total = (size_t)n * nz;
  for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(q % n), k = (int)(q / n);
    const int2 ds = lst[e];   // (dst, rank << 28 | src) as tile*slab + cell of a 2-D array; the level offset is added here
    const size_t dt = (size_t)ds.x / slab, dc = (size_t)ds.x % slab;
    const size_t src = (size_t)(ds.y & 0x0FFFFFFF), st = src / slab, sc = src % slab;
    f[dc + slab * (k + (size_t)nz * dt)] = peer_load(f, pd, (ds.y >> 28) & 7, sc + slab * (k + (size_t)nz * st));
  }
}

// Rank barrier on the stream: tell every rank "I am here" (sequence number into its flag word for me), wait until
// every rank has said the same.  Flags: 8 words at arena + 2048 (zeroed by mitgcm_b200_comm_connect_).
#ifndef E2_SPIN_LIMIT
#define E2_SPIN_LIMIT (1LL << 31)
#endif
__global__ void e2_barrier_kernel(unsigned long long *flags, PeerDeltas pd, int nRanks, int myRank, unsigned long long seq, int *err) {
  const int r = threadIdx.x;
  __threadfence_system();
  if (r < nRanks) {
    unsigned long long *peer = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(flags) + pd.d[r]);
    *reinterpret_cast<volatile unsigned long long *>(peer + myRank) = seq;
    long long spins = 0;
    while (*reinterpret_cast<volatile unsigned long long *>(flags + r) < seq)
      if (++spins > E2_SPIN_LIMIT) { if (err) *err = 83; break; }
  }
  __threadfence_system();
}

static bool e2_distributed() { return ctx().nRanks > 1; }

static PeerDeltas e2_deltas() {
  PeerDeltas pd;
  for (int r = 0; r < 8; r++) pd.d[r] = ctx().arenaDelta[r];
  return pd;
}

static bool e2_barrier() {
  Ctx &c = ctx();
  if (!halo_connected()) return fail(70, "exch2 across ranks: peers not connected (mitgcm_b200_comm_connect_)");
  if (!halo_check_error()) return false;
  c.launches++;
  e2_barrier_kernel<<<1, 32, 0, c.stream>>>(reinterpret_cast<unsigned long long *>(c.arena + 2048), e2_deltas(), c.nRanks, c.myRank,
                                          ++c.e2Seq, halo_error_word());
  MG_CUDA(cudaGetLastError());
  return true;
}

// ---- vector pairs on the C grid: EXCH2_UV_3D_RX -------------------------------------------------------
// EXCH2_GET_UV_BOUNDS with fCode 'Cg' (exch2_get_uv_bounds.F:60-262): target ranges of the two components
// through neighbour entry n of tile t and the index offsets of the source entry.
struct UvBounds { int r1[4], r2[4], o[4]; };
static UvBounds uv_bounds(const E2Tables &T, int n, int t, int eW, bool corners) {
  UvBounds b;
  const int tIlo = T.at(T.iLo, n, t), tIhi = T.at(T.iHi, n, t), tJlo = T.at(T.jLo, n, t), tJhi = T.at(T.jHi, n, t);
  const int s = T.at(T.nid, n, t) - 1, m = T.at(T.opp, n, t) - 1;
  int oi1 = T.at(T.oi, m, s), oj1 = T.at(T.oj, m, s);
  const int p0 = T.P(0, m, s), p1 = T.P(1, m, s), p2 = T.P(2, m, s), p3 = T.P(3, m, s);
  int i0 = 0, i1 = 0, j0 = 0, j1 = 0;
  const int grow = corners ? eW - 1 : -1;
  if (tIlo == tIhi && tIlo == 0) { i0 = 1 - eW; i1 = 0; j0 = tJlo - grow; j1 = tJhi + grow; }
  if (tIlo == tIhi && tIlo > 1) { i0 = tIlo; i1 = tIhi + eW - 1; j0 = tJlo - grow; j1 = tJhi + grow; }
  if (tJlo == tJhi && tJlo == 0) { j0 = 1 - eW; j1 = 0; i0 = tIlo - grow; i1 = tIhi + grow; }
  if (tJlo == tJhi && tJlo > 1) { j0 = tJlo; j1 = tJhi + eW - 1; i0 = tIlo - grow; i1 = tIhi + grow; }
  int a[4] = {i0, i1, j0, j1}, c[4] = {i0, i1, j0, j1};
  int oi2 = oi1, oj2 = oj1;
  if (p0 == -1) oi1++;
  if (p2 == -1) oj1++;
  if (p1 == -1) oi2++;
  if (p3 == -1) oj2++;
  if (corners) {
    if (p0 == -1 || p2 == -1) a[0]++;
    if (p1 == -1 || p3 == -1) c[2]++;
    if (tIlo == tIhi && tIlo > 1) {           // east edge of a tile on the face S / N edge
      if (T.isS[t] == 1) { a[2] = tJlo + 1; c[2] = tJlo + 1; }
      if (T.isN[t] == 1) { a[3] = tJhi - 1; c[3] = tJhi; }
    }
    if (tJlo == tJhi && tJlo > 1) {           // north edge of a tile on the face W / E edge
      if (T.isW[t] == 1) { a[0] = tIlo + 1; c[0] = tIlo + 1; }
      if (T.isE[t] == 1) { a[1] = tIhi; c[1] = tIhi - 1; }
    }
  } else {
    if (p0 == -1 || p2 == -1) { a[0]++; a[1]++; }
    if (p1 == -1 || p3 == -1) { c[2]++; c[3]++; }
  }
  for (int q = 0; q < 4; q++) { b.r1[q] = a[q]; b.r2[q] = c[q]; }
  b.o[0] = oi1; b.o[1] = oj1; b.o[2] = oi2; b.o[3] = oj2;
  return b;
}

struct Prov { int arr, cell, sign; };   // which array (0 = u, 1 = v), flat cell, +-1

// Both passes of EXCH2_RX2_CUBE (exch2_rx2_cube.template, exch2_put_rx2.template:150-260: the value put
// for a target cell is sa1*u(src) + sa2*v(src) with (sa1, sa2) = (pij1, pij3) for the first component and
// (pij2, pij4) for the second, absolute values when withSigns is false) and the four cube-corner fix-ups of
// exch2_uv_3d_rx.template:117-230, composed cell by cell into provenances.
static bool compile_uv(const E2Tables &T, int sNx, int sNy, int OL, bool withSigns, std::vector<Prov> pr[2]) {
  const int PX = sNx + 2 * OL, PY = sNy + 2 * OL;
  auto flat = [&](int t, int i, int j) { return (t * PY + (j + OL - 1)) * PX + (i + OL - 1); };
  const size_t n = (size_t)T.nT * PX * PY;
  for (int a = 0; a < 2; a++) {
    pr[a].resize(n);
    for (size_t q = 0; q < n; q++) pr[a][q] = Prov{a, (int)q, 1};
  }
  for (int pass = 0; pass < 2; pass++) {
    std::vector<Prov> nw[2] = {pr[0], pr[1]};
    for (int t = 0; t < T.nT; t++)
      for (int e = 0; e < T.N(t); e++) {
        const int s = T.at(T.nid, e, t) - 1, m = T.at(T.opp, e, t) - 1;
        const UvBounds b = uv_bounds(T, e, t, OL, pass == 1);
        const int p0 = T.P(0, m, s), p1 = T.P(1, m, s), p2 = T.P(2, m, s), p3 = T.P(3, m, s);
        for (int comp = 0; comp < 2; comp++) {
          const int *r = comp == 0 ? b.r1 : b.r2;
          const int oi = b.o[2 * comp], oj = b.o[2 * comp + 1];
          int sa1 = comp == 0 ? p0 : p1, sa2 = comp == 0 ? p2 : p3;
          if (!withSigns) { sa1 = abs(sa1); sa2 = abs(sa2); }
          if ((sa1 != 0) == (sa2 != 0)) return fail(73, "exch2: vector index map is not a signed permutation");
          for (int j = r[2]; j <= r[3]; j++)
            for (int i = r[0]; i <= r[1]; i++) {
              const int ic = i + T.bx[t], jc = j + T.by[t];
              const int si = p0 * ic + p1 * jc + oi - T.bx[s], sj = p2 * ic + p3 * jc + oj - T.by[s];
              if (si < 1 - OL || si > sNx + OL || sj < 1 - OL || sj > sNy + OL || i < 1 - OL || i > sNx + OL ||
                  j < 1 - OL || j > sNy + OL)
                return fail(71, "exch2: vector index map leaves the tile array");
              Prov src = pr[sa1 != 0 ? 0 : 1][flat(s, si, sj)];
              src.sign *= (sa1 != 0 ? sa1 : sa2);
              nw[comp][flat(t, i, j)] = src;
            }
        }
      }
    pr[0].swap(nw[0]);
    pr[1].swap(nw[1]);
  }
  if (OL >= 2) {
    const int sg = withSigns ? -1 : 1;
    auto cp = [&](int da, int t, int di, int dj, int sa, int si, int sj, int sign) {
      Prov v = pr[sa][flat(t, si, sj)];
      v.sign *= sign;
      pr[da][flat(t, di, dj)] = v;
    };
    for (int t = 0; t < T.nT; t++) {
      const bool W = T.isW[t] == 1, E = T.isE[t] == 1, S = T.isS[t] == 1, N = T.isN[t] == 1;
      if (W && S) { cp(0, t, 0, 0, 1, 1, 0, 1); cp(1, t, 0, 0, 0, 0, 1, 1); }
      if (W && N) { cp(0, t, 0, sNy + 1, 1, 1, sNy + 2, sg); cp(1, t, 0, sNy + 2, 0, 0, sNy, sg); }
      if (E && S) { cp(0, t, sNx + 2, 0, 1, sNx, 0, sg); cp(1, t, sNx + 1, 0, 0, sNx + 2, 1, sg); }
      if (E && N) { cp(0, t, sNx + 2, sNy + 1, 1, sNx, sNy + 2, 1); cp(1, t, sNx + 1, sNy + 2, 0, sNx + 2, sNy, 1); }
    }
  }
  return true;
}

// list entries: 4 ints (dst array, dst tile*slab+cell, src array << 1 | (sign < 0), owner rank << 28 | src tile*slab+cell);
// only destinations in tiles of rank `me` (me < 0: all)
static void uv_list(const std::vector<Prov> pr[2], const std::vector<int> &local, const std::vector<int> &owner, int me,
                    size_t slab, std::vector<int> &lst) {
  lst.clear();
  for (int a = 0; a < 2; a++)
    for (size_t q = 0; q < pr[a].size(); q++) {
      const Prov &v = pr[a][q];
      if (v.arr == a && v.cell == (int)q && v.sign == 1) continue;
      if (me >= 0 && owner[q / slab] != me) continue;
      const int st = v.cell / (int)slab;
      lst.push_back(a);
      lst.push_back(local[q / slab] * (int)slab + (int)(q % slab));
      lst.push_back((v.arr << 1) | (v.sign < 0 ? 1 : 0));
      lst.push_back((local[st] * (int)slab + v.cell % (int)slab) | ((me >= 0 ? owner[st] : 0) << 28));
    }
}

__global__ void gather_uv_kernel(double *u, double *v, const int4 *lst, int n, int nz, size_t slab, PeerDeltas pd) {
  const size_t total = (size_t)n * nz;
  for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(q % n), k = (int)(q / n);
    const int4 d = lst[e];
    const size_t s0 = (size_t)(d.w & 0x0FFFFFFF);
    const size_t dt = (size_t)d.y / slab, dc = (size_t)d.y % slab, st = s0 / slab, sc = s0 % slab;
    const double *src = (d.z >> 1) ? v : u;
    double val = peer_load(src, pd, (d.w >> 28) & 7, sc + slab * (k + (size_t)nz * st));
    if (d.z & 1) val = -val;
    (d.x ? v : u)[dc + slab * (k + (size_t)nz * dt)] = val;
  }
}

bool exch2_uv_field(double *u, double *v, int nz, bool withSigns) {
  Ctx &c = ctx();
  const int w = withSigns ? 1 : 0;
  const size_t total = (size_t)c.e2UvCount[w] * nz;
  int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)c.numSMs * 16);
  const bool dist = e2_distributed();
  if (dist) {
    if (!in_arena(u) || !in_arena(v)) return fail(70, "exch2 across ranks: the vector pair must be mirrors in the peer arena (uVel / vVel)");
    if (!e2_barrier()) return false;
  }
  c.launches++;
  gather_uv_kernel<<<std::max(blocks, 1), 256, 0, c.stream>>>(u, v, reinterpret_cast<const int4 *>(c.e2UvList[w]), c.e2UvCount[w], nz, c.g.slab,
                                                              e2_deltas());
  MG_CUDA(cudaGetLastError());
  return !dist || e2_barrier();
}

bool exch2_active() { return ctx().e2Count > 0; }

bool exch2_field(double *f, int nz) {
  Ctx &c = ctx();
  const size_t total = (size_t)c.e2Count * nz;
  int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)c.numSMs * 16);
  const bool dist = e2_distributed();
  if (dist) {
    if (!in_arena(f)) return fail(70, "exch2 across ranks: the field must be a mirror in the peer arena (uVel, vVel, wVel, theta, salt, etaN, etaH, cg2d_x, aC2d)");
    if (!e2_barrier()) return false;
  }
  c.launches++;
  gather_kernel<<<std::max(blocks, 1), 256, 0, c.stream>>>(f, reinterpret_cast<const int2 *>(c.e2List), c.e2Count, nz, c.g.slab, e2_deltas());
  MG_CUDA(cudaGetLastError());
  return !dist || e2_barrier();
}

// What one rank needs of a tile graph, compiled on the host (no device involved): the scalar gather list, the
// width-1 push table of its tiles and the two vector-pair gather lists.  g: sizes of one rank (nTiles = its tiles).
struct E2Lists { std::vector<int> scalar, push, uv[2]; };
static bool compile_lists(const E2Tables &T, const Geom &g, int me, bool dist, E2Lists &L) {
  const int slab = (int)g.slab;
  const size_t nGlobal = (size_t)g.slab * T.nT;
  auto enc = [&](int tile, int cell) { return (T.local[tile] * slab + cell) | ((dist ? T.owner[tile] : 0) << 28); };
  // ---- full-width scalar exchange: gather list over halo cells -------------------------------
  std::vector<int> prov;
  if (!compile_gather(T, g.sNx, g.sNy, g.OLx, g.OLx, true, prov)) return false;
  // the gather runs in place and in parallel on every rank (f[dst] = f[src]): that equals the buffered two-pass exchange
  // only if no source of the composed map is itself a destination (sources must be cells no entry writes: interior cells)
  {
    std::vector<char> isDst(nGlobal, 0);
    for (size_t q = 0; q < prov.size(); q++) isDst[q] = prov[q] != (int)q;
    for (size_t q = 0; q < prov.size(); q++)
      if (isDst[q] && isDst[prov[q]]) return fail(72, "exch2: a halo cell is both source and destination of the composed exchange");
  }
  std::vector<int> &lst = L.scalar;
  lst.clear();
  for (size_t q = 0; q < prov.size(); q++)
    if (prov[q] != (int)q) {
      // tile ids -> positions of the tiles in their owner's (bi,bj) order; destinations: my tiles only
      const int dt = (int)(q / g.slab), st = prov[q] / slab;
      if (T.owner[dt] != me) continue;
      lst.push_back(T.local[dt] * slab + (int)(q % g.slab));
      lst.push_back(enc(st, prov[q] % slab));
    }
  // ---- width-1 exchange of CG2D: the push table ------------------------------------------------
  std::vector<int> p1;
  if (!compile_gather(T, g.sNx, g.sNy, 1, 1, false, p1)) return false;
  const int per = 2 * g.sNy + 2 * g.sNx, P1X = g.sNx + 2, P1Y = g.sNy + 2;
  std::vector<int> &tab = L.push;
  tab.assign((size_t)per * g.nTiles, -1);
  auto idx = [&](int i, int j, int tile) {
    return (int)((size_t)(i + g.OLx - 1) + (size_t)g.PX * (size_t)(j + g.OLy - 1) + g.slab * (size_t)tile);
  };
  for (int t = 0; t < T.nT; t++)
    for (int n = 0; n < T.N(t); n++) {
      const int s = T.at(T.nid, n, t) - 1, m = T.at(T.opp, n, t) - 1;
      const int sdir = T.at(T.ndir, m, s);     // edge of the source tile this entry leaves through
      int r[4];
      target_range(T, n, t, 1, false, r);
      for (int j = r[2]; j <= r[3]; j++)
        for (int i = r[0]; i <= r[1]; i++) {
          const int src = p1[(t * P1Y + j) * P1X + i];
          if (src / (P1X * P1Y) != s) return fail(72, "exch2: width-1 map is not a single-source copy");
          if (T.owner[s] != me) continue;      // the rank that owns the source edge pushes it
          const int si = src % P1X, sj = (src / P1X) % P1Y;   // 0-based in the (0:sNx+1) frame = Fortran index
          int slot;
          if (sdir == 4) slot = sj - 1;                         // west edge, by j
          else if (sdir == 3) slot = g.sNy + sj - 1;            // east edge
          else if (sdir == 2) slot = 2 * g.sNy + si - 1;        // south edge, by i
          else slot = 2 * g.sNy + g.sNx + si - 1;               // north edge
          int &e = tab[(size_t)per * T.local[s] + slot];
          if (e != -1) return fail(72, "exch2: an edge point feeds two halo cells through one edge");
          e = idx(i, j, T.local[t]) | ((dist ? T.owner[t] : 0) << 28);      // into the halo of tile t, on its owner
        }
    }
  for (int v : tab)
    if (v < 0) return fail(72, "exch2: an edge point has no neighbour (open edges are not supported)");
  // ---- vector-pair exchange (EXCH_UV_XY / EXCH_UV_XYZ), unsigned and signed ----------------------------
  for (int w = 0; w < 2; w++) {
    std::vector<Prov> pr[2];
    if (!compile_uv(T, g.sNx, g.sNy, g.OLx, w == 1, pr)) return false;
    {      // same in-place condition for the vector pair: (array, cell) written by one entry must not be read by another
      std::vector<char> isDst[2] = {std::vector<char>(nGlobal, 0), std::vector<char>(nGlobal, 0)};
      for (int a = 0; a < 2; a++)
        for (size_t q = 0; q < pr[a].size(); q++) {
          const Prov &v = pr[a][q];
          isDst[a][q] = !(v.arr == a && v.cell == (int)q && v.sign == 1);
        }
      for (int a = 0; a < 2; a++)
        for (size_t q = 0; q < pr[a].size(); q++)
          if (isDst[a][q] && isDst[pr[a][q].arr][pr[a][q].cell])
            return fail(72, "exch2: a halo cell is both source and destination of the composed vector exchange");
    }
    uv_list(pr, T.local, T.owner, dist ? me : -1, g.slab, L.uv[w]);
  }
  return true;
}

static bool set_topology(const E2Tables &T) {
  Ctx &c = ctx();
  const bool dist = c.nRanks > 1;
  E2Lists L;
  if (!compile_lists(T, c.g, c.myRank, dist, L)) return false;
  auto upload = [&](int **dev, const std::vector<int> &v, size_t minInts) {
    if (*dev) cudaFree(*dev);
    *dev = nullptr;
    MG_CUDA(cudaMalloc(dev, std::max(v.size(), minInts) * sizeof(int)));
    MG_CUDA(cudaMemcpy(*dev, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
    return true;
  };
  if (!upload(&c.e2List, L.scalar, 2)) return false;
  c.e2Count = (int)(L.scalar.size() / 2);
  MG_CUDA(cudaMemcpy(c.pushTab, L.push.data(), L.push.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (dist && !cg2d_comm_rank_slots()) return false;
  for (int w = 0; w < 2; w++) {
    if (!upload(&c.e2UvList[w], L.uv[w], 4)) return false;
    c.e2UvCount[w] = (int)(L.uv[w].size() / 4);
  }
  return true;
}

static bool load_tables(E2Tables &T, int nT, int maxN, int nLocal, const int *nNeighbours, const int *neighbourId,
                        const int *opposingSend, const int *neighbourDir, const int *pij, const int *oi, const int *oj,
                        const int *iLo, const int *iHi, const int *jLo, const int *jHi, const int *tBasex,
                        const int *tBasey, const int *isNedge, const int *isSedge, const int *isEedge,
                        const int *isWedge, const int *myTileList, int nRanks, int myRank,
                        const std::vector<int> &tileOwner) {
  T.nT = nT; T.maxN = maxN;
  const size_t nn = (size_t)T.nT * T.maxN;
  T.nN.assign(nNeighbours, nNeighbours + T.nT);
  T.nid.assign(neighbourId, neighbourId + nn); T.opp.assign(opposingSend, opposingSend + nn);
  T.ndir.assign(neighbourDir, neighbourDir + nn); T.pij.assign(pij, pij + 4 * nn);
  T.oi.assign(oi, oi + nn); T.oj.assign(oj, oj + nn);
  T.iLo.assign(iLo, iLo + nn); T.iHi.assign(iHi, iHi + nn); T.jLo.assign(jLo, jLo + nn); T.jHi.assign(jHi, jHi + nn);
  T.bx.assign(tBasex, tBasex + T.nT); T.by.assign(tBasey, tBasey + T.nT);
  T.isN.assign(isNedge, isNedge + T.nT); T.isS.assign(isSedge, isSedge + T.nT);
  T.isE.assign(isEedge, isEedge + T.nT); T.isW.assign(isWedge, isWedge + T.nT);
  T.local.assign(T.nT, -1);
  T.owner.assign(T.nT, 0);
  const bool dist = nRanks > 1;
  if (dist) {
    // W2_tileProc (mitgcm_b200_set_exch2_tile_proc_): tiles are numbered on their owner in increasing tile id, as
    // W2_SET_MAP_TILES fills W2_myTileList; every rank holds the same number of tiles (same arena layout)
    if ((int)tileOwner.size() != T.nT) return fail(70, "exch2 topology across ranks: call mitgcm_b200_set_exch2_tile_proc_ first (W2_tileProc for every tile)");
    if (T.nT != nLocal * nRanks) return fail(70, "exch2 topology across ranks: every rank must hold nTiles / nRanks tiles");
    std::vector<int> cnt(nRanks, 0);
    for (int t = 0; t < T.nT; t++) {
      const int r = tileOwner[t];
      if (r < 0 || r >= nRanks) return fail(70, "exch2 topology: W2_tileProc out of range");
      T.owner[t] = r;
      T.local[t] = cnt[r]++;
    }
    for (int r = 0; r < nRanks; r++)
      if (cnt[r] != nLocal) return fail(70, "exch2 topology across ranks: every rank must hold nTiles / nRanks tiles");
    for (int l = 0; l < nLocal; l++) {
      const int id = myTileList[l];
      if (id < 1 || id > T.nT || T.owner[id - 1] != myRank || T.local[id - 1] != l)
        return fail(70, "exch2 topology: W2_myTileList must list this rank's tiles (W2_tileProc) in increasing tile id");
    }
  } else {
    for (int l = 0; l < nLocal; l++) {
      const int id = myTileList[l];
      if (id < 1 || id > T.nT || T.local[id - 1] != -1) return fail(70, "exch2 topology: bad W2_myTileList");
      T.local[id - 1] = l;
    }
  }
  for (int t = 0; t < T.nT; t++) {
    if (T.local[t] < 0) return fail(70, "exch2 topology: every tile must be in W2_myTileList (one process)");
    if (T.nN[t] < 0 || T.nN[t] > T.maxN) return fail(70, "exch2 topology: bad exch2_nNeighbours");
    for (int n = 0; n < T.nN[t]; n++) {
      const int s = T.at(T.nid, n, t), m = T.at(T.opp, n, t);
      if (s < 1 || s > T.nT || m < 1 || m > T.nN[s - 1]) return fail(70, "exch2 topology: bad neighbour tables");
    }
  }
  return true;
}

}  // namespace mg

using namespace mg;

extern "C" void mitgcm_b200_set_exch2_topology_(
    const int *nTiles, const int *maxNeighbours, const int *nNeighbours, const int *neighbourId,
    const int *opposingSend, const int *neighbourDir, const int *pij, const int *oi, const int *oj,
    const int *iLo, const int *iHi, const int *jLo, const int *jHi, const int *tBasex, const int *tBasey,
    const int *isNedge, const int *isSedge, const int *isEedge, const int *isWedge,
    const int *myTileList, int *ierr) {
  Ctx &c = ctx();
  *ierr = 1;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  const Geom &g = c.g;
  if (c.nRanks > 1 && !halo_connected()) { fail(70, "exch2 topology across ranks: connect the peers first (mitgcm_b200_comm_connect_)"); return; }
  if (*nTiles != g.nTiles * c.nRanks) { fail(70, "exch2 topology: nTiles must equal nSx*nSy of this process times the number of ranks"); return; }
  if (g.OLx != g.OLy) { fail(70, "exch2 topology: OLx must equal OLy"); return; }
  if ((size_t)g.n2 >= ((size_t)1 << 31)) { fail(70, "exch2 topology: tile2d array too large for the gather list"); return; }
  E2Tables T;
  if (!load_tables(T, *nTiles, *maxNeighbours, g.nTiles, nNeighbours, neighbourId, opposingSend, neighbourDir, pij, oi, oj,
                   iLo, iHi, jLo, jHi, tBasex, tBasey, isNedge, isSedge, isEedge, isWedge, myTileList, c.nRanks, c.myRank,
                   c.e2Owner)) return;
  if (!set_topology(T)) return;
  *ierr = 0;
}

// W2_tileProc (W2_EXCH2_TOPOLOGY.h:153-165, filled by w2_map_procs.F:91): 1-based process number owning every tile.
// Needed before mitgcm_b200_set_exch2_topology_ when the tile graph is spread over several ranks.
extern "C" void mitgcm_b200_set_exch2_tile_proc_(const int *nTiles, const int *tileProc, int *ierr) {
  Ctx &c = ctx();
  *ierr = 1;
  if (!c.ready) { fail(30, "mitgcm_b200_init_ not called"); return; }
  if (*nTiles < 1) { fail(70, "set_exch2_tile_proc: nTiles < 1"); return; }
  c.e2Owner.resize(*nTiles);
  for (int t = 0; t < *nTiles; t++) {
    if (tileProc[t] < 1 || tileProc[t] > c.nRanks) { c.e2Owner.clear(); fail(70, "set_exch2_tile_proc: W2_tileProc must be 1 .. number of ranks"); return; }
    c.e2Owner[t] = tileProc[t] - 1;
  }
  *ierr = 0;
}

// Host-only: the compiled vector-pair exchange as a list of (dst array, dst flat cell, src array << 1 | negate,
// src flat cell) over (nTiles, sNy+2*OL, sNx+2*OL) arrays, tiles in id order.  No device needed: lets the host
// set-up code (grid metrics, operator halos) and the CPU tests use exactly what the GPU runs.
extern "C" void mitgcm_b200_exch2_uv_map_(
    const int *dims3, const int *withSigns, const int *nTiles, const int *maxNeighbours, const int *nNeighbours,
    const int *neighbourId, const int *opposingSend, const int *neighbourDir, const int *pij, const int *oi,
    const int *oj, const int *iLo, const int *iHi, const int *jLo, const int *jHi, const int *tBasex,
    const int *tBasey, const int *isNedge, const int *isSedge, const int *isEedge, const int *isWedge,
    const int *maxEntries, int *nEntries, int *out4, int *ierr) {
  *ierr = 1;
  E2Tables T;
  std::vector<int> ids(*nTiles);
  for (int t = 0; t < *nTiles; t++) ids[t] = t + 1;
  if (!load_tables(T, *nTiles, *maxNeighbours, *nTiles, nNeighbours, neighbourId, opposingSend, neighbourDir, pij, oi, oj,
                   iLo, iHi, jLo, jHi, tBasex, tBasey, isNedge, isSedge, isEedge, isWedge, ids.data(), 1, 0, {})) return;
  std::vector<Prov> pr[2];
  if (!compile_uv(T, dims3[0], dims3[1], dims3[2], *withSigns != 0, pr)) return;
  std::vector<int> lst;
  uv_list(pr, T.local, T.owner, -1, (size_t)(dims3[0] + 2 * dims3[2]) * (dims3[1] + 2 * dims3[2]), lst);
  *nEntries = (int)(lst.size() / 4);
  if (*nEntries > *maxEntries) { fail(74, "exch2_uv_map: output buffer too small"); return; }
  std::copy(lst.begin(), lst.end(), out4);
  *ierr = 0;
}

// Host-only (no device needed): what rank `myRank` of `nRanks` would be given for a tile graph spread over ranks --
// the scalar gather list (2 ints per entry: dst, owner << 28 | src), the width-1 push table of its tiles (per tile
// W | E | S | N edge points: owner << 28 | halo index) and the vector-pair gather list (4 ints per entry, source tagged
// the same way), all indices relative to the owning rank's (nTiles / nRanks, sNy+2*OL, sNx+2*OL) arrays.  The CPU tests
// drive a many-rank exchange through these lists in numpy against the literal exch2 algorithm.
// dims3 = (sNx, sNy, OL); sizes(3) in: capacities in ints, out: ints written (scalar, push, uv).
extern "C" void mitgcm_b200_exch2_dist_lists_(
    const int *dims3, const int *nRanks, const int *myRank, const int *tileProc, const int *withSigns, const int *nTiles,
    const int *maxNeighbours, const int *nNeighbours, const int *neighbourId, const int *opposingSend,
    const int *neighbourDir, const int *pij, const int *oi, const int *oj, const int *iLo, const int *iHi, const int *jLo,
    const int *jHi, const int *tBasex, const int *tBasey, const int *isNedge, const int *isSedge, const int *isEedge,
    const int *isWedge, int *sizes, int *scalarOut, int *pushOut, int *uvOut, int *ierr) {
  *ierr = 1;
  if (*nRanks < 1 || *nRanks > 8 || *myRank < 0 || *myRank >= *nRanks || *nTiles % *nRanks) { fail(70, "exch2_dist_lists: bad rank layout"); return; }
  Geom g{};
  g.sNx = dims3[0]; g.sNy = dims3[1]; g.OLx = g.OLy = dims3[2];
  g.PX = g.sNx + 2 * g.OLx; g.PY = g.sNy + 2 * g.OLy; g.nTiles = *nTiles / *nRanks;
  g.slab = (size_t)g.PX * g.PY;
  std::vector<int> owner(*nTiles), mine;
  for (int t = 0; t < *nTiles; t++) {
    owner[t] = tileProc[t] - 1;
    if (owner[t] == *myRank) mine.push_back(t + 1);
  }
  if ((int)mine.size() != g.nTiles) { fail(70, "exch2_dist_lists: every rank must hold nTiles / nRanks tiles"); return; }
  E2Tables T;
  if (!load_tables(T, *nTiles, *maxNeighbours, g.nTiles, nNeighbours, neighbourId, opposingSend, neighbourDir, pij, oi, oj,
                   iLo, iHi, jLo, jHi, tBasex, tBasey, isNedge, isSedge, isEedge, isWedge, mine.data(), *nRanks, *myRank, owner)) return;
  E2Lists L;
  if (!compile_lists(T, g, *myRank, *nRanks > 1, L)) return;
  const std::vector<int> &uv = L.uv[*withSigns != 0 ? 1 : 0];
  if ((int)L.scalar.size() > sizes[0] || (int)L.push.size() > sizes[1] || (int)uv.size() > sizes[2]) { fail(74, "exch2_dist_lists: output buffer too small"); return; }
  std::copy(L.scalar.begin(), L.scalar.end(), scalarOut);
  std::copy(L.push.begin(), L.push.end(), pushOut);
  std::copy(uv.begin(), uv.end(), uvOut);
  sizes[0] = (int)L.scalar.size(); sizes[1] = (int)L.push.size(); sizes[2] = (int)uv.size();
  *ierr = 0;
}
