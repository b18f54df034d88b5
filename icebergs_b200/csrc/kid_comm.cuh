// kid_comm.cuh -- NCCL, loaded at run time, and the device-side pack/unpack of migrating
// bergs (send_bergs_to_other_pes F:2997, pack_berg_into_buffer2 F:3250,
// unpack_berg_from_buffer2 F:3468).
#pragma once
#include <dlfcn.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <vector>

#include "kid_geom.cuh"

namespace kid {

// the slice of the NCCL API this library uses (types per nccl.h 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclInt8_ = 0, ncclInt32_ = 2, ncclInt64_ = 4, ncclFloat64_ = 8 };

struct NcclApi {
  void* dl = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, void*) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok() const { return dl && GetUniqueId && CommInitRank && Send && Recv && GroupStart && GroupEnd && AllGather; }
};

// ncclConfig_t as of NCCL 2.28 (nccl.h: ncclConfig_v22800); used only when the loaded library reports >= 2.28
struct NcclConfig2280 {
  size_t size; unsigned int magic; unsigned int version;
  int blocking, cgaClusterSize, minCTAs, maxCTAs; const char* netName; int splitShare, trafficClass; const char* commName;
  int collnetEnable, CTAPolicy, shrinkShare, nvlsCTAs, nChannelsPerNetPeer, nvlinkCentricSched;
};

inline NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // prefer the NCCL already in the process (torch's), then the system one
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.dl = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (api.dl) break; }
    if (!api.dl) for (const char* n : names) { api.dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.dl) break; }
    if (api.dl) {
      api.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(api.dl, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(api.dl, "ncclCommInitRank");
      api.CommInitRankConfig = (int (*)(ncclComm_t*, int, ncclUniqueId, int, void*))dlsym(api.dl, "ncclCommInitRankConfig");
      api.GetVersion = (int (*)(int*))dlsym(api.dl, "ncclGetVersion");
      api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.dl, "ncclCommDestroy");
      api.GroupStart = (int (*)())dlsym(api.dl, "ncclGroupStart");
      api.GroupEnd = (int (*)())dlsym(api.dl, "ncclGroupEnd");
      api.Send = (int (*)(const void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.dl, "ncclSend");
      api.Recv = (int (*)(void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.dl, "ncclRecv");
      api.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(api.dl, "ncclAllGather");
      api.GetErrorString = (const char* (*)(int))dlsym(api.dl, "ncclGetErrorString");
    }
  }
  return api;
}

// ---- exchange record: one migrating berg = PACK_W fp64 words.  Integers travel as
// their own bit patterns (the reference round-trips them through float()/nint(),
// F:3401/F:3425, exact below 2^53; the id is split in two halves F:3299-3301 -- here
// the 64-bit id is carried whole).
enum PackSlot : int {
  PK_F64_0 = 0,                       // C_LON .. C_FL_K in BergCol order (C_NBASE words)
  PK_ID = C_NBASE, PK_INE_JNE, PK_YEAR_FLAGS,
  PACK_W
};

// What follows the PACK_W words of a record (pack_berg_into_buffer2 F:3250-3354: buffer_width grows with max_bonds,
// mts and dem, F:1266-1292): 3 words per half-bond (other id, other ine/jne, length); with mts the environment cache
// and the fast accelerations (+ ang_vel, ang_accel, rot with dem) = the columns C_NINTER..ncols-1; with dem the bond
// history and saved pair forces (BD_N words) and the broken flag per half-bond.  *_old columns do not travel: the
// receiver resets them from the current state (F:3574-3577).
struct RecLayout {
  int32_t ncols;      // allocated fp64 columns: C_NBASE, C_NINTER, C_NMTS or C_NDEM
  int32_t mb;         // max_bonds
  int32_t dem;
  int32_t w;          // words per record
  int32_t off_bonds, off_extra, off_bdem;
};
__host__ __device__ __forceinline__ RecLayout make_rec_layout(int ncols, int mb, int dem, int mts) {
  RecLayout L;
  L.ncols = ncols; L.mb = mb; L.dem = dem;
  L.off_bonds = PACK_W;
  L.off_extra = L.off_bonds + 3 * mb;
  L.off_bdem = L.off_extra + (ncols > (int)C_NINTER ? ncols - (int)C_NINTER : 0);
  L.w = L.off_bdem + (dem ? (BD_N + 1) * mb : 0);
  (void)mts;
  return L;
}
// everything of the berg in slot s except the cell index / flags words (the callers differ there)
__device__ __forceinline__ void pack_berg(const DevBergs& b, long long s, double* __restrict__ rec, const RecLayout& L) {
#pragma unroll
  for (int c = 0; c < C_NBASE; c++) rec[PK_F64_0 + c] = b.f64[c][s];
  rec[PK_ID] = __longlong_as_double(b.id[s]);
  for (int k = 0; k < L.mb; k++) {          // the bonds travel with the berg, F:3336-3354
    long long slot = (long long)k * b.capacity + s;
    double* br = rec + L.off_bonds + 3 * k;
    br[0] = __longlong_as_double(b.bond_other_id[slot]);
    br[1] = __longlong_as_double(((long long)(unsigned)b.bond_other_ine[slot] << 32) | (unsigned)b.bond_other_jne[slot]);
    br[2] = b.bond_length[slot];
    if (L.dem) {
      double* bd = rec + L.off_bdem + (BD_N + 1) * k;
      for (int q = 0; q < BD_N; q++) bd[q] = b.bond_dem[q][slot];
      bd[BD_N] = (double)b.bond_broken[slot];
    }
  }
  for (int c = C_NINTER; c < L.ncols; c++) rec[L.off_extra + (c - C_NINTER)] = b.f64[c][s];
}
__device__ __forceinline__ void unpack_berg(const DevBergs& b, long long s, const double* __restrict__ rec, const RecLayout& L) {
#pragma unroll
  for (int c = 0; c < C_NBASE; c++) b.f64[c][s] = rec[PK_F64_0 + c];
  if (b.f64[C_UVEL_OLD]) {                  // F:3574-3577
    b.f64[C_UVEL_OLD][s] = rec[PK_F64_0 + C_UVEL]; b.f64[C_VVEL_OLD][s] = rec[PK_F64_0 + C_VVEL];
    b.f64[C_LON_OLD][s] = rec[PK_F64_0 + C_LON]; b.f64[C_LAT_OLD][s] = rec[PK_F64_0 + C_LAT];
  }
  b.id[s] = __double_as_longlong(rec[PK_ID]);
  for (int k = 0; k < L.mb; k++) {
    long long slot = (long long)k * b.capacity + s;
    const double* br = rec + L.off_bonds + 3 * k;
    b.bond_other_id[slot] = __double_as_longlong(br[0]);
    long long oij = __double_as_longlong(br[1]);
    b.bond_other_ine[slot] = (int)(oij >> 32); b.bond_other_jne[slot] = (int)(oij & 0xffffffffll);
    b.bond_length[slot] = br[2];
    b.bond_other_slot[slot] = -1;
    if (L.dem) {
      const double* bd = rec + L.off_bdem + (BD_N + 1) * k;
      for (int q = 0; q < BD_N; q++) b.bond_dem[q][slot] = bd[q];
      b.bond_broken[slot] = (int32_t)bd[BD_N];
    }
  }
  for (int c = C_NINTER; c < L.ncols; c++) b.f64[c][s] = rec[L.off_extra + (c - C_NINTER)];
}

}  // namespace kid

namespace kid {

// ---------------------------------------------------------------- layout
// layout of the ranks over the global grid: the tile in column px, row py of the layout owns columns
// xs[px]..xs[px+1]-1 and rows ys[py]..ys[py+1]-1.  kid_init builds xs, ys and the (px,py) -> rank table from the
// compute domains the ranks were actually given (mpp_define_domains: any mpp_compute_extent split, masked-out
// PEs = -1 in the table); without a table (host-side kid_owner_rank) rank = px + lx*py.
#define KID_MAX_DIV 64
struct DevLayout {
  int32_t lx, ly, gni, gnj, cyclic_x, cyclic_y, rank, nranks;
  int32_t fold_north, pad;    // FOLD_NORTH_EDGE: cell (i, gnj+k) is cell (gni+1-i, gnj+1-k)
  int32_t xs[KID_MAX_DIV + 1], ys[KID_MAX_DIV + 1];
  const int32_t* pe_at;       // [lx*ly] in device memory, or nullptr
};

// owner of global cell (i,j); -1 = outside the model (NULL_PE).  i may be any number of periods off.
__host__ __device__ __forceinline__ int owner_rank(const DevLayout& L, int i, int j) {
  if (L.fold_north && j > L.gnj && j <= 2 * L.gnj) { i = L.gni + 1 - i; j = 2 * L.gnj + 1 - j; }
  if (i < 1 || i > L.gni) { if (!L.cyclic_x) return -1; i = ((i - 1) % L.gni + L.gni) % L.gni + 1; }
  if (j < 1 || j > L.gnj) { if (!L.cyclic_y) return -1; j = ((j - 1) % L.gnj + L.gnj) % L.gnj + 1; }
  int px = 0, py = 0;
  while (px + 1 < L.lx && i >= L.xs[px + 1]) px++;
  while (py + 1 < L.ly && j >= L.ys[py + 1]) py++;
#ifdef __CUDA_ARCH__
  if (L.pe_at) return L.pe_at[px + L.lx * py];
#endif
  return px + L.lx * py;
}

// ------------------------------------------------- berg migration kernels
// k_step appended the slot of every berg that left the tile to leaver_list (count in
// *list_count).  Pass 1: destination rank of each and the per-destination counts
// (send_bergs_to_other_pes F:3024-3050, all directions at once: NVSwitch reaches any rank, so
// the reference's E/W-then-N/S relay F:3103-3106 is not needed).
__global__ void k_leaver_dest(const __grid_constant__ DevLayout L, const int32_t* __restrict__ ine,
                              const int32_t* __restrict__ jne, const int32_t* __restrict__ list,
                              const unsigned long long* __restrict__ list_count, int32_t list_cap,
                              int32_t* __restrict__ dest, int32_t* __restrict__ counts /* [nranks], zeroed */) {
  long long n = (long long)*list_count;
  if (n > list_cap) n = list_cap;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    int32_t s = list[k];
    int d = owner_rank(L, ine[s], jne[s]);
    if (d == L.rank) d = -1;                 // cannot happen (route_berg); never send to self
    dest[k] = d;
    if (d >= 0) atomicAdd(&counts[d], 1);
  }
}

// pass 2: pack_berg_into_buffer2 (F:3250) into the per-destination regions; the slot is freed
// (delete_iceberg_from_list F:3040)
__global__ void k_pack_leavers(const __grid_constant__ DevLayout L, const __grid_constant__ DevBergs b, const int32_t* __restrict__ list,
                               const int32_t* __restrict__ dest, const unsigned long long* __restrict__ list_count,
                               int32_t list_cap, const int32_t* __restrict__ offsets /* [nranks] */,
                               int32_t* __restrict__ cursor /* [nranks], zeroed */, double* __restrict__ sendbuf,
                               const __grid_constant__ RecLayout RL) {
  long long n = (long long)*list_count;
  if (n > list_cap) n = list_cap;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    int32_t s = list[k];
    int d = dest[k];
    uint8_t f = b.flags[s];
    b.flags[s] = 0;
    if (d < 0) continue;                     // left the model through an open boundary
    int pos = offsets[d] + atomicAdd(&cursor[d], 1);
    double* rec = sendbuf + (size_t)pos * RL.w;
    pack_berg(b, s, rec, RL);
    for (int k = 0; k < b.max_bonds; k++) b.bond_other_id[(long long)k * b.capacity + s] = 0;
    int ci = b.ine[s], cj = b.jne[s];   // the owner's own index of the cell: one period off when the berg crossed the seam
    if (L.fold_north && cj > L.gnj) { ci = L.gni + 1 - ci; cj = 2 * L.gnj + 1 - cj; }      // ... mirrored when it crossed the fold
    if (L.cyclic_x && (ci < 1 || ci > L.gni)) ci = ((ci - 1) % L.gni + L.gni) % L.gni + 1;
    rec[PK_INE_JNE] = __longlong_as_double(((long long)(unsigned)ci << 32) | (unsigned)cj);
    rec[PK_YEAR_FLAGS] = __longlong_as_double(((long long)(unsigned)b.start_year[s] << 32) | (unsigned)(f & ~BF_LEAVER));
  }
}

// unpack_berg_from_buffer2 (F:3468): the berg keeps its lon/lat, its cell is re-found on this
// rank's grid (check_and_find_cell F:5973: the sender's indices first -- shifted by one period when
// the berg crossed the cyclic seam, which is the cell the reference's scan F:6002 finds -- then the
// wide search F:3640), xi,yj are recomputed (F:3638) and *_old reset (F:3574-3577).  The berg is
// flagged BF_ARRIVAL: its thermodynamics of this step runs here (k_thermo_range).
__global__ void k_unpack_arrivals(const __grid_constant__ DevGrid g, const __grid_constant__ DevBergs b,
                                  const __grid_constant__ DevParams p, DevCounters* __restrict__ cnt,
                                  const double* __restrict__ recvbuf, long long n_recv, long long s0,
                                  const __grid_constant__ RecLayout RL) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_recv) return;
  if (k == 0) cnt->n_slots = (unsigned long long)(s0 + n_recv);     // the append cursor moves past the arrivals
  long long s = s0 + k;
  const double* rec = recvbuf + (size_t)k * RL.w;
  unpack_berg(b, s, rec, RL);
  double lon = rec[PK_F64_0 + C_LON], lat = rec[PK_F64_0 + C_LAT];
  long long ij = __double_as_longlong(rec[PK_INE_JNE]), yf = __double_as_longlong(rec[PK_YEAR_FLAGS]);
  int i = (int)(ij >> 32), j = (int)(ij & 0xffffffffll);
  b.start_year[s] = (int32_t)(yf >> 32);
  uint8_t f = (uint8_t)(yf & 0xff);
  bool found = false;
  int oi = i, oj = j;
  if (cell_on_pe(g, oi, oj)) found = is_point_in_cell(g, p, lon, lat, oi, oj, &cnt->error_flags);
  if (!found && g.cyclic_x) {
    for (int sh = -1; sh <= 1 && !found; sh += 2) {
      oi = i + sh * g.gni;
      if (cell_on_pe(g, oi, oj)) found = is_point_in_cell(g, p, lon, lat, oi, oj, &cnt->error_flags);
    }
  }
  if (!found) found = find_cell_wide(g, p, lon, lat, &oi, &oj, &cnt->error_flags);
  if (!found) { atomicOr(&cnt->error_flags, (unsigned)KID_DEVERR_LOST_BERG); b.flags[s] = 0; return; }
  double xi, yj;
  pos_within_cell(g, p, lon, lat, oi, oj, &xi, &yj, &cnt->error_flags);
  b.f64[C_XI][s] = xi; b.f64[C_YJ][s] = yj;
  b.ine[s] = oi; b.jne[s] = oj;
  b.halo_code[s] = 0;
  b.flags[s] = (uint8_t)((f | BF_ALIVE | BF_ARRIVAL) & ~(BF_LEAVER | BF_HALO));
}

// ------------------------------------------------------- halo strips
// mpp_update_domains (F:1058-1066, I:5321-5355) for a set of data-domain fields: the strip of
// compute cells next to each of the 8 neighbours is packed, exchanged and written into the
// matching halo strip.  dir = (dx+1) + 3*(dy+1), dir 4 unused.
struct HaloStrips {
  int32_t i0[9], j0[9], ni[9], nj[9];   // strip origin (global indices) and size, per direction
  long long off[9];                      // offset of the strip in the buffer, in cells (x nf fields)
  int32_t nf, pad;
};
struct HaloFields { double* f[16]; };

__global__ void k_halo_pack(const __grid_constant__ DevGrid g, const __grid_constant__ HaloStrips st,
                            const __grid_constant__ HaloFields fl, double* __restrict__ buf, int unpack) {
  int dir = blockIdx.y;
  if (dir == 4) return;
  long long ncell = (long long)st.ni[dir] * st.nj[dir];
  long long n = ncell * st.nf;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    int q = (int)(k / ncell);
    long long c = k - (long long)q * ncell;
    int ii = (int)(c % st.ni[dir]), jj = (int)(c / st.ni[dir]);
    size_t cell = gidx(g, st.i0[dir] + ii, st.j0[dir] + jj);
    double* slot = buf + (size_t)st.off[dir] * st.nf + k;
    if (unpack) fl.f[q][cell] = *slot; else *slot = fl.f[q][cell];
  }
}

// ------------------------------------------------------- tripolar fold
// mpp_update_domains across FOLD_NORTH_EDGE (F:649, F:933; the tripolar grid): the halo rows beyond the folded northern
// edge are the top rows of the model read backwards.  For a point at position (sx, sy) -- (0,0) cell centre, (1,1) NE
// corner, (1,0) east face, (0,1) north face --
//     f(i, gnj+k) = sign * f(gni+1-sx-i, gnj+1-sy-k),  k = 1..halo, i cyclic,
// sign = -1 for the components of a true vector.  For a true vector ON the fold (sy = 1) the eastern half of row gnj is
// the western half mirrored with the sign, and the two pole points of a corner field are zero.  (FMS itself is outside
// the reference tree: its documented semantics restated; oracle/kid_oracle.c halo_update_pos, pinned by the analytic
// continuation of a bipolar cap, tests/test_fold_oracle.py.)
// The ranks of the top row share their top halo+1 rows: every such rank holds the whole strip, piece by piece in the
// order of the layout columns (piece px: [nf][halo+1][nic_px], rows gnj-halo .. gnj), and fills its own halo from it.
struct FoldKind { int8_t sx, sy, sign, vector; };
struct FoldArgs {
  int32_t nf, w, gni, gnj, isd, ied, isc, iec, lx, pad;
  int32_t xs[KID_MAX_DIV + 1];
  double* f[16];
  FoldKind kind[16];
};
__device__ __forceinline__ double fold_strip_at(const FoldArgs& a, const double* __restrict__ strip, int q, int r, int si) {
  int px = 0;
  while (px + 1 < a.lx && si >= a.xs[px + 1]) px++;
  int nicp = a.xs[px + 1] - a.xs[px];
  size_t base = (size_t)a.nf * (a.w + 1) * (size_t)(a.xs[px] - 1);
  return strip[base + ((size_t)q * (a.w + 1) + r) * nicp + (si - a.xs[px])];
}
// my piece of the strip
__global__ void k_fold_pack(const __grid_constant__ DevGrid g, const __grid_constant__ FoldArgs a, double* __restrict__ strip) {
  int nic = a.iec - a.isc + 1;
  long long n = (long long)a.nf * (a.w + 1) * nic;
  size_t base = (size_t)a.nf * (a.w + 1) * (size_t)(a.isc - 1);
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    int ii = (int)(k % nic);
    long long t = k / nic;
    int r = (int)(t % (a.w + 1)), q = (int)(t / (a.w + 1));
    strip[base + k] = a.f[q][gidx(g, a.isc + ii, a.gnj - a.w + r)];
  }
}
__global__ void k_fold_fill(const __grid_constant__ DevGrid g, const __grid_constant__ FoldArgs a, const double* __restrict__ strip) {
  int nid = a.ied - a.isd + 1;
  long long n = (long long)a.nf * (a.w + 1) * nid;
  for (long long kk = (long long)blockIdx.x * blockDim.x + threadIdx.x; kk < n; kk += (long long)gridDim.x * blockDim.x) {
    int ii = (int)(kk % nid);
    long long t = kk / nid;
    int k = (int)(t % (a.w + 1)), q = (int)(t / (a.w + 1));
    const FoldKind fk = a.kind[q];
    int i = a.isd + ii;
    if (k == 0) {                               // the fold row itself
      if (!(fk.vector && fk.sy == 1)) continue;
      int iw = ((i - 1) % a.gni + a.gni) % a.gni + 1;
      if (fk.sx == 1 && (iw == a.gni / 2 || iw == a.gni)) { a.f[q][gidx(g, i, a.gnj)] = 0.; continue; }
      if (iw > a.gni / 2 && iw <= a.gni - fk.sx)
        a.f[q][gidx(g, i, a.gnj)] = (double)fk.sign * fold_strip_at(a, strip, q, a.w, a.gni + 1 - fk.sx - iw);
      continue;
    }
    int si = a.gni + 1 - fk.sx - i;
    si = ((si - 1) % a.gni + a.gni) % a.gni + 1;
    int sj = a.gnj + 1 - fk.sy - k;             // in gnj-w .. gnj
    a.f[q][gidx(g, i, a.gnj + k)] = (double)fk.sign * fold_strip_at(a, strip, q, sj - (a.gnj - a.w), si);
  }
}

// ----------------------------------------------------- in-process group
// Ranks that live in one process (one host thread per rank, e.g. several tiles on one GPU, or a
// host without NCCL): buffers are exchanged with device-to-device copies after a rendezvous.
struct PostedMsg { int dst, tag; const void* ptr; size_t bytes; int device; };
struct LocalGroup {
  int nranks = 0;
  std::mutex m;
  std::condition_variable cv;
  int waiting = 0;
  unsigned long long gen = 0;
  bool failed = false;
  std::vector<std::vector<PostedMsg>> posted;
  std::vector<std::vector<int32_t>> counts;
  explicit LocalGroup(int n) : nranks(n), posted(n), counts(n) {}
  bool barrier() {
    std::unique_lock<std::mutex> lk(m);
    if (failed) return false;
    unsigned long long g0 = gen;
    if (++waiting == nranks) { waiting = 0; gen++; cv.notify_all(); return true; }
    if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return gen != g0 || failed; })) { failed = true; cv.notify_all(); return false; }
    return !failed;
  }
};

}  // namespace kid
