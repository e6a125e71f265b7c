// pt_api.cu -- extern "C" layer of libsisi4s_pt (see include/sisi4s_pt.h).
// Host logic only: handle lifetime, uploads + one-time packing, the triple / orbit work lists,
// the hole-block walk for shapes that exceed one GPU, launches, and the final fixed-order sums.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "pt_common.cuh"

using namespace pt;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

}  // namespace

// library-internal: lets the other translation units (upt.cu) report through pt_last_error()
namespace pt {
int record_error(int code, const char* message) { return fail(code, "%s", message); }
// catch (...) handler of every extern "C" entry point: no C++ exception crosses the C ABI
int on_exception(const char* where) {
  try {
    throw;
  } catch (const std::bad_alloc&) {
    return fail(PT_ERR_NOMEM, "%s: host memory exhausted", where);
  } catch (const std::exception& e) {
    return fail(PT_ERR_INVALID, "%s: %s", where, e.what());
  } catch (...) {
    return fail(PT_ERR_INVALID, "%s: unknown C++ exception", where);
  }
}
}  // namespace pt

namespace {

#define CU(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(e_ == cudaErrorMemoryAllocation ? PT_ERR_NOMEM : PT_ERR_CUDA,          \
                  "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
  } while (0)
#define RC(call)                 \
  do {                           \
    if (int rc_ = (call)) return rc_; \
  } while (0)

// stream-ordered scratch memory of one call: allocation and release are enqueued on the handle's
// stream (cudaMallocAsync / cudaFreeAsync), so a setter never synchronises the device
template <typename T>
struct StreamScratch {
  T* p = nullptr;
  cudaStream_t s = nullptr;
  StreamScratch() = default;
  StreamScratch(const StreamScratch&) = delete;
  StreamScratch& operator=(const StreamScratch&) = delete;
  ~StreamScratch() { if (p) cudaFreeAsync(p, s); }
  cudaError_t alloc(size_t n, cudaStream_t st) { s = st; return cudaMallocAsync((void**)&p, n * sizeof(T), st); }
  T* release() { T* q = p; p = nullptr; return q; }
  operator T*() const { return p; }
};

}  // namespace

struct PtHandle_ {
  Dims d{};             // dims of the device-resident problem; hole-block mode: o = most active holes of a group
  int o_full = 0;       // holes of the full problem (= d.ol)
  int device = 0;
  int sm_count = 0;
  int engine = PT_ENGINE_FUSED;
  int keep_raw = 0;
  int grid = 0;
  int order = 1;
  int item_sync = 0;    // optional item-round barrier between the CTAs of the fused kernel (forces L2 reuse of the PPPH tiles; measured: 6.8x less DRAM traffic but 5 % slower than the free-running equal-cost order)
  int class_sort = 1;   // launch list ordered generic triples first (equal-cost items keep the CTAs in step without a barrier)
  int tile_holes = 0;   // optional: launch list grouped by hole blocks of this width (0 = reference order; measured: no effect on DRAM traffic)
  int debug = 0;
  int async_upload = 0; // 1: setters only enqueue (host buffers must stay valid until pt_sync / pt_run returns)
  int pin_host = 0;     // 1: page-lock the caller's large host tensors (hole-block mode) for full-speed DMA
  unsigned int* d_sync = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_up0 = nullptr, ev_up1 = nullptr;
  bool up_pending = false, up_started = false;
  // raw device tensors
  double *epsi = nullptr, *epsa = nullptr, *t1 = nullptr, *pphh = nullptr, *qsum = nullptr;
  double *t1b = nullptr, *pphhb = nullptr, *qsumb = nullptr;   // optional second singles term (pt_set_singles_pair)
  double *t2_raw = nullptr, *hhhp_raw = nullptr, *ppph_raw = nullptr;  // keep_raw only (hhhp_raw also: hole-block mode)
  // packed
  double *Tt = nullptr, *T2h = nullptr, *Vt = nullptr, *Ut = nullptr;
  double* slab_stage = nullptr;  // one raw PPPH slab (host-PPPH sources)
  std::vector<char> slab_set;    // a source for slab k has been given
  // hole-blocked PPPH residency (option slab_slots = S < o, implied by hole_block): Vt holds S slab
  // slots, slabs are (re)built on demand from the resident vertex or a caller-owned host tensor
  int slab_slots = 0;
  std::vector<int> slot_of;      // hole -> slot of its packed slab, -1 = not resident
  std::vector<int> hole_in;      // slot -> hole, -1 = free
  std::vector<long long> slot_tick;
  long long tick = 0;
  int* d_vslot = nullptr;        // active hole -> slot
  // resident K-major image of the CoulombVertex (pt_pack.cu)
  double* gp = nullptr;
  int g_nf = 0, g_np = 0;
  const double* host_ppph = nullptr;        // caller-owned PPPHCoulombIntegrals[v,v,v,o]
  // all-resident engines with asynchronous setters: the slabs of a host PPPH tensor are uploaded by pt_run,
  // only those its triples touch, on a second stream while earlier waves of triples already compute
  bool lazy_ppph = false;
  cudaStream_t copy_stream = nullptr;
  // hole-block mode (option hole_block = b: BASELINE configs[4]): T2 / PPHH stay in caller-owned host
  // memory, the sorted triples are walked by hole-block triples (I<=J<=K), each group's <= 3b active
  // holes are staged into the buffers above (sized for 3b holes once) before its launch
  int hole_block = 0;
  const double *host_t2 = nullptr, *host_pphh = nullptr;
  std::vector<double> host_epsi, host_t1;
  double* stage_raw = nullptr;   // raw staging of T2 blocks: max(v^2 na^2, v^2 o) doubles
  int* d_hmap = nullptr;         // active hole -> hole of the full problem
  std::vector<int> cur_holes;    // holes currently staged
  std::vector<void*> registered; // cudaHostRegister'ed caller buffers
  bool pphh_from_vertex = false, hhhp_from_vertex = false;
  bool ppph_from_vertex = false;   // slabs come from the resident vertex (pt_set_vertex was the last PPPH source given)
  // holes of the triple enumeration = holes that own a PPPH slab: the active holes of the engine, or --
  // in hole-block mode, where the active set changes from group to group -- all holes of the problem
  int oh() const { return hole_block ? o_full : d.o; }
  int nslots() const { return (slab_slots > 0 && slab_slots < oh()) ? slab_slots : oh(); }
  bool blocked() const { return nslots() < oh(); }
  bool use_vslot() const { return hole_block || blocked(); }
  bool have_t2h = false;
  bool have_eps = false, have_t1 = false, have_t2 = false, have_pphh = false, have_hhhp = false;
  // work lists + per-run scratch (grow-only)
  uchar4* d_orbits = nullptr;
  int norbits = 0;
  int4* d_list = nullptr; size_t cap_list = 0;
  double* d_e = nullptr;  size_t cap_e = 0;
  double* d_item = nullptr; size_t cap_item = 0;
  PtStats stats{};
  double bytes_alloc = 0;

  template <typename T>
  cudaError_t alloc(T** p, size_t n) {
    cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) bytes_alloc += (double)(n * sizeof(T));
    return e;
  }
  template <typename T>
  cudaError_t grow(T** p, size_t* cap, size_t n) {
    if (n <= *cap) return cudaSuccess;
    if (*p) { cudaFree(*p); bytes_alloc -= (double)(*cap * sizeof(T)); *p = nullptr; *cap = 0; }
    cudaError_t e = alloc(p, n);
    if (e == cudaSuccess) *cap = n;
    return e;
  }
};

namespace {

struct Triple { int i, j, k; };

// reference enumeration order, CcsdPerturbativeTriples.cxx:156-158
void enumerate_triples(int o, std::vector<Triple>& out) {
  out.clear();
  out.reserve((size_t)o * (o + 1) / 2 * (o + 2) / 3);   // an absurd o fails here, at once (std::length_error / bad_alloc)
  for (int i = 0; i < o; ++i)
    for (int j = i; j < o; ++j)
      for (int k = j; k < o; ++k) out.push_back({i, j, k});
}
inline int triple_class(const Triple& t) { return (t.i == t.j ? 1 : 0) + (t.j == t.k ? 2 : 0); }
inline int triple_weight(const Triple& t) {
  static const int w[4] = {6, 3, 3, 1};
  return w[triple_class(t)];
}

// hole-block walk: a sorted triple belongs to the group of the hole blocks (I<=J<=K) of width bw it touches;
// K runs fastest in the key, so consecutive groups share the blocks I and J
inline long long block_key(int i, int j, int k, int bw, long long nb) {
  return ((long long)(i / bw) * nb + j / bw) * nb + k / bw;
}
// ... and needs the holes of those (at most three distinct) blocks, ascending
std::vector<int> group_holes(int i, int j, int k, int bw, int o) {
  std::vector<int> holes;
  const int first[3] = {i / bw * bw, j / bw * bw, k / bw * bw};
  for (int m = 0; m < 3; ++m)
    for (int z = first[m]; z < std::min(first[m] + bw, o); ++z)
      if (std::find(holes.begin(), holes.end(), z) == holes.end()) holes.push_back(z);
  std::sort(holes.begin(), holes.end());
  return holes;
}

// device time of the uploads + packing: one event window per synchronous setter, one window over
// all of them in async mode (closed by pt_sync / pt_run)
struct UploadScope {
  pt_handle_t h;
  explicit UploadScope(pt_handle_t h_) : h(h_) {
    if (!h->async_upload) cudaEventRecord(h->ev0, h->stream);
    else if (!h->up_started) { cudaEventRecord(h->ev_up0, h->stream); h->up_started = true; }
  }
  int done() {
    if (!h->async_upload) {
      CU(cudaEventRecord(h->ev1, h->stream));
      CU(cudaEventSynchronize(h->ev1));
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
      h->stats.seconds_upload += ms * 1e-3;
    } else {
      CU(cudaEventRecord(h->ev_up1, h->stream));
      h->up_pending = true;
    }
    return PT_OK;
  }
};

int sync_uploads(pt_handle_t h) {
  if (h->up_pending) {
    CU(cudaEventSynchronize(h->ev_up1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev_up0, h->ev_up1));
    h->stats.seconds_upload += ms * 1e-3;
    h->up_pending = false;
    h->up_started = false;
  }
  return PT_OK;
}

int upload(pt_handle_t h, double* dst, const double* src, size_t n) {
  CU(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->stats.bytes_h2d += (double)(n * sizeof(double));
  return PT_OK;
}

// page-lock a caller-owned tensor for full-speed DMA (option pin_host), as ONE region (a copy must not
// straddle two registrations).  Read-only registration first: it is the one the kernel grants for
// shared file mappings (/dev/shm).  When the memory is already page-locked by the caller, or cannot
// be locked, the copies simply stay as they are (staged by the driver).
void pin(pt_handle_t h, const double* p, size_t n) {
  if (!h->pin_host || !p) return;
  const size_t bytes = n * sizeof(double);
  cudaError_t e = cudaHostRegister((void*)p, bytes, cudaHostRegisterReadOnly);
  if (e != cudaSuccess) { cudaGetLastError(); e = cudaHostRegister((void*)p, bytes, cudaHostRegisterDefault); }
  if (e != cudaSuccess) { cudaGetLastError(); return; }
  h->registered.push_back((void*)p);
  h->stats.bytes_pinned += (double)bytes;
}

int configure_kernels() {
  CU(fused_configure(nullptr));
  CU(vertex_gemm_configure());
  return PT_OK;
}

int init_handle(pt_handle_t h, int o, int o_all, int v, int device, int sm_count) {
  h->d = make_dims(o, v, o_all);
  h->o_full = o_all;
  h->device = device;
  h->sm_count = sm_count;
  h->stats.sm_count = sm_count;
  h->slab_set.assign(o, 0);
  h->slot_of.assign(o, -1);
  if (h->d.nr > 255) return fail(PT_ERR_UNSUPPORTED, "pt_create: v too large (nr=%d > 255)", h->d.nr);
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&h->ev0));
  CU(cudaEventCreate(&h->ev1));
  CU(cudaEventCreate(&h->ev_up0));
  CU(cudaEventCreate(&h->ev_up1));
  RC(configure_kernels());
  // orbit list: A >= B >= C, C fastest; generic orbits (A>B>C, 18 steps per item) first, degenerate
  // ones after: the CTAs of the fused kernel then advance through equal-cost items, see pt_fused.cu
  std::vector<uchar4> orb;
  for (int A = 0; A < h->d.nr; ++A)
    for (int B = 0; B <= A; ++B)
      for (int C = 0; C <= B; ++C) {
        uchar4 u;
        u.x = (unsigned char)A; u.y = (unsigned char)B; u.z = (unsigned char)C;
        u.w = (unsigned char)((A == B ? 1 : 0) + (B == C ? 2 : 0));
        orb.push_back(u);
      }
  std::stable_sort(orb.begin(), orb.end(), [](const uchar4& a, const uchar4& b) { return a.w < b.w; });
  h->norbits = (int)orb.size();
  CU(h->alloc(&h->d_orbits, orb.size()));
  CU(cudaMemcpy(h->d_orbits, orb.data(), orb.size() * sizeof(uchar4), cudaMemcpyHostToDevice));
  return PT_OK;
}

// the Vabci / Vabij / Vijka GEMMs of CoulombIntegralsFromVertex.cxx on the resident vertex image
VgParams vg_base(pt_handle_t h) {
  VgParams p{};
  p.gp = h->gp;
  p.rows_padded = vertex_rows_padded(h->g_np);
  p.kp4 = vertex_kp4(h->g_nf);
  p.np = h->g_np;
  p.a0 = h->g_np - h->d.v;
  p.v = h->d.v;
  p.nr = h->d.nr;
  p.nk4 = h->d.nk4;
  p.nb0 = p.nb1 = 1;
  p.alpha = 1.0;
  return p;
}
// Vabci["abci"] = G["Gac"] G["Gbi"] (:430-431): slab z, kernel naming V[b,c,d,z] = sum_G G[G,b,d] G[G,c,z],
// written straight into the packed layout
int build_slab_packed(pt_handle_t h, int z, double* vt_slab) {
  VgParams p = vg_base(h);
  p.mode = VG_PACKED;
  p.a_base = p.a0;                                  // + 16 Q + np (a0 + d)
  p.b_base = p.a0 + (long long)p.np * z;            // + c
  p.out = vt_slab;
  CU(launch_vertex_gemm(p, h->stream));
  h->stats.kernel_launches += 1;
  return PT_OK;
}
// the same slab, raw column-major [v,v,v] (a + v (b + v c)): batch c, M = a, N = b
int build_slab_raw(pt_handle_t h, int z, double* raw) {
  VgParams p = vg_base(h);
  p.mode = VG_STRIDED;
  p.nb0 = p.v;                                       // batch c
  p.a_base = p.a0 + (long long)p.np * p.a0; p.a_s0 = p.np;     // rows (a, c)
  p.b_base = p.a0 + (long long)p.np * z;                       // rows (b, z)
  p.o_s0 = (long long)p.v * p.v;
  p.M = p.N = p.v; p.sm = 1; p.sn = p.v;
  p.out = raw;
  CU(launch_vertex_gemm(p, h->stream));
  h->stats.kernel_launches += 1;
  return PT_OK;
}
// Vabij["abij"] = G["Gai"] G["Gbj"] (:402-403) for na x na hole pairs (optionally a hole subset `map`),
// output [v,v,na,na]
int build_pphh(pt_handle_t h, int na, const int* d_map, double* out) {
  VgParams p = vg_base(h);
  p.mode = VG_STRIDED;
  p.nb0 = p.nb1 = na; p.map0 = p.map1 = d_map;       // batch (i, j)
  p.a_base = p.a0; p.a_s0 = p.np;                    // rows (a, i)
  p.b_base = p.a0; p.b_s1 = p.np;                    // rows (b, j)
  p.o_s0 = (long long)p.v * p.v; p.o_s1 = (long long)p.v * p.v * na;
  p.M = p.N = p.v; p.sm = 1; p.sn = p.v;
  p.out = out;
  CU(launch_vertex_gemm(p, h->stream));
  h->stats.kernel_launches += 1;
  return PT_OK;
}
// Vijka["ijka"] = G["Gik"] G["Gaj"] (:416-417), full [o,o,o,v]: batch (k, j), M = i, N = a
int build_hhhp(pt_handle_t h, double* out) {
  VgParams p = vg_base(h);
  const long long o = h->o_full;
  p.mode = VG_STRIDED;
  p.nb0 = p.nb1 = (int)o;                            // b0 = k, b1 = j
  p.a_base = 0; p.a_s0 = p.np;                       // rows (i, k)
  p.b_base = p.a0; p.b_s1 = p.np;                    // rows (a, j)
  p.o_s0 = o * o; p.o_s1 = o;
  p.M = (int)o; p.N = p.v; p.sm = 1; p.sn = o * o * o;
  p.out = out;
  CU(launch_vertex_gemm(p, h->stream));
  h->stats.kernel_launches += 1;
  return PT_OK;
}

}  // namespace

extern "C" {

const char* pt_last_error(void) { return g_last_error.c_str(); }
const char* pt_version(void) { return "sisi4s_b200 (T) 0.2 sm_100a"; }

int64_t pt_num_triples(int o) { return (int64_t)o * (o + 1) * (o + 2) / 6; }

int pt_partition(int o, int nranks, int rank, int64_t* begin, int64_t* end) try {
  if (o < 1 || nranks < 1 || rank < 0 || rank >= nranks || !begin || !end)
    return fail(PT_ERR_INVALID, "pt_partition: bad arguments");
  std::vector<Triple> tr;
  enumerate_triples(o, tr);
  // contiguous chunks of (nearly) equal weight = number of W blocks to build
  long long total = 0;
  for (auto& t : tr) total += triple_weight(t);
  auto cut = [&](int r) -> int64_t {
    if (r <= 0) return 0;
    if (r >= nranks) return (int64_t)tr.size();
    const double target = (double)total * r / nranks;
    long long acc = 0;
    for (size_t n = 0; n < tr.size(); ++n) {
      if ((double)acc >= target) return (int64_t)n;
      acc += triple_weight(tr[n]);
    }
    return (int64_t)tr.size();
  };
  *begin = cut(rank);
  *end = cut(rank + 1);
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_partition");
}

int64_t pt_estimate_device_bytes(int o, int v, int slab_slots, int hole_block) {
  if (o < 1 || v < 1) return 0;
  const int oa = hole_block > 0 ? std::min(3 * hole_block, o) : o;   // active holes on the device
  const Dims d = make_dims(oa, v, o);
  int slots = hole_block > 0 ? std::max(slab_slots, oa) : slab_slots;
  if (slots <= 0 || slots > o) slots = o;
  const double vv = (double)v * v;
  double n = (double)vt_slab_elems(d) * slots + (double)tt_elems(d) + (double)t2h_elems(d) + (double)ut_elems(d)
             + 2.0 * vv * oa * oa          /* PPHH + its pre-added pair sums */
             + (double)v * oa + oa + v     /* T1, eigenenergies */
             + vv * v;                     /* one raw slab stage (PPPH given as a tensor) */
  if (hole_block > 0) n += vv * std::max((double)oa * oa, (double)o) + (double)o * o * o * v;   /* T2 staging, full HHHP */
  const double nr = d.nr;
  return (int64_t)(8.0 * n + 4.0 * nr * (nr + 1) * (nr + 2) / 6);
}

int pt_plan_hole_blocks(int o, int hole_block, int64_t begin, int64_t end, int64_t* n_groups, int32_t* max_active_holes,
                        int64_t* slab_loads) try {
  if (o < 1 || hole_block < 1 || begin < 0 || begin > end || end > pt_num_triples(o))
    return fail(PT_ERR_INVALID, "pt_plan_hole_blocks: bad arguments");
  std::vector<Triple> tr;
  enumerate_triples(o, tr);
  const int bw = std::min(hole_block, o);
  const long long nb = (o + bw - 1) / bw;
  const int slots = std::min(3 * bw, o);
  std::vector<std::pair<long long, Triple>> keyed;
  for (int64_t n = begin; n < end; ++n)
    if (triple_class(tr[n]) != 3) keyed.push_back({block_key(tr[n].i, tr[n].j, tr[n].k, bw, nb), tr[n]});
  std::stable_sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  // the walk pt_run does: one group per key; slabs kept in `slots` slots with LRU replacement
  int64_t groups = 0, loads = 0;
  int32_t max_active = 0;
  std::vector<long long> tick(o, -1);   // last use of a resident slab, -1 = not resident
  int resident = 0;
  long long now = 0;
  for (size_t g0 = 0; g0 < keyed.size();) {
    size_t g1 = g0;
    while (g1 < keyed.size() && keyed[g1].first == keyed[g0].first) ++g1;
    const std::vector<int> holes = group_holes(keyed[g0].second.i, keyed[g0].second.j, keyed[g0].second.k, bw, o);
    max_active = std::max<int32_t>(max_active, (int32_t)holes.size());
    ++now;
    for (int z : holes)
      if (tick[z] >= 0) tick[z] = now;
    for (int z : holes) {
      if (tick[z] >= 0) continue;
      if (resident == slots) {   // evict the least recently used slab that this group does not need
        int victim = -1;
        for (int q = 0; q < o; ++q)
          if (tick[q] >= 0 && tick[q] != now && (victim < 0 || tick[q] < tick[victim])) victim = q;
        if (victim < 0) return fail(PT_ERR_INVALID, "pt_plan_hole_blocks: %zu slabs needed, %d slots", holes.size(), slots);
        tick[victim] = -1;
        --resident;
      }
      tick[z] = now;
      ++resident;
      ++loads;
    }
    ++groups;
    g0 = g1;
  }
  if (n_groups) *n_groups = groups;
  if (max_active_holes) *max_active_holes = max_active;
  if (slab_loads) *slab_loads = loads;
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_plan_hole_blocks");
}

int pt_create(pt_handle_t* out, int o, int v, int device) { return pt_create_ex(out, o, o, v, device); }

int pt_create_ex(pt_handle_t* out, int o, int o_all, int v, int device) try {
  if (!out || o < 1 || v < 1 || o_all < o) return fail(PT_ERR_INVALID, "pt_create: need 1 <= o_act <= o_all, v >= 1");
  if ((v + TILE - 1) / TILE > 255) return fail(PT_ERR_UNSUPPORTED, "pt_create: v too large (%d particle ranges > 255)", (v + TILE - 1) / TILE);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(PT_ERR_CUDA, "pt_create: no CUDA device (%s); this library has no CPU fallback",
                cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(PT_ERR_INVALID, "pt_create: device %d of %d", device, ndev);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(PT_ERR_UNSUPPORTED, "pt_create: device sm_%d%d, built for sm_100a only", prop.major,
                prop.minor);
  // every failure path below releases what was created so far
  std::unique_ptr<PtHandle_, int (*)(pt_handle_t)> h(new PtHandle_(), pt_destroy);
  RC(init_handle(h.get(), o, o_all, v, device, prop.multiProcessorCount));
  *out = h.release();
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_create_ex");
}

int pt_destroy(pt_handle_t h) try {
  if (!h) return PT_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  double* ptrs[] = {h->t1b, h->pphhb, h->qsumb, h->epsi, h->epsa, h->t1, h->pphh, h->qsum, h->t2_raw, h->hhhp_raw, h->ppph_raw,
                    h->Tt, h->T2h, h->Vt, h->Ut, h->slab_stage, h->gp, h->stage_raw, h->d_e, h->d_item};
  for (double* p : ptrs)
    if (p) cudaFree(p);
  if (h->d_orbits) cudaFree(h->d_orbits);
  if (h->d_vslot) cudaFree(h->d_vslot);
  if (h->d_hmap) cudaFree(h->d_hmap);
  if (h->d_sync) cudaFree(h->d_sync);
  if (h->d_list) cudaFree(h->d_list);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  for (void* p : h->registered) cudaHostUnregister(p);
  for (cudaEvent_t ev : {h->ev0, h->ev1, h->ev_up0, h->ev_up1})
    if (ev) cudaEventDestroy(ev);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_destroy");
}

int pt_set_option(pt_handle_t h, const char* key, int64_t value) try {
  if (!h || !key) return fail(PT_ERR_INVALID, "pt_set_option: null");
  const bool any_input = h->have_eps || h->have_t1 || h->have_t2 || h->have_pphh || h->have_hhhp || h->Vt || h->gp;
  if (!strcmp(key, "engine")) {
    if (value != PT_ENGINE_FUSED && value != PT_ENGINE_NAIVE) return fail(PT_ERR_INVALID, "engine %lld", (long long)value);
    h->engine = (int)value;
  } else if (!strcmp(key, "keep_raw")) {
    h->keep_raw = value != 0;
  } else if (!strcmp(key, "grid")) {
    if (value < 0) return fail(PT_ERR_INVALID, "grid %lld", (long long)value);
    h->grid = (int)value;
  } else if (!strcmp(key, "slab_slots")) {
    if (value < 0 || (value > 0 && value < 3)) return fail(PT_ERR_INVALID, "slab_slots %lld (0 = all resident, else >= 3)", (long long)value);
    if (h->Vt) return fail(PT_ERR_INVALID, "slab_slots must be set before the PPPH integrals / vertex");
    if (h->hole_block && value > 0 && value < h->d.o)
      return fail(PT_ERR_INVALID, "slab_slots %lld < %d active holes of a hole-block group", (long long)value, h->d.o);
    h->slab_slots = (int)value;
  } else if (!strcmp(key, "hole_block")) {
    if (value < 0) return fail(PT_ERR_INVALID, "hole_block %lld", (long long)value);
    if (any_input) return fail(PT_ERR_INVALID, "hole_block must be set before any input tensor");
    if (h->d.o != h->o_full) return fail(PT_ERR_INVALID, "hole_block needs pt_create (o_act == o_all)");
    h->hole_block = (int)std::min<int64_t>(value, h->o_full);
    const int oa = h->hole_block > 0 ? std::min(3 * h->hole_block, h->o_full) : h->o_full;
    h->d = make_dims(oa, h->d.v, h->o_full);
    if (h->hole_block > 0 && (h->slab_slots == 0 || h->slab_slots < oa)) h->slab_slots = oa;
    h->slab_set.assign(h->oh(), 0);
    h->slot_of.assign(h->oh(), -1);
  } else if (!strcmp(key, "particle_contraction")) {
    if (value < 1) return fail(PT_ERR_INVALID, "particle_contraction %lld", (long long)value);
    if (any_input) return fail(PT_ERR_INVALID, "particle_contraction must be set before any input tensor");
    if (h->hole_block) return fail(PT_ERR_UNSUPPORTED, "particle_contraction and hole_block are mutually exclusive");
    h->d = make_dims(h->d.o, h->d.v, h->d.ol, (int)value);
  } else if (!strcmp(key, "async_upload")) {
    h->async_upload = value != 0;
  } else if (!strcmp(key, "pin_host")) {
    h->pin_host = value != 0;
  } else if (!strcmp(key, "item_sync")) {
    if (value < 0) return fail(PT_ERR_INVALID, "item_sync %lld", (long long)value);
    h->item_sync = (int)value;   // 0 = off, N = barrier before every N-th item round
  } else if (!strcmp(key, "class_sort")) {
    h->class_sort = value != 0;
  } else if (!strcmp(key, "tile_holes")) {
    if (value < 0) return fail(PT_ERR_INVALID, "tile_holes %lld", (long long)value);
    h->tile_holes = (int)value;
  } else if (!strcmp(key, "debug")) {
    h->debug = (int)value;
  } else if (!strcmp(key, "order")) {
    if (value != 0 && value != 1) return fail(PT_ERR_INVALID, "order %lld", (long long)value);
    h->order = (int)value;
  } else {
    return fail(PT_ERR_INVALID, "pt_set_option: unknown key '%s'", key);
  }
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_set_option");
}

int pt_sync(pt_handle_t h) try {
  if (!h) return fail(PT_ERR_INVALID, "pt_sync: null");
  CU(cudaSetDevice(h->device));
  RC(sync_uploads(h));
  CU(cudaStreamSynchronize(h->stream));
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_sync");
}

int pt_set_eigenenergies(pt_handle_t h, const double* epsi, const double* epsa) try {
  if (!h || !epsi || !epsa) return fail(PT_ERR_INVALID, "pt_set_eigenenergies: null");
  CU(cudaSetDevice(h->device));
  UploadScope up(h);
  if (!h->epsi) CU(h->alloc(&h->epsi, h->d.o));
  if (!h->epsa) CU(h->alloc(&h->epsa, h->d.v));
  if (h->hole_block) h->host_epsi.assign(epsi, epsi + h->o_full);   // sliced per group
  else RC(upload(h, h->epsi, epsi, h->d.o));
  RC(upload(h, h->epsa, epsa, h->d.v));
  RC(up.done());
  h->have_eps = true;
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_set_eigenenergies");
}

int pt_set_singles(pt_handle_t h, const double* t1) try {
  if (!h || !t1) return fail(PT_ERR_INVALID, "pt_set_singles: null");
  CU(cudaSetDevice(h->device));
  UploadScope up(h);
  const size_t n = (size_t)h->d.v * h->d.o;
  if (!h->t1) CU(h->alloc(&h->t1, n));
  if (h->hole_block) h->host_t1.assign(t1, t1 + (size_t)h->d.v * h->o_full);
  else RC(upload(h, h->t1, t1, n));
  RC(up.done());
  h->have_t1 = true;
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_set_singles");
}

int pt_set_pphh(pt_handle_t h, const double* vabij) try {
  if (!h || !vabij) return fail(PT_ERR_INVALID, "pt_set_pphh: null");
  CU(cudaSetDevice(h->device));
  UploadScope up(h);
  const size_t n = (size_t)h->d.v * h->d.v * h->d.o * h->d.o;
  if (!h->pphh) CU(h->alloc(&h->pphh, n));
  if (!h->qsum) CU(h->alloc(&h->qsum, n));
  h->pphh_from_vertex = false;
  if (h->hole_block) {
    h->host_pphh = vabij;   // caller-owned [v,v,o,o]; blocks are staged per group
    pin(h, vabij, (size_t)h->d.v * h->d.v * h->o_full * h->o_full);
    h->cur_holes.clear();
  } else {
    RC(upload(h, h->pphh, vabij, n));
    CU(launch_pphh_symsum(h->pphh, h->qsum, h->d, h->stream));
    h->stats.kernel_launches += 1;
  }
  RC(up.done());
  h->have_pphh = true;
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_set_pphh");
}

int pt_set_singles_pair(pt_handle_t h, const double* t1b, const double* vabij_b) try {
  if (!h || !t1b || !vabij_b) return fail(PT_ERR_INVALID, "pt_set_singles_pair: null");
  if (h->hole_block) return fail(PT_ERR_UNSUPPORTED, "pt_set_singles_pair: not in hole_block mode");
  CU(cudaSetDevice(h->device));
  UploadScope up(h);
  const size_t n1 = (size_t)h->d.v * h->d.o, n = (size_t)h->d.v * h->d.v * h->d.o * h->d.o;
  if (!h->t1b) CU(h->alloc(&h->t1b, n1));
  if (!h->pphhb) CU(h->alloc(&h->pphhb, n));
  if (!h->qsumb) CU(h->alloc(&h->qsumb, n));
  RC(upload(h, h->t1b, t1b, n1));
  RC(upload(h, h->pphhb, vabij_b, n));
  CU(launch_pphh_symsum(h->pphhb, h->qsumb, h->d, h->stream));
  h->stats.kernel_launches += 1;
  RC(up.done());
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_set_singles_pair");
}

int pt_set_doubles(pt_handle_t h, const double* t2) try {
  if (!h || !t2) return fail(PT_ERR_INVALID, "pt_set_doubles: null");
  CU(cudaSetDevice(h->device));
  UploadScope up(h);
  const size_t n = (size_t)h->d.v * h->d.vd * h->d.o * h->d.o;
  if (!h->Tt) CU(h->alloc(&h->Tt, tt_elems(h->d)));
  if (h->hole_block) {
    if (!h->T2h) CU(h->alloc(&h->T2h, t2h_elems(h->d)));
    if (!h->stage_raw) CU(h->alloc(&h->stage_raw, (size_t)h->d.v * h->d.v * std::max((size_t)h->d.o * h->d.o, (size_t)h->o_full)));
    h->host_t2 = t2;        // caller-owned [v,v,o,o]
    pin(h, t2, (size_t)h->d.v * h->d.v * h->o_full * h->o_full);
    h->cur_holes.clear();
    h->have_t2h = true;
  } else {
    StreamScratch<double> tmp;
    double* raw = h->t2_raw;
    if (!raw) {
      if (h->keep_raw) { CU(h->alloc(&h->t2_raw, n)); raw = h->t2_raw; }
      else { CU(tmp.alloc(n, h->stream)); raw = tmp; }
    }
    RC(upload(h, raw, t2, n));
    CU(launch_pack_tt(raw, h->Tt, h->d, h->stream));
    h->stats.kernel_launches += 1;
    if (h->d.ol == h->d.o) {  // the same tensor serves the hole term
      if (!h->T2h) CU(h->alloc(&h->T2h, t2h_elems(h->d)));
      CU(launch_pack_t2h(raw, h->T2h, h->d, h->stream));
      h->stats.kernel_launches += 1;
      h->have_t2h = true;
    }
  }
  RC(up.done());
  h->have_t2 = true;
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_set_doubles");
}

int pt_set_doubles_hole(pt_handle_t h, const double* t2_xl) try {
  if (!h || !t2_xl) return fail(PT_ERR_INVALID, "pt_set_doubles_hole: null");
  if (h->hole_block) return fail(PT_ERR_INVALID, "pt_set_doubles_hole: not used in hole_block mode (pt_set_doubles takes the full tensor)");
  CU(cudaSetDevice(h->device));
  UploadScope up(h);
  const size_t n = (size_t)h->d.v * h->d.v * h->d.o * h->d.ol;   // [v,v,o_act,o_all]
  StreamScratch<double> raw;
  CU(raw.alloc(n, h->stream));
  RC(upload(h, raw, t2_xl, n));
  if (!h->T2h) CU(h->alloc(&h->T2h, t2h_elems(h->d)));
  CU(launch_pack_t2h(raw, h->T2h, h->d, h->stream));
  h->stats.kernel_launches += 1;
  RC(up.done());
  h->have_t2h = true;
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_set_doubles_hole");
}

int pt_set_hhhp(pt_handle_t h, const double* vijka) try {
  if (!h || !vijka) return fail(PT_ERR_INVALID, "pt_set_hhhp: null");
  CU(cudaSetDevice(h->device));
  UploadScope up(h);
  if (!h->Ut) CU(h->alloc(&h->Ut, ut_elems(h->d)));
  h->hhhp_from_vertex = false;
  if (h->hole_block) {
    // the full tensor [o,o,o,v] stays on the device (o^3 v: 6.4 GB at o=100, v=800); Ut is packed per group
    const size_t n = (size_t)h->o_full * h->o_full * h->o_full * h->d.v;
    if (!h->hhhp_raw) CU(h->alloc(&h->hhhp_raw, n));
    RC(upload(h, h->hhhp_raw, vijka, n));
    h->cur_holes.clear();
  } else {
    const size_t n = (size_t)h->d.o * h->d.o * h->d.ol * h->d.v;   // [o,o,o_all,v]
    StreamScratch<double> tmp;
    double* raw = h->hhhp_raw;
    if (!raw) {
      if (h->keep_raw) { CU(h->alloc(&h->hhhp_raw, n)); raw = h->hhhp_raw; }
      else { CU(tmp.alloc(n, h->stream)); raw = tmp; }
    }
    RC(upload(h, raw, vijka, n));
    CU(launch_pack_ut(raw, h->Ut, h->d, nullptr, h->stream));
    h->stats.kernel_launches += 1;
  }
  RC(up.done());
  h->have_hhhp = true;
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_set_hhhp");
}

static int ensure_ppph_buffers(pt_handle_t h, bool need_stage) {
  const size_t slab = (size_t)h->d.v * h->d.v * h->d.vd;
  if (h->blocked() && h->keep_raw) return fail(PT_ERR_INVALID, "slab_slots / hole_block and keep_raw are mutually exclusive");
  if (!h->Vt) {
    CU(h->alloc(&h->Vt, vt_slab_elems(h->d) * (size_t)h->nslots()));
    h->hole_in.assign(h->nslots(), -1);
    h->slot_tick.assign(h->nslots(), 0);
  }
  if (need_stage && !h->keep_raw && !h->slab_stage) CU(h->alloc(&h->slab_stage, slab));
  if (h->keep_raw && !h->ppph_raw) CU(h->alloc(&h->ppph_raw, slab * h->d.o));
  if (h->use_vslot() && !h->d_vslot) CU(h->alloc(&h->d_vslot, (size_t)h->d.o));
  return PT_OK;
}

static void note_slot(pt_handle_t h, int k, int slot) {
  if (h->hole_in[slot] >= 0) h->slot_of[h->hole_in[slot]] = -1;
  h->hole_in[slot] = k;
  h->slot_of[k] = slot;
  h->slot_tick[slot] = ++h->tick;
}

// pack the raw slab `src` (device) of hole k into slot `slot`
static int pack_into_slot(pt_handle_t h, const double* src, int k, int slot) {
  CU(launch_pack_vt_slab(src, h->Vt + vt_slab_elems(h->d) * (size_t)slot, h->d, h->stream));
  h->stats.kernel_launches += 1;
  note_slot(h, k, slot);
  return PT_OK;
}

int pt_set_ppph_slabs(pt_handle_t h, int k0, int k1, const double* slabs) try {
  if (!h || !slabs) return fail(PT_ERR_INVALID, "pt_set_ppph_slabs: null");
  if (k0 < 0 || k1 > h->oh() || k0 >= k1) return fail(PT_ERR_INVALID, "pt_set_ppph_slabs: range [%d,%d) of %d", k0, k1, h->oh());
  if (h->blocked())
    return fail(PT_ERR_INVALID, "pt_set_ppph_slabs: with slab_slots < o the slabs are fetched on demand; "
                                "use pt_set_ppph_host or pt_set_vertex");
  CU(cudaSetDevice(h->device));
  RC(ensure_ppph_buffers(h, true));
  h->ppph_from_vertex = false;
  UploadScope up(h);
  const size_t slab = (size_t)h->d.v * h->d.v * h->d.vd;
  for (int k = k0; k < k1; ++k) {
    double* dst = h->keep_raw ? h->ppph_raw + slab * k : h->slab_stage;
    RC(upload(h, dst, slabs + slab * (size_t)(k - k0), slab));
    RC(pack_into_slot(h, dst, k, k));
    h->slab_set[k] = 1;
  }
  RC(up.done());
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_set_ppph_slabs");
}

int pt_set_ppph_host(pt_handle_t h, const double* vabci) try {
  if (!h || !vabci) return fail(PT_ERR_INVALID, "pt_set_ppph_host: null");
  if (!h->blocked() && !(h->async_upload && !h->keep_raw && !h->hole_block)) return pt_set_ppph_slabs(h, 0, h->oh(), vabci);
  CU(cudaSetDevice(h->device));
  RC(ensure_ppph_buffers(h, true));
  h->lazy_ppph = !h->blocked();   // everything fits: pt_run uploads the slabs its triples touch, in waves
  h->host_ppph = vabci;
  h->ppph_from_vertex = false;    // a vertex given earlier stays resident (PPHH / HHHP may come from it) but no longer feeds the slabs
  pin(h, vabci, (size_t)h->d.v * h->d.v * h->d.vd * h->oh());
  std::fill(h->slab_set.begin(), h->slab_set.end(), 1);
  std::fill(h->slot_of.begin(), h->slot_of.end(), -1);
  std::fill(h->hole_in.begin(), h->hole_in.end(), -1);
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_set_ppph_host");
}

int pt_set_vertex(pt_handle_t h, int nf, int np, const double* gre, const double* gim) try {
  if (!h || !gre || !gim) return fail(PT_ERR_INVALID, "pt_set_vertex: null");
  if (nf < 1 || np < h->oh() + h->d.v) return fail(PT_ERR_INVALID, "pt_set_vertex: nf=%d np=%d (o+v=%d)", nf, np, h->oh() + h->d.v);
  if (h->d.vd != h->d.v) return fail(PT_ERR_UNSUPPORTED, "pt_set_vertex: not with a stacked particle contraction");
  CU(cudaSetDevice(h->device));
  RC(ensure_ppph_buffers(h, false));
  UploadScope up(h);
  const size_t n = (size_t)nf * np * np;
  const size_t gpn = (size_t)vertex_rows_padded(np) * vertex_kp4(nf) * 4;
  if (h->gp) {
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaFree(h->gp));
    h->bytes_alloc -= (double)((size_t)vertex_rows_padded(h->g_np) * vertex_kp4(h->g_nf) * 4 * sizeof(double));
    h->gp = nullptr;
  }
  CU(h->alloc(&h->gp, gpn));
  h->g_nf = nf; h->g_np = np;
  {
    // Re / Im parts are only needed to build the K-major image (fromComplexTensor + the GEMM layout)
    StreamScratch<double> re, im;
    CU(re.alloc(n, h->stream));
    CU(im.alloc(n, h->stream));
    RC(upload(h, re, gre, n));
    RC(upload(h, im, gim, n));
    CU(launch_pack_vertex(re, im, nf, np, h->gp, h->stream));
    h->stats.kernel_launches += 1;
  }
  h->host_ppph = nullptr;   // the vertex is the PPPH source from now on
  h->ppph_from_vertex = true;
  std::fill(h->slab_set.begin(), h->slab_set.end(), 1);
  std::fill(h->slot_of.begin(), h->slot_of.end(), -1);
  std::fill(h->hole_in.begin(), h->hole_in.end(), -1);
  if (!h->blocked()) {
    // all slabs resident: build them now, straight into the packed layout
    const size_t slab = (size_t)h->d.v * h->d.v * h->d.vd;
    for (int k = 0; k < h->oh(); ++k) {
      RC(build_slab_packed(h, k, h->Vt + vt_slab_elems(h->d) * (size_t)k));
      note_slot(h, k, k);
      if (h->keep_raw) RC(build_slab_raw(h, k, h->ppph_raw + slab * k));
    }
  }
  RC(up.done());
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_set_vertex");
}

int pt_use_vertex_integrals(pt_handle_t h) try {
  if (!h) return fail(PT_ERR_INVALID, "pt_use_vertex_integrals: null");
  if (!h->gp) return fail(PT_ERR_MISSING, "Missing argument: CoulombVertex (pt_set_vertex before pt_use_vertex_integrals)");
  if (!h->hole_block && h->d.o != h->d.ol)
    return fail(PT_ERR_UNSUPPORTED, "pt_use_vertex_integrals: not for hole-subset engines (pt_create_ex)");
  CU(cudaSetDevice(h->device));
  UploadScope up(h);
  const size_t n = (size_t)h->d.v * h->d.v * h->d.o * h->d.o;
  if (!h->pphh) CU(h->alloc(&h->pphh, n));
  if (!h->qsum) CU(h->alloc(&h->qsum, n));
  if (!h->Ut) CU(h->alloc(&h->Ut, ut_elems(h->d)));
  const size_t nh = (size_t)h->o_full * h->o_full * h->o_full * h->d.v;
  if (h->hole_block) {
    if (!h->hhhp_raw) CU(h->alloc(&h->hhhp_raw, nh));
    RC(build_hhhp(h, h->hhhp_raw));
    h->host_pphh = nullptr;   // PPHH blocks are built per group
    h->cur_holes.clear();
  } else {
    RC(build_pphh(h, h->d.o, nullptr, h->pphh));
    CU(launch_pphh_symsum(h->pphh, h->qsum, h->d, h->stream));
    StreamScratch<double> tmp;
    double* raw = h->hhhp_raw;
    if (!raw) {
      if (h->keep_raw) { CU(h->alloc(&h->hhhp_raw, nh)); raw = h->hhhp_raw; }
      else { CU(tmp.alloc(nh, h->stream)); raw = tmp; }
    }
    RC(build_hhhp(h, raw));
    CU(launch_pack_ut(raw, h->Ut, h->d, nullptr, h->stream));
    h->stats.kernel_launches += 2;
  }
  h->pphh_from_vertex = h->hhhp_from_vertex = true;
  RC(up.done());
  h->have_pphh = h->have_hhhp = true;
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_use_vertex_integrals");
}

int pt_vertex_integrals(pt_handle_t h, const char* block, double* out) try {
  if (!h || !block || !out) return fail(PT_ERR_INVALID, "pt_vertex_integrals: null");
  if (!h->gp) return fail(PT_ERR_MISSING, "Missing argument: CoulombVertex");
  if (!h->hole_block && h->d.o != h->d.ol)
    return fail(PT_ERR_UNSUPPORTED, "pt_vertex_integrals: not for hole-subset engines (pt_create_ex)");
  CU(cudaSetDevice(h->device));
  RC(sync_uploads(h));
  const size_t o = h->o_full, v = h->d.v;
  if (!strcmp(block, "PPHH")) {
    const size_t n = v * v * o * o;
    StreamScratch<double> tmp;
    CU(tmp.alloc(n, h->stream));
    RC(build_pphh(h, (int)o, nullptr, tmp));
    CU(cudaMemcpyAsync(out, tmp, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    h->stats.bytes_d2h += (double)(n * sizeof(double));
  } else if (!strcmp(block, "HHHP")) {
    const size_t n = o * o * o * v;
    StreamScratch<double> tmp;
    CU(tmp.alloc(n, h->stream));
    RC(build_hhhp(h, tmp));
    CU(cudaMemcpyAsync(out, tmp, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    h->stats.bytes_d2h += (double)(n * sizeof(double));
  } else if (!strcmp(block, "PPPH")) {
    const size_t slab = v * v * v;
    StreamScratch<double> tmp;
    CU(tmp.alloc(slab, h->stream));
    for (size_t z = 0; z < o; ++z) {   // one hole slab at a time: the v^3 o tensor never exists on the device
      RC(build_slab_raw(h, (int)z, tmp));
      CU(cudaMemcpyAsync(out + slab * z, tmp, slab * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      h->stats.bytes_d2h += (double)(slab * sizeof(double));
    }
  } else {
    return fail(PT_ERR_INVALID, "pt_vertex_integrals: unknown block '%s' (PPHH, HHHP, PPPH)", block);
  }
  CU(cudaStreamSynchronize(h->stream));
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_vertex_integrals");
}

// blocked mode: make the slabs of all holes in `need` resident (LRU replacement among the slots
// that hold none of them), then publish the active hole -> slot table to the device
static int ensure_slabs(pt_handle_t h, const std::vector<int>& need) {
  const size_t slab = (size_t)h->d.v * h->d.v * h->d.vd;
  std::vector<char> pinned(h->nslots(), 0);
  for (int z : need)
    if (h->slot_of[z] >= 0) { pinned[h->slot_of[z]] = 1; h->slot_tick[h->slot_of[z]] = ++h->tick; }
  for (int z : need) {
    if (h->slot_of[z] >= 0) continue;
    int victim = -1;
    for (int s = 0; s < h->nslots(); ++s) {
      if (pinned[s]) continue;
      if (h->hole_in[s] < 0) { victim = s; break; }
      if (victim < 0 || h->slot_tick[s] < h->slot_tick[victim]) victim = s;
    }
    if (victim < 0) return fail(PT_ERR_INVALID, "ensure_slabs: %zu slabs needed, %d slots", need.size(), h->nslots());
    if (h->gp && h->ppph_from_vertex) {
      RC(build_slab_packed(h, z, h->Vt + vt_slab_elems(h->d) * (size_t)victim));
      note_slot(h, z, victim);
    } else if (h->host_ppph) {
      RC(upload(h, h->slab_stage, h->host_ppph + slab * (size_t)z, slab));
      RC(pack_into_slot(h, h->slab_stage, z, victim));
    } else {
      return fail(PT_ERR_MISSING, "Missing argument: PPPHCoulombIntegrals (or CoulombVertex)");
    }
    pinned[victim] = 1;
    h->stats.slab_loads += 1;
  }
  return PT_OK;
}

// hole-block mode: stage everything the launch of one group needs for its active holes U (ascending
// holes of the full problem): eps_i[U], T1[:,U], T2[:,:,U,U] -> Tt, T2[:,:,U,:] -> T2h,
// PPHH[:,:,U,U] (+ pair sums), HHHP[U,U,:,:] -> Ut.  T2 / PPHH come from the caller's host tensors in
// contiguous v^2 blocks (one run of consecutive holes = one copy); everything is enqueued on the
// handle's stream behind the previous group's launch, into buffers allocated once.
static int stage_group(pt_handle_t h, const std::vector<int>& U) {
  if (U == h->cur_holes) return PT_OK;
  const int na = (int)U.size(), v = h->d.v, o = h->o_full;
  const size_t vv = (size_t)v * v;
  Dims d = make_dims(na, v, o);
  if (!h->d_hmap) CU(h->alloc(&h->d_hmap, (size_t)h->d.o));
  CU(cudaMemcpyAsync(h->d_hmap, U.data(), (size_t)na * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  // eigenenergies and singles of the active holes (tiny; pageable sources are staged before the call returns)
  std::vector<double> ei(na), t1((size_t)v * na);
  for (int l = 0; l < na; ++l) {
    ei[l] = h->host_epsi[U[l]];
    std::copy(h->host_t1.begin() + (size_t)v * U[l], h->host_t1.begin() + (size_t)v * (U[l] + 1), t1.begin() + (size_t)v * l);
  }
  RC(upload(h, h->epsi, ei.data(), na));
  RC(upload(h, h->t1, t1.data(), t1.size()));
  // runs of consecutive holes in U: [(first local index, length)]
  std::vector<std::pair<int, int>> runs;
  for (int l = 0; l < na;) {
    int e = l + 1;
    while (e < na && U[e] == U[e - 1] + 1) ++e;
    runs.push_back({l, e - l});
    l = e;
  }
  auto gather_pairs = [&](const double* host, double* dev) -> int {   // dev[:,:,xl,yl] = host[:,:,U[xl],U[yl]]
    for (int yl = 0; yl < na; ++yl)
      for (auto& r : runs)
        RC(upload(h, dev + vv * ((size_t)r.first + (size_t)na * yl), host + vv * ((size_t)U[r.first] + (size_t)o * U[yl]),
                  vv * r.second));
    return PT_OK;
  };
  // particle-term doubles T2[a,d,x,y], x, y in U
  RC(gather_pairs(h->host_t2, h->stage_raw));
  CU(launch_pack_tt(h->stage_raw, h->Tt, d, h->stream));
  // hole-term doubles T2[a,b,x,l], x in U, all l: one active hole at a time through the staging buffer
  Dims d1 = d;
  d1.o = 1;
  for (int xl = 0; xl < na; ++xl) {
    CU(cudaMemcpy2DAsync(h->stage_raw, vv * sizeof(double), h->host_t2 + vv * U[xl], vv * o * sizeof(double),
                         vv * sizeof(double), (size_t)o, cudaMemcpyHostToDevice, h->stream));
    h->stats.bytes_h2d += (double)(vv * o * sizeof(double));
    CU(launch_pack_t2h(h->stage_raw, h->T2h + t2h_block_off(d, xl, 0, 0), d1, h->stream));
  }
  // PPHH[b,c,j,k], j, k in U: from the host tensor, or rebuilt from the resident vertex
  if (h->host_pphh) RC(gather_pairs(h->host_pphh, h->pphh));
  else RC(build_pphh(h, na, h->d_hmap, h->pphh));
  CU(launch_pphh_symsum(h->pphh, h->qsum, d, h->stream));
  // hole-term integrals of the active (y, z) pairs from the resident full HHHP tensor
  CU(launch_pack_ut(h->hhhp_raw, h->Ut, d, h->d_hmap, h->stream));
  h->stats.kernel_launches += 3 + na;
  h->stats.groups_staged += 1;
  h->cur_holes = U;
  return PT_OK;
}

static int check_inputs(pt_handle_t h) {
  if (!h->have_eps) return fail(PT_ERR_MISSING, "Missing argument: HoleEigenEnergies/ParticleEigenEnergies");
  if (!h->have_t1) return fail(PT_ERR_MISSING, "Missing argument: CcsdSinglesAmplitudes");
  if (!h->have_t2) return fail(PT_ERR_MISSING, "Missing argument: CcsdDoublesAmplitudes");
  if (!h->have_t2h) return fail(PT_ERR_MISSING, "Missing argument: CcsdDoublesAmplitudes (hole term, pt_set_doubles_hole)");
  if (!h->have_pphh) return fail(PT_ERR_MISSING, "Missing argument: PPHHCoulombIntegrals");
  if (!h->have_hhhp) return fail(PT_ERR_MISSING, "Missing argument: HHHPCoulombIntegrals");
  if (h->slab_set.empty()) return fail(PT_ERR_MISSING, "Missing argument: PPPHCoulombIntegrals slab 0 (or CoulombVertex)");
  for (size_t k = 0; k < h->slab_set.size(); ++k)
    if (!h->slab_set[k]) return fail(PT_ERR_MISSING, "Missing argument: PPPHCoulombIntegrals slab %zu (or CoulombVertex)", k);
  return PT_OK;
}

static FusedParams make_params(pt_handle_t h) {
  FusedParams p{};
  p.d = h->d;
  p.Tt = h->Tt; p.T2h = h->T2h; p.Vt = h->Vt; p.Ut = h->Ut;
  p.t1 = h->t1; p.pphh = h->pphh; p.qsum = h->qsum; p.epsi = h->epsi; p.epsa = h->epsa;
  p.t1b = h->t1b; p.pphhb = h->pphhb; p.qsumb = h->qsumb;
  p.orbits = h->d_orbits;
  p.norbits = h->norbits;
  return p;
}

static int run_naive(pt_handle_t h, const std::vector<Triple>& tr, std::vector<double>& e_out) {
  if (!h->t2_raw || !h->ppph_raw || !h->hhhp_raw)
    return fail(PT_ERR_INVALID, "PT_ENGINE_NAIVE needs option keep_raw=1 set before the tensors");
  if (h->d.ol != h->d.o || h->d.vd != h->d.v) return fail(PT_ERR_UNSUPPORTED, "PT_ENGINE_NAIVE needs o_act == o_all and no stacked contraction");
  const size_t n3 = (size_t)h->d.v * h->d.v * h->d.v;
  StreamScratch<double> w, d_e;
  CU(w.alloc(6 * n3, h->stream));
  CU(d_e.alloc(std::max<size_t>(tr.size(), 1), h->stream));
  CU(cudaMemsetAsync(d_e, 0, tr.size() * sizeof(double), h->stream));
  static const int perm[6][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}, {0, 2, 1}, {2, 0, 1}, {2, 1, 0}};
  for (size_t n = 0; n < tr.size(); ++n) {
    const int hh[3] = {tr[n].i, tr[n].j, tr[n].k};
    const double* w6[6];
    int hp[6][3];
    for (int p = 0; p < 6; ++p) {
      for (int m = 0; m < 3; ++m) hp[p][m] = hh[perm[p][m]];
      int q = 0;
      for (; q < p; ++q)
        if (hp[q][0] == hp[p][0] && hp[q][1] == hp[p][1] && hp[q][2] == hp[p][2]) break;
      w6[p] = w + n3 * p;
      if (q == p) {
        CU(launch_naive_w(h->t2_raw, h->ppph_raw, h->hhhp_raw, h->d, hp[p][0], hp[p][1], hp[p][2],
                          w + n3 * p, h->stream));
        h->stats.kernel_launches += 1;
      }
    }
    CU(launch_naive_energy(w6, h->t1, h->pphh, h->epsi, h->epsa, h->d, hh[0], hh[1], hh[2], d_e + n,
                           h->stream));
    h->stats.kernel_launches += 1;
  }
  e_out.resize(tr.size());
  CU(cudaMemcpyAsync(e_out.data(), d_e, tr.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->stats.bytes_d2h += (double)(tr.size() * sizeof(double));
  return PT_OK;
}

// the triples `tr` (any order) -> e[n]; shared by pt_run and pt_run_list
static int run_triples(pt_handle_t h, const std::vector<Triple>& tr, double* e_triples, double* e_per_triple);

int pt_run(pt_handle_t h, int64_t begin, int64_t end, double* e_triples, double* e_per_triple) try {
  if (!h || !e_triples) return fail(PT_ERR_INVALID, "pt_run: null");
  CU(cudaSetDevice(h->device));
  RC(check_inputs(h));
  std::vector<Triple> all;
  enumerate_triples(h->oh(), all);
  if (begin < 0 || end > (int64_t)all.size() || begin > end)
    return fail(PT_ERR_INVALID, "pt_run: triple range [%lld,%lld) of %zu", (long long)begin, (long long)end, all.size());
  std::vector<Triple> tr(all.begin() + begin, all.begin() + end);
  return run_triples(h, tr, e_triples, e_per_triple);
} catch (...) {
  return pt::on_exception("pt_run");
}

int pt_run_list(pt_handle_t h, int64_t n, const int64_t* triples, double* e_triples, double* e_per_triple) try {
  if (!h || !e_triples || n < 0 || (n > 0 && !triples)) return fail(PT_ERR_INVALID, "pt_run_list: null");
  CU(cudaSetDevice(h->device));
  RC(check_inputs(h));
  std::vector<Triple> all;
  enumerate_triples(h->oh(), all);
  std::vector<Triple> tr((size_t)n);
  for (int64_t m = 0; m < n; ++m) {
    if (triples[m] < 0 || triples[m] >= (int64_t)all.size())
      return fail(PT_ERR_INVALID, "pt_run_list: triple %lld of %zu", (long long)triples[m], all.size());
    tr[(size_t)m] = all[(size_t)triples[m]];
  }
  return run_triples(h, tr, e_triples, e_per_triple);
} catch (...) {
  return pt::on_exception("pt_run_list");
}

static double wall_now() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
#define PT_TRACE(tag)                                                                            \
  do {                                                                                           \
    if (trace) { if (trace == 1) cudaStreamSynchronize(h->stream); fprintf(stderr, "[pt trace] %-18s %9.3f ms\n", tag, (wall_now() - t_trace) * 1e3); } \
  } while (0)

static int run_triples(pt_handle_t h, const std::vector<Triple>& tr, double* e_triples, double* e_per_triple) {
  const int trace = getenv("PT_TRACE") ? atoi(getenv("PT_TRACE")) : 0;   // 1: marks after a stream sync, 2: host-side marks only
  const double t_trace = wall_now();
  RC(sync_uploads(h));
  std::vector<double> e(tr.size(), 0.0);
  CU(cudaEventRecord(h->ev0, h->stream));
  long long weight = 0;
  for (auto& t : tr) weight += triple_weight(t);
  const int o = h->oh();

  if (h->engine == PT_ENGINE_NAIVE) {
    if (h->hole_block) return fail(PT_ERR_UNSUPPORTED, "PT_ENGINE_NAIVE is not available in hole_block mode");
    RC(run_naive(h, tr, e));
  } else {
    // i=j=k triples contribute exactly zero (sum of the spin factors over S3 vanishes), the
    // reference only accumulates rounding noise there (CcsdPerturbativeTriples.cxx:156-158)
    struct Entry { int4 t; int where; long long key; };
    std::vector<Entry> ent;
    // hole-blocked residency: triples are grouped by the hole blocks (I<=J<=K) of width bw they touch;
    // one launch per group with the slabs (and, in hole_block mode, the T2 / PPHH blocks) of those
    // holes resident.  K runs fastest, so consecutive groups share the slabs of blocks I and J.
    const bool grouped = h->blocked() || h->hole_block;
    const int bw = h->hole_block ? h->hole_block : (h->blocked() ? h->nslots() / 3 : o);
    const long long nb = (o + bw - 1) / bw;
    // Lazy upload of a host PPPH tensor (all-resident engine, asynchronous setters): only the slabs of the
    // holes these triples touch are uploaded, in ascending hole order on the copy stream, and the triples
    // run in WAVES -- wave w = the triples whose largest hole k lies among the first c_w missing slabs
    // (i <= j <= k, so their other slabs came earlier) -- so that all but the first small wave's upload
    // hides behind the kernel.  wave_of[z] = wave that makes slab z resident (-1: already there).
    std::vector<int> wave_of, wave_end;   // wave_end[w] = number of missing slabs loaded up to wave w
    std::vector<int> missing;
    if (!grouped && h->lazy_ppph) {
      std::vector<char> touched(o, 0);
      for (auto& t : tr)
        if (triple_class(t) != 3) touched[t.i] = touched[t.j] = touched[t.k] = 1;
      for (int z = 0; z < o; ++z)
        if (touched[z] && h->slot_of[z] < 0) missing.push_back(z);
      wave_of.assign(o, -1);
      const int m = (int)missing.size();
      // waves pay (a few small launches) only when the copies are a visible share of the run: ~40 GB/s of
      // host->device bandwidth against ~30 TFLOP/s of kernel; otherwise one wave = plain need-only upload
      const double copy_s = (double)m * h->d.v * h->d.v * h->d.vd * 8.0 / 4.0e10;
      const double kernel_s = 2.0 * h->d.v * h->d.v * (double)h->d.v * (h->d.vd + h->d.ol) * (double)weight / 3.0e13;
      const bool split = copy_s > 0.02 * kernel_s;
      for (int c : {(m + 7) / 8, (m + 3) / 4, (m + 1) / 2, m})
        if (c > 0 && (split || c == m) && (wave_end.empty() || c > wave_end.back())) wave_end.push_back(c);
      for (int q = 0, w = 0; q < m; ++q) {
        while (q >= wave_end[w]) ++w;
        wave_of[missing[q]] = w;
      }
    }
    for (size_t n = 0; n < tr.size(); ++n) {
      const int c = triple_class(tr[n]);
      if (c == 3) continue;
      long long key = grouped ? block_key(tr[n].i, tr[n].j, tr[n].k, bw, nb) : 0;
      if (!missing.empty()) key = std::max(0, std::max(wave_of[tr[n].i], std::max(wave_of[tr[n].j], wave_of[tr[n].k])));
      ent.push_back({make_int4(tr[n].i, tr[n].j, tr[n].k, c), (int)n, key});
    }
    // equal-cost items: within a group generic triples (i<j<k) first, then i=j, then j=k (stable:
    // consecutive entries keep sharing their leading holes).  CTAs that all run equal-cost items stay
    // in step, so the PPPH tiles they share are read within the L2's residency window.
    if (grouped || !missing.empty() || h->class_sort)
      std::stable_sort(ent.begin(), ent.end(), [&](const Entry& a, const Entry& b) {
        if (a.key != b.key) return a.key < b.key;
        return h->class_sort ? a.t.w < b.t.w : false;
      });
    if (h->tile_holes > 1) {
      const int tb = h->tile_holes;
      std::stable_sort(ent.begin(), ent.end(), [tb](const Entry& a, const Entry& b) {
        if (a.key != b.key) return a.key < b.key;
        const int ka[3] = {a.t.x / tb, a.t.y / tb, a.t.z / tb}, kb[3] = {b.t.x / tb, b.t.y / tb, b.t.z / tb};
        for (int m = 0; m < 3; ++m)
          if (ka[m] != kb[m]) return ka[m] < kb[m];
        return false;
      });
    }
    h->stats.seconds_kernel = 0.0;
    if (!ent.empty()) {
      // groups, their active holes, and the list of LOCAL (active-index) triples
      struct Group { size_t g0, g1; std::vector<int> holes; };
      std::vector<Group> groups;
      std::vector<int4> list(ent.size());
      size_t max_items = 0;
      for (size_t g0 = 0; g0 < ent.size();) {
        size_t g1 = g0;
        while (g1 < ent.size() && ent[g1].key == ent[g0].key) ++g1;
        Group g{g0, g1, {}};
        if (grouped) {
          g.holes = group_holes(ent[g0].t.x, ent[g0].t.y, ent[g0].t.z, bw, o);
        }
        for (size_t n = g0; n < g1; ++n) {
          int4 t = ent[n].t;
          if (h->hole_block) {   // hole of the full problem -> position among the group's active holes
            t.x = (int)(std::lower_bound(g.holes.begin(), g.holes.end(), t.x) - g.holes.begin());
            t.y = (int)(std::lower_bound(g.holes.begin(), g.holes.end(), t.y) - g.holes.begin());
            t.z = (int)(std::lower_bound(g.holes.begin(), g.holes.end(), t.z) - g.holes.begin());
          }
          list[n] = t;
        }
        max_items = std::max(max_items, (g1 - g0) * (size_t)h->norbits);
        groups.push_back(std::move(g));
        g0 = g1;
      }
      PT_TRACE("host lists built");
      // grow-only scratch with headroom: the partitions of one problem differ by a few per cent in their
      // number of triples, and a reallocation in the middle of a run costs a device synchronisation
      const size_t room_l = list.size() + list.size() / 4 + 64, room_i = max_items + max_items / 4 + 64;
      if (list.size() > h->cap_list) CU(h->grow(&h->d_list, &h->cap_list, room_l));
      if (list.size() > h->cap_e) CU(h->grow(&h->d_e, &h->cap_e, room_l));
      if (max_items > h->cap_item) CU(h->grow(&h->d_item, &h->cap_item, room_i));
      PT_TRACE("scratch ready");
      CU(cudaMemcpyAsync(h->d_list, list.data(), list.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
      PT_TRACE("lists ready");
      // waves: slab copies go to the copy stream, into the raw staging area, one event per wave.  Wave 0 is
      // issued now, wave w + 1 right after the kernel of wave w has been launched (with pageable host
      // memory a copy blocks the calling thread, so it must not be issued ahead of that launch).
      std::vector<cudaEvent_t> wev(wave_end.size(), nullptr);
      struct WevGuard { std::vector<cudaEvent_t>& v; ~WevGuard() { for (auto e : v) if (e) cudaEventDestroy(e); } } wevguard{wev};
      size_t waves_issued = 0;
      const size_t slab3 = (size_t)h->d.v * h->d.v * h->d.vd;
      // raw slab q waits for its packing in the (still empty) slot of the NEXT missing slab -- a packed slab
      // is never smaller than a raw one, and slot missing[q+1] is written only by pack(q+1), after pack(q)
      // has read it -- the last one in the one-slab staging buffer: no extra device memory
      auto raw_home = [&](size_t q) -> double* {
        return q + 1 < missing.size() ? h->Vt + vt_slab_elems(h->d) * (size_t)missing[q + 1] : h->slab_stage;
      };
      auto issue_waves = [&](size_t upto) -> int {   // enqueue the copies of all waves < upto
        for (; waves_issued < std::min(upto, wave_end.size()); ++waves_issued) {
          const size_t w = waves_issued;
          for (size_t q = w ? wave_end[w - 1] : 0; q < (size_t)wave_end[w]; ++q) {
            CU(cudaMemcpyAsync(raw_home(q), h->host_ppph + slab3 * (size_t)missing[q], slab3 * sizeof(double),
                               cudaMemcpyHostToDevice, h->copy_stream));
            h->stats.bytes_h2d += (double)(slab3 * sizeof(double));
          }
          CU(cudaEventCreateWithFlags(&wev[w], cudaEventDisableTiming));
          CU(cudaEventRecord(wev[w], h->copy_stream));
        }
        return PT_OK;
      };
      if (!missing.empty()) {
        if (!h->copy_stream) CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        // the copy stream must not overtake earlier work of the compute stream that still reads these areas
        // (the previous run's last pack out of slab_stage)
        CU(cudaEventRecord(h->ev1, h->stream));
        CU(cudaStreamWaitEvent(h->copy_stream, h->ev1, 0));
        RC(issue_waves(1));
      }
      std::vector<cudaEvent_t> kev(2 * groups.size(), nullptr);
      struct EvGuard { std::vector<cudaEvent_t>& v; ~EvGuard() { for (auto e : v) if (e) cudaEventDestroy(e); } } evguard{kev};
      for (auto& ev : kev) CU(cudaEventCreate(&ev));
      for (size_t gi = 0; gi < groups.size(); ++gi) {
        const Group& g = groups[gi];
        FusedParams p = make_params(h);
        if (h->hole_block) {
          RC(stage_group(h, g.holes));
          p.d = make_dims((int)g.holes.size(), h->d.v, h->o_full);
        }
        const int wave = missing.empty() ? -1 : (int)ent[g.g0].key;
        if (wave >= 0) {
          // the slabs up to this wave have arrived (or the stream waits for them): pack them into their slots
          RC(issue_waves((size_t)wave + 1));
          for (int ww = 0; ww <= wave; ++ww) {     // waves without triples of their own still have to be packed
            if (!wev[ww]) continue;
            CU(cudaStreamWaitEvent(h->stream, wev[ww], 0));
            for (int q = ww ? wave_end[ww - 1] : 0; q < wave_end[ww]; ++q)
              RC(pack_into_slot(h, raw_home((size_t)q), missing[q], missing[q]));
            h->stats.slab_loads += wave_end[ww] - (ww ? wave_end[ww - 1] : 0);
            CU(cudaEventDestroy(wev[ww]));
            wev[ww] = nullptr;
          }
        }
        if (grouped) {
          if (h->blocked()) RC(ensure_slabs(h, g.holes));
          // active hole -> slot (pageable source: staged before the call returns; ordered on the
          // stream behind the previous launch that read the table)
          std::vector<int> vs(h->d.o, 0);
          if (h->hole_block) for (size_t l = 0; l < g.holes.size(); ++l) vs[l] = h->slot_of[g.holes[l]];
          else for (int z = 0; z < h->d.o; ++z) vs[z] = std::max(h->slot_of[z], 0);
          CU(cudaMemcpyAsync(h->d_vslot, vs.data(), (size_t)h->d.o * sizeof(int), cudaMemcpyHostToDevice, h->stream));
          p.vslot = h->d_vslot;
        }
        p.triples = h->d_list + g.g0;
        p.ntriples = (int)(g.g1 - g.g0);
        p.order = h->order;
        p.debug = h->debug;
        p.nitems = (long long)(g.g1 - g.g0) * h->norbits;
        p.e_item = h->d_item;
        int grid = h->grid > 0 ? h->grid : h->sm_count;
        if ((long long)grid > p.nitems) grid = (int)p.nitems;
        if (h->item_sync && grid <= h->sm_count && !h->debug) {
          if (!h->d_sync) CU(h->alloc(&h->d_sync, 1));
          CU(cudaMemsetAsync(h->d_sync, 0, sizeof(unsigned int), h->stream));
          p.sync_ctr = h->d_sync;
          p.sync_every = h->item_sync;
        }
        CU(cudaEventRecord(kev[2 * gi], h->stream));
        PT_TRACE("before launch");
        CU(launch_fused(p, grid, h->stream));
        PT_TRACE("launched");
        CU(cudaEventRecord(kev[2 * gi + 1], h->stream));
        if (wave >= 0) RC(issue_waves((size_t)wave + 2));   // the next wave's copies run behind this kernel
        PT_TRACE("fused kernel done");
        CU(launch_reduce_items(h->d_item, p.ntriples, h->norbits, p.order, h->d_e + g.g0, h->stream));
        h->stats.kernel_launches += 2;
      }
      PT_TRACE("reduce done");
      std::vector<double> el(list.size());
      CU(cudaMemcpyAsync(el.data(), h->d_e, list.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      for (size_t gi = 0; gi < groups.size(); ++gi) {
        float kms = 0;
        CU(cudaEventElapsedTime(&kms, kev[2 * gi], kev[2 * gi + 1]));
        h->stats.seconds_kernel += kms * 1e-3;
      }
      h->stats.bytes_h2d += (double)(list.size() * sizeof(int4));
      h->stats.bytes_d2h += (double)(list.size() * sizeof(double));
      for (size_t n = 0; n < list.size(); ++n) e[ent[n].where] = el[n];
    }
  }
  CU(cudaEventRecord(h->ev1, h->stream));
  CU(cudaEventSynchronize(h->ev1));
  PT_TRACE("run done");
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->stats.seconds_run = ms * 1e-3;
  if (h->engine == PT_ENGINE_NAIVE) h->stats.seconds_kernel = h->stats.seconds_run;
  // fixed-order summation in extended precision
  long double sum = 0.0L;
  for (double x : e) sum += (long double)x;
  *e_triples = (double)sum;
  if (e_per_triple) std::copy(e.begin(), e.end(), e_per_triple);
  const double of = h->d.ol, v = h->d.v, vd = h->d.vd;
  h->stats.flops_algorithmic = 2.0 * v * v * v * (vd + of) * (double)weight;
  h->stats.triples_run = (int64_t)tr.size();
  return PT_OK;
}

int pt_get_stats(pt_handle_t h, PtStats* s) try {
  if (!h || !s) return fail(PT_ERR_INVALID, "pt_get_stats: null");
  h->stats.device_bytes = h->bytes_alloc;
  *s = h->stats;
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_get_stats");
}

int pt_debug_w_tile(pt_handle_t h, int x, int y, int z, int ra, int rb, int rc, double* out) try {
  if (!h || !out) return fail(PT_ERR_INVALID, "pt_debug_w_tile: null");
  if (h->hole_block || h->blocked()) return fail(PT_ERR_UNSUPPORTED, "pt_debug_w_tile: all-resident engines only");
  CU(cudaSetDevice(h->device));
  RC(check_inputs(h));
  RC(sync_uploads(h));
  const int o = h->d.o, nr = h->d.nr;
  if (x < 0 || y < 0 || z < 0 || x >= o || y >= o || z >= o || ra < 0 || rb < 0 || rc < 0 || ra >= nr || rb >= nr || rc >= nr)
    return fail(PT_ERR_INVALID, "pt_debug_w_tile: index out of range");
  StreamScratch<double> d_out;
  CU(d_out.alloc(XT_DBL, h->stream));
  FusedParams p = make_params(h);
  WTileJob job{x, y, z, ra, rb, rc};
  CU(launch_w_tile(p, job, d_out, h->stream));
  h->stats.kernel_launches += 1;
  CU(cudaMemcpyAsync(out, d_out, XT_DBL * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_debug_w_tile");
}

int pt_bench_vertex_gemm(pt_handle_t h, int what, int reps, double* seconds, double* flop) try {
  if (!h || !seconds || !flop || reps < 1) return fail(PT_ERR_INVALID, "pt_bench_vertex_gemm: args");
  if (!h->gp || !h->Vt) return fail(PT_ERR_MISSING, "Missing argument: CoulombVertex");
  CU(cudaSetDevice(h->device));
  RC(sync_uploads(h));
  const double v = h->d.v, o = h->o_full, k2 = 2.0 * h->g_nf;
  StreamScratch<double> tmp;
  if (what == 1) CU(tmp.alloc((size_t)(v * v * o * o), h->stream));
  else if (what != 0) return fail(PT_ERR_INVALID, "pt_bench_vertex_gemm: what = 0 (packed PPPH slab) or 1 (PPHH)");
  // slot 0 is rebuilt with its own slab, so the engine's inputs stay intact
  const int z = h->hole_in.empty() || h->hole_in[0] < 0 ? 0 : h->hole_in[0];
  for (int r = 0; r <= reps; ++r) {
    if (r == 1) CU(cudaEventRecord(h->ev0, h->stream));   // r = 0: warm-up
    if (what == 0) RC(build_slab_packed(h, z, h->Vt));
    else RC(build_pphh(h, (int)o, nullptr, tmp));
  }
  CU(cudaEventRecord(h->ev1, h->stream));
  CU(cudaEventSynchronize(h->ev1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  *seconds = ms * 1e-3 / reps;
  *flop = what == 0 ? 2.0 * k2 * v * v * v : 2.0 * k2 * v * v * o * o;
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_bench_vertex_gemm");
}

int pt_bench_fp64(pt_handle_t h, int mode, int warps_per_sm, int iters, double* tflops, double* sm_mhz_est) try {
  if (!h || !tflops) return fail(PT_ERR_INVALID, "pt_bench_fp64: null");
  if (warps_per_sm < 1 || warps_per_sm > 32 || iters < 1) return fail(PT_ERR_INVALID, "pt_bench_fp64: args");
  CU(cudaSetDevice(h->device));
  StreamScratch<double> sink;
  StreamScratch<unsigned long long> cyc;
  const int blocks = h->sm_count;
  CU(sink.alloc(1, h->stream));
  CU(cyc.alloc(blocks, h->stream));
  CU(launch_bench_fp64(mode, blocks, warps_per_sm, iters / 8 + 1, sink, cyc, h->stream));  // warm-up
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaEventRecord(h->ev0, h->stream));
  CU(launch_bench_fp64(mode, blocks, warps_per_sm, iters, sink, cyc, h->stream));
  CU(cudaEventRecord(h->ev1, h->stream));
  CU(cudaEventSynchronize(h->ev1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  const double sec = ms * 1e-3;
  h->stats.kernel_launches += 2;
  std::vector<unsigned long long> hc(blocks);
  CU(cudaMemcpyAsync(hc.data(), cyc, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  unsigned long long mx = 0;
  for (auto c : hc) mx = std::max(mx, c);
  // mode 0: 8 DMMA (256 FMA each) per warp per iteration; mode 1: 16 DFMA per thread per iteration
  const double fma_per_warp_iter = mode == 0 ? 8.0 * 256.0 : 16.0 * 32.0;
  const double flops = 2.0 * fma_per_warp_iter * (double)iters * warps_per_sm * blocks;
  *tflops = flops / sec * 1e-12;
  if (sm_mhz_est) *sm_mhz_est = (double)mx / sec * 1e-6;
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_bench_fp64");
}

}  // extern "C"
