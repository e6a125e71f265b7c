// pt_api.cu -- extern "C" layer of libsisi4s_pt (see include/sisi4s_pt.h).
// Host logic only: handle lifetime, uploads + one-time packing, the triple /
// orbit work lists, launches, and the final fixed-order summation.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstring>
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

#define CU(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(e_ == cudaErrorMemoryAllocation ? PT_ERR_NOMEM : PT_ERR_CUDA,          \
                  "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
  } while (0)

// scratch device memory / events of one call: released on every exit path (the CU macro returns early)
template <typename T>
struct Scratch {
  T* p = nullptr;
  Scratch() = default;
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
  ~Scratch() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc((void**)&p, n * sizeof(T)); }
  T* release() { T* q = p; p = nullptr; return q; }
  operator T*() const { return p; }
};
struct ScratchEvent {
  cudaEvent_t e = nullptr;
  ~ScratchEvent() { if (e) cudaEventDestroy(e); }
  cudaError_t create() { return cudaEventCreate(&e); }
  operator cudaEvent_t() const { return e; }
};

}  // namespace

struct PtHandle_ {
  Dims d{};
  int device = 0;
  int sm_count = 0;
  int engine = PT_ENGINE_FUSED;
  int keep_raw = 0;
  int grid = 0;
  int order = 1;
  int item_sync = 0;    // optional item-round barrier between the CTAs of the fused kernel (forces L2 reuse of the PPPH tiles; measured: 6.8x less DRAM traffic but 5 % slower than the free-running equal-cost order)
  int class_sort = 1;   // launch list ordered generic triples first (equal-cost items keep the CTAs in step without a barrier)
  unsigned int* d_sync = nullptr;
  int tile_holes = 0;   // optional: launch list grouped by hole blocks of this width (0 = reference order; measured: no effect on DRAM traffic, profiles/r01h_step_traffic_th*.csv)
  int debug = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // raw device tensors
  double *epsi = nullptr, *epsa = nullptr, *t1 = nullptr, *pphh = nullptr, *qsum = nullptr;
  double *t2_raw = nullptr, *hhhp_raw = nullptr, *ppph_raw = nullptr;  // keep_raw only
  // packed
  double *Tt = nullptr, *T2h = nullptr, *Vt = nullptr, *Ut = nullptr;
  double* slab_stage = nullptr;  // one raw PPPH slab
  std::vector<char> slab_set;    // a source for slab k has been given
  // hole-blocked PPPH residency (option slab_slots = S < o): Vt holds S slab slots, slabs are
  // (re)built on demand from a resident vertex or a caller-owned host tensor
  int slab_slots = 0;
  std::vector<int> slot_of;      // hole -> slot of its packed slab, -1 = not resident
  std::vector<int> hole_in;      // slot -> hole, -1 = free
  std::vector<long long> slot_tick;
  long long tick = 0;
  int* d_vslot = nullptr;
  double *g_re = nullptr, *g_im = nullptr;  // resident CoulombVertex parts (blocked mode)
  int g_nf = 0, g_np = 0;
  const double* host_ppph = nullptr;        // caller-owned PPPHCoulombIntegrals[v,v,v,o]
  int nslots() const { return (slab_slots > 0 && slab_slots < d.o) ? slab_slots : d.o; }
  bool blocked() const { return nslots() < d.o; }
  bool have_t2h = false;
  bool have_eps = false, have_t1 = false, have_t2 = false, have_pphh = false, have_hhhp = false;
  // work lists
  uchar4* d_orbits = nullptr;
  int norbits = 0;
  PtStats stats{};
  double bytes_alloc = 0;

  template <typename T>
  cudaError_t alloc(T** p, size_t n) {
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e == cudaSuccess) bytes_alloc += (double)(n * sizeof(T));
    return e;
  }
};

namespace {

struct Triple { int i, j, k; };

// reference enumeration order, CcsdPerturbativeTriples.cxx:156-158
void enumerate_triples(int o, std::vector<Triple>& out) {
  out.clear();
  for (int i = 0; i < o; ++i)
    for (int j = i; j < o; ++j)
      for (int k = j; k < o; ++k) out.push_back({i, j, k});
}
inline int triple_class(const Triple& t) { return (t.i == t.j ? 1 : 0) + (t.j == t.k ? 2 : 0); }
inline int triple_weight(const Triple& t) {
  static const int w[4] = {6, 3, 3, 1};
  return w[triple_class(t)];
}

struct Timer {
  cudaEvent_t a, b;
  cudaStream_t s;
  Timer(cudaEvent_t a_, cudaEvent_t b_, cudaStream_t s_) : a(a_), b(b_), s(s_) { cudaEventRecord(a, s); }
  double stop() {
    cudaEventRecord(b, s);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms * 1e-3;
  }
};

int upload(pt_handle_t h, double* dst, const double* src, size_t n) {
  CU(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->stats.bytes_h2d += (double)(n * sizeof(double));
  return PT_OK;
}

}  // namespace

extern "C" {

const char* pt_last_error(void) { return g_last_error.c_str(); }
const char* pt_version(void) { return "sisi4s_b200 (T) 0.1 sm_100a"; }

int64_t pt_num_triples(int o) { return (int64_t)o * (o + 1) * (o + 2) / 6; }

int pt_partition(int o, int nranks, int rank, int64_t* begin, int64_t* end) {
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
}

int pt_create(pt_handle_t* out, int o, int v, int device) { return pt_create_ex(out, o, o, v, device); }

int pt_create_ex(pt_handle_t* out, int o, int o_all, int v, int device) {
  if (!out || o < 1 || v < 1 || o_all < o) return fail(PT_ERR_INVALID, "pt_create: need 1 <= o_act <= o_all, v >= 1");
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
  pt_handle_t h = new PtHandle_();
  h->d = make_dims(o, v, o_all);
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->stats.sm_count = prop.multiProcessorCount;
  h->slab_set.assign(o, 0);
  h->slot_of.assign(o, -1);
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&h->ev0));
  CU(cudaEventCreate(&h->ev1));
  CU(fused_configure(nullptr));
  // orbit list: A >= B >= C, C fastest
  std::vector<uchar4> orb;
  for (int A = 0; A < h->d.nr; ++A)
    for (int B = 0; B <= A; ++B)
      for (int C = 0; C <= B; ++C) {
        uchar4 u;
        u.x = (unsigned char)A; u.y = (unsigned char)B; u.z = (unsigned char)C;
        u.w = (unsigned char)((A == B ? 1 : 0) + (B == C ? 2 : 0));
        orb.push_back(u);
      }
  if (h->d.nr > 255) return fail(PT_ERR_UNSUPPORTED, "pt_create: v too large (nr=%d > 255)", h->d.nr);
  // generic orbits (A>B>C, 18 steps per item) first, degenerate ones after: the CTAs of the fused
  // kernel advance in rounds of equal-cost items (item-round barrier), see pt_fused.cu
  std::stable_sort(orb.begin(), orb.end(), [](const uchar4& a, const uchar4& b) { return a.w < b.w; });
  h->norbits = (int)orb.size();
  CU(h->alloc(&h->d_orbits, orb.size()));
  CU(cudaMemcpy(h->d_orbits, orb.data(), orb.size() * sizeof(uchar4), cudaMemcpyHostToDevice));
  *out = h;
  return PT_OK;
}

int pt_destroy(pt_handle_t h) {
  if (!h) return PT_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  double* ptrs[] = {h->epsi, h->epsa, h->t1, h->pphh, h->qsum, h->t2_raw, h->hhhp_raw, h->ppph_raw,
                    h->Tt, h->T2h, h->Vt, h->Ut, h->slab_stage, h->g_re, h->g_im};
  for (double* p : ptrs)
    if (p) cudaFree(p);
  if (h->d_orbits) cudaFree(h->d_orbits);
  if (h->d_vslot) cudaFree(h->d_vslot);
  if (h->d_sync) cudaFree(h->d_sync);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  cudaStreamDestroy(h->stream);
  delete h;
  return PT_OK;
}

int pt_set_option(pt_handle_t h, const char* key, int64_t value) {
  if (!h || !key) return fail(PT_ERR_INVALID, "pt_set_option: null");
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
    h->slab_slots = (int)value;
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
}

int pt_set_eigenenergies(pt_handle_t h, const double* epsi, const double* epsa) {
  if (!h || !epsi || !epsa) return fail(PT_ERR_INVALID, "pt_set_eigenenergies: null");
  CU(cudaSetDevice(h->device));
  Timer tm(h->ev0, h->ev1, h->stream);
  if (!h->epsi) CU(h->alloc(&h->epsi, h->d.o));
  if (!h->epsa) CU(h->alloc(&h->epsa, h->d.v));
  if (int rc = upload(h, h->epsi, epsi, h->d.o)) return rc;
  if (int rc = upload(h, h->epsa, epsa, h->d.v)) return rc;
  h->stats.seconds_upload += tm.stop();
  h->have_eps = true;
  return PT_OK;
}

int pt_set_singles(pt_handle_t h, const double* t1) {
  if (!h || !t1) return fail(PT_ERR_INVALID, "pt_set_singles: null");
  CU(cudaSetDevice(h->device));
  Timer tm(h->ev0, h->ev1, h->stream);
  const size_t n = (size_t)h->d.v * h->d.o;
  if (!h->t1) CU(h->alloc(&h->t1, n));
  if (int rc = upload(h, h->t1, t1, n)) return rc;
  h->stats.seconds_upload += tm.stop();
  h->have_t1 = true;
  return PT_OK;
}

int pt_set_pphh(pt_handle_t h, const double* vabij) {
  if (!h || !vabij) return fail(PT_ERR_INVALID, "pt_set_pphh: null");
  CU(cudaSetDevice(h->device));
  Timer tm(h->ev0, h->ev1, h->stream);
  const size_t n = (size_t)h->d.v * h->d.v * h->d.o * h->d.o;
  if (!h->pphh) CU(h->alloc(&h->pphh, n));
  if (int rc = upload(h, h->pphh, vabij, n)) return rc;
  if (!h->qsum) CU(h->alloc(&h->qsum, n));
  CU(launch_pphh_symsum(h->pphh, h->qsum, h->d, h->stream));
  h->stats.kernel_launches += 1;
  h->stats.seconds_upload += tm.stop();
  h->have_pphh = true;
  return PT_OK;
}

int pt_set_doubles(pt_handle_t h, const double* t2) {
  if (!h || !t2) return fail(PT_ERR_INVALID, "pt_set_doubles: null");
  CU(cudaSetDevice(h->device));
  Timer tm(h->ev0, h->ev1, h->stream);
  const size_t n = (size_t)h->d.v * h->d.v * h->d.o * h->d.o;
  Scratch<double> tmp;
  double* raw = h->t2_raw;
  if (!raw) {
    CU(tmp.alloc(n));
    raw = tmp;
    if (h->keep_raw) { h->t2_raw = tmp.release(); h->bytes_alloc += (double)(n * sizeof(double)); }
  }
  if (int rc = upload(h, raw, t2, n)) return rc;
  if (!h->Tt) CU(h->alloc(&h->Tt, tt_elems(h->d)));
  CU(launch_pack_tt(raw, h->Tt, h->d, h->stream));
  h->stats.kernel_launches += 1;
  if (h->d.ol == h->d.o) {  // the same tensor serves the hole term
    if (!h->T2h) CU(h->alloc(&h->T2h, t2h_elems(h->d)));
    CU(launch_pack_t2h(raw, h->T2h, h->d, h->stream));
    h->stats.kernel_launches += 1;
    h->have_t2h = true;
  }
  h->stats.seconds_upload += tm.stop();
  h->have_t2 = true;
  return PT_OK;
}

int pt_set_doubles_hole(pt_handle_t h, const double* t2_xl) {
  if (!h || !t2_xl) return fail(PT_ERR_INVALID, "pt_set_doubles_hole: null");
  CU(cudaSetDevice(h->device));
  Timer tm(h->ev0, h->ev1, h->stream);
  const size_t n = (size_t)h->d.v * h->d.v * h->d.o * h->d.ol;   // [v,v,o_act,o_all]
  Scratch<double> raw;
  CU(raw.alloc(n));
  if (int rc = upload(h, raw, t2_xl, n)) return rc;
  if (!h->T2h) CU(h->alloc(&h->T2h, t2h_elems(h->d)));
  CU(launch_pack_t2h(raw, h->T2h, h->d, h->stream));
  h->stats.kernel_launches += 1;
  h->stats.seconds_upload += tm.stop();   // synchronises: raw may be released
  h->have_t2h = true;
  return PT_OK;
}

int pt_set_hhhp(pt_handle_t h, const double* vijka) {
  if (!h || !vijka) return fail(PT_ERR_INVALID, "pt_set_hhhp: null");
  CU(cudaSetDevice(h->device));
  Timer tm(h->ev0, h->ev1, h->stream);
  const size_t n = (size_t)h->d.o * h->d.o * h->d.ol * h->d.v;   // [o,o,o_all,v]
  Scratch<double> tmp;
  double* raw = h->hhhp_raw;
  if (!raw) {
    CU(tmp.alloc(n));
    raw = tmp;
    if (h->keep_raw) { h->hhhp_raw = tmp.release(); h->bytes_alloc += (double)(n * sizeof(double)); }
  }
  if (int rc = upload(h, raw, vijka, n)) return rc;
  if (!h->Ut) CU(h->alloc(&h->Ut, ut_elems(h->d)));
  CU(launch_pack_ut(raw, h->Ut, h->d, h->stream));
  h->stats.kernel_launches += 1;
  h->stats.seconds_upload += tm.stop();
  h->have_hhhp = true;
  return PT_OK;
}

static int ensure_ppph_buffers(pt_handle_t h) {
  const size_t slab = (size_t)h->d.v * h->d.v * h->d.v;
  if (h->blocked() && h->keep_raw) return fail(PT_ERR_INVALID, "slab_slots and keep_raw are mutually exclusive");
  if (!h->Vt) {
    CU(h->alloc(&h->Vt, vt_slab_elems(h->d) * (size_t)h->nslots()));
    h->hole_in.assign(h->nslots(), -1);
    h->slot_tick.assign(h->nslots(), 0);
  }
  if (!h->slab_stage) CU(h->alloc(&h->slab_stage, slab));
  if (h->keep_raw && !h->ppph_raw) CU(h->alloc(&h->ppph_raw, slab * h->d.o));
  if (h->blocked() && !h->d_vslot) CU(h->alloc(&h->d_vslot, (size_t)h->d.o));
  return PT_OK;
}

// pack the raw slab `src` (device) of hole k into slot `slot`
static int pack_into_slot(pt_handle_t h, const double* src, int k, int slot) {
  CU(launch_pack_vt_slab(src, h->Vt + vt_slab_elems(h->d) * (size_t)slot, h->d, h->stream));
  h->stats.kernel_launches += 1;
  if (h->hole_in[slot] >= 0) h->slot_of[h->hole_in[slot]] = -1;
  h->hole_in[slot] = k;
  h->slot_of[k] = slot;
  h->slot_tick[slot] = ++h->tick;
  return PT_OK;
}

int pt_set_ppph_slabs(pt_handle_t h, int k0, int k1, const double* slabs) {
  if (!h || !slabs) return fail(PT_ERR_INVALID, "pt_set_ppph_slabs: null");
  if (k0 < 0 || k1 > h->d.o || k0 >= k1) return fail(PT_ERR_INVALID, "pt_set_ppph_slabs: range [%d,%d) of %d", k0, k1, h->d.o);
  if (h->blocked())
    return fail(PT_ERR_INVALID, "pt_set_ppph_slabs: with slab_slots < o the slabs are fetched on demand; "
                                "use pt_set_ppph_host or pt_set_vertex");
  CU(cudaSetDevice(h->device));
  if (int rc = ensure_ppph_buffers(h)) return rc;
  Timer tm(h->ev0, h->ev1, h->stream);
  const size_t slab = (size_t)h->d.v * h->d.v * h->d.v;
  for (int k = k0; k < k1; ++k) {
    double* dst = h->keep_raw ? h->ppph_raw + slab * k : h->slab_stage;
    if (int rc = upload(h, dst, slabs + slab * (size_t)(k - k0), slab)) return rc;
    if (int rc = pack_into_slot(h, dst, k, k)) return rc;
    h->slab_set[k] = 1;
  }
  h->stats.seconds_upload += tm.stop();
  return PT_OK;
}

int pt_set_ppph_host(pt_handle_t h, const double* vabci) {
  if (!h || !vabci) return fail(PT_ERR_INVALID, "pt_set_ppph_host: null");
  if (!h->blocked()) return pt_set_ppph_slabs(h, 0, h->d.o, vabci);
  CU(cudaSetDevice(h->device));
  if (int rc = ensure_ppph_buffers(h)) return rc;
  h->host_ppph = vabci;
  std::fill(h->slab_set.begin(), h->slab_set.end(), 1);
  std::fill(h->slot_of.begin(), h->slot_of.end(), -1);
  std::fill(h->hole_in.begin(), h->hole_in.end(), -1);
  return PT_OK;
}

int pt_set_vertex(pt_handle_t h, int nf, int np, const double* gre, const double* gim) {
  if (!h || !gre || !gim) return fail(PT_ERR_INVALID, "pt_set_vertex: null");
  if (nf < 1 || np < h->d.o + h->d.v) return fail(PT_ERR_INVALID, "pt_set_vertex: nf=%d np=%d (o+v=%d)", nf, np, h->d.o + h->d.v);
  CU(cudaSetDevice(h->device));
  if (int rc = ensure_ppph_buffers(h)) return rc;
  Timer tm(h->ev0, h->ev1, h->stream);
  const size_t n = (size_t)nf * np * np, slab = (size_t)h->d.v * h->d.v * h->d.v;
  if (h->g_re) { CU(cudaFree(h->g_re)); h->g_re = nullptr; }
  if (h->g_im) { CU(cudaFree(h->g_im)); h->g_im = nullptr; }
  CU(cudaMalloc((void**)&h->g_re, n * sizeof(double)));
  CU(cudaMalloc((void**)&h->g_im, n * sizeof(double)));
  h->g_nf = nf; h->g_np = np;
  if (int rc = upload(h, h->g_re, gre, n)) return rc;
  if (int rc = upload(h, h->g_im, gim, n)) return rc;
  if (h->blocked()) {
    // vertex-direct mode: the vertex stays resident, slabs are built when a launch needs them
    h->host_ppph = nullptr;
    h->bytes_alloc += 2.0 * (double)(n * sizeof(double));
    std::fill(h->slab_set.begin(), h->slab_set.end(), 1);
    std::fill(h->slot_of.begin(), h->slot_of.end(), -1);
    std::fill(h->hole_in.begin(), h->hole_in.end(), -1);
    h->stats.seconds_upload += tm.stop();
    return PT_OK;
  }
  for (int k = 0; k < h->d.o; ++k) {
    double* dst = h->keep_raw ? h->ppph_raw + slab * k : h->slab_stage;
    CU(launch_ppph_slab_from_vertex(h->g_re, h->g_im, nf, np, k, dst, h->d, h->stream));
    h->stats.kernel_launches += 1;
    if (int rc = pack_into_slot(h, dst, k, k)) return rc;
    h->slab_set[k] = 1;
  }
  h->stats.seconds_upload += tm.stop();
  CU(cudaFree(h->g_re)); h->g_re = nullptr;
  CU(cudaFree(h->g_im)); h->g_im = nullptr;
  return PT_OK;
}

// blocked mode: make the slabs of all holes in `need` resident (LRU replacement among the slots
// that hold none of them), then publish the hole -> slot table to the device
static int ensure_slabs(pt_handle_t h, const std::vector<int>& need) {
  const size_t slab = (size_t)h->d.v * h->d.v * h->d.v;
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
    if (h->g_re) {
      CU(launch_ppph_slab_from_vertex(h->g_re, h->g_im, h->g_nf, h->g_np, z, h->slab_stage, h->d, h->stream));
      h->stats.kernel_launches += 1;
    } else if (h->host_ppph) {
      if (int rc = upload(h, h->slab_stage, h->host_ppph + slab * (size_t)z, slab)) return rc;
    } else {
      return fail(PT_ERR_MISSING, "Missing argument: PPPHCoulombIntegrals (or CoulombVertex)");
    }
    if (int rc = pack_into_slot(h, h->slab_stage, z, victim)) return rc;
    pinned[victim] = 1;
    h->stats.slab_loads += 1;
  }
  // pageable source: the copy is staged before the call returns, and it is ordered on the
  // stream behind the previous launch that read the table
  CU(cudaMemcpyAsync(h->d_vslot, h->slot_of.data(), (size_t)h->d.o * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  return PT_OK;
}

static int check_inputs(pt_handle_t h) {
  if (!h->have_eps) return fail(PT_ERR_MISSING, "Missing argument: HoleEigenEnergies/ParticleEigenEnergies");
  if (!h->have_t1) return fail(PT_ERR_MISSING, "Missing argument: CcsdSinglesAmplitudes");
  if (!h->have_t2) return fail(PT_ERR_MISSING, "Missing argument: CcsdDoublesAmplitudes");
  if (!h->have_t2h) return fail(PT_ERR_MISSING, "Missing argument: CcsdDoublesAmplitudes (hole term, pt_set_doubles_hole)");
  if (!h->have_pphh) return fail(PT_ERR_MISSING, "Missing argument: PPHHCoulombIntegrals");
  if (!h->have_hhhp) return fail(PT_ERR_MISSING, "Missing argument: HHHPCoulombIntegrals");
  for (int k = 0; k < h->d.o; ++k)
    if (!h->slab_set[k]) return fail(PT_ERR_MISSING, "Missing argument: PPPHCoulombIntegrals slab %d (or CoulombVertex)", k);
  return PT_OK;
}

static FusedParams make_params(pt_handle_t h) {
  FusedParams p{};
  p.d = h->d;
  p.Tt = h->Tt; p.T2h = h->T2h; p.Vt = h->Vt; p.Ut = h->Ut;
  p.t1 = h->t1; p.pphh = h->pphh; p.qsum = h->qsum; p.epsi = h->epsi; p.epsa = h->epsa;
  p.orbits = h->d_orbits;
  p.norbits = h->norbits;
  return p;
}

static int run_naive(pt_handle_t h, const std::vector<Triple>& tr, std::vector<double>& e_out) {
  if (!h->t2_raw || !h->ppph_raw || !h->hhhp_raw)
    return fail(PT_ERR_INVALID, "PT_ENGINE_NAIVE needs option keep_raw=1 set before the tensors");
  if (h->d.ol != h->d.o) return fail(PT_ERR_UNSUPPORTED, "PT_ENGINE_NAIVE needs o_act == o_all");
  const size_t n3 = (size_t)h->d.v * h->d.v * h->d.v;
  Scratch<double> w, d_e;
  CU(w.alloc(6 * n3));
  CU(d_e.alloc(tr.size()));
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

int pt_run(pt_handle_t h, int64_t begin, int64_t end, double* e_triples, double* e_per_triple) {
  if (!h || !e_triples) return fail(PT_ERR_INVALID, "pt_run: null");
  CU(cudaSetDevice(h->device));
  if (int rc = check_inputs(h)) return rc;
  std::vector<Triple> all;
  enumerate_triples(h->d.o, all);
  if (begin < 0 || end > (int64_t)all.size() || begin > end)
    return fail(PT_ERR_INVALID, "pt_run: triple range [%lld,%lld) of %zu", (long long)begin, (long long)end, all.size());
  std::vector<Triple> tr(all.begin() + begin, all.begin() + end);
  return run_triples(h, tr, e_triples, e_per_triple);
}

int pt_run_list(pt_handle_t h, int64_t n, const int64_t* triples, double* e_triples, double* e_per_triple) {
  if (!h || !e_triples || n < 0 || (n > 0 && !triples)) return fail(PT_ERR_INVALID, "pt_run_list: null");
  CU(cudaSetDevice(h->device));
  if (int rc = check_inputs(h)) return rc;
  std::vector<Triple> all;
  enumerate_triples(h->d.o, all);
  std::vector<Triple> tr((size_t)n);
  for (int64_t m = 0; m < n; ++m) {
    if (triples[m] < 0 || triples[m] >= (int64_t)all.size())
      return fail(PT_ERR_INVALID, "pt_run_list: triple %lld of %zu", (long long)triples[m], all.size());
    tr[(size_t)m] = all[(size_t)triples[m]];
  }
  return run_triples(h, tr, e_triples, e_per_triple);
}

static int run_triples(pt_handle_t h, const std::vector<Triple>& tr, double* e_triples, double* e_per_triple) {
  std::vector<double> e(tr.size(), 0.0);
  Timer tm(h->ev0, h->ev1, h->stream);
  long long weight = 0;
  for (auto& t : tr) weight += triple_weight(t);

  if (h->engine == PT_ENGINE_NAIVE) {
    if (int rc = run_naive(h, tr, e)) return rc;
  } else {
    // i=j=k triples contribute exactly zero (sum of the spin factors over S3 vanishes), the
    // reference only accumulates rounding noise there (CcsdPerturbativeTriples.cxx:156-158)
    struct Entry { int4 t; int where; long long key; };
    std::vector<Entry> ent;
    // hole-blocked residency: triples are grouped by the hole blocks (I<=J<=K) of width
    // b = slots/3 they touch; one launch per group with the <= 3b slabs of those blocks resident
    const int bw = h->blocked() ? h->nslots() / 3 : h->d.o;
    for (size_t n = 0; n < tr.size(); ++n) {
      const int c = triple_class(tr[n]);
      if (c == 3) continue;
      const long long nb = (h->d.o + bw - 1) / bw;
      const long long key = h->blocked() ? ((long long)(tr[n].i / bw) * nb + tr[n].j / bw) * nb + tr[n].k / bw : 0;
      ent.push_back({make_int4(tr[n].i, tr[n].j, tr[n].k, c), (int)n, key});
    }
    if (h->blocked())
      std::stable_sort(ent.begin(), ent.end(), [](const Entry& a, const Entry& b) { return a.key < b.key; });
    // L2 locality within a launch: the CTAs that run concurrently take CONSECUTIVE list entries (of
    // one particle-range orbit), so the list is ordered by hole blocks of width tile_holes: any ~150
    // consecutive triples then touch ~12-20 PPPH slabs instead of up to o, and their tiles stay in L2.
    // (E_t is independent of the order; results are scattered back through `where`.)
    // equal-cost items: generic triples (i<j<k) first, then i=j, then j=k (stable: consecutive
    // entries keep sharing their leading holes).  CTAs that all run equal-cost items stay in step, so
    // the PPPH tiles they share are read within the L2's residency window.
    if (h->class_sort)
      std::stable_sort(ent.begin(), ent.end(), [](const Entry& a, const Entry& b) {
        return a.key != b.key ? a.key < b.key : a.t.w < b.t.w;
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
      std::vector<int4> list(ent.size());
      for (size_t n = 0; n < ent.size(); ++n) list[n] = ent[n].t;
      Scratch<int4> d_list;
      Scratch<double> d_e;
      CU(d_list.alloc(list.size()));
      CU(d_e.alloc(list.size()));
      CU(cudaMemcpyAsync(d_list, list.data(), list.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
      CU(cudaMemsetAsync(d_e, 0, list.size() * sizeof(double), h->stream));
      ScratchEvent k0, k1;
      CU(k0.create());
      CU(k1.create());
      for (size_t g0 = 0; g0 < ent.size();) {
        size_t g1 = g0;
        while (g1 < ent.size() && ent[g1].key == ent[g0].key) ++g1;
        FusedParams p = make_params(h);
        if (h->blocked()) {
          std::vector<int> need;
          const int first[3] = {ent[g0].t.x / bw * bw, ent[g0].t.y / bw * bw, ent[g0].t.z / bw * bw};
          for (int m = 0; m < 3; ++m)
            for (int z = first[m]; z < std::min(first[m] + bw, h->d.o); ++z)
              if (std::find(need.begin(), need.end(), z) == need.end()) need.push_back(z);
          if (int rc = ensure_slabs(h, need)) return rc;
          p.vslot = h->d_vslot;
        }
        p.triples = d_list + g0;
        p.ntriples = (int)(g1 - g0);
        p.order = h->order;
        p.debug = h->debug;
        p.nitems = (long long)(g1 - g0) * h->norbits;
        p.e_triple = d_e + g0;
        int grid = h->grid > 0 ? h->grid : h->sm_count;
        if ((long long)grid > p.nitems) grid = (int)p.nitems;
        if (h->item_sync && grid <= h->sm_count && !h->debug) {
          if (!h->d_sync) CU(h->alloc(&h->d_sync, 1));
          CU(cudaMemsetAsync(h->d_sync, 0, sizeof(unsigned int), h->stream));
          p.sync_ctr = h->d_sync;
          p.sync_every = h->item_sync;
        }
        CU(cudaEventRecord(k0, h->stream));
        CU(launch_fused(p, grid, h->stream));
        CU(cudaEventRecord(k1, h->stream));
        CU(cudaEventSynchronize(k1));
        float kms = 0;
        CU(cudaEventElapsedTime(&kms, k0, k1));
        h->stats.seconds_kernel += kms * 1e-3;
        h->stats.kernel_launches += 1;
        g0 = g1;
      }
      std::vector<double> el(list.size());
      CU(cudaMemcpyAsync(el.data(), d_e, list.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      h->stats.bytes_h2d += (double)(list.size() * sizeof(int4));
      h->stats.bytes_d2h += (double)(list.size() * sizeof(double));
      for (size_t n = 0; n < list.size(); ++n) e[ent[n].where] = el[n];
    }
  }
  h->stats.seconds_run = tm.stop();
  if (h->engine == PT_ENGINE_NAIVE) h->stats.seconds_kernel = h->stats.seconds_run;
  // fixed-order summation in extended precision
  long double sum = 0.0L;
  for (double x : e) sum += (long double)x;
  *e_triples = (double)sum;
  if (e_per_triple) std::copy(e.begin(), e.end(), e_per_triple);
  const double o = h->d.ol, v = h->d.v;
  h->stats.flops_algorithmic = 2.0 * v * v * v * (v + o) * (double)weight;
  h->stats.triples_run = (int64_t)tr.size();
  return PT_OK;
}

int pt_get_stats(pt_handle_t h, PtStats* s) {
  if (!h || !s) return fail(PT_ERR_INVALID, "pt_get_stats: null");
  h->stats.device_bytes = h->bytes_alloc;
  *s = h->stats;
  return PT_OK;
}

int pt_debug_w_tile(pt_handle_t h, int x, int y, int z, int ra, int rb, int rc, double* out) {
  if (!h || !out) return fail(PT_ERR_INVALID, "pt_debug_w_tile: null");
  CU(cudaSetDevice(h->device));
  if (int r = check_inputs(h)) return r;
  const int o = h->d.o, nr = h->d.nr;
  if (x < 0 || y < 0 || z < 0 || x >= o || y >= o || z >= o || ra < 0 || rb < 0 || rc < 0 || ra >= nr || rb >= nr || rc >= nr)
    return fail(PT_ERR_INVALID, "pt_debug_w_tile: index out of range");
  Scratch<double> d_out;
  CU(d_out.alloc(XT_DBL));
  FusedParams p = make_params(h);
  WTileJob job{x, y, z, ra, rb, rc};
  CU(launch_w_tile(p, job, d_out, h->stream));
  h->stats.kernel_launches += 1;
  CU(cudaMemcpyAsync(out, d_out, XT_DBL * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return PT_OK;
}

int pt_bench_fp64(pt_handle_t h, int mode, int warps_per_sm, int iters, double* tflops, double* sm_mhz_est) {
  if (!h || !tflops) return fail(PT_ERR_INVALID, "pt_bench_fp64: null");
  if (warps_per_sm < 1 || warps_per_sm > 32 || iters < 1) return fail(PT_ERR_INVALID, "pt_bench_fp64: args");
  CU(cudaSetDevice(h->device));
  Scratch<double> sink;
  Scratch<unsigned long long> cyc;
  const int blocks = h->sm_count;
  CU(sink.alloc(1));
  CU(cyc.alloc(blocks));
  CU(launch_bench_fp64(mode, blocks, warps_per_sm, iters / 8 + 1, sink, cyc, h->stream));  // warm-up
  CU(cudaStreamSynchronize(h->stream));
  Timer tm(h->ev0, h->ev1, h->stream);
  CU(launch_bench_fp64(mode, blocks, warps_per_sm, iters, sink, cyc, h->stream));
  const double sec = tm.stop();
  h->stats.kernel_launches += 2;
  std::vector<unsigned long long> hc(blocks);
  CU(cudaMemcpy(hc.data(), cyc, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  unsigned long long mx = 0;
  for (auto c : hc) mx = std::max(mx, c);
  // mode 0: 8 DMMA (256 FMA each) per warp per iteration; mode 1: 16 DFMA per thread per iteration
  const double fma_per_warp_iter = mode == 0 ? 8.0 * 256.0 : 16.0 * 32.0;
  const double flops = 2.0 * fma_per_warp_iter * (double)iters * warps_per_sm * blocks;
  *tflops = flops / sec * 1e-12;
  if (sm_mhz_est) *sm_mhz_est = (double)mx / sec * 1e-6;
  return PT_OK;
}

}  // extern "C"
