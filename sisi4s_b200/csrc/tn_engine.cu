// tn_engine.cu -- dense FP64 tensor contractions on the device, behind include/sisi4s_tn.h.
//
// The steps next to the (T) path (N1 all integral blocks of
// CoulombIntegralsFromVertex.cxx:390-560, N3 the CCSD residuum of
// CcsdEnergyFromCoulombIntegralsReference.cxx:29-295) are written in the reference as Cyclops CTF
// index-string contractions, `C["abij"] += alpha * A["acik"] * B["cbkj"]`.  This engine executes one
// such statement on one GPU as transpose - transpose - GEMM (- transpose):
//
//   1. each operand is gathered ONCE into the K-major image  P[kc][row][4]  the integrals-from-vertex
//      GEMM consumes (pt_pack.cu): rows = the operand's free indices, K = the contracted indices,
//      zero-padded -- HBM-bound, one read + one write per element;
//   2. vertex_gemm_kernel (FP64 tensor pipe, DMMA.8x8x4, cp.async.bulk operand ring) multiplies the
//      two images: C[m,n] = alpha * sum_K A[m,K] B[n,K] (+ beta * C);
//   3. when the left-hand side's index order is not (free indices of A)(free indices of B) or the
//      other way round, the product goes through a scratch matrix and a permuted add.
//
// No cuBLAS / cuTENSOR: the GEMM is the library's own kernel.
#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "pt_common.cuh"
#include "../../include/sisi4s_tn.h"

using namespace pt;

namespace {

thread_local std::string g_tn_error;
int tn_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_tn_error = buf;
  return code;
}
}  // namespace

namespace pt {
// library-internal (ccsd_solver.cu): messages for tn_last_error(), and the catch (...) handler of the entry points
int tn_record_error(int code, const char* message) { return tn_fail(code, "%s", message); }
int tn_on_exception(const char* where) {
  try {
    throw;
  } catch (const std::bad_alloc&) {
    return tn_fail(TN_ERR_NOMEM, "%s: host memory exhausted", where);
  } catch (const std::exception& e) {
    return tn_fail(TN_ERR_INVALID, "%s: %s", where, e.what());
  } catch (...) {
    return tn_fail(TN_ERR_INVALID, "%s: unknown C++ exception", where);
  }
}
}  // namespace pt

namespace {

#define TCU(call)                                                                                          \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess)                                                                                 \
      return tn_fail(e_ == cudaErrorMemoryAllocation ? TN_ERR_NOMEM : TN_ERR_CUDA, "%s:%d %s: %s", __FILE__, \
                     __LINE__, #call, cudaGetErrorString(e_));                                             \
  } while (0)
#define TRC(call)                       \
  do {                                  \
    if (int rc_ = (call)) return rc_;   \
  } while (0)

constexpr int TN_MAXD = 8;

// a multi-index (first dimension fastest) and where each of its dimensions lives in some tensor
struct IndexMap {
  int nd;
  long long dim[TN_MAXD];
  long long stride[TN_MAXD];
};

// offset of multi-index number `lin` (first dimension fastest).  Index counts fit 32 bits (checked on the
// host), so the digits are peeled off with 32-bit divisions -- a 64-bit division costs ~80 instructions and
// made the gathers ALU-bound at ~60 GB/s.
__device__ __forceinline__ long long map_offset(const IndexMap& m, unsigned lin) {
  long long off = 0;
#pragma unroll 1
  for (int d = 0; d < m.nd; ++d) {
    const unsigned dim = (unsigned)m.dim[d];
    const unsigned q = lin / dim;
    off += (long long)(lin - q * dim) * m.stride[d];
    lin = q;
  }
  return off;
}

// P[kc][r][kk] = src[row r, K-element 4 kc + kk]; rows >= nrows and K-elements >= K are zero
__global__ void __launch_bounds__(256) tn_pack_kernel(const double* __restrict__ src, double* __restrict__ P,
                                                      IndexMap rows, IndexMap ks, long long nrows, long long K,
                                                      long long rows_padded, int kp4, int row_fast) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows_padded * kp4) return;
  long long r;
  int kc;
  if (rows_padded * kp4 < (1LL << 32)) {          // 32-bit split of the thread index (the common case)
    const unsigned g = (unsigned)gid, rp = (unsigned)rows_padded;
    if (row_fast) { kc = (int)(g / rp); r = g - (unsigned)kc * rp; }
    else { const unsigned q = g / (unsigned)kp4; kc = (int)(g - q * (unsigned)kp4); r = q; }
  } else if (row_fast) { r = gid % rows_padded; kc = (int)(gid / rows_padded); }
  else { kc = (int)(gid % kp4); r = gid / kp4; }
  double out[4] = {0.0, 0.0, 0.0, 0.0};
  if (r < nrows) {
    const long long ro = map_offset(rows, (unsigned)r);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const long long kf = 4LL * kc + kk;
      if (kf < K) out[kk] = src[ro + map_offset(ks, (unsigned)kf)];
    }
  }
  double2* dst = reinterpret_cast<double2*>(P + ((size_t)kc * rows_padded + r) * 4);
  dst[0] = make_double2(out[0], out[1]);
  dst[1] = make_double2(out[2], out[3]);
}

// dst[offset of the same multi-index in dst] = alpha * src[lin] + beta * dst[..]   (src contiguous)
__global__ void __launch_bounds__(256) tn_permute_add_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                                             IndexMap m, long long n, double alpha, double beta) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n) return;
  double* o = dst + map_offset(m, (unsigned)gid);
  *o = beta == 0.0 ? alpha * src[gid] : alpha * src[gid] + beta * *o;
}

// fixed-shape two-pass dot product: bitwise reproducible
constexpr int DOT_BLOCKS = 592;
__global__ void __launch_bounds__(256) tn_dot_partial_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                             long long n, double* __restrict__ part) {
  __shared__ double red[256];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)DOT_BLOCKS * 256) s += a[i] * b[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
__global__ void tn_dot_final_kernel(const double* __restrict__ part, double* __restrict__ out) {
  double s = 0.0;
  for (int i = 0; i < DOT_BLOCKS; ++i) s += part[i];
  *out = s;
}

// estimateAmplitudesFromResiduum (ClusterSinglesDoublesAlgorithm.cxx:302-331) with the excitation
// energies of calculateExcitationEnergies (:343-365): R = -(R - shift T) / (sum eps_a - sum eps_i + shift)
__global__ void __launch_bounds__(256) tn_excitation_divide_kernel(double* __restrict__ R, const double* __restrict__ T,
                                                                   const double* __restrict__ epsi, const double* __restrict__ epsa,
                                                                   int v, int o, int level, long long n, double shift) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n) return;
  long long lin = gid;
  double d = 0.0;
  for (int p = 0; p < level; ++p) { d += epsa[lin % v]; lin /= v; }
  for (int p = 0; p < level; ++p) { d -= epsi[lin % o]; lin /= o; }
  R[gid] = -(R[gid] - shift * T[gid]) / (d + shift);
}

struct TnTensor {
  double* d = nullptr;
  int nd = 0;
  long long len[TN_MAXD] = {0};
  long long n = 0;
};

}  // namespace

struct TnHandle_ {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<TnTensor> t;
  double *wa = nullptr, *wb = nullptr, *wc = nullptr, *dot = nullptr;
  size_t cap_a = 0, cap_b = 0, cap_c = 0;
  double flops = 0, bytes = 0;
  int64_t launches = 0;
};

namespace {

int grow(double** p, size_t* cap, size_t n) {
  if (n <= *cap) return TN_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  const size_t want = n + n / 8;
  TCU(cudaMalloc((void**)p, want * sizeof(double)));
  *cap = want;
  return TN_OK;
}

int get(tn_handle_t h, int id, TnTensor** out) {
  if (!h) return tn_fail(TN_ERR_INVALID, "null handle");
  if (id < 0 || id >= (int)h->t.size() || !h->t[id].d) return tn_fail(TN_ERR_INVALID, "no tensor with id %d", id);
  *out = &h->t[id];
  return TN_OK;
}

// strides of tensor `t` (column-major) for the letters `letters` taken from its index string `idx`
int map_of(const TnTensor& t, const char* idx, const std::string& letters, IndexMap* m) {
  m->nd = (int)letters.size();
  for (int q = 0; q < m->nd; ++q) {
    const char* p = strchr(idx, letters[q]);
    if (!p) return tn_fail(TN_ERR_INVALID, "index '%c' not in \"%s\"", letters[q], idx);
    const int d = (int)(p - idx);
    long long s = 1;
    for (int e = 0; e < d; ++e) s *= t.len[e];
    m->dim[q] = t.len[d];
    m->stride[q] = s;
  }
  return TN_OK;
}

long long count(const IndexMap& m) {
  long long n = 1;
  for (int d = 0; d < m.nd; ++d) n *= m.dim[d];
  return n;
}

int check_indices(const TnTensor& t, const char* idx, const char* what) {
  if (!idx || (int)strlen(idx) != t.nd) return tn_fail(TN_ERR_INVALID, "%s: index string \"%s\" for an order-%d tensor", what, idx ? idx : "", t.nd);
  for (int a = 0; a < t.nd; ++a)
    for (int b = a + 1; b < t.nd; ++b)
      if (idx[a] == idx[b]) return tn_fail(TN_ERR_UNSUPPORTED, "%s: repeated index in \"%s\"", what, idx);
  return TN_OK;
}

inline unsigned nblocks(long long n) { return (unsigned)((n + 255) / 256); }

// gather operand X (index string ix) into its K-major image: rows = `rows` letters, K = `ks` letters
int pack_operand(tn_handle_t h, const TnTensor& X, const char* ix, const std::string& rows, const std::string& ks,
                 double** ws, size_t* cap, long long* rows_padded, int* kp4) {
  IndexMap rm, km;
  TRC(map_of(X, ix, rows, &rm));
  TRC(map_of(X, ix, ks, &km));
  const long long nrows = count(rm), K = count(km);
  *rows_padded = nrows + 256;
  *kp4 = (int)(((K + 15) / 16) * 4);
  const size_t need = (size_t)*rows_padded * *kp4 * 4;
  TRC(grow(ws, cap, need));
  // threads run along the rows when the operand's fastest dimension is a row index (coalesced reads)
  const int row_fast = rows.find(ix[0]) != std::string::npos;
  tn_pack_kernel<<<nblocks(*rows_padded * *kp4), 256, 0, h->stream>>>(X.d, *ws, rm, km, nrows, K, *rows_padded, *kp4, row_fast);
  TCU(cudaGetLastError());
  h->launches += 1;
  h->bytes += 8.0 * (double)X.n + 8.0 * (double)need;
  return TN_OK;
}

}  // namespace

extern "C" {

const char* tn_last_error(void) { return g_tn_error.c_str(); }

int tn_create(tn_handle_t* out, int device) try {
  if (!out) return tn_fail(TN_ERR_INVALID, "tn_create: null");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return tn_fail(TN_ERR_CUDA, "tn_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return tn_fail(TN_ERR_INVALID, "tn_create: device %d of %d", device, ndev);
  TCU(cudaSetDevice(device));
  TCU(vertex_gemm_configure());
  tn_handle_t h = new TnHandle_();
  h->device = device;
  cudaError_t e1 = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  cudaError_t e2 = e1 == cudaSuccess ? cudaMalloc((void**)&h->dot, (DOT_BLOCKS + 1) * sizeof(double)) : e1;
  if (e2 != cudaSuccess) {
    tn_destroy(h);
    return tn_fail(TN_ERR_CUDA, "tn_create: %s", cudaGetErrorString(e2));
  }
  *out = h;
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("tn_create");
}

int tn_destroy(tn_handle_t h) try {
  if (!h) return TN_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& t : h->t)
    if (t.d) cudaFree(t.d);
  for (double* p : {h->wa, h->wb, h->wc, h->dot})
    if (p) cudaFree(p);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("tn_destroy");
}

int tn_tensor(tn_handle_t h, int ndim, const int64_t* lens, int* id) try {
  if (!h || !id || ndim < 0 || ndim > TN_MAXD || (ndim > 0 && !lens)) return tn_fail(TN_ERR_INVALID, "tn_tensor: arguments");
  TCU(cudaSetDevice(h->device));
  TnTensor t;
  t.nd = ndim;
  t.n = 1;
  for (int d = 0; d < ndim; ++d) {
    if (lens[d] < 1) return tn_fail(TN_ERR_INVALID, "tn_tensor: length %lld of dimension %d", (long long)lens[d], d);
    t.len[d] = lens[d];
    t.n *= lens[d];
  }
  if (t.n >= (1LL << 32)) return tn_fail(TN_ERR_UNSUPPORTED, "tn_tensor: %lld elements (the index arithmetic is 32-bit)", t.n);
  TCU(cudaMalloc((void**)&t.d, (size_t)t.n * sizeof(double)));
  TCU(cudaMemsetAsync(t.d, 0, (size_t)t.n * sizeof(double), h->stream));
  // reuse a free slot
  for (size_t q = 0; q < h->t.size(); ++q)
    if (!h->t[q].d) { h->t[q] = t; *id = (int)q; return TN_OK; }
  h->t.push_back(t);
  *id = (int)h->t.size() - 1;
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("tn_tensor");
}

int tn_free(tn_handle_t h, int id) try {
  TnTensor* t;
  TRC(get(h, id, &t));
  TCU(cudaSetDevice(h->device));
  TCU(cudaStreamSynchronize(h->stream));
  TCU(cudaFree(t->d));
  *t = TnTensor();
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("tn_free");
}

int tn_upload(tn_handle_t h, int id, const double* host) try {
  TnTensor* t;
  TRC(get(h, id, &t));
  if (!host) return tn_fail(TN_ERR_INVALID, "tn_upload: null");
  TCU(cudaSetDevice(h->device));
  TCU(cudaMemcpyAsync(t->d, host, (size_t)t->n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  TCU(cudaStreamSynchronize(h->stream));
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("tn_upload");
}

int tn_download(tn_handle_t h, int id, double* host) try {
  TnTensor* t;
  TRC(get(h, id, &t));
  if (!host) return tn_fail(TN_ERR_INVALID, "tn_download: null");
  TCU(cudaSetDevice(h->device));
  TCU(cudaMemcpyAsync(host, t->d, (size_t)t->n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  TCU(cudaStreamSynchronize(h->stream));
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("tn_download");
}

int tn_add(tn_handle_t h, double alpha, int a, const char* ia, double beta, int c, const char* ic) try {
  TnTensor *A, *C;
  TRC(get(h, a, &A));
  TRC(get(h, c, &C));
  TRC(check_indices(*A, ia, "tn_add"));
  TRC(check_indices(*C, ic, "tn_add"));
  if (A->nd != C->nd) return tn_fail(TN_ERR_INVALID, "tn_add: \"%s\" -> \"%s\"", ia, ic);
  if (A->d == C->d && strcmp(ia, ic) != 0) return tn_fail(TN_ERR_UNSUPPORTED, "tn_add: in-place permutation");
  IndexMap m;
  TRC(map_of(*C, ic, std::string(ia), &m));   // where A's dimensions (in A's own order) live in C
  for (int d = 0; d < A->nd; ++d)
    if (m.dim[d] != A->len[d]) return tn_fail(TN_ERR_INVALID, "tn_add: extent of index '%c' differs", ia[d]);
  TCU(cudaSetDevice(h->device));
  tn_permute_add_kernel<<<nblocks(A->n), 256, 0, h->stream>>>(A->d, C->d, m, A->n, alpha, beta);
  TCU(cudaGetLastError());
  h->launches += 1;
  h->bytes += 8.0 * (double)A->n * (beta == 0.0 ? 2.0 : 3.0);
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("tn_add");
}

int tn_contract(tn_handle_t h, double alpha, int a, const char* ia, int b, const char* ib, double beta, int c,
                const char* ic) try {
  TnTensor *A, *B, *C;
  TRC(get(h, a, &A));
  TRC(get(h, b, &B));
  TRC(get(h, c, &C));
  TRC(check_indices(*A, ia, "tn_contract"));
  TRC(check_indices(*B, ib, "tn_contract"));
  TRC(check_indices(*C, ic, "tn_contract"));
  if (C->d == A->d || C->d == B->d) return tn_fail(TN_ERR_UNSUPPORTED, "tn_contract: the result aliases an operand");
  // index classes (Einstein convention of the reference's CTF strings)
  std::string Ms, Ns, Ks;
  for (const char* p = ic; *p; ++p) {
    const bool inA = strchr(ia, *p), inB = strchr(ib, *p);
    if (inA && inB) return tn_fail(TN_ERR_UNSUPPORTED, "tn_contract: index '%c' appears in both operands and the result", *p);
    if (!inA && !inB) return tn_fail(TN_ERR_INVALID, "tn_contract: result index '%c' in neither operand", *p);
    (inA ? Ms : Ns) += *p;
  }
  for (const char* p = ia; *p; ++p) {
    if (strchr(ic, *p)) continue;
    if (!strchr(ib, *p)) return tn_fail(TN_ERR_UNSUPPORTED, "tn_contract: index '%c' of \"%s\" is neither contracted nor kept", *p, ia);
    Ks += *p;
  }
  for (const char* p = ib; *p; ++p)
    if (!strchr(ic, *p) && !strchr(ia, *p))
      return tn_fail(TN_ERR_UNSUPPORTED, "tn_contract: index '%c' of \"%s\" is neither contracted nor kept", *p, ib);
  // extents must agree
  IndexMap ma, mb, mc;
  TRC(map_of(*A, ia, Ks, &ma));
  TRC(map_of(*B, ib, Ks, &mb));
  for (int d = 0; d < ma.nd; ++d)
    if (ma.dim[d] != mb.dim[d]) return tn_fail(TN_ERR_INVALID, "tn_contract: extent of contracted index '%c' differs", Ks[d]);
  TRC(map_of(*A, ia, Ms, &ma));
  TRC(map_of(*C, ic, Ms, &mc));
  for (int d = 0; d < ma.nd; ++d)
    if (ma.dim[d] != mc.dim[d]) return tn_fail(TN_ERR_INVALID, "tn_contract: extent of index '%c' differs", Ms[d]);
  TRC(map_of(*B, ib, Ns, &mb));
  TRC(map_of(*C, ic, Ns, &mc));
  for (int d = 0; d < mb.nd; ++d)
    if (mb.dim[d] != mc.dim[d]) return tn_fail(TN_ERR_INVALID, "tn_contract: extent of index '%c' differs", Ns[d]);
  TCU(cudaSetDevice(h->device));

  // the GEMM writes C directly when the result is ordered (M indices)(N indices) -- or (N)(M), with the
  // operands' roles swapped; otherwise through a scratch matrix and a permuted add
  const std::string sc(ic);
  const TnTensor *L = A, *R = B;
  const char *il = ia, *ir = ib;
  std::string Ls = Ms, Rs = Ns;
  bool direct = sc == Ms + Ns;
  if (!direct && sc == Ns + Ms) {
    std::swap(L, R);
    std::swap(il, ir);
    std::swap(Ls, Rs);
    direct = true;
  }
  long long rpl, rpr;
  int kpl, kpr;
  TRC(pack_operand(h, *L, il, Ls, Ks, &h->wa, &h->cap_a, &rpl, &kpl));
  TRC(pack_operand(h, *R, ir, Rs, Ks, &h->wb, &h->cap_b, &rpr, &kpr));
  const long long M = rpl - 256, N = rpr - 256;
  if (M > 2147483647LL || N > 2147483647LL || M * N >= (1LL << 32))
    return tn_fail(TN_ERR_UNSUPPORTED, "tn_contract: result of %lld x %lld elements (the index arithmetic is 32-bit)", M, N);
  if ((N + 63) / 64 > 65535) return tn_fail(TN_ERR_UNSUPPORTED, "tn_contract: free dimension of the right operand too large (%lld)", N);
  VgParams p{};
  p.mode = VG_STRIDED;
  p.gp = h->wa; p.rows_padded = rpl;
  p.gpb = h->wb; p.rows_padded_b = rpr;
  p.kp4 = kpl;
  p.nb0 = p.nb1 = 1;
  p.M = (int)M; p.N = (int)N;
  p.sm = 1; p.sn = M;
  if (direct) {
    p.out = C->d;
    p.alpha = alpha; p.beta = beta; p.accumulate = beta != 0.0;
  } else {
    TRC(grow(&h->wc, &h->cap_c, (size_t)(M * N)));
    p.out = h->wc;
    p.alpha = 1.0; p.beta = 0.0; p.accumulate = 0;
  }
  TCU(launch_vertex_gemm(p, h->stream));
  h->launches += 1;
  long long K = 1;
  IndexMap mk;
  TRC(map_of(*A, ia, Ks, &mk));
  K = count(mk);
  h->flops += 2.0 * (double)M * (double)N * (double)K;
  if (!direct) {
    IndexMap m;
    TRC(map_of(*C, ic, Ls + Rs, &m));
    tn_permute_add_kernel<<<nblocks(M * N), 256, 0, h->stream>>>(h->wc, C->d, m, M * N, alpha, beta);
    TCU(cudaGetLastError());
    h->launches += 1;
    h->bytes += 8.0 * (double)(M * N) * 3.0;
  }
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("tn_contract");
}

int tn_dot(tn_handle_t h, int a, int b, double* out) try {
  TnTensor *A, *B;
  TRC(get(h, a, &A));
  TRC(get(h, b, &B));
  if (!out || A->n != B->n) return tn_fail(TN_ERR_INVALID, "tn_dot: sizes %lld and %lld", (long long)A->n, (long long)B->n);
  TCU(cudaSetDevice(h->device));
  tn_dot_partial_kernel<<<DOT_BLOCKS, 256, 0, h->stream>>>(A->d, B->d, A->n, h->dot);
  tn_dot_final_kernel<<<1, 1, 0, h->stream>>>(h->dot, h->dot + DOT_BLOCKS);
  TCU(cudaGetLastError());
  h->launches += 2;
  TCU(cudaMemcpyAsync(out, h->dot + DOT_BLOCKS, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  TCU(cudaStreamSynchronize(h->stream));
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("tn_dot");
}

int tn_excitation_divide(tn_handle_t h, int r, int t, int epsi, int epsa, double shift) try {
  TnTensor *R, *T, *Ei, *Ea;
  TRC(get(h, r, &R));
  TRC(get(h, t, &T));
  TRC(get(h, epsi, &Ei));
  TRC(get(h, epsa, &Ea));
  const int level = R->nd / 2;
  if (R->nd % 2 || level < 1 || R->n != T->n || Ei->nd != 1 || Ea->nd != 1)
    return tn_fail(TN_ERR_INVALID, "tn_excitation_divide: shapes");
  for (int p = 0; p < level; ++p)
    if (R->len[p] != Ea->len[0] || R->len[level + p] != Ei->len[0])
      return tn_fail(TN_ERR_INVALID, "tn_excitation_divide: residuum must be [v,..,o,..]");
  TCU(cudaSetDevice(h->device));
  tn_excitation_divide_kernel<<<nblocks(R->n), 256, 0, h->stream>>>(R->d, T->d, Ei->d, Ea->d, (int)Ea->len[0],
                                                                    (int)Ei->len[0], level, R->n, shift);
  TCU(cudaGetLastError());
  h->launches += 1;
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("tn_excitation_divide");
}

int tn_get_stats(tn_handle_t h, double* flops, double* bytes, int64_t* launches) try {
  if (!h) return tn_fail(TN_ERR_INVALID, "tn_get_stats: null");
  if (flops) *flops = h->flops;
  if (bytes) *bytes = h->bytes;
  if (launches) *launches = h->launches;
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("tn_get_stats");
}

}  // extern "C"
