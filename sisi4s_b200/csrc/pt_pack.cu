// pt_pack.cu -- one-time gather of the reference's CTF (column-major) tensors
// into the pre-tiled device layouts described in pt_common.cuh.  This replaces
// CcsdPerturbativeTriples::sliceTensors (reference
// src/algorithms/CcsdPerturbativeTriples.cxx:32-79, SlicedCtfTensor.hpp:33-49):
// instead of 3 o^2 + 4 o + 2 materialised CTF slices, each tensor is re-laid
// once so that every later access is a contiguous TMA bulk copy.
//
// All kernels are HBM-bound copies: reads are coalesced along the fastest raw
// index (a, or b for the PPPH slab), writes are contiguous 32-byte rows.
#include "pt_common.cuh"

namespace pt {

static inline unsigned blocks_for(size_t n, int per) { return (unsigned)((n + per - 1) / per); }

// raw_slab[b + v*(c + v*d)] = Vppph[b,c,d,z], d < vd  ->  Vt slab [Q][R][dc][n][kk]
__global__ void __launch_bounds__(256) pack_vt_slab_kernel(const double* __restrict__ raw,
                                                           double* __restrict__ vt, Dims d) {
  __shared__ double buf[4][256 + 8];
  const int dc = blockIdx.x;
  const int Q = blockIdx.y / d.nr, R = blockIdx.y % d.nr;
  const int tid = threadIdx.x;
  const int bl = tid & 15, cl = tid >> 4;
  const int b = TILE * Q + bl, c = TILE * R + cl;
  const size_t v = d.v;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int dd = 4 * dc + kk;
    double val = 0.0;
    if (b < d.v && c < d.v && dd < d.vd) val = raw[b + v * (c + v * dd)];
    buf[kk][bl * 16 + cl] = val;
  }
  __syncthreads();
  double* dst = vt + ((size_t)(Q * d.nr + R) * d.nk4 + dc) * 1024;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int e = tid + 256 * r;
    dst[e] = buf[e & 3][e >> 2];
  }
}

// Tt[y][x][P][dc][m][kk] = T2[a=16P+m, d=4dc+kk, x, y] (raw [v,vd,o,o]); one thread per (.., m) row
__global__ void __launch_bounds__(256) pack_tt_kernel(const double* __restrict__ t2,
                                                      double* __restrict__ tt, Dims d) {
  const size_t rows = (size_t)d.o * d.o * d.nr * d.nk4 * 16;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows) return;
  const int m = gid & 15;
  size_t r = gid >> 4;
  const int dc = r % d.nk4; r /= d.nk4;
  const int P = r % d.nr; r /= d.nr;
  const int x = r % d.o;
  const int y = r / d.o;
  const int a = TILE * P + m;
  const size_t v = d.v;
  double out[4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int dd = 4 * dc + kk;
    out[kk] = (a < d.v && dd < d.vd) ? t2[a + v * (dd + (size_t)d.vd * (x + (size_t)d.o * y))] : 0.0;
  }
  double2* dst = reinterpret_cast<double2*>(tt + gid * 4);
  dst[0] = make_double2(out[0], out[1]);
  dst[1] = make_double2(out[2], out[3]);
}

// T2h[x][P][Q][lc][g][b8][m][kk] = T2[a=16P+m, b=16Q+8g+b8, x, l=4lc+kk]   (raw [v,v,o,ol]: x active, l all)
__global__ void __launch_bounds__(256) pack_t2h_kernel(const double* __restrict__ t2,
                                                       double* __restrict__ t2h, Dims d) {
  const size_t rows = (size_t)d.o * d.nr * d.nr * d.nl4 * 256;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows) return;
  const int m = gid & 15;
  const int b16 = (gid >> 4) & 15;  // 8g + b8
  size_t r = gid >> 8;
  const int lc = r % d.nl4; r /= d.nl4;
  const int Q = r % d.nr; r /= d.nr;
  const int P = r % d.nr;
  const int x = r / d.nr;
  const int a = TILE * P + m, b = TILE * Q + b16;
  const size_t v = d.v;
  double out[4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int l = 4 * lc + kk;
    out[kk] = (a < d.v && b < d.v && l < d.ol) ? t2[a + v * (b + v * (x + (size_t)d.o * l))] : 0.0;
  }
  double2* dst = reinterpret_cast<double2*>(t2h + gid * 4);
  dst[0] = make_double2(out[0], out[1]);
  dst[1] = make_double2(out[2], out[3]);
}

// Ut[z][y][R][lc][c][kk] = -Vhhhp[y, z, l=4lc+kk, c=16R+c]   (raw [o,o,ol,v]; with `hmap` the source is
// the FULL tensor [ol,ol,ol,v] and the active holes y, z are looked up: hole-blocked mode)
__global__ void __launch_bounds__(256) pack_ut_kernel(const double* __restrict__ hhhp,
                                                      double* __restrict__ ut, Dims d, const int* __restrict__ hmap) {
  const size_t rows = (size_t)d.o * d.o * d.nr * d.nl4 * 16;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows) return;
  const int c16 = gid & 15;
  size_t r = gid >> 4;
  const int lc = r % d.nl4; r /= d.nl4;
  const int R = r % d.nr; r /= d.nr;
  const int y = r % d.o;
  const int z = r / d.o;
  const int c = TILE * R + c16;
  const size_t o = hmap ? d.ol : d.o;
  const size_t ys = hmap ? hmap[y] : y, zs = hmap ? hmap[z] : z;
  double out[4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int l = 4 * lc + kk;
    out[kk] = (c < d.v && l < d.ol) ? -hhhp[ys + o * (zs + o * (l + (size_t)d.ol * c))] : 0.0;
  }
  double2* dst = reinterpret_cast<double2*>(ut + gid * 4);
  dst[0] = make_double2(out[0], out[1]);
  dst[1] = make_double2(out[2], out[3]);
}

// Qsum[b,c,j,k] = Vpphh[b,c,j,k] + Vpphh[c,b,k,j]: the two singles-term operands that the distinct
// hole permutations (j,k) and (k,j) contribute to one point (getSinglesContribution,
// CcsdPerturbativeTriples.cxx:81-85, summed as in :209-211), pre-added once so that the fused
// epilogue fetches ONE value per operand.  Same column-major layout as the raw tensor.
__global__ void __launch_bounds__(256) pphh_symsum_kernel(const double* __restrict__ pphh,
                                                          double* __restrict__ qsum, Dims d) {
  __shared__ double tile[32][33];
  // block: 32 x 32 (b, c) patch of one (j, k) pair; the transposed patch of (k, j) is read coalesced too
  const int v = d.v, o = d.o;
  const int nb = (v + 31) / 32;
  const int pb = blockIdx.x % nb, pc = blockIdx.x / nb;
  const int j = blockIdx.y, k = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const size_t vv = (size_t)v;
  const double* T = pphh + vv * vv * ((size_t)k + (size_t)o * j);  // [c,b,k,j]: transposed source
  const double* S = pphh + vv * vv * ((size_t)j + (size_t)o * k);
  double* Q = qsum + vv * vv * ((size_t)j + (size_t)o * k);
  for (int r = ty; r < 32; r += 8) {
    const int c = 32 * pc + tx, b = 32 * pb + r;  // element T[c + v*b]
    tile[r][tx] = (c < v && b < v) ? T[c + vv * b] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int b = 32 * pb + tx, c = 32 * pc + r;
    if (b < v && c < v) Q[b + vv * c] = S[b + vv * c] + tile[tx][r];
  }
}

// =====================================================================================
// Coulomb integrals from the vertex on the FP64 tensor pipe (reference
// src/algorithms/CoulombIntegralsFromVertex.cxx:402-403 (Vabij), :416-417 (Vijka), :430-431
// (Vabci): V = Re G . Re G + Im G . Im G with the index strings quoted at each launcher).
//
// Every block is a (batched) GEMM  C[m,n] = sum_K A[rowA(m)][K] B[rowB(n)][K]  whose operand
// rows are rows (p,q) of ONE K-major image of the vertex:
//
//   Gp[kc][r][kk] = K-element 4 kc + kk of row r = p + Np q,   K = (Re G[0..NF), Im G[0..NF), 0-pad)
//
// so Re.Re + Im.Im is a single contraction of length 2 NF, a run of 16 rows x 4 K-values is one
// contiguous 512-byte bulk copy, and a DMMA fragment load is `base + lane` (conflict-free), the
// same trick as the (T) kernel's layouts.  pack_vertex_kernel builds the image once per vertex.
//
// vertex_gemm_kernel: CTA tile 64 (M) x 128 (N), K stages of 16 through a 4-deep mbarrier ring
// filled by one producer warp (20 cp.async.bulk copies per stage, one per lane), 8 consumer warps
// with a 32 x 32 register tile each (4 x 4 DMMA.8x8x4 accumulators), two CTAs per SM so one CTA's
// prologue/epilogue overlaps the other's main loop.  Output modes:
//   VG_STRIDED : out[outoff(batch) + m sm + n sn], masked to m < M, n < N
//   VG_PACKED  : straight into the packed PPPH layout Vt[Q][R][dc][16 b][16 c][4 d] of one hole slab
//                (M tile = 16 b x 4 d, N tile = 128 c), zero for padded b/c/d -- no raw slab round trip.
namespace {

__device__ __forceinline__ uint32_t vg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void vg_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void vg_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vg_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void vg_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void vg_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

}  // namespace

constexpr int VG_BM = 64, VG_BN = 128, VG_STAGES = 4;                   // VG_BN: the wide N tile; BN = 64 variant below
constexpr int VG_STAGE_DBL = (VG_BM + VG_BN) * 16;                      // K = 16 per stage: 24 KB (sized for BN = 128)
constexpr int VG_THREADS = 9 * 32;                                      // 8 consumer warps + 1 producer warp
constexpr int VG_SMEM_BYTES = VG_STAGES * VG_STAGE_DBL * 8 + 2 * VG_STAGES * 8;

// Gp[kc][r][kk] from G[F + nf (p + np q)] (Re, Im); rows r >= np^2 (padding) and K >= 2 nf are zero.
// One thread per (r, kc): reads 4 consecutive F (coalesced across the kc-fastest thread index),
// writes one 32-byte row.
__global__ void __launch_bounds__(256) pack_vertex_kernel(const double* __restrict__ gre,
                                                          const double* __restrict__ gim, int nf, long long nrows,
                                                          long long rows_padded, int kp4, double* __restrict__ gp) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows_padded * kp4) return;
  const int kc = (int)(gid % kp4);
  const long long r = gid / kp4;
  double out[4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int kf = 4 * kc + kk;
    double x = 0.0;
    if (r < nrows) {
      if (kf < nf) x = gre[(size_t)r * nf + kf];
      else if (kf < 2 * nf) x = gim[(size_t)r * nf + (kf - nf)];
    }
    out[kk] = x;
  }
  double2* dst = reinterpret_cast<double2*>(gp + ((size_t)kc * rows_padded + r) * 4);
  dst[0] = make_double2(out[0], out[1]);
  dst[1] = make_double2(out[2], out[3]);
}

// BN = 128: warp tile 32 x 32 (4 x 4 DMMA); BN = 64: warp tile 32 x 16 (4 x 2) for N extents that would waste
// more than 10 % of a 128-wide tiling (v = 300: 384 vs 320 columns).
template <int MODE, int BN>
__global__ void __launch_bounds__(VG_THREADS, 2) vertex_gemm_kernel(const VgParams p) {
  constexpr int NJ = BN / 32;   // 8-column DMMA fragments per warp
  extern __shared__ __align__(128) unsigned char vg_smem[];
  double* ring = reinterpret_cast<double*>(vg_smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + VG_STAGES * VG_STAGE_DBL);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ring_u = vg_smem_u32(ring), full_u = vg_smem_u32(bars), empty_u = vg_smem_u32(bars + VG_STAGES);
  if (tid == 0) {
    for (int s = 0; s < VG_STAGES; ++s) {
      vg_mbar_init(full_u + 8 * s, 1);
      vg_mbar_init(empty_u + 8 * s, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- which tile
  const int mt = blockIdx.x, nt = blockIdx.y, batch = blockIdx.z;
  // first vertex row of 16-row group g of the M tile (producer lane l copies group l & 3), of the N tile
  const int g_ = lane & 3;
  long long rowA, rowB;
  long long outoff = 0;
  int Q = 0, dc = 0;
  if (MODE == VG_PACKED) {
    Q = mt / p.nk4;
    dc = mt - Q * p.nk4;
    const int d = min(4 * dc + g_, p.v - 1);                     // padded d: clamped source, masked output
    rowA = p.a_base + 16 * Q + (long long)p.np * (p.a0 + d);     // rows (b, d): G[., a0 + b, a0 + d]
    rowB = p.b_base + (long long)BN * nt;                        // rows (c, z): G[., a0 + c, z]
  } else {
    const int b0 = batch % p.nb0, b1 = batch / p.nb0;
    const long long s0 = p.map0 ? p.map0[b0] : b0, s1 = p.map1 ? p.map1[b1] : b1;
    rowA = p.a_base + p.a_s0 * s0 + p.a_s1 * s1 + (long long)VG_BM * mt + 16 * g_;
    rowB = p.b_base + p.b_s0 * s0 + p.b_s1 * s1 + (long long)BN * nt;
    outoff = p.o_s0 * b0 + p.o_s1 * b1;
  }
  const int nstage = (p.kp4 + 3) >> 2;   // kp4 is a multiple of 4 (K padded to 16)

  if (warp == 8) {
    // ===== producer warp: lane l < 16 copies A group (l & 3) of K-chunk (l >> 2); lanes 16..19 the B chunk
    const bool isA = lane < 16, act = lane < 20;
    const int kcl = isA ? (lane >> 2) : (lane - 16);
    const long long row = isA ? rowA : rowB;
    const uint32_t bytes = isA ? 512u : (uint32_t)(BN * 32);
    const uint32_t soff = isA ? (uint32_t)((kcl * VG_BM + 16 * (lane & 3)) * 32) : (uint32_t)(VG_BM * 128 + kcl * BN * 32);
    const double* src = ((isA || !p.gpb) ? p.gp : p.gpb) + (size_t)row * 4;
    const size_t kstride = (size_t)((isA || !p.gpb) ? p.rows_padded : p.rows_padded_b) * 4;   // doubles per K chunk of 4
    for (int s = 0; s < nstage; ++s) {
      const int slot = s % VG_STAGES;
      const uint32_t ph = (uint32_t)(s / VG_STAGES) & 1u;
      vg_mbar_wait(empty_u + 8 * slot, ph ^ 1u);
      if (lane == 0) vg_mbar_expect_tx(full_u + 8 * slot, (uint32_t)((VG_BM + BN) * 16 * 8));
      __syncwarp();
      if (act)
        vg_bulk_g2s(ring_u + slot * (VG_STAGE_DBL * 8) + soff, src + (size_t)(4 * s + kcl) * kstride, bytes, full_u + 8 * slot);
    }
    return;
  }

  // ===== consumer warps: warp tile 32 x 32
  const int wm = warp & 1, wn = warp >> 1;
  double acc[4][NJ][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int s = 0; s < nstage; ++s) {
    const int slot = s % VG_STAGES;
    const uint32_t ph = (uint32_t)(s / VG_STAGES) & 1u;
    vg_mbar_wait(full_u + 8 * slot, ph);
    const double* As = ring + slot * VG_STAGE_DBL + 32 * wm * 4 + lane;
    const double* Bs = ring + slot * VG_STAGE_DBL + VG_BM * 16 + 8 * NJ * wn * 4 + lane;
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      double af[4], bf[NJ];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = As[kc * VG_BM * 4 + 32 * i];
#pragma unroll
      for (int j = 0; j < NJ; ++j) bf[j] = Bs[kc * BN * 4 + 32 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncwarp();
    if (lane == 0) vg_mbar_arrive(empty_u + 8 * slot);
  }

  // ---- output.  C fragment (i, j): rows 8 i + (lane >> 2), columns 8 j + 2 (lane & 3) + {0, 1}
  const int g = lane >> 2, t2 = 2 * (lane & 3);
  if (MODE == VG_PACKED) {
    // tile row m = 16 kk + bl (kk = d - 4 dc), tile column n = c - 128 nt
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = 32 * wm + 8 * i + g, kk = m >> 4, bl = m & 15;
      const bool mok = (16 * Q + bl < p.v) && (4 * dc + kk < p.v);
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = BN * nt + 8 * NJ * wn + 8 * j + t2 + e;
          const int R = c >> 4, cl = c & 15;
          if (R < p.nr)
            p.out[((size_t)(Q * p.nr + R) * p.nk4 + dc) * 1024 + (size_t)(16 * bl + cl) * 4 + kk] =
                (mok && c < p.v) ? acc[i][j][e] : 0.0;
        }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = VG_BM * mt + 32 * wm + 8 * i + g;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int n = BN * nt + 8 * NJ * wn + 8 * j + t2 + e;
          if (n < p.N) {
            double* o = p.out + outoff + (long long)m * p.sm + (long long)n * p.sn;
            *o = p.accumulate ? p.alpha * acc[i][j][e] + p.beta * *o : p.alpha * acc[i][j][e];
          }
        }
    }
  }
}

// fixed-order second pass of the (T) energy: E_t = sum over the orbits of triple t of the per-item
// partial sums the fused kernel wrote -- bitwise reproducible for any grid size / CTA timing.
__global__ void __launch_bounds__(256) reduce_items_kernel(const double* __restrict__ e_item, int ntriples,
                                                           int norbits, int order, double* __restrict__ e_triple) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntriples) return;
  double s = 0.0;
  if (order == 0) {
    const double* q = e_item + (size_t)t * norbits;
    for (int orb = 0; orb < norbits; ++orb) s += q[orb];
  } else {
    for (int orb = 0; orb < norbits; ++orb) s += e_item[(size_t)orb * ntriples + t];
  }
  e_triple[t] = s;
}


cudaError_t launch_pack_vt_slab(const double* raw_slab, double* vt_slab, Dims d, cudaStream_t s) {
  dim3 grid(d.nk4, d.nr * d.nr);
  pack_vt_slab_kernel<<<grid, 256, 0, s>>>(raw_slab, vt_slab, d);
  return cudaGetLastError();
}
cudaError_t launch_pack_tt(const double* t2, double* tt, Dims d, cudaStream_t s) {
  const size_t rows = (size_t)d.o * d.o * d.nr * d.nk4 * 16;
  pack_tt_kernel<<<blocks_for(rows, 256), 256, 0, s>>>(t2, tt, d);
  return cudaGetLastError();
}
cudaError_t launch_pack_t2h(const double* t2, double* t2h, Dims d, cudaStream_t s) {
  const size_t rows = (size_t)d.o * d.nr * d.nr * d.nl4 * 256;
  pack_t2h_kernel<<<blocks_for(rows, 256), 256, 0, s>>>(t2, t2h, d);
  return cudaGetLastError();
}
cudaError_t launch_pack_ut(const double* hhhp, double* ut, Dims d, const int* hmap, cudaStream_t s) {
  const size_t rows = (size_t)d.o * d.o * d.nr * d.nl4 * 16;
  pack_ut_kernel<<<blocks_for(rows, 256), 256, 0, s>>>(hhhp, ut, d, hmap);
  return cudaGetLastError();
}
cudaError_t launch_pphh_symsum(const double* pphh, double* qsum, Dims d, cudaStream_t s) {
  const int nb = (d.v + 31) / 32;
  dim3 grid(nb * nb, d.o, d.o);
  pphh_symsum_kernel<<<grid, 256, 0, s>>>(pphh, qsum, d);
  return cudaGetLastError();
}
cudaError_t vertex_gemm_configure() {
  for (const void* f : {(const void*)vertex_gemm_kernel<VG_STRIDED, 128>, (const void*)vertex_gemm_kernel<VG_STRIDED, 64>,
                        (const void*)vertex_gemm_kernel<VG_PACKED, 128>, (const void*)vertex_gemm_kernel<VG_PACKED, 64>}) {
    cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, VG_SMEM_BYTES);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
cudaError_t launch_pack_vertex(const double* gre, const double* gim, int nf, int np, double* gp, cudaStream_t s) {
  const long long nrows = (long long)np * np, rp = vertex_rows_padded(np);
  const int kp4 = vertex_kp4(nf);
  pack_vertex_kernel<<<blocks_for((size_t)rp * kp4, 256), 256, 0, s>>>(gre, gim, nf, nrows, rp, kp4, gp);
  return cudaGetLastError();
}
cudaError_t launch_vertex_gemm(const VgParams& p, cudaStream_t s) {
  // the 64-wide N tile when the 128-wide tiling would compute > 10 % more padded columns
  const long long n = p.mode == VG_PACKED ? (long long)p.nr * 16 : p.N;
  const long long w128 = (n + 127) / 128 * 128, w64 = (n + 63) / 64 * 64;
  const bool narrow = w64 * 10 < w128 * 9;
  const unsigned ny = (unsigned)((narrow ? w64 : w128) / (narrow ? 64 : 128));
  if (p.mode == VG_PACKED) {
    dim3 grid((unsigned)(p.nr * p.nk4), ny, 1);
    if (narrow) vertex_gemm_kernel<VG_PACKED, 64><<<grid, VG_THREADS, VG_SMEM_BYTES, s>>>(p);
    else vertex_gemm_kernel<VG_PACKED, 128><<<grid, VG_THREADS, VG_SMEM_BYTES, s>>>(p);
  } else {
    dim3 grid((unsigned)((p.M + VG_BM - 1) / VG_BM), ny, (unsigned)(p.nb0 * p.nb1));
    if (narrow) vertex_gemm_kernel<VG_STRIDED, 64><<<grid, VG_THREADS, VG_SMEM_BYTES, s>>>(p);
    else vertex_gemm_kernel<VG_STRIDED, 128><<<grid, VG_THREADS, VG_SMEM_BYTES, s>>>(p);
  }
  return cudaGetLastError();
}
cudaError_t launch_reduce_items(const double* e_item, int ntriples, int norbits, int order, double* e_triple,
                                cudaStream_t s) {
  if (ntriples <= 0) return cudaSuccess;
  reduce_items_kernel<<<blocks_for((size_t)ntriples, 256), 256, 0, s>>>(e_item, ntriples, norbits, order, e_triple);
  return cudaGetLastError();
}

}  // namespace pt
