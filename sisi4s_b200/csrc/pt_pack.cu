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

// raw_slab[b + v*(c + v*d)] = Vppph[b,c,d,z]  ->  Vt slab [Q][R][dc][n][kk]
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
    if (b < d.v && c < d.v && dd < d.v) val = raw[b + v * (c + v * dd)];
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

// Tt[y][x][P][dc][m][kk] = T2[a=16P+m, d=4dc+kk, x, y]; one thread per (.., m) row
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
    out[kk] = (a < d.v && dd < d.v) ? t2[a + v * (dd + v * (x + (size_t)d.o * y))] : 0.0;
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

// Ut[z][y][R][lc][c][kk] = -Vhhhp[y, z, l=4lc+kk, c=16R+c]   (raw [o,o,ol,v])
__global__ void __launch_bounds__(256) pack_ut_kernel(const double* __restrict__ hhhp,
                                                      double* __restrict__ ut, Dims d) {
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
  const size_t o = d.o;
  double out[4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int l = 4 * lc + kk;
    out[kk] = (c < d.v && l < d.ol) ? -hhhp[y + o * (z + o * (l + (size_t)d.ol * c))] : 0.0;
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

// PPPH slab z from the Coulomb vertex (reference CoulombIntegralsFromVertex.cxx:430-431,
// Vabci["abci"] = ReG["Gac"] ReG["Gbi"] + ImG["Gac"] ImG["Gbi"], particles = last v states):
//   slab[a + v*(b + v*c)] = sum_F G[F,a0+a,a0+c] G[F,a0+b,z]   (Re.Re + Im.Im)
// G is column-major [F + nf*(p + np*q)], i.e. both operands are K-contiguous.  A GEMM with
// M = (a,c) pairs, N = b, K = 2 NF on the FP64 tensor pipe: 128 x 64 output tile per CTA, K chunks
// of 16 staged through shared memory (k-major, padded: conflict-free fragment reads), each of the
// 8 warps owns 32 x 32 of the tile = 4 x 4 DMMA.8x8x4 accumulators.
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

constexpr int VG_M = 128, VG_N = 64, VG_K = 16;

__global__ void __launch_bounds__(256) ppph_slab_from_vertex_kernel(
    const double* __restrict__ gre, const double* __restrict__ gim, int nf, int np, int z,
    double* __restrict__ slab, Dims d) {
  __shared__ double As[VG_K][VG_M + 1];
  __shared__ double Bs[VG_K][VG_N + 1];
  const int v = d.v, a0 = np - v;
  const long long M = (long long)v * v;  // m = a + v*c
  const long long m0 = (long long)blockIdx.x * VG_M;
  const int n0 = blockIdx.y * VG_N;      // n = b
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  double acc[4][4][2] = {};
  // global -> shared assignment: thread loads k = tid & 15 of rows (tid >> 4) + 16 r
  const int lk = tid & 15, lr = tid >> 4;
  size_t arow[VG_M / 16], brow[VG_N / 16];
  bool aok[VG_M / 16], bok[VG_N / 16];
#pragma unroll
  for (int r = 0; r < VG_M / 16; ++r) {
    const long long m = m0 + lr + 16 * r;
    aok[r] = m < M;
    const int a = aok[r] ? (int)(m % v) : 0, c = aok[r] ? (int)(m / v) : 0;
    arow[r] = (size_t)nf * ((a0 + a) + (size_t)np * (a0 + c));
  }
#pragma unroll
  for (int r = 0; r < VG_N / 16; ++r) {
    const int b = n0 + lr + 16 * r;
    bok[r] = b < v;
    brow[r] = (size_t)nf * ((a0 + (bok[r] ? b : 0)) + (size_t)np * z);
  }
  for (int part = 0; part < 2; ++part) {
    const double* G = part == 0 ? gre : gim;
    for (int k0 = 0; k0 < nf; k0 += VG_K) {
      const bool kok = k0 + lk < nf;
#pragma unroll
      for (int r = 0; r < VG_M / 16; ++r) As[lk][lr + 16 * r] = (kok && aok[r]) ? G[arow[r] + k0 + lk] : 0.0;
#pragma unroll
      for (int r = 0; r < VG_N / 16; ++r) Bs[lk][lr + 16 * r] = (kok && bok[r]) ? G[brow[r] + k0 + lk] : 0.0;
      __syncthreads();
#pragma unroll
      for (int ks = 0; ks < VG_K; ks += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) af[i] = As[ks + fk][wm + 8 * i + fr];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = Bs[ks + fk][wn + 8 * j + fr];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
      __syncthreads();
    }
  }
  // C fragment: rows 8i + (lane>>2), columns 8j + 2 (lane&3) + {0,1}
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + wm + 8 * i + fr;
    if (m >= M) continue;
    const int a = (int)(m % v), c = (int)(m / v);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int b = n0 + wn + 8 * j + 2 * fk + e;
        if (b < v) slab[a + (size_t)v * (b + (size_t)v * c)] = acc[i][j][e];
      }
  }
}

static inline unsigned blocks_for(size_t n, int per) { return (unsigned)((n + per - 1) / per); }

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
cudaError_t launch_pack_ut(const double* hhhp, double* ut, Dims d, cudaStream_t s) {
  const size_t rows = (size_t)d.o * d.o * d.nr * d.nl4 * 16;
  pack_ut_kernel<<<blocks_for(rows, 256), 256, 0, s>>>(hhhp, ut, d);
  return cudaGetLastError();
}
cudaError_t launch_pphh_symsum(const double* pphh, double* qsum, Dims d, cudaStream_t s) {
  const int nb = (d.v + 31) / 32;
  dim3 grid(nb * nb, d.o, d.o);
  pphh_symsum_kernel<<<grid, 256, 0, s>>>(pphh, qsum, d);
  return cudaGetLastError();
}
cudaError_t launch_ppph_slab_from_vertex(const double* gre, const double* gim, int nf, int np,
                                         int z, double* raw_slab, Dims d, cudaStream_t s) {
  const long long M = (long long)d.v * d.v;
  dim3 grid((unsigned)((M + VG_M - 1) / VG_M), (unsigned)((d.v + VG_N - 1) / VG_N));
  ppph_slab_from_vertex_kernel<<<grid, 256, 0, s>>>(gre, gim, nf, np, z, raw_slab, d);
  return cudaGetLastError();
}

}  // namespace pt
