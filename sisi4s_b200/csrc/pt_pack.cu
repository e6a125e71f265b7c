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

// T2h[x][P][Q][lc][g][b8][m][kk] = T2[a=16P+m, b=16Q+8g+b8, x, l=4lc+kk]
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
    out[kk] = (a < d.v && b < d.v && l < d.o) ? t2[a + v * (b + v * (x + (size_t)d.o * l))] : 0.0;
  }
  double2* dst = reinterpret_cast<double2*>(t2h + gid * 4);
  dst[0] = make_double2(out[0], out[1]);
  dst[1] = make_double2(out[2], out[3]);
}

// Ut[z][y][R][lc][c][kk] = -Vhhhp[y, z, l=4lc+kk, c=16R+c]
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
    out[kk] = (c < d.v && l < d.o) ? -hhhp[y + o * (z + o * (l + o * c))] : 0.0;
  }
  double2* dst = reinterpret_cast<double2*>(ut + gid * 4);
  dst[0] = make_double2(out[0], out[1]);
  dst[1] = make_double2(out[2], out[3]);
}

// PPPH slab z from the Coulomb vertex (reference CoulombIntegralsFromVertex.cxx:430-431,
// Vabci["abci"] = ReG["Gac"] ReG["Gbi"] + ImG["Gac"] ImG["Gbi"], particles = last v states):
//   slab[a + v*(b + v*c)] = sum_F G[F,a0+a,a0+c] G[F,a0+b,z]   (Re.Re + Im.Im)
// G is column-major [F + nf*(p + np*q)].  64x64 output tile, K chunks of 16.
__global__ void __launch_bounds__(256) ppph_slab_from_vertex_kernel(
    const double* __restrict__ gre, const double* __restrict__ gim, int nf, int np, int z,
    double* __restrict__ slab, Dims d) {
  __shared__ double As[16][64 + 1];
  __shared__ double Bs[16][64 + 1];
  const int v = d.v, a0 = np - v;
  const long long M = (long long)v * v;  // m = a + v*c
  const long long m0 = (long long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;        // n = b
  const int tid = threadIdx.x;
  const int tm = (tid & 15) * 4, tn = (tid >> 4) * 4;
  double acc[4][4] = {};
  for (int part = 0; part < 2; ++part) {
    const double* G = part == 0 ? gre : gim;
    for (int k0 = 0; k0 < nf; k0 += 16) {
      // 64 rows x 16 k per operand: thread loads 4 elements of each
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int e = tid + 256 * r;  // 0..1023
        const int kk = e & 15, row = e >> 4;
        const int k = k0 + kk;
        double av = 0.0, bv = 0.0;
        const long long m = m0 + row;
        if (k < nf && m < M) {
          const int a = (int)(m % v), c = (int)(m / v);
          av = G[k + (size_t)nf * ((a0 + a) + (size_t)np * (a0 + c))];
        }
        const int b = n0 + row;
        if (k < nf && b < v) bv = G[k + (size_t)nf * ((a0 + b) + (size_t)np * z)];
        As[kk][row] = av;
        Bs[kk][row] = bv;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        double a4[4], b4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a4[u] = As[kk][tm + u]; b4[u] = Bs[kk][tn + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int w = 0; w < 4; ++w) acc[u][w] = fma(a4[u], b4[w], acc[u][w]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long m = m0 + tm + u;
    if (m >= M) continue;
    const int a = (int)(m % v), c = (int)(m / v);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int b = n0 + tn + w;
      if (b < v) slab[a + (size_t)v * (b + (size_t)v * c)] = acc[u][w];
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
cudaError_t launch_ppph_slab_from_vertex(const double* gre, const double* gim, int nf, int np,
                                         int z, double* raw_slab, Dims d, cudaStream_t s) {
  const long long M = (long long)d.v * d.v;
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((d.v + 63) / 64));
  ppph_slab_from_vertex_kernel<<<grid, 256, 0, s>>>(gre, gim, nf, np, z, raw_slab, d);
  return cudaGetLastError();
}

}  // namespace pt
