// pt_naive.cu -- literal on-device restatement of the reference loop body
// (src/algorithms/CcsdPerturbativeTriples.cxx:81-117,159-216), one thread per
// (a,b,c) element, no tensor cores, no tiling tricks.  It exists to validate the
// fused kernel on the device at sizes the CPU oracle cannot reach (sampled
// triples at o=40,v=300) and is selected only by PT_ENGINE_NAIVE.  It works on
// the raw (unpacked) tensors, so it also cross-checks the packing kernels.
#include "pt_common.cuh"

namespace pt {

// Permutation<3>(p).images (reference src/math/Permutation.hpp:52-62)
__constant__ int c_perm[6][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}, {0, 2, 1}, {2, 0, 1}, {2, 1, 0}};
// spinAndFermiFactors[invariantElementsCount] (reference :143,202) per sigma_s
__constant__ double c_sf[6] = {8.0, -4.0, 2.0, -4.0, 2.0, -4.0};

// getDoublesContribution (:87-96) with V_bcdk = PPPH[b,c,d,k]:
//   W[a,b,c] = sum_d T2[a,d,x,y] V[b,c,d,z] - sum_l T2[a,b,x,l] Vhhhp[y,z,l,c]
__global__ void __launch_bounds__(256) naive_w_kernel(const double* __restrict__ t2,
                                                      const double* __restrict__ ppph,
                                                      const double* __restrict__ hhhp, Dims d,
                                                      int x, int y, int z, double* __restrict__ w) {
  const size_t v = d.v, o = d.o;
  const size_t n = v * v * v;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n) return;
  const size_t a = gid % v, b = (gid / v) % v, c = gid / (v * v);
  const double* t = t2 + v * v * (x + o * y);       // T2[:, :, x, y]
  const double* vs = ppph + v * v * v * z;          // V[:, :, :, z]
  double acc = 0.0;
  for (size_t dd = 0; dd < v; ++dd) acc = fma(t[a + v * dd], vs[b + v * (c + v * dd)], acc);
  for (size_t l = 0; l < o; ++l)
    acc = fma(-t2[a + v * (b + v * (x + o * l))], hhhp[y + o * (z + o * (l + o * c))], acc);
  w[gid] = acc;
}

struct NaiveEnergyArgs {
  const double* w[6];   // piDVabc[p]; duplicates alias the first occurrence (:170-173)
  int distinct[6];      // givesDistinctIndexPermutation[p]
  int hp[6][3];         // (i,j,k) o pi_p
};

__global__ void __launch_bounds__(256) naive_energy_kernel(NaiveEnergyArgs A,
                                                           const double* __restrict__ t1,
                                                           const double* __restrict__ pphh,
                                                           const double* __restrict__ epsi,
                                                           const double* __restrict__ epsa, Dims d,
                                                           int i, int j, int k, double* e_out) {
  const size_t v = d.v, o = d.o;
  const size_t n = v * v * v;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (gid < n) {
    size_t xx[3] = {gid % v, (gid / v) % v, gid / (v * v)};
    // DVabc["abc"] += piDVabc[p]["abc" o pi]   (:179)
    double X = 0.0;
    for (int p = 0; p < 6; ++p) {
      const size_t c0 = xx[c_perm[p][0]], c1 = xx[c_perm[p][1]], c2 = xx[c_perm[p][2]];
      X += A.w[p][c0 + v * (c1 + v * c2)];
    }
    // divide by the energy denominator (:98-117,183-191)
    const double D = epsi[i] + epsi[j] + epsi[k] - epsa[xx[0]] - epsa[xx[1]] - epsa[xx[2]];
    X = X / D;
    for (int p = 0; p < 6; ++p) {
      if (!A.distinct[p]) continue;
      double Y = 0.0;
      for (int s = 0; s < 6; ++s) {
        // index string ("abc" o sigma) o pi:  coordinate m is x[sigma(pi(m))]
        size_t cc[3];
        for (int m = 0; m < 3; ++m) cc[m] = xx[c_perm[s][c_perm[p][m]]];
        const double wv = A.w[p][cc[0] + v * (cc[1] + v * cc[2])];
        // getSinglesContribution(i o pi)(:81-85): 0.5 T1[a,i'] Vabij[b,c,j',k']
        const double sv = 0.5 * t1[cc[0] + v * A.hp[p][0]] *
                          pphh[cc[1] + v * (cc[2] + v * (A.hp[p][1] + o * A.hp[p][2]))];
        Y += c_sf[s] * (wv + sv);
      }
      e += X * Y;  // energy[""] += DVabc["abc"] * Tabc["abc"]  (:214)
    }
  }
  // block reduction
  __shared__ double red[8];
  for (int off = 16; off > 0; off >>= 1) e += __shfl_down_sync(0xffffffffu, e, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(e_out, s);
  }
}

cudaError_t launch_naive_w(const double* t2, const double* ppph, const double* hhhp, Dims d,
                           int x, int y, int z, double* w, cudaStream_t s) {
  const size_t n = (size_t)d.v * d.v * d.v;
  naive_w_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(t2, ppph, hhhp, d, x, y, z, w);
  return cudaGetLastError();
}

cudaError_t launch_naive_energy(const double* const* w6, const double* t1, const double* pphh,
                                const double* epsi, const double* epsa, Dims d, int i, int j, int k,
                                double* e_out, cudaStream_t s) {
  static const int perm[6][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}, {0, 2, 1}, {2, 0, 1}, {2, 1, 0}};
  const int h[3] = {i, j, k};
  NaiveEnergyArgs A;
  for (int p = 0; p < 6; ++p) {
    for (int m = 0; m < 3; ++m) A.hp[p][m] = h[perm[p][m]];
    int q = 0;
    for (; q < p; ++q)
      if (A.hp[q][0] == A.hp[p][0] && A.hp[q][1] == A.hp[p][1] && A.hp[q][2] == A.hp[p][2]) break;
    A.distinct[p] = (q == p);
    A.w[p] = w6[q];  // q == p for distinct permutations
  }
  const size_t n = (size_t)d.v * d.v * d.v;
  naive_energy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(A, t1, pphh, epsi, epsa, d, i, j, k,
                                                                 e_out);
  return cudaGetLastError();
}

}  // namespace pt
