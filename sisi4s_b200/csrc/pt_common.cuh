// pt_common.cuh -- shared definitions of libsisi4s_pt (host + device).
//
// Data layout in HBM (all FP64).  "Raw" arrays keep the reference's CTF
// column-major layouts; "packed" arrays are pre-tiled images of exactly what one
// pipeline stage of the fused kernel stages in shared memory, so that every
// stage is a handful of contiguous cp.async.bulk (TMA engine, SASS UBLKCP)
// copies and every DMMA fragment load is a conflict-free 256-byte warp read.
//
//   TILE = 16 particle indices per range, nr = ceil(v/16) ranges, vp = 16 nr
//   nk4 = ceil(vd/4) chunks of the particle contraction index d (vd = v unless Re/Im parts are stacked)
//   nl4 = ceil(ol/4) chunks of the hole contraction index l (ol = o unless pt_create_ex)
//
//   Vt [z][Q][R][dc][n=256][kk=4]      = Vppph[b=16Q+n/16, c=16R+n%16, d=4dc+kk, z]
//   Tt [y][x][P][dc][m=16][kk=4]       = T2[a=16P+m, d=4dc+kk, x, y]
//   T2h[x][P][Q][lc][g=2][b8=8][m=16][kk=4] = T2[a=16P+m, b=16Q+8g+b8, x, l=4lc+kk]
//   Ut [z][y][R][lc][c=16][kk=4]       = -Vhhhp[y, z, l=4lc+kk, c=16R+c]
//
// Out-of-range elements (a,b,c,d >= v or l >= o) are stored as zeros, so the
// kernels need no boundary handling in the contraction loops.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/sisi4s_pt.h"

namespace pt {

constexpr int TILE = 16;
constexpr int STAGE_DBL = 2304;            // doubles per pipeline stage (18432 B: K = 8 of the particle contraction)
constexpr int NSTAGE = 8;               // operand ring depth (X tiles live in TMEM, so the ring owns the smem)
constexpr int XT_DBL = TILE * TILE * TILE; // one X tile
constexpr int NCONSUMER_WARPS = 8;
constexpr int NPRODUCER_WARPS = 4;          // stage j of the operand stream is issued by producer warp j % 4
constexpr int FUSED_THREADS = (NCONSUMER_WARPS + NPRODUCER_WARPS) * 32;

struct Dims {
  int o, v;  // ACTIVE holes (the hole indices of the triples that are run), particles
  int ol;    // holes of the contraction sum_l of getDoublesContribution; = o unless the engine holds a
             // hole SUBSET of a larger problem (pt_create_ex) or a stacked complex problem
  int vd;    // length of the particle contraction sum_d; = v unless real and imaginary parts are
             // stacked along d (complex triples: W_re = [T_re | -T_im] . [V_re ; V_im])
  int nr;    // particle ranges
  int vp;    // 16*nr
  int nk4;   // ceil(v/4)
  int nl4;   // ceil(o/4)
};

__host__ __device__ inline Dims make_dims(int o, int v, int ol = 0, int vd = 0) {
  Dims d;
  d.o = o; d.v = v;
  d.ol = ol > 0 ? ol : o;
  d.vd = vd > 0 ? vd : v;
  d.nr = (v + TILE - 1) / TILE;
  d.vp = d.nr * TILE;
  d.nk4 = (d.vd + 3) / 4;
  d.nl4 = (d.ol + 3) / 4;
  return d;
}

__host__ __device__ inline size_t vt_slab_elems(const Dims& d) { return (size_t)d.nr * d.nr * d.nk4 * 1024; }
__host__ __device__ inline size_t vt_elems(const Dims& d) { return vt_slab_elems(d) * d.o; }
__host__ __device__ inline size_t tt_elems(const Dims& d) { return (size_t)d.o * d.o * d.nr * d.nk4 * 64; }
__host__ __device__ inline size_t t2h_elems(const Dims& d) { return (size_t)d.o * d.nr * d.nr * d.nl4 * 1024; }
__host__ __device__ inline size_t ut_elems(const Dims& d) { return (size_t)d.o * d.o * d.nr * d.nl4 * 64; }

__host__ __device__ inline size_t vt_tile_off(const Dims& d, int z, int Q, int R) {
  return (((size_t)z * d.nr + Q) * d.nr + R) * d.nk4 * 1024;
}
__host__ __device__ inline size_t tt_panel_off(const Dims& d, int x, int y, int P) {
  return (((size_t)y * d.o + x) * d.nr + P) * d.nk4 * 64;
}
__host__ __device__ inline size_t t2h_block_off(const Dims& d, int x, int P, int Q) {
  return (((size_t)x * d.nr + P) * d.nr + Q) * d.nl4 * 1024;
}
__host__ __device__ inline size_t ut_panel_off(const Dims& d, int y, int z, int R) {
  return (((size_t)z * d.o + y) * d.nr + R) * d.nl4 * 64;
}

// ---- parameters of the fused kernel ---------------------------------------
struct FusedParams {
  Dims d;
  const double* Tt;
  const double* T2h;
  const double* Vt;
  const double* Ut;
  const int* vslot;    // hole z -> slot of its PPPH slab in Vt; nullptr = identity (all slabs resident)
  const double* t1;    // raw [v,o]
  const double* pphh;  // raw [v,v,o,o]
  const double* qsum;  // pphh[b,c,j,k] + pphh[c,b,k,j], same layout
  // optional SECOND singles term (complex triples: S_re = 1/2 (t_re P_re - t_im P_im) is a sum of two
  // outer products): Sd += 1/2 t1b (x) pphhb, same layouts; nullptr = one term
  const double* t1b;
  const double* pphhb;
  const double* qsumb;
  const double* epsi;
  const double* epsa;
  const int4* triples;    // (i,j,k,class) of the sorted triples of this run
  const uchar4* orbits;   // (A,B,C,class), A>=B>=C
  int norbits;
  int ntriples;
  int order;              // 0: triple-major (orbit fastest), 1: orbit-major (triple fastest; L2 reuse of PPPH tiles)
  int debug;              // measurement switches (wrong results): 1 = consumers skip LDS/DMMA (operand-feed ceiling), 4 = skip the scatter, 8 = skip the epilogue point loops
  long long nitems;       // ntriples * norbits
  double* e_item;         // [nitems]: per-item partial sums (one plain store per item; summed per triple in a
                          // fixed order by reduce_items_kernel, so E_t is bitwise reproducible)
  unsigned int* sync_ctr; // item-round barrier counter (zeroed per launch); nullptr = CTAs run free
  int sync_every;         // barrier before every sync_every-th item round
};

// item -> (triple, orbit).  Orbit-major order makes the CTAs that run concurrently work on
// the SAME particle-range orbit of neighbouring hole triples, so they stream the same PPPH
// tiles (z; Q,R) at the same time and those are served by L2 instead of HBM.
__device__ __forceinline__ void decode_item(const FusedParams& p, long long item, int& t, int& orb) {
  if (p.order == 0) {
    t = (int)(item / p.norbits);
    orb = (int)(item - (long long)t * p.norbits);
  } else {
    orb = (int)(item / p.ntriples);
    t = (int)(item - (long long)orb * p.ntriples);
  }
}

// one W tile job of the debug kernel
struct WTileJob { int x, y, z, P, Q, R; };

// ---- host-side launchers (defined in the .cu files) ------------------------
cudaError_t launch_pack_vt_slab(const double* raw_slab, double* vt_slab, Dims d, cudaStream_t s);
cudaError_t launch_pack_tt(const double* t2, double* tt, Dims d, cudaStream_t s);
cudaError_t launch_pack_t2h(const double* t2, double* t2h, Dims d, cudaStream_t s);
cudaError_t launch_pack_ut(const double* hhhp, double* ut, Dims d, const int* hmap, cudaStream_t s);
cudaError_t launch_pphh_symsum(const double* pphh, double* qsum, Dims d, cudaStream_t s);
cudaError_t launch_reduce_items(const double* e_item, int ntriples, int norbits, int order, double* e_triple,
                                cudaStream_t s);

// ---- integrals from the vertex (pt_pack.cu): one K-major image of the vertex, batched DMMA GEMMs
enum { VG_STRIDED = 0, VG_PACKED = 1 };
struct VgParams {
  const double* gp;        // Gp[kc][r][4], r = p + np q   (A operand; also the B operand unless gpb is set)
  long long rows_padded;
  const double* gpb;       // optional separate K-major image of the B operand (tensor engine), rows_padded_b rows
  long long rows_padded_b;
  double alpha, beta;      // VG_STRIDED: out = alpha * acc (+ beta * out when accumulate)
  int accumulate;
  int kp4;                 // K chunks of 4 (K = 2 nf padded to a multiple of 16)
  int mode;
  // VG_STRIDED: batch (b0, b1) -> first rows / output offset; optional index maps (hole subsets)
  int nb0, nb1;
  const int *map0, *map1;
  long long a_base, a_s0, a_s1, b_base, b_s0, b_s1, o_s0, o_s1;
  int M, N;
  long long sm, sn;
  // VG_PACKED: one PPPH hole slab in the Vt layout
  int np, a0, v, nr, nk4;
  double* out;
};
inline long long vertex_rows_padded(int np) { return (long long)np * np + 256; }
inline int vertex_kp4(int nf) { return ((2 * nf + 15) / 16) * 4; }
cudaError_t vertex_gemm_configure();
cudaError_t launch_pack_vertex(const double* gre, const double* gim, int nf, int np, double* gp, cudaStream_t s);
cudaError_t launch_vertex_gemm(const VgParams& p, cudaStream_t s);

cudaError_t fused_configure(int* smem_bytes_out);
cudaError_t launch_fused(const FusedParams& p, int grid, cudaStream_t s);
cudaError_t launch_w_tile(const FusedParams& p, WTileJob job, double* d_out, cudaStream_t s);

cudaError_t launch_naive_w(const double* t2, const double* ppph, const double* hhhp, Dims d,
                           int x, int y, int z, double* w, cudaStream_t s);
cudaError_t launch_naive_energy(const double* const* w6, const double* t1, const double* pphh,
                                const double* epsi, const double* epsa, Dims d, int i, int j, int k,
                                double* e_out, cudaStream_t s);

cudaError_t launch_bench_fp64(int mode, int blocks, int warps, int iters, double* d_sink,
                              unsigned long long* d_cycles, cudaStream_t s);

}  // namespace pt
