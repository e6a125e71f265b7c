// pt_fused.cu -- the product kernel of the (T) step.
//
// One persistent CTA per SM.  A work item is (sorted hole triple i<=j<=k, orbit
// {A>=B>=C} of 16-wide particle ranges).  For an item the CTA
//   1. runs the table-driven list of stacked GEMM steps (tools/gen_tables.py):
//      each step contracts up to two 16-row T2 panels with ONE shared 256-pair
//      tile of a PPPH slab (K = v) plus the hole term (K = o) on FP64 tensor
//      cores (mma.sync m8n8k4.f64 -> SASS DMMA.8x8x4), operands streamed from the
//      pre-tiled HBM layouts by cp.async.bulk (TMA engine, SASS UBLKCP) through an
//      8-stage mbarrier ring of 18 KB stages (K = 8) filled by four producer warps; the
//      consumers interleave the fragment loads and barrier traffic of the next K-chunk
//      between the DMMAs of the current one;
//   2. adds each 16^3 W tile, index-permuted, into the orbit's six X tiles, which
//      live in TENSOR MEMORY (6 x 32 KB of the SM's 256 KB TMEM, tcgen05.ld/st ->
//      SASS LDTM/STTM) for the whole item, so that shared memory belongs to the
//      operand ring -- the v^3 triples blocks are never written to HBM.  The index
//      permutation goes through a per-group 32 KB staging tile in shared memory;
//   3. epilogue: the X tiles are copied from TMEM into the (drained) ring region, then
//      permutational symmetrisation (generic orbits: a whole S3 point orbit per thread, six
//      reads, the S3 group table and ONE reciprocal for six points; degenerate orbits: six
//      permuted reads per point), singles term, eigenvalue denominator, warp-shuffle
//      reduction, one store per item (summed per triple in a fixed order by reduce_items_kernel).
//
// This replaces, per sorted triple, getDoublesContribution / the permutation
// accumulate / divide / spin-factor symmetrise / energy dot of the reference
// (src/algorithms/CcsdPerturbativeTriples.cxx:87-96,161-216), i.e. ~150
// collective CTF operations on v^3 tensors.
//
// Fragment conventions (PTX ISA, mma.m8n8k4 .f64):
//   A (8x4, row):  a  = A[lane>>2][lane&3]
//   B (4x8, col):  b  = B[lane&3][lane>>2]
//   C (8x8):       c0,c1 = C[lane>>2][2*(lane&3) + {0,1}]
// The 8 consumer warps form two groups of four; group h = warp>>2 owns half h of the
// stacked 32 x (16b x 16c) output (one 16^3 W tile), warp wq = warp&3 of the group the
// columns b in {wq, wq+4, wq+8, wq+12}:  acc[mf][bi][co][e] is
//   W_h[la = 8 mf + (lane>>2), lb = wq + 4 bi, lc = 8 co + 2 (lane&3) + e].
#include "pt_common.cuh"
#include "pt_tables.h"

namespace pt {

__constant__ PtClassTable c_tab[4][4] = PT_TABLES_INIT;

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
// one non-blocking probe; the result is consumed later (mbar_wait_tok), so that the SYNCS latency
// hides behind the DMMA block issued in between
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait_tok(uint32_t bar, uint32_t parity, uint32_t ok) {
  if (!ok) mbar_wait(bar, parity);
}
template <int OFF>
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double x;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(x) : "r"(addr), "n"(OFF));
  return x;
}
// 1-D bulk copy global -> shared, completion counted on an mbarrier (TMA engine)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void consumer_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(NCONSUMER_WARPS * 32) : "memory");
}
__device__ __forceinline__ void group_barrier(int grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(2 + grp), "n"(NCONSUMER_WARPS * 16) : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// ---- tensor memory (TMEM) as the home of the X accumulator tiles ------------------------
// tcgen05.ld/st .32x32b: thread i of warp w addresses TMEM lane 32*(w%4)+i and N consecutive
// 32-bit columns (verified on hardware by tools/probes/tmem_probe.cu).  A double occupies two
// columns (lo, hi).  X tile tau, element n = x0 + 16 x1 + 256 x2 lives in lane n & 127,
// columns 64 tau + 2 (n >> 7) + {0,1}.
#define TM_R16(r, o) "=r"(r[o+0]),"=r"(r[o+1]),"=r"(r[o+2]),"=r"(r[o+3]),"=r"(r[o+4]),"=r"(r[o+5]),"=r"(r[o+6]),"=r"(r[o+7]),"=r"(r[o+8]),"=r"(r[o+9]),"=r"(r[o+10]),"=r"(r[o+11]),"=r"(r[o+12]),"=r"(r[o+13]),"=r"(r[o+14]),"=r"(r[o+15])
#define TM_I16(r, o) "r"(r[o+0]),"r"(r[o+1]),"r"(r[o+2]),"r"(r[o+3]),"r"(r[o+4]),"r"(r[o+5]),"r"(r[o+6]),"r"(r[o+7]),"r"(r[o+8]),"r"(r[o+9]),"r"(r[o+10]),"r"(r[o+11]),"r"(r[o+12]),"r"(r[o+13]),"r"(r[o+14]),"r"(r[o+15])
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : TM_R16(r, 0), TM_R16(r, 16)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      TM_I16(r, 0), TM_I16(r, 16)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
constexpr int TMEM_COLS = 512;
// register split (setmaxnreg): 384 threads x 168 at launch = 8 x 32 x 224 + 4 x 32 x 56
#ifndef PT_CONSUMER_REGS
#define PT_CONSUMER_REGS 224
#define PT_PRODUCER_REGS 56
#endif
// unroll factor of the epilogue's x2 point loops (16 = full; experiment: smaller bodies against the
// instruction-cache misses ncu shows there)
#ifndef PT_EPI_UNROLL
#define PT_EPI_UNROLL 16
#endif
constexpr int EPI_UNROLL = PT_EPI_UNROLL;
constexpr int CONSUMER_REGS = PT_CONSUMER_REGS;
constexpr int PRODUCER_REGS = PT_PRODUCER_REGS;
static_assert(NCONSUMER_WARPS * 32 * CONSUMER_REGS + NPRODUCER_WARPS * 32 * PRODUCER_REGS <= 65536, "register file");
constexpr int TMEM_COLS_PER_TILE = 64;

__device__ __forceinline__ int sel3(int a, int b, int c, int idx) {
  return idx == 0 ? a : (idx == 1 ? b : c);
}
// X-tile element index with the XOR swizzle (found by exhaustive search over XOR-linear maps of the three nibbles):
// low nibble x0 ^ x1 ^ bitswap13(x2); at most 2-way bank conflicts for every
// permuted accumulate / read pattern of the kernel.
__device__ __forceinline__ int xt_index(int x0, int x1, int x2) {
  const int s2 = (x2 & 5) | ((x2 & 2) << 2) | ((x2 & 8) >> 2);
  return (x0 ^ x1 ^ s2) + 16 * x1 + 256 * x2;
}

// ------------------------------------------------------------ step descriptors
struct StepSrc {
  const double* v;      // Vt tile            (nk4 chunks of 1024)
  const double* t[2];   // Tt panels          (nk4 chunks of 64)
  const double* hh[2];  // T2h blocks         (nl4 * 2 chunks of 512)
  const double* u[2];   // Ut panels          (nl4 chunks of 64)
  int en[2];
};

__device__ __forceinline__ StepSrc make_step_src(const FusedParams& p, const PtStep& st, int h0, int h1,
                                                 int h2, int rg0, int rg1, int rg2) {
  StepSrc s;
  const int z = sel3(h0, h1, h2, st.zs);
  const int P = sel3(rg0, rg1, rg2, st.r0);
  const int Q = sel3(rg0, rg1, rg2, st.r1);
  const int R = sel3(rg0, rg1, rg2, st.r2);
  // hole-blocked PPPH residency: slab z lives in slot vslot[z] of Vt (identity when all slabs are resident)
  s.v = p.Vt + vt_tile_off(p.d, p.vslot ? p.vslot[z] : z, Q, R);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int x = sel3(h0, h1, h2, st.h[h].tx);
    const int y = sel3(h0, h1, h2, st.h[h].ty);
    s.en[h] = st.h[h].en;
    s.t[h] = p.Tt + tt_panel_off(p.d, x, y, P);
    s.hh[h] = p.T2h + t2h_block_off(p.d, x, P, Q);
    s.u[h] = p.Ut + ut_panel_off(p.d, y, z, R);
  }
  return s;
}

struct Pipe {
  uint32_t ring;   // shared address of stage 0
  uint32_t full;   // shared address of full[0]
  uint32_t empty;  // shared address of empty[0]
  uint32_t slot;
  uint32_t phase;
  __device__ __forceinline__ void advance() {
    if (++slot == NSTAGE) { slot = 0; phase ^= 1; }
  }
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// Stage j of a step (STAGE_DBL = 2304 doubles):
//   j < nk8: particle chunks 2j, 2j+1 (K = 8): [V chunk 2j | V chunk 2j+1 | T0: 2 x 64 | T1: 2 x 64]
//            = one 16 KB slab-tile copy + one 1 KB copy per enabled T2 panel (half of that for
//            the odd tail chunk);
//   else hole chunk lc = j - nk8 (K = 4, both b-halves g): [T2h half 0: 1024 | T2h half 1: 1024 |
//            U0: 64 (+64 unused) | U1: 64] = one 8 KB + one 512 B copy per enabled half.
__device__ __forceinline__ void issue_stage(const StepSrc& s, int j, int nk4, uint32_t stage, uint32_t full) {
  const uint32_t nen = (uint32_t)(s.en[0] + s.en[1]);
  const int nk8 = (nk4 + 1) >> 1;
  if (j < nk8) {
    const int c0 = 2 * j;
    const uint32_t two = (c0 + 1 < nk4) ? 2u : 1u;
    mbar_expect_tx(full, two * (8192u + 512u * nen));
    bulk_g2s(stage, s.v + (size_t)c0 * 1024, two * 8192u, full);
    if (s.en[0]) bulk_g2s(stage + 16384u, s.t[0] + (size_t)c0 * 64, two * 512u, full);
    if (s.en[1]) bulk_g2s(stage + 17408u, s.t[1] + (size_t)c0 * 64, two * 512u, full);
  } else {
    const int lc = j - nk8;
    mbar_expect_tx(full, 8704u * nen);
    if (s.en[0]) {
      bulk_g2s(stage, s.hh[0] + (size_t)lc * 1024, 8192u, full);
      bulk_g2s(stage + 16384u, s.u[0] + (size_t)lc * 64, 512u, full);
    }
    if (s.en[1]) {
      bulk_g2s(stage + 8192u, s.hh[1] + (size_t)lc * 1024, 8192u, full);
      bulk_g2s(stage + 17408u, s.u[1] + (size_t)lc * 64, 512u, full);
    }
  }
}
// Position in the CTA's stream of operand stages (item -> step -> stage); warp-uniform.
struct StageIter {
  long long item;
  int s, j, nsteps;
  bool valid;
  int4 tr;
  uchar4 ob;
  int tc, oc;
  StepSrc src;
  __device__ __forceinline__ void load_item(const FusedParams& p) {
    valid = item < p.nitems;
    if (!valid) return;
    int t, orb;
    decode_item(p, item, t, orb);
    tr = p.triples[t];
    ob = p.orbits[orb];
    tc = tr.w;
    oc = ob.w;
    nsteps = c_tab[tc][oc].nsteps;
    s = 0;
    j = 0;
    src = make_step_src(p, c_tab[tc][oc].steps[0], tr.x, tr.y, tr.z, ob.x, ob.y, ob.z);
  }
  // move n stages forward in the stream
  __device__ __forceinline__ void advance(const FusedParams& p, int n, int nst, int stride) {
    j += n;
    while (valid && j >= nst) {
      const int jj = j - nst;
      if (++s < nsteps) {
        src = make_step_src(p, c_tab[tc][oc].steps[s], tr.x, tr.y, tr.z, ob.x, ob.y, ob.z);
      } else {
        item += stride;
        load_item(p);
      }
      j = jj;
    }
  }
};


// Producer warps.  One thread's issue rate (mbarrier wait + expect_tx + three or four bulk
// copies + stream bookkeeping, ~500 clk per stage) cannot feed the DMMA pipe, so
// NPRODUCER_WARPS warps share the stream: warp pw issues the stages n = pw (mod NPRODUCER_WARPS)
// of the CTA's global stage sequence.  All 32 lanes run the (uniform) bookkeeping, one elected
// lane issues the copies.
__device__ __forceinline__ void producer_loop(const FusedParams& p, const Pipe& pp0, uint32_t go_bar, int pw,
                                              long long first_item, int stride) {
  const int nk4 = p.d.nk4, nst = ((nk4 + 1) >> 1) + p.d.nl4;
  StageIter ld;
  ld.item = first_item;
  ld.load_item(p);
  if (ld.valid) ld.advance(p, pw, nst, stride);
  const bool leader = elect_one();
  uint32_t n = (uint32_t)pw;  // global stage number -> ring slot n % NSTAGE, phase (n / NSTAGE) & 1
  uint32_t items_started = 0;
  long long cur_item = -1;
  while (ld.valid) {
    if (ld.item != cur_item) {
      // the epilogue of the previous item copies the X tiles over the ring: wait until it is done
      if (items_started > 0) {
        mbar_wait(go_bar, (items_started - 1) & 1);
        // Item-round barrier: no CTA starts its n-th item before every CTA has finished its
        // (n-1)-th.  Concurrent CTAs work on the same particle-range orbit of neighbouring hole
        // triples and share two of their three PPPH slabs; kept in step they read the same tiles
        // within the L2's residency window, which turns most DRAM re-reads into L2 hits.  (All
        // CTAs are co-resident: cooperative launch, one CTA per SM.)
        if (p.sync_ctr && items_started % (uint32_t)p.sync_every == 0) {
          const unsigned target = items_started * gridDim.x;
          unsigned seen;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.sync_ctr) : "memory");
            if (seen < target) __nanosleep(200);
          } while (seen < target);
        }
      }
      ++items_started;
      cur_item = ld.item;
    }
    const uint32_t slot = n % NSTAGE, phase = (n / NSTAGE) & 1;
    const uint32_t full = pp0.full + 8 * slot, stage = pp0.ring + slot * (STAGE_DBL * 8);
    mbar_wait(pp0.empty + 8 * slot, phase ^ 1);
    if (leader) issue_stage(ld.src, ld.j, nk4, stage, full);
    n += NPRODUCER_WARPS;
    ld.advance(p, NPRODUCER_WARPS, nst, stride);
  }
}

// consumer side of one step for one warp.  The 8 consumer warps form two groups of four:
// group h = warp>>2 owns half h of the stacked step (one 16 x 256 W tile), warp wq = warp&3 of
// the group owns the columns b in {wq, wq+4, wq+8, wq+12}:  acc[mf][bi][co][e] is
//   W_h[la = 8 mf + (lane>>2), lb = wq + 4 bi, lc = 8 co + 2 (lane&3) + e].
// The two warps that share an SM sub-partition (w, w+4) therefore belong to different groups
// and scatter into different X tiles, so no CTA barrier is needed between steps and one
// group's scatter overlaps the other group's DMMA stream.
//
// Software pipeline (per warp): a "unit" is one K=4 chunk (16 DMMA, 10 LDS.64) or one hole
// half-chunk (8 DMMA, 4-6 LDS.64); a stage holds two units.  The loads of unit n+1 (into the other
// register buffer), the full-barrier wait of its stage, the release of a finished stage and the
// probe of the next stage's barrier are all INTERLEAVED between the DMMAs of unit n, one
// instruction per DMMA issue slot.  A warp therefore never leaves the DMMA stream: the two warps
// that share an SM sub-partition's FP64 tensor pipe (issue interval 16 clk) cannot fall into a
// common "fetch" phase during which the pipe would idle.  All shared addresses are one
// per-thread base + slot * stage bytes + immediates.
struct FragP { double b[4][2]; double a[2]; };
struct FragH { double a[2][2]; };
enum { NEXT_NONE = 0, NEXT_P0 = 1, NEXT_P1 = 2, NEXT_H0 = 3, NEXT_H1 = 4 };

struct Consumer {
  Pipe& pp;
  uint32_t offB, offA, offH;  // per-thread fragment bases (shared byte addresses of stage 0)
  uint32_t B, A, H;           // ... of the current stage
  uint32_t tok;
  bool lane0;
  static constexpr uint32_t SB = STAGE_DBL * 8;

  __device__ __forceinline__ void bind() {
    const uint32_t sb = pp.slot * SB;
    B = offB + sb; A = offA + sb; H = offH + sb;
  }
  __device__ __forceinline__ void probe() { tok = mbar_try(pp.full + 8 * pp.slot, pp.phase); }
  __device__ __forceinline__ void await() {
    mbar_wait_tok(pp.full + 8 * pp.slot, pp.phase, tok);
    bind();
  }
  __device__ __forceinline__ void release() {  // this warp is done reading the current stage
    __syncwarp();
    if (lane0) mbar_arrive(pp.empty + 8 * pp.slot);
    pp.advance();
    probe();
  }
  // load n of the next unit (K = particle chunk half 0/1, hole half-chunk 0/1)
  template <int NEXT, int N, int NB = 4>
  __device__ __forceinline__ void ld(FragP& fp, FragH& fh, double (&u)[2]) {
    if constexpr (NEXT == NEXT_P0 || NEXT == NEXT_P1) {
      constexpr int C = (NEXT == NEXT_P1) ? 1 : 0;
      if constexpr (N < 8) { if constexpr ((N >> 1) < NB) fp.b[N >> 1][N & 1] = lds_f64<C * 8192 + (N >> 1) * 2048 + (N & 1) * 256>(B); }
      else if constexpr (N < 10) fp.a[N - 8] = lds_f64<C * 512 + (N - 8) * 256>(A);
    } else if constexpr (NEXT == NEXT_H0 || NEXT == NEXT_H1) {
      constexpr int G = (NEXT == NEXT_H1) ? 1 : 0;
      if constexpr (N < 4) { if constexpr (2 * G + (N >> 1) < NB) fh.a[N >> 1][N & 1] = lds_f64<G * 4096 + (N >> 1) * 2048 + (N & 1) * 256>(H); }
      else if constexpr (N < 6 && NEXT == NEXT_H0) u[N - 4] = lds_f64<(N - 4) * 256>(A);
    }
  }
  template <int NEXT>
  __host__ __device__ static constexpr int nloads() {
    return (NEXT == NEXT_P0 || NEXT == NEXT_P1) ? 10 : (NEXT == NEXT_H0 ? 6 : (NEXT == NEXT_H1 ? 4 : 0));
  }

  // 16 DMMAs of particle chunk `cur`, with the next unit fetched in between.  NB < 4: the step's b range is
  // the padded last particle range and only its first NB column groups (b = wq + 4 bi, bi < NB) hold
  // valid particles; the DMMAs and fragment loads of the others are not issued (their accumulators stay 0).
  template <int NEXT, int NB>
  __device__ __forceinline__ void block_p(double (&acc)[2][4][2][2], const FragP& cur, FragP& fp, FragH& fh,
                                          double (&u)[2], bool last) {
    constexpr int NL = nloads<NEXT>();
    constexpr bool first = (NEXT == NEXT_P0 || NEXT == NEXT_H0);
#define PT_DM(n) do { if constexpr ((((n) >> 1) & 3) < NB)                                                      \
      dmma(acc[(n) >> 3][((n) >> 1) & 3][(n) & 1][0], acc[(n) >> 3][((n) >> 1) & 3][(n) & 1][1],               \
           cur.a[(n) >> 3], cur.b[((n) >> 1) & 3][(n) & 1]); } while (0)
    PT_DM(0);
    if constexpr (first) await();
    ld<NEXT, 0, NB>(fp, fh, u); PT_DM(1);
    ld<NEXT, 1, NB>(fp, fh, u); PT_DM(2);
    ld<NEXT, 2, NB>(fp, fh, u); PT_DM(3);
    ld<NEXT, 3, NB>(fp, fh, u); PT_DM(4);
    ld<NEXT, 4, NB>(fp, fh, u); PT_DM(5);
    ld<NEXT, 5, NB>(fp, fh, u); PT_DM(6);
    ld<NEXT, 6, NB>(fp, fh, u); PT_DM(7);
    ld<NEXT, 7, NB>(fp, fh, u); PT_DM(8);
    ld<NEXT, 8, NB>(fp, fh, u); PT_DM(9);
    ld<NEXT, 9, NB>(fp, fh, u); PT_DM(10);
    if (NL > 0 && last) release();
    PT_DM(11); PT_DM(12); PT_DM(13); PT_DM(14); PT_DM(15);
#undef PT_DM
  }
  // 8 DMMAs of hole half-chunk g (columns bi = 2g + j)
  template <int NEXT, int G, int NB>
  __device__ __forceinline__ void block_h(double (&acc)[2][4][2][2], const FragH& cur, const double (&uc)[2],
                                          FragP& fp, FragH& fh, double (&u)[2]) {
    constexpr bool first = (NEXT == NEXT_H0);
#define PT_DH(n) do { if constexpr (2 * G + ((n) >> 2) < NB)                                               \
      dmma(acc[(n) & 1][2 * G + ((n) >> 2)][((n) >> 1) & 1][0],                                            \
           acc[(n) & 1][2 * G + ((n) >> 2)][((n) >> 1) & 1][1], cur.a[(n) >> 2][(n) & 1],                  \
           uc[((n) >> 1) & 1]); } while (0)
    PT_DH(0);
    if constexpr (first) await();
    ld<NEXT, 0, NB>(fp, fh, u); PT_DH(1);
    ld<NEXT, 1, NB>(fp, fh, u); PT_DH(2);
    ld<NEXT, 2, NB>(fp, fh, u); PT_DH(3);
    ld<NEXT, 3, NB>(fp, fh, u); PT_DH(4);
    ld<NEXT, 4, NB>(fp, fh, u); PT_DH(5);
    ld<NEXT, 5, NB>(fp, fh, u);
    if constexpr (NEXT == NEXT_H1) release();
    PT_DH(6); PT_DH(7);
#undef PT_DH
  }
};

template <int NB>
__device__ __forceinline__ void consume_step(double (&acc)[2][4][2][2], Pipe& pp, int nk4, int nl4, int en,
                                             int h, int wq, int lane) {
#pragma unroll
  for (int mf = 0; mf < 2; ++mf)
#pragma unroll
    for (int bi = 0; bi < 4; ++bi)
#pragma unroll
      for (int co = 0; co < 2; ++co) acc[mf][bi][co][0] = acc[mf][bi][co][1] = 0.0;
  const int nk8 = (nk4 + 1) >> 1;

  if (!en) {
    // disabled half: keep the stage accounting going
    const int n = nk8 + nl4;
    for (int s = 0; s < n; ++s) {
      mbar_wait(pp.full + 8 * pp.slot, pp.phase);
      __syncwarp();
      if (lane == 0) mbar_arrive(pp.empty + 8 * pp.slot);
      pp.advance();
    }
    return;
  }

  Consumer c{pp};
  c.offB = pp.ring + (uint32_t)(wq * 64 + lane) * 8u;            // + chunk*8192 + bi*2048 + co*256
  c.offA = pp.ring + (uint32_t)(2048 + h * 128 + lane) * 8u;     // + chunk*512 + mf*256  (U: + co*256)
  c.offH = pp.ring + (uint32_t)(h * 1024 + wq * 64 + lane) * 8u; // + g*4096 + j*2048 + mf*256
  c.lane0 = lane == 0;

  FragP fA, fB;
  FragH hA, hB;
  double uf[2], un[2];
  // prologue: chunk 0 of the first stage
  c.probe();
  c.await();
  c.ld<NEXT_P0, 0>(fA, hA, uf); c.ld<NEXT_P0, 1>(fA, hA, uf); c.ld<NEXT_P0, 2>(fA, hA, uf);
  c.ld<NEXT_P0, 3>(fA, hA, uf); c.ld<NEXT_P0, 4>(fA, hA, uf); c.ld<NEXT_P0, 5>(fA, hA, uf);
  c.ld<NEXT_P0, 6>(fA, hA, uf); c.ld<NEXT_P0, 7>(fA, hA, uf); c.ld<NEXT_P0, 8>(fA, hA, uf);
  c.ld<NEXT_P0, 9>(fA, hA, uf);
  if (nk4 == 1) c.release();
  // ---- particle contraction: W[a,b,c] += sum_d T2[a,d,x,y] V[b,c,d,z]
  int dc = 0;
  // bulk: two full stages per iteration, no tail conditions (halves the loop back-edges, whose
  // branch resolution otherwise costs ~5 % of the consumers' time)
  for (; dc + 9 < nk4; dc += 8) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      c.block_p<NEXT_P1, NB>(acc, fA, fB, hB, un, true);    // chunk dc+2r;   fetch dc+2r+1, stage done
      c.block_p<NEXT_P0, NB>(acc, fB, fA, hA, uf, false);   // chunk dc+2r+1; fetch dc+2r+2 (dc+2r+3 exists)
    }
  }
  for (; dc + 5 < nk4; dc += 4) {
    c.block_p<NEXT_P1, NB>(acc, fA, fB, hB, un, true);    // chunk dc;   fetch dc+1, stage done
    c.block_p<NEXT_P0, NB>(acc, fB, fA, hA, uf, false);   // chunk dc+1; fetch dc+2
    c.block_p<NEXT_P1, NB>(acc, fA, fB, hB, un, true);    // chunk dc+2; fetch dc+3, stage done
    c.block_p<NEXT_P0, NB>(acc, fB, fA, hA, uf, false);   // chunk dc+3; fetch dc+4 (dc+5 exists)
  }
  for (; dc + 1 < nk4; dc += 2) {
    c.block_p<NEXT_P1, NB>(acc, fA, fB, hB, un, true);                      // chunk dc; fetch dc+1, stage done
    if (dc + 2 < nk4) c.block_p<NEXT_P0, NB>(acc, fB, fA, hA, uf, dc + 3 >= nk4);  // chunk dc+1; fetch dc+2
    else c.block_p<NEXT_H0, NB>(acc, fB, fA, hA, uf, false);                // chunk dc+1; fetch hole (0, g=0)
  }
  if (nk4 & 1) c.block_p<NEXT_H0, NB>(acc, fA, fB, hA, uf, false);          // tail chunk; fetch hole (0, g=0)
  // ---- hole contraction: W[a,b,c] += sum_l T2[a,b,x,l] (-Vhhhp[y,z,l,c]); half-chunk g covers
  //      b = 8g .. 8g+7, of which this warp owns b = 8g + wq + 4j  (bi = 2g + j)
  int lc = 0;
  for (; lc + 4 < nl4; lc += 4) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      c.block_h<NEXT_H1, 0, NB>(acc, hA, uf, fA, hB, un);
      c.block_h<NEXT_H0, 1, NB>(acc, hB, uf, fA, hA, un);
      uf[0] = un[0];
      uf[1] = un[1];
    }
  }
  for (; lc + 1 < nl4; ++lc) {
    c.block_h<NEXT_H1, 0, NB>(acc, hA, uf, fA, hB, un);
    c.block_h<NEXT_H0, 1, NB>(acc, hB, uf, fA, hA, un);
    uf[0] = un[0];
    uf[1] = un[1];
  }
  c.block_h<NEXT_H1, 0, NB>(acc, hA, uf, fA, hB, un);
  c.block_h<NEXT_NONE, 1, NB>(acc, hB, uf, fA, hA, un);
}

// composition table of Permutation<3> indices: S3_MUL[mu][nu] = index of m -> mu(nu(m)), i.e.
// (x o mu) o nu = x o S3_MUL[mu][nu] (images p0..p5 = 012, 102, 120, 021, 201, 210; this is the
// `nbr` table of the generic orbit class, checked by static_assert-free tests/test_tables.py)
__device__ constexpr int8_t S3_MUL[6][6] = {{0, 1, 2, 3, 4, 5}, {1, 0, 3, 2, 5, 4}, {2, 5, 4, 1, 0, 3},
                                            {3, 4, 5, 0, 1, 2}, {4, 3, 0, 5, 2, 1}, {5, 2, 1, 4, 3, 0}};

__host__ __device__ constexpr int bitswap13(int v) { return (v & 5) | ((v & 2) << 2) | ((v & 8) >> 2); }
__host__ __device__ constexpr int csel3(int a, int b, int c, int idx) { return idx == 0 ? a : (idx == 1 ? b : c); }

// S[x] = W_h[w] with x_n = w_{q[n]}, q compile-time: the warp's W fragment is written,
// index-permuted, into the group's staging tile (same XOR-swizzled layout as the X tiles of the
// epilogue).  A W coordinate splits into a per-thread part (g, wq, 2*t4) and a per-register part
// (8mf, 4bi, 8co+e) with disjoint bits, and the swizzle of xt_index is XOR-linear, so
// idx = (tl ^ cl) + tbase + cbase with cl, cbase immediates.
template <int Q0, int Q1, int Q2>
__device__ __forceinline__ void stage_store_q(double* S, const double (&a)[2][4][2][2], int wq, int lane) {
  const int g = lane >> 2, t2 = 2 * (lane & 3);
  const int tx0 = csel3(g, wq, t2, Q0), tx1 = csel3(g, wq, t2, Q1), tx2 = csel3(g, wq, t2, Q2);
  const int tl = tx0 ^ tx1 ^ bitswap13(tx2);
  double* Sb = S + 16 * tx1 + 256 * tx2;
#pragma unroll
  for (int mf = 0; mf < 2; ++mf)
#pragma unroll
    for (int bi = 0; bi < 4; ++bi)
#pragma unroll
      for (int co = 0; co < 2; ++co)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c0 = csel3(8 * mf, 4 * bi, 8 * co + e, Q0), c1 = csel3(8 * mf, 4 * bi, 8 * co + e, Q1),
                    c2 = csel3(8 * mf, 4 * bi, 8 * co + e, Q2);
          const int cl = c0 ^ c1 ^ bitswap13(c2);
          Sb[(tl ^ cl) + 16 * c1 + 256 * c2] = a[mf][bi][co][e];
        }
}

// X_tau += (permuted W_h) for one group of four warps:
//   registers -> staging tile (permuted) -> group barrier -> every thread adds the 32 elements
//   that live in its TMEM lane (tcgen05.ld, DADD, tcgen05.st).
// The staging read-back is conflict-free: lanes 0-15 / 16-31 of a warp read 16 consecutive
// doubles each.
__device__ __forceinline__ void scatter_add(double* stg, uint32_t tmem_lane_base, const double (&a)[2][4][2][2],
                                            int tau, int q0, int q1, int grp, int wq, int lane) {
  switch (q0 * 3 + q1) {
    case 1: stage_store_q<0, 1, 2>(stg, a, wq, lane); break;
    case 2: stage_store_q<0, 2, 1>(stg, a, wq, lane); break;
    case 3: stage_store_q<1, 0, 2>(stg, a, wq, lane); break;
    case 5: stage_store_q<1, 2, 0>(stg, a, wq, lane); break;
    case 6: stage_store_q<2, 0, 1>(stg, a, wq, lane); break;
    default: stage_store_q<2, 1, 0>(stg, a, wq, lane); break;
  }
  group_barrier(grp);
  tmem_fence_after();
  // owned elements: n = L + 128 d, L = 32 wq + lane  ->  x0 = L & 15, x1 = (L >> 4) + 8 (d & 1), x2 = d >> 1
  const int x0 = lane & 15, x1lo = 2 * wq + (lane >> 4);
  const uint32_t tcol = tmem_lane_base + (uint32_t)(tau * TMEM_COLS_PER_TILE);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld32(tcol + 32 * c, r);
    double sv[16];
#pragma unroll
    for (int dd = 0; dd < 16; ++dd) {
      const int d = 16 * c + dd;
      sv[dd] = stg[xt_index(x0, x1lo + 8 * (d & 1), d >> 1)];
    }
    tmem_wait_ld();
#pragma unroll
    for (int dd = 0; dd < 16; ++dd) {
      const double x = __hiloint2double((int)r[2 * dd + 1], (int)r[2 * dd]) + sv[dd];
      r[2 * dd] = (uint32_t)__double2loint(x);
      r[2 * dd + 1] = (uint32_t)__double2hiint(x);
    }
    tmem_st32(tcol + 32 * c, r);
  }
  tmem_wait_st();
  tmem_fence_before();
}

// staging values of one X tile of the epilogue (singles-term operands + eigenvalues) into
// registers; issued one tile ahead so the global-load latency hides behind the point loop.
//   Sd = 1/2 (t_i[a] Qa[b,c] + t_j[b] Qb[a,c] + t_k[c] Qc[a,b])
// (getSinglesContribution, CcsdPerturbativeTriples.cxx:81-85, summed over the distinct hole
// permutations; pm = mask of those)
template <int SET = 0>
__device__ __forceinline__ void epi_stage_load(const FusedParams& p, const PtClassTable& tab, uchar4 ob, int tl,
                                               int hi, int hj, int hk, int pm, int tid, double& sq0,
                                               double& sq1, double& sq2, double& stv0, double& stv1) {
  const int v = p.d.v, o = p.d.o;
  const int ga0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[tl][0]);
  const int gb0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[tl][1]);
  const int gc0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[tl][2]);
  const int u = tid & 15, w_ = tid >> 4;
  const double* P = SET == 0 ? p.pphh : p.pphhb;
  const double* S = SET == 0 ? p.qsum : p.qsumb;  // P[b,c,j,k] + P[c,b,k,j]
  const double* T1 = SET == 0 ? p.t1 : p.t1b;
  const size_t vv = (size_t)v;
  sq0 = sq1 = sq2 = 0.0;
  if (p.debug & 32) return;  // measurement switch: no global loads in the epilogue (wrong results)
  // ONE load per operand and no arithmetic on the loaded values here: the results stay in flight
  // until the staging store, across the barriers in between.  Both hole permutations of a pair
  // distinct -> the pre-added Qsum, else the one raw entry that exists.
  auto pick = [&](int both, int m1, int m2, size_t r, size_t c, int h1, int h2) -> const double* {
    if ((pm & both) == both) return S + r + vv * (c + vv * (h1 + (size_t)o * h2));
    if (pm & m1) return P + r + vv * (c + vv * (h1 + (size_t)o * h2));
    if (pm & m2) return P + c + vv * (r + vv * (h2 + (size_t)o * h1));
    return nullptr;
  };
  {
    const size_t g1 = gb0 + u, g2 = gc0 + w_;
    const double* q = pick(1 | 8, 1, 8, g1, g2, hj, hk);
    if (q && g1 < vv && g2 < vv) sq0 = __ldg(q);
  }
  {
    const size_t g0 = ga0 + u, g2 = gc0 + w_;
    const double* q = pick(2 | 4, 2, 4, g0, g2, hi, hk);
    if (q && g0 < vv && g2 < vv) sq1 = __ldg(q);
  }
  {
    const size_t g0 = ga0 + u, g1 = gb0 + w_;
    const double* q = pick(16 | 32, 16, 32, g0, g1, hi, hj);
    if (q && g0 < vv && g1 < vv) sq2 = __ldg(q);
  }
  if (tid < 48) {
    const int which = tid >> 4, uu = tid & 15;
    const int g = (which == 0 ? ga0 : (which == 1 ? gb0 : gc0)) + uu;
    const int hh = which == 0 ? hi : (which == 1 ? hj : hk);
    stv0 = g < v ? __ldg(T1 + g + (size_t)v * hh) : 0.0;
    stv1 = g < v ? __ldg(p.epsa + g) : 0.0;
  }
}

// shared memory: [ring: NSTAGE stages][staging: one 16^3 tile per consumer group][Qs][tv][red]
// [mbarriers: full[NSTAGE], empty[NSTAGE], go][tmem base].  In the epilogue the six X tiles
// (6 x 4096 doubles) are copied from TMEM to the start of the buffer, over the drained ring
// and the first part of the staging tiles.
constexpr int RING_DBL = NSTAGE * STAGE_DBL;
constexpr int STG_DBL = 2 * XT_DBL;
static_assert(RING_DBL + STG_DBL >= 6 * XT_DBL, "epilogue X copy must fit into ring + staging");
__device__ __forceinline__ void carve_smem(unsigned char* raw, double*& Xs, double*& ring, double*& stg,
                                           double*& Qs, double*& tv, double*& red, uint64_t*& bars,
                                           uint32_t*& tmem_slot) {
  Xs = reinterpret_cast<double*>(raw);
  ring = Xs;
  stg = ring + RING_DBL;
  Qs = stg + STG_DBL;
  tv = Qs + 3 * 256;
  red = tv + 96;
  bars = reinterpret_cast<uint64_t*>(red + 8);
  tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 1);
}
constexpr int FUSED_SMEM_BYTES = (RING_DBL + STG_DBL + 3 * 256 + 96 + 8) * 8 + (2 * NSTAGE + 1) * 8 + 16;
static_assert(FUSED_SMEM_BYTES <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");

// ------------------------------------------------------------------ the kernel
// NS = number of singles terms: 1 for the real closed-shell step; 2 adds a second pass of the singles
// part with (t1b, pphhb) for the stacked complex problem (pt_fused_kernel2; the real kernel's code is
// not touched by it).
template <int NS>
__device__ __forceinline__ void fused_body(const FusedParams& p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *Xs, *ring, *stg, *Qs, *tv, *red;
  uint64_t* bars;
  uint32_t* tmem_slot;
  carve_smem(smem_raw, Xs, ring, stg, Qs, tv, red, bars, tmem_slot);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk4 = p.d.nk4, nl4 = p.d.nl4, v = p.d.v, o = p.d.o;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(smem_u32(bars + s), 1);
      mbar_init(smem_u32(bars + NSTAGE + s), NCONSUMER_WARPS);
    }
    mbar_init(smem_u32(bars + 2 * NSTAGE), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    // the whole tensor memory of the SM (one CTA per SM): 6 X tiles use 384 of the 512 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  Pipe pp;
  pp.ring = smem_u32(ring);
  pp.full = smem_u32(bars);
  pp.empty = smem_u32(bars + NSTAGE);
  pp.slot = 0;
  pp.phase = 0;
  const uint32_t go_bar = smem_u32(bars + 2 * NSTAGE);

  if (warp >= NCONSUMER_WARPS) {
    // ===== producer warps (one warpgroup): hand registers to the consumers =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
    producer_loop(p, pp, go_bar, warp - NCONSUMER_WARPS, blockIdx.x, gridDim.x);
    return;
  }

  // ===== consumer warps (two warpgroups) =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS));
  const int grp = warp >> 2, wq = warp & 3;
  const uint32_t tmem_lane_base = tmem_base + ((uint32_t)(32 * wq) << 16);
  double* my_stg = stg + grp * XT_DBL;
  // the two warps that share a TMEM lane quadrant (w, w+4) split the 64 columns of a tile
  const uint32_t my_cols = 32 * grp;
  {
    uint32_t z[32];
#pragma unroll
    for (int n = 0; n < 32; ++n) z[n] = 0u;
    for (int tau = 0; tau < 6; ++tau) tmem_st32(tmem_lane_base + tau * TMEM_COLS_PER_TILE + my_cols, z);
    tmem_wait_st();
    tmem_fence_before();
    consumer_barrier();
    tmem_fence_after();
  }
  // without CTA barriers between steps, accesses of one X element / staging word by different
  // warps are ordered through the stage ring (a warp can run at most NSTAGE stages ahead of
  // the slowest one); that needs steps of more than NSTAGE stages
  const bool step_sync_always = (((nk4 + 1) >> 1) + nl4) < 2 * NSTAGE;
  for (long long item = blockIdx.x; item < p.nitems; item += gridDim.x) {
    int t, orb;
    decode_item(p, item, t, orb);
    const int4 tr = p.triples[t];
    const uchar4 ob = p.orbits[orb];
    const PtClassTable& tab = c_tab[tr.w][ob.w];
    const int hi = tr.x, hj = tr.y, hk = tr.z;

    for (int s = 0; s < tab.nsteps; ++s) {
      const PtStep& st = tab.steps[s];
      const PtHalf& hf = st.h[grp];
      double acc[2][4][2][2];
#ifdef PT_SKIP_PADDED_B
      // the step's b range (r1) is the padded last particle range with <= 12 valid particles: skip column group 3
      if (sel3(ob.x, ob.y, ob.z, st.r1) == p.d.nr - 1 && p.d.v - TILE * (p.d.nr - 1) <= 12)
        consume_step<3>(acc, pp, nk4, nl4, (p.debug & 1) ? 0 : hf.en, grp, wq, lane);
      else
#endif
      consume_step<4>(acc, pp, nk4, nl4, (p.debug & 1) ? 0 : hf.en, grp, wq, lane);
      // the halves of a step target different X tiles except in orbits with coinciding
      // ranges; only then (or for very short steps) are the two scatters separated by barriers
      const bool sync = step_sync_always || (st.h[0].en && st.h[1].en && st.h[0].tau == st.h[1].tau);
      if (p.debug & 4) {
        // measurement switch: drop the scatter (results are wrong), keeps the accumulators alive
        if (acc[0][0][0][0] == 1.2345e300) my_stg[tid] = acc[1][3][1][1];
      } else if (!sync) {
        if (hf.en) scatter_add(my_stg, tmem_lane_base, acc, hf.tau, hf.q0, hf.q1, grp, wq, lane);
      } else {
        if (grp == 0 && hf.en) scatter_add(my_stg, tmem_lane_base, acc, hf.tau, hf.q0, hf.q1, grp, wq, lane);
        tmem_fence_before();
        consumer_barrier();
        tmem_fence_after();
        if (grp == 1 && hf.en) scatter_add(my_stg, tmem_lane_base, acc, hf.tau, hf.q0, hf.q1, grp, wq, lane);
        tmem_fence_before();
        consumer_barrier();
        tmem_fence_after();
      }
    }

    // ---- epilogue: E_item = sum_tiles sum_x (Xd + Sd)[x] * (sum_nu c_nu Xd[x o nu]) / D[x]
    const double e3 = p.epsi[hi] + p.epsi[hj] + p.epsi[hk];
    const int pm = tab.pmask;
    double e_acc = 0.0;
    const bool sym = tab.ntiles == 6 && !(p.debug & 16);   // generic orbit A > B > C: symmetric epilogue
    // Staging values (singles-term operands, T1 columns, eigenvalues) come from global memory with
    // DRAM latency.  Generic orbit: all six tiles' values are fetched into registers here, before
    // the X copy, so the latency hides behind the copy and phase 1; otherwise one tile ahead.
    double sq[6][3], stv[6][2];
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      stv[u][0] = stv[u][1] = 0.0;
      if (u == 0 || sym)
        epi_stage_load(p, tab, ob, u, hi, hj, hk, pm, tid, sq[u][0], sq[u][1], sq[u][2], stv[u][0], stv[u][1]);
    }

    // ---- X tiles: TMEM -> shared memory (over the drained ring), and clear them for the next item
    tmem_fence_before();
    consumer_barrier();  // all scatters of the item are done, every ring stage has been consumed
    tmem_fence_after();
    {
      const int x0 = lane & 15, x1lo = 2 * wq + (lane >> 4);
      for (int tau = 0; tau < tab.ntiles; ++tau) {
        uint32_t r[32];
        const uint32_t taddr = tmem_lane_base + tau * TMEM_COLS_PER_TILE + my_cols;
        tmem_ld32(taddr, r);
        tmem_wait_ld();
        double* Xt = Xs + tau * XT_DBL;
#pragma unroll
        for (int dd = 0; dd < 16; ++dd) {
          const int d = 16 * grp + dd;
          Xt[xt_index(x0, x1lo + 8 * (d & 1), d >> 1)] = __hiloint2double((int)r[2 * dd + 1], (int)r[2 * dd]);
        }
#pragma unroll
        for (int n = 0; n < 32; ++n) r[n] = 0u;
        tmem_st32(taddr, r);
      }
      tmem_wait_st();
    }


    const double c0 = tab.coef[0], c1 = tab.coef[1], c2 = tab.coef[2], c3 = tab.coef[3], c4 = tab.coef[4],
                 c5 = tab.coef[5];
    const int x0 = tid & 15, x1 = tid >> 4;
    // xt_index(a,b,c) = (a ^ b ^ s(c)) + 16 b + 256 c with s = bitswap13 (XOR-linear): the six
    // permuted reads share three low nibbles, each a per-thread constant XOR a function of x2
    const int l01 = x0 ^ x1, l25 = x1 ^ bitswap13(x0), l34 = x0 ^ bitswap13(x1);
    auto put_stage = [&](double* Qb, double* tb, const double (&q)[3], const double (&t)[2]) {
      Qb[tid] = q[0];
      Qb[256 + tid] = q[1];
      Qb[512 + tid] = q[2];
      if (tid < 48) {
        tb[tid] = t[0];
        tb[48 + tid] = t[1];
      }
    };

    if (p.debug & 8) {
      // measurement switch: no point loops
    } else if (sym) {
      // ================= generic orbit: the six tiles are the images of tile 0 under S3 =================
      // two staging buffers: the Qs/tv area and the (now free) tail of the scatter staging tiles
      double* const Qb[2] = {Qs, Xs + 6 * XT_DBL};
      double* const tb[2] = {tv, Xs + 6 * XT_DBL + 768};
      put_stage(Qb[0], tb[0], sq[0], stv[0]);
      consumer_barrier();  // publishes Xs and tile 0's staging values
      if (!(p.debug & 64)) {
        // ---- phase 1.  D is symmetric, so the point orbit {x o mu} is handled at once: six reads
        // v_pi = Xd[x o pi] (one per tile), Z_mu = sum_nu c_nu v_{mu nu}, ONE reciprocal, the Xd.Z/D
        // part of all six points, and Y[x o mu] = Z_mu / D written back over Xd for the singles part.
        const int ga0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[0][0]);
        const int gb0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[0][1]);
        const int gc0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[0][2]);
        const bool valid01 = (ga0 + x0 < v) && (gb0 + x1 < v);
        const int nv2 = min(16, v - gc0);
        const double d01 = e3 - tv[48 + x0] - tv[64 + x1];
        double* R0 = Xs + 16 * x1;
        double* R1 = Xs + 1 * XT_DBL + 16 * x0;
        double* R2 = Xs + 2 * XT_DBL + 256 * x0;
        double* R3 = Xs + 3 * XT_DBL + 256 * x1;
        double* R4 = Xs + 4 * XT_DBL + 16 * x0 + 256 * x1;
        double* R5 = Xs + 5 * XT_DBL + 16 * x1 + 256 * x0;
        const double cf[6] = {c0, c1, c2, c3, c4, c5};
#pragma unroll EPI_UNROLL
        for (int x2 = 0; x2 < 16; ++x2) {
          const int a01 = l01 ^ bitswap13(x2), a25 = l25 ^ x2, a34 = l34 ^ x2;
          double* A[6] = {R0 + a01 + 256 * x2, R1 + a01 + 256 * x2, R2 + a25 + 16 * x2,
                          R3 + a34 + 16 * x2, R4 + a34, R5 + a25};
          double vv[6];
#pragma unroll
          for (int n = 0; n < 6; ++n) vv[n] = *A[n];
          const double rd = 1.0 / (d01 - tv[80 + x2]);
          double dot = 0.0;
#pragma unroll
          for (int mu = 0; mu < 6; ++mu) {
            double z = 0.0;
#pragma unroll
            for (int nu = 0; nu < 6; ++nu) z += cf[nu] * vv[S3_MUL[mu][nu]];
            dot += vv[mu] * z;
            *A[mu] = z * rd;
          }
          if (valid01 && x2 < nv2) e_acc += dot * rd;
        }
      }
      // ---- phase 2, per tile: singles part sum_x Sd[x] Y[x]; one barrier per tile (ping-pong staging)
#pragma unroll
      for (int tl = 0; tl < ((p.debug & 128) ? 0 : 6); ++tl) {
        consumer_barrier();  // tl = 0: every Y is in place; tl > 0: staging of tile tl is published
        if (tl + 1 < 6) put_stage(Qb[(tl + 1) & 1], tb[(tl + 1) & 1], sq[(tl + 1) % 6], stv[(tl + 1) % 6]);
        const double* Qc = Qb[tl & 1];
        const double* tc = tb[tl & 1];
        const int ga0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[tl][0]);
        const int gb0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[tl][1]);
        const int gc0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[tl][2]);
        const bool valid01 = (ga0 + x0 < v) && (gb0 + x1 < v);
        const int nv2 = min(16, v - gc0);
        const double t0 = 0.5 * tc[x0], t1v = 0.5 * tc[16 + x1];
        const double q2c = 0.5 * Qc[512 + x0 + 16 * x1];
        const double* Q0 = Qc + x1;
        const double* Q1 = Qc + 256 + x0;
        const double* P0 = Xs + tl * XT_DBL + 16 * x1;
#pragma unroll EPI_UNROLL
        for (int x2 = 0; x2 < 16; ++x2) {
          const int a01 = l01 ^ bitswap13(x2);
          const double sd = t0 * Q0[16 * x2] + t1v * Q1[16 * x2] + tc[32 + x2] * q2c;
          if (valid01 && x2 < nv2) e_acc += sd * P0[a01 + 256 * x2];
        }
      }
      if (NS == 2) {
        // second singles term: sum_x Sd2[x] Y[x] with the operands of set 2, tile by tile
        for (int tl = 0; tl < 6; ++tl) {
          double q2[3], t2[2] = {0.0, 0.0};
          epi_stage_load<1>(p, tab, ob, tl, hi, hj, hk, pm, tid, q2[0], q2[1], q2[2], t2[0], t2[1]);
          consumer_barrier();                 // the previous readers of the staging buffer are done
          put_stage(Qb[0], tb[0], q2, t2);
          consumer_barrier();
          const double* Qc = Qb[0];
          const double* tc = tb[0];
          const int ga0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[tl][0]);
          const int gb0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[tl][1]);
          const int gc0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[tl][2]);
          const bool valid01 = (ga0 + x0 < v) && (gb0 + x1 < v);
          const int nv2 = min(16, v - gc0);
          const double t0 = 0.5 * tc[x0], t1v = 0.5 * tc[16 + x1];
          const double q2c = 0.5 * Qc[512 + x0 + 16 * x1];
          const double* Q0 = Qc + x1;
          const double* Q1 = Qc + 256 + x0;
          const double* P0 = Xs + tl * XT_DBL + 16 * x1;
#pragma unroll
          for (int x2 = 0; x2 < 16; ++x2) {
            const int a01 = l01 ^ bitswap13(x2);
            const double sd = t0 * Q0[16 * x2] + t1v * Q1[16 * x2] + tc[32 + x2] * q2c;
            if (valid01 && x2 < nv2) e_acc += sd * P0[a01 + 256 * x2];
          }
        }
      }
    } else {
      // ================= degenerate orbits (<= 3 tiles): per tile, six permuted reads per point =================
      double q[3] = {sq[0][0], sq[0][1], sq[0][2]}, t[2] = {stv[0][0], stv[0][1]};
      for (int tl = 0; tl < tab.ntiles; ++tl) {
        const int ga0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[tl][0]);
        const int gb0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[tl][1]);
        const int gc0 = TILE * sel3(ob.x, ob.y, ob.z, tab.tile_slots[tl][2]);
        put_stage(Qs, tv, q, t);
        consumer_barrier();
        if (tl + 1 < tab.ntiles) epi_stage_load(p, tab, ob, tl + 1, hi, hj, hk, pm, tid, q[0], q[1], q[2], t[0], t[1]);
        const int8_t* nb = tab.nbr[tl];
        const bool valid01 = (ga0 + x0 < v) && (gb0 + x1 < v);
        const double d01 = e3 - tv[48 + x0] - tv[64 + x1];
        const double t0 = 0.5 * tv[x0], t1v = 0.5 * tv[16 + x1];
        const double q2c = 0.5 * Qs[512 + x0 + 16 * x1];
        const double* Q0 = Qs + x1;
        const double* Q1 = Qs + 256 + x0;
        const int nv2 = min(16, v - gc0);   // valid x2 of this tile (>= 1)
        const double* P0 = Xs + tl * XT_DBL + 16 * x1;
        const double* P1 = Xs + nb[1] * XT_DBL + 16 * x0;
        const double* P2 = Xs + nb[2] * XT_DBL + 256 * x0;
        const double* P3 = Xs + nb[3] * XT_DBL + 256 * x1;
        const double* P4 = Xs + nb[4] * XT_DBL + 16 * x0 + 256 * x1;
        const double* P5 = Xs + nb[5] * XT_DBL + 16 * x1 + 256 * x0;
#pragma unroll EPI_UNROLL
        for (int x2 = 0; x2 < 16; ++x2) {
          // nu = Permutation<3>(n): (x o nu)_m = x_{nu(m)}; nbr[tl][0] == tl (identity)
          const int a01 = l01 ^ bitswap13(x2), a25 = l25 ^ x2, a34 = l34 ^ x2;
          const double xd = P0[a01 + 256 * x2];
          double zn = c0 * xd;
          zn += c1 * P1[a01 + 256 * x2];
          zn += c2 * P2[a25 + 16 * x2];
          zn += c3 * P3[a34 + 16 * x2];
          zn += c4 * P4[a34];
          zn += c5 * P5[a25];
          const double sd = t0 * Q0[16 * x2] + t1v * Q1[16 * x2] + tv[32 + x2] * q2c;
          const double dd = d01 - tv[80 + x2];
          if (valid01 && x2 < nv2) e_acc += (xd + sd) * zn / dd;
        }
        consumer_barrier();
        if (NS == 2) {
          // second singles term of this tile: sum_x Sd2[x] Z[x] / D[x] with the operands of set 2
          double q2[3], t2[2] = {0.0, 0.0};
          epi_stage_load<1>(p, tab, ob, tl, hi, hj, hk, pm, tid, q2[0], q2[1], q2[2], t2[0], t2[1]);
          put_stage(Qs, tv, q2, t2);
          consumer_barrier();
          const double t0b = 0.5 * tv[x0], t1b = 0.5 * tv[16 + x1];
          const double q2b = 0.5 * Qs[512 + x0 + 16 * x1];
#pragma unroll
          for (int x2 = 0; x2 < 16; ++x2) {
            const int a01 = l01 ^ bitswap13(x2), a25 = l25 ^ x2, a34 = l34 ^ x2;
            double zn = c0 * P0[a01 + 256 * x2];
            zn += c1 * P1[a01 + 256 * x2];
            zn += c2 * P2[a25 + 16 * x2];
            zn += c3 * P3[a34 + 16 * x2];
            zn += c4 * P4[a34];
            zn += c5 * P5[a25];
            const double sd = t0b * Q0[16 * x2] + t1b * Q1[16 * x2] + tv[32 + x2] * q2b;
            const double dd = d01 - tv[80 + x2];
            if (valid01 && x2 < nv2) e_acc += sd * zn / dd;
          }
          consumer_barrier();
        }
      }
    }
    // warp-shuffle reduction, then one atomic per item
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) e_acc += __shfl_down_sync(0xffffffffu, e_acc, off);
    if (lane == 0) red[warp] = e_acc;
    tmem_fence_before();
    consumer_barrier();
    tmem_fence_after();
    if (tid == 0) {
      mbar_arrive(go_bar);  // Xs has been read by everyone: the producer may refill the ring
      if (p.sync_ctr) {
        __threadfence();
        atomicAdd(p.sync_ctr, 1u);  // this CTA has finished one more item (item-round barrier)
      }
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NCONSUMER_WARPS; ++w) s += red[w];
      p.e_item[item] = s;   // one plain store per item; per-triple sums are a fixed-order second pass
    }
  }
  consumer_barrier();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

__global__ void __launch_bounds__(FUSED_THREADS, 1) pt_fused_kernel(const FusedParams p) { fused_body<1>(p); }
__global__ void __launch_bounds__(FUSED_THREADS, 1) pt_fused_kernel2(const FusedParams p) { fused_body<2>(p); }

// ------------------------------------------------ debug: one W tile via the main loop
__global__ void __launch_bounds__(FUSED_THREADS, 1) pt_w_tile_kernel(const FusedParams p, WTileJob job,
                                                                    double* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *Xs, *ring, *stg, *Qs, *tv, *red;
  uint64_t* bars;
  uint32_t* tmem_slot;
  carve_smem(smem_raw, Xs, ring, stg, Qs, tv, red, bars, tmem_slot);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(smem_u32(bars + s), 1);
      mbar_init(smem_u32(bars + NSTAGE + s), NCONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  Pipe pp;
  pp.ring = smem_u32(ring);
  pp.full = smem_u32(bars);
  pp.empty = smem_u32(bars + NSTAGE);
  pp.slot = 0;
  pp.phase = 0;
  StepSrc src;
  src.v = p.Vt + vt_tile_off(p.d, p.vslot ? p.vslot[job.z] : job.z, job.Q, job.R);
  src.t[0] = src.t[1] = p.Tt + tt_panel_off(p.d, job.x, job.y, job.P);
  src.hh[0] = src.hh[1] = p.T2h + t2h_block_off(p.d, job.x, job.P, job.Q);
  src.u[0] = src.u[1] = p.Ut + ut_panel_off(p.d, job.y, job.z, job.R);
  src.en[0] = 1;
  src.en[1] = 0;
  if (warp > NCONSUMER_WARPS) return;
  if (warp == NCONSUMER_WARPS) {
    const bool leader = elect_one();
    const int nst = ((p.d.nk4 + 1) >> 1) + p.d.nl4;
    for (int j = 0; j < nst; ++j) {
      mbar_wait(pp.empty + 8 * pp.slot, pp.phase ^ 1);
      if (leader) issue_stage(src, j, p.d.nk4, pp.ring + pp.slot * (STAGE_DBL * 8), pp.full + 8 * pp.slot);
      pp.advance();
    }
    return;
  }
  const int grp = warp >> 2, wq = warp & 3;
  double acc[2][4][2][2];
  consume_step<4>(acc, pp, p.d.nk4, p.d.nl4, grp == 0, grp, wq, lane);
  if (grp != 0) return;
#pragma unroll
  for (int mf = 0; mf < 2; ++mf)
#pragma unroll
    for (int bi = 0; bi < 4; ++bi)
#pragma unroll
      for (int co = 0; co < 2; ++co)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int la = 8 * mf + (lane >> 2), lb = wq + 4 * bi, lc = 8 * co + 2 * (lane & 3) + e;
          out[la + 16 * (lb + 16 * lc)] = acc[mf][bi][co][e];
        }
}

// ------------------------------------------------------------------- launchers
cudaError_t fused_configure(int* smem_bytes_out) {
  cudaError_t e = cudaFuncSetAttribute(pt_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       FUSED_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pt_fused_kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pt_w_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           FUSED_SMEM_BYTES);
  if (smem_bytes_out) *smem_bytes_out = FUSED_SMEM_BYTES;
  return e;
}

cudaError_t launch_fused(const FusedParams& p, int grid, cudaStream_t s) {
  if (p.nitems <= 0) return cudaSuccess;
  if (p.sync_ctr) {
    // the item-round barrier spins on other CTAs: co-residency must be guaranteed, not assumed
    FusedParams pc = p;
    void* args[] = {&pc};
    return cudaLaunchCooperativeKernel(p.t1b ? (const void*)pt_fused_kernel2 : (const void*)pt_fused_kernel, dim3(grid),
                                       dim3(FUSED_THREADS), args, FUSED_SMEM_BYTES, s);
  }
  if (p.t1b) pt_fused_kernel2<<<grid, FUSED_THREADS, FUSED_SMEM_BYTES, s>>>(p);
  else pt_fused_kernel<<<grid, FUSED_THREADS, FUSED_SMEM_BYTES, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_w_tile(const FusedParams& p, WTileJob job, double* d_out, cudaStream_t s) {
  pt_w_tile_kernel<<<1, FUSED_THREADS, FUSED_SMEM_BYTES, s>>>(p, job, d_out);
  return cudaGetLastError();
}

// ------------------------------------------------ FP64 issue-rate microbenchmarks
// mode 0: DMMA.8x8x4 (8 independent accumulator fragments per warp)
// mode 1: DFMA (16 independent chains per thread)
__global__ void __launch_bounds__(1024) bench_fp64_kernel(int mode, int iters, double* sink,
                                                          unsigned long long* cycles) {
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  double c[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = 0.0;
  const long long t0 = clock64();
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma(c[j][0], c[j][1], a, b);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        c[j][0] = fma(a, c[j][0], b);
        c[j][1] = fma(b, c[j][1], a);
      }
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  if (s == 123.456) sink[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

cudaError_t launch_bench_fp64(int mode, int blocks, int warps, int iters, double* d_sink,
                              unsigned long long* d_cycles, cudaStream_t s) {
  bench_fp64_kernel<<<blocks, warps * 32, 0, s>>>(mode, iters, d_sink, d_cycles);
  return cudaGetLastError();
}

}  // namespace pt
