/* sisi4s_pt.h -- C ABI of the B200-native closed-shell CCSD(T) perturbative
 * triples step (libsisi4s_pt.so).
 *
 * This is the drop-in boundary for ONE step of alejandrogallo/sisi4s: the
 * algorithm `CcsdPerturbativeTriples`
 *   (reference src/algorithms/CcsdPerturbativeTriples.cxx:119-248, YAML contract
 *    integration-tests/bench/todo/CcsdPerturbativeTriples/in.yaml)
 * and its PPPH-integral spelling `PerturbativeTriples`
 *   (reference src/algorithms/PerturbativeTriples.cxx:172-239, YAML contract
 *    integration-tests/bench/todo/PerturbativeTriples/in.yaml).
 * A sisi4s `Algorithm` subclass gathers the CTF tensors once with
 * `Tensor::read_all` and hands the dense buffers to these entry points
 * (sisi4s_b200/csrc/CcsdPerturbativeTriplesGpu.cxx).
 *
 * Conventions
 *  - every array is IEEE FP64, dense, COLUMN-MAJOR in the reference's CTF index
 *    order (first index fastest; docs/manual.org:320), caller-owned host memory;
 *  - o = number of holes (HoleEigenEnergies->lens[0]), v = number of particles;
 *  - functions return PT_OK (0) or a negative PtStatus; pt_last_error() gives a
 *    message for the calling thread.  No C++ types or exceptions cross the ABI;
 *  - one handle per host thread / GPU.  The library owns all device memory;
 *  - there is no CPU fallback: without a CUDA device pt_create fails.
 */
#ifndef SISI4S_PT_H
#define SISI4S_PT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct PtHandle_ *pt_handle_t;

typedef enum PtStatus {
  PT_OK = 0,
  PT_ERR_INVALID = -1,   /* bad argument / call order                          */
  PT_ERR_CUDA = -2,      /* CUDA runtime error (message in pt_last_error)      */
  PT_ERR_MISSING = -3,   /* an input tensor was not set before pt_run          */
  PT_ERR_NOMEM = -4,     /* device or host memory exhausted                    */
  PT_ERR_UNSUPPORTED = -5
} PtStatus;

/* which device implementation pt_run uses */
typedef enum PtEngine {
  PT_ENGINE_FUSED = 0,   /* product path: DMMA tiles + on-chip symmetrise/divide/reduce */
  PT_ENGINE_NAIVE = 1    /* literal on-device restatement (validation only; needs keep_raw) */
} PtEngine;

typedef struct PtStats {
  double seconds_run;        /* device time of the last pt_run (CUDA events; incl. list copies) */
  double seconds_kernel;     /* device time of the triples kernel(s) alone in the last pt_run   */
  double seconds_upload;     /* host->device copies + packing since pt_create          */
  double flops_algorithmic;  /* 2 o^3 v^3 (v+o) scaled to the triples of the last run */
  double bytes_h2d;          /* host->device bytes since pt_create                     */
  double bytes_d2h;          /* device->host bytes since pt_create                     */
  double device_bytes;       /* device memory currently held by the handle             */
  int64_t kernel_launches;   /* kernels launched by this handle since pt_create        */
  int64_t triples_run;       /* sorted triples processed by the last pt_run            */
  int32_t sm_count;
  int32_t reserved;
  int64_t slab_loads;        /* PPPH slabs (re)built or re-uploaded on demand (slab_slots < o) */
  int64_t groups_staged;     /* hole-block groups whose T2 / PPHH / HHHP blocks were staged (hole_block)  */
  double bytes_pinned;       /* caller-owned host memory page-locked by the library (pin_host)            */
} PtStats;

/* ---- lifecycle ---------------------------------------------------------- */
/* o, v: dimensions; device: CUDA ordinal.  Replaces the constructor +
 * sliceTensors() of the reference class (CcsdPerturbativeTriples.cxx:16-79).  */
int pt_create(pt_handle_t *out, int o, int v, int device);
/* Engine for a hole SUBSET of a larger problem (out-of-core driver for shapes whose T2 / PPHH
 * tensors exceed one GPU, BASELINE configs[4]): o_act active holes -- the hole indices i,j,k of
 * the triples that are run, in ascending order of the full problem -- and o_all holes in the
 * contraction sum_l of getDoublesContribution (:94).  Shapes then are
 *   HoleEigenEnergies[o_act], CcsdSinglesAmplitudes[v,o_act], PPHHCoulombIntegrals[v,v,o_act,o_act],
 *   CcsdDoublesAmplitudes[v,v,o_act,o_act]        (pt_set_doubles: particle term T2[a,d,x,y]),
 *   CcsdDoublesAmplitudes[v,v,o_act,o_all]        (pt_set_doubles_hole: hole term T2[a,b,x,l]),
 *   HHHPCoulombIntegrals[o_act,o_act,o_all,v], PPPHCoulombIntegrals[v,v,v,o_act].
 * The caller slices the full tensors; E_t of a triple is the same number as in the full problem.
 * pt_create(o, v) == pt_create_ex(o, o, v).  PT_ENGINE_NAIVE needs o_act == o_all.             */
int pt_create_ex(pt_handle_t *out, int o_act, int o_all, int v, int device);
int pt_destroy(pt_handle_t h);
const char *pt_last_error(void);
const char *pt_version(void);

/* options: "engine" (PtEngine), "keep_raw" (0/1: keep unpacked copies on the
 * device, required by PT_ENGINE_NAIVE and the debug entry points; set BEFORE
 * the tensors), "grid" (CTAs of the fused kernel, 0 = one per SM),
 * "slab_slots" (S: hole-blocked residency of the PPPH integrals -- only S >= 3
 * of the o slabs V[:,:,:,k] are resident at a time, for shapes whose v^3 o
 * tensor exceeds HBM (BASELINE configs[4]); pt_run then walks the sorted
 * triples by hole blocks of width S/3 and (re)builds slabs on demand from the
 * resident vertex or re-uploads them from the pt_set_ppph_host tensor; 0 = all
 * resident; set BEFORE the PPPH integrals / vertex);
 * "hole_block" (b: hole-blocked OUT-OF-CORE mode for shapes whose hole-indexed
 * tensors exceed one GPU -- BASELINE configs[4], o=100 v=800: T2, its second
 * packing and PPHH are 51 GB each, PPPH 410 GB.  Set FIRST, on a pt_create
 * handle.  The setters then take the FULL tensors of the problem but keep only
 * what fits: pt_set_doubles / pt_set_pphh / pt_set_ppph_host record the
 * caller-owned host pointers (they must stay valid until pt_destroy),
 * pt_set_hhhp uploads the o^3 v tensor once, pt_set_vertex keeps the vertex
 * image resident.  pt_run / pt_run_list walk their triples by hole-block
 * triples (I<=J<=K) of width b: before each group's launch the T2 / PPHH /
 * HHHP blocks of its <= 3b active holes are staged into device buffers that
 * were allocated once for 3b holes, and the group's PPPH slabs are made
 * resident (rebuilt from the vertex, like the reference does per triple,
 * CcsdPerturbativeTriples.cxx:89-92, or uploaded from the host tensor; LRU
 * over slab_slots >= 3b slots, K fastest so consecutive groups share the slabs
 * of blocks I and J).  E_t of a triple is bitwise the number the all-resident
 * run gives.  pt_partition ranges work unchanged: a contiguous range of the
 * (i,j,k) enumeration touches few I blocks);
 * "async_upload" (1: the pt_set_* calls only enqueue their copies and packing
 * kernels on the handle's stream and return; the host buffers must then stay
 * valid and unmodified until pt_sync or pt_run returns.  0 (default): every
 * setter returns after its copy has completed);
 * "pin_host" (1: page-lock the caller's large tensors in hole_block mode so
 * the per-group block copies run at full DMA speed);
 * "particle_contraction" (vd: length of the particle contraction sum_d when it differs from v --
 * the doubles then have shape [v,vd,o,o] and the PPPH slabs [v,v,vd]; used by the complex triples
 * driver, which stacks real and imaginary parts along d; set before any input).              */
int pt_set_option(pt_handle_t h, const char *key, int64_t value);
/* waits for everything the setters enqueued (async_upload) and reports their errors */
int pt_sync(pt_handle_t h);
/* Device memory (bytes) a handle for (o, v) will hold with the given slab_slots / hole_block
 * options (0 = unset), PPPH given as a tensor; host-only, no GPU needed.  This is the number the
 * plugin's dryRun reports in place of the reference's CTF estimate
 * (CcsdPerturbativeTriples.cxx:250-284).                                      */
int64_t pt_estimate_device_bytes(int o, int v, int slab_slots, int hole_block);

/* ---- inputs (names = the reference's YAML argument keys) ------------------ */
/* HoleEigenEnergies[o], ParticleEigenEnergies[v]
 * (getEnergyDenominator, CcsdPerturbativeTriples.cxx:98-117)                  */
int pt_set_eigenenergies(pt_handle_t h, const double *epsi, const double *epsa);
/* CcsdSinglesAmplitudes[v,o]  ("ai", ClusterSinglesDoublesAlgorithm.cxx:42-45) */
int pt_set_singles(pt_handle_t h, const double *t1);
/* Optional SECOND singles term: the step then uses S = 1/2 (T1 (x) PPHH + t1b (x) vabij_b) in
 * getSinglesContribution (same shapes as pt_set_singles / pt_set_pphh).  The complex closed-shell
 * step (reference CcsdPerturbativeTriplesComplex.cxx:142-148) needs it: the real part of
 * 1/2 Tai Vabij is 1/2 (Re T Re V - Im T Im V).                                   */
int pt_set_singles_pair(pt_handle_t h, const double *t1b, const double *vabij_b);
/* CcsdDoublesAmplitudes[v,v,o,o] ("abij")                                     */
int pt_set_doubles(pt_handle_t h, const double *t2);
/* pt_create_ex engines only: the doubles amplitudes of the HOLE term, T2[a,b,x,l] with x over
 * the active and l over all holes, [v,v,o_act,o_all].  (With o_act == o_all pt_set_doubles
 * serves both terms and this call is not needed.)                              */
int pt_set_doubles_hole(pt_handle_t h, const double *t2_xl);
/* PPHHCoulombIntegrals[v,v,o,o] (getSinglesContribution, :81-85)              */
int pt_set_pphh(pt_handle_t h, const double *vabij);
/* HHHPCoulombIntegrals[o,o,o,v] (hole term of getDoublesContribution, :94)    */
int pt_set_hhhp(pt_handle_t h, const double *vijka);
/* PPPHCoulombIntegrals[v,v,v,o] slabs [:,:,:,k0:k1) -- `slab` points at the
 * first element of slab k0 (v^3 doubles per slab).  May be called repeatedly
 * with disjoint ranges so the caller never holds more than a few slabs
 * (PerturbativeTriples.cxx:176,190).                                          */
int pt_set_ppph_slabs(pt_handle_t h, int k0, int k1, const double *slab);
/* The whole PPPHCoulombIntegrals[v,v,v,o] tensor in caller-owned HOST memory.
 * Without slab_slots it is uploaded at once (= pt_set_ppph_slabs(h,0,o,..)).
 * With slab_slots < o only the pointer is recorded: it must stay valid until
 * pt_destroy, and pt_run uploads the slabs its current hole blocks need.
 * With "async_upload" = 1 (and everything resident) the upload is deferred to
 * pt_run as well: only the slabs of the holes its triples touch are copied
 * (a rank of a multi-GPU run needs the slabs k >= its smallest hole only), on
 * a second stream, and the triples run in waves ordered by their largest hole
 * so that all but the first small wave's copies hide behind the kernel.       */
int pt_set_ppph_host(pt_handle_t h, const double *vabci);
/* Alternative to pt_set_ppph_slabs: CoulombVertex Gamma[NF,Np,Np] complex,
 * given as separate real and imaginary parts (fromComplexTensor,
 * CcsdPerturbativeTriples.cxx:48-78).  PPPH is built on the device exactly as
 * CoulombIntegralsFromVertex.cxx:430-431; particles are the last v states.
 * With slab_slots < o the vertex stays resident on the device and slabs are
 * rebuilt when needed, like the reference does per triple (:89-92).           */
int pt_set_vertex(pt_handle_t h, int nf, int np, const double *gamma_re, const double *gamma_im);
/* After pt_set_vertex: build PPHHCoulombIntegrals (Vabij["abij"] = G["Gai"] G["Gbj"],
 * CoulombIntegralsFromVertex.cxx:402-403) and HHHPCoulombIntegrals (Vijka["ijka"] = G["Gik"] G["Gaj"],
 * :416-417) on the device from the resident vertex and use them as the step's inputs, so the
 * CoulombVertex contract needs neither tensor from the host (in hole_block mode PPHH is rebuilt
 * per group).  Replaces pt_set_pphh + pt_set_hhhp.                              */
int pt_use_vertex_integrals(pt_handle_t h);
/* CoulombIntegralsFromVertex on the device: after pt_set_vertex, computes one
 * real integral block from the resident vertex with the FP64 tensor-core GEMM and copies it to
 * `out` (caller-owned host memory, column-major):
 *   "PPHH" [v,v,o,o] (:402-403), "HHHP" [o,o,o,v] (:416-417), "PPPH" [v,v,v,o] (:430-431, one hole
 *   slab at a time).  Not available on hole-subset engines (pt_create_ex).        */
int pt_vertex_integrals(pt_handle_t h, const char *block, double *out);

/* ---- run ------------------------------------------------------------------ */
/* number of sorted hole triples i<=j<=k = o(o+1)(o+2)/6, enumerated in the
 * reference's loop order (CcsdPerturbativeTriples.cxx:156-158)                */
int64_t pt_num_triples(int o);
/* contiguous share [begin,end) of the sorted-triple enumeration for `rank` of
 * `nranks`, balanced by the number of distinct hole permutations (6/3/3/1)    */
int pt_partition(int o, int nranks, int rank, int64_t *begin, int64_t *end);
/* Host-only preview of the hole-block walk pt_run does with option "hole_block" = b over the sorted
 * triples [begin,end): number of groups (= launches), the largest number of active holes of a group
 * (<= 3b: what the device buffers are sized for) and the number of PPPH slab (re)builds with 3b slots.
 * No GPU needed; used to size a run (at o=100, b=6: 969 groups for the whole problem).          */
int pt_plan_hole_blocks(int o, int hole_block, int64_t begin, int64_t end, int64_t *n_groups,
                        int32_t *max_active_holes, int64_t *slab_loads);
/* Computes sum_{t in [begin,end)} E_t, the (T) energy contribution of those
 * sorted triples (the body of the reference loop, :159-216).  e_triples: the
 * sum; e_per_triple: NULL or end-begin doubles.  The caller adds CcsdEnergy
 * (:241-247) and, across GPUs, all-reduces the scalar.                        */
int pt_run(pt_handle_t h, int64_t begin, int64_t end, double *e_triples, double *e_per_triple);
/* Same for an explicit list of n sorted-triple indices (reference enumeration over the ACTIVE
 * holes), in any order, e.g. the triples of one hole-block triple.  e_per_triple: NULL or n.   */
int pt_run_list(pt_handle_t h, int64_t n, const int64_t *triples, double *e_triples, double *e_per_triple);
int pt_get_stats(pt_handle_t h, PtStats *stats);

/* ---- complex closed-shell step -------------------------------- */
/* Re E(T) of CcsdPerturbativeTriplesComplex::Calculator<complex>::calculate (reference
 * src/algorithms/CcsdPerturbativeTriplesComplex.cxx:166-271, particle term :341-348, hole term from
 * PHHHCoulombIntegrals["clkj"] :135-140, conj(DV / Delta) :224-230).  Every complex tensor is given as its
 * real and imaginary part (column-major, the reference's index order): T1[v,o], T2[v,v,o,o],
 * PPHH[v,v,o,o], PHHH[v,o,o,o], CoulombVertex[NF,Np,Np].  One call = two passes of the real step's fused
 * kernel with Re/Im stacked along the contracted indices (4x the real step's work).  e_per_triple: NULL or
 * o(o+1)(o+2)/6 doubles.                                                                         */
int pt_complex_triples(int o, int v, int device, const double *epsi, const double *epsa, const double *t1_re,
                       const double *t1_im, const double *t2_re, const double *t2_im, const double *pphh_re,
                       const double *pphh_im, const double *phhh_re, const double *phhh_im, int nf, int np,
                       const double *gamma_re, const double *gamma_im, double *e_triples, double *e_per_triple);

/* Spin-orbital (unrestricted) triples, UPerturbativeTriples::run (reference
 * src/algorithms/UPerturbativeTriples.cxx:19-305): full-tensor form on ANTISYMMETRISED integrals
 * PPHH[v,v,o,o], HHHP[o,o,o,v], PPPH[v,v,v,o] and spin-orbital amplitudes; returns the triples energy alone
 * (the reference's PerturbativeTriplesEnergy, :305).  Three v^3 o^3 device tensors: small systems, as in the
 * reference.                                                                                     */
int pt_spin_orbital_triples(int o, int v, int device, const double *epsi, const double *epsa, const double *tai,
                            const double *tabij, const double *vabij, const double *vijka, const double *vabci,
                            double *e_triples);

/* ---- debug / measurement helpers (not part of the drop-in contract) ------- */
/* one 16x16x16 tile of W_{xyz}[a,b,c] (getDoublesContribution) computed by the
 * fused kernel's own main loop; out[la + 16*(lb + 16*lc)]                     */
int pt_debug_w_tile(pt_handle_t h, int x, int y, int z, int ra, int rb, int rc, double *out);
/* FP64 issue-rate microbenchmarks on the handle's device: mode 0 = DMMA.8x8x4,
 * mode 1 = DFMA.  Returns achieved TFLOP/s over `iters` inner iterations.     */
int pt_bench_fp64(pt_handle_t h, int mode, int warps_per_sm, int iters, double *tflops, double *sm_mhz_est);
/* kernel-only timing of the integrals-from-vertex GEMM on the resident vertex: what = 0 one packed
 * PPPH hole slab (2 * 2NF * v^3 FLOP), 1 = the PPHH block (2 * 2NF * v^2 o^2 FLOP); average device
 * seconds per build over `reps` launches (CUDA events) and the algorithmic FLOP of one build.   */
int pt_bench_vertex_gemm(pt_handle_t h, int what, int reps, double *seconds, double *flop);

#ifdef __cplusplus
}
#endif
#endif /* SISI4S_PT_H */
