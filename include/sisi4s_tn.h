/* sisi4s_tn.h -- C ABI of the device tensor-contraction engine inside libsisi4s_pt.so.
 *
 * It executes the index-string tensor statements the reference writes against Cyclops CTF
 *   C["abij"] += alpha * A["acik"] * B["cbkj"];          (CTF::Tensor::operator[] / Idx_Tensor)
 * on one B200: FP64, dense, column-major tensors that live in device memory.  It serves the steps
 * next to the (T) hot path:
 *   N1  every block of CoulombIntegralsFromVertex  (reference src/algorithms/
 *       CoulombIntegralsFromVertex.cxx:390-560),
 *   N3  the closed-shell CCSD residuum and solver loop (src/algorithms/
 *       CcsdEnergyFromCoulombIntegralsReference.cxx:29-295, ClusterSinglesDoublesAlgorithm.cxx:37-128,
 *       302-331, src/mixers/DiisMixer.cxx:103-181) -- driven from sisi4s_b200/ccsd.py.
 * Contractions run on the library's own FP64 tensor-core GEMM (pt_pack.cu: vertex_gemm_kernel); no
 * cuBLAS / cuTENSOR, no CPU fallback.
 *
 * Conventions: tensors are addressed by small integer ids; index strings name one letter per
 * dimension, first dimension fastest (the CTF global layout, docs/manual.org:320); functions return
 * TN_OK or a negative TnStatus, tn_last_error() has the message for the calling thread.
 */
#ifndef SISI4S_TN_H
#define SISI4S_TN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct TnHandle_ *tn_handle_t;

typedef enum TnStatus {
  TN_OK = 0,
  TN_ERR_INVALID = -1,
  TN_ERR_CUDA = -2,
  TN_ERR_NOMEM = -4,
  TN_ERR_UNSUPPORTED = -5
} TnStatus;

int tn_create(tn_handle_t *out, int device);
int tn_destroy(tn_handle_t h);
const char *tn_last_error(void);

/* a zero-initialised dense tensor of `ndim` <= 8 dimensions (ndim = 0: a scalar); *id names it */
int tn_tensor(tn_handle_t h, int ndim, const int64_t *lens, int *id);
int tn_free(tn_handle_t h, int id);
/* whole-tensor copies from / to caller-owned host memory (column-major) */
int tn_upload(tn_handle_t h, int id, const double *host);
int tn_download(tn_handle_t h, int id, double *host);

/* C[ic] = alpha * sum A[ia] * B[ib] + beta * C[ic], summed over the indices that appear in both
 * operands and not in the result (the reference's CTF statements `C[..] = / += a * A[..] * B[..]`
 * with beta = 0 / 1).  Every index must appear in exactly two of the three strings. */
int tn_contract(tn_handle_t h, double alpha, int a, const char *ia, int b, const char *ib, double beta, int c,
                const char *ic);
/* C[ic] = alpha * A[ia] + beta * C[ic]: index permutation / axpy / scaling (`C["aibj"] = A["abij"]`) */
int tn_add(tn_handle_t h, double alpha, int a, const char *ia, double beta, int c, const char *ic);
/* sum_x A[x] B[x] over the storage order (FockVector::dot), fixed summation order */
int tn_dot(tn_handle_t h, int a, int b, double *out);
/* R = -(R - shift * T) / (sum eps_a - sum eps_i + shift) for R, T of shape [v,(v,)o,(o)]:
 * ClusterSinglesDoublesAlgorithm::estimateAmplitudesFromResiduum (:302-331) */
int tn_excitation_divide(tn_handle_t h, int r, int t, int epsi, int epsa, double shift);
/* algorithmic FLOP of the GEMMs, bytes moved by the gathers / permuted adds, kernels launched */
int tn_get_stats(tn_handle_t h, double *flops, double *bytes, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* SISI4S_TN_H */
