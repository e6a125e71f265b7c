/* sisi4s_ccsd.h -- C ABI of the device closed-shell CCSD solver inside libsisi4s_pt.so
 * (the step in front of the (T) path).
 *
 * Replaces, for real closed-shell inputs, the reference's
 *   CcsdEnergyFromCoulombIntegralsReference::getResiduum   (src/algorithms/
 *       CcsdEnergyFromCoulombIntegralsReference.cxx:29-295, Hirata et al.),
 *   ClusterSinglesDoublesAlgorithm::run / getEnergy / estimateAmplitudesFromResiduum
 *       (src/algorithms/ClusterSinglesDoublesAlgorithm.cxx:37-128, 130-205, 302-331),
 *   LinearMixer / DiisMixer (src/mixers/LinearMixer.cxx:31-49, DiisMixer.cxx:103-181).
 * Every CTF statement of the residuum is one statement of the device tensor engine
 * (include/sisi4s_tn.h) with the same index strings; the solver loop, the mixers and the convergence
 * test run on the host side of the library, all tensors stay on the GPU.
 *
 * A sisi4s Algorithm subclass gathers its CTF tensors with Tensor::read_all and calls these entry points
 * (sisi4s_b200/csrc/CcsdEnergyFromCoulombIntegralsGpu.cxx); sisi4s_b200/ccsd.py binds the same ABI.
 * Arrays: FP64, dense, column-major in the CTF index order, caller-owned host memory.  Functions return
 * 0 or a negative TnStatus (include/sisi4s_tn.h); tn_last_error() has the message.  No CPU fallback.
 */
#ifndef SISI4S_CCSD_H
#define SISI4S_CCSD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CcsdHandle_ *ccsd_handle_t;

typedef enum CcsdMixer { CCSD_LINEAR_MIXER = 0, CCSD_DIIS_MIXER = 1 } CcsdMixer;

typedef struct CcsdOptions {
  int32_t mixer;                  /* CcsdMixer; the reference's default is the LinearMixer (:48)      */
  int32_t max_residua;            /* DiisMixer: maxResidua (DiisMixer.cxx:55), default 4              */
  double mixing_ratio;            /* LinearMixer: mixingRatio (LinearMixer.cxx:15), default 1.0        */
  int32_t max_iterations;         /* maxIterations, default 16                                        */
  int32_t reserved;
  double energy_convergence;      /* |(e - e_prev) / e| <  energyConvergence      (:103), default 1e-6 */
  double amplitudes_convergence;  /* dT.dT / T.T        <  amplitudesConvergence^2 (:104-106), 1e-5    */
  double level_shift;             /* levelShift (:305), default 0                                     */
} CcsdOptions;

typedef struct CcsdResult {
  double energy;                  /* CcsdEnergy: direct + exchange (getEnergy :160-178)               */
  double direct, exchange;
  int32_t iterations;
  int32_t converged;              /* 0: maxIterations reached -- a WARNING in the reference (:120-124), not an error */
  double flops;                   /* algorithmic FLOP of the contractions                             */
  int64_t kernel_launches;
} CcsdResult;

int ccsd_create(ccsd_handle_t *out, int o, int v, int device);
int ccsd_destroy(ccsd_handle_t h);
void ccsd_default_options(CcsdOptions *opt);

/* HoleEigenEnergies[o], ParticleEigenEnergies[v] (calculateExcitationEnergies :343-365) */
int ccsd_set_eigenenergies(ccsd_handle_t h, const double *epsi, const double *epsa);
/* one of the integral blocks getResiduum reads (:49,136-140): name in
 * "PPHH" [v,v,o,o], "PHPH" [v,o,v,o], "HHHH" [o,o,o,o], "HHHP" [o,o,o,v], "PPPH" [v,v,v,o], "PPPP" [v,v,v,v] */
int ccsd_set_integrals(ccsd_handle_t h, const char *name, const double *block);
/* Alternative to the six ccsd_set_integrals calls: CoulombVertex[NF,Np,Np] as real and imaginary parts;
 * the six blocks are built on the device with the index strings of CoulombIntegralsFromVertex.cxx:395-431 */
int ccsd_set_vertex(ccsd_handle_t h, int nf, int np, const double *gamma_re, const double *gamma_im);
/* copies one of the six blocks back (the CoulombIntegralsFromVertex step's outputs) */
int ccsd_get_integrals(ccsd_handle_t h, const char *name, double *block);

/* initialSinglesAmplitudes / initialDoublesAmplitudes (createAmplitudes :207-237); default zero */
int ccsd_set_amplitudes(ccsd_handle_t h, const double *t1, const double *t2);
/* one getResiduum(iteration, amplitudes) evaluation on the current amplitudes (testing / custom loops):
 * r1[v,o], r2[v,v,o,o] to host memory */
int ccsd_residuum(ccsd_handle_t h, int iteration, double *r1, double *r2);
/* ClusterSinglesDoublesAlgorithm::run<double> (:37-128) */
int ccsd_solve(ccsd_handle_t h, const CcsdOptions *opt, CcsdResult *result);
/* CcsdSinglesAmplitudes[v,o], CcsdDoublesAmplitudes[v,v,o,o] (storeAmplitudes :289-300); either may be NULL */
int ccsd_get_amplitudes(ccsd_handle_t h, double *t1, double *t2);

#ifdef __cplusplus
}
#endif
#endif /* SISI4S_CCSD_H */
