// ccsd_solver.cu -- closed-shell CCSD amplitude solver on the device (include/sisi4s_ccsd.h).
//
// Host logic only: every tensor statement goes through the device tensor engine (tn_engine.cu, the
// library's FP64 tensor-core GEMM); this file holds the reference's statement list and solver loop.
//
//   residuum      CcsdEnergyFromCoulombIntegralsReference::getResiduum   (reference
//                 src/algorithms/CcsdEnergyFromCoulombIntegralsReference.cxx:29-295): each CTF statement is
//                 one line of the table-like code below, same index strings, same order.  Products of three
//                 tensors (V * Tai * Tai) are evaluated pairwise through the intermediates Y / Zki, or through
//                 Xabij = Tabij + Tai Tbj (the tensor the reference itself builds at :73-74), which merges the
//                 `.. * Tabij` and `.. * Tai * Tai` statements that differ only in that factor.
//   loop          ClusterSinglesDoublesAlgorithm::run (:37-128), getEnergy (:160-178),
//                 estimateAmplitudesFromResiduum (:302-331)
//   mixers        LinearMixer (src/mixers/LinearMixer.cxx:31-49), DiisMixer (src/mixers/DiisMixer.cxx:103-181;
//                 its (count+1)^2 dsysv_ solve is a small Gaussian elimination on the host)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/sisi4s_ccsd.h"
#include "../../include/sisi4s_tn.h"

namespace pt {
int tn_record_error(int code, const char* message);   // tn_engine.cu: what tn_last_error() returns
int tn_on_exception(const char* where);
}  // namespace pt

namespace {

#define CRC(call)                     \
  do {                                \
    if (int rc_ = (call)) return rc_; \
  } while (0)

struct Shape { int nd; int64_t len[8]; };

// solve the symmetric system B x = rhs (DiisMixer.cxx:16-41 uses dsysv_); partial pivoting
bool solve_dense(std::vector<double> a, std::vector<double>& x, int n) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(a[r * n + c]) > std::fabs(a[piv * n + c])) piv = r;
    if (a[piv * n + c] == 0.0) return false;
    if (piv != c) {
      for (int k = 0; k < n; ++k) std::swap(a[c * n + k], a[piv * n + k]);
      std::swap(x[c], x[piv]);
    }
    for (int r = c + 1; r < n; ++r) {
      const double f = a[r * n + c] / a[c * n + c];
      for (int k = c; k < n; ++k) a[r * n + k] -= f * a[c * n + k];
      x[r] -= f * x[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    for (int k = r + 1; k < n; ++k) x[r] -= a[r * n + k] * x[k];
    x[r] /= a[r * n + r];
  }
  return true;
}

}  // namespace

struct CcsdHandle_ {
  tn_handle_t tn = nullptr;
  int o = 0, v = 0;
  int epsi = -1, epsa = -1;
  std::map<std::string, int> V;   // integral blocks
  int Vx = -1;                    // Vabij["baij"]
  int T1 = -1, T2 = -1;           // current amplitudes
  bool initial_doubles = false;
  // intermediates of the residuum
  int X = -1, Kac = -1, Mac = -1, Lac = -1, Kki = -1, Mki = -1, Lki = -1, Kck = -1, Zki = -1, Y = -1, Xakic = -1,
      Xakci = -1, Xklij = -1, Xabcd = -1, S = -1;

  int tensor(std::initializer_list<int64_t> lens, int* id) {
    std::vector<int64_t> l(lens);
    return tn_tensor(tn, (int)l.size(), l.data(), id);
  }
  int C(double alpha, int a, const char* ia, int b, const char* ib, double beta, int c, const char* ic) {
    return tn_contract(tn, alpha, a, ia, b, ib, beta, c, ic);
  }
  int A(double alpha, int a, const char* ia, double beta, int c, const char* ic) { return tn_add(tn, alpha, a, ia, beta, c, ic); }
};

namespace {

Shape block_shape(const CcsdHandle_* h, const std::string& name) {
  Shape s{4, {0}};
  for (int d = 0; d < 4; ++d) s.len[d] = name[d] == 'P' ? h->v : h->o;
  return s;
}

int ensure_block(ccsd_handle_t h, const std::string& name, int* id) {
  static const char* names[] = {"PPHH", "PHPH", "HHHH", "HHHP", "PPPH", "PPPP"};
  bool ok = false;
  for (const char* n : names) ok = ok || name == n;
  if (!ok) return pt::tn_record_error(TN_ERR_INVALID, ("unknown integral block " + name + " (PPHH, PHPH, HHHH, HHHP, PPPH, PPPP)").c_str());
  auto it = h->V.find(name);
  if (it == h->V.end()) {
    Shape s = block_shape(h, name);
    int t;
    CRC(tn_tensor(h->tn, 4, s.len, &t));
    it = h->V.emplace(name, t).first;
  }
  *id = it->second;
  return TN_OK;
}

int build_x(ccsd_handle_t h, int Tai, int Tabij) {   // Xabij["abij"] = Tabij["abij"] + Tai["ai"] Tai["bj"]   (:73-74)
  CRC(h->A(1.0, Tabij, "abij", 0.0, h->X, "abij"));
  return h->C(1.0, Tai, "ai", Tai, "bj", 1.0, h->X, "abij");
}

// getResiduum(i, amplitudes) into Rai, Rabij (:29-295)
int residuum(ccsd_handle_t h, int iteration, int Tai, int Tabij, int Rai, int Rabij) {
  for (const char* n : {"PPHH", "PHPH", "HHHH", "HHHP", "PPPH", "PPPP"})
    if (!h->V.count(n)) return pt::tn_record_error(TN_ERR_INVALID, (std::string("Missing argument: ") + n + "CoulombIntegrals").c_str());
  const int Vabij = h->V["PPHH"], Vaibj = h->V["PHPH"], Vijkl = h->V["HHHH"], Vijka = h->V["HHHP"], Vabci = h->V["PPPH"],
            Vabcd = h->V["PPPP"];
  if (iteration == 0 && !h->initial_doubles) {
    CRC(h->A(0.0, Tai, "ai", 0.0, Rai, "ai"));
    return h->A(1.0, Vabij, "abij", 0.0, Rabij, "abij");                       // :52-57: MP2 amplitudes
  }
  const int X = h->X, Y = h->Y;
  CRC(build_x(h, Tai, Tabij));
  // Kac (:169-173), with Tabij + Tai Tai = X
  CRC(h->C(-2.0, Vabij, "cdkl", X, "adkl", 0.0, h->Kac, "ac"));
  CRC(h->C(1.0, Vabij, "dckl", X, "adkl", 1.0, h->Kac, "ac"));
  // Lac - Kac (:177-178), Lac (:176)
  CRC(h->C(2.0, Vabci, "cdak", Tai, "dk", 0.0, h->Mac, "ac"));
  CRC(h->C(-1.0, Vabci, "dcak", Tai, "dk", 1.0, h->Mac, "ac"));
  CRC(h->A(1.0, h->Kac, "ac", 0.0, h->Lac, "ac"));
  CRC(h->A(1.0, h->Mac, "ac", 1.0, h->Lac, "ac"));
  // Kki (:181-184)
  CRC(h->C(2.0, Vabij, "cdkl", X, "cdil", 0.0, h->Kki, "ki"));
  CRC(h->C(-1.0, Vabij, "dckl", X, "cdil", 1.0, h->Kki, "ki"));
  // Lki - Kki (:188-189), Lki (:187)
  CRC(h->C(2.0, Vijka, "klic", Tai, "cl", 0.0, h->Mki, "ki"));
  CRC(h->C(-1.0, Vijka, "lkic", Tai, "cl", 1.0, h->Mki, "ki"));
  CRC(h->A(1.0, h->Kki, "ki", 0.0, h->Lki, "ki"));
  CRC(h->A(1.0, h->Mki, "ki", 1.0, h->Lki, "ki"));
  // :192-201
  CRC(h->C(1.0, h->Lac, "ac", Tabij, "cbij", 0.0, Rabij, "abij"));
  CRC(h->C(-1.0, h->Lki, "ki", Tabij, "abkj", 1.0, Rabij, "abij"));
  CRC(h->C(1.0, Vabci, "baci", Tai, "cj", 1.0, Rabij, "abij"));
  CRC(h->C(1.0, Vaibj, "bkci", Tai, "cj", 0.0, Y, "bkij"));                    // :198 = -(Vaibj Tai) Tai
  CRC(h->C(-1.0, Y, "bkij", Tai, "ak", 1.0, Rabij, "abij"));
  CRC(h->C(-1.0, Vijka, "jika", Tai, "bk", 1.0, Rabij, "abij"));
  CRC(h->C(1.0, Vabij, "acik", Tai, "cj", 0.0, Y, "aikj"));                    // :201
  CRC(h->C(-1.0, Y, "aikj", Tai, "bk", 1.0, Rabij, "abij"));
  // Xakic (:204-210)
  CRC(h->A(1.0, Vabij, "acik", 0.0, h->Xakic, "akic"));
  CRC(h->C(-1.0, Vijka, "lkic", Tai, "al", 1.0, h->Xakic, "akic"));
  CRC(h->C(1.0, Vabci, "acdk", Tai, "di", 1.0, h->Xakic, "akic"));
  CRC(h->C(-0.5, Vabij, "dclk", Tabij, "dail", 1.0, h->Xakic, "akic"));
  CRC(h->C(1.0, Vabij, "dclk", Tai, "di", 0.0, Y, "clki"));                    // :208
  CRC(h->C(-1.0, Y, "clki", Tai, "al", 1.0, h->Xakic, "akic"));
  CRC(h->C(1.0, Vabij, "dclk", Tabij, "adil", 1.0, h->Xakic, "akic"));
  CRC(h->C(-0.5, Vabij, "cdlk", Tabij, "adil", 1.0, h->Xakic, "akic"));
  // Xakci (:213-217)
  CRC(h->A(1.0, Vaibj, "akci", 0.0, h->Xakci, "akci"));
  CRC(h->C(-1.0, Vijka, "klic", Tai, "al", 1.0, h->Xakci, "akci"));
  CRC(h->C(1.0, Vabci, "adck", Tai, "di", 1.0, h->Xakci, "akci"));
  CRC(h->C(-0.5, Vabij, "cdlk", Tabij, "dail", 1.0, h->Xakci, "akci"));
  CRC(h->C(1.0, Vabij, "cdlk", Tai, "di", 0.0, Y, "clki"));                    // :217
  CRC(h->C(-1.0, Y, "clki", Tai, "al", 1.0, h->Xakci, "akci"));
  // :220-224
  CRC(h->C(2.0, h->Xakic, "akic", Tabij, "cbkj", 1.0, Rabij, "abij"));
  CRC(h->C(-1.0, h->Xakic, "akic", Tabij, "bckj", 1.0, Rabij, "abij"));
  CRC(h->C(-1.0, h->Xakci, "akci", Tabij, "cbkj", 1.0, Rabij, "abij"));
  CRC(h->C(-1.0, h->Xakci, "bkci", Tabij, "ackj", 1.0, Rabij, "abij"));
  // permutation operator (:228-229), bare integrals (:238)
  CRC(h->A(1.0, Rabij, "abij", 0.0, h->S, "abij"));
  CRC(h->A(1.0, h->S, "baji", 1.0, Rabij, "abij"));
  CRC(h->A(1.0, Vabij, "abij", 1.0, Rabij, "abij"));
  // Xklij (:241-245) and its contractions (:248-251)
  CRC(h->A(1.0, Vijkl, "klij", 0.0, h->Xklij, "klij"));
  CRC(h->C(1.0, Vijka, "klic", Tai, "cj", 1.0, h->Xklij, "klij"));
  CRC(h->C(1.0, Vijka, "lkjc", Tai, "ci", 1.0, h->Xklij, "klij"));
  CRC(h->C(1.0, Vabij, "cdkl", X, "cdij", 1.0, h->Xklij, "klij"));
  CRC(h->C(1.0, h->Xklij, "klij", X, "abkl", 1.0, Rabij, "abij"));
  // Xabcd (:254-256) and its contractions (:259-260)
  CRC(h->A(1.0, Vabcd, "abcd", 0.0, h->Xabcd, "abcd"));
  CRC(h->C(-1.0, Vabci, "cdak", Tai, "bk", 1.0, h->Xabcd, "abcd"));
  CRC(h->C(-1.0, Vabci, "dcbk", Tai, "ak", 1.0, h->Xabcd, "abcd"));
  CRC(h->C(1.0, h->Xabcd, "abcd", X, "cdij", 1.0, Rabij, "abij"));
  // T1 equations (:270-293)
  CRC(h->C(1.0, h->Kac, "ac", Tai, "ci", 0.0, Rai, "ai"));
  CRC(h->C(-1.0, h->Kki, "ki", Tai, "ak", 1.0, Rai, "ai"));
  CRC(h->C(2.0, Vabij, "cdkl", Tai, "dl", 0.0, h->Kck, "ck"));
  CRC(h->C(-1.0, Vabij, "cdlk", Tai, "dl", 1.0, h->Kck, "ck"));
  CRC(h->C(2.0, h->Kck, "ck", Tabij, "caki", 1.0, Rai, "ai"));
  CRC(h->C(-1.0, h->Kck, "ck", Tabij, "caik", 1.0, Rai, "ai"));
  CRC(h->C(1.0, h->Kck, "ck", Tai, "ci", 0.0, h->Zki, "ki"));                  // :280
  CRC(h->C(1.0, h->Zki, "ki", Tai, "ak", 1.0, Rai, "ai"));
  CRC(h->C(2.0, Vabij, "acik", Tai, "ck", 1.0, Rai, "ai"));
  CRC(h->C(-1.0, Vaibj, "akci", Tai, "ck", 1.0, Rai, "ai"));
  CRC(h->C(2.0, Vabci, "cdak", Tabij, "cdik", 1.0, Rai, "ai"));
  CRC(h->C(-1.0, Vabci, "dcak", Tabij, "cdik", 1.0, Rai, "ai"));
  CRC(h->C(1.0, h->Mac, "ac", Tai, "ci", 1.0, Rai, "ai"));                     // :286-287 = (Lac - Kac) Tai
  CRC(h->C(-2.0, Vijka, "klic", Tabij, "ackl", 1.0, Rai, "ai"));
  CRC(h->C(1.0, Vijka, "lkic", Tabij, "ackl", 1.0, Rai, "ai"));
  CRC(h->C(-1.0, h->Mki, "ki", Tai, "ak", 1.0, Rai, "ai"));                    // :290-291 = -(Lki - Kki) Tai
  return TN_OK;
}

// getEnergy (:160-178), spins = 2: direct 2 X.V, exchange -X.V["baij"]
int energy(ccsd_handle_t h, int Tai, int Tabij, double* e, double* dire, double* exce) {
  CRC(build_x(h, Tai, Tabij));
  double d = 0, x = 0;
  CRC(tn_dot(h->tn, h->X, h->V["PPHH"], &d));
  CRC(tn_dot(h->tn, h->X, h->Vx, &x));
  *dire = 2.0 * d;
  *exce = -x;
  *e = *dire + *exce;
  return TN_OK;
}

int alloc_pair(ccsd_handle_t h, int out[2]) {
  CRC(h->tensor({h->v, h->o}, &out[0]));
  return h->tensor({h->v, h->v, h->o, h->o}, &out[1]);
}
void free_pair(ccsd_handle_t h, int p[2]) {
  for (int q = 0; q < 2; ++q)
    if (p[q] >= 0) { tn_free(h->tn, p[q]); p[q] = -1; }
}
const char* IDX[2] = {"ai", "abij"};

}  // namespace

extern "C" {

void ccsd_default_options(CcsdOptions* opt) {
  if (!opt) return;
  memset(opt, 0, sizeof *opt);
  opt->mixer = CCSD_LINEAR_MIXER;       // ClusterSinglesDoublesAlgorithm.cxx:48
  opt->max_residua = 4;
  opt->mixing_ratio = 1.0;
  opt->max_iterations = 16;             // ClusterSinglesDoublesAlgorithm.hpp defaults
  opt->energy_convergence = 1e-6;
  opt->amplitudes_convergence = 1e-5;
  opt->level_shift = 0.0;
}

int ccsd_create(ccsd_handle_t* out, int o, int v, int device) try {
  if (!out || o < 1 || v < 1) return pt::tn_record_error(TN_ERR_INVALID, "ccsd_create: bad arguments");
  ccsd_handle_t h = new CcsdHandle_();
  h->o = o; h->v = v;
  int rc = tn_create(&h->tn, device);
  auto t = [&](std::initializer_list<int64_t> l, int* id) { if (!rc) rc = h->tensor(l, id); };
  t({o}, &h->epsi); t({v}, &h->epsa);
  t({v, o}, &h->T1); t({v, v, o, o}, &h->T2);
  t({v, v, o, o}, &h->Vx); t({v, v, o, o}, &h->X); t({v, v, o, o}, &h->S);
  t({v, v}, &h->Kac); t({v, v}, &h->Mac); t({v, v}, &h->Lac);
  t({o, o}, &h->Kki); t({o, o}, &h->Mki); t({o, o}, &h->Lki);
  t({v, o}, &h->Kck); t({o, o}, &h->Zki); t({v, o, o, o}, &h->Y);
  t({v, o, o, v}, &h->Xakic); t({v, o, v, o}, &h->Xakci); t({o, o, o, o}, &h->Xklij); t({v, v, v, v}, &h->Xabcd);
  if (rc) { ccsd_destroy(h); return rc; }
  *out = h;
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("ccsd_create");
}

int ccsd_destroy(ccsd_handle_t h) try {
  if (!h) return TN_OK;
  if (h->tn) tn_destroy(h->tn);   // frees every tensor of the engine
  delete h;
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("ccsd_destroy");
}

int ccsd_set_eigenenergies(ccsd_handle_t h, const double* epsi, const double* epsa) try {
  if (!h || !epsi || !epsa) return pt::tn_record_error(TN_ERR_INVALID, "ccsd_set_eigenenergies: bad arguments");
  CRC(tn_upload(h->tn, h->epsi, epsi));
  return tn_upload(h->tn, h->epsa, epsa);
} catch (...) {
  return pt::tn_on_exception("ccsd_set_eigenenergies");
}

int ccsd_set_integrals(ccsd_handle_t h, const char* name, const double* block) try {
  if (!h || !name || !block) return pt::tn_record_error(TN_ERR_INVALID, "ccsd_set_integrals: bad arguments");
  int id;
  CRC(ensure_block(h, name, &id));
  CRC(tn_upload(h->tn, id, block));
  if (!strcmp(name, "PPHH")) CRC(h->A(1.0, id, "baij", 0.0, h->Vx, "abij"));   // exchange operand of getEnergy
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("ccsd_set_integrals");
}

int ccsd_get_integrals(ccsd_handle_t h, const char* name, double* block) try {
  if (!h || !name || !block || !h->V.count(name)) return pt::tn_record_error(TN_ERR_INVALID, "ccsd_get_integrals: bad arguments");
  return tn_download(h->tn, h->V[name], block);
} catch (...) {
  return pt::tn_on_exception("ccsd_get_integrals");
}

int ccsd_set_vertex(ccsd_handle_t h, int nf, int np, const double* gre, const double* gim) try {
  if (!h || !gre || !gim || nf < 1 || np < h->o + h->v) return pt::tn_record_error(TN_ERR_INVALID, "ccsd_set_vertex: bad arguments");
  const int o = h->o, v = h->v, a0 = np - v;
  // the three vertex blocks the reference slices (CoulombIntegralsFromVertex.cxx:121-136), Re and Im apart
  // (fromComplexTensor); gathered on the host: O(NF Np^2), the blocks themselves are built on the device
  struct Part { int p0, np_, q0, nq; int id[2]; } parts[3] = {{0, o, 0, o, {-1, -1}}, {a0, v, 0, o, {-1, -1}}, {a0, v, a0, v, {-1, -1}}};
  for (auto& pt : parts)
    for (int c = 0; c < 2; ++c) {
      const double* g = c == 0 ? gre : gim;
      std::vector<double> buf((size_t)nf * pt.np_ * pt.nq);
      for (int q = 0; q < pt.nq; ++q)
        for (int p = 0; p < pt.np_; ++p)
          memcpy(&buf[(size_t)nf * (p + (size_t)pt.np_ * q)], g + (size_t)nf * ((pt.p0 + p) + (size_t)np * (pt.q0 + q)),
                 sizeof(double) * nf);
      int rc = h->tensor({nf, pt.np_, pt.nq}, &pt.id[c]);
      if (!rc) rc = tn_upload(h->tn, pt.id[c], buf.data());
      if (rc) return rc;
    }
  enum { IJ = 0, AI = 1, AB = 2 };
  struct Def { const char* name; int p1; const char* i1; int p2; const char* i2; const char* out; } defs[] = {
      {"PPHH", AI, "Gai", AI, "Gbj", "abij"},    // :402-403
      {"HHHH", IJ, "Gik", IJ, "Gjl", "ijkl"},    // :409-410
      {"HHHP", IJ, "Gik", AI, "Gaj", "ijka"},    // :416-417
      {"PPPP", AB, "Gac", AB, "Gbd", "abcd"},    // :423-424
      {"PPPH", AB, "Gac", AI, "Gbi", "abci"},    // :430-431
      {"PHPH", AB, "Gab", IJ, "Gij", "aibj"}};   // :395-396
  for (auto& d : defs) {
    int id;
    CRC(ensure_block(h, d.name, &id));
    for (int c = 0; c < 2; ++c)   // Re.Re, then += Im.Im
      CRC(h->C(1.0, parts[d.p1].id[c], d.i1, parts[d.p2].id[c], d.i2, c ? 1.0 : 0.0, id, d.out));
  }
  CRC(h->A(1.0, h->V["PPHH"], "baij", 0.0, h->Vx, "abij"));
  for (auto& pt : parts)
    for (int c = 0; c < 2; ++c) tn_free(h->tn, pt.id[c]);
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("ccsd_set_vertex");
}

int ccsd_set_amplitudes(ccsd_handle_t h, const double* t1, const double* t2) try {
  if (!h) return pt::tn_record_error(TN_ERR_INVALID, "ccsd_set_amplitudes: bad arguments");
  if (t1) CRC(tn_upload(h->tn, h->T1, t1));
  if (t2) { CRC(tn_upload(h->tn, h->T2, t2)); h->initial_doubles = true; }
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("ccsd_set_amplitudes");
}

int ccsd_get_amplitudes(ccsd_handle_t h, double* t1, double* t2) try {
  if (!h) return pt::tn_record_error(TN_ERR_INVALID, "ccsd_get_amplitudes: bad arguments");
  if (t1) CRC(tn_download(h->tn, h->T1, t1));
  if (t2) CRC(tn_download(h->tn, h->T2, t2));
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("ccsd_get_amplitudes");
}

int ccsd_residuum(ccsd_handle_t h, int iteration, double* r1, double* r2) try {
  if (!h || !r1 || !r2) return pt::tn_record_error(TN_ERR_INVALID, "ccsd_residuum: bad arguments");
  int R[2] = {-1, -1};
  int rc = alloc_pair(h, R);
  if (!rc) rc = residuum(h, iteration, h->T1, h->T2, R[0], R[1]);
  if (!rc) rc = tn_download(h->tn, R[0], r1);
  if (!rc) rc = tn_download(h->tn, R[1], r2);
  free_pair(h, R);
  return rc;
} catch (...) {
  return pt::tn_on_exception("ccsd_residuum");
}

int ccsd_solve(ccsd_handle_t h, const CcsdOptions* opt_in, CcsdResult* res) try {
  if (!h || !res) return pt::tn_record_error(TN_ERR_INVALID, "ccsd_solve: bad arguments");
  CcsdOptions opt;
  if (opt_in) opt = *opt_in; else ccsd_default_options(&opt);
  if (opt.mixer != CCSD_LINEAR_MIXER && opt.mixer != CCSD_DIIS_MIXER) return pt::tn_record_error(TN_ERR_INVALID, "ccsd_solve: Mixer not implemented");   // (:50-54)
  memset(res, 0, sizeof *res);
  const int N = opt.mixer == CCSD_DIIS_MIXER ? std::max(1, opt.max_residua) : 0;
  // DIIS ring (DiisMixer.cxx:55-100): amplitudes, residua, overlap matrix bordered by -1
  std::vector<int> ringA(2 * N, -1), ringR(2 * N, -1);
  std::vector<double> B((N + 1) * (N + 1), 0.0);
  for (int q = 1; q <= N; ++q) B[q] = B[q * (N + 1)] = -1.0;
  int next_index = 0, count = 0;
  double e = 0, dire = 0, exce = 0, prev = 0;
  int it = 0, rc = TN_OK;
  bool converged = false;
  for (; it < opt.max_iterations && !rc; ++it) {
    int R[2] = {-1, -1}, D[2] = {-1, -1};
    rc = alloc_pair(h, R);
    if (!rc) rc = alloc_pair(h, D);
    if (!rc) rc = residuum(h, it, h->T1, h->T2, R[0], R[1]);
    // estimateAmplitudesFromResiduum (:302-331), then amplitudesChange = estimate - amplitudes (:96-97)
    const int T[2] = {h->T1, h->T2};
    double dd = 0, tt = 0;
    for (int q = 0; q < 2 && !rc; ++q) {
      rc = tn_excitation_divide(h->tn, R[q], T[q], h->epsi, h->epsa, opt.level_shift);
      if (!rc) rc = h->A(1.0, R[q], IDX[q], 0.0, D[q], IDX[q]);
      if (!rc) rc = h->A(-1.0, T[q], IDX[q], 1.0, D[q], IDX[q]);
      double x = 0;
      if (!rc) rc = tn_dot(h->tn, D[q], D[q], &x);
      dd += x;
    }
    if (rc) { free_pair(h, R); free_pair(h, D); break; }
    if (opt.mixer == CCSD_LINEAR_MIXER) {
      // LinearMixer::append (:31-45): next = ratio * estimate + (1 - ratio) * last  (last = the current amplitudes
      // from the second iteration on; the first estimate is taken as it is)
      for (int q = 0; q < 2 && !rc; ++q) {
        if (it > 0) rc = h->A(1.0 - opt.mixing_ratio, T[q], IDX[q], opt.mixing_ratio, R[q], IDX[q]);
        if (!rc) rc = h->A(1.0, R[q], IDX[q], 0.0, T[q], IDX[q]);
      }
      free_pair(h, R);
      free_pair(h, D);
    } else {
      // DiisMixer::append (:103-181)
      for (int q = 0; q < 2; ++q) {
        if (ringA[2 * next_index + q] >= 0) tn_free(h->tn, ringA[2 * next_index + q]);
        if (ringR[2 * next_index + q] >= 0) tn_free(h->tn, ringR[2 * next_index + q]);
        ringA[2 * next_index + q] = R[q];
        ringR[2 * next_index + q] = D[q];
      }
      for (int i = 0; i < N && !rc; ++i) {
        if (ringR[2 * i] < 0) continue;
        double ov = 0;
        for (int q = 0; q < 2 && !rc; ++q) {
          double x = 0;
          rc = tn_dot(h->tn, ringR[2 * i + q], D[q], &x);
          ov += x;
        }
        B[(next_index + 1) * (N + 1) + (i + 1)] = B[(i + 1) * (N + 1) + (next_index + 1)] = 2.0 * ov;   // :126
      }
      if (rc) break;
      if (count < N) ++count;
      const int dim = count + 1;
      std::vector<double> a(dim * dim), col(dim, 0.0);
      for (int r = 0; r < dim; ++r)
        for (int c = 0; c < dim; ++c) a[r * dim + c] = B[r * (N + 1) + c];
      col[0] = -1.0;
      if (!solve_dense(a, col, dim)) { rc = pt::tn_record_error(TN_ERR_INVALID, "DiisMixer: singular B matrix"); break; }   // "problem diagonalization" (:37-39)
      for (int q = 0; q < 2 && !rc; ++q) {                             // next = sum_j w_j amplitudes_j (:161-172)
        rc = h->A(0.0, T[q], IDX[q], 0.0, T[q], IDX[q]);
        for (int j = 0; j < count && !rc; ++j) {
          const int i = (next_index + N - j) % N;
          rc = h->A(col[i + 1], ringA[2 * i + q], IDX[q], 1.0, T[q], IDX[q]);
        }
      }
      next_index = (next_index + 1) % N;
    }
    if (rc) break;
    rc = energy(h, h->T1, h->T2, &e, &dire, &exce);
    for (int q = 0; q < 2 && !rc; ++q) {
      double x = 0;
      rc = tn_dot(h->tn, T[q], T[q], &x);
      tt += x;
    }
    if (rc) break;
    if (std::fabs((e - prev) / e) < std::fabs(opt.energy_convergence) &&
        std::fabs(dd / tt) < std::fabs(opt.amplitudes_convergence * opt.amplitudes_convergence)) {   // :103-107
      converged = true;
      ++it;
      break;
    }
    prev = e;
  }
  for (int id : ringA) if (id >= 0) tn_free(h->tn, id);
  for (int id : ringR) if (id >= 0) tn_free(h->tn, id);
  if (rc) return rc;
  if (opt.max_iterations == 0) CRC(energy(h, h->T1, h->T2, &e, &dire, &exce));   // :116-119
  res->energy = e; res->direct = dire; res->exchange = exce;
  res->iterations = it;
  res->converged = converged ? 1 : 0;
  double bytes = 0;
  tn_get_stats(h->tn, &res->flops, &bytes, &res->kernel_launches);
  return TN_OK;
} catch (...) {
  return pt::tn_on_exception("ccsd_solve");
}

}  // extern "C"
