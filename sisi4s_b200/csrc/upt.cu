// upt.cu -- spin-orbital (unrestricted) perturbative triples on the device tensor engine
// (include/sisi4s_pt.h: pt_spin_orbital_triples).  Host logic only.
//
// Reference: UPerturbativeTriples::run (src/algorithms/UPerturbativeTriples.cxx:19-305): the full-tensor form
// on antisymmetrised integrals; every CTF statement below is one tn_contract / tn_add with the same index
// strings.  Like the reference it holds three v^3 o^3 tensors, i.e. it is meant for small systems.
#include <vector>

#include "../../include/sisi4s_pt.h"
#include "../../include/sisi4s_tn.h"

namespace pt {
int record_error(int code, const char* message);
int on_exception(const char* where);   // pt_api.cu
}

extern "C" int pt_spin_orbital_triples(int o, int v, int device, const double* epsi, const double* epsa, const double* tai,
                                       const double* tabij, const double* vabij, const double* vijka,
                                       const double* vabci, double* e_triples) try {
  if (o < 1 || v < 1 || !epsi || !epsa || !tai || !tabij || !vabij || !vijka || !vabci || !e_triples)
    return pt::record_error(PT_ERR_INVALID, "pt_spin_orbital_triples: bad arguments");
  tn_handle_t tn = nullptr;
  if (tn_create(&tn, device)) return pt::record_error(PT_ERR_CUDA, tn_last_error());
  struct Guard { tn_handle_t h; ~Guard() { tn_destroy(h); } } guard{tn};
  int rc = 0;
  auto tensor = [&](std::initializer_list<int64_t> lens, const double* data) -> int {
    std::vector<int64_t> l(lens);
    int id = -1;
    if (!rc) rc = tn_tensor(tn, (int)l.size(), l.data(), &id);
    if (!rc && data) rc = tn_upload(tn, id, data);
    return id;
  };
  const int ei = tensor({o}, epsi), ea = tensor({v}, epsa), t1 = tensor({v, o}, tai), t2 = tensor({v, v, o, o}, tabij);
  const int pphh = tensor({v, v, o, o}, vabij), hhhp = tensor({o, o, o, v}, vijka), ppph = tensor({v, v, v, o}, vabci);
  const int T = tensor({v, v, v, o, o, o}, nullptr), DV = tensor({v, v, v, o, o, o}, nullptr), SV = tensor({v, v, v, o, o, o}, nullptr);
  if (rc) return pt::record_error(PT_ERR_CUDA, tn_last_error());
  struct Term { double sign; const char* idx; };
  auto accumulate = [&](int src, const Term (&terms)[9], const char* into) {
    for (const Term& t : terms)
      if (!rc) rc = tn_add(tn, t.sign, src, t.idx, 1.0, T, into);
  };
  // VABCI part (:112-124)
  rc = tn_contract(tn, 1.0, t2, "adij", ppph, "bcdk", 0.0, DV, "abcijk");
  const Term vabci_terms[9] = {{+1, "defjki"}, {-1, "edfjki"}, {-1, "fedjki"}, {+1, "defkij"}, {-1, "edfkij"},
                               {-1, "fedkij"}, {+1, "defijk"}, {-1, "edfijk"}, {-1, "fedijk"}};
  accumulate(DV, vabci_terms, "defjki");
  // VIJKA part (:127-137)
  if (!rc) rc = tn_contract(tn, 1.0, t2, "deok", hhhp, "ijof", 0.0, DV, "defkij");
  const Term vijka_terms[9] = {{+1, "defkij"}, {-1, "dfekij"}, {-1, "fedkij"}, {-1, "defjik"}, {+1, "dfejik"},
                               {+1, "fedjik"}, {-1, "defikj"}, {+1, "dfeikj"}, {+1, "fedikj"}};
  accumulate(DV, vijka_terms, "defjki");
  if (!rc) rc = tn_add(tn, 1.0, T, "abcijk", 0.0, DV, "abcijk");                  // :140 the antisymmetrised doubles part
  // singles part (:143-153)
  if (!rc) rc = tn_contract(tn, 1.0, t1, "dk", pphh, "efij", 0.0, SV, "defkij");
  const Term singles_terms[9] = {{+1, "defkij"}, {-1, "edfkij"}, {-1, "fedkij"}, {-1, "defjik"}, {+1, "edfjik"},
                                 {+1, "fedjik"}, {-1, "defikj"}, {+1, "edfikj"}, {+1, "fedikj"}};
  accumulate(SV, singles_terms, "defkij");
  // T / (eps_i + eps_j + eps_k - eps_a - eps_b - eps_c) (:270-283), energy (1/36) DV . T (:286)
  if (!rc) rc = tn_excitation_divide(tn, T, T, ei, ea, 0.0);
  double dot = 0.0;
  if (!rc) rc = tn_dot(tn, DV, T, &dot);
  if (rc) return pt::record_error(PT_ERR_CUDA, tn_last_error());
  *e_triples = dot / 36.0;
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_spin_orbital_triples");
}
