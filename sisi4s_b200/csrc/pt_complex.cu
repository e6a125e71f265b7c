// pt_complex.cu -- complex closed-shell perturbative triples through the real step's machinery
// (include/sisi4s_pt.h: pt_complex_triples).  Host logic only.
//
// Reference: CcsdPerturbativeTriplesComplex::Calculator<complex>::calculate
// (src/algorithms/CcsdPerturbativeTriplesComplex.cxx:166-271): particle term from
// conj(GammaFab)["Fdb"] GammaFai["Fck"] (:341-348), hole term from PHHHCoulombIntegrals["clkj"] (:135-140),
// DV <- conj(DV / Delta) (:224-230), real part of the energy (:76).
//
// With W = W_r + i W_i the triples block of one hole permutation, S the singles term, and B(W; S) the bilinear
// form the fused epilogue evaluates, Re E(T) = B(W_r; S_r) + B(W_i; S_i): conj() flips the
// sign of one factor's imaginary part and the permutation / spin-factor algebra is real.  Each pass is a REAL
// (T) evaluation whose contractions have real and imaginary parts stacked along the contracted index,
//     W_r = [T_r | -T_i] . [V_r ; V_i]      W_i = [T_i | T_r] . [V_r ; V_i]      (particle: 2v, hole: 2o)
//     S_r = 1/2 (t_r (x) P_r - t_i (x) P_i)  S_i = 1/2 (t_r (x) P_i + t_i (x) P_r)  (two singles terms)
// so the big stacked integrals are packed once and shared by both passes; 4x the real step's work.
#include <cstring>
#include <vector>

#include "../../include/sisi4s_pt.h"
#include "../../include/sisi4s_tn.h"

namespace pt {
int record_error(int code, const char* message);
int on_exception(const char* where);   // pt_api.cu: what pt_last_error() returns
}

namespace {

int tn_failed() { return pt::record_error(PT_ERR_CUDA, tn_last_error()); }

struct Guard {
  pt_handle_t pt = nullptr;
  tn_handle_t tn = nullptr;
  ~Guard() {
    if (pt) pt_destroy(pt);
    if (tn) tn_destroy(tn);
  }
};

#define PRC(call)                     \
  do {                                \
    if (int rc_ = (call)) return rc_; \
  } while (0)

// out[first ; second] stacked along dimension `dim` of a column-major tensor with extents lens[0..nd)
void stack(const double* first, double s1, const double* second, double s2, const int64_t* lens, int nd, int dim,
           std::vector<double>& out) {
  int64_t inner = 1, outer = 1;
  for (int d = 0; d < dim; ++d) inner *= lens[d];
  for (int d = dim + 1; d < nd; ++d) outer *= lens[d];
  const int64_t n = lens[dim], chunk = inner * n;
  out.resize((size_t)(2 * chunk * outer));
  for (int64_t q = 0; q < outer; ++q) {
    double* o = out.data() + 2 * chunk * q;
    const double *a = first + chunk * q, *b = second + chunk * q;
    for (int64_t x = 0; x < chunk; ++x) o[x] = s1 * a[x];
    for (int64_t x = 0; x < chunk; ++x) o[chunk + x] = s2 * b[x];
  }
}

}  // namespace

extern "C" int pt_complex_triples(int o, int v, int device, const double* epsi, const double* epsa, const double* t1_re,
                                  const double* t1_im, const double* t2_re, const double* t2_im, const double* pphh_re,
                                  const double* pphh_im, const double* phhh_re, const double* phhh_im, int nf, int np,
                                  const double* gamma_re, const double* gamma_im, double* e_triples,
                                  double* e_per_triple) try {
  if (o < 1 || v < 1 || nf < 1 || np < o + v || !epsi || !epsa || !t1_re || !t1_im || !t2_re || !t2_im || !pphh_re ||
      !pphh_im || !phhh_re || !phhh_im || !gamma_re || !gamma_im || !e_triples)
    return pt::record_error(PT_ERR_INVALID, "pt_complex_triples: bad arguments (need o, v, nf >= 1, np >= o + v, non-null arrays)");
  Guard g;
  const int a0 = np - v;
  const size_t vv = (size_t)v * v, n4 = vv * v * o;
  // ---- complex PPPH block V[b,c,d,k] = sum_F conj(G[F,d,b]) G[F,c,k] on the device (:341-348)
  std::vector<double> vr(n4), vi(n4);
  {
    if (tn_create(&g.tn, device)) return tn_failed();
    auto block = [&](const double* src, int p0, int npart, int q0, int nq, int* id) -> int {   // G[:, p0:p0+npart, q0:q0+nq]
      std::vector<double> buf((size_t)nf * npart * nq);
      for (int q = 0; q < nq; ++q)
        for (int p = 0; p < npart; ++p)
          memcpy(&buf[(size_t)nf * (p + (size_t)npart * q)], src + (size_t)nf * ((p0 + p) + (size_t)np * (q0 + q)), sizeof(double) * nf);
      const int64_t lens[3] = {nf, npart, nq};
      if (tn_tensor(g.tn, 3, lens, id)) return tn_failed();
      return tn_upload(g.tn, *id, buf.data()) ? tn_failed() : PT_OK;
    };
    int abr, abi, air, aii, tr, ti;
    PRC(block(gamma_re, a0, v, a0, v, &abr));
    PRC(block(gamma_im, a0, v, a0, v, &abi));
    PRC(block(gamma_re, a0, v, 0, o, &air));
    PRC(block(gamma_im, a0, v, 0, o, &aii));
    const int64_t l4[4] = {v, v, v, o};
    if (tn_tensor(g.tn, 4, l4, &tr) || tn_tensor(g.tn, 4, l4, &ti)) return tn_failed();
    int rc = tn_contract(g.tn, 1.0, abr, "Fdb", air, "Fck", 0.0, tr, "bcdk");      // Re: ab_r ai_r + ab_i ai_i
    if (!rc) rc = tn_contract(g.tn, 1.0, abi, "Fdb", aii, "Fck", 1.0, tr, "bcdk");
    if (!rc) rc = tn_contract(g.tn, 1.0, abr, "Fdb", aii, "Fck", 0.0, ti, "bcdk"); // Im: ab_r ai_i - ab_i ai_r
    if (!rc) rc = tn_contract(g.tn, -1.0, abi, "Fdb", air, "Fck", 1.0, ti, "bcdk");
    if (!rc) rc = tn_download(g.tn, tr, vr.data());
    if (!rc) rc = tn_download(g.tn, ti, vi.data());
    if (rc) return tn_failed();
    tn_destroy(g.tn);
    g.tn = nullptr;
  }
  // ---- the stacked, shared integrals
  std::vector<double> ppph_e, hhhp_e((size_t)o * o * 2 * o * v);
  {
    const int64_t l4[4] = {v, v, v, o};
    stack(vr.data(), 1.0, vi.data(), 1.0, l4, 4, 2, ppph_e);                         // [v,v,2v,o]: [V_r ; V_i] along d
    std::vector<double>().swap(vr);
    std::vector<double>().swap(vi);
    // U[y,z,l,c] = Vphhh[c,l,z,y]; [U_r ; U_i] along l: [o,o,2o,v]
    for (int c = 0; c < v; ++c)
      for (int l = 0; l < o; ++l)
        for (int z = 0; z < o; ++z)
          for (int y = 0; y < o; ++y) {
            const size_t src = c + (size_t)v * (l + (size_t)o * (z + (size_t)o * y));
            const size_t dst = y + (size_t)o * (z + (size_t)o * ((size_t)2 * o * c));
            hhhp_e[dst + (size_t)o * o * l] = phhh_re[src];
            hhhp_e[dst + (size_t)o * o * (o + l)] = phhh_im[src];
          }
  }
  PRC(pt_create_ex(&g.pt, o, 2 * o, v, device));
  PRC(pt_set_option(g.pt, "particle_contraction", 2 * v));
  PRC(pt_set_eigenenergies(g.pt, epsi, epsa));
  PRC(pt_set_hhhp(g.pt, hhhp_e.data()));
  PRC(pt_set_ppph_slabs(g.pt, 0, o, ppph_e.data()));
  std::vector<double>().swap(ppph_e);
  const int64_t ntr = pt_num_triples(o);
  std::vector<double> per((size_t)ntr, 0.0), per_pass((size_t)ntr), t2d, t2h, neg_t1i((size_t)v * o);
  for (size_t q = 0; q < neg_t1i.size(); ++q) neg_t1i[q] = -t1_im[q];
  double total = 0.0;
  const int64_t l4[4] = {v, v, o, o};
  for (int pass = 0; pass < 2; ++pass) {
    // pass 0: W_r, S_r from [T_r | -T_i];  pass 1: W_i, S_i from [T_i | T_r]
    const double *ta = pass == 0 ? t2_re : t2_im, *tb = pass == 0 ? t2_im : t2_re;
    const double sb = pass == 0 ? -1.0 : 1.0;
    stack(ta, 1.0, tb, sb, l4, 4, 1, t2d);   // particle term: [v,2v,o,o]
    stack(ta, 1.0, tb, sb, l4, 4, 3, t2h);   // hole term:     [v,v,o,2o]
    PRC(pt_set_doubles(g.pt, t2d.data()));
    PRC(pt_set_doubles_hole(g.pt, t2h.data()));
    if (pass == 0) {   // S_r = 1/2 (t_r P_r - t_i P_i)
      PRC(pt_set_singles(g.pt, t1_re));
      PRC(pt_set_pphh(g.pt, pphh_re));
      PRC(pt_set_singles_pair(g.pt, neg_t1i.data(), pphh_im));
    } else {           // S_i = 1/2 (t_r P_i + t_i P_r)
      PRC(pt_set_singles(g.pt, t1_re));
      PRC(pt_set_pphh(g.pt, pphh_im));
      PRC(pt_set_singles_pair(g.pt, t1_im, pphh_re));
    }
    double e = 0.0;
    PRC(pt_run(g.pt, 0, ntr, &e, per_pass.data()));
    total += e;
    for (int64_t t = 0; t < ntr; ++t) per[(size_t)t] += per_pass[(size_t)t];
  }
  *e_triples = total;
  if (e_per_triple) memcpy(e_per_triple, per.data(), sizeof(double) * (size_t)ntr);
  return PT_OK;
} catch (...) {
  return pt::on_exception("pt_complex_triples");
}
