// TEST HARNESS ONLY - not part of the product.  Compiles the device headers
// (dvt_circuits_b200/csrc/*.cuh) for the host with plain g++ so that `-m "not gpu"` tests can check
// the exact per-thread logic of the CUDA kernels (formulas, codecs, Horner, window tables)
// against the oracle on a machine without a GPU.  The only thing not covered here is the PTX
// carry-chain Montgomery product, which the `-m gpu` tests pin on a real B200.
// Nothing under dvt_circuits_b200/ links or loads this file.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../dvt_circuits_b200/csrc/vm.cuh"
#include "../../dvt_circuits_b200/csrc/fdiff.cuh"

using namespace dkgv;

// host emulation of the fixed-base table: 16-bit windows (50 MB), one lazily allocated copy, only the entries a scalar touches are
// computed (gtab_entry, the slow definition the builder of dkgv.cu is checked against)
static const uint32_t HE_GTAB_BITS = 16;
static GTab he_gtab_for(const uint32_t* s_raw) {
  static std::vector<uint32_t> gtab;
  if (gtab.empty()) gtab.assign(gtab_words(HE_GTAB_BITS), 0);
  GTab g{gtab.data(), HE_GTAB_BITS, gtab_windows(HE_GTAB_BITS)};
  uint32_t u[8];
  gtab_scalar(u, s_raw);
  for (uint32_t w = 0; w < g.windows; w++) {
    bool ng;
    uint32_t idx = gtab_index(g.bits, g.windows, u, w, &ng);
    if (gtab[(size_t)idx * 24] | gtab[(size_t)idx * 24 + 1]) continue;
    G1Aff e = gtab_entry(g.bits, idx);
    for (int i = 0; i < 12; i++) {
      gtab[(size_t)idx * 24 + i] = e.x.l[i];
      gtab[(size_t)idx * 24 + 12 + i] = e.y.l[i];
    }
  }
  return g;
}

extern "C" {

// canonical 48-byte big-endian a, b -> canonical a*b mod p, a+b, a-b
void he_fp_ops(const uint8_t* a48, const uint8_t* b48, uint8_t* mul48, uint8_t* add48, uint8_t* sub48, uint8_t* inv48) {
  Fp a, b;
  fp_raw_from_be48(a.l, a48);
  fp_raw_from_be48(b.l, b48);
  a = to_mont(a);
  b = to_mont(b);
  fp_raw_to_be48(mul48, from_mont(mul(a, b)).l);
  fp_raw_to_be48(add48, from_mont(add(a, b)).l);
  fp_raw_to_be48(sub48, from_mont(sub(a, b)).l);
  fp_raw_to_be48(inv48, from_mont(fp_inv(a)).l);
}
// canonical a -> a^-1 by the binary extended Euclid routine (field.cuh fp_inv_bgcd)
void he_fp_inv_bgcd(const uint8_t* a48, uint8_t* inv48) {
  Fp a;
  fp_raw_from_be48(a.l, a48);
  fp_raw_to_be48(inv48, from_mont(fp_inv_bgcd(to_mont(a))).l);
}

void he_fr_mul(const uint8_t* a32, const uint8_t* b32, uint8_t* out32) {
  Fr a, b;
  fr_raw_from_be32(a.l, a32);
  fr_raw_from_be32(b.l, b32);
  Fr r = from_mont(mul(to_mont(a), to_mont(b)));
  for (int i = 0; i < 8; i++) {
    uint8_t* q = out32 + 28 - 4 * i;
    q[0] = r.l[i] >> 24; q[1] = r.l[i] >> 16; q[2] = r.l[i] >> 8; q[3] = r.l[i];
  }
}

uint32_t he_g1_decompress(const uint8_t* in48, uint8_t* recompressed48) {
  G1Aff a;
  uint32_t st = g1_decompress(in48, &a, true);
  if (st == G1_DEC_OK) g1_compress(a, recompressed48);
  return st;
}

// out = compress(P + Q), compress(2P), compress([k]P) through the projective formulas
uint32_t he_g1_ops(const uint8_t* p48, const uint8_t* q48, uint32_t k, uint8_t* add48, uint8_t* madd48, uint8_t* dbl48, uint8_t* mul48) {
  G1Aff p, q;
  if (g1_decompress(p48, &p, false) || g1_decompress(q48, &q, false)) return 1;
  G1Proj pp = g1_from_affine(p), qq = g1_from_affine(q);
  g1_compress(g1_to_affine(g1_add(pp, qq)), add48);
  g1_compress(g1_to_affine(g1_add_mixed(pp, q)), madd48);
  g1_compress(g1_to_affine(g1_dbl(pp)), dbl48);
  g1_compress(g1_to_affine(g1_mul_small(pp, k)), mul48);
  return g1_eq(g1_add(pp, qq), g1_add_mixed(pp, q)) ? 0 : 2;
}

// One emulated "thread" of k_share_verify / k_feldman_eval for dealer 0 of a 1-dealer session.
// vv: [t][48]; returns status; eval48 = compress(evaluate_polynomial), pk48 = compress(G*s)
uint32_t he_share_check(const uint8_t* vv, uint32_t t, uint32_t id, const uint8_t* secret32, uint8_t* eval48, uint8_t* pk48) {
  const uint32_t n_pad = 32;
  uint32_t tt = t ? t : 1;
  std::vector<uint32_t> limbs((size_t)tt * 24 * n_pad, 0);
  std::vector<uint8_t> inf((size_t)tt * n_pad, 1);
  bool bad = false;
  for (uint32_t k = 0; k < t; k++) {
    G1Aff a;
    if (g1_decompress(vv + (size_t)k * 48, &a, true) != G1_DEC_OK) bad = true;
    vv_store(limbs.data(), inf.data(), n_pad, k, 0, a);
  }
  VVView view{limbs.data(), inf.data(), n_pad};
  // the table entries this scalar touches, computed by the same routine as k_build_gtab
  uint32_t s[8];
  fr_raw_from_be32(s, secret32);
  GTab gtab = he_gtab_for(s);
  g1_compress(g1_to_affine(feldman_eval(view, t, 0, id)), eval48);
  g1_compress(g1_to_affine(fixed_base_mul(gtab, s)), pk48);
  uint32_t st = share_check(view, t, 0, id, secret32, gtab, bad);
  // the operand-file (vm.cuh) formulation used by the hot kernel must agree
  const uint32_t NT = 4, me = 2;  // pretend to be thread 2 of a 4-thread block
  std::vector<U4> file((size_t)VM_SLOTS * 3 * NT);
  OpFile f{file.data() + me, NT};
  uint32_t st_vm = vm_share_check(f, view, t, 0, id, secret32, gtab, bad);
  if (st_vm != st) return 0x100 | st_vm;
  vm_feldman_eval(f, view, t, 0, id);
  uint8_t ev2[48];
  g1_compress(g1_to_affine(vm_get_point(f, AX)), ev2);
  if (memcmp(ev2, eval48, 48)) return 0x200;
  vm_fixed_base_mul(f, gtab, s);
  g1_compress(g1_to_affine(vm_get_point(f, BX)), ev2);
  if (memcmp(ev2, pk48, 48)) return 0x300;
  return st;
}

}

// ---- finite-difference share path (fdiff.cuh): the exact sequence of k_fd_seed / k_fd_init /
// k_fd_ext / k_fd_digits / k_fd_combine items share_fd.cu launches, for one dealer sitting in lane
// `d` of a 32-dealer plane.
extern "C" {
// plan6 = use, parts, h, lo, hi, steps.  m_force 0 = planner's choice; lo_force != 0x7fffffff overrides
// the planned window start.  out48[j] = compress(f(j + 1)), j = 0..n_r-1.  Returns 0, or < 0 when no
// plan exists for the shape.
int he_fd_row(const uint8_t* vv, uint32_t t, uint32_t n_r, uint32_t m_force, int32_t lo_force, int32_t* plan6, uint8_t* out48) {
  const uint32_t n_pad = 32, d = 5;
  FdPlan plan = fd_make_plan(t, n_r, m_force);
  plan6[0] = plan.use;
  plan6[1] = (int32_t)plan.m;
  plan6[2] = (int32_t)plan.h;
  plan6[3] = plan.lo;
  plan6[4] = plan.hi;
  plan6[5] = (int32_t)plan.steps;
  if (plan.cost_fd == ~0ull) return -1;
  const uint32_t m = plan.m, h = plan.h;
  if (lo_force != 0x7fffffff) {
    plan.lo = lo_force;
    plan.hi = lo_force + (int32_t)h - 1;
    if (plan.lo > 1 || plan.hi < 1 || (uint32_t)plan.hi >= n_r) return -2;
    plan.steps = n_r - (uint32_t)plan.hi;
  }
  std::vector<uint32_t> limbs((size_t)t * 24 * n_pad, 0);
  std::vector<uint8_t> inf((size_t)t * n_pad, 1);
  for (uint32_t k = 0; k < t; k++) {
    G1Aff a;
    if (g1_decompress(vv + (size_t)k * 48, &a, true) != G1_DEC_OK) return -3;
    vv_store(limbs.data(), inf.data(), n_pad, k, d, a);
  }
  VVView view{limbs.data(), inf.data(), n_pad};
  const uint32_t NT = 4, me = 1;
  std::vector<U4> file((size_t)VM_SLOTS * 3 * NT);
  OpFile f{file.data() + me, NT};
  const uint32_t n_padv = n_pad * m;
  const size_t ent = (size_t)36 * n_padv;
  const size_t n_evals = (size_t)((int64_t)n_r - plan.lo + 1);
  std::vector<uint32_t> evals(n_evals * ent, 0xdeadbeef), p0((size_t)h * ent), p1((size_t)h * ent), da((size_t)h * ent),
      db((size_t)h * ent);
  for (uint32_t part = 0; part < m; part++)
    for (int32_t x = plan.lo; x <= plan.hi; x++) {
      fd_seed_eval(f, view, t, d, x, part * h, h);
      fd_store(f, AX, fd_entry(evals.data(), n_padv, (size_t)(x - plan.lo), part * n_pad + d), n_padv);
    }
  const size_t e_hi = (size_t)(plan.hi - plan.lo);
  memcpy(da.data(), evals.data() + e_hi * ent, ent * 4);
  memcpy(db.data(), evals.data() + e_hi * ent, ent * 4);
  uint32_t* pp[2] = {p0.data(), p1.data()};
  uint32_t* dd[2] = {da.data(), db.data()};
  for (uint32_t part = 0; part < m; part++) {  // a virtual dealer = one column of the planes
    const uint32_t vd = part * n_pad + d;
    const uint32_t* src = evals.data();
    for (uint32_t r = 1; r < h; r++) {
      uint32_t* dst = pp[r & 1];
      for (uint32_t i = 0; i + r < h; i++) fd_init_item(f, src, dst, da.data(), db.data(), n_padv, h, r, i, vd);
      src = dst;
    }
    for (uint32_t tick = 1; tick <= plan.steps + h - 2; tick++) {
      int32_t k_lo, k_hi;
      fd_ext_band(h, plan.steps, tick, &k_lo, &k_hi);
      // items of one tick are independent: run them in descending order to catch any accidental
      // dependence on the ascending order
      for (int32_t k = k_hi; k >= k_lo; k--)
        fd_ext_item(f, dd[(tick & 1) ^ 1], dd[tick & 1], evals.data(), n_padv, h, tick, (uint32_t)k, e_hi, vd);
    }
  }
  std::vector<int8_t> dig(fd_dig_bytes(m));
  std::vector<uint32_t> tab((size_t)(m > 1 ? m - 1 : 1) * FD_TAB_SLOTS * 36 * n_pad, 0xdeadbeef);
  for (uint32_t j = 0; j < n_r; j++) {
    int top = m > 1 ? fd_comb_digits(j + 1, h, m, dig.data()) : -1;
    fd_combine_eval(f, evals.data(), n_padv, n_pad, m, (size_t)((int64_t)(j + 1) - plan.lo), d, dig.data(), top, tab.data());
    g1_compress(g1_to_affine(vm_get_point(f, AX)), out48 + (size_t)j * 48);
  }
  return 0;
}

// signed digits of x^(h i) mod r: out[((i-1)*2 + half) * 132 + b]; returns top
int he_fd_digits(uint32_t x, uint32_t h, uint32_t m, int8_t* out) { return fd_comb_digits(x, h, m, out); }
}

// ---- tower / pairing / hash-to-G2 (tower.cuh, h2c.cuh) -------------------------------------------
#include "../../dvt_circuits_b200/csrc/h2c.cuh"
extern "C" {
uint32_t he_g2_decompress(const uint8_t* in96, uint8_t* out96) {
  G2Aff a;
  uint32_t st = g2_decompress(in96, &a, true);
  if (st == G1_DEC_OK) g2_compress(&a, out96);
  return st;
}
void he_hash_to_g2(const uint8_t* msg, size_t len, uint8_t* out96) {
  G2Aff h;
  hash_to_g2(&h, msg, len);
  g2_compress(&h, out96);
}
// 1 valid, 0 invalid, negative: -48 bad pk, -49 bad sig/hm
int he_bls_verify_hm(const uint8_t* pk48, const uint8_t* sig96, const uint8_t* hm96) {
  G1Aff pk;
  G2Aff sig, hm;
  if (g1_decompress(pk48, &pk, true) != G1_DEC_OK) return -48;
  if (g2_decompress(sig96, &sig, true) != G1_DEC_OK) return -49;
  if (g2_decompress(hm96, &hm, true) != G1_DEC_OK) return -49;
  int plain = bls_verify_precomputed(&pk, &sig, &hm) ? 1 : 0;
  if (!hm.inf) {  // the prepared-lines route of k_bls_verify must agree
    std::vector<G2Line> lines(G2_PREP_LINES);
    g2_prepare(lines.data(), &hm);
    int prep = bls_verify_prepared(&pk, &sig, &hm, lines.data()) ? 1 : 0;
    if (prep != plain) return -100;
  }
  return plain;
}
// e(P,Q)^3 as 576 canonical big-endian bytes, same layout as orc_pairing_bytes
int he_pairing_bytes(const uint8_t* p48, const uint8_t* q96, uint8_t* out576) {
  G1Aff p;
  G2Aff q;
  if (g1_decompress(p48, &p, true) != G1_DEC_OK || g2_decompress(q96, &q, true) != G1_DEC_OK) return -1;
  Fp12 f = fp12_one(), e;
  miller_loop_acc(&f, &p, &q);
  fp12_conj(&f, &f);
  final_exponentiation(&e, &f);
  const Fp* c = &e.c0.c0.c0;
  for (int i = 0; i < 12; i++) {
    Fp v = from_mont(c[i]);
    fp_raw_to_be48(out576 + 48 * i, v.l);
  }
  return 0;
}
}
// ---- the pairing VM (pairing_vm.cuh): the interpreter and the generated program, all PVM_R roles of ONE check run level by level
#include "../../dvt_circuits_b200/csrc/pairing_vm.cuh"
extern "C" {
// status as k_pairing_vm (0 ok, 7 invalid, 48 bad pk, 49 bad sig); out576: the final value of RA, canonical big-endian, or nullptr
int he_pairing_vm(const uint8_t* pk48, const uint8_t* sig96, const uint8_t* hm96, uint8_t* out576) {
  G1Aff pk;
  G2Aff sig, hm;
  if (g2_decompress(hm96, &hm, true) != G1_DEC_OK) return -1;
  uint32_t sst = g2_decompress(sig96, &sig, true), pst = g1_decompress(pk48, &pk, true);
  std::vector<G2Line> lines(G2_PREP_LINES), ref_lines(G2_PREP_LINES);
  const int lane = 5;  // any lane of the 32-wide file
  std::vector<U4> file((size_t)PVM_SLOTS * 6 * PVM_LANES);
  std::vector<U4> scratch((size_t)2 * 6 * 6);
  auto run_segment = [&](const PvmCtx& cc, const PvmCall& k) -> bool {
    uint32_t pc[PVM_R];
    for (uint32_t r = 0; r < PVM_R; r++) pc[r] = pvm_seg_start[k.a][r];
    for (;;) {  // one level: every role up to its barrier
      uint32_t ended = 0;
      for (uint32_t r = 0; r < PVM_R; r++) {
        pc[r] = pvm_exec(cc, pc[r], k.b);
        ended += (pc[r] & PVM_END_FLAG) ? 1 : 0;
      }
      if (ended == PVM_R) return true;
      if (ended != 0) return false;  // the roles of a segment must agree on the number of levels
    }
  };
  if (!hm.inf) {
    // the lines of the hashed message by the VM's own preparation program (k_g2_prepare_vm), which must reproduce tower.cuh's
    // g2_prepare bit for bit (same formulas, same projective representatives)
    g2_prepare(ref_lines.data(), &hm);
    PvmCtx pc{file.data() + lane, nullptr, nullptr, (const uint32_t*)&hm, (uint32_t*)lines.data()};
    pvm_init_point(pc);
    for (uint32_t ci = 0; ci < PVM_N_PREP_CALLS; ci++)
      if (!run_segment(pc, pvm_prep_call(ci))) return -2;
    if (memcmp(lines.data(), ref_lines.data(), sizeof(G2Line) * G2_PREP_LINES) != 0) return -3;
  }
  PvmCtx c{file.data() + lane, (const uint32_t*)lines.data(), (const uint32_t*)&pk, (const uint32_t*)&sig, nullptr};
  pvm_init_point(c);
  for (uint32_t ci = 0; ci < PVM_N_CALLS; ci++) {
    PvmCall k = pvm_call(ci);
    if (k.kind == 0) {
      if (!run_segment(c, k)) return -2;
    } else {
      for (uint32_t s2 = 0; s2 < 6; s2++)
        for (uint32_t ch = 0; ch < 6; ch++) {
          if (k.kind == 1) c.file[(size_t)((pvm_reg_slot(k.a) + s2) * 6 + ch) * PVM_LANES] = c.file[(size_t)((pvm_reg_slot(k.b) + s2) * 6 + ch) * PVM_LANES];
          else if (k.kind == 2) scratch[(k.a * 6 + s2) * 6 + ch] = c.file[(size_t)((pvm_reg_slot(k.b) + s2) * 6 + ch) * PVM_LANES];
          else c.file[(size_t)((pvm_reg_slot(k.a) + s2) * 6 + ch) * PVM_LANES] = scratch[(k.b * 6 + s2) * 6 + ch];
        }
    }
  }
  if (out576)
    for (uint32_t k = 0; k < 6; k++) {
      Fp a, b;
      pvm_ld_slot(c, SLOT_A0 + k, a, b);
      fp_raw_to_be48(out576 + 96 * k, from_mont(a).l);
      fp_raw_to_be48(out576 + 96 * k + 48, from_mont(b).l);
    }
  if (sst != G1_DEC_OK) return DKGV_PANIC_BAD_G2;
  if (pst != G1_DEC_OK) return DKGV_PANIC_BAD_G1;
  return pvm_status(pk.inf != 0, sig.inf != 0, hm.inf != 0, pvm_result_is_one(c));
}
void he_sha256(const uint8_t* msg, size_t len, uint8_t* out32) {
  Sha256 s;
  sha_init(&s);
  sha_update(&s, msg, len);
  sha_finish(&s, out32);
}

// G * s through a table of `bits`-bit windows holding just the entries s touches (digit logic for every width the ctx accepts)
void he_fixed_base_bits(uint32_t bits, const uint8_t* scalar32, uint8_t* out48) {
  uint32_t s[8], u[8];
  fr_raw_from_be32(s, scalar32);
  gtab_scalar(u, s);
  // a compact table: window w keeps only its one touched entry, the lookup is redirected through a private GTab per window
  G1Proj acc = g1_identity();
  const uint32_t W = gtab_windows(bits);
  for (uint32_t w = 0; w < W; w++) {
    bool ng;
    uint32_t idx = gtab_index(bits, W, u, w, &ng);
    G1Aff e = gtab_entry(bits, idx);
    if (ng) e.y = neg(e.y);
    acc = w ? g1_add_mixed_nz(acc, e.x, e.y) : g1_from_affine(e);
  }
  g1_compress(g1_to_affine(acc), out48);
}
// the table builder's per-thread routines (gtab_base, gtab_fill_run) for the run holding entry (w, m) against gtab_entry: 0 = all
// GTAB_RUN entries of the run agree
int he_gtab_builder(uint32_t bits, uint32_t w, uint32_t m) {
  std::vector<uint32_t> base(48 * (size_t)gtab_windows(bits), 0), tab(24 * (size_t)GTAB_RUN, 0);
  gtab_base(bits, w, base.data());
  const uint32_t m0 = m - m % GTAB_RUN;
  // gtab_fill_run addresses the full table: hand it a pointer offset so that its run lands in `tab`
  uint32_t* fake = tab.data() - ((size_t)(w << (bits - 1)) + m0) * 24;
  gtab_fill_run(bits, base.data(), w, m0, fake);
  int bad = 0;
  for (uint32_t i = 0; i < GTAB_RUN; i++) {
    G1Aff e = gtab_entry(bits, (w << (bits - 1)) + m0 + i);
    for (int k = 0; k < 12; k++) bad += tab[(size_t)i * 24 + k] != e.x.l[k] || tab[(size_t)i * 24 + 12 + k] != e.y.l[k];
  }
  return bad;
}

// prev - j * a mod r through fr_submul_small (canonical 32-byte big-endian in / out)
static void be32_to_fr(Fr& f, const uint8_t* b) { fr_raw_from_be32(f.l, b); }
static void fr_to_be32(uint8_t* out32, const Fr& r) {
  for (int i = 0; i < 8; i++) {
    uint8_t* q = out32 + 28 - 4 * i;
    q[0] = r.l[i] >> 24; q[1] = r.l[i] >> 16; q[2] = r.l[i] >> 8; q[3] = r.l[i];
  }
}
void he_fr_submul_small(const uint8_t* prev32, const uint8_t* a32, uint32_t j, uint8_t* out32) {
  Fr p, a;
  be32_to_fr(p, prev32);
  be32_to_fr(a, a32);
  fr_to_be32(out32, fr_submul_small(p, a, j));
}

// One block of k_fd_difftab (share_fd.cu) with its threads run in lockstep: per round every thread publishes, then (the
// barrier) every thread steps - the same dt1_* / dt2_* per-thread routines, predicates and double buffer as the kernel.
// shares: n_r canonical values s(1..n_r) [n_r][32]; ifact: 1/k! canonical [t][32] (the kernel's table holds them in
// Montgomery form).  Returns 0 and coef [t][32] when every t-th difference vanishes, 1 otherwise.
int he_difftab(const uint8_t* shares, uint32_t n_r, uint32_t t, const uint8_t* ifact, uint8_t* coef) {
  const uint32_t nt = (((n_r + 1) / 2 + 31) / 32) * 32;
  std::vector<DtPair> th(nt);
  std::vector<uint32_t> pub(18 * (size_t)nt);
  std::vector<Fr> E(t);
  for (uint32_t i = 0; i < nt; i++) {
    Fr a = zero<FrParams>(), b = zero<FrParams>();
    if (2 * i < n_r) be32_to_fr(a, shares + (size_t)(2 * i) * 32);
    if (2 * i + 1 < n_r) be32_to_fr(b, shares + (size_t)(2 * i + 1) * 32);
    th[i].a = lz_from(a);
    th[i].b = lz_from(b);
  }
  uint32_t left = DT1_PERIOD;
  for (uint32_t r = 1; r <= t; r++) {
    uint32_t* pr = pub.data() + (size_t)(r & 1) * 9 * nt;
    for (uint32_t i = 0; i < nt; i++)
      if (dt1_publishes(i, r)) lz_publish(pr, nt, i, th[i].b);
    const bool red = --left == 0;
    if (red) left = DT1_PERIOD;
    for (uint32_t i = 0; i < nt; i++)
      if (dt1_active(i, r)) dt1_step(th[i], i, r, red, pr, nt);
  }
  bool bad = false;
  for (uint32_t i = 0; i < nt; i++) {
    dt1_finish(th[i]);
    uint32_t k0 = 2 * i, k1 = 2 * i + 1;
    bad |= (k0 >= t && k0 < n_r && !lz_is_zero(th[i].a)) || (k1 >= t && k1 < n_r && !lz_is_zero(th[i].b));
  }
  if (bad) return 1;
  for (uint32_t i = 0; i < nt; i++) {
    Fr f;
    if (2 * i < t) {
      be32_to_fr(f, ifact + (size_t)(2 * i) * 32);
      E[2 * i] = dt2_signed_e(mul(lz_low(th[i].a), to_mont(f)), t, 2 * i);
    }
    if (2 * i + 1 < t) {
      be32_to_fr(f, ifact + (size_t)(2 * i + 1) * 32);
      E[2 * i + 1] = dt2_signed_e(mul(lz_low(th[i].b), to_mont(f)), t, 2 * i + 1);
    }
  }
  for (uint32_t i = 0; i < nt; i++) {
    th[i].a = i == 0 ? lz_from(E[t - 1]) : lz_zero();
    th[i].b = lz_zero();
  }
  const uint32_t period = dt2_period(t);
  left = period;
  for (uint32_t j = t - 1; j >= 1; j--) {
    uint32_t* pr = pub.data() + (size_t)(j & 1) * 9 * nt;
    for (uint32_t i = 0; i < nt; i++)
      if (dt2_active(i, j, t)) lz_publish(pr, nt, i, th[i].b);
    const bool red = --left == 0;
    if (red) left = period;
    for (uint32_t i = 0; i < nt; i++)
      if (dt2_active(i, j, t)) dt2_step(th[i], i, j, red, pr, nt, E.data());
  }
  for (uint32_t i = 0; i < nt; i++) {
    dt2_finish(th[i], i, t);
    if (2 * i < t) fr_to_be32(coef + (size_t)(2 * i) * 32, lz_low(th[i].a));
    if (2 * i + 1 < t) fr_to_be32(coef + (size_t)(2 * i + 1) * 32, lz_low(th[i].b));
  }
  return 0;
}
// lz_reduce on a 36-byte little-endian-limb value (9 x u32, host order): out = the canonical residue, 32 bytes big-endian
void he_lz_reduce(const uint32_t* limbs9, int is_signed, uint8_t* out32) {
  Lz v;
  for (int i = 0; i < 9; i++) v.l[i] = limbs9[i];
  lz_reduce(v, is_signed != 0);
  out32[0] = v.l[8] ? 0xff : 0;  // poisons the answer if the top limb is not clear
  uint8_t tmp[32];
  fr_to_be32(tmp, lz_low(v));
  for (int i = 0; i < 32; i++) out32[i] = (i == 0 ? out32[0] : 0) | tmp[i];
}

// Condition (3) of the consistency shortcut against COMPRESSED commitments (fdiff.cuh fd_coef_point + fd_coef_signs, the
// per-thread routines of k_fd_coefpoint / k_fd_coefsign): n (scalar, encoding) pairs of one dealer, sign halves in batches of
// FD_SIGN_K as in the kernel.  out[i] = 1 when compress(G * scalar_i) == enc_i according to the two halves.
void he_coef_bytes_check(const uint8_t* scalars32, const uint8_t* enc48, uint32_t n, uint8_t* out) {
  const uint32_t NT = 4, me = 1;
  std::vector<U4> file((size_t)VM_SLOTS * 3 * NT);
  OpFile f{file.data() + me, NT};
  std::vector<Fp> ys(n), zs(n);
  std::vector<uint8_t> same(n);
  for (uint32_t i = 0; i < n; i++) {
    uint32_t sc[8];
    fr_raw_from_be32(sc, scalars32 + (size_t)i * 32);
    GTab gtab = he_gtab_for(sc);
    same[i] = fd_coef_point(f, gtab, sc, enc48 + (size_t)i * 48, &ys[i], &zs[i]);
  }
  for (uint32_t k0 = 0; k0 < n; k0 += FD_SIGN_K) {
    int cnt = (int)std::min<uint32_t>(FD_SIGN_K, n - k0);
    uint8_t fs[FD_SIGN_K];
    for (int i = 0; i < cnt; i++) fs[i] = (enc48[(size_t)(k0 + i) * 48] >> 5) & 1;
    // the kernel's verdict is per dealer (all batches); here each point separately as well: batch verdict with the other
    // points' signs forced right, so that one wrong sign is attributed to its own point
    for (int i = 0; i < cnt; i++) {
      uint8_t g[FD_SIGN_K];
      for (int j = 0; j < cnt; j++) {
        Fp ya = mul(ys[k0 + j], fp_inv(zs[k0 + j]));
        g[j] = j == i ? fs[j] : (uint8_t)fp_lex_largest(ya);
      }
      bool ok = fd_coef_signs<FD_SIGN_K>(zs.data() + k0, ys.data() + k0, g, cnt);
      out[k0 + i] = same[k0 + i] && ok;
    }
  }
}
}
