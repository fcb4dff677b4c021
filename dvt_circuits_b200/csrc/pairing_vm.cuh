// Pairing VM: the batched BLS check e(pk, H(m)) == e(G1, sig) (crates/dkg/src/crypto/bls_common.rs:26-35) executed by R
// cooperating warps per 32 checks, every Fp2 of a check in a shared-memory slot - nothing in local memory.
//
// Why: round 1's k_bls_verify was one thread per check with the Fp12 values (576 B each, ~9 KB of live state per thread) in local
// memory; ncu showed 3.9 GB read + 33.4 GB written to DRAM for 32 768 checks (~1.1 MB of spill traffic per check against 145
// algorithmic bytes) and one check's latency (~40 ms) as the floor of any small batch.  Here
//   * a block is PVM_R warps ("roles"), lane = check, so the 32 lanes of a role always run the same instruction on the same slot:
//     slot s, 16-byte chunk c of lane l sits at file[(s * 6 + c) * 32 + l] - conflict-free LDS.128 / STS.128;
//   * the program (tools/gen_pairing_vm.py -> pairing_prog.inc) is a per-role stream of register-machine instructions over two Fp2
//     registers X, Y: signed sums of slots into X / Y, X <- X * Y, X <- X^2, small multiples, store.  The roles meet at a block
//     barrier after every dependency level of the formulas; the generator schedules the Fp2 products of a level over the roles;
//   * each product routine exists once (this interpreter loop), the program itself sits in constant memory (warp-uniform fetch);
//   * 32 slots x 96 B per check = 96 KB per block: two blocks (2 x PVM_R warps) per SM;
//   * a check's latency is ~1/PVM_R of the serial one: finalization's 1 024 checks no longer wait for one thread.
// The formulas are tower.cuh's (same line functions, same final exponentiation f^(3 (p^12 - 1) / r)), the two lines of a Miller
// step merged into one sparse element first (6 + 17 Fp2 products instead of 2 x 13).  The generator proves the program against
// the Python restatement of the reference before it writes it; tests/test_pairing_vm.py runs THIS interpreter on the host.
#pragma once
#include "tower.cuh"
#include "vm.cuh"

// program, call list and constants: constant memory on the device (warp-uniform fetch), plain arrays for the host emulation
#if defined(__CUDACC__)
#define PVM_CONST __constant__ const
#pragma nv_diag_suppress 20091  // the DKGV_HD interpreter below is only ever RUN on the device in nvcc builds
#else
#define PVM_CONST static const
#endif
#ifndef PVM_PROG_FILE
#define PVM_PROG_FILE "pairing_prog.inc"
#endif
#include PVM_PROG_FILE

namespace dkgv {

constexpr uint32_t PVM_END_FLAG = 0x80000000u;
constexpr int PVM_LANES = 32;  // checks per block

struct PvmCtx {
  U4* file;              // this lane's chunk 0 of slot 0 (stride PVM_LANES between chunks)
  const uint32_t* line;  // this check's prepared lines of H(m): G2Line[G2_PREP_LINES] as words (72 per line)
  const uint32_t* pk;    // G1Aff words: x[12] y[12]
  const uint32_t* sig;   // G2Aff words: x.c0[12] x.c1[12] y.c0[12] y.c1[12]
  uint32_t* line_out;    // line preparation only: where STXL writes (same layout as `line`)
};

DKGV_HD void pvm_ld_slot(const PvmCtx& c, uint32_t s, Fp& a, Fp& b) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    U4 v = c.file[(size_t)(s * 6 + k) * PVM_LANES], w = c.file[(size_t)(s * 6 + 3 + k) * PVM_LANES];
    a.l[4 * k] = v.x, a.l[4 * k + 1] = v.y, a.l[4 * k + 2] = v.z, a.l[4 * k + 3] = v.w;
    b.l[4 * k] = w.x, b.l[4 * k + 1] = w.y, b.l[4 * k + 2] = w.z, b.l[4 * k + 3] = w.w;
  }
}
DKGV_HD void pvm_st_slot(const PvmCtx& c, uint32_t s, const Fp& a, const Fp& b) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    U4 v{a.l[4 * k], a.l[4 * k + 1], a.l[4 * k + 2], a.l[4 * k + 3]}, w{b.l[4 * k], b.l[4 * k + 1], b.l[4 * k + 2], b.l[4 * k + 3]};
    c.file[(size_t)(s * 6 + k) * PVM_LANES] = v;
    c.file[(size_t)(s * 6 + 3 + k) * PVM_LANES] = w;
  }
}
DKGV_HD void pvm_ld_words(const uint32_t* p, Fp& a, Fp& b) {  // 24 consecutive words, 16-byte aligned
  const U4* q = (const U4*)p;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    U4 v = q[k], w = q[3 + k];
    a.l[4 * k] = v.x, a.l[4 * k + 1] = v.y, a.l[4 * k + 2] = v.z, a.l[4 * k + 3] = v.w;
    b.l[4 * k] = w.x, b.l[4 * k + 1] = w.y, b.l[4 * k + 2] = w.z, b.l[4 * k + 3] = w.w;
  }
}
DKGV_HD void pvm_ld_const(uint32_t k, Fp& a, Fp& b) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    a.l[i] = pvm_consts[k][i];
    b.l[i] = pvm_consts[k][12 + i];
  }
}
// input k of the check as an Fp2: 0 = (xp, yp) of the key, 1 = sig.x, 2 = sig.y
DKGV_HD void pvm_ld_input(const PvmCtx& c, uint32_t k, Fp& a, Fp& b) { pvm_ld_words(k == 0 ? c.pk : c.sig + (k - 1) * 24, a, b); }

// Executes this role's stream from pc up to (and including) the next BAR or END.  Returns the next pc, | PVM_END_FLAG at END.
// line_idx: which prepared line the LDXL / ADDXL of this segment refer to.
DKGV_HD uint32_t pvm_exec(const PvmCtx& c, uint32_t pc, uint32_t line_idx) {
  Fp x0 = zero<FpParams>(), x1 = x0, y0 = x0, y1 = x0, t0, t1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (;;) {
    const uint32_t w = pvm_prog[pc++], arg = w >> 8;
    switch (w & 0xff) {
      case PVM_END: return pc | PVM_END_FLAG;
      case PVM_BAR: return pc;
      case PVM_LDX: pvm_ld_slot(c, arg, x0, x1); break;
      case PVM_ADDX: pvm_ld_slot(c, arg, t0, t1); x0 = add(x0, t0); x1 = add(x1, t1); break;
      case PVM_SUBX: pvm_ld_slot(c, arg, t0, t1); x0 = sub(x0, t0); x1 = sub(x1, t1); break;
      case PVM_LDY: pvm_ld_slot(c, arg, y0, y1); break;
      case PVM_ADDY: pvm_ld_slot(c, arg, t0, t1); y0 = add(y0, t0); y1 = add(y1, t1); break;
      case PVM_SUBY: pvm_ld_slot(c, arg, t0, t1); y0 = sub(y0, t0); y1 = sub(y1, t1); break;
      case PVM_STX: pvm_st_slot(c, arg, x0, x1); break;
      case PVM_LDXK: pvm_ld_const(arg, x0, x1); break;
      case PVM_LDYK: pvm_ld_const(arg, y0, y1); break;
      case PVM_LDXL: pvm_ld_words(c.line + ((size_t)line_idx * 3 + arg) * 24, x0, x1); break;
      case PVM_ADDXL: pvm_ld_words(c.line + ((size_t)line_idx * 3 + arg) * 24, t0, t1); x0 = add(x0, t0); x1 = add(x1, t1); break;
      case PVM_LDXIN: pvm_ld_input(c, arg, x0, x1); break;
      case PVM_ADDXIN: pvm_ld_input(c, arg, t0, t1); x0 = add(x0, t0); x1 = add(x1, t1); break;
      case PVM_SUBXIN: pvm_ld_input(c, arg, t0, t1); x0 = sub(x0, t0); x1 = sub(x1, t1); break;
      case PVM_LDYIN: pvm_ld_input(c, arg, y0, y1); break;
      case PVM_ADDYIN: pvm_ld_input(c, arg, t0, t1); y0 = add(y0, t0); y1 = add(y1, t1); break;
      case PVM_LDYS: {  // Y = (xp or yp, 0): scaling of a prepared line coefficient by the key
        pvm_ld_input(c, 0, t0, t1);
        y0 = arg ? t1 : t0;
        y1 = zero<FpParams>();
        break;
      }
      case PVM_STXL: {  // coefficient arg of the line being prepared
        U4* q = (U4*)(c.line_out + ((size_t)line_idx * 3 + arg) * 24);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          q[k] = U4{x0.l[4 * k], x0.l[4 * k + 1], x0.l[4 * k + 2], x0.l[4 * k + 3]};
          q[3 + k] = U4{x1.l[4 * k], x1.l[4 * k + 1], x1.l[4 * k + 2], x1.l[4 * k + 3]};
        }
        break;
      }
      case PVM_MUL: {  // (x0 y0 - x1 y1, x0 y1 + x1 y0): two fused sum-of-two-products routines
        t0 = neg(y1);
        mul2add_pair(t0, t1, x0, y0, x1, t0, x0, y1, x1, y0);  // the two sums interleaved: four carry chains in flight
        x0 = t0;
        x1 = t1;
        break;
      }
      case PVM_SQR: {  // ((x0 + x1)(x0 - x1), 2 x0 x1)
        t0 = add(x0, x1);
        t1 = sub(x0, x1);
        mul_pair(t0, t1, t0, t1, x0, x1);  // the two products interleaved
        x0 = t0;
        x1 = dbl(t1);
        break;
      }
      case PVM_XI: t0 = sub(x0, x1); x1 = add(x0, x1); x0 = t0; break;
      case PVM_NEGX: x0 = neg(x0); x1 = neg(x1); break;
      case PVM_DBLX: x0 = dbl(x0); x1 = dbl(x1); break;
      case PVM_TPLX: x0 = add(dbl(x0), x0); x1 = add(dbl(x1), x1); break;
      case PVM_CONJX: x1 = neg(x1); break;
      case PVM_INVX: {  // 1 / (x0 + x1 u) = (x0 - x1 u) / (x0^2 + x1^2); 0 -> 0
        t0 = add(mul(x0, x0), mul(x1, x1));
        t0 = fp_inv_bgcd(t0);
        x0 = mul(x0, t0);
        x1 = neg(mul(x1, t0));
        break;
      }
      default: return pc | PVM_END_FLAG;  // not reachable: the generator only emits the opcodes above
    }
  }
}

// the six Fp2 of register RA (reg 0: slots SLOT_A0..) / RB (reg 1: SLOT_B0..)
DKGV_HD uint32_t pvm_reg_slot(uint32_t reg) { return reg ? (uint32_t)SLOT_B0 : (uint32_t)SLOT_A0; }

// decoding of one entry of the driver's call list
struct PvmCall {
  uint32_t kind, a, b;  // 0: segment a, line index b | 1: COPY reg a <- reg b | 2: SPILL global a <- reg b | 3: FILL reg a <- global b
};
DKGV_HD PvmCall pvm_call(uint32_t i) {
  uint32_t w = pvm_calls[i];
  return PvmCall{w & 15u, (w >> 4) & 0xffu, w >> 12};
}
DKGV_HD PvmCall pvm_prep_call(uint32_t i) {  // call list of the line preparation (one point step per Miller step)
  uint32_t w = pvm_prep_calls[i];
  return PvmCall{w & 15u, (w >> 4) & 0xffu, w >> 12};
}

// final verdict of a check from RA: the pairing product is 1
DKGV_HD bool pvm_result_is_one(const PvmCtx& c) {
  Fp a, b;
  pvm_ld_slot(c, SLOT_A0, a, b);
  bool ok = eq(a, one<FpParams>()) && is_zero(b);
  for (uint32_t k = 1; k < 6; k++) {
    pvm_ld_slot(c, SLOT_A0 + k, a, b);
    ok = ok && is_zero(a) && is_zero(b);
  }
  return ok;
}

// T <- (sig.x, sig.y, 1): the running point of the Miller loop starts at the signature
DKGV_HD void pvm_init_point(const PvmCtx& c) {
  Fp a, b;
  pvm_ld_input(c, 1, a, b);
  pvm_st_slot(c, SLOT_TX, a, b);
  pvm_ld_input(c, 2, a, b);
  pvm_st_slot(c, SLOT_TY, a, b);
  pvm_st_slot(c, SLOT_TZ, one<FpParams>(), zero<FpParams>());
}

// status of a check whose arguments decoded (bls12_381::pairing semantics for identity arguments):
// a pairing with an identity argument is the Gt identity, and e(P, Q) = 1 for subgroup points only when one of them is the identity
DKGV_HD uint8_t pvm_status(bool pk_inf, bool sig_inf, bool hm_inf, bool product_is_one) {
  const bool lhs_one = pk_inf || hm_inf, rhs_one = sig_inf;
  const bool ok = (lhs_one || rhs_one) ? (lhs_one && rhs_one) : product_is_one;
  return ok ? DKGV_OK : DKGV_SLASHABLE_SIG_INVALID;
}

}  // namespace dkgv
