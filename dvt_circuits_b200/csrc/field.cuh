// Montgomery prime-field arithmetic on 32-bit limbs for BLS12-381 Fp (12 limbs) and Fr (8 limbs).
//
// Replaces, on the GPU, what the reference gets from the un-vendored `bls12_381` crate
// (`Fp`, `Scalar`; call sites crates/dkg/src/dkg_math.rs:43-98,114-127).  Written from the
// textbook algorithms (operand-scanning Montgomery product with parity-split accumulators so
// that every 32x32->64 product is one lo/hi `mad` pair on a single carry chain).
//
// Device path: inline PTX `mad.lo.cc.u32` / `madc.hi.cc.u32` chains (ptxas fuses each pair into
// IMAD.WIDE.U32[.X] on sm_100a).  Host path (DKGV_HOST_EMU builds of the test harness only, never
// the product): portable CIOS on uint64_t, bit-identical results.
//
// Representation: fully reduced values in [0, mod), Montgomery form (x*R mod p, R = 2^(32N)).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DKGV_HD __host__ __device__ __forceinline__
#define DKGV_D __device__ __forceinline__
#else
#define DKGV_HD inline
#define DKGV_D inline
#endif

namespace dkgv {

#include "constants.cuh"

#if defined(__CUDA_ARCH__)
namespace ptx {
// acc[j], acc[j+1] = a[j] * b  for j = 0, 2, ..., n-2 (independent wide products)
template <int n>
DKGV_D void mul_n(uint32_t* acc, const uint32_t* a, uint32_t b) {
#pragma unroll
  for (int j = 0; j < n; j += 2)
    asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(acc[j]), "=r"(acc[j + 1]) : "r"(a[j]), "r"(b));
}
// acc[j], acc[j+1] += a[j] * b on one carry chain; carry-out is left in CC
template <int n>
DKGV_D void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(b));
#pragma unroll
  for (int j = 2; j < n; j += 2)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[j]), "+r"(acc[j + 1]) : "r"(a[j]), "r"(b));
}
// same, but the multiplicand limbs are compile-time constants PR::mod(j + off)
template <class PR, int off>
DKGV_D void cmad_mod(uint32_t* acc, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(PR::mod(off)), "r"(b));
#pragma unroll
  for (int j = 2; j < PR::N; j += 2)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                 : "+r"(acc[j]), "+r"(acc[j + 1])
                 : "r"(PR::mod(j + off < PR::N ? j + off : 0)), "r"(b));
}
// acc[j], acc[j+1] = a[j] * b + acc[j+2], acc[j+3] (+ incoming CC); top pair gets product + carry
template <int n>
DKGV_D void madc_n_rshift(uint32_t* acc, const uint32_t* a, uint32_t b) {
#pragma unroll
  for (int j = 0; j < n - 2; j += 2)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                 : "=r"(acc[j]), "=r"(acc[j + 1])
                 : "r"(a[j]), "r"(b), "r"(acc[j + 2]), "r"(acc[j + 3]));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;" : "=r"(acc[n - 2]), "=r"(acc[n - 1]) : "r"(a[n - 2]), "r"(b));
}
// EXPERIMENT (-DDKGV_RED_TWO_PIPE, off by default: measured slower, profiles/r1_fp30_experiment.md).
// Reduction row on TWO pipes.  m * mod[k] + E[k] never overflows 64 bits, so the 12 products are
// carry-free wide multiply-adds (IMAD.WIDE.U32 at the full FMA-pipe rate, bench/imad_carry.cu) whose
// low words ARE the new E[k]; the high words are folded into the next column by one add/addc chain
// on the otherwise idle ALU pipe.  E: column-aligned array (12 limbs), O: one column up; the row
// m * mod spans columns 0..12, so the last high word and the chain's carry land in O[N-1].
// Same sums as cmad_mod<PR,1>(O, m); cmad_mod<PR,0>(E, m); addc O[N-1].
template <class PR>
DKGV_D void red_row_alu(uint32_t* E, uint32_t* O, uint32_t m) {
  constexpr int N = PR::N;
  uint32_t hi[N];
#pragma unroll
  for (int k = 0; k < N; k++) {
    unsigned long long w = (unsigned long long)m * PR::mod(k) + E[k];
    E[k] = (uint32_t)w;
    hi[k] = (uint32_t)(w >> 32);
  }
  asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(E[1]) : "r"(hi[0]));
#pragma unroll
  for (int k = 2; k < N; k++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(E[k]) : "r"(hi[k - 1]));
  asm volatile("addc.u32 %0, %0, %1;" : "+r"(O[N - 1]) : "r"(hi[N - 1]));
}
}  // namespace ptx
#endif

template <class PR>
struct alignas(16) Mont {
  static constexpr int N = PR::N;
  uint32_t l[N];
};

// ---- comparisons / selects (plain C, both paths)
template <class PR>
DKGV_HD bool is_zero(const Mont<PR>& a) {
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < PR::N; i++) x |= a.l[i];
  return x == 0;
}
template <class PR>
DKGV_HD bool eq(const Mont<PR>& a, const Mont<PR>& b) {
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < PR::N; i++) x |= a.l[i] ^ b.l[i];
  return x == 0;
}
template <class PR>
DKGV_HD Mont<PR> select(const Mont<PR>& a, const Mont<PR>& b, bool take_b) {
  Mont<PR> r;
#pragma unroll
  for (int i = 0; i < PR::N; i++) r.l[i] = take_b ? b.l[i] : a.l[i];
  return r;
}
template <class PR>
DKGV_HD Mont<PR> zero() {
  Mont<PR> r;
#pragma unroll
  for (int i = 0; i < PR::N; i++) r.l[i] = 0;
  return r;
}
template <class PR>
DKGV_HD Mont<PR> one() {
  Mont<PR> r;
#pragma unroll
  for (int i = 0; i < PR::N; i++) r.l[i] = PR::one(i);
  return r;
}

// raw (non-modular) helpers ---------------------------------------------------------------
// r = a - mod if a >= mod else a      (a < 2*mod, optional extra top carry bit)
template <class PR>
DKGV_HD void cond_sub_mod(uint32_t* a, uint32_t top) {
  constexpr int N = PR::N;
  uint32_t t[N];
#if defined(__CUDA_ARCH__)
  uint32_t borrow;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(t[0]) : "r"(a[0]), "r"(PR::mod(0)));
#pragma unroll
  for (int i = 1; i < N; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t[i]) : "r"(a[i]), "r"(PR::mod(i)));
  asm volatile("subc.u32 %0, %1, 0;" : "=r"(borrow) : "r"(top));
  bool ge = (borrow == 0);  // top:a - mod did not underflow
#else
  uint64_t br = 0;
  for (int i = 0; i < N; i++) {
    uint64_t d = (uint64_t)a[i] - PR::mod(i) - br;
    t[i] = (uint32_t)d;
    br = (d >> 32) & 1;
  }
  bool ge = ((uint64_t)top - br) >> 63 == 0;
#endif
#pragma unroll
  for (int i = 0; i < N; i++) a[i] = ge ? t[i] : a[i];
}

template <class PR>
DKGV_HD Mont<PR> add(const Mont<PR>& a, const Mont<PR>& b) {
  constexpr int N = PR::N;
  Mont<PR> r;
#if defined(__CUDA_ARCH__)
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(a.l[0]), "r"(b.l[0]));
#pragma unroll
  for (int i = 1; i < N; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.l[i]) : "r"(a.l[i]), "r"(b.l[i]));
#else
  uint64_t c = 0;
  for (int i = 0; i < N; i++) {
    c += (uint64_t)a.l[i] + b.l[i];
    r.l[i] = (uint32_t)c;
    c >>= 32;
  }
#endif
  cond_sub_mod<PR>(r.l, 0);  // both moduli leave a spare top bit: a + b < 2^(32N)
  return r;
}

template <class PR>
DKGV_HD Mont<PR> sub(const Mont<PR>& a, const Mont<PR>& b) {
  constexpr int N = PR::N;
  Mont<PR> r;
#if defined(__CUDA_ARCH__)
  uint32_t borrow;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(a.l[0]), "r"(b.l[0]));
#pragma unroll
  for (int i = 1; i < N; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r.l[i]) : "r"(a.l[i]), "r"(b.l[i]));
  asm volatile("subc.u32 %0, 0, 0;" : "=r"(borrow));
  uint32_t m = borrow;  // 0xffffffff when a < b
  asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(r.l[0]) : "r"(PR::mod(0) & m));
#pragma unroll
  for (int i = 1; i < N; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r.l[i]) : "r"(PR::mod(i) & m));
#else
  uint64_t br = 0;
  for (int i = 0; i < N; i++) {
    uint64_t d = (uint64_t)a.l[i] - b.l[i] - br;
    r.l[i] = (uint32_t)d;
    br = (d >> 32) & 1;
  }
  uint32_t m = br ? 0xffffffffu : 0u;
  uint64_t c = 0;
  for (int i = 0; i < N; i++) {
    c += (uint64_t)r.l[i] + (PR::mod(i) & m);
    r.l[i] = (uint32_t)c;
    c >>= 32;
  }
#endif
  return r;
}

template <class PR>
DKGV_HD Mont<PR> neg(const Mont<PR>& a) {
  return sub(zero<PR>(), a);
}
template <class PR>
DKGV_HD Mont<PR> dbl(const Mont<PR>& a) {
  return add(a, a);
}

// Montgomery product a*b*R^-1 mod p
template <class PR>
DKGV_HD Mont<PR> mul(const Mont<PR>& a, const Mont<PR>& b) {
  constexpr int N = PR::N;
  Mont<PR> r;
#if defined(__CUDA_ARCH__)
  // Parity-split operand scanning: `even` holds columns c, c+1, ... and `odd` columns c+1, ...
  // of the running sum; after each reduction step the lowest column vanishes and the two
  // arrays swap roles (a one-limb shift turns odd alignment into even alignment).
  uint32_t ev[N], od[N];
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    // ---- row i : ev is column-aligned
    if (i == 0) {
      ptx::mul_n<N>(od, a.l + 1, b.l[0]);
      ptx::mul_n<N>(ev, a.l, b.l[0]);
    } else {
      asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(ev[0]) : "r"(od[1]));
      ptx::madc_n_rshift<N>(od, a.l + 1, b.l[i]);
      ptx::cmad_n<N>(ev, a.l, b.l[i]);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(od[N - 1]));
    }
    {
      uint32_t m = ev[0] * PR::INV;
#ifndef DKGV_RED_TWO_PIPE
      ptx::cmad_mod<PR, 1>(od, m);
      ptx::cmad_mod<PR, 0>(ev, m);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(od[N - 1]));
#else
      ptx::red_row_alu<PR>(ev, od, m);
#endif
    }
    // ---- row i+1 : roles swapped (od is column-aligned)
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(od[0]) : "r"(ev[1]));
    ptx::madc_n_rshift<N>(ev, a.l + 1, b.l[i + 1]);
    ptx::cmad_n<N>(od, a.l, b.l[i + 1]);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(ev[N - 1]));
    {
      uint32_t m = od[0] * PR::INV;
#ifndef DKGV_RED_TWO_PIPE
      ptx::cmad_mod<PR, 1>(ev, m);
      ptx::cmad_mod<PR, 0>(od, m);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(ev[N - 1]));
#else
      ptx::red_row_alu<PR>(od, ev, m);
#endif
    }
  }
  // merge: the last row (index N-1, odd) ran with `od` column-aligned, so od[0] == 0 is the
  // vanished column and `ev` sits one column above:  r[k] = od[k+1] + ev[k]
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(od[1]), "r"(ev[0]));
#pragma unroll
  for (int k = 1; k < N - 1; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.l[k]) : "r"(od[k + 1]), "r"(ev[k]));
  asm volatile("addc.u32 %0, %1, 0;" : "=r"(r.l[N - 1]) : "r"(ev[N - 1]));
  cond_sub_mod<PR>(r.l, 0);
#else
  uint32_t t[N + 2];
  for (int i = 0; i < N + 2; i++) t[i] = 0;
  for (int i = 0; i < N; i++) {
    uint64_t c = 0;
    for (int j = 0; j < N; j++) {
      uint64_t uv = (uint64_t)a.l[j] * b.l[i] + t[j] + c;
      t[j] = (uint32_t)uv;
      c = uv >> 32;
    }
    uint64_t uv = (uint64_t)t[N] + c;
    t[N] = (uint32_t)uv;
    t[N + 1] = (uint32_t)(uv >> 32);
    uint32_t m = t[0] * PR::INV;
    uv = (uint64_t)m * PR::mod(0) + t[0];
    c = uv >> 32;
    for (int j = 1; j < N; j++) {
      uv = (uint64_t)m * PR::mod(j) + t[j] + c;
      t[j - 1] = (uint32_t)uv;
      c = uv >> 32;
    }
    uv = (uint64_t)t[N] + c;
    t[N - 1] = (uint32_t)uv;
    t[N] = t[N + 1] + (uint32_t)(uv >> 32);
  }
  for (int i = 0; i < N; i++) r.l[i] = t[i];
  cond_sub_mod<PR>(r.l, t[N]);
#endif
  return r;
}

// (a*b + c*d) * R^-1 mod p with ONE interleaved Montgomery reduction: per row the two multiplicand rows
// a*b_i and c*d_i are accumulated before the single reduction row m*p, i.e. 3 x N^2 + N wide products
// instead of the 4 x N^2 + 2N of two separate products and an addition (444 vs 600 for Fp).  The running
// value stays below 3p + 1 (a, c < p), which the spare top bits of both moduli absorb: 3 * 0x1a0111eb < 2^32
// for Fp, so no chain can carry out of the top limb; two conditional subtractions bring the result below p.
// Same value as add(mul(a, b), mul(c, d)) - used for the "sum of two products" lines of the RCB formulas.
template <class PR>
DKGV_HD Mont<PR> mul2add(const Mont<PR>& a, const Mont<PR>& b, const Mont<PR>& c, const Mont<PR>& d) {
#if defined(__CUDA_ARCH__)
  constexpr int N = PR::N;
  Mont<PR> r;
  uint32_t ev[N], od[N];
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    // ---- row i : ev is column-aligned
    if (i == 0) {
      ptx::mul_n<N>(od, a.l + 1, b.l[0]);
      ptx::mul_n<N>(ev, a.l, b.l[0]);
    } else {
      asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(ev[0]) : "r"(od[1]));
      ptx::madc_n_rshift<N>(od, a.l + 1, b.l[i]);
      ptx::cmad_n<N>(ev, a.l, b.l[i]);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(od[N - 1]));
    }
    ptx::cmad_n<N>(od, c.l + 1, d.l[i]);
    ptx::cmad_n<N>(ev, c.l, d.l[i]);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(od[N - 1]));
    {
      uint32_t m = ev[0] * PR::INV;
      ptx::cmad_mod<PR, 1>(od, m);
      ptx::cmad_mod<PR, 0>(ev, m);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(od[N - 1]));
    }
    // ---- row i+1 : roles swapped (od is column-aligned)
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(od[0]) : "r"(ev[1]));
    ptx::madc_n_rshift<N>(ev, a.l + 1, b.l[i + 1]);
    ptx::cmad_n<N>(od, a.l, b.l[i + 1]);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(ev[N - 1]));
    ptx::cmad_n<N>(ev, c.l + 1, d.l[i + 1]);
    ptx::cmad_n<N>(od, c.l, d.l[i + 1]);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(ev[N - 1]));
    {
      uint32_t m = od[0] * PR::INV;
      ptx::cmad_mod<PR, 1>(ev, m);
      ptx::cmad_mod<PR, 0>(od, m);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(ev[N - 1]));
    }
  }
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(od[1]), "r"(ev[0]));
#pragma unroll
  for (int k = 1; k < N - 1; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r.l[k]) : "r"(od[k + 1]), "r"(ev[k]));
  asm volatile("addc.u32 %0, %1, 0;" : "=r"(r.l[N - 1]) : "r"(ev[N - 1]));
  cond_sub_mod<PR>(r.l, 0);
  cond_sub_mod<PR>(r.l, 0);
  return r;
#else
  return add(mul(a, b), mul(c, d));
#endif
}

// Two independent sums of two products, (a*b + c*d) and (e*f + g*h), with their rows INTERLEAVED in program order: each product is
// two carry chains (even / odd columns) whose links depend on one another through the carry flag, so one product alone offers the
// scheduler two-way instruction-level parallelism; side by side the two give four independent chains.  For the Fp2 product of the
// pairing VM (pairing_vm.cuh), where a warp is alone on its scheduler much of the time.  Same values as two mul2add calls.
template <class PR>
DKGV_HD void mul2add_pair(Mont<PR>& r0, Mont<PR>& r1, const Mont<PR>& a, const Mont<PR>& b, const Mont<PR>& c, const Mont<PR>& d,
                          const Mont<PR>& e, const Mont<PR>& f, const Mont<PR>& g, const Mont<PR>& h) {
#if defined(__CUDA_ARCH__)
  constexpr int N = PR::N;
  uint32_t ev[N], od[N], fv[N], fd[N];
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    if (i == 0) {
      ptx::mul_n<N>(od, a.l + 1, b.l[0]);
      ptx::mul_n<N>(fd, e.l + 1, f.l[0]);
      ptx::mul_n<N>(ev, a.l, b.l[0]);
      ptx::mul_n<N>(fv, e.l, f.l[0]);
    } else {
      asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(ev[0]) : "r"(od[1]));
      ptx::madc_n_rshift<N>(od, a.l + 1, b.l[i]);
      ptx::cmad_n<N>(ev, a.l, b.l[i]);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(od[N - 1]));
      asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(fv[0]) : "r"(fd[1]));
      ptx::madc_n_rshift<N>(fd, e.l + 1, f.l[i]);
      ptx::cmad_n<N>(fv, e.l, f.l[i]);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(fd[N - 1]));
    }
    ptx::cmad_n<N>(od, c.l + 1, d.l[i]);
    ptx::cmad_n<N>(ev, c.l, d.l[i]);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(od[N - 1]));
    ptx::cmad_n<N>(fd, g.l + 1, h.l[i]);
    ptx::cmad_n<N>(fv, g.l, h.l[i]);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(fd[N - 1]));
    {
      uint32_t m = ev[0] * PR::INV, m2 = fv[0] * PR::INV;
      ptx::cmad_mod<PR, 1>(od, m);
      ptx::cmad_mod<PR, 0>(ev, m);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(od[N - 1]));
      ptx::cmad_mod<PR, 1>(fd, m2);
      ptx::cmad_mod<PR, 0>(fv, m2);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(fd[N - 1]));
    }
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(od[0]) : "r"(ev[1]));
    ptx::madc_n_rshift<N>(ev, a.l + 1, b.l[i + 1]);
    ptx::cmad_n<N>(od, a.l, b.l[i + 1]);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(ev[N - 1]));
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(fd[0]) : "r"(fv[1]));
    ptx::madc_n_rshift<N>(fv, e.l + 1, f.l[i + 1]);
    ptx::cmad_n<N>(fd, e.l, f.l[i + 1]);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(fv[N - 1]));
    ptx::cmad_n<N>(ev, c.l + 1, d.l[i + 1]);
    ptx::cmad_n<N>(od, c.l, d.l[i + 1]);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(ev[N - 1]));
    ptx::cmad_n<N>(fv, g.l + 1, h.l[i + 1]);
    ptx::cmad_n<N>(fd, g.l, h.l[i + 1]);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(fv[N - 1]));
    {
      uint32_t m = od[0] * PR::INV, m2 = fd[0] * PR::INV;
      ptx::cmad_mod<PR, 1>(ev, m);
      ptx::cmad_mod<PR, 0>(od, m);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(ev[N - 1]));
      ptx::cmad_mod<PR, 1>(fv, m2);
      ptx::cmad_mod<PR, 0>(fd, m2);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(fv[N - 1]));
    }
  }
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r0.l[0]) : "r"(od[1]), "r"(ev[0]));
#pragma unroll
  for (int k = 1; k < N - 1; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r0.l[k]) : "r"(od[k + 1]), "r"(ev[k]));
  asm volatile("addc.u32 %0, %1, 0;" : "=r"(r0.l[N - 1]) : "r"(ev[N - 1]));
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r1.l[0]) : "r"(fd[1]), "r"(fv[0]));
#pragma unroll
  for (int k = 1; k < N - 1; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r1.l[k]) : "r"(fd[k + 1]), "r"(fv[k]));
  asm volatile("addc.u32 %0, %1, 0;" : "=r"(r1.l[N - 1]) : "r"(fv[N - 1]));
  cond_sub_mod<PR>(r0.l, 0);
  cond_sub_mod<PR>(r0.l, 0);
  cond_sub_mod<PR>(r1.l, 0);
  cond_sub_mod<PR>(r1.l, 0);
#else
  Mont<PR> t0 = mul2add(a, b, c, d), t1 = mul2add(e, f, g, h);
  r0 = t0;
  r1 = t1;
#endif
}

// two independent products a*b and e*f with interleaved rows (see mul2add_pair)
template <class PR>
DKGV_HD void mul_pair(Mont<PR>& r0, Mont<PR>& r1, const Mont<PR>& a, const Mont<PR>& b, const Mont<PR>& e, const Mont<PR>& f) {
#if defined(__CUDA_ARCH__)
  constexpr int N = PR::N;
  uint32_t ev[N], od[N], fv[N], fd[N];
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    if (i == 0) {
      ptx::mul_n<N>(od, a.l + 1, b.l[0]);
      ptx::mul_n<N>(fd, e.l + 1, f.l[0]);
      ptx::mul_n<N>(ev, a.l, b.l[0]);
      ptx::mul_n<N>(fv, e.l, f.l[0]);
    } else {
      asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(ev[0]) : "r"(od[1]));
      ptx::madc_n_rshift<N>(od, a.l + 1, b.l[i]);
      ptx::cmad_n<N>(ev, a.l, b.l[i]);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(od[N - 1]));
      asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(fv[0]) : "r"(fd[1]));
      ptx::madc_n_rshift<N>(fd, e.l + 1, f.l[i]);
      ptx::cmad_n<N>(fv, e.l, f.l[i]);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(fd[N - 1]));
    }
    {
      uint32_t m = ev[0] * PR::INV, m2 = fv[0] * PR::INV;
      ptx::cmad_mod<PR, 1>(od, m);
      ptx::cmad_mod<PR, 0>(ev, m);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(od[N - 1]));
      ptx::cmad_mod<PR, 1>(fd, m2);
      ptx::cmad_mod<PR, 0>(fv, m2);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(fd[N - 1]));
    }
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(od[0]) : "r"(ev[1]));
    ptx::madc_n_rshift<N>(ev, a.l + 1, b.l[i + 1]);
    ptx::cmad_n<N>(od, a.l, b.l[i + 1]);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(ev[N - 1]));
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(fd[0]) : "r"(fv[1]));
    ptx::madc_n_rshift<N>(fv, e.l + 1, f.l[i + 1]);
    ptx::cmad_n<N>(fd, e.l, f.l[i + 1]);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(fv[N - 1]));
    {
      uint32_t m = od[0] * PR::INV, m2 = fd[0] * PR::INV;
      ptx::cmad_mod<PR, 1>(ev, m);
      ptx::cmad_mod<PR, 0>(od, m);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(ev[N - 1]));
      ptx::cmad_mod<PR, 1>(fv, m2);
      ptx::cmad_mod<PR, 0>(fd, m2);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(fv[N - 1]));
    }
  }
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r0.l[0]) : "r"(od[1]), "r"(ev[0]));
#pragma unroll
  for (int k = 1; k < N - 1; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r0.l[k]) : "r"(od[k + 1]), "r"(ev[k]));
  asm volatile("addc.u32 %0, %1, 0;" : "=r"(r0.l[N - 1]) : "r"(ev[N - 1]));
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r1.l[0]) : "r"(fd[1]), "r"(fv[0]));
#pragma unroll
  for (int k = 1; k < N - 1; k++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r1.l[k]) : "r"(fd[k + 1]), "r"(fv[k]));
  asm volatile("addc.u32 %0, %1, 0;" : "=r"(r1.l[N - 1]) : "r"(fv[N - 1]));
  cond_sub_mod<PR>(r0.l, 0);
  cond_sub_mod<PR>(r1.l, 0);
#else
  Mont<PR> t0 = mul(a, b), t1 = mul(e, f);
  r0 = t0;
  r1 = t1;
#endif
}

template <class PR>
DKGV_HD Mont<PR> sqr(const Mont<PR>& a) {
  return mul(a, a);
}

// canonical integer (little-endian limbs) <-> Montgomery form
template <class PR>
DKGV_HD Mont<PR> to_mont(const Mont<PR>& raw) {
  Mont<PR> r2;
#pragma unroll
  for (int i = 0; i < PR::N; i++) r2.l[i] = PR::r2(i);
  return mul(raw, r2);
}
template <class PR>
DKGV_HD Mont<PR> from_mont(const Mont<PR>& a) {
  Mont<PR> o = zero<PR>();
  o.l[0] = 1;
  return mul(a, o);
}
// raw < mod ?
template <class PR>
DKGV_HD bool raw_lt_mod(const uint32_t* a) {
  // lexicographic compare from the top limb
  bool lt = false, decided = false;
#pragma unroll
  for (int i = PR::N - 1; i >= 0; i--) {
    uint32_t m = PR::mod(i);
    if (!decided && a[i] != m) {
      lt = a[i] < m;
      decided = true;
    }
  }
  return lt;
}

// a^e for a public exponent given as little-endian 32-bit limbs (MSB-first square & multiply;
// control flow depends only on the exponent, which is a compile-time constant at every call site)
template <class PR, class ExpFn>
DKGV_HD Mont<PR> pow_const(const Mont<PR>& a, ExpFn e, int nlimbs) {
  Mont<PR> r = one<PR>();
  bool started = false;
#pragma unroll 1
  for (int i = nlimbs - 1; i >= 0; i--) {
    uint32_t w = e(i);
#pragma unroll 1
    for (int b = 31; b >= 0; b--) {
      if (started) r = sqr(r);
      if ((w >> b) & 1) {
        r = started ? mul(r, a) : a;
        started = true;
      }
    }
  }
  return r;
}

using Fp = Mont<FpParams>;
using Fr = Mont<FrParams>;

struct ExpPm2 { DKGV_HD uint32_t operator()(int i) const { return consts::P_MINUS_2(i); } };
struct ExpSqrt { DKGV_HD uint32_t operator()(int i) const { return consts::P_PLUS_1_DIV4(i); } };

DKGV_HD Fp fp_inv(const Fp& a) { return pow_const<FpParams>(a, ExpPm2(), 12); }  // 0 -> 0
// candidate square root a^((p+1)/4); caller checks s*s == a
DKGV_HD Fp fp_sqrt_candidate(const Fp& a) { return pow_const<FpParams>(a, ExpSqrt(), 12); }

// Inversion by the binary extended Euclidean algorithm instead of Fermat's a^(p-2) (607 dependent products, ~316 k instructions):
// invariants u = b x, v = c x (mod p) with v odd; each round subtracts the smaller of (u, v) from the larger when u is odd, then strips
// up to 32 factors of two from u at once (count-trailing-zeros) while b is divided by the same power of two modulo p in one
// multiply-add pass (b + m p with m = -b / p mod 2^k).  ~270 rounds of ~130 plain ALU instructions on 381-bit inputs: about 8x
// fewer instructions than the exponentiation and none of them on the multiplier pipe.  Control flow depends on the DATA - fine for
// public values (commitments, verification results); every lane of a warp loops until its own u reaches 1.
// Input and output in Montgomery form: a = x R  ->  x^-1 R;  0 -> 0 (as fp_inv).
DKGV_HD Fp fp_inv_bgcd(const Fp& a) {
  if (is_zero(a)) return a;
  uint32_t u[12], v[12], b[12], c[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    u[i] = a.l[i];
    v[i] = FpParams::mod(i);
    b[i] = i == 0 ? 1u : 0u;
    c[i] = 0;
  }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (;;) {
    if (u[0] & 1u) {
      uint32_t rest = u[0] ^ 1u;
#pragma unroll
      for (int i = 1; i < 12; i++) rest |= u[i];
      if (rest == 0) break;  // u == 1: b = x^-1
      // d = u - v
      uint32_t d[12], nb[12];
      uint64_t br = 0;
#pragma unroll
      for (int i = 0; i < 12; i++) {
        uint64_t w = (uint64_t)u[i] - v[i] - br;
        d[i] = (uint32_t)w;
        br = (w >> 32) & 1;
      }
      const bool lt = br != 0;  // u < v: (u, v) <- (v - u, u), (b, c) <- (c - b, b)
      // nb = (lt ? c - b : b - c) mod p
      uint64_t bb = 0;
#pragma unroll
      for (int i = 0; i < 12; i++) {
        uint32_t hi = lt ? c[i] : b[i], lo = lt ? b[i] : c[i];
        uint64_t w = (uint64_t)hi - lo - bb;
        nb[i] = (uint32_t)w;
        bb = (w >> 32) & 1;
      }
      uint32_t mask = bb ? 0xffffffffu : 0u;
      uint64_t cy = 0;
#pragma unroll
      for (int i = 0; i < 12; i++) {
        cy += (uint64_t)nb[i] + (FpParams::mod(i) & mask);
        nb[i] = (uint32_t)cy;
        cy >>= 32;
      }
      // u <- |d|, v <- min(u, v)
      uint64_t ng = lt ? 1 : 0;
#pragma unroll
      for (int i = 0; i < 12; i++) {
        uint32_t old_u = u[i];
        uint64_t w = (uint64_t)(lt ? ~d[i] : d[i]) + ng;
        u[i] = (uint32_t)w;
        ng = lt ? (w >> 32) : 0;
        v[i] = lt ? old_u : v[i];
        c[i] = lt ? b[i] : c[i];
        b[i] = nb[i];
      }
    }
    // u is even here: strip k <= 32 factors of two; b <- b / 2^k mod p
    {
      uint32_t k = 32;
      if (u[0]) {
        k = 0;
        uint32_t t = u[0];
        while (!(t & 1u)) {
          t >>= 1;
          k++;
        }
      }
#if defined(__CUDA_ARCH__)
      if (u[0]) k = (uint32_t)(__ffs((int)u[0]) - 1);
#endif
      uint32_t m = b[0] * FpParams::INV;
      if (k < 32) m &= (1u << k) - 1u;
      uint32_t t13[13];
      uint64_t cy = 0;
#pragma unroll
      for (int i = 0; i < 12; i++) {
        uint64_t w = (uint64_t)m * FpParams::mod(i) + b[i] + cy;
        t13[i] = (uint32_t)w;
        cy = w >> 32;
      }
      t13[12] = (uint32_t)cy;
      if (k == 32) {
#pragma unroll
        for (int i = 0; i < 12; i++) {
          b[i] = t13[i + 1];
          u[i] = i < 11 ? u[i + 1] : 0u;
        }
      } else if (k) {
#pragma unroll
        for (int i = 0; i < 12; i++) {
          b[i] = (t13[i] >> k) | (t13[i + 1] << (32 - k));
          u[i] = (u[i] >> k) | ((i < 11 ? u[i + 1] : 0u) << (32 - k));
        }
      }
    }
  }
  Fp y, r2;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    y.l[i] = b[i];
    r2.l[i] = FpParams::r2(i);
  }
  return mul(mul(y, r2), r2);  // plain inverse (x R)^-1 = x^-1 R^-1  ->  x^-1 R
}

// lexicographically-largest test on a Montgomery-form y: canonical(y) > (p-1)/2
DKGV_HD bool fp_lex_largest(const Fp& y_mont) {
  Fp c = from_mont(y_mont);
  bool gt = false, decided = false;
#pragma unroll
  for (int i = 11; i >= 0; i--) {
    uint32_t h = consts::P_MINUS_1_HALF(i);
    if (!decided && c.l[i] != h) {
      gt = c.l[i] > h;
      decided = true;
    }
  }
  return gt;
}

// 48 big-endian bytes (flag bits already masked off) -> raw limbs
DKGV_HD void fp_raw_from_be48(uint32_t* l, const uint8_t* b) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const uint8_t* q = b + 44 - 4 * i;
    l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
  }
}
DKGV_HD void fp_raw_to_be48(uint8_t* b, const uint32_t* l) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint8_t* q = b + 44 - 4 * i;
    q[0] = (uint8_t)(l[i] >> 24);
    q[1] = (uint8_t)(l[i] >> 16);
    q[2] = (uint8_t)(l[i] >> 8);
    q[3] = (uint8_t)l[i];
  }
}

}  // namespace dkgv
