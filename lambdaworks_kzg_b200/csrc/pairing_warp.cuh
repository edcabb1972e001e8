// Warp-cooperative optimal-ate pairing check.
//
// pairing.cuh runs the whole pairing in one thread: ~20 k dependent Fp
// multiplications = ~30 ms on a B200 (one warp cannot issue IMAD.WIDE faster
// than its scheduler's share of the pipe).  Here one warp shares the work at the
// granularity that needs no intra-multiplication communication: an Fp12 product
// is 18 independent Fp2 products (Karatsuba over Fp6 over Fp12), so 18 lanes
// compute one Fp2 product each (3 Fp multiplications deep instead of 54), and 6
// lanes recombine the 6 output coefficients.  Fp12 values live in shared memory
// as 6 Fp2 coefficients in the memory order of struct Fp12
// (c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2).
//
// Every step is written as a per-lane phase function taking the lane id, so the
// host emulation (tests/host_emul) can run the phases lane by lane and check them
// against the single-thread implementation and the Python oracle.
#pragma once
#include "pairing.cuh"

namespace lw {

// The cooperating group is one block of LW_PAIR_LANES = 64 threads (two warps): an Fp12 product is 54 independent Fp
// products, and with 64 lanes they are ONE multiplication deep.  Every function below must be called by all 64
// threads of the block with the same arguments (the phases are separated by block barriers).
constexpr int LW_PAIR_LANES = 64;

struct WarpFp12 {
  Fp2 c[6];
};
struct WarpScratch {
  Fp2 prod[18];
  Fp2 coef[9];   // coefficient i of the k-th Fp6 product at coef[3 k + i]
  // the 18 Fp2 products as 54 Fp products (Karatsuba: a0 b0, a1 b1, (a0 + a1)(b0 + b1)): operands and results
  Fp opa[54], opb[54], pr[54];
};

// ---- phase 1 of dst = a * b: lane L < 18 prepares the operands of one Karatsuba product (1a), 54 lanes do one of
// the 54 Fp multiplications behind the 18 Fp2 products each -- one multiplication deep instead of three (1b) -- and
// lane L < 18 assembles its Fp2 product (1c)
LW_COLD void wfp12_mul_phase1a(WarpScratch& sc, const WarpFp12& a, const WarpFp12& b, int lane) {
  if (lane >= 18) return;
  const int k = lane / 6, m = lane % 6;
  const int sel = (m == 0) ? 1 : (m == 1) ? 2 : (m == 2) ? 4 : (m == 3) ? 6 : (m == 4) ? 3 : 5;  // subset of {x0,x1,x2}
  // A, B = sums of the selected coefficients; the first term is copied, not added to zero (the additions are on
  // the critical path of every Fp12 product)
  Fp2 A = fp2_zero(), B = fp2_zero();
  bool first = true;
  for (int i = 0; i < 3; i++) {
    if (!((sel >> i) & 1)) continue;
    if (k == 0 || k == 2) {
      if (first) { A = a.c[i]; B = b.c[i]; first = false; }
      else { A = fp2_add(A, a.c[i]); B = fp2_add(B, b.c[i]); }
    }
    if (k == 1 || k == 2) {
      if (first) { A = a.c[3 + i]; B = b.c[3 + i]; first = false; }
      else { A = fp2_add(A, a.c[3 + i]); B = fp2_add(B, b.c[3 + i]); }
    }
  }
  sc.opa[3 * lane] = A.c0;     sc.opb[3 * lane] = B.c0;
  sc.opa[3 * lane + 1] = A.c1; sc.opb[3 * lane + 1] = B.c1;
  sc.opa[3 * lane + 2] = fp_add(A.c0, A.c1);
  sc.opb[3 * lane + 2] = fp_add(B.c0, B.c1);
}
LW_COLD void wfp12_mul_phase1b(WarpScratch& sc, int lane) {
  for (int idx = lane; idx < 54; idx += LW_PAIR_LANES) {
    Fp t;
    fp_mul_ni(t, sc.opa[idx], sc.opb[idx]);
    sc.pr[idx] = t;
  }
}
LW_COLD void wfp12_mul_phase1c(WarpScratch& sc, int lane) {
  if (lane >= 18) return;
  const Fp t0 = sc.pr[3 * lane], t1 = sc.pr[3 * lane + 1], t2 = sc.pr[3 * lane + 2];
  Fp2 r;   // as fp2_mul
  r.c0 = fp_sub(t0, t1);
  r.c1 = fp_sub(fp_sub(t2, t0), t1);
  sc.prod[lane] = r;
}
// coefficient i of the k-th Fp6 product from its six Karatsuba pieces
LW_COLD Fp2 wfp6_coeff(const WarpScratch& sc, int k, int i) {
  const Fp2* v = sc.prod + 6 * k;
  if (i == 0) return fp2_add(v[0], fp2_mul_xi(fp2_sub(fp2_sub(v[3], v[1]), v[2])));
  if (i == 1) return fp2_add(fp2_sub(fp2_sub(v[4], v[0]), v[1]), fp2_mul_xi(v[2]));
  return fp2_add(fp2_sub(fp2_sub(v[5], v[0]), v[2]), v[1]);
}
// ---- phase 2a: lane L < 9 recombines coefficient L % 3 of Fp6 product L / 3 (nine lanes instead of each of the six
// output lanes recomputing up to three of them)
LW_COLD void wfp12_mul_phase2a(WarpScratch& sc, int lane) {
  if (lane >= 9) return;
  sc.coef[lane] = wfp6_coeff(sc, lane / 3, lane % 3);
}
// ---- phase 2b: lane o < 6 writes output coefficient o
LW_COLD void wfp12_mul_phase2b(WarpFp12& dst, const WarpScratch& sc, int lane) {
  if (lane >= 6) return;
  Fp2 r;
  if (lane < 3) {
    // r.c0 = T0 + v * T1,  v * (c0, c1, c2) = (xi c2, c0, c1)
    Fp2 t1 = sc.coef[3 + (lane + 2) % 3];
    if (lane == 0) t1 = fp2_mul_xi(t1);
    r = fp2_add(sc.coef[lane], t1);
  } else {
    int i = lane - 3;  // r.c1 = T2 - T0 - T1
    r = fp2_sub(fp2_sub(sc.coef[6 + i], sc.coef[i]), sc.coef[3 + i]);
  }
  dst.c[lane] = r;
}

// conjugation (p^6 Frobenius): negate the w-odd half
LW_COLD void wfp12_conj_lane(WarpFp12& dst, const WarpFp12& a, int lane) {
  if (lane >= 6) return;
  dst.c[lane] = (lane < 3) ? a.c[lane] : fp2_neg(a.c[lane]);
}
// p-power Frobenius: coefficient of w^e -> conj(.) * xi^(e (p-1)/6); memory slot s holds w^(2 s) (s < 3) or w^(2 (s-3) + 1)
LW_COLD void wfp12_frobenius_lane(WarpFp12& dst, const WarpFp12& a, int lane) {
  if (lane >= 6) return;
  const int e = (lane < 3) ? 2 * lane : 2 * (lane - 3) + 1;
  Fp2 v = fp2_conj(a.c[lane]);
  dst.c[lane] = (e == 0) ? v : fp2_mul(v, frob_gamma(e));
}
LW_COLD void wfp12_copy_lane(WarpFp12& dst, const WarpFp12& a, int lane) {
  if (lane < 6) dst.c[lane] = a.c[lane];
}
// the sparse line element  mu + (-lambda xP) v + yP v w  as a full Fp12 (slots 0, 1 and 4)
LW_COLD void wfp12_line_lane(WarpFp12& dst, const G2Line& ln, const G1Affine& p, int lane) {
  if (lane >= 6) return;
  Fp2 r = fp2_zero();
  if (lane == 0) r = ln.mu;
  if (lane == 1) r = fp2_neg(fp2_mul_fp(ln.lambda, p.x));
  if (lane == 4) { r.c0 = p.y; r.c1 = fp_zero(); }
  dst.c[lane] = r;
}

#if defined(LWKZG_HOST_EMUL)
#define LW_WARP_SYNC()
#define LW_FOR_LANES(lane) for (int lane = 0; lane < LW_PAIR_LANES; lane++)
#else
#define LW_WARP_SYNC() __syncthreads()
#define LW_FOR_LANES(lane) for (int lane = (int)threadIdx.x, _once = 1; _once; _once = 0)
#endif

// dst = a * b (dst may alias a and/or b)
LW_COLD void wfp12_mul(WarpFp12& dst, const WarpFp12& a, const WarpFp12& b, WarpScratch& sc) {
  LW_FOR_LANES(lane) wfp12_mul_phase1a(sc, a, b, lane);
  LW_WARP_SYNC();
  LW_FOR_LANES(lane) wfp12_mul_phase1b(sc, lane);
  LW_WARP_SYNC();
  LW_FOR_LANES(lane) wfp12_mul_phase1c(sc, lane);
  LW_WARP_SYNC();
  LW_FOR_LANES(lane) wfp12_mul_phase2a(sc, lane);
  LW_WARP_SYNC();
  LW_FOR_LANES(lane) wfp12_mul_phase2b(dst, sc, lane);
  LW_WARP_SYNC();
}
LW_COLD void wfp12_conj(WarpFp12& dst, const WarpFp12& a) {
  LW_FOR_LANES(lane) wfp12_conj_lane(dst, a, lane);
  LW_WARP_SYNC();
}
LW_COLD void wfp12_frobenius(WarpFp12& dst, const WarpFp12& a) {
  LW_FOR_LANES(lane) wfp12_frobenius_lane(dst, a, lane);
  LW_WARP_SYNC();
}
LW_COLD void wfp12_copy(WarpFp12& dst, const WarpFp12& a) {
  LW_FOR_LANES(lane) wfp12_copy_lane(dst, a, lane);
  LW_WARP_SYNC();
}
// single-lane pieces (inherently sequential or tiny): run by lane 0
LW_COLD void wfp12_set_one(WarpFp12& dst) {
  LW_FOR_LANES(lane) if (lane == 0) { Fp12 o = fp12_one(); dst = *reinterpret_cast<WarpFp12*>(&o); }
  LW_WARP_SYNC();
}
LW_COLD void wfp12_inv(WarpFp12& dst, const WarpFp12& a) {
  LW_FOR_LANES(lane) if (lane == 0) {
    Fp12 t = fp12_inv(*reinterpret_cast<const Fp12*>(&a));
    dst = *reinterpret_cast<WarpFp12*>(&t);
  }
  LW_WARP_SYNC();
}

// acc = g^|x|, conjugated (x < 0); g unitary
LW_COLD void wfp12_pow_x(WarpFp12& acc, const WarpFp12& g, WarpFp12& tmp, WarpScratch& sc) {
  wfp12_copy(tmp, g);  // tmp = g (acc may alias g)
  wfp12_copy(acc, tmp);
  for (int bit = 62; bit >= 0; bit--) {
    wfp12_mul(acc, acc, acc, sc);
    if ((k::BLS_X_ABS >> bit) & 1ull) wfp12_mul(acc, acc, tmp, sc);
  }
  wfp12_conj(acc, acc);
}

struct WarpPairingMem {
  WarpFp12 f, g, a, b, c, t, u;
  WarpScratch sc;
};

// product over pairs of Miller loops, then the final exponentiation (same
// exponent 3 (p^12 - 1)/r as pairing.cuh); result left in m.f.  All 32 lanes of the
// warp must call this with the same arguments; `m` is warp-shared memory.
LW_COLD void warp_pairing_product(WarpPairingMem& m, const G1Affine* ps, const G2Prepared* const* qs, int npairs) {
  wfp12_set_one(m.f);
  int n = 0;
  for (int bit = 62; bit >= 0; bit--) {
    wfp12_mul(m.f, m.f, m.f, m.sc);
    for (int i = 0; i < npairs; i++) {
      if (g1a_is_inf(ps[i]) || qs[i]->infinity) continue;
      LW_FOR_LANES(lane) wfp12_line_lane(m.t, qs[i]->line[n], ps[i], lane);
      LW_WARP_SYNC();
      wfp12_mul(m.f, m.f, m.t, m.sc);
    }
    n++;
    if ((k::BLS_X_ABS >> bit) & 1ull) {
      for (int i = 0; i < npairs; i++) {
        if (g1a_is_inf(ps[i]) || qs[i]->infinity) continue;
        LW_FOR_LANES(lane) wfp12_line_lane(m.t, qs[i]->line[n], ps[i], lane);
        LW_WARP_SYNC();
        wfp12_mul(m.f, m.f, m.t, m.sc);
      }
      n++;
    }
  }
  wfp12_conj(m.f, m.f);  // x < 0
  // ---- easy part: g = f^((p^6-1)(p^2+1))
  wfp12_inv(m.t, m.f);
  wfp12_conj(m.g, m.f);
  wfp12_mul(m.g, m.g, m.t, m.sc);
  wfp12_frobenius(m.t, m.g);
  wfp12_frobenius(m.t, m.t);
  wfp12_mul(m.g, m.t, m.g, m.sc);
  // ---- hard part: (x-1)^2 (x+p) (x^2+p^2-1) + 3
  wfp12_pow_x(m.a, m.g, m.u, m.sc);
  wfp12_conj(m.t, m.g);
  wfp12_mul(m.a, m.a, m.t, m.sc);            // g^(x-1)
  wfp12_pow_x(m.b, m.a, m.u, m.sc);
  wfp12_conj(m.t, m.a);
  wfp12_mul(m.a, m.b, m.t, m.sc);            // g^((x-1)^2)
  wfp12_pow_x(m.b, m.a, m.u, m.sc);
  wfp12_frobenius(m.t, m.a);
  wfp12_mul(m.b, m.b, m.t, m.sc);            // a^(x+p)
  wfp12_pow_x(m.c, m.b, m.u, m.sc);
  wfp12_pow_x(m.c, m.c, m.u, m.sc);          // b^(x^2)
  wfp12_frobenius(m.t, m.b);
  wfp12_frobenius(m.t, m.t);
  wfp12_mul(m.c, m.c, m.t, m.sc);
  wfp12_conj(m.t, m.b);
  wfp12_mul(m.c, m.c, m.t, m.sc);            // b^(x^2+p^2-1)
  wfp12_mul(m.t, m.g, m.g, m.sc);
  wfp12_mul(m.t, m.t, m.g, m.sc);            // g^3
  wfp12_mul(m.f, m.c, m.t, m.sc);
}

LW_COLD bool warp_fp12_is_one(const WarpFp12& a) { return fp12_is_one(*reinterpret_cast<const Fp12*>(&a)); }

}  // namespace lw
