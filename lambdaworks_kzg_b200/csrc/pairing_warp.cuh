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
// v if keep, else 0 -- limb selects, no branch: the phases below must not diverge (a warp executes BOTH sides of every
// branch its lanes disagree on, and each side is dozens of multi-limb additions on the critical path of every Fp12 product)
LW_INL Fp2 fp2_keep_if(const Fp2& v, bool keep) {
  Fp2 r;
  for (int i = 0; i < 12; i++) { r.c0.l[i] = keep ? v.c0.l[i] : 0u; r.c1.l[i] = keep ? v.c1.l[i] : 0u; }
  return r;
}
// Phase 1a, data-driven: Karatsuba product L = 6 k + m multiplies the sum of the coefficients x_i, i in sel(m), of
// the c0 half (k = 0), of the c1 half (k = 1) or of both (k = 2): at most four terms.  Every lane runs the same code --
// four loads (unused terms read slot 0 and are zeroed), a two-level sum, c0 + c1 -- and the A-side (lanes 0..17, from
// a) and the B-side (lanes 32..49, from b) run on different warps instead of one after the other.
LW_COLD void wfp12_mul_phase1a(WarpScratch& sc, const WarpFp12& a, const WarpFp12& b, int lane) {
  const int L = lane & 31, side = lane >> 5;
  if (L >= 18 || side > 1) return;
  const WarpFp12& src = side ? b : a;
  const int k = L / 6, m = L % 6;
  const int sel = (m == 0) ? 1 : (m == 1) ? 2 : (m == 2) ? 4 : (m == 3) ? 6 : (m == 4) ? 3 : 5;  // subset of {x0,x1,x2}
  const int i0 = (sel & 1) ? 0 : (sel & 2) ? 1 : 2;                    // lowest coefficient of the subset
  const int i1 = (sel == 6) ? 2 : (sel == 3) ? 1 : (sel == 5) ? 2 : -1;  // the other one, if any
  // term t: coefficient index or -1
  const int t0 = (k == 1) ? 3 + i0 : i0;
  const int t1 = (k == 2) ? 3 + i0 : -1;
  const int t2 = (i1 < 0) ? -1 : (k == 1) ? 3 + i1 : i1;
  const int t3 = (i1 < 0 || k != 2) ? -1 : 3 + i1;
  const Fp2 x0 = src.c[t0];
  const Fp2 x1 = fp2_keep_if(src.c[t1 < 0 ? 0 : t1], t1 >= 0);
  const Fp2 x2 = fp2_keep_if(src.c[t2 < 0 ? 0 : t2], t2 >= 0);
  const Fp2 x3 = fp2_keep_if(src.c[t3 < 0 ? 0 : t3], t3 >= 0);
  const Fp2 X = fp2_add(fp2_add(x0, x1), fp2_add(x2, x3));
  Fp* dst = side ? sc.opb : sc.opa;
  dst[3 * L] = X.c0;
  dst[3 * L + 1] = X.c1;
  dst[3 * L + 2] = fp_add(X.c0, X.c1);
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
// ---- phase 2a: lane L < 9 recombines coefficient i = L % 3 of Fp6 product L / 3 from its six Karatsuba pieces v0..v5:
//   i = 0: v0 + xi (v3 - v1 - v2)      i = 1: v4 - v0 - v1 + xi v2      i = 2: v5 + v1 - v0 - v2
// written as ONE formula (P0 + P1) - (N0 + N1) + xi (Q - (M0 + M1)) with unused terms zeroed, so the nine lanes do not
// diverge (three different return paths used to be executed one after the other)
LW_COLD void wfp12_mul_phase2a(WarpScratch& sc, int lane) {
  if (lane >= 9) return;
  const int i = lane % 3;
  const Fp2* v = sc.prod + 6 * (lane / 3);
  const Fp2 P0 = v[i == 0 ? 0 : i == 1 ? 4 : 5];
  const Fp2 P1 = fp2_keep_if(v[1], i == 2);
  const Fp2 N0 = fp2_keep_if(v[0], i != 0);
  const Fp2 N1 = fp2_keep_if(v[i == 1 ? 1 : 2], i != 0);
  const Fp2 Q = fp2_keep_if(v[i == 0 ? 3 : 2], i != 2);
  const Fp2 M0 = fp2_keep_if(v[1], i == 0);
  const Fp2 M1 = fp2_keep_if(v[2], i == 0);
  const Fp2 s1 = fp2_sub(fp2_add(P0, P1), fp2_add(N0, N1));
  const Fp2 s2 = fp2_sub(Q, fp2_add(M0, M1));
  sc.coef[lane] = fp2_add(s1, fp2_mul_xi(s2));
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
