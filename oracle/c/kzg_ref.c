/*
 * kzg_ref.c -- CPU restatement of lambdaworks_kzg's commit / proof path in C.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call this; it is never
 * linked into liblwkzg_b200.so.  "Restatement, not the Rust binary": the Rust
 * toolchain and the un-vendored lambdaworks-math / lambdaworks-crypto crates
 * (Cargo.toml:15-16, no rev pin) are not available, so this file follows the
 * reference's algorithmic structure step for step:
 *
 *   blob_to_polynomial            src/utils.rs:27-41   BE parse, reduce mod r, trim
 *   kzgsettings_to_structured_reference_string   src/srs.rs:258-280   PER-CALL
 *        re-hydration of 4096 G1 points (Montgomery conversion + curve check)
 *   KZG::commit = msm(coeffs, srs[..len])          src/lib.rs:269-270; App. D.2/D.3:
 *        sequential Pippenger, UNSIGNED window of floor(0.8 log2 n) bits (9 for
 *        n = 4096), 2^w - 1 buckets, homogeneous projective coordinates
 *   Polynomial::evaluate (Horner), KZG::open (Ruffini)   src/lib.rs:320-329, 389-394
 *   compress / decompress + naive [r]P subgroup check     src/compression.rs:22-103
 *   compute_challenge (SHA-256)                            src/utils.rs:120-154
 *
 * Arithmetic: 6 x u64 (Fp) / 4 x u64 (Fr) Montgomery CIOS with unsigned
 * __int128 -- a different limb width, coordinate system and MSM algorithm from
 * the CUDA path it checks.
 *
 * Parity pins: tests/test_oracle.py checks this file against the reference's
 * own tests (tests/lib_test.rs, src/compression.rs unit tests) and against
 * oracle/py (itself pinned by the 208 c-kzg YAML vectors).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>
#include "consts.h"

typedef unsigned __int128 u128;
typedef struct { uint64_t l[6]; } fp;
typedef struct { uint64_t l[4]; } fr;
typedef struct { fp x, y, z; } g1p; /* homogeneous projective, neutral = (0:1:0) */

#define NPTS 4096
#define BLOB_BYTES (4096 * 32)

/* ------------------------------------------------------------------ bigint */
static int ge_n(const uint64_t *a, const uint64_t *b, int n) {
  for (int i = n - 1; i >= 0; i--) {
    if (a[i] > b[i]) return 1;
    if (a[i] < b[i]) return 0;
  }
  return 1;
}
static void sub_n(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
  uint64_t borrow = 0;
  for (int i = 0; i < n; i++) {
    u128 t = (u128)a[i] - b[i] - borrow;
    r[i] = (uint64_t)t;
    borrow = (uint64_t)(t >> 64) & 1;
  }
}
static uint64_t add_n(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
  uint64_t c = 0;
  for (int i = 0; i < n; i++) {
    u128 t = (u128)a[i] + b[i] + c;
    r[i] = (uint64_t)t;
    c = (uint64_t)(t >> 64);
  }
  return c;
}
/* Montgomery CIOS: r = a b / 2^(64 n) mod m */
static inline __attribute__((always_inline)) void mont_mul_n(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m, uint64_t inv, int n) {
  uint64_t t[8] = {0};
  for (int i = 0; i < n; i++) {
    u128 c = 0;
    for (int j = 0; j < n; j++) {
      c += (u128)a[j] * b[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[n];
    t[n] = (uint64_t)c;
    t[n + 1] = (uint64_t)(c >> 64);
    uint64_t q = t[0] * inv;
    c = (u128)q * m[0] + t[0];
    c >>= 64;
    for (int j = 1; j < n; j++) {
      c += (u128)q * m[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[n];
    t[n - 1] = (uint64_t)c;
    t[n] = t[n + 1] + (uint64_t)(c >> 64);
  }
  if (t[n] || ge_n(t, m, n)) sub_n(r, t, m, n);
  else memcpy(r, t, 8 * n);
}

/* ------------------------------------------------------------------ Fp */
static void fp_mul(fp *r, const fp *a, const fp *b) { mont_mul_n(r->l, a->l, b->l, FP_P, FP_INV, 6); }
static void fp_sqr(fp *r, const fp *a) { fp_mul(r, a, a); }
static void fp_add(fp *r, const fp *a, const fp *b) {
  uint64_t c = add_n(r->l, a->l, b->l, 6);
  if (c || ge_n(r->l, FP_P, 6)) sub_n(r->l, r->l, FP_P, 6);
}
static void fp_sub(fp *r, const fp *a, const fp *b) {
  if (ge_n(a->l, b->l, 6)) sub_n(r->l, a->l, b->l, 6);
  else { uint64_t t[6]; sub_n(t, b->l, a->l, 6); sub_n(r->l, FP_P, t, 6); }
}
static int fp_is_zero(const fp *a) { uint64_t o = 0; for (int i = 0; i < 6; i++) o |= a->l[i]; return o == 0; }
static int fp_eq(const fp *a, const fp *b) { return memcmp(a, b, sizeof(fp)) == 0; }
static void fp_neg(fp *r, const fp *a) { if (fp_is_zero(a)) *r = *a; else sub_n(r->l, FP_P, a->l, 6); }
static void fp_set_one(fp *r) { memcpy(r->l, FP_ONE, 48); }
static void fp_to_mont(fp *r, const fp *a) { fp r2; memcpy(r2.l, FP_R2, 48); fp_mul(r, a, &r2); }
static void fp_from_mont(fp *r, const fp *a) { fp one = {{1, 0, 0, 0, 0, 0}}; fp_mul(r, a, &one); }
static void fp_pow(fp *r, const fp *a, const uint64_t *e, int n) {
  fp acc; fp_set_one(&acc);
  for (int i = n * 64 - 1; i >= 0; i--) {
    fp_sqr(&acc, &acc);
    if ((e[i / 64] >> (i % 64)) & 1) fp_mul(&acc, &acc, a);
  }
  *r = acc;
}
static void fp_inv(fp *r, const fp *a) { fp_pow(r, a, FP_PM2, 6); }
/* from_bytes_be: 48 bytes, silently reduced mod p (App. D.1) */
static void fp_from_be(fp *r, const uint8_t *b) {
  fp t;
  for (int i = 0; i < 6; i++) {
    uint64_t w = 0;
    for (int k = 0; k < 8; k++) w = (w << 8) | b[(5 - i) * 8 + k];
    t.l[i] = w;
  }
  while (ge_n(t.l, FP_P, 6)) sub_n(t.l, t.l, FP_P, 6);
  fp_to_mont(r, &t);
}
static void fp_to_be(uint8_t *b, const fp *a_mont) {
  fp c; fp_from_mont(&c, a_mont);
  for (int i = 0; i < 6; i++) for (int k = 0; k < 8; k++) b[(5 - i) * 8 + k] = (uint8_t)(c.l[i] >> (56 - 8 * k));
}

/* ------------------------------------------------------------------ Fr */
static void fr_mul(fr *r, const fr *a, const fr *b) { mont_mul_n(r->l, a->l, b->l, FR_P, FR_INV, 4); }
static void fr_add(fr *r, const fr *a, const fr *b) {
  uint64_t c = add_n(r->l, a->l, b->l, 4);
  if (c || ge_n(r->l, FR_P, 4)) sub_n(r->l, r->l, FR_P, 4);
}
static void fr_sub(fr *r, const fr *a, const fr *b) {
  if (ge_n(a->l, b->l, 4)) sub_n(r->l, a->l, b->l, 4);
  else { uint64_t t[4]; sub_n(t, b->l, a->l, 4); sub_n(r->l, FR_P, t, 4); }
}
static int fr_is_zero(const fr *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static void fr_to_mont(fr *r, const fr *a) { fr r2; memcpy(r2.l, FR_R2, 32); fr_mul(r, a, &r2); }
static void fr_from_mont(fr *r, const fr *a) { fr one = {{1, 0, 0, 0}}; fr_mul(r, a, &one); }
/* FE::from_bytes_be: 32 bytes BE -> Montgomery, reduced mod r */
static void fr_from_be(fr *r, const uint8_t *b) {
  fr t;
  for (int i = 0; i < 4; i++) {
    uint64_t w = 0;
    for (int k = 0; k < 8; k++) w = (w << 8) | b[(3 - i) * 8 + k];
    t.l[i] = w;
  }
  while (ge_n(t.l, FR_P, 4)) sub_n(t.l, t.l, FR_P, 4);
  fr_to_mont(r, &t);
}
static void fr_to_be(uint8_t *b, const fr *a_mont) {
  fr c; fr_from_mont(&c, a_mont);
  for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) b[(3 - i) * 8 + k] = (uint8_t)(c.l[i] >> (56 - 8 * k));
}

/* ------------------------------------------------------------------ G1, homogeneous projective */
static void g1_neutral(g1p *r) { memset(r, 0, sizeof(*r)); fp_set_one(&r->y); }
static int g1_is_neutral(const g1p *p) { return fp_is_zero(&p->z); }
static void g1_double(g1p *r, const g1p *p) {
  /* a = 0: w = 3 x^2, s = y z, b = x y s, h = w^2 - 8 b */
  if (g1_is_neutral(p) || fp_is_zero(&p->y)) { g1_neutral(r); return; }
  fp xx, w, s, ss, sss, ys, b, h, t, t2, yy;
  fp_sqr(&xx, &p->x);
  fp_add(&w, &xx, &xx); fp_add(&w, &w, &xx);
  fp_mul(&s, &p->y, &p->z);
  fp_sqr(&ss, &s); fp_mul(&sss, &ss, &s);
  fp_mul(&ys, &p->y, &s);
  fp_mul(&b, &p->x, &ys);
  fp_sqr(&h, &w);
  fp_add(&t, &b, &b); fp_add(&t, &t, &t); fp_add(&t2, &t, &t); /* t = 4b, t2 = 8b */
  fp_sub(&h, &h, &t2);
  g1p o;
  fp_mul(&o.x, &h, &s); fp_add(&o.x, &o.x, &o.x);              /* 2 h s */
  fp_sub(&t, &t, &h);                                          /* 4b - h */
  fp_mul(&t, &w, &t);
  fp_sqr(&yy, &ys); fp_add(&yy, &yy, &yy); fp_add(&yy, &yy, &yy); fp_add(&yy, &yy, &yy); /* 8 (y s)^2 */
  fp_sub(&o.y, &t, &yy);
  fp_add(&o.z, &sss, &sss); fp_add(&o.z, &o.z, &o.z); fp_add(&o.z, &o.z, &o.z);        /* 8 s^3 */
  *r = o;
}
static void g1_add(g1p *r, const g1p *p, const g1p *q) {
  if (g1_is_neutral(p)) { *r = *q; return; }
  if (g1_is_neutral(q)) { *r = *p; return; }
  fp u1, u2, v1, v2;
  fp_mul(&u1, &q->y, &p->z); fp_mul(&u2, &p->y, &q->z);
  fp_mul(&v1, &q->x, &p->z); fp_mul(&v2, &p->x, &q->z);
  if (fp_eq(&v1, &v2)) {
    if (!fp_eq(&u1, &u2) || fp_is_zero(&p->y)) { g1_neutral(r); return; }
    g1_double(r, p);
    return;
  }
  fp u, v, w, vv, vvv, uu, a, t;
  fp_sub(&u, &u1, &u2); fp_sub(&v, &v1, &v2);
  fp_mul(&w, &p->z, &q->z);
  fp_sqr(&vv, &v); fp_mul(&vvv, &vv, &v);
  fp_sqr(&uu, &u);
  fp_mul(&a, &uu, &w); fp_sub(&a, &a, &vvv);
  fp_mul(&t, &vv, &v2); fp_sub(&a, &a, &t); fp_sub(&a, &a, &t);   /* a = u^2 w - v^3 - 2 v^2 v2 */
  g1p o;
  fp_mul(&o.x, &v, &a);
  fp_sub(&t, &t, &a); fp_mul(&t, &u, &t);                         /* u (v^2 v2 - a) */
  fp t3; fp_mul(&t3, &vvv, &u2);
  fp_sub(&o.y, &t, &t3);
  fp_mul(&o.z, &vvv, &w);
  *r = o;
}
static void g1_neg(g1p *r, const g1p *p) { *r = *p; fp_neg(&r->y, &p->y); }
/* operate_with_self: LSB-first double-and-add over a little-endian u64 scalar */
static void g1_mul_u64s(g1p *r, const g1p *p, const uint64_t *k, int n) {
  g1p acc, base = *p;
  g1_neutral(&acc);
  for (int i = 0; i < n * 64; i++) {
    if ((k[i / 64] >> (i % 64)) & 1) g1_add(&acc, &acc, &base);
    g1_double(&base, &base);
  }
  *r = acc;
}
static int g1_on_curve_affine(const fp *x, const fp *y) {
  fp l, r3, four = {{4, 0, 0, 0, 0, 0}}, b;
  fp_to_mont(&b, &four);
  fp_sqr(&l, y);
  fp_sqr(&r3, x); fp_mul(&r3, &r3, x); fp_add(&r3, &r3, &b);
  return fp_eq(&l, &r3);
}
static int g1_to_affine(fp *x, fp *y, const g1p *p) {
  if (g1_is_neutral(p)) return 0;
  fp zi; fp_inv(&zi, &p->z);
  fp_mul(x, &p->x, &zi); fp_mul(y, &p->y, &zi);
  return 1;
}
static int g1_eq(const g1p *a, const g1p *b) {
  int na = g1_is_neutral(a), nb = g1_is_neutral(b);
  if (na || nb) return na && nb;
  fp l, r;
  fp_mul(&l, &a->x, &b->z); fp_mul(&r, &b->x, &a->z);
  if (!fp_eq(&l, &r)) return 0;
  fp_mul(&l, &a->y, &b->z); fp_mul(&r, &b->y, &a->z);
  return fp_eq(&l, &r);
}

/* ------------------------------------------------------------------ codecs (src/compression.rs) */
static void compress_g1(uint8_t out[48], const g1p *p) {
  fp x, y;
  if (!g1_to_affine(&x, &y, p)) { memset(out, 0, 48); out[0] = 0xC0; return; }
  fp_to_be(out, &x);
  out[0] |= 0x80;
  fp yc, ny, nyc;
  fp_from_mont(&yc, &y); fp_neg(&ny, &y); fp_from_mont(&nyc, &ny);
  if (!ge_n(nyc.l, yc.l, 6)) out[0] |= 0x20;  /* (-y) < y */
}
static int in_subgroup(const g1p *p) {
  g1p t; g1_mul_u64s(&t, p, FR_P, 4);  /* [r]P == O, compression.rs:22-27 */
  return g1_is_neutral(&t);
}
static int decompress_g1(g1p *out, const uint8_t in[48]) {
  uint8_t b0 = in[0];
  if (!(b0 & 0x80)) return 0;
  if (b0 & 0x40) { g1_neutral(out); return 1; }
  uint8_t tmp[48]; memcpy(tmp, in, 48); tmp[0] = b0 & 0x1F;
  fp x, y2, y, four = {{4, 0, 0, 0, 0, 0}}, b, chk;
  fp_from_be(&x, tmp);
  fp_to_mont(&b, &four);
  fp_sqr(&y2, &x); fp_mul(&y2, &y2, &x); fp_add(&y2, &y2, &b);
  fp_pow(&y, &y2, FP_SQRT_EXP, 6);
  fp_sqr(&chk, &y);
  if (!fp_eq(&chk, &y2)) return 0;
  fp yc, ny, nyc;
  fp_from_mont(&yc, &y); fp_neg(&ny, &y); fp_from_mont(&nyc, &ny);
  int y_is_larger = !ge_n(nyc.l, yc.l, 6);
  if (((b0 & 0x20) != 0) != y_is_larger) y = ny;
  out->x = x; out->y = y; fp_set_one(&out->z);
  return in_subgroup(out);
}

/* ------------------------------------------------------------------ SHA-256 */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
static void sha_block(uint32_t h[8], const uint8_t *p) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = ROR(w[i - 15], 7) ^ ROR(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = ROR(w[i - 2], 17) ^ ROR(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t t1 = hh + (ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
    uint32_t t2 = (ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
static void sha256(uint8_t out[32], const uint8_t *msg, size_t len) {
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  size_t off = 0;
  for (; off + 64 <= len; off += 64) sha_block(h, msg + off);
  uint8_t tail[128] = {0};
  size_t rem = len - off;
  memcpy(tail, msg + off, rem);
  tail[rem] = 0x80;
  size_t tl = rem + 9 <= 64 ? 64 : 128;
  uint64_t bits = (uint64_t)len * 8;
  for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
  for (size_t o = 0; o < tl; o += 64) sha_block(h, tail + o);
  for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
}

/* ------------------------------------------------------------------ KZGSettings-shaped global setup */
/* g1_values exactly as the reference stores them: canonical integers, u64 limbs
 * most-significant first, x | y | z (144 bytes per point) -- src/srs.rs:131-153 */
static uint64_t *g_g1_values = NULL;

static void store_blst_fp(uint64_t *dst, const fp *a_mont) {
  fp c; fp_from_mont(&c, a_mont);
  for (int k = 0; k < 6; k++) dst[k] = c.l[5 - k];
}

/* tiny pthread fan-out: each worker pulls item indices from a shared counter */
static int hw_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
static void run_parallel(void *(*fn)(void *), void *arg, int nthreads) {
  if (nthreads <= 0) nthreads = hw_threads();
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, fn, arg);
  fn(arg);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}
struct load_job { const uint8_t *in; uint64_t *vals; int next; int bad; };
static void *load_worker(void *a) {
  struct load_job *j = (struct load_job *)a;
  for (;;) {
    int i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    if (i >= NPTS) break;
    g1p p;
    if (!decompress_g1(&p, j->in + 48 * i)) { __atomic_store_n(&j->bad, 1, __ATOMIC_RELAXED); continue; }
    fp x, y;
    uint64_t *d = j->vals + (size_t)i * 18;
    if (g1_to_affine(&x, &y, &p)) { store_blst_fp(d, &x); store_blst_fp(d + 6, &y); }
    d[17] = 1; /* z = [0,0,0,0,0,1] */
  }
  return NULL;
}

int oracle_load_setup_g1(const uint8_t *g1_compressed, int n) {
  /* load_trusted_setup: decompress + validate every point (lib.rs:728-733) */
  if (n != NPTS) return 1;
  uint64_t *vals = (uint64_t *)calloc((size_t)NPTS * 18, 8);
  struct load_job job = {g1_compressed, vals, 0, 0};
  run_parallel(load_worker, &job, 0);
  if (job.bad) { free(vals); return 2; }
  free(g_g1_values);
  g_g1_values = vals;
  return 0;
}
const uint64_t *oracle_g1_values(void) { return g_g1_values; }

/* kzgsettings_to_structured_reference_string (src/srs.rs:258-280, 155-172):
 * rebuild every G1 point from its canonical limbs on EVERY call */
static int rehydrate_srs(g1p *srs) {
  if (!g_g1_values) return 0;
  for (int i = 0; i < NPTS; i++) {
    const uint64_t *d = g_g1_values + (size_t)i * 18;
    uint8_t be[96];
    for (int k = 0; k < 6; k++) for (int b = 0; b < 8; b++) { be[k * 8 + b] = (uint8_t)(d[k] >> (56 - 8 * b)); be[48 + k * 8 + b] = (uint8_t)(d[6 + k] >> (56 - 8 * b)); }
    fp_from_be(&srs[i].x, be);
    fp_from_be(&srs[i].y, be + 48);
    if (!g1_on_curve_affine(&srs[i].x, &srs[i].y)) return 0; /* from_affine */
    fp_set_one(&srs[i].z);
  }
  return 1;
}

/* msm::pippenger::msm (App. D.3): unsigned window w = max(2, floor(0.8 log2 n)) */
static void msm_pippenger(g1p *out, const fr *scalars_canon, const g1p *pts, int n) {
  g1p total; g1_neutral(&total);
  if (n == 0) { *out = total; return; }
  int w = (int)floor(0.8 * log2((double)n));
  if (w < 2) w = 2;
  int nwin = (256 - 1) / w + 1;
  int nb = (1 << w) - 1;
  g1p *buckets = (g1p *)malloc(sizeof(g1p) * nb);
  for (int win = nwin - 1; win >= 0; win--) {
    for (int k = 0; k < w; k++) g1_double(&total, &total);
    for (int b = 0; b < nb; b++) g1_neutral(&buckets[b]);
    for (int i = 0; i < n; i++) {
      int bit = win * w;
      uint64_t d = scalars_canon[i].l[bit / 64] >> (bit % 64);
      if (bit % 64 + w > 64 && bit / 64 + 1 < 4) d |= scalars_canon[i].l[bit / 64 + 1] << (64 - bit % 64);
      d &= (uint64_t)nb;
      if (d) g1_add(&buckets[d - 1], &buckets[d - 1], &pts[i]);
    }
    g1p run, acc;
    g1_neutral(&run); g1_neutral(&acc);
    for (int b = nb - 1; b >= 0; b--) {
      g1_add(&run, &run, &buckets[b]);
      g1_add(&acc, &acc, &run);
    }
    g1_add(&total, &total, &acc);
  }
  free(buckets);
  *out = total;
}

/* blob_to_polynomial: Montgomery coefficients, trailing zeros trimmed */
static int blob_to_poly(fr *coeffs, const uint8_t *blob) {
  int len = 0;
  for (int i = 0; i < NPTS; i++) {
    fr_from_be(&coeffs[i], blob + 32 * i);
    if (!fr_is_zero(&coeffs[i])) len = i + 1;
  }
  return len;
}
static void poly_eval(fr *y, const fr *c, int len, const fr *z) {
  fr acc = {{0, 0, 0, 0}};
  for (int i = len - 1; i >= 0; i--) { fr_mul(&acc, &acc, z); fr_add(&acc, &acc, &c[i]); }
  *y = acc;
}
/* KZG::commit: msm(coeffs.representative(), srs[..len]) */
static void commit_poly(g1p *out, const fr *c_mont, int len, const g1p *srs) {
  fr *canon = (fr *)malloc(sizeof(fr) * (len ? len : 1));
  for (int i = 0; i < len; i++) fr_from_mont(&canon[i], &c_mont[i]);
  msm_pippenger(out, canon, srs, len);
  free(canon);
}
/* KZG::open: (p - y) / (X - z) by Ruffini, then commit */
static void open_poly(g1p *out, const fr *c, int len, const fr *z, const g1p *srs) {
  fr *q = (fr *)malloc(sizeof(fr) * NPTS);
  int qlen = 0;
  if (len > 1) {
    fr acc = {{0, 0, 0, 0}};
    for (int k = len - 1; k >= 1; k--) { fr_mul(&acc, &acc, z); fr_add(&acc, &acc, &c[k]); q[k - 1] = acc; }
    for (int i = 0; i < len - 1; i++) if (!fr_is_zero(&q[i])) qlen = i + 1;
  }
  commit_poly(out, q, qlen, srs);
  free(q);
}
static void challenge(fr *z, const uint8_t *blob, const g1p *commitment) {
  uint8_t *msg = (uint8_t *)malloc(32 + BLOB_BYTES + 48);
  memcpy(msg, "FSBLOBVERIFY_V1_", 16);
  memset(msg + 16, 0, 16);
  msg[17] = 0x10; /* le64(4096) */
  memcpy(msg + 32, blob, BLOB_BYTES);
  compress_g1(msg + 32 + BLOB_BYTES, commitment);
  uint8_t dg[32];
  sha256(dg, msg, 32 + BLOB_BYTES + 48);
  free(msg);
  fr_from_be(z, dg);
}

/* return codes: 0 = C_KZG_OK, 2 = C_KZG_ERROR */
int oracle_blob_to_kzg_commitment(uint8_t out[48], const uint8_t *blob) {
  fr *c = (fr *)malloc(sizeof(fr) * NPTS);
  g1p *srs = (g1p *)malloc(sizeof(g1p) * NPTS);
  int len = blob_to_poly(c, blob);
  int rc = 2;
  if (rehydrate_srs(srs)) {
    g1p cm; commit_poly(&cm, c, len, srs);
    compress_g1(out, &cm);
    rc = 0;
  }
  free(c); free(srs);
  return rc;
}
int oracle_compute_kzg_proof(uint8_t proof[48], uint8_t y_out[32], const uint8_t *blob, const uint8_t z_bytes[32]) {
  fr *c = (fr *)malloc(sizeof(fr) * NPTS);
  g1p *srs = (g1p *)malloc(sizeof(g1p) * NPTS);
  int len = blob_to_poly(c, blob);
  fr z, y;
  fr_from_be(&z, z_bytes);
  poly_eval(&y, c, len, &z);
  int rc = 2;
  if (rehydrate_srs(srs)) {
    g1p pr; open_poly(&pr, c, len, &z, srs);
    compress_g1(proof, &pr);
    fr_to_be(y_out, &y);
    rc = 0;
  }
  free(c); free(srs);
  return rc;
}
int oracle_compute_blob_kzg_proof(uint8_t out[48], const uint8_t *blob, const uint8_t commitment[48]) {
  g1p cm;
  if (!decompress_g1(&cm, commitment)) return 2;
  fr *c = (fr *)malloc(sizeof(fr) * NPTS);
  g1p *srs = (g1p *)malloc(sizeof(g1p) * NPTS);
  int len = blob_to_poly(c, blob);
  fr z, y;
  challenge(&z, blob, &cm);
  poly_eval(&y, c, len, &z);
  int rc = 2;
  if (rehydrate_srs(srs)) {
    g1p pr; open_poly(&pr, c, len, &z, srs);
    compress_g1(out, &pr);
    rc = 0;
  }
  free(c); free(srs);
  return rc;
}
/* n blobs, one blob per thread (the reference is single-threaded inside a call) */
struct batch_job { uint8_t *c, *p; const uint8_t *blobs; int n, next, bad; };
static void *batch_worker(void *a) {
  struct batch_job *j = (struct batch_job *)a;
  for (;;) {
    int i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    if (i >= j->n) break;
    if (oracle_blob_to_kzg_commitment(j->c + 48 * i, j->blobs + (size_t)BLOB_BYTES * i) ||
        oracle_compute_blob_kzg_proof(j->p + 48 * i, j->blobs + (size_t)BLOB_BYTES * i, j->c + 48 * i))
      __atomic_store_n(&j->bad, 1, __ATOMIC_RELAXED);
  }
  return NULL;
}
int oracle_commit_and_prove_batch(uint8_t *commitments, uint8_t *proofs, const uint8_t *blobs, int n, int nthreads) {
  struct batch_job job = {commitments, proofs, blobs, n, 0, 0};
  run_parallel(batch_worker, &job, nthreads);
  return job.bad ? 2 : 0;
}
int oracle_max_threads(void) { return hw_threads(); }
/* helpers used by the tests */
int oracle_g1_decompress_check(const uint8_t in[48], uint8_t recompressed[48]) {
  g1p p;
  if (!decompress_g1(&p, in)) return 0;
  compress_g1(recompressed, &p);
  return 1;
}
void oracle_sha256(uint8_t out[32], const uint8_t *msg, size_t len) { sha256(out, msg, len); }

/* ---- bench.py helpers (the reference arm must not load the product library) ---- */
/* SURVEY 8(d) synthetic blob k: word i = four big-endian u64 of SplitMix64(seed 0xB2004844 ^ (k*4096+i)),
 * byte[0] &= 0x3f (the generator of csrc/misc.cu, restated) */
static uint64_t splitmix64(uint64_t *st) {
  *st += 0x9E3779B97F4A7C15ull;
  uint64_t z = *st;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
void oracle_synth_blob(uint8_t *blob, uint64_t k) {
  for (uint32_t i = 0; i < NPTS; i++) {
    uint64_t st = 0xB2004844ull ^ (k * 4096ull + (uint64_t)i);
    uint8_t *o = blob + 32 * i;
    for (int u = 0; u < 4; u++) {
      uint64_t v = splitmix64(&st);
      for (int b = 0; b < 8; b++) o[8 * u + b] = (uint8_t)(v >> (56 - 8 * b));
    }
    o[0] &= 0x3f;
  }
}

/* g1_lincomb (src/lib.rs:241-243): msm(scalars, points) with the dependency's Pippenger.
 * points: n x (x || y) 48-byte big-endian canonical, all-zero = infinity; scalars: n x 32-byte big-endian
 * (reduced mod r like from_bytes_be).  Returns 1 for a point off the curve. */
int oracle_g1_lincomb(uint8_t out[48], const uint8_t *points_xy_be, const uint8_t *scalars_be, int n) {
  g1p *pts = (g1p *)malloc(sizeof(g1p) * (n ? n : 1));
  fr *sc = (fr *)malloc(sizeof(fr) * (n ? n : 1));
  int bad = 0;
  for (int i = 0; i < n; i++) {
    const uint8_t *b = points_xy_be + (size_t)96 * i;
    int zero = 1;
    for (int k = 0; k < 96; k++) if (b[k]) { zero = 0; break; }
    if (zero) { g1_neutral(&pts[i]); }
    else {
      fp_from_be(&pts[i].x, b);
      fp_from_be(&pts[i].y, b + 48);
      if (!g1_on_curve_affine(&pts[i].x, &pts[i].y)) bad = 1;
      fp_set_one(&pts[i].z);
    }
    fr m;
    fr_from_be(&m, scalars_be + (size_t)32 * i);
    fr_from_mont(&sc[i], &m);
  }
  if (!bad) {
    g1p r;
    msm_pippenger(&r, sc, pts, n);
    compress_g1(out, &r);
  }
  free(pts); free(sc);
  return bad;
}

/* n synthetic points (running sums of the first SRS point: P_i = (i + 1) g1[0]) and SplitMix64 scalars, written in
 * the formats oracle_g1_lincomb takes -- the cost of an MSM does not depend on which points it sums */
int oracle_synth_msm_inputs(uint8_t *points_xy_be, uint8_t *scalars_be, int n, uint64_t seed) {
  g1p *srs = (g1p *)malloc(sizeof(g1p) * NPTS);
  if (!rehydrate_srs(srs)) { free(srs); return 2; }
  g1p acc = srs[0];
  for (int i = 0; i < n; i++) {
    fp x, y;
    if (!g1_to_affine(&x, &y, &acc)) { free(srs); return 2; }
    fp_to_be(points_xy_be + (size_t)96 * i, &x);
    fp_to_be(points_xy_be + (size_t)96 * i + 48, &y);
    g1_add(&acc, &acc, &srs[0]);
    uint64_t st = 0xB2004844ull ^ (seed * 4096ull + (uint64_t)i);
    uint8_t *o = scalars_be + (size_t)32 * i;
    for (int u = 0; u < 4; u++) {
      uint64_t v = splitmix64(&st);
      for (int b = 0; b < 8; b++) o[8 * u + b] = (uint8_t)(v >> (56 - 8 * b));
    }
    o[0] &= 0x3f;
  }
  free(srs);
  return 0;
}
