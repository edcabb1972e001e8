// SHA-256 (FIPS 180-4) for the Fiat-Shamir challenges.
// Replaces the `sha256` + `hex` crates used by
// /root/reference/src/utils.rs:148-154 (hash_field_unsafe): the reference
// hex-encodes the digest and decodes it again; we keep the raw 32 bytes.
#pragma once
#include "ptx.cuh"

namespace lw {

LW_CONST uint32_t SHA_K[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

struct Sha256State {
  uint32_t h[8];
};

LW_INL uint32_t sha_rotr(uint32_t x, int n) {
#if defined(LWKZG_HOST_EMUL)
  return (x >> n) | (x << (32 - n));
#else
  return __funnelshift_r(x, x, n);
#endif
}

LW_INL void sha256_init(Sha256State& s) {
  s.h[0] = 0x6a09e667u; s.h[1] = 0xbb67ae85u; s.h[2] = 0x3c6ef372u; s.h[3] = 0xa54ff53au;
  s.h[4] = 0x510e527fu; s.h[5] = 0x9b05688cu; s.h[6] = 0x1f83d9abu; s.h[7] = 0x5be0cd19u;
}

// One compression; w[16] = the block as big-endian words (already byte-swapped).
LW_INL void sha256_compress(Sha256State& s, uint32_t* w) {
  uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], h = s.h[7];
#pragma unroll
  for (int i = 0; i < 64; i++) {
    if (i >= 16) {
      uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
      uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
      w[i & 15] = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
    }
    // The hash is one dependency chain; written so that the chain through e is three instructions per round
    // (rotate -> xor3 -> add3): everything that does not need this round's e is summed beforehand.
    const uint32_t dhk = d + h + (SHA_K[i] + w[i & 15]);
    const uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
    const uint32_t ch = g ^ (e & (f ^ g));
    const uint32_t e_new = dhk + S1 + ch;                 // d + t1
    const uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
    const uint32_t mj = (a & b) | (c & (a | b));
    const uint32_t t2md = S0 + mj - d;                    // t2 - d
    h = g; g = f; f = e; d = c; c = b; b = a; a = e_new + t2md; e = e_new;   // a = t1 + t2
  }
  s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += h;
}

// Rounds only, with W[i] + K[i] precomputed (warp-cooperative path below).
// wk is laid out [64][32]: word i of the block prepared by lane b at wk[i * 32 + b].
//
// One lane runs the rounds, so the warp's issue rate is the bound: a round is 6 funnel shifts + 4 three-input
// logic ops + 7 additions, and on the ALU pipe (one warp instruction every two cycles) that is the measured 32
// cycles per round.  The additions are therefore written as multiply-adds by a run-time 1 (`one`: a value ptxas
// cannot fold), which issue on the FMA pipe beside the shifts and logic ops.
#if defined(LWKZG_HOST_EMUL)
LW_INL uint32_t sha_add(uint32_t a, uint32_t b, uint32_t) { return a + b; }
LW_INL uint32_t sha_runtime_one() { return 1u; }
#else
static __device__ uint32_t g_sha_one = 1u;
LW_INL uint32_t sha_add(uint32_t a, uint32_t b, uint32_t one) { return a * one + b; }
LW_INL uint32_t sha_runtime_one() { return *(volatile uint32_t*)&g_sha_one; }
#endif
LW_INL void sha256_rounds_wk(Sha256State& s, const uint32_t* wk, int b, uint32_t one) {
  uint32_t a = s.h[0], bb = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], h = s.h[7];
  // all 64 schedule words are requested before the first round: left to itself ptxas issues each shared-memory
  // load right before its use, which puts the ~30-cycle load on the dependency chain of EVERY round
  uint32_t kw[64];
#pragma unroll
  for (int i = 0; i < 64; i++) kw[i] = wk[i * 32 + b];
#pragma unroll
  for (int i = 0; i < 64; i++) {
    const uint32_t dhk = sha_add(sha_add(d, h, one), kw[i], one);   // off the e-chain (see sha256_compress)
    const uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
    const uint32_t ch = g ^ (e & (f ^ g));
    const uint32_t e_new = sha_add(S1, sha_add(dhk, ch, one), one);
    const uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
    const uint32_t mj = (a & bb) | (c & (a | bb));
    const uint32_t t2md = sha_add(S0, mj - d, one);
    h = g; g = f; f = e; d = c; c = bb; bb = a; a = sha_add(e_new, t2md, one); e = e_new;
  }
  s.h[0] += a; s.h[1] += bb; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += h;
}
LW_INL void sha256_rounds_wk(Sha256State& s, const uint32_t* wk, int b) { sha256_rounds_wk(s, wk, b, sha_runtime_one()); }
// Message-schedule expansion of one block into wk (adds the round constants).
LW_INL void sha256_expand_wk(uint32_t* wk, int b, uint32_t* w /* 16 words in, 64 used */) {
#pragma unroll
  for (int i = 0; i < 64; i++) {
    if (i >= 16) {
      uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
      uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
      w[i & 15] = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
    }
    wk[i * 32 + b] = w[i & 15] + SHA_K[i];
  }
}
#if !defined(LWKZG_HOST_EMUL)
// Warp-cooperative hashing of `nblocks` consecutive full blocks: SHA-256 is
// sequential in the state but the message schedule (60 % of the work) is not, so
// the 32 lanes expand 32 blocks in parallel and lane 0 then only runs the rounds.
// load(blk, w16) must fill the 16 big-endian words of block blk.  The state is
// lane 0's; wk is 64*32 words of shared memory owned by this warp.
template <class Load>
__device__ __forceinline__ void sha256_warp_blocks(Sha256State& s, int nblocks, Load load, uint32_t* wk) {
  const int lane = threadIdx.x & 31;
  const uint32_t one = sha_runtime_one();
  for (int base = 0; base < nblocks; base += 32) {
    const int blk = base + lane;
    if (blk < nblocks) {
      uint32_t w[16];
      load(blk, w);
      sha256_expand_wk(wk, lane, w);
    }
    __syncwarp();
    if (lane == 0) {
      const int cnt = (nblocks - base < 32) ? (nblocks - base) : 32;
      for (int b = 0; b < cnt; b++) sha256_rounds_wk(s, wk, b, one);
    }
    __syncwarp();
  }
}
#endif

LW_INL void sha256_digest_bytes(uint8_t* out32, const Sha256State& s) {
  for (int i = 0; i < 8; i++) {
    out32[4 * i] = (uint8_t)(s.h[i] >> 24); out32[4 * i + 1] = (uint8_t)(s.h[i] >> 16);
    out32[4 * i + 2] = (uint8_t)(s.h[i] >> 8); out32[4 * i + 3] = (uint8_t)s.h[i];
  }
}

// Generic byte-stream hash (cold path: batch challenge, tests).
LW_DEV inline void sha256_oneshot(uint8_t* out32, const uint8_t* msg, size_t len) {
  Sha256State s;
  sha256_init(s);
  uint32_t w[16];
  size_t off = 0;
  for (; off + 64 <= len; off += 64) {
    for (int i = 0; i < 16; i++) {
      const uint8_t* q = msg + off + 4 * i;
      w[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
    sha256_compress(s, w);
  }
  uint8_t tail[128];
  size_t rem = len - off;
  for (size_t i = 0; i < rem; i++) tail[i] = msg[off + i];
  tail[rem] = 0x80;
  size_t tl = (rem + 9 <= 64) ? 64 : 128;
  for (size_t i = rem + 1; i < tl; i++) tail[i] = 0;
  unsigned long long bits = (unsigned long long)len * 8ull;
  for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
  for (size_t o = 0; o < tl; o += 64) {
    for (int i = 0; i < 16; i++) {
      const uint8_t* q = tail + o + 4 * i;
      w[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
    sha256_compress(s, w);
  }
  sha256_digest_bytes(out32, s);
}

}  // namespace lw
