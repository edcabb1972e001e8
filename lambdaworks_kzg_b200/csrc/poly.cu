// Fiat-Shamir challenge (SHA-256) and polynomial evaluate / quotient kernels.
//   compute_challenge        /root/reference/src/utils.rs:120-154   (App. A.5)
//   Polynomial::evaluate + KZG::open's Ruffini division (src/lib.rs:320-329, 389-394; App. A.4)
#include "field.cuh"
#include "frpoly.cuh"
#include "kernels.h"
#include "sha256.cuh"

namespace lw {

__device__ __forceinline__ void load_block_words(uint32_t* w, const uint8_t* p16aligned) {
  const uint4* q = reinterpret_cast<const uint4*>(p16aligned);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint4 v = __ldg(q + i);
    w[4 * i + 0] = bswap32(v.x); w[4 * i + 1] = bswap32(v.y); w[4 * i + 2] = bswap32(v.z); w[4 * i + 3] = bswap32(v.w);
  }
}

// words 4..7 of block 0: the 16 "degree" bytes after the domain.  le64(4096) || le64(0) in the reference and in
// MODE_CKZG_LE; int.to_bytes(4096, 16, 'big') in MODE_DENEB (consensus-specs deneb compute_challenge)
__device__ __forceinline__ void challenge_header_words(uint32_t* w, int be_header) {
  w[0] = 0x4653424cu; w[1] = 0x4f425645u; w[2] = 0x52494659u; w[3] = 0x5f56315fu;  // "FSBL" "OBVE" "RIFY" "_V1_"
  w[4] = be_header ? 0u : 0x00100000u;   // 4096 little-endian: 00 10 00 00 00 00 00 00
  w[5] = 0u; w[6] = 0u;
  w[7] = be_header ? 0x00001000u : 0u;
}

// The hashed message is  "FSBLOBVERIFY_V1_" || le64(4096) || le64(0) || blob || compress(C)
// = 131152 bytes = 2049 full blocks + 16 bytes.  Everything up to byte 131072
// (2048 blocks: 32-byte header + blob[0..131040)) does not depend on the
// commitment, so this midstate runs concurrently with the commitment MSM.
// SHA-256 is inherently sequential per message: one thread per blob.
__global__ void __launch_bounds__(32) challenge_midstate_kernel(Sha256State* __restrict__ states, const uint8_t* __restrict__ blobs, int n, int be_header) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  const uint8_t* blob = blobs + (size_t)b * BLOB_BYTES;
  Sha256State s;
  sha256_init(s);
  uint32_t w[16];
  // block 0: domain (16) + le64(4096) + le64(0) + blob[0..32)
  challenge_header_words(w, be_header);
  {
    const uint4* q = reinterpret_cast<const uint4*>(blob);
    uint4 v0 = __ldg(q), v1 = __ldg(q + 1);
    w[8] = bswap32(v0.x); w[9] = bswap32(v0.y); w[10] = bswap32(v0.z); w[11] = bswap32(v0.w);
    w[12] = bswap32(v1.x); w[13] = bswap32(v1.y); w[14] = bswap32(v1.z); w[15] = bswap32(v1.w);
  }
  sha256_compress(s, w);
  for (int k = 1; k < 2048; k++) {
    load_block_words(w, blob + 32 + (size_t)(k - 1) * 64);
    sha256_compress(s, w);
  }
  states[b] = s;
}

// Latency variant for small batches: one WARP per blob, message schedules expanded by
// all lanes (sha256_warp_blocks).  Same result, ~2x lower latency per blob; not used for
// large batches where a thread per blob keeps the issue slots free for the MSM.
__global__ void __launch_bounds__(32) challenge_midstate_warp_kernel(Sha256State* __restrict__ states, const uint8_t* __restrict__ blobs, int n, int be_header) {
  __shared__ uint32_t wk[64 * 32];
  const int b = blockIdx.x;
  const uint8_t* blob = blobs + (size_t)b * BLOB_BYTES;
  Sha256State s;
  sha256_init(s);
  sha256_warp_blocks(s, 2048, [&](int blk, uint32_t* w) {
    if (blk == 0) {
      challenge_header_words(w, be_header);
      const uint4* q = reinterpret_cast<const uint4*>(blob);
      uint4 v0 = __ldg(q), v1 = __ldg(q + 1);
      w[8] = bswap32(v0.x); w[9] = bswap32(v0.y); w[10] = bswap32(v0.z); w[11] = bswap32(v0.w);
      w[12] = bswap32(v1.x); w[13] = bswap32(v1.y); w[14] = bswap32(v1.z); w[15] = bswap32(v1.w);
    } else {
      load_block_words(w, blob + 32 + (size_t)(blk - 1) * 64);
    }
  }, wk);
  if ((threadIdx.x & 31) == 0) states[b] = s;
}

// Throughput/latency middle ground for batches where nothing else hides the hash (blob proofs for given
// commitments, batched verification): one warp serves G blobs.  In each iteration the 32 lanes expand the
// message schedules of 32 / G consecutive blocks of each blob, then lanes 0 .. G-1 run the rounds of their blob.
// Per blob this is as fast as the warp-per-blob kernel (the rounds are the critical path either way) but it
// issues G times fewer warp instructions: 4096 warp-per-blob hashes saturate the issue slots of the whole GPU
// (measured: a 4096-blob verification spent 17 ms there), 512 warps of this kernel do not.
template <int G>
__global__ void __launch_bounds__(32) challenge_midstate_group_kernel(Sha256State* __restrict__ states, const uint8_t* __restrict__ blobs, int n, int be_header) {
  __shared__ uint32_t wk[64 * 32];
  constexpr int PER = 32 / G;   // blocks per blob and iteration; 2048 % PER == 0
  const int lane = threadIdx.x;
  const int blob_e = blockIdx.x * G + lane / PER;   // the blob this lane expands schedules for
  const int j = lane % PER;
  const int blob_r = blockIdx.x * G + lane;         // the blob whose rounds this lane runs (lanes < G)
  const uint8_t* blob = blobs + (size_t)(blob_e < n ? blob_e : 0) * BLOB_BYTES;
  Sha256State s;
  sha256_init(s);
  const uint32_t one = sha_runtime_one();
  for (int base = 0; base < 2048; base += PER) {
    if (blob_e < n) {
      const int blk = base + j;
      uint32_t w[16];
      if (blk == 0) {
        challenge_header_words(w, be_header);
        const uint4* q = reinterpret_cast<const uint4*>(blob);
        uint4 v0 = __ldg(q), v1 = __ldg(q + 1);
        w[8] = bswap32(v0.x); w[9] = bswap32(v0.y); w[10] = bswap32(v0.z); w[11] = bswap32(v0.w);
        w[12] = bswap32(v1.x); w[13] = bswap32(v1.y); w[14] = bswap32(v1.z); w[15] = bswap32(v1.w);
      } else {
        load_block_words(w, blob + 32 + (size_t)(blk - 1) * 64);
      }
      sha256_expand_wk(wk, lane, w);
    }
    __syncwarp();
    if (lane < G && blob_r < n) {
      for (int t = 0; t < PER; t++) sha256_rounds_wk(s, wk, lane * PER + t, one);
    }
    __syncwarp();
  }
  if (lane < G && blob_r < n) states[blob_r] = s;
}

// blocks 2048 (blob tail 32 B + commitment[0..32)) and 2049 (commitment[32..48) + padding)
__global__ void __launch_bounds__(32) challenge_finish_kernel(uint32_t* __restrict__ z_out, const Sha256State* __restrict__ states,
                                                               const uint8_t* __restrict__ blobs, const uint8_t* __restrict__ commit48, int n,
                                                               int le_digest) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  Sha256State s = states[b];
  uint32_t w[16];
  const uint8_t* tail = blobs + (size_t)b * BLOB_BYTES + (BLOB_BYTES - 32);
  const uint8_t* c = commit48 + (size_t)b * 48;
  {
    const uint4* q = reinterpret_cast<const uint4*>(tail);
    uint4 v0 = __ldg(q), v1 = __ldg(q + 1);
    w[0] = bswap32(v0.x); w[1] = bswap32(v0.y); w[2] = bswap32(v0.z); w[3] = bswap32(v0.w);
    w[4] = bswap32(v1.x); w[5] = bswap32(v1.y); w[6] = bswap32(v1.z); w[7] = bswap32(v1.w);
  }
  for (int i = 0; i < 8; i++) w[8 + i] = ((uint32_t)c[4 * i] << 24) | ((uint32_t)c[4 * i + 1] << 16) | ((uint32_t)c[4 * i + 2] << 8) | c[4 * i + 3];
  sha256_compress(s, w);
  for (int i = 0; i < 4; i++) w[i] = ((uint32_t)c[32 + 4 * i] << 24) | ((uint32_t)c[33 + 4 * i] << 16) | ((uint32_t)c[34 + 4 * i] << 8) | c[35 + 4 * i];
  w[4] = 0x80000000u;
  for (int i = 5; i < 14; i++) w[i] = 0;
  w[14] = 0;
  w[15] = 131152u * 8u;
  sha256_compress(s, w);
  // digest read big-endian, reduced mod r (hash_field_unsafe, utils.rs:148-154)
  // (MODE_CKZG_LE reads the digest little-endian: SURVEY App. B)
  Fr z;
  for (int i = 0; i < 8; i++) z.l[i] = le_digest ? bswap32(s.h[i]) : s.h[7 - i];
  mod_reduce_small<FrCfg, 2>(z.l);
  for (int i = 0; i < 8; i++) z_out[b * 8 + i] = z.l[i];
}

__global__ void fr_from_be_kernel(uint32_t* __restrict__ out, const uint8_t* __restrict__ in, int n) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  Fr z = fr_canon_from_be32(in + (size_t)b * 32);
  for (int i = 0; i < 8; i++) out[b * 8 + i] = z.l[i];
}

__device__ __forceinline__ Fr shfl_down_fr(const Fr& v, int d) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], d);
  return r;
}

// One warp per blob; lane t owns coefficients [128 t, 128 t + 128).
constexpr int POLY_WARPS = 4;
__global__ void __launch_bounds__(POLY_WARPS * 32) poly_eval_quot_kernel(uint32_t* __restrict__ q_out, uint32_t* __restrict__ y_out,
                                                                          uint8_t* __restrict__ y_be_out, const uint8_t* __restrict__ blobs,
                                                                          const uint32_t* __restrict__ zs, int n) {
  const int lane = threadIdx.x & 31;
  const int blob = blockIdx.x * POLY_WARPS + (threadIdx.x >> 5);
  if (blob >= n) return;  // whole warp exits together
  constexpr int M = N_POINTS / 32;  // 128 coefficients per lane
  const uint8_t* base = blobs + (size_t)blob * BLOB_BYTES + (size_t)lane * M * 32;
  Fr zc;
  for (int i = 0; i < 8; i++) zc.l[i] = zs[blob * 8 + i];
  const Fr z = fr_to_mont(zc);

  auto load = [&](int k) {
    const uint4* p = reinterpret_cast<const uint4*>(base + (size_t)k * 32);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    return fr_canon_from_be_words(w);
  };

  // pass 1: local Horner value, then inclusive suffix scan over lanes
  Fr H = chunk_horner(load, M, z);
  Fr pw = fr_pow2k(z, 7);  // z^128
  for (int d = 1; d < 32; d <<= 1) {
    Fr o = shfl_down_fr(H, d);
    if (lane + d < 32) H = fr_add(H, fr_mul(pw, o));
    pw = fr_sqr(pw);
  }
  Fr carry = shfl_down_fr(H, 1);
  if (lane == 31) carry = fr_zero();

  // pass 2: exact quotient coefficients
  uint32_t* q = q_out ? q_out + (size_t)blob * N_POINTS * 8 : nullptr;
  const int start = lane * M;
  Fr v = chunk_sweep(
      load,
      [&](int k, const Fr& val) {
        int g = start + k;
        if (q && g >= 1) {
          uint4* dst = reinterpret_cast<uint4*>(q + (size_t)(g - 1) * 8);
          dst[0] = make_uint4(val.l[0], val.l[1], val.l[2], val.l[3]);
          dst[1] = make_uint4(val.l[4], val.l[5], val.l[6], val.l[7]);
        }
      },
      M, z, carry);
  if (q && lane == 31) {
    uint4* dst = reinterpret_cast<uint4*>(q + (size_t)(N_POINTS - 1) * 8);
    dst[0] = make_uint4(0, 0, 0, 0);
    dst[1] = make_uint4(0, 0, 0, 0);
  }
  if (lane == 0) {
    if (y_out) for (int i = 0; i < 8; i++) y_out[blob * 8 + i] = v.l[i];
    if (y_be_out) fr_canon_to_be32(y_be_out + (size_t)blob * 32, v);
  }
}

// latency = true: nothing else runs beside the hash (blob proofs for given commitments, verification), so the
// warp-per-blob kernel's ~2x shorter critical path is what counts; false: the hash hides under a commitment MSM and
// the thread-per-blob kernel leaves the issue slots to it.
void launch_challenge_midstate(void* d_states, const void* d_blobs, int n, cudaStream_t st, bool latency, bool be_header) {
  LW_SAME_CARVEOUT(challenge_midstate_warp_kernel);
  LW_SAME_CARVEOUT(challenge_midstate_group_kernel<8>);
  LW_SAME_CARVEOUT(challenge_midstate_kernel);
  if (n <= 0) return;
  const int bh = be_header ? 1 : 0;
  if (n <= 64)
    challenge_midstate_warp_kernel<<<n, 32, 0, st>>>((Sha256State*)d_states, (const uint8_t*)d_blobs, n, bh);
  else if (latency)
    challenge_midstate_group_kernel<8><<<(n + 7) / 8, 32, 0, st>>>((Sha256State*)d_states, (const uint8_t*)d_blobs, n, bh);
  else
    challenge_midstate_kernel<<<(n + 31) / 32, 32, 0, st>>>((Sha256State*)d_states, (const uint8_t*)d_blobs, n, bh);
  count_launch();
}
void launch_challenge_finish(void* d_z, const void* d_states, const void* d_blobs, const void* d_commit48, int n, cudaStream_t st, bool le_digest) {
  if (n <= 0) return;
  LW_SAME_CARVEOUT(challenge_finish_kernel);
  challenge_finish_kernel<<<(n + 31) / 32, 32, 0, st>>>((uint32_t*)d_z, (const Sha256State*)d_states, (const uint8_t*)d_blobs, (const uint8_t*)d_commit48, n,
                                                       le_digest ? 1 : 0);
  count_launch();
}
void launch_fr_from_be(void* d_z, const void* d_z_be32, int n, cudaStream_t st) {
  if (n <= 0) return;
  fr_from_be_kernel<<<(n + 63) / 64, 64, 0, st>>>((uint32_t*)d_z, (const uint8_t*)d_z_be32, n);
  count_launch();
}
void launch_poly_eval_quot(void* d_q, void* d_y, void* d_y_be32, const void* d_blobs, const void* d_z, int n, cudaStream_t st) {
  if (n <= 0) return;
  LW_SAME_CARVEOUT(poly_eval_quot_kernel);
  poly_eval_quot_kernel<<<(n + POLY_WARPS - 1) / POLY_WARPS, POLY_WARPS * 32, 0, st>>>((uint32_t*)d_q, (uint32_t*)d_y, (uint8_t*)d_y_be32,
                                                                                      (const uint8_t*)d_blobs, (const uint32_t*)d_z, n);
  count_launch();
}

}  // namespace lw
