// Carry-chain primitives for multi-limb integer arithmetic on sm_100a.
//
// Device: one PTX instruction each (add.cc / addc / mad.lo.cc / madc.hi.cc ...),
// ptxas maps the CC flag onto predicate carries and fuses lo/hi pairs into
// IMAD.WIDE where it can.
//
// Host (LWKZG_HOST_EMUL, test builds only): the same call sequences are
// emulated with a thread-local carry flag so that every algorithm built on
// these primitives (Montgomery multiplication, point formulas, pairing) can be
// unit-tested with g++ in a container without a GPU.  The shipped library
// never compiles this branch.
#pragma once
#include <stdint.h>

#if defined(LWKZG_HOST_EMUL)
#define LW_DEV
#define LW_HD
#define LW_INL inline
#define LW_CONST static const
#define LW_COLD static inline
#else
#define LW_DEV __device__
#define LW_HD __host__ __device__
#define LW_INL __device__ __forceinline__
#define LW_CONST __device__ __constant__ const
// cold, code-size-heavy helpers: out of line so the tower-field / pairing code
// does not explode at compile time
#define LW_COLD static __device__ __noinline__
#endif

namespace lw {
namespace ptx {

#if defined(LWKZG_HOST_EMUL)

static thread_local uint32_t g_cf = 0;

inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; g_cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + g_cf; g_cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + g_cf; }
// PTX: after sub.cc the CC.CF flag holds the BORROW; subc computes a - (b + CC.CF).
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; g_cf = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - g_cf; g_cf = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - g_cf; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_lo(a, b), c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_lo(a, b), c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_lo(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }
// borrow-out of the last sub chain as 0/1 (1 = borrowed)
inline uint32_t borrow_flag() { return g_cf; }

#else

LW_INL uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LW_INL uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LW_INL uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LW_INL uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LW_INL uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LW_INL uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LW_INL uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LW_INL uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
LW_INL uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
LW_INL uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
LW_INL uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
LW_INL uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
LW_INL uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
LW_INL uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
// borrow-out of the last sub chain as 0/1 (1 = borrowed): subc 0-0-borrow = 0 or 0xffffffff
LW_INL uint32_t borrow_flag() { uint32_t r; asm volatile("subc.u32 %0, 0, 0;" : "=r"(r)); return r & 1u; }

#endif

}  // namespace ptx
}  // namespace lw
