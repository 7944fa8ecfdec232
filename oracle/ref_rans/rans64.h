/* rans64.h -- TEST INFRASTRUCTURE (oracle/): a restatement of the 64-bit rANS primitives of rygorous/ryg_rans
 * (rans64.h @ c9d162d996fd600315af9ae8eb89d832576cb32d, public domain, Fabian Giesen 2014), the un-vendored dependency the
 * reference's pMCTF/cpp/rans/rans.cpp includes (cpp/3rdparty/ryg_rans/CMakeLists.txt.in:8-9; no network here).  Only the
 * entry points rans.cpp calls are provided (rans.cpp:37-73,143,153,161,269,293,301).  Published algorithm: state x in
 * [2^31, 2^63); encoding a symbol of interval [start, start+freq) at `scale_bits` bits first emits the low 32 bits of x
 * (downwards in memory) when x would overflow, then x' = ((x / freq) << scale_bits) + (x % freq) + start; decoding inverts
 * it and refills 32 bits when x drops below 2^31.
 * With this header oracle/Makefile compiles the reference's OWN rans.cpp / py_rans.cpp / ops.cpp into oracle/_ref/, the
 * checker tests/test_rans.py holds the product's streams against.  Nothing under learned-pmctf_b200/ includes it. */
#ifndef RANS64_RESTATED_H
#define RANS64_RESTATED_H

#include <stdint.h>

#ifndef Rans64Assert
#include <assert.h>
#define Rans64Assert assert
#endif

#define RANS64_L (1ull << 31)

typedef uint64_t Rans64State;

static inline void Rans64EncInit(Rans64State *r) { *r = RANS64_L; }

static inline void Rans64EncPut(Rans64State *r, uint32_t **pptr, uint32_t start, uint32_t freq, uint32_t scale_bits)
{
    Rans64Assert(freq != 0);
    uint64_t x = *r;
    const uint64_t x_max = ((RANS64_L >> scale_bits) << 32) * freq;
    if (x >= x_max) {
        *pptr -= 1;
        **pptr = (uint32_t)x;
        x >>= 32;
        Rans64Assert(x < x_max);
    }
    *r = ((x / freq) << scale_bits) + (x % freq) + start;
}

static inline void Rans64EncFlush(Rans64State *r, uint32_t **pptr)
{
    const uint64_t x = *r;
    *pptr -= 2;
    (*pptr)[0] = (uint32_t)(x >> 0);
    (*pptr)[1] = (uint32_t)(x >> 32);
}

static inline void Rans64DecInit(Rans64State *r, uint32_t **pptr)
{
    uint64_t x = (uint64_t)((*pptr)[0]) << 0;
    x |= (uint64_t)((*pptr)[1]) << 32;
    *pptr += 2;
    *r = x;
}

static inline uint32_t Rans64DecGet(Rans64State *r, uint32_t scale_bits) { return (uint32_t)(*r & ((1u << scale_bits) - 1)); }

static inline void Rans64DecAdvance(Rans64State *r, uint32_t **pptr, uint32_t start, uint32_t freq, uint32_t scale_bits)
{
    const uint64_t mask = (1ull << scale_bits) - 1;
    uint64_t x = *r;
    x = freq * (x >> scale_bits) + (x & mask) - start;
    if (x < RANS64_L) {
        x = (x << 32) | **pptr;
        *pptr += 1;
        Rans64Assert(x >= RANS64_L);
    }
    *r = x;
}

#endif /* RANS64_RESTATED_H */
