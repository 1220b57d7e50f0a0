// hash.cuh -- blake2b-512 single-block compression and the decimal-ASCII leaf
// encoding, as the reference's Merkle tree uses them:
//   leaf  = blake2b(str(value).encode())          code/merkle.py:13-14, code/algebra.py:53-57
//   node  = blake2b(left || right)                code/merkle.py:6-11
// A leaf message is 1..39 bytes and a node message exactly 128 bytes, so both
// are ONE compression of the final block (RFC 7693; hashlib.blake2b defaults:
// 64-byte digest, no key, fanout = depth = 1).
// __host__ __device__ so tests/emu can run the same code on the CPU.
#pragma once
#include "field.cuh"

namespace sa {

SA_HD uint64_t b2_rotr(uint64_t x, int r) {
#if defined(__CUDA_ARCH__)
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if (r == 32) return ((uint64_t)lo << 32) | hi;
    if (r == 24) {
        // bytes of (hi:lo) rotated right by 3
        uint32_t nlo = __byte_perm(lo, hi, 0x6543), nhi = __byte_perm(lo, hi, 0x2107);
        return ((uint64_t)nhi << 32) | nlo;
    }
    if (r == 16) {
        uint32_t nlo = __byte_perm(lo, hi, 0x5432), nhi = __byte_perm(lo, hi, 0x1076);
        return ((uint64_t)nhi << 32) | nlo;
    }
    // r == 63: rotate left by one
    uint32_t nlo = __funnelshift_l(hi, lo, 1), nhi = __funnelshift_l(lo, hi, 1);
    return ((uint64_t)nhi << 32) | nlo;
#else
    return (x >> r) | (x << (64 - r));
#endif
}

#define SA_B2_G(a, b, c, d, x, y)   \
    do {                            \
        a = a + b + (x);            \
        d = b2_rotr(d ^ a, 32);     \
        c = c + d;                  \
        b = b2_rotr(b ^ c, 24);     \
        a = a + b + (y);            \
        d = b2_rotr(d ^ a, 16);     \
        c = c + d;                  \
        b = b2_rotr(b ^ c, 63);     \
    } while (0)

#define SA_B2_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    do {                                                                                   \
        SA_B2_G(v0, v4, v8, v12, m[s0], m[s1]);                                            \
        SA_B2_G(v1, v5, v9, v13, m[s2], m[s3]);                                            \
        SA_B2_G(v2, v6, v10, v14, m[s4], m[s5]);                                           \
        SA_B2_G(v3, v7, v11, v15, m[s6], m[s7]);                                           \
        SA_B2_G(v0, v5, v10, v15, m[s8], m[s9]);                                           \
        SA_B2_G(v1, v6, v11, v12, m[s10], m[s11]);                                         \
        SA_B2_G(v2, v7, v8, v13, m[s12], m[s13]);                                          \
        SA_B2_G(v3, v4, v9, v14, m[s14], m[s15]);                                          \
    } while (0)

// digest (8 words) of a message that fits one 128-byte block; len = message bytes,
// m[] zero padded.  h0 = IV ^ parameter block (0x01010040 into word 0).
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
void blake2b_single_block(uint64_t out[8], const uint64_t m[16], uint32_t len) {
    const uint64_t iv0 = 0x6a09e667f3bcc908ULL, iv1 = 0xbb67ae8584caa73bULL, iv2 = 0x3c6ef372fe94f82bULL,
                   iv3 = 0xa54ff53a5f1d36f1ULL, iv4 = 0x510e527fade682d1ULL, iv5 = 0x9b05688c2b3e6c1fULL,
                   iv6 = 0x1f83d9abfb41bd6bULL, iv7 = 0x5be0cd19137e2179ULL;
    const uint64_t h0 = iv0 ^ 0x01010040ULL;
    uint64_t v0 = h0, v1 = iv1, v2 = iv2, v3 = iv3, v4 = iv4, v5 = iv5, v6 = iv6, v7 = iv7;
    uint64_t v8 = iv0, v9 = iv1, v10 = iv2, v11 = iv3, v12 = iv4 ^ (uint64_t)len, v13 = iv5, v14 = ~iv6,
             v15 = iv7;
    SA_B2_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    SA_B2_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    SA_B2_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
    SA_B2_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
    SA_B2_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
    SA_B2_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
    SA_B2_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
    SA_B2_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
    SA_B2_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
    SA_B2_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
    SA_B2_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    SA_B2_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    out[0] = h0 ^ v0 ^ v8;
    out[1] = iv1 ^ v1 ^ v9;
    out[2] = iv2 ^ v2 ^ v10;
    out[3] = iv3 ^ v3 ^ v11;
    out[4] = iv4 ^ v4 ^ v12;
    out[5] = iv5 ^ v5 ^ v13;
    out[6] = iv6 ^ v6 ^ v14;
    out[7] = iv7 ^ v7 ^ v15;
}

// str(value).encode(): decimal ASCII, most significant digit first, no leading
// zeros ("0" for zero), packed little-endian into w[0..4] (40 bytes, zero padded).
// Returns the length in bytes (1..39).   code/algebra.py:53-57
//
// Round 2's second version (the first one, 650 instructions of a leaf's 2 650, divided by 1e9 with 64-bit integer
// reciprocals and peeled 39 digits off with /10, one compare and one byte insertion each - all on the ALU pipe, the
// one blake2b saturates).  Now: base-1e8 limbs, so that a limb is two 4-digit groups = two aligned 32-bit words of
// the output; the long division estimates each quotient word with the FP64 unit (idle in this kernel) from BELOW and
// repairs it with one compare; a 4-digit group becomes four ASCII bytes with multiplies (the FMA pipe, idle too);
// the length comes from the top non-zero limb instead of from every digit.

// quotient and remainder of rem_in * 2^32 + lo by 1e8 (rem_in < 1e8, so the quotient fits 32 bits).
// K = 1e-8 (1 - 1e-13): the estimate cur * K is below the true quotient q by less than 2^32 * 1.01e-13 + rounding
// (three roundings of 2^-53 relative) < 1e-3, so its integer part is q or q - 1: one repair step, never an overflow.
SA_HD uint32_t dec_div_1e8(uint32_t rem_in, uint32_t lo, uint32_t &rem_out) {
    const uint64_t cur = ((uint64_t)rem_in << 32) | lo;
    uint32_t d = (uint32_t)((double)cur * 9.999999999999e-9);
    uint32_t r = lo - d * 100000000u;  // cur - d * 1e8 in [0, 2e8): the low word is the whole value
    if (r >= 100000000u) {
        d++;
        r -= 100000000u;
    }
    rem_out = r;
    return d;
}
// v < 10000 -> its four decimal digits as ASCII, most significant digit in the low byte.
// (v * 5243) >> 19 = v / 100 for v < 43699, (x * 205) >> 11 = x / 10 for x < 1029; a pair "ab" of x = 10 a + b is
// a + 256 b = 256 x - 2559 a, and the two '0' offsets ride along in the same multiply-add.
SA_HD uint32_t dec4_ascii(uint32_t v) {
    const uint32_t d = (v * 5243u) >> 19, r = v - d * 100u;
    const uint32_t td = (d * 205u) >> 11, tr = (r * 205u) >> 11;
    const uint32_t pd = d * 256u + 0x3030u - td * 2559u;
    const uint32_t pr = r * 256u + 0x3030u - tr * 2559u;
    return pd + (pr << 16);
}
SA_HD uint32_t fe_decimal_words(uint64_t w[5], const fe &x) {
    // base-1e8 limbs, least significant first; 2^128 < 1e40, so five limbs of eight digits.  After k limbs have been
    // taken off, the quotient is below 2^128 / 1e8^k = 2^101.5, 2^74.9, 2^48.3, 2^21.7: the long divisions run over
    // 4, 4, 3 and 2 words and the last limb is the remaining quotient itself.  The first word of every long division
    // has no remainder coming in: a 32-bit division by a constant.
    uint32_t q[4] = {x.v[0], x.v[1], x.v[2], x.v[3]};
    uint32_t chunk[5];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; k++) {
        const int top = k <= 1 ? 3 : (k == 2 ? 2 : 1);
        uint32_t rem = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 3; i >= 0; i--) {
            if (i > top) continue;
            if (i == top) {
                const uint32_t d = q[i] / 100000000u;
                rem = q[i] - d * 100000000u;
                q[i] = d;
            } else {
                q[i] = dec_div_1e8(rem, q[i], rem);
            }
        }
        chunk[k] = rem;
    }
    chunk[4] = q[0];  // < 3 402 824
    // the 40-character zero-padded string, most significant limb first: word j = limb 4 - j
    uint64_t t[6];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 5; j++) {
        const uint32_t c = chunk[4 - j];
        const uint32_t hi4 = c / 10000u, lo4 = c - hi4 * 10000u;
        t[j] = (uint64_t)dec4_ascii(hi4) | ((uint64_t)dec4_ascii(lo4) << 32);
    }
    t[5] = 0;
    // significant digits: eight per limb below the top non-zero one plus that limb's own
    uint32_t kt = 0, v = chunk[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 1; k < 5; k++) {
        if (chunk[k] != 0) {
            kt = (uint32_t)k;
            v = chunk[k];
        }
    }
    const uint32_t nd = 8u * kt + 1u + (v >= 10u) + (v >= 100u) + (v >= 1000u) + (v >= 10000u) + (v >= 100000u) +
                        (v >= 1000000u) + (v >= 10000000u);
    // drop the leading (40 - nd) bytes: out byte i = string byte (40 - nd + i); the tail fills with zeros
    const uint32_t lz = 40u - nd;
    const uint32_t ws = lz >> 3, bs = (lz & 7u) * 8u;
    // whole-word shift by ws in {0..4} as three conditional stages (1, 2, 4 words)
    if (ws & 1u) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 5; i++) t[i] = t[i + 1];
    }
    if (ws & 2u) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 5; i++) t[i] = (i + 2 < 6) ? t[i + 2] : 0;
    }
    if (ws & 4u) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 5; i++) t[i] = (i + 4 < 6) ? t[i + 4] : 0;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 5; i++) {
        const uint64_t lo = t[i] >> bs;
        const uint64_t hi = bs ? (t[i + 1] << (64u - bs)) : 0;
        w[i] = lo | hi;
    }
    return nd;
}

// leaf digest of one field element
SA_HD void merkle_leaf_digest(uint64_t out[8], const fe &x) {
    uint64_t m[16];
    const uint32_t len = fe_decimal_words(m, x);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 5; i < 16; i++) m[i] = 0;
    blake2b_single_block(out, m, len);
}

// node digest of two child digests
// (out may alias left or right)
SA_HD void merkle_node_digest(uint64_t out[8], const uint64_t left[8], const uint64_t right[8]) {
    uint64_t m[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 8; i++) {
        m[i] = left[i];
        m[8 + i] = right[i];
    }
    blake2b_single_block(out, m, 128u);
}

// ---- four lanes per compression (latency-bound upper tree levels) ----------------------------
// The 4x4 state is split by columns: lane j of a quad holds (a, b, c, d) = (v[j], v[4+j], v[8+j],
// v[12+j]).  A round is the column G on the lane's own column, a rotation of b, c, d by 1, 2, 3
// lanes (which puts the diagonals in the lanes), the diagonal G, and the rotation back.  The
// message words are read from `msg` (16 contiguous words = left || right child digests, in shared
// memory on the device) by index, so no lane needs all sixteen in registers.  The dependent chain
// per compression is the same 24 G steps, but a warp that has only a few nodes left to hash keeps
// all its lanes busy and issues a quarter of the instructions per node.
// sigma rows packed as sixteen nibbles, nibble i = sigma[r][i]
SA_HD uint64_t b2_sigma_packed(int r) {
    const uint64_t s[10] = {0xfedcba9876543210ULL, 0x357b20c16df984aeULL, 0x491763eadf250c8bULL,
                            0x8f04a562ebcd1397ULL, 0xd386cb1efa427509ULL, 0x91ef57d438b0a6c2ULL,
                            0xb8293670a4def15cULL, 0xa2684f05931ce7bdULL, 0x5a417d2c803b9ef6ULL,
                            0x0dc3e9bf5167482aULL};
    return s[r % 10];
}
SA_HD void b2_coop4_init(int j, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d, uint32_t len) {
    const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                            0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                            0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    a = iv[j] ^ (j == 0 ? 0x01010040ULL : 0);
    b = iv[4 + j];
    c = iv[j];
    d = iv[4 + j];
    if (j == 0) d ^= (uint64_t)len;
    if (j == 2) d = ~d;
}
// one G on the lane's current (a, b, c, d) with message words x, y
SA_HD void b2_coop4_g(uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d, uint64_t x, uint64_t y) {
    SA_B2_G(a, b, c, d, x, y);
}
// the two digest words lane j ends up with: out[j] and out[4 + j]
SA_HD void b2_coop4_final(int j, uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t &lo, uint64_t &hi) {
    const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                            0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                            0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    lo = (iv[j] ^ (j == 0 ? 0x01010040ULL : 0)) ^ a ^ c;
    hi = iv[4 + j] ^ b ^ d;
}

#if defined(__CUDACC__)
// device: all 32 lanes of the warp call this together; lane j = threadIdx.x & 3 of each quad
__device__ __forceinline__ void blake2b_coop4_node(uint64_t &out_lo, uint64_t &out_hi, const uint64_t *msg, int j) {
    uint64_t a, b, c, d;
    b2_coop4_init(j, a, b, c, d, 128u);
#pragma unroll
    for (int r = 0; r < 12; r++) {
        const uint64_t sg = b2_sigma_packed(r) >> (8 * j);
        b2_coop4_g(a, b, c, d, msg[sg & 15], msg[(sg >> 4) & 15]);
        b = __shfl_sync(0xffffffffu, b, (j + 1) & 3, 4);
        c = __shfl_sync(0xffffffffu, c, (j + 2) & 3, 4);
        d = __shfl_sync(0xffffffffu, d, (j + 3) & 3, 4);
        b2_coop4_g(a, b, c, d, msg[(sg >> 32) & 15], msg[(sg >> 36) & 15]);
        b = __shfl_sync(0xffffffffu, b, (j + 3) & 3, 4);
        c = __shfl_sync(0xffffffffu, c, (j + 2) & 3, 4);
        d = __shfl_sync(0xffffffffu, d, (j + 1) & 3, 4);
    }
    b2_coop4_final(j, a, b, c, d, out_lo, out_hi);
}
#endif
// host model of the same four-lane schedule (tests/emu): out = blake2b(msg[0..16))
inline void blake2b_coop4_node_host(uint64_t out[8], const uint64_t msg[16]) {
    uint64_t a[4], b[4], c[4], d[4], t[4];
    for (int j = 0; j < 4; j++) b2_coop4_init(j, a[j], b[j], c[j], d[j], 128u);
    for (int r = 0; r < 12; r++) {
        for (int j = 0; j < 4; j++) {
            const uint64_t sg = b2_sigma_packed(r) >> (8 * j);
            b2_coop4_g(a[j], b[j], c[j], d[j], msg[sg & 15], msg[(sg >> 4) & 15]);
        }
        for (int j = 0; j < 4; j++) t[j] = b[(j + 1) & 3];
        for (int j = 0; j < 4; j++) b[j] = t[j];
        for (int j = 0; j < 4; j++) t[j] = c[(j + 2) & 3];
        for (int j = 0; j < 4; j++) c[j] = t[j];
        for (int j = 0; j < 4; j++) t[j] = d[(j + 3) & 3];
        for (int j = 0; j < 4; j++) d[j] = t[j];
        for (int j = 0; j < 4; j++) {
            const uint64_t sg = b2_sigma_packed(r) >> (8 * j);
            b2_coop4_g(a[j], b[j], c[j], d[j], msg[(sg >> 32) & 15], msg[(sg >> 36) & 15]);
        }
        for (int j = 0; j < 4; j++) t[j] = b[(j + 3) & 3];
        for (int j = 0; j < 4; j++) b[j] = t[j];
        for (int j = 0; j < 4; j++) t[j] = c[(j + 2) & 3];
        for (int j = 0; j < 4; j++) c[j] = t[j];
        for (int j = 0; j < 4; j++) t[j] = d[(j + 1) & 3];
        for (int j = 0; j < 4; j++) d[j] = t[j];
    }
    for (int j = 0; j < 4; j++) b2_coop4_final(j, a[j], b[j], c[j], d[j], out[j], out[4 + j]);
}

}  // namespace sa
