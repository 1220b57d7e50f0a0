// ntt_tile.cuh -- the NTT building block: L-point transforms (L = 2^LOGL <= 1024)
// on a tile of 8 columns, natural order in, natural order out.
// Reproduces code/ntt.py:3-18 (out[i] = sum_j v[j] * w^(i*j)); bit-exact because
// every operation is exact arithmetic on canonical residues (field.cuh).
//
// Algorithm (Blackwell-first, not the reference's recursion): mixed-radix
// decimation with radix-16 register blocks.  A thread owns 16 elements of one
// column, runs a complete 16-point (or 2/4/8-point) transform in registers,
// applies the inter-stage twiddle, and exchanges through shared memory only
// between register blocks (two exchanges for L = 1024).  All stages work in
// place on tile row p = K*M' + d*M + m (d = the digit being transformed), so the
// first block loads straight from global memory into registers and the last one
// stores straight from registers, digit-reversed, to the natural output index.
//
// A tile row is 8 columns x 16 bytes = one 128-byte line: lanes 0-7 of every
// quarter warp cover a full row, so every shared-memory access is conflict free
// and every global access is a full 128-byte (or 64-byte, see ntt.cu) segment.
//
// The functions are __host__ __device__: tests/emu runs exactly this code on the
// CPU, thread by thread and phase by phase.
#pragma once
#include "field.cuh"

namespace sa {

constexpr int TILE_C = 8;

struct TileArgs {
    const fe *in;
    fe *out;
    const fe *tw;   // tw[e] = w_L^e (Montgomery form), e < L
    const fe *twb;  // optional output twiddle matrix [k * twb_stride + col]; nullptr = none
    long long in_sr, in_sc, in_sb;     // element strides of (row, column, batch) on input
    long long out_sr, out_sc, out_sb;  // ... and on output
    long long twb_stride;
    int ncols;   // valid columns per batch item
    int nbatch;  // batch items
    int has_scale;
    fe scale;   // Montgomery-form scalar applied to every output when has_scale
    fe cst[8];  // cst[k] = w_Rmax^k (Montgomery form), Rmax = min(16, L), k < Rmax/2
};

template <int LOGL>
struct TilePlan {
    static constexpr int L = 1 << LOGL;
    static constexpr int NST = (LOGL + 3) / 4;
    static constexpr int E = L >= 16 ? 16 : L;       // elements per thread
    static constexpr int TPT = (L / E) * TILE_C;     // threads per tile
    static constexpr int TPC = TPT >= 128 ? 1 : 128 / TPT;  // tiles per CTA
    static constexpr int RMAX = L >= 16 ? 16 : L;
    // log2 radix of stage i: as many radix-16 blocks as fit, then the remainder
    SA_HDC int rlog(int i) { return i < LOGL / 4 ? 4 : LOGL % 4; }
    // log2 of M_i = L / (R_0 * ... * R_i)
    SA_HDC int mlog(int i) {
        int s = 0;
        for (int j = 0; j <= i; j++) s += rlog(j);
        return LOGL - s;
    }
    SA_HDC size_t smem_bytes() { return NST > 1 ? (size_t)TPC * L * TILE_C * sizeof(fe) : 0; }
};

SA_HDC int tile_bitrev(int i, int r) {
    int j = 0;
    for (int b = 1, bb = r >> 1; b < r; b <<= 1, bb >>= 1)
        if (i & b) j |= bb;
    return j;
}

// one radix-2 decimation-in-time level (span LEN) of an R-point transform held in registers
template <int R, int LEN>
SA_HD void dft_level(fe *x, const fe *cst, int cstep) {
    constexpr int HALF = LEN / 2, STEP = R / LEN;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int g = 0; g < R; g += LEN) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < HALF; k++) {
            const fe t = (k == 0) ? x[g + k + HALF] : fe_montmul(x[g + k + HALF], cst[k * STEP * cstep]);
            const fe e = x[g + k];
            x[g + k] = fe_add(e, t);
            x[g + k + HALF] = fe_sub(e, t);
        }
    }
    if constexpr (LEN < R) dft_level<R, LEN * 2>(x, cst, cstep);
}

// natural-order R-point transform of x[0..R) in registers; w_R^j = cst[j * cstep]
template <int R>
SA_HD void dft_regs(fe *x, const fe *cst, int cstep) {
    // bit-reversal permutation (register renaming once unrolled)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < R; i++) {
        const int j = tile_bitrev(i, R);
        if (i < j) {
            const fe tmp = x[i];
            x[i] = x[j];
            x[j] = tmp;
        }
    }
    if constexpr (R >= 2) dft_level<R, 2>(x, cst, cstep);
}

template <int LOGL>
SA_HD int tile_digit_reverse(int p) {
    using P = TilePlan<LOGL>;
    int o = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < P::NST; i++) {
        const int k = (p >> P::mlog(i)) & ((1 << P::rlog(i)) - 1);
        o |= k << (LOGL - P::mlog(i) - P::rlog(i));
    }
    return o;
}

SA_HD fe tile_ld(const fe *p) {
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    return fe_make(v.x, v.y, v.z, v.w);
#else
    return *p;
#endif
}
SA_HD fe tile_ldg(const fe *p) {
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    return fe_make(v.x, v.y, v.z, v.w);
#else
    return *p;
#endif
}
SA_HD void tile_st(fe *p, const fe &x) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4 *>(p) = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
#else
    *p = x;
#endif
}

// One register-block stage for thread t (0 <= t < TPT) of one tile.
//   sm    : this tile's shared rows, L x 8 elements
//   b,col0: batch item and first column of the tile; valid = tile exists
template <int LOGL, int ST>
SA_HD void ntt_tile_stage(int t, fe *sm, const TileArgs &a, long long b, int col0, bool valid) {
    using P = TilePlan<LOGL>;
    constexpr int RL = P::rlog(ST), R = 1 << RL, ML = P::mlog(ST), M = 1 << ML;
    constexpr int U = P::E / R;  // units (independent R-point transforms) per thread
    constexpr bool FIRST = ST == 0, LAST = ST == P::NST - 1;
    constexpr int WLOG = LOGL - ML - RL;  // log2 of the product of the earlier radices
    constexpr int CSTEP = P::RMAX / R;
    const int c = t & (TILE_C - 1), q = t >> 3;
    const int col = col0 + c;
    const bool active = valid && col < a.ncols;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int s = 0; s < U; s++) {
        const int u = q * U + s;
        const int K = u >> ML, m = u & (M - 1);
        const int row0 = (K << (ML + RL)) + m;
        fe x[R];
        if (FIRST) {
            const fe *src = a.in + b * a.in_sb + (long long)col * a.in_sc;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int d = 0; d < R; d++)
                x[d] = active ? tile_ld(src + (long long)(row0 + d * M) * a.in_sr) : fe_zero();
        } else {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int d = 0; d < R; d++) x[d] = tile_ld(sm + (row0 + d * M) * TILE_C + c);
        }
        dft_regs<R>(x, a.cst, CSTEP);
        if (!LAST) {
            // multiply output k by w_{M*R}^(k*m) = w_L^((k*m) << WLOG), then park it in row row0 + k*M
            tile_st(sm + row0 * TILE_C + c, x[0]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 1; k < R; k++) {
                const fe w = tile_ldg(a.tw + ((k * m) << WLOG));
                tile_st(sm + (row0 + k * M) * TILE_C + c, fe_montmul(x[k], w));
            }
        } else {
            fe *dst = a.out + b * a.out_sb + (long long)col * a.out_sc;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < R; k++) {
                const int o = tile_digit_reverse<LOGL>(row0 + k * M);
                fe v = x[k];
                if (a.twb != nullptr && active) v = fe_montmul(v, tile_ldg(a.twb + (long long)o * a.twb_stride + col));
                if (a.has_scale) v = fe_montmul(v, a.scale);
                if (active) tile_st(dst + (long long)o * a.out_sr, v);
            }
        }
    }
}

}  // namespace sa
