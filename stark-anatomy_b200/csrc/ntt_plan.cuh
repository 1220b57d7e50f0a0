// ntt_plan.cuh -- how an n-point transform (code/ntt.py:3-30) is cut into tile passes.
// Shared by the CUDA host code (sa_b200.cu) and the CPU emulation (tests/emu).
//
//   log_n <= 10 : one pass, every transform is one tile column.
//   log_n 11..20: four-step split n = n1 * n2 (n1 = 2^l1 >= n2 = 2^l2):
//       pass 1: for every column j2, an n1-point transform over j1 (stride n2),
//               times w^(k1*j2) [* n^-1 for intt], written to tmp[k1][j2];
//       pass 2: for every row k1 of tmp, an n2-point transform over j2,
//               written to out[k1 + n1*k2].
//     With j = j1*n2 + j2 and k = k1 + n1*k2:  w^(jk) = (w^n2)^(j1 k1) * w^(j2 k1) * (w^n1)^(j2 k2).
//   log_n 21..26: the same split applied twice, n = n1 * n2 * n3, j = j1*(n2 n3) + j2*n3 + j3,
//       k = k1 + n1*k2 + n1*n2*k3:
//       pass 1: n1-point transforms over j1 (stride n2 n3), times w^(k1*m), m = j2*n3 + j3;
//       pass 2: inside every k1 block, n2-point transforms over j2 (stride n3), times (w^n1)^(k2*j3);
//       pass 3: n3-point transforms over j3 (contiguous), written to out[k1 + n1*k2 + n1*n2*k3].
#pragma once
#include "ntt_tile.cuh"

namespace sa {

constexpr int NTT_MAX_LOG_N = 26;  // the pass-1 twiddle matrix has n entries (1 GiB at 2^26)

struct NttShape {
    int log_n, l1, l2, l3;  // l3 > 0: three passes
};
// force3: split into three digits even when two would do (tests exercise the 3-pass path small)
SA_HD NttShape ntt_shape(int log_n, bool force3 = false) {
    NttShape s;
    s.log_n = log_n;
    s.l3 = 0;
    if (log_n <= 10 && !force3) {
        s.l1 = log_n;
        s.l2 = 0;
    } else if (log_n <= 20 && !force3) {
        s.l2 = log_n / 2;
        s.l1 = log_n - s.l2;
    } else {
        s.l3 = log_n / 3;
        s.l1 = (log_n + 2) / 3;
        s.l2 = log_n - s.l1 - s.l3;
    }
    return s;
}

// cst[k] = w_Rmax^k (Montgomery), Rmax = min(16, L), where wL_m generates the L-point transform
inline void ntt_fill_cst(fe cst[8], const fe &wL_m, int L) {
    const int rmax = L >= 16 ? 16 : L;
    const fe wr = fe_mont_pow_u64(wL_m, (uint64_t)(L / rmax));
    fe acc = fe_mont_one();
    for (int k = 0; k < 8; k++) {
        cst[k] = (k < rmax / 2 || k == 0) ? acc : fe_mont_one();
        acc = fe_montmul(acc, wr);
    }
}

// single pass: `batch` contiguous transforms of n = 2^log_n elements, one tile column each
inline void ntt_fill_common(TileArgs &a) {
    a.in_sb2 = a.out_sb2 = 0;
    a.inner = 1;
    a.has_scale = 0;
    a.scale = fe_mont_one();
    a.twb = nullptr;
    a.twb_stride = 0;
    a.npeer = 0;
    a.mc_out = nullptr;
}
// three-pass split (see the header comment).  tmp: n * batch workspace; out doubles as the first
// intermediate (a tile reads all of its elements before it writes them, so in == out is fine).
inline void ntt_fill_3pass_a(TileArgs &a, const fe *in, fe *mid, const NttShape &s, size_t batch, const fe *tw1,
                             const fe *twb1, const fe cst1[8]) {
    const long long n = 1ll << s.log_n, m = 1ll << (s.l2 + s.l3);
    ntt_fill_common(a);
    a.in = in; a.out = mid; a.tw = tw1;
    a.twb = twb1; a.twb_stride = m;
    a.in_sr = m; a.in_sc = 1; a.in_sb = n;
    a.out_sr = m; a.out_sc = 1; a.out_sb = n;
    a.ncols = (int)m; a.nbatch = (int)batch;
    for (int k = 0; k < 8; k++) a.cst[k] = cst1[k];
}
inline void ntt_fill_3pass_b(TileArgs &a, const fe *mid, fe *tmp, const NttShape &s, size_t batch, const fe *tw2,
                             const fe *twb2, const fe cst2[8]) {
    const long long n1 = 1ll << s.l1, n3 = 1ll << s.l3, m = 1ll << (s.l2 + s.l3);
    ntt_fill_common(a);
    a.in = mid; a.out = tmp; a.tw = tw2;
    a.twb = twb2; a.twb_stride = n3;
    a.in_sr = n3; a.in_sc = 1; a.in_sb = m;   // batch item = (b, k1): contiguous blocks of m
    a.out_sr = n3; a.out_sc = 1; a.out_sb = m;
    a.ncols = (int)n3; a.nbatch = (int)(batch * n1);
    for (int k = 0; k < 8; k++) a.cst[k] = cst2[k];
}
inline void ntt_fill_3pass_c(TileArgs &a, const fe *tmp, fe *out, const NttShape &s, size_t batch, const fe *tw3,
                             const fe cst3[8]) {
    const long long n = 1ll << s.log_n, n1 = 1ll << s.l1, n2 = 1ll << s.l2, n3 = 1ll << s.l3, m = n2 * n3;
    ntt_fill_common(a);
    a.in = tmp; a.out = out; a.tw = tw3;
    // batch item = (b, k2); column = k1 (adjacent k1 -> adjacent outputs); row = j3 / k3
    a.in_sr = 1; a.in_sc = m; a.in_sb = n; a.in_sb2 = n3;
    a.out_sr = n1 * n2; a.out_sc = 1; a.out_sb = n; a.out_sb2 = n1;
    a.inner = (int)n2;
    a.ncols = (int)n1; a.nbatch = (int)(batch * n2);
    for (int k = 0; k < 8; k++) a.cst[k] = cst3[k];
}

inline void ntt_fill_single(TileArgs &a, const fe *in, fe *out, int log_n, size_t batch, const fe *tw,
                            const fe cst[8], int has_scale, const fe &scale_m) {
    const long long n = 1ll << log_n;
    ntt_fill_common(a);
    a.in = in;
    a.out = out;
    a.tw = tw;
    a.twb = nullptr;
    a.twb_stride = 0;
    a.in_sr = 1; a.in_sc = n; a.in_sb = 0;
    a.out_sr = 1; a.out_sc = n; a.out_sb = 0;
    a.ncols = (int)batch;
    a.nbatch = 1;
    a.has_scale = has_scale;
    a.scale = scale_m;
    for (int k = 0; k < 8; k++) a.cst[k] = cst[k];
}
inline void ntt_fill_pass1(TileArgs &a, const fe *in, fe *tmp, const NttShape &s, size_t batch, const fe *tw1,
                           const fe *twb, const fe cst1[8]) {
    const long long n = 1ll << s.log_n, n2 = 1ll << s.l2;
    ntt_fill_common(a);
    a.in = in;
    a.out = tmp;
    a.tw = tw1;
    a.twb = twb;
    a.twb_stride = n2;
    a.in_sr = n2; a.in_sc = 1; a.in_sb = n;
    a.out_sr = n2; a.out_sc = 1; a.out_sb = n;
    a.ncols = (int)n2;
    a.nbatch = (int)batch;
    a.has_scale = 0;
    a.scale = fe_mont_one();
    for (int k = 0; k < 8; k++) a.cst[k] = cst1[k];
}
inline void ntt_fill_pass2(TileArgs &a, const fe *tmp, fe *out, const NttShape &s, size_t batch, const fe *tw2,
                           const fe cst2[8]) {
    const long long n = 1ll << s.log_n, n1 = 1ll << s.l1, n2 = 1ll << s.l2;
    ntt_fill_common(a);
    a.in = tmp;
    a.out = out;
    a.tw = tw2;
    a.twb = nullptr;
    a.twb_stride = 0;
    a.in_sr = 1; a.in_sc = n2; a.in_sb = n;      // column = k1 (a row of tmp), row = j2
    a.out_sr = n1; a.out_sc = 1; a.out_sb = n;   // out[k1 + n1 * k2]
    a.ncols = (int)n1;
    a.nbatch = (int)batch;
    a.has_scale = 0;
    a.scale = fe_mont_one();
    for (int k = 0; k < 8; k++) a.cst[k] = cst2[k];
}

}  // namespace sa
