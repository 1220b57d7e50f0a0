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
#pragma once
#include "ntt_tile.cuh"

namespace sa {

struct NttShape {
    int log_n, l1, l2;
};
SA_HD NttShape ntt_shape(int log_n) {
    NttShape s;
    s.log_n = log_n;
    if (log_n <= 10) {
        s.l1 = log_n;
        s.l2 = 0;
    } else {
        s.l2 = log_n / 2;
        s.l1 = log_n - s.l2;
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
inline void ntt_fill_single(TileArgs &a, const fe *in, fe *out, int log_n, size_t batch, const fe *tw,
                            const fe cst[8], int has_scale, const fe &scale_m) {
    const long long n = 1ll << log_n;
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
