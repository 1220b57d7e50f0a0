// fri_merkle.cuh -- per-thread pieces of the FRI round kernel: split-and-fold
// (code/fri.py:85) and the chunked Merkle reduction (code/merkle.py:6-14).
// __host__ __device__ so tests/emu can drive the same code phase by phase.
#pragma once
#include "field.cuh"
#include "hash.cuh"

namespace sa {

constexpr int MK_THREADS = 256;  // threads per Merkle CTA
constexpr int MK_MAX_IPT_LOG = 3;  // a thread reduces at most 8 bottom nodes privately

// c'[i] = 2^-1 (a + b) + (alpha * 2^-1 * x_i^-1) (a - b)   ==  fri.py:85
//   inv2_m : 2^-1 in Montgomery form
//   t_m    : alpha * 2^-1 * (offset * omega^i)^-1 in Montgomery form
SA_HD fe fri_fold_one(const fe &a, const fe &b, const fe &t_m, const fe &inv2_m) {
    const fe s = fe_montmul(fe_add(a, b), inv2_m);
    const fe d = fe_montmul(fe_sub(a, b), t_m);
    return fe_add(s, d);
}

struct MerkleArgs {
    uint64_t *tree;      // heap layout, 8 words per node
    long long width;     // number of bottom nodes of this launch
    int chunk;           // bottom nodes per CTA (power of two, <= MK_THREADS << ipt_log)
    int ipt_log;         // log2 of the bottom nodes one thread reduces privately (0..3)
    int mode;            // 0: bottom digests already in tree; 1: leaves from `values`; 2: leaves from a fold
    const fe *values;    // mode 1: the codeword (width elements)
    const fe *prev;      // mode 2: the codeword being folded (2 * width elements)
    fe *next;            // mode 2: receives the folded codeword (width elements)
    const fe *xinv;      // mode 2: xinv[i] = omega^-i in Montgomery form, i < width
    fe s_m;              // mode 2: alpha * 2^-1 * offset^-1 in Montgomery form
    fe inv2_m;           // mode 2: 2^-1 in Montgomery form
};

// launch shape for a level of `width` bottom nodes: small levels are latency bound (one node per
// thread, as many CTAs as possible), big ones throughput bound (four leaves and their three
// parents per thread, no barrier in between)
SA_HD void merkle_shape(MerkleArgs &a) {
    if (a.width <= MK_THREADS) {  // one CTA finishes the tree
        a.ipt_log = 0;
        a.chunk = (int)a.width;
    } else if (a.width < (1 << 17)) {  // <= one wave of 256-node CTAs: shortest dependency chain
        a.ipt_log = 0;
        a.chunk = MK_THREADS;
    } else if (a.width < (1 << 18)) {
        a.ipt_log = 1;
        a.chunk = MK_THREADS << 1;
    } else {  // throughput bound: 4 bottom nodes + their 3 parents per thread, barrier free
        a.ipt_log = 2;
        a.chunk = MK_THREADS << 2;
    }
}

// digest of bottom node g of this launch: loads it (mode 0) or hashes the leaf (modes 1, 2) and
// writes it to the tree (and next[] in mode 2)
SA_HD void merkle_bottom(uint64_t d[8], const MerkleArgs &a, long long g) {
    uint64_t *node = a.tree + (a.width + g) * 8;
    if (a.mode == 0) {
        for (int i = 0; i < 8; i++) d[i] = node[i];
        return;
    }
    fe v;
    if (a.mode == 1) {
        v = a.values[g];
    } else {
        const fe t_m = fe_montmul(a.xinv[g], a.s_m);
        v = fri_fold_one(a.prev[g], a.prev[a.width + g], t_m, a.inv2_m);
        a.next[g] = v;
    }
    merkle_leaf_digest(d, v);
    for (int i = 0; i < 8; i++) node[i] = d[i];
}

// Private phase of thread t of CTA blk: reduce its 2^ipt_log bottom nodes to one digest (streaming:
// a slot per height), writing every node it creates to the tree.  root = the subtree digest.
SA_HD void merkle_private(uint64_t root[8], const MerkleArgs &a, long long blk, int t) {
    uint64_t slot[MK_MAX_IPT_LOG][8];
    const int ipt = 1 << a.ipt_log;
    uint64_t d[8];
    for (int j = 0; j < ipt; j++) {
        const long long g = blk * a.chunk + (long long)t * ipt + j;
        merkle_bottom(d, a, g);
        long long idx = a.width + g;
        int h = 0;
        while ((j >> h) & 1) {  // a left sibling of this height is waiting: combine
            uint64_t l[8];
            for (int i = 0; i < 8; i++) l[i] = slot[h][i];
            merkle_node_digest(d, l, d);
            idx >>= 1;
            for (int i = 0; i < 8; i++) a.tree[idx * 8 + i] = d[i];
            h++;
        }
        if (h < a.ipt_log)
            for (int i = 0; i < 8; i++) slot[h][i] = d[i];
    }
    for (int i = 0; i < 8; i++) root[i] = d[i];
}

}  // namespace sa
