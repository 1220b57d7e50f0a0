// pipebench.cu -- issue-rate microbenchmarks of the integer instructions the field arithmetic is made of.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipebench tools/pipebench.cu
// Prints one JSON line per variant: cycles of pipe time per warp instruction per SM sub-partition
// (time * clock * 4 SMSPs * SMs / warp instructions issued).  Every thread runs ILP independent chains
// of the instruction under test; 8 warps per SMSP, so latency is hidden and the number is the issue rate.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

constexpr int ILP = 8;
constexpr int ITERS = 2048;
constexpr int UNROLL = 4;

enum {
    V_WIDE_RRR = 0,   // mad.wide.u32 d, a, b, d          IMAD.WIDE.U32 R, R, R, R
    V_WIDE_IMM,       // mad.wide.u32 d, a, 0xCB800000, d  IMAD.WIDE.U32 R, R, imm, R
    V_WIDE_CONST,     // multiplier from the constant bank (kernel parameter)
    V_WIDE_NOADD,     // mul.wide.u32 d, a, b
    V_WIDE_CARRY,     // mad.lo.cc / madc.hi.cc pairs: IMAD.WIDE.U32(.X) with predicate carries
    V_IMAD_LO,        // mad.lo.u32
    V_IMAD_HI,        // mad.hi.u32
    V_IADD3,          // add.u32 (IADD3)
    V_IADD_CARRY,     // add.cc / addc chains (IADD3.X)
    V_LOP3,           // xor
    V_MIX_WIDE_IADD,  // one IMAD.WIDE + one IADD3 per slot: do the two pipes overlap?
    V_MIX_WIDE_2IADD, // one IMAD.WIDE + two IADD3
    V_MIX_LO_IADD,    // one IMAD + one IADD3
    V_DFMA,           // fma.rn.f64
    V_MIX_WIDE_DFMA,  // one IMAD.WIDE + one DFMA
    V_IND_WIDE_2LOP,  // one IMAD.WIDE (own chain) + two LOP3 on an INDEPENDENT chain: structural overlap of the pipes
    V_IND_WIDE_4LOP,  // ... + four LOP3
    V_IND_WIDEC_2LOP, // IMAD.WIDE with fused 64-bit addend (carry-chain form) + two independent LOP3
    V_IND_LO_2LOP,    // one IMAD (lo) + two independent LOP3 (reference point)
    V_COUNT
};
static const char *NAMES[V_COUNT] = {"imad.wide rrr", "imad.wide imm", "imad.wide const", "imul.wide (no addend)",
                                      "imad.wide carry chain (lo.cc+madc.hi)", "imad lo", "imad hi", "iadd3",
                                      "iadd3.x carry chain", "lop3", "mix wide+iadd3", "mix wide+2 iadd3",
                                      "mix imad.lo+iadd3", "dfma", "mix wide+dfma", "independent: wide + 2 lop3",
                                      "independent: wide + 4 lop3", "independent: wide(carry form) + 2 lop3",
                                      "independent: imad.lo + 2 lop3"};
static const int INSTR_PER_SLOT[V_COUNT] = {1, 1, 1, 1, 1, 1, 1, 2, 4, 2, 2, 3, 2, 1, 2, 3, 5, 3, 3};

template <int V>
__global__ void __launch_bounds__(256) k_pipe(uint64_t *sink, uint32_t mult, int iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc[ILP];
    uint32_t a[ILP], s[ILP];
    double df[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        acc[i] = t * 2654435761u + i;
        a[i] = t * 40503u + 977u * i + 1u;
        s[i] = t + i;
        df[i] = 1.0 + 1e-9 * (t + i);
    }
    const double dm = 1.0000001, da = 1e-12;
    for (int it = 0; it < iters; it += UNROLL) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                // every chain feeds its own result back as the multiplicand / addend, so ptxas cannot
                // strength-reduce a loop-invariant product (it does: acc += 2 * (a * s) ...)
                const uint32_t lo0 = (uint32_t)acc[i], hi0 = (uint32_t)(acc[i] >> 32);
                if (V == V_WIDE_RRR) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(lo0), "r"(s[i]));
                if (V == V_WIDE_IMM) asm volatile("mad.wide.u32 %0, %1, 0xCB800000, %0;" : "+l"(acc[i]) : "r"(lo0));
                if (V == V_WIDE_CONST) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(lo0), "r"(mult));
                if (V == V_WIDE_NOADD) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[i]) : "r"(hi0), "r"(s[i]));
                if (V == V_WIDE_CARRY) {
                    uint32_t lo = lo0, hi = hi0;
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                                 : "+r"(lo), "+r"(hi) : "r"(hi0), "r"(s[i]));
                    acc[i] = ((uint64_t)hi << 32) | lo;
                }
                if (V == V_IMAD_LO) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(s[i]), "r"(mult));
                if (V == V_IMAD_HI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(s[i]), "r"(mult));
                if (V == V_IADD3) {  // a += s; s += a (two dependent IADD3 per slot)
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(s[i]));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(a[i]));
                }
                if (V == V_IADD_CARRY) {  // 64-bit Fibonacci: two IADD3 + two IADD3.X per slot
                    uint32_t lo = lo0, hi = hi0;
                    asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(a[i]), "r"(s[i]));
                    asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(s[i]) : "r"(lo), "r"(hi));
                    acc[i] = ((uint64_t)hi << 32) | lo;
                }
                if (V == V_LOP3) {  // (a | b) ^ c, twice per slot
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x56;" : "+r"(a[i]) : "r"(s[i]), "r"(mult));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x56;" : "+r"(s[i]) : "r"(a[i]), "r"(mult));
                }
                if (V == V_MIX_WIDE_IADD || V == V_MIX_WIDE_2IADD) {
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(lo0), "r"(mult));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(s[i]));
                    if (V == V_MIX_WIDE_2IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(a[i]));
                }
                if (V == V_MIX_LO_IADD) {
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(s[i]), "r"(mult));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(a[i]));
                }
                if (V == V_IND_WIDE_2LOP || V == V_IND_WIDE_4LOP) {
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[i]) : "r"(hi0), "r"(mult));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x56;" : "+r"(a[i]) : "r"(s[i]), "r"(mult));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x56;" : "+r"(s[i]) : "r"(a[i]), "r"(mult));
                    if (V == V_IND_WIDE_4LOP) {
                        asm volatile("lop3.b32 %0, %0, %1, %2, 0x56;" : "+r"(a[i]) : "r"(s[i]), "r"(mult));
                        asm volatile("lop3.b32 %0, %0, %1, %2, 0x56;" : "+r"(s[i]) : "r"(a[i]), "r"(mult));
                    }
                }
                if (V == V_IND_WIDEC_2LOP) {
                    uint32_t lo = lo0, hi = hi0;
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                                 : "+r"(lo), "+r"(hi) : "r"(hi0), "r"(mult));
                    acc[i] = ((uint64_t)hi << 32) | lo;
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x56;" : "+r"(a[i]) : "r"(s[i]), "r"(mult));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x56;" : "+r"(s[i]) : "r"(a[i]), "r"(mult));
                }
                if (V == V_IND_LO_2LOP) {
                    uint32_t lo = lo0;
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo) : "r"(mult), "r"(hi0));
                    acc[i] = ((uint64_t)hi0 << 32) | lo;
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x56;" : "+r"(a[i]) : "r"(s[i]), "r"(mult));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x56;" : "+r"(s[i]) : "r"(a[i]), "r"(mult));
                }
                if (V == V_DFMA || V == V_MIX_WIDE_DFMA) {
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(df[i]) : "d"(dm), "d"(da));
                    if (V == V_MIX_WIDE_DFMA) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(lo0), "r"(mult));
                }
            }
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r += acc[i] + a[i] + s[i] + (uint64_t)df[i];
    if (r == 0x1234567ull) sink[t] = r;
}

template <int V>
static void run(int sms, double clock_ghz, uint64_t *sink) {
    const int blocks = sms * 4, threads = 256;  // 1024 threads per SM = 8 warps per SMSP
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_pipe<V><<<blocks, threads>>>(sink, 0x9E3779B1u, ITERS / 8);
    cudaEventRecord(e0);
    k_pipe<V><<<blocks, threads>>>(sink, 0x9E3779B1u, ITERS);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double slots = (double)ITERS * ILP;                     // per thread
    const double warp_slots = slots * blocks * threads / 32.0;   // whole chip
    const double cycles = ms * 1e-3 * clock_ghz * 1e9;
    const double per_slot = cycles * (sms * 4) / warp_slots;      // cycles per slot per SMSP
    printf("{\"variant\": \"%s\", \"ms\": %.4f, \"cycles_per_slot_per_smsp\": %.3f, \"instr_per_slot\": %d, "
           "\"cycles_per_warp_instr\": %.3f}\n",
           NAMES[V], ms, per_slot, INSTR_PER_SLOT[V], per_slot / INSTR_PER_SLOT[V]);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

template <int V>
static void run_all(int sms, double clock_ghz, uint64_t *sink) {
    run<V>(sms, clock_ghz, sink);
    if constexpr (V + 1 < V_COUNT) run_all<V + 1>(sms, clock_ghz, sink);
}

int main(int argc, char **argv) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) {
        printf("{\"error\": \"no device\"}\n");
        return 1;
    }
    const double clock_ghz = argc > 1 ? atof(argv[1]) : 1.965;  // SM clock under load (bench.py's clocks line)
    uint64_t *sink = nullptr;
    cudaMalloc(&sink, sizeof(uint64_t) * prop.multiProcessorCount * 4 * 256);
    printf("{\"device\": \"%s\", \"sms\": %d, \"assumed_clock_ghz\": %.3f, \"ilp\": %d, \"warps_per_smsp\": 8}\n", prop.name,
           prop.multiProcessorCount, clock_ghz, ILP);
    run_all<0>(prop.multiProcessorCount, clock_ghz, sink);
    cudaFree(sink);
    return 0;
}
