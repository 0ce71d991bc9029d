// fp64_issue_overlap.cu -- does FP64 pipe time hide the issue slots of everything else on a B200 SM sub-partition?
//
// The hex8 operator kernel issues about one non-FP64 instruction (LDS, SHFL, IMAD / IADD3 / LOP3, STG, mbarrier) per FP64
// instruction.  Its time per x step fits "FP64 pipe time + issue slots of the rest" much better than "max of the two"
// (DESIGN 3.1).  This microbenchmark measures that directly: per loop trip a warp issues NF independent DFMAs and NO
// independent instructions of another class (integer multiply-add, logic / add, shared-memory load, warp shuffle); the three
// timings -- FP64 only, the other class only, both interleaved -- say how much of the cheaper stream hides behind the other.
//     overlap = (t_fp64 + t_other - t_both) / min(t_fp64, t_other)      1 = perfect overlap (max model), 0 = none (sum model)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_issue_overlap fp64_issue_overlap.cu
#include <cstdio>
#include <cuda_runtime.h>

enum { OTHER_IMAD = 0, OTHER_LOGIC = 1, OTHER_LDS = 2, OTHER_SHFL = 3, OTHER_LOP3 = 4, OTHER_MINMAX = 5, OTHER_LEA = 6, OTHER_SHF = 7, OTHER_SETP_SEL = 8 };

template <int NF, int NO, int OTHER, bool DO_F, bool DO_O>
__global__ void k_mix(double* out, int iters, double a, double b, int seed) {
    __shared__ double sh[1024 + 64];
    double acc[8];
    unsigned x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc[i] = threadIdx.x * 1e-9 + i;
        x[i] = threadIdx.x * 7u + i + seed;
    }
    for (int i = threadIdx.x; i < 1024 + 64; i += blockDim.x) sh[i] = i;
    __syncthreads();
    const unsigned shp = (unsigned)__cvta_generic_to_shared(sh + (threadIdx.x & 1023));
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (DO_F) {
#pragma unroll
                for (int i = 0; i < NF; ++i) acc[i] = fma(acc[i], a, b);
            }
            if (DO_O) {
#pragma unroll
                for (int i = 0; i < NO; ++i) {
                    if (OTHER == OTHER_IMAD) x[i] = x[i] * x[i] + 12345u;          // one IMAD, not collapsible across trips
                    if (OTHER == OTHER_LOGIC) x[i] = x[i] + x[(i + 1) & 7];        // one IADD3
                    if (OTHER == OTHER_LDS) {                                       // one LDS.64, consumed by an empty asm
                        double v;
                        asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(shp + 8 * ((i + r) & 7)) : "memory");
                        asm volatile("" ::"d"(v));
                    }
                    if (OTHER == OTHER_SHFL) x[i] = __shfl_up_sync(0xffffffffu, x[i], 1);
                    if (OTHER == OTHER_LOP3) x[i] = x[i] ^ (x[(i + 1) & 7] & x[(i + 3) & 7]);   // one LOP3
                    if (OTHER == OTHER_LEA) x[i] = (x[i] << 3) + x[(i + 1) & 7];                  // LEA
                    if (OTHER == OTHER_SHF) x[i] = __funnelshift_l(x[i], x[(i + 1) & 7], 5);      // SHF
                    if (OTHER == OTHER_SETP_SEL) x[i] = ((int)x[i] < it) ? x[(i + 1) & 7] : x[(i + 3) & 7];  // ISETP + SEL
                    if (OTHER == OTHER_MINMAX) x[i] = max(x[i] + 1u, x[(i + 1) & 7]);            // compare / select class
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i] + (double)x[i];
    if (s == 12345.678) out[0] = s;
}

// the operator kernel takes its 45 mode-space constants from kernel parameters: every use is an LDCU (uniform load from the constant bank)
// feeding a DFMA with a uniform-register operand.  Same DFMA count with register operands (CONST = false) or parameter operands (true).
struct ConstBlock {
    double c[48];
};
template <bool CONST>
__global__ void k_const(double* out, int iters, const __grid_constant__ ConstBlock blk, double a, double b) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-9 + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = CONST ? fma(acc[i], blk.c[(r * 8 + i) % 45], b) : fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

template <int NF, int NO, int OTHER>
static void run(const char* name, int sms, int threads, double* d, int iters, bool last) {
    const float tf = time_ms([&] { k_mix<NF, NO, OTHER, true, false><<<sms, threads>>>(d, iters, 1.0000001, 1e-9, 1); });
    const float to = time_ms([&] { k_mix<NF, NO, OTHER, false, true><<<sms, threads>>>(d, iters, 1.0000001, 1e-9, 1); });
    const float tb = time_ms([&] { k_mix<NF, NO, OTHER, true, true><<<sms, threads>>>(d, iters, 1.0000001, 1e-9, 1); });
    const float mn = tf < to ? tf : to;
    printf(" \"%s, %d warps per scheduler, %d DFMA + %d other per trip\": {\"ms_fp64_only\": %.3f, \"ms_other_only\": %.3f, \"ms_both\": %.3f, "
           "\"overlap\": %.2f}%s\n",
           name, threads / 128, NF, NO, tf, to, tb, (tf + to - tb) / mn, last ? "" : ",");
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* d;
    cudaMalloc(&d, 64);
    const int iters = 20000;
    printf("{\"device\": \"%s\", \"sms\": %d, \"iters\": %d, \"trip\": \"8 x (NF DFMA + NO other), independent chains\",\n", p.name, sms, iters);
    for (int threads : {256, 512, 1024}) {
        if (threads == 256) {
            run<8, 8, OTHER_IMAD>("IMAD", sms, 256, d, iters, false);
            run<8, 8, OTHER_LOGIC>("IADD3", sms, 256, d, iters, false);
            run<8, 8, OTHER_LDS>("LDS.64", sms, 256, d, iters, false);
            run<8, 8, OTHER_SHFL>("SHFL", sms, 256, d, iters, false);
            run<8, 4, OTHER_LOGIC>("IADD3", sms, 256, d, iters, false);
            run<8, 8, OTHER_LOP3>("LOP3", sms, 256, d, iters, false);
            run<8, 4, OTHER_LOP3>("LOP3", sms, 256, d, iters, false);
            run<8, 8, OTHER_MINMAX>("VIADD+VIMNMX", sms, 256, d, iters, false);
            run<8, 4, OTHER_LEA>("LEA", sms, 256, d, iters, false);
            run<8, 4, OTHER_SHF>("SHF", sms, 256, d, iters, false);
            run<8, 4, OTHER_SETP_SEL>("ISETP+SEL", sms, 256, d, iters, false);
            {
                ConstBlock blk;
                for (int i = 0; i < 48; ++i) blk.c[i] = 1.0000001 + 1e-9 * i;
                const float tr = time_ms([&] { k_const<false><<<sms, 256>>>(d, iters, blk, 1.0000001, 1e-9); });
                const float tc = time_ms([&] { k_const<true><<<sms, 256>>>(d, iters, blk, 1.0000001, 1e-9); });
                printf(" \"64 DFMA per trip with register operands / with kernel-parameter operands (LDCU + uniform register), 2 warps per scheduler\": "
                       "{\"ms_register\": %.3f, \"ms_constant\": %.3f},\n", tr, tc);
            }
        } else if (threads == 512) {
            run<8, 8, OTHER_IMAD>("IMAD", sms, 512, d, iters, false);
            run<8, 8, OTHER_LOGIC>("IADD3", sms, 512, d, iters, false);
            run<8, 8, OTHER_LDS>("LDS.64", sms, 512, d, iters, false);
        } else {
            run<8, 8, OTHER_IMAD>("IMAD", sms, 1024, d, iters, false);
            run<8, 8, OTHER_LOGIC>("IADD3", sms, 1024, d, iters, false);
            run<8, 8, OTHER_LDS>("LDS.64", sms, 1024, d, iters, false);
            run<8, 8, OTHER_SHFL>("SHFL", sms, 1024, d, iters, true);
        }
    }
    printf("}\n");
    return 0;
}
