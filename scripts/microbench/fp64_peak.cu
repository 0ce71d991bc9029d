// fp64_peak.cu -- one-minute FP64 roof of a B200: DFMA / DADD vector pipe, DMMA (mma.sync m8n8k4 f64), both together,
// and the DFMA rate as a function of warps per SM and independent chains per thread (latency / occupancy the operator
// kernel needs).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>  // OP 0: DFMA, 1: DADD, 2: alternating DADD/DFMA
__global__ void k_vec(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) acc[i] = fma(acc[i], a, b);
                if (OP == 1) acc[i] = acc[i] + a;
                if (OP == 2) acc[i] = (r & 1) ? fma(acc[i], a, b) : acc[i] + a;
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

// DMMA m8n8k4: A 8x4 (1 reg), B 4x8 (1 reg), C/D 8x8 (2 regs) per thread; 512 flop per warp instruction
template <int NACC, bool WITH_DFMA>
__global__ void k_dmma(double* out, int iters, double a, double b) {
    double c[NACC][2];
    double x[4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = i + threadIdx.x * 1e-9;
    double fa = a, fb = b;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < NACC; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1])
                             : "d"(fa), "d"(fb));
            if (WITH_DFMA) {
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = fma(x[i], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz_attr\": %.0f,\n", p.name, sms, clk_khz / 1e3);
    double* d;
    cudaMalloc(&d, 64);
    const int iters = 20000;
    // ---- vector pipe peaks: 1024 threads/SM x 2 CTAs, 8 chains
    auto vec = [&](const char* name, auto kern, int ilp, int threads, int ctas_per_sm, double flop_per_op) {
        float ms = time_ms([&] { kern<<<sms * ctas_per_sm, threads>>>(d, iters, 1.0000001, 1e-9); });
        double ops = (double)sms * ctas_per_sm * threads * iters * 8.0 * ilp;
        printf(" \"%s\": {\"ms\": %.3f, \"tera_ops\": %.3f, \"tflops\": %.3f, \"lane_ops_per_clk_per_sm_at_attr_clock\": %.2f},\n", name, ms, ops / ms / 1e9,
               ops * flop_per_op / ms / 1e9, ops / (ms * 1e-3) / sms / (clk_khz * 1e3));
    };
    vec("dfma_ilp8_1024x2", k_vec<8, 0>, 8, 1024, 2, 2);
    vec("dadd_ilp8_1024x2", k_vec<8, 1>, 8, 1024, 2, 1);
    vec("dadd_dfma_mix_ilp8_1024x2", k_vec<8, 2>, 8, 1024, 2, 1.5);
    // ---- occupancy / ILP sweep of DFMA (what the operator kernel needs to keep the pipe busy)
    vec("dfma_ilp1_128x1 (1 warp/SMSP)", k_vec<1, 0>, 1, 128, 1, 2);
    vec("dfma_ilp2_128x1", k_vec<2, 0>, 2, 128, 1, 2);
    vec("dfma_ilp4_128x1", k_vec<4, 0>, 4, 128, 1, 2);
    vec("dfma_ilp8_128x1", k_vec<8, 0>, 8, 128, 1, 2);
    vec("dfma_ilp1_256x1 (2 warps/SMSP)", k_vec<1, 0>, 1, 256, 1, 2);
    vec("dfma_ilp2_256x1", k_vec<2, 0>, 2, 256, 1, 2);
    vec("dfma_ilp4_256x1", k_vec<4, 0>, 4, 256, 1, 2);
    vec("dfma_ilp1_512x1 (4 warps/SMSP)", k_vec<1, 0>, 1, 512, 1, 2);
    vec("dfma_ilp2_512x1", k_vec<2, 0>, 2, 512, 1, 2);
    vec("dfma_ilp1_1024x1 (8 warps/SMSP)", k_vec<1, 0>, 1, 1024, 1, 2);
    // ---- DMMA
    auto mm = [&](const char* name, auto kern, int nacc, int threads, int ctas_per_sm, bool with_dfma) {
        float ms = time_ms([&] { kern<<<sms * ctas_per_sm, threads>>>(d, iters, 1.0000001, 1e-9); });
        double warps = (double)sms * ctas_per_sm * threads / 32;
        double mma_flop = warps * iters * 8.0 * nacc * 512.0;
        double dfma_flop = with_dfma ? warps * 32 * iters * 8.0 * 4 * 2.0 : 0.0;
        printf(" \"%s\": {\"ms\": %.3f, \"dmma_tflops\": %.3f, \"dfma_tflops\": %.3f, \"sum_tflops\": %.3f},\n", name, ms, mma_flop / ms / 1e9,
               dfma_flop / ms / 1e9, (mma_flop + dfma_flop) / ms / 1e9);
    };
    mm("dmma_m8n8k4_acc4_1024x2", k_dmma<4, false>, 4, 1024, 2, false);
    mm("dmma_m8n8k4_acc8_512x2", k_dmma<8, false>, 8, 512, 2, false);
    mm("dmma_acc4_plus_dfma4_1024x2 (same warps)", k_dmma<4, true>, 4, 1024, 2, true);
    printf(" \"note\": \"nominal B200 FP64 vector = 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s; lane_ops/clk/SM of 64 = full rate\"}\n");
    return 0;
}
