// tma_rows.cu -- what the operator kernel's producer warp relies on, checked on the device:
//  (1) correctness of 1-D tiled TMA boxes (34 doubles = 272 B) that start at ARBITRARY element offsets of a vector whose
//      rows are only 8-byte aligned (nz + 1 odd), incl. negative and past-the-end coordinates (zero fill);
//  (2) throughput of streaming such row boxes (51 per node plane) versus one 3-D box per plane on a z-padded layout,
//      one CTA per SM, 4-slot ring, consumers only touch the data.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_rows tma_rows.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    return (EncodeTiled)fn;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(void* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(void* b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const CUtensorMap* tm, int c0, void* bar) {
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(smem_u32(dst)), "l"(tm),
                 "r"(c0), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, void* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// non-tensor bulk copy: src and dst 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, void* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int BOX = 34;

// (1) correctness: one box per listed coordinate
__global__ void k_check(const __grid_constant__ CUtensorMap tm, const int* coords, int n, double* out) {
    __shared__ __align__(128) double buf[BOX + 2];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int i = 0; i < n; ++i) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, BOX * 8);
            tma_load_1d(buf, &tm, coords[i], &bar);
        }
        mbar_wait(&bar, i & 1);
        if (threadIdx.x < BOX) out[i * BOX + threadIdx.x] = buf[threadIdx.x];
        __syncthreads();
    }
}

// (2) throughput.  MODE 0: 51 row boxes (1-D) per plane, MODE 1: one 3-D box {34, 17, 3} per plane (padded layout)
constexpr int ROWS = 17, SLOTS = 4, NCONS = 16;
struct Ring {
    double in[SLOTS][3][ROWS][BOX];
    unsigned long long full[SLOTS], empty[SLOTS];
};
template <int MODE>
__global__ void __launch_bounds__(32 * (NCONS + 1), 1) k_stream(const __grid_constant__ CUtensorMap tm, int nplanes, int ny1, int nzs, int nseg,
                                                                 int ytiles, double* sink) {
    extern __shared__ __align__(128) unsigned char raw[];
    Ring& R = *reinterpret_cast<Ring*>(raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(&R.full[s], 1);
            mbar_init(&R.empty[s], NCONS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ytile = blockIdx.x % ytiles, xchunk = blockIdx.x / ytiles;
    const int Y0 = ytile * (ROWS - 1);
    const int p0 = xchunk * nplanes;
    const int total = nseg * nplanes;
    if (warp == NCONS) {  // producer
        for (int q = 0; q < total; ++q) {
            const int slot = q & (SLOTS - 1), seg = q / nplanes, pl = p0 + q % nplanes;
            if (q >= SLOTS) mbar_wait(&R.empty[slot], ((q / SLOTS) - 1) & 1);
            if (lane == 0) mbar_expect_tx(&R.full[slot], 3 * ROWS * BOX * 8);
            __syncwarp();
            if (MODE == 0) {
                for (int r = lane; r < 3 * ROWS; r += 32) {
                    const int c = r / ROWS, jr = r % ROWS;
                    const long long e = ((long long)(pl * 3 + c) * ny1 + (Y0 + jr)) * nzs + seg * 32;
                    tma_load_1d(&R.in[slot][c][jr][0], &tm, (int)e, &R.full[slot]);
                }
            } else if (lane == 0) {
                tma_load_3d(&R.in[slot][0][0][0], &tm, seg * 32, Y0, pl * 3, &R.full[slot]);
            }
        }
    } else {
        double acc = 0;
        for (int q = 0; q < total; ++q) {
            const int slot = q & (SLOTS - 1);
            mbar_wait(&R.full[slot], (q / SLOTS) & 1);
            acc += R.in[slot][0][warp][lane] + R.in[slot][1][warp + 1][lane + 1] + R.in[slot][2][warp][lane + 1];
            __syncwarp();
            if (lane == 0) mbar_arrive(&R.empty[slot]);
        }
        if (acc == 1.2345) sink[0] = acc;
    }
}


// generic single-box check: DIM-d box of `box_bytes` bytes at the given coordinates, copied out as doubles
template <int DIM>
__global__ void k_box(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int box_bytes, double* out) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, box_bytes);
        if (DIM == 1) tma_load_1d(raw, &tm, c0, &bar);
        if (DIM == 2) tma_load_2d(raw, &tm, c0, c1, &bar);
        if (DIM == 3) tma_load_3d(raw, &tm, c0, c1, c2, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < box_bytes / 8; i += blockDim.x) out[i] = reinterpret_cast<double*>(raw)[i];
}
__global__ void k_bulk(const double* src, int bytes, int dst_off_bytes, double* out) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, bytes);
        bulk_load(raw + dst_off_bytes, src, bytes, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes / 8; i += blockDim.x) out[i] = reinterpret_cast<double*>(raw + dst_off_bytes)[i];
}
// streaming with non-tensor bulk row copies on a padded layout (rows 16-byte aligned): 51 rows of 272 B per plane
__global__ void __launch_bounds__(32 * (NCONS + 1), 1) k_stream_bulk(const double* v, int nplanes, int ny1, int nzs, int nseg, int ytiles, double* sink) {
    extern __shared__ __align__(128) unsigned char raw[];
    Ring& R = *reinterpret_cast<Ring*>(raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(&R.full[s], 1);
            mbar_init(&R.empty[s], NCONS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ytile = blockIdx.x % ytiles, xchunk = blockIdx.x / ytiles;
    const int Y0 = ytile * (ROWS - 1), p0 = xchunk * nplanes, total = nseg * nplanes;
    if (warp == NCONS) {
        for (int q = 0; q < total; ++q) {
            const int slot = q & (SLOTS - 1), seg = q / nplanes, pl = p0 + q % nplanes;
            if (q >= SLOTS) mbar_wait(&R.empty[slot], ((q / SLOTS) - 1) & 1);
            if (lane == 0) mbar_expect_tx(&R.full[slot], 3 * ROWS * BOX * 8);
            __syncwarp();
            for (int r = lane; r < 3 * ROWS; r += 32) {
                const int c = r / ROWS, jr = r % ROWS;
                const long long e = ((long long)(pl * 3 + c) * ny1 + (Y0 + jr)) * nzs + seg * 32;
                bulk_load(&R.in[slot][c][jr][0], v + e, BOX * 8, &R.full[slot]);
            }
        }
    } else {
        double acc = 0;
        for (int q = 0; q < total; ++q) {
            const int slot = q & (SLOTS - 1);
            mbar_wait(&R.full[slot], (q / SLOTS) & 1);
            acc += R.in[slot][0][warp][lane] + R.in[slot][1][warp + 1][lane + 1] + R.in[slot][2][warp][lane + 1];
            __syncwarp();
            if (lane == 0) mbar_arrive(&R.empty[slot]);
        }
        if (acc == 1.2345) sink[0] = acc;
    }
}

static int verdict(const char* name, cudaError_t e, int bad, int total) {
    printf("{\"variant\": \"%s\", \"cuda_error\": \"%s\", \"mismatches\": %d, \"checked\": %d}\n", name, cudaGetErrorString(e), bad, total);
    return 0;
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    EncodeTiled enc = get_encode();
    if (!enc) return 1;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int NY1 = 257, planes = 6;
    const int NZS = (variant == 5 || variant == 6 || variant == 10 || variant == 11) ? 260 : 257;
    const long long n = (long long)planes * NY1 * NZS;
    std::vector<double> h(n);
    for (long long i = 0; i < n; ++i) h[i] = 1.0 + (double)i;
    double *d, *dout;
    cudaMalloc(&d, n * 8 + 4096);
    cudaMalloc(&dout, 1 << 20);
    cudaMemset(dout, 0, 1 << 20);
    cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
    alignas(64) CUtensorMap tm;
    std::vector<double> ho(1 << 17);
    auto fetch = [&]() { cudaMemcpy(ho.data(), dout, 1 << 20, cudaMemcpyDeviceToHost); };
    auto enc_nd = [&](int rank, CUtensorMapDataType dt, cuuint64_t* dim, cuuint64_t* stride, cuuint32_t* box) {
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tm, dt, rank, d, dim, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) printf("{\"encode_error\": %d}\n", (int)r);
        return r;
    };
    if (variant == 0) {  // sanity: 2-D f32 box 64 x 8 on a [rows][1024-float] view
        cuuint64_t dim[2] = {1024, (cuuint64_t)(n * 2 / 1024)}, stride[1] = {4096};
        cuuint32_t box[2] = {64, 8};
        if (enc_nd(2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, dim, stride, box)) return 0;
        k_box<2><<<1, 64, 4096>>>(tm, 64, 3, 0, 64 * 8 * 4, dout);
        cudaError_t e = cudaDeviceSynchronize();
        fetch();
        int bad = 0;
        for (int r = 0; r < 8; ++r)
            for (int k = 0; k < 32; ++k) bad += ho[r * 32 + k] != h[((3 + r) * 1024 + 64) / 2 + k];
        return verdict("2d_f32_64x8_sanity", e, bad, 256);
    }
    if (variant == 1 || variant == 2 || variant == 3 || variant == 4) {
        const int coord = variant == 2 ? 259 : (variant == 3 ? 261 : (variant == 4 ? 2 * 263 : 0));
        if (variant == 3) {  // rank-2 view {N, 1}
            cuuint64_t dim[2] = {(cuuint64_t)n, 1}, stride[1] = {(cuuint64_t)((n * 8 + 15) / 16 * 16)};
            cuuint32_t box[2] = {BOX, 1};
            if (enc_nd(2, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, dim, stride, box)) return 0;
            k_box<2><<<1, 64, 4096>>>(tm, coord, 0, 0, BOX * 8, dout);
        } else if (variant == 4) {  // the doubles seen as pairs of floats: box 68 floats
            cuuint64_t dim[1] = {(cuuint64_t)n * 2}, stride[1] = {0};
            cuuint32_t box[1] = {2 * BOX};
            if (enc_nd(1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, dim, stride, box)) return 0;
            k_box<1><<<1, 64, 4096>>>(tm, coord, 0, 0, BOX * 8, dout);
        } else {
            cuuint64_t dim[1] = {(cuuint64_t)n}, stride[1] = {0};
            cuuint32_t box[1] = {BOX};
            if (enc_nd(1, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, dim, stride, box)) return 0;
            k_box<1><<<1, 64, 4096>>>(tm, coord, 0, 0, BOX * 8, dout);
        }
        cudaError_t e = cudaDeviceSynchronize();
        fetch();
        int bad = 0;
        const long long g0 = variant == 4 ? coord / 2 : coord;
        for (int k = 0; k < BOX; ++k) bad += ho[k] != h[g0 + k];
        const char* names[5] = {"", "1d_f64_aligned_start", "1d_f64_odd_start", "2d_f64_Nx1_odd_start", "1d_f32_pairs_odd_start"};
        return verdict(names[variant], e, bad, BOX);
    }
    if (variant == 5) {  // 3-D box {34, 17, 3} on the z-padded layout, start (32, 15, 3), and a clipped one at (-1.., 256-ish)
        cuuint64_t dim[3] = {(cuuint64_t)NZS, (cuuint64_t)NY1, (cuuint64_t)planes}, stride[2] = {(cuuint64_t)NZS * 8, (cuuint64_t)NZS * NY1 * 8};
        cuuint32_t box[3] = {BOX, ROWS, 3};
        if (enc_nd(3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, dim, stride, box)) return 0;
        int bad = 0, tot = 0;
        cudaError_t e = cudaSuccess;
        const int starts[3][3] = {{32, 15, 3}, {224, -1, 0}, {256, 255, 4}};
        for (int t = 0; t < 3 && e == cudaSuccess; ++t) {
            k_box<3><<<1, 128, 3 * ROWS * BOX * 8>>>(tm, starts[t][0], starts[t][1], starts[t][2], 3 * ROWS * BOX * 8, dout);
            e = cudaDeviceSynchronize();
            fetch();
            for (int c = 0; c < 3; ++c)
                for (int j = 0; j < ROWS; ++j)
                    for (int k = 0; k < BOX; ++k, ++tot) {
                        int gz = starts[t][0] + k, gy = starts[t][1] + j, gp = starts[t][2] + c;
                        double want = (gz >= 0 && gz < NZS && gy >= 0 && gy < NY1 && gp >= 0 && gp < planes) ? h[((long long)gp * NY1 + gy) * NZS + gz] : 0.0;
                        bad += ho[(c * ROWS + j) * BOX + k] != want;
                    }
        }
        return verdict("3d_f64_padded_box_34x17x3_incl_oob", e, bad, tot);
    }
    if (variant == 6 || variant == 7) {  // non-tensor bulk copy of a 272 B row; 6: 16 B-aligned source, dst pitch 272; 7: 8 B-aligned source
        const long long e0 = variant == 6 ? (long long)(3 * NY1 + 17) * NZS + 64 : 259;
        k_bulk<<<1, 64, 4096>>>(d + e0, BOX * 8, 272, dout);
        cudaError_t e = cudaDeviceSynchronize();
        fetch();
        int bad = 0;
        for (int k = 0; k < BOX; ++k) bad += ho[k] != h[e0 + k];
        return verdict(variant == 6 ? "bulk_272B_src16B_aligned_dst16B_aligned" : "bulk_272B_src8B_aligned", e, bad, BOX);
    }
    if (variant >= 10) {  // throughput: 10 = 3-D plane boxes (padded), 11 = bulk rows (padded), 12 = 1-D f32-pair rows (unpadded)
        const int NPL = 8 * 24, nzs = variant == 12 ? 257 : 260;
        const long long nn = (long long)NPL * 3 * NY1 * nzs + 64;
        double* v;
        cudaMalloc(&v, nn * 8);
        cudaMemset(v, 0, nn * 8);
        alignas(64) CUtensorMap t2;
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = CUDA_SUCCESS;
        if (variant == 10) {
            cuuint64_t dim[3] = {(cuuint64_t)nzs, (cuuint64_t)NY1, (cuuint64_t)NPL * 3}, stride[2] = {(cuuint64_t)nzs * 8, (cuuint64_t)nzs * NY1 * 8};
            cuuint32_t box[3] = {BOX, ROWS, 3};
            r = enc(&t2, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, v, dim, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        const int ytiles = 16, xchunks = 8, nplanes = 24, nseg = 8;
        size_t smem = sizeof(Ring);
        double* sink;
        cudaMalloc(&sink, 8);
        cudaFuncSetAttribute(k_stream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_stream_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        auto launch = [&]() {
            if (variant == 10) k_stream<1><<<ytiles * xchunks, 32 * (NCONS + 1), smem>>>(t2, nplanes, NY1, nzs, nseg, ytiles, sink);
            if (variant == 11) k_stream_bulk<<<ytiles * xchunks, 32 * (NCONS + 1), smem>>>(v, nplanes, NY1, nzs, nseg, ytiles, sink);
        };
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        launch();
        cudaError_t e2 = cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int rep = 0; rep < 5; ++rep) launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= 5;
        double bytes = (double)ytiles * xchunks * nseg * nplanes * 3 * ROWS * BOX * 8;
        printf("{\"variant\": \"stream_%s\", \"encode\": %d, \"cuda_error\": \"%s\", \"ctas\": %d, \"ms\": %.4f, \"box_bytes_GBps\": %.1f, \"GBps_if_148_ctas\": %.1f}\n",
               variant == 10 ? "3d_plane_boxes_padded" : "bulk_rows_272B_padded", (int)r, cudaGetErrorString(e2), ytiles * xchunks, ms, bytes / ms / 1e6,
               bytes / ms / 1e6 * 148.0 / (ytiles * xchunks));
    }
    return 0;
}
