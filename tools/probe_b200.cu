// probe_b200.cu — hardware probes behind the round-2 search design: not part of the library.
//   A. TMA (cp.async.bulk.tensor.2d) on a u8 plane with byte-granular, partly out-of-range box origins
//   B. TMA staging throughput for the box shapes of the sliding search kernel
//   C. effective L2 capacity: repeated whole-buffer reads at 16..200 MB
//   D. ALU-pipe mix: VABSDIFF4 alone / with LDS / with PRMT / with IMAD
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe_b200.bin tools/probe_b200.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                             \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess) {                                                          \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                      \
        }                                                                                 \
    } while (0)

static PFN_cuTensorMapEncodeTiled getEncode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    return (PFN_cuTensorMapEncodeTiled)fn;
}

static CUtensorMap makeMap(void* base, int W, int H, int pitch, int boxW, int boxH) {
    static PFN_cuTensorMapEncodeTiled enc = getEncode();
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)boxW, (cuuint32_t)boxH};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("cuTensorMapEncodeTiled failed %d (W %d H %d pitch %d box %dx%d)\n", (int)r, W, H, pitch, boxW, boxH);
        exit(1);
    }
    return m;
}

__device__ __forceinline__ void mbarInit(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, unsigned phase) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void tmaLoad2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)),
                 "l"(map), "r"(x), "r"(y), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// ---- A: correctness ----------------------------------------------------------------------------------
__global__ void tmaCheckKernel(const __grid_constant__ CUtensorMap map, int x0, int y0, int boxW, int boxH, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbarInit(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbarExpectTx(&bar, boxW * boxH);
        tmaLoad2d(sm, &map, x0, y0, &bar);
    }
    mbarWait(&bar, 0);
    for (int i = threadIdx.x; i < boxW * boxH; i += blockDim.x) out[i] = sm[i];
}

// ---- B: staging throughput: every CTA stages nBox boxes (columns side by side) and touches them ----------------
__global__ void tmaTputKernel(const __grid_constant__ CUtensorMap map, int boxW, int boxH, int tileW, int tileH, int shiftX, int shiftY, unsigned* sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint64_t bar;
    const int nBox = tileW / boxW;
    if (threadIdx.x == 0) {
        mbarInit(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbarExpectTx(&bar, boxW * boxH * nBox);
        for (int b = 0; b < nBox; ++b) tmaLoad2d(sm + b * ((boxW * boxH + 127) & ~127), &map, blockIdx.x * tileW + b * boxW + shiftX, blockIdx.y * tileH + shiftY, &bar);
    }
    mbarWait(&bar, 0);
    unsigned acc = 0;
    const uint32_t* w = (const uint32_t*)sm;
    for (int i = threadIdx.x; i < ((boxW * boxH + 127) & ~127) * nBox / 4; i += blockDim.x) acc += w[i];
    if (acc == 0x12345678u) *sink = acc;
}

// ---- C: L2 capacity ----------------------------------------------------------------------------------
__global__ void readKernel(const uint4* __restrict__ p, size_t n, unsigned* sink) {
    unsigned acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(p + i);
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

// ---- D: ALU mixes --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
template <int MODE> __global__ void __launch_bounds__(256) mixKernel(unsigned* out, int iters, unsigned seed) {
    __shared__ uint32_t s[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) s[i] = i * 2654435761u + seed;
    __syncthreads();
    uint32_t acc[8], x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        acc[k] = 0;
        x[k] = threadIdx.x * 747796405u + k * 2891336453u + seed;
    }
    const uint32_t* q = s + (threadIdx.x & 31);
    uint32_t sel = 0x4321 + (seed & 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t a = x[k];
            if (MODE == 1) a = q[(k * 32 + (it & 31) * 32) & 2047];                         // 1 LDS per SAD
            if (MODE == 2) a = __byte_perm(x[k], x[(k + 1) & 7], sel);                    // 1 PRMT per SAD
            if (MODE == 3) a = x[k] * 3u + seed;                                          // 1 IMAD per SAD
            if (MODE == 4) a = __byte_perm(q[(k * 32 + (it & 31) * 32) & 2047], q[(k * 32 + 32 + (it & 31) * 32) & 2047], sel);  // 2 LDS + PRMT
            acc[k] = sad4(a, x[(k + 3) & 7], acc[k]);
            if (MODE == 5) acc[k] = sad4(a, x[(k + 5) & 7], acc[k]);                      // SAD only, 2 per k
        }
        if (MODE == 3 || MODE == 2) x[0] += acc[0];
    }
    unsigned r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += acc[k];
    out[blockIdx.x * 256 + threadIdx.x] = r;
}

template <int MODE> static void runMix(const char* name, unsigned* dOut, int sms) {
    const int iters = 4096, grid = sms * 8;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    mixKernel<MODE><<<grid, 256>>>(dOut, 64, 1);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    mixKernel<MODE><<<grid, 256>>>(dOut, iters, 2);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double sads = (double)grid * 256 * iters * 8 * (MODE == 5 ? 2 : 1);
    printf("D mix %-28s %8.3f ms  %7.2f G lane-SAD/s  (%.1f lanes/clk/SM at 1.965 GHz)\n", name, ms, sads / ms * 1e-6, sads / (ms * 1e-3) / (sms * 1.965e9));
}

int main(int argc, char** argv) {
    const char* only = argc > 1 ? argv[1] : "ABCD";
    auto want = [&](char c) { return strchr(only, c) != nullptr; };
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, L2 %d MB, smem/SM %zu, smem/block optin %zu\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20, prop.sharedMemPerMultiprocessor,
           prop.sharedMemPerBlockOptin);
    const int W = 3840, H = 2160, pitch = 3840;
    std::vector<uint8_t> host((size_t)pitch * H);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t* dPlane;
    CK(cudaMalloc(&dPlane, host.size()));
    CK(cudaMemcpy(dPlane, host.data(), host.size(), cudaMemcpyHostToDevice));
    unsigned* dSink;
    CK(cudaMalloc(&dSink, 4 << 20));

    // ---- A
    if (want('A')) {
        const int boxW = 128, boxH = 241;
        CUtensorMap map = makeMap(dPlane, W, H, pitch, boxW, boxH);
        uint8_t* dOut;
        CK(cudaMalloc(&dOut, boxW * boxH));
        CK(cudaFuncSetAttribute(tmaCheckKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, boxW * boxH));
        int origins[8][2] = {{0, 0}, {5, 7}, {1283, 1001}, {-3, -64}, {3840 - 100, 2160 - 200}, {-130, 50}, {3, 2159}, {16, 3}};
        int nOrg = 8;
        if (argc > 3) { origins[0][0] = atoi(argv[2]); origins[0][1] = atoi(argv[3]); nOrg = 1; }
        for (int oi = 0; oi < nOrg; ++oi) {
            const int* o = origins[oi];
            CK(cudaMemset(dOut, 0xAB, boxW * boxH));
            tmaCheckKernel<<<1, 128, boxW * boxH>>>(map, o[0], o[1], boxW, boxH, dOut);
            CK(cudaDeviceSynchronize());
            std::vector<uint8_t> got(boxW * boxH);
            CK(cudaMemcpy(got.data(), dOut, got.size(), cudaMemcpyDeviceToHost));
            size_t bad = 0, zeros = 0;
            for (int r = 0; r < boxH; ++r)
                for (int c = 0; c < boxW; ++c) {
                    const int x = o[0] + c, y = o[1] + r;
                    const bool in = x >= 0 && x < W && y >= 0 && y < H;
                    const uint8_t want = in ? host[(size_t)y * pitch + x] : 0;
                    if (!in) ++zeros;
                    if (got[r * boxW + c] != want) ++bad;
                }
            printf("A tma box 128x241 at (%5d,%5d): %zu mismatches, %zu out-of-range bytes expected zero\n", o[0], o[1], bad, zeros);
        }
        CK(cudaFree(dOut));
    }

    // ---- B
    if (want('B')) {
        struct Shape {
            int boxW, boxH, tileW, tileH, threads;
        };
        const Shape shapes[] = {{128, 241, 128, 128, 128}, {128, 177, 128, 64, 64},  {64, 177, 128, 64, 128}, {32, 145, 128, 32, 128},
                                {16, 145, 128, 32, 128},   {128, 128, 128, 128, 128}, {256, 241, 256, 128, 256}, {128, 145, 128, 32, 32}};
        for (auto& s : shapes) {
            CUtensorMap map = makeMap(dPlane, W, H, pitch, s.boxW, s.boxH);
            const int nBox = s.tileW / s.boxW;
            const size_t smem = (size_t)((s.boxW * s.boxH + 127) & ~127) * nBox;
            CK(cudaFuncSetAttribute(tmaTputKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dim3 grid(W / s.tileW, (H + s.tileH - 1) / s.tileH);
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            for (int i = 0; i < 3; ++i) tmaTputKernel<<<grid, s.threads, smem>>>(map, s.boxW, s.boxH, s.tileW, s.tileH, 16, -40, dSink);
            CK(cudaDeviceSynchronize());
            const int reps = 20;
            CK(cudaEventRecord(a));
            for (int i = 0; i < reps; ++i) tmaTputKernel<<<grid, s.threads, smem>>>(map, s.boxW, s.boxH, s.tileW, s.tileH, 16 * (i & 3), -40 + i, dSink);
            CK(cudaEventRecord(b));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, a, b));
            const double bytes = (double)grid.x * grid.y * smem;
            printf("B tma box %3dx%3d x%d per %3dx%3d tile, %3d thr, %5d CTAs, smem %6zu: %7.2f us/pass, %7.1f GB/s staged\n", s.boxW, s.boxH, nBox, s.tileW, s.tileH, s.threads,
                   grid.x * grid.y, smem, ms / reps * 1e3, bytes / (ms / reps * 1e-3) * 1e-9);
        }
    }

    // ---- C
    if (want('C')) {
        const size_t maxBytes = (size_t)400 << 20;
        uint4* buf;
        CK(cudaMalloc(&buf, maxBytes));
        CK(cudaMemset(buf, 1, maxBytes));
        const int sizesMB[] = {8, 16, 25, 33, 50, 66, 80, 100, 120, 133, 160, 200, 400};
        for (int mb : sizesMB) {
            const size_t n = ((size_t)mb << 20) / 16;
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            for (int i = 0; i < 3; ++i) readKernel<<<prop.multiProcessorCount * 8, 256>>>(buf, n, dSink);
            CK(cudaDeviceSynchronize());
            const int reps = 10;
            CK(cudaEventRecord(a));
            for (int i = 0; i < reps; ++i) readKernel<<<prop.multiProcessorCount * 8, 256>>>(buf, n, dSink);
            CK(cudaEventRecord(b));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, a, b));
            printf("C repeated read of %4d MB: %8.2f us/pass, %8.1f GB/s\n", mb, ms / reps * 1e3, (double)mb * 1.048576e6 / (ms / reps * 1e-3) * 1e-9);
        }
        // two buffers alternating (the X / Y pass pattern): A, B, A, B ...
        for (int mb : {25, 33, 50, 66}) {
            const size_t n = ((size_t)mb << 20) / 16;
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            for (int i = 0; i < 4; ++i) readKernel<<<prop.multiProcessorCount * 8, 256>>>(buf + (i & 1) * n, n, dSink);
            CK(cudaDeviceSynchronize());
            const int reps = 10;
            CK(cudaEventRecord(a));
            for (int i = 0; i < reps; ++i) readKernel<<<prop.multiProcessorCount * 8, 256>>>(buf + (i & 1) * n, n, dSink);
            CK(cudaEventRecord(b));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, a, b));
            printf("C alternating 2 x %3d MB: %8.2f us/pass, %8.1f GB/s\n", mb, ms / reps * 1e3, (double)mb * 1.048576e6 / (ms / reps * 1e-3) * 1e-9);
        }
        CK(cudaFree(buf));
    }

    // ---- D
    if (want('D')) {
        unsigned* dOut;
        CK(cudaMalloc(&dOut, (size_t)prop.multiProcessorCount * 8 * 256 * 4));
        runMix<0>("VABSDIFF4 only", dOut, prop.multiProcessorCount);
        runMix<5>("VABSDIFF4 only (x2)", dOut, prop.multiProcessorCount);
        runMix<1>("VABSDIFF4 + 1 LDS", dOut, prop.multiProcessorCount);
        runMix<2>("VABSDIFF4 + 1 PRMT", dOut, prop.multiProcessorCount);
        runMix<3>("VABSDIFF4 + 1 IMAD", dOut, prop.multiProcessorCount);
        runMix<4>("VABSDIFF4 + 2 LDS + PRMT", dOut, prop.multiProcessorCount);
        CK(cudaFree(dOut));
    }
    printf("probe done\n");
    return 0;
}
