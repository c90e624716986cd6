// tcgen05 probe: D[128x64] (f32, TMEM) = A[128x64] (bf16, K-major, SWIZZLE_128B) x B[64x64]^T (bf16, K-major).
// Validates descriptors/layouts for the tensor-core DCT experiment.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu && ./tc_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address
    d |= (uint64_t)0 << 16;                    // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;             // 128 rows x 128 B = 16 KB
    uint8_t* sB = smem + 16384;     // 64 rows x 128 B = 8 KB
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5;
    // fill A row t, B row t (t < 64) with the 128B swizzle: 16-byte chunk c of row r at r*128 + ((c ^ (r&7)) << 4)
    for (int c = 0; c < 8; ++c) {
        uint4 v = *reinterpret_cast<const uint4*>(A + t * 64 + c * 8);
        *reinterpret_cast<uint4*>(sA + t * 128 + ((c ^ (t & 7)) << 4)) = v;
        if (t < 64) {
            uint4 w = *reinterpret_cast<const uint4*>(B + t * 64 + c * 8);
            *reinterpret_cast<uint4*>(sB + t * 128 + ((c ^ (t & 7)) << 4)) = w;
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    }
    asm volatile("fence.proxy.async.shared::cta;");   // generic-proxy smem writes -> visible to the async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = tmem_base;
    if (t == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (8u << 17) | (8u << 24);  // f32 acc, bf16 x bf16, N=64, M=128
        uint64_t da = make_desc(smem_u32(sA)), db = make_desc(smem_u32(sB));
        for (int k = 0; k < 4; ++k) {
            uint32_t acc = k > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tm), "l"(da + (uint64_t)(2 * k)), "l"(db + (uint64_t)(2 * k)), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
    }
    // wait for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u));
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t r[64];
    uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, "
        "%46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int n = 0; n < 64; ++n) D[t * 64 + n] = __uint_as_float(r[n]);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm));
}

int main() {
    std::vector<__nv_bfloat16> hA(128 * 64), hB(64 * 64);
    std::vector<float> fA(128 * 64), fB(64 * 64);
    srand(7);
    for (int i = 0; i < 128 * 64; ++i) { fA[i] = (float)(rand() % 256 - 128); hA[i] = __float2bfloat16(fA[i]); }
    for (int i = 0; i < 64 * 64; ++i) { fB[i] = (float)(rand() % 255 - 127) / 128.0f; hB[i] = __float2bfloat16(fB[i]); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * 64 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576 + 1024);
    probe<<<1, 128, 24576 + 1024>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> hD(128 * 64);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            double s = 0;
            for (int k = 0; k < 64; ++k) s += (double)fA[m * 64 + k] * fB[n * 64 + k];
            double d = fabs(s - hD[m * 64 + n]);
            if (d > maxerr) maxerr = d;
            if (d > 1e-3) { if (bad < 5) printf("mismatch m=%d n=%d got %f want %f\n", m, n, hD[m * 64 + n], s); ++bad; }
        }
    printf("max abs err %.3e, mismatches %d of %d\n", maxerr, bad, 128 * 64);
    return bad != 0;
}
