// TEST TOOL (not part of the product library): one M128 N64 K64 tcgen05 tile with exactly the operand layout, instruction
// descriptor and MMA order of k_transform_tc (jb_transform.cu) -- fp16 operands, fp32 accumulation in TMEM, 4 K-chunks of
// the first B matrix then 4 of the second -- returning the raw fp32 accumulators.  tests/test_gpu_tc_model.py feeds it
// fixed-point operands and compares with exact integer arithmetic: this pins the two hardware properties on which the
// near-tie band of the tensor-core transform (jb_tables.cpp: build_tc_matrices) is derived.
//   nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o tests/_build/libtcprobe.so tests/tools/tc_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
}

// A: [128][64] fp16 bit patterns (row = block, column = sample); B0, B1: [64][64] fp16 bit patterns (row = output n,
// column = sample k); D: [128][64] fp32.  n_first / n_second: K-chunks (of 16) taken from B0 / B1, in that order.
__global__ void __launch_bounds__(128, 1) k_probe(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B0,
                                                  const uint16_t* __restrict__ B1, float* __restrict__ D, int n_first, int n_second) {
    extern __shared__ __align__(1024) uint8_t raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t *tA = smem, *tB0 = smem + 16384, *tB1 = smem + 16384 + 8192;
    const int tid = threadIdx.x;
    auto put_row = [](uint8_t* tile, int r, const uint16_t* src) {  // 64 fp16 = 8 chunks of 16 bytes, 128-byte swizzle
        for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(src + r * 64 + c * 8);
    };
    put_row(tA, tid, A);
    if (tid < 64) put_row(tB0, tid, B0); else put_row(tB1, tid - 64, B1);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_mbar)));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = (1u << 4) | (8u << 17) | (8u << 24);  // f32 += fp16 x fp16, N = 64, M = 128 (as the product)
    if (tid == 0) {
        const uint64_t da = umma_desc(smem_u32(tA)), d0 = umma_desc(smem_u32(tB0)), d1 = umma_desc(smem_u32(tB1));
        int issued = 0;
        for (int k = 0; k < n_first; ++k) umma_f16(tmem, da + 2 * k, d0 + 2 * k, idesc, issued++ ? 1u : 0u);
        for (int k = 0; k < n_second; ++k) umma_f16(tmem, da + 2 * k, d1 + 2 * k, idesc, issued++ ? 1u : 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_mbar)) : "memory");
    }
    mbar_wait(smem_u32(&s_mbar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
    for (int c = 0; c < 64; c += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[tid * 64 + c + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}
}  // namespace

// host pointers in, host pointer out; returns the CUDA error code (0 = ok)
extern "C" int tc_probe(const uint16_t* A, const uint16_t* B0, const uint16_t* B1, float* D, int n_first, int n_second) {
    uint16_t *dA, *dB0, *dB1;
    float* dD;
    cudaError_t e;
    if ((e = cudaMalloc(&dA, 128 * 64 * 2)) || (e = cudaMalloc(&dB0, 64 * 64 * 2)) || (e = cudaMalloc(&dB1, 64 * 64 * 2)) ||
        (e = cudaMalloc(&dD, 128 * 64 * 4)))
        return (int)e;
    cudaMemcpy(dA, A, 128 * 64 * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB0, B0, 64 * 64 * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB1, B1, 64 * 64 * 2, cudaMemcpyHostToDevice);
    const int smem = 16384 + 2 * 8192 + 1024;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_probe<<<1, 128, smem>>>(dA, dB0, dB1, dD, n_first, n_second);
    e = cudaDeviceSynchronize();
    if (!e) cudaMemcpy(D, dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    cudaFree(dA); cudaFree(dB0); cudaFree(dB1); cudaFree(dD);
    return (int)e;
}
