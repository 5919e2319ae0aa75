// tcgen05.mma issue-rate microbenchmark (dev aid): cycles per MMA for SS operands as a function of M, N and the
// A-operand layout.  One CTA, one issuing thread, `reps` MMAs back to back into one accumulator, then a commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rust-birdnet-onnx_b200/csrc tools/mma_microbench.cu -o tools/_build/mma_microbench
#include "tc_common.cuh"
#include <cstdio>
#include <vector>
using namespace bn::tc;

struct Cfg { int M, N, layout, a_stride, b_same, alt, bg; };   // bg: background smem traffic from the other warps (0 none, 1 cp.async 16 B, 2 st.shared.v4, 3 ld.shared.v4)   // alt: 1 = pairs (N=2n into acc, N=n into acc+n) like the conv kernel; 2 = same but one idesc   // layout: 0 no-swizzle cells, 1 SW128, 2 SW64, 3 SW32

__global__ void __launch_bounds__(512, 1) k_bench(Cfg c, int reps, long long* out, const uint4* gsrc, volatile int* stop) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_holder;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x;
    for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (tid < 32) tmem_alloc(&tmem_holder, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_holder;
    if (tid == 0) {
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
        const uint32_t idesc = umma_idesc_f16(c.M, c.N);
        uint64_t da0, db0;
        const uint32_t rb = c.layout == 1 ? 128u : (c.layout == 2 ? 64u : 32u);
        if (c.layout == 0) { da0 = umma_desc_nosw(a_base, 4096, 128); db0 = umma_desc_nosw(b_base, 4096, 128); }
        else { da0 = umma_desc_kmajor(a_base, rb); db0 = umma_desc_kmajor(b_base, rb); }
        // warm
        umma_f16(tmem, da0, db0, idesc, 0);
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t0 = clock64();
        if (c.alt) {
            const uint32_t idesc2 = umma_idesc_f16(c.M, 2 * c.N), idesc1 = c.alt == 1 ? umma_idesc_f16(c.M, c.N) : idesc2;
            const uint32_t second = c.alt == 3 ? 256u : (uint32_t)c.N;      // alt 3: the second MMA accumulates into columns far away
            for (int i = 0; i < reps / 2; ++i) {
                const uint32_t off = ((uint32_t)i * (uint32_t)c.a_stride) & 0x7FFFu & ~0xFu;
                umma_f16(tmem, da0 + (off >> 4), db0, idesc2, 1);
                umma_f16(tmem + second, da0 + ((off + 0x8000u) >> 4), db0, c.alt == 3 ? umma_idesc_f16(c.M, c.N) : idesc1, 1);
            }
        } else
        for (int i = 0; i < reps; ++i) {
            // walk the A start address like a K loop would (a_stride bytes per step, wrapping inside 64 KB)
            const uint32_t off = ((uint32_t)i * (uint32_t)c.a_stride) & 0xFFFFu & ~0xFu;
            umma_f16(tmem, da0 + (off >> 4), c.b_same ? db0 : db0 + ((off & 0x3FFFu) >> 4), idesc, 1);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 1);
        const long long t1 = clock64();
        out[0] = t1 - t0;
        *stop = 1;
        __threadfence();
    }
    if (tid >= 128 && c.bg) {
        // background traffic into / out of a 48 KB region behind the B tile until the issuing thread is done
        __shared__ volatile int s_stop;
        if (tid == 128) s_stop = 0;
        uint8_t* reg = smem + 110 * 1024;
        const int t = tid - 128;                       // 0..383
        uint4 acc4 = make_uint4(0, 0, 0, 0);
        long long n = 0;
        for (int it = 0; it < 100000 && !*stop; ++it) {
            const uint32_t o = (uint32_t)((t * 16 + (it & 7) * 6144) % (48 * 1024));
            if (c.bg == 1) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(reg + o)), "l"(gsrc + ((t + it * 384) & 0xFFFF)) : "memory");
                if ((it & 3) == 3) asm volatile("cp.async.wait_all;" ::: "memory");
            } else if (c.bg == 2) {
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(reg + o)), "r"(acc4.x), "r"(acc4.y), "r"(acc4.z), "r"(acc4.w) : "memory");
            } else {
                uint4 v;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(reg + o)) : "memory");
                acc4.x += v.x;
            }
            ++n;
        }
        if (c.bg == 1) asm volatile("cp.async.wait_all;" ::: "memory");
        if (acc4.x == 12345u) out[2] = n;
        if (t == 0) out[1] = n;
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 512);
}

// same measurement with the CUTLASS-style issue: the whole warp runs the loop (warp-uniform values stay in uniform
// registers), only the MMA itself is predicated on one elected lane
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
__global__ void __launch_bounds__(128, 1) k_bench_uniform(Cfg c, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_holder;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x;
    for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (tid < 32) tmem_alloc(&tmem_holder, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_holder;
    if (tid < 32) {
        const uint32_t leader = elect_one();
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
        const uint64_t da0 = umma_desc_nosw(a_base, 4096, 128), db0 = umma_desc_nosw(b_base, 4096, 128);
        const uint32_t idesc2 = umma_idesc_f16(c.M, 2 * c.N), idesc1 = umma_idesc_f16(c.M, c.N);
        if (leader) { umma_f16(tmem, da0, db0, idesc1, 0); umma_commit(&bar); }
        mbar_wait(&bar, 0);
        const long long t0 = clock64();
        if (c.alt) {
            for (int i = 0; i < reps / 2; ++i) {
                const uint32_t off = ((uint32_t)i * (uint32_t)c.a_stride) & 0x7FFFu & ~0xFu;
                if (leader) {
                    umma_f16(tmem, da0 + (off >> 4), db0, idesc2, 1);
                    umma_f16(tmem + (uint32_t)c.N, da0 + ((off + 0x8000u) >> 4), db0, idesc1, 1);
                }
            }
        } else {
            for (int i = 0; i < reps; ++i) {
                const uint32_t off = ((uint32_t)i * (uint32_t)c.a_stride) & 0xFFFFu & ~0xFu;
                if (leader) umma_f16(tmem, da0 + (off >> 4), db0, idesc1, 1);
            }
        }
        if (leader) umma_commit(&bar);
        mbar_wait(&bar, 1);
        const long long t1 = clock64();
        if (leader) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 32);
    uint4* gsrc;
    cudaMalloc(&gsrc, 65536 * 16 + 65536);
    cudaMemset(gsrc, 0, 65536 * 16);
    int* stop;
    cudaMalloc(&stop, 4);
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int reps = 2000;
    const char* lname[] = {"nosw", "sw128", "sw64", "sw32"};
    cudaFuncSetAttribute(k_bench_uniform, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int alt : {0, 1})
        for (int M : {64, 128})
            for (int N : {16, 32, 64, 128, 256}) {
                if (alt && N > 128) continue;
                Cfg c{M, N, 0, 32, 1, alt, 0};
                k_bench_uniform<<<1, 128, 200 * 1024>>>(c, 20000, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("uniform: %s\n", cudaGetErrorString(e)); return 1; }
                long long cyc = 0;
                cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
                printf("uniform-issue %s M=%3d N=%3d: %7.1f cycles/MMA\n", alt ? "pairs (N=2n, N=n)" : "single", M, N, (double)cyc / 20000);
            }
    for (int bg : {0}) {
        Cfg c{128, 16, 0, 32, 1, 1, bg};
        cudaMemset(stop, 0, 4);
        k_bench<<<1, 512, 200 * 1024>>>(c, 20000, d, gsrc, stop);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("bg: %s\n", cudaGetErrorString(e)); return 1; }
        long long r[2] = {0, 0};
        cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost);
        printf("pairs n=16 with background %d (0 none, 1 cp.async16, 2 st.v4, 3 ld.v4; 12 warps): %7.1f cycles/MMA; background 16-byte ops per thread %lld -> %.1f B/clk\n",
               bg, (double)r[0] / 20000, r[1], bg ? (double)r[1] * 384 * 16 / (double)r[0] : 0.0);
    }
    for (int alt : {1, 2, 3})
        for (int N : {16, 32, 64}) {
            Cfg c{128, N, 0, 32, 1, alt, 0};
            cudaMemset(stop, 0, 4);
            k_bench<<<1, 128, 200 * 1024>>>(c, reps, d, gsrc, stop);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("alt: %s\n", cudaGetErrorString(e)); return 1; }
            long long cyc = 0;
            cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
            printf("alt=%d (1: N=2n then N=n into acc+n; 2: both N=2n; 3: N=2n then N=n into far columns) n=%d: %7.1f cycles/MMA\n", alt, N, (double)cyc / reps);
        }
    for (int layout : {0, 1})
        for (int M : {64, 128})
            for (int N : {16, 64, 128, 256}) {
                for (int a_stride : {0, 32, 1024}) {
                    if (layout != 0 && a_stride == 1024 && false) continue;
                    Cfg c{M, N, layout, a_stride, 1, 0, 0};
                    cudaMemset(stop, 0, 4);
                    k_bench<<<1, 128, 200 * 1024>>>(c, reps, d, gsrc, stop);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("%s M=%d N=%d stride=%d: %s\n", lname[layout], M, N, a_stride, cudaGetErrorString(e)); return 1; }
                    long long cyc = 0;
                    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
                    printf("%-5s M=%3d N=%3d a_stride=%4d: %7.1f cycles/MMA  (math floor %d)\n", lname[layout], M, N, a_stride,
                           (double)cyc / reps, (M < 128 ? 128 : M) * N / 256);
                }
            }
    return 0;
}
