// membw.cu -- read-bandwidth probes for the scan kernels' access patterns (diagnostic tool).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/membw tools/membw.cu && /tmp/membw
#include <cstdio>
#include <cuda_runtime.h>
#include "../patmatchdocker_b200/csrc/packed.cuh"

__global__ void __launch_bounds__(256) k_read_ldg(const uint4 *__restrict__ p, long long n16, unsigned *out)
{
    unsigned acc = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = __ldg(p + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// three planes, warp tiles of 128 words each, like k_scan_packed (uint4 + uint2 halo per lane)
__global__ void __launch_bounds__(256) k_read_planes(const unsigned *hi, const unsigned *lo, const unsigned *xx, long long ntiles, unsigned *out)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned acc = 0;
    for (long long t = warp; t < ntiles; t += nwarps) {
        const long long q0 = t * 128 + 4 * lane;
        const uint4 a = __ldg((const uint4 *)(hi + q0)), b = __ldg((const uint4 *)(lo + q0)), c = __ldg((const uint4 *)(xx + q0));
        acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// the TMA ring of k_scan_packed_exact with a trivial consumer
__global__ void __launch_bounds__(256) k_read_tma(const unsigned *hi, const unsigned *lo, const unsigned *xx, long long nbt, unsigned *out)
{
    extern __shared__ __align__(128) unsigned char ex_smem[];
    unsigned *stage_base = reinterpret_cast<unsigned *>(ex_smem);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(ex_smem + EX_STAGES * EX_STAGE_BYTES);
    unsigned long long *empty = full + EX_STAGES;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long my = blockIdx.x < nbt ? (nbt - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < EX_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long it) {
        const int s = (int)(it % EX_STAGES);
        const long long q = (blockIdx.x + it * gridDim.x) * EX_WORDS;
        unsigned *dst = stage_base + (size_t)s * (3 * EX_ROW);
        mbar_expect_tx(&full[s], EX_STAGE_BYTES);
        tma_load_1d(dst, hi + q, EX_ROW * 4, &full[s]);
        tma_load_1d(dst + EX_ROW, lo + q, EX_ROW * 4, &full[s]);
        tma_load_1d(dst + 2 * EX_ROW, xx + q, EX_ROW * 4, &full[s]);
    };
    if (threadIdx.x == 0) for (long long it = 0; it < my && it < EX_STAGES - 1; it++) issue(it);
    unsigned acc = 0;
    for (long long it = 0; it < my; it++) {
        const int s = (int)(it % EX_STAGES);
        if (threadIdx.x == 0) {
            const long long nx = it + EX_STAGES - 1;
            if (nx < my) { if (nx >= EX_STAGES) mbar_wait(&empty[nx % EX_STAGES], (unsigned)(((nx / EX_STAGES) - 1) & 1)); issue(nx); }
        }
        mbar_wait(&full[s], (unsigned)((it / EX_STAGES) & 1));
        const unsigned *sp = stage_base + (size_t)s * (3 * EX_ROW) + wib * 128 + 4 * lane;
        const uint4 a = *(const uint4 *)sp, b = *(const uint4 *)(sp + EX_ROW), c = *(const uint4 *)(sp + 2 * EX_ROW);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main()
{
    const long long nw = 96LL * 1024 * 1024 + 1024;      // words per plane (384 MiB each)
    unsigned *buf, *out;
    cudaMalloc(&buf, (size_t)nw * 4 * 3);
    cudaMalloc(&out, 64);
    cudaMemset(buf, 1, (size_t)nw * 4 * 3);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const long long ntiles = (nw - 1024) / 128, nbt = ntiles / 8;
    const double bytes = (double)ntiles * 128 * 4 * 3;
    const size_t smem = EX_STAGES * EX_STAGE_BYTES + 2 * EX_STAGES * 8;
    cudaFuncSetAttribute(k_read_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int variant = 0; variant < 3; variant++)
        for (int mult : {2, 4, 8, 16}) {
            float best = 1e9;
            for (int rep = 0; rep < 5; rep++) {
                cudaEventRecord(e0);
                if (variant == 0) k_read_ldg<<<sms * mult, 256>>>((const uint4 *)buf, (long long)(bytes / 16), out);
                else if (variant == 1) k_read_planes<<<sms * mult, 256>>>(buf, buf + nw, buf + 2 * nw, ntiles, out);
                else k_read_tma<<<sms * (mult > 4 ? 4 : mult), 256, smem>>>(buf, buf + nw, buf + 2 * nw, nbt, out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("%s blocks/SM=%2d  %.3f ms  %.0f GB/s  (%s)\n", variant == 0 ? "ldg-linear " : variant == 1 ? "ldg-3planes" : "tma-ring   ", mult, best,
                   bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
