// Latency floor of one synchronous "host block in -> kernel -> host block out" call on this box (benchmark tooling only):
// what a single launch + completion costs, what reading the block over PCIe from inside the kernel costs, what a cluster
// launch and a cluster barrier cost.  nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/lat_floor.cu -o gpurun_out/lat_floor
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void k_empty() {}

// reads nF2 float2 from `in` (host), writes one float per thread block to `out` (host); stamps[0..3] = start, after the read, end
__global__ void k_rw(const float2* in, float* out, int nF2, unsigned long long* stamps, volatile unsigned* flag, unsigned seq)
{
    const unsigned long long t0 = gtime();
    float s = 0.f;
    for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < nF2; i += blockDim.x * gridDim.x) { const float2 v = in[i]; s += v.x + v.y; }
    __shared__ float red[32];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    const unsigned long long t1 = gtime();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        out[blockIdx.x] = t;
        if (flag) { __threadfence_system(); *flag = seq; }
        if (stamps && blockIdx.x == 0) { stamps[0] = t0; stamps[1] = t1; stamps[2] = gtime(); }
    }
}

__global__ void k_cluster(unsigned long long* stamps)
{
    const unsigned long long t0 = gtime();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const unsigned long long t1 = gtime();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const unsigned long long t2 = gtime();
    if (stamps && blockIdx.x == 0 && threadIdx.x == 0) { stamps[0] = t0; stamps[1] = t1; stamps[2] = t2; }
}

template <class F> static double p50_us(F&& f, int n = 2000, int warm = 200)
{
    std::vector<double> v;
    for (int i = 0; i < warm + n; ++i) {
        auto a = std::chrono::steady_clock::now();
        f();
        auto b = std::chrono::steady_clock::now();
        if (i >= warm) v.push_back(std::chrono::duration<double, std::micro>(b - a).count());
    }
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
}

int main()
{
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    float2* hin; float* hout; unsigned* flag; unsigned long long *dst, hst[4];
    CK(cudaHostAlloc(&hin, 1 << 20, cudaHostAllocDefault));
    CK(cudaHostAlloc(&hout, 4096, cudaHostAllocDefault));
    CK(cudaHostAlloc(&flag, 64, cudaHostAllocDefault));
    CK(cudaMalloc(&dst, 64));
    memset(hin, 0, 1 << 20); *flag = 0;
    printf("empty kernel + cudaStreamSynchronize            p50 %.2f us\n", p50_us([&] { k_empty<<<1, 32, 0, st>>>(); cudaStreamSynchronize(st); }));
    printf("launch only (no sync, queue kept short)         p50 %.2f us\n", p50_us([&] { k_empty<<<1, 32, 0, st>>>(); }, 200, 20));
    CK(cudaStreamSynchronize(st));
    for (int bytes : {4096, 12800, 65536}) {
        for (int ctas : {1, 8}) {
            const int nF2 = bytes / 8;
            double a = p50_us([&] { k_rw<<<ctas, 512, 0, st>>>(hin, hout, nF2, dst, nullptr, 0); cudaStreamSynchronize(st); });
            CK(cudaMemcpy(hst, dst, 32, cudaMemcpyDeviceToHost));
            unsigned seq = 0;
            double b = p50_us([&] { ++seq; k_rw<<<ctas, 512, 0, st>>>(hin, hout, nF2, nullptr, flag, seq); while (*(volatile unsigned*)flag != seq) { } });
            CK(cudaStreamSynchronize(st));
            printf("host read %6d B, %d CTA(s): sync p50 %.2f us, flag-poll p50 %.2f us; in-kernel: read %.2f us, tail %.2f us\n", bytes, ctas, a, b,
                   (hst[1] - hst[0]) * 1e-3, (hst[2] - hst[1]) * 1e-3);
        }
    }
    {
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(8); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        double a = p50_us([&] { cudaLaunchKernelEx(&cfg, k_cluster, dst); cudaStreamSynchronize(st); });
        CK(cudaGetLastError());
        CK(cudaMemcpy(hst, dst, 32, cudaMemcpyDeviceToHost));
        printf("cluster of 8 x 512 threads, two cluster barriers + sync: p50 %.2f us; in-kernel: first barrier %.2f us, second %.2f us\n", a,
               (hst[1] - hst[0]) * 1e-3, (hst[2] - hst[1]) * 1e-3);
    }
    return 0;
}
