// pcie_probe.cu -- which mechanism moves the MUTABLE fields of packed 316-byte agent records over PCIe fastest?
// (measurement behind the e2e / strict-mode transfer design).  nvcc -O3 -arch=sm_100a -o pcie_probe pcie_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int ITEM = 316;
struct Span { int off, len; };
__constant__ Span c_spans[3] = {{0, 32}, {124, 80}, {260, 40}};

// zero-copy: words of the three spans of every record, host-mapped <-> compact device buffer (38 words per record)
__global__ void k_read_spans(const uint32_t *__restrict__ host, uint32_t *__restrict__ dev, int n) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long rec = g / 38; const int w = (int)(g % 38);
    if (rec >= n) return;
    const int src = w < 8 ? w : (w < 28 ? 31 + (w - 8) : 65 + (w - 28));
    dev[g] = host[rec * (ITEM / 4) + src];
}
__global__ void k_write_spans(uint32_t *__restrict__ host, const uint32_t *__restrict__ dev, int n) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long rec = g / 38; const int w = (int)(g % 38);
    if (rec >= n) return;
    const int dst = w < 8 ? w : (w < 28 ? 31 + (w - 8) : 65 + (w - 28));
    host[rec * (ITEM / 4) + dst] = dev[g];
}
int main() {
    const int n = 1000000;
    const size_t bytes = (size_t)n * ITEM;
    uint8_t *h = nullptr, *d = nullptr; uint32_t *dc = nullptr, *hm = nullptr;
    CK(cudaHostAlloc((void **)&h, bytes, cudaHostAllocMapped));
    memset(h, 1, bytes);
    CK(cudaMalloc((void **)&d, bytes + 64));
    CK(cudaMalloc((void **)&dc, (size_t)n * 38 * 4));
    CK(cudaHostGetDevicePointer((void **)&hm, h, 0));
    cudaStream_t s0, s1; CK(cudaStreamCreate(&s0)); CK(cudaStreamCreate(&s1));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const Span spans[3] = {{0, 32}, {124, 80}, {260, 40}};
    auto timeit = [&](const char *name, double useful_mb, auto fn) {
        float best = 1e9f;
        for (int r = 0; r < 5; ++r) {
            cudaDeviceSynchronize();
            cudaEventRecord(e0, s0);
            fn();
            cudaEventRecord(e1, s0);
            cudaEventSynchronize(e1); cudaStreamSynchronize(s1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0 && ms < best) best = ms;
        }
        printf("%-46s %8.3f ms  %7.1f GB/s useful\n", name, best, useful_mb / best);
        return 0;
    };
    timeit("H2D whole records (316 B)", bytes / 1e6, [&] { cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s0); });
    timeit("D2H whole records (316 B)", bytes / 1e6, [&] { cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s0); });
    timeit("H2D 3 spans, cudaMemcpy2DAsync (152 B)", n * 152 / 1e6, [&] {
        for (auto sp : spans) cudaMemcpy2DAsync(d + sp.off, ITEM, h + sp.off, ITEM, sp.len, n, cudaMemcpyHostToDevice, s0); });
    timeit("D2H 3 spans, cudaMemcpy2DAsync (152 B)", n * 152 / 1e6, [&] {
        for (auto sp : spans) cudaMemcpy2DAsync(h + sp.off, ITEM, d + sp.off, ITEM, sp.len, n, cudaMemcpyDeviceToHost, s0); });
    timeit("H2D 1 span 0..300, cudaMemcpy2DAsync", n * 300 / 1e6, [&] { cudaMemcpy2DAsync(d, ITEM, h, ITEM, 300, n, cudaMemcpyHostToDevice, s0); });
    const int blocks = (int)(((long long)n * 38 + 255) / 256);
    timeit("H2D zero-copy kernel read of 3 spans (152 B)", n * 152 / 1e6, [&] { k_read_spans<<<blocks, 256, 0, s0>>>(hm, dc, n); });
    timeit("D2H zero-copy kernel write of 3 spans (152 B)", n * 152 / 1e6, [&] { k_write_spans<<<blocks, 256, 0, s0>>>(hm, dc, n); });
    timeit("H2D compact buffer (152 B, contiguous)", n * 152 / 1e6, [&] { cudaMemcpyAsync(dc, h, (size_t)n * 152, cudaMemcpyHostToDevice, s0); });
    // full duplex: whole H2D on s0 while whole D2H on s1
    uint8_t *h2 = nullptr, *d2 = nullptr;
    CK(cudaHostAlloc((void **)&h2, bytes, cudaHostAllocDefault)); CK(cudaMalloc((void **)&d2, bytes));
    timeit("H2D whole || D2H whole (two streams)", 2 * bytes / 1e6, [&] {
        cudaMemcpyAsync(h2, d2, bytes, cudaMemcpyDeviceToHost, s1);
        cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s0); });
    // chunked H2D (8 chunks) to see the per-call overhead
    timeit("H2D whole in 8 chunks", bytes / 1e6, [&] {
        for (int c = 0; c < 8; ++c) cudaMemcpyAsync(d + c * (bytes / 8), h + c * (bytes / 8), bytes / 8, cudaMemcpyHostToDevice, s0); });
    return 0;
}
