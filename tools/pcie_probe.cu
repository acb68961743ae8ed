// PCIe host->device probe for the host entry points (sq_implicit_loss_host*): how fast do the rows an ImplicitLoss call
// needs (every 4th 256-byte row of 256 8-bit 256x256 depth maps = 4.2 MB) reach HBM
//   (a) as one strided copy-engine transfer (cudaMemcpy2DAsync, pitch 1024 -> 256),
//   (b) as a contiguous copy of the same number of bytes, (c) as a contiguous copy of the whole images (16.8 MB),
//   (d) through a zero-copy gather kernel (16-byte loads from mapped pinned memory), (e) = (a) and (d) half each, concurrently.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pcie_probe tools/pcie_probe.cu && /tmp/pcie_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void gather_rows(const uint4* __restrict__ src, uint4* __restrict__ dst, int rows, int row_vec, int pitch_vec) {
    const int n = rows * row_vec;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / row_vec, c = i - r * row_vec;
        dst[i] = src[(size_t)r * pitch_vec + c];
    }
}

int main() {
    const int B = 256, H = 256, W = 256, R = 64, rows = B * R;
    const size_t img_bytes = (size_t)B * H * W, need = (size_t)rows * W;
    unsigned char *h, *d, *hmap;
    CK(cudaHostAlloc(&h, img_bytes, cudaHostAllocMapped));
    for (size_t i = 0; i < img_bytes; ++i) h[i] = (unsigned char)(i * 2654435761u >> 24);
    CK(cudaHostGetDevicePointer(&hmap, h, 0));
    CK(cudaMalloc(&d, img_bytes));
    cudaStream_t s1, s2;
    CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    const int reps = 50;
    auto report = [&](const char* name, size_t bytes, float ms) {
        printf("%-58s %8.1f us  %6.1f GB/s\n", name, ms * 1e3 / reps, bytes * reps / (ms * 1e-3) / 1e9);
    };
    for (int pass = 0; pass < 2; ++pass) {
        float ms;
        CK(cudaEventRecord(e0, s1));
        for (int i = 0; i < reps; ++i) CK(cudaMemcpy2DAsync(d, W, h, 4 * W, W, rows, cudaMemcpyHostToDevice, s1));
        CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (pass) report("(a) copy engine, 16384 rows of 256 B, pitch 1024", need, ms);
        CK(cudaEventRecord(e0, s1));
        for (int i = 0; i < reps; ++i) CK(cudaMemcpyAsync(d, h, need, cudaMemcpyHostToDevice, s1));
        CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (pass) report("(b) copy engine, 4.2 MB contiguous", need, ms);
        CK(cudaEventRecord(e0, s1));
        for (int i = 0; i < reps; ++i) CK(cudaMemcpyAsync(d, h, img_bytes, cudaMemcpyHostToDevice, s1));
        CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (pass) report("(c) copy engine, whole images 16.8 MB contiguous", img_bytes, ms);
        for (int blocks = 148; blocks <= 148 * 8; blocks *= 2) {
            CK(cudaEventRecord(e0, s1));
            for (int i = 0; i < reps; ++i)
                gather_rows<<<blocks, 256, 0, s1>>>((const uint4*)hmap, (uint4*)d, rows, W / 16, 4 * W / 16);
            CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            char nm[96]; snprintf(nm, sizeof nm, "(d) zero-copy gather kernel, %d blocks x 256", blocks);
            if (pass) report(nm, need, ms);
        }
        // (e) half the rows by the copy engine, half by the kernel, concurrently on two streams
        CK(cudaEventRecord(e0, s1)); CK(cudaStreamWaitEvent(s2, e0, 0));
        for (int i = 0; i < reps; ++i) {
            CK(cudaMemcpy2DAsync(d, W, h, 4 * W, W, rows / 2, cudaMemcpyHostToDevice, s1));
            gather_rows<<<296, 256, 0, s2>>>((const uint4*)(hmap + (size_t)(rows / 2) * 4 * W), (uint4*)(d + (size_t)(rows / 2) * W),
                                              rows / 2, W / 16, 4 * W / 16);
        }
        CK(cudaEventRecord(e2, s2)); CK(cudaStreamWaitEvent(s1, e2, 0));
        CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (pass) report("(e) half copy engine + half zero-copy kernel, concurrent", need, ms);
        // (f) the strided transfer split over two streams (two copy engines?)
        CK(cudaEventRecord(e0, s1)); CK(cudaStreamWaitEvent(s2, e0, 0));
        for (int i = 0; i < reps; ++i) {
            CK(cudaMemcpy2DAsync(d, W, h, 4 * W, W, rows / 2, cudaMemcpyHostToDevice, s1));
            CK(cudaMemcpy2DAsync(d + (size_t)(rows / 2) * W, W, h + (size_t)(rows / 2) * 4 * W, 4 * W, W, rows / 2, cudaMemcpyHostToDevice, s2));
        }
        CK(cudaEventRecord(e2, s2)); CK(cudaStreamWaitEvent(s1, e2, 0));
        CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (pass) report("(f) strided transfer split over two streams", need, ms);
    }
    return 0;
}
