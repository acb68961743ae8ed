// Accuracy of the MUFU approximations on this GPU against fp64 (tools; not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe mufu_probe.cu && ./mufu_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
__device__ float ex2f_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ float lg2f_(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ float rcpf_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// res: [0] max abs err, [1] max rel err, [2] sum abs err, [3] sum signed err
__global__ void probe(int op, float lo, float hi, int n, double* res) {
    double mabs = 0, mrel = 0, sabs = 0, ssgn = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = lo + (hi - lo) * ((float)i + 0.5f) / (float)n;
        double ref, got;
        if (op == 0) { got = ex2f_(x); ref = exp2((double)x); }
        else if (op == 1) { got = lg2f_(x); ref = log2((double)x); }
        else { got = rcpf_(x); ref = 1.0 / (double)x; }
        const double e = got - ref;
        mabs = fmax(mabs, fabs(e)); mrel = fmax(mrel, fabs(e) / fmax(fabs(ref), 1e-300)); sabs += fabs(e); ssgn += e;
    }
    atomicMax((unsigned long long*)&res[0], (unsigned long long)__double_as_longlong(mabs));
    atomicMax((unsigned long long*)&res[1], (unsigned long long)__double_as_longlong(mrel));
    atomicAdd(&res[2], sabs); atomicAdd(&res[3], ssgn);
}
int main() {
    double* d; cudaMalloc(&d, 32);
    struct { const char* name; int op; float lo, hi; } cases[] = {
        {"ex2 [-1,0]", 0, -1.f, 0.f}, {"ex2 [0,1]", 0, 0.f, 1.f}, {"ex2 [-0.2,0.2]", 0, -0.2f, 0.2f}, {"ex2 [-60,-20]", 0, -60.f, -20.f},
        {"lg2 [1,2]", 1, 1.f, 2.f}, {"lg2 [0.5,1]", 1, 0.5f, 1.f}, {"lg2 [0.9,1.1]", 1, 0.9f, 1.1f}, {"lg2 [0.01,0.5]", 1, 0.01f, 0.5f}, {"lg2 [2,30]", 1, 2.f, 30.f},
        {"rcp [1,2]", 2, 1.f, 2.f}, {"rcp [0.01,1]", 2, 0.01f, 1.f}};
    const int n = 1 << 24;
    printf("{\n");
    for (auto& c : cases) {
        cudaMemset(d, 0, 32);
        probe<<<592, 256>>>(c.op, c.lo, c.hi, n, d);
        double h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf(" \"%s\": {\"max_abs\": %.3e, \"max_rel\": %.3e, \"mean_abs\": %.3e, \"mean_signed\": %.3e},\n", c.name, h[0], h[1], h[2] / n, h[3] / n);
    }
    printf(" \"n\": %d\n}\n", n);
    return 0;
}
