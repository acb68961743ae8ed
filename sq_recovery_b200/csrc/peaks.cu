// Pipe-throughput microbenchmarks for B200 (sm_100a): the roofline denominators of the SQ loss kernels.
//
// The loss kernels are bound by the MUFU (SFU) pipe and by FP32 issue slots, not by HBM or tensor cores
// (SURVEY.md 8d), and MEASURED_PEAKS.json only holds HBM and bf16 figures.  This program measures, on the
// box it runs on, thread-ops per clock per SM for the instruction classes those kernels use, alone and mixed,
// and prints one JSON object.  bench.py runs it once and uses "mufu_per_clk_sm" x SMs x the SM clock observed
// during the timed region as the MUFU peak.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peaks peaks.cu && ./peaks
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ILP = 8;          // independent chains per thread
constexpr int INNER = 64;       // unrolled ops per chain per outer iteration

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float y; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }
__device__ __forceinline__ float fadd(float a, float b) { float y; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b)); return y; }
__device__ __forceinline__ float fmnmx(float a, float b) { float y; asm volatile("max.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b)); return y; }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long y; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(y) : "l"(a), "l"(b), "l"(c)); return y; }
__device__ __forceinline__ double dfma(double a, double b, double c) { double y; asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(y) : "d"(a), "d"(b), "d"(c)); return y; }
__device__ __forceinline__ float d2f(double a) { float y; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(y) : "d"(a)); return y; }
__device__ __forceinline__ float i2f(int a) { float y; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(y) : "r"(a)); return y; }
__device__ __forceinline__ float fsel(float a, float b) { float y; asm volatile("{ .reg .pred p; setp.gt.f32 p, %1, %2; selp.f32 %0, %1, %2, p; }" : "=f"(y) : "f"(a), "f"(b)); return y; }

__device__ __forceinline__ int popc(int a) { int y; asm volatile("popc.b32 %0, %1;" : "=r"(y) : "r"(a)); return y; }

enum Op { EX2, LG2, RCP, MUFU_MIX, FFMA, FADD, FFMA2, FMNMX, FSEL, DFMA, D2F, I2F, SHFL,
          POPC, EX2_POPC, VOTE, EX2_VOTE, REDUX, EX2_REDUX,     // what the pool appends of the column kernel add to the walk
          EX2_FFMA1, EX2_FFMA2, EX2_FFMA4, EX2_FFMA6, EX2_FFMA8, EX2_FFMA2X4, EX2_D2F, EX2_DFMA, EX2_FMNMX4, POW_CHAIN };

// Each variant: ILP independent dependency chains, INNER steps per outer iteration.  "ops" counted per thread
// per outer iteration is returned by ops_per_iter() on the host side.
template <int OP>
__global__ void __launch_bounds__(256) bench_kernel(float* out, int iters, float seed, long long* cycles) {
    float v[ILP];
    double dv[ILP];
    unsigned long long pv[ILP];
    int iv[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
        v[j] = seed + 0.001f * (threadIdx.x + j);
        dv[j] = v[j];
        iv[j] = threadIdx.x + j;
        float2 t = make_float2(v[j], v[j] + 0.5f);
        pv[j] = *reinterpret_cast<unsigned long long*>(&t);
    }
    float acc = 0.f;
    const float c1 = 0.999f, c2 = 1e-3f;
    const double d1 = 0.999, d2 = 1e-3;
    float2 pc1 = make_float2(0.999f, 0.999f), pc2 = make_float2(1e-3f, 1e-3f);
    unsigned long long pC1 = *reinterpret_cast<unsigned long long*>(&pc1), pC2 = *reinterpret_cast<unsigned long long*>(&pc2);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < INNER; ++k) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                if (OP == EX2) v[j] = ex2f(v[j]);
                else if (OP == LG2) v[j] = lg2f(v[j]);
                else if (OP == RCP) v[j] = fadd(rcpf(v[j]), c2);   // + FADD (volatile asm, like every op here) keeps x near 1
                else if (OP == MUFU_MIX) { v[j] = (k % 3 == 0) ? ex2f(v[j]) : (k % 3 == 1) ? lg2f(v[j]) : rcpf(v[j]); }
                else if (OP == FFMA) v[j] = ffma(v[j], c1, c2);
                else if (OP == FADD) v[j] = fadd(v[j], c2);
                else if (OP == FFMA2) pv[j] = ffma2(pv[j], pC1, pC2);
                else if (OP == FMNMX) v[j] = fmnmx(v[j], c1);
                else if (OP == FSEL) v[j] = fsel(v[j], c1);
                else if (OP == DFMA) dv[j] = dfma(dv[j], d1, d2);
                else if (OP == D2F) { v[j] = d2f(dv[j]); dv[j] += (double)k; acc += v[j]; }   // 1 cvt + 1 DADD + 1 FADD
                else if (OP == I2F) { v[j] += i2f(iv[j]); iv[j] += k; }                        // 1 cvt + 1 FADD + 1 IADD
                else if (OP == SHFL) v[j] = __shfl_xor_sync(0xffffffffu, v[j], 1);
                else if (OP == POPC) iv[j] = popc(iv[j]) + k;                                  // 1 popc + 1 IADD
                else if (OP == EX2_POPC) { v[j] = ex2f(v[j]); iv[j] = popc(iv[j]) + k; }
                else if (OP == VOTE) iv[j] += (int)__ballot_sync(0xffffffffu, iv[j] & 1);      // 1 vote + setp + IADD
                else if (OP == EX2_VOTE) { v[j] = ex2f(v[j]); iv[j] += (int)__ballot_sync(0xffffffffu, iv[j] & 1); }
                else if (OP == REDUX) iv[j] = __reduce_add_sync(0xffffffffu, iv[j] & 3) + k;   // 1 redux + LOP + IADD
                else if (OP == EX2_REDUX) { v[j] = ex2f(v[j]); iv[j] = __reduce_add_sync(0xffffffffu, iv[j] & 3) + k; }
                else if (OP == EX2_FFMA1) { v[j] = ex2f(v[j]); v[j] = ffma(v[j], c1, c2); }
                else if (OP == EX2_FFMA2) { v[j] = ex2f(v[j]); v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2); }
                else if (OP == EX2_FFMA4) { v[j] = ex2f(v[j]);
                    v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2); }
                else if (OP == EX2_FFMA6) { v[j] = ex2f(v[j]);
                    v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2);
                    v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2); }
                else if (OP == EX2_FFMA8) { v[j] = ex2f(v[j]);
                    v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2);
                    v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2); v[j] = ffma(v[j], c1, c2); }
                else if (OP == EX2_FFMA2X4) { v[j] = ex2f(v[j]);      // 1 MUFU + 4 packed FMA (= 8 scalar FMA) per step
                    pv[j] = ffma2(pv[j], pC1, pC2); pv[j] = ffma2(pv[j], pC1, pC2);
                    pv[j] = ffma2(pv[j], pC1, pC2); pv[j] = ffma2(pv[j], pC1, pC2); }
                else if (OP == EX2_D2F) { v[j] = ex2f(v[j]); acc += d2f(dv[j]); dv[j] += (double)k; }
                else if (OP == EX2_DFMA) { v[j] = ex2f(v[j]); dv[j] = dfma(dv[j], d1, d2); }
                else if (OP == EX2_FMNMX4) { v[j] = ex2f(v[j]);
                    v[j] = fmnmx(v[j], c1); v[j] = fmnmx(v[j], c2); v[j] = fmnmx(v[j], c1); v[j] = fmnmx(v[j], c2); }
                else if (OP == POW_CHAIN) {   // the dependent lg2 -> fma -> ex2 -> add pattern of one pow stage
                    v[j] = ex2f(ffma(lg2f(v[j]), c1, c2)) + c2; }
            }
        }
    }
    long long t1 = clock64();
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
        float2 t = *reinterpret_cast<float2*>(&pv[j]);
        acc += v[j] + (float)dv[j] + t.x + t.y + (float)iv[j];
    }
    if (acc == 123.456f) out[0] = acc;    // keep the chains alive
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Spec { const char* name; int op; double mufu, fp32, other; };   // thread-ops per chain step

template <int OP>
static void run(const Spec& s, int sms, int blocks_per_sm, int threads, float* d_out, long long* d_cyc, std::string& json) {
    const int grid = sms * blocks_per_sm;
    const int iters = 200;
    bench_kernel<OP><<<grid, threads>>>(d_out, 8, 1.0f, d_cyc);            // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best_ms = 1e30f; double best_cyc = 0;
    std::vector<long long> cyc(grid);
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(a));
        bench_kernel<OP><<<grid, threads>>>(d_out, iters, 1.0f, d_cyc);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best_ms) {
            best_ms = ms;
            CK(cudaMemcpy(cyc.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
            double m = 0; for (long long c : cyc) m = c > m ? c : m;
            best_cyc = m;
        }
    }
    const double steps = (double)iters * INNER * ILP * threads * blocks_per_sm;     // chain steps per SM
    const double per_clk = steps / best_cyc;                                         // steps/clk/SM (in-kernel clock)
    const double mhz = best_cyc / (best_ms * 1e3);                                    // effective SM clock
    char buf[512];
    snprintf(buf, sizeof buf,
             "  \"%s\": {\"steps_per_clk_sm\": %.3f, \"mufu_per_clk_sm\": %.3f, \"fp32_per_clk_sm\": %.3f, "
             "\"other_per_clk_sm\": %.3f, \"ms\": %.4f, \"sm_mhz\": %.0f, \"threads_per_sm\": %d},\n",
             s.name, per_clk, per_clk * s.mufu, per_clk * s.fp32, per_clk * s.other, best_ms, mhz, threads * blocks_per_sm);
    json += buf;
    CK(cudaEventDestroy(a)); CK(cudaEventDestroy(b));
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    const int sms = p.multiProcessorCount;
    float* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, 1024)); CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * 16));
    std::string json = "{\n";
    char hdr[768];
    snprintf(hdr, sizeof hdr, "  \"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_rate_mhz\": %.0f,\n",
             p.name, sms, p.major, p.minor, clk_khz / 1e3);
    json += hdr;
    const int bps = 4, th = 256;      // 1024 threads/SM = 8 warps per SMSP, ILP 8 each
#define RUN(OPNAME, m, f, o) { Spec s{#OPNAME, OPNAME, m, f, o}; run<OPNAME>(s, sms, bps, th, d_out, d_cyc, json); }
    const bool quick = argc > 2 && strcmp(argv[2], "--quick") == 0;      // MUFU + FP32 peaks only (bench.py)
    RUN(EX2, 1, 0, 0) RUN(LG2, 1, 0, 0) RUN(MUFU_MIX, 1, 0, 0) RUN(FFMA, 0, 1, 0)
    if (!quick) {
    RUN(RCP, 1, 1, 0)
    RUN(FADD, 0, 1, 0) RUN(FFMA2, 0, 2, 0) RUN(FMNMX, 0, 0, 1) RUN(FSEL, 0, 0, 2)
    RUN(DFMA, 0, 0, 1) RUN(D2F, 0, 1, 2) RUN(I2F, 0, 1, 2) RUN(SHFL, 0, 0, 1)
    RUN(POPC, 0, 0, 2) RUN(EX2_POPC, 1, 0, 2) RUN(VOTE, 0, 0, 3) RUN(EX2_VOTE, 1, 0, 3) RUN(REDUX, 0, 0, 3) RUN(EX2_REDUX, 1, 0, 3)
    RUN(EX2_FFMA1, 1, 1, 0) RUN(EX2_FFMA2, 1, 2, 0) RUN(EX2_FFMA4, 1, 4, 0) RUN(EX2_FFMA6, 1, 6, 0) RUN(EX2_FFMA8, 1, 8, 0)
    RUN(EX2_FFMA2X4, 1, 8, 0) RUN(EX2_D2F, 1, 1, 2) RUN(EX2_DFMA, 1, 0, 1) RUN(EX2_FMNMX4, 1, 0, 4) RUN(POW_CHAIN, 2, 2, 0)
    }
    json += "  \"note\": \"thread-ops per clock per SM from in-kernel clock64(); sm_mhz = cycles / event time\"\n}\n";
    fputs(json.c_str(), stdout);
    if (argc > 1) { FILE* f = fopen(argv[1], "w"); if (f) { fputs(json.c_str(), f); fclose(f); } }
    return 0;
}
