// Microbenchmark: issue rate of FFMA / FADD / FMUL versus their packed f32x2 forms on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed)
{
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + i + threadIdx.x; b[i] = seed * 0.5f + i; }
    const float c = seed * 1.0001f, d = seed * 0.9999f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) {            // 2 scalar FFMA
                a[i] = fmaf(a[i], c, d);
                b[i] = fmaf(b[i], c, d);
            } else if (MODE == 1) {     // 1 packed FFMA2
                unsigned long long x, cc, dd;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(b[i]));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(dd) : "f"(d));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(cc), "l"(dd));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(b[i]) : "l"(x));
            } else if (MODE == 2) {     // 2 scalar FADD
                a[i] = a[i] + c;
                b[i] = b[i] + d;
            } else if (MODE == 3) {     // 1 packed FADD2
                unsigned long long x, cc;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(b[i]));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c), "f"(d));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(cc));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(b[i]) : "l"(x));
            } else if (MODE == 4) {     // FADD + FFMA mix typical of butterflies: a' = a + b ; b' = fma(b, c, a)
                float t = a[i] + b[i];
                b[i] = fmaf(b[i], c, a[i]);
                a[i] = t;
            } else if (MODE == 5) {     // same on packed pairs (a[i], a[i^1])... pairs (a,b) as one vector each
                unsigned long long x, y, cc, t;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[(i + 4) & 7]));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b[i]), "f"(b[(i + 4) & 7]));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(x), "l"(y));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y) : "l"(cc), "l"(x));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[(i + 4) & 7]) : "l"(t));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(b[i]), "=f"(b[(i + 4) & 7]) : "l"(y));
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, double flops_per_iter_thread)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    float* out;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double total = flops_per_iter_thread * iters * (double)blocks * threads;
    printf("%-28s %8.3f ms  %8.2f Tflop/s  (%s)\n", name, ms, total / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    run<0>("FFMA scalar", 8 * 2 * 2.0);
    run<1>("FFMA2 packed", 8 * 2 * 2.0);
    run<2>("FADD scalar", 8 * 2 * 1.0);
    run<3>("FADD2 packed", 8 * 2 * 1.0);
    run<4>("FADD+FFMA scalar", 8 * 3.0);
    run<5>("FADD2+FFMA2 packed", 8 * 3.0);
    return 0;
}
