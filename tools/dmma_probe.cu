// dmma_probe.cu -- issue-rate probe of the FP64 mma.sync shapes on sm_100a (not part of the product).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int NACC>
__global__ void probe(double *out, int iters)
{
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 2e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (SHAPE == 884)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
            if (SHAPE == 1684)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            if (SHAPE == 1688)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            if (SHAPE == 16816)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int NACC>
void run(const char *name, double flops_per_mma, int warps_per_cta)
{
    double *out;
    cudaMalloc(&out, 148 * 4 * 1024 * sizeof(double));
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ctas = 1; ctas <= 2; ++ctas) {
        probe<SHAPE, NACC><<<148 * ctas, warps_per_cta * 32>>>(out, 10);
        cudaEventRecord(e0);
        probe<SHAPE, NACC><<<148 * ctas, warps_per_cta * 32>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double total = (double)148 * ctas * warps_per_cta * iters * NACC * flops_per_mma;
        printf("%-10s acc/warp=%2d warps/cta=%2d ctas/sm=%d : %.2f TFLOP/s  err=%s\n", name, NACC, warps_per_cta, ctas,
               total / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(out);
}

int main()
{
    run<884, 8>("m8n8k4", 2.0 * 8 * 8 * 4, 8);
    run<884, 16>("m8n8k4", 2.0 * 8 * 8 * 4, 8);
    run<884, 8>("m8n8k4", 2.0 * 8 * 8 * 4, 16);
    run<1684, 8>("m16n8k4", 2.0 * 16 * 8 * 4, 8);
    run<1688, 8>("m16n8k8", 2.0 * 16 * 8 * 8, 8);
    run<16816, 8>("m16n8k16", 2.0 * 16 * 8 * 16, 8);
    run<16816, 4>("m16n8k16", 2.0 * 16 * 8 * 16, 8);
    run<16816, 8>("m16n8k16", 2.0 * 16 * 8 * 16, 4);
    return 0;
}
