// Latency / issue-interval microbenchmarks for the ops on the error-diffusion critical path.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int OP, int ILP>
__global__ void k(double *out, long long *cyc, double seed, float fseed)
{
    double d[ILP];
    float f[ILP];
    int ii[ILP];
    for (int j = 0; j < ILP; ++j) { d[j] = seed + j; f[j] = fseed + j; ii[j] = (int)fseed + j; }
    __shared__ double sm[1024];
    for (int j = threadIdx.x; j < 1024; j += blockDim.x) sm[j] = (double)((j * 7) & 1023);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N / 8; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                if (OP == 0) d[j] = __dadd_rn(d[j], seed);
                if (OP == 1) d[j] = __dmul_rn(d[j], seed);
                if (OP == 2) d[j] = __fma_rn(d[j], seed, seed);
                if (OP == 3) { f[j] = __double2float_rn(d[j]); d[j] = (double)f[j]; }  // 2 cvts
                if (OP == 4) f[j] = __fadd_rn(f[j], fseed);
                if (OP == 5) { ii[j] = __float2int_rz(f[j]); f[j] = __int_as_float(ii[j] | 0x3f800000); }
                if (OP == 6) f[j] = __shfl_up_sync(0xffffffffu, f[j], 1);
                if (OP == 7) { ii[j] = (int)sm[ii[j] & 1023]; }   // LDS.64 + F2I.F64
                if (OP == 8) { ii[j] = ((int *)sm)[ii[j] & 2047]; }   // LDS.32 chain
                if (OP == 9) { ii[j] = min(ii[j], (int)fseed) + 1; }
                if (OP == 10) { d[j] = __dadd_rn((double)f[j], d[j]); f[j] = __double2float_rn(d[j]); }  // acc_f64
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < ILP; ++j) s += d[j] + f[j] + ii[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP, int ILP>
void run(const char *name, int warps = 1)
{
    double *out; long long *cyc;
    cudaMalloc(&out, 8 * 1024 * 64); cudaMalloc(&cyc, 8 * 64);
    k<OP, ILP><<<1, 32 * warps>>>(out, cyc, 1.000001, 1.5f);
    k<OP, ILP><<<1, 32 * warps>>>(out, cyc, 1.000001, 1.5f);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s ILP=%d warps=%d: %.2f cycles per op-slot (per warp: %.2f / op)\n", name, ILP, warps,
           (double)h / N, (double)h / N / ILP);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<0, 1>("DADD"); run<0, 8>("DADD"); run<0, 8>("DADD", 4); run<0, 8>("DADD", 16);
    run<1, 1>("DMUL"); run<1, 8>("DMUL");
    run<2, 1>("DFMA"); run<2, 8>("DFMA"); run<2, 8>("DFMA", 16);
    run<3, 1>("F2F.32<-64 + F2F.64<-32"); run<3, 8>("F2F.32<-64 + F2F.64<-32"); run<3, 8>("F2F pair", 4); run<3, 8>("F2F pair", 16);
    run<4, 1>("FADD"); run<4, 8>("FADD");
    run<5, 1>("F2I + LOP"); run<5, 8>("F2I + LOP");
    run<6, 1>("SHFL.UP"); run<6, 8>("SHFL.UP");
    run<7, 1>("LDS.64 + F2I.F64"); run<7, 8>("LDS.64 + F2I.F64");
    run<8, 1>("LDS.32 chain"); run<8, 8>("LDS.32 chain");
    run<9, 1>("IMNMX + IADD"); run<9, 8>("IMNMX + IADD");
    run<10, 1>("acc_f64 (F2F,DADD,F2F)"); run<10, 8>("acc_f64"); run<10, 8>("acc_f64", 4);
    return 0;
}
