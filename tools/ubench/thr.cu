// Per-SM throughput of the integer ops the threshold kernels are made of (one block, 32 warps).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o thr thr.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
template <int OP>
__global__ void k(unsigned *out, long long *cyc, unsigned s0, unsigned s1)
{
    __shared__ unsigned sm[4096];
    for (int j = threadIdx.x; j < 4096; j += blockDim.x) sm[j] = (j * 2654435761u) >> 20;
    __syncthreads();
    unsigned x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = s0 + threadIdx.x * 977u + j * 131u;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (OP == 0) x[j] = __dp4a(x[j], s1, x[j]);
                if (OP == 1) x[j] = __vabsdiffu4(x[j], s1);
                if (OP == 2) x[j] = (unsigned)min((int)x[j], (int)(s1 + j));
                if (OP == 3) x[j] = (unsigned)min(min((int)x[j], (int)(s1 + j)), (int)(s0 ^ j));
                if (OP == 4) x[j] = __byte_perm(x[j], s1, 0x4441 + j);
                if (OP == 5) x[j] = (x[j] & s1) | (s0 & ~s1) ^ j;
                if (OP == 6) x[j] = __funnelshift_l(x[j], s1, 7);
                if (OP == 7) x[j] = x[j] * s1 + s0;
                if (OP == 8) x[j] = __float_as_uint(__int2float_rn((int)x[j]));
                if (OP == 9) x[j] = sm[x[j] & 4095];
                if (OP == 10) x[j] = sm[(x[j] & 15) + 16 * j] + x[j];
                if (OP == 11) x[j] = __float_as_uint(fmaf(__uint_as_float(x[j]), 1.0001f, 0.5f));
                if (OP == 12) x[j] = x[j] + s1 + j;
            }
        }
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char *name)
{
    unsigned *out; long long *cyc;
    cudaMalloc(&out, 4 * 1024 * 4); cudaMalloc(&cyc, 8 * 4);
    for (int r = 0; r < 2; ++r) k<OP><<<1, 1024>>>(out, cyc, 12345u, 0x01020304u);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double ops = 1024.0 * N * 4 * 8;
    printf("%-22s %.1f lane-ops/clk/SM  (%.2f warp-instr/clk/SM)\n", name, ops / h, ops / h / 32);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<0>("IDP.4A"); run<1>("VABSDIFF4"); run<2>("VIMNMX"); run<3>("VIMNMX3"); run<4>("PRMT");
    run<5>("LOP3"); run<6>("SHF"); run<7>("IMAD"); run<8>("I2FP"); run<9>("LDS.32 random");
    run<10>("LDS.32 16 entries"); run<11>("FFMA"); run<12>("IADD3");
    return 0;
}
