// fp64_pipes.cu -- micro-benchmark: FP64 FMA pipe vs FP64 tensor-core MMA (mma.sync m8n8k4 f64) on sm_100a, alone and mixed.
// Question it answers: the accumulation acc[pair][T] += w[pair][node] * B[node][T] of the temperature-shared
// kernel builder is a small dense FP64 contraction; does DMMA run at (at least) the DFMA rate on B200, and do the two
// overlap when issued from the same warp / from different warps of an SM?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode 0: DFMA only (8 chains).  1: DMMA only (NACC independent accumulator tiles).  2: both in every warp, 8 DFMA per DMMA
// (equal flop).  3: warp-specialised: even warps DMMA, odd warps DFMA.  4: both in every warp, 16 DFMA per DMMA.
template <int MODE, int NACC>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    const bool do_mma = MODE == 1 || MODE == 2 || MODE == 4 || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 0);
    const bool do_fma = MODE == 0 || MODE == 2 || MODE == 4 || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 1);
    const double fa = a + 1e-9 * threadIdx.x, fb = b;
    for (int it = 0; it < iters; ++it) {
        if (do_mma) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], fa, fb);
        }
        if (do_fma) {
            constexpr int R = (MODE == 4) ? 2 * NACC : NACC;
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NACC>
static void run(const char* name, int sms, double* d, int blocks_per_sm) {
    const int blocks = sms * blocks_per_sm, threads = 256, iters = 1 << 13;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE, NACC><<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double warps = (double)blocks * threads / 32.0;
    double mma_w = (MODE == 1 || MODE == 2 || MODE == 4) ? warps : (MODE == 3 ? warps / 2 : 0);
    double fma_w = (MODE == 0 || MODE == 2 || MODE == 4) ? warps : (MODE == 3 ? warps / 2 : 0);
    const double fl_mma = mma_w * iters * NACC * 512.0;
    const double fl_fma = fma_w * iters * ((MODE == 4) ? 2 * NACC : NACC) * 8.0 * 64.0;
    printf("%-44s blocks/SM %d  %8.3f ms   DMMA %7.2f TF/s   DFMA %7.2f TF/s   sum %7.2f TF/s  (%s)\n", name, blocks_per_sm, best,
           fl_mma / best / 1e9, fl_fma / best / 1e9, (fl_mma + fl_fma) / best / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs\n", p.name, sms);
    double* d; cudaMalloc(&d, sizeof(double) * sms * 8 * 256);
    for (int bps : {2, 4, 8}) {
        run<0, 8>("DFMA only (8 chains x 8)", sms, d, bps);
        run<1, 4>("DMMA only, 4 acc tiles/warp", sms, d, bps);
        run<1, 8>("DMMA only, 8 acc tiles/warp", sms, d, bps);
        run<1, 16>("DMMA only, 16 acc tiles/warp", sms, d, bps);
        run<2, 8>("mixed same warp, 8 DMMA + 64 DFMA", sms, d, bps);
        run<4, 8>("mixed same warp, 8 DMMA + 128 DFMA", sms, d, bps);
        run<3, 8>("warp-specialised (even DMMA, odd DFMA)", sms, d, bps);
    }
    cudaFree(d);
    return 0;
}
