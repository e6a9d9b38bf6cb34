// Microbenchmark: cost of a cooperative-groups grid barrier on this GPU for the launch shapes the
// routing kernel uses (CTAs of 256 threads, 45 KB static shared memory, 128 registers).
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(256, 2) k(int n, double *sink)
{
    __shared__ double pad[5600];
    pad[threadIdx.x] = threadIdx.x;
    cg::grid_group g = cg::this_grid();
    double acc = 0;
    for (int i = 0; i < n; i++) {
        acc += pad[(threadIdx.x + i) & 255];
        g.sync();
    }
    if (acc == -1.0) sink[0] = acc;
}

int main()
{
    double *sink; cudaMalloc(&sink, 8);
    int sizes[] = {1, 8, 78, 148, 296};
    for (int b : sizes) {
        int n = 2000;
        void *args[] = {&n, &sink};
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaLaunchCooperativeKernel((void *)k, dim3(b), dim3(256), args, 0, 0);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        cudaError_t err = cudaLaunchCooperativeKernel((void *)k, dim3(b), dim3(256), args, 0, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("blocks %3d: %.2f us per grid.sync (%s)\n", b, 1000.0 * ms / n, cudaGetErrorString(err));
    }
    return 0;
}
