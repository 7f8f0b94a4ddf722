// launch_floor.cu -- what a chain of dependent kernels costs per link inside a CUDA graph on this GPU,
// with and without programmatic dependent launch: the floor under the per-step API (one launch per
// env step).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o launch_floor launch_floor.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void link_kernel(unsigned long long *x, int n, int pdl)
{
    if (pdl) {
        asm volatile("griddepcontrol.launch_dependents;");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += 1ull;
}

static float run(int n, bool pdl, int links, int reps)
{
    unsigned long long *x;
    cudaMalloc(&x, sizeof(unsigned long long) * n);
    cudaMemset(x, 0, sizeof(unsigned long long) * n);
    cudaStream_t s;
    cudaStreamCreate(&s);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
    for (int k = 0; k < links; ++k) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((n + 255) / 256); cfg.blockDim = dim3(256); cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
        cudaLaunchKernelEx(&cfg, link_kernel, x, n, pdl ? 1 : 0);
    }
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, s);
    cudaStreamSynchronize(s);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, s);
    for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, s);
    cudaEventRecord(b, s);
    cudaStreamSynchronize(s);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    unsigned long long h = 0;
    cudaMemcpy(&h, x, sizeof h, cudaMemcpyDeviceToHost);
    if (h != (unsigned long long)links * (reps + 1)) printf("{\"error\": \"chain result %llu\"}\n", h);
    cudaFree(x);
    return ms * 1e3f / (links * reps);
}

int main()
{
    printf("{");
    const int sizes[3] = {16384, 65536, 1048576};
    for (int i = 0; i < 3; ++i)
        printf("%s\"n%d_us_per_link\": {\"plain\": %.3f, \"pdl\": %.3f}", i ? ", " : "", sizes[i],
               run(sizes[i], false, 64, 20), run(sizes[i], true, 64, 20));
    printf(", \"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
