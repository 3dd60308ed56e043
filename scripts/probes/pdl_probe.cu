// pdl_probe.cu -- does programmatic dependent launch overlap kernels on this box (eager stream and
// captured graph)?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a pdl_probe.cu -o pdl_probe
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void spin_us(float us)
{
    const long long t0 = clock64();
    const long long dt = (long long)(us * 1900.0f);
    while (clock64() - t0 < dt) {}
}

__global__ void kernelA(int trigger, float us, int *sink)
{
    if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    spin_us(us);
    if (sink && threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}

__global__ void kernelB(int trigger, float pre_us, float post_us, int *sink)
{
    if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    spin_us(pre_us);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    spin_us(post_us);
    if (sink && threadIdx.x == 0 && blockIdx.x == 0) *sink = 2;
}

template <typename... KArgs, typename... Args>
static void launch(bool pdl, void (*k)(KArgs...), dim3 g, dim3 b, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = g; cfg.blockDim = b; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k, KArgs(args)...);
}

static void chain(bool pdl, int trigger, cudaStream_t s, int *sink, int gridA, int gridB)
{
    launch(pdl, kernelA, dim3(gridA), dim3(256), s, trigger, 20.0f, sink);
    launch(pdl, kernelB, dim3(gridB), dim3(256), s, trigger, 15.0f, 5.0f, sink);
    launch(pdl, kernelB, dim3(gridB), dim3(256), s, trigger, 15.0f, 5.0f, sink);
    launch(pdl, kernelB, dim3(8), dim3(1024), s, trigger, 2.0f, 10.0f, sink);
}

int main()
{
    int *sink; cudaMalloc(&sink, 4);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 200;
    for (int gridA : {148, 600, 2400}) {
        for (int mode = 0; mode < 3; ++mode) {     // 0: plain, 1: PDL attr without trigger, 2: PDL + trigger
            const bool pdl = mode > 0; const int trig = mode == 2;
            for (int i = 0; i < 20; ++i) chain(pdl, trig, s, sink, gridA, 296);
            cudaStreamSynchronize(s);
            cudaEventRecord(e0, s);
            for (int i = 0; i < reps; ++i) chain(pdl, trig, s, sink, gridA, 296);
            cudaEventRecord(e1, s);
            cudaStreamSynchronize(s);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            // graph
            cudaGraph_t g; cudaGraphExec_t ge;
            cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
            cudaMemsetAsync(sink, 0, 4, s);
            chain(pdl, trig, s, sink, gridA, 296);
            cudaStreamEndCapture(s, &g);
            cudaError_t err = cudaGraphInstantiate(&ge, g, 0);
            for (int i = 0; i < 20; ++i) cudaGraphLaunch(ge, s);
            cudaStreamSynchronize(s);
            cudaEventRecord(e0, s);
            for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, s);
            cudaEventRecord(e1, s);
            cudaStreamSynchronize(s);
            float gms; cudaEventElapsedTime(&gms, e0, e1);
            printf("gridA %4d mode %d (%s): eager %.1f us/chain, graph(+memset) %.1f us/chain  [%s] serial sum = 20+20+20+12 = 72 us\n",
                   gridA, mode, mode == 0 ? "plain" : mode == 1 ? "pdl attr, no trigger" : "pdl attr + trigger",
                   ms * 1e3f / reps, gms * 1e3f / reps, cudaGetErrorString(err));
            cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
