// pdl_probe2.cu -- does a kernel that executes griddepcontrol.launch_dependents let the NEXT stream
// operation start early when that operation was NOT launched as a programmatic dependent?
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

__device__ __forceinline__ void spin_us(float us)
{
    const long long t0 = clock64();
    const long long dt = (long long)(us * 1900.0f);
    while (clock64() - t0 < dt) {}
}

__global__ void producer(int trigger, int *out, int value)
{
    if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    spin_us(10.0f + (blockIdx.x % 7) * 5.0f);
    out[blockIdx.x * blockDim.x + threadIdx.x] = value;
}

__global__ void consumer(const int *in, int *out2, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out2[i] = in[i];
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

int main()
{
    const int G = 600, T = 256, N = G * T;
    int *out, *out2, *h; cudaMalloc(&out, N * 4); cudaMalloc(&out2, N * 4); cudaMallocHost(&h, N * 4);
    for (int use_legacy = 0; use_legacy < 2; ++use_legacy) {
        cudaStream_t s = 0;
        if (!use_legacy) cudaStreamCreate(&s);
        for (int trigger = 0; trigger < 2; ++trigger)
            for (int attr = 0; attr < 2; ++attr)
                for (int next = 0; next < 3; ++next) {   // 0: <<<>>> kernel, 1: D2H memcpy, 2: kernel via LaunchKernelEx w/o attr
                    int bad_runs = 0;
                    for (int it = 1; it <= 50; ++it) {
                        cudaMemsetAsync(out, 0, N * 4, s);
                        launch(attr, producer, dim3(G), dim3(T), s, trigger, out, it);
                        if (next == 0) { consumer<<<G, T, 0, s>>>(out, out2, N); cudaMemcpyAsync(h, out2, N * 4, cudaMemcpyDeviceToHost, s); }
                        else if (next == 1) cudaMemcpyAsync(h, out, N * 4, cudaMemcpyDeviceToHost, s);
                        else { launch(false, consumer, dim3(G), dim3(T), s, (const int *)out, out2, N); cudaMemcpyAsync(h, out2, N * 4, cudaMemcpyDeviceToHost, s); }
                        cudaStreamSynchronize(s);
                        int bad = 0;
                        for (int i = 0; i < N; ++i) bad += h[i] != it;
                        bad_runs += bad != 0;
                    }
                    printf("stream %s trigger %d producer-attr %d next %s: %d / 50 runs read stale data\n",
                           use_legacy ? "legacy-0" : "created", trigger, attr,
                           next == 0 ? "kernel<<<>>>" : next == 1 ? "memcpyD2H" : "kernelEx(no attr)", bad_runs);
                }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
