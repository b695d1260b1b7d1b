// Ping-pong latency between two CTAs (different SMs) through global memory, for several flavours of the
// polling load and the publishing store.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pingpong pingpong.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int LD>
__device__ __forceinline__ unsigned long long poll(const unsigned long long* p) {
    unsigned long long v;
    if (LD == 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (LD == 1) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (LD == 2) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (LD == 3) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (LD == 4) asm volatile("atom.global.add.u64 %0, [%1], 0;" : "=l"(v) : "l"(p) : "memory");
    if (LD == 5) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (LD == 6) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
template <int ST>
__device__ __forceinline__ void publish(unsigned long long* p, unsigned long long v) {
    if (ST == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (ST == 1) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (ST == 2) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (ST == 3) asm volatile("atom.global.exch.b64 %1, [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (ST == 4) asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (ST == 5) asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}

// CTA 0 and CTA `peer` bounce a counter: a[0] written by CTA 0, a[32] (another line) by the peer
template <int LD, int ST>
__global__ void pingpong(unsigned long long* a, int iters, int peer, long long* cycles, int pollers) {
    if (blockIdx.x != 0 && blockIdx.x != peer) {
        // optional background pollers (flood): every other CTA polls a[0] with all its threads
        if ((int)blockIdx.x < pollers + 1 && blockIdx.x != peer)
            while (poll<LD>(a) < (unsigned long long)iters) {}
        return;
    }
    if (threadIdx.x != 0) return;
    const bool first = blockIdx.x == 0;
    unsigned long long* mine = first ? a : a + 32;
    const unsigned long long* theirs = first ? a + 32 : a;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (first) {
            publish<ST>(mine, (unsigned long long)i);
            while (poll<LD>(theirs) < (unsigned long long)i) {}
        } else {
            while (poll<LD>(theirs) < (unsigned long long)i) {}
            publish<ST>(mine, (unsigned long long)i);
        }
    }
    if (first) *cycles = clock64() - t0;
}

template <int LD, int ST>
void run(const char* name, unsigned long long* a, long long* cyc, int pollers) {
    const int iters = 2000;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int peer : {1, sms / 2, sms - 1}) {
        cudaMemset(a, 0, 1024);
        void* args[] = {&a, (void*)&iters, &peer, &cyc, &pollers};
        cudaLaunchCooperativeKernel((void*)pingpong<LD, ST>, dim3(sms), dim3(ST == 5 ? 32 : 128), args, 0, 0);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-34s peer %3d pollers %3d: %7.0f cycles per round trip (2 hops)%s\n", name, peer, pollers, (double)c / iters,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}

int main() {
    unsigned long long* a;
    long long* cyc;
    cudaMalloc(&a, 1024);
    cudaMalloc(&cyc, 8);
    for (int pollers : {0, 140}) {
        run<0, 0>("ld.relaxed.gpu / st.relaxed.gpu", a, cyc, pollers);
        run<1, 1>("ld.acquire.gpu / st.release.gpu", a, cyc, pollers);
        run<2, 2>("ld.volatile / st.volatile", a, cyc, pollers);
        run<3, 0>("ld.cv / st.relaxed.gpu", a, cyc, pollers);
        run<4, 0>("atom.add 0 / st.relaxed.gpu", a, cyc, pollers);
        run<0, 3>("ld.relaxed.gpu / atom.exch", a, cyc, pollers);
        run<4, 3>("atom.add 0 / atom.exch", a, cyc, pollers);
        run<6, 4>("ld.cg / st.cg", a, cyc, pollers);
        run<5, 0>("ld.relaxed.sys / st.relaxed.gpu", a, cyc, pollers);
    }
    return 0;
}
