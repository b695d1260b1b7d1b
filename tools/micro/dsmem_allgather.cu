// All-gather of a state vector between the CTAs of a thread-block cluster through distributed shared memory: cycles per
// exchange for the ways of moving the bytes.  Every CTA owns U doubles and needs all C*U of them, every step.
//   V0  one st.async (8 bytes + complete_tx) per thread and peer           (recur_cluster.cuh, first version)
//   V1  the same, but lanes pair up: one 16-byte st.async.v2 per two units
//   V2  one cp.async.bulk shared::cta -> shared::cluster per peer (bulk copy of the CTA's slice, issued by C-1 threads
//       after a CTA barrier + proxy fence)
//   V3  V0 with a scattered placement (slot = bit-reversed-ish permutation): what a bank-conflict-avoiding layout costs
//   V4  V0 with the units of a warp permuted inside their own 32-slot segment (a placement that keeps a warp's stores
//       inside one 256-byte segment)
//   V5/V6/V7  V0 sending only to the peers that need the value: every (unit, peer) pair with probability 0.7 / 0.5 / 0.3
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_allgather dsmem_allgather.cu ; run without arguments.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ unsigned cta_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

template <int V>
__global__ void __launch_bounds__(512, 1) allgather(int U, int C, int iters, long long* cycles, double* check) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const unsigned c = cta_rank();
    const int n_pad = (C * U + 15) / 16 * 16 + 16;
    double* buf = reinterpret_cast<double*>(smem);                     // [3][n_pad]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(buf + 3 * n_pad);
    const unsigned buf_addr = smem_u32(buf), bar_addr = smem_u32(bars), copy_bytes = n_pad * 8u;
    if (tid < 3) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar_addr + 8u * tid), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    __syncthreads();
    cluster_sync();
    const bool owner = tid < U;
    const int unit = c * U + tid;
    int slot = unit;
    if (V == 3) {   // scatter inside the CTA's slice: stride-17 walk (a permutation of [0,U) when gcd(17,U)=1, else natural)
        const int s = (U % 17) ? (int)(((long long)tid * 17) % U) : tid;
        slot = c * U + s;
    }
    if (V == 4) slot = c * U + (tid & ~31) + ((tid & 31) * 5 + 3) % 32;   // U a multiple of 32 here
    // V5..V7: which peers get this unit (hash of (unit, peer) against the density)
    unsigned peer_mask = 0xffu;
    if (V >= 5) {
        const unsigned thr = V == 5 ? 179u : (V == 6 ? 128u : 77u);   // of 256
        peer_mask = 0;
        for (int q = 0; q < C; ++q) {
            unsigned h = (unsigned)unit * 2654435761u + (unsigned)q * 40503u;
            h ^= h >> 15, h *= 2246822519u, h ^= h >> 13;
            if ((h & 255u) < thr) peer_mask |= 1u << q;
        }
    }
    __shared__ unsigned s_expect;
    if (tid == 0) s_expect = 0;
    __syncthreads();
    double v = unit;
    unsigned cur = 0, phase_bits = 0;
    unsigned expect = (unsigned)(C - 1) * U * 8u;
    if (V >= 5) {   // what the peers will send me: recompute their masks
        unsigned mine = 0;
        for (int i = tid; i < C * U; i += blockDim.x) {
            if (i / U == (int)c) continue;
            const unsigned thr = V == 5 ? 179u : (V == 6 ? 128u : 77u);
            unsigned h = (unsigned)i * 2654435761u + c * 40503u;
            h ^= h >> 15, h *= 2246822519u, h ^= h >> 13;
            if ((h & 255u) < thr) mine += 8u;
        }
        atomicAdd(&s_expect, mine);
        __syncthreads();
        expect = s_expect;
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const unsigned nxt = cur == 2 ? 0u : cur + 1u;
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_addr + 8u * nxt), "r"(expect) : "memory");
        const unsigned dst = buf_addr + nxt * copy_bytes;
        v = v * 1.0000001 + 1.0;
        if (owner) {
            asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(dst + (unsigned)slot * 8u), "d"(v) : "memory");
            if (V == 0 || V >= 3) {
                for (unsigned q = 0; q < (unsigned)C; ++q)
                    if (q != c && ((peer_mask >> q) & 1u))
                        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];\n" ::"r"(mapa(dst + (unsigned)slot * 8u, q)),
                                     "l"(__double_as_longlong(v)), "r"(mapa(bar_addr + 8u * nxt, q)) : "memory");
            }
        }
        if (V == 1) {
            const double other = __shfl_down_sync(0xffffffffu, v, 1);
            if (owner && !(tid & 1)) {
                if (tid + 1 < U) {
                    for (unsigned q = 0; q < (unsigned)C; ++q)
                        if (q != c)
                            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];\n" ::"r"(mapa(dst + (unsigned)slot * 8u, q)),
                                         "l"(__double_as_longlong(v)), "l"(__double_as_longlong(other)), "r"(mapa(bar_addr + 8u * nxt, q)) : "memory");
                } else {
                    for (unsigned q = 0; q < (unsigned)C; ++q)
                        if (q != c)
                            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];\n" ::"r"(mapa(dst + (unsigned)slot * 8u, q)),
                                         "l"(__double_as_longlong(v)), "r"(mapa(bar_addr + 8u * nxt, q)) : "memory");
                }
            }
        }
        if (V == 2) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            __syncthreads();
            if (tid < C && (unsigned)tid != c) {
                const unsigned src = dst + c * U * 8u;
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(mapa(src, tid)),
                             "r"(src), "r"((unsigned)U * 8u), "r"(mapa(bar_addr + 8u * nxt, tid)) : "memory");
            }
        }
        mbar_wait(bar_addr + 8u * nxt, (phase_bits >> nxt) & 1u);
        phase_bits ^= 1u << nxt;
        __syncthreads();
        // consume something from the new copy so the loads are real
        v += reinterpret_cast<double*>(smem)[nxt * n_pad + (tid * 7 + it) % (C * U)] * 1e-30;
        cur = nxt;
    }
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    // correctness: after the last exchange every CTA holds the same vector
    if (tid == 0) {
        double s = 0;
        for (int i = 0; i < C * U; ++i) s += buf[cur * n_pad + i];
        check[blockIdx.x] = s;
    }
    cluster_sync();
}

template <int V>
void run(int U, int C, int clusters, const char* name) {
    const int iters = 4000;
    long long* d_cyc;
    double* d_chk;
    cudaMalloc(&d_cyc, sizeof(long long) * C * clusters);
    cudaMalloc(&d_chk, sizeof(double) * C * clusters);
    const int n_pad = (C * U + 15) / 16 * 16 + 16;
    const size_t smem = 3ull * n_pad * 8 + 64;
    cudaFuncSetAttribute(allgather<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C * clusters), cfg.blockDim = dim3(512), cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, allgather<V>, U, C, iters, d_cyc, d_chk);
        if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) {
            printf("%-34s U=%3d C=%d clusters=%2d: %s\n", name, U, C, clusters, cudaGetErrorString(e));
            return;
        }
    }
    long long cyc[256];
    double chk[256];
    cudaMemcpy(cyc, d_cyc, sizeof(long long) * C * clusters, cudaMemcpyDeviceToHost);
    cudaMemcpy(chk, d_chk, sizeof(double) * C * clusters, cudaMemcpyDeviceToHost);
    bool same = true;
    for (int i = 1; i < C; ++i) same = same && chk[i] == chk[0];
    long long mx = 0;
    for (int i = 0; i < C * clusters; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
    printf("%-34s U=%3d C=%d clusters=%2d: %7.0f cycles per exchange  (%5.1f B/cycle if everything is sent)  %s\n", name, U, C, clusters,
           (double)mx / iters, (double)(C - 1) * U * 8 / ((double)mx / iters), V >= 5 ? "" : (same ? "ok" : "MISMATCH"));
    cudaFree(d_cyc), cudaFree(d_chk);
}

int main() {
    const int shapes[][3] = {{512, 4, 1}, {512, 4, 16}, {128, 8, 1}, {256, 8, 1}, {256, 8, 16}, {512, 8, 8}, {512, 2, 1}};
    for (auto& s : shapes) {
        run<0>(s[0], s[1], s[2], "V0 st.async b64 per thread+peer");
        run<1>(s[0], s[1], s[2], "V1 st.async v2.b64 per lane pair");
        run<2>(s[0], s[1], s[2], "V2 cp.async.bulk per peer");
        run<3>(s[0], s[1], s[2], "V3 V0, scattered slots");
        run<4>(s[0], s[1], s[2], "V4 V0, permuted inside 32-slot segments");
        run<5>(s[0], s[1], s[2], "V5 V0, 70% of (unit, peer) pairs");
        run<6>(s[0], s[1], s[2], "V6 V0, 50% of (unit, peer) pairs");
        run<7>(s[0], s[1], s[2], "V7 V0, 30% of (unit, peer) pairs");
    }
    return 0;
}
