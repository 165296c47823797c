// Micro-benchmark (development aid): what does this HBM give a stream that READS a bytes and WRITES b bytes per unit, at the sizes of
// the BASELINE workloads, with NO arithmetic?  The knot kernel's roofline denominator (MEASURED_PEAKS.json) is a 1:1 copy of 2 GiB;
// the Jacobian workloads are write-dominated (C2 40 B in / 160 B out per knot, C3 68 / 884, C4 144 / 1728) and ~200 MB per launch.
//
//   mode 0  "ldst":  grid-stride 16-byte loads / stores from registers (flat arrays, ratio in:out kept by the index arithmetic)
//   mode 1  "bulk":  the knot kernel's skeleton — persistent CTAs, one TMA bulk load of a tile's input rows into a double-buffered smem
//                    image (mbarrier), one TMA bulk store of the tile's output image (written once at start, never recomputed), the
//                    next tile's smem "assembly" replaced by nothing but the wait for the previous store's smem read
//   mode 2  "bulk + touch": as 1, but every thread also re-writes its part of the output image per tile (st.shared + fence.proxy.async),
//                    i.e. everything the real kernel does except the arithmetic
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/stream_mix scripts/micro/stream_mix.cu && /tmp/stream_mix
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ uint64_t policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t policy_evict_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void bulk_load_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_store_hint(void* dst, uint32_t src, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// mode 0: per 16-byte unit of input, R = out/in units of output
__global__ void __launch_bounds__(256) ldst_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long n_in16, int ratio_num, int ratio_den) {
    // units of work: blocks of `ratio_den` input vectors and `ratio_num` output vectors
    const long long nblk = n_in16 / ratio_den;
    for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < nblk; b += (long long)gridDim.x * blockDim.x) {
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int i = 0; i < ratio_den; ++i) { const uint4 v = in[(long long)i * nblk + b]; acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w; }
        for (int i = 0; i < ratio_num; ++i) out[(long long)i * nblk + b] = acc;
    }
}

template <int MODE>
__global__ void bulk_kernel(const char* __restrict__ in, char* __restrict__ out, long long ntiles, int in_bytes, int out_bytes, int nbuf_out) {
    extern __shared__ __align__(128) unsigned char smem[];
    // layout: in0 | in1 | out images (nbuf_out) | barriers
    const int inb = (in_bytes + 127) & ~127, outb = (out_bytes + 127) & ~127;
    unsigned char* in_img[2] = {smem, smem + inb};
    unsigned char* out_img = smem + 2 * inb;
    const uint32_t bar0 = smem_u32(smem + 2 * inb + nbuf_out * outb);
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid * 16; i < nbuf_out * outb; i += blockDim.x * 16) *reinterpret_cast<uint4*>(out_img + i) = make_uint4(i, tid, 3, 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    long long tile = blockIdx.x;
    if (tile < ntiles && tid == 0) { mbar_expect_tx(bar0, in_bytes); bulk_load(smem_u32(in_img[0]), in + tile * in_bytes, in_bytes, bar0); }
    uint32_t phase[2] = {0, 0};
    unsigned acc = 0;
    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int s = it & 1;
        const long long nxt = tile + gridDim.x;
        if (tid == 0 && nxt < ntiles) { mbar_expect_tx(bar0 + 8 * (s ^ 1), in_bytes); bulk_load(smem_u32(in_img[s ^ 1]), in + nxt * in_bytes, in_bytes, bar0 + 8 * (s ^ 1)); }
        while (!mbar_try_wait(bar0 + 8 * s, phase[s])) {}
        phase[s] ^= 1;
        acc += in_img[s][(tid * 4) % in_bytes];
        unsigned char* oimg = out_img + (it % nbuf_out) * outb;
        if (MODE == 2) {
            // wait until the store that last used this image has read it, then rewrite it
            if (tid == 0) {
                if (nbuf_out == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            __syncthreads();
            for (int i = tid * 16; i < out_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(oimg + i) = make_uint4(acc, tid, it, 4);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) { bulk_store(out + tile * out_bytes, smem_u32(oimg), out_bytes); asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (acc == 0xFFFFFFFFu) out[0] = 1;
}

// the skeleton with the launch / cache options under test (all run-time, PDL always on).  opt bits: 2 = stores with L2 evict_first,
// 4 = loads with L2 evict_first, 8 = L2 prefetch of the first tile BEFORE griddepcontrol.wait, 32 = also L2-prefetch tile i+2 every iteration
__global__ void opt_kernel(const char* __restrict__ in, char* __restrict__ out, long long ntiles, int in_bytes, int out_bytes, int nimg, int opt) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int inb = (in_bytes + 127) & ~127, outb = (out_bytes + 127) & ~127;
    unsigned char* in_img[2] = {smem, smem + inb};
    unsigned char* out_img = smem + 2 * inb;
    const uint32_t bar0 = smem_u32(smem + 2 * inb + nimg * outb);
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    const uint64_t pf = policy_evict_first();
    long long tile = blockIdx.x;
    if (opt & 8) { if (tid == 0 && tile < ntiles) bulk_prefetch_l2(in + tile * in_bytes, in_bytes); }
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    auto load = [&](long long t, int buf) {
        mbar_expect_tx(bar0 + 8 * buf, in_bytes);
        if (opt & 4) bulk_load_hint(smem_u32(in_img[buf]), in + t * in_bytes, in_bytes, bar0 + 8 * buf, pf);
        else bulk_load(smem_u32(in_img[buf]), in + t * in_bytes, in_bytes, bar0 + 8 * buf);
    };
    if (tile < ntiles && tid == 0) load(tile, 0);
    uint32_t phase[2] = {0, 0};
    unsigned acc = 0;
    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int s = it & 1;
        const long long nxt = tile + gridDim.x;
        if (tid == 0 && nxt < ntiles) load(nxt, s ^ 1);
        if ((opt & 32) && tid == 0 && nxt + gridDim.x < ntiles) bulk_prefetch_l2(in + (nxt + gridDim.x) * in_bytes, in_bytes);
        while (!mbar_try_wait(bar0 + 8 * s, phase[s])) {}
        phase[s] ^= 1;
        acc += in_img[s][(tid * 4) % in_bytes];
        unsigned char* oimg = out_img + (it % nimg) * outb;
        if (tid == 0) { if (nimg == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
        __syncthreads();
        for (int i = tid * 16; i < out_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(oimg + i) = make_uint4(acc, tid, it, 4);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            if (opt & 2) bulk_store_hint(out + tile * out_bytes, smem_u32(oimg), out_bytes, pf);
            else bulk_store(out + tile * out_bytes, smem_u32(oimg), out_bytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (acc == 0xFFFFFFFFu) out[0] = 1;
}
void launch_opt(const char* in, char* out, long long ntiles, int ib, int ob, int nimg, int opt, int grid, int nthr, size_t smem) {
    static bool once = false;
    if (!once) { CK(cudaFuncSetAttribute(opt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); once = true; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(nthr); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, opt_kernel, in, out, ntiles, ib, ob, nimg, opt));
}

struct Case { const char* name; long long N; int in_b, out_b; };

int main(int argc, char** argv) {
    const Case cases[] = {{"C2 cartpole fp64 (40 in / 160 out), N=2^20", 1 << 20, 40, 160},
                          {"C3 quadrotor fp32 (68 in / 884 out), N=262144", 262144, 68, 884},
                          {"C4 satellite fp64 (144 in / 1728 out), N=2^20", 1 << 20, 144, 1728},
                          {"1:1 copy, 100 B + 100 B per unit, N=2^20", 1 << 20, 96, 96},
                          {"1:1 copy, 1 KiB + 1 KiB per unit, N=2^20", 1 << 20, 1024, 1024},
                          {"write-only-ish (16 in / 1024 out), N=2^20", 1 << 20, 16, 1024}};
    const int steps = 50, nsets = 4;
    const bool quick = argc > 1 && !strcmp(argv[1], "opt");
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(bulk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(bulk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (const Case& c : cases) {
        if (quick && c.in_b != 40 && c.in_b != 68) continue;
        const size_t inB = size_t(c.N) * c.in_b, outB = size_t(c.N) * c.out_b;
        const int sets = (inB + outB) * nsets > (size_t(12) << 30) ? 2 : nsets;
        std::vector<char*> in(sets), out(sets);
        for (int i = 0; i < sets; ++i) { CK(cudaMalloc(&in[i], inB)); CK(cudaMalloc(&out[i], outB)); CK(cudaMemset(in[i], 1, inB)); CK(cudaMemset(out[i], 0, outB)); }
        printf("== %s: %.1f MB per launch\n", c.name, (inB + outB) / 1e6);
        auto time_it = [&](const char* label, auto launch) {
            for (int i = 0; i < 5; ++i) launch(i % sets);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            for (int i = 0; i < steps; ++i) launch(i % sets);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double us = ms * 1e3 / steps;
            printf("   %-58s %8.1f us  %7.0f GB/s\n", label, us, (inB + outB) / (us * 1e-6) / 1e9);
        };
        // mode 0: 16-byte units, ratio reduced
        if (!quick) {
            int num = c.out_b / 4, den = c.in_b / 4;
            auto g = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
            const int gg = g(num, den); num /= gg; den /= gg;
            const long long n_in16 = (long long)(inB / 16) / den * den;
            for (int ctas : {4, 8}) {
                char label[128]; snprintf(label, sizeof label, "ld/st 16 B per thread, %d x 256 threads per SM (%d:%d)", ctas, den, num);
                time_it(label, [&](int s) { ldst_kernel<<<148 * ctas, 256>>>((const uint4*)in[s], (uint4*)out[s], n_in16, num, den); });
            }
        }
        if (!quick) for (int tile : {32, 64, 128, 256}) {
            for (int nbuf : {1, 2}) {
                const int ib = tile * c.in_b, ob = tile * c.out_b;
                if (ib % 16 || ob % 16) continue;
                const size_t smem = 2 * ((ib + 127) & ~127) + nbuf * ((ob + 127) & ~127) + 16;
                if (smem > 200 * 1024) continue;
                int per_sm = int((220 * 1024) / (smem + 1024)); if (per_sm > 16) per_sm = 16; if (per_sm < 1) continue;
                const long long ntiles = c.N / tile;
                const int grid = int(ntiles < 148LL * per_sm ? ntiles : 148LL * per_sm);
                char label[128];
                snprintf(label, sizeof label, "bulk skeleton, tile %3d, %d out image(s), %2d CTAs/SM", tile, nbuf, per_sm);
                if (nbuf == 1) time_it(label, [&](int s) { bulk_kernel<1><<<grid, 64, smem>>>(in[s], out[s], ntiles, ib, ob, nbuf); });
                snprintf(label, sizeof label, "bulk + smem rewrite, tile %3d, %d out image(s), %2d CTAs/SM", tile, nbuf, per_sm);
                time_it(label, [&](int s) { bulk_kernel<2><<<grid, 64, smem>>>(in[s], out[s], ntiles, ib, ob, nbuf); });
            }
        }
        if (quick || c.in_b <= 144) for (int tile : {64, 128, 256}) for (int nthr : {64, 128}) for (int nimg : {1, 2}) for (int cap : {4, 8, 16}) {
            const int ib = tile * c.in_b, ob = tile * c.out_b;
            const size_t smem = 2 * ((ib + 127) & ~127) + nimg * ((ob + 127) & ~127) + 16;
            if (smem > 200 * 1024) continue;
            int per_sm = int((225 * 1024) / (smem + 1024)); if (per_sm > 2048 / nthr) per_sm = 2048 / nthr;
            if (per_sm < cap && cap != 16) continue;          // this cap is not reachable
            if (per_sm > cap) per_sm = cap;
            if (per_sm < 1) continue;
            const long long ntiles = c.N / tile;
            const int grid = int(ntiles < 148LL * per_sm ? ntiles : 148LL * per_sm);
            for (int opt : {0, 2, 8, 10, 42}) {
                double best = 1e30, med[5];
                for (int rep = 0; rep < 5; ++rep) {
                    for (int i = 0; i < 3; ++i) launch_opt(in[i % sets], out[i % sets], ntiles, ib, ob, nimg, opt, grid, nthr, smem);
                    CK(cudaDeviceSynchronize());
                    CK(cudaEventRecord(e0));
                    for (int i = 0; i < steps; ++i) launch_opt(in[i % sets], out[i % sets], ntiles, ib, ob, nimg, opt, grid, nthr, smem);
                    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    med[rep] = ms * 1e3 / steps; if (med[rep] < best) best = med[rep];
                }
                std::sort(med, med + 5);
                printf("   opt tile %3d x %3d thr, %d img, %2d CTAs/SM, %-32s min %6.1f  median %6.1f us  %5.0f GB/s\n", tile, nthr, nimg, per_sm,
                       opt == 0 ? "PDL" : opt == 2 ? "PDL + st evict_first" : opt == 8 ? "PDL + prefetch" : opt == 10 ? "PDL + evict_first + prefetch" : "PDL + ef + prefetch + deep pf",
                       best, med[2], (inB + outB) / (med[2] * 1e-6) / 1e9);
            }
        }
        for (int i = 0; i < sets; ++i) { CK(cudaFree(in[i])); CK(cudaFree(out[i])); }
    }
    return 0;
}
