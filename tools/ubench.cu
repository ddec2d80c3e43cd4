// ubench.cu — memory-system micro-benchmarks that size the scan kernel's stages on B200 (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench tools/ubench.cu
// Measures (CUDA events, best of 5): random 4-byte gathers from an L2-sized filter, random 32-byte bucket
// gathers and 64-byte record gathers from a DRAM-sized table (one lane per item vs several lanes per item),
// random 64-bit atomicMin, with and without an L2 persisting window on the filter.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t v) {
    v *= 0xD6E8FEB86659FD93ull; v ^= v >> 32; v *= 0xD6E8FEB86659FD93ull; v ^= v >> 32;
    return v;
}

// each thread: `per` lookups, U in flight
template <int U>
__global__ void gather4(const uint32_t *__restrict__ buf, uint32_t words_mask, uint64_t per, uint64_t *out) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint64_t i = 0; i < per; i += U) {
        uint32_t v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = __ldg(buf + ((uint32_t)mix64(tid * per + i + u + 1) & words_mask));
#pragma unroll
        for (int u = 0; u < U; u++) acc += v[u];
    }
    if (acc == 0x1234567) out[0] = acc;
}

// clustered variant: groups of G consecutive lanes hit the same 32-byte sector (different words)
template <int U, int G>
__global__ void gather4_clustered(const uint32_t *__restrict__ buf, uint32_t words_mask, uint64_t per, uint64_t *out) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint64_t i = 0; i < per; i += U) {
        uint32_t v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t sector = (uint32_t)mix64((tid / G) * per + i + u + 1) & words_mask & ~7u;
            v[u] = __ldg(buf + sector + (tid & 7));
        }
#pragma unroll
        for (int u = 0; u < U; u++) acc += v[u];
    }
    if (acc == 0x1234567) out[0] = acc;
}

struct __align__(32) u32x8 { uint32_t v[8]; };
__device__ __forceinline__ u32x8 ld256(const void *p) {
    u32x8 r;
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}
// one lane per item of BYTES bytes (32/64/128), 256-bit loads
template <int U, int BYTES>
__global__ void gather_ld256(const uint8_t *__restrict__ buf, uint32_t item_mask, uint64_t per, uint64_t *out) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    constexpr int V = BYTES / 32;
    for (uint64_t i = 0; i < per; i += U) {
        u32x8 a[U][V];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t r = (uint32_t)mix64(tid * per + i + u + 1) & item_mask;
#pragma unroll
            for (int v = 0; v < V; v++) a[u][v] = ld256(buf + (size_t)r * BYTES + 32 * v);
        }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int v = 0; v < V; v++) acc += a[u][v].v[0] + a[u][v].v[7];
    }
    if (acc == 0x1234567) out[0] = acc;
}
// verify-stage model: load a 64-byte record (2 x 256-bit), then atomicMin on its second 8 bytes for a fraction of items
template <int U>
__global__ void record_then_atomic(uint8_t *buf, uint32_t item_mask, uint64_t per, uint32_t atomic_per_256, uint64_t *out) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint64_t i = 0; i < per; i += U) {
        u32x8 a[U][2];
        uint64_t h[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            h[u] = mix64(tid * per + i + u + 1);
            const uint32_t r = (uint32_t)h[u] & item_mask;
            a[u][0] = ld256(buf + (size_t)r * 64); a[u][1] = ld256(buf + (size_t)r * 64 + 32);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            acc += a[u][0].v[0] + a[u][1].v[7];
            if (((h[u] >> 40) & 255) < atomic_per_256)
                atomicMin((long long *)(buf + (size_t)((uint32_t)h[u] & item_mask) * 64 + 8), (long long)(h[u] >> 9));
        }
    }
    if (acc == 0x1234567) out[0] = acc;
}

// one lane per 32-byte bucket: two 16-byte loads
template <int U>
__global__ void gather32(const uint4 *__restrict__ buf, uint32_t bucket_mask, uint64_t per, uint64_t *out) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint64_t i = 0; i < per; i += U) {
        uint4 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t bk = (uint32_t)mix64(tid * per + i + u + 1) & bucket_mask;
            a[u] = __ldg(buf + (size_t)bk * 2); b[u] = __ldg(buf + (size_t)bk * 2 + 1);
        }
#pragma unroll
        for (int u = 0; u < U; u++) acc += a[u].x + b[u].w;
    }
    if (acc == 0x1234567) out[0] = acc;
}

// two lanes per 32-byte bucket: one 16-byte load each
template <int U>
__global__ void gather32_pair(const uint4 *__restrict__ buf, uint32_t bucket_mask, uint64_t per, uint64_t *out) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint64_t i = 0; i < per; i += U) {
        uint4 a[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t bk = (uint32_t)mix64((tid >> 1) * per + i + u + 1) & bucket_mask;
            a[u] = __ldg(buf + (size_t)bk * 2 + (tid & 1));
        }
#pragma unroll
        for (int u = 0; u < U; u++) acc += a[u].x;
    }
    if (acc == 0x1234567) out[0] = acc;
}

// 64-byte records: LANES lanes per record, each 64/LANES bytes as uint4 loads
template <int U, int LANES>
__global__ void gather64(const uint4 *__restrict__ buf, uint32_t rec_mask, uint64_t per, uint64_t *out) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    constexpr int V = 4 / LANES;
    for (uint64_t i = 0; i < per; i += U) {
        uint4 a[U][V];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t r = (uint32_t)mix64((tid / LANES) * per + i + u + 1) & rec_mask;
#pragma unroll
            for (int v = 0; v < V; v++) a[u][v] = __ldg(buf + (size_t)r * 4 + (tid % LANES) * V + v);
        }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int v = 0; v < V; v++) acc += a[u][v].x;
    }
    if (acc == 0x1234567) out[0] = acc;
}

__global__ void atomic_min64(long long *buf, uint32_t mask, uint64_t per, int stride_words) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t i = 0; i < per; i++) {
        const uint64_t h = mix64(tid * per + i + 1);
        atomicMin(buf + (size_t)((uint32_t)h & mask) * stride_words, (long long)(h >> 8));
    }
}

__global__ void atomic_or32(uint32_t *buf, uint32_t mask, uint64_t per) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t i = 0; i < per; i++) {
        const uint64_t h = mix64(tid * per + i + 1);
        atomicOr(buf + ((uint32_t)h & mask), 1u << (h >> 59));
    }
}

__global__ void atomic_cas64(unsigned long long *buf, uint32_t mask, uint64_t per) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t i = 0; i < per; i++) {
        const uint64_t h = mix64(tid * per + i + 1);
        atomicCAS(buf + ((uint32_t)h & mask), 0xFFFFFFFFFFFFFFFFull, h);
    }
}

// pure integer work per "position": the cost of a window hash (no memory)
__global__ void hash_only(uint64_t per, uint64_t *out) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    uint32_t a = (uint32_t)tid, b = a * 3u, c = a * 7u;
    for (uint64_t i = 0; i < per; i++) {
        a = __funnelshift_r(a, b, 1); b = __funnelshift_r(b, c, 1); c += 0x9E3779B9u;
        uint64_t v = ((uint64_t)b << 32 | a) ^ ((uint64_t)(a & b) * 0x9E3779B97F4A7C15ull);
        v = mix64(v);
        acc += v >> 37;
    }
    if (acc == 0x1234567) out[0] = acc;
}

template <class F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm %d.%d SMs %d L2 %d MB persistingL2max %d MB accessPolicyMaxWindow %d MB smem/SM %zu KB\n", p.name, p.major, p.minor,
           p.multiProcessorCount, p.l2CacheSize >> 20, p.persistingL2CacheMaxSize >> 20, p.accessPolicyMaxWindowSize >> 20,
           p.sharedMemPerMultiprocessor >> 10);
    const int SM = p.multiProcessorCount;
    if (getenv("UB_L2_FETCH")) {
        size_t g = 0;
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(getenv("UB_L2_FETCH")));
        cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
        printf("L2 fetch granularity request %s -> %s, now %zu\n", getenv("UB_L2_FETCH"), cudaGetErrorString(e), g);
    } else {
        size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity default %zu\n", g);
    }
    if (getenv("UB_SHORT")) {
        uint64_t *out; CK(cudaMalloc(&out, 64));
        const size_t big = 1ull << 30; void *table; CK(cudaMalloc(&table, big)); CK(cudaMemset(table, 0xFF, big));
        const uint64_t N = 1ull << 27; const int blocks = SM * 8, threads = 256;
        const uint64_t per = N / ((uint64_t)blocks * threads); const double items = (double)per * blocks * threads;
        float a = time_ms([&] { gather_ld256<2, 32><<<blocks, threads>>>((uint8_t *)table, (uint32_t)(big / 32 - 1), per, out); });
        float b = time_ms([&] { gather64<2, 2><<<blocks, threads>>>((uint4 *)table, (uint32_t)(big / 64 - 1), per, out); });
        float c = time_ms([&] { gather4<4><<<blocks, threads>>>((uint32_t *)table, (uint32_t)(big / 4 - 1), per, out); });
        printf("1 GB table: 32 B items (LDG.256) %.1f G/s | 64 B items (lane pair) %.1f G/s | 4 B gathers %.1f G/s\n", items / a / 1e6, items / 2 / b / 1e6, items / c / 1e6);
        return 0;
    }
    uint64_t *out; CK(cudaMalloc(&out, 64));
    const size_t big = 1ull << 30;      // 1 GB "table"
    void *table; CK(cudaMalloc(&table, big)); CK(cudaMemset(table, 0xFF, big));
    const uint64_t N = 1ull << 27;      // items per test (134 M)
    const int blocks = SM * 8, threads = 256;
    const uint64_t per = N / ((uint64_t)blocks * threads);
    const double items = (double)per * blocks * threads;

    for (int fmb : {8, 16, 32, 64}) {
        void *filter; const size_t fb = (size_t)fmb << 20;
        CK(cudaMalloc(&filter, fb)); CK(cudaMemset(filter, 0, fb));
        const uint32_t wm = (uint32_t)(fb / 4 - 1);
        float t1 = time_ms([&] { gather4<1><<<blocks, threads>>>((uint32_t *)filter, wm, per, out); });
        float t4 = time_ms([&] { gather4<4><<<blocks, threads>>>((uint32_t *)filter, wm, per, out); });
        float t8 = time_ms([&] { gather4<8><<<blocks, threads>>>((uint32_t *)filter, wm, per, out); });
        float c4 = time_ms([&] { gather4_clustered<4, 4><<<blocks, threads>>>((uint32_t *)filter, wm, per, out); });
        float c8 = time_ms([&] { gather4_clustered<4, 8><<<blocks, threads>>>((uint32_t *)filter, wm, per, out); });
        printf("filter %2d MB  gather4 U1 %.1f  U4 %.1f  U8 %.1f G/s | clustered x4 %.1f  x8 %.1f G lookups/s\n", fmb,
               items / t1 / 1e6, items / t4 / 1e6, items / t8 / 1e6, items / c4 / 1e6, items / c8 / 1e6);
        CK(cudaFree(filter));
    }
    {
        const uint32_t bm = (uint32_t)(big / 32 - 1);
        float a1 = time_ms([&] { gather32<1><<<blocks, threads>>>((uint4 *)table, bm, per, out); });
        float a2 = time_ms([&] { gather32<2><<<blocks, threads>>>((uint4 *)table, bm, per, out); });
        float a4 = time_ms([&] { gather32<4><<<blocks, threads>>>((uint4 *)table, bm, per, out); });
        printf("bucket 32 B from 1 GB, 1 lane: U1 %.1f  U2 %.1f  U4 %.1f G buckets/s  (%.0f GB/s at U4)\n", items / a1 / 1e6,
               items / a2 / 1e6, items / a4 / 1e6, items * 32 / a4 / 1e6);
        float b1 = time_ms([&] { gather32_pair<1><<<blocks, threads>>>((uint4 *)table, bm, per, out); });
        float b4 = time_ms([&] { gather32_pair<4><<<blocks, threads>>>((uint4 *)table, bm, per, out); });
        printf("bucket 32 B from 1 GB, 2 lanes: U1 %.1f  U4 %.1f G buckets/s\n", items / 2 / b1 / 1e6, items / 2 / b4 / 1e6);
    }
    {
        const uint32_t rm = (uint32_t)(big / 64 - 1);
        float a1 = time_ms([&] { gather64<1, 1><<<blocks, threads>>>((uint4 *)table, rm, per, out); });
        float a2 = time_ms([&] { gather64<2, 1><<<blocks, threads>>>((uint4 *)table, rm, per, out); });
        float b1 = time_ms([&] { gather64<1, 2><<<blocks, threads>>>((uint4 *)table, rm, per, out); });
        float b2 = time_ms([&] { gather64<2, 2><<<blocks, threads>>>((uint4 *)table, rm, per, out); });
        float c1 = time_ms([&] { gather64<1, 4><<<blocks, threads>>>((uint4 *)table, rm, per, out); });
        float c4 = time_ms([&] { gather64<4, 4><<<blocks, threads>>>((uint4 *)table, rm, per, out); });
        printf("record 64 B from 1 GB: 1 lane U1 %.1f U2 %.1f | 2 lanes U1 %.1f U2 %.1f | 4 lanes U1 %.1f U4 %.1f G records/s (%.0f GB/s best 4-lane)\n",
               items / a1 / 1e6, items / a2 / 1e6, items / 2 / b1 / 1e6, items / 2 / b2 / 1e6, items / 4 / c1 / 1e6, items / 4 / c4 / 1e6,
               items / 4 * 64 / (c4 < c1 ? c4 : c1) / 1e6);
    }
    {
        float a = time_ms([&] { gather_ld256<2, 32><<<blocks, threads>>>((uint8_t *)table, (uint32_t)(big / 32 - 1), per, out); });
        float b = time_ms([&] { gather_ld256<2, 64><<<blocks, threads>>>((uint8_t *)table, (uint32_t)(big / 64 - 1), per, out); });
        float c = time_ms([&] { gather_ld256<2, 128><<<blocks, threads>>>((uint8_t *)table, (uint32_t)(big / 128 - 1), per, out); });
        float b1 = time_ms([&] { gather_ld256<1, 64><<<blocks, threads>>>((uint8_t *)table, (uint32_t)(big / 64 - 1), per, out); });
        float b4 = time_ms([&] { gather_ld256<4, 64><<<blocks, threads>>>((uint8_t *)table, (uint32_t)(big / 64 - 1), per, out); });
        printf("LDG.256 one lane per item from 1 GB: 32 B %.1f | 64 B U1 %.1f U2 %.1f U4 %.1f | 128 B %.1f G items/s\n", items / a / 1e6, items / b1 / 1e6,
               items / b / 1e6, items / b4 / 1e6, items / c / 1e6);
        for (uint32_t frac : {0u, 64u, 128u, 256u}) {
            float t = time_ms([&] { record_then_atomic<2><<<blocks, threads>>>((uint8_t *)table, (uint32_t)(big / 64 - 1), per, frac, out); });
            printf("record 64 B + atomicMin on %u/256 of them: %.1f G records/s\n", frac, items / t / 1e6);
        }
        // smaller footprints: how much of a 256 MB / 128 MB table does the L2 absorb?
        for (size_t mb : {512, 256, 128, 64}) {
            float t = time_ms([&] { gather_ld256<2, 64><<<blocks, threads>>>((uint8_t *)table, (uint32_t)((mb << 20) / 64 - 1), per, out); });
            printf("64 B items from a %zu MB table: %.1f G items/s\n", mb, items / t / 1e6);
        }
    }
    {
        const uint64_t per_a = per / 4;
        const double it = (double)per_a * blocks * threads;
        float m1 = time_ms([&] { atomic_min64<<<blocks, threads>>>((long long *)table, (uint32_t)(big / 64 - 1), per_a, 8); });
        float m2 = time_ms([&] { atomic_min64<<<blocks, threads>>>((long long *)table, (uint32_t)((64u << 20) / 8 - 1), per_a, 1); });
        float o1 = time_ms([&] { atomic_or32<<<blocks, threads>>>((uint32_t *)table, (uint32_t)((32u << 20) / 4 - 1), per_a); });
        float cs = time_ms([&] { atomic_cas64<<<blocks, threads>>>((unsigned long long *)table, (uint32_t)(big / 8 - 1), per_a); });
        printf("atomicMin64 random over 1 GB (64 B stride) %.1f G/s | over 64 MB %.1f G/s | atomicOr32 over 32 MB %.1f G/s | atomicCAS64 over 1 GB %.1f G/s\n",
               it / m1 / 1e6, it / m2 / 1e6, it / o1 / 1e6, it / cs / 1e6);
    }
    {
        float h = time_ms([&] { hash_only<<<blocks, threads>>>(per * 4, out); });
        printf("hash_only %.1f G hashes/s\n", items * 4 / h / 1e6);
    }
    // L2 persistence: filter 32 MB pinned while a DRAM-sized stream of bucket gathers runs concurrently
    {
        const size_t fb = 32u << 20;
        void *filter; CK(cudaMalloc(&filter, fb)); CK(cudaMemset(filter, 0, fb));
        const uint32_t wm = (uint32_t)(fb / 4 - 1), bm = (uint32_t)(big / 32 - 1);
        cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
        auto both = [&](const char *label) {
            cudaEvent_t a, b, c; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); CK(cudaEventCreate(&c));
            float bestf = 1e30f, bestg = 1e30f;
            for (int r = 0; r < 4; r++) {
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(a, s1)); CK(cudaEventRecord(c, s2));
                gather4<4><<<SM * 4, threads, 0, s1>>>((uint32_t *)filter, wm, per * 2, out);
                gather32<2><<<SM * 4, threads, 0, s2>>>((uint4 *)table, bm, per / 2, out);
                CK(cudaEventRecord(b, s1));
                cudaEvent_t d; CK(cudaEventCreate(&d)); CK(cudaEventRecord(d, s2));
                CK(cudaDeviceSynchronize());
                float f, g; CK(cudaEventElapsedTime(&f, a, b)); CK(cudaEventElapsedTime(&g, c, d));
                if (f < bestf) bestf = f;
                if (g < bestg) bestg = g;
            }
            printf("%s: concurrent filter gathers %.1f G/s, bucket gathers %.1f G/s\n", label, items / bestf / 1e6, items / 4 / bestg / 1e6);
        };
        both("no persistence");
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)p.persistingL2CacheMaxSize));
        cudaStreamAttrValue attr = {};
        attr.accessPolicyWindow.base_ptr = filter;
        attr.accessPolicyWindow.num_bytes = fb;
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CK(cudaStreamSetAttribute(s1, cudaStreamAttributeAccessPolicyWindow, &attr));
        both("filter persisting in L2");
    }
    return 0;
}
