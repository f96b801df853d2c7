// Microbenchmark: order-0 histogram of 256 KiB streams, one CTA per stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hist_variants hist_variants.cu
// V0 = the run-merging per-warp-bin kernel of kernels.cu (round 2), V1 = one conflict-free column per lane
// (table [256][32] shared by the CTA, one shared-memory atomic per byte), V3 = run boundaries by SWAR, one atomic per run.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int T = 256;
constexpr uint32_t S = 262144;

__device__ __forceinline__ uint4 ldg_u128(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// ---------------- V0
__device__ __forceinline__ void hist16_v0(uint4 q, uint32_t F_s) {
    uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t prev = w[0] & 0xff, cnt = 0;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            uint32_t c = (w[a] >> (8 * b)) & 0xff;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, %1;\n\t@p red.shared.add.u32 [%2], %3;\n\t}"
                         ::"r"(c), "r"(prev), "r"(F_s + prev * 4), "r"(cnt) : "memory");
            cnt = (c == prev) ? cnt + 1 : 1;
            prev = c;
        }
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(F_s + prev * 4), "r"(cnt) : "memory");
}
__global__ void __launch_bounds__(T) k_v0(const uint8_t *in, uint32_t *out) {
    __shared__ uint32_t Fw[8][256];
    const int tid = threadIdx.x, wid = tid >> 5;
    for (int j = tid; j < 8 * 256; j += T) (&Fw[0][0])[j] = 0;
    __syncthreads();
    const uint4 *v = (const uint4 *)(in + (size_t)blockIdx.x * S);
    const uint32_t nv = S / 16;
    const uint32_t F_s = (uint32_t)__cvta_generic_to_shared(Fw[wid]);
    for (uint32_t i = tid; i + 3 * T < nv; i += 4 * T) {
        uint4 q0 = ldg_u128(v + i), q1 = ldg_u128(v + i + T), q2 = ldg_u128(v + i + 2 * T), q3 = ldg_u128(v + i + 3 * T);
        hist16_v0(q0, F_s); hist16_v0(q1, F_s); hist16_v0(q2, F_s); hist16_v0(q3, F_s);
    }
    __syncthreads();
    uint32_t f = 0;
    for (int w = 0; w < 8; w++) f += Fw[w][tid];
    out[(size_t)blockIdx.x * 256 + tid] = f;
}

// ---------------- V1: table [256 symbols][32 lanes], every lane owns a bank
template <int U>
__global__ void __launch_bounds__(T) k_v1(const uint8_t *in, uint32_t *out) {
    extern __shared__ __align__(16) uint32_t tab[];         // 8192 words
    const int tid = threadIdx.x, lane = tid & 31;
    for (int j = tid; j < 2048; j += T) ((uint4 *)tab)[j] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const uint4 *v = (const uint4 *)(in + (size_t)blockIdx.x * S);
    const uint32_t nv = S / 16;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab) + 4 * lane;
    auto add = [&](uint32_t w) {
        uint32_t a0 = ((w << 7) & 0x7f80) + base, a1 = ((w >> 1) & 0x7f80) + base, a2 = ((w >> 9) & 0x7f80) + base,
                 a3 = ((w >> 17) & 0x7f80) + base;
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a0) : "memory");
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a1) : "memory");
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a2) : "memory");
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a3) : "memory");
    };
    for (uint32_t i = tid; i + (U - 1) * T < nv; i += U * T) {
        uint4 q[U];
#pragma unroll
        for (int u = 0; u < U; u++) q[u] = ldg_u128(v + i + u * T);
#pragma unroll
        for (int u = 0; u < U; u++) { add(q[u].x); add(q[u].y); add(q[u].z); add(q[u].w); }
    }
    __syncthreads();
    uint32_t f = 0;
#pragma unroll 8
    for (int k = 0; k < 32; k++) f += tab[tid * 32 + ((k + tid) & 31)];
    out[(size_t)blockIdx.x * 256 + tid] = f;
}

// ---------------- V2: as V1 without atomics: the table is private to a warp (u16 pairs), 4 warps per CTA
__global__ void __launch_bounds__(128) k_v2(const uint8_t *in, uint32_t *out) {
    extern __shared__ __align__(16) uint32_t tab[];         // 4 warps x 128 rows x 32 lanes
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int j = tid; j < 4096; j += 128) ((uint4 *)tab)[j] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const uint4 *v = (const uint4 *)(in + (size_t)blockIdx.x * S);
    const uint32_t nv = S / 16;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab) + wid * 16384 + 4 * lane;
    auto rmw = [&](uint32_t sym, uint32_t cnt) {           // row sym >> 1, half sym & 1
        uint32_t a = ((sym << 6) & 0x3f80) + base, inc = cnt << ((sym & 1) * 16), x;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(a) : "memory");
        x += inc;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory");
    };
    auto h16 = [&](uint4 q) {
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t prev = w[0] & 0xff, cnt = 0;
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                uint32_t c = (w[a] >> (8 * b)) & 0xff;
                if (c != prev) { rmw(prev, cnt); cnt = 0; }
                cnt++;
                prev = c;
            }
        rmw(prev, cnt);
    };
    for (uint32_t i = tid; i + 7 * 128 < nv; i += 8 * 128) {
        uint4 q[8];
#pragma unroll
        for (int u = 0; u < 8; u++) q[u] = ldg_u128(v + i + u * 128);
#pragma unroll
        for (int u = 0; u < 8; u++) h16(q[u]);
    }
    __syncthreads();
    uint32_t lo = 0, hi = 0;                                // thread t: symbols 2t, 2t+1
    for (int w = 0; w < 4; w++)
#pragma unroll 8
        for (int k = 0; k < 32; k++) { uint32_t x = tab[w * 4096 + tid * 32 + ((k + tid) & 31)]; lo += x & 0xffff; hi += x >> 16; }
    out[(size_t)blockIdx.x * 256 + 2 * tid] = lo;
    out[(size_t)blockIdx.x * 256 + 2 * tid + 1] = hi;
}

// ---------------- V3: run starts by SWAR, one atomic per run, per-warp bins
__device__ __forceinline__ uint32_t nz7(uint32_t d) { return (((d & 0x7f7f7f7fu) + 0x7f7f7f7fu) | d) & 0x80808080u; }
__global__ void __launch_bounds__(T) k_v3(const uint8_t *in, uint32_t *out) {
    __shared__ uint32_t Fw[8][256];
    const int tid = threadIdx.x, wid = tid >> 5;
    for (int j = tid; j < 8 * 256; j += T) (&Fw[0][0])[j] = 0;
    __syncthreads();
    const uint4 *v = (const uint4 *)(in + (size_t)blockIdx.x * S);
    const uint32_t nv = S / 16;
    const uint32_t F_s = (uint32_t)__cvta_generic_to_shared(Fw[wid]);
    auto h16 = [&](uint4 q) {
        // byte j starts a run iff it differs from byte j-1 (byte 0 always does)
        uint32_t m0 = nz7(q.x ^ (q.x << 8)), m1 = nz7(q.y ^ __funnelshift_l(q.x, q.y, 8)),
                 m2 = nz7(q.z ^ __funnelshift_l(q.y, q.z, 8)), m3 = nz7(q.w ^ __funnelshift_l(q.z, q.w, 8));
        // one bit per byte, stream order: bit j of m
        uint32_t m = (((m0 >> 7) * 0x00204081u) >> 21 & 0xf) | ((((m1 >> 7) * 0x00204081u) >> 21 & 0xf) << 4) |
                     ((((m2 >> 7) * 0x00204081u) >> 21 & 0xf) << 8) | ((((m3 >> 7) * 0x00204081u) >> 21 & 0xf) << 12);
        m |= 1u;
        m |= 1u << 16;                                      // sentinel: end of the 16 bytes
        uint32_t at = 0;
        m &= m - 1;                                         // drop the start at 0
        while (true) {
            uint32_t nx = __ffs(m) - 1;                     // next run start (or 16)
            uint32_t wsel = at >> 2;
            uint32_t word = wsel == 0 ? q.x : wsel == 1 ? q.y : wsel == 2 ? q.z : q.w;
            uint32_t sym = (word >> ((at & 3) * 8)) & 0xff;
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(F_s + sym * 4), "r"(nx - at) : "memory");
            if (nx == 16) break;
            at = nx;
            m &= m - 1;
        }
    };
    for (uint32_t i = tid; i + 3 * T < nv; i += 4 * T) {
        uint4 q0 = ldg_u128(v + i), q1 = ldg_u128(v + i + T), q2 = ldg_u128(v + i + 2 * T), q3 = ldg_u128(v + i + 3 * T);
        h16(q0); h16(q1); h16(q2); h16(q3);
    }
    __syncthreads();
    uint32_t f = 0;
    for (int w = 0; w < 8; w++) f += Fw[w][tid];
    out[(size_t)blockIdx.x * 256 + tid] = f;
}

int main(int argc, char **argv) {
    const int nstreams = argc > 1 ? atoi(argv[1]) : 3815;
    const size_t n = (size_t)nstreams * S;
    std::vector<uint8_t> h(n);
    uint64_t x = 88172645463325252ull;
    uint8_t cur = 30;
    for (size_t i = 0; i < n; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        if ((x & 7) == 0) cur = 2 + (uint8_t)((x >> 8) % 39);
        h[i] = cur;
    }
    std::vector<uint32_t> ref((size_t)nstreams * 256, 0);
    for (size_t i = 0; i < n; i++) ref[(i / S) * 256 + h[i]]++;
    uint8_t *d; uint32_t *o;
    CK(cudaMalloc(&d, n)); CK(cudaMalloc(&o, (size_t)nstreams * 1024));
    CK(cudaMemcpy(d, h.data(), n, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_v1<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    CK(cudaFuncSetAttribute(k_v1<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    CK(cudaFuncSetAttribute(k_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<uint32_t> got((size_t)nstreams * 256);
    auto run = [&](const char *name, auto launch) {
        CK(cudaMemset(o, 0xff, (size_t)nstreams * 1024));
        for (int w = 0; w < 3; w++) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 10; r++) launch();
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 10;
        CK(cudaMemcpy(got.data(), o, (size_t)nstreams * 1024, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < got.size(); i++) bad += got[i] != ref[i];
        printf("{\"variant\": \"%s\", \"ms\": %.4f, \"gbs\": %.1f, \"mismatches\": %zu}\n", name, ms, n / ms / 1e6, bad);
    };
    run("v0_runmerge_warpbins", [&] { k_v0<<<nstreams, T>>>(d, o); });
    run("v1_lanecolumns_u4", [&] { k_v1<4><<<nstreams, T, 32768>>>(d, o); });
    run("v1_lanecolumns_u8", [&] { k_v1<8><<<nstreams, T, 32768>>>(d, o); });
    run("v2_private_u16", [&] { k_v2<<<nstreams, 128, 65536>>>(d, o); });
    run("v3_swar_runs", [&] { k_v3<<<nstreams, T>>>(d, o); });
    return 0;
}
