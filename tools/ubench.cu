// tools/ubench.cu -- instruction-throughput probes used to size the TNC kernel (not product code).
// Prints lane-ops per clock per SM for a few integer ops on the actual device.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
template <int OP>
__global__ void k(uint32_t *out, uint32_t seed)
{
    uint32_t a = threadIdx.x + seed, b = a * 2654435761u, c = b ^ 0x9e3779b9u, d = a + 77;
    uint32_t e = a * 3 + 1, f = b * 5 + 2, g = c * 7 + 3, h = d * 11 + 4;
    __shared__ uint32_t sm[32 * 64];
    if (OP >= 5) { for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) sm[i] = 0; __syncthreads(); }
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (OP == 0) { a = (a & b) ^ c; e = (e & f) ^ g; b = (b | d) ^ a; f = (f | h) ^ e; }                 // LOP3 x4
            if (OP == 1) { a += __popc(b ^ a); e += __popc(f ^ e); c += __popc(d ^ c); g += __popc(h ^ g); }     // POPC+LOP+IADD x4
            if (OP == 2) { a = __byte_perm(a, b, c); e = __byte_perm(e, f, g); c = __byte_perm(c, d, a); g = __byte_perm(g, h, e); }
            if (OP == 3) { a = a * b + c; e = e * f + g; c = c * d + a; g = g * h + e; }                          // IMAD x4
            if (OP == 4) { a = __funnelshift_l(a, b, 7) + 1; e = __funnelshift_l(e, f, 9) + 1; c = __funnelshift_l(c, d, 3) + 1; g = __funnelshift_l(g, h, 5) + 1; }
            if (OP == 5) { atomicAdd(&sm[(a & 63) * 32 + (threadIdx.x & 31)], 1); a = a * 1664525u + 1013904223u; e ^= a; c += e; g ^= c; }   // ATOMS, lane-private bank
            if (OP == 6) { atomicAdd(&sm[a & 2047], 1); a = a * 1664525u + 1013904223u; e ^= a; c += e; g ^= c; }                              // ATOMS, random
            if (OP == 7) { uint32_t idx = (a & 63) * 32 + (threadIdx.x & 31); sm[idx] = sm[idx] + 1; a = a * 1664525u + 1013904223u; e ^= a; c += e; g ^= c; }  // LDS+STS private
        }
    }
    if (OP >= 5) { __syncthreads(); a += sm[threadIdx.x]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ e ^ c ^ g;
}

template <int OP> void run(const char *name, double ops_per_inner, int sms, double clock_hz)
{
    uint32_t *out; cudaMalloc(&out, sizeof(uint32_t) * 1024 * 1024 * 4);
    int blocks = sms * 4, threads = 512;
    k<OP><<<blocks, threads>>>(out, 1); cudaDeviceSynchronize();
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
    cudaEventRecord(t0); k<OP><<<blocks, threads>>>(out, 2); cudaEventRecord(t1); cudaEventSynchronize(t1);
    float ms; cudaEventElapsedTime(&ms, t0, t1);
    double ops = (double)blocks * threads * ITER * 8 * ops_per_inner;
    printf("%-28s %8.3f ms  %7.1f lane-ops/clk/SM (at %.0f MHz)\n", name, ms, ops / (ms * 1e-3) / clock_hz / sms, clock_hz / 1e6);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double hz = khz * 1e3;
    printf("%s: %d SMs, clock attr %d kHz, L2 %d MB\n", p.name, p.multiProcessorCount, khz, p.l2CacheSize >> 20);
    run<0>("LOP3 (x4 per inner)", 4, p.multiProcessorCount, hz);
    run<1>("POPC+LOP3+IADD (x4)", 4, p.multiProcessorCount, hz);
    run<2>("PRMT (x4)", 4, p.multiProcessorCount, hz);
    run<3>("IMAD (x4)", 4, p.multiProcessorCount, hz);
    run<4>("SHF+IADD (x4)", 4, p.multiProcessorCount, hz);
    run<5>("ATOMS lane-private bank", 1, p.multiProcessorCount, hz);
    run<6>("ATOMS random 2048 bins", 1, p.multiProcessorCount, hz);
    run<7>("LDS+STS private RMW", 1, p.multiProcessorCount, hz);
    return 0;
}
