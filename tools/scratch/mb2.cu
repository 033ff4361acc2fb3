#include <cstdio>
constexpr int kIter = 2048, kChains = 8;
template <int OP>
__global__ void __launch_bounds__(1024) probe2(float seed, long long *cycles, float *sink)
{
    __shared__ float4 sm[1024 + 64];
    float lo[kChains], hi[kChains];
    for (int i = 0; i < kChains; ++i) { lo[i] = seed + i + threadIdx.x; hi[i] = lo[i] * 0.5f; }
    const float c = seed * 0.5f, c2 = seed * 0.25f;
    unsigned long long qk;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(qk) : "f"(c), "f"(c2));
    sm[threadIdx.x] = make_float4(c, c2, c, c2);
    __syncthreads();
    const float4 *sp = sm + threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < kIter; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (OP == 0) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(lo[i]) : "f"(c), "f"(hi[i]));
            if (OP == 1) {  // packed ACS: 1 FADD2 + 2 FMNMX
                unsigned long long p;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(lo[i]), "f"(hi[i]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(qk));
                float a, b;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p));
                asm volatile("max.f32 %0, %1, %2;" : "=f"(lo[i]) : "f"(a), "f"(c));
                asm volatile("max.f32 %0, %1, %2;" : "=f"(hi[i]) : "f"(b), "f"(c2));
            }
            if (OP == 2) {  // 1 FADD + 1 FMNMX
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(lo[i]) : "f"(c));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(hi[i]) : "f"(c2));
            }
            if (OP == 3) {  // LDS.128
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(sp + (i & 1))) : "memory");
                lo[i] += v.x + v.w; hi[i] += v.y + v.z;
            }
            if (OP == 4) {  // butterfly packed: 2 FADD2 + 2 FMNMX (4 adds, 2 max)
                unsigned long long p, q;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(lo[i]), "f"(lo[i]));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(hi[i]), "f"(hi[i]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(qk));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q) : "l"(qk));
                float a, b, d, e;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(d), "=f"(e) : "l"(q));
                asm volatile("max.f32 %0, %1, %2;" : "=f"(lo[i]) : "f"(a), "f"(e));
                asm volatile("max.f32 %0, %1, %2;" : "=f"(hi[i]) : "f"(b), "f"(d));
            }
            if (OP == 5) {  // sub.f32x2 alone for reference
                unsigned long long p;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(lo[i]), "f"(hi[i]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(qk));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo[i]), "=f"(hi[i]) : "l"(p));
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
    for (int i = 0; i < kChains; ++i) acc += lo[i] + hi[i];
    if (acc == 123.456f) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char *name, double instr_per_iter) {
    long long *d; float *s; cudaMalloc(&d, 8 * 148); cudaMalloc(&s, 4);
    probe2<OP><<<148, 1024>>>(1.f, d, s); probe2<OP><<<148, 1024>>>(1.f, d, s);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%-28s %.3f warp-instr/clk/SMSP (%.2f clk per chain-iter per warp)\n", name, 32.0 * kIter * kChains * instr_per_iter / avg / 4.0 / 32.0 * 1.0, avg / (kIter * kChains) / 8.0);
}
int main() {
    run<0>("FMNMX3", 1); run<1>("FADD2+2FMNMX", 3); run<2>("FADD+FMNMX", 2); run<3>("LDS.128(+4FADD)", 5);
    run<4>("2FADD2+2FMNMX", 4); run<5>("FADD2", 1);
    return 0;
}
