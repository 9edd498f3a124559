// FP32 / FP64 pipe probe for sm_100a: what the filter cascade's arithmetic can reach per SM per clock.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_probe tools/fma_probe.cu
// Every kernel runs register-only math (no memory in the loop); reported: FMA lane-operations per clock per SM
// at the clock measured with clock64 / globaltimer inside the launch.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define ITERS 4096

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// ---- raw instruction streams: NCH independent chains per thread
template <int NCH>
__global__ void k_ffma(float* out, float a, float b) {
    float x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) x[i] = threadIdx.x * 1e-3f + i;
    float aa = a + threadIdx.x * 1e-9f, bb = b;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) x[i] = fmaf(aa, x[i], bb);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;
}

template <int NCH>
__global__ void k_ffma2(float* out, float a, float b) {
    float2 x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    float2 aa = make_float2(a + threadIdx.x * 1e-9f, a), bb = make_float2(b, b + 1e-3f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) x[i] = ffma2(aa, x[i], bb);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += x[i].x + x[i].y;
    if (s == 123.456f) out[0] = s;
}

template <int NCH>
__global__ void k_dfma(float* out, double a, double b) {
    double x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) x[i] = threadIdx.x * 1e-3 + i;
    double aa = a + threadIdx.x * 1e-9, bb = b;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) x[i] = fma(aa, x[i], bb);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += x[i];
    if (s == 123.456) out[0] = (float)s;
}

// NF packed-FP32 chains and ND FP64 chains interleaved: do the two pipes run concurrently?
template <int NF, int ND>
__global__ void k_mix(float* out, float a, float b) {
    float2 x[NF];
    double y[ND];
#pragma unroll
    for (int i = 0; i < NF; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i);
#pragma unroll
    for (int i = 0; i < ND; ++i) y[i] = threadIdx.x * 1e-3 + i;
    float2 aa = make_float2(a + threadIdx.x * 1e-9f, a), bb = make_float2(b, b + 1e-3f);
    double da = a + threadIdx.x * 1e-9, db = b;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < (NF > ND ? NF : ND); ++i) {
            if (i < NF) x[i] = ffma2(aa, x[i], bb);
            if (i < ND) y[i] = fma(da, y[i], db);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += x[i].x + x[i].y;
#pragma unroll
    for (int i = 0; i < ND; ++i) s += y[i];
    if (s == 123.456) out[0] = (float)s;
}

// scalar FP32 chains + FP64 chains
template <int NF, int ND>
__global__ void k_mix1(float* out, float a, float b) {
    float x[NF];
    double y[ND];
#pragma unroll
    for (int i = 0; i < NF; ++i) x[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < ND; ++i) y[i] = threadIdx.x * 1e-3 + i;
    float aa = a + threadIdx.x * 1e-9f, bb = b;
    double da = a + threadIdx.x * 1e-9, db = b;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < (NF > ND ? NF : ND); ++i) {
            if (i < NF) x[i] = fmaf(aa, x[i], bb);
            if (i < ND) y[i] = fma(da, y[i], db);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < ND; ++i) s += y[i];
    if (s == 123.456) out[0] = (float)s;
}

// ---- the cascade's section step, registers only: NS sections, R rows per iteration in wavefront order
struct Sec2 { float2 nc, al, g, s1, s2; };
struct Sec1 { float nc, al, g, s1, s2; };
struct SecD { double nc, al, g, s1, s2; };

__device__ __forceinline__ float2 step2(float2 x, Sec2& r) {
    const float2 neg1 = make_float2(-1.0f, -1.0f);
    const float2 xs = ffma2(r.s2, neg1, x);
    const float2 e = ffma2(r.nc, r.s1, xs);
    const float2 bp = ffma2(r.al, e, r.s1);
    r.s1 = ffma2(r.al, e, bp);
    const float2 lp = ffma2(r.g, bp, r.s2);
    r.s2 = ffma2(r.g, bp, lp);
    return lp;
}
__device__ __forceinline__ float step1(float x, Sec1& r) {
    const float xs = x - r.s2;
    const float e = fmaf(r.nc, r.s1, xs);
    const float bp = fmaf(r.al, e, r.s1);
    r.s1 = fmaf(r.al, e, bp);
    const float lp = fmaf(r.g, bp, r.s2);
    r.s2 = fmaf(r.g, bp, lp);
    return lp;
}
__device__ __forceinline__ double stepd(double x, SecD& r) {
    const double xs = x - r.s2;
    const double e = fma(r.nc, r.s1, xs);
    const double bp = fma(r.al, e, r.s1);
    r.s1 = fma(r.al, e, bp);
    const double lp = fma(r.g, bp, r.s2);
    r.s2 = fma(r.g, bp, lp);
    return lp;
}

struct Sec5 { float2 nc, al, a2, g, g2, s1, s2; };
// FORM 1: five coefficients, the two updates that share (e, s1) / (bp, s2) written back to back (operand reuse)
// FORM 2: transposed direct form II with the low-pass numerator folded (5 instructions; reference for op count only)
template <int FORM>
__device__ __forceinline__ float2 step5(float2 x, Sec5& r) {
    if (FORM == 1) {
        const float2 xs = __fadd2_rn(x, make_float2(-r.s2.x, -r.s2.y));
        const float2 e = ffma2(r.nc, r.s1, xs);
        const float2 bp = ffma2(r.al, e, r.s1);
        r.s1 = ffma2(r.a2, e, r.s1);
        const float2 lp = ffma2(r.g, bp, r.s2);
        r.s2 = ffma2(r.g2, bp, r.s2);
        return lp;
    } else if (FORM == 3) {
        // three coefficients, state updates as 2 bp - s1 / 2 lp - s2 with an immediate 2.0 (two register reads each)
        const float2 two = make_float2(2.0f, 2.0f);
        const float2 xs = __fadd2_rn(x, make_float2(-r.s2.x, -r.s2.y));
        const float2 e = ffma2(r.nc, r.s1, xs);
        const float2 bp = ffma2(r.al, e, r.s1);
        r.s1 = ffma2(bp, two, make_float2(-r.s1.x, -r.s1.y));
        const float2 lp = ffma2(r.g, bp, r.s2);
        r.s2 = ffma2(lp, two, make_float2(-r.s2.x, -r.s2.y));
        return lp;
    } else {
        const float2 t = __fmul2_rn(r.g, x);
        const float2 y = __fadd2_rn(t, r.s1);
        const float2 u = ffma2(r.nc, y, r.s2);
        r.s1 = ffma2(make_float2(2.0f, 2.0f), t, u);
        r.s2 = ffma2(r.al, y, t);
        return y;
    }
}

// FORM 4 / 5: the two lanes of a packed register hold two TIME PIECES of ONE channel, so every coefficient is a scalar
// that FFMA2 takes in its broadcast (.F32) form -- half the operand bytes of a coefficient pair.  4: five coefficients,
// 5: three coefficients + 2.0.
struct Sec5s { float nc, al, a2, g, g2; float2 s1, s2; };
template <int FORM>
__device__ __forceinline__ float2 step5s(float2 x, Sec5s& r) {
    const float2 xs = __fadd2_rn(x, make_float2(-r.s2.x, -r.s2.y));
    const float2 e = ffma2(make_float2(r.nc, r.nc), r.s1, xs);
    const float2 bp = ffma2(make_float2(r.al, r.al), e, r.s1);
    if (FORM == 4) r.s1 = ffma2(make_float2(r.a2, r.a2), e, r.s1);
    else r.s1 = ffma2(bp, make_float2(2.0f, 2.0f), make_float2(-r.s1.x, -r.s1.y));
    const float2 lp = ffma2(make_float2(r.g, r.g), bp, r.s2);
    if (FORM == 4) r.s2 = ffma2(make_float2(r.g2, r.g2), bp, r.s2);
    else r.s2 = ffma2(lp, make_float2(2.0f, 2.0f), make_float2(-r.s2.x, -r.s2.y));
    return lp;
}

template <int NS, int R, int FORM>
__global__ void k_sec5s(float* out, float g) {
    Sec5s s[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const float gg = g * (1.0f + 0.01f * i + 1e-4f * threadIdx.x);
        s[i].g = gg; s[i].g2 = 2 * gg; s[i].nc = -(1.4f + gg); s[i].al = gg * 0.9f; s[i].a2 = gg * 1.8f;
        s[i].s1 = s[i].s2 = make_float2(0.f, 0.f);
    }
    float2 acc = make_float2(0.f, 0.f);
    for (int it = 0; it < ITERS / R; ++it) {
        float2 x[R];
#pragma unroll
        for (int k = 0; k < R; ++k) x[k] = make_float2(__int_as_float(0x3f800000 | ((it * R + k) * 2654435 & 0x7fffff)) - 1.5f, 0.25f);
#pragma unroll
        for (int d = 0; d < R + NS - 1; ++d)
#pragma unroll
            for (int i = 0; i < NS; ++i) { const int r = d - i; if (r >= 0 && r < R) x[r] = step5s<FORM>(x[r], s[i]); }
#pragma unroll
        for (int k = 0; k < R; ++k) { acc.x += x[k].x; acc.y += x[k].y; }
    }
    if (acc.x + acc.y == 123.456f) out[0] = acc.x;
}

// raw FFMA2 chains whose multiplier is a scalar taken in broadcast form
template <int NCH>
__global__ void k_ffma2s(float* out, float a, float b) {
    float2 x[NCH];
    float aa[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, i); aa[i] = a + (threadIdx.x + i) * 1e-9f; }
    float2 bb = make_float2(b, b + 1e-3f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) x[i] = ffma2(make_float2(aa[i], aa[i]), x[i], bb);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += x[i].x + x[i].y;
    if (s == 123.456f) out[0] = s;
}
// ... and the same with a distinct packed multiplier per chain (three distinct 64-bit operands per instruction)
template <int NCH>
__global__ void k_ffma2d(float* out, float a, float b) {
    float2 x[NCH], aa[NCH], bb[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, i); aa[i] = make_float2(a + (threadIdx.x + i) * 1e-9f, a - i * 1e-9f); bb[i] = make_float2(b + i * 1e-6f, b); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) x[i] = ffma2(aa[i], x[i], bb[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += x[i].x + x[i].y;
    if (s == 123.456f) out[0] = s;
}

template <int NS, int R, int FORM>
__global__ void k_sec5(float* out, float g) {
    Sec5 s[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const float gg = g * (1.0f + 0.01f * i + 1e-4f * threadIdx.x);
        s[i].g = make_float2(gg, gg * 1.1f);
        s[i].g2 = make_float2(2 * gg, gg * 2.2f);
        s[i].nc = make_float2(-(1.4f + gg), -(1.3f + gg));
        s[i].al = make_float2(gg * 0.9f, gg * 0.8f);
        s[i].a2 = make_float2(gg * 1.8f, gg * 1.6f);
        s[i].s1 = s[i].s2 = make_float2(0.f, 0.f);
    }
    float2 acc = make_float2(0.f, 0.f);
    for (int it = 0; it < ITERS / R; ++it) {
        float2 x[R];
#pragma unroll
        for (int k = 0; k < R; ++k) x[k] = make_float2(__int_as_float(0x3f800000 | ((it * R + k) * 2654435 & 0x7fffff)) - 1.5f, 0.25f);
#pragma unroll
        for (int d = 0; d < R + NS - 1; ++d)
#pragma unroll
            for (int i = 0; i < NS; ++i) { const int r = d - i; if (r >= 0 && r < R) x[r] = step5<FORM>(x[r], s[i]); }
#pragma unroll
        for (int k = 0; k < R; ++k) { acc.x += x[k].x; acc.y += x[k].y; }
    }
    if (acc.x + acc.y == 123.456f) out[0] = acc.x;
}

template <int NS, int R>
__global__ void k_sec2(float* out, float g) {
    Sec2 s[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const float gg = g * (1.0f + 0.01f * i + 1e-4f * threadIdx.x);
        s[i].g = make_float2(gg, gg * 1.1f);
        s[i].nc = make_float2(-(1.4f + gg), -(1.3f + gg));
        s[i].al = make_float2(gg * 0.9f, gg * 0.8f);
        s[i].s1 = s[i].s2 = make_float2(0.f, 0.f);
    }
    float2 acc = make_float2(0.f, 0.f);
    for (int it = 0; it < ITERS / R; ++it) {
        float2 x[R];
#pragma unroll
        for (int k = 0; k < R; ++k) x[k] = make_float2(__int_as_float(0x3f800000 | ((it * R + k) * 2654435 & 0x7fffff)) - 1.5f, 0.25f);
#pragma unroll
        for (int d = 0; d < R + NS - 1; ++d)
#pragma unroll
            for (int i = 0; i < NS; ++i) { const int r = d - i; if (r >= 0 && r < R) x[r] = step2(x[r], s[i]); }
#pragma unroll
        for (int k = 0; k < R; ++k) { acc.x += x[k].x; acc.y += x[k].y; }
    }
    if (acc.x + acc.y == 123.456f) out[0] = acc.x;
}

// two channels per thread, scalar FFMA
template <int NS, int R>
__global__ void k_sec1(float* out, float g) {
    Sec1 s[2][NS];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const float gg = g * (1.0f + 0.01f * i + 1e-4f * threadIdx.x + 0.1f * c);
        s[c][i].g = gg; s[c][i].nc = -(1.4f + gg); s[c][i].al = gg * 0.9f; s[c][i].s1 = s[c][i].s2 = 0.f;
    }
    float acc = 0.f;
    for (int it = 0; it < ITERS / R; ++it) {
        float x[2][R];
#pragma unroll
        for (int k = 0; k < R; ++k) { x[0][k] = __int_as_float(0x3f800000 | ((it * R + k) * 2654435 & 0x7fffff)) - 1.5f; x[1][k] = 0.25f; }
#pragma unroll
        for (int d = 0; d < R + NS - 1; ++d)
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const int r = d - i;
                if (r >= 0 && r < R) { x[0][r] = step1(x[0][r], s[0][i]); x[1][r] = step1(x[1][r], s[1][i]); }
            }
#pragma unroll
        for (int k = 0; k < R; ++k) acc += x[0][k] + x[1][k];
    }
    if (acc == 123.456f) out[0] = acc;
}

// two channels per thread: NF sections packed FP32 followed by ND sections FP64 (two scalar double chains)
template <int NF, int ND, int R>
__global__ void k_secmix(float* out, float g) {
    Sec2 s[NF];
    SecD t[2][ND];
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        const float gg = g * (1.0f + 0.01f * i + 1e-4f * threadIdx.x);
        s[i].g = make_float2(gg, gg * 1.1f);
        s[i].nc = make_float2(-(1.4f + gg), -(1.3f + gg));
        s[i].al = make_float2(gg * 0.9f, gg * 0.8f);
        s[i].s1 = s[i].s2 = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int i = 0; i < ND; ++i) {
        const double gg = g * (1.0 + 0.01 * i + 1e-4 * threadIdx.x + 0.1 * c);
        t[c][i].g = gg; t[c][i].nc = -(1.4 + gg); t[c][i].al = gg * 0.9; t[c][i].s1 = t[c][i].s2 = 0.0;
    }
    double acc = 0.0;
    for (int it = 0; it < ITERS / R; ++it) {
        float2 x[R];
        double y[2][R];
#pragma unroll
        for (int k = 0; k < R; ++k) x[k] = make_float2(__int_as_float(0x3f800000 | ((it * R + k) * 2654435 & 0x7fffff)) - 1.5f, 0.25f);
#pragma unroll
        for (int d = 0; d < R + NF + ND - 1; ++d)
#pragma unroll
            for (int i = 0; i < NF + ND; ++i) {
                const int r = d - i;
                if (r >= 0 && r < R) {
                    if (i < NF) {
                        x[r] = step2(x[r], s[i]);
                        if (i == NF - 1) { y[0][r] = (double)x[r].x; y[1][r] = (double)x[r].y; }
                    } else {
                        y[0][r] = stepd(y[0][r], t[0][i - NF]);
                        y[1][r] = stepd(y[1][r], t[1][i - NF]);
                    }
                }
            }
#pragma unroll
        for (int k = 0; k < R; ++k) acc += y[0][k] + y[1][k];
    }
    if (acc == 123.456) out[0] = (float)acc;
}

// FORM "delta": the low-pass section as an all-pole difference recurrence + its two Nyquist zeros as plain adds
// (3 FFMA2 + 2 FADD2 per two channel-samples instead of FADD2 + 5 FFMA2), two coefficient pairs per section.
struct SecDl { float2 a, be, D, Z, P; };
__device__ __forceinline__ float2 stepdl(float2 x, SecDl& r, const float2 m4) {
    const float2 w = ffma2(m4, r.Z, x);
    r.D = ffma2(r.a, r.D, w);
    const float2 zn = ffma2(r.be, r.D, r.Z);
    const float2 p = __fadd2_rn(zn, r.Z);
    r.Z = zn;
    const float2 o = __fadd2_rn(p, r.P);
    r.P = p;
    return o;
}
template <int NS, int R>
__global__ void k_secdl(float* out, float g) {
    SecDl s[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const float gg = g * (1.0f + 0.01f * i + 1e-4f * threadIdx.x);
        s[i].a = make_float2(1.0f - gg, 1.0f - 1.1f * gg);
        s[i].be = make_float2(gg * gg * 0.25f, gg * gg * 0.3f);
        s[i].D = s[i].Z = s[i].P = make_float2(0.f, 0.f);
    }
    float2 m4 = make_float2(-4.0f, -4.0f);
    asm volatile("" : "+f"(m4.x), "+f"(m4.y));
    float2 acc = make_float2(0.f, 0.f);
    for (int it = 0; it < ITERS / R; ++it) {
        float2 x[R];
#pragma unroll
        for (int k = 0; k < R; ++k) x[k] = make_float2(__int_as_float(0x3f800000 | ((it * R + k) * 2654435 & 0x7fffff)) - 1.5f, 0.25f);
#pragma unroll
        for (int d = 0; d < R + NS - 1; ++d)
#pragma unroll
            for (int i = 0; i < NS; ++i) { const int r = d - i; if (r >= 0 && r < R) x[r] = stepdl(x[r], s[i], m4); }
#pragma unroll
        for (int k = 0; k < R; ++k) { acc.x += x[k].x; acc.y += x[k].y; }
    }
    if (acc.x + acc.y == 123.456f) out[0] = acc.x;
}

// delta form with TIME-PAIR lanes: the two lanes of a packed register are two time pieces of ONE channel, so the two
// coefficients are scalars taken in broadcast form (fewer register-file words per instruction)
struct SecDls { float a, be; float2 D, Z, P; };
__device__ __forceinline__ float2 stepdls(float2 x, SecDls& r, const float m4) {
    const float2 w = ffma2(make_float2(m4, m4), r.Z, x);
    r.D = ffma2(make_float2(r.a, r.a), r.D, w);
    const float2 zn = ffma2(make_float2(r.be, r.be), r.D, r.Z);
    const float2 p = __fadd2_rn(zn, r.Z);
    r.Z = zn;
    const float2 o = __fadd2_rn(p, r.P);
    r.P = p;
    return o;
}
template <int NS, int R>
__global__ void k_secdls(float* out, float g) {
    SecDls s[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const float gg = g * (1.0f + 0.01f * i + 1e-4f * threadIdx.x);
        s[i].a = 1.0f - gg;
        s[i].be = gg * gg * 0.25f;
        s[i].D = s[i].Z = s[i].P = make_float2(0.f, 0.f);
    }
    float m4 = -4.0f;
    asm volatile("" : "+f"(m4));
    float2 acc = make_float2(0.f, 0.f);
    for (int it = 0; it < ITERS / R; ++it) {
        float2 x[R];
#pragma unroll
        for (int k = 0; k < R; ++k) x[k] = make_float2(__int_as_float(0x3f800000 | ((it * R + k) * 2654435 & 0x7fffff)) - 1.5f, 0.25f);
#pragma unroll
        for (int d = 0; d < R + NS - 1; ++d)
#pragma unroll
            for (int i = 0; i < NS; ++i) { const int r = d - i; if (r >= 0 && r < R) x[r] = stepdls(x[r], s[i], m4); }
#pragma unroll
        for (int k = 0; k < R; ++k) { acc.x += x[k].x; acc.y += x[k].y; }
    }
    if (acc.x + acc.y == 123.456f) out[0] = acc.x;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
    return ms / 5.0;
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, 4);
    const double clk = khz * 1e3;
    printf("SMs %d, clock %.0f MHz (attribute)\n", sms, khz / 1e3);
    const int threads = 256;
#define REPORT(name, per_thread_iter_fma, per_thread_iter_inst, ...)                                                   \
    for (int bps = 1; bps <= 4; bps *= 2) {                                                                                \
        const int blocks = sms * bps;                                                                                  \
        double ms = time_ms([&] { __VA_ARGS__; });                                                                     \
        double cyc = ms * 1e-3 * clk;                                                                                  \
        double fma = (double)(per_thread_iter_fma) * ITERS * threads * bps / cyc;                                      \
        double inst = (double)(per_thread_iter_inst) * ITERS * (threads / 32) * bps / cyc;                             \
        printf("%-34s warps/SM %2d  %8.3f ms  %7.1f FMA/clk/SM  %5.2f warp-inst/clk/SM\n", name, threads / 32 * bps, ms, fma, inst); \
    }
    REPORT("FFMA x16 chains", 16, 16, (k_ffma<16><<<blocks, threads>>>(out, 0.999f, 0.001f)))
    REPORT("FFMA2 x8 chains", 16, 8, (k_ffma2<8><<<blocks, threads>>>(out, 0.999f, 0.001f)))
    REPORT("FFMA2 x16 chains", 32, 16, (k_ffma2<16><<<blocks, threads>>>(out, 0.999f, 0.001f)))
    REPORT("DFMA x8 chains", 8, 8, (k_dfma<8><<<blocks, threads>>>(out, 0.999, 0.001)))
    REPORT("DFMA x16 chains", 16, 16, (k_dfma<16><<<blocks, threads>>>(out, 0.999, 0.001)))
    REPORT("FFMA2 x8 + DFMA x8", 24, 16, (k_mix<8, 8><<<blocks, threads>>>(out, 0.999f, 0.001f)))
    REPORT("FFMA2 x8 + DFMA x16", 32, 24, (k_mix<8, 16><<<blocks, threads>>>(out, 0.999f, 0.001f)))
    REPORT("FFMA x16 + DFMA x8", 24, 24, (k_mix1<16, 8><<<blocks, threads>>>(out, 0.999f, 0.001f)))
    REPORT("FFMA x16 + DFMA x16", 32, 32, (k_mix1<16, 16><<<blocks, threads>>>(out, 0.999f, 0.001f)))
    // cascades: FMA per thread per row = sections * 6 * 2 channels
    REPORT("cascade 8 sec packed f32, R=4", 96, 48, (k_sec2<8, 4><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec packed 5-coef, R=4", 96, 48, (k_sec5<8, 4, 1><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec packed 5-coef, R=2", 96, 48, (k_sec5<8, 2, 1><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec packed 5-coef, R=8", 96, 48, (k_sec5<8, 8, 1><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec packed 3-coef imm2, R=4", 96, 48, (k_sec5<8, 4, 3><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec packed 3-coef imm2, R=8", 96, 48, (k_sec5<8, 8, 3><<<blocks, threads>>>(out, 0.05f)))
    REPORT("FFMA2 x8, scalar-broadcast multiplier", 16, 8, (k_ffma2s<8><<<blocks, threads>>>(out, 0.999f, 0.001f)))
    REPORT("FFMA2 x8, 3 distinct packed operands", 16, 8, (k_ffma2d<8><<<blocks, threads>>>(out, 0.999f, 0.001f)))
    REPORT("cascade 8 sec, time-pair lanes, 5 scalar coef, R=8", 96, 48, (k_sec5s<8, 8, 4><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec, time-pair lanes, 5 scalar coef, R=4", 96, 48, (k_sec5s<8, 4, 4><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec, time-pair lanes, 3 scalar coef, R=8", 96, 48, (k_sec5s<8, 8, 5><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec packed DF2T, R=4", 80, 40, (k_sec5<8, 4, 2><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec packed DELTA (3 FFMA2 + 2 FADD2), R=4", 80, 40, (k_secdl<8, 4><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec packed DELTA (3 FFMA2 + 2 FADD2), R=8", 80, 40, (k_secdl<8, 8><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec DELTA, time-pair lanes (scalar coef), R=4", 80, 40, (k_secdls<8, 4><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec DELTA, time-pair lanes (scalar coef), R=8", 80, 40, (k_secdls<8, 8><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec packed f32, R=8", 96, 48, (k_sec2<8, 8><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 8 sec scalar f32, R=4", 96, 96, (k_sec1<8, 4><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 6 f32x2 + 2 f64, R=4", 96, 36 + 24, (k_secmix<6, 2, 4><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 5 f32x2 + 3 f64, R=4", 96, 30 + 36, (k_secmix<5, 3, 4><<<blocks, threads>>>(out, 0.05f)))
    REPORT("cascade 4 f32x2 + 4 f64, R=4", 96, 24 + 48, (k_secmix<4, 4, 4><<<blocks, threads>>>(out, 0.05f)))
    return 0;
}
