// Device helpers shared by the kernels of libsigb200.so: oscillator waveforms with the reference's exact
// float64 phase arithmetic, the Q0.32 phase-word fast paths, and the state-variable filter section.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "sigb200.h"
#include "sigb_internal.h"

namespace sigb_dev {

// ------------------------------------------------------------------------------------------
// oscillators
// ------------------------------------------------------------------------------------------

// sin(2*pi*r) for r in [-0.5, 0.5], float32.  Measured on B200 over 3M points (tests/test_gpu_parity.py
// ::test_sine_error_budget): variant 0 (MUFU.SIN) 3.39e-7 max-abs, variant 1 (folded to [-1/4,1/4]
// first) 3.39e-7 -- folding buys nothing --, variant 2 (folded FP32 polynomial) 1.73e-7.
template <int VARIANT>
__device__ __forceinline__ float sin2pi(float r) {
    if (VARIANT == 0) {
        return __sinf(r * 6.2831853071795864f);
    } else {
        float f = copysignf(0.5f, r) - r;          // exact (Sterbenz) when |r| >= 1/4
        r = fabsf(r) > 0.25f ? f : r;
        if (VARIANT == 1) {
            return __sinf(r * 6.2831853071795864f);
        } else {
            float u = r * r;                        // x * P4(x^2), max error 1.7e-7
            float p = 39.53670883178711f;
            p = fmaf(p, u, -76.5497817993164f);
            p = fmaf(p, u, 81.60100555419922f);
            p = fmaf(p, u, -41.34165573120117f);
            p = fmaf(p, u, 6.283185005187988f);
            return p * r;
        }
    }
}

#ifndef SIGB_SIN_VARIANT
#define SIGB_SIN_VARIANT 0
#endif

// numpy float remainder np.mod(a, b) for b in {1, 0.5}: fmod, then shift negatives up by b,
// and +0.0 for an exact zero (npy_divmod semantics; osc.py:49,55,61-62 rely on them).
template <int HALF>
__device__ __forceinline__ double np_mod(double a) {
    double m = HALF ? a - trunc(a * 2.0) * 0.5 : a - trunc(a);   // == fmod(a, b), exact
    if (m != 0.0) {
        if (m < 0.0) m = __dadd_rn(m, HALF ? 0.5 : 1.0);
    } else {
        m = 0.0;
    }
    return m;
}

__device__ __forceinline__ double np_sign(double v) {
    return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : (v == 0.0 ? 0.0 : v));   // NaN stays NaN
}

// osc.py:32 with the reference's exact float64 op order and no FMA contraction.
__device__ __forceinline__ double osc_cycles(double tn, double hertz, double phase) {
    return __dadd_rn(__dmul_rn(tn, hertz), phase);
}

__device__ __forceinline__ float osc_wave(int wave, double cyc) {
    switch (wave) {
        case SIGB_WAVE_SINE: {
            double r = cyc - rint(cyc);                                   // exact
            return sin2pi<SIGB_SIN_VARIANT>((float)r);
        }
        case SIGB_WAVE_SQUARE:                                           // osc.py:49
            return (float)np_sign(__dadd_rn(0.5, -np_mod<0>(cyc)));
        case SIGB_WAVE_SAWTOOTH:                                              // osc.py:55
            return (float)__dadd_rn(__dmul_rn(2.0, np_mod<0>(__dadd_rn(cyc, -0.5))), -1.0);
        default: {                                                        // osc.py:61-62
            double t = __dadd_rn(cyc, -0.25);
            double a = __dadd_rn(__dmul_rn(4.0, np_mod<1>(t)), -1.0);
            double s = np_sign(__dadd_rn(np_mod<0>(t), -0.5));
            return (float)__dmul_rn(a, s);
        }
    }
}

// sine from the top 32 bits of a Q0.64 phase accumulator, read as a signed fraction of a cycle
// (one I2F, one FMUL by 2*pi*2^-32, then __sinf = FMUL.RZ by 1/2pi + MUFU.SIN)
__device__ __forceinline__ float sine_q32(int hi) {
    if (SIGB_SIN_VARIANT == 0) return __sinf((float)hi * 1.4629180792671596e-9f);
    return sin2pi<SIGB_SIN_VARIANT>((float)hi * 2.3283064365386963e-10f);
}


// ------------------------------------------------------------------------------------------
// zero-delay-feedback state-variable section (one 2nd-order Butterworth factor)
//   hp = d (x - c s1 - s2);  bp = g hp + s1;  s1' = g hp + bp;  lp = g bp + s2;  s2' = g bp + lp
// first-order section: v = G (x - s1); lp = v + s1; s1' = lp + v; hp = x - lp
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float svf_lp2(float x, float g, float c, float d, float& s1, float& s2) {
    float t = fmaf(-c, s1, x);
    float hp = (t - s2) * d;
    float bp = fmaf(g, hp, s1);
    s1 = fmaf(g, hp, bp);
    float lp = fmaf(g, bp, s2);
    s2 = fmaf(g, bp, lp);
    return lp;
}

__device__ __forceinline__ float svf_any(int kind, float x, float g, float c, float d, float& s1, float& s2) {
    if (kind & SEC_FIRST_ORDER) {
        float v = (x - s1) * g;
        float lp = v + s1;
        s1 = lp + v;
        return (kind & SEC_HP) ? x - lp : lp;
    }
    float t = fmaf(-c, s1, x);
    float hp = (t - s2) * d;
    float bp = fmaf(g, hp, s1);
    s1 = fmaf(g, hp, bp);
    float lp = fmaf(g, bp, s2);
    s2 = fmaf(g, bp, lp);
    return (kind & SEC_HP) ? hp : lp;
}

constexpr float kTwoPiQ32 = 1.4629180792671596e-9f;   // 2*pi*2^-32

// waveform from the signed top word of the Q0.64 phase (fraction of a cycle in [-1/2, 1/2) after the
// two's-complement reading): osc.py:43, 49, 55, 61-62 away from their discontinuities
template <int WAVE>
__device__ __forceinline__ float wave_q32(int w) {
    if (WAVE == SIGB_WAVE_SINE) return __sinf((float)w * kTwoPiQ32);
    if (WAVE == SIGB_WAVE_SQUARE) return __int_as_float(0x3f800000 | (w & (int)0x80000000));   // frac < 1/2 -> +1: 1.0 with w's sign bit (one LOP3)
    if (WAVE == SIGB_WAVE_SAWTOOTH) return (float)w * 4.656612873077393e-10f;      // 2 frac (- 2 past 1/2)
    return fmaf(-fabsf((float)(w - 0x40000000)), 9.313225746154785e-10f, 1.0f);    // 1 - 4 |frac - 1/4|
}

// true when phase word w lies within `guard` (units of 2^-32 cycles) of a point where wave_q32 may
// disagree with the reference's float64 evaluation: the jumps of Square (frac 0 and 1/2) and Sawtooth
// (1/2), and Triangle's trough (3/4), where the reference yields -0.0 instead of -1 (sign(0) = 0, osc.py:62)
template <int WAVE>
__device__ __forceinline__ bool wave_near_edge(int w, int guard) {
    if (WAVE == SIGB_WAVE_SQUARE) return ((unsigned)(w + guard) & 0x7fffffffu) < 2u * (unsigned)guard;
    if (WAVE == SIGB_WAVE_SAWTOOTH) return ((unsigned)w ^ 0x80000000u) + (unsigned)guard < 2u * (unsigned)guard;
    if (WAVE == SIGB_WAVE_TRIANGLE) return (unsigned)w - 0xC0000000u + (unsigned)guard < 2u * (unsigned)guard;
    return false;
}

// K consecutive samples from phase word w advancing by dhi; returns true when any of them needs the
// float64 path (lies within `guard` of a point where wave_q32 may disagree with the reference, see wave_near_edge).
// The edge test rides on what the waveform computes anyway, one instruction per sample:
//   Sawtooth / Triangle: the jump / the trough sits where the (shifted) phase word wraps, |v| -> 2^31, and v = (float)w
//     is already there: |v| >= 2^31 - (guard + 128).  The float is within 64 of the integer up there, so the test can
//     only fire early (a wider band, never a missed edge).
//   Square: both jumps (frac 0 and 1/2) are the wrap of the DOUBLED word: (2 w + 2 guard) < 4 guard as unsigned,
//     exactly wave_near_edge's test, as one shift-add and one compare.
template <int WAVE, int K>
__device__ __forceinline__ bool gen_tile(int w, int dhi, int guard, float (&x)[K]) {
    // The edge tests accumulate as a running minimum / maximum (one instruction per sample, compared once per tile) instead
    // of a compare and a select per sample.
    bool near = false;
    if (WAVE == SIGB_WAVE_SINE) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            x[k] = wave_q32<SIGB_WAVE_SINE>(w);
            w += dhi;
        }
    } else if (WAVE == SIGB_WAVE_SQUARE) {
        const unsigned g2 = 2u * (unsigned)guard, g4 = 4u * (unsigned)guard;
        unsigned umin = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            x[k] = wave_q32<SIGB_WAVE_SQUARE>(w);
            umin = min(umin, ((unsigned)w << 1) + g2);
            w += dhi;
        }
        near = umin < g4;
    } else {
        const float thr = 2147483648.0f - (float)(guard + 128);
        if (WAVE == SIGB_WAVE_TRIANGLE) w -= 0x40000000;      // the trough (frac 3/4) becomes the wrap point
        float vmax = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float v = (float)w;
            x[k] = WAVE == SIGB_WAVE_SAWTOOTH ? v * 4.656612873077393e-10f                  // 2 frac (- 2 past 1/2)
                                              : fmaf(-fabsf(v), 9.313225746154785e-10f, 1.0f);   // 1 - 4 |frac - 1/4|
            vmax = fmaxf(vmax, fabsf(v));
            w += dhi;
        }
        near = vmax >= thr;
    }
    return near;
}

// guard band for wave_near_edge, in units of 2^-32 cycles (host and device agree on the formula):
// in-tile drift of the rounded increment (<= K/2), rounding of the top word (1), and the float64
// rounding of the reference's own phase (3 roundings of relative size 2^-53 on `cycles` cycles)

}  // namespace sigb_dev
