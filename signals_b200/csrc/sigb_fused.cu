// Fused render + mix-down kernels of libsigb200.so (sm_100a): the voice / partial blocks of these
// configurations are never materialised in HBM.
//
//   k_bank          oscillator bank under a GroupSum (BASELINE config C3: 65,536 sine partials -> 64
//                   channels).  Lane = time sample; the partials of a group are walked from a shared
//                   memory tile of {phase word, phase increment, amplitude}.  Bound: the MUFU pipe.
//   k_voices        voice bank under a PanSum (config C5: 1M osc -> filter -> gain -> pan instances ->
//                   stereo).  Thread = M voices with phase, filter state and weights in registers,
//                   walking time in 16-row tiles; per tile the CTA reduces its voices to one (16, 2)
//                   partial in a fixed order.
//   k_voices_finish fixed-order sum of the per-CTA partials into the (frames, 2) output.
//
// Reference semantics (file:line under /root/reference/src/signals/chain): osc.py:26-62 (phase from the
// absolute frame index, four waveforms), fx.py:49-52 (Gain), fx.py:85-151 (order <= 2 Butterworth
// low/high-pass).  The reference has no working N -> 1 mix-down (shape.py:32-57 is broken); GroupSum /
// PanSum are defined by oracle/np_oracle.py::group_sum / pan_sum.
#include <cuda_runtime.h>
#include <stdint.h>

#include "sigb200.h"
#include "sigb_internal.h"
#include "sigb_device.cuh"

namespace {

using namespace sigb_dev;


// ------------------------------------------------------------------------------------------
// k_bank
// ------------------------------------------------------------------------------------------
constexpr int BANK_THREADS = 256;
constexpr int BANK_WARPS = BANK_THREADS / 32;
constexpr int BANK_TN = 256;       // rows per tile: 8 per lane
constexpr int BANK_J = BANK_TN / 32;
constexpr int BANK_GB = 8;         // groups per work item (one 32-byte sector of an output row)
constexpr int BANK_PCHUNK = 1024;  // partials staged in shared memory at a time

// Two pipes instead of one.  A lane owns rows lane, lane + 32, ..., lane + 224 of the tile.  Evaluating every sample
// with MUFU.SIN makes the kernel MUFU-bound (16 results / clk / SM; round 1: 4.0e12 partial-samples/s at 0.86 of that
// pipe while the FMA pipe idled at 28 %).  Here only rows lane + 32 and lane + 160 (j = 1 and 5) are evaluated
// transcendentally -- sine AND cosine -- and their neighbours j = 0, 2, 3 / 4, 6, 7 come from the angle-addition
// rotation by the partial's 32-row phase advance D32, whose (cos, sin) the host tabulates in float64 -> float32
// once per plan:   sin(t +- D32) = sin t cos D32 +- cos t sin D32,  cos(t + D32) = cos t cos D32 - sin t sin D32.
// The two groups ride in the two lanes of packed f32x2 registers, so the eight samples cost 4 MUFU + ~10 FFMA2-class
// instructions: the MUFU load halves and the kernel becomes FMA-pipe / issue bound.  At most two rotation steps
// separate a sample from a transcendental evaluation (error: tests/test_gpu_parity.py::test_bank_*).
template <int UNROLL>           // partials per loop iteration: 1 -> 64 registers, 4 CTAs per SM; 2 -> 80 registers, 3 CTAs per SM
__global__ void __launch_bounds__(BANK_THREADS, UNROLL == 1 ? 4 : 3) k_bank(const BankDev a, int n_items, int gblocks) {
    __shared__ int4 par[BANK_PCHUNK];                         // {phase word at tile row 0, per-row increment, amplitude bits, 0}
    __shared__ float4 rot[BANK_PCHUNK];                       // {cos D32, sin D32, -sin D32, 0}
    __shared__ float red[BANK_WARPS][BANK_TN];
    __shared__ float outtile[BANK_TN][BANK_GB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = a.P / a.groups;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item / gblocks, gb = item - tile * gblocks;
        const int n0 = tile * BANK_TN;
        const int g_first = gb * BANK_GB;
        const int g_count = min(BANK_GB, a.groups - g_first);
        // the phase word is exact at the centre row of the tile; rows are +-128 away, so the rounded
        // increment drifts by at most 128 * 2^-33 cycles (9.4e-8 rad)
        const unsigned long long n_mid = (unsigned long long)(a.position + n0 + BANK_TN / 2);

        for (int gi = 0; gi < g_count; ++gi) {
            const int g = g_first + gi;
            float2 acc[BANK_J / 2];                           // acc[j] = rows (lane + 32 j, lane + 32 (j + 4))
#pragma unroll
            for (int j = 0; j < BANK_J / 2; ++j) acc[j] = make_float2(0.0f, 0.0f);

            for (int p0 = 0; p0 < per; p0 += BANK_PCHUNK) {
                const int cnt = min(BANK_PCHUNK, per - p0);
                __syncthreads();                              // previous chunk fully consumed
                for (int q = tid; q < cnt; q += BANK_THREADS) {
                    const int p = g * per + p0 + q;
                    const unsigned long long dth = a.dtheta[p];
                    const unsigned long long th = a.theta0[p] + n_mid * dth;      // exact mod 2^64
                    const int B = (int)((th + 0x80000000ull) >> 32);
                    const int D = (int)((dth + 0x80000000ull) >> 32);
                    const float amp = a.gain ? a.gain[p] : 1.0f;
                    par[q] = make_int4(B - (BANK_TN / 2) * D, D, __float_as_int(amp), 0);
                    const float2 cs = a.rot32[p];
                    rot[q] = make_float4(cs.x, cs.y, -cs.y, 0.0f);
                }
                __syncthreads();
#pragma unroll UNROLL
                for (int q = warp; q < cnt; q += BANK_WARPS) {
                    const int4 e = par[q];                    // broadcast reads
                    const float4 cs = rot[q];
                    const int step = e.y << 5;
                    const int wA = e.x + lane * e.y + step;   // row lane + 32   (j = 1)
                    const int wB = wA + 4 * step;             // row lane + 160  (j = 5)
                    const float amp = __int_as_float(e.z);
                    const float2 amp2 = make_float2(amp, amp);
                    const float2 CC = make_float2(cs.x, cs.x), SS = make_float2(cs.y, cs.y), NS = make_float2(cs.z, cs.z);
                    const float2 r = __fmul2_rn(make_float2((float)wA, (float)wB), make_float2(kTwoPiQ32, kTwoPiQ32));
                    const float2 S1 = __fmul2_rn(make_float2(__sinf(r.x), __sinf(r.y)), amp2);
                    const float2 C1 = __fmul2_rn(make_float2(__cosf(r.x), __cosf(r.y)), amp2);
                    // j = 0 / 4: one step back;  j = 2 / 6 and 3 / 7: one and two steps forward
                    acc[0] = __ffma2_rn(C1, NS, __ffma2_rn(S1, CC, acc[0]));
                    acc[1] = __fadd2_rn(acc[1], S1);
                    const float2 S2 = __ffma2_rn(C1, SS, __fmul2_rn(S1, CC));
                    const float2 C2 = __ffma2_rn(S1, NS, __fmul2_rn(C1, CC));
                    acc[2] = __fadd2_rn(acc[2], S2);
                    acc[3] = __ffma2_rn(C2, SS, __ffma2_rn(S2, CC, acc[3]));
                }
            }
            // cross-warp sum in a fixed order
#pragma unroll
            for (int j = 0; j < BANK_J / 2; ++j) {
                red[warp][lane + 32 * j] = acc[j].x;
                red[warp][lane + 32 * (j + 4)] = acc[j].y;
            }
            __syncthreads();
            {
                float s = 0.0f;
#pragma unroll
                for (int wv = 0; wv < BANK_WARPS; ++wv) s += red[wv][tid];
                outtile[tid][gi] = s;
            }
        }
        __syncthreads();
        const int rows = min(BANK_TN, a.frames - n0);
        for (int i = tid; i < rows * g_count; i += BANK_THREADS) {
            const int r = i / g_count, gi = i - r * g_count;
            a.out[(int64_t)(n0 + r) * a.ld_out + g_first + gi] = outtile[r][gi];
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_voices
// ------------------------------------------------------------------------------------------
constexpr int VT = SIGB_VOICE_THREADS;

// filter one tile of M voices: the M recurrences are independent, so they are written interleaved
// (row-major over k, then m) and the scheduler overlaps their dependency chains
template <int KIND, int M, int VK>
__device__ __forceinline__ void filt_tile(float (&x)[M][VK], const float (&g)[M], const float (&c)[M], const float (&d)[M],
                                          float (&s1)[M], float (&s2)[M], float (&s3)[M], int kmax) {
#ifndef SIGB_VOICES_SCALAR
    constexpr bool PACKED = M == 4 && !(KIND & SEC_FIRST_ORDER);
#else
    constexpr bool PACKED = false;
#endif
    if constexpr (PACKED) {
        // second-order sections of four voices as two packed f32x2 recurrences in DELTA FORM (sigb_reg.cu, k_cascade_delta):
        // the thread holds (a, F/4 | D, Z, P) for a low-pass voice and (-Q, -F | D, -4 Z) for a high-pass voice instead of
        // (g, c, d | s1, s2) -- k_voices converts at the ends of a piece -- 5 / 4 FFMA2-class instructions per two
        // voice-samples instead of 6 / 7; the high-pass output scale d is folded into the voice's (L, R) weights.
        //   low-pass:   w = x - 4 Z;  D' = a D + w;  Z' = Z + (F/4) D';  p = Z' + Z;  lp = p + P;  P' = p
        //   high-pass:  w = x + (-4 Z);  t = w - Q D;  D' = D + t;  (-4 Z)' = (-4 Z) - F D';  hp = d t
        float2 ca[2], cb[2], pD[2], pZ[2], pP[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m0 = 2 * h, m1 = 2 * h + 1;
            ca[h] = make_float2(g[m0], g[m1]);
            cb[h] = make_float2(c[m0], c[m1]);
            pD[h] = make_float2(s1[m0], s1[m1]);
            pZ[h] = make_float2(s2[m0], s2[m1]);
            pP[h] = make_float2(s3[m0], s3[m1]);
        }
        const float2 m4 = make_float2(-4.0f, -4.0f);
        auto row = [&](int k) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 xp = make_float2(x[2 * h][k], x[2 * h + 1][k]);
                float2 y;
                if (KIND & SEC_HP) {
                    const float2 w = __fadd2_rn(xp, pZ[h]);
                    y = __ffma2_rn(ca[h], pD[h], w);
                    pD[h] = __fadd2_rn(pD[h], y);
                    pZ[h] = __ffma2_rn(cb[h], pD[h], pZ[h]);
                } else {
                    const float2 w = __ffma2_rn(m4, pZ[h], xp);
                    pD[h] = __ffma2_rn(ca[h], pD[h], w);
                    const float2 zn = __ffma2_rn(cb[h], pD[h], pZ[h]);
                    const float2 p = __fadd2_rn(zn, pZ[h]);
                    pZ[h] = zn;
                    y = __fadd2_rn(p, pP[h]);
                    pP[h] = p;
                }
                x[2 * h][k] = y.x;
                x[2 * h + 1][k] = y.y;
            }
        };
        if (kmax == VK) {                                  // every tile but the ragged last one of a launch: no per-row test
#pragma unroll
            for (int k = 0; k < VK; ++k) row(k);
        } else {
#pragma unroll
            for (int k = 0; k < VK; ++k)
                if (k < kmax) row(k);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            s1[2 * h] = pD[h].x; s1[2 * h + 1] = pD[h].y;
            s2[2 * h] = pZ[h].x; s2[2 * h + 1] = pZ[h].y;
            s3[2 * h] = pP[h].x; s3[2 * h + 1] = pP[h].y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < VK; ++k) {
            if (k < kmax) {
#pragma unroll
                for (int m = 0; m < M; ++m) x[m][k] = svf_any(KIND, x[m][k], g[m], c[m], d[m], s1[m], s2[m]);
            }
        }
    }
}

template <int WAVE, int M, int VK>
__device__ __forceinline__ unsigned gen_tiles(const int (&w)[M], const int (&dhi)[M], int guard, float (&x)[M][VK]) {
    unsigned near = 0u;
#pragma unroll
    for (int m = 0; m < M; ++m)
        if (gen_tile<WAVE, VK>(w[m], dhi[m], guard, x[m])) near |= 1u << m;
    return near;
}

// M voices per thread, VK rows per tile (M = 4: VK = 8 keeps x[M][VK], the accumulators and the voices'
// phase / filter / weight registers inside 128 registers).
//
// Work decomposition: a *group* is the VT * M voices one CTA holds in registers (all of one segment, i.e. one
// wave x filter kind); the (group, VK-row block) space, group-major, is cut into `npieces` equal contiguous pieces,
// one per CTA, so a bank of ANY size fills the machine with equally long pieces (C5 on 8 GPUs leaves 131,072
// instances = 132 groups per GPU for 296 CTA slots).  A piece is walked as sub-ranges [b0, b1) of one group each;
// a sub-range that starts inside a group begins `warm_rows` earlier from zero filter state without storing (the
// bank's decay horizon, sized by the host to 2^-40), oscillators need no warm-up at all; the sub-range that reaches
// the end of the launch hands the filter state to the next call.
template <int M, int VK>
__global__ void __launch_bounds__(VT, M == 4 ? 2 : 3) k_voices(const __grid_constant__ VoicesDev a) {
    __shared__ float2 red[VK * VT];
    const int tid = threadIdx.x;
    const int bpg = (a.frames + VK - 1) / VK;                  // blocks per group
    const int64_t total = (int64_t)a.ngroups * bpg;
    int64_t blk = total * blockIdx.x / a.npieces;
    const int64_t blk_end = total * (blockIdx.x + 1) / a.npieces;
    const double rate = (double)a.rate;
    constexpr int PARTS = VT / VK;          // threads that share one row in the CTA reduction

  while (blk < blk_end) {
    const int grp = (int)(blk / bpg);
    const int b0 = (int)(blk - (int64_t)grp * bpg);
    const int b1 = (int)min((int64_t)bpg, b0 + (blk_end - blk));
    blk += b1 - b0;
    int si = 0;
    for (int i = 1; i < a.nseg; ++i)
        if (grp >= a.seg[i].cta0) si = i;
    const VoiceSeg& sg = a.seg[si];
    const int cta = grp - sg.cta0;
    const int row_store = b0 * VK;
    const int row_end = min(a.frames, b1 * VK);
    const int row_begin = b0 == 0 ? 0 : max(0, row_store - (sg.nsec ? (a.warm_rows + VK - 1) / VK * VK : 0));
    const bool first_seg = row_begin == 0;
    const int wave = sg.wave, guard = sg.guard;
    const int fk = sg.nsec == 0 ? -1 : sg.sec_kind;
    const size_t C = (size_t)sg.C;

    unsigned long long th[M], dK[M];
    int dhi[M], chan[M];
    float g[M], cf[M], d[M], s1[M], s2[M], s3[M];
    float2 wt[M];
#ifndef SIGB_VOICES_SCALAR
    const bool dform = M == 4 && fk >= 0 && !(fk & SEC_FIRST_ORDER);       // delta-form registers (filt_tile)
#else
    const bool dform = false;
#endif
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int c = (cta * M + m) * VT + tid;
        const bool live = c < sg.C;
        const int cc = live ? c : sg.C - 1;
        chan[m] = live ? c : -1;
        const unsigned long long dth = sg.dtheta[cc];
        th[m] = sg.theta0[cc] + (unsigned long long)(a.position + row_begin) * dth + 0x80000000ull;   // + 1/2 ulp of the top word
        dK[m] = dth * (unsigned long long)VK;
        dhi[m] = (int)((dth + 0x80000000ull) >> 32);
        wt[m] = live ? make_float2(sg.wl[cc], sg.wr[cc]) : make_float2(0.0f, 0.0f);
        g[m] = cf[m] = d[m] = s1[m] = s2[m] = s3[m] = 0.0f;
        if (fk >= 0) {
            g[m] = sg.coef[0 * C + cc];
            cf[m] = sg.coef[1 * C + cc];
            d[m] = sg.coef[2 * C + cc];
            s1[m] = first_seg ? (float)sg.state[0 * C + cc] : 0.0f;
            s2[m] = first_seg ? (float)sg.state[1 * C + cc] : 0.0f;
            if (dform) {
                // (g, c, d | s1, s2) -> delta-form registers; same formulas as delta_coef / delta_state_in of sigb_reg.cu
                const double G = (double)g[m], Dd = (double)d[m], R2 = (double)cf[m] - G;
                const double gd2 = 2.0 * G * Dd, F = 4.0 * G * G * Dd;
                const double S1 = first_seg ? sg.state[0 * C + cc] : 0.0, S2 = first_seg ? sg.state[1 * C + cc] : 0.0;
                const double Dst = S1 / gd2;
                s1[m] = (float)Dst;
                if (fk & SEC_HP) {
                    wt[m].x *= d[m];                                  // the high-pass output scale rides on the weights
                    wt[m].y *= d[m];
                    g[m] = (float)(-R2 * gd2);                        // -Q
                    cf[m] = (float)(-F);
                    s2[m] = (float)(0.5 * (double)cf[m] * Dst - S2);   // -4 Z = -(s2 + (F/2) D), with the float32 F the store inverts
                } else {
                    g[m] = (float)(1.0 - R2 * gd2);                   // a
                    cf[m] = (float)(0.25 * F);
                    s2[m] = (float)(0.25 * (S2 + 2.0 * (double)cf[m] * Dst));
                    s3[m] = (float)(0.5 * S2);
                }
            }
        }
    }
    float2* part_out = reinterpret_cast<float2*>(a.partial) + (size_t)grp * a.frames;

    for (int n0 = row_begin; n0 < row_end; n0 += VK) {
        const int kmax = min(VK, row_end - n0);
        float x[M][VK];
        int w[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            w[m] = (int)(th[m] >> 32);
            th[m] += dK[m];
        }
        unsigned near;
        switch (wave) {
            case SIGB_WAVE_SINE: near = gen_tiles<SIGB_WAVE_SINE, M, VK>(w, dhi, guard, x); break;
            case SIGB_WAVE_SQUARE: near = gen_tiles<SIGB_WAVE_SQUARE, M, VK>(w, dhi, guard, x); break;
            case SIGB_WAVE_SAWTOOTH: near = gen_tiles<SIGB_WAVE_SAWTOOTH, M, VK>(w, dhi, guard, x); break;
            default: near = gen_tiles<SIGB_WAVE_TRIANGLE, M, VK>(w, dhi, guard, x); break;
        }
        if (near) {
            // a sample within `guard` of a discontinuity: redo that voice's tile with the reference's own
            // float64 arithmetic (osc.py:32), so the jump lands on the same sample as in numpy
#pragma unroll
            for (int m = 0; m < M; ++m) {
                if (((near >> m) & 1u) && chan[m] >= 0) {
                    const double hz = sg.hertz[chan[m]], ph = sg.phase[chan[m]];
#pragma unroll
                    for (int k = 0; k < VK; ++k)
                        x[m][k] = osc_wave(wave, osc_cycles(__ddiv_rn((double)(a.position + n0 + k), rate), hz, ph));
                }
            }
        }
        switch (fk) {
            case 0: filt_tile<0, M, VK>(x, g, cf, d, s1, s2, s3, kmax); break;
            case SEC_HP: filt_tile<SEC_HP, M, VK>(x, g, cf, d, s1, s2, s3, kmax); break;
            case SEC_FIRST_ORDER: filt_tile<SEC_FIRST_ORDER, M, VK>(x, g, cf, d, s1, s2, s3, kmax); break;
            case SEC_FIRST_ORDER | SEC_HP: filt_tile<SEC_FIRST_ORDER | SEC_HP, M, VK>(x, g, cf, d, s1, s2, s3, kmax); break;
            default: break;
        }
        if (n0 + VK <= row_store) continue;             // warm-up tile: nothing to reduce or store (uniform over the CTA)
        // CTA reduction in a fixed order: VK rows x 256 threads -> VK float2
#pragma unroll
        for (int k = 0; k < VK; ++k) {
            float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int m = 0; m < M; ++m) acc = __ffma2_rn(wt[m], make_float2(x[m][k], x[m][k]), acc);
            red[k * VT + tid] = acc;
        }
        __syncthreads();
        {
            const int k = tid / PARTS, part = tid % PARTS;
            float2 s = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int i = 0; i < VT / PARTS; ++i) s = __fadd2_rn(s, red[k * VT + i * PARTS + part]);
#pragma unroll
            for (int o = PARTS / 2; o > 0; o >>= 1) {
                s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
                s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
            }
            if (part == 0 && k < kmax && n0 + k >= row_store) part_out[n0 + k] = s;
        }
        __syncthreads();
    }
    if (fk >= 0 && row_end == a.frames) {      // the sub-range that ends the launch hands the filter state to the next call
#pragma unroll
        for (int m = 0; m < M; ++m) {
            if (chan[m] >= 0) {
                double S1 = (double)s1[m], S2 = (double)s2[m];
                if (dform) {
                    // back to the plan's state-variable convention (delta_state_out of sigb_reg.cu); d[m] still holds d
                    const double gd2 = 2.0 * (double)sg.coef[0 * C + chan[m]] * (double)d[m];
                    S1 = gd2 * (double)s1[m];
                    S2 = (fk & SEC_HP) ? 0.5 * (double)cf[m] * (double)s1[m] - (double)s2[m]
                                       : 4.0 * (double)s2[m] - 2.0 * (double)cf[m] * (double)s1[m];
                }
                sg.state_out[0 * C + chan[m]] = S1;
                sg.state_out[1 * C + chan[m]] = S2;
            }
        }
    }
  }
}

__global__ void __launch_bounds__(256) k_voices_finish(const float* __restrict__ partial, int nparts, int frames,
                                                       float* __restrict__ out, int64_t ld_out) {
    const int64_t total = (int64_t)frames * 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int p = 0; p < nparts; ++p) s += (double)__ldg(partial + (int64_t)p * total + i);
        out[(i >> 1) * ld_out + (i & 1)] = (float)s;
    }
}

}  // namespace

static int g_bank_unroll = 2;      // measured on C3: 5.45e12 (1) / 5.99e12 (2) partial-samples/s
extern "C" void sigb_set_bank_unroll(int n) { g_bank_unroll = n; }

extern "C" int sigb_launch_bank(const BankDev* a, void* stream) {
    if (a->frames <= 0 || a->P <= 0 || a->groups <= 0) return 0;
    const int tiles = (a->frames + BANK_TN - 1) / BANK_TN;
    const int gblocks = (a->groups + BANK_GB - 1) / BANK_GB;
    const long long items = (long long)tiles * gblocks;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_bank_unroll == 2) {
        const int grid = (int)(items < (long long)sms * 3 ? items : (long long)sms * 3);
        k_bank<2><<<grid, BANK_THREADS, 0, (cudaStream_t)stream>>>(*a, (int)items, gblocks);
    } else {
        const int grid = (int)(items < (long long)sms * 4 ? items : (long long)sms * 4);
        k_bank<1><<<grid, BANK_THREADS, 0, (cudaStream_t)stream>>>(*a, (int)items, gblocks);
    }
    return (int)cudaGetLastError();
}

extern "C" int sigb_voices_ctas(int channels, int M) { return (channels + VT * M - 1) / (VT * M); }

extern "C" int sigb_voices_block_rows(int M) { return M == 4 ? 8 : 16; }

extern "C" int sigb_voices_slots(int M) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * (M == 4 ? 2 : 3);
}

extern "C" int sigb_launch_voices(const VoicesDev* a, void* stream) {
    if (a->frames <= 0 || a->ngroups <= 0 || a->npieces <= 0) return 0;
    if (a->M == 4) k_voices<4, 8><<<a->npieces, VT, 0, (cudaStream_t)stream>>>(*a);
    else k_voices<1, 16><<<a->npieces, VT, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int sigb_launch_voices_finish(const float* partial, int nparts, int frames, float* out, int64_t ld_out, void* stream) {
    if (frames <= 0) return 0;
    const int64_t total = (int64_t)frames * 2;
    const int blocks = (int)(total + 255) / 256 < 148 * 8 ? (int)((total + 255) / 256) : 148 * 8;
    k_voices_finish<<<blocks, 256, 0, (cudaStream_t)stream>>>(partial, nparts, frames, out, ld_out);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// k_param_eval: block-rate parameter graphs (an LFO on hertz / phase / gain / mix / exponent)
//
// The reference samples such parameters once per request, at the request's first frame, in float64
// (forward_at_block_rate, chain/__init__.py:305-306; osc.py:28-30, fx.py:39,52,59).  One thread per
// channel interprets the (tiny) program for its own channel in float64 -- width-1 rows are recomputed by
// every thread, so there is no cross-thread dependency -- and publishes each row as float64 (oscillator
// hertz / phase) and float32 (gain / mix / exponent).
// ---------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ double osc_wave_f64(int wave, double cyc) {
    switch (wave) {
        case SIGB_WAVE_SINE: return sin(__dmul_rn(__dmul_rn(cyc, 2.0), 3.141592653589793));     // np.sin(t * 2 * np.pi), osc.py:43
        case SIGB_WAVE_SQUARE: return np_sign(__dadd_rn(0.5, -np_mod<0>(cyc)));
        case SIGB_WAVE_SAWTOOTH: return __dadd_rn(__dmul_rn(2.0, np_mod<0>(__dadd_rn(cyc, -0.5))), -1.0);
        default: {
            const double t = __dadd_rn(cyc, -0.25);
            return __dmul_rn(__dadd_rn(__dmul_rn(4.0, np_mod<1>(t)), -1.0), np_sign(__dadd_rn(np_mod<0>(t), -0.5)));
        }
    }
}

__global__ void __launch_bounds__(128) k_param_eval(const ParamInstr* __restrict__ prog, int n_instr, int n_rows, double* drows,
                                                    float* frows, int row_stride, int64_t position, const int64_t* pos_ptr, int rate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= row_stride) return;
    if (pos_ptr) position = *pos_ptr;                     // realtime graphs: the block header the host rewrites per launch
    double v[SIGB_PARAM_ROWS];
    for (int r = 0; r < n_rows; ++r) v[r] = drows[(size_t)r * row_stride + c];      // constants are preloaded (replicated)
    const double tn = __ddiv_rn((double)position, (double)rate);
    for (int i = 0; i < n_instr; ++i) {
        const ParamInstr in = prog[i];
        double y;
        switch (in.op) {
            case PRM_OSC: y = osc_wave_f64(in.wave, osc_cycles(tn, v[in.a], v[in.b])); break;               // osc.py:32
            case PRM_MUL: y = __dmul_rn(v[in.a], v[in.b]); break;                                             // fx.py:46, 52
            case PRM_MIX: y = __dadd_rn(__dmul_rn(v[in.c], v[in.a]), __dmul_rn(__dadd_rn(1.0, -v[in.c]), v[in.b])); break;   // fx.py:40
            case PRM_AMP: y = copysign(pow(v[in.a], v[in.b]), v[in.a]); break;                                // fx.py:60
            default: y = v[in.a]; break;
        }
        v[in.dst] = y;
        drows[(size_t)in.dst * row_stride + c] = y;
        frows[(size_t)in.dst * row_stride + c] = (float)y;
    }
}
}  // namespace

extern "C" int sigb_launch_param_eval(const ParamInstr* prog_dev, int n_instr, int n_rows, double* drows, float* frows, int row_stride,
                                      int64_t position, const int64_t* pos_ptr, int rate, void* stream) {
    if (n_instr <= 0) return 0;
    k_param_eval<<<(row_stride + 127) / 128, 128, 0, (cudaStream_t)stream>>>(prog_dev, n_instr, n_rows, drows, frows, row_stride, position, pos_ptr, rate);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// k_design: per-request filter design for MODULATED cutoffs.  The reference samples a filter's cutoff once per
// request (SingleCritFilter._eval -> forward_at_block_rate, chain/fx.py:124-129) and designs butter(N, Wn) per
// channel per block (fx.py:98-102); here one thread per channel turns the cutoff row of the parameter program
// into the {g, c, d} coefficients of the filter's sections -- sigb_butter_sections / sigb_section_coef
// (sigb_design.cpp) restated in float64 on the device -- straight into the chain's coefficient table.
// Wn is clipped to [0, 1] as in fx.py:100-101; scipy then rejects Wn outside (0, 1) with a ValueError
// ("Digital filter critical frequencies must be 0 < Wn < 1").  A device-side design cannot raise: it sets
// *err_flag (page-locked host memory the runtime checks after the render, SIGB_ECRIT) and designs that channel just
// inside the open interval so that the block itself stays finite.
// ---------------------------------------------------------------------------------------------
namespace {
// one sample of a section in float64 (sigb_section_step of sigb_design.cpp)
__device__ __forceinline__ double design_step(int kind, double g, double r2, double x, double& s1, double& s2) {
    if (kind & SEC_FIRST_ORDER) {
        const double G = g / (1.0 + g);
        const double v = (x - s1) * G;
        const double lp = v + s1;
        s1 = lp + v;
        return (kind & SEC_HP) ? x - lp : lp;
    }
    const double d = 1.0 / (1.0 + r2 * g + g * g);
    const double hp = (x - (r2 + g) * s1 - s2) * d;
    const double bp = g * hp + s1;
    s1 = g * hp + bp;
    const double lp = g * bp + s2;
    s2 = g * bp + lp;
    return (kind & SEC_HP) ? hp : lp;
}

__global__ void __launch_bounds__(128) k_design(const DesignDev a) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int C = a.C;
    double warm = 0.0;
    if (c < C) {
        const double pi = 3.14159265358979323846;
        double wn = a.cutoff[c] / ((double)a.rate / 2.0);
        if (!(wn > 0.0 && wn < 1.0)) {                 // also catches NaN
            if (a.err_flag) *a.err_flag = 1;
            wn = wn >= 1.0 ? 1.0 - 1e-9 : 1e-9;
        }
        const double g = tan(pi * wn / 2.0);
        const int nsec = a.order / 2 + (a.order & 1);
        for (int k = 0; k < nsec; ++k) {
            const bool first = k == a.order / 2;        // the odd order's first-order section comes last
            const int kind = (a.highpass ? SEC_HP : 0) | (first ? SEC_FIRST_ORDER : 0);
            const double r2 = first ? 0.0 : 2.0 * sin(pi * (2.0 * k + 1.0) / (2.0 * a.order));
            const size_t s = (size_t)(a.s0 + k);
            if (first) {
                a.coef[(s * 3 + 0) * C + c] = (float)(g / (1.0 + g));
                a.coef[(s * 3 + 1) * C + c] = 0.0f;
                a.coef[(s * 3 + 2) * C + c] = 0.0f;
            } else {
                a.coef[(s * 3 + 0) * C + c] = (float)g;
                a.coef[(s * 3 + 1) * C + c] = (float)(r2 + g);
                a.coef[(s * 3 + 2) * C + c] = (float)(1.0 / (1.0 + r2 * g + g * g));
            }
            if (!a.apow) continue;
            // the linear-system tables of the time-parallel kernels (make_chain of sigb_plan.cu, on the device):
            // images of the two unit states over SIGB_SCAN_L rows of zero input
            double a1 = 1.0, a2 = 0.0, b1 = 0.0, b2 = 1.0, m1[4] = {1.0, 0.0, 0.0, 1.0}, ya0 = 0.0, yb0 = 0.0;
            for (int r = 0; r < SIGB_SCAN_L; ++r) {
                const double ya = design_step(kind, g, r2, 0.0, a1, a2);
                const double yb = design_step(kind, g, r2, 0.0, b1, b2);
                a.ztab[((s * SIGB_SCAN_L + r) * 2 + 0) * C + c] = (float)ya;
                a.ztab[((s * SIGB_SCAN_L + r) * 2 + 1) * C + c] = (float)yb;
                if (r == 0) { ya0 = ya; yb0 = yb; }
                if (r == 1) {       // first differences of the zero-input output (delta form of its recurrence)
                    a.hrec[(s * 4 + 2) * C + c] = (float)(ya - ya0);
                    a.hrec[(s * 4 + 3) * C + c] = (float)(yb - yb0);
                }
                if (r == 0) { m1[0] = a1; m1[1] = b1; m1[2] = a2; m1[3] = b2; }
                if (r == SIGB_SCAN_L / 2 - 1) {
                    const double mh[4] = {a1, b1, a2, b2};
                    for (int j = 0; j < 4; ++j) {
                        a.apow_h[(s * 4 + j) * C + c] = mh[j];
                        a.m8[(s * 4 + j) * C + c] = (float)mh[j];
                    }
                }
            }
            a.apow[(s * 4 + 0) * C + c] = a1; a.apow[(s * 4 + 1) * C + c] = b1;
            a.apow[(s * 4 + 2) * C + c] = a2; a.apow[(s * 4 + 3) * C + c] = b2;
            const double tr = m1[0] + m1[3], det = m1[0] * m1[3] - m1[1] * m1[2];
            a.hrec[(s * 4 + 0) * C + c] = (float)det;
            a.hrec[(s * 4 + 1) * C + c] = (float)(tr - 1.0 - det);
            // decay horizon (sigb_section_decay_rows): rows until the zero-input response is below 2^-40, doubled
            const double disc = tr * tr - 4.0 * det;
            const double rho = disc < 0.0 ? sqrt(fabs(det)) : fmax(fabs(tr + sqrt(disc)), fabs(tr - sqrt(disc))) / 2.0;
            warm += !(rho < 1.0) ? 1e9 : (rho < 1e-12 ? 2.0 : 2.0 * (40.0 * 0.6931471805599453 / -log(rho)) + 16.0);
        }
    }
    if (a.warm_out) {
        // sections in series: a channel's decays are budgeted one after another; the filter's horizon is the
        // slowest channel's (the host adds the horizons of the chain's filters)
        int w = (int)fmin(ceil(warm), 1.0e9);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
        if ((threadIdx.x & 31) == 0 && w > 0) atomicMax(a.warm_out, w);
    }
}
}  // namespace

// k_osc_tables: a Sine whose hertz / phase are driven by emitters (vibrato).  The rows are constant within a request, so the
// exact Q0.64 phase theta0 + n dtheta of the constant oscillator holds for the request: one thread per channel restates
// frac_q64 / ratio_q64 of sigb_plan.cu (exact integer arithmetic on the float64 mantissa) and the (cos, sin) of the one-row
// advance, once per request, and the sine fast paths run unchanged.
namespace {
__device__ unsigned long long dev_frac_q64(double x) {
    if (!isfinite(x)) return 0ull;
    const double f = x - floor(x);
    if (!(f < 1.0)) return 0ull;
    const double scaled = ldexp(f, 32);
    const double hi = floor(scaled);
    const double lo = scaled - hi;
    return ((unsigned long long)hi << 32) | (unsigned long long)floor(ldexp(lo, 32));
}
__device__ unsigned long long dev_ratio_q64(double hertz, int rate) {
    if (!isfinite(hertz) || rate <= 0 || hertz == 0.0) return 0ull;
    int e;
    const double m = frexp(fabs(hertz), &e);
    const unsigned long long mant = (unsigned long long)ldexp(m, 53);
    e -= 53;
    const unsigned __int128 R = (unsigned __int128)(unsigned)rate;
    unsigned long long q;
    if (e >= 0) {
        unsigned __int128 r = (unsigned __int128)mant % R;
        for (int i = 0; i < e; ++i) r = (r << 1) % R;
        q = (unsigned long long)((r << 64) / R);
    } else {
        const int k = -e;
        unsigned __int128 num;
        if (k <= 64) num = (unsigned __int128)mant << (64 - k);
        else num = (k - 64 >= 64) ? 0 : ((unsigned __int128)mant >> (k - 64));
        q = (unsigned long long)(num / R);
    }
    return hertz < 0.0 ? (0ull - q) : q;
}
__global__ void __launch_bounds__(128) k_osc_tables(int C, const double* __restrict__ hertz, const double* __restrict__ phase, int rate,
                                                    unsigned long long* __restrict__ theta0, unsigned long long* __restrict__ dtheta,
                                                    float* __restrict__ rot1) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const unsigned long long dt = dev_ratio_q64(hertz[c], rate);
    theta0[c] = dev_frac_q64(phase[c]);
    dtheta[c] = dt;
    if (rot1) {
        const double ang = 6.283185307179586476925 * ldexp((double)(long long)dt, -64);
        rot1[2 * c + 0] = (float)cos(ang);
        rot1[2 * c + 1] = (float)sin(ang);
    }
}
}  // namespace

extern "C" int sigb_launch_osc_tables(int C, const double* hertz, const double* phase, int rate, unsigned long long* theta0,
                                      unsigned long long* dtheta, float* rot1, void* stream) {
    if (C <= 0) return 0;
    k_osc_tables<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(C, hertz, phase, rate, theta0, dtheta, rot1);
    return (int)cudaGetLastError();
}

// k_gain_rows: a modulated Gain at the end of a chain (tremolo; fx.py:49-52, `right` sampled once per request): the chain's
// per-channel output gain for this request = the folded constant gains x the sampled row.
namespace {
__global__ void __launch_bounds__(256) k_gain_rows(int C, const float* __restrict__ gain_const, const double* __restrict__ row,
                                                   float* __restrict__ gain_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    gain_out[c] = (float)((gain_const ? (double)gain_const[c] : 1.0) * row[c]);
}
}  // namespace

extern "C" int sigb_launch_gain_rows(int C, const float* gain_const, const double* row, float* gain_out, void* stream) {
    if (C <= 0) return 0;
    k_gain_rows<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(C, gain_const, row, gain_out);
    return (int)cudaGetLastError();
}

// k_pan_weights: PanSum with a MODULATED pan fused with its voices -- the (L, R) weights gain * (1 - pan), gain * pan of one
// segment from the request's pan row (float64, as the host derives them for a constant pan), once per request.
namespace {
__global__ void __launch_bounds__(256) k_pan_weights(int C, const double* __restrict__ gain, const double* __restrict__ pan,
                                                     float* __restrict__ wl, float* __restrict__ wr) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double g = gain[c], p = pan[c];
    wl[c] = (float)(g * (1.0 - p));
    wr[c] = (float)(g * p);
}
}  // namespace

extern "C" int sigb_launch_pan_weights(int C, const double* gain, const double* pan, float* wl, float* wr, void* stream) {
    if (C <= 0) return 0;
    k_pan_weights<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(C, gain, pan, wl, wr);
    return (int)cudaGetLastError();
}

extern "C" int sigb_launch_design(const DesignDev* a, void* stream) {
    if (a->C <= 0) return 0;
    k_design<<<(a->C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// probe: write-only streaming fill (the practical HBM ceiling of a store-only kernel such as C2's)
// ---------------------------------------------------------------------------------------------
namespace {
// store flavour of the probes: 0 = st.global.cs (streaming), 1 = plain st.global (write-back), 2 = st.global.cg, 3 = st.global.wt
int g_probe_store = 0;
__device__ __forceinline__ void probe_store(float4* p, float4 v, int how) {
    if (how == 1) *p = v;
    else if (how == 2) __stcg(p, v);
    else if (how == 3) __stwt(p, v);
    else __stcs(p, v);
}
__global__ void __launch_bounds__(256) k_probe_fill(float4* __restrict__ out, int64_t n4, float v, int how) {
    const float4 val = make_float4(v, v, v, v);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) probe_store(out + i, val, how);
}
}  // namespace

namespace {
// tiled fill: the C2 store pattern -- a warp owns a (rows_per_tile x width_bytes) tile of a (frames, 4096) float
// block and walks time; `width` floats per tile row (32 = the scan kernel's 128-byte rows)
__global__ void __launch_bounds__(256) k_probe_fill_tiled(float* __restrict__ out, int frames, int C, int width, int rows_per_tile, float v, int how) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int tiles_x = C / width;
    const int steps = frames / rows_per_tile;
    const long long total = (long long)tiles_x * steps;
    const float4 val = make_float4(v, v, v, v);
    const int lanes_per_row = width / 4;                 // float4 lanes covering one tile row
    const int rows_per_instr = 32 / lanes_per_row;
    // tile-major order cut into equal contiguous pieces per warp (as the scan kernel's persistent pieces)
    const long long quota = (total + nwarps - 1) / nwarps;
    for (long long it = (long long)warp * quota; it < min(total, (long long)(warp + 1) * quota); ++it) {
        const int tx = (int)(it / steps), st = (int)(it % steps);
        float* base = out + (size_t)st * rows_per_tile * C + (size_t)tx * width;
        for (int r = lane / lanes_per_row; r < rows_per_tile; r += rows_per_instr)
            probe_store(reinterpret_cast<float4*>(base + (size_t)r * C) + (lane % lanes_per_row), val, how);
    }
}
}  // namespace

// mode 1: CTA b owns column stripe b % tiles_x over time piece b / tiles_x and its 8 warps interleave the piece's steps,
// so CTAs b, b + 1, ... write ADJACENT stripes of the same rows at about the same time (what a cluster of CTAs on adjacent
// tiles of the scan kernel would do); mode 2: the same with the stripes permuted (adjacent CTAs far apart in columns).
namespace {
__global__ void __launch_bounds__(256) k_probe_fill_lockstep(float* __restrict__ out, int frames, int C, int width, int rows_per_tile, int mode, float v) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tiles_x = C / width;
    const int steps = frames / rows_per_tile;
    const int pieces_t = gridDim.x / tiles_x;
    if (mode != 3 && (pieces_t == 0 || (int)blockIdx.x >= pieces_t * tiles_x)) return;
    if (mode == 3) {
        // time-major: all warps of the launch sweep the same narrow band of rows together (item = step * tiles_x + stripe),
        // so the launch touches a few 2 MB pages at a time instead of one page per warp
        const int nw = gridDim.x * 8, gw = blockIdx.x * 8 + w;
        const float4 val3 = make_float4(v, v, v, v);
        const int lpr = width / 4, rpi = 32 / lpr;
        for (long long it = gw; it < (long long)tiles_x * steps; it += nw) {
            const int st = (int)(it / tiles_x), txx = (int)(it % tiles_x);
            float* base = out + (size_t)st * rows_per_tile * C + (size_t)txx * width;
            for (int r = lane / lpr; r < rows_per_tile; r += rpi)
                __stcs(reinterpret_cast<float4*>(base + (size_t)r * C) + (lane % lpr), val3);
        }
        return;
    }
    int tx = blockIdx.x % tiles_x;
    if (mode == 2) tx = (int)(((long long)tx * 37) % tiles_x);
    const int pt = blockIdx.x / tiles_x;
    const int s0 = (int)((long long)steps * pt / pieces_t), s1 = (int)((long long)steps * (pt + 1) / pieces_t);
    const float4 val = make_float4(v, v, v, v);
    const int lanes_per_row = width / 4, rows_per_instr = 32 / lanes_per_row;
    for (int st = s0 + w; st < s1; st += 8) {
        float* base = out + (size_t)st * rows_per_tile * C + (size_t)tx * width;
        for (int r = lane / lanes_per_row; r < rows_per_tile; r += rows_per_instr)
            __stcs(reinterpret_cast<float4*>(base + (size_t)r * C) + (lane % lanes_per_row), val);
    }
}
}  // namespace

extern "C" void sigb_probe_set_store(int how) { g_probe_store = how; }

extern "C" int sigb_probe_fill_tiled(float* out_dev, int32_t frames, int32_t C, int32_t width, int32_t rows_per_tile, int32_t blocks, void* stream) {
    if (width < 4 || width > 128 || (width & (width - 1)) != 0 || C % width != 0 || rows_per_tile < 1) return (int)cudaErrorInvalidValue;
    k_probe_fill_tiled<<<blocks, 256, 0, (cudaStream_t)stream>>>(out_dev, frames, C, width, rows_per_tile, 1.0f, g_probe_store);
    return (int)cudaGetLastError();
}

extern "C" int sigb_probe_fill_lockstep(float* out_dev, int32_t frames, int32_t C, int32_t width, int32_t rows_per_tile, int32_t blocks, int32_t mode, void* stream) {
    if (width < 4 || width > 128 || (width & (width - 1)) != 0 || C % width != 0 || rows_per_tile < 1) return (int)cudaErrorInvalidValue;
    k_probe_fill_lockstep<<<blocks, 256, 0, (cudaStream_t)stream>>>(out_dev, frames, C, width, rows_per_tile, mode, 1.0f);
    return (int)cudaGetLastError();
}

extern "C" int sigb_probe_fill(float* out_dev, int64_t n_floats, float value, int32_t blocks, void* stream) {
    k_probe_fill<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(out_dev), n_floats / 4, value, g_probe_store);
    return (int)cudaGetLastError();
}
