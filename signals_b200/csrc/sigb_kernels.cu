// Hand-written sm_100a kernels for the `signals` block-render hot path.
//
//   k_chain_seq   one thread per channel walking time sequentially (state in registers);
//                 stateless chains (no filter sections) are tiled over time as well.
//   k_chain_scan2 / k_chain_scan3
//                 the time-parallel kernels: a CTA owns a tile of 32 / 64 channels and walks time in
//                 steps; each worker warp renders one sub-chunk of a step from zero filter state,
//                 a scanner warp chains the sub-chunk end states with the 2x2 transition
//                 A^L of each state-variable section, and the workers add the zero-input
//                 response of the true initial state before storing.
//   k_ewise       binary / pointwise nodes on materialised blocks (Mix, RingMod, Amp, Gain, copy).
//   k_reduce      channel reductions (GroupSum, PanSum) of a materialised block.
//
// Reference semantics being reproduced (file:line under /root/reference/src/signals/chain):
//   osc.py:26-62   cycles = frame_range / rate * hertz + phase (float64, that op order), waveforms
//   fx.py:35-60    Mix / RingMod / Gain / Amp
//   fx.py:85-121   per-channel Butterworth low/high-pass == cascade of bilinear 2nd-order sections;
//                  run here as zero-delay-feedback state-variable sections (same transfer function,
//                  far better float32 behaviour at low cutoffs than the direct form scipy uses)
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <string.h>

#include "sigb200.h"
#include "sigb_internal.h"
#include "sigb_device.cuh"

namespace {

using namespace sigb_dev;

constexpr int L = SIGB_SCAN_L;
int g_scan_tma = 1;   // staged TMA tensor stores in k_chain_scan (0: direct STG)
int g_scan_split = 1; // k_chain_scan2: split tiles along time across SMs (decay warm-up)
int g_scan_rot = 1;   // k_chain_scan3: sine rows from two-pipe rotations (one sin/cos pair per four rows) instead of one MUFU.SIN per row

int sm_count() {
    static int n = [] {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v > 0 ? v : 148;
    }();
    return n;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// one (L x 32) staging tile -> global memory with a single TMA tensor store (UTMASTG); the tensor
// map clips columns past C, so ragged last tiles take this path too
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* map, int x, int y, const float* ssrc) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x), "r"(y),
                 "r"(smem_u32(ssrc))
                 : "memory");
}

__device__ __forceinline__ float load_src(const ChainDev& a, int64_t row, int c) {
    if (row >= a.src_rows) return 0.0f;
    return __ldg(a.src + row * a.src_ld + (int64_t)c * a.src_cs);
}

// ------------------------------------------------------------------------------------------
// k_chain_seq
// ------------------------------------------------------------------------------------------
template <int SRC, int NSEC>
__global__ void __launch_bounds__(128) k_chain_seq(const ChainDev a, int rows_per_seg) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.C) return;
    int r0 = 0, r1 = a.frames;
    if (NSEC == 0) {                      // stateless: time is tiled over blockIdx.y
        r0 = blockIdx.y * rows_per_seg;
        r1 = min(a.frames, r0 + rows_per_seg);
    }
    float g[NSEC > 0 ? NSEC : 1], cc[NSEC > 0 ? NSEC : 1], d[NSEC > 0 ? NSEC : 1];
    float s1[NSEC > 0 ? NSEC : 1], s2[NSEC > 0 ? NSEC : 1];
#pragma unroll
    for (int s = 0; s < NSEC; ++s) {
        g[s] = a.coef[(size_t)(s * 3 + 0) * a.C + c];
        cc[s] = a.coef[(size_t)(s * 3 + 1) * a.C + c];
        d[s] = a.coef[(size_t)(s * 3 + 2) * a.C + c];
        s1[s] = (float)a.state[(size_t)(s * 2 + 0) * a.C + c];
        s2[s] = (float)a.state[(size_t)(s * 2 + 1) * a.C + c];
    }
    const float gain = a.gain ? a.gain[c] : 1.0f;
    double hz = 0.0, ph = 0.0;
    float cv = 0.0f;
    if (SRC == SRC_OSC) { hz = a.hertz[c]; ph = a.phase[c]; }
    if (SRC == SRC_CONST) cv = a.constv[c];
    const double rate = (double)a.rate;
    const int64_t position = a.pos_ptr ? *a.pos_ptr : a.position;     // realtime graphs read the block header
    float* outp = a.out + (int64_t)r0 * a.ld_out + c;
    // fused Mix / RingMod epilogue (stateless chains only): the other operand never leaves registers when it is
    // an oscillator, and is read once when it is a materialised block
    const int epi = a.epi_op;
    double hz2 = 0.0, ph2 = 0.0;
    float g2 = 1.0f, mixp = 0.0f;
    if (epi) {
        if (a.epi_wave >= 0) { hz2 = a.epi_hertz[c]; ph2 = a.epi_phase[c]; g2 = a.epi_gain ? a.epi_gain[c] : 1.0f; }
        if (a.epi_p) mixp = a.epi_p[c];
    }
    for (int r = r0; r < r1; ++r) {
        float x;
        double tn = 0.0;
        if (SRC == SRC_OSC || epi) tn = __ddiv_rn((double)(position + r), rate);
        if (SRC == SRC_OSC) {
            x = osc_wave(a.wave, osc_cycles(tn, hz, ph));
        } else if (SRC == SRC_BUF) {
            x = load_src(a, r, c);
        } else {
            x = cv;
        }
#pragma unroll
        for (int s = 0; s < NSEC; ++s) x = svf_any(a.sec_kind[s], x, g[s], cc[s], d[s], s1[s], s2[s]);
        x *= gain;
        if (epi) {
            float o;
            if (a.epi_wave >= 0) o = osc_wave(a.epi_wave, osc_cycles(tn, hz2, ph2)) * g2;
            else o = (a.epi_rows >= 0 && r >= a.epi_rows) ? 0.0f : __ldg(a.epi_buf + (int64_t)r * a.epi_ld + (int64_t)c * a.epi_cs);
            const float left = a.epi_side ? o : x, right = a.epi_side ? x : o;
            x = epi == EW_MIX ? mixp * left + (1.0f - mixp) * right : left * right;      // fx.py:40, 46
        }
        __stcs(outp, x);
        outp += a.ld_out;
    }
#pragma unroll
    for (int s = 0; s < NSEC; ++s) {
        a.state[(size_t)(s * 2 + 0) * a.C + c] = (double)s1[s];
        a.state[(size_t)(s * 2 + 1) * a.C + c] = (double)s2[s];
    }
}

// ------------------------------------------------------------------------------------------
// k_chain_rt: the latency kernel -- one WARP per channel.
//
// k_chain_seq walks a channel's rows in one thread: source (float64 phase, division, waveform) and filter recurrence
// alternate in a single dependency chain, ~400 cycles per row for an oscillator source (measured: a 512-frame
// block of Triangle -> Gain -> LowPass took 110 us).  For the audio callback's blocks (SinkDevice._callback,
// chain/dev.py:167-179: a few channels x 128..1024 rows) the machine is empty, so a warp per channel splits the two:
// the 32 lanes evaluate the source of 32 consecutive rows in parallel, then every lane runs the SAME sequential
// recurrence over those 32 samples (broadcast by shuffle; redundant lanes cost nothing under SIMT) and lane k keeps
// row k.  What remains on the critical path is the recurrence itself, ~20 cycles per row.
// Same arithmetic in the same order as k_chain_seq: the two kernels agree bit for bit.
// ------------------------------------------------------------------------------------------
template <int SRC, int NSEC>
__global__ void __launch_bounds__(128) k_chain_rt(const ChainDev a) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= a.C) return;
    float g[NSEC], cc[NSEC], d[NSEC], s1[NSEC], s2[NSEC];
#pragma unroll
    for (int s = 0; s < NSEC; ++s) {
        g[s] = a.coef[(size_t)(s * 3 + 0) * a.C + c];
        cc[s] = a.coef[(size_t)(s * 3 + 1) * a.C + c];
        d[s] = a.coef[(size_t)(s * 3 + 2) * a.C + c];
        s1[s] = (float)a.state[(size_t)(s * 2 + 0) * a.C + c];
        s2[s] = (float)a.state[(size_t)(s * 2 + 1) * a.C + c];
    }
    const float gain = a.gain ? a.gain[c] : 1.0f;
    double hz = 0.0, ph = 0.0;
    float cv = 0.0f;
    if (SRC == SRC_OSC) { hz = a.hertz[c]; ph = a.phase[c]; }
    if (SRC == SRC_CONST) cv = a.constv[c];
    const double rate = (double)a.rate;
    const int64_t position = a.pos_ptr ? *a.pos_ptr : a.position;     // realtime graphs read the block header
    for (int r0 = 0; r0 < a.frames; r0 += 32) {
        const int r = r0 + lane;
        float x = 0.0f;
        if (r < a.frames) {
            if (SRC == SRC_OSC) x = osc_wave(a.wave, osc_cycles(__ddiv_rn((double)(position + r), rate), hz, ph));
            else if (SRC == SRC_BUF) x = load_src(a, r, c);
            else x = cv;
        }
        const int n = min(32, a.frames - r0);
        float mine = 0.0f;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if (k < n) {                                   // uniform over the warp
                float v = __shfl_sync(0xffffffffu, x, k);
#pragma unroll
                for (int s = 0; s < NSEC; ++s) v = svf_any(a.sec_kind[s], v, g[s], cc[s], d[s], s1[s], s2[s]);
                if (lane == k) mine = v;
            }
        }
        if (r < a.frames) a.out[(int64_t)r * a.ld_out + c] = mine * gain;
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NSEC; ++s) {
            a.state[(size_t)(s * 2 + 0) * a.C + c] = (double)s1[s];
            a.state[(size_t)(s * 2 + 1) * a.C + c] = (double)s2[s];
        }
    }
}

// Decay horizon of a chain with modulated cutoffs, in steps: k_design left each modulated filter's horizon (rows, slowest
// channel) in device memory a moment ago on this stream, so the launch needs no host round trip; the host only had an
// estimate (the previous request's value) to decide whether to cut tiles along time at all.
__device__ __forceinline__ int device_warm_steps(const ChainDev& a, int step_rows) {
    long long w = a.warm_rows > 0 ? a.warm_rows : 0;                 // the filters with constant cutoffs
    for (int i = 0; i < a.n_warm_dev; ++i) w += a.warm_dev[i];
    w = min(w, (long long)1 << 30);
    return (int)((w + step_rows - 1) / step_rows);
}

// ------------------------------------------------------------------------------------------
// named barriers of the time-parallel kernels: NG groups of WG worker warps + 1 scanner warp.  Group g renders steps
// g, g+NG, g+2NG, ...; a step is WG sub-chunks of L rows.  Barrier 1+2g: "end states of group g published",
// 2+2g: "initial states for group g published".
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// ------------------------------------------------------------------------------------------
// k_chain_scan2: the packed (f32x2) time-parallel kernel
//
// Decomposition: a CTA owns a tile of 32 channels and walks time in steps; each worker warp renders one sub-chunk of a
// step from zero filter state, a scanner warp chains the sub-chunk end states with the fp64 2x2 transition A^L of each
// state-variable section, and the workers add the zero-input response of the true initial state before storing.
// Every worker thread renders its 16-row sub-chunk as TWO
// independent 8-row halves from zero state, carried in the two lanes of packed float2 registers, so
// the state-variable section, the zero-input correction and the gain run as FFMA2/FMUL2/FADD2 (one
// issue slot per two samples).  The halves are stitched inside the thread with the fp32 transition
// A^8; the scanner warp still chains whole sub-chunks with the fp64 transition A^16.  The
// zero-input response is advanced with its own 2-term recurrence h[k] = tr(A) h[k-1] - det(A) h[k-2]
// (Cayley-Hamilton on the section's transition), so the correction needs no table traffic.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 pk(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 pk1(float a) { return make_float2(a, a); }

struct SecPar {            // per-thread parameters of one section (packed where the math is packed)
    float2 g, nc, d;       // (g,g) (-c,-c) (d,d); first-order: g = (G,G)
    float2 gd, gd2, g2;    // (g d) (2 g d) (2 g): six-instruction low-pass form
    float2 al, be;         // zero-input recurrence in delta form: al = det(A), be = tr(A) - 1 - det(A)
    float m8[4];           // A^8, row-major
    float p0, r0, p1, r1;  // zero-input output at sample 0 (p0, r0) and its first difference (p1, r1), per unit state
};

// With e = x - c s1 - s2:  hp = d e;  bp = s1 + g d e;  s1' = s1 + 2 g d e;  lp = s2 + g bp;  s2' = s2 + 2 g bp
// (the same section as svf_lp2, regrouped: six packed instructions for low-pass, seven for high-pass).
template <int H, bool HP>
__device__ __forceinline__ void svf2_second_order(const SecPar& c, float2 (&v)[H], float2& s1, float2& s2) {
    const float2 neg1 = pk1(-1.0f);
#pragma unroll
    for (int k = 0; k < H; ++k) {
        const float2 xs = __ffma2_rn(s2, neg1, v[k]);
        const float2 e = __ffma2_rn(c.nc, s1, xs);
        const float2 bp = __ffma2_rn(c.gd, e, s1);
        s1 = __ffma2_rn(c.gd2, e, s1);
        const float2 lp = __ffma2_rn(c.g, bp, s2);
        s2 = __ffma2_rn(c.g2, bp, s2);
        v[k] = HP ? __fmul2_rn(e, c.d) : lp;
    }
}

template <int H, bool HP>
__device__ __forceinline__ void svf2_first_order(const SecPar& c, float2 (&v)[H], float2& s1) {
    const float2 neg1 = pk1(-1.0f);
#pragma unroll
    for (int k = 0; k < H; ++k) {
        float2 t = __ffma2_rn(s1, neg1, v[k]);        // x - s1
        float2 w = __fmul2_rn(t, c.g);
        float2 lp = __fadd2_rn(w, s1);
        s1 = __fadd2_rn(lp, w);
        v[k] = HP ? __ffma2_rn(lp, neg1, v[k]) : lp;
    }
}

// the section kind is uniform over the launch: branch once per block of samples, not per sample
template <int H>
__device__ __forceinline__ void svf2_block(int kind, const SecPar& c, float2 (&v)[H], float2& s1, float2& s2) {
    if (kind == 0) svf2_second_order<H, false>(c, v, s1, s2);
    else if (kind == SEC_HP) svf2_second_order<H, true>(c, v, s1, s2);
    else if (kind == SEC_FIRST_ORDER) svf2_first_order<H, false>(c, v, s1);
    else svf2_first_order<H, true>(c, v, s1);
}

template <int SRC, int NSEC, int NG, int WG, bool FASTSINE, int PIPE = 0>
__global__ void __launch_bounds__((NG * WG + 1) * 32, 1)
k_chain_scan2(const ChainDev a, int nsteps, int warm_steps, const __grid_constant__ CUtensorMap out_map, int use_tma) {
    if (a.warm_dev && warm_steps > 0) warm_steps = device_warm_steps(a, WG * L);
    constexpr int NW = NG * WG;
    constexpr int STEP = WG * L;
    constexpr int H = L / 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage = reinterpret_cast<float*>(smem_raw);                        // [NW][L][32] output tiles (TMA source)
    float2* zs = reinterpret_cast<float2*>(stage + NW * L * 32);             // [NW][32]
    float2* si = zs + NW * 32;                                                // [NW][32]
    float4* par = reinterpret_cast<float4*>(si + NW * 32);                    // [NSEC][4][32]
    double* tnb = reinterpret_cast<double*>(par + NSEC * 4 * 32);             // [NW][L]
    // deep cascades: the scanner's transition tables live in shared memory (in registers they spill)
    constexpr bool SMEM_SCAN = NSEC > 2;
    double* smg = tnb + NW * L;                                               // [NSEC][4][32]  A^(16 WG), float64
    float4* smq = reinterpret_cast<float4*>(smg + (SMEM_SCAN ? NSEC * 4 * 32 : 0));   // [NSEC][WG][32] A^(16 q), float32

    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const size_t C = (size_t)a.C;
    const bool bulk = use_tma != 0;
    const double rate = (double)a.rate;

    // Work = (tile, step) pairs in tile-major order, cut into gridDim.x equal contiguous pieces, so a
    // launch with fewer tiles than SMs (C2: 128 tiles, 148 SMs) still fills the machine.  A piece that
    // starts inside a tile first re-renders `warm_steps` steps from zero state without storing: the
    // host sizes them so the filters' memory of the unknown true state has decayed below 2^-40.
    const int tiles = (a.C + 31) / 32;
    const long long total = (long long)tiles * nsteps;
    const long long quota = (total + gridDim.x - 1) / gridDim.x;
    long long lin = (long long)blockIdx.x * quota;
    const long long lin_end = min(total, lin + quota);

    while (lin < lin_end) {
        const int tile_idx = (int)(lin / nsteps);
        const int s0 = (int)(lin - (long long)tile_idx * nsteps);
        const int s1 = (int)min((long long)nsteps, s0 + (lin_end - lin));
        const int w0 = max(0, s0 - warm_steps);
        lin += s1 - s0;

        const int c = tile_idx * 32 + lane;
        const bool live = c < a.C;
        const int cc = live ? c : a.C - 1;

        // per-channel section parameters -> shared memory
        for (int i = threadIdx.x; i < NSEC * 32; i += blockDim.x) {
            const int l = i & 31, s = i >> 5;
            const int ch = min(tile_idx * 32 + l, a.C - 1);
            par[(s * 4 + 0) * 32 + l] = make_float4(a.coef[(size_t)(s * 3 + 0) * C + ch], a.coef[(size_t)(s * 3 + 1) * C + ch],
                                                    a.coef[(size_t)(s * 3 + 2) * C + ch], 0.0f);
            par[(s * 4 + 1) * 32 + l] = make_float4(a.m8[(size_t)(s * 4 + 0) * C + ch], a.m8[(size_t)(s * 4 + 1) * C + ch],
                                                    a.m8[(size_t)(s * 4 + 2) * C + ch], a.m8[(size_t)(s * 4 + 3) * C + ch]);
            par[(s * 4 + 2) * 32 + l] = make_float4(a.hrec[(size_t)(s * 4 + 0) * C + ch], a.hrec[(size_t)(s * 4 + 1) * C + ch], 0.0f, 0.0f);
            // zero-input output at sample 0 (row 0 of the response table) and its first difference (float64 on the host)
            par[(s * 4 + 3) * 32 + l] = make_float4(a.ztab[((size_t)(s * L + 0) * 2 + 0) * C + ch], a.ztab[((size_t)(s * L + 0) * 2 + 1) * C + ch],
                                                    a.hrec[(size_t)(s * 4 + 2) * C + ch], a.hrec[(size_t)(s * 4 + 3) * C + ch]);
        }
        __syncthreads();

        if (w == NW) {
            // ---------------- scanner warp: lane = channel ----------------
            // Per event (one step of one group, one section) the only work on the step-to-step critical
            // path is the float64 carry update c' = A^(16 WG) c + P_WG (two dependent DFMAs).  The prefix
            // offsets P_q = sum_{j<q} A^(16 (q-1-j)) z_j do not depend on the carry, and the initial states
            // s_q = A^(16 q) c + P_q of all sub-chunks are independent of each other.
            double c1[NSEC], c2[NSEC];                 // carries
            if (SMEM_SCAN) {
#pragma unroll 1
                for (int s = 0; s < NSEC; ++s) {
                    double m0[4], acc[4] = {1.0, 0.0, 0.0, 1.0};
#pragma unroll
                    for (int k = 0; k < 4; ++k) m0[k] = a.apow[(size_t)(s * 4 + k) * C + cc];
#pragma unroll
                    for (int q = 0; q < WG; ++q) {
                        smq[(s * WG + q) * 32 + lane] = make_float4((float)acc[0], (float)acc[1], (float)acc[2], (float)acc[3]);
                        const double n0 = m0[0] * acc[0] + m0[1] * acc[2], n1 = m0[0] * acc[1] + m0[1] * acc[3];
                        const double n2 = m0[2] * acc[0] + m0[3] * acc[2], n3 = m0[2] * acc[1] + m0[3] * acc[3];
                        acc[0] = n0; acc[1] = n1; acc[2] = n2; acc[3] = n3;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) smg[(s * 4 + k) * 32 + lane] = acc[k];
                }
#pragma unroll
                for (int s = 0; s < NSEC; ++s) {
                    c1[s] = w0 == 0 ? a.state[(size_t)(s * 2 + 0) * C + cc] : 0.0;
                    c2[s] = w0 == 0 ? a.state[(size_t)(s * 2 + 1) * C + cc] : 0.0;
                }
                __syncwarp();
                // events in (round, section, group) order: all NG groups advance section by section, so a
                // group never waits for another group's whole 8-section chain
                for (int base = w0; base < s1; base += NG) {
#pragma unroll
                    for (int s = 0; s < NSEC; ++s) {
#pragma unroll 1
                      for (int grp = 0; grp < NG; ++grp) {
                        if (base + grp >= s1) break;
                        const float f1 = (float)c1[s], f2 = (float)c2[s];
                        const float4 a16 = smq[(s * WG + 1) * 32 + lane];
                        const double g0 = smg[(s * 4 + 0) * 32 + lane], g1 = smg[(s * 4 + 1) * 32 + lane];
                        const double g2 = smg[(s * 4 + 2) * 32 + lane], g3 = smg[(s * 4 + 3) * 32 + lane];
                        bar_sync(1 + 2 * grp, (WG + 1) * 32);
                        float p1 = 0.0f, p2 = 0.0f;
#pragma unroll
                        for (int q = 0; q < WG; ++q) {
                            const float4 m = smq[(s * WG + q) * 32 + lane];
                            const float2 z = zs[(grp * WG + q) * 32 + lane];
                            si[(grp * WG + q) * 32 + lane] = make_float2(fmaf(m.x, f1, fmaf(m.y, f2, p1)), fmaf(m.z, f1, fmaf(m.w, f2, p2)));
                            const float n1 = fmaf(a16.x, p1, fmaf(a16.y, p2, z.x));
                            const float n2 = fmaf(a16.z, p1, fmaf(a16.w, p2, z.y));
                            p1 = n1;
                            p2 = n2;
                        }
                        bar_arrive(2 + 2 * grp, (WG + 1) * 32);
                        const double n1 = fma(g0, c1[s], fma(g1, c2[s], (double)p1));
                        const double n2 = fma(g2, c1[s], fma(g3, c2[s], (double)p2));
                        c1[s] = n1;
                        c2[s] = n2;
                      }
                    }
                }
            } else {
            double mg[NSEC][4];                        // A^(16 WG) in float64
            float mq[NSEC][WG][4];                     // A^(16 q), q = 0..WG-1, float32
#pragma unroll
            for (int s = 0; s < NSEC; ++s) {
                double m0[4], acc[4] = {1.0, 0.0, 0.0, 1.0};
#pragma unroll
                for (int k = 0; k < 4; ++k) m0[k] = a.apow[(size_t)(s * 4 + k) * C + cc];
#pragma unroll
                for (int q = 0; q < WG; ++q) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) mq[s][q][k] = (float)acc[k];
                    const double n0 = m0[0] * acc[0] + m0[1] * acc[2], n1 = m0[0] * acc[1] + m0[1] * acc[3];
                    const double n2 = m0[2] * acc[0] + m0[3] * acc[2], n3 = m0[2] * acc[1] + m0[3] * acc[3];
                    acc[0] = n0; acc[1] = n1; acc[2] = n2; acc[3] = n3;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) mg[s][k] = acc[k];
                // a piece that starts at the beginning of a tile continues the carried state
                c1[s] = w0 == 0 ? a.state[(size_t)(s * 2 + 0) * C + cc] : 0.0;
                c2[s] = w0 == 0 ? a.state[(size_t)(s * 2 + 1) * C + cc] : 0.0;
            }
            for (int base = w0; base < s1; base += NG) {
#pragma unroll
                for (int s = 0; s < NSEC; ++s) {
#pragma unroll 1
                  for (int grp = 0; grp < NG; ++grp) {
                    if (base + grp >= s1) break;
                    const float f1 = (float)c1[s], f2 = (float)c2[s];
                    bar_sync(1 + 2 * grp, (WG + 1) * 32);
                    float2 z[WG];
#pragma unroll
                    for (int q = 0; q < WG; ++q) z[q] = zs[(grp * WG + q) * 32 + lane];
                    float p1 = 0.0f, p2 = 0.0f;
#pragma unroll
                    for (int q = 0; q < WG; ++q) {
                        si[(grp * WG + q) * 32 + lane] = make_float2(fmaf(mq[s][q][0], f1, fmaf(mq[s][q][1], f2, p1)),
                                                                     fmaf(mq[s][q][2], f1, fmaf(mq[s][q][3], f2, p2)));
                        const float n1 = fmaf(mq[s][1][0], p1, fmaf(mq[s][1][1], p2, z[q].x));
                        const float n2 = fmaf(mq[s][1][2], p1, fmaf(mq[s][1][3], p2, z[q].y));
                        p1 = n1;
                        p2 = n2;
                    }
                    bar_arrive(2 + 2 * grp, (WG + 1) * 32);
                    const double n1 = fma(mg[s][0], c1[s], fma(mg[s][1], c2[s], (double)p1));
                    const double n2 = fma(mg[s][2], c1[s], fma(mg[s][3], c2[s], (double)p2));
                    c1[s] = n1;
                    c2[s] = n2;
                  }
                }
            }
            }
            if (live && s1 == nsteps) {   // the piece that finishes a tile hands its state to the next launch
#pragma unroll
                for (int s = 0; s < NSEC; ++s) {
                    a.state_out[(size_t)(s * 2 + 0) * C + c] = c1[s];
                    a.state_out[(size_t)(s * 2 + 1) * C + c] = c2[s];
                }
            }
        } else if (w < NW) {
            // ---------------- worker warps: lane = channel, warp = sub-chunk ----------------
            const int grp = w / WG, q = w % WG;
            const float gain = a.gain ? a.gain[cc] : 1.0f;
            const float2 gain2 = pk1(gain);
            int64_t row = ((int64_t)w0 + grp) * STEP + (int64_t)q * L;
            const int64_t row_stride = (int64_t)NG * STEP;
            float* outp = a.out + row * a.ld_out + c;
            const int64_t out_stride = row_stride * a.ld_out;
            float* tile = stage + w * (L * 32);

            unsigned long long th = 0, th_step = 0;
            int dhi = 0, dhi_h = 0;
            double hz = 0.0, ph = 0.0;
            float cv = 0.0f;
            if (SRC == SRC_OSC) {
                if (FASTSINE) {
                    const unsigned long long dth = a.dtheta[cc];
                    th = a.theta0[cc] + (unsigned long long)(a.position + row) * dth;
                    th_step = dth * (unsigned long long)row_stride;
                    dhi = (int)((dth + 0x80000000ull) >> 32);
                    dhi_h = (int)((dth * (unsigned long long)H + 0x80000000ull) >> 32);   // half-chunk jump, rounded once
                } else {
                    hz = a.hertz[cc];
                    ph = a.phase[cc];
                }
            }
            if (SRC == SRC_CONST) cv = a.constv[cc];

            auto load_par = [&](int s, SecPar& p) {
                const float4 q0 = par[(s * 4 + 0) * 32 + lane], q1 = par[(s * 4 + 1) * 32 + lane];
                const float4 q2 = par[(s * 4 + 2) * 32 + lane], q3 = par[(s * 4 + 3) * 32 + lane];
                p.g = pk1(q0.x); p.nc = pk1(-q0.y); p.d = pk1(q0.z);
                p.gd = pk1(q0.x * q0.z); p.gd2 = pk1(2.0f * (q0.x * q0.z)); p.g2 = pk1(2.0f * q0.x);
                p.m8[0] = q1.x; p.m8[1] = q1.y; p.m8[2] = q1.z; p.m8[3] = q1.w;
                p.al = pk1(q2.x); p.be = pk1(q2.y);
                p.p0 = q3.x; p.r0 = q3.y; p.p1 = q3.z; p.r1 = q3.w;
            };
            SecPar p0;
            if (NSEC == 1) load_par(0, p0);

            // source samples of the sub-chunk starting at row `grow`: v[k] = (row k of the first half, row k of the second half)
            auto gen = [&](float2 (&v)[H], int64_t grow) {
                if (SRC == SRC_OSC) {
                    if (FASTSINE) {
                        int ha = (int)(th >> 32);
                        int hb = ha + dhi_h;
#pragma unroll
                        for (int k = 0; k < H; ++k) {
                            // (I2F, I2F) -> one FMUL2 by 2*pi*2^-32 -> two __sinf (FMUL.RZ by 1/2pi + MUFU.SIN)
                            const float2 r = __fmul2_rn(pk((float)ha, (float)hb), pk1(1.4629180792671596e-9f));
                            v[k] = pk(__sinf(r.x), __sinf(r.y));
                            ha += dhi;
                            hb += dhi;
                        }
                        th += th_step;
                    } else {
                        if (lane < L) tnb[w * L + lane] = __ddiv_rn((double)(a.position + grow + lane), rate);
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < H; ++k)
                            v[k] = pk(osc_wave(a.wave, osc_cycles(tnb[w * L + k], hz, ph)),
                                      osc_wave(a.wave, osc_cycles(tnb[w * L + H + k], hz, ph)));
                        __syncwarp();
                    }
                } else if (SRC == SRC_BUF) {
                    if (live && grow + L <= a.src_rows) {          // whole sub-chunk inside the source: plain loads
                        const float* sp = a.src + grow * a.src_ld + (int64_t)c * a.src_cs;
#pragma unroll
                        for (int k = 0; k < H; ++k) v[k] = pk(__ldg(sp + (int64_t)k * a.src_ld), __ldg(sp + (int64_t)(H + k) * a.src_ld));
                    } else {
#pragma unroll
                        for (int k = 0; k < H; ++k)
                            v[k] = live ? pk(load_src(a, grow + k, c), load_src(a, grow + H + k, c)) : pk1(0.0f);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < H; ++k) v[k] = pk1(cv);
                }
            };
            // add the zero-input response of the true initial state `ia` (both halves; the second half's
            // initial state follows from the first half's zero-state end state (za1, za2))
            auto correct = [&](const SecPar& ps, float2 (&v)[H], float2 ia, float za1, float za2) {
                const float ib1 = fmaf(ps.m8[0], ia.x, fmaf(ps.m8[1], ia.y, za1));
                const float ib2 = fmaf(ps.m8[2], ia.x, fmaf(ps.m8[3], ia.y, za2));
                // h[k] = tr h[k-1] - det h[k-2] in DELTA form: dh[k] = det dh[k-1] + (tr - 1 - det) h[k-1], h[k] = h[k-1] + dh[k].
                // With tr ~ 2 and det ~ 1 at low cutoffs the direct form amplifies float32 rounding like k^2 (6.6e-6 on a
                // 300 Hz section over 16 rows); the delta form only adds one rounding of h per row.
                float2 h = pk(fmaf(ps.p0, ia.x, ps.r0 * ia.y), fmaf(ps.p0, ib1, ps.r0 * ib2));
                float2 dh = pk(fmaf(ps.p1, ia.x, ps.r1 * ia.y), fmaf(ps.p1, ib1, ps.r1 * ib2));
                v[0] = __fadd2_rn(v[0], h);
#pragma unroll
                for (int k = 1; k < H; ++k) {
                    if (k > 1) dh = __ffma2_rn(ps.al, dh, __fmul2_rn(ps.be, h));
                    h = __fadd2_rn(h, dh);
                    v[k] = __fadd2_rn(v[k], h);
                }
            };
            auto store = [&](float2 (&v)[H], int64_t srow, float* soutp) {
                if (bulk) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < H; ++k) {
                        const float2 o = __fmul2_rn(v[k], gain2);
                        tile[k * 32 + lane] = o.x;
                        tile[(H + k) * 32 + lane] = o.y;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_tile(&out_map, tile_idx * 32, (int)srow, tile);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else if (live) {
#pragma unroll
                    for (int k = 0; k < H; ++k) {
                        const float2 o = __fmul2_rn(v[k], gain2);
                        __stcs(soutp + (int64_t)k * a.ld_out, o.x);
                        __stcs(soutp + (int64_t)(H + k) * a.ld_out, o.y);
                    }
                }
            };

            if (PIPE && NSEC == 1) {
                // Software-pipelined single-section loop: the zero-state render of the NEXT step is issued between
                // publishing this step's end state and waiting for its true initial state, so the scanner's
                // turn-around is hidden behind useful work instead of stalling the warp at the barrier.
                const int kind = a.sec_kind[0];
                int step = w0 + grp;
                float2 vn[H];
                float2 s1n = pk1(0.0f), s2n = pk1(0.0f);
                if (step < s1) {
                    gen(vn, row);
                    svf2_block<H>(kind, p0, vn, s1n, s2n);
                }
                while (step < s1) {
                    zs[w * 32 + lane] = make_float2(fmaf(p0.m8[0], s1n.x, fmaf(p0.m8[1], s2n.x, s1n.y)),
                                                    fmaf(p0.m8[2], s1n.x, fmaf(p0.m8[3], s2n.x, s2n.y)));
                    bar_arrive(1 + 2 * grp, (WG + 1) * 32);
                    float2 v[H];
#pragma unroll
                    for (int k = 0; k < H; ++k) v[k] = vn[k];
                    const float za1 = s1n.x, za2 = s2n.x;
                    const int next = step + NG;
                    if (next < s1) {
                        s1n = pk1(0.0f);
                        s2n = pk1(0.0f);
                        gen(vn, row + row_stride);
                        svf2_block<H>(kind, p0, vn, s1n, s2n);
                    }
                    bar_sync(2 + 2 * grp, (WG + 1) * 32);
                    correct(p0, v, si[w * 32 + lane], za1, za2);
                    if (step >= s0) store(v, row, outp);
                    outp += out_stride;
                    row += row_stride;
                    step = next;
                }
            } else {
            for (int step = w0 + grp; step < s1; step += NG) {
                float2 v[H];
                gen(v, row);
#pragma unroll
                for (int s = 0; s < NSEC; ++s) {
                    SecPar ps;
                    if (NSEC == 1) ps = p0; else load_par(s, ps);
                    const int kind = a.sec_kind[s];
                    float2 s1v = pk1(0.0f), s2v = pk1(0.0f);
                    svf2_block<H>(kind, ps, v, s1v, s2v);
                    // end state of the whole 16-row sub-chunk from zero state: A^8 * z_a + z_b
                    const float zx = fmaf(ps.m8[0], s1v.x, fmaf(ps.m8[1], s2v.x, s1v.y));
                    const float zy = fmaf(ps.m8[2], s1v.x, fmaf(ps.m8[3], s2v.x, s2v.y));
                    zs[w * 32 + lane] = make_float2(zx, zy);
                    bar_arrive(1 + 2 * grp, (WG + 1) * 32);
                    bar_sync(2 + 2 * grp, (WG + 1) * 32);
                    correct(ps, v, si[w * 32 + lane], s1v.x, s2v.x);
                }
                if (step >= s0) store(v, row, outp);           // warm-up steps only advance the state
                outp += out_stride;
                row += row_stride;
            }
            }
        }
        __syncthreads();     // piece boundary: shared parameters and the barrier protocol restart
    }
    if (bulk && w < NW && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// k_chain_scan3: the channel-pair variant of the packed time-parallel kernel (single section)
//
// Measured on B200 (tools/probe_fill.py): a store-only kernel writing 128-byte rows of a (frames, 4096)
// block tops out at 4.57 TB/s, with 256-byte rows at 6.0 TB/s -- k_chain_scan2's 32-channel tiles sit AT that
// first ceiling.  Here a tile is 64 adjacent channels: the two lanes of every packed register are the adjacent
// channels 2 l and 2 l + 1 of the SAME rows (instead of two 8-row halves of one channel), a sub-chunk is R3 rows, and
// each worker hands the TMA engine (R3 x 64) tiles, i.e. 256-byte rows (one STS.64 per lane per row).  There is nothing to stitch; the scanner
// chains the float64 carry through the sub-chunks directly, c_{q+1} = A^8 c_q + z_q, and publishes c_q as the
// true initial state of sub-chunk q.  Workers are software-pipelined as in k_chain_scan2.
// ------------------------------------------------------------------------------------------
template <int SRC, int NG, int WG, bool FASTSINE, int R3, bool PIPE3 = true, bool F32CARRY = false>      // R3 = rows per sub-chunk: 8 or 16
__global__ void __launch_bounds__((NG * WG + 1) * 32, 1)
k_chain_scan3(const ChainDev a, int nsteps, int warm_steps, const __grid_constant__ CUtensorMap out_map, int use_tma, int rot) {
    if (a.warm_dev && warm_steps > 0) warm_steps = device_warm_steps(a, WG * R3);
    constexpr int NW = NG * WG;
    constexpr int STEP = WG * R3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage = reinterpret_cast<float*>(smem_raw);                        // [NW][R3][64] output tiles (TMA source)
    float4* zs = reinterpret_cast<float4*>(stage + NW * R3 * 64);            // [NW][32] end states of the lane's two channels
    float4* si = zs + NW * 32;                                                // [NW][32] true initial states
    double* tnb = reinterpret_cast<double*>(si + NW * 32);                    // [NW][R3] n / rate (generic oscillators)
    // per-lane constants each worker phase needs only briefly (correction: zero-input tables and recurrence; store: gain;
    // source: rotation) live in shared memory, one copy per CTA (every worker of the CTA has the same channels), and are
    // re-read per step: the kernel sits on its register limit (72 at 896 threads) and spilled inside the loop
    float2* cst = reinterpret_cast<float2*>(tnb + NW * R3);                   // [CST_N][32]
    constexpr int CST_ZP0 = 0, CST_ZR0 = 1, CST_ZP1 = 2, CST_ZR1 = 3, CST_AL = 4, CST_BE = 5, CST_GAIN = 6, CST_ROTC = 7, CST_ROTS = 8,
                  CST_G = 9, CST_NC = 10, CST_D = 11, CST_GD = 12, CST_GD2 = 13, CST_G2 = 14;       // the section itself

    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const size_t C = (size_t)a.C;
    const bool bulk = use_tma != 0;
    const double rate = (double)a.rate;

    const int tiles = (a.C + 63) / 64;
    const long long total = (long long)tiles * nsteps;
    const long long quota = (total + gridDim.x - 1) / gridDim.x;
    long long lin = (long long)blockIdx.x * quota;
    const long long lin_end = min(total, lin + quota);

    while (lin < lin_end) {
        const int tile_idx = (int)(lin / nsteps);
        const int s0 = (int)(lin - (long long)tile_idx * nsteps);
        const int s1 = (int)min((long long)nsteps, s0 + (lin_end - lin));
        const int w0 = max(0, s0 - warm_steps);
        lin += s1 - s0;

        const int cA = tile_idx * 64 + 2 * lane, cB = cA + 1;       // a lane's packed registers hold two ADJACENT channels
        const bool liveA = cA < a.C, liveB = cB < a.C;
        const int ccA = liveA ? cA : a.C - 1, ccB = liveB ? cB : a.C - 1;

        if (w == 0) {
            cst[CST_ZP0 * 32 + lane] = pk(a.ztab[((size_t)0 * 2 + 0) * C + ccA], a.ztab[((size_t)0 * 2 + 0) * C + ccB]);
            cst[CST_ZR0 * 32 + lane] = pk(a.ztab[((size_t)0 * 2 + 1) * C + ccA], a.ztab[((size_t)0 * 2 + 1) * C + ccB]);
            cst[CST_ZP1 * 32 + lane] = pk(a.hrec[(size_t)2 * C + ccA], a.hrec[(size_t)2 * C + ccB]);
            cst[CST_ZR1 * 32 + lane] = pk(a.hrec[(size_t)3 * C + ccA], a.hrec[(size_t)3 * C + ccB]);
            cst[CST_AL * 32 + lane] = pk(a.hrec[(size_t)0 * C + ccA], a.hrec[(size_t)0 * C + ccB]);       // det(A)
            cst[CST_BE * 32 + lane] = pk(a.hrec[(size_t)1 * C + ccA], a.hrec[(size_t)1 * C + ccB]);       // tr(A) - 1 - det(A)
            cst[CST_GAIN * 32 + lane] = a.gain ? pk(a.gain[ccA], a.gain[ccB]) : pk1(1.0f);
            if (FASTSINE && rot) {
                const float2 ra = a.rot1[ccA], rb = a.rot1[ccB];
                cst[CST_ROTC * 32 + lane] = pk(ra.x, rb.x);
                cst[CST_ROTS * 32 + lane] = pk(ra.y, rb.y);
            }
            const float gA = a.coef[(size_t)0 * C + ccA], gB = a.coef[(size_t)0 * C + ccB];
            const float dA = a.coef[(size_t)2 * C + ccA], dB = a.coef[(size_t)2 * C + ccB];
            cst[CST_G * 32 + lane] = pk(gA, gB);
            cst[CST_NC * 32 + lane] = pk(-a.coef[(size_t)1 * C + ccA], -a.coef[(size_t)1 * C + ccB]);
            cst[CST_D * 32 + lane] = pk(dA, dB);
            cst[CST_GD * 32 + lane] = pk(gA * dA, gB * dB);
            cst[CST_GD2 * 32 + lane] = pk(2.0f * (gA * dA), 2.0f * (gB * dB));
            cst[CST_G2 * 32 + lane] = pk(2.0f * gA, 2.0f * gB);
        }
        __syncthreads();
        const unsigned cst_addr = smem_u32(cst + lane);
        auto cst_ld = [&](int which) {                       // volatile: not to be hoisted back into registers
            float2 v2;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v2.x), "=f"(v2.y) : "r"(cst_addr + which * 256u));
            return v2;
        };

        if (w == NW) {
            // ---------------- scanner warp: lane = channels (2 l, 2 l + 1) of the tile, float64 carry chain ----------------
            double mA[4], mB[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                mA[k] = (R3 == SIGB_SCAN_L ? a.apow : a.apow_h)[(size_t)k * C + ccA];
                mB[k] = (R3 == SIGB_SCAN_L ? a.apow : a.apow_h)[(size_t)k * C + ccB];
            }
            double a1 = w0 == 0 ? a.state[(size_t)0 * C + ccA] : 0.0, a2 = w0 == 0 ? a.state[(size_t)1 * C + ccA] : 0.0;
            double b1 = w0 == 0 ? a.state[(size_t)0 * C + ccB] : 0.0, b2 = w0 == 0 ? a.state[(size_t)1 * C + ccB] : 0.0;
            if (F32CARRY) {
                // float32 chain, both channels packed: c_{q+1} = A c_q + z_q is contractive and rounds once per 16 rows,
                // i.e. far less often than the workers' own float32 recurrence (4 FFMA2 per sub-chunk instead of 8 DFMA)
                const float2 m0 = pk((float)mA[0], (float)mB[0]), m1 = pk((float)mA[1], (float)mB[1]);
                const float2 m2 = pk((float)mA[2], (float)mB[2]), m3 = pk((float)mA[3], (float)mB[3]);
                float2 c1 = pk((float)a1, (float)b1), c2 = pk((float)a2, (float)b2);
                for (int step = w0; step < s1; ++step) {
                    const int grp = (step - w0) % NG;
                    bar_sync(1 + 2 * grp, (WG + 1) * 32);
                    float4 z[WG];
#pragma unroll
                    for (int q = 0; q < WG; ++q) z[q] = zs[(grp * WG + q) * 32 + lane];
#pragma unroll
                    for (int q = 0; q < WG; ++q) {
                        si[(grp * WG + q) * 32 + lane] = make_float4(c1.x, c1.y, c2.x, c2.y);
                        const float2 n1 = __ffma2_rn(m0, c1, __ffma2_rn(m1, c2, pk(z[q].x, z[q].y)));
                        const float2 n2 = __ffma2_rn(m2, c1, __ffma2_rn(m3, c2, pk(z[q].z, z[q].w)));
                        c1 = n1;
                        c2 = n2;
                    }
                    bar_arrive(2 + 2 * grp, (WG + 1) * 32);
                }
                a1 = c1.x; b1 = c1.y; a2 = c2.x; b2 = c2.y;
            } else {
            for (int step = w0; step < s1; ++step) {
                const int grp = (step - w0) % NG;
                bar_sync(1 + 2 * grp, (WG + 1) * 32);
                float4 z[WG];
#pragma unroll
                for (int q = 0; q < WG; ++q) z[q] = zs[(grp * WG + q) * 32 + lane];
#pragma unroll
                for (int q = 0; q < WG; ++q) {
                    si[(grp * WG + q) * 32 + lane] = make_float4((float)a1, (float)b1, (float)a2, (float)b2);
                    const double na1 = fma(mA[0], a1, fma(mA[1], a2, (double)z[q].x));
                    const double na2 = fma(mA[2], a1, fma(mA[3], a2, (double)z[q].z));
                    const double nb1 = fma(mB[0], b1, fma(mB[1], b2, (double)z[q].y));
                    const double nb2 = fma(mB[2], b1, fma(mB[3], b2, (double)z[q].w));
                    a1 = na1; a2 = na2; b1 = nb1; b2 = nb2;
                }
                bar_arrive(2 + 2 * grp, (WG + 1) * 32);
            }
            }
            if (s1 == nsteps) {       // the piece that finishes a tile hands its state to the next launch
                if (liveA) { a.state_out[(size_t)0 * C + cA] = a1; a.state_out[(size_t)1 * C + cA] = a2; }
                if (liveB) { a.state_out[(size_t)0 * C + cB] = b1; a.state_out[(size_t)1 * C + cB] = b2; }
            }
        } else {
            // ---------------- worker warps: lane = channels (2 l, 2 l + 1) of the tile, warp = 8-row sub-chunk ----------------
            const int grp = w / WG, q = w % WG;
            int64_t row = ((int64_t)w0 + grp) * STEP + (int64_t)q * R3;
            const int64_t row_stride = (int64_t)NG * STEP;
            float* outp = a.out + row * a.ld_out + cA;
            const int64_t out_stride = row_stride * a.ld_out;
            float* tile = stage + w * (R3 * 64);
            const int kind = a.sec_kind[0];

            // the section's coefficients (both lanes of every field: the lane's two adjacent channels) are read from
            // shared memory where the section runs, so that they occupy registers only there
            auto section = [&](float2 (&vv)[R3], float2& t1, float2& t2) {
                SecPar ps;
                ps.g = cst_ld(CST_G); ps.nc = cst_ld(CST_NC); ps.gd = cst_ld(CST_GD); ps.gd2 = cst_ld(CST_GD2); ps.g2 = cst_ld(CST_G2);
                ps.d = (kind & (SEC_HP | SEC_FIRST_ORDER)) ? cst_ld(CST_D) : pk1(0.0f);
                svf2_block<R3>(kind, ps, vv, t1, t2);
            };

            unsigned long long thA = 0, thB = 0, stepA = 0, stepB = 0;
            int dhiA = 0, dhiB = 0;
            double hzA = 0.0, hzB = 0.0, phA = 0.0, phB = 0.0;
            float2 cv = pk1(0.0f);
            if (SRC == SRC_OSC) {
                if (FASTSINE) {
                    const unsigned long long dA = a.dtheta[ccA], dB = a.dtheta[ccB];
                    thA = a.theta0[ccA] + (unsigned long long)(a.position + row) * dA;
                    thB = a.theta0[ccB] + (unsigned long long)(a.position + row) * dB;
                    stepA = dA * (unsigned long long)row_stride;
                    stepB = dB * (unsigned long long)row_stride;
                    dhiA = (int)((dA + 0x80000000ull) >> 32);
                    dhiB = (int)((dB + 0x80000000ull) >> 32);
                } else {
                    hzA = a.hertz[ccA]; hzB = a.hertz[ccB];
                    phA = a.phase[ccA]; phB = a.phase[ccB];
                }
            }
            if (SRC == SRC_CONST) cv = pk(a.constv[ccA], a.constv[ccB]);

            auto gen = [&](float2 (&v)[R3], int64_t grow) {
                if (SRC == SRC_OSC) {
                    if (FASTSINE) {
                        int ha = (int)(thA >> 32), hb = (int)(thB >> 32);
                        if (rot) {
                            // k_bank's two-pipe evaluation: rows 1, 5, 9, 13 of the sub-chunk get a sine AND a
                            // cosine from the SFU, their neighbours (one row back, two rows forward) the angle-addition
                            // rotation by the channel's one-row phase advance (cos, sin tabulated in float64 on the host):
                            // 16 instructions per 4 rows x 2 channels instead of 36
                            const float2 rotC = cst_ld(CST_ROTC), rotS = cst_ld(CST_ROTS);
                            const float2 NS = pk(-rotS.x, -rotS.y);
#pragma unroll
                            for (int q4 = 0; q4 + 3 < R3; q4 += 4) {
                                const float2 r = __fmul2_rn(pk((float)(ha + dhiA), (float)(hb + dhiB)), pk1(1.4629180792671596e-9f));
                                const float2 S1 = pk(__sinf(r.x), __sinf(r.y)), C1 = pk(__cosf(r.x), __cosf(r.y));
                                const float2 t = __fmul2_rn(S1, rotC);
                                v[q4 + 0] = __ffma2_rn(C1, NS, t);
                                v[q4 + 1] = S1;
                                const float2 S2 = __ffma2_rn(C1, rotS, t);
                                const float2 C2 = __ffma2_rn(S1, NS, __fmul2_rn(C1, rotC));
                                v[q4 + 2] = S2;
                                v[q4 + 3] = __ffma2_rn(C2, rotS, __fmul2_rn(S2, rotC));
                                ha += 4 * dhiA;
                                hb += 4 * dhiB;
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < R3; ++k) {
                                const float2 r = __fmul2_rn(pk((float)ha, (float)hb), pk1(1.4629180792671596e-9f));
                                v[k] = pk(__sinf(r.x), __sinf(r.y));
                                ha += dhiA;
                                hb += dhiB;
                            }
                        }
                        thA += stepA;
                        thB += stepB;
                    } else {
                        if (lane < R3) tnb[w * R3 + lane] = __ddiv_rn((double)(a.position + grow + lane), rate);
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < R3; ++k)
                            v[k] = pk(osc_wave(a.wave, osc_cycles(tnb[w * R3 + k], hzA, phA)),
                                      osc_wave(a.wave, osc_cycles(tnb[w * R3 + k], hzB, phB)));
                        __syncwarp();
                    }
                } else if (SRC == SRC_BUF) {
#pragma unroll
                    for (int k = 0; k < R3; ++k)
                        v[k] = pk(liveA ? load_src(a, grow + k, cA) : 0.0f, liveB ? load_src(a, grow + k, cB) : 0.0f);
                } else {
#pragma unroll
                    for (int k = 0; k < R3; ++k) v[k] = cv;
                }
            };

            int step = w0 + grp;
            float2 vn[PIPE3 ? R3 : 1];
            float2 s1n = pk1(0.0f), s2n = pk1(0.0f);
            if (PIPE3 && step < s1) {
                gen(reinterpret_cast<float2(&)[R3]>(vn), row);
                section(reinterpret_cast<float2(&)[R3]>(vn), s1n, s2n);
            }
            while (step < s1) {
                float2 v[R3];
                const int next = step + NG;
                if (PIPE3) {
                    zs[w * 32 + lane] = make_float4(s1n.x, s1n.y, s2n.x, s2n.y);       // (s1 of both channels, s2 of both channels)
                    bar_arrive(1 + 2 * grp, (WG + 1) * 32);
#pragma unroll
                    for (int k = 0; k < R3; ++k) v[k] = vn[PIPE3 ? k : 0];
                    if (next < s1) {      // the next step's zero-state render hides the scanner's turn-around
                        s1n = pk1(0.0f);
                        s2n = pk1(0.0f);
                        gen(reinterpret_cast<float2(&)[R3]>(vn), row + row_stride);
                        section(reinterpret_cast<float2(&)[R3]>(vn), s1n, s2n);
                    }
                } else {
                    s1n = pk1(0.0f);
                    s2n = pk1(0.0f);
                    gen(v, row);
                    section(v, s1n, s2n);
                    zs[w * 32 + lane] = make_float4(s1n.x, s1n.y, s2n.x, s2n.y);       // (s1 of both channels, s2 of both channels)
                    bar_arrive(1 + 2 * grp, (WG + 1) * 32);
                }
                bar_sync(2 + 2 * grp, (WG + 1) * 32);
                const float4 ia = si[w * 32 + lane];
                {   // zero-input response of the true initial state, advanced by its 2-term recurrence in delta form
                    // (dh[k] = det dh[k-1] + (tr - 1 - det) h[k-1], h[k] = h[k-1] + dh[k]: see k_chain_scan2's `correct`)
                    const float2 i1 = pk(ia.x, ia.y), i2 = pk(ia.z, ia.w);
                    float2 h = __ffma2_rn(cst_ld(CST_ZP0), i1, __fmul2_rn(cst_ld(CST_ZR0), i2));
                    float2 dh = __ffma2_rn(cst_ld(CST_ZP1), i1, __fmul2_rn(cst_ld(CST_ZR1), i2));
                    const float2 al = cst_ld(CST_AL), be = cst_ld(CST_BE);
                    v[0] = __fadd2_rn(v[0], h);
#pragma unroll
                    for (int k = 1; k < R3; ++k) {
                        if (k > 1) dh = __ffma2_rn(al, dh, __fmul2_rn(be, h));
                        h = __fadd2_rn(h, dh);
                        v[k] = __fadd2_rn(v[k], h);
                    }
                }
                if (step >= s0) {
                    const float2 gain2 = cst_ld(CST_GAIN);
                    if (bulk) {
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < R3; ++k) {
                            const float2 o = __fmul2_rn(v[k], gain2);
                            reinterpret_cast<float2*>(tile)[k * 32 + lane] = o;        // one STS.64 per row
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_tile(&out_map, tile_idx * 64, (int)row, tile);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < R3; ++k) {
                            const float2 o = __fmul2_rn(v[k], gain2);
                            if (liveA) __stcs(outp + (int64_t)k * a.ld_out, o.x);
                            if (liveB) __stcs(outp + (int64_t)k * a.ld_out + 1, o.y);
                        }
                    }
                }
                outp += out_stride;
                row += row_stride;
                step = next;
            }
        }
        __syncthreads();     // piece boundary: the barrier protocol restarts
    }
    if (bulk && w < NW && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// k_ewise / k_reduce
// ------------------------------------------------------------------------------------------
// Amp (fx.py:60): copysign(input ** exp, input).  Small integer exponents -- the usual waveshaping settings -- as products
// (one or two roundings instead of powf's ~4 ulp and ~30 instructions); everything else, NaN for a negative input under a
// fractional exponent included, through powf.
__device__ __forceinline__ float ew_amp(float x, float p) {
    if (p == 2.0f) return x * fabsf(x);
    if (p == 1.0f) return x;
    if (p == 3.0f) return x * x * x;
    if (p == 4.0f) { const float x2 = x * x; return copysignf(x2 * x2, x); }
    return copysignf(powf(x, p), x);
}

__device__ __forceinline__ float ew_load(const float* p, int64_t ld, int cs, int64_t rows, int64_t r, int c) {
    if (rows >= 0 && r >= rows) return 0.0f;
    return __ldg(p + r * ld + (int64_t)c * cs);
}

__global__ void __launch_bounds__(256) k_ewise(const EwiseDev a) {
    const int64_t total = (int64_t)a.frames * a.C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / a.C;
        const int c = (int)(i - r * a.C);
        const float x = ew_load(a.a, a.lda, a.acs, a.a_rows, r, c);
        float y;
        switch (a.op) {
            case EW_COPY: y = x; break;
            case EW_GAIN: y = x * a.p[c]; break;                                   // fx.py:52
            case EW_MIX: {                                                          // fx.py:40
                const float m = a.p[c];
                y = m * x + (1.0f - m) * ew_load(a.b, a.ldb, a.bcs, a.b_rows, r, c);
                break;
            }
            case EW_RINGMOD: y = x * ew_load(a.b, a.ldb, a.bcs, a.b_rows, r, c); break;   // fx.py:46
            default: y = ew_amp(x, a.p[c]); break;                                  // fx.py:60
        }
        a.out[r * a.ld_out + c] = y;
    }
}

// Vector variant: every operand is a full-width block whose rows are 16-byte aligned, so a thread handles four
// adjacent channels with 128-bit loads and one 128-bit streaming store, and rows are walked without a division.
__device__ __forceinline__ float ew_apply(int op, float x, float b, float p) {
    switch (op) {
        case EW_COPY: return x;
        case EW_GAIN: return x * p;
        case EW_MIX: return p * x + (1.0f - p) * b;
        case EW_RINGMOD: return x * b;
        default: return ew_amp(x, p);
    }
}

__global__ void __launch_bounds__(256) k_ewise_v4(const EwiseDev a, int c4, int rows_per_block) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;            // group of four channels
    if (q >= c4) return;
    const bool has_b = a.op == EW_MIX || a.op == EW_RINGMOD;
    float4 p = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (a.p) p = __ldg(reinterpret_cast<const float4*>(a.p) + q);
    const int r0 = blockIdx.y * rows_per_block, r1 = min(a.frames, r0 + rows_per_block);
    for (int r = r0; r < r1; ++r) {
        const bool in_a = a.a_rows < 0 || r < a.a_rows, in_b = a.b_rows < 0 || r < a.b_rows;
        const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const float4 x = in_a ? __ldcs(reinterpret_cast<const float4*>(a.a + (int64_t)r * a.lda) + q) : zero;
        const float4 b = has_b && in_b ? __ldcs(reinterpret_cast<const float4*>(a.b + (int64_t)r * a.ldb) + q) : zero;
        float4 y;
        y.x = ew_apply(a.op, x.x, b.x, p.x);
        y.y = ew_apply(a.op, x.y, b.y, p.y);
        y.z = ew_apply(a.op, x.z, b.z, p.z);
        y.w = ew_apply(a.op, x.w, b.w, p.w);
        __stcs(reinterpret_cast<float4*>(a.out + (int64_t)r * a.ld_out) + q, y);
    }
}

// one warp per (row, group): float32 lane partials, float64 cross-lane tree (fixed order)
__global__ void __launch_bounds__(256) k_reduce(const ReduceDev a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int ngroups = a.pan ? 1 : a.groups;
    const int per = a.C / ngroups;
    for (int64_t item = warp; item < (int64_t)a.frames * ngroups; item += nwarps) {
        const int64_t r = item / ngroups;
        const int gidx = (int)(item - r * ngroups);
        double acc0 = 0.0, acc1 = 0.0;
        // whole 1,024-channel chunks of contiguous, 16-byte aligned rows: eight independent 128-bit loads per lane in flight
        // (the scalar loop below keeps one 4-byte load per lane in flight and measured 1 TB/s); 32 float32 terms per lane, then float64
        const bool vec = a.ics == 1 && (per & 1023) == 0 && (a.ld_in & 3) == 0 && (reinterpret_cast<uintptr_t>(a.in) & 15) == 0 &&
                         (!a.pan || (reinterpret_cast<uintptr_t>(a.w) & 15) == 0) && (a.in_rows < 0 || r < a.in_rows);
        if (vec) {
            const float4* rowp = reinterpret_cast<const float4*>(a.in + r * a.ld_in + (int64_t)gidx * per);
            const float4* wp = a.pan ? reinterpret_cast<const float4*>(a.w + (int64_t)gidx * per) : nullptr;
            for (int q0 = 0; q0 < per / 4; q0 += 256) {
                float4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = __ldcs(rowp + q0 + lane + 32 * k);
                float p0 = 0.0f, p1 = 0.0f;
                if (a.pan) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 w = __ldg(wp + q0 + lane + 32 * k);
                        p0 = fmaf(v[k].x, 1.0f - w.x, p0); p1 = fmaf(v[k].x, w.x, p1);
                        p0 = fmaf(v[k].y, 1.0f - w.y, p0); p1 = fmaf(v[k].y, w.y, p1);
                        p0 = fmaf(v[k].z, 1.0f - w.z, p0); p1 = fmaf(v[k].z, w.z, p1);
                        p0 = fmaf(v[k].w, 1.0f - w.w, p0); p1 = fmaf(v[k].w, w.w, p1);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) p0 += (v[k].x + v[k].y) + (v[k].z + v[k].w);
                }
                acc0 += (double)p0;
                acc1 += (double)p1;
            }
        } else
        for (int j0 = 0; j0 < per; j0 += 32 * 32) {
            float p0 = 0.0f, p1 = 0.0f;                 // <=32 terms in float32, then float64
            for (int j = j0 + lane; j < min(per, j0 + 32 * 32); j += 32) {
                const int c = gidx * per + j;
                const float x = ew_load(a.in, a.ld_in, a.ics, a.in_rows, r, c);
                if (a.pan) {
                    const float pn = a.w[c];
                    p0 = fmaf(x, 1.0f - pn, p0);
                    p1 = fmaf(x, pn, p1);
                } else {
                    p0 += x;
                }
            }
            acc0 += (double)p0;
            acc1 += (double)p1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
            acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
        }
        if (lane == 0) {
            if (a.pan) {
                a.out[r * a.ld_out + 0] = (float)acc0;
                a.out[r * a.ld_out + 1] = (float)acc1;
            } else {
                a.out[r * a.ld_out + gidx] = (float)acc0;
            }
        }
    }
}

__global__ void k_probe_sin(const double* r, int n, float* out, int variant) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = (float)r[i];
    out[i] = variant == 0 ? sin2pi<0>(x) : (variant == 1 ? sin2pi<1>(x) : sin2pi<2>(x));
}

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
template <int SRC>
cudaError_t launch_seq_src(const ChainDev& a, cudaStream_t st) {
    dim3 block(128);
    dim3 grid((a.C + 127) / 128, 1);
    int rows_per_seg = a.frames;
    int nsec = a.nsec;
    if (nsec == 0) {
        // stateless: tile time so the grid fills the machine
        int want = (148 * 16 + (int)grid.x - 1) / (int)grid.x;
        int segs = max(1, min(want, (a.frames + 63) / 64));
        segs = min(segs, 65535);
        rows_per_seg = (a.frames + segs - 1) / segs;
        grid.y = (a.frames + rows_per_seg - 1) / rows_per_seg;
    }
    // few channels: a warp per channel keeps the source off the recurrence's critical path (k_chain_rt)
    if (nsec >= 1 && a.epi_op == 0 && (int64_t)a.C * 32 <= (int64_t)sm_count() * 2048) {
        dim3 rgrid((a.C * 32 + 127) / 128);
#define RT_CASE(N) k_chain_rt<SRC, N><<<rgrid, block, 0, st>>>(a)
        if (nsec == 1) RT_CASE(1);
        else if (nsec == 2) RT_CASE(2);
        else if (nsec <= 4) RT_CASE(4);
        else if (nsec <= 8) RT_CASE(8);
        else RT_CASE(16);
#undef RT_CASE
        return cudaGetLastError();
    }
#define SEQ_CASE(N) k_chain_seq<SRC, N><<<grid, block, 0, st>>>(a, rows_per_seg)
    if (nsec == 0) SEQ_CASE(0);
    else if (nsec == 1) SEQ_CASE(1);
    else if (nsec == 2) SEQ_CASE(2);
    else if (nsec <= 4) SEQ_CASE(4);
    else if (nsec <= 8) SEQ_CASE(8);
    else SEQ_CASE(16);
#undef SEQ_CASE
    return cudaGetLastError();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cudaGetLastError();
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// 2-D float32 tensor map over the output block: dim0 = channels (contiguous), dim1 = rows; box 32 x L
bool make_out_map(const ChainDev& a, int rows, CUtensorMap* map) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(a.out) & 15) != 0 || ((a.ld_out * 4) & 15) != 0 || rows <= 0) return false;
    cuuint64_t dims[2] = {(cuuint64_t)a.C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)a.ld_out * 4};
    cuuint32_t box[2] = {32u, (cuuint32_t)L};
    cuuint32_t estr[2] = {1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int SRC, int NSEC, int NG, int WG, bool FASTSINE, int PIPE = 0>
cudaError_t launch_scan2_t(const ChainDev& a, cudaStream_t st, int* rows_done) {
    constexpr int NW = NG * WG;
    constexpr int STEP = WG * L;
    const int nsteps = a.frames / STEP;
    *rows_done = nsteps * STEP;
    if (nsteps == 0) return cudaSuccess;
    size_t smem = (size_t)NW * L * 32 * sizeof(float) + (size_t)NW * 32 * sizeof(float2) * 2 +
                  (size_t)NSEC * 4 * 32 * sizeof(float4) + (size_t)NW * L * sizeof(double) +
                  (NSEC > 2 ? (size_t)NSEC * 4 * 32 * sizeof(double) + (size_t)NSEC * WG * 32 * sizeof(float4) : 0);
    auto kern = k_chain_scan2<SRC, NSEC, NG, WG, FASTSINE, PIPE>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    const int use_tma = g_scan_tma && make_out_map(a, nsteps * STEP, &map) ? 1 : 0;
    // one persistent CTA per SM, each taking an equal share of the (tile, step) work, when a mid-tile
    // start costs little warm-up; otherwise one CTA per tile
    const int tiles = (a.C + 31) / 32;
    const int sms = sm_count();
    int grid_x = tiles, warm_steps = 0;
    const int warm_host = a.warm_dev ? a.warm_est : a.warm_rows;      // modulated cutoffs: last request's horizon as estimate
    if (g_scan_split && warm_host >= 0) {
        const int ws = std::max(1, (warm_host + STEP - 1) / STEP);
        const long long total = (long long)tiles * nsteps;
        const long long quota = (total + sms - 1) / sms;
        if (tiles % sms != 0 && quota >= 8ll * ws && quota >= 4) {
            grid_x = (int)((total + quota - 1) / quota);
            warm_steps = ws;
        }
    }
    dim3 grid(grid_x), block((NW + 1) * 32);
    kern<<<grid, block, smem, st>>>(a, nsteps, warm_steps, map, use_tma);
    return cudaGetLastError();
}

template <int NSEC, int NG, int WG, int PIPE = 0>
cudaError_t launch_scan2_n(const ChainDev& a, cudaStream_t st, int* rows_done) {
    const bool fast = a.src_kind == SRC_OSC && a.wave == SIGB_WAVE_SINE && a.theta0 != nullptr;
    switch (a.src_kind) {
        case SRC_OSC:
            return fast ? launch_scan2_t<SRC_OSC, NSEC, NG, WG, true, PIPE>(a, st, rows_done)
                        : launch_scan2_t<SRC_OSC, NSEC, NG, WG, false, PIPE>(a, st, rows_done);
        case SRC_BUF: return launch_scan2_t<SRC_BUF, NSEC, NG, WG, false, PIPE>(a, st, rows_done);
        default: return launch_scan2_t<SRC_CONST, NSEC, NG, WG, false, PIPE>(a, st, rows_done);
    }
}

template <int SRC, int NG, int WG, bool FASTSINE, int R3, bool PIPE3, bool F32CARRY>
cudaError_t launch_scan3_t(const ChainDev& a, cudaStream_t st, int* rows_done) {
    constexpr int NW = NG * WG;
    constexpr int STEP = WG * R3;
    const int nsteps = a.frames / STEP;
    *rows_done = nsteps * STEP;
    if (nsteps == 0) return cudaSuccess;
    const size_t smem = (size_t)NW * R3 * 64 * sizeof(float) + (size_t)NW * 32 * sizeof(float4) * 2 + (size_t)NW * R3 * sizeof(double) +
                        (size_t)15 * 32 * sizeof(float2);
    auto kern = k_chain_scan3<SRC, NG, WG, FASTSINE, R3, PIPE3, F32CARRY>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    // 2-D tensor map with a 64 x 8 box: 256-byte rows per TMA store
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    int use_tma = 0;
    EncodeTiledFn enc = encode_tiled_fn();
    if (g_scan_tma && enc && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 && ((a.ld_out * 4) & 15) == 0) {
        cuuint64_t dims[2] = {(cuuint64_t)a.C, (cuuint64_t)(nsteps * STEP)};
        cuuint64_t strides[1] = {(cuuint64_t)a.ld_out * 4};
        cuuint32_t box[2] = {64u, (cuuint32_t)R3};
        cuuint32_t estr[2] = {1u, 1u};
        use_tma = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    const int tiles = (a.C + 63) / 64;
    const int sms = sm_count();
    int grid_x = tiles, warm_steps = 0;
    const int warm_host = a.warm_dev ? a.warm_est : a.warm_rows;      // modulated cutoffs: last request's horizon as estimate
    if (g_scan_split && warm_host >= 0) {
        const int ws = std::max(1, (warm_host + STEP - 1) / STEP);
        const long long total = (long long)tiles * nsteps;
        const long long quota = (total + sms - 1) / sms;
        if (tiles % sms != 0 && quota >= 8ll * ws && quota >= 4) {
            grid_x = (int)((total + quota - 1) / quota);
            warm_steps = ws;
        }
    }
    dim3 grid(grid_x), block((NW + 1) * 32);
    kern<<<grid, block, smem, st>>>(a, nsteps, warm_steps, map, use_tma, (FASTSINE && g_scan_rot && a.rot1 != nullptr) ? 1 : 0);
    return cudaGetLastError();
}

template <int NG, int WG, int R3, bool PIPE3 = true, bool F32CARRY = false>
cudaError_t launch_scan3_n(const ChainDev& a, cudaStream_t st, int* rows_done) {
    const bool fast = a.src_kind == SRC_OSC && a.wave == SIGB_WAVE_SINE && a.theta0 != nullptr;
    switch (a.src_kind) {
        case SRC_OSC:
            return fast ? launch_scan3_t<SRC_OSC, NG, WG, true, R3, PIPE3, F32CARRY>(a, st, rows_done) : launch_scan3_t<SRC_OSC, NG, WG, false, R3, PIPE3, F32CARRY>(a, st, rows_done);
        case SRC_BUF: return launch_scan3_t<SRC_BUF, NG, WG, false, R3, PIPE3, F32CARRY>(a, st, rows_done);
        default: return launch_scan3_t<SRC_CONST, NG, WG, false, R3, PIPE3, F32CARRY>(a, st, rows_done);
    }
}

}  // namespace

// Scan variants (A/B; every one computes the same transfer function):
//   9   k_chain_scan2  32-channel tiles, packed two half sub-chunks per thread (the kernel for 2 sections, for deep
//                      cascades forced onto the scan, and for chains of at most 32 channels)
//   16  k_chain_scan3  64-channel tiles, float64 carry chain in the scanner      (single section)
//   18  k_chain_scan3  64-channel tiles, packed float32 carry chain (default)    (single section)
// Other values fall to the nearest of these (geometries measured in round 1 and dropped: DESIGN section 9).
static int scan_normalise(int nsec, int C, int variant) {
    if (nsec != 1 || C <= 32) return 9;          // one 32-channel tile: the 64-channel kernel would idle half its lanes
    return variant == 16 ? 16 : variant >= 13 ? 18 : 9;
}

// scan geometry: (groups, worker warps per group).  Deep cascades keep the block at 512 threads
// so the scanner's fp64 transition matrices stay in registers.
static void scan_geometry(int nsec, int variant, int* ng, int* wg) {
    if (nsec > 2) { *ng = 4; *wg = 7; return; }
    if (nsec == 2) { *ng = 5; *wg = 6; return; }
    if (variant >= 13) { *ng = 3; *wg = 9; return; }
    *ng = 3; *wg = 8;
}

extern "C" int sigb_scan_rows_per_step(int nsec, int variant) {
    int ng, wg;
    scan_geometry(nsec, scan_normalise(nsec, 64, variant), &ng, &wg);
    return wg * L;
}

extern "C" void sigb_set_scan_tma(int on) { g_scan_tma = on; }
extern "C" void sigb_set_scan_split(int on) { g_scan_split = on; }
extern "C" void sigb_set_scan_rot(int on) { g_scan_rot = on; }

extern "C" int sigb_launch_chain_seq(const ChainDev* a, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (a->frames <= 0 || a->C <= 0) return 0;
    switch (a->src_kind) {
        case SRC_OSC: return (int)launch_seq_src<SRC_OSC>(*a, st);
        case SRC_BUF: return (int)launch_seq_src<SRC_BUF>(*a, st);
        default: return (int)launch_seq_src<SRC_CONST>(*a, st);
    }
}

// Renders the leading whole steps of `a` with the time-parallel kernel; *rows_done tells the
// caller how many rows were covered (the tail goes through k_chain_seq).
extern "C" int sigb_launch_chain_scan(const ChainDev* a, int variant, void* stream, int* rows_done) {
    cudaStream_t st = (cudaStream_t)stream;
    *rows_done = 0;
    if (a->frames <= 0 || a->C <= 0 || a->nsec < 1 || a->nsec > 8) return 0;
    variant = scan_normalise(a->nsec, a->C, variant);
    if (a->nsec > 4) return (int)launch_scan2_n<8, 4, 7>(*a, st, rows_done);
    if (a->nsec > 2) return (int)launch_scan2_n<4, 4, 7>(*a, st, rows_done);
    if (a->nsec == 2) return (int)launch_scan2_n<2, 5, 6>(*a, st, rows_done);
    if (variant == 16) return (int)launch_scan3_n<3, 9, 16, false>(*a, st, rows_done);
    if (variant == 18) return (int)launch_scan3_n<3, 9, 16, false, true>(*a, st, rows_done);
    return (int)launch_scan2_n<1, 3, 8, 1>(*a, st, rows_done);       // single section: software-pipelined workers
}

extern "C" int sigb_launch_ewise(const EwiseDev* a, void* stream) {
    if (a->frames <= 0 || a->C <= 0) return 0;
    auto aligned = [](const void* ptr, int64_t ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld & 3) == 0; };
    const bool need_b = a->op == EW_MIX || a->op == EW_RINGMOD;
    const bool vec = (a->C & 3) == 0 && aligned(a->out, a->ld_out) && a->acs == 1 && aligned(a->a, a->lda) &&
                     (!need_b || (a->bcs == 1 && aligned(a->b, a->ldb))) && (!a->p || (reinterpret_cast<uintptr_t>(a->p) & 15) == 0);
    if (vec) {
        const int c4 = a->C / 4;
        dim3 grid((c4 + 255) / 256, 1);
        const int want = (148 * 8 + (int)grid.x - 1) / (int)grid.x;               // row blocks to fill the machine
        const int rows_per_block = max(8, (a->frames + want - 1) / want);
        grid.y = (a->frames + rows_per_block - 1) / rows_per_block;
        k_ewise_v4<<<grid, 256, 0, (cudaStream_t)stream>>>(*a, c4, rows_per_block);
        return (int)cudaGetLastError();
    }
    const int64_t total = (int64_t)a->frames * a->C;
    int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
    k_ewise<<<blocks, 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int sigb_launch_reduce(const ReduceDev* a, void* stream) {
    if (a->frames <= 0 || a->C <= 0) return 0;
    const int64_t items = (int64_t)a->frames * (a->pan ? 1 : a->groups);
    int blocks = (int)min((int64_t)148 * 8, (items + 7) / 8);
    k_reduce<<<blocks, 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int sigb_launch_probe_sin(const double* r, int n, float* out, int variant, void* stream) {
    k_probe_sin<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(r, n, out, variant);
    return (int)cudaGetLastError();
}
