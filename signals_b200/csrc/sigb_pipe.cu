// k_cascade_pipe: section-pipelined filter cascade (sm_100a) for deep cascades on many channels
// (BASELINE config C4: 8 chained Butterworth low-pass biquads on 16,384 channels x 60 s).
//
// The time-parallel scan kernel pays, per section, a zero-input correction and a scanner hand-shake;
// for an 8-section cascade that is ~70 instructions per channel-sample and the launch is FP32-issue
// bound far below the HBM roofline.  Here the parallelism comes from the SECTIONS instead of from
// time: a CTA owns a tile of 64 adjacent channels and runs one warp per section (a lane = 2 adjacent
// channels, packed f32x2 math).  Warp s filters chunk (t - s) at iteration t and hands its output to
// warp s+1 through a double-buffered shared-memory tile, so every warp walks time sequentially with
// its section's state in registers: no scan, no correction, no recompute -- 7 packed FP32 instructions
// + one LDS.64 + one STS.64 per section per two channel-samples.
//
//   source   SRC_BUF: every thread streams the (16 x 64) input chunks of the tile into a shared-memory
//            ring with cp.async, PRE chunks ahead (enough bytes in flight per SM to cover HBM latency);
//            SRC_OSC / SRC_CONST: an extra warp renders the source chunk one iteration ahead.
//   sink     the last section's warp multiplies by the folded gain and stores 256-byte rows.
//
// Reference semantics: CritFilter._filter, /root/reference/src/signals/chain/fx.py:85-121 (per-channel
// Butterworth sections from zero state); the sections are the same zero-delay-feedback state-variable
// sections as in sigb_kernels.cu (sigb_design.h).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <type_traits>

#include "sigb200.h"
#include "sigb_internal.h"
#include "sigb_device.cuh"

namespace {

using namespace sigb_dev;

constexpr int PR = 16;          // rows per chunk
constexpr int PC = 64;          // channels per tile (2 per lane)
__host__ __device__ constexpr int pre_chunks(int spw) { return spw == 2 ? 2 : 3; }   // cp.async chunks in flight per CTA (4 KB each)
constexpr int CHUNK_FLOATS = PR * PC;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct SecReg {
    float2 nc, a2, al, g, g2;   // second order: (-c) (2 g d) (g d) (g) (2 g); first order: g = G; d kept in a2 for HP
    float2 d;                   // high-pass output scale
    float2 s1, s2;
};

// One sample (two channels) of a zero-delay-feedback state-variable section.  With e = x - c s1 - s2:
//   hp = d e;  bp = s1 + g d e;  s1' = s1 + 2 g d e;  lp = s2 + g bp;  s2' = s2 + 2 g bp
// -- six packed FP32 instructions for a low-pass section (seven for high-pass), and only two of them
// (e, s1') on the s1 -> s1' dependency chain.
template <int KIND>
__device__ __forceinline__ float2 pipe_step(float2 x, SecReg& r) {
    const float2 neg1 = make_float2(-1.0f, -1.0f);
    if (KIND & SEC_FIRST_ORDER) {
        const float2 t = __ffma2_rn(r.s1, neg1, x);
        const float2 w = __fmul2_rn(t, r.g);
        const float2 lp = __fadd2_rn(w, r.s1);
        r.s1 = __fadd2_rn(lp, w);
        return (KIND & SEC_HP) ? __ffma2_rn(lp, neg1, x) : lp;
    }
#ifdef SIGB_PIPE_SCALAR
    // two independent scalar chains (4-cycle FFMA latency each) instead of one packed chain
    float2 y;
    {
        const float xs = x.x - r.s2.x;
        const float e = fmaf(r.nc.x, r.s1.x, xs);
        const float bp = fmaf(r.al.x, e, r.s1.x);
        r.s1.x = fmaf(r.a2.x, e, r.s1.x);
        const float lp = fmaf(r.g.x, bp, r.s2.x);
        r.s2.x = fmaf(r.g2.x, bp, r.s2.x);
        y.x = (KIND & SEC_HP) ? e * r.d.x : lp;
    }
    {
        const float xs = x.y - r.s2.y;
        const float e = fmaf(r.nc.y, r.s1.y, xs);
        const float bp = fmaf(r.al.y, e, r.s1.y);
        r.s1.y = fmaf(r.a2.y, e, r.s1.y);
        const float lp = fmaf(r.g.y, bp, r.s2.y);
        r.s2.y = fmaf(r.g2.y, bp, r.s2.y);
        y.y = (KIND & SEC_HP) ? e * r.d.y : lp;
    }
    return y;
#else
    const float2 xs = __ffma2_rn(r.s2, neg1, x);          // x - s2 (off the s1 chain)
    const float2 e = __ffma2_rn(r.nc, r.s1, xs);
    const float2 bp = __ffma2_rn(r.al, e, r.s1);
    r.s1 = __ffma2_rn(r.a2, e, r.s1);
    const float2 lp = __ffma2_rn(r.g, bp, r.s2);
    r.s2 = __ffma2_rn(r.g2, bp, r.s2);
    return (KIND & SEC_HP) ? __fmul2_rn(e, r.d) : lp;
#endif
}

// one chunk through SPW consecutive sections of the same kind, in place in the shared-memory slot (a lane owns
// its two channels of every row, so there is no cross-thread hazard inside a stage).  Two sections per warp
// halve the shared-memory traffic of the pipeline (the binding resource at one section per warp) and give the
// scheduler two dependency chains to overlap.
template <int KIND, int SPW>
__device__ __forceinline__ void pipe_chunk(float* __restrict__ slot, int rows, int lane, SecReg (&r)[2]) {
    float2* __restrict__ p = reinterpret_cast<float2*>(slot) + lane;
    if (rows == PR) {
        constexpr int HR = PR / 2;                 // two half-chunks of 8 rows keep the kernel at 64 registers
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float2 x[HR];
#pragma unroll
            for (int k = 0; k < HR; ++k) x[k] = p[(h * HR + k) * (PC / 2)];
#pragma unroll
            for (int k = 0; k < HR; ++k) {
                x[k] = pipe_step<KIND>(x[k], r[0]);
                if (SPW == 2) x[k] = pipe_step<KIND>(x[k], r[1]);
            }
#pragma unroll
            for (int k = 0; k < HR; ++k) p[(h * HR + k) * (PC / 2)] = x[k];
        }
    } else {
        for (int k = 0; k < rows; ++k) {     // ragged last chunk: state must stop at the last real row
            float2 y = pipe_step<KIND>(p[k * (PC / 2)], r[0]);
            if (SPW == 2) y = pipe_step<KIND>(y, r[1]);
            p[k * (PC / 2)] = y;
        }
    }
}

template <int SPW>
__device__ __forceinline__ void pipe_chunk_kind(int kind, float* slot, int rows, int lane, SecReg (&r)[2]) {
    switch (kind) {
        case 0: pipe_chunk<0, SPW>(slot, rows, lane, r); break;
        case SEC_HP: pipe_chunk<SEC_HP, SPW>(slot, rows, lane, r); break;
        case SEC_FIRST_ORDER: pipe_chunk<SEC_FIRST_ORDER, SPW>(slot, rows, lane, r); break;
        default: pipe_chunk<SEC_FIRST_ORDER | SEC_HP, SPW>(slot, rows, lane, r); break;
    }
}

// source warp: chunk c of an oscillator / constant source for the tile's 64 channels -> smem
template <int WAVE>
__device__ __forceinline__ void pipe_source_wave(const ChainDev& a, int tile, int64_t n0, int lane, float2* dst) {
    float x[2][PR];
    bool near[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int cc = min(tile * PC + 2 * lane + h, a.C - 1);
        const unsigned long long dth = a.dtheta[cc];
        const unsigned long long th = a.theta0[cc] + (unsigned long long)n0 * dth + 0x80000000ull;
        const int w = (int)(th >> 32), dhi = (int)((dth + 0x80000000ull) >> 32);
        near[h] = gen_tile<WAVE, PR>(w, dhi, a.guard, x[h]);
    }
#pragma unroll
    for (int k = 0; k < PR; ++k) dst[k * (PC / 2) + lane] = make_float2(x[0][k], x[1][k]);
    if (WAVE != SIGB_WAVE_SINE && (near[0] || near[1])) {
        // within the guard band of a discontinuity: redo the lane's column with the reference's float64
        // arithmetic (osc.py:32), written straight to shared memory (keeps x[] in registers)
        float* col = reinterpret_cast<float*>(dst + lane);
        for (int h = 0; h < 2; ++h) {
            if (!near[h]) continue;
            const int cc = min(tile * PC + 2 * lane + h, a.C - 1);
            const double hz = a.hertz[cc], ph = a.phase[cc], rate = (double)a.rate;
            for (int k = 0; k < PR; ++k)
                col[k * PC + h] = osc_wave(WAVE, osc_cycles(__ddiv_rn((double)(n0 + k), rate), hz, ph));
        }
    }
}

__device__ __forceinline__ void pipe_source_chunk(const ChainDev& a, int tile, int c, int lane, float* dstf) {
    const int64_t n0 = a.position + (int64_t)c * PR;
    float2* dst = reinterpret_cast<float2*>(dstf);
    if (a.src_kind == SRC_CONST) {
        const float2 v = make_float2(a.constv[min(tile * PC + 2 * lane, a.C - 1)], a.constv[min(tile * PC + 2 * lane + 1, a.C - 1)]);
#pragma unroll
        for (int k = 0; k < PR; ++k) dst[k * (PC / 2) + lane] = v;
        return;
    }
    switch (a.wave) {
        case SIGB_WAVE_SINE: pipe_source_wave<SIGB_WAVE_SINE>(a, tile, n0, lane, dst); break;
        case SIGB_WAVE_SQUARE: pipe_source_wave<SIGB_WAVE_SQUARE>(a, tile, n0, lane, dst); break;
        case SIGB_WAVE_SAWTOOTH: pipe_source_wave<SIGB_WAVE_SAWTOOTH>(a, tile, n0, lane, dst); break;
        default: pipe_source_wave<SIGB_WAVE_TRIANGLE>(a, tile, n0, lane, dst); break;
    }
}

// Work item = (tile of 64 channels, time segment).  Segment j > 0 starts `warm_chunks` chunks early from
// zero state without storing: the host sizes the warm-up so that the cascade's memory of the unknown true
// state has decayed below 2^-40 (same contract as the time-split pieces of k_chain_scan2).  Segment 0
// continues the carried state exactly; the last segment hands its state to the next launch.
//
// Iteration t of an item:  every thread issues its 16-byte granule of source chunk t + PRE (cp.async) and
// stores its granule of finished chunk t - nsec (x gain) to global memory; warp w filters chunk t - w in
// place.  All section warps do identical work, so nobody waits at the barrier for a straggler.
template <bool BUF, int SPW>
__global__ void __launch_bounds__(BUF ? 256 : 288, BUF ? (SPW == 2 ? 3 : 4) : 2)
k_cascade_pipe(const ChainDev a, int nchunks, int tiles, int nseg, int seg_chunks, int warm_chunks, int nslot, int src_fast, int out_fast) {
    extern __shared__ __align__(16) float ring[];        // [nslot] chunks of (PR x PC); a chunk stays in its slot
                                                         // from its load until its rows have been stored
    constexpr int PRE = pre_chunks(SPW);
    const int nsec = a.nsec / SPW;          // pipeline stages (warps): SPW sections each
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const size_t C = (size_t)a.C;
    const int nthreads = blockDim.x;
    constexpr int GRAN = PR * (PC / 4);                   // 16-byte granules per chunk (256)
    const unsigned ring_base = (unsigned)__cvta_generic_to_shared(ring);
    const unsigned ring_end = ring_base + (unsigned)nslot * CHUNK_FLOATS * 4u;

    for (int item = blockIdx.x; item < tiles * nseg; item += gridDim.x) {
        const int tile = item / nseg, seg = item - tile * nseg;
        const int c_store = seg * seg_chunks;                            // first chunk this item stores
        const int c_end = min(nchunks, c_store + seg_chunks);
        const int c_first = max(0, c_store - (seg > 0 ? warm_chunks : 0));
        const int c0 = tile * PC + 2 * lane;
        const bool live0 = c0 < a.C, live1 = c0 + 1 < a.C;
        SecReg r[2];
        int kind = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) r[j].nc = r[j].a2 = r[j].al = r[j].g = r[j].g2 = r[j].d = r[j].s1 = r[j].s2 = make_float2(0.0f, 0.0f);
        if (w < nsec) {
            const int ca = min(c0, a.C - 1), cb = min(c0 + 1, a.C - 1);
            kind = a.sec_kind[w * SPW];
#pragma unroll
            for (int j = 0; j < SPW; ++j) {
                const int sx = w * SPW + j;
                const float ga = a.coef[(size_t)(sx * 3 + 0) * C + ca], gb = a.coef[(size_t)(sx * 3 + 0) * C + cb];
                const float da = a.coef[(size_t)(sx * 3 + 2) * C + ca], db = a.coef[(size_t)(sx * 3 + 2) * C + cb];
                r[j].g = make_float2(ga, gb);
                r[j].g2 = make_float2(2.0f * ga, 2.0f * gb);
                r[j].nc = make_float2(-a.coef[(size_t)(sx * 3 + 1) * C + ca], -a.coef[(size_t)(sx * 3 + 1) * C + cb]);
                r[j].d = make_float2(da, db);
                r[j].al = make_float2(ga * da, gb * db);
                r[j].a2 = make_float2(2.0f * (ga * da), 2.0f * (gb * db));
                if (c_first == 0) {
                    r[j].s1 = make_float2((float)a.state[(size_t)(sx * 2 + 0) * C + ca], (float)a.state[(size_t)(sx * 2 + 0) * C + cb]);
                    r[j].s2 = make_float2((float)a.state[(size_t)(sx * 2 + 1) * C + ca], (float)a.state[(size_t)(sx * 2 + 1) * C + cb]);
                }
            }
        }

        // ---- per-thread granule: row g_k, channels g_ch .. g_ch + 3 (hoisted out of the time loop)
        const int g_k = tid >> 4, g_col = (tid & 15) * 4;
        const int g_ch = tile * PC + g_col;
        const bool g_in = g_ch + 3 < a.C;
        const unsigned g_off = (unsigned)(g_k * PC + g_col) * 4u;
        const int64_t full_rows = min(a.src_rows, (int64_t)a.frames);
        // a thread owns gpt granules of every chunk: rows g_k, g_k + nthreads/16, ... (same channels)
        const int gpt = (GRAN % nthreads == 0 && nthreads % 16 == 0) ? GRAN / nthreads : 0;
        const bool one_granule = gpt > 0;
        const unsigned row_skip_smem = (unsigned)(nthreads / 16) * PC * 4u;
        const int64_t row_skip_src = (int64_t)(nthreads / 16) * a.src_ld, row_skip_dst = (int64_t)(nthreads / 16) * a.ld_out;
        // source side
        const int fast_end = (BUF && src_fast && one_granule && g_in) ? min(c_end, (int)min(full_rows / PR, (int64_t)nchunks)) : c_first;
        const float* g_src = BUF ? a.src + ((int64_t)c_first * PR + g_k) * a.src_ld + g_ch : nullptr;
        const int64_t src_step = (int64_t)PR * a.src_ld;
        int issue_c = c_first;
        unsigned issue_addr = ring_base;
        auto issue_chunk = [&]() {
            if (issue_c < fast_end) {
                for (int j = 0; j < gpt; ++j)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(issue_addr + g_off + j * row_skip_smem), "l"(g_src + j * row_skip_src) : "memory");
            } else if (issue_c < c_end) {
                float* slot = ring + (issue_addr - ring_base) / 4u;
                const int64_t row0 = (int64_t)issue_c * PR;
                for (int i = tid; i < GRAN; i += nthreads) {
                    const int k = i >> 4, col = (i & 15) * 4;
                    const int ch = tile * PC + col;
                    float* dst = slot + k * PC + col;
                    const int64_t row = row0 + k;
                    if (src_fast && row < full_rows && ch + 3 < a.C) {
                        cp_async16(dst, a.src + row * a.src_ld + ch);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dst[j] = (row < full_rows && ch + j < a.C) ? __ldg(a.src + row * a.src_ld + (int64_t)(ch + j) * a.src_cs) : 0.0f;
                    }
                }
            }
            cp_async_commit();
            ++issue_c;
            g_src += src_step;
            issue_addr += CHUNK_FLOATS * 4u;
            if (issue_addr == ring_end) issue_addr = ring_base;
        };
        // sink side: chunk d = c_first + t - nsec leaves the ring at iteration t
        float4 gain4 = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
        if (a.gain && one_granule) {
            gain4.x = a.gain[min(g_ch, a.C - 1)];
            gain4.y = a.gain[min(g_ch + 1, a.C - 1)];
            gain4.z = a.gain[min(g_ch + 2, a.C - 1)];
            gain4.w = a.gain[min(g_ch + 3, a.C - 1)];
        }
        const bool sink_fast = out_fast && one_granule && g_in;
        int drain_c = c_first - nsec;
        float* g_dst = a.out + ((int64_t)drain_c * PR + g_k) * a.ld_out + g_ch;
        const int64_t dst_step = (int64_t)PR * a.ld_out;
        unsigned drain_addr = ring_base;
        auto drain_chunk = [&]() {
            if (drain_c >= c_store) {
                if (sink_fast && (drain_c + 1) * PR <= a.frames) {
                    for (int j = 0; j < gpt; ++j) {
                        float4 v;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(drain_addr + g_off + j * row_skip_smem) : "memory");
                        v.x *= gain4.x; v.y *= gain4.y; v.z *= gain4.z; v.w *= gain4.w;
                        __stcs(reinterpret_cast<float4*>(g_dst + j * row_skip_dst), v);
                    }
                } else {
                    const float* slot = ring + (drain_addr - ring_base) / 4u;
                    for (int i = tid; i < GRAN; i += nthreads) {
                        const int k = i >> 4, col = (i & 15) * 4;
                        const int64_t row = (int64_t)drain_c * PR + k;
                        if (row >= a.frames) continue;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int ch = tile * PC + col + j;
                            if (ch < a.C) a.out[row * a.ld_out + ch] = slot[k * PC + col + j] * (a.gain ? a.gain[ch] : 1.0f);
                        }
                    }
                }
            }
            if (drain_c >= c_first) {
                drain_addr += CHUNK_FLOATS * 4u;
                if (drain_addr == ring_end) drain_addr = ring_base;
            }
            ++drain_c;
            g_dst += dst_step;
        };

        if (BUF) {
            for (int c = 0; c < PRE; ++c) issue_chunk();
        } else if (w == nsec) {
            pipe_source_chunk(a, tile, c_first, lane, ring);
        }

        // warp w handles chunk c_first + t - w at iteration t; its slot pointer advances with it
        float* my_slot = ring;
        float* src_slot = ring;                           // source warp: slot of the chunk it rendered last
        float* const ring_last = ring + (size_t)nslot * CHUNK_FLOATS;
        int c = c_first - w;                              // this warp's chunk at iteration t
        const int iters = (c_end - c_first) + nsec;       // one extra iteration drains the last chunk
        // one generic iteration: every path (ragged rows, slow granules, idle warps at the pipeline's ends)
        auto iteration = [&](int t) {
            if (BUF) {
                issue_chunk();                   // chunk c_first + t + PRE
                cp_async_wait<PRE>();            // this thread's granules of chunk c_first + t have landed
            }
            __syncthreads();
            drain_chunk();                       // chunk c_first + t - nsec: every section finished it last iteration
            if (w < nsec) {
                if (c >= c_first && c < c_end) {
                    pipe_chunk_kind<SPW>(kind, my_slot, min(PR, a.frames - c * PR), lane, r);
                    my_slot += CHUNK_FLOATS;
                    if (my_slot == ring_last) my_slot = ring;
                }
            } else if (!BUF && c_first + t + 1 < c_end) {
                src_slot += CHUNK_FLOATS;
                if (src_slot == ring_last) src_slot = ring;
                pipe_source_chunk(a, tile, c_first + t + 1, lane, src_slot);
            }
            ++c;
        };
        // Steady state [t_lo, t_hi): every warp holds a full chunk, the chunk being issued takes the cp.async
        // fast path and the chunk being drained is a stored, full one -- a branch-free body with the section
        // kind resolved outside the loop.
        int t_lo = iters, t_hi = iters;
        const bool tile_in = (tile + 1) * PC <= a.C;                        // CTA-uniform: no ragged channels in this tile
        if (BUF && gpt >= 1 && src_fast && out_fast && tile_in) {
            const int full_end = min(c_end, a.frames / PR);                 // chunks with all PR rows
            t_lo = max(nsec, (c_store - c_first) + nsec);
            t_hi = min(fast_end - c_first - PRE, full_end - c_first);
            if (t_hi <= t_lo) t_lo = t_hi = iters;
        }
        int t = 0;
        for (; t < t_lo && t < iters; ++t) iteration(t);
        if (t_lo < t_hi) {
            auto steady = [&](auto kind_tag, auto one_tag) {
                constexpr int KIND = decltype(kind_tag)::value;
                constexpr bool ONE = decltype(one_tag)::value;        // one granule per thread (256-thread CTA): no inner loops
                unsigned my_addr = ring_base + (unsigned)(my_slot - ring) * 4u;
                for (; t < t_hi; ++t) {
                    if (ONE) {
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(issue_addr + g_off), "l"(g_src) : "memory");
                    } else {
                        for (int j = 0; j < gpt; ++j)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(issue_addr + g_off + j * row_skip_smem), "l"(g_src + j * row_skip_src) : "memory");
                    }
                    cp_async_commit();
                    g_src += src_step;
                    issue_addr += CHUNK_FLOATS * 4u;
                    if (issue_addr == ring_end) issue_addr = ring_base;
                    cp_async_wait<PRE>();
                    __syncthreads();
                    for (int j = 0; j < (ONE ? 1 : gpt); ++j) {
                        float4 v;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(drain_addr + g_off + j * row_skip_smem) : "memory");
                        v.x *= gain4.x; v.y *= gain4.y; v.z *= gain4.z; v.w *= gain4.w;
                        __stcs(reinterpret_cast<float4*>(g_dst + j * row_skip_dst), v);
                    }
                    g_dst += dst_step;
                    drain_addr += CHUNK_FLOATS * 4u;
                    if (drain_addr == ring_end) drain_addr = ring_base;
                    if (w < nsec) {
                        pipe_chunk<KIND, SPW>(ring + (my_addr - ring_base) / 4u, PR, lane, r);
                        my_addr += CHUNK_FLOATS * 4u;
                        if (my_addr == ring_end) my_addr = ring_base;
                    }
                }
                const int done = t_hi - t_lo;
                issue_c += done;
                drain_c += done;
                c += done;
                my_slot = ring + (my_addr - ring_base) / 4u;
            };
            if (gpt == 1) {
                switch (kind) {
                    case 0: steady(std::integral_constant<int, 0>{}, std::true_type{}); break;
                    case SEC_HP: steady(std::integral_constant<int, SEC_HP>{}, std::true_type{}); break;
                    case SEC_FIRST_ORDER: steady(std::integral_constant<int, SEC_FIRST_ORDER>{}, std::true_type{}); break;
                    default: steady(std::integral_constant<int, SEC_FIRST_ORDER | SEC_HP>{}, std::true_type{}); break;
                }
            } else {
                switch (kind) {
                    case 0: steady(std::integral_constant<int, 0>{}, std::false_type{}); break;
                    case SEC_HP: steady(std::integral_constant<int, SEC_HP>{}, std::false_type{}); break;
                    case SEC_FIRST_ORDER: steady(std::integral_constant<int, SEC_FIRST_ORDER>{}, std::false_type{}); break;
                    default: steady(std::integral_constant<int, SEC_FIRST_ORDER | SEC_HP>{}, std::false_type{}); break;
                }
            }
        }
        for (; t < iters; ++t) iteration(t);
        if (BUF) cp_async_wait<0>();
        if (w < nsec && c_end == nchunks) {       // the segment that finishes the launch carries the state on
#pragma unroll
            for (int j = 0; j < SPW; ++j) {
                const int sx = w * SPW + j;
                if (live0) {
                    a.state_out[(size_t)(sx * 2 + 0) * C + c0] = (double)r[j].s1.x;
                    a.state_out[(size_t)(sx * 2 + 1) * C + c0] = (double)r[j].s2.x;
                }
                if (live1) {
                    a.state_out[(size_t)(sx * 2 + 0) * C + c0 + 1] = (double)r[j].s1.y;
                    a.state_out[(size_t)(sx * 2 + 1) * C + c0 + 1] = (double)r[j].s2.y;
                }
            }
        }
        __syncthreads();       // item boundary: the ring restarts
    }
}

}  // namespace

// Whether the pipelined kernel can take this chain: a static property of the chain (never of a
// particular call's pointers), so a stream keeps one kernel -- and one state convention -- for its life.
extern "C" int sigb_cascade_pipe_ok(const ChainDev* a) {
    if (a->nsec < 1 || a->nsec > 8 || a->C <= 0) return 0;
    if (a->src_kind == SRC_OSC && (!a->theta0 || !a->dtheta)) return 0;
    if (a->src_kind == SRC_OSC && a->osc_mod && a->wave != SIGB_WAVE_SINE) return 0;      // the guard band is a host maximum here
    return 1;
}

// resident CTAs per SM the launch is sized for: limited by warps (48 of 64 slots), shared memory and 16 CTAs
static int ctas_per_sm(int warps, size_t smem) {
    int n = 48 / warps;
    const int by_smem = (int)((size_t)220 * 1024 / (smem + 1024));
    if (n > by_smem) n = by_smem;
    if (n > 16) n = 16;
    return n < 1 ? 1 : n;
}

// Work items (tiles x time segments) the launch would run: the planner prefers the scan kernel for
// shallow chains when this cannot fill the machine.
extern "C" int sigb_cascade_pipe_items(const ChainDev* a, int max_segments) {
    const int nchunks = (a->frames + PR - 1) / PR;
    const int tiles = (a->C + PC - 1) / PC;
    int nseg = 1;
    if (max_segments > 1 && a->warm_rows >= 0) {
        const int warm_chunks = (a->warm_rows + PR - 1) / PR;
        const int fit = warm_chunks > 0 ? nchunks / (4 * warm_chunks) : nchunks;
        nseg = std::max(1, std::min(fit, max_segments));
    }
    return tiles * nseg;
}

extern "C" int sigb_launch_cascade_pipe(const ChainDev* a, int max_segments, int sections_per_warp, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (a->frames <= 0) return 0;
    const int nchunks = (a->frames + PR - 1) / PR;
    const int tiles = (a->C + PC - 1) / PC;
    const bool buf = a->src_kind == SRC_BUF;
    // two sections per warp when the cascade is an even number of sections of one kind
    bool same = true;
    for (int k = 1; k < a->nsec; ++k) same = same && a->sec_kind[k] == a->sec_kind[0];
    // (measured on C4: 3.30e11 vs 3.35e11 channel-samples/s with one section per warp -- kept selectable)
    const int spw = (sections_per_warp == 2 && a->nsec % 2 == 0 && same && !(a->sec_kind[0] & SEC_FIRST_ORDER)) ? 2 : 1;
    const int nstage = a->nsec / spw;
    const int warps = nstage + (buf ? 0 : 1);
    // a chunk occupies its slot while it is in flight (PRE), while each section works on it (nsec) and while
    // it is stored (1), plus one iteration of slack so a slot is never rewritten right after its last read
    const int nslot = (buf ? pre_chunks(spw) : 1) + nstage + 2;
    const size_t smem = (size_t)nslot * CHUNK_FLOATS * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        const int max_smem = (3 + 8 + 2) * CHUNK_FLOATS * (int)sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(k_cascade_pipe<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_cascade_pipe<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_cascade_pipe<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_cascade_pipe<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    // fast paths: 16-byte cp.async granules of the source, 16-byte stores of the output
    const int src_fast = buf && a->src_cs == 1 && (reinterpret_cast<uintptr_t>(a->src) & 15) == 0 && (a->src_ld & 3) == 0;
    const int out_fast = (reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && (a->ld_out & 3) == 0;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // time segments: enough items to fill the machine in one wave, as long as the warm-up stays below 1/4 of a segment
    const int per_sm = ctas_per_sm(warps, smem);
    int nseg = 1, warm_chunks = 0;
    if (max_segments > 1 && a->warm_rows >= 0) {
        warm_chunks = (a->warm_rows + PR - 1) / PR;
        const int want = std::max(1, sms * per_sm / tiles);          // whole items must fit in one wave
        const int fit = warm_chunks > 0 ? nchunks / (4 * warm_chunks) : nchunks;
        nseg = want < fit ? want : fit;
        if (nseg > max_segments) nseg = max_segments;
        if (nseg < 1) nseg = 1;
    }
    const int seg_chunks = (nchunks + nseg - 1) / nseg;
    nseg = (nchunks + seg_chunks - 1) / seg_chunks;
    const long long items = (long long)tiles * nseg;
    const long long slots = (long long)sms * per_sm;
    const int grid = (int)(items < slots ? items : slots);
#define PIPE_LAUNCH(B, S) k_cascade_pipe<B, S><<<grid, warps * 32, smem, st>>>(*a, nchunks, tiles, nseg, seg_chunks, warm_chunks, nslot, src_fast, out_fast)
    if (buf) { if (spw == 2) PIPE_LAUNCH(true, 2); else PIPE_LAUNCH(true, 1); }
    else { if (spw == 2) PIPE_LAUNCH(false, 2); else PIPE_LAUNCH(false, 1); }
#undef PIPE_LAUNCH
    return (int)cudaGetLastError();
}
