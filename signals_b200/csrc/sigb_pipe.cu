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

#include "sigb200.h"
#include "sigb_internal.h"
#include "sigb_device.cuh"

namespace {

using namespace sigb_dev;

constexpr int PR = 16;          // rows per chunk
constexpr int PC = 64;          // channels per tile (2 per lane)
constexpr int PRE = 6;          // cp.async chunks in flight per CTA (6 x 4 KB)
constexpr int NSLOT = PRE + 2;     // a slot is rewritten two iterations after it was read (one barrier in between)
constexpr int CHUNK_FLOATS = PR * PC;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct SecReg {
    float2 g, nc, d;      // second order: (g) (-c) (d); first order: g = G
    float2 s1, s2;
};

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
    const float2 t = __ffma2_rn(r.nc, r.s1, x);
    const float2 u = __ffma2_rn(r.s2, neg1, t);
    const float2 hp = __fmul2_rn(u, r.d);
    const float2 bp = __ffma2_rn(r.g, hp, r.s1);
    r.s1 = __ffma2_rn(r.g, hp, bp);
    const float2 lp = __ffma2_rn(r.g, bp, r.s2);
    r.s2 = __ffma2_rn(r.g, bp, lp);
    return (KIND & SEC_HP) ? hp : lp;
}

// one chunk of one section: in -> (section) -> smem out, or -> gain -> global rows when LAST
template <int KIND, bool LAST>
__device__ __forceinline__ void pipe_chunk(const float* in, float* out_s, float* out_g, int64_t ld_out, int rows, int lane,
                                           bool live0, bool live1, float2 gain, SecReg& r) {
    if (rows == PR) {
#pragma unroll
        for (int k = 0; k < PR; ++k) {
            const float2 x = *reinterpret_cast<const float2*>(in + k * PC + 2 * lane);
            float2 y = pipe_step<KIND>(x, r);
            if (LAST) {
                y = __fmul2_rn(y, gain);
                float* o = out_g + (int64_t)k * ld_out;
                if (live1) *reinterpret_cast<float2*>(o) = y;
                else if (live0) *o = y.x;
            } else {
                *reinterpret_cast<float2*>(out_s + k * PC + 2 * lane) = y;
            }
        }
    } else {
        for (int k = 0; k < rows; ++k) {     // ragged last chunk: state must stop at the last real row
            const float2 x = *reinterpret_cast<const float2*>(in + k * PC + 2 * lane);
            float2 y = pipe_step<KIND>(x, r);
            if (LAST) {
                y = __fmul2_rn(y, gain);
                float* o = out_g + (int64_t)k * ld_out;
                if (live1) *reinterpret_cast<float2*>(o) = y;
                else if (live0) *o = y.x;
            } else {
                *reinterpret_cast<float2*>(out_s + k * PC + 2 * lane) = y;
            }
        }
    }
}

template <bool LAST>
__device__ __forceinline__ void pipe_chunk_kind(int kind, const float* in, float* out_s, float* out_g, int64_t ld_out, int rows,
                                                int lane, bool live0, bool live1, float2 gain, SecReg& r) {
    switch (kind) {
        case 0: pipe_chunk<0, LAST>(in, out_s, out_g, ld_out, rows, lane, live0, live1, gain, r); break;
        case SEC_HP: pipe_chunk<SEC_HP, LAST>(in, out_s, out_g, ld_out, rows, lane, live0, live1, gain, r); break;
        case SEC_FIRST_ORDER: pipe_chunk<SEC_FIRST_ORDER, LAST>(in, out_s, out_g, ld_out, rows, lane, live0, live1, gain, r); break;
        default: pipe_chunk<SEC_FIRST_ORDER | SEC_HP, LAST>(in, out_s, out_g, ld_out, rows, lane, live0, live1, gain, r); break;
    }
}

// source warp: chunk c of an oscillator / constant source for the tile's 64 channels -> smem
__device__ __noinline__ void pipe_source_chunk(const ChainDev& a, int tile, int c, int lane, float* dst) {
    const int64_t n0 = a.position + (int64_t)c * PR;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        const int ch = tile * PC + 2 * lane + h;
        const int cc = min(ch, a.C - 1);
        float x[PR];
        if (a.src_kind == SRC_CONST) {
            const float v = a.constv[cc];
#pragma unroll
            for (int k = 0; k < PR; ++k) x[k] = v;
        } else {
            const unsigned long long dth = a.dtheta[cc];
            const unsigned long long th = a.theta0[cc] + (unsigned long long)n0 * dth + 0x80000000ull;
            const int w = (int)(th >> 32), dhi = (int)((dth + 0x80000000ull) >> 32);
            bool near;
            switch (a.wave) {
                case SIGB_WAVE_SINE: near = gen_tile<SIGB_WAVE_SINE, PR>(w, dhi, a.guard, x); break;
                case SIGB_WAVE_SQUARE: near = gen_tile<SIGB_WAVE_SQUARE, PR>(w, dhi, a.guard, x); break;
                case SIGB_WAVE_SAWTOOTH: near = gen_tile<SIGB_WAVE_SAWTOOTH, PR>(w, dhi, a.guard, x); break;
                default: near = gen_tile<SIGB_WAVE_TRIANGLE, PR>(w, dhi, a.guard, x); break;
            }
            if (near) {   // within the guard band of a discontinuity: the reference's float64 arithmetic (osc.py:32)
                const double hz = a.hertz[cc], ph = a.phase[cc], rate = (double)a.rate;
                for (int k = 0; k < PR; ++k) x[k] = osc_wave(a.wave, osc_cycles(__ddiv_rn((double)(n0 + k), rate), hz, ph));
            }
        }
#pragma unroll
        for (int k = 0; k < PR; ++k) dst[k * PC + 2 * lane + h] = x[k];
    }
}

template <bool BUF>
__global__ void __launch_bounds__(288, 2) k_cascade_pipe(const ChainDev a, int nchunks, int tiles) {
    extern __shared__ __align__(16) float psm[];
    const int nsec = a.nsec;
    float* ring = psm;                                              // BUF: [NSLOT] chunks; else [2] chunks
    float* stage = psm + (BUF ? NSLOT : 2) * CHUNK_FLOATS;          // [nsec - 1][2] chunks: inputs of sections 1..
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const size_t C = (size_t)a.C;

    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int c0 = tile * PC + 2 * lane;
        const bool live0 = c0 < a.C, live1 = c0 + 1 < a.C;
        SecReg r;
        r.g = r.nc = r.d = r.s1 = r.s2 = make_float2(0.0f, 0.0f);
        int kind = 0;
        float2 gain = make_float2(1.0f, 1.0f);
        if (w < nsec) {
            const int ca = min(c0, a.C - 1), cb = min(c0 + 1, a.C - 1);
            kind = a.sec_kind[w];
            r.g = make_float2(a.coef[(size_t)(w * 3 + 0) * C + ca], a.coef[(size_t)(w * 3 + 0) * C + cb]);
            r.nc = make_float2(-a.coef[(size_t)(w * 3 + 1) * C + ca], -a.coef[(size_t)(w * 3 + 1) * C + cb]);
            r.d = make_float2(a.coef[(size_t)(w * 3 + 2) * C + ca], a.coef[(size_t)(w * 3 + 2) * C + cb]);
            r.s1 = make_float2((float)a.state[(size_t)(w * 2 + 0) * C + ca], (float)a.state[(size_t)(w * 2 + 0) * C + cb]);
            r.s2 = make_float2((float)a.state[(size_t)(w * 2 + 1) * C + ca], (float)a.state[(size_t)(w * 2 + 1) * C + cb]);
            if (a.gain) gain = make_float2(a.gain[ca], a.gain[cb]);
        }
        float* out_g = a.out + c0;

        // 16-byte granule `i` of chunk c: row i / 16, channels 4 (i % 16) .. +3 of the tile
        auto issue_chunk = [&](int c) {
            if (c < nchunks) {
                float* slot = ring + (c % NSLOT) * CHUNK_FLOATS;
                for (int i = tid; i < PR * (PC / 4); i += blockDim.x) {
                    const int k = i >> 4, col = (i & 15) * 4;
                    const int64_t row = (int64_t)c * PR + k;
                    const int ch = tile * PC + col;
                    float* dst = slot + k * PC + col;
                    if (row < a.src_rows && row < a.frames && ch + 3 < a.C) {
                        cp_async16(dst, a.src + row * a.src_ld + ch);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dst[j] = (row < a.src_rows && row < a.frames && ch + j < a.C) ? __ldg(a.src + row * a.src_ld + ch + j) : 0.0f;
                    }
                }
            }
            cp_async_commit();
        };

        if (BUF) {
            for (int c = 0; c < PRE; ++c) issue_chunk(c);
        } else if (w == nsec) {
            pipe_source_chunk(a, tile, 0, lane, ring);
        }

        const int iters = nchunks + nsec - 1;
        for (int t = 0; t < iters; ++t) {
            if (BUF) {
                issue_chunk(t + PRE);
                cp_async_wait<PRE>();            // this thread's granules of chunk t have landed
            }
            __syncthreads();
            if (w < nsec) {
                const int c = t - w;
                if (c >= 0 && c < nchunks) {
                    const int rows = min(PR, a.frames - c * PR);
                    const float* in = w == 0 ? ring + (BUF ? (c % NSLOT) : (c & 1)) * CHUNK_FLOATS
                                             : stage + ((w - 1) * 2 + (c & 1)) * CHUNK_FLOATS;
                    if (w == nsec - 1)
                        pipe_chunk_kind<true>(kind, in, nullptr, out_g + (int64_t)c * PR * a.ld_out, a.ld_out, rows, lane, live0, live1, gain, r);
                    else
                        pipe_chunk_kind<false>(kind, in, stage + (w * 2 + (c & 1)) * CHUNK_FLOATS, nullptr, 0, rows, lane, live0, live1, gain, r);
                }
            } else if (!BUF && t + 1 < nchunks) {
                pipe_source_chunk(a, tile, t + 1, lane, ring + ((t + 1) & 1) * CHUNK_FLOATS);
            }
        }
        if (BUF) cp_async_wait<0>();
        if (w < nsec) {
            if (live0) {
                a.state[(size_t)(w * 2 + 0) * C + c0] = (double)r.s1.x;
                a.state[(size_t)(w * 2 + 1) * C + c0] = (double)r.s2.x;
            }
            if (live1) {
                a.state[(size_t)(w * 2 + 0) * C + c0 + 1] = (double)r.s1.y;
                a.state[(size_t)(w * 2 + 1) * C + c0 + 1] = (double)r.s2.y;
            }
        }
        __syncthreads();       // tile boundary: the ring and the stage buffers restart
    }
}

}  // namespace

// Whether the pipelined kernel can take this chain (alignment of the packed 8-byte stores and the
// 16-byte cp.async granules); the planner falls back to the scan kernel otherwise.
extern "C" int sigb_cascade_pipe_ok(const ChainDev* a) {
    if (a->nsec < 2 || a->nsec > 8 || a->frames <= 0 || a->C <= 0) return 0;
    if ((reinterpret_cast<uintptr_t>(a->out) & 7) != 0 || (a->ld_out & 1) != 0) return 0;
    if (a->src_kind == SRC_BUF) {
        if (a->src_cs != 1 || (reinterpret_cast<uintptr_t>(a->src) & 15) != 0 || (a->src_ld & 3) != 0) return 0;
    } else if (a->src_kind == SRC_OSC) {
        if (!a->theta0 || !a->dtheta) return 0;
    }
    return 1;
}

extern "C" int sigb_launch_cascade_pipe(const ChainDev* a, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int nchunks = (a->frames + PR - 1) / PR;
    const int tiles = (a->C + PC - 1) / PC;
    const bool buf = a->src_kind == SRC_BUF;
    const int warps = a->nsec + (buf ? 0 : 1);
    const size_t smem = ((buf ? NSLOT : 2) + (size_t)(a->nsec - 1) * 2) * CHUNK_FLOATS * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        const int max_smem = (NSLOT + 7 * 2) * CHUNK_FLOATS * (int)sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(k_cascade_pipe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_cascade_pipe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = tiles < sms * 2 ? tiles : sms * 2;
    if (buf) k_cascade_pipe<true><<<grid, warps * 32, smem, st>>>(*a, nchunks, tiles);
    else k_cascade_pipe<false><<<grid, warps * 32, smem, st>>>(*a, nchunks, tiles);
    return (int)cudaGetLastError();
}
