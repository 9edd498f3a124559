// k_cascade_delta / k_osc_delta / k_cascade_reg / k_osc_reg: register-resident filter cascades (sm_100a) for chains of two and
// more sections on many channels (BASELINE config C4: HBM buffer -> 8 chained Butterworth low-pass biquads, 16,384 channels
// x 60 s).  The *_delta kernels run second-order sections in DELTA FORM (5 / 4 operations per low- / high-pass section, see
// below) and are the default; the *_reg kernels keep the state-variable section (6 / 7 operations) for cascades with
// first-order sections and for ragged or unaligned blocks.  What follows describes the decomposition they all share.
//
// k_cascade_pipe hands every chunk from section to section through shared memory (one LDS.64 + one STS.64
// per section per two channel-samples) and measured 0.42 of the HBM roofline.  Here a thread owns TWO adjacent
// channels (packed f32x2 math) and keeps ALL sections of them in registers: per row one 8-byte source read
// (k_cascade_reg: through a warp-private cp.async ring; k_osc_reg: an oscillator evaluated in the thread),
// NSEC x (FADD2 + 5 FFMA2), one multiply by the folded gain, one 8-byte streaming store.  Instruction-level
// parallelism comes from the sections: a block of R rows is evaluated in wavefront order (section s of row r
// next to section s-1 of row r+1).  Thread-level parallelism comes from channels (one warp = 64 adjacent
// channels = 256-byte rows) and from TIME: the (tile, block) space is cut into one equal piece per warp slot of
// the machine; a piece that starts inside a tile begins `warm` rows early from zero state without storing --
// the host sizes the warm-up so that the cascade's memory of the unknown true state has decayed below 2^-40
// (same contract as k_cascade_pipe / k_chain_scan2).  A piece that starts at row 0 continues the carried state
// exactly; the piece that finishes a tile hands its state to the next launch.
//
// What bounds it (tools/fma_probe.cu, profiles/r01_fma_probe.txt): every multiply-add of the section reads
// three different registers, so the loop is limited by register-file operand reads at ~96 FMA/clk/SM, not by
// the 128-lane FMA datapath; the five-coefficient form below lets ptxas mark the operands shared by
// consecutive instructions .reuse (85 -> 96 FMA/clk/SM register-only).
//
// Section arithmetic (pipe_step of sigb_pipe.cu, same state convention):  with e = x - c s1 - s2,
//   bp = s1 + g d e;   s1' = s1 + 2 g d e;   lp = s2 + g bp;   s2' = s2 + 2 g bp;   hp = d e
//
// Reference semantics: CritFilter._filter, /root/reference/src/signals/chain/fx.py:85-121 (per-channel
// Butterworth sections run from zero state over the whole request).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <type_traits>

#include "sigb200.h"
#include "sigb_internal.h"
#include "sigb_device.cuh"

namespace {

using namespace sigb_dev;

int g_reg_pieces = 1;           // pieces per warp slot of the machine (A/B: more pieces even out the tail, each costs a warm-up)
int g_osc_pieces_pct = 100;     // A/B ("osc_pieces_pct"): time pieces of the register kernels as a percentage of the resident warp slots

constexpr int RC = 64;          // channels per warp (2 per lane)
constexpr int RWARPS = 4;       // warps per CTA: 256 adjacent channels, 1 KB of every row

struct RegSec {
    float2 nc, al, a2, g, g2;   // (-c) (g d) (2 g d) (g) (2 g)
    float2 d;                   // high-pass output scale (dead code for low-pass cascades)
    float2 s1, s2;
};

// One sample (two channels) of a zero-delay-feedback state-variable section, the arithmetic of pipe_step in
// sigb_pipe.cu.  The two updates that share (e, s1) and the two that share (bp, s2) are written back to back:
// ptxas marks the shared operands .reuse, which matters because this loop is bound by register-file reads
// (tools/fma_probe.cu: 96 FMA/clk/SM in this form against 85 with three coefficients and no shared operands).
template <int KIND>
__device__ __forceinline__ float2 reg_step(float2 x, RegSec& r) {
    const float2 xs = __fadd2_rn(x, make_float2(-r.s2.x, -r.s2.y));
    const float2 e = __ffma2_rn(r.nc, r.s1, xs);
    const float2 bp = __ffma2_rn(r.al, e, r.s1);
    r.s1 = __ffma2_rn(r.a2, e, r.s1);
    const float2 lp = __ffma2_rn(r.g, bp, r.s2);
    r.s2 = __ffma2_rn(r.g2, bp, r.s2);
    return (KIND & SEC_HP) ? __fmul2_rn(e, r.d) : lp;
}

// a value ptxas must keep in its register (it would rather recompute 2g and 2gd inside the loop than hold them)
__device__ __forceinline__ float keep(float v) {
    asm volatile("" : "+f"(v));
    return v;
}

// Coefficients of section s for channels (ca, cb) from the plan's {g, c, d} table.  A FIRST-ORDER section (odd Butterworth
// orders; table entry {G, 0, 0}, G = g / (1 + g): v = G (x - s1), lp = s1 + v, s1' = s1 + 2 v, hp = x - lp) runs on the very
// same six instructions with coefficients chosen here, once: c = 1, g d = G, and for a low-pass g = 1, 2g = 0 (so that
// lp = s2 + 1 bp = bp and s2 stays 0), for a high-pass d = 1 - G (hp = (1 - G) e) and 2g = 0.  The state convention is the
// other kernels' (s1 as in svf_any, s2 = 0), so streams can still pass between kernels.
__device__ __forceinline__ void load_section(const ChainDev& a, int s, int ca, int cb, RegSec& r) {
    const size_t C = (size_t)a.C;
    const float ga = a.coef[(size_t)(s * 3 + 0) * C + ca], gb = a.coef[(size_t)(s * 3 + 0) * C + cb];
    if (a.sec_kind[s] & SEC_FIRST_ORDER) {
        r.nc = make_float2(-1.0f, -1.0f);
        r.al = make_float2(ga, gb);
        r.a2 = make_float2(keep(2.0f * ga), keep(2.0f * gb));
        r.g = make_float2(1.0f, 1.0f);
        r.g2 = make_float2(keep(0.0f), keep(0.0f));
        r.d = make_float2(1.0f - ga, 1.0f - gb);
        return;
    }
    const float da = a.coef[(size_t)(s * 3 + 2) * C + ca], db = a.coef[(size_t)(s * 3 + 2) * C + cb];
    r.g = make_float2(ga, gb);
    r.nc = make_float2(-a.coef[(size_t)(s * 3 + 1) * C + ca], -a.coef[(size_t)(s * 3 + 1) * C + cb]);
    r.d = make_float2(da, db);
    r.g2 = make_float2(keep(2.0f * ga), keep(2.0f * gb));
    r.al = make_float2(ga * da, gb * db);
    r.a2 = make_float2(keep(2.0f * (ga * da)), keep(2.0f * (gb * db)));
}

// ---------------------------------------------------------------------------------------------------------
// DELTA FORM of a second-order low-pass section (default for all-low-pass cascades of second-order sections).
//
// The bilinear Butterworth section is  b0 (1 + z^-1)^2 / (1 + a1 z^-1 + a2 z^-2),  b0 = (1 + a1 + a2) / 4.  The six
// multiply-adds of the state-variable form all read three different registers, and the cascade is bound by exactly
// that (register-file operand reads, ~96 lane-ops/clk/SM).  Here the all-pole part runs as a difference recurrence
//     d[n] = a d[n-1] + F (x[n] - y[n-1]),   y[n] = y[n-1] + d[n],      F = 1 + a1 + a2 = 4 g^2 d,  a = a2 = 1 - 2 r2 g d
// (F and 1 - a are the SMALL numbers at low cutoffs and are held as such, which is what keeps float32 exact there -- a
// direct form holds a1 ~ -2 and loses 1e-3 at 20 Hz), and the two zeros at Nyquist are two plain adds:
//     lp[n] = (y[n] + 2 y[n-1] + y[n-2]) / 4.
// With Z = y / 4 and D = d / F:   w = x - 4 Z;  D' = a D + w;  Z' = Z + (F/4) D';  p = Z' + Z;  lp = p + P;  P' = p
// = 3 FFMA2 + 2 FADD2 per two channel-samples (5 instead of 6 operations, two of them with two operands), two
// coefficient pairs instead of five, three state pairs instead of two: 10 registers per section instead of 14.
//
// The states are the state-variable section's own, re-scaled (bp = (y[n] - y[n-2]) / 4g, hp = (y[n] - 2 y[n-1] +
// y[n-2]) / 4 g^2  =>  s1 = bp[n-1] + g hp[n-1] = (y[n-1] - y[n-2]) / 2g,  s2 = lp[n-1] + g bp[n-1] = (y[n-1] + y[n-2]) / 2):
//     P = s2 / 2,   D = s1 / (2 g d),   Z = (s2 + g s1) / 4;       s2 = 2 P = 4 Z - (F/2) D,   s1 = 2 g d D
// so a stream passes between this form and every other kernel of the plan with a pointwise conversion.
// Measured float32 error against the float64 cascade, 8 sections x 60 s (tools/delta_form_sim.c): 8.1e-7 at 200 Hz
// cutoffs (state-variable form 8.0e-7), 5.7e-6 at 20 Hz (1.3e-6), 1.2e-6 at 20 kHz (7.7e-7): inside the 1e-4 bar.
// ---------------------------------------------------------------------------------------------------------
struct DeltaSec {
    float2 a, be;               // low-pass: (a2) (F / 4);   high-pass and mixed cascades: (-Q) (-F)
    float2 c;                   // high-pass and mixed cascades: scale of the value handed to the next section
    float2 D, Z, P;             // high-pass and mixed cascades: Z holds -4 Z (and P, the low-pass memory, -4 P)
};
constexpr int SEC_MIXED = 4;    // kernel template value: low- and high-pass sections in one cascade (kind per section at run time)

// HIGH-PASS sections in the same variables (all-high-pass cascades): hp[n] = (y[n] - 2 y[n-1] + y[n-2]) / 4 g^2 is the second
// difference of the all-pole output, and the recurrence hands it over without a subtraction of neighbours:
//     t = w - Q D,   D' = D + t  (= a D + w),   hp = d t          with w = x - 4 Z,  Q = 1 - a2 = 2 r2 g d,  d = F / 4 g^2.
// The output scale d rides on the NEXT section's first instruction (w' = d t - 4 Z' as one FFMA2 on the state -4 Z'), the
// last one's on the gain: 4 operations per section (FFMA2, FFMA2, FADD2, FFMA2) against 7 in state-variable form.
// Same states, same conversion.  Float32 error against the float64 cascade, 8 sections x 30 s (tools/delta_form_sim_hp.c):
// 3.1e-7 at 200 Hz (state-variable form 4.5e-7), 5.3e-7 at 20 Hz (4.9e-7), 2.7e-6 at 20 kHz (1.5e-6).
// MIXED cascades (low- and high-pass sections in any order) run both forms on the shared states in the scaling of the
// high-pass form: the four instructions of the high-pass section, the low-pass section's two adds (on -4 Z, so the value is
// -4 lp and the scale handed on is -1/4), and a select by the section's kind -- 6 FP32-pipe operations + a select on the
// ALU pipe per section, the cost of a state-variable section, instead of the section-pipelined kernel's shared-memory
// hand-offs.
template <int KIND>
__device__ __forceinline__ float2 delta_step(float2 x, DeltaSec& r, const float2 m4, bool first, const float2 cprev, bool is_hp) {
    if (KIND & (SEC_HP | SEC_MIXED)) {
        const float2 w = first ? __fadd2_rn(x, r.Z) : __ffma2_rn(cprev, x, r.Z);
        const float2 t = __ffma2_rn(r.a, r.D, w);
        r.D = __fadd2_rn(r.D, t);
        const float2 zn = __ffma2_rn(r.be, r.D, r.Z);
        if (KIND & SEC_MIXED) {
            const float2 p = __fadd2_rn(zn, r.Z);
            const float2 o = __fadd2_rn(p, r.P);
            r.P = p;
            r.Z = zn;
            return is_hp ? t : o;
        }
        r.Z = zn;
        return t;
    }
    const float2 w = __ffma2_rn(m4, r.Z, x);
    r.D = __ffma2_rn(r.a, r.D, w);
    const float2 zn = __ffma2_rn(r.be, r.D, r.Z);
    const float2 p = __fadd2_rn(zn, r.Z);
    r.Z = zn;
    const float2 o = __fadd2_rn(p, r.P);
    r.P = p;
    return o;
}

// coefficients from the plan's float32 {g, c, d} entries (the very numbers the other kernels filter with), derived in
// float64 so that each carries one rounding
template <int KIND>
__device__ __forceinline__ void delta_coef(float g, float c, float d, bool is_hp, float& a, float& be, float& sc) {
    const double G = (double)g, Dd = (double)d, R2 = (double)c - G;
    if (KIND & (SEC_HP | SEC_MIXED)) {
        a = (float)(-2.0 * R2 * G * Dd);
        be = (float)(-4.0 * G * G * Dd);
        sc = is_hp ? d : -0.25f;
    } else {
        a = (float)(1.0 - 2.0 * R2 * G * Dd);
        be = (float)(G * G * Dd);
    }
}
// State hand-over.  D and Z are the recurrent states and survive a store / load round trip bit for bit (s1 = 2 g d D and
// s2 = 4 Z - (F/2) D are formed in float64 from float32 factors, and the load inverts them with the same float32 F);
// the low-pass P is only the one-row memory of the second zero and comes back as s2 / 2, equal to the running value up to
// one float32 rounding -- so a low-pass stream cut into several requests differs from the uncut one by rounding noise
// (~1e-7), not bit for bit as with the state-variable kernels.
template <int KIND>
__device__ __forceinline__ void delta_state_in(float g, float d, float be, double s1, double s2, float& D, float& Z, float& P) {
    const double Dd = s1 / (2.0 * (double)g * (double)d);
    D = (float)Dd;
    if (KIND & (SEC_HP | SEC_MIXED)) {
        Z = (float)(0.5 * (double)be * Dd - s2);            // -4 Z = -(s2 + (F/2) D), be = -F
        if (KIND & SEC_MIXED) P = (float)(-2.0 * s2);       // -4 P
    } else {
        Z = (float)(0.25 * (s2 + 2.0 * (double)be * Dd));
        P = (float)(0.5 * s2);
    }
}
template <int KIND>
__device__ __forceinline__ void delta_state_out(float g, float d, float be, float D, float Z, double& s1, double& s2) {
    s1 = 2.0 * (double)g * (double)d * (double)D;
    if (KIND & (SEC_HP | SEC_MIXED)) s2 = 0.5 * (double)be * (double)D - (double)Z;
    else s2 = 4.0 * (double)Z - 2.0 * (double)be * (double)D;
}

constexpr int RING_D = 4;       // blocks of R rows per warp in the cp.async ring (RING_D - 1 in flight)

__device__ __forceinline__ void cp_async8(unsigned smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float2 lds_f2(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}


// R rows through NSEC sections in wavefront order
template <int NSEC, int KIND, int R>
__device__ __forceinline__ void reg_block(float2 (&x)[R], RegSec (&sec)[NSEC]) {
#pragma unroll
    for (int dgl = 0; dgl < R + NSEC - 1; ++dgl) {
#pragma unroll
        for (int s = 0; s < NSEC; ++s) {
            const int r = dgl - s;
            if (r >= 0 && r < R) x[r] = reg_step<KIND>(x[r], sec[s]);
        }
    }
}

// FAST: every lane of every warp owns two live channels, the source covers all rows of the launch, and both
// blocks take 8-byte accesses (host-checked) -- the block loop is loads, FFMA2 and stores only.
// resident CTAs per SM the kernel is compiled for: five coefficient and two state register pairs per section
// (14 registers) -- up to five sections fit 128 registers, deeper cascades get 168 (at 128 ptxas recomputes
// 2g and 2gd in the loop)
__host__ __device__ constexpr int reg_min_blocks(int nsec, int rows) { return (nsec <= 5 && rows <= 4) ? 4 : 3; }

template <int NSEC, int KIND, int R, bool FAST>
__global__ void __launch_bounds__(RWARPS * 32, reg_min_blocks(NSEC, R))
k_cascade_reg(const ChainDev a, int tiles, int npieces, int warm_rows) {
    // Work decomposition: the (tile, block-of-R-rows) space, tile-major, is cut into `npieces` equal contiguous
    // pieces, one per warp, so every warp slot of the machine gets the same number of rows whatever the tile count
    // (C4: 256 tiles on 148 x 12 slots).  A piece is walked as sub-ranges [b0, b1) of one tile each.
    __shared__ __align__(16) float2 ring[FAST ? RWARPS * RING_D * R * 32 : 1];
    const unsigned ring_base = (unsigned)__cvta_generic_to_shared(ring);
    const int lane = threadIdx.x & 31;
    const int piece = blockIdx.x * RWARPS + (threadIdx.x >> 5);
    if (piece >= npieces) return;
    const size_t C = (size_t)a.C;
    const int bpt = (a.frames + R - 1) / R;                                // blocks per tile (the last one may be ragged)
    const int64_t total = (int64_t)tiles * bpt;
    int64_t blk = total * piece / npieces;
    const int64_t blk_end = total * (piece + 1) / npieces;
    const int64_t full_rows = min(a.src_rows, (int64_t)a.frames);          // rows beyond read as zero
    // FAST only: byte offsets of the block's rows as 32-bit values held in registers (host-checked range)
    uint32_t src_off[R], out_off[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        src_off[k] = (uint32_t)k * ((uint32_t)a.src_ld * 4u);
        out_off[k] = (uint32_t)k * ((uint32_t)a.ld_out * 4u);
        if (FAST && k > 0) { asm volatile("" : "+r"(src_off[k])); asm volatile("" : "+r"(out_off[k])); }
    }
  while (blk < blk_end) {
    const int tile = (int)(blk / bpt);
    const int b0 = (int)(blk - (int64_t)tile * bpt);
    const int b1 = (int)min((int64_t)bpt, b0 + (blk_end - blk));
    blk += b1 - b0;
    const int c0 = tile * RC + 2 * lane;
    const bool live0 = c0 < a.C, live1 = c0 + 1 < a.C;
    const int ca = min(c0, a.C - 1), cb = min(c0 + 1, a.C - 1);
    const int row_store = b0 * R;                                          // first row this sub-range stores
    const int row_end = min(a.frames, b1 * R);
    const int row_first = max(0, row_store - warm_rows);                   // warm_rows is a multiple of R

    RegSec sec[NSEC];
#pragma unroll
    for (int s = 0; s < NSEC; ++s) {
        load_section(a, s, ca, cb, sec[s]);
        if (row_first == 0) {
            sec[s].s1 = make_float2((float)a.state[(size_t)(s * 2 + 0) * C + ca], (float)a.state[(size_t)(s * 2 + 0) * C + cb]);
            sec[s].s2 = make_float2((float)a.state[(size_t)(s * 2 + 1) * C + ca], (float)a.state[(size_t)(s * 2 + 1) * C + cb]);
        } else {
            sec[s].s1 = sec[s].s2 = make_float2(0.0f, 0.0f);
        }
    }
    float2 gain = make_float2(1.0f, 1.0f);
    if (a.gain) gain = make_float2(a.gain[ca], a.gain[cb]);

    const float* srcp = a.src + (int64_t)row_first * a.src_ld + (int64_t)ca * a.src_cs;
    const int64_t cs1 = live1 ? (int64_t)a.src_cs : 0;
    float* outp = a.out + (int64_t)row_first * a.ld_out + c0;

    // rows [row, row + R) of the source (guarded path); the caller guarantees row + R <= row_end
    auto load_block = [&](int row, const float* sp, float2 (&x)[R]) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const bool in = row + k < full_rows;
            x[k].x = in ? __ldg(sp + (int64_t)k * a.src_ld) : 0.0f;
            x[k].y = in ? __ldg(sp + (int64_t)k * a.src_ld + cs1) : 0.0f;
        }
    };

    int row = row_first;
    const int nfull = (row_end - row_first) / R;                           // whole blocks of R rows
    if (FAST) {
        // the source runs RING_D - 1 blocks ahead of the math through a warp-private shared-memory ring filled by
        // cp.async: a lane only ever reads the 8 bytes per row it copied itself, so no barrier is involved, and
        // the prefetch depth costs no registers (ncu: waiting on the loads was the top stall with one block ahead)
        const unsigned my = ring_base + (unsigned)((threadIdx.x >> 5) * (RING_D * R * 32) + lane) * 8u;
        const char* ip = reinterpret_cast<const char*>(srcp);
        const int64_t step_b = (int64_t)R * a.src_ld * 4;
        int in_blk = 0;
        unsigned in_addr = my, out_addr = my;                              // slot addresses advance with wrap-around
        const unsigned my_end = my + RING_D * R * 256u;
#pragma unroll
        for (int j = 0; j < RING_D - 1; ++j) {
            if (in_blk < nfull) {
#pragma unroll
                for (int k = 0; k < R; ++k) cp_async8(in_addr + k * 256u, ip + src_off[k]);
                ip += step_b;
                ++in_blk;
                in_addr += R * 256u;
                if (in_addr == my_end) in_addr = my;
            }
            cp_async_commit();
        }
        for (int b = 0; b < nfull; ++b) {
            if (in_blk < nfull) {
#pragma unroll
                for (int k = 0; k < R; ++k) cp_async8(in_addr + k * 256u, ip + src_off[k]);
                ip += step_b;
                ++in_blk;
                in_addr += R * 256u;
                if (in_addr == my_end) in_addr = my;
            }
            cp_async_commit();
            cp_async_wait<RING_D - 1>();
            float2 x[R];
#pragma unroll
            for (int k = 0; k < R; ++k) x[k] = lds_f2(out_addr + k * 256u);
            out_addr += R * 256u;
            if (out_addr == my_end) out_addr = my;
            reg_block<NSEC, KIND, R>(x, sec);
            if (row >= row_store) {
#pragma unroll
                for (int k = 0; k < R; ++k)
                    __stcs(reinterpret_cast<float2*>(reinterpret_cast<char*>(outp) + out_off[k]), __fmul2_rn(x[k], gain));
            }
            outp += (int64_t)R * a.ld_out;
            row += R;
        }
        cp_async_wait<0>();
        srcp += (int64_t)nfull * R * a.src_ld;
    } else {
        float2 nxt[R];
        if (nfull > 0) load_block(row, srcp, nxt);
        for (int b = 0; b < nfull; ++b) {
            float2 x[R];
#pragma unroll
            for (int k = 0; k < R; ++k) x[k] = nxt[k];
            srcp += (int64_t)R * a.src_ld;
            if (b + 1 < nfull) load_block(row + R, srcp, nxt);             // prefetch the next block behind the math
            reg_block<NSEC, KIND, R>(x, sec);
            if (row >= row_store) {
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    if (live0) outp[(int64_t)k * a.ld_out] = x[k].x * gain.x;
                    if (live1) outp[(int64_t)k * a.ld_out + 1] = x[k].y * gain.y;
                }
            }
            outp += (int64_t)R * a.ld_out;
            row += R;
        }
    }
    // ragged tail (< R rows; only the segment that ends the launch has one): the state stops at the last real row
    for (; row < row_end; ++row) {
        float2 x;
        const bool in = row < full_rows;
        x.x = in ? __ldg(srcp) : 0.0f;
        x.y = in ? __ldg(srcp + cs1) : 0.0f;
#pragma unroll
        for (int s = 0; s < NSEC; ++s) x = reg_step<KIND>(x, sec[s]);
        if (live0) outp[0] = x.x * gain.x;
        if (live1) outp[1] = x.y * gain.y;
        srcp += a.src_ld;
        outp += a.ld_out;
    }
    if (row_end == a.frames) {                // the segment that finishes the launch carries the state on
#pragma unroll
        for (int s = 0; s < NSEC; ++s) {
            if (live0) {
                a.state_out[(size_t)(s * 2 + 0) * C + c0] = (double)sec[s].s1.x;
                a.state_out[(size_t)(s * 2 + 1) * C + c0] = (double)sec[s].s2.x;
            }
            if (live1) {
                a.state_out[(size_t)(s * 2 + 0) * C + c0 + 1] = (double)sec[s].s1.y;
                a.state_out[(size_t)(s * 2 + 1) * C + c0 + 1] = (double)sec[s].s2.y;
            }
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// k_cascade_delta: k_cascade_reg's decomposition (two channels per thread, every section in registers, blocks of 8
// rows in wavefront order, equal time pieces per warp slot, warp-private cp.async ring) with the sections in DELTA
// FORM (above).  FAST layout only (host-checked), second-order low-pass sections only.
// ---------------------------------------------------------------------------------------------------------
// resident CTAs per SM: 10 registers per section + the block's rows.  8 sections at 4 CTAs (128 registers) spill; measured on
// C4 (same box, alternating): 4 / 3 / 2 CTAs per SM = 0.672-0.678 / 0.694-0.697 and 0.719-0.726 / 0.732 of the HBM peak -- the
// section chains supply the instruction-level parallelism, the cp.async ring hides the loads, and fewer warp slots mean
// fewer time pieces, i.e. less warm-up
// 7 and 8 sections run 16-ROW blocks at 2 CTAs per SM: with 8 warps per SM the wavefront order is what supplies the
// instruction-level parallelism, and a 16 x 8 parallelogram holds 5.6 independent section steps per diagonal against 4.3 for
// 8 x 8 -- C4 0.695-0.701 -> 0.735 of the HBM peak on the same box (profiles/r02_c4_delta.txt, call r02y).
__host__ __device__ constexpr int delta_min_blocks(int nsec, int kind = 0) {
    return (kind & 4) ? (nsec <= 4 ? 4 : nsec <= 6 ? 3 : 2)          // mixed cascades: 12 registers per section
                      : (nsec <= 4 ? 4 : nsec <= 6 ? 3 : 2);
}

template <int NSEC, int R, int MINB, int WR = R, int KIND = 0>
__global__ void __launch_bounds__(RWARPS * 32, MINB)
k_cascade_delta(const ChainDev a, int tiles, int npieces, int warm_rows) {
    constexpr int RD = R >= 16 ? 3 : RING_D;             // ring depth in blocks (16-row blocks: 48 KB of static shared memory at depth 3)
    __shared__ __align__(16) float2 ring[RWARPS * RD * R * 32];
    const unsigned ring_base = (unsigned)__cvta_generic_to_shared(ring);
    const int lane = threadIdx.x & 31;
    const int piece = blockIdx.x * RWARPS + (threadIdx.x >> 5);
    if (piece >= npieces) return;
    const size_t C = (size_t)a.C;
    const int bpt = (a.frames + R - 1) / R;
    const int64_t total = (int64_t)tiles * bpt;
    int64_t blk = total * piece / npieces;
    const int64_t blk_end = total * (piece + 1) / npieces;
    float2 m4 = make_float2(-4.0f, -4.0f);
    asm volatile("" : "+f"(m4.x), "+f"(m4.y));            // one register pair, not an immediate per use
    const uint32_t src_ldb = (uint32_t)a.src_ld * 4u, out_ldb = (uint32_t)a.ld_out * 4u;   // row strides in bytes (host-checked range)
  while (blk < blk_end) {
    const int tile = (int)(blk / bpt);
    const int b0 = (int)(blk - (int64_t)tile * bpt);
    const int b1 = (int)min((int64_t)bpt, b0 + (blk_end - blk));
    blk += b1 - b0;
    const int c0 = tile * RC + 2 * lane;
    const int row_store = b0 * R;
    const int row_end = min(a.frames, b1 * R);
    const int row_first = max(0, row_store - warm_rows);                   // warm_rows is a multiple of R
    const int nfull = (row_end - row_first) / R;

    DeltaSec sec[NSEC];
#pragma unroll
    for (int s = 0; s < NSEC; ++s) {
        const float2 g = *reinterpret_cast<const float2*>(a.coef + (size_t)(s * 3 + 0) * C + c0);
        const float2 c = *reinterpret_cast<const float2*>(a.coef + (size_t)(s * 3 + 1) * C + c0);
        const float2 d = *reinterpret_cast<const float2*>(a.coef + (size_t)(s * 3 + 2) * C + c0);
        const bool hp_s = (KIND & SEC_HP) || ((KIND & SEC_MIXED) && (a.sec_kind[s] & SEC_HP));
        delta_coef<KIND>(g.x, c.x, d.x, hp_s, sec[s].a.x, sec[s].be.x, sec[s].c.x);
        delta_coef<KIND>(g.y, c.y, d.y, hp_s, sec[s].a.y, sec[s].be.y, sec[s].c.y);
        if (row_first == 0) {
            delta_state_in<KIND>(g.x, d.x, sec[s].be.x, a.state[(size_t)(s * 2 + 0) * C + c0], a.state[(size_t)(s * 2 + 1) * C + c0], sec[s].D.x, sec[s].Z.x, sec[s].P.x);
            delta_state_in<KIND>(g.y, d.y, sec[s].be.y, a.state[(size_t)(s * 2 + 0) * C + c0 + 1], a.state[(size_t)(s * 2 + 1) * C + c0 + 1], sec[s].D.y, sec[s].Z.y, sec[s].P.y);
        } else {
            sec[s].D = sec[s].Z = sec[s].P = make_float2(0.0f, 0.0f);
        }
    }
    float2 gain = make_float2(1.0f, 1.0f);
    if (a.gain) gain = *reinterpret_cast<const float2*>(a.gain + c0);
    if (KIND & (SEC_HP | SEC_MIXED)) gain = __fmul2_rn(gain, sec[NSEC - 1].c);   // the last section's output scale

    const char* ip = reinterpret_cast<const char*>(a.src + (int64_t)row_first * a.src_ld + c0);
    char* op = reinterpret_cast<char*>(a.out + (int64_t)row_first * a.ld_out + c0);
    const unsigned my = ring_base + (unsigned)((threadIdx.x >> 5) * (RD * R * 32) + lane) * 8u;
    const unsigned my_end = my + RD * R * 256u;
    unsigned in_addr = my, out_addr = my;
    int in_blk = 0;
    auto prefetch = [&]() {
        if (in_blk < nfull) {
#pragma unroll
            for (int k = 0; k < R; ++k) cp_async8(in_addr + k * 256u, ip + (uint32_t)k * src_ldb);
            ip += (int64_t)R * src_ldb;
            ++in_blk;
            in_addr += R * 256u;
            if (in_addr == my_end) in_addr = my;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int j = 0; j < RD - 1; ++j) prefetch();
    int row = row_first;
    for (int b = 0; b < nfull; ++b) {
        prefetch();
        cp_async_wait<RD - 1>();
        float2 x[R];
#pragma unroll
        for (int k = 0; k < R; ++k) x[k] = lds_f2(out_addr + k * 256u);
        out_addr += R * 256u;
        if (out_addr == my_end) out_addr = my;
#pragma unroll
        for (int h = 0; h < R; h += WR) {                 // wavefronts of WR rows
#pragma unroll
            for (int dgl = 0; dgl < WR + NSEC - 1; ++dgl) {
#pragma unroll
                for (int s = 0; s < NSEC; ++s) {
                    const int r = dgl - s;
                    if (r >= 0 && r < WR) x[h + r] = delta_step<KIND>(x[h + r], sec[s], m4, s == 0, sec[s > 0 ? s - 1 : 0].c, (a.sec_kind[s] & SEC_HP) != 0);
                }
            }
        }
        if (row >= row_store) {
            if (KIND == 0 && !a.gain) {            // a low-pass cascade without a Gain behind it: nothing to scale
#pragma unroll
                for (int k = 0; k < R; ++k) __stcs(reinterpret_cast<float2*>(op + (uint32_t)k * out_ldb), x[k]);
            } else {
#pragma unroll
                for (int k = 0; k < R; ++k) __stcs(reinterpret_cast<float2*>(op + (uint32_t)k * out_ldb), __fmul2_rn(x[k], gain));
            }
        }
        op += (int64_t)R * out_ldb;
        row += R;
    }
    cp_async_wait<0>();
    // ragged tail (< R rows; only the sub-range that ends the launch has one)
    {
        const float* srcp = a.src + (int64_t)row * a.src_ld + c0;
        float* outp = a.out + (int64_t)row * a.ld_out + c0;
        for (; row < row_end; ++row) {
            float2 x = *reinterpret_cast<const float2*>(srcp);
#pragma unroll
            for (int s = 0; s < NSEC; ++s) x = delta_step<KIND>(x, sec[s], m4, s == 0, sec[s > 0 ? s - 1 : 0].c, (a.sec_kind[s] & SEC_HP) != 0);
            *reinterpret_cast<float2*>(outp) = __fmul2_rn(x, gain);
            srcp += a.src_ld;
            outp += a.ld_out;
        }
    }
    if (row_end == a.frames) {                // the sub-range that finishes the launch carries the state on
#pragma unroll
        for (int s = 0; s < NSEC; ++s) {
            const float2 g = *reinterpret_cast<const float2*>(a.coef + (size_t)(s * 3 + 0) * C + c0);
            const float2 d = *reinterpret_cast<const float2*>(a.coef + (size_t)(s * 3 + 2) * C + c0);
            double s1, s2;
            delta_state_out<KIND>(g.x, d.x, sec[s].be.x, sec[s].D.x, sec[s].Z.x, s1, s2);
            a.state_out[(size_t)(s * 2 + 0) * C + c0] = s1;
            a.state_out[(size_t)(s * 2 + 1) * C + c0] = s2;
            delta_state_out<KIND>(g.y, d.y, sec[s].be.y, sec[s].D.y, sec[s].Z.y, s1, s2);
            a.state_out[(size_t)(s * 2 + 0) * C + c0 + 1] = s1;
            a.state_out[(size_t)(s * 2 + 1) * C + c0 + 1] = s2;
        }
    }
  }
}

int g_delta_probe = 0;          // A/B (8 low-pass sections only): 0 default geometry (16-row blocks, 2 CTAs/SM); 1: 4-row blocks, 4 CTAs/SM;
                                // 3 / 6 / 7: 8-row blocks, 4 / 3 / 2 CTAs/SM

// high-pass sections keep two coefficient pairs, an output scale and two states (the 10 registers of a low-pass section);
// mixed cascades three coefficient pairs and three states
template <int KIND>
int delta_launch_nsec(const ChainDev* a, dim3 grid, int tiles, int npieces, int warm, cudaStream_t st) {
    switch (a->nsec) {
        case 2: k_cascade_delta<2, 8, delta_min_blocks(2, KIND), 8, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm); break;
        case 3: k_cascade_delta<3, 8, delta_min_blocks(3, KIND), 8, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm); break;
        case 4: k_cascade_delta<4, 8, delta_min_blocks(4, KIND), 8, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm); break;
        case 5: k_cascade_delta<5, 8, delta_min_blocks(5, KIND), 8, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm); break;
        case 6: k_cascade_delta<6, 8, delta_min_blocks(6, KIND), 8, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm); break;
        case 7: k_cascade_delta<7, 16, 2, 16, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm); break;
        default:
            if (KIND == 0 && g_delta_probe == 1) k_cascade_delta<8, 4, 4><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm);
            else if (KIND == 0 && g_delta_probe == 3) k_cascade_delta<8, 8, 4><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm);
            else if (KIND == 0 && g_delta_probe == 6) k_cascade_delta<8, 8, 3><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm);
            else if (KIND == 0 && g_delta_probe == 7) k_cascade_delta<8, 8, 2><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm);
            else k_cascade_delta<8, 16, 2, 16, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm);
            break;
    }
    return (int)cudaGetLastError();
}

template <int NSEC, int KIND, int R, bool FAST>
int reg_launch(const ChainDev* a, dim3 grid, int tiles, int npieces, int warm, cudaStream_t st) {
    k_cascade_reg<NSEC, KIND, R, FAST><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm);
    return (int)cudaGetLastError();
}

template <int KIND, int R, bool FAST>
int reg_launch_nsec(const ChainDev* a, dim3 grid, int tiles, int npieces, int warm, cudaStream_t st) {
    switch (a->nsec) {
        case 3: return reg_launch<3, KIND, R, FAST>(a, grid, tiles, npieces, warm, st);
        case 4: return reg_launch<4, KIND, R, FAST>(a, grid, tiles, npieces, warm, st);
        case 5: return reg_launch<5, KIND, R, FAST>(a, grid, tiles, npieces, warm, st);
        case 6: return reg_launch<6, KIND, R, FAST>(a, grid, tiles, npieces, warm, st);
        case 7: return reg_launch<7, KIND, R, FAST>(a, grid, tiles, npieces, warm, st);
        default: return reg_launch<8, KIND, R, FAST>(a, grid, tiles, npieces, warm, st);
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_osc_reg: the same register-resident cascade fed by an OSCILLATOR evaluated in the thread (config C2's shape:
// osc -> sections -> gain, write-only).  Per 8-row block a lane takes the exact Q0.64 phase of its two channels at
// the block's first row (theta0 + n dtheta mod 2^64, carried by one 64-bit add per block), steps the top word by
// the rounded increment inside the block (drift <= 8 * 2^-33 cycles), evaluates the waveform from the phase word
// (wave_q32) and -- for the discontinuous waveforms -- redoes a column with the reference's float64 arithmetic
// (osc.py:32) when it passes within the guard band of a jump, exactly as k_cascade_pipe's source warp does.
// No scan, no barrier, no shared memory: time pieces with decay warm-up supply the parallelism, and the warm-up
// costs arithmetic only (nothing is read, nothing is stored), of which a write-bound chain has plenty.
// ---------------------------------------------------------------------------------------------------------
constexpr int OR = 8;           // rows per block
// resident CTAs per SM: shallow chains are write-bound and want warps (20 per SM), deep ones need the registers
__host__ __device__ constexpr int osc_min_blocks(int nsec) { return nsec <= 2 ? 5 : nsec <= 6 ? 3 : 2; }

// phase_guard (sigb_plan.cu) for the two channels of a thread at absolute row `last_row`: in-tile drift of the rounded
// increment, rounding of the top word, and the float64 rounding of the reference's own phase
__device__ __forceinline__ int osc_guard(const ChainDev& a, int ca, int cb, int64_t last_row) {
    const double hz = fmax(fabs(a.hertz[ca]), fabs(a.hertz[cb])), ph = fmax(fabs(a.phase[ca]), fabs(a.phase[cb]));
    const double cyc = hz * (double)last_row / (double)a.rate + ph + 1.0;
    const double g = 17.0 + cyc * (3.0 * 4294967296.0 / 9007199254740992.0);
    return g < 1073741823.0 ? (int)g + 1 : 0x3fffffff;
}

template <int WAVE>
__device__ __forceinline__ void osc_rows_g(const ChainDev& a, int guard, int c, unsigned long long th, unsigned long long dth, int64_t n0, float (&x)[OR]) {
    const int w = (int)((th + 0x80000000ull) >> 32), dhi = (int)((dth + 0x80000000ull) >> 32);
    const bool near = gen_tile<WAVE, OR>(w, dhi, guard, x);
    if (WAVE != SIGB_WAVE_SINE && near) {
        const double hz = a.hertz[c], ph = a.phase[c], rate = (double)a.rate;
#pragma unroll
        for (int k = 0; k < OR; ++k) x[k] = osc_wave(WAVE, osc_cycles(__ddiv_rn((double)(n0 + k), rate), hz, ph));
    }
}

// ... with explicit hertz / phase tables (the second oscillator of a fused Mix / RingMod)
template <int WAVE>
__device__ __forceinline__ void osc_rows_p(const double* hertz, const double* phase, int rate, int guard, int c, unsigned long long th,
                                           unsigned long long dth, int64_t n0, float (&x)[OR]) {
    const int w = (int)((th + 0x80000000ull) >> 32), dhi = (int)((dth + 0x80000000ull) >> 32);
    const bool near = gen_tile<WAVE, OR>(w, dhi, guard, x);
    if (WAVE != SIGB_WAVE_SINE && near) {
        const double hz = hertz[c], ph = phase[c], r = (double)rate;
#pragma unroll
        for (int k = 0; k < OR; ++k) x[k] = osc_wave(WAVE, osc_cycles(__ddiv_rn((double)(n0 + k), r), hz, ph));
    }
}

template <int WAVE>
__device__ __forceinline__ void osc_rows(const ChainDev& a, int c, unsigned long long th, unsigned long long dth, int64_t n0, float (&x)[OR]) {
    const int w = (int)((th + 0x80000000ull) >> 32), dhi = (int)((dth + 0x80000000ull) >> 32);
    const bool near = gen_tile<WAVE, OR>(w, dhi, a.guard, x);
    if (WAVE != SIGB_WAVE_SINE && near) {
        const double hz = a.hertz[c], ph = a.phase[c], rate = (double)a.rate;
#pragma unroll
        for (int k = 0; k < OR; ++k) x[k] = osc_wave(WAVE, osc_cycles(__ddiv_rn((double)(n0 + k), rate), hz, ph));
    }
}

template <int NSEC, int KIND, int WAVE>
__global__ void __launch_bounds__(RWARPS * 32, osc_min_blocks(NSEC))
k_osc_reg(const ChainDev a, int tiles, int npieces, int warm_rows, int fast) {
    const int lane = threadIdx.x & 31;
    const int piece = blockIdx.x * RWARPS + (threadIdx.x >> 5);
    if (piece >= npieces) return;
    const size_t C = (size_t)a.C;
    const int bpt = (a.frames + OR - 1) / OR;
    const int64_t total = (int64_t)tiles * bpt;
    int64_t blk = total * piece / npieces;
    const int64_t blk_end = total * (piece + 1) / npieces;
    while (blk < blk_end) {
        const int tile = (int)(blk / bpt);
        const int b0 = (int)(blk - (int64_t)tile * bpt);
        const int b1 = (int)min((int64_t)bpt, b0 + (blk_end - blk));
        blk += b1 - b0;
        const int c0 = tile * RC + 2 * lane;
        const bool live0 = c0 < a.C, live1 = c0 + 1 < a.C;
        const int ca = min(c0, a.C - 1), cb = min(c0 + 1, a.C - 1);
        const int row_store = b0 * OR;
        const int row_end = min(a.frames, b1 * OR);
        const int row_first = max(0, row_store - warm_rows);               // warm_rows is a multiple of OR

        RegSec sec[NSEC];
#pragma unroll
        for (int s = 0; s < NSEC; ++s) {
            load_section(a, s, ca, cb, sec[s]);
            if (row_first == 0) {
                sec[s].s1 = make_float2((float)a.state[(size_t)(s * 2 + 0) * C + ca], (float)a.state[(size_t)(s * 2 + 0) * C + cb]);
                sec[s].s2 = make_float2((float)a.state[(size_t)(s * 2 + 1) * C + ca], (float)a.state[(size_t)(s * 2 + 1) * C + cb]);
            } else {
                sec[s].s1 = sec[s].s2 = make_float2(0.0f, 0.0f);
            }
        }
        float2 gain = make_float2(1.0f, 1.0f);
        if (a.gain) gain = make_float2(a.gain[ca], a.gain[cb]);
        const unsigned long long dtha = a.dtheta[ca], dthb = a.dtheta[cb];
        int64_t n = a.position + row_first;
        unsigned long long tha = a.theta0[ca] + (unsigned long long)n * dtha, thb = a.theta0[cb] + (unsigned long long)n * dthb;
        float* outp = a.out + (int64_t)row_first * a.ld_out + c0;
        const bool vec = fast && live1;

        for (int row = row_first; row < row_end; row += OR) {
            float xa[OR], xb[OR];
            osc_rows<WAVE>(a, ca, tha, dtha, n, xa);
            osc_rows<WAVE>(a, cb, thb, dthb, n, xb);
            tha += (unsigned long long)OR * dtha;
            thb += (unsigned long long)OR * dthb;
            n += OR;
            float2 x[OR];
#pragma unroll
            for (int k = 0; k < OR; ++k) x[k] = make_float2(xa[k], xb[k]);
            if (row + OR <= row_end) {
                reg_block<NSEC, KIND, OR>(x, sec);
                if (row >= row_store) {
                    if (vec) {
#pragma unroll
                        for (int k = 0; k < OR; ++k) __stcs(reinterpret_cast<float2*>(outp + (int64_t)k * a.ld_out), __fmul2_rn(x[k], gain));
                    } else {
#pragma unroll
                        for (int k = 0; k < OR; ++k) {
                            if (live0) outp[(int64_t)k * a.ld_out] = x[k].x * gain.x;
                            if (live1) outp[(int64_t)k * a.ld_out + 1] = x[k].y * gain.y;
                        }
                    }
                }
            } else {
                // ragged last block of the launch: the state stops at the last real row
                for (int k = 0; k < row_end - row; ++k) {
                    float2 y = x[0];
#pragma unroll
                    for (int j = 1; j < OR; ++j) if (j == k) y = x[j];
#pragma unroll
                    for (int s = 0; s < NSEC; ++s) y = reg_step<KIND>(y, sec[s]);
                    if (live0) outp[(int64_t)k * a.ld_out] = y.x * gain.x;
                    if (live1) outp[(int64_t)k * a.ld_out + 1] = y.y * gain.y;
                }
            }
            outp += (int64_t)OR * a.ld_out;
        }
        if (row_end == a.frames) {
#pragma unroll
            for (int s = 0; s < NSEC; ++s) {
                if (live0) {
                    a.state_out[(size_t)(s * 2 + 0) * C + c0] = (double)sec[s].s1.x;
                    a.state_out[(size_t)(s * 2 + 1) * C + c0] = (double)sec[s].s2.x;
                }
                if (live1) {
                    a.state_out[(size_t)(s * 2 + 0) * C + c0 + 1] = (double)sec[s].s1.y;
                    a.state_out[(size_t)(s * 2 + 1) * C + c0 + 1] = (double)sec[s].s2.y;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_osc_delta: k_osc_reg with the sections in DELTA FORM (low-pass 5, high-pass 4, mixed 6 operations + a select per section
// instead of 6 / 7; see k_cascade_delta) for oscillator-fed chains of 2..8 SECOND-ORDER sections of any kinds.  The waveform
// is a run-time switch per block (uniform over the launch), so one instantiation per (sections, kind class) serves all four.
// ---------------------------------------------------------------------------------------------------------
// Resident CTAs per SM = time pieces per SM / 4.  Measured on C2's shape (4,096 voices = 64 tiles x 480,000 rows, where every
// piece pays a warm-up of ~2,400 rows; tools/bench_osc_sections.py, profiles/r02_osc_sections.txt): 2 sections 5 / 3 / 2 CTAs
// per SM = 9.1 / 9.4 / 9.3e11 voice-samples/s, 3 sections 4 / 3 / 2 = 7.1 / 9.0 / 8.8e11, 4 sections 5.8 / 7.4 / 7.3e11 --
// 8 to 12 warps saturate the FP32 pipe, more of them only add warm-up rows and concurrent write streams.
__host__ __device__ constexpr int osc_delta_min_blocks(int nsec, int kind) {
    return nsec <= 4 ? 3 : 2;
}

template <int NSEC, int KIND>
__global__ void __launch_bounds__(RWARPS * 32, osc_delta_min_blocks(NSEC, KIND))
k_osc_delta(const ChainDev a, int tiles, int npieces, int warm_rows, int fast) {
    const int lane = threadIdx.x & 31;
    const int piece = blockIdx.x * RWARPS + (threadIdx.x >> 5);
    if (piece >= npieces) return;
    const size_t C = (size_t)a.C;
    const int bpt = (a.frames + OR - 1) / OR;
    const int64_t total = (int64_t)tiles * bpt;
    int64_t blk = total * piece / npieces;
    const int64_t blk_end = total * (piece + 1) / npieces;
    const float2 m4 = make_float2(-4.0f, -4.0f);
    const int wave = a.wave;
    while (blk < blk_end) {
        const int tile = (int)(blk / bpt);
        const int b0 = (int)(blk - (int64_t)tile * bpt);
        const int b1 = (int)min((int64_t)bpt, b0 + (blk_end - blk));
        blk += b1 - b0;
        const int c0 = tile * RC + 2 * lane;
        const bool live0 = c0 < a.C, live1 = c0 + 1 < a.C;
        const int ca = min(c0, a.C - 1), cb = min(c0 + 1, a.C - 1);
        const int row_store = b0 * OR;
        const int row_end = min(a.frames, b1 * OR);
        const int row_first = max(0, row_store - warm_rows);               // warm_rows is a multiple of OR

        DeltaSec sec[NSEC];
#pragma unroll
        for (int s = 0; s < NSEC; ++s) {
            const float ga = a.coef[(size_t)(s * 3 + 0) * C + ca], gb = a.coef[(size_t)(s * 3 + 0) * C + cb];
            const float cca = a.coef[(size_t)(s * 3 + 1) * C + ca], ccb = a.coef[(size_t)(s * 3 + 1) * C + cb];
            const float da = a.coef[(size_t)(s * 3 + 2) * C + ca], db = a.coef[(size_t)(s * 3 + 2) * C + cb];
            const bool hp_s = (KIND & SEC_HP) || ((KIND & SEC_MIXED) && (a.sec_kind[s] & SEC_HP));
            delta_coef<KIND>(ga, cca, da, hp_s, sec[s].a.x, sec[s].be.x, sec[s].c.x);
            delta_coef<KIND>(gb, ccb, db, hp_s, sec[s].a.y, sec[s].be.y, sec[s].c.y);
            if (row_first == 0) {
                delta_state_in<KIND>(ga, da, sec[s].be.x, a.state[(size_t)(s * 2 + 0) * C + ca], a.state[(size_t)(s * 2 + 1) * C + ca], sec[s].D.x, sec[s].Z.x, sec[s].P.x);
                delta_state_in<KIND>(gb, db, sec[s].be.y, a.state[(size_t)(s * 2 + 0) * C + cb], a.state[(size_t)(s * 2 + 1) * C + cb], sec[s].D.y, sec[s].Z.y, sec[s].P.y);
            } else {
                sec[s].D = sec[s].Z = sec[s].P = make_float2(0.0f, 0.0f);
            }
        }
        float2 gain = make_float2(1.0f, 1.0f);
        if (a.gain) gain = make_float2(a.gain[ca], a.gain[cb]);
        if (KIND & (SEC_HP | SEC_MIXED)) gain = __fmul2_rn(gain, sec[NSEC - 1].c);   // the last section's output scale
        const unsigned long long dtha = a.dtheta[ca], dthb = a.dtheta[cb];
        const bool rot = wave == SIGB_WAVE_SINE && a.rot1 != nullptr;
        float2 rotC = make_float2(1.0f, 1.0f), rotS = make_float2(0.0f, 0.0f);
        if (rot) {
            const float2 ra = a.rot1[ca], rb = a.rot1[cb];
            rotC = make_float2(ra.x, rb.x);
            rotS = make_float2(ra.y, rb.y);
        }
        const int dhiA = (int)((dtha + 0x80000000ull) >> 32), dhiB = (int)((dthb + 0x80000000ull) >> 32);
        // guard band around the waveform's discontinuities from the thread's OWN channels (phase_guard of sigb_plan.cu at the
        // request's last row): needs no host-side maximum, so it also serves oscillators whose hertz / phase are sampled per
        // request on the device (k_osc_tables)
        const int guard = osc_guard(a, ca, cb, a.position + a.frames);
        int64_t n = a.position + row_first;
        unsigned long long tha = a.theta0[ca] + (unsigned long long)n * dtha, thb = a.theta0[cb] + (unsigned long long)n * dthb;
        float* outp = a.out + (int64_t)row_first * a.ld_out + c0;
        const bool vec = fast && live1;

        for (int row = row_first; row < row_end; row += OR) {
            float2 x[OR];
            if (rot) {
                // Sine by k_chain_scan3's two-pipe evaluation: rows 1 and 5 of the block get a sine AND a
                // cosine from the SFU, their neighbours (one row back, two rows forward) the angle-addition rotation by the
                // channel's one-row phase advance (cos, sin tabulated in float64 on the host): 18 instructions per 8 rows x 2
                // channels instead of 40
                int ha = (int)((tha + 0x80000000ull) >> 32), hb = (int)((thb + 0x80000000ull) >> 32);
                const float2 NS = make_float2(-rotS.x, -rotS.y);
#pragma unroll
                for (int q4 = 0; q4 + 3 < OR; q4 += 4) {
                    const float2 r = __fmul2_rn(make_float2((float)(ha + dhiA), (float)(hb + dhiB)), make_float2(kTwoPiQ32, kTwoPiQ32));
                    const float2 S1 = make_float2(__sinf(r.x), __sinf(r.y)), C1 = make_float2(__cosf(r.x), __cosf(r.y));
                    const float2 t = __fmul2_rn(S1, rotC);
                    x[q4 + 0] = __ffma2_rn(C1, NS, t);
                    x[q4 + 1] = S1;
                    const float2 S2 = __ffma2_rn(C1, rotS, t);
                    const float2 C2 = __ffma2_rn(S1, NS, __fmul2_rn(C1, rotC));
                    x[q4 + 2] = S2;
                    x[q4 + 3] = __ffma2_rn(C2, rotS, __fmul2_rn(S2, rotC));
                    ha += 4 * dhiA;
                    hb += 4 * dhiB;
                }
            } else {
                float xa[OR], xb[OR];
                switch (wave) {
                    case SIGB_WAVE_SINE: osc_rows_g<SIGB_WAVE_SINE>(a, guard, ca, tha, dtha, n, xa); osc_rows_g<SIGB_WAVE_SINE>(a, guard, cb, thb, dthb, n, xb); break;
                    case SIGB_WAVE_SQUARE: osc_rows_g<SIGB_WAVE_SQUARE>(a, guard, ca, tha, dtha, n, xa); osc_rows_g<SIGB_WAVE_SQUARE>(a, guard, cb, thb, dthb, n, xb); break;
                    case SIGB_WAVE_SAWTOOTH: osc_rows_g<SIGB_WAVE_SAWTOOTH>(a, guard, ca, tha, dtha, n, xa); osc_rows_g<SIGB_WAVE_SAWTOOTH>(a, guard, cb, thb, dthb, n, xb); break;
                    default: osc_rows_g<SIGB_WAVE_TRIANGLE>(a, guard, ca, tha, dtha, n, xa); osc_rows_g<SIGB_WAVE_TRIANGLE>(a, guard, cb, thb, dthb, n, xb); break;
                }
#pragma unroll
                for (int k = 0; k < OR; ++k) x[k] = make_float2(xa[k], xb[k]);
            }
            tha += (unsigned long long)OR * dtha;
            thb += (unsigned long long)OR * dthb;
            n += OR;
            if (row + OR <= row_end) {
#pragma unroll
                for (int dgl = 0; dgl < OR + NSEC - 1; ++dgl) {
#pragma unroll
                    for (int s = 0; s < NSEC; ++s) {
                        const int r = dgl - s;
                        if (r >= 0 && r < OR) x[r] = delta_step<KIND>(x[r], sec[s], m4, s == 0, sec[s > 0 ? s - 1 : 0].c, (a.sec_kind[s] & SEC_HP) != 0);
                    }
                }
                if (row >= row_store) {
                    if (vec) {
#pragma unroll
                        for (int k = 0; k < OR; ++k) __stcs(reinterpret_cast<float2*>(outp + (int64_t)k * a.ld_out), __fmul2_rn(x[k], gain));
                    } else {
#pragma unroll
                        for (int k = 0; k < OR; ++k) {
                            if (live0) outp[(int64_t)k * a.ld_out] = x[k].x * gain.x;
                            if (live1) outp[(int64_t)k * a.ld_out + 1] = x[k].y * gain.y;
                        }
                    }
                }
            } else {
                // ragged last block of the launch: the state stops at the last real row
                for (int k = 0; k < row_end - row; ++k) {
                    float2 y = x[0];
#pragma unroll
                    for (int j = 1; j < OR; ++j) if (j == k) y = x[j];
#pragma unroll
                    for (int s = 0; s < NSEC; ++s) y = delta_step<KIND>(y, sec[s], m4, s == 0, sec[s > 0 ? s - 1 : 0].c, (a.sec_kind[s] & SEC_HP) != 0);
                    if (live0) outp[(int64_t)k * a.ld_out] = y.x * gain.x;
                    if (live1) outp[(int64_t)k * a.ld_out + 1] = y.y * gain.y;
                }
            }
            outp += (int64_t)OR * a.ld_out;
        }
        if (row_end == a.frames) {
#pragma unroll
            for (int s = 0; s < NSEC; ++s) {
                const float ga = a.coef[(size_t)(s * 3 + 0) * C + ca], gb = a.coef[(size_t)(s * 3 + 0) * C + cb];
                const float da = a.coef[(size_t)(s * 3 + 2) * C + ca], db = a.coef[(size_t)(s * 3 + 2) * C + cb];
                double s1, s2;
                if (live0) {
                    delta_state_out<KIND>(ga, da, sec[s].be.x, sec[s].D.x, sec[s].Z.x, s1, s2);
                    a.state_out[(size_t)(s * 2 + 0) * C + c0] = s1;
                    a.state_out[(size_t)(s * 2 + 1) * C + c0] = s2;
                }
                if (live1) {
                    delta_state_out<KIND>(gb, db, sec[s].be.y, sec[s].D.y, sec[s].Z.y, s1, s2);
                    a.state_out[(size_t)(s * 2 + 0) * C + c0 + 1] = s1;
                    a.state_out[(size_t)(s * 2 + 1) * C + c0 + 1] = s2;
                }
            }
        }
    }
}

template <int KIND>
int osc_delta_launch_nsec(const ChainDev* a, dim3 grid, int tiles, int npieces, int warm, int fast, cudaStream_t st) {
    switch (a->nsec) {
        case 1: if (KIND & SEC_MIXED) return (int)cudaErrorInvalidValue;
                k_osc_delta<1, KIND & SEC_HP><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm, fast); break;
        case 2: k_osc_delta<2, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm, fast); break;
        case 3: k_osc_delta<3, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm, fast); break;
        case 4: k_osc_delta<4, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm, fast); break;
        case 5: k_osc_delta<5, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm, fast); break;
        case 6: k_osc_delta<6, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm, fast); break;
        case 7: k_osc_delta<7, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm, fast); break;
        default: k_osc_delta<8, KIND><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm, fast); break;
    }
    return (int)cudaGetLastError();
}

template <int NSEC, int KIND>
int osc_launch(const ChainDev* a, dim3 grid, int tiles, int npieces, int warm, int fast, cudaStream_t st) {
    switch (a->wave) {
        case SIGB_WAVE_SINE: k_osc_reg<NSEC, KIND, SIGB_WAVE_SINE><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm, fast); break;
        case SIGB_WAVE_SQUARE: k_osc_reg<NSEC, KIND, SIGB_WAVE_SQUARE><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm, fast); break;
        case SIGB_WAVE_SAWTOOTH: k_osc_reg<NSEC, KIND, SIGB_WAVE_SAWTOOTH><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm, fast); break;
        default: k_osc_reg<NSEC, KIND, SIGB_WAVE_TRIANGLE><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, npieces, warm, fast); break;
    }
    return (int)cudaGetLastError();
}

template <int KIND>
int osc_launch_nsec(const ChainDev* a, dim3 grid, int tiles, int npieces, int warm, int fast, cudaStream_t st) {
    switch (a->nsec) {
        case 1: return osc_launch<1, KIND>(a, grid, tiles, npieces, warm, fast, st);
        case 2: return osc_launch<2, KIND>(a, grid, tiles, npieces, warm, fast, st);
        case 3: return osc_launch<3, KIND>(a, grid, tiles, npieces, warm, fast, st);
        case 4: return osc_launch<4, KIND>(a, grid, tiles, npieces, warm, fast, st);
        case 5: return osc_launch<5, KIND>(a, grid, tiles, npieces, warm, fast, st);
        case 6: return osc_launch<6, KIND>(a, grid, tiles, npieces, warm, fast, st);
        case 7: return osc_launch<7, KIND>(a, grid, tiles, npieces, warm, fast, st);
        default: return osc_launch<8, KIND>(a, grid, tiles, npieces, warm, fast, st);
    }
}

}  // namespace

// Whether the register-resident kernel can take this chain: a static property of the chain (never of a
// particular call's pointers): a materialised source and 3..8 second-order sections of one kind.
extern "C" void sigb_set_reg_pieces(int n) { g_reg_pieces = n; }
extern "C" void sigb_set_delta_probe(int n) { g_delta_probe = n; }

// 8-byte loads / stores and no ragged edges: unit channel stride, even leading dimensions, 8-byte aligned bases,
// whole 64-channel tiles, a source that covers every row
static bool reg_fast_layout(const ChainDev* a) {
    return a->src_cs == 1 && (reinterpret_cast<uintptr_t>(a->src) & 7) == 0 && (a->src_ld & 1) == 0 &&
           (reinterpret_cast<uintptr_t>(a->out) & 7) == 0 && (a->ld_out & 1) == 0 && a->C % RC == 0 &&
           a->src_rows >= (int64_t)a->frames && a->src_ld > 0 && a->src_ld < (1 << 26) && a->ld_out > 0 && a->ld_out < (1 << 26);
}

extern "C" int sigb_cascade_reg_ok(const ChainDev* a) {
    if (a->src_kind != SRC_BUF || a->nsec < 2 || a->nsec > 8 || a->C <= 0) return 0;
    bool mixed = false, any_first = false;
    for (int k = 0; k < a->nsec; ++k) {
        mixed |= (a->sec_kind[k] & SEC_HP) != (a->sec_kind[0] & SEC_HP);
        any_first |= (a->sec_kind[k] & SEC_FIRST_ORDER) != 0;
    }
    // one kind: always (first-order sections are welcome, load_section); low- and high-pass sections mixed: k_cascade_delta
    // only, i.e. second-order sections on the aligned whole-tile layout of THIS call (k_cascade_pipe takes the rest)
    if (mixed || a->nsec == 2) return !any_first && reg_fast_layout(a);     // (two sections: the planner also asks sigb_cascade_reg_fill)
    return 1;
}

// variant 0 (default): blocks of 8 rows; variant 1: blocks of 4 rows (measured on C4: 4.59e11 vs 4.48e11
// channel-samples/s).  max_segments bounds the pieces per tile (1: never cut along time).
struct RegGeometry { bool fast, wide, delta, mixed; int R, tiles, npieces, warm; int64_t slots; };

static RegGeometry cascade_reg_geometry(const ChainDev* a, int max_segments, int variant) {
    RegGeometry q;
    q.fast = reg_fast_layout(a);
    bool any_first = false;
    q.mixed = false;
    for (int k = 0; k < a->nsec; ++k) {
        any_first |= (a->sec_kind[k] & SEC_FIRST_ORDER) != 0;
        q.mixed |= (a->sec_kind[k] & SEC_HP) != (a->sec_kind[0] & SEC_HP);
    }
    if (q.mixed || a->nsec == 2) variant = 0;    // only k_cascade_delta runs both kinds in one cascade, or two sections (sigb_cascade_reg_ok checked the layout)
    q.wide = q.fast && variant != 1;
    // delta form (5 operations per low-pass section, 4 per high-pass section, instead of 6 / 7): second-order sections only;
    // variant 4 keeps the state-variable form in 8-row blocks for A/B
    q.delta = q.fast && variant != 1 && variant != 4 && !any_first;
    const int dprobe = (q.delta && a->nsec == 8 && !q.mixed && !(a->sec_kind[0] & SEC_HP)) ? g_delta_probe : 0;
    q.R = (q.delta && a->nsec >= 7 && dprobe == 0) ? 16 : (q.wide && dprobe != 1) ? 8 : 4;     // 7 and 8 sections: 16-row blocks (delta_min_blocks)
    const int R = q.R;
    q.tiles = (a->C + RC - 1) / RC;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int warps_per_sm = (q.delta ? (dprobe == 1 || dprobe == 3 ? 4 : dprobe == 6 ? 3 : delta_min_blocks(a->nsec, q.mixed ? SEC_MIXED : 0)) : reg_min_blocks(a->nsec, R)) * RWARPS;
    // pieces: one per warp slot of the machine, as long as the warm-up of a piece that starts inside a tile stays
    // below 1/4 of the piece; never fewer than one per tile
    const int bpt = (a->frames + R - 1) / R;
    q.slots = (int64_t)sms * warps_per_sm * std::max(1, g_reg_pieces) * g_osc_pieces_pct / 100;
    int64_t want = q.tiles;
    if (max_segments > 1 && a->warm_rows >= 0) {
        q.warm = (a->warm_rows + R - 1) / R * R;
        const int64_t fit = (int64_t)q.tiles * bpt / std::max(1, 4 * q.warm / R);
        want = std::max<int64_t>(q.tiles, std::min<int64_t>(std::min(q.slots, fit), (int64_t)q.tiles * max_segments));
    } else {
        q.warm = bpt * R;            // unknown decay: a piece never starts inside a tile (npieces == tiles)
    }
    q.npieces = (int)want;
    return q;
}

// share of the machine's warp slots the launch would occupy, in 1/1024 (see sigb_osc_reg_fill)
extern "C" int sigb_cascade_reg_fill(const ChainDev* a, int max_segments, int variant) {
    const RegGeometry q = cascade_reg_geometry(a, max_segments, variant);
    return (int)std::min<int64_t>(1024, (int64_t)q.npieces * 1024 / std::max<int64_t>(1, q.slots));
}

extern "C" int sigb_launch_cascade_reg(const ChainDev* a, int max_segments, int variant, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (a->frames <= 0) return 0;
    const RegGeometry q = cascade_reg_geometry(a, max_segments, variant);
    const bool fast = q.fast, wide = q.wide, delta = q.delta, mixed = q.mixed;
    const int tiles = q.tiles, npieces = q.npieces, warm = q.warm;
    const dim3 grid((unsigned)((npieces + RWARPS - 1) / RWARPS));
    const bool hp = (a->sec_kind[0] & SEC_HP) != 0;
    if (delta) return mixed ? delta_launch_nsec<SEC_MIXED>(a, grid, tiles, npieces, warm, st)
                      : (a->sec_kind[0] & SEC_HP) ? delta_launch_nsec<SEC_HP>(a, grid, tiles, npieces, warm, st)
                                                  : delta_launch_nsec<0>(a, grid, tiles, npieces, warm, st);
    if (wide) return hp ? reg_launch_nsec<SEC_HP, 8, true>(a, grid, tiles, npieces, warm, st)
                        : reg_launch_nsec<0, 8, true>(a, grid, tiles, npieces, warm, st);
    if (fast) return hp ? reg_launch_nsec<SEC_HP, 4, true>(a, grid, tiles, npieces, warm, st)
                        : reg_launch_nsec<0, 4, true>(a, grid, tiles, npieces, warm, st);
    return hp ? reg_launch_nsec<SEC_HP, 4, false>(a, grid, tiles, npieces, warm, st)
              : reg_launch_nsec<0, 4, false>(a, grid, tiles, npieces, warm, st);
}

// ---------------------------------------------------------------------------------------------------------
// k_osc_fill: STATELESS oscillator chains (osc -> gain, no filter) on many channels.  k_chain_seq evaluates every sample of
// such a chain in float64 (division, multiply-add, waveform): 6.0e11 voice-samples/s on C2's shape, 0.37 of the HBM
// roofline for the simplest graph there is.  Here a thread owns two adjacent channels, takes the exact Q0.64 phase at the
// first row of every 8-row block (theta0 + n dtheta mod 2^64) and generates the block from the phase word -- sine rows by
// rotation (one sin/cos pair per four rows), the discontinuous waveforms from the word with the float64 redo inside the
// guard band -- exactly as k_osc_delta's source does.
//
// Block invariance (the reference's oscillators do not depend on block boundaries, osc.py:26-33): the 8-row blocks are
// aligned to the ABSOLUTE sample index (n mod 8 == 0), the guard band is a function of the block's own position, so a
// sample is computed by the same instructions on the same operands whatever request it falls into: requests cut anywhere
// give the same bits.  A partial block at either end of a request is generated whole and stored in part.
// ---------------------------------------------------------------------------------------------------------
template <int WAVE>
__global__ void __launch_bounds__(RWARPS * 32, 4)
k_osc_fill(const ChainDev a, int tiles, int chunks, int blocks_per_chunk, int fast) {
    const int lane = threadIdx.x & 31;
    const int wid = blockIdx.x * RWARPS + (threadIdx.x >> 5);
    if (wid >= tiles * chunks) return;
    const int tile = wid / chunks, chunk = wid - tile * chunks;
    const int64_t pos = a.position, end = a.position + a.frames;
    const int64_t nb_first = pos >> 3, nb_last = (end + 7) >> 3;             // absolute 8-row blocks [nb_first, nb_last)
    int64_t nb = nb_first + (int64_t)chunk * blocks_per_chunk;
    const int64_t nb_end = min(nb_last, nb + blocks_per_chunk);
    if (nb >= nb_end) return;
    const int c0 = tile * RC + 2 * lane;
    const bool live0 = c0 < a.C, live1 = c0 + 1 < a.C;
    const int ca = min(c0, a.C - 1), cb = min(c0 + 1, a.C - 1);
    float2 gain = make_float2(1.0f, 1.0f);
    if (a.gain) gain = make_float2(a.gain[ca], a.gain[cb]);
    const unsigned long long dtha = a.dtheta[ca], dthb = a.dtheta[cb];
    const bool rot = WAVE == SIGB_WAVE_SINE && a.rot1 != nullptr;
    float2 rotC = make_float2(1.0f, 1.0f), rotS = make_float2(0.0f, 0.0f);
    if (rot) {
        const float2 ra = a.rot1[ca], rb = a.rot1[cb];
        rotC = make_float2(ra.x, rb.x);
        rotS = make_float2(ra.y, rb.y);
    }
    const int dhiA = (int)((dtha + 0x80000000ull) >> 32), dhiB = (int)((dthb + 0x80000000ull) >> 32);
    int64_t n = nb << 3;
    unsigned long long tha = a.theta0[ca] + (unsigned long long)n * dtha, thb = a.theta0[cb] + (unsigned long long)n * dthb;
    const bool vec = fast && live1;
    // fused Mix / RingMod with a second oscillator (fx.py:35-46): its phase words ride along in registers
    const int epi = a.epi_op;
    unsigned long long eta = 0, etb = 0, edta = 0, edtb = 0;
    float2 g2 = make_float2(1.0f, 1.0f), mixp = make_float2(0.0f, 0.0f);
    if (epi) {
        edta = a.epi_dtheta[ca]; edtb = a.epi_dtheta[cb];
        eta = a.epi_theta0[ca] + (unsigned long long)n * edta;
        etb = a.epi_theta0[cb] + (unsigned long long)n * edtb;
        if (a.epi_gain) g2 = make_float2(a.epi_gain[ca], a.epi_gain[cb]);
        if (a.epi_p) mixp = make_float2(a.epi_p[ca], a.epi_p[cb]);
    }
    for (; nb < nb_end; ++nb) {
        float2 x[OR];
        if (rot) {
            int ha = (int)((tha + 0x80000000ull) >> 32), hb = (int)((thb + 0x80000000ull) >> 32);
            const float2 NS = make_float2(-rotS.x, -rotS.y);
#pragma unroll
            for (int q4 = 0; q4 + 3 < OR; q4 += 4) {
                const float2 r = __fmul2_rn(make_float2((float)(ha + dhiA), (float)(hb + dhiB)), make_float2(kTwoPiQ32, kTwoPiQ32));
                const float2 S1 = make_float2(__sinf(r.x), __sinf(r.y)), C1 = make_float2(__cosf(r.x), __cosf(r.y));
                const float2 t = __fmul2_rn(S1, rotC);
                x[q4 + 0] = __ffma2_rn(C1, NS, t);
                x[q4 + 1] = S1;
                const float2 S2 = __ffma2_rn(C1, rotS, t);
                const float2 C2 = __ffma2_rn(S1, NS, __fmul2_rn(C1, rotC));
                x[q4 + 2] = S2;
                x[q4 + 3] = __ffma2_rn(C2, rotS, __fmul2_rn(S2, rotC));
                ha += 4 * dhiA;
                hb += 4 * dhiB;
            }
        } else {
            // guard band of THIS block (phase_guard at the block's last row, from the thread's own channels): a function of the
            // absolute position only
            const int guard = osc_guard(a, ca, cb, n + OR);
            float xa[OR], xb[OR];
            osc_rows_g<WAVE>(a, guard, ca, tha, dtha, n, xa);
            osc_rows_g<WAVE>(a, guard, cb, thb, dthb, n, xb);
#pragma unroll
            for (int k = 0; k < OR; ++k) x[k] = make_float2(xa[k], xb[k]);
        }
#pragma unroll
        for (int k = 0; k < OR; ++k) x[k] = __fmul2_rn(x[k], gain);
        if (epi) {
            // the second oscillator's block, then out = mix * left + (1 - mix) * right or left * right (k_chain_seq's epilogue)
            int g2nd;
            {
                const double hz = fmax(fabs(a.epi_hertz[ca]), fabs(a.epi_hertz[cb])), ph = fmax(fabs(a.epi_phase[ca]), fabs(a.epi_phase[cb]));
                const double cyc = hz * (double)(n + OR) / (double)a.rate + ph + 1.0;
                const double gg = 17.0 + cyc * (3.0 * 4294967296.0 / 9007199254740992.0);
                g2nd = gg < 1073741823.0 ? (int)gg + 1 : 0x3fffffff;
            }
            float ya[OR], yb[OR];
            switch (a.epi_wave) {
                case SIGB_WAVE_SINE:
                    osc_rows_p<SIGB_WAVE_SINE>(a.epi_hertz, a.epi_phase, a.rate, g2nd, ca, eta, edta, n, ya);
                    osc_rows_p<SIGB_WAVE_SINE>(a.epi_hertz, a.epi_phase, a.rate, g2nd, cb, etb, edtb, n, yb);
                    break;
                case SIGB_WAVE_SQUARE:
                    osc_rows_p<SIGB_WAVE_SQUARE>(a.epi_hertz, a.epi_phase, a.rate, g2nd, ca, eta, edta, n, ya);
                    osc_rows_p<SIGB_WAVE_SQUARE>(a.epi_hertz, a.epi_phase, a.rate, g2nd, cb, etb, edtb, n, yb);
                    break;
                case SIGB_WAVE_SAWTOOTH:
                    osc_rows_p<SIGB_WAVE_SAWTOOTH>(a.epi_hertz, a.epi_phase, a.rate, g2nd, ca, eta, edta, n, ya);
                    osc_rows_p<SIGB_WAVE_SAWTOOTH>(a.epi_hertz, a.epi_phase, a.rate, g2nd, cb, etb, edtb, n, yb);
                    break;
                default:
                    osc_rows_p<SIGB_WAVE_TRIANGLE>(a.epi_hertz, a.epi_phase, a.rate, g2nd, ca, eta, edta, n, ya);
                    osc_rows_p<SIGB_WAVE_TRIANGLE>(a.epi_hertz, a.epi_phase, a.rate, g2nd, cb, etb, edtb, n, yb);
                    break;
            }
            eta += (unsigned long long)OR * edta;
            etb += (unsigned long long)OR * edtb;
            const float2 one_m = make_float2(1.0f - mixp.x, 1.0f - mixp.y);
#pragma unroll
            for (int k = 0; k < OR; ++k) {
                const float2 o = __fmul2_rn(make_float2(ya[k], yb[k]), g2);
                const float2 left = a.epi_side ? o : x[k], right = a.epi_side ? x[k] : o;
                x[k] = epi == EW_MIX ? __fadd2_rn(__fmul2_rn(mixp, left), __fmul2_rn(one_m, right)) : __fmul2_rn(left, right);
            }
        }
        float* outp = a.out + (n - pos) * a.ld_out + c0;                    // row n of the stream (may lie before the request)
        if (n >= pos && n + OR <= end && vec) {
#pragma unroll
            for (int k = 0; k < OR; ++k) __stcs(reinterpret_cast<float2*>(outp + (int64_t)k * a.ld_out), x[k]);
        } else {
#pragma unroll
            for (int k = 0; k < OR; ++k) {
                if (n + k >= pos && n + k < end) {
                    if (live0) outp[(int64_t)k * a.ld_out] = x[k].x;
                    if (live1) outp[(int64_t)k * a.ld_out + 1] = x[k].y;
                }
            }
        }
        tha += (unsigned long long)OR * dtha;
        thb += (unsigned long long)OR * dthb;
        n += OR;
    }
}

// Stateless oscillator chains (no filter; a fused Mix / RingMod with a second oscillator is welcome) from 128 channels on, Q0.64 phase tables present (built by the
// host for constant hertz / phase, by k_osc_tables per request for modulated ones): a static property of the chain, so that
// every request of a plan takes the same kernel (block invariance).
extern "C" int sigb_osc_fill_ok(const ChainDev* a) {
    if (a->epi_op && !(a->epi_wave >= 0 && a->epi_theta0 && a->epi_dtheta && a->epi_hertz && a->epi_phase && (a->epi_op != EW_MIX || a->epi_p)))
        return 0;                                  // a fused Mix / RingMod only with a second oscillator (not a materialised block)
    return a->src_kind == SRC_OSC && a->nsec == 0 && a->theta0 && a->dtheta && a->hertz && a->phase &&
           a->pos_ptr == nullptr && a->C >= 128;
}

extern "C" int sigb_launch_osc_fill(const ChainDev* a, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (a->frames <= 0) return 0;
    const int fast = (reinterpret_cast<uintptr_t>(a->out) & 7) == 0 && (a->ld_out & 1) == 0;
    const int tiles = (a->C + RC - 1) / RC;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t nblk = ((a->position + a->frames + 7) >> 3) - (a->position >> 3);
    // ~2 waves of the 16 resident warps per SM, each warp streaming its own run of blocks (at least 16 blocks per warp)
    int64_t chunks = std::max<int64_t>(1, std::min<int64_t>(((int64_t)sms * 32 + tiles - 1) / std::max(1, tiles), (nblk + 15) / 16));
    const int bpc = (int)((nblk + chunks - 1) / chunks);
    chunks = (nblk + bpc - 1) / bpc;
    const int64_t warps = (int64_t)tiles * chunks;
    const dim3 grid((unsigned)((warps + RWARPS - 1) / RWARPS));
    switch (a->wave) {
        case SIGB_WAVE_SINE: k_osc_fill<SIGB_WAVE_SINE><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, (int)chunks, bpc, fast); break;
        case SIGB_WAVE_SQUARE: k_osc_fill<SIGB_WAVE_SQUARE><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, (int)chunks, bpc, fast); break;
        case SIGB_WAVE_SAWTOOTH: k_osc_fill<SIGB_WAVE_SAWTOOTH><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, (int)chunks, bpc, fast); break;
        default: k_osc_fill<SIGB_WAVE_TRIANGLE><<<grid, RWARPS * 32, 0, st>>>(*a, tiles, (int)chunks, bpc, fast); break;
    }
    return (int)cudaGetLastError();
}

// Oscillator-fed chains: 1..8 sections, unmodulated oscillator (Q0.64 phase tables present).  Second-order sections of any
// kinds run in delta form (k_osc_delta, from 2 sections); chains with first-order sections (odd Butterworth orders) need the
// state-variable kernel and one kind.
static void osc_chain_kinds(const ChainDev* a, bool& mixed, bool& any_first) {
    mixed = any_first = false;
    for (int k = 0; k < a->nsec; ++k) {
        mixed |= (a->sec_kind[k] & SEC_HP) != (a->sec_kind[0] & SEC_HP);
        any_first |= (a->sec_kind[k] & SEC_FIRST_ORDER) != 0;
    }
}

static bool osc_use_delta_static(const ChainDev* a, bool mixed, bool any_first, int allow_delta) {
    return (allow_delta || mixed) && !any_first;
}

// allow_delta = 0: state-variable sections (k_osc_reg) even where the delta form applies (plan option "osc_delta", A/B)
extern "C" int sigb_osc_reg_ok(const ChainDev* a, int allow_delta) {
    if (a->src_kind != SRC_OSC || !a->theta0 || !a->dtheta || a->nsec < 1 || a->nsec > 8 || a->C <= 0) return 0;
    bool mixed, any_first;
    osc_chain_kinds(a, mixed, any_first);
    if (mixed) return !any_first && allow_delta && a->nsec >= 2;
    // a discontinuous waveform with device-sampled hertz / phase: only k_osc_delta derives the guard band from its own channels
    if (a->osc_mod && a->wave != SIGB_WAVE_SINE && !osc_use_delta_static(a, mixed, any_first, allow_delta)) return 0;
    return 1;
}

extern "C" void sigb_set_osc_pieces_pct(int n) { g_osc_pieces_pct = std::max(1, n); }

// launch geometry shared by the launch and by the planner's "does it fill the machine" question
static void osc_reg_geometry(const ChainDev* a, int max_segments, bool delta, int kind, int& tiles, int& npieces, int& warm, int64_t& slots) {
    tiles = (a->C + RC - 1) / RC;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int warps_per_sm = (delta ? osc_delta_min_blocks(a->nsec, kind) : osc_min_blocks(a->nsec)) * RWARPS;
    slots = (int64_t)sms * warps_per_sm * g_osc_pieces_pct / 100;
    const int bpt = (a->frames + OR - 1) / OR;
    int64_t want = tiles;
    if (max_segments > 1 && a->warm_rows >= 0) {
        warm = (a->warm_rows + OR - 1) / OR * OR;
        const int64_t fit = (int64_t)tiles * bpt / std::max(1, 4 * warm / OR);
        want = std::max<int64_t>(tiles, std::min<int64_t>(std::min(slots, fit), (int64_t)tiles * max_segments));
    } else {
        warm = bpt * OR;
    }
    npieces = (int)want;
}

static bool osc_use_delta(const ChainDev* a, bool mixed, bool any_first, int allow_delta) {
    return osc_use_delta_static(a, mixed, any_first, allow_delta);
}

// share of the machine's warp slots the launch would occupy, in 1/1024 (shallow chains only go register-resident when the
// time pieces their decay horizon allows fill the machine; otherwise the time-parallel scan kernels are the better choice)
extern "C" int sigb_osc_reg_fill(const ChainDev* a, int max_segments, int allow_delta) {
    bool mixed, any_first;
    osc_chain_kinds(a, mixed, any_first);
    const bool delta = osc_use_delta(a, mixed, any_first, allow_delta);
    int tiles, npieces, warm;
    int64_t slots;
    osc_reg_geometry(a, max_segments, delta, mixed ? SEC_MIXED : (a->sec_kind[0] & SEC_HP), tiles, npieces, warm, slots);
    return (int)std::min<int64_t>(1024, (int64_t)npieces * 1024 / std::max<int64_t>(1, slots));
}

extern "C" int sigb_launch_osc_reg(const ChainDev* a, int max_segments, int allow_delta, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (a->frames <= 0) return 0;
    const int fast = (reinterpret_cast<uintptr_t>(a->out) & 7) == 0 && (a->ld_out & 1) == 0;
    bool mixed, any_first;
    osc_chain_kinds(a, mixed, any_first);
    const bool delta = osc_use_delta(a, mixed, any_first, allow_delta);
    const bool hp = (a->sec_kind[0] & SEC_HP) != 0;
    int tiles, npieces, warm;
    int64_t slots;
    osc_reg_geometry(a, max_segments, delta, mixed ? SEC_MIXED : (hp ? SEC_HP : 0), tiles, npieces, warm, slots);
    const dim3 grid((unsigned)((npieces + RWARPS - 1) / RWARPS));
    if (delta) return mixed ? osc_delta_launch_nsec<SEC_MIXED>(a, grid, tiles, npieces, warm, fast, st)
                      : hp  ? osc_delta_launch_nsec<SEC_HP>(a, grid, tiles, npieces, warm, fast, st)
                            : osc_delta_launch_nsec<0>(a, grid, tiles, npieces, warm, fast, st);
    return hp ? osc_launch_nsec<SEC_HP>(a, grid, tiles, npieces, warm, fast, st)
              : osc_launch_nsec<0>(a, grid, tiles, npieces, warm, fast, st);
}
