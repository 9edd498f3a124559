// Internal structures shared by the host planner and the kernels of libsigb200.so.
#pragma once
#include <stdint.h>
#include <vector_types.h>

#define SIGB_MAX_SEC 16       // sections per fused chain launch (longer cascades are split)
#ifndef SIGB_SCAN_L
#define SIGB_SCAN_L 16        // rows per sub-chunk in the time-parallel scan kernel
#endif

enum { SRC_OSC = 0, SRC_BUF = 1, SRC_CONST = 2 };

// section kind bits (uniform over the channels of a chain)
enum { SEC_HP = 1, SEC_FIRST_ORDER = 2 };

enum { EW_COPY = 0, EW_GAIN = 1, EW_MIX = 2, EW_RINGMOD = 3, EW_AMP = 4 };

// One fused per-channel chain:  source -> nsec state-variable sections -> gain -> store.
// All per-channel tables are dense arrays of C entries (the host replicates broadcasts).
struct ChainDev {
    int32_t C;
    int32_t src_kind;
    int32_t wave;
    int32_t nsec;
    int32_t rate;
    int32_t frames;                 // rows this launch covers
    int32_t warm_rows;              // rows after which the filters forget their initial state (< 2^-40); -1: unknown
    int32_t warm_est;               // modulated cutoffs: the host's estimate of the horizon (previous request), -1: none
    int32_t n_warm_dev;             // ... and the per-filter horizons k_design wrote for THIS request (scan kernels read them)
    const int* warm_dev;
    int32_t guard;                  // phase-word guard band around waveform discontinuities (host maximum; constant oscillators only)
    int32_t osc_mod;                // hertz / phase sampled per request on the device (k_osc_tables): `guard` is unknown, only the
                                    // kernels that derive the band from their own channels may take a discontinuous waveform
    int64_t position;               // absolute index of row 0
    const int64_t* pos_ptr;         // realtime graphs (k_chain_seq only): when set, row 0 is *pos_ptr -- a block header the
                                    // host rewrites before every launch of the captured CUDA graph
    uint8_t sec_kind[SIGB_MAX_SEC];
    // SRC_OSC
    const double* hertz;            // [C]
    const double* phase;            // [C]
    const unsigned long long* theta0;   // [C] frac(phase)    in Q0.64 (sine fast path)
    const unsigned long long* dtheta;   // [C] frac(hertz/rate) in Q0.64
    const float2* rot1;                 // [C] (cos, sin) of the one-row phase advance 2 pi frac(hertz / rate) (k_chain_scan3's rotation)
    // SRC_BUF: src[row*src_ld + c*src_cs]; rows >= src_rows read as zero
    const float* src;
    int64_t src_ld;
    int32_t src_cs;
    int64_t src_rows;
    // SRC_CONST
    const float* constv;            // [C]
    // sections: coef[(s*3+k)*C + c], k: 0=g 1=c(=r2+g) 2=d(=1/(1+r2 g+g^2));
    //           first-order: 0=G(=g/(1+g))
    const float* coef;
    const float* gain;              // [C] or nullptr
    double* state;                  // [(s*2+k)*C + c] integrator states (carried across launches)
    double* state_out;              // k_chain_scan2 writes end states here (a piece of another CTA may still read `state`)
    // scan helpers (time-parallel kernel): transition A^L per section, and the zero-input
    // output response of each state over one sub-chunk
    const double* apow;             // [(s*4+k)*C + c], row-major 2x2
    const double* apow_h;           // [(s*4+k)*C + c]  A^(L/2) in float64 (channel-pair kernel: carry chained per 8-row sub-chunk)
    const float* ztab;              // [((s*L + k)*2 + j)*C + c]
    const float* m8;                // [(s*4+k)*C + c]  float32 A^(L/2) (packed kernel: stitches the two halves)
    const float* hrec;              // [(s*4+k)*C + c]  zero-input output recurrence in delta form: k: 0 = det(A), 1 = tr(A) - 1 - det(A),
                                    //                  2, 3 = first difference of the zero-input output per unit state
    float* out;
    int64_t ld_out;
    // epilogue fused into a stateless chain (k_chain_seq): out = op(chain value, other operand), fx.py:35-46.
    // The other operand is a materialised block (epi_buf) or a second oscillator evaluated in registers.
    int32_t epi_op;                 // 0 none, EW_MIX, EW_RINGMOD
    int32_t epi_side;               // 0: the chain is `left`, 1: the chain is `right`
    const float* epi_p;             // EW_MIX: mix[C]
    const float* epi_buf; int64_t epi_ld; int32_t epi_cs; int64_t epi_rows;
    int32_t epi_wave;               // second oscillator: SIGB_WAVE_*, -1 when the other operand is epi_buf
    const double* epi_hertz;        // [C]
    const double* epi_phase;        // [C]
    const float* epi_gain;          // [C] or nullptr
    const unsigned long long* epi_theta0;   // [C] Q0.64 phase / increment of the second oscillator (k_osc_fill)
    const unsigned long long* epi_dtheta;
};

struct EwiseDev {
    int32_t op;
    int32_t C;
    int32_t frames;
    float* out; int64_t ld_out;
    const float* a; int64_t lda; int32_t acs; int64_t a_rows;   // a_rows<0: unlimited
    const float* b; int64_t ldb; int32_t bcs; int64_t b_rows;
    const float* p;                 // per-channel parameter [C] (gain / mix / exp) or nullptr
};

struct ReduceDev {
    int32_t C;          // input channels
    int32_t groups;     // GROUPSUM: output channels; PANSUM: 2
    int32_t frames;
    int32_t pan;        // 0 = group sum, 1 = pan sum
    const float* in; int64_t ld_in; int32_t ics; int64_t in_rows;
    const float* w;     // PANSUM: pan[C]
    float* out; int64_t ld_out;
};


// Oscillator bank fused with GroupSum (config C3): out[n][g] = sum_{p in group g} gain[p] * sin(2 pi theta_p(n)),
// theta_p(n) = theta0[p] + (position + n) * dtheta[p] in Q0.64.
struct BankDev {
    int32_t P;                          // partials = input channels
    int32_t groups;                     // output channels; partial p belongs to group p / (P/groups)
    int32_t frames;
    int64_t position;
    const unsigned long long* theta0;   // [P]
    const unsigned long long* dtheta;   // [P]
    const float* gain;                  // [P] or nullptr (unit amplitude)
    const float2* rot32;                // [P] (cos, sin) of the phase advance over 32 rows, 2 pi frac(32 hertz / rate)
    float* out;
    int64_t ld_out;
};

// One homogeneous segment of a voice bank fused with PanSum (config C5): per channel
// osc(wave) -> at most one filter section -> gain -> (L, R) weights; channels are summed.
struct VoiceSeg {
    int32_t C;
    int32_t wave;                       // SIGB_WAVE_*
    int32_t nsec;                       // 0 or 1
    int32_t sec_kind;                   // SEC_* bits of the section
    int32_t cta0;                       // first voice group (= partial index) of this segment in the launch
    int32_t guard;                      // phase-word guard band around waveform discontinuities
    const unsigned long long* theta0;   // [C] Q0.64 phase / increment (all waves)
    const unsigned long long* dtheta;
    const double* hertz;                // [C] reference float64 path near discontinuities
    const double* phase;
    const float* coef;                  // [(k)*C + c], k: g c d
    const float* wl;                    // [C] gain * (1 - pan)
    const float* wr;                    // [C] gain * pan
    double* state;                      // [(k)*C + c]  read by the CTAs of the first time segment
    double* state_out;                  // written by the CTAs of the last time segment (the other copy)
};

// Block-rate parameter program (modulated parameters): evaluated in float64 for ONE frame per request, at the
// request's position, before the block is rendered (BoundPort.forward_at_block_rate, chain/__init__.py:305-306).
enum { PRM_OSC = 1, PRM_MUL = 2, PRM_MIX = 3, PRM_AMP = 4, PRM_COPY = 5 };
struct ParamInstr {
    int32_t op;        // PRM_*
    int32_t wave;      // PRM_OSC: SIGB_WAVE_*
    int32_t dst;       // row written
    int32_t a, b, c;   // operand rows (OSC: hertz, phase; MUL: left, right; MIX: left, right, mix; AMP: left, exp)
    int32_t width;     // channels of dst (operands are 1 or `width` wide)
    int32_t wa, wb, wc;
};
// Per-request design of ONE filter whose cutoff is modulated (k_design): sections [s0, s0 + ceil(order / 2)) of a chain
// of C channels.  coef is always written; the scan tables only when apow != nullptr.
struct DesignDev {
    int32_t C, s0, order, highpass, rate;
    const double* cutoff;               // [C] the parameter-program row
    float* coef;                        // chain tables, laid out as in ChainDev
    double* apow; double* apow_h; float* ztab; float* m8; float* hrec;
    int* err_flag;                      // set to 1 when a cutoff leaves (0, rate/2) (scipy's ValueError, fx.py:102)
    int* warm_out;                      // atomicMax: decay horizon of this filter in rows (slowest channel)
};

#define SIGB_PARAM_ROWS 96              // rows per parameter program (values live in a per-thread array)

#define SIGB_VOICE_SEGS 24              // segments per k_voices launch
#define SIGB_VOICE_THREADS 256

struct VoicesDev {
    int32_t nseg;
    int32_t rate;
    int32_t frames;
    int32_t M;                          // channels per thread (1 or 4)
    int32_t ngroups;                    // CTA-sized voice groups (= rows of `partial`) in this launch
    int32_t npieces;                    // CTAs: equal contiguous pieces of the (group, row block) space, group-major
    int32_t warm_rows;                  // rows a piece that starts inside a group re-renders from zero state without storing
    int64_t position;
    float* partial;                     // [nparts][frames][2]
    VoiceSeg seg[SIGB_VOICE_SEGS];
};

#ifdef __cplusplus
extern "C" {
#endif
// launch wrappers implemented in sigb_kernels.cu; return cudaError_t as int
int sigb_launch_chain_seq(const ChainDev* a, void* stream);
int sigb_launch_chain_scan(const ChainDev* a, int variant, void* stream, int* rows_done);
int sigb_cascade_pipe_ok(const ChainDev* a);
int sigb_cascade_pipe_items(const ChainDev* a, int max_segments);
int sigb_launch_cascade_pipe(const ChainDev* a, int max_segments, int sections_per_warp, void* stream);
int sigb_cascade_reg_ok(const ChainDev* a);
int sigb_launch_cascade_reg(const ChainDev* a, int max_segments, int variant, void* stream);
int sigb_cascade_reg_fill(const ChainDev* a, int max_segments, int variant);
void sigb_set_reg_pieces(int n);
void sigb_set_delta_probe(int n);
int sigb_osc_reg_ok(const ChainDev* a, int allow_delta);
int sigb_launch_osc_reg(const ChainDev* a, int max_segments, int allow_delta, void* stream);
int sigb_osc_reg_fill(const ChainDev* a, int max_segments, int allow_delta);
int sigb_osc_fill_ok(const ChainDev* a);
int sigb_launch_osc_fill(const ChainDev* a, void* stream);
void sigb_set_osc_pieces_pct(int n);
int sigb_launch_ewise(const EwiseDev* a, void* stream);
int sigb_launch_reduce(const ReduceDev* a, void* stream);
int sigb_scan_rows_per_step(int nsec, int variant);
void sigb_set_scan_tma(int on);
void sigb_set_scan_split(int on);
void sigb_set_scan_rot(int on);
int sigb_launch_bank(const BankDev* a, void* stream);
void sigb_set_bank_unroll(int n);
int sigb_voices_ctas(int channels, int M);                         // CTAs (= partials) a segment of `channels` needs
int sigb_voices_block_rows(int M);                                 // rows per block of the (group, block) space
int sigb_voices_slots(int M);                                      // CTAs of k_voices resident on the device at once
int sigb_launch_voices(const VoicesDev* a, void* stream);
int sigb_launch_voices_finish(const float* partial, int nparts, int frames, float* out, int64_t ld_out, void* stream);
int sigb_launch_param_eval(const ParamInstr* prog_dev, int n_instr, int n_rows, double* drows, float* frows, int row_stride,
                           int64_t position, const int64_t* pos_ptr, int rate, void* stream);
int sigb_launch_design(const DesignDev* a, void* stream);
int sigb_launch_osc_tables(int C, const double* hertz, const double* phase, int rate, unsigned long long* theta0,
                           unsigned long long* dtheta, float* rot1, void* stream);
int sigb_launch_gain_rows(int C, const float* gain_const, const double* row, float* gain_out, void* stream);
int sigb_launch_pan_weights(int C, const double* gain, const double* pan, float* wl, float* wr, void* stream);
int sigb_launch_probe_sin(const double* r, int n, float* out, int variant, void* stream);
#ifdef __cplusplus
}
#endif
