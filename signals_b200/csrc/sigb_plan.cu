// libsigb200 host runtime: graph records -> fused launch plan -> render.
//
// The reference evaluates its node graph by Python recursion, one numpy temporary per edge per
// block (/root/reference/src/signals/chain/__init__.py:245-315).  Here the host compiles the
// topologically sorted node records once into a short list of kernel launches:
//   * every maximal linear run  Osc|block -> {Gain, LowPass, HighPass}*  becomes ONE chain launch
//     (all gains folded into one per-channel factor: the filters are linear and time-invariant),
//   * Mix / RingMod / Amp / Merge / GroupSum / PanSum run on materialised float32 blocks,
//   * filter state lives in the plan and is carried from one render call to the next.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "sigb200.h"
#include "sigb_design.h"
#include "sigb_internal.h"

namespace {

thread_local std::string g_last_error;

// defaults applied to plans created afterwards (sigb_set_default_option)
int64_t g_default_fuse_reduce = 1;
int64_t g_default_voices_m = 0;
int64_t g_default_fuse_pointwise = 1;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(SIGB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
    } while (0)

enum ValKind { VK_NONE = 0, VK_CONST, VK_BUF, VK_EXT };

struct Val {
    ValKind kind = VK_NONE;
    int channels = 1;
    std::vector<double> cv;   // VK_CONST
    int buf = -1;             // VK_BUF: index into Plan::bufs (-1 = the caller's `out`)
    int64_t const_off = -1;   // VK_CONST: float offset in the arena once staged
};

struct BufInfo {
    int channels = 0;
    float* ptr = nullptr;
    int64_t cap_rows = 0;
};

struct ExtBinding {
    const float* ptr = nullptr;
    int64_t first_row = 0;      // absolute frame position of row 0 of the bound memory
    int64_t rows = 0;
};

// A table staged in the arena: byte offset + size; resolved to a device pointer after upload.
struct Table {
    int64_t off = -1;
    template <class T> const T* dev(const unsigned char* base) const {
        return off < 0 ? nullptr : reinterpret_cast<const T*>(base + off);
    }
};

struct ChainSpec {
    int C = 0;
    int src_kind = SRC_CONST;
    int wave = 0;
    int nsec = 0;                // padded section count the kernels run
    int nsec_real = 0;
    uint8_t sec_kind[SIGB_MAX_SEC] = {0};
    Table hertz, phase, theta0, dtheta, rot1, constv, coef, gain, apow, apow_h, ztab, m8, hrec;
    int src_node = -1;           // SRC_BUF: node whose value is read
    int src_osc_node = -1;       // SRC_OSC: the oscillator node
    int64_t state_off = 0;       // doubles into the state arena
    int state_cur = 0;           // which copy of the state arena holds the live state
    int warm_rows = -1;          // rows until the cascade forgets its initial state (see sigb_section_decay_rows)
    double warm_static = 0.0;    // ... the share of the filters with constant cutoffs (modulated ones add theirs per request)
    int dst_node = -1;
    int dst_coff = 0;             // first column of the destination block this chain writes (a Merge rendered in place)
    int gain_row = -1;            // a MODULATED Gain at the end of the chain (tremolo): row of the parameter program; k_gain_rows
    Table gain_dev;               // ... writes gain_const[c] * row[c] here at every request's first frame, and the kernels read it
    int hertz_row = -1, phase_row = -1;   // SRC_OSC with modulated hertz / phase: rows of the parameter program
    bool osc_tables_dev = false;          // ... and (Sine) the Q0.64 phase tables are re-derived from those rows on the device per request
    // filters whose cutoff is driven by an emitter: their sections are re-designed on the device once per request
    // (k_design) from row `row` of the parameter program
    struct ModFilter { int s0, order, row, highpass, slot; };   // slot: index of this filter's decay horizon in d_warm
    std::vector<ModFilter> mods;
    // Mix / RingMod fused as an epilogue of a stateless chain (k_chain_seq)
    int epi_op = 0, epi_side = 0, epi_node = -1, epi_wave = -1, epi_p_row = -1;
    Table epi_p, epi_hertz, epi_phase, epi_gain, epi_theta0, epi_dtheta;
    std::vector<double> gain_d;  // folded gain in float64 (fused reductions derive their weights from it)
    bool has_gain = false;
    double max_abs_hertz = 0.0, max_abs_phase = 0.0;   // SRC_OSC: sizes the phase-word guard band
};

struct EwiseSpec {
    int op = EW_COPY;
    int C = 0;
    int a_node = -1, b_node = -1;   // -1 = zeros(1,1)
    Table p;
    int p_row = -1;                 // modulated parameter: row of the block-rate parameter program instead of `p`
    int dst_node = -1;
    int dst_coff = 0;
};

struct ReduceSpec {
    int C = 0, groups = 0, pan = 0;
    int in_node = -1;
    Table w;
    int w_row = -1;              // PanSum with a modulated pan: row of the parameter program holding pan[C] for this request
    int dst_node = -1;
};

// oscillator bank fused with its GroupSum (k_bank)
struct BankSpec {
    ChainSpec ch;
    Table rot32;                 // (cos, sin) of each partial's phase advance over 32 rows (k_bank's rotation)
    int groups = 0;
    int dst_node = -1;
};

// voice chains fused with their PanSum (k_voices): one segment per leaf of the Merge tree
struct VoiceSegSpec {
    ChainSpec ch;
    Table wl, wr;
    Table gain64;              // modulated pan only: the folded gain in float64 (k_pan_weights re-derives wl / wr per request)
    int coff = 0;              // first channel of this segment in the PanSum's input (= column of the pan row)
};
struct VoicesSpec {
    std::vector<VoiceSegSpec> segs;
    int M = 1;
    int nparts = 0;
    int partial_buf = -1;      // index into Plan::bufs (2 * nparts "channels")
    int state_cur = 0;         // which copy of the state arena holds the live filter state
    int dst_node = -1;
    int pan_row = -1;          // pan driven by an emitter: row of the parameter program holding pan[C] for the request
};

enum LaunchKind { LK_CHAIN, LK_EWISE, LK_REDUCE, LK_BANK, LK_VOICES };
struct Launch {
    LaunchKind kind;
    int idx;
};

}  // namespace

struct sigb_plan {
    int channels = 0, rate = 0, root = -1;
    std::vector<sigb_node> nodes;
    std::vector<double> data;
    std::vector<Val> vals;
    std::vector<int> uses;
    std::vector<int> node_ctx;      // warm-up frames a seek needs at this node (fx.py:82-83, summed down a cascade)
    std::vector<BufInfo> bufs;
    std::vector<ExtBinding> ext;        // per node
    std::vector<ChainSpec> chains;
    std::vector<EwiseSpec> ewises;
    std::vector<ReduceSpec> reduces;
    std::vector<BankSpec> banks;
    std::vector<VoicesSpec> voices;
    std::vector<Launch> launches;
    // block-rate parameter program (modulated parameters), evaluated once per request at its position
    std::vector<ParamInstr> pprog;
    std::vector<std::vector<double>> prow_const;   // per row: constant values (empty: computed row)
    std::vector<int> prow_width;                   // natural channel count of the row's node
    std::vector<int> pnode_row;                    // node -> row (-1: not in the program)
    int pwidth = 1;                                // row stride = widest consumer
    ParamInstr* d_pprog = nullptr;
    double* d_prow_d = nullptr;
    float* d_prow_f = nullptr;
    std::vector<unsigned char> arena;   // host image of all parameter tables
    unsigned char* d_arena = nullptr;
    int64_t n_state = 0;
    double* d_state = nullptr;          // two copies of the state arena: [0, n_state) and [n_state, 2 n_state)
    int context = 0;
    int zero_const_node = -1;
    Val zero_val;
    // options
    int64_t opt_scan_variant = 18;
    int64_t opt_force_seq = 0;
    int64_t opt_scan_max_tiles = 148 * 6;
    int64_t opt_slab_frames = 0;
    int64_t opt_host_slab_bytes = 64ll << 20;
    int64_t opt_buffer_budget = 6ll << 30;
    int64_t opt_cascade_pipe = -1;      // -1: automatic choice; 0: never the section-pipelined kernel; n > 0: always from n sections
    int64_t opt_cascade_reg = -1;       // -1: register-resident cascade kernel whenever it can take the chain; 0: never
    int64_t opt_osc_reg = 2;            // oscillator-fed chains of >= n sections run register-resident (k_osc_delta / k_osc_reg); 0: never
    int64_t opt_osc_fill = 1;           // 0: stateless oscillator chains stay on k_chain_seq whatever their width (A/B)
    int64_t opt_osc_delta = 1;          // 0: oscillator-fed register chains keep the state-variable sections (A/B)
    bool osc_reg_user = false;          // set through the option: honoured as given; the default's 2-section rule is for unmodulated chains that fill the machine
    int64_t opt_reg_variant = 0;        // register cascades: 0 = delta form where it applies, else 8-row blocks; 1 = 4-row blocks; 4 = state-variable form in 8-row blocks
    int64_t opt_pipe_spw = 1;           // sections per warp in k_cascade_pipe (2: halves the shared-memory traffic)
    int64_t opt_pipe_segments = 64;     // upper bound on the time segments per tile of k_cascade_pipe (1: never split)
    int64_t opt_fuse_reduce = 1;        // 0: GroupSum / PanSum always run on materialised blocks
    int64_t opt_fuse_pointwise = 1;     // 0: Mix / RingMod always run on materialised blocks
    int64_t opt_voices_segments = 0;    // k_voices: 0 auto (equal pieces per CTA slot), 1 one piece per voice group, n > 1: n pieces per group
    int64_t opt_voices_pieces = 16;     // k_voices, automatic mode: pieces per resident CTA slot (measured 1 / 2 / 4 / 8 / 16 on C5:
                                        // 1.40 / 1.50 / 1.58 / 1.65 / 1.67e12 voice-samples/s at 1M instances, 1.35 / 1.41 / 1.46 / 1.48 / 1.48e12 at 131,072)
    int64_t opt_voices_m = 0;           // 0: auto; 1 or 4: channels per thread in k_voices
    // runtime
    bool uploaded = false;
    bool have_pos = false;
    int64_t next_pos = 0;
    int64_t launch_count = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
    float* scratch = nullptr;
    int64_t scratch_floats = 0;
    // host-render resources
    cudaStream_t s_render = nullptr, s_copy = nullptr;
    float* stage[2] = {nullptr, nullptr};
    int64_t stage_floats = 0;
    cudaEvent_t ev_rendered[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    cudaEvent_t ev_caller = nullptr;
    // modulated cutoffs: per-filter decay horizons written by k_design, and the "cutoff outside (0, Nyquist)" flag
    int n_mods = 0;
    bool warm_on_device = false;        // this request's k_design launches wrote the horizons to d_warm
    int* d_warm = nullptr;
    int* h_warm = nullptr;              // page-locked copy
    int* h_err = nullptr;               // page-locked, device-mapped: k_design stores 1 here
    // realtime block path (sigb_render_block): one captured CUDA graph per block length, position read from a
    // page-locked block header, output written straight into page-locked staging
    struct RtGraph { int frames; cudaGraph_t graph; cudaGraphExec_t exec; int nodes; uint64_t state_sig; };
    std::vector<RtGraph> rt_graphs;
    int64_t* rt_hdr = nullptr;          // page-locked, device-mapped: [0] = position of the block
    float* rt_stage = nullptr;          // page-locked, device-mapped (frames, channels) block
    int64_t rt_stage_floats = 0;
    const int64_t* rt_pos_ptr = nullptr;    // set while run_slab records / runs a realtime block
    int rt_state = 0;                   // 0 unknown, 1 eligible, -1 not (falls back to sigb_render_host)
    int64_t opt_rt_graph = 1;           // 0: realtime blocks launch their kernels directly (A/B)
    int64_t opt_rt_max_bytes = 1 << 20; // larger blocks go through sigb_render_host (DMA copies)
    int64_t opt_restart = 0;            // 1: every request restarts the filters from zero state + context (the reference's
                                        // own blockwise behaviour, fx.py:82-83, 93-105), for A/B against it
    int64_t graph_launches = 0;
    // taps (Wave / Spec / FileWriter inputs kept materialised for sigb_plan_read_tap)
    std::vector<int> tap_nodes;
    int64_t last_position = 0, last_frames = 0, last_slab = 0;
};

namespace {

int64_t arena_put(sigb_plan* p, const void* src, size_t bytes) {
    size_t off = (p->arena.size() + 255) & ~size_t(255);
    p->arena.resize(off + bytes);
    std::memcpy(p->arena.data() + off, src, bytes);
    return (int64_t)off;
}

template <class T>
Table put_vec(sigb_plan* p, const std::vector<T>& v) {
    Table t;
    t.off = arena_put(p, v.data(), v.size() * sizeof(T));
    return t;
}

bool is_chain_kind(int kind) { return kind == SIGB_NODE_OSC || kind == SIGB_NODE_GAIN || kind == SIGB_NODE_FILTER; }

// value of a block-rate / constant port: nullptr when the node is not a compile-time constant
const std::vector<double>* const_of(sigb_plan* p, int idx) {
    if (idx < 0) return &p->zero_val.cv;
    const Val& v = p->vals[idx];
    return v.kind == VK_CONST ? &v.cv : nullptr;
}

// replicate a (1,1)/(1,C) constant row to C entries; false when not broadcast-compatible
bool rep(const std::vector<double>& v, int C, std::vector<double>* out) {
    if ((int)v.size() == C) { *out = v; return true; }
    if (v.size() == 1) { out->assign(C, v[0]); return true; }
    return false;
}

int section_count(int order) { return order / 2 + (order & 1); }

// Per-channel table building for large banks (C5: a million instances) on all host cores: fn(begin, end, worker) over
// disjoint channel ranges.  Small banks run inline.
template <class Fn>
void parallel_channels(int n, Fn fn, int* n_workers = nullptr) {
    const int hw = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    const int workers = n < 32768 ? 1 : std::min(hw, n / 8192);
    if (n_workers) *n_workers = workers;
    if (workers <= 1) {
        fn(0, n, 0);
        return;
    }
    std::vector<std::thread> th;
    for (int w = 0; w < workers; ++w)
        th.emplace_back([=]() { fn((int)((int64_t)n * w / workers), (int)((int64_t)n * (w + 1) / workers), w); });
    for (std::thread& t : th) t.join();
}
constexpr int kMaxWorkers = 32;

// guard band (units of 2^-32 cycles) around waveform discontinuities for the phase-word fast paths:
// in-tile drift of the rounded increment, the rounding of the top word, and the float64 rounding of
// the reference's own phase (3 roundings of relative size 2^-53 on `cycles` cycles)
int phase_guard(double max_abs_hertz, double max_abs_phase, int64_t last_row, int rate) {
    const double cyc = max_abs_hertz * (double)last_row / rate + max_abs_phase + 1.0;
    const double g = 16.0 + cyc * 3.0 * 4294967296.0 / 9007199254740992.0;
    return g < 1073741823.0 ? (int)std::ceil(g) : 0x3fffffff;
}

int pad_sections(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return n == 0 ? 0 : p;
}

unsigned long long frac_q64(double x) {
    // frac(x) in Q0.64 for finite x (float64 has at most 53 significant bits, so this is exact
    // whenever x - floor(x) is; a tiny negative x rounds to 1.0 == 0 cycles)
    if (!std::isfinite(x)) return 0ull;
    const double f = x - std::floor(x);
    if (!(f < 1.0)) return 0ull;
    const double scaled = std::ldexp(f, 32);
    const double hi = std::floor(scaled);
    const double lo = scaled - hi;              // exact, in [0,1)
    return ((unsigned long long)hi << 32) | (unsigned long long)std::floor(std::ldexp(lo, 32));
}

// floor(frac(hertz / rate) * 2^64), exact integer arithmetic on the float64 mantissa.
// Negative frequencies map to the two's-complement (phase runs backwards).
unsigned long long ratio_q64(double hertz, int rate) {
    if (!std::isfinite(hertz) || rate <= 0 || hertz == 0.0) return 0ull;
    int e;
    const double m = std::frexp(std::fabs(hertz), &e);                        // |hertz| = m 2^e
    const unsigned long long mant = (unsigned long long)std::ldexp(m, 53);    // 53-bit integer
    e -= 53;                                                                  // |hertz| = mant 2^e
    const unsigned __int128 R = (unsigned __int128)(unsigned)rate;
    unsigned long long q;
    if (e >= 0) {
        unsigned __int128 r = (unsigned __int128)mant % R;                    // whole cycles drop out
        for (int i = 0; i < e; ++i) r = (r << 1) % R;
        q = (unsigned long long)((r << 64) / R);
    } else {
        const int k = -e;
        unsigned __int128 num;
        if (k <= 64) num = (unsigned __int128)mant << (64 - k);
        else num = (k - 64 >= 64) ? 0 : ((unsigned __int128)mant >> (k - 64));
        q = (unsigned long long)(num / R);                                    // low 64 bits = mod 2^64
    }
    return hertz < 0.0 ? (0ull - q) : q;
}

struct Builder {
    sigb_plan* p;
    int err = SIGB_OK;

    int node_C(int i) const { return i == p->root ? p->channels : p->nodes[i].channels; }

    int ensure(int i);            // materialise node i's value (emits launches); returns status
    int param_row(int idx, int* row);          // value of node idx in the block-rate parameter program
    int param_port(int idx, int C, int* row);  // ... checked against a consumer of C channels
    bool gain_is_const(int i) const { return p->nodes[i].kind != SIGB_NODE_GAIN || const_of(p, p->nodes[i].in[1]) != nullptr; }
    int build_chain(int i);
    int make_chain(int i, ChainSpec& ch, bool scan_tables, int width = 0);
    bool stateless_osc(int i) const;           // a pure oscillator (+ constant gains) chain nobody else consumes
    int build_fused_pointwise(int i, bool* done);
    bool pure_osc_run(int i, int* nsec, int* wave) const;
    bool collect_voice_leaves(int idx, std::vector<int>* leaves) const;
    int build_bank(int i);
    int build_voices(int i, const std::vector<int>& leaves);
    int build_ewise(int i);
    int build_merge(int i);
    int build_reduce(int i);
    void mark_consumed(int idx) {       // the run ending at idx lives inside a fused launch: never materialised
        for (int cur = idx; cur >= 0;) {
            p->vals[cur].kind = VK_BUF;
            p->vals[cur].channels = p->nodes[cur].channels;
            p->vals[cur].buf = -2;
            if (p->nodes[cur].kind == SIGB_NODE_OSC) break;
            cur = p->nodes[cur].in[0];
        }
    }
    int new_buf(int i) {
        if (i == p->root) return -1;
        BufInfo b;
        b.channels = p->nodes[i].channels;
        p->bufs.push_back(b);
        return (int)p->bufs.size() - 1;
    }
};

int Builder::param_row(int idx, int* row) {
    if (p->pnode_row.empty()) p->pnode_row.assign(p->nodes.size() + 1, -1);
    const size_t slot = idx < 0 ? p->nodes.size() : (size_t)idx;       // the shared zeros(1,1) of unconnected ports
    if (p->pnode_row[slot] >= 0) {
        *row = p->pnode_row[slot];
        return SIGB_OK;
    }
    if ((int)p->prow_const.size() >= SIGB_PARAM_ROWS)
        return fail(SIGB_EUNSUPPORTED, "block-rate parameter graph larger than " + std::to_string(SIGB_PARAM_ROWS) + " nodes");
    const std::vector<double>* cv = const_of(p, idx);
    auto new_row = [&](int width) {
        p->prow_const.emplace_back();
        p->prow_width.push_back(width);
        return (int)p->prow_const.size() - 1;
    };
    if (cv) {
        const int r = new_row((int)cv->size());
        p->prow_const[r] = *cv;
        p->pnode_row[slot] = *row = r;
        return SIGB_OK;
    }
    const sigb_node& n = p->nodes[idx];
    ParamInstr in;
    std::memset(&in, 0, sizeof(in));
    int nops = 0;
    switch (n.kind) {
        case SIGB_NODE_OSC: in.op = PRM_OSC; in.wave = n.subtype; nops = 2; break;     // hertz, phase
        case SIGB_NODE_GAIN: case SIGB_NODE_RINGMOD: in.op = PRM_MUL; nops = 2; break;
        case SIGB_NODE_MIX: in.op = PRM_MIX; nops = 3; break;                           // left, right, mix
        case SIGB_NODE_AMP: in.op = PRM_AMP; nops = 2; break;
        default:
            return fail(SIGB_EUNSUPPORTED, "node " + std::to_string(idx) + ": this node kind cannot drive a block-rate parameter "
                        "(supported: Fixed, oscillators, Gain, Mix, RingMod, Amp)");
    }
    int rows[3] = {0, 0, 0}, width = 1;
    for (int k = 0; k < nops; ++k) {
        int st = param_row(n.in[k], &rows[k]);
        if (st != SIGB_OK) return st;
        const int w = p->prow_width[rows[k]];
        if (w != 1 && width != 1 && w != width)
            return fail(SIGB_ESHAPE, "node " + std::to_string(idx) + ": parameter operands of " + std::to_string(w) + " and " + std::to_string(width) + " channels");
        width = std::max(width, w);
    }
    in.a = rows[0]; in.b = rows[1]; in.c = rows[2];
    in.dst = new_row(width);
    in.width = width;
    p->pprog.push_back(in);
    p->pnode_row[slot] = *row = in.dst;
    return SIGB_OK;
}

int Builder::param_port(int idx, int C, int* row) {
    int st = param_row(idx, row);
    if (st != SIGB_OK) return st;
    const int w = p->prow_width[*row];
    if (w != 1 && w != C)
        return fail(SIGB_ESHAPE, "node " + std::to_string(idx) + ": block-rate parameter with " + std::to_string(w) + " channels incompatible with " + std::to_string(C));
    p->pwidth = std::max(p->pwidth, std::max(C, w));
    return SIGB_OK;
}

int Builder::ensure(int i) {
    if (i < 0) return SIGB_OK;
    if (p->vals[i].kind != VK_NONE) return SIGB_OK;
    const sigb_node& n = p->nodes[i];
    switch (n.kind) {
        case SIGB_NODE_GAIN:
            if (!gain_is_const(i)) {
                // a modulated Gain (tremolo) at the END of a chain nobody else reads multiplies the chain's output, per request:
                // it rides on the chain's gain table, re-derived on the device at every request's first frame (k_gain_rows).
                // (A modulated Gain in FRONT of a filter does not commute with it across requests and stays a k_ewise launch.)
                const int u = n.in[0];
                if (u >= 0 && is_chain_kind(p->nodes[u].kind) && p->uses[u] == 1 && p->vals[u].kind == VK_NONE && gain_is_const(u) &&
                    p->nodes[u].channels == node_C(i) && p->opt_fuse_pointwise != 0) {
                    ChainSpec ch;
                    int st = make_chain(u, ch, true);
                    if (st != SIGB_OK) return st;
                    st = param_port(n.in[1], ch.C, &ch.gain_row);
                    if (st != SIGB_OK) return st;
                    ch.gain_dev = put_vec(p, std::vector<float>(ch.C, 0.0f));
                    ch.dst_node = i;
                    p->vals[u].kind = VK_BUF;
                    p->vals[u].channels = ch.C;
                    p->vals[u].buf = -2;
                    Val v;
                    v.kind = VK_BUF;
                    v.channels = ch.C;
                    v.buf = new_buf(i);
                    if (v.buf >= 0) p->bufs[v.buf].channels = ch.C;
                    p->vals[i] = v;
                    p->chains.push_back(ch);
                    p->launches.push_back({LK_CHAIN, (int)p->chains.size() - 1});
                    return SIGB_OK;
                }
                return build_ewise(i);                         // modulated gain elsewhere: a pointwise launch
            }
            return build_chain(i);
        case SIGB_NODE_OSC:
        case SIGB_NODE_FILTER: return build_chain(i);
        case SIGB_NODE_MIX:
        case SIGB_NODE_RINGMOD:
        case SIGB_NODE_AMP: return build_ewise(i);
        case SIGB_NODE_MERGE: return build_merge(i);
        case SIGB_NODE_GROUPSUM:
        case SIGB_NODE_PANSUM: return build_reduce(i);
        case SIGB_NODE_TAP: {
            // a pass-through side-effect node: its value IS its input's (chain/__init__.py:409-417), kept materialised
            // in a plan buffer so that sigb_plan_read_tap can hand the block to the host after the render
            if (n.in[0] < 0) {
                p->vals[i] = p->zero_val;
            } else {
                int st = ensure(n.in[0]);
                if (st != SIGB_OK) return st;
                p->vals[i] = p->vals[n.in[0]];
            }
            return SIGB_OK;
        }
        default: return fail(SIGB_EINVAL, "node " + std::to_string(i) + ": unexpected kind");
    }
}

int Builder::build_chain(int i) {
    ChainSpec ch;
    int st = make_chain(i, ch, true);
    if (st != SIGB_OK) return st;
    const int C = ch.C;
    Val v;
    v.kind = VK_BUF;
    v.channels = C;
    v.buf = new_buf(i);
    if (v.buf >= 0) p->bufs[v.buf].channels = C;
    p->vals[i] = v;
    p->chains.push_back(ch);
    p->launches.push_back({LK_CHAIN, (int)p->chains.size() - 1});
    return SIGB_OK;
}

// Walks the linear run ending at node i down to its source and fills `ch` with the per-channel tables
// (no launch is recorded).  scan_tables = false skips the tables only the time-parallel kernels read.
int Builder::make_chain(int i, ChainSpec& ch, bool scan_tables, int width) {
    const int C = width > 0 ? width : node_C(i);
    ch.C = C;
    ch.dst_node = i;
    std::vector<double> gain(C, 1.0);
    bool has_gain = false;
    std::vector<int> filters;     // sink-to-source order
    int nsec = 0;
    int cur = i;
    std::vector<double> tmp;
    for (;;) {
        const sigb_node& n = p->nodes[cur];
        if (n.kind == SIGB_NODE_OSC) {
            const std::vector<double>* hz = const_of(p, n.in[0]);
            const std::vector<double>* ph = const_of(p, n.in[1]);
            if (!hz || !ph) {
                // modulated frequency / phase: sampled once per request in float64 by the parameter program
                // (osc.py:28-30); the chain kernels then take their float64 waveform path
                int st = param_port(n.in[0], C, &ch.hertz_row);
                if (st == SIGB_OK) st = param_port(n.in[1], C, &ch.phase_row);
                if (st != SIGB_OK) return st;
                ch.src_kind = SRC_OSC;
                ch.wave = n.subtype;
                if (scan_tables) {
                    // vibrato: hertz / phase are constant within a request, so the exact Q0.64 phase (theta0 + n dtheta) holds for
                    // the request; k_osc_tables re-derives theta0 / dtheta (and a Sine's rot1) from the sampled rows at every
                    // request's first frame and the phase-word fast paths (k_osc_fill, k_chain_scan3's sine, k_osc_delta) run as
                    // for a constant oscillator
                    ch.theta0 = put_vec(p, std::vector<unsigned long long>(C, 0ull));
                    ch.dtheta = put_vec(p, std::vector<unsigned long long>(C, 0ull));
                    if (n.subtype == SIGB_WAVE_SINE) ch.rot1 = put_vec(p, std::vector<float>((size_t)C * 2, 0.0f));
                    ch.osc_tables_dev = true;
                }
                break;
            }
            std::vector<double> hzv, phv;
            if (!rep(*hz, C, &hzv) || !rep(*ph, C, &phv))
                return fail(SIGB_ESHAPE, "node " + std::to_string(cur) + ": hertz/phase channels incompatible with " + std::to_string(C));
            ch.src_kind = SRC_OSC;
            ch.src_osc_node = cur;
            ch.wave = n.subtype;
            ch.hertz = put_vec(p, hzv);
            ch.phase = put_vec(p, phv);
            for (int c = 0; c < C; ++c) {
                ch.max_abs_hertz = std::max(ch.max_abs_hertz, std::fabs(hzv[c]));
                ch.max_abs_phase = std::max(ch.max_abs_phase, std::fabs(phv[c]));
            }
            {   // Q0.64 phase / increment: sine fast path of the chain kernels, every wave in k_voices
                std::vector<unsigned long long> t0(C), dt(C);
                const int rate = p->rate;
                parallel_channels(C, [&](int c0, int c1, int) {
                    for (int c = c0; c < c1; ++c) {
                        t0[c] = frac_q64(phv[c]);
                        dt[c] = ratio_q64(hzv[c], rate);
                    }
                });
                ch.theta0 = put_vec(p, t0);
                ch.dtheta = put_vec(p, dt);
                if (scan_tables && n.subtype == SIGB_WAVE_SINE) {
                    // (cos, sin) of the one-row phase advance, from the exact Q0.64 increment in float64
                    std::vector<float> rot((size_t)C * 2);
                    for (int c = 0; c < C; ++c) {
                        const double ang = 6.283185307179586476925 * std::ldexp((double)(long long)dt[c], -64);
                        rot[2 * c + 0] = (float)std::cos(ang);
                        rot[2 * c + 1] = (float)std::sin(ang);
                    }
                    ch.rot1 = put_vec(p, rot);
                }
            }
            break;
        }
        int up;
        if (n.kind == SIGB_NODE_GAIN) {
            const std::vector<double>* g = const_of(p, n.in[1]);
            if (!g) return fail(SIGB_EUNSUPPORTED, "node " + std::to_string(cur) + ": Gain.right driven by a non-constant emitter");
            if (!rep(*g, C, &tmp)) return fail(SIGB_ESHAPE, "node " + std::to_string(cur) + ": gain channels incompatible");
            for (int c = 0; c < C; ++c) gain[c] *= tmp[c];
            has_gain = true;
            up = n.in[0];
        } else {   // FILTER
            if (n.subtype != SIGB_FILT_LOWPASS && n.subtype != SIGB_FILT_HIGHPASS)
                return fail(SIGB_EUNSUPPORTED, "node " + std::to_string(cur) + ": band filters are unreachable in the reference (fx.py:99)");
            if (n.order < 1) return fail(SIGB_EINVAL, "node " + std::to_string(cur) + ": filter order < 1");
            filters.push_back(cur);
            nsec += section_count(n.order);
            up = n.in[0];
        }
        if (up < 0) {
            if (!filters.empty() && C > 1)
                return fail(SIGB_EINDEX, "node " + std::to_string(cur) + ": filter input narrower than the request (fx.py:105)");
            ch.src_kind = SRC_CONST;
            ch.constv = put_vec(p, std::vector<float>(C, 0.0f));
            break;
        }
        const sigb_node& u = p->nodes[up];
        bool fusable = is_chain_kind(u.kind) && p->uses[up] == 1 && p->vals[up].kind == VK_NONE && gain_is_const(up);
        if (fusable && u.kind == SIGB_NODE_FILTER && nsec + section_count(u.order) > SIGB_MAX_SEC) fusable = false;
        // a filter needs its input at full width (fx.py:105 indexes input_[:, i])
        if (n.kind == SIGB_NODE_FILTER && C > 1 && u.channels != C)
            return fail(SIGB_EINDEX, "node " + std::to_string(cur) + ": filter input narrower than the request (fx.py:105)");
        if (fusable) {
            cur = up;
            continue;
        }
        int st = ensure(up);
        if (st != SIGB_OK) return st;
        const Val& uv = p->vals[up];
        if (uv.channels != 1 && uv.channels != C)
            return fail(SIGB_ESHAPE, "node " + std::to_string(up) + ": block with " + std::to_string(uv.channels) + " channels incompatible with requested " + std::to_string(C));
        if (uv.kind == VK_CONST) {
            std::vector<double> cvd;
            rep(uv.cv, C, &cvd);
            std::vector<float> cvf(cvd.begin(), cvd.end());
            ch.src_kind = SRC_CONST;
            ch.constv = put_vec(p, cvf);
        } else {
            ch.src_kind = SRC_BUF;
            ch.src_node = up;
        }
        break;
    }
    // sections, source-to-sink
    std::reverse(filters.begin(), filters.end());
    ch.nsec_real = nsec;
    ch.nsec = pad_sections(nsec);
    if (ch.nsec > SIGB_MAX_SEC) return fail(SIGB_EUNSUPPORTED, "filter cascade longer than 16 sections in one node");
    if (ch.nsec > 0) {
        const size_t CS = scan_tables ? (size_t)C : 0;      // scan-only tables are left empty when unused
        std::vector<float> coef((size_t)ch.nsec * 3 * C, 0.0f);
        std::vector<double> apow((size_t)ch.nsec * 4 * CS, 0.0);
        std::vector<double> apow_h((size_t)ch.nsec * 4 * CS, 0.0);
        std::vector<float> ztab((size_t)ch.nsec * SIGB_SCAN_L * 2 * CS, 0.0f);
        std::vector<float> m8((size_t)ch.nsec * 4 * CS, 0.0f);
        std::vector<float> hrec((size_t)ch.nsec * 4 * CS, 0.0f);
        for (int s = 0; s < ch.nsec; ++s) {   // identity padding: high-pass with g = 0 passes x through
            ch.sec_kind[s] = SEC_HP;
            for (int c = 0; c < C; ++c) {
                coef[((size_t)s * 3 + 2) * C + c] = 1.0f;
                if (!scan_tables) continue;
                apow[((size_t)s * 4 + 0) * C + c] = 1.0;
                apow[((size_t)s * 4 + 3) * C + c] = 1.0;
                apow_h[((size_t)s * 4 + 0) * C + c] = 1.0;
                apow_h[((size_t)s * 4 + 3) * C + c] = 1.0;
                m8[((size_t)s * 4 + 0) * C + c] = 1.0f;
                m8[((size_t)s * 4 + 3) * C + c] = 1.0f;
                hrec[((size_t)s * 4 + 0) * C + c] = 1.0f;      // identity: det = 1, tr - 1 - det = 0, differences 0
            }
        }
        int s0 = 0;
        double warm = 0.0;
        std::vector<double> rho_max(scan_tables ? C : 0, 0.0), tau_sum(scan_tables ? C : 0, 0.0);
        for (int f : filters) {
            const sigb_node& n = p->nodes[f];
            const std::vector<double>* cut = const_of(p, n.in[1]);
            if (!cut) {
                // modulated cutoff: sampled once per request by the parameter program, sections designed on the device
                ChainSpec::ModFilter mf;
                mf.s0 = s0;
                mf.order = n.order;
                mf.highpass = n.subtype == SIGB_FILT_HIGHPASS;
                mf.slot = p->n_mods++;
                int st = param_port(n.in[1], C, &mf.row);
                if (st != SIGB_OK) return st;
                if (p->prow_width[mf.row] != C)   // crit_1[0, i] is not broadcast (fx.py:99)
                    return fail(SIGB_EINDEX, "node " + std::to_string(f) + ": cutoff has " + std::to_string(p->prow_width[mf.row]) + " channels, request has " + std::to_string(C));
                std::vector<SvfSection> secs = sigb_butter_sections(n.subtype == SIGB_FILT_HIGHPASS, n.order, 0.5);
                const int ns = (int)secs.size();
                for (int k = 0; k < ns; ++k) {
                    float cf[3];
                    sigb_section_coef(secs[k], cf);          // placeholder until the first request designs them
                    ch.sec_kind[s0 + k] = (uint8_t)secs[k].kind;
                    for (int c = 0; c < C; ++c)
                        for (int j = 0; j < 3; ++j) coef[((size_t)(s0 + k) * 3 + j) * C + c] = cf[j];
                }
                ch.mods.push_back(mf);
                s0 += ns;                                    // (its decay horizon is designed per request: k_design)
                continue;
            }
            if ((int)cut->size() != C)   // crit_1[0, i] is not broadcast (fx.py:99)
                return fail(SIGB_EINDEX, "node " + std::to_string(f) + ": cutoff has " + std::to_string(cut->size()) + " channels, request has " + std::to_string(C));
            const int ns = section_count(n.order);
            std::vector<double> sec_warm(ns, 0.0);
            std::vector<double> warm_w((size_t)kMaxWorkers * ns, 0.0);       // per-worker maxima (merged below)
            std::vector<uint8_t> kinds(ns, 0);
            int bad[kMaxWorkers] = {0};
            const int rate = p->rate;
            const bool hp_filter = n.subtype == SIGB_FILT_HIGHPASS;
            const int order = n.order;
            parallel_channels(C, [&](int cbeg, int cend, int worker) {
              float tab[SIGB_SCAN_L * 2];
              double* sec_warm = warm_w.data() + (size_t)worker * ns;        // shadows the merged vector
              for (int c = cbeg; c < cend; ++c) {
                double wn = (*cut)[c] / (rate / 2.0);
                wn = std::min(std::max(wn, 0.0), 1.0);   // fx.py:100-101
                if (!(wn > 0.0 && wn < 1.0)) {            // scipy.signal.butter: "0 < Wn < 1"
                    bad[worker] = 1;
                    return;
                }
                std::vector<SvfSection> secs = sigb_butter_sections(hp_filter, order, wn);
                for (int k = 0; k < ns; ++k) {
                    float cf[3];
                    double m[4];
                    sigb_section_coef(secs[k], cf);
                    const int s = s0 + k;
                    if (c == cbeg) kinds[k] = (uint8_t)secs[k].kind;         // uniform over the channels
                    for (int j = 0; j < 3; ++j) coef[((size_t)s * 3 + j) * C + c] = cf[j];
                    if (!scan_tables) {
                        double m1[4];
                        sigb_section_transition(secs[k], 1, m1);
                        sec_warm[k] = std::max(sec_warm[k], sigb_section_decay_rows(m1));
                        continue;
                    }
                    sigb_section_transition(secs[k], SIGB_SCAN_L, m);
                    sigb_section_zero_input(secs[k], SIGB_SCAN_L, tab);
                    for (int j = 0; j < 4; ++j) apow[((size_t)s * 4 + j) * C + c] = m[j];
                    double mh[4], m1[4];
                    sigb_section_transition(secs[k], SIGB_SCAN_L / 2, mh);
                    sigb_section_transition(secs[k], 1, m1);
                    for (int j = 0; j < 4; ++j) m8[((size_t)s * 4 + j) * C + c] = (float)mh[j];
                    for (int j = 0; j < 4; ++j) apow_h[((size_t)s * 4 + j) * C + c] = mh[j];
                    {   // zero-input output recurrence h[k] = tr h[k-1] - det h[k-2] in delta form (see k_chain_scan2)
                        const double tr = m1[0] + m1[3], det = m1[0] * m1[3] - m1[1] * m1[2];
                        double a1 = 1.0, a2 = 0.0, b1 = 0.0, b2 = 1.0;      // outputs at samples 0 and 1 per unit state, float64
                        const double ya0 = sigb_section_step(secs[k], 0.0, a1, a2), yb0 = sigb_section_step(secs[k], 0.0, b1, b2);
                        const double ya1 = sigb_section_step(secs[k], 0.0, a1, a2), yb1 = sigb_section_step(secs[k], 0.0, b1, b2);
                        hrec[((size_t)s * 4 + 0) * C + c] = (float)det;
                        hrec[((size_t)s * 4 + 1) * C + c] = (float)(tr - 1.0 - det);
                        hrec[((size_t)s * 4 + 2) * C + c] = (float)(ya1 - ya0);
                        hrec[((size_t)s * 4 + 3) * C + c] = (float)(yb1 - yb0);
                    }
                    sec_warm[k] = std::max(sec_warm[k], sigb_section_decay_rows(m1));
                    const double rho = sigb_section_radius(m1);
                    rho_max[c] = std::max(rho_max[c], rho);
                    tau_sum[c] += rho < 1.0 ? 1.0 / (1.0 - rho) : 1e12;
                    for (int r = 0; r < SIGB_SCAN_L; ++r)
                        for (int j = 0; j < 2; ++j)
                            ztab[(((size_t)s * SIGB_SCAN_L + r) * 2 + j) * C + c] = tab[r * 2 + j];
                }
              }
            });
            for (int w = 0; w < kMaxWorkers; ++w) {
                if (bad[w])
                    return fail(SIGB_ECRIT, "node " + std::to_string(f) + ": Digital filter critical frequencies must be 0 < Wn < 1");
                for (int k = 0; k < ns; ++k) sec_warm[k] = std::max(sec_warm[k], warm_w[(size_t)w * ns + k]);
            }
            for (int k = 0; k < ns; ++k) ch.sec_kind[s0 + k] = kinds[k];
            s0 += ns;
            for (double wv : sec_warm) warm += wv;   // sections in series: budget the decays one after another
        }
        ch.coef = put_vec(p, coef);
        ch.warm_static = warm;
        ch.warm_rows = (warm < 1e8 && ch.mods.empty()) ? (int)std::ceil(warm) : -1;
        if (scan_tables) {
            ch.apow = put_vec(p, apow);
            ch.apow_h = put_vec(p, apow_h);
            ch.ztab = put_vec(p, ztab);
            if (nsec >= 2 && ch.warm_rows >= 0 && ch.mods.empty()) {
                // Cascades: the per-section budgets add up far too conservatively.  Simulate, in float64, the
                // zero-input decay (to 2^-44) of the slowest channels -- largest pole radius, largest sum of
                // section time constants -- and keep a 10 % margin over the worst of them.
                std::vector<int> cand;
                for (int pass = 0; pass < 2; ++pass) {
                    const std::vector<double>& key = pass == 0 ? rho_max : tau_sum;
                    std::vector<int> idx(C);
                    for (int c = 0; c < C; ++c) idx[c] = c;
                    const int top = std::min(C, 4);
                    std::partial_sort(idx.begin(), idx.begin() + top, idx.end(), [&](int x, int y) { return key[x] > key[y]; });
                    cand.insert(cand.end(), idx.begin(), idx.begin() + top);
                }
                int worst = 0;
                for (int c : cand) {
                    std::vector<SvfSection> all;
                    for (int f : filters) {
                        const sigb_node& n = p->nodes[f];
                        double wn = (*const_of(p, n.in[1]))[c] / (p->rate / 2.0);
                        std::vector<SvfSection> secs = sigb_butter_sections(n.subtype == SIGB_FILT_HIGHPASS, n.order, wn);
                        all.insert(all.end(), secs.begin(), secs.end());
                    }
                    const int rows = sigb_cascade_decay_rows(all, 44, 1 << 20);
                    worst = rows < 0 ? (1 << 30) : std::max(worst, rows);
                }
                if (worst < (1 << 30)) ch.warm_rows = std::min(ch.warm_rows, (int)(worst * 1.1) + 64);
                ch.warm_static = ch.warm_rows;
            }
            ch.m8 = put_vec(p, m8);
            ch.hrec = put_vec(p, hrec);
        }
        ch.state_off = p->n_state;
        p->n_state += (int64_t)ch.nsec * 2 * C;
    }
    if (has_gain) {
        std::vector<float> gf(gain.begin(), gain.end());
        ch.gain = put_vec(p, gf);
    }
    ch.gain_d = gain;
    ch.has_gain = has_gain;
    return SIGB_OK;
}

// true when the linear run ending at node i consists only of Gain / LowPass / HighPass nodes, each
// consumed once, down to an oscillator (so the whole run can live inside a fused render+reduce kernel)
bool Builder::pure_osc_run(int i, int* nsec, int* wave) const {
    int cur = i;
    *nsec = 0;
    for (;;) {
        if (cur < 0 || p->vals[cur].kind != VK_NONE || p->uses[cur] != 1) return false;
        const sigb_node& n = p->nodes[cur];
        if (n.kind == SIGB_NODE_OSC) {
            *wave = n.subtype;
            return const_of(p, n.in[0]) && const_of(p, n.in[1]);     // modulated oscillators are not fused
        }
        if (n.kind == SIGB_NODE_FILTER) {
            if (n.order < 1 || !const_of(p, n.in[1])) return false;      // modulated cutoffs are not fused
            *nsec += section_count(n.order);
        } else if (n.kind != SIGB_NODE_GAIN || !gain_is_const(cur)) {
            return false;
        }
        cur = n.in[0];
    }
}

bool Builder::collect_voice_leaves(int idx, std::vector<int>* leaves) const {
    if (idx < 0 || p->vals[idx].kind != VK_NONE || p->uses[idx] != 1) return false;
    const sigb_node& n = p->nodes[idx];
    if (n.kind == SIGB_NODE_MERGE)
        return collect_voice_leaves(n.in[0], leaves) && collect_voice_leaves(n.in[1], leaves);
    int nsec = 0, wave = 0;
    if (!pure_osc_run(idx, &nsec, &wave) || nsec > 1) return false;
    leaves->push_back(idx);
    return true;
}

int Builder::build_bank(int i) {
    const sigb_node& n = p->nodes[i];
    BankSpec b;
    int st = make_chain(n.in[0], b.ch, false);
    if (st != SIGB_OK) return st;
    b.groups = n.order;
    b.dst_node = i;
    {   // 32-row phase advance per partial: frac(32 hertz / rate) exactly (Q0.64), then cos / sin in float64
        const std::vector<double>& hz = *const_of(p, p->nodes[b.ch.src_osc_node].in[0]);
        std::vector<float> rot((size_t)b.ch.C * 2);
        for (int c = 0; c < b.ch.C; ++c) {
            const unsigned long long d32 = ratio_q64(hz[hz.size() == 1 ? 0 : c], p->rate) << 5;
            const double ang = 6.283185307179586476925 * std::ldexp((double)(long long)d32, -64);
            rot[2 * c + 0] = (float)std::cos(ang);
            rot[2 * c + 1] = (float)std::sin(ang);
        }
        b.rot32 = put_vec(p, rot);
    }
    if (b.groups < 1 || b.ch.C % b.groups != 0)
        return fail(SIGB_ESHAPE, "node " + std::to_string(i) + ": " + std::to_string(b.ch.C) + " channels do not split into " + std::to_string(b.groups) + " groups");
    if (i == p->root && b.groups != p->channels)
        return fail(SIGB_ESHAPE, "reduction yields " + std::to_string(b.groups) + " channels, request has " + std::to_string(p->channels));
    Val v;
    v.kind = VK_BUF;
    v.channels = b.groups;
    v.buf = new_buf(i);
    if (v.buf >= 0) p->bufs[v.buf].channels = b.groups;
    p->vals[i] = v;
    p->vals[n.in[0]].kind = VK_BUF;      // consumed inside the fused kernel: never materialised
    p->vals[n.in[0]].channels = b.ch.C;
    p->vals[n.in[0]].buf = -2;
    p->banks.push_back(b);
    p->launches.push_back({LK_BANK, (int)p->banks.size() - 1});
    return SIGB_OK;
}

int Builder::build_voices(int i, const std::vector<int>& leaves) {
    const sigb_node& n = p->nodes[i];
    const int Cin = p->nodes[n.in[0]].channels;
    // a pan driven by an emitter is a row of the parameter program: the (L, R) weight tables are then re-derived on the
    // device at every request's first frame (k_pan_weights, run_params) instead of once here
    const std::vector<double>* pan = const_of(p, n.in[1]);
    std::vector<double> pv;
    VoicesSpec vs;
    if (pan) {
        if (!rep(*pan, Cin, &pv)) return fail(SIGB_ESHAPE, "node " + std::to_string(i) + ": pan channels incompatible");
    } else {
        int st = param_port(n.in[1], Cin, &vs.pan_row);
        if (st != SIGB_OK) return st;
        pv.assign(Cin, 0.5);
    }
    if (i == p->root && p->channels != 2)
        return fail(SIGB_ESHAPE, "reduction yields 2 channels, request has " + std::to_string(p->channels));
    vs.dst_node = i;
    int coff = 0;
    long long total = 0;
    for (int leaf : leaves) {
        VoiceSegSpec sg;
        int st = make_chain(leaf, sg.ch, false);
        if (st != SIGB_OK) return st;
        const int C = sg.ch.C;
        if (coff + C > Cin) return fail(SIGB_ESHAPE, "node " + std::to_string(i) + ": merged voices exceed the input width");
        std::vector<float> wl(C), wr(C);
        for (int c = 0; c < C; ++c) {
            const double g = sg.ch.has_gain ? sg.ch.gain_d[c] : 1.0;
            wl[c] = (float)(g * (1.0 - pv[coff + c]));
            wr[c] = (float)(g * pv[coff + c]);
        }
        sg.wl = put_vec(p, wl);
        sg.wr = put_vec(p, wr);
        sg.coff = coff;
        if (vs.pan_row >= 0) {
            std::vector<double> g64(C);
            for (int c = 0; c < C; ++c) g64[c] = sg.ch.has_gain ? sg.ch.gain_d[c] : 1.0;
            sg.gain64 = put_vec(p, g64);
        }

        sg.ch.gain_d.clear();
        coff += C;
        total += C;
        vs.segs.push_back(std::move(sg));
    }
    if (coff != Cin) return fail(SIGB_ESHAPE, "node " + std::to_string(i) + ": merged voices do not cover the input width");
    vs.M = p->opt_voices_m == 1 || p->opt_voices_m == 4 ? (int)p->opt_voices_m
                                                         : (total >= 8ll * SIGB_VOICE_THREADS * 4 ? 4 : 1);
    for (const VoiceSegSpec& sg : vs.segs) vs.nparts += sigb_voices_ctas(sg.ch.C, vs.M);
    BufInfo pb;
    pb.channels = 2 * vs.nparts;
    p->bufs.push_back(pb);
    vs.partial_buf = (int)p->bufs.size() - 1;
    Val v;
    v.kind = VK_BUF;
    v.channels = 2;
    v.buf = new_buf(i);
    if (v.buf >= 0) p->bufs[v.buf].channels = 2;
    p->vals[i] = v;
    // everything below the PanSum lives inside the fused kernel
    std::vector<int> stack{n.in[0]};
    while (!stack.empty()) {
        const int k = stack.back();
        stack.pop_back();
        p->vals[k].kind = VK_BUF;
        p->vals[k].channels = p->nodes[k].channels;
        p->vals[k].buf = -2;
        if (p->nodes[k].kind == SIGB_NODE_MERGE) {
            stack.push_back(p->nodes[k].in[0]);
            stack.push_back(p->nodes[k].in[1]);
        }
    }
    p->voices.push_back(std::move(vs));
    p->launches.push_back({LK_VOICES, (int)p->voices.size() - 1});
    return SIGB_OK;
}

bool Builder::stateless_osc(int i) const {
    int nsec = 0, wave = 0;
    return i >= 0 && pure_osc_run(i, &nsec, &wave) && nsec == 0;
}

// Mix / RingMod with a stateless oscillator chain on one side: that chain's launch evaluates the node as its
// epilogue (the other side is a second oscillator in registers, or a block read once), so "gain and mix fuse
// into their producers" instead of costing two materialised operands and a pointwise pass.
int Builder::build_fused_pointwise(int i, bool* done) {
    *done = false;
    const sigb_node& n = p->nodes[i];
    if (!p->opt_fuse_pointwise || (n.kind != SIGB_NODE_MIX && n.kind != SIGB_NODE_RINGMOD)) return SIGB_OK;
    const int C = node_C(i);
    const bool a_st = stateless_osc(n.in[0]), b_st = stateless_osc(n.in[1]);
    if (!a_st && !b_st) return SIGB_OK;
    if (n.in[0] == n.in[1]) return SIGB_OK;
    const int side = b_st ? 1 : 0;                       // which operand becomes the chain
    const int f = n.in[side], o = n.in[side ^ 1];
    const bool o_st = side == 1 ? a_st : false;          // both stateless: the other one is the in-register oscillator
    for (int k : {f, o}) {
        if (k < 0) continue;
        const int kc = p->nodes[k].channels;
        if (kc != 1 && kc != C)
            return fail(SIGB_ESHAPE, "node " + std::to_string(k) + ": block with " + std::to_string(kc) + " channels incompatible with requested " + std::to_string(C));
    }
    ChainSpec ch;
    int st = make_chain(f, ch, false, C);
    if (st != SIGB_OK) return st;
    ch.dst_node = i;
    ch.epi_op = n.kind == SIGB_NODE_MIX ? EW_MIX : EW_RINGMOD;
    ch.epi_side = side;
    if (n.kind == SIGB_NODE_MIX) {
        const std::vector<double>* m = const_of(p, n.in[2]);
        if (m) {
            std::vector<double> pv;
            if (!rep(*m, C, &pv)) return fail(SIGB_ESHAPE, "node " + std::to_string(i) + ": mix channels incompatible");
            ch.epi_p = put_vec(p, std::vector<float>(pv.begin(), pv.end()));
        } else {
            st = param_port(n.in[2], C, &ch.epi_p_row);
            if (st != SIGB_OK) return st;
        }
    }
    if (o_st) {
        ChainSpec oc;
        st = make_chain(o, oc, false, C);
        if (st != SIGB_OK) return st;
        if (oc.hertz_row >= 0) return fail(SIGB_EUNSUPPORTED, "internal: modulated oscillator in a fused epilogue");
        ch.epi_wave = oc.wave;
        ch.epi_hertz = oc.hertz;
        ch.epi_phase = oc.phase;
        ch.epi_gain = oc.gain;
        ch.epi_theta0 = oc.theta0;
        ch.epi_dtheta = oc.dtheta;
        mark_consumed(o);
    } else if (o >= 0) {
        st = ensure(o);
        if (st != SIGB_OK) return st;
        ch.epi_node = o;
    } else {
        ch.epi_node = -1;                                // unconnected: zeros(1,1)
    }
    mark_consumed(f);
    Val v;
    v.kind = VK_BUF;
    v.channels = C;
    v.buf = new_buf(i);
    if (v.buf >= 0) p->bufs[v.buf].channels = C;
    p->vals[i] = v;
    p->chains.push_back(ch);
    p->launches.push_back({LK_CHAIN, (int)p->chains.size() - 1});
    *done = true;
    return SIGB_OK;
}

int Builder::build_ewise(int i) {
    {
        bool done = false;
        int st = build_fused_pointwise(i, &done);
        if (st != SIGB_OK || done) return st;
    }
    const sigb_node& n = p->nodes[i];
    const int C = node_C(i);
    EwiseSpec e;
    e.C = C;
    e.dst_node = i;
    e.a_node = n.in[0];
    e.b_node = n.in[1];
    std::vector<double> pv;
    // block-rate parameter port: a constant row, or (modulated) a row of the parameter program
    auto param = [&](int port_node, const char* what) -> int {
        const std::vector<double>* m = const_of(p, port_node);
        if (!m) return param_port(port_node, C, &e.p_row);
        if (!rep(*m, C, &pv)) return fail(SIGB_ESHAPE, "node " + std::to_string(i) + ": " + what + " channels incompatible");
        return SIGB_OK;
    };
    if (n.kind == SIGB_NODE_MIX) {
        e.op = EW_MIX;
        int st = param(n.in[2], "mix");
        if (st != SIGB_OK) return st;
    } else if (n.kind == SIGB_NODE_RINGMOD) {
        e.op = EW_RINGMOD;
    } else {
        e.op = n.kind == SIGB_NODE_GAIN ? EW_GAIN : EW_AMP;
        e.b_node = -1;
        int st = param(n.in[1], n.kind == SIGB_NODE_GAIN ? "gain" : "exponent");
        if (st != SIGB_OK) return st;
    }
    for (int opnd : {e.a_node, e.b_node}) {
        if (opnd < 0) continue;
        int st = ensure(opnd);
        if (st != SIGB_OK) return st;
        const int oc = p->vals[opnd].channels;
        if (oc != 1 && oc != C)
            return fail(SIGB_ESHAPE, "node " + std::to_string(opnd) + ": block with " + std::to_string(oc) + " channels incompatible with requested " + std::to_string(C));
    }
    if (!pv.empty()) {
        std::vector<float> pf(pv.begin(), pv.end());
        e.p = put_vec(p, pf);
    }
    Val v;
    v.kind = VK_BUF;
    v.channels = C;
    v.buf = new_buf(i);
    if (v.buf >= 0) p->bufs[v.buf].channels = C;
    p->vals[i] = v;
    p->ewises.push_back(e);
    p->launches.push_back({LK_EWISE, (int)p->ewises.size() - 1});
    return SIGB_OK;
}

int Builder::build_merge(int i) {
    const sigb_node& n = p->nodes[i];
    if (n.in[0] < 0 || n.in[1] < 0)
        return fail(SIGB_ESHAPE, "node " + std::to_string(i) + ": Merge with an unplugged input (shape.py:69-72)");
    const int cl = p->nodes[n.in[0]].channels, cr = p->nodes[n.in[1]].channels;
    const int C = cl + cr;
    if (i == p->root && C != p->channels)
        return fail(SIGB_ESHAPE, "Merge yields " + std::to_string(C) + " channels, request has " + std::to_string(p->channels));
    Val v;
    v.kind = VK_BUF;
    v.channels = C;
    v.buf = new_buf(i);
    if (v.buf >= 0) p->bufs[v.buf].channels = C;
    // np.hstack (shape.py:73-74) rendered IN PLACE: a tree of Merge nodes is flattened into its leaves (a nested Merge nobody else
    // reads needs no block of its own), and a leaf that is the end of a chain nobody else reads writes its column range of the
    // merged block directly (out + first column, leading dimension of the merged block) -- one write per sample instead of a
    // write and a copy per Merge level.  Every other leaf is materialised as before and copied into place once.
    struct Leaf { int node, coff, width; };
    std::vector<Leaf> leaves;
    std::vector<int> inner;                                   // flattened Merge nodes below i
    {
        std::vector<Leaf> stack{{n.in[1], cl, cr}, {n.in[0], 0, cl}};
        while (!stack.empty()) {
            const Leaf t = stack.back();
            stack.pop_back();
            const sigb_node& m = p->nodes[t.node];
            if (m.kind == SIGB_NODE_MERGE && p->uses[t.node] == 1 && p->vals[t.node].kind == VK_NONE && m.in[0] >= 0 && m.in[1] >= 0 &&
                p->nodes[m.in[0]].channels + p->nodes[m.in[1]].channels == t.width) {
                const int wl = p->nodes[m.in[0]].channels;
                inner.push_back(t.node);
                stack.push_back({m.in[1], t.coff + wl, t.width - wl});
                stack.push_back({m.in[0], t.coff, wl});
            } else {
                leaves.push_back(t);
            }
        }
    }
    p->vals[i] = v;                                           // (the in-place chains resolve their destination through vals[i])
    for (const Leaf& lf : leaves) {
        const sigb_node& m = p->nodes[lf.node];
        const bool in_place = is_chain_kind(m.kind) && p->uses[lf.node] == 1 && p->vals[lf.node].kind == VK_NONE &&
                              m.channels == lf.width && (m.kind != SIGB_NODE_GAIN || gain_is_const(lf.node));
        if (in_place) {
            ChainSpec ch;
            int st = make_chain(lf.node, ch, true);
            if (st != SIGB_OK) return st;
            ch.dst_node = i;
            ch.dst_coff = lf.coff;
            p->vals[lf.node].kind = VK_BUF;                   // lives inside the merged block: never read on its own (uses == 1)
            p->vals[lf.node].channels = lf.width;
            p->vals[lf.node].buf = -2;
            p->chains.push_back(ch);
            p->launches.push_back({LK_CHAIN, (int)p->chains.size() - 1});
            continue;
        }
        int st = ensure(lf.node);
        if (st != SIGB_OK) return st;
        EwiseSpec e;
        e.op = EW_COPY;
        e.C = lf.width;
        if (p->vals[lf.node].channels != 1 && p->vals[lf.node].channels != e.C)
            return fail(SIGB_ESHAPE, "node " + std::to_string(lf.node) + ": channels incompatible with Merge slot");
        e.a_node = lf.node;
        e.dst_node = i;
        e.dst_coff = lf.coff;
        p->ewises.push_back(e);
        p->launches.push_back({LK_EWISE, (int)p->ewises.size() - 1});
    }
    for (int k : inner) {
        p->vals[k].kind = VK_BUF;
        p->vals[k].channels = p->nodes[k].channels;
        p->vals[k].buf = -2;
    }
    return SIGB_OK;
}

int Builder::build_reduce(int i) {
    const sigb_node& n = p->nodes[i];
    if (n.in[0] < 0) return fail(SIGB_ESHAPE, "node " + std::to_string(i) + ": reduction without input");
    // a pan driven by an emitter (block-rate modulation: an LFO sweeping the stereo position) is a row of the parameter
    // program, re-sampled at every request's first frame like every block-rate port (fused: build_voices; here: k_reduce
    // reads the row)
    const bool pan_modulated = n.kind == SIGB_NODE_PANSUM && const_of(p, n.in[1]) == nullptr;
    if (p->opt_fuse_reduce) {
        if (n.kind == SIGB_NODE_GROUPSUM) {
            int nsec = 0, wave = 0;
            if (pure_osc_run(n.in[0], &nsec, &wave) && nsec == 0 && wave == SIGB_WAVE_SINE) return build_bank(i);
        } else {
            std::vector<int> leaves;
            if (collect_voice_leaves(n.in[0], &leaves)) return build_voices(i, leaves);
        }
    }
    int st = ensure(n.in[0]);
    if (st != SIGB_OK) return st;
    ReduceSpec r;
    r.in_node = n.in[0];
    r.C = p->nodes[n.in[0]].channels;
    r.dst_node = i;
    const int vc = p->vals[n.in[0]].channels;
    if (vc != 1 && vc != r.C) return fail(SIGB_ESHAPE, "reduction input channels inconsistent");
    int outC;
    if (n.kind == SIGB_NODE_PANSUM) {
        r.pan = 1;
        r.groups = 2;
        outC = 2;
        if (pan_modulated) {
            st = param_port(n.in[1], r.C, &r.w_row);
            if (st != SIGB_OK) return st;
        } else {
            const std::vector<double>* pan = const_of(p, n.in[1]);
            std::vector<double> pv;
            if (!rep(*pan, r.C, &pv)) return fail(SIGB_ESHAPE, "node " + std::to_string(i) + ": pan channels incompatible");
            std::vector<float> pf(pv.begin(), pv.end());
            r.w = put_vec(p, pf);
        }
    } else {
        r.groups = n.order;
        outC = n.order;
        if (r.groups < 1 || r.C % r.groups != 0)
            return fail(SIGB_ESHAPE, "node " + std::to_string(i) + ": " + std::to_string(r.C) + " channels do not split into " + std::to_string(r.groups) + " groups");
    }
    if (i == p->root && outC != p->channels)
        return fail(SIGB_ESHAPE, "reduction yields " + std::to_string(outC) + " channels, request has " + std::to_string(p->channels));
    Val v;
    v.kind = VK_BUF;
    v.channels = outC;
    v.buf = new_buf(i);
    if (v.buf >= 0) p->bufs[v.buf].channels = outC;
    p->vals[i] = v;
    p->reduces.push_back(r);
    p->launches.push_back({LK_REDUCE, (int)p->reduces.size() - 1});
    return SIGB_OK;
}

// ---------------------------------------------------------------------------------------------
// runtime
// ---------------------------------------------------------------------------------------------

int upload(sigb_plan* p) {
    if (p->uploaded) return SIGB_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(SIGB_ECUDA, "no CUDA device: libsigb200 has no CPU fallback");
    // constants read at frame rate live in the arena as float rows
    for (size_t i = 0; i < p->vals.size(); ++i) {
        Val& v = p->vals[i];
        if (v.kind == VK_CONST && v.const_off < 0) {
            std::vector<float> f(v.cv.begin(), v.cv.end());
            v.const_off = arena_put(p, f.data(), f.size() * sizeof(float));
        }
    }
    {
        std::vector<float> z(1, 0.0f);
        p->zero_val.const_off = arena_put(p, z.data(), sizeof(float));
    }
    CUDA_TRY(cudaMalloc(&p->d_arena, std::max<size_t>(p->arena.size(), 256)));
    CUDA_TRY(cudaMemcpy(p->d_arena, p->arena.data(), p->arena.size(), cudaMemcpyHostToDevice));
    if (p->n_state > 0) {
        CUDA_TRY(cudaMalloc(&p->d_state, 2 * p->n_state * sizeof(double)));
        CUDA_TRY(cudaMemset(p->d_state, 0, 2 * p->n_state * sizeof(double)));
    }
    if (!p->pprog.empty()) {
        const size_t nrows = p->prow_const.size(), ps = (size_t)p->pwidth;
        std::vector<double> img(nrows * ps, 0.0);
        for (size_t r = 0; r < nrows; ++r) {
            const std::vector<double>& cv = p->prow_const[r];
            if (cv.empty()) continue;
            for (size_t c = 0; c < ps; ++c) img[r * ps + c] = cv[std::min(c, cv.size() - 1)];     // replicated to the row stride
        }
        CUDA_TRY(cudaMalloc(&p->d_prow_d, img.size() * sizeof(double)));
        CUDA_TRY(cudaMalloc(&p->d_prow_f, img.size() * sizeof(float)));
        CUDA_TRY(cudaMemcpy(p->d_prow_d, img.data(), img.size() * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemset(p->d_prow_f, 0, img.size() * sizeof(float)));
        CUDA_TRY(cudaMalloc(&p->d_pprog, p->pprog.size() * sizeof(ParamInstr)));
        CUDA_TRY(cudaMemcpy(p->d_pprog, p->pprog.data(), p->pprog.size() * sizeof(ParamInstr), cudaMemcpyHostToDevice));
    }
    if (p->n_mods > 0) {
        CUDA_TRY(cudaMalloc(&p->d_warm, p->n_mods * sizeof(int)));
        CUDA_TRY(cudaHostAlloc(&p->h_warm, p->n_mods * sizeof(int), cudaHostAllocDefault));
        std::memset(p->h_warm, 0, p->n_mods * sizeof(int));
        CUDA_TRY(cudaHostAlloc(&p->h_err, sizeof(int), cudaHostAllocMapped));
        *p->h_err = 0;
    }
    CUDA_TRY(cudaEventCreate(&p->ev0));
    CUDA_TRY(cudaEventCreate(&p->ev1));
    CUDA_TRY(cudaEventCreateWithFlags(&p->ev_caller, cudaEventDisableTiming));
    // the uploads above ran on the legacy default stream; renders may use any (non-blocking) stream
    CUDA_TRY(cudaDeviceSynchronize());
    p->uploaded = true;
    return SIGB_OK;
}

struct Operand {
    const float* ptr;
    int64_t ld;
    int cs;
    int64_t rows;   // <0: unlimited
};

// where node `idx`'s value for the current slab can be read
Operand operand_of(sigb_plan* p, int idx, int64_t abs_row0, const float* out, int64_t ld_out) {
    Operand o{nullptr, 0, 0, -1};
    if (idx < 0) {
        o.ptr = reinterpret_cast<const float*>(p->d_arena + p->zero_val.const_off);
        return o;
    }
    const Val& v = p->vals[idx];
    if (v.kind == VK_CONST) {
        o.ptr = reinterpret_cast<const float*>(p->d_arena + v.const_off);
        o.cs = v.channels == 1 ? 0 : 1;
    } else if (v.kind == VK_EXT) {
        const ExtBinding& e = p->ext[idx];
        o.ld = v.channels;
        o.cs = v.channels == 1 ? 0 : 1;
        const int64_t rel = abs_row0 - e.first_row;      // rows outside the bound window read as zero
        o.rows = rel < 0 ? 0 : std::max<int64_t>(0, e.rows - rel);
        o.ptr = e.ptr ? e.ptr + std::min(std::max<int64_t>(rel, 0), e.rows) * o.ld : nullptr;
        if (!e.ptr) o.rows = 0;
    } else if (v.buf < 0) {   // the root itself (only read by Merge-into-root bookkeeping; not expected)
        o.ptr = out;
        o.ld = ld_out;
        o.cs = 1;
    } else {
        o.ptr = p->bufs[v.buf].ptr;
        o.ld = v.channels;
        o.cs = v.channels == 1 ? 0 : 1;
    }
    return o;
}

int ensure_bufs(sigb_plan* p, int64_t rows) {
    for (BufInfo& b : p->bufs) {
        if (b.cap_rows >= rows) continue;
        if (b.ptr) cudaFree(b.ptr);
        b.ptr = nullptr;
        CUDA_TRY(cudaMalloc(&b.ptr, (size_t)rows * b.channels * sizeof(float)));
        b.cap_rows = rows;
    }
    return SIGB_OK;
}

// run every launch of the plan for rows [abs_row0, abs_row0 + rows) into `out`
int run_slab(sigb_plan* p, int64_t abs_row0, int rows, float* out, int64_t ld_out, cudaStream_t st) {
    const unsigned char* base = p->d_arena;
    for (const Launch& l : p->launches) {
        if (l.kind == LK_CHAIN) {
            ChainSpec& ch = p->chains[l.idx];
            ChainDev a;
            std::memset(&a, 0, sizeof(a));
            a.C = ch.C;
            a.src_kind = ch.src_kind;
            a.wave = ch.wave;
            a.nsec = ch.nsec;
            a.rate = p->rate;
            a.frames = rows;
            a.warm_rows = ch.warm_rows;
            a.warm_est = -1;
            a.position = abs_row0;
            a.pos_ptr = p->rt_pos_ptr;
            std::memcpy(a.sec_kind, ch.sec_kind, sizeof(a.sec_kind));
            a.hertz = ch.hertz_row >= 0 ? p->d_prow_d + (size_t)ch.hertz_row * p->pwidth : ch.hertz.dev<double>(base);
            a.phase = ch.phase_row >= 0 ? p->d_prow_d + (size_t)ch.phase_row * p->pwidth : ch.phase.dev<double>(base);
            a.theta0 = ch.theta0.dev<unsigned long long>(base);
            a.dtheta = ch.dtheta.dev<unsigned long long>(base);
            a.rot1 = ch.rot1.dev<float2>(base);
            a.constv = ch.constv.dev<float>(base);
            a.coef = ch.coef.dev<float>(base);
            a.gain = ch.gain_row >= 0 ? ch.gain_dev.dev<float>(base) : ch.gain.dev<float>(base);
            a.apow = ch.apow.dev<double>(base);
            a.apow_h = ch.apow_h.dev<double>(base);
            a.ztab = ch.ztab.dev<float>(base);
            a.m8 = ch.m8.dev<float>(base);
            a.hrec = ch.hrec.dev<float>(base);
            a.state = p->d_state ? p->d_state + ch.state_cur * p->n_state + ch.state_off : nullptr;
            a.state_out = p->d_state ? p->d_state + (ch.state_cur ^ 1) * p->n_state + ch.state_off : nullptr;
            a.src_rows = INT64_MAX;
            if (ch.src_kind == SRC_BUF) {
                Operand o = operand_of(p, ch.src_node, abs_row0, out, ld_out);
                a.src = o.ptr;
                a.src_ld = o.ld;
                a.src_cs = o.cs;
                a.src_rows = o.rows < 0 ? INT64_MAX : o.rows;
            }
            const Val& dv = p->vals[ch.dst_node];
            if (dv.buf < 0) { a.out = out + ch.dst_coff; a.ld_out = ld_out; }
            else { a.out = p->bufs[dv.buf].ptr + ch.dst_coff; a.ld_out = dv.channels; }
            a.epi_op = ch.epi_op;
            a.epi_wave = -1;
            if (ch.epi_op) {
                a.epi_side = ch.epi_side;
                a.epi_p = ch.epi_p_row >= 0 ? p->d_prow_f + (size_t)ch.epi_p_row * p->pwidth : ch.epi_p.dev<float>(base);
                if (ch.epi_wave >= 0) {
                    a.epi_wave = ch.epi_wave;
                    a.epi_hertz = ch.epi_hertz.dev<double>(base);
                    a.epi_phase = ch.epi_phase.dev<double>(base);
                    a.epi_gain = ch.epi_gain.dev<float>(base);
                    a.epi_theta0 = ch.epi_theta0.dev<unsigned long long>(base);
                    a.epi_dtheta = ch.epi_dtheta.dev<unsigned long long>(base);
                } else {
                    Operand o = operand_of(p, ch.epi_node, abs_row0, out, ld_out);
                    a.epi_buf = o.ptr; a.epi_ld = o.ld; a.epi_cs = o.cs; a.epi_rows = o.rows;
                }
                if (!p->opt_force_seq && p->rt_pos_ptr == nullptr && p->opt_osc_fill != 0 && ch.hertz_row < 0 && sigb_osc_fill_ok(&a)) {
                    int e = sigb_launch_osc_fill(&a, st);        // both oscillators from their phase words, in registers
                    if (e) return fail(SIGB_ECUDA, std::string("k_osc_fill: ") + cudaGetErrorString((cudaError_t)e));
                    p->launch_count++;
                    continue;
                }
                int e = sigb_launch_chain_seq(&a, st);           // epilogue chains are stateless: the sequential kernel tiles time
                if (e) return fail(SIGB_ECUDA, std::string("k_chain_seq: ") + cudaGetErrorString((cudaError_t)e));
                p->launch_count++;
                continue;
            }
            int done = 0;
            const bool force_seq = p->opt_force_seq || p->rt_pos_ptr != nullptr;    // realtime blocks: one sequential launch per chain
            if (ch.src_kind == SRC_OSC && ch.hertz_row < 0) a.guard = phase_guard(ch.max_abs_hertz, ch.max_abs_phase, abs_row0 + rows, p->rate);
            a.osc_mod = ch.src_kind == SRC_OSC && ch.hertz_row >= 0;
            // kernel choice: cascades of >= 3 sections run section-pipelined (k_cascade_pipe); shallower
            // chains stay on the time-parallel scan kernel, which measured faster for them (C2: 1.07e12 vs
            // 0.82e12 voice-samples/s) unless "cascade_pipe" forces the pipeline from n sections
            if (!force_seq && p->opt_cascade_pipe != 0 && ch.nsec_real >= 1 && ch.nsec_real <= 8) {
                ChainDev t = a;
                t.nsec = ch.nsec_real;           // identity padding sections are not run
                const bool deep = ch.nsec_real >= 3;
                const bool forced = p->opt_cascade_pipe > 0 && ch.nsec_real >= (int)p->opt_cascade_pipe;
                // deep cascades on a materialised block keep every section in registers (k_cascade_reg)
                // (two sections: only chains whose decay horizon the host knows -- unmodulated, or modulated in a large request, run_params --
                // and whose time pieces fill at least half of the machine's warp slots)
                if (p->opt_cascade_reg != 0 && sigb_cascade_reg_ok(&t) &&
                    (ch.nsec_real >= 3 || (t.warm_rows >= 0 && sigb_cascade_reg_fill(&t, (int)p->opt_pipe_segments, (int)p->opt_reg_variant) >= 512))) {
                    int e = sigb_launch_cascade_reg(&t, (int)p->opt_pipe_segments, (int)p->opt_reg_variant, st);
                    if (e) return fail(SIGB_ECUDA, std::string("k_cascade_reg: ") + cudaGetErrorString((cudaError_t)e));
                    p->launch_count++;
                    ch.state_cur ^= 1;           // the kernel wrote the other copy of the state
                    continue;
                }
                // oscillator-fed chains from "osc_reg" sections on; 2-section chains only when their time pieces fill at least
                // half of the machine's warp slots (else the time-parallel scan kernel is the better choice)
                // ... and ONE section behind a Square / Sawtooth / Triangle (the scan kernels evaluate those in float64 per sample:
                // 4.5-6.8e11 voice-samples/s on C2's shape against 1.58e12 for a Sine) under the same condition
                // ... or behind any oscillator on more channels than the scan kernels take ("scan_max_tiles"): there the alternative
                // is k_chain_seq's float64 oscillator, and with at least one tile per warp slot no time piece needs a warm-up
                const bool one_nonsine = ch.nsec_real == 1 && (ch.wave != SIGB_WAVE_SINE || (ch.C + 31) / 32 > p->opt_scan_max_tiles) &&
                                         p->opt_osc_reg == 2 && p->opt_osc_delta != 0;
                if (p->opt_osc_reg > 0 && (ch.nsec_real >= (int)p->opt_osc_reg || one_nonsine) && sigb_osc_reg_ok(&t, (int)p->opt_osc_delta) &&
                    (ch.nsec_real >= 3 || (p->osc_reg_user && !one_nonsine) ||
                     (t.warm_rows >= 0 && sigb_osc_reg_fill(&t, (int)p->opt_pipe_segments, (int)p->opt_osc_delta) >= 512))) {
                    int e = sigb_launch_osc_reg(&t, (int)p->opt_pipe_segments, (int)p->opt_osc_delta, st);
                    if (e) return fail(SIGB_ECUDA, std::string("k_osc_reg: ") + cudaGetErrorString((cudaError_t)e));
                    p->launch_count++;
                    ch.state_cur ^= 1;
                    continue;
                }
                if (sigb_cascade_pipe_ok(&t) &&
                    (deep || forced)) {
                    int e = sigb_launch_cascade_pipe(&t, (int)p->opt_pipe_segments, (int)p->opt_pipe_spw, st);
                    if (e) return fail(SIGB_ECUDA, std::string("k_cascade_pipe: ") + cudaGetErrorString((cudaError_t)e));
                    p->launch_count++;
                    ch.state_cur ^= 1;           // the kernel wrote the other copy of the state
                    continue;
                }
            }
            // stateless oscillator chains on many channels: from the Q0.64 phase word instead of float64 per sample (k_osc_fill)
            if (!force_seq && p->opt_osc_fill != 0 && (ch.hertz_row < 0 || ch.osc_tables_dev) && sigb_osc_fill_ok(&a)) {
                int e = sigb_launch_osc_fill(&a, st);
                if (e) return fail(SIGB_ECUDA, std::string("k_osc_fill: ") + cudaGetErrorString((cudaError_t)e));
                p->launch_count++;
                continue;
            }
            const int tiles = (ch.C + 31) / 32;
            // (a chain with a modulated cutoff gets its scan tables from k_design, once per request)
            const bool scan_ok = !force_seq && ch.nsec >= 1 && ch.nsec <= 8 && tiles <= p->opt_scan_max_tiles;
            if (scan_ok) {
                if (!ch.mods.empty() && p->warm_on_device && ch.warm_static < 1e8) {
                    // the scan kernels read the request's decay horizon on the device (k_design wrote it on this stream);
                    // the host's estimate -- what earlier requests copied back, possibly stale -- only decides whether
                    // to cut tiles along time
                    a.warm_rows = (int)std::ceil(ch.warm_static);
                    a.warm_dev = p->d_warm + ch.mods.front().slot;
                    a.n_warm_dev = (int)ch.mods.size();
                    double est = ch.warm_static;
                    for (const ChainSpec::ModFilter& mf : ch.mods) est += p->h_warm[mf.slot];
                    a.warm_est = (int)std::min(est, 1e9);
                }
                int e = sigb_launch_chain_scan(&a, (int)p->opt_scan_variant, st, &done);
                if (e) return fail(SIGB_ECUDA, std::string("k_chain_scan: ") + cudaGetErrorString((cudaError_t)e));
                if (done > 0) {
                    p->launch_count++;
                    ch.state_cur ^= 1;                // the time-parallel kernels write the other copy of the state
                    std::swap(a.state, a.state_out);
                }
            }
            if (done < rows) {
                ChainDev t = a;
                t.frames = rows - done;
                t.position = abs_row0 + done;
                t.out = a.out + (int64_t)done * a.ld_out;
                if (ch.src_kind == SRC_BUF) {
                    t.src = a.src + (int64_t)done * a.src_ld;
                    t.src_rows = a.src_rows == INT64_MAX ? INT64_MAX : std::max<int64_t>(0, a.src_rows - done);
                }
                int e = sigb_launch_chain_seq(&t, st);
                if (e) return fail(SIGB_ECUDA, std::string("k_chain_seq: ") + cudaGetErrorString((cudaError_t)e));
                p->launch_count++;
            }
        } else if (l.kind == LK_EWISE) {
            const EwiseSpec& e = p->ewises[l.idx];
            EwiseDev a;
            std::memset(&a, 0, sizeof(a));
            a.op = e.op;
            a.C = e.C;
            a.frames = rows;
            const Val& dv = p->vals[e.dst_node];
            if (dv.buf < 0) { a.out = out + e.dst_coff; a.ld_out = ld_out; }
            else { a.out = p->bufs[dv.buf].ptr + e.dst_coff; a.ld_out = dv.channels; }
            Operand oa = operand_of(p, e.a_node, abs_row0, out, ld_out);
            Operand ob = operand_of(p, e.b_node, abs_row0, out, ld_out);
            a.a = oa.ptr; a.lda = oa.ld; a.acs = oa.cs; a.a_rows = oa.rows;
            a.b = ob.ptr; a.ldb = ob.ld; a.bcs = ob.cs; a.b_rows = ob.rows;
            a.p = e.p_row >= 0 ? p->d_prow_f + (size_t)e.p_row * p->pwidth : e.p.dev<float>(base);
            int err = sigb_launch_ewise(&a, st);
            if (err) return fail(SIGB_ECUDA, std::string("k_ewise: ") + cudaGetErrorString((cudaError_t)err));
            p->launch_count++;
        } else if (l.kind == LK_BANK) {
            const BankSpec& b = p->banks[l.idx];
            BankDev a;
            std::memset(&a, 0, sizeof(a));
            a.P = b.ch.C;
            a.groups = b.groups;
            a.frames = rows;
            a.position = abs_row0;
            a.theta0 = b.ch.theta0.dev<unsigned long long>(base);
            a.dtheta = b.ch.dtheta.dev<unsigned long long>(base);
            a.gain = b.ch.gain.dev<float>(base);
            a.rot32 = b.rot32.dev<float2>(base);
            const Val& dv = p->vals[b.dst_node];
            if (dv.buf < 0) { a.out = out; a.ld_out = ld_out; }
            else { a.out = p->bufs[dv.buf].ptr; a.ld_out = dv.channels; }
            int err = sigb_launch_bank(&a, st);
            if (err) return fail(SIGB_ECUDA, std::string("k_bank: ") + cudaGetErrorString((cudaError_t)err));
            p->launch_count++;
        } else if (l.kind == LK_VOICES) {
            VoicesSpec& vs = p->voices[l.idx];
            float* partial = p->bufs[vs.partial_buf].ptr;
            // decay horizon of the bank's filters (2^-40): what a piece that starts inside a voice group re-renders
            // from zero state before its first stored row; oscillators need no warm-up at all
            int warm = 0;
            for (const VoiceSegSpec& sg : vs.segs) {
                if (sg.ch.nsec_real == 0) continue;
                warm = sg.ch.warm_rows < 0 ? -1 : (warm < 0 ? -1 : std::max(warm, sg.ch.warm_rows));
                if (warm < 0) break;
            }
            // ch.warm_rows budgets 2 x 40 bits of decay (doubled for near-defective cascades).  The voices of this kernel
            // have ONE Butterworth section each (complex pole pair of radius rho: the forgotten state decays like rho^k times
            // a transient factor <= ~1/theta, 2^6 at 100 Hz), and a voice enters the mix with weight ~1/sqrt(N): 36 bits
            // leave every voice's start-up error below 2^-30 of its own amplitude, and cut the warm-up (12 % of the
            // instructions at 131,072 instances per GPU with the 80-bit budget) by more than half.
            if (warm > 0) warm = (int)std::ceil(warm * 0.45) + 16;
            const int VK = sigb_voices_block_rows(vs.M);
            const int64_t bpg = (rows + VK - 1) / VK;
            int part0 = 0;
            for (size_t s0 = 0; s0 < vs.segs.size(); s0 += SIGB_VOICE_SEGS) {
                VoicesDev a;
                std::memset(&a, 0, sizeof(a));
                a.nseg = (int)std::min<size_t>(SIGB_VOICE_SEGS, vs.segs.size() - s0);
                a.rate = p->rate;
                a.frames = rows;
                a.M = vs.M;
                a.position = abs_row0;
                a.partial = partial + (int64_t)part0 * rows * 2;
                a.warm_rows = std::max(warm, 0);
                int groups = 0;
                for (int k = 0; k < a.nseg; ++k) {
                    const VoiceSegSpec& sg = vs.segs[s0 + k];
                    VoiceSeg& d = a.seg[k];
                    d.C = sg.ch.C;
                    d.wave = sg.ch.wave;
                    d.nsec = sg.ch.nsec_real;
                    d.sec_kind = sg.ch.sec_kind[0];
                    d.cta0 = groups;
                    d.theta0 = sg.ch.theta0.dev<unsigned long long>(base);
                    d.dtheta = sg.ch.dtheta.dev<unsigned long long>(base);
                    d.hertz = sg.ch.hertz.dev<double>(base);
                    d.phase = sg.ch.phase.dev<double>(base);
                    d.coef = sg.ch.coef.dev<float>(base);
                    d.wl = sg.wl.dev<float>(base);
                    d.wr = sg.wr.dev<float>(base);
                    d.state = p->d_state ? p->d_state + vs.state_cur * p->n_state + sg.ch.state_off : nullptr;
                    d.state_out = p->d_state ? p->d_state + (vs.state_cur ^ 1) * p->n_state + sg.ch.state_off : nullptr;
                    d.guard = phase_guard(sg.ch.max_abs_hertz, sg.ch.max_abs_phase, abs_row0 + rows, p->rate);
                    groups += sigb_voices_ctas(sg.ch.C, vs.M);
                }
                // pieces: equal contiguous shares of the (group, row block) space.  With a known decay horizon the
                // launch is cut into `voices_pieces` pieces per resident CTA slot (more than one, so that the hardware
                // scheduler evens out the cost differences between wave / filter kinds), as long as the warm-up of a
                // piece that starts inside a group stays below 1/8 of the piece; otherwise one piece per group.
                a.ngroups = groups;
                int64_t np = groups;
                if (warm >= 0 && p->opt_voices_segments != 1) {
                    const int64_t total = (int64_t)groups * bpg;
                    const int64_t target = p->opt_voices_segments > 1 ? (int64_t)groups * p->opt_voices_segments
                                                                      : (int64_t)sigb_voices_slots(vs.M) * std::max<int64_t>(1, p->opt_voices_pieces);
                    const int64_t fit = warm > 0 ? std::max<int64_t>(1, total * VK / (8ll * warm)) : total;
                    np = std::max<int64_t>(1, std::min(std::min(target, fit), total));
                    if (fit < groups) np = groups;       // too short to cut inside a group: pieces = whole groups, no warm-up
                }
                a.npieces = (int)np;
                int err = sigb_launch_voices(&a, st);
                if (err) return fail(SIGB_ECUDA, std::string("k_voices: ") + cudaGetErrorString((cudaError_t)err));
                p->launch_count++;
                part0 += groups;
            }
            vs.state_cur ^= 1;
            const Val& dv = p->vals[vs.dst_node];
            float* o = dv.buf < 0 ? out : p->bufs[dv.buf].ptr;
            const int64_t ldo = dv.buf < 0 ? ld_out : dv.channels;
            int err = sigb_launch_voices_finish(partial, vs.nparts, rows, o, ldo, st);
            if (err) return fail(SIGB_ECUDA, std::string("k_voices_finish: ") + cudaGetErrorString((cudaError_t)err));
            p->launch_count++;
        } else {
            const ReduceSpec& r = p->reduces[l.idx];
            ReduceDev a;
            std::memset(&a, 0, sizeof(a));
            a.C = r.C;
            a.groups = r.groups;
            a.frames = rows;
            a.pan = r.pan;
            Operand oi = operand_of(p, r.in_node, abs_row0, out, ld_out);
            a.in = oi.ptr; a.ld_in = oi.ld; a.ics = oi.cs; a.in_rows = oi.rows;
            a.w = r.w_row >= 0 ? p->d_prow_f + (size_t)r.w_row * p->pwidth : r.w.dev<float>(base);
            const Val& dv = p->vals[r.dst_node];
            if (dv.buf < 0) { a.out = out; a.ld_out = ld_out; }
            else { a.out = p->bufs[dv.buf].ptr; a.ld_out = dv.channels; }
            int err = sigb_launch_reduce(&a, st);
            if (err) return fail(SIGB_ECUDA, std::string("k_reduce: ") + cudaGetErrorString((cudaError_t)err));
            p->launch_count++;
        }
    }
    return SIGB_OK;
}

int64_t slab_rows(sigb_plan* p, int64_t frames) {
    if (p->opt_slab_frames > 0) return std::min<int64_t>(frames, p->opt_slab_frames);
    int64_t chsum = 0;
    for (const BufInfo& b : p->bufs) chsum += b.channels;
    if (chsum == 0) return frames;
    int64_t rows = p->opt_buffer_budget / (4 * chsum);
    const int64_t step = 7 * SIGB_SCAN_L * 4;    // keep slabs a multiple of every scan step size
    rows = std::max<int64_t>(step, rows / step * step);
    return std::min(frames, rows);
}

// Block-rate parameters of one request: the parameter program at `position` (k_param_eval), then -- design = true --
// the per-request design of every filter whose cutoff is modulated (k_design).  `frames` sizes the decision whether
// the request is long enough for the time-parallel kernels: only then is the decay horizon read back (one
// synchronisation of `st`), otherwise those chains run unsegmented.
int run_params(sigb_plan* p, int64_t position, int64_t frames, cudaStream_t st, bool design = true) {
    if (p->pprog.empty()) return SIGB_OK;
    int e = sigb_launch_param_eval(p->d_pprog, (int)p->pprog.size(), (int)p->prow_const.size(), p->d_prow_d, p->d_prow_f, p->pwidth,
                                   position, p->rt_pos_ptr, p->rate, st);
    if (e) return fail(SIGB_ECUDA, std::string("k_param_eval: ") + cudaGetErrorString((cudaError_t)e));
    p->launch_count++;
    for (const ChainSpec& ch : p->chains) {
        if (ch.gain_row < 0) continue;
        const unsigned char* base = p->d_arena;
        e = sigb_launch_gain_rows(ch.C, ch.gain.dev<float>(base), p->d_prow_d + (size_t)ch.gain_row * p->pwidth,
                                  const_cast<float*>(ch.gain_dev.dev<float>(base)), st);
        if (e) return fail(SIGB_ECUDA, std::string("k_gain_rows: ") + cudaGetErrorString((cudaError_t)e));
        p->launch_count++;
    }
    for (const ChainSpec& ch : p->chains) {
        if (!ch.osc_tables_dev) continue;
        const unsigned char* base = p->d_arena;
        e = sigb_launch_osc_tables(ch.C, p->d_prow_d + (size_t)ch.hertz_row * p->pwidth, p->d_prow_d + (size_t)ch.phase_row * p->pwidth, p->rate,
                                   const_cast<unsigned long long*>(ch.theta0.dev<unsigned long long>(base)),
                                   const_cast<unsigned long long*>(ch.dtheta.dev<unsigned long long>(base)),
                                   const_cast<float*>(ch.rot1.dev<float>(base)), st);
        if (e) return fail(SIGB_ECUDA, std::string("k_osc_tables: ") + cudaGetErrorString((cudaError_t)e));
        p->launch_count++;
    }
    for (const VoicesSpec& vs : p->voices) {
        if (vs.pan_row < 0) continue;
        const unsigned char* base = p->d_arena;
        for (const VoiceSegSpec& sg : vs.segs) {
            e = sigb_launch_pan_weights(sg.ch.C, sg.gain64.dev<double>(base), p->d_prow_d + (size_t)vs.pan_row * p->pwidth + sg.coff,
                                        const_cast<float*>(sg.wl.dev<float>(base)), const_cast<float*>(sg.wr.dev<float>(base)), st);
            if (e) return fail(SIGB_ECUDA, std::string("k_pan_weights: ") + cudaGetErrorString((cudaError_t)e));
            p->launch_count++;
        }
    }
    if (!design || p->n_mods == 0) return SIGB_OK;
    const bool want_warm = p->rt_pos_ptr == nullptr && !p->opt_force_seq && frames >= 4096;
    p->warm_on_device = want_warm;
    if (want_warm) CUDA_TRY(cudaMemsetAsync(p->d_warm, 0, p->n_mods * sizeof(int), st));
    bool need_host = false;         // deep cascades run the register / pipelined kernels, whose launch geometry needs the value
    for (ChainSpec& ch : p->chains) {
        for (const ChainSpec::ModFilter& mf : ch.mods) {
            DesignDev d;
            std::memset(&d, 0, sizeof(d));
            d.C = ch.C; d.s0 = mf.s0; d.order = mf.order; d.highpass = mf.highpass; d.rate = p->rate;
            d.cutoff = p->d_prow_d + (size_t)mf.row * p->pwidth;
            d.coef = reinterpret_cast<float*>(p->d_arena + ch.coef.off);
            if (ch.apow.off >= 0) {
                d.apow = reinterpret_cast<double*>(p->d_arena + ch.apow.off);
                d.apow_h = reinterpret_cast<double*>(p->d_arena + ch.apow_h.off);
                d.ztab = reinterpret_cast<float*>(p->d_arena + ch.ztab.off);
                d.m8 = reinterpret_cast<float*>(p->d_arena + ch.m8.off);
                d.hrec = reinterpret_cast<float*>(p->d_arena + ch.hrec.off);
            }
            d.err_flag = p->h_err;
            d.warm_out = want_warm ? p->d_warm + mf.slot : nullptr;
            e = sigb_launch_design(&d, st);
            if (e) return fail(SIGB_ECUDA, std::string("k_design: ") + cudaGetErrorString((cudaError_t)e));
            p->launch_count++;
        }
        if (!ch.mods.empty()) {
            ch.warm_rows = -1;
            if (ch.nsec_real >= 3 || p->opt_cascade_pipe > 0 || (p->osc_reg_user && p->opt_osc_reg > 0 && ch.nsec_real >= (int)p->opt_osc_reg)) need_host = true;
            // two sections: the register kernels are ~1.9x the scan kernel's rate, worth one synchronisation when
            // the request is large (>= 2^28 samples, a render of >= 0.25 ms)
            if (ch.nsec_real == 2 && p->opt_cascade_pipe != 0 && frames * (int64_t)ch.C >= (1ll << 28)) need_host = true;
        }
    }
    if (want_warm) {
        // the copy also refreshes the host's estimate for the NEXT request's scan launches; only deep cascades wait for it
        CUDA_TRY(cudaMemcpyAsync(p->h_warm, p->d_warm, p->n_mods * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (need_host) {
            CUDA_TRY(cudaStreamSynchronize(st));
            for (ChainSpec& ch : p->chains) {
                if (ch.mods.empty()) continue;
                double w = ch.warm_static;
                for (const ChainSpec::ModFilter& mf : ch.mods) w += p->h_warm[mf.slot];
                ch.warm_rows = w < 1e8 ? (int)std::ceil(w) : -1;
            }
        }
    }
    return SIGB_OK;
}

// scipy's ValueError for a modulated cutoff outside (0, Nyquist) (fx.py:102), raised by the first call that finds the
// flag k_design left in page-locked memory
int check_design_error(sigb_plan* p) {
    if (p->h_err && *reinterpret_cast<volatile int*>(p->h_err)) {
        *p->h_err = 0;
        p->have_pos = false;
        return fail(SIGB_ECRIT, "Digital filter critical frequencies must be 0 < Wn < 1 (modulated cutoff)");
    }
    return SIGB_OK;
}

// device-slab loop of one request (no seek handling, no parameters)
int run_rows(sigb_plan* p, int64_t position, int64_t frames, float* out, int64_t ld_out, cudaStream_t st) {
    const int64_t slab = slab_rows(p, frames);
    int e = ensure_bufs(p, slab);
    if (e != SIGB_OK) return e;
    for (int64_t r = 0; r < frames; r += slab) {
        const int rows = (int)std::min(slab, frames - r);
        e = run_slab(p, position + r, rows, out + r * ld_out, ld_out, st);
        if (e != SIGB_OK) return e;
    }
    p->last_slab = slab;
    return SIGB_OK;
}

// start of a request: seek handling + the block-rate parameters, sampled ONCE at the request's first frame
// (forward_at_block_rate, chain/__init__.py:305-306) however the request is later cut into slabs
int begin_request(sigb_plan* p, int64_t position, int64_t frames, cudaStream_t st) {
    if (!p->have_pos || p->next_pos != position || p->opt_restart) {
        // seek: zero state, then warm the filters up on `context` frames (fx.py:93-105)
        if (p->d_state) CUDA_TRY(cudaMemsetAsync(p->d_state, 0, 2 * p->n_state * sizeof(double), st));
        int64_t pre = std::min<int64_t>(p->context, position);
        if (pre > 0 && p->n_state > 0) {
            const int64_t need = std::min<int64_t>(pre, slab_rows(p, pre)) * p->channels;
            if (p->scratch_floats < need) {
                if (p->scratch) cudaFree(p->scratch);
                p->scratch = nullptr;
                CUDA_TRY(cudaMalloc(&p->scratch, need * sizeof(float)));
                p->scratch_floats = need;
            }
            // the filter samples its cutoff at the REQUEST's position (fx.py:124-129) and runs context + block with
            // that one design; the context request it sends upstream samples ITS parameters at its own position
            int e = run_params(p, position, frames, st, true);
            if (e == SIGB_OK) e = run_params(p, position - pre, frames, st, false);
            if (e != SIGB_OK) return e;
            // warm-up in pieces no longer than a device slab (the intermediate buffers hold one slab)
            const int64_t slab = slab_rows(p, pre);
            e = ensure_bufs(p, slab);
            if (e != SIGB_OK) return e;
            for (int64_t r = 0; r < pre; r += slab) {
                e = run_slab(p, position - pre + r, (int)std::min(slab, pre - r), p->scratch, p->channels, st);
                if (e != SIGB_OK) return e;
            }
        }
    }
    return run_params(p, position, frames, st);
}

void end_request(sigb_plan* p, int64_t position, int64_t frames) {
    p->have_pos = true;
    p->next_pos = position + frames;
    p->last_position = position;
    p->last_frames = frames;
}

// the body of sigb_render: seek handling + parameters + slab loop, all on `st`
int render_range(sigb_plan* p, int64_t position, int64_t frames, float* out, int64_t ld_out, cudaStream_t st) {
    int e = begin_request(p, position, frames, st);
    if (e == SIGB_OK) e = run_rows(p, position, frames, out, ld_out, st);
    if (e != SIGB_OK) return e;
    end_request(p, position, frames);
    return SIGB_OK;
}

int ensure_host_streams(sigb_plan* plan) {
    if (plan->s_render) return SIGB_OK;
    CUDA_TRY(cudaStreamCreateWithFlags(&plan->s_render, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&plan->s_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&plan->ev_rendered[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&plan->ev_copied[i], cudaEventDisableTiming));
    }
    return SIGB_OK;
}

void rt_drop_graphs(sigb_plan* p) {
    for (sigb_plan::RtGraph& g : p->rt_graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
    }
    p->rt_graphs.clear();
}

// Whether the realtime block path can take this plan: every launch must be position-independent once the position
// comes from the block header (k_chain_seq, k_ewise, k_reduce, k_param_eval, k_design), i.e. no fused bank / voices
// kernels and no Buffer source (its window pointer depends on the position).
// which copy of the double-buffered filter state every launch currently reads: a captured graph holds these pointers
uint64_t rt_state_sig(const sigb_plan* p) {
    uint64_t h = 1469598103934665603ull;
    for (const ChainSpec& ch : p->chains) h = (h ^ (uint64_t)(ch.state_cur + 1)) * 1099511628211ull;
    for (const VoicesSpec& v : p->voices) h = (h ^ (uint64_t)(v.state_cur + 1)) * 1099511628211ull;
    return h;
}

bool rt_eligible(sigb_plan* p) {
    if (p->rt_state == 0) {
        bool ok = true;
        for (const Launch& l : p->launches)
            if (l.kind == LK_BANK || l.kind == LK_VOICES) ok = false;
        for (const Val& v : p->vals)
            if (v.kind == VK_EXT) ok = false;
        for (const ChainSpec& ch : p->chains)
            if (ch.nsec > SIGB_MAX_SEC) ok = false;
        p->rt_state = ok ? 1 : -1;
    }
    return p->rt_state > 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------

extern "C" int sigb_abi_version(void) { return SIGB_ABI_VERSION; }

extern "C" const char* sigb_last_error(void) { return g_last_error.c_str(); }

extern "C" const char* sigb_strerror(int status) {
    switch (status) {
        case SIGB_OK: return "ok";
        case SIGB_EINVAL: return "invalid argument";
        case SIGB_ESHAPE: return "block shape incompatible with the requested shape";
        case SIGB_EINDEX: return "filter cutoff/input narrower than the requested channels";
        case SIGB_ECRIT: return "Digital filter critical frequencies must be 0 < Wn < 1";
        case SIGB_EUNSUPPORTED: return "graph not supported by the B200 evaluator";
        case SIGB_ECUDA: return "CUDA failure";
        case SIGB_ENOMEM: return "out of memory";
        case SIGB_ESTATE: return "plan in invalid state";
        default: return "unknown status";
    }
}

extern "C" int sigb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int sigb_plan_create(const sigb_node* nodes, int32_t n_nodes, int32_t root, const double* data,
                                int64_t n_data, int32_t channels, int32_t rate, sigb_plan** out_plan) {
    if (!nodes || !out_plan || n_nodes <= 0 || root < 0 || root >= n_nodes || channels < 1 || rate < 1)
        return fail(SIGB_EINVAL, "sigb_plan_create: bad arguments");
    std::unique_ptr<sigb_plan> p(new sigb_plan());
    p->opt_fuse_reduce = g_default_fuse_reduce;
    p->opt_voices_m = g_default_voices_m;
    p->opt_fuse_pointwise = g_default_fuse_pointwise;
    p->channels = channels;
    p->rate = rate;
    p->root = root;
    p->nodes.assign(nodes, nodes + n_nodes);
    if (data && n_data > 0) p->data.assign(data, data + n_data);
    p->vals.resize(n_nodes);
    p->uses.assign(n_nodes, 0);
    p->node_ctx.assign(n_nodes, 0);
    p->ext.resize(n_nodes);
    p->zero_val.kind = VK_CONST;
    p->zero_val.channels = 1;
    p->zero_val.cv.assign(1, 0.0);
    for (int i = 0; i < n_nodes; ++i) {
        const sigb_node& n = p->nodes[i];
        if (n.kind < SIGB_NODE_ZERO || n.kind > SIGB_NODE_TAP) return fail(SIGB_EINVAL, "node " + std::to_string(i) + ": unknown kind");
        if (n.channels < 1) return fail(SIGB_ESHAPE, "node " + std::to_string(i) + ": channels < 1");
        for (int k = 0; k < 3; ++k)
            if (n.in[k] >= i || n.in[k] < -1) return fail(SIGB_EINVAL, "node " + std::to_string(i) + ": inputs must precede the node (topological order)");
        Val& v = p->vals[i];
        if (n.kind == SIGB_NODE_ZERO) {
            v = p->zero_val;
        } else if (n.kind == SIGB_NODE_FIXED) {
            if (n.rows != 1)
                return fail(SIGB_EUNSUPPORTED, "node " + std::to_string(i) + ": Fixed with " + std::to_string(n.rows) + " rows (frame-rate tables go through a Buffer node)");
            if (n.data_off < 0 || n.data_off + n.channels > (int64_t)p->data.size())
                return fail(SIGB_EINVAL, "node " + std::to_string(i) + ": data_off out of range");
            v.kind = VK_CONST;
            v.channels = n.channels;
            v.cv.assign(p->data.begin() + n.data_off, p->data.begin() + n.data_off + n.channels);
        } else if (n.kind == SIGB_NODE_BUFFER) {
            v.kind = VK_EXT;
            v.channels = n.channels;
        } else if (n.kind == SIGB_NODE_TAP) {
            p->tap_nodes.push_back(i);            // taps are numbered in record order
        }
        // frame-rate edges (block-rate parameter ports are constants and never materialise)
        int nsig = 0;
        switch (n.kind) {
            case SIGB_NODE_GAIN: case SIGB_NODE_AMP: case SIGB_NODE_FILTER: case SIGB_NODE_TAP:
            case SIGB_NODE_GROUPSUM: case SIGB_NODE_PANSUM: nsig = 1; break;
            case SIGB_NODE_MIX: case SIGB_NODE_RINGMOD: case SIGB_NODE_MERGE: nsig = 2; break;
            default: break;
        }
        int ctx = 0;
        for (int k = 0; k < nsig; ++k)
            if (n.in[k] >= 0) {
                p->uses[n.in[k]]++;
                ctx = std::max(ctx, p->node_ctx[n.in[k]]);
            }
        p->node_ctx[i] = ctx + (n.kind == SIGB_NODE_FILTER ? std::max(0, n.context) : 0);
    }
    p->context = p->node_ctx[root];
    p->uses[root]++;
    const sigb_node& rn = p->nodes[root];
    if (rn.kind == SIGB_NODE_TAP) return fail(SIGB_EINVAL, "a tap cannot be the root (the root block is the caller's `out`)");
    if (rn.channels != 1 && rn.channels != channels)
        return fail(SIGB_ESHAPE, "root block with " + std::to_string(rn.channels) + " channels incompatible with requested " + std::to_string(channels));
    Builder b{p.get()};
    if (p->vals[root].kind == VK_NONE) {
        int st = b.ensure(root);
        if (st != SIGB_OK) return st;
    } else {
        // constant or external root: broadcast-copy it into the output
        EwiseSpec e;
        e.op = EW_COPY;
        e.C = channels;
        e.a_node = root;
        e.dst_node = root;
        p->ewises.push_back(e);
        p->launches.push_back({LK_EWISE, (int)p->ewises.size() - 1});
    }
    *out_plan = p.release();
    return SIGB_OK;
}


extern "C" int sigb_plan_bind_buffer(sigb_plan* plan, int32_t node, const float* dev_ptr, int64_t rows) {
    if (!plan || node < 0 || node >= (int)plan->nodes.size() || plan->nodes[node].kind != SIGB_NODE_BUFFER)
        return fail(SIGB_EINVAL, "sigb_plan_bind_buffer: not a Buffer node");
    plan->ext[node].ptr = dev_ptr;
    plan->ext[node].first_row = 0;
    plan->ext[node].rows = rows;
    return SIGB_OK;
}

extern "C" int sigb_plan_bind_buffer_window(sigb_plan* plan, int32_t node, const float* dev_ptr, int64_t first_row, int64_t rows) {
    int e = sigb_plan_bind_buffer(plan, node, dev_ptr, rows);
    if (e != SIGB_OK) return e;
    plan->ext[node].first_row = first_row;
    return SIGB_OK;
}

extern "C" int sigb_render(sigb_plan* plan, int64_t position, int32_t frames, float* out, int64_t ld_out, void* stream) {
    if (!plan || (!out && frames != 0) || frames < 0 || position < 0 || ld_out < plan->channels)
        return fail(SIGB_EINVAL, "sigb_render: bad arguments");
    if (frames == 0) return SIGB_OK;          // an empty block (shape (0, C)): nothing to render, the stream position stands
    int e = upload(plan);
    if (e == SIGB_OK) e = check_design_error(plan);
    if (e != SIGB_OK) return e;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaEventRecord(plan->ev0, st));
    e = render_range(plan, position, frames, out, ld_out, st);
    if (e != SIGB_OK) return e;
    CUDA_TRY(cudaEventRecord(plan->ev1, st));
    plan->ev_valid = true;
    return SIGB_OK;
}

extern "C" int sigb_render_host(sigb_plan* plan, int64_t position, int32_t frames, float* out_host, int64_t ld_out, void* after_stream) {
    if (!plan || (!out_host && frames != 0) || frames < 0 || position < 0 || ld_out < plan->channels)
        return fail(SIGB_EINVAL, "sigb_render_host: bad arguments");
    if (frames == 0) return SIGB_OK;
    int e = upload(plan);
    if (e == SIGB_OK) e = check_design_error(plan);
    if (e != SIGB_OK) return e;
    e = ensure_host_streams(plan);
    if (e != SIGB_OK) return e;
    // work the caller queued on `after_stream` (the upload of a bound Buffer, ...) is ordered before the render
    CUDA_TRY(cudaEventRecord(plan->ev_caller, (cudaStream_t)after_stream));
    CUDA_TRY(cudaStreamWaitEvent(plan->s_render, plan->ev_caller, 0));
    const int C = plan->channels;
    int64_t rows = std::max<int64_t>(1, plan->opt_host_slab_bytes / (4ll * C));
    const int64_t step = 7 * SIGB_SCAN_L * 4;
    if (rows > step) rows = rows / step * step;
    rows = std::min<int64_t>(rows, frames);
    if (plan->stage_floats < rows * C) {
        for (int i = 0; i < 2; ++i) {
            if (plan->stage[i]) cudaFree(plan->stage[i]);
            plan->stage[i] = nullptr;
            CUDA_TRY(cudaMalloc(&plan->stage[i], (size_t)rows * C * sizeof(float)));
        }
        plan->stage_floats = rows * C;
    }
    // ONE request: seek handling and the block-rate parameters happen once, at the request's first frame; the host
    // slabs below only cut the rows
    e = begin_request(plan, position, frames, plan->s_render);
    if (e != SIGB_OK) return e;
    int slot = 0;
    int64_t n_slabs = 0;
    for (int64_t r = 0; r < frames; r += rows, slot ^= 1, ++n_slabs) {
        const int64_t nr = std::min(rows, frames - r);
        if (n_slabs >= 2) CUDA_TRY(cudaStreamWaitEvent(plan->s_render, plan->ev_copied[slot], 0));
        e = run_rows(plan, position + r, nr, plan->stage[slot], C, plan->s_render);
        if (e != SIGB_OK) return e;
        CUDA_TRY(cudaEventRecord(plan->ev_rendered[slot], plan->s_render));
        CUDA_TRY(cudaStreamWaitEvent(plan->s_copy, plan->ev_rendered[slot], 0));
        if (ld_out == C) {
            CUDA_TRY(cudaMemcpyAsync(out_host + r * ld_out, plan->stage[slot], (size_t)nr * C * sizeof(float),
                                     cudaMemcpyDeviceToHost, plan->s_copy));
        } else {
            CUDA_TRY(cudaMemcpy2DAsync(out_host + r * ld_out, ld_out * sizeof(float), plan->stage[slot],
                                       (size_t)C * sizeof(float), (size_t)C * sizeof(float), nr,
                                       cudaMemcpyDeviceToHost, plan->s_copy));
        }
        CUDA_TRY(cudaEventRecord(plan->ev_copied[slot], plan->s_copy));
    }
    end_request(plan, position, frames);
    if (n_slabs > 1) plan->last_slab = 0;          // intermediate buffers hold the last host slab only
    CUDA_TRY(cudaStreamSynchronize(plan->s_copy));
    CUDA_TRY(cudaStreamSynchronize(plan->s_render));
    return check_design_error(plan);
}

// The audio callback's block (SinkDevice._callback, chain/dev.py:167-179): small, latency-bound.  The plan's kernels
// for a block of `frames` rows are captured ONCE into a CUDA graph; every later block of that length is one
// cudaGraphLaunch -- the kernels read the block position from a page-locked header and the root launch stores
// straight into page-locked staging (no copy node) -- one stream synchronisation and one host memcpy into `out_host`.
// Seeks, plans the graph cannot express (see rt_eligible) and blocks above "rt_max_bytes" take sigb_render_host.
extern "C" int sigb_render_block(sigb_plan* plan, int64_t position, int32_t frames, float* out_host, int64_t ld_out) {
    if (!plan || (!out_host && frames != 0) || frames < 0 || position < 0 || ld_out < plan->channels)
        return fail(SIGB_EINVAL, "sigb_render_block: bad arguments");
    if (frames == 0) return SIGB_OK;
    int e = upload(plan);
    if (e == SIGB_OK) e = check_design_error(plan);
    if (e != SIGB_OK) return e;
    const int C = plan->channels;
    const int64_t bytes = (int64_t)frames * C * 4;
    const bool seek = !plan->have_pos || plan->next_pos != position || plan->opt_restart;
    const bool mine = rt_eligible(plan) && bytes <= plan->opt_rt_max_bytes && slab_rows(plan, frames) >= frames;
    if (seek || !mine) {
        // a seek of a stream this path serves runs the same sequential kernels as its graphs, so that a block does not
        // depend (in the last bit) on whether it followed a seek
        const int64_t saved = plan->opt_force_seq;
        if (mine) plan->opt_force_seq = 1;
        const int r = sigb_render_host(plan, position, frames, out_host, ld_out, nullptr);
        plan->opt_force_seq = saved;
        return r;
    }
    e = ensure_host_streams(plan);
    if (e != SIGB_OK) return e;
    if (!plan->rt_hdr) CUDA_TRY(cudaHostAlloc(&plan->rt_hdr, 64, cudaHostAllocMapped));
    if (plan->rt_stage_floats < (int64_t)frames * C) {
        rt_drop_graphs(plan);                              // captured launches point into the old staging
        if (plan->rt_stage) cudaFreeHost(plan->rt_stage);
        plan->rt_stage = nullptr;
        const int64_t want = std::max<int64_t>((int64_t)frames * C, 4096);
        CUDA_TRY(cudaHostAlloc(&plan->rt_stage, want * sizeof(float), cudaHostAllocMapped));
        plan->rt_stage_floats = want;
    }
    cudaStream_t st = plan->s_render;
    plan->rt_hdr[0] = position;
    e = ensure_bufs(plan, frames);
    if (e != SIGB_OK) return e;
    auto record = [&]() -> int {                           // the block's launches, position taken from the header
        plan->rt_pos_ptr = plan->rt_hdr;
        int r = run_params(plan, position, frames, st);
        if (r == SIGB_OK) r = run_slab(plan, position, frames, plan->rt_stage, C, st);
        plan->rt_pos_ptr = nullptr;
        return r;
    };
    if (plan->opt_rt_graph) {
        sigb_plan::RtGraph* g = nullptr;
        const uint64_t sig = rt_state_sig(plan);
        for (sigb_plan::RtGraph& c : plan->rt_graphs)
            if (c.frames == frames) g = &c;
        if (g && g->state_sig != sig) {                    // a render on another path moved the live state copy
            rt_drop_graphs(plan);
            g = nullptr;
        }
        if (!g) {
            sigb_plan::RtGraph ng{frames, nullptr, nullptr, 0, sig};
            const int64_t launches0 = plan->launch_count;
            CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            e = record();
            cudaError_t ce = cudaStreamEndCapture(st, &ng.graph);
            plan->launch_count = launches0;                // recorded, not launched
            if (e != SIGB_OK) { if (ng.graph) cudaGraphDestroy(ng.graph); return e; }
            if (ce != cudaSuccess) return fail(SIGB_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
            CUDA_TRY(cudaGraphInstantiate(&ng.exec, ng.graph, 0));
            size_t nodes = 0;
            cudaGraphGetNodes(ng.graph, nullptr, &nodes);
            ng.nodes = (int)nodes;
            if (plan->rt_graphs.size() >= 8) rt_drop_graphs(plan);
            plan->rt_graphs.push_back(ng);
            g = &plan->rt_graphs.back();
        }
        CUDA_TRY(cudaGraphLaunch(g->exec, st));
        plan->launch_count += g->nodes;
        plan->graph_launches++;
    } else {
        e = record();
        if (e != SIGB_OK) return e;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    if (ld_out == C) {
        std::memcpy(out_host, plan->rt_stage, (size_t)bytes);
    } else {
        for (int r = 0; r < frames; ++r) std::memcpy(out_host + (int64_t)r * ld_out, plan->rt_stage + (int64_t)r * C, (size_t)C * 4);
    }
    plan->last_slab = frames;
    end_request(plan, position, frames);
    return check_design_error(plan);
}

// Blocks of the most recent request as seen by tap `tap` (a Wave / Spec / FileWriter of the graph; its value is the
// input it passes through, chain/__init__.py:409-417): copied from the buffer the render left in HBM -- no second
// render.  SIGB_ESTATE when the request was cut into several slabs (the buffer then holds the last one only).
extern "C" int sigb_plan_tap_count(const sigb_plan* plan) { return plan ? (int)plan->tap_nodes.size() : 0; }

extern "C" int sigb_plan_read_tap(sigb_plan* plan, int32_t tap, float* out_host, int64_t ld_out, int32_t* channels) {
    if (!plan || tap < 0 || tap >= (int)plan->tap_nodes.size()) return fail(SIGB_EINVAL, "sigb_plan_read_tap: no such tap");
    const Val& v = plan->vals[plan->tap_nodes[tap]];
    if (channels) *channels = v.channels;
    if (!out_host) return SIGB_OK;
    if (ld_out < v.channels) return fail(SIGB_EINVAL, "sigb_plan_read_tap: ld_out < channels");
    const int64_t frames = plan->last_frames;
    if (frames <= 0 || plan->last_slab < frames) return fail(SIGB_ESTATE, "sigb_plan_read_tap: the last request was rendered in several slabs");
    cudaStream_t st = plan->s_render ? plan->s_render : nullptr;
    if (v.kind == VK_CONST) {
        for (int64_t r = 0; r < frames; ++r)
            for (int c = 0; c < v.channels; ++c) out_host[r * ld_out + c] = (float)v.cv[c];
        return SIGB_OK;
    }
    if (v.kind != VK_BUF || v.buf < 0 || !plan->bufs[v.buf].ptr) return fail(SIGB_ESTATE, "sigb_plan_read_tap: tap value not materialised");
    CUDA_TRY(cudaMemcpy2DAsync(out_host, ld_out * sizeof(float), plan->bufs[v.buf].ptr, (size_t)v.channels * sizeof(float),
                               (size_t)v.channels * sizeof(float), frames, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return SIGB_OK;
}

extern "C" int sigb_state_reset(sigb_plan* plan) {
    if (!plan) return fail(SIGB_EINVAL, "null plan");
    plan->have_pos = false;
    return SIGB_OK;
}

extern "C" int sigb_plan_destroy(sigb_plan* plan) {
    if (!plan) return SIGB_OK;
    if (plan->uploaded) cudaDeviceSynchronize();
    for (BufInfo& b : plan->bufs)
        if (b.ptr) cudaFree(b.ptr);
    if (plan->d_arena) cudaFree(plan->d_arena);
    if (plan->d_pprog) cudaFree(plan->d_pprog);
    if (plan->d_prow_d) cudaFree(plan->d_prow_d);
    if (plan->d_prow_f) cudaFree(plan->d_prow_f);
    if (plan->d_state) cudaFree(plan->d_state);
    if (plan->scratch) cudaFree(plan->scratch);
    for (int i = 0; i < 2; ++i) {
        if (plan->stage[i]) cudaFree(plan->stage[i]);
        if (plan->ev_rendered[i]) cudaEventDestroy(plan->ev_rendered[i]);
        if (plan->ev_copied[i]) cudaEventDestroy(plan->ev_copied[i]);
    }
    rt_drop_graphs(plan);
    if (plan->rt_hdr) cudaFreeHost(plan->rt_hdr);
    if (plan->rt_stage) cudaFreeHost(plan->rt_stage);
    if (plan->d_warm) cudaFree(plan->d_warm);
    if (plan->h_warm) cudaFreeHost(plan->h_warm);
    if (plan->h_err) cudaFreeHost(plan->h_err);
    if (plan->ev_caller) cudaEventDestroy(plan->ev_caller);
    if (plan->ev0) cudaEventDestroy(plan->ev0);
    if (plan->ev1) cudaEventDestroy(plan->ev1);
    if (plan->s_render) cudaStreamDestroy(plan->s_render);
    if (plan->s_copy) cudaStreamDestroy(plan->s_copy);
    delete plan;
    return SIGB_OK;
}

extern "C" int64_t sigb_plan_describe(const sigb_plan* plan, char* buf, int64_t cap) {
    if (!plan) return 0;
    static const char* src_names[] = {"osc", "block", "const"};
    static const char* wave_names[] = {"sine", "square", "sawtooth", "triangle"};
    static const char* ew_names[] = {"copy", "gain", "mix", "ringmod", "amp"};
    std::string s = "{\"channels\": " + std::to_string(plan->channels) + ", \"rate\": " + std::to_string(plan->rate) +
                    ", \"context\": " + std::to_string(plan->context) + ", \"buffers\": " + std::to_string(plan->bufs.size()) +
                    ", \"state_doubles\": " + std::to_string(plan->n_state) + ", \"param_bytes\": " + std::to_string(plan->arena.size()) +
                    ", \"modulated_parameters\": " + std::to_string(plan->pprog.size()) +
                    ", \"launches\": [";
    bool first = true;
    for (const Launch& l : plan->launches) {
        if (!first) s += ", ";
        first = false;
        if (l.kind == LK_CHAIN) {
            const ChainSpec& c = plan->chains[l.idx];
            s += "{\"kind\": \"chain\", \"node\": " + std::to_string(c.dst_node) + ", \"channels\": " + std::to_string(c.C) +
                 ", \"source\": \"" + src_names[c.src_kind] + "\"";
            if (c.src_kind == SRC_OSC) s += std::string(", \"wave\": \"") + wave_names[c.wave & 3] + "\"";
            if (c.epi_op)
                s += std::string(", \"epilogue\": \"") + (c.epi_op == EW_MIX ? "mix" : "ringmod") + "\", \"other\": \"" +
                     (c.epi_wave >= 0 ? "osc" : "block") + "\"";
            s += ", \"sections\": " + std::to_string(c.nsec_real) + ", \"sections_padded\": " + std::to_string(c.nsec) +
                 ", \"warm_rows\": " + std::to_string(c.warm_rows) +
                 ", \"modulated_cutoffs\": " + std::to_string(c.mods.size()) +
                 ", \"gain\": " + (c.gain.off >= 0 ? "true" : "false") + "}";
        } else if (l.kind == LK_EWISE) {
            const EwiseSpec& e = plan->ewises[l.idx];
            s += "{\"kind\": \"ewise\", \"node\": " + std::to_string(e.dst_node) + ", \"op\": \"" + ew_names[e.op] +
                 "\", \"channels\": " + std::to_string(e.C) + ", \"column\": " + std::to_string(e.dst_coff) + "}";
        } else if (l.kind == LK_BANK) {
            const BankSpec& b = plan->banks[l.idx];
            s += "{\"kind\": \"bank\", \"node\": " + std::to_string(b.dst_node) + ", \"partials\": " + std::to_string(b.ch.C) +
                 ", \"groups\": " + std::to_string(b.groups) + ", \"gain\": " + (b.ch.gain.off >= 0 ? "true" : "false") + "}";
        } else if (l.kind == LK_VOICES) {
            const VoicesSpec& v = plan->voices[l.idx];
            long long total = 0;
            for (const VoiceSegSpec& sg : v.segs) total += sg.ch.C;
            s += "{\"kind\": \"voices\", \"node\": " + std::to_string(v.dst_node) + ", \"segments\": " + std::to_string(v.segs.size()) +
                 ", \"channels\": " + std::to_string(total) + ", \"channels_per_thread\": " + std::to_string(v.M) +
                 ", \"partials\": " + std::to_string(v.nparts) + "}";
        } else {
            const ReduceSpec& r = plan->reduces[l.idx];
            s += "{\"kind\": \"reduce\", \"node\": " + std::to_string(r.dst_node) + ", \"op\": \"" +
                 (r.pan ? "pansum" : "groupsum") + "\", \"channels\": " + std::to_string(r.C) +
                 ", \"groups\": " + std::to_string(r.groups) + "}";
        }
    }
    s += "]}";
    if (buf && cap > 0) {
        const int64_t n = std::min<int64_t>(cap - 1, (int64_t)s.size());
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int64_t)s.size() + 1;
}

extern "C" int sigb_plan_set_option(sigb_plan* plan, const char* key, int64_t value) {
    if (!plan || !key) return fail(SIGB_EINVAL, "null");
    const std::string k(key);
    if (k == "scan_variant") plan->opt_scan_variant = value;
    else if (k == "force_seq") plan->opt_force_seq = value;
    else if (k == "scan_max_tiles") plan->opt_scan_max_tiles = value;
    else if (k == "slab_frames") plan->opt_slab_frames = value;
    else if (k == "host_slab_bytes") plan->opt_host_slab_bytes = value;
    else if (k == "buffer_budget") plan->opt_buffer_budget = value;
    else if (k == "cascade_pipe") plan->opt_cascade_pipe = value;
    else if (k == "pipe_segments") plan->opt_pipe_segments = value;
    else if (k == "pipe_spw") plan->opt_pipe_spw = value;
    else if (k == "cascade_reg") plan->opt_cascade_reg = value;
    else if (k == "reg_variant") plan->opt_reg_variant = value;
    else if (k == "osc_reg") { plan->opt_osc_reg = value; plan->osc_reg_user = true; }
    else if (k == "voices_segments") plan->opt_voices_segments = value;
    else if (k == "voices_pieces") plan->opt_voices_pieces = value;
    else if (k == "rt_graph") plan->opt_rt_graph = value;
    else if (k == "rt_max_bytes") plan->opt_rt_max_bytes = value;
    else if (k == "blockwise_reference") { plan->opt_restart = value; plan->have_pos = false; }
    else if (k == "scan_tma") sigb_set_scan_tma((int)value);   // process-wide switch (A/B testing)
    else if (k == "scan_split") sigb_set_scan_split((int)value);
    else if (k == "scan_rot") sigb_set_scan_rot((int)value);       // process-wide switch (A/B testing)
    else if (k == "reg_pieces") sigb_set_reg_pieces((int)value);   // process-wide switch (A/B testing)
    else if (k == "delta_probe") sigb_set_delta_probe((int)value); // process-wide switch (A/B testing)
    else if (k == "osc_delta") plan->opt_osc_delta = value;
    else if (k == "osc_fill") plan->opt_osc_fill = value;
    else if (k == "osc_pieces_pct") sigb_set_osc_pieces_pct((int)value);   // process-wide switch (A/B testing)
    else if (k == "bank_unroll") sigb_set_bank_unroll((int)value); // process-wide switch (A/B testing)
    else return fail(SIGB_EINVAL, "unknown option " + k);
    rt_drop_graphs(plan);            // captured launches embody the old choice
    return SIGB_OK;
}

extern "C" int64_t sigb_plan_graph_launches(const sigb_plan* plan) { return plan ? plan->graph_launches : 0; }

extern "C" int sigb_set_default_option(const char* key, int64_t value) {
    if (!key) return fail(SIGB_EINVAL, "null");
    const std::string k(key);
    if (k == "fuse_reduce") g_default_fuse_reduce = value;
    else if (k == "voices_m") g_default_voices_m = value;
    else if (k == "fuse_pointwise") g_default_fuse_pointwise = value;
    else return fail(SIGB_EINVAL, "unknown default option " + k);
    return SIGB_OK;
}

extern "C" int64_t sigb_plan_launch_count(const sigb_plan* plan) { return plan ? plan->launch_count : 0; }

extern "C" int sigb_plan_last_kernel_ms(sigb_plan* plan, float* ms) {
    if (!plan || !ms || !plan->ev_valid) return fail(SIGB_ESTATE, "no render recorded");
    CUDA_TRY(cudaEventElapsedTime(ms, plan->ev0, plan->ev1));
    return SIGB_OK;
}

extern "C" int sigb_design_butter(int32_t subtype, int32_t order, double wn, double* coef, int32_t cap_sections) {
    if (!coef || order < 1) return fail(SIGB_EINVAL, "sigb_design_butter: bad arguments");
    if (!(wn > 0.0 && wn < 1.0)) return fail(SIGB_ECRIT, "Digital filter critical frequencies must be 0 < Wn < 1");
    std::vector<SvfSection> secs = sigb_butter_sections(subtype == SIGB_FILT_HIGHPASS, order, wn);
    if ((int)secs.size() > cap_sections) return fail(SIGB_EINVAL, "sigb_design_butter: buffer too small");
    for (size_t i = 0; i < secs.size(); ++i) {
        coef[i * 4 + 0] = secs[i].g;
        coef[i * 4 + 1] = secs[i].r2;
        coef[i * 4 + 2] = (double)secs[i].kind;
        coef[i * 4 + 3] = 0.0;
    }
    return (int)secs.size();
}

extern "C" int sigb_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return fail(SIGB_EINVAL, "sigb_host_alloc: bad arguments");
    CUDA_TRY(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
    return SIGB_OK;
}

extern "C" int sigb_host_free(void* ptr) {
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return SIGB_OK;
}

// test hooks -----------------------------------------------------------------------------------
extern "C" int sigb_probe_sin(const double* r_dev, int32_t n, float* out_dev, int32_t variant, void* stream) {
    int e = sigb_launch_probe_sin(r_dev, n, out_dev, variant, stream);
    if (e) return fail(SIGB_ECUDA, cudaGetErrorString((cudaError_t)e));
    return SIGB_OK;
}

extern "C" uint64_t sigb_probe_ratio_q64(double hertz, int32_t rate) { return ratio_q64(hertz, rate); }
extern "C" uint64_t sigb_probe_frac_q64(double x) { return frac_q64(x); }
