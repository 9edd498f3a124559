/*
 * sigb200.h -- C ABI of libsigb200.so, the B200 (sm_100a) block-render engine for the
 * `signals` node graph.
 *
 * The reference (noah-aviel-dove/signals) has no FFI: its render is the Python recursion
 *   Receiver.BoundPort.request(loc) -> Emitter.respond(request) -> _eval(request)
 * (src/signals/chain/__init__.py:296-300, 253-258) rooted at
 *   block = self.input.request(loc)            (src/signals/chain/dev.py:173).
 * This header is the boundary a maintainer binds instead of that recursion: the Python
 * host topologically sorts the graph into `sigb_node` records (one per node, inputs by
 * index), and every sigb_render() call is one block request
 * BlockLoc{position, rate, shape=(frames, channels)} (src/signals/chain/__init__.py:107-125).
 *
 * Plain C types only; no torch types.  All device pointers are caller-owned CUDA global
 * memory on the plan's device; `stream` is a cudaStream_t passed as void* (NULL = legacy
 * default stream).  One render in flight per plan (the reference has a single audio thread,
 * src/signals/chain/dev.py:167-179).  Functions return 0 (SIGB_OK) or a negative SIGB_E*.
 * There is no CPU fallback: without a CUDA device every render entry point fails with
 * SIGB_ECUDA.
 */
#ifndef SIGB200_H
#define SIGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIGB_ABI_VERSION 2

/* status codes; the Python shim maps them onto the reference's exception types
 * (src/signals/chain/__init__.py:21, 87-104). */
enum {
    SIGB_OK = 0,
    SIGB_EINVAL = -1,       /* malformed record / NULL pointer                    -> ValueError          */
    SIGB_ESHAPE = -2,       /* block not broadcast-compatible with the request    -> BadShape   (:292)   */
    SIGB_EINDEX = -3,       /* filter cutoff/input narrower than the request      -> IndexError (fx.py:99,105) */
    SIGB_ECRIT = -4,        /* critical frequency not in (0, Nyquist)             -> ValueError (scipy butter via fx.py:102) */
    SIGB_EUNSUPPORTED = -5, /* node/graph shape this engine does not lower        -> ChainLayerError     */
    SIGB_ECUDA = -6,        /* CUDA runtime failure / no device                   -> RuntimeError        */
    SIGB_ENOMEM = -7,
    SIGB_ESTATE = -8        /* plan used while another render is in flight, etc.  */
};

/* node kinds: one per reference node class on the hot path */
enum {
    SIGB_NODE_ZERO = 0,   /* Emitter.empty_result(): zeros((1,1)) -- unconnected port or disabled node
                             (src/signals/chain/__init__.py:249-254, 296-298) */
    SIGB_NODE_FIXED = 1,  /* chain/fixed.py:21-39; rows==1: per-channel constants                 */
    SIGB_NODE_OSC = 2,    /* chain/osc.py:18-62;   in[0]=hertz  in[1]=phase (block rate)          */
    SIGB_NODE_GAIN = 3,   /* chain/fx.py:49-52;    in[0]=left   in[1]=right (block rate)          */
    SIGB_NODE_MIX = 4,    /* chain/fx.py:35-40;    in[0]=left   in[1]=right in[2]=mix (block rate)*/
    SIGB_NODE_RINGMOD = 5,/* chain/fx.py:43-46;    in[0]=left   in[1]=right                       */
    SIGB_NODE_AMP = 6,    /* chain/fx.py:55-60;    in[0]=left   in[1]=right=exp (block rate)      */
    SIGB_NODE_FILTER = 7, /* chain/fx.py:63-151;   in[0]=input  in[1]=cutoff (block rate)         */
    SIGB_NODE_MERGE = 8,  /* chain/shape.py:60-74; in[0]=left   in[1]=right                       */
    SIGB_NODE_GROUPSUM = 9,  /* extension (the reference's Flatten is broken, shape.py:32-35):
                                in[0]=input (C ch) -> `order` groups of C/order adjacent channels */
    SIGB_NODE_PANSUM = 10,   /* extension: in[0]=input (C ch), in[1]=pan (block rate) -> 2 ch:
                                L = sum((1-pan) y), R = sum(pan y)                                */
    SIGB_NODE_BUFFER = 11,   /* extension: HBM-resident sample source, (rows, channels) floats,
                                row index = absolute frame position; zeros past the end           */
    SIGB_NODE_TAP = 12       /* pass-through side-effect node (chain/vis.py:61-64 Wave / Spec, chain/files.py:89-102
                                FileWriter): in[0]=input; value = its input, kept materialised so that the host can
                                read the block after the render (sigb_plan_read_tap).  Never the root.            */
};

/* OSC subtype */
enum { SIGB_WAVE_SINE = 0, SIGB_WAVE_SQUARE = 1, SIGB_WAVE_SAWTOOTH = 2, SIGB_WAVE_TRIANGLE = 3 };
/* FILTER subtype (CritFilter.Type, chain/fx.py:68-72); band types are unreachable in the
 * reference (fx.py:99 raises TypeError) and are rejected with SIGB_EUNSUPPORTED. */
enum { SIGB_FILT_LOWPASS = 0, SIGB_FILT_HIGHPASS = 1 };

typedef struct sigb_node {
    int32_t kind;       /* SIGB_NODE_*                                                         */
    int32_t subtype;    /* OSC: SIGB_WAVE_*; FILTER: SIGB_FILT_*                               */
    int32_t channels;   /* natural output channel count of this node (1 = broadcasts)          */
    int32_t in[3];      /* indices of the nodes on this node's ports, -1 = unconnected         */
    int32_t order;      /* FILTER: Butterworth order N (CritFilter.order, fx.py:66); GROUPSUM: groups */
    int32_t context;    /* FILTER: context_frames() (fx.py:82-83): zero-state warm-up after a seek */
    int32_t rows;       /* FIXED / BUFFER: rows of the value table                             */
    int32_t reserved;
    int64_t data_off;   /* FIXED: offset (in doubles) into `data`, rows*channels row-major.
                           BUFFER: offset (in floats) is given separately via sigb_plan_bind_buffer */
} sigb_node;

typedef struct sigb_plan sigb_plan;

/* Build a plan for the graph `nodes[0..n_nodes)` (topologically sorted: in[] < own index),
 * rendering node `root` as a (frames, channels) float32 block at sample rate `rate`.
 * Pure host work (validation, Butterworth design, fusion); CUDA is first touched by render. */
int sigb_plan_create(const sigb_node* nodes, int32_t n_nodes, int32_t root,
                     const double* data, int64_t n_data,
                     int32_t channels, int32_t rate, sigb_plan** out_plan);

/* Attach device memory to a SIGB_NODE_BUFFER node (caller keeps ownership). */
int sigb_plan_bind_buffer(sigb_plan* plan, int32_t node, const float* dev_ptr, int64_t rows);
/* Streaming variant: the bound memory holds frames [first_row, first_row + rows) of the source (a
 * FileReader-style window, chain/files.py:70-87); frames outside the window read as zero.  Re-bind
 * before each sigb_render call of a stream; filter state is carried as usual. */
int sigb_plan_bind_buffer_window(sigb_plan* plan, int32_t node, const float* dev_ptr, int64_t first_row, int64_t rows);

/* Render frames [position, position+frames) into `out` (device, row-major, leading dimension
 * `ld_out` floats >= channels).  Filter state is carried when `position` continues the previous
 * call; otherwise (a seek) state is zeroed and the whole graph is warmed up once from
 * `position` - (sum of the `context` frames along the deepest filter chain).  For ONE filter on a path that
 * is what CritFilter._filter does for every block (fx.py:93-105; golden `lowpass_blockwise`, 8e-8).  For
 * chained filters the reference nests its restarts (the upstream filter's context block is itself restarted
 * 100 frames earlier); both are approximations of the stream from position 0 and differ from each other by
 * the upstream transient (golden `cascade2_seek`: 8.3e-5 at 900..3000 Hz cutoffs, more at lower ones). */
int sigb_render(sigb_plan* plan, int64_t position, int32_t frames,
                float* out, int64_t ld_out, void* stream);

/* Same block, delivered to HOST memory (page-locked for full speed): renders in time slabs on the
 * plan's own stream and overlaps each slab's device->host copy with the next slab's kernels.
 * This is the call that replaces `block = self.input.request(loc)` + the copy into `outdata`
 * (src/signals/chain/dev.py:173,178).  Blocks until `out_host` is complete. */
int sigb_render_host(sigb_plan* plan, int64_t position, int32_t frames,
                     float* out_host, int64_t ld_out, void* after_stream);
/* `after_stream` (cudaStream_t as void*, NULL = legacy default stream): work the caller has queued there -- the
 * upload of a bound Buffer, a previous sigb_render -- is ordered before the render, which runs on the plan's own
 * streams.  Block-rate parameters are sampled ONCE, at `position`, however the request is cut into slabs. */

/* The audio callback's block (SinkDevice._callback, src/signals/chain/dev.py:167-179): the latency path.  The
 * launches of a block of `frames` rows are captured once into a CUDA graph; every later contiguous block of that
 * length is ONE cudaGraphLaunch (the kernels read the position from a page-locked block header, the root launch
 * writes page-locked staging directly), one stream synchronisation and one memcpy into `out_host` (any host
 * memory).  Seeks, plans with Buffer sources or fused bank / voice reductions, and blocks above "rt_max_bytes"
 * are served by sigb_render_host.  Same results as sigb_render_host for the same calls. */
int sigb_render_block(sigb_plan* plan, int64_t position, int32_t frames,
                      float* out_host, int64_t ld_out);
/* CUDA graphs launched by sigb_render_block since the plan was created. */
int64_t sigb_plan_graph_launches(const sigb_plan* plan);

/* Taps: the blocks the SIGB_NODE_TAP nodes saw during the most recent request, copied from the buffers the render
 * left in HBM (no second render).  Taps are numbered in plan order (topological order of their records).
 * out_host == NULL only reports `channels`.  SIGB_ESTATE when the request was rendered in several slabs. */
int sigb_plan_tap_count(const sigb_plan* plan);
int sigb_plan_read_tap(sigb_plan* plan, int32_t tap, float* out_host, int64_t ld_out, int32_t* channels);

int sigb_state_reset(sigb_plan* plan);          /* zero all filter state; next render is a seek  */
int sigb_plan_destroy(sigb_plan* plan);

/* Plan introspection (host only): JSON description of the fused launches; returns bytes needed. */
int64_t sigb_plan_describe(const sigb_plan* plan, char* buf, int64_t cap);
/* Tuning knobs (A/B testing; every setting computes the same transfer function); unknown key -> SIGB_EINVAL:
 *   "force_seq" (1: k_chain_seq only), "scan_variant" (geometry of the time-parallel scan kernel),
 *   "scan_max_tiles", "slab_frames", "host_slab_bytes", "buffer_budget",
 *   "cascade_reg" (-1: register-resident cascade kernel whenever it can take the chain, 0: never),
 *   "reg_variant" (0: k_cascade_delta wherever it applies, else k_cascade_reg in 8-row blocks; 1: k_cascade_reg in 4-row blocks;
 *   4: k_cascade_reg in 8-row blocks), "osc_reg" (oscillator-fed chains of >= n sections run
 *   register-resident, 0: never; default 2, where 2-section chains qualify only when unmodulated and filling the machine),
 *   "osc_delta" (0: those chains keep state-variable sections), "delta_probe" / "osc_pieces_pct" / "reg_pieces" (process-wide A/B
 *   switches behind profiles/r02_c4_delta.txt and r02_osc_sections.txt: geometry of k_cascade_delta for 8 low-pass sections, time
 *   pieces of the register kernels as a percentage of / per resident warp slot), "cascade_pipe" (-1 auto, 0 never, n: from n sections),
 *   "pipe_spw", "pipe_segments" (upper bound on the time pieces per tile of the cascade kernels; 1: never cut),
 *   "voices_segments" (k_voices: 0 equal pieces per resident CTA, 1 one piece per voice group, n pieces per group),
 *   "voices_pieces" (automatic mode: pieces per resident CTA slot, default 16), "voices_m",
 *   "rt_graph" (0: sigb_render_block launches its kernels directly instead of through a captured graph),
 *   "rt_max_bytes" (largest block sigb_render_block serves itself),
 *   "blockwise_reference" (1: EVERY request restarts the filters from zero state and `context` warm-up frames, as the
 *   reference itself does block by block, fx.py:82-83, 93-105 -- for A/B against the reference's blockwise output;
 *   default 0 carries the true state across contiguous requests). */
int sigb_plan_set_option(sigb_plan* plan, const char* key, int64_t value);
/* Defaults for plans created AFTERWARDS (decisions taken while the plan is built):
 * "fuse_reduce" (1: GroupSum / PanSum over oscillator chains run as one fused render+reduce kernel,
 * 0: always on materialised blocks), "voices_m" (0 auto, 1 or 4 voices per thread in the fused kernel),
 * "fuse_pointwise" (1: Mix / RingMod ride on a stateless oscillator chain as its epilogue, 0: always k_ewise). */
int sigb_set_default_option(const char* key, int64_t value);
/* Kernels launched by this plan since creation (bench.py's gpu_launches claim). */
int64_t sigb_plan_launch_count(const sigb_plan* plan);
/* Device time (ms) of the kernels of the most recent sigb_render call, measured with CUDA events
 * recorded on the launching stream; valid after that stream has been synchronised. */
int sigb_plan_last_kernel_ms(sigb_plan* plan, float* ms);

/* Host-side filter design used by the plan (exposed for tests): 2nd-order sections of the order-N
 * Butterworth low/high-pass at normalised frequency wn = cutoff/(rate/2), as the zero-delay-feedback
 * state-variable sections the kernels run.  coef receives n_sections*4 doubles {g, r2, kind, 0}. */
int sigb_design_butter(int32_t subtype, int32_t order, double wn, double* coef, int32_t cap_sections);

/* Page-locked host memory for sigb_render_host outputs. */
int sigb_host_alloc(void** ptr, int64_t bytes);
int sigb_host_free(void* ptr);

const char* sigb_strerror(int status);
const char* sigb_last_error(void);   /* thread-local detail for the last failure */
int sigb_abi_version(void);
int sigb_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SIGB200_H */
