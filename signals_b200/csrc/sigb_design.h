// Host-side filter design for libsigb200: Butterworth low/high-pass of order N as a cascade of
// zero-delay-feedback state-variable sections, plus the per-section linear-system tables the
// time-parallel scan kernel needs.  float64 throughout.
//
// Reference: CritFilter._filter / _get_sos, /root/reference/src/signals/chain/fx.py:85-121, which
// calls scipy.signal.butter(N, Wn, btype, output='sos') per channel.  butter() = analog Butterworth
// prototype (poles exp(j*pi*(2k+N+1)/(2N))), frequency pre-warp 2*fs*tan(pi*Wn/2), bilinear
// transform.  A trapezoidal (TPT) state-variable section with g = tan(pi*Wn/2) and damping
// r2 = 2*sin(pi*(2k+1)/(2N)) realises exactly the bilinear transform of the prototype's k-th
// conjugate pole pair, so the cascade has the same transfer function as the scipy sos cascade.
#pragma once
#include <vector>

struct SvfSection {
    int kind;      // SEC_HP | SEC_FIRST_ORDER bits
    double g;      // tan(pi*wn/2)
    double r2;     // 2*sin(pi*(2k+1)/(2N)); unused for first-order sections
};

// Sections of the order-N Butterworth (highpass != 0 -> high-pass).  wn = cutoff / (rate/2), 0<wn<1.
std::vector<SvfSection> sigb_butter_sections(int highpass, int order, double wn);

// float32 kernel coefficients {g, c, d} of one section (first-order: {G, 0, 0}).
void sigb_section_coef(const SvfSection& s, float out[3]);

// One sample of the section in float64: returns the output, updates (s1, s2).
double sigb_section_step(const SvfSection& s, double x, double& s1, double& s2);

// State transition over `len` samples of zero input: m[4] row-major 2x2 with (s1,s2)' = m (s1,s2).
void sigb_section_transition(const SvfSection& s, int len, double m[4]);

// Zero-input output response: tab[k*2+j] = output at sample k (k < len) when state j starts at 1.
void sigb_section_zero_input(const SvfSection& s, int len, float* tab);

// Rows after which the section's zero-input response has decayed below 2^-40 of the initial state,
// from the spectral radius of the one-sample transition m1 (row-major 2x2); 1e9 if not contractive.
double sigb_section_decay_rows(const double m1[4]);

// Spectral radius of a one-sample transition.
double sigb_section_radius(const double m1[4]);

// Rows after which the zero-input response of the whole cascade (every state and the output, from every
// unit initial state) stays below 2^-bits: simulated in float64.  -1 when it has not decayed by max_rows.
int sigb_cascade_decay_rows(const std::vector<SvfSection>& secs, int bits, int max_rows);
