#include "sigb_design.h"

#include <cmath>

#include "sigb_internal.h"

std::vector<SvfSection> sigb_butter_sections(int highpass, int order, double wn) {
    std::vector<SvfSection> secs;
    const double pi = 3.14159265358979323846;
    const double g = std::tan(pi * wn / 2.0);
    const int hp = highpass ? SEC_HP : 0;
    for (int k = 0; k < order / 2; ++k) {
        SvfSection s;
        s.kind = hp;
        s.g = g;
        s.r2 = 2.0 * std::sin(pi * (2.0 * k + 1.0) / (2.0 * order));
        secs.push_back(s);
    }
    if (order & 1) {
        SvfSection s;
        s.kind = hp | SEC_FIRST_ORDER;
        s.g = g;
        s.r2 = 0.0;
        secs.push_back(s);
    }
    return secs;
}

void sigb_section_coef(const SvfSection& s, float out[3]) {
    if (s.kind & SEC_FIRST_ORDER) {
        out[0] = (float)(s.g / (1.0 + s.g));
        out[1] = 0.0f;
        out[2] = 0.0f;
    } else {
        out[0] = (float)s.g;
        out[1] = (float)(s.r2 + s.g);
        out[2] = (float)(1.0 / (1.0 + s.r2 * s.g + s.g * s.g));
    }
}

double sigb_section_step(const SvfSection& s, double x, double& s1, double& s2) {
    if (s.kind & SEC_FIRST_ORDER) {
        const double G = s.g / (1.0 + s.g);
        const double v = (x - s1) * G;
        const double lp = v + s1;
        s1 = lp + v;
        return (s.kind & SEC_HP) ? x - lp : lp;
    }
    const double d = 1.0 / (1.0 + s.r2 * s.g + s.g * s.g);
    const double hp = (x - (s.r2 + s.g) * s1 - s2) * d;
    const double bp = s.g * hp + s1;
    s1 = s.g * hp + bp;
    const double lp = s.g * bp + s2;
    s2 = s.g * bp + lp;
    return (s.kind & SEC_HP) ? hp : lp;
}

void sigb_section_transition(const SvfSection& s, int len, double m[4]) {
    double a1 = 1.0, a2 = 0.0, b1 = 0.0, b2 = 1.0;   // images of the two unit states
    for (int k = 0; k < len; ++k) {
        sigb_section_step(s, 0.0, a1, a2);
        sigb_section_step(s, 0.0, b1, b2);
    }
    m[0] = a1; m[1] = b1;
    m[2] = a2; m[3] = b2;
}

void sigb_section_zero_input(const SvfSection& s, int len, float* tab) {
    double a1 = 1.0, a2 = 0.0, b1 = 0.0, b2 = 1.0;
    for (int k = 0; k < len; ++k) {
        tab[k * 2 + 0] = (float)sigb_section_step(s, 0.0, a1, a2);
        tab[k * 2 + 1] = (float)sigb_section_step(s, 0.0, b1, b2);
    }
}

double sigb_section_decay_rows(const double m1[4]) {
    const double tr = m1[0] + m1[3], det = m1[0] * m1[3] - m1[1] * m1[2];
    const double disc = tr * tr - 4.0 * det;
    double rho;
    if (disc < 0.0) rho = std::sqrt(std::fabs(det));
    else rho = std::max(std::fabs(tr + std::sqrt(disc)), std::fabs(tr - std::sqrt(disc))) / 2.0;
    if (!(rho < 1.0)) return 1e9;
    if (rho < 1e-12) return 2.0;
    // 2^-40 decay, doubled: a near-defective pair decays like k rho^k
    return 2.0 * (40.0 * std::log(2.0) / -std::log(rho)) + 16.0;
}

double sigb_section_radius(const double m1[4]) {
    const double tr = m1[0] + m1[3], det = m1[0] * m1[3] - m1[1] * m1[2];
    const double disc = tr * tr - 4.0 * det;
    if (disc < 0.0) return std::sqrt(std::fabs(det));
    return std::max(std::fabs(tr + std::sqrt(disc)), std::fabs(tr - std::sqrt(disc))) / 2.0;
}

int sigb_cascade_decay_rows(const std::vector<SvfSection>& secs, int bits, int max_rows) {
    const int S = (int)secs.size();
    const double eps = std::ldexp(1.0, -bits);
    int worst = 0;
    std::vector<double> s1(S), s2(S);
    for (int j = 0; j < 2 * S; ++j) {
        std::fill(s1.begin(), s1.end(), 0.0);
        std::fill(s2.begin(), s2.end(), 0.0);
        if (j & 1) {
            if (secs[j / 2].kind & SEC_FIRST_ORDER) continue;     // no second state
            s2[j / 2] = 1.0;
        } else {
            s1[j / 2] = 1.0;
        }
        int last = 0, quiet = 0;
        for (int k = 0; k < max_rows; ++k) {
            double x = 0.0, big = 0.0;
            for (int s = 0; s < S; ++s) {
                x = sigb_section_step(secs[s], x, s1[s], s2[s]);
                big = std::max(big, std::max(std::fabs(x), std::max(std::fabs(s1[s]), std::fabs(s2[s]))));
            }
            if (big >= eps) { last = k + 1; quiet = 0; }
            else if (++quiet > 256) break;                        // below the threshold for good
            if (k + 1 == max_rows) return -1;
        }
        worst = std::max(worst, last);
    }
    return worst;
}
