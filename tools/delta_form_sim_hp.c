// precision of the 4-op delta HIGH-PASS section vs the SVF high-pass in float32, against a float64 SVF cascade
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#define NS 8
int main(int argc, char** argv) {
    double rate = 48000.0; int secs = argc > 1 ? atoi(argv[1]) : 10; long n = (long)(rate * secs);
    double cuts[][NS] = {
        {20,20,20,20,20,20,20,20}, {200,200,200,200,200,200,200,200}, {20,8000,50,3000,100,12000,20,500},
        {8000,8000,8000,8000,8000,8000,8000,8000}, {20000,20000,20000,20000,20000,20000,20000,20000}, {23000,23000,23000,23000,20,20,20,20},
        {1000,2000,300,5000,700,250,7000,400}, {5,5,5,5,5,5,5,5}};
    int ncase = sizeof(cuts) / sizeof(cuts[0]);
    for (int ci = 0; ci < ncase; ++ci) {
        double g[NS], c[NS], d[NS]; float gf[NS], ncf[NS], alf[NS], a2f[NS], g2f[NS], df[NS]; float nQ[NS], nbe4[NS], dc[NS];
        for (int s = 0; s < NS; ++s) {
            double wn = cuts[ci][s] / (rate / 2); g[s] = tan(M_PI * wn / 2); double r2 = sqrt(2.0);
            c[s] = r2 + g[s]; d[s] = 1.0 / (1.0 + r2 * g[s] + g[s] * g[s]);
            gf[s] = (float)g[s]; float cf = (float)c[s]; df[s] = (float)d[s];
            ncf[s] = -cf; alf[s] = gf[s] * df[s]; a2f[s] = 2 * alf[s]; g2f[s] = 2 * gf[s];
            double G = gf[s], Cc = cf, Dd = df[s]; double R2 = Cc - G; double Q = 2 * R2 * G * Dd, F = 4 * G * G * Dd;
            nQ[s] = (float)(-Q); nbe4[s] = (float)(-F); dc[s] = df[s];
        }
        double s1[NS] = {0}, s2[NS] = {0}; float t1[NS] = {0}, t2[NS] = {0}; float D[NS] = {0}, Zm[NS] = {0};
        double e1 = 0, e2 = 0, ymax = 0; unsigned long long rs = 88172645463325252ull;
        for (long i = 0; i < n; ++i) {
            rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17;
            float xf = (float)((double)(rs >> 11) / 9007199254740992.0 * 2 - 1);
            xf = 0.5f * xf + 0.5f * (float)sin(2 * M_PI * 7.0 * i / rate);
            double x = xf; float xa = xf; float tprev = xf, cprev = 1.0f;
            for (int s = 0; s < NS; ++s) {
                double e = x - c[s] * s1[s] - s2[s]; double hp = d[s] * e; double bp = s1[s] + g[s] * hp; s1[s] = s1[s] + 2 * g[s] * hp;
                double lp = s2[s] + g[s] * bp; s2[s] = s2[s] + 2 * g[s] * bp; x = hp;
                float xs = xa - t2[s]; float ee = fmaf(ncf[s], t1[s], xs); float bpf = fmaf(alf[s], ee, t1[s]); t1[s] = fmaf(a2f[s], ee, t1[s]);
                t2[s] = fmaf(g2f[s], bpf, t2[s]); xa = ee * df[s];
                float w = fmaf(cprev, tprev, Zm[s]); float t = fmaf(nQ[s], D[s], w); D[s] = D[s] + t; Zm[s] = fmaf(nbe4[s], D[s], Zm[s]);
                tprev = t; cprev = dc[s];
            }
            float xb = cprev * tprev;
            if (fabs(x) > ymax) ymax = fabs(x);
            if (fabs(xa - x) > e1) e1 = fabs(xa - x);
            if (fabs(xb - x) > e2) e2 = fabs(xb - x);
        }
        double ds = 0; for (int s = 0; s < NS; ++s) { ds = fmax(ds, fabs(2 * g[s] * d[s] * D[s] - s1[s])); ds = fmax(ds, fabs((s2[s] + g[s] * s1[s]) / 4 - (-Zm[s] / 4))); }
        printf("case %d cut0 %g: max|y| %.3g  err svf32 %.3g  err delta32 %.3g  state-identity dev %.3g\n", ci, cuts[ci][0], ymax, e1, e2, ds);
    }
    return 0;
}
