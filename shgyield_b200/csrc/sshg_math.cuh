// FP64 complex arithmetic and the per-column physics of the SSHG yield
// (wave vectors, Fresnel factors, three-layer multiple reflection, chi2 contraction).
//
// Follows the formulas of the reference shgyield/shg.py (PRB 94, 115314):
//   wvec 137-141, frefs/frefp 144-156, ftrans/ftranp 159-171, mrc2w 174-185, mrc1w 188-198,
//   rad_pp 201-270, rad_ps 273-314, rad_sp 317-352, rad_ss 355-378, prefactor 428-436.
// The evaluation is re-organised for the GPU: everything that does not depend on the azimuth
// phi is done once per (energy, theta, thickness) column ("column setup"), each r(phi) is then
// a 7-term (sS: 4-term) real-basis expansion with complex coefficients.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

// The physics is plain FP64 arithmetic, so it is declared __host__ __device__: the product only
// ever calls it from kernels; tests/hostmath_harness.cu runs the same functions on the CPU to
// check the formulas against the golden vectors without a GPU.
#define SSHG_HD __host__ __device__ __forceinline__

namespace sshg {

struct cplx {
    double re, im;
};

SSHG_HD cplx mk(double r, double i) { return cplx{r, i}; }
SSHG_HD cplx operator+(cplx a, cplx b) { return mk(a.re + b.re, a.im + b.im); }
SSHG_HD cplx operator-(cplx a, cplx b) { return mk(a.re - b.re, a.im - b.im); }
SSHG_HD cplx operator-(cplx a) { return mk(-a.re, -a.im); }
SSHG_HD cplx operator*(cplx a, cplx b) {
    return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
SSHG_HD cplx operator*(double s, cplx a) { return mk(s * a.re, s * a.im); }
SSHG_HD cplx operator+(double s, cplx a) { return mk(s + a.re, a.im); }
SSHG_HD cplx operator-(double s, cplx a) { return mk(s - a.re, -a.im); }
SSHG_HD cplx csq(cplx a) { return mk(a.re * a.re - a.im * a.im, 2.0 * a.re * a.im); }

// 1/b with one real division.
SSHG_HD cplx crecip(cplx b) {
    double n = 1.0 / (b.re * b.re + b.im * b.im);
    return mk(b.re * n, -b.im * n);
}
SSHG_HD cplx operator/(cplx a, cplx b) { return a * crecip(b); }

// Principal square root with the C99 / NumPy branch-cut convention: for negative real parts the
// sign of the imaginary part (including a signed zero) selects the half plane.
SSHG_HD cplx csqrt_(cplx z) {
    double a = z.re, b = z.im;
    if (a == 0.0 && b == 0.0) return mk(0.0, b);
    double t = sqrt(0.5 * (fabs(a) + sqrt(a * a + b * b)));
    if (a >= 0.0) return mk(t, b / (2.0 * t));
    return mk(fabs(b) / (2.0 * t), copysign(t, b));
}

// exp(i z)
SSHG_HD cplx cexp_i(cplx z) {
    double s, c;
    sincos(z.re, &s, &c);
    double m = exp(-z.im);
    return mk(m * c, m * s);
}

// sin(z) for complex z
SSHG_HD cplx csin_(cplx z) {
    double s, c;
    sincos(z.re, &s, &c);
    return mk(s * cosh(z.im), c * sinh(z.im));
}

struct Physics {
    double kd;            // 1e-9 / (h c):  E[eV] * d[nm] * kd = d / lambda_vacuum
    double pref;          // 100 / (hbar * sqrt(2 eps0 c^3)):  sqrt(prefactor) = pref * E / |cos(theta)|
    int mref;
    int sinc_normalized;
    int zzz_phi_quirk;
    int raw;              // 1: leave sqrt(prefactor) / n_l out of the Gamma's (amplitudes of rad_pp .. rad_ss)
    int radians;          // 1: theta / phi of point lists are in radians (the reference's helper functions)
};

// Everything of one (energy, theta, thickness) column that the four polarisations share.
struct Column {
    cplx Gpp, Gps, Gsp, Gss;   // sqrt(prefactor) * Gamma / n_l for each polarisation
    cplx A;                    // (1 - R^M_p) W_l          x/y part of the outgoing P field
    cplx Bz;                   // (1 + R^M_p) sin(theta)   z part of the outgoing P field
    cplx Q;                    // (1 + R^M_p) (1 + r^M_p)^2          (chi_zzz quirk term)
    cplx a2, ab, b2;           // (1-r^M_p)^2 w_l^2, (1+r^M_p)(1-r^M_p) w_l sin(theta), (1+r^M_p)^2 sin^2(theta)
};

struct Fresnel {
    cplx r_p_vl, r_p_lb, r_s_vl, r_s_lb;   // reflection, vacuum|layer and layer|bulk
    cplx t_p, t_s;                         // transmission vacuum -> layer
    cplx w_l;                              // wave vector in the layer
};

// shg.py:137-171 for one frequency.  Shares the denominators between r and t.
SSHG_HD Fresnel fresnel_set(cplx ev, cplx el, cplx eb, double s2) {
    Fresnel f;
    cplx wv = csqrt_(mk(ev.re - s2, ev.im));
    cplx wl = csqrt_(mk(el.re - s2, el.im));
    cplx wb = csqrt_(mk(eb.re - s2, eb.im));
    f.w_l = wl;
    // s polarisation
    cplx isvl = crecip(wv + wl);
    f.r_s_vl = (wv - wl) * isvl;
    f.t_s = (2.0 * wv) * isvl;
    f.r_s_lb = (wl - wb) / (wl + wb);
    // p polarisation
    cplx wvel = wv * el, wlev = wl * ev;
    cplx ipvl = crecip(wvel + wlev);
    f.r_p_vl = (wvel - wlev) * ipvl;
    f.t_p = ((2.0 * wv) * csqrt_(ev * el)) * ipvl;      // sqrt of the PRODUCT (shg.py:170)
    cplx wleb = wl * eb, wbel = wb * el;
    f.r_p_lb = (wleb - wbel) / (wleb + wbel);
    return f;
}

// Column setup: shg.py:174-216, 277-288, 321-332, 359-370 and the prefactor 428-430.
SSHG_HD Column column_setup(const cplx e1[3], const cplx e2[3], double energy,
                                               double sin_t, double cos_t, double thick,
                                               const Physics& ph) {
    const double s2 = sin_t * sin_t;
    Fresnel f1 = fresnel_set(e1[0], e1[1], e1[2], s2);   // fundamental
    Fresnel f2 = fresnel_set(e2[0], e2[1], e2[2], s2);   // second harmonic

    cplx Rp = f2.r_p_lb, Rs = f2.r_s_lb, rp = f1.r_p_lb, rs = f1.r_s_lb;
    if (ph.mref) {
        const double x = energy * thick * ph.kd;         // d / lambda
        const double PI = 3.141592653589793;
        // 2w: delta = 8 pi x W_l;  R^M = R_lb e^{i delta/2} / (1 + R_vl R_lb e^{i delta}) * sinc(delta/2)
        cplx hd = (4.0 * PI * x) * f2.w_l;               // delta / 2
        cplx eh = cexp_i(hd);
        cplx ed = csq(eh);
        cplx arg = hd;
        if (arg.re == 0.0 && arg.im == 0.0) arg = mk(1.0e-20, 0.0);   // np.sinc: 0 -> 1e-20
        if (ph.sinc_normalized) arg = PI * arg;
        cplx snc = csin_(arg) / arg;
        Rp = ((f2.r_p_lb * eh) / (1.0 + f2.r_p_vl * f2.r_p_lb * ed)) * snc;
        Rs = ((f2.r_s_lb * eh) / (1.0 + f2.r_s_vl * f2.r_s_lb * ed)) * snc;
        // 1w: varphi = 4 pi x w_l;  r^M = r_lb e^{i varphi} / (1 + r_vl r_lb e^{i varphi})
        cplx ev = cexp_i((4.0 * PI * x) * f1.w_l);
        rp = (f1.r_p_lb * ev) / (1.0 + f1.r_p_vl * f1.r_p_lb * ev);
        rs = (f1.r_s_lb * ev) / (1.0 + f1.r_s_vl * f1.r_s_lb * ev);
    }

    const cplx inl = crecip(csqrt_(e1[1]));              // 1 / n_l(w)
    const cplx iNl = crecip(csqrt_(e2[1]));              // 1 / N_l(2w)
    const cplx q = ph.raw ? mk(1.0, 0.0) : (ph.pref * energy / fabs(cos_t)) * inl;   // sqrt(prefactor) / n_l
    const cplx out_p = f2.t_p * iNl;                     // T_p / N_l
    const cplx out_s = f2.t_s * (1.0 + Rs);              // T_s (1 + R^M_s)
    const cplx in_p = csq(f1.t_p * inl);                 // (t_p / n_l)^2
    const cplx in_s = csq(f1.t_s * (1.0 + rs));          // (t_s (1 + r^M_s))^2

    Column c;
    c.Gpp = q * (out_p * in_p);
    c.Gps = q * (out_s * in_p);
    c.Gsp = q * (out_p * in_s);
    c.Gss = q * (out_s * in_s);
    const cplx onepR = 1.0 + Rp, onepr = 1.0 + rp, onemr = 1.0 - rp;
    c.A = (1.0 - Rp) * f2.w_l;
    c.Bz = sin_t * onepR;
    c.Q = ph.zzz_phi_quirk ? onepR * csq(onepr) : sin_t * s2 * (onepR * csq(onepr));
    const cplx a = onemr * f1.w_l;
    const cplx b = sin_t * onepr;
    c.a2 = csq(a);
    c.ab = a * b;
    c.b2 = csq(b);
    return c;
}

// chi2 rows in canonical order (include/sshg.h)
enum { XXX = 0, XYY, XZZ, XYZ, XXZ, XXY, YXX, YYY, YZZ, YYZ, YXZ, YXY, ZXX, ZYY, ZZZ, ZYZ, ZXZ, ZXY };

// phi basis: F = {1, c^2, s c, c^3, s c^2, s^2 c, s^3} with c = cos(phi), s = sin(phi).
// The first three span the even harmonics (0, 2), the last four the odd ones (1, 3):
// F_even(phi + pi) = F_even(phi), F_odd(phi + pi) = -F_odd(phi).
//
// `fold` maps the monomial coefficients of the reference's sums onto F using
// s^2 = 1 - c^2, c = c^3 + s^2 c, s = s c^2 + s^3, and scales by G.
SSHG_HD void fold(cplx G, cplx m1, cplx mc2, cplx msc, cplx ms2, cplx mc3,
                                     cplx msc2, cplx ms2c, cplx ms3, cplx mc, cplx ms, cplx k[7]) {
    k[0] = G * (m1 + ms2);
    k[1] = G * (mc2 - ms2);
    k[2] = G * msc;
    k[3] = G * (mc3 + mc);
    k[4] = G * (msc2 + ms);
    k[5] = G * (ms2c + mc);
    k[6] = G * (ms3 + ms);
}

template <typename ChiFn>   // chi(i) returns row i of the rotated chi2 at this energy
SSHG_HD void coeffs_pp(const Column& c, ChiFn chi, const Physics& ph, cplx k[7]) {
    const cplx Aa2 = c.A * c.a2, Aab = c.A * c.ab, Ab2 = c.A * c.b2;
    const cplx Ba2 = c.Bz * c.a2, Bab = c.Bz * c.ab;
    const cplx zero = mk(0.0, 0.0);
    cplx mc3 = -(Aa2 * chi(XXX));
    cplx msc2 = -(Aa2 * (2.0 * chi(XXY) + chi(YXX)));
    cplx ms2c = -(Aa2 * (chi(XYY) + 2.0 * chi(YXY)));
    cplx ms3 = -(Aa2 * chi(YYY));
    cplx m1 = zero;
    if (ph.zzz_phi_quirk) ms3 = ms3 + c.Q * chi(ZZZ);    // shg.py:269
    else m1 = c.Q * chi(ZZZ);
    cplx mc2 = Ba2 * chi(ZXX) - 2.0 * (Aab * chi(XXZ));
    cplx msc = 2.0 * (Ba2 * chi(ZXY)) - 2.0 * (Aab * (chi(XYZ) + chi(YXZ)));
    cplx ms2 = Ba2 * chi(ZYY) - 2.0 * (Aab * chi(YYZ));
    cplx mc = 2.0 * (Bab * chi(ZXZ)) - Ab2 * chi(XZZ);
    cplx ms = 2.0 * (Bab * chi(ZYZ)) - Ab2 * chi(YZZ);
    fold(c.Gpp, m1, mc2, msc, ms2, mc3, msc2, ms2c, ms3, mc, ms, k);
}

template <typename ChiFn>
SSHG_HD void coeffs_ps(const Column& c, ChiFn chi, cplx k[7]) {
    const cplx zero = mk(0.0, 0.0);
    cplx mc3 = c.a2 * chi(YXX);
    cplx msc2 = c.a2 * (2.0 * chi(YXY) - chi(XXX));
    cplx ms2c = c.a2 * (chi(YYY) - 2.0 * chi(XXY));
    cplx ms3 = -(c.a2 * chi(XYY));
    cplx mc2 = 2.0 * (c.ab * chi(YXZ));
    cplx msc = 2.0 * (c.ab * (chi(YYZ) - chi(XXZ)));
    cplx ms2 = -2.0 * (c.ab * chi(XYZ));
    cplx mc = c.b2 * chi(YZZ);
    cplx ms = -(c.b2 * chi(XZZ));
    fold(c.Gps, zero, mc2, msc, ms2, mc3, msc2, ms2c, ms3, mc, ms, k);
}

template <typename ChiFn>
SSHG_HD void coeffs_sp(const Column& c, ChiFn chi, cplx k[7]) {
    const cplx zero = mk(0.0, 0.0);
    cplx mc3 = -(c.A * chi(XYY));
    cplx msc2 = c.A * (2.0 * chi(XXY) - chi(YYY));
    cplx ms2c = c.A * (2.0 * chi(YXY) - chi(XXX));
    cplx ms3 = -(c.A * chi(YXX));
    cplx mc2 = c.Bz * chi(ZYY);
    cplx msc = -2.0 * (c.Bz * chi(ZXY));
    cplx ms2 = c.Bz * chi(ZXX);
    fold(c.Gsp, zero, mc2, msc, ms2, mc3, msc2, ms2c, ms3, zero, zero, k);
}

// sS has no even part: k[0..3] multiply {c^3, s c^2, s^2 c, s^3}.
template <typename ChiFn>
SSHG_HD void coeffs_ss(const Column& c, ChiFn chi, cplx k[4]) {
    k[0] = c.Gss * chi(YYY);
    k[1] = -(c.Gss * (chi(XYY) + 2.0 * chi(YXY)));
    k[2] = c.Gss * (2.0 * chi(XXY) + chi(YXX));
    k[3] = -(c.Gss * chi(XXX));
}

// basis row for one azimuth: b[0..5] = {c^2, s c, c^3, s c^2, s^2 c, s^3}
SSHG_HD void phi_basis_rad(double phi, double b[6]) {
    double s, c;
    sincos(phi, &s, &c);
    b[0] = c * c;
    b[1] = s * c;
    b[2] = b[0] * c;
    b[3] = b[1] * c;
    b[4] = b[1] * s;
    b[5] = s * s * s;
}

SSHG_HD void phi_basis(double phi_deg, double b[6]) { phi_basis_rad(phi_deg * (3.141592653589793 / 180.0), b); }

// ---- harmonic form of the yield (fused output broadening, K3) --------------------------------------
// |r(phi)|^2 of a 7-term r is a real trigonometric polynomial of degree 6:
//   y(phi) = C0 + sum_{q=1..6} (C_q cos(q phi) + S_q sin(q phi)).
// The Gaussian broadening of the yields along the energy axis (broad(., sigma_out), shg.py:26-28,
// 433-436) is linear, so it commutes with this expansion: broadening the 13 coefficient spectra of a
// (theta, thickness) column once replaces broadening every one of its phi rows.
//
// With r = sum_{n=-3..3} rho_n e^{i n phi}:  y = sum_q Y_q e^{i q phi},  Y_q = sum_n rho_n conj(rho_{n-q}),
// Y_{-q} = conj(Y_q), hence C0 = Y_0, C_q = 2 Re Y_q, S_q = -2 Im Y_q.
//
// Layout of H[13] (and of a phi-table item, without the leading 1):
//   H[0] = C0 | H[1..6] = C2 S2 C4 S4 C6 S6 (even under phi -> phi + 180 deg) | H[7..12] = C1 S1 C3 S3 C5 S5 (odd).
constexpr int NH = 13;
constexpr int NH_SS = 7;      // sS has odd rho only: q = 0, 2, 4, 6

// F-basis coefficients k[7] (see `fold`) -> rho[n + 3], n = -3..3:
//   c^2 = (1 + cos 2phi)/2, s c = sin 2phi / 2, c^3 = (3 cos phi + cos 3phi)/4, s c^2 = (sin phi + sin 3phi)/4,
//   s^2 c = (cos phi - cos 3phi)/4, s^3 = (3 sin phi - sin 3phi)/4.
SSHG_HD void rho_from_coeffs(const cplx k[7], cplx rho[7]) {
    const cplx a0 = k[0] + 0.5 * k[1];
    const cplx a2 = 0.5 * k[1], b2 = 0.5 * k[2];
    const cplx a1 = 0.25 * (3.0 * k[3] + k[5]), b1 = 0.25 * (k[4] + 3.0 * k[6]);
    const cplx a3 = 0.25 * (k[3] - k[5]), b3 = 0.25 * (k[4] - k[6]);
    // alpha cos + beta sin = (alpha - i beta)/2 e^{+i n phi} + (alpha + i beta)/2 e^{-i n phi}
    rho[3] = a0;
    rho[4] = mk(0.5 * (a1.re + b1.im), 0.5 * (a1.im - b1.re));
    rho[2] = mk(0.5 * (a1.re - b1.im), 0.5 * (a1.im + b1.re));
    rho[5] = mk(0.5 * (a2.re + b2.im), 0.5 * (a2.im - b2.re));
    rho[1] = mk(0.5 * (a2.re - b2.im), 0.5 * (a2.im + b2.re));
    rho[6] = mk(0.5 * (a3.re + b3.im), 0.5 * (a3.im - b3.re));
    rho[0] = mk(0.5 * (a3.re - b3.im), 0.5 * (a3.im + b3.re));
}

// ODD_ONLY: rho_0 = rho_{+-2} = 0 (sS); only H[0..6] are written.
template <bool ODD_ONLY>
SSHG_HD void harmonics_from_rho(const cplx rho[7], double* H) {
#pragma unroll
    for (int q = 0; q <= 6; ++q) {
        if (ODD_ONLY && (q & 1)) continue;
        double re = 0.0, im = 0.0;
#pragma unroll
        for (int n = q - 3; n <= 3; ++n) {
            if (ODD_ONLY && !(n & 1)) continue;
            const cplx a = rho[n + 3], b = rho[n - q + 3];
            re = fma(a.re, b.re, fma(a.im, b.im, re));
            if (q) im = fma(a.im, b.re, fma(-a.re, b.im, im));
        }
        if (q == 0) H[0] = re;
        else if (q & 1) { H[q + 6] = 2.0 * re; H[q + 7] = -2.0 * im; }
        else { H[q - 1] = 2.0 * re; H[q] = -2.0 * im; }
    }
}

SSHG_HD void harmonics7(const cplx k[7], double H[NH]) {
    cplx rho[7];
    rho_from_coeffs(k, rho);
    harmonics_from_rho<false>(rho, H);
}

// sS: k[0..3] multiply {c^3, s c^2, s^2 c, s^3}
SSHG_HD void harmonics4(const cplx k[4], double H[NH_SS]) {
    const cplx zero = mk(0.0, 0.0);
    const cplx k7[7] = {zero, zero, zero, k[0], k[1], k[2], k[3]};
    cplx rho[7];
    rho_from_coeffs(k7, rho);
    harmonics_from_rho<true>(rho, H);
}

// phi-table item of the harmonic form: t[0..5] = cos/sin of 2, 4, 6 phi; t[6..11] = cos/sin of 1, 3, 5 phi.
SSHG_HD void phi_harmonics(double phi_deg, double t[12]) {
#pragma unroll
    for (int q = 1; q <= 6; ++q) {
        double s, c;
        sincospi(fmod((double)q * phi_deg, 360.0) / 180.0, &s, &c);
        const int j = (q & 1) ? q + 5 : q - 2;
        t[j] = c;
        t[j + 1] = s;
    }
}

// U = |C0| + sum_q (|C_q| + |S_q|) >= |C0| + sum_q sqrt(C_q^2 + S_q^2) >= max over phi of the harmonic sum.  The rounding error of that sum is
// <= 5e-14 U, so a value below NODE_FRACTION * U (the yield next to an azimuthal node, or an exact symmetry zero,
// where the sum returns rounding noise of either sign) is no longer accurate to 1e-10 relative: K3 flags its row
// and re-evaluates it the reference's way (|r|^2 per energy, then the FIR).  With 1e-3 everything that is NOT
// flagged keeps an element-wise error below 5e-11.
constexpr double NODE_FRACTION = 1e-3;
// A flagged row whose values are all below DEEP_FRACTION * U is an exact node (a symmetry zero; the reference returns
// ~1e-32 of the maximum there, K3 rounding noise of ~1e-15 U, stored as max(noise, 0)).  It is not re-evaluated:
// U <= 13 max_phi y, so such a value is below 1.3e-7 of the maximum of any array that contains its column, i.e. outside
// the element-wise part of the parity metric, and the `1e-10 * max` part holds with four orders to spare.
constexpr double DEEP_FRACTION = 1e-8;

SSHG_HD double amplitude_bound(const double* H, bool odd_only) {
    // |C| + |S| >= sqrt(C^2 + S^2): a bound that needs no square roots (at most sqrt(2) above the tight one)
    double u = fabs(H[0]);
#pragma unroll
    for (int j = 0; j < 6; ++j)
        if (!odd_only || j < 3) u += fabs(H[1 + 2 * j]) + fabs(H[2 + 2 * j]);
    return u;
}

SSHG_HD double harmonic_yield(const double* H, const double t[12], bool odd_only) {
    double ye = H[0];
#pragma unroll
    for (int j = 0; j < 6; ++j) ye = fma(H[1 + j], t[j], ye);
    if (odd_only) return ye;
    double yo = H[7] * t[6];
#pragma unroll
    for (int j = 1; j < 6; ++j) yo = fma(H[7 + j], t[6 + j], yo);
    return ye + yo;
}

// ---- sparse-phi form -------------------------------------------------------------------------
// For sweeps with few azimuths per column (thickness / angle-of-incidence scans) the chi2 tensor is
// contracted with the phi-dependent unit vectors FIRST, once per (energy, phi):
//   r_pP = G_pP ( -A (a2 X0 + 2 ab X1 + b2 X2) + Bz (a2 X3 + 2 ab X4) + Q X5 )
//   r_pS = G_pS (  a2 X6 + 2 ab X7 + b2 X8 )
//   r_sP = G_sP ( -A X9 + Bz X10 )
//   r_sS = G_sS X11
// (the same 18/12/9/6 terms of shg.py:218-269, 290-313, 334-351, 372-377, grouped by their
// theta/thickness-dependent factor), so a column costs 12 complex multiplies instead of the
// 7-coefficient assembly per polarisation.
constexpr int NX = 12;

template <typename ChiFn>
SSHG_HD void contract_phi(ChiFn chi, double s, double c, int zzz_phi_quirk, cplx X[NX]) {
    const double c2 = c * c, sc = s * c, s2 = s * s;
    const double c3 = c2 * c, sc2 = sc * c, s2c = sc * s, s3 = s2 * s;
    X[0] = c3 * chi(XXX) + sc2 * (2.0 * chi(XXY) + chi(YXX)) + s2c * (chi(XYY) + 2.0 * chi(YXY)) + s3 * chi(YYY);
    X[1] = c2 * chi(XXZ) + sc * (chi(XYZ) + chi(YXZ)) + s2 * chi(YYZ);
    X[2] = c * chi(XZZ) + s * chi(YZZ);
    X[3] = c2 * chi(ZXX) + (2.0 * sc) * chi(ZXY) + s2 * chi(ZYY);
    X[4] = c * chi(ZXZ) + s * chi(ZYZ);
    X[5] = zzz_phi_quirk ? s3 * chi(ZZZ) : chi(ZZZ);          // shg.py:269
    X[6] = c3 * chi(YXX) + sc2 * (2.0 * chi(YXY) - chi(XXX)) + s2c * (chi(YYY) - 2.0 * chi(XXY)) - s3 * chi(XYY);
    X[7] = c2 * chi(YXZ) + sc * (chi(YYZ) - chi(XXZ)) - s2 * chi(XYZ);
    X[8] = c * chi(YZZ) - s * chi(XZZ);
    X[9] = c3 * chi(XYY) + sc2 * (chi(YYY) - 2.0 * chi(XXY)) + s2c * (chi(XXX) - 2.0 * chi(YXY)) + s3 * chi(YXX);
    X[10] = c2 * chi(ZYY) - (2.0 * sc) * chi(ZXY) + s2 * chi(ZXX);
    X[11] = c3 * chi(YYY) - sc2 * (chi(XYY) + 2.0 * chi(YXY)) + s2c * (2.0 * chi(XXY) + chi(YXX)) - s3 * chi(XXX);
}

// The part of the column setup that does not depend on the thickness (shg.py:137-171 and the
// prefactor): kept per (energy, theta) while a thread walks the thickness axis.
struct ColumnStatic {
    cplx R_p, R_s, P2p, P2s, W;        // 2w: r^{lb}, r^{vl} r^{lb} (p, s), wave vector in the layer
    cplx r_p, r_s, P1p, P1s, w;        // 1w
    cplx out_p, T_s, in_p, t_s, q;     // T_p / N_l, T_s, (t_p / n_l)^2, t_s, sqrt(prefactor) / n_l
    double sin_t, s3;                  // sin(theta), sin^3(theta)
};

SSHG_HD ColumnStatic column_static(const cplx e1[3], const cplx e2[3], double energy, double sin_t, double cos_t,
                                   const Physics& ph) {
    const double s2 = sin_t * sin_t;
    const Fresnel f1 = fresnel_set(e1[0], e1[1], e1[2], s2);
    const Fresnel f2 = fresnel_set(e2[0], e2[1], e2[2], s2);
    const cplx inl = crecip(csqrt_(e1[1]));
    const cplx iNl = crecip(csqrt_(e2[1]));
    ColumnStatic c;
    c.R_p = f2.r_p_lb; c.R_s = f2.r_s_lb; c.P2p = f2.r_p_vl * f2.r_p_lb; c.P2s = f2.r_s_vl * f2.r_s_lb; c.W = f2.w_l;
    c.r_p = f1.r_p_lb; c.r_s = f1.r_s_lb; c.P1p = f1.r_p_vl * f1.r_p_lb; c.P1s = f1.r_s_vl * f1.r_s_lb; c.w = f1.w_l;
    c.out_p = f2.t_p * iNl;
    c.T_s = f2.t_s;
    c.in_p = csq(f1.t_p * inl);
    c.t_s = f1.t_s;
    c.q = ph.raw ? mk(1.0, 0.0) : (ph.pref * energy / fabs(cos_t)) * inl;
    c.sin_t = sin_t;
    c.s3 = sin_t * s2;
    return c;
}

// The thickness-dependent part (shg.py:174-198) and the four yields of one azimuth from the
// phi-contracted chi2.
struct ColumnDynamic {
    cplx Gpp, Gps, Gsp, Gss, A, Bz, Q, a2, ab2, b2;     // ab2 = 2 a b
};

// The three thickness-dependent phase factors of a column (shg.py:180-183, 193-195):
//   eh = e^{i delta/2}, ev = e^{i varphi}, snc = sinc(delta/2) in the reference's (np.sinc) convention.
struct Phases {
    cplx eh, ev, snc;
};

SSHG_HD cplx sinc_arg(const ColumnStatic& c, double energy, double thick, const Physics& ph) {
    const double PI = 3.141592653589793;
    cplx arg = (4.0 * PI * (energy * thick * ph.kd)) * c.W;      // delta / 2
    if (arg.re == 0.0 && arg.im == 0.0) arg = mk(1.0e-20, 0.0);  // np.sinc: 0 -> 1e-20
    return ph.sinc_normalized ? PI * arg : arg;
}

SSHG_HD Phases phases_exact(const ColumnStatic& c, double energy, double thick, const Physics& ph) {
    const double PI = 3.141592653589793;
    const double x = energy * thick * ph.kd;
    Phases p;
    p.eh = cexp_i((4.0 * PI * x) * c.W);
    p.ev = cexp_i((4.0 * PI * x) * c.w);
    const cplx arg = sinc_arg(c, energy, thick, ph);
    p.snc = csin_(arg) / arg;
    return p;
}

// Uniform thickness grids: the three exponentials e^{i delta/2}, e^{i varphi}, e^{i arg} advance by a
// constant complex factor per step, so a run of columns needs one complex multiply each instead of
// sincos + exp; sin(arg) = (e^{i arg} - e^{-i arg}) / 2i.  The run is re-seeded exactly at its first
// column (<= 32 steps: accumulated error ~1e-14), and |arg| < 0.5, where that difference cancels,
// takes the direct complex sine.
struct PhaseWalk {
    cplx eh, ev, ep;      // current exponentials
    cplx sh, sv, sp;      // per-step factors
};

SSHG_HD void walk_seed(PhaseWalk& w, const ColumnStatic& c, double energy, double thick, double dstep, const Physics& ph) {
    const double PI = 3.141592653589793;
    const double k = 4.0 * PI * energy * ph.kd;
    const double ks = ph.sinc_normalized ? PI * k : k;
    w.eh = cexp_i((k * thick) * c.W);
    w.ev = cexp_i((k * thick) * c.w);
    w.ep = cexp_i((ks * thick) * c.W);
    w.sh = cexp_i((k * dstep) * c.W);
    w.sv = cexp_i((k * dstep) * c.w);
    w.sp = cexp_i((ks * dstep) * c.W);
}

SSHG_HD void walk_advance(PhaseWalk& w) {
    w.eh = w.eh * w.sh;
    w.ev = w.ev * w.sv;
    w.ep = w.ep * w.sp;
}

SSHG_HD Phases walk_phases(const PhaseWalk& w, const ColumnStatic& c, double energy, double thick, const Physics& ph) {
    Phases p;
    p.eh = w.eh;
    p.ev = w.ev;
    const cplx arg = sinc_arg(c, energy, thick, ph);
    if (arg.re * arg.re + arg.im * arg.im < 0.25) {
        p.snc = csin_(arg) / arg;
    } else {
        const cplx em = crecip(w.ep);
        const cplx diff = w.ep - em;                             // 2i sin(arg)
        p.snc = mk(0.5 * diff.im, -0.5 * diff.re) / arg;
    }
    return p;
}

SSHG_HD ColumnDynamic column_from_phases(const ColumnStatic& c, const Phases& f, const Physics& ph) {
    cplx Rp = c.R_p, Rs = c.R_s, rp = c.r_p, rs = c.r_s;
    if (ph.mref) {
        const cplx ed = csq(f.eh);
        Rp = ((c.R_p * f.eh) / (1.0 + c.P2p * ed)) * f.snc;
        Rs = ((c.R_s * f.eh) / (1.0 + c.P2s * ed)) * f.snc;
        rp = (c.r_p * f.ev) / (1.0 + c.P1p * f.ev);
        rs = (c.r_s * f.ev) / (1.0 + c.P1s * f.ev);
    }
    const cplx out_s = c.T_s * (1.0 + Rs);
    const cplx in_s = csq(c.t_s * (1.0 + rs));
    ColumnDynamic d;
    d.Gpp = c.q * (c.out_p * c.in_p);
    d.Gps = c.q * (out_s * c.in_p);
    d.Gsp = c.q * (c.out_p * in_s);
    d.Gss = c.q * (out_s * in_s);
    const cplx onepR = 1.0 + Rp, onepr = 1.0 + rp;
    d.A = (1.0 - Rp) * c.W;
    d.Bz = c.sin_t * onepR;
    d.Q = ph.zzz_phi_quirk ? onepR * csq(onepr) : c.s3 * (onepR * csq(onepr));
    const cplx a = (1.0 - rp) * c.w;
    const cplx b = c.sin_t * onepr;
    d.a2 = csq(a);
    d.ab2 = 2.0 * (a * b);
    d.b2 = csq(b);
    return d;
}

SSHG_HD ColumnDynamic column_dynamic(const ColumnStatic& c, double energy, double thick, const Physics& ph) {
    Phases f{};
    if (ph.mref) f = phases_exact(c, energy, thick, ph);
    return column_from_phases(c, f, ph);
}

SSHG_HD void sparse_yields(const ColumnDynamic& d, const cplx X[NX], double y[4]) {
    const cplx pp = d.Gpp * (d.Bz * (d.a2 * X[3] + d.ab2 * X[4]) + d.Q * X[5] - d.A * (d.a2 * X[0] + d.ab2 * X[1] + d.b2 * X[2]));
    const cplx ps = d.Gps * (d.a2 * X[6] + d.ab2 * X[7] + d.b2 * X[8]);
    const cplx sp = d.Gsp * (d.Bz * X[10] - d.A * X[9]);
    const cplx ss = d.Gss * X[11];
    y[0] = fma(pp.re, pp.re, pp.im * pp.im);
    y[1] = fma(ps.re, ps.re, ps.im * ps.im);
    y[2] = fma(sp.re, sp.re, sp.im * sp.im);
    y[3] = fma(ss.re, ss.re, ss.im * ss.im);
}

// ---- the reference's small helpers, one formula each (shg.py:137-198) -----------------------
SSHG_HD cplx wvec_(cplx eps, double sin_t) { return csqrt_(mk(eps.re - sin_t * sin_t, eps.im)); }
SSHG_HD cplx frefs_(cplx ei, cplx ej, double st) {
    cplx wi = wvec_(ei, st), wj = wvec_(ej, st);
    return (wi - wj) / (wi + wj);
}
SSHG_HD cplx frefp_(cplx ei, cplx ej, double st) {
    cplx wi = wvec_(ei, st), wj = wvec_(ej, st);
    return (wi * ej - wj * ei) / (wi * ej + wj * ei);
}
SSHG_HD cplx ftrans_(cplx ei, cplx ej, double st) {
    cplx wi = wvec_(ei, st), wj = wvec_(ej, st);
    return (2.0 * wi) / (wi + wj);
}
SSHG_HD cplx ftranp_(cplx ei, cplx ej, double st) {
    cplx wi = wvec_(ei, st), wj = wvec_(ej, st);
    return ((2.0 * wi) * csqrt_(ei * ej)) / (wi * ej + wj * ei);
}
// mrc2w (harmonic = 2, shg.py:174-185) / mrc1w (harmonic = 1, shg.py:188-198)
SSHG_HD cplx multiref_(int harmonic, double energy, cplx eps_l, cplx f1, cplx f2, double st, double thick,
                       const Physics& ph) {
    if (!ph.mref) return f1;
    const double PI = 3.141592653589793;
    const double x = energy * thick * ph.kd;
    const cplx w = wvec_(eps_l, st);
    if (harmonic == 1) {
        cplx ev = cexp_i((4.0 * PI * x) * w);
        return (f1 * ev) / (1.0 + f2 * f1 * ev);
    }
    cplx hd = (4.0 * PI * x) * w;
    cplx eh = cexp_i(hd), ed = cexp_i(2.0 * hd);
    cplx arg = hd;
    if (arg.re == 0.0 && arg.im == 0.0) arg = mk(1.0e-20, 0.0);
    if (ph.sinc_normalized) arg = PI * arg;
    return ((f1 * eh) / (1.0 + f2 * f1 * ed)) * (csin_(arg) / arg);
}

}  // namespace sshg
