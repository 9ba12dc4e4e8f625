// TEST HARNESS (not part of libsshg.so): runs the __host__ __device__ physics of
// shgyield_b200/csrc/sshg_math.cuh on the CPU so that the formulas can be checked against the
// golden vectors in the GPU-less build container.  Built by tests/test_hostmath.py.
#include "../shgyield_b200/csrc/sshg_math.cuh"
#include "../include/sshg.h"
#include <math.h>

using namespace sshg;

extern "C" int hostmath_yield_points(const sshg_points* p, double* out) {
    Physics ph;
    ph.kd = 1e-9 / (p->phys.h_eVs * p->phys.c);
    ph.pref = 100.0 / (p->phys.hbar_eVs * sqrt(2.0 * p->phys.eps0 * p->phys.c * p->phys.c * p->phys.c));
    ph.mref = p->phys.mref;
    ph.sinc_normalized = p->phys.sinc_normalized;
    ph.zzz_phi_quirk = p->phys.zzz_phi_quirk;
    ph.raw = 0;
    ph.radians = 0;
    const long long nE = p->nE, n = p->n;
    for (long long i = 0; i < n; ++i) {
        const long long e = p->eidx[i];
        double st, ct;
        sincos(p->theta_deg[i] * (3.141592653589793 / 180.0), &st, &ct);
        cplx e1[3], e2[3];
        for (int m = 0; m < 3; ++m) {
            e1[m] = mk(p->eps1w[2 * (m * nE + e)], p->eps1w[2 * (m * nE + e) + 1]);
            e2[m] = mk(p->eps2w[2 * (m * nE + e)], p->eps2w[2 * (m * nE + e) + 1]);
        }
        const Column c = column_setup(e1, e2, p->energy_eV[e], st, ct, p->thick_nm[i], ph);
        double b[6];
        phi_basis(p->phi_deg[i], b);
        auto chi = [&](int r) { return mk(p->chi2[2 * (r * nE + e)], p->chi2[2 * (r * nE + e) + 1]); };
        for (int pol = 0; pol < 3; ++pol) {
            cplx k[7];
            if (pol == 0) coeffs_pp(c, chi, ph, k);
            else if (pol == 1) coeffs_ps(c, chi, k);
            else coeffs_sp(c, chi, k);
            double re = k[0].re, im = k[0].im;
            for (int j = 0; j < 6; ++j) {
                re = fma(k[j + 1].re, b[j], re);
                im = fma(k[j + 1].im, b[j], im);
            }
            out[pol * n + i] = fma(re, re, im * im);
        }
        cplx k[4];
        coeffs_ss(c, chi, k);
        double re = k[0].re * b[2], im = k[0].im * b[2];
        for (int j = 1; j < 4; ++j) {
            re = fma(k[j].re, b[j + 2], re);
            im = fma(k[j].im, b[j + 2], im);
        }
        out[3 * n + i] = fma(re, re, im * im);
    }
    return 0;
}

// ---- K1 register-window FIR (sshg_fir.cuh) on the CPU: one "thread" = SR consecutive outputs -------
#include "../shgyield_b200/csrc/sshg_fir.cuh"
#include <vector>

// x: n samples, w: r + 1 taps (w[0] centre); out: n outputs with the 'reflect' boundary.
extern "C" int hostmath_fir(const double* x, long long n, const double* w, int r, double* out) {
    const int SR = sshg::SR;
    const int ntap = 2 * r + 1 + 3 * SR;
    std::vector<double> wext(ntap);
    for (int k = 0; k < ntap; ++k) {
        const int j = k - (SR - 1);
        wext[k] = (j >= 0 && j <= 2 * r) ? w[j <= r ? r - j : j - r] : 0.0;
    }
    auto reflect = [&](long long i) {
        long long m = i % (2 * n);
        if (m < 0) m += 2 * n;
        return m >= n ? 2 * n - 1 - m : m;
    };
    std::vector<double> xw(2 * r + 3 * SR);
    for (long long base = 0; base < n; base += SR) {
        for (int t = 0; t < (int)xw.size(); ++t) {
            const long long i = base - r + t;
            xw[t] = (t < 2 * r + SR && i < n + r) ? x[reflect(i)] : 0.0;
        }
        double acc[sshg::SR], wr[sshg::SR];
        for (int q = 0; q < SR; ++q) {
            acc[q] = 0.0;
            wr[q] = wext[q];
        }
        sshg::fir_window(xw.data(), wext.data(), r, acc, wr);
        for (int q = 0; q < SR && base + q < n; ++q) out[base + q] = acc[q];
    }
    return 0;
}

// ---- sparse-phi form (contract_phi, column_static / column_dynamic, sparse_yields) on the CPU --------
extern "C" int hostmath_sparse_points(const sshg_points* p, double* out) {
    Physics ph;
    ph.kd = 1e-9 / (p->phys.h_eVs * p->phys.c);
    ph.pref = 100.0 / (p->phys.hbar_eVs * sqrt(2.0 * p->phys.eps0 * p->phys.c * p->phys.c * p->phys.c));
    ph.mref = p->phys.mref;
    ph.sinc_normalized = p->phys.sinc_normalized;
    ph.zzz_phi_quirk = p->phys.zzz_phi_quirk;
    ph.raw = 0;
    ph.radians = 0;
    const long long nE = p->nE, n = p->n;
    for (long long i = 0; i < n; ++i) {
        const long long e = p->eidx[i];
        double st, ct, s, c;
        sincos(p->theta_deg[i] * (3.141592653589793 / 180.0), &st, &ct);
        sincos(p->phi_deg[i] * (3.141592653589793 / 180.0), &s, &c);
        cplx e1[3], e2[3];
        for (int m = 0; m < 3; ++m) {
            e1[m] = mk(p->eps1w[2 * (m * nE + e)], p->eps1w[2 * (m * nE + e) + 1]);
            e2[m] = mk(p->eps2w[2 * (m * nE + e)], p->eps2w[2 * (m * nE + e) + 1]);
        }
        auto chi = [&](int r) { return mk(p->chi2[2 * (r * nE + e)], p->chi2[2 * (r * nE + e) + 1]); };
        cplx X[NX];
        contract_phi(chi, s, c, ph.zzz_phi_quirk, X);
        const ColumnStatic cs = column_static(e1, e2, p->energy_eV[e], st, ct, ph);
        const ColumnDynamic cd = column_dynamic(cs, p->energy_eV[e], p->thick_nm[i], ph);
        double y[4];
        sparse_yields(cd, X, y);
        for (int k = 0; k < 4; ++k) out[k * n + i] = y[k];
    }
    return 0;
}

// ---- uniform thickness grids: phase factors by recurrence (PhaseWalk) on the CPU ---------------------
// One (energy column e, theta, phi); thick = d0 + k * dstep, k < nd; out[pol][k].
extern "C" int hostmath_sparse_walk(const sshg_points* p, long long e, double theta_deg, double phi_deg, double d0,
                                    double dstep, long long nd, double* out) {
    Physics ph;
    ph.kd = 1e-9 / (p->phys.h_eVs * p->phys.c);
    ph.pref = 100.0 / (p->phys.hbar_eVs * sqrt(2.0 * p->phys.eps0 * p->phys.c * p->phys.c * p->phys.c));
    ph.mref = p->phys.mref;
    ph.sinc_normalized = p->phys.sinc_normalized;
    ph.zzz_phi_quirk = p->phys.zzz_phi_quirk;
    ph.raw = 0;
    ph.radians = 0;
    const long long nE = p->nE;
    double st, ct, s, c;
    sincos(theta_deg * (3.141592653589793 / 180.0), &st, &ct);
    sincos(phi_deg * (3.141592653589793 / 180.0), &s, &c);
    cplx e1[3], e2[3];
    for (int m = 0; m < 3; ++m) {
        e1[m] = mk(p->eps1w[2 * (m * nE + e)], p->eps1w[2 * (m * nE + e) + 1]);
        e2[m] = mk(p->eps2w[2 * (m * nE + e)], p->eps2w[2 * (m * nE + e) + 1]);
    }
    auto chi = [&](int r) { return mk(p->chi2[2 * (r * nE + e)], p->chi2[2 * (r * nE + e) + 1]); };
    cplx X[NX];
    contract_phi(chi, s, c, ph.zzz_phi_quirk, X);
    const double energy = p->energy_eV[e];
    const ColumnStatic cs = column_static(e1, e2, energy, st, ct, ph);
    PhaseWalk w;
    for (long long k = 0; k < nd; ++k) {
        const double thick = d0 + dstep * (double)k;
        if (k % 32 == 0) walk_seed(w, cs, energy, thick, dstep, ph);     // a run of the kernel is at most 32 columns
        else walk_advance(w);
        const ColumnDynamic cd = column_from_phases(cs, walk_phases(w, cs, energy, thick, ph), ph);
        double y[4];
        sparse_yields(cd, X, y);
        for (int q = 0; q < 4; ++q) out[q * nd + k] = y[q];
    }
    return 0;
}

// ---- K3 on the CPU: harmonic form of one (theta, thickness) column, broadened along the energy axis ----
// p: a point list whose spectra columns are the energy axis (eidx unused); theta / thickness scalars;
// phi: nP azimuths; w: r + 1 Gaussian taps (w[0] centre); out[pol][nP][nE].
// strict_nodes: 1 = as the product: a value between DEEP_FRACTION and NODE_FRACTION of its column's amplitude bound
// flags its (polarisation, row, 256-energy tile), which is then re-evaluated row-wise (|r|^2 per energy, FIR), values
// below DEEP_FRACTION (exact symmetry zeros) are stored as max(noise, 0); 0 = the harmonic sums alone.
extern "C" int hostmath_blur_column(const sshg_points* p, double theta_deg, double thick_nm, const double* phi_deg,
                                    long long nP, const double* w, int r, int strict_nodes, double* out) {
    Physics ph;
    ph.kd = 1e-9 / (p->phys.h_eVs * p->phys.c);
    ph.pref = 100.0 / (p->phys.hbar_eVs * sqrt(2.0 * p->phys.eps0 * p->phys.c * p->phys.c * p->phys.c));
    ph.mref = p->phys.mref;
    ph.sinc_normalized = p->phys.sinc_normalized;
    ph.zzz_phi_quirk = p->phys.zzz_phi_quirk;
    ph.raw = 0;
    ph.radians = 0;
    const long long nE = p->nE;
    double st, ct;
    sincos(theta_deg * (3.141592653589793 / 180.0), &st, &ct);
    const int off[5] = {0, NH, 2 * NH, 3 * NH, 3 * NH + NH_SS};
    const int NA = off[4];
    std::vector<double> raw((size_t)NA * nE), blur((size_t)NA * nE);
    for (long long e = 0; e < nE; ++e) {
        cplx e1[3], e2[3];
        for (int m = 0; m < 3; ++m) {
            e1[m] = mk(p->eps1w[2 * (m * nE + e)], p->eps1w[2 * (m * nE + e) + 1]);
            e2[m] = mk(p->eps2w[2 * (m * nE + e)], p->eps2w[2 * (m * nE + e) + 1]);
        }
        const Column c = column_setup(e1, e2, p->energy_eV[e], st, ct, thick_nm, ph);
        auto chi = [&](int q) { return mk(p->chi2[2 * (q * nE + e)], p->chi2[2 * (q * nE + e) + 1]); };
        for (int pol = 0; pol < 3; ++pol) {
            cplx k[7];
            if (pol == 0) coeffs_pp(c, chi, ph, k);
            else if (pol == 1) coeffs_ps(c, chi, k);
            else coeffs_sp(c, chi, k);
            double H[NH];
            harmonics7(k, H);
            for (int m = 0; m < NH; ++m) raw[(size_t)(off[pol] + m) * nE + e] = H[m];
        }
        cplx k[4];
        coeffs_ss(c, chi, k);
        double H[NH_SS];
        harmonics4(k, H);
        for (int m = 0; m < NH_SS; ++m) raw[(size_t)(off[3] + m) * nE + e] = H[m];
    }
    auto reflect = [&](long long i) {
        long long m = i % (2 * nE);
        if (m < 0) m += 2 * nE;
        return m >= nE ? 2 * nE - 1 - m : m;
    };
    for (int a = 0; a < NA; ++a)
        for (long long e = 0; e < nE; ++e) {
            double acc = 0.0;
            for (int j = -r; j <= r; ++j) acc = fma(raw[(size_t)a * nE + reflect(e + j)], w[j < 0 ? -j : j], acc);
            blur[(size_t)a * nE + e] = acc;
        }
    // un-broadened coefficients of every energy, kept for the strict re-evaluation of flagged row segments
    std::vector<cplx> kk((size_t)4 * 7 * nE);
    for (long long e = 0; e < nE; ++e) {
        cplx e1[3], e2[3];
        for (int m = 0; m < 3; ++m) {
            e1[m] = mk(p->eps1w[2 * (m * nE + e)], p->eps1w[2 * (m * nE + e) + 1]);
            e2[m] = mk(p->eps2w[2 * (m * nE + e)], p->eps2w[2 * (m * nE + e) + 1]);
        }
        const Column c = column_setup(e1, e2, p->energy_eV[e], st, ct, thick_nm, ph);
        auto chi = [&](int q) { return mk(p->chi2[2 * (q * nE + e)], p->chi2[2 * (q * nE + e) + 1]); };
        cplx k[7];
        coeffs_pp(c, chi, ph, k); for (int j = 0; j < 7; ++j) kk[(size_t)(0 * 7 + j) * nE + e] = k[j];
        coeffs_ps(c, chi, k);     for (int j = 0; j < 7; ++j) kk[(size_t)(1 * 7 + j) * nE + e] = k[j];
        coeffs_sp(c, chi, k);     for (int j = 0; j < 7; ++j) kk[(size_t)(2 * 7 + j) * nE + e] = k[j];
        coeffs_ss(c, chi, k);     for (int j = 0; j < 4; ++j) kk[(size_t)(3 * 7 + j) * nE + e] = k[j];
    }
    const long long FXT = 256;                                  // energies per CTA of the kernel (wide tiles)
    for (long long ip = 0; ip < nP; ++ip) {
        double t[12], bb[6];
        phi_harmonics(phi_deg[ip], t);
        phi_basis(phi_deg[ip], bb);
        for (int pol = 0; pol < 4; ++pol) {
            std::vector<char> flag((size_t)((nE + FXT - 1) / FXT), 0);
            for (long long e = 0; e < nE; ++e) {
                double H[NH];
                const int nh = pol == 3 ? NH_SS : NH;
                for (int m = 0; m < nh; ++m) H[m] = blur[(size_t)(off[pol] + m) * nE + e];
                double y = harmonic_yield(H, t, pol == 3);
                if (strict_nodes) {
                    const double u = amplitude_bound(H, pol == 3);
                    if (u < 1.0e300 && y < NODE_FRACTION * u) {
                        if (y > DEEP_FRACTION * u) flag[(size_t)(e / FXT)] = 1;   // next to a node: the tile's row is re-evaluated
                        else y = fmax(y, 0.0);                                    // exact node: rounding noise, never negative
                    }
                }
                out[((size_t)pol * nP + ip) * nE + e] = y;
            }
            for (size_t tl = 0; tl < flag.size(); ++tl) {
                if (!flag[tl]) continue;
                auto raw_y = [&](long long e) {
                    const cplx* k = &kk[(size_t)(pol * 7) * nE + e];
                    double re, im;
                    if (pol == 3) {
                        re = k[0].re * bb[2]; im = k[0].im * bb[2];
                        for (int j = 1; j < 4; ++j) { re = fma(k[(size_t)j * nE].re, bb[j + 2], re); im = fma(k[(size_t)j * nE].im, bb[j + 2], im); }
                    } else {
                        re = k[0].re; im = k[0].im;
                        for (int j = 0; j < 6; ++j) { re = fma(k[(size_t)(j + 1) * nE].re, bb[j], re); im = fma(k[(size_t)(j + 1) * nE].im, bb[j], im); }
                    }
                    return fma(re, re, im * im);
                };
                for (long long e = (long long)tl * FXT; e < nE && e < (long long)(tl + 1) * FXT; ++e) {
                    double acc = raw_y(e) * w[0];
                    for (int j = r; j >= 1; --j) acc = fma(raw_y(reflect(e - j)) + raw_y(reflect(e + j)), w[j], acc);
                    out[((size_t)pol * nP + ip) * nE + e] = acc;
                }
            }
        }
    }
    return 0;
}
