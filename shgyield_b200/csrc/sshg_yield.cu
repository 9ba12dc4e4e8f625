// K2: the fused SSHG yield kernel (sm_100a, FP64).
//
// One thread owns two adjacent energies of one (theta, thickness) column.  It
//   (A) evaluates the column setup once (Fresnel factors, multiple reflection, Gamma's:
//       reference shg.py:137-216) and parks the ten shared complex factors in shared memory,
//   (B) per polarisation contracts the rotated chi2 with them into 7 (sS: 4) complex
//       coefficients held in registers (shg.py:218-269, 290-313, 334-351, 372-377), and
//   (C) sweeps the azimuth: r(phi) = sum_k coef_k F_k(phi), yield = |r|^2 (shg.py:433-436),
//       written with streaming 128-bit stores, energy fastest.
// The phi basis F is a small table in shared memory, read as warp broadcasts.  When the phi
// grid contains phi and phi + 180 deg the even/odd split of F gives both rows from one
// evaluation (20 instead of 28 FP64 instructions per energy and row pair; sS: 10).
//
// HBM traffic: 8 B per yield point written once; inputs are 384 B per energy, L2 resident.
#include "sshg_internal.h"
#include "sshg_math.cuh"

#include <math.h>
#include <stdlib.h>

using namespace sshg;

namespace {

constexpr int NT_BIG = 256;    // threads per CTA when the column setup dominates (few phi rows per column)
constexpr int NT_SMALL = 64;   // ... when the stores dominate: 128-energy tiles, up to 7 CTAs per SM
constexpr int EPT = 2;         // energies per thread
constexpr int CHUNK_BIG = 384;    // phi items per shared-memory table chunk (18 KB)
#ifndef SSHG_K2_CHUNK
#define SSHG_K2_CHUNK 256
#endif
constexpr int CHUNK_SMALL = SSHG_K2_CHUNK;  // ... 12 KB for the 64-thread CTAs (32 KB of shared memory each)
constexpr int NCOL = 20;       // doubles per parked Column (10 complex)
#ifndef SSHG_NSEG_MAX
#define SSHG_NSEG_MAX 4
#endif
#ifndef SSHG_NSEG_WAVES
#define SSHG_NSEG_WAVES 4.5
#endif

struct YieldArgs {
    const double2* eps1w;      // [3][nE]
    const double2* eps2w;      // [3][nE]
    const double2* chi;        // [18][nE]
    const double* energy;      // [nE]
    const double* theta_deg;   // [nT]
    const double* thick_nm;    // [nD]
    const double* tab;         // [1 + ...]: header (H as double) then items x 6 doubles
    double* out;
    long long nE, nP;
    long long col0;            // first local column of this launch (out points at it)
    int nTs, nDs, t0, d0;      // shard extents / offsets
    int n_tiles;
    int n_seg;                 // the phi items of a column are cut into n_seg segments, one CTA each (short launches)
    int vec_ok;                // 128-bit stores allowed (nE even, out 16-B aligned, even strides)
    // output layout (K2 / K2s; K3 always writes the cube layout): yield (tl, dl, pol, row, e) of the shard goes to
    // out[tl t_stride + dl d_stride - out_off + pol pol_stride + row row_stride + e].  Cube layout [T][D][pol][P][E]:
    // d_stride = 4 nP nE, t_stride = nDs d_stride, pol_stride = nP nE, row_stride = nE, out_off = col0 d_stride.
    long long t_stride, d_stride, pol_stride, row_stride, out_off;
    Physics ph;
};

// Two adjacent yields of one row.  The store width is known at compile time: the sweeps are instantiated for both, so
// that their inner loops carry no branch on it (a run-time flag costs a BSSY / BRA / BSYNC triple per store).
template <bool VEC>
__device__ __forceinline__ void store2t(double* p, double y0, double y1, bool valid1) {
    if (VEC) {
        __stcs(reinterpret_cast<double2*>(p), make_double2(y0, y1));
    } else {
        __stcs(p, y0);
        if (valid1) __stcs(p + 1, y1);
    }
}

// The phi item list of a sweep, built on the device (exact comparisons in degrees):
//   tab[0] = H      (phi, phi + 180 deg) row pairs: rows j and j + H, j < H           (0: the grid has no such pairing)
//   tab[1] = Q      quad items: phi_j, 180 - phi_j, phi_j + 180 and 360 - phi_j are all on the grid -- rows j, H - j,
//                   j + H, 2 H - j for j = 1 .. Q (true for linspace(0, 360, 721), arange(0, 360), ...)
//   tab[2] = NPI    pair items, tab[3] = PSTEP: pair item k owns the rows k PSTEP and k PSTEP + H (with quads: the
//                   self-mirrored angles 0 and 90 deg; without: every j < H)
//   items from tab[4] (16-byte aligned): the quads, then the pairs, then the un-paired rows 2 H, 2 H + 1, ...
// One evaluation of an item gives all its rows.  K2's basis {1, c^2, sc | c^3, sc^2, s^2 c, s^3}: the first three are
// unchanged and the last four flip under phi -> phi + 180 deg; under phi -> 180 - phi (c -> -c) sc, c^3 and s^2 c flip.
// K3's harmonics: cos(q phi) is even and sin(q phi) odd under phi -> -phi, even-q harmonics are unchanged and odd-q
// ones flip under phi -> phi + 180 deg.
constexpr int TAB_HDR = 4;
struct PhiItems {
    long long H, Q, NPI, PSTEP, NS;
    __device__ __forceinline__ long long total() const { return Q + NPI + NS; }
};
__device__ __forceinline__ PhiItems read_phi_items(const double* __restrict__ tab, long long nP) {
    PhiItems it;
    it.H = (long long)tab[0];
    it.Q = (long long)tab[1];
    it.NPI = (long long)tab[2];
    it.PSTEP = (long long)tab[3];
    it.NS = nP - 2 * it.H;
    return it;
}
// base-angle row of item i
__device__ __forceinline__ long long phi_item_row(const PhiItems& it, long long i) {
    if (i < it.Q) return i + 1;
    if (i < it.Q + it.NPI) return (i - it.Q) * it.PSTEP;
    return 2 * it.H + (i - it.Q - it.NPI);
}

// Decides the structure of the grid and writes the header; every thread of the (single) CTA returns the same PhiItems.
__device__ __forceinline__ PhiItems phi_items_build(const double* __restrict__ phi, long long nP, double* __restrict__ tab,
                                                    int allow_pairs, int allow_quads) {
    __shared__ int bad_pair, bad_quad;
    const long long H = nP / 2;
    if (threadIdx.x == 0) {
        bad_pair = (allow_pairs && H > 0) ? 0 : 1;
        bad_quad = (allow_quads && H >= 3) ? 0 : 1;
    }
    __syncthreads();
    if (allow_pairs && H > 0) {
        int bp = 0, bq = 0;
        for (long long j = threadIdx.x; j < H; j += blockDim.x) {
            if (phi[j + H] != phi[j] + 180.0) bp = 1;
            if (j >= 1 && phi[H - j] != 180.0 - phi[j]) bq = 1;
        }
        if (bp) bad_pair = 1;
        if (bq) bad_quad = 1;
    }
    __syncthreads();
    PhiItems it;
    it.H = bad_pair ? 0 : H;
    const bool quads = !bad_pair && !bad_quad;
    it.Q = quads ? (H - 1) / 2 : 0;
    it.NPI = quads ? H - 2 * it.Q : it.H;
    it.PSTEP = quads ? (it.NPI > 1 ? H / 2 : 1) : 1;
    it.NS = nP - 2 * it.H;
    if (threadIdx.x == 0) {
        tab[0] = (double)it.H;
        tab[1] = (double)it.Q;
        tab[2] = (double)it.NPI;
        tab[3] = (double)it.PSTEP;
    }
    return it;
}

// K2's table: items of 6 doubles {c^2, s c, c^3, s c^2, s^2 c, s^3} of the item's base angle.
__global__ void phi_table_kernel(const double* __restrict__ phi, long long nP, double* __restrict__ tab,
                                 int allow_pairs, int allow_quads) {
    const PhiItems it = phi_items_build(phi, nP, tab, allow_pairs, allow_quads);
    const long long items = it.total();
    for (long long i = threadIdx.x; i < items; i += blockDim.x) {
        double b[6];
        phi_basis(phi[phi_item_row(it, i)], b);
#pragma unroll
        for (int k = 0; k < 6; ++k) tab[TAB_HDR + i * 6 + k] = b[k];
    }
}

struct Coef7 {
    double r[7], i[7];
};

// (C) for a 7-term polarisation, two energies per thread.
template <bool VEC>
__device__ __forceinline__ void sweep7(const Coef7& k0, const Coef7& k1, const double* __restrict__ tb,
                                       int n_pair, int n_single, double* plo, double* phi_, double* psg,
                                       long long stride, long long pair_stride, bool valid1) {
#pragma unroll 2
    for (int i = 0; i < n_pair; ++i) {
        const double b0 = tb[0], b1 = tb[1], b2 = tb[2], b3 = tb[3], b4 = tb[4], b5 = tb[5];
        tb += 6;
        double er0 = fma(k0.r[2], b1, fma(k0.r[1], b0, k0.r[0]));
        double ei0 = fma(k0.i[2], b1, fma(k0.i[1], b0, k0.i[0]));
        double or0 = fma(k0.r[6], b5, fma(k0.r[5], b4, fma(k0.r[4], b3, k0.r[3] * b2)));
        double oi0 = fma(k0.i[6], b5, fma(k0.i[5], b4, fma(k0.i[4], b3, k0.i[3] * b2)));
        double er1 = fma(k1.r[2], b1, fma(k1.r[1], b0, k1.r[0]));
        double ei1 = fma(k1.i[2], b1, fma(k1.i[1], b0, k1.i[0]));
        double or1 = fma(k1.r[6], b5, fma(k1.r[5], b4, fma(k1.r[4], b3, k1.r[3] * b2)));
        double oi1 = fma(k1.i[6], b5, fma(k1.i[5], b4, fma(k1.i[4], b3, k1.i[3] * b2)));
        double ar0 = er0 + or0, ai0 = ei0 + oi0, sr0 = er0 - or0, si0 = ei0 - oi0;
        double ar1 = er1 + or1, ai1 = ei1 + oi1, sr1 = er1 - or1, si1 = ei1 - oi1;
        store2t<VEC>(plo, fma(ar0, ar0, ai0 * ai0), fma(ar1, ar1, ai1 * ai1), valid1);
        store2t<VEC>(phi_, fma(sr0, sr0, si0 * si0), fma(sr1, sr1, si1 * si1), valid1);
        plo += pair_stride;
        phi_ += pair_stride;
    }
#pragma unroll 2
    for (int i = 0; i < n_single; ++i) {
        const double b0 = tb[0], b1 = tb[1], b2 = tb[2], b3 = tb[3], b4 = tb[4], b5 = tb[5];
        tb += 6;
        double r0 = fma(k0.r[6], b5, fma(k0.r[5], b4, fma(k0.r[4], b3, fma(k0.r[3], b2,
                    fma(k0.r[2], b1, fma(k0.r[1], b0, k0.r[0]))))));
        double i0 = fma(k0.i[6], b5, fma(k0.i[5], b4, fma(k0.i[4], b3, fma(k0.i[3], b2,
                    fma(k0.i[2], b1, fma(k0.i[1], b0, k0.i[0]))))));
        double r1 = fma(k1.r[6], b5, fma(k1.r[5], b4, fma(k1.r[4], b3, fma(k1.r[3], b2,
                    fma(k1.r[2], b1, fma(k1.r[1], b0, k1.r[0]))))));
        double i1 = fma(k1.i[6], b5, fma(k1.i[5], b4, fma(k1.i[4], b3, fma(k1.i[3], b2,
                    fma(k1.i[2], b1, fma(k1.i[1], b0, k1.i[0]))))));
        store2t<VEC>(psg, fma(r0, r0, i0 * i0), fma(r1, r1, i1 * i1), valid1);
        psg += stride;
    }
}

// (C) quad items of a 7-term polarisation: the four rows phi (a), 180 - phi (b), phi + 180 (c), 360 - phi (d) from the
// four parts of r -- even / odd under phi + 180 x even / odd under 180 - phi: 36 FP64 instructions per energy for four
// points (9 per point instead of 10).  Rows ascend from pa and pc and descend from pb and pd.
template <bool VEC>
__device__ __forceinline__ void sweep7_quads(const Coef7& k0, const Coef7& k1, const double* __restrict__ tb, int n,
                                             double* pa, double* pb, double* pc, double* pd, long long stride, bool valid1) {
    auto rows4 = [](const Coef7& k, double b0, double b1, double b2, double b3, double b4, double b5, double& ya, double& yb,
                    double& yc, double& yd) {
        const double epr = fma(k.r[1], b0, k.r[0]), epi = fma(k.i[1], b0, k.i[0]);           // 1, c^2
        const double emr = k.r[2] * b1, emi = k.i[2] * b1;                                   // s c
        const double opr = fma(k.r[6], b5, k.r[4] * b3), opi = fma(k.i[6], b5, k.i[4] * b3); // s c^2, s^3
        const double omr = fma(k.r[5], b4, k.r[3] * b2), omi = fma(k.i[5], b4, k.i[3] * b2); // c^3, s^2 c
        const double Epr = epr + emr, Epi = epi + emi, Emr = epr - emr, Emi = epi - emi;
        const double Opr = opr + omr, Opi = opi + omi, Omr = opr - omr, Omi = opi - omi;
        const double ar = Epr + Opr, ai = Epi + Opi, cr = Epr - Opr, ci = Epi - Opi;
        const double br = Emr + Omr, bi = Emi + Omi, dr = Emr - Omr, di = Emi - Omi;
        ya = fma(ar, ar, ai * ai);
        yb = fma(br, br, bi * bi);
        yc = fma(cr, cr, ci * ci);
        yd = fma(dr, dr, di * di);
    };
#pragma unroll 2
    for (int i = 0; i < n; ++i) {
        const double b0 = tb[0], b1 = tb[1], b2 = tb[2], b3 = tb[3], b4 = tb[4], b5 = tb[5];
        tb += 6;
        double a0, b0y, c0, d0, a1, b1y, c1, d1;
        rows4(k0, b0, b1, b2, b3, b4, b5, a0, b0y, c0, d0);
        rows4(k1, b0, b1, b2, b3, b4, b5, a1, b1y, c1, d1);
        store2t<VEC>(pa, a0, a1, valid1);
        store2t<VEC>(pb, b0y, b1y, valid1);
        store2t<VEC>(pc, c0, c1, valid1);
        store2t<VEC>(pd, d0, d1, valid1);
        pa += stride;
        pc += stride;
        pb -= stride;
        pd -= stride;
    }
}

struct Coef4 {
    double r[4], i[4];
};

// (C) for sS: odd part only, so the row at phi + 180 deg equals the row at phi.
template <bool VEC>
__device__ __forceinline__ void sweep4(const Coef4& k0, const Coef4& k1, const double* __restrict__ tb,
                                       int n_pair, int n_single, double* plo, double* phi_, double* psg,
                                       long long stride, long long pair_stride, bool valid1) {
#pragma unroll 2
    for (int i = 0; i < n_pair + n_single; ++i) {
        const double b2 = tb[2], b3 = tb[3], b4 = tb[4], b5 = tb[5];
        tb += 6;
        double r0 = fma(k0.r[3], b5, fma(k0.r[2], b4, fma(k0.r[1], b3, k0.r[0] * b2)));
        double i0 = fma(k0.i[3], b5, fma(k0.i[2], b4, fma(k0.i[1], b3, k0.i[0] * b2)));
        double r1 = fma(k1.r[3], b5, fma(k1.r[2], b4, fma(k1.r[1], b3, k1.r[0] * b2)));
        double i1 = fma(k1.i[3], b5, fma(k1.i[2], b4, fma(k1.i[1], b3, k1.i[0] * b2)));
        const double y0 = fma(r0, r0, i0 * i0), y1 = fma(r1, r1, i1 * i1);
        if (i < n_pair) {
            store2t<VEC>(plo, y0, y1, valid1);
            store2t<VEC>(phi_, y0, y1, valid1);
            plo += pair_stride;
            phi_ += pair_stride;
        } else {
            store2t<VEC>(psg, y0, y1, valid1);
            psg += stride;
        }
    }
}

// (C) quad items of sS (odd part only): rows a and c are equal, and so are b and d.
template <bool VEC>
__device__ __forceinline__ void sweep4_quads(const Coef4& k0, const Coef4& k1, const double* __restrict__ tb, int n,
                                             double* pa, double* pb, double* pc, double* pd, long long stride, bool valid1) {
    auto rows2 = [](const Coef4& k, double b2, double b3, double b4, double b5, double& ya, double& yb) {
        const double opr = fma(k.r[3], b5, k.r[1] * b3), opi = fma(k.i[3], b5, k.i[1] * b3); // s c^2, s^3
        const double omr = fma(k.r[2], b4, k.r[0] * b2), omi = fma(k.i[2], b4, k.i[0] * b2); // c^3, s^2 c
        const double ar = opr + omr, ai = opi + omi, br = opr - omr, bi = opi - omi;
        ya = fma(ar, ar, ai * ai);
        yb = fma(br, br, bi * bi);
    };
#pragma unroll 2
    for (int i = 0; i < n; ++i) {
        const double b2 = tb[2], b3 = tb[3], b4 = tb[4], b5 = tb[5];
        tb += 6;
        double a0, b0y, a1, b1y;
        rows2(k0, b2, b3, b4, b5, a0, b0y);
        rows2(k1, b2, b3, b4, b5, a1, b1y);
        store2t<VEC>(pa, a0, a1, valid1);
        store2t<VEC>(pc, a0, a1, valid1);
        store2t<VEC>(pb, b0y, b1y, valid1);
        store2t<VEC>(pd, b0y, b1y, valid1);
        pa += stride;
        pc += stride;
        pb -= stride;
        pd -= stride;
    }
}

__device__ __forceinline__ cplx ldc(const double2* p) {
    double2 v = __ldg(p);
    return mk(v.x, v.y);
}

// QUADS: the item list may contain four-row items (instantiated apart so that the default pair form keeps its registers)
template <int NT, int CHUNK, bool QUADS>
__global__ void __launch_bounds__(NT, (NT >= 256 ? 2 : 7)) yield_grid_kernel(const YieldArgs a) {
    extern __shared__ double smem[];
    double* tabs = smem;                        // [CHUNK][6]
    double* cols = smem + CHUNK * 6;            // [EPT][NCOL][NT]

    const long long blk = blockIdx.x;
    const int tile = (int)(blk % a.n_tiles);
    const int seg = (int)((blk / a.n_tiles) % a.n_seg);
    const long long lcol = blk / a.n_tiles / a.n_seg;   // column within this launch
    const long long col = a.col0 + lcol;        // local column of the shard = tl * nDs + dl
    const int tl = (int)(col / a.nDs), dl = (int)(col % a.nDs);
    const int tid = threadIdx.x;
    const long long e0 = ((long long)tile * NT + tid) * EPT;
    const bool active = e0 < a.nE;
    const bool valid1 = e0 + 1 < a.nE;
    const long long nE = a.nE;
    const long long ee[2] = {active ? e0 : nE - 1, valid1 ? e0 + 1 : nE - 1};

    // ---- (A) column setup, one energy at a time to keep register pressure down -------------
    {
        double st, ct;
        sincos(a.theta_deg[a.t0 + tl] * (3.141592653589793 / 180.0), &st, &ct);
        const double thick = a.thick_nm[a.d0 + dl];
#pragma unroll 1
        for (int k = 0; k < EPT; ++k) {
            const long long e = ee[k];
            cplx e1[3], e2[3];
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                e1[m] = ldc(a.eps1w + m * nE + e);
                e2[m] = ldc(a.eps2w + m * nE + e);
            }
            Column c = column_setup(e1, e2, a.energy[e], st, ct, thick, a.ph);
            double* d = cols + (k * NCOL) * NT + tid;
            const cplx v[10] = {c.Gpp, c.Gps, c.Gsp, c.Gss, c.A, c.Bz, c.Q, c.a2, c.ab, c.b2};
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                d[(2 * j) * NT] = v[j].re;
                d[(2 * j + 1) * NT] = v[j].im;
            }
        }
    }

    const PhiItems it = read_phi_items(a.tab, a.nP);
    const long long items = it.total();
    const bool vec = a.vec_ok != 0;
    const long long rows_per_pol = a.pol_stride, rs = a.row_stride;
    double* out_col = a.out + ((long long)tl * a.t_stride + (long long)dl * a.d_stride - a.out_off) + e0;

    auto col_field = [&](int k, int j) -> cplx {
        const double* d = cols + (k * NCOL) * NT + tid;
        return mk(d[(2 * j) * NT], d[(2 * j + 1) * NT]);
    };
    auto load_column = [&](int k) -> Column {
        Column c;
        c.Gpp = col_field(k, 0); c.Gps = col_field(k, 1); c.Gsp = col_field(k, 2); c.Gss = col_field(k, 3);
        c.A = col_field(k, 4); c.Bz = col_field(k, 5); c.Q = col_field(k, 6);
        c.a2 = col_field(k, 7); c.ab = col_field(k, 8); c.b2 = col_field(k, 9);
        return c;
    };

    const long long seg_len = (items + a.n_seg - 1) / a.n_seg;
    const long long seg_hi = min(items, (seg + 1) * seg_len);
    for (long long base = seg * seg_len; base < seg_hi; base += CHUNK) {
        const int n_items = (int)min((long long)CHUNK, seg_hi - base);
        __syncthreads();                        // previous chunk fully consumed; columns parked
        for (int i = tid; i < n_items * 6; i += NT) tabs[i] = a.tab[TAB_HDR + base * 6 + i];
        __syncthreads();
        if (!active) continue;
        // the kinds of item in [base, base + n_items): quads, then pair items, then un-paired rows
        const long long q0 = QUADS ? min(base, it.Q) : 0, q1 = QUADS ? min(base + n_items, it.Q) : 0;
        const long long p0 = min(max(base - it.Q, 0LL), it.NPI), p1 = min(max(base + n_items - it.Q, 0LL), it.NPI);
        const long long s0 = max(base - it.Q - it.NPI, 0LL);
        const int nq = (int)(q1 - q0), np = (int)(p1 - p0), ns = n_items - nq - np;
        const double* tq = tabs;
        const double* tp = tabs + 6 * nq;       // pair items, then singles

#pragma unroll 1
        for (int pol = 0; pol < 4; ++pol) {
            double* prow = out_col + pol * rows_per_pol;
            double* pa = prow + (q0 + 1) * rs;                           // quads: rows j, H - j, j + H, 2 H - j
            double* pb = prow + (it.H - q0 - 1) * rs;
            double* pc = prow + (q0 + 1 + it.H) * rs;
            double* pd = prow + (2 * it.H - q0 - 1) * rs;
            const long long pj = QUADS ? p0 * it.PSTEP : p0;
            double* plo = prow + pj * rs;                                // pair item k: rows k PSTEP and k PSTEP + H
            double* phi_ = prow + (pj + it.H) * rs;
            double* psg = prow + (2 * it.H + s0) * rs;                   // un-paired rows
            const long long pstride = QUADS ? it.PSTEP * rs : rs;
            if (pol == 3) {
                Coef4 k[2];
#pragma unroll
                for (int q = 0; q < EPT; ++q) {
                    const long long e = ee[q];
                    Column c = load_column(q);
                    cplx kk[4];
                    coeffs_ss(c, [&](int r) { return ldc(a.chi + r * nE + e); }, kk);
#pragma unroll
                    for (int j = 0; j < 4; ++j) { k[q].r[j] = kk[j].re; k[q].i[j] = kk[j].im; }
                }
                if (vec) {
                    if (QUADS) sweep4_quads<true>(k[0], k[1], tq, nq, pa, pb, pc, pd, rs, valid1);
                    sweep4<true>(k[0], k[1], tp, np, ns, plo, phi_, psg, rs, pstride, valid1);
                } else {
                    if (QUADS) sweep4_quads<false>(k[0], k[1], tq, nq, pa, pb, pc, pd, rs, valid1);
                    sweep4<false>(k[0], k[1], tp, np, ns, plo, phi_, psg, rs, pstride, valid1);
                }
            } else {
                Coef7 k[2];
#pragma unroll
                for (int q = 0; q < EPT; ++q) {
                    const long long e = ee[q];
                    Column c = load_column(q);
                    auto chi = [&](int r) { return ldc(a.chi + r * nE + e); };
                    cplx kk[7];
                    if (pol == 0) coeffs_pp(c, chi, a.ph, kk);
                    else if (pol == 1) coeffs_ps(c, chi, kk);
                    else coeffs_sp(c, chi, kk);
#pragma unroll
                    for (int j = 0; j < 7; ++j) { k[q].r[j] = kk[j].re; k[q].i[j] = kk[j].im; }
                }
                if (vec) {
                    if (QUADS) sweep7_quads<true>(k[0], k[1], tq, nq, pa, pb, pc, pd, rs, valid1);
                    sweep7<true>(k[0], k[1], tp, np, ns, plo, phi_, psg, rs, pstride, valid1);
                } else {
                    if (QUADS) sweep7_quads<false>(k[0], k[1], tq, nq, pa, pb, pc, pd, rs, valid1);
                    sweep7<false>(k[0], k[1], tp, np, ns, plo, phi_, psg, rs, pstride, valid1);
                }
            }
        }
    }
}


#include "sshg_blur.cuh"   // K3: yield_blur_kernel (the sweep with the output broadening in one pass)

// ---- sparse-phi sweeps (few azimuths per column: thickness / incidence-angle scans) ---------------
// chi2 is contracted with the phi-dependent unit vectors once per (energy, phi) (`phi_contract_kernel`,
// X[p][12][nE]); `yield_sparse_kernel` then gives one thread one energy and a run of consecutive
// columns (theta-major, thickness fastest): the thickness-independent Fresnel factors are computed once
// per (energy, theta) and only the exponentials / multiple-reflection coefficients per thickness.
struct SparseArgs {
    const double2* eps1w;
    const double2* eps2w;
    const double2* chi;
    const double2* X;          // [nP][NX][nE]
    const double* energy;
    const double* theta_deg;
    const double* thick_nm;
    const double* phi_deg;
    double* out;               // column col0 of the shard is at out[0]
    long long nE, nP;
    long long col0, ncols;     // columns of this launch (col = tl * nDs + dl)
    int nDs, t0, d0;
    int n_tiles, chunk;        // energies tiles of SNT threads; columns per CTA
    const double* dinfo;       // [2]: {1.0 if thick_nm is uniformly spaced, its step} (thick_uniform_kernel)
    long long t_stride, d_stride, pol_stride, row_stride, out_off;     // output layout, as in YieldArgs
    Physics ph;
};

constexpr int SNT = 128;

// Is the thickness axis uniformly spaced?  (Decided on the device so that device-resident inputs need
// no host round trip.)  info[0] = 1.0 / 0.0, info[1] = step.
__global__ void thick_uniform_kernel(const double* __restrict__ thick, long long nD, double* __restrict__ info,
                                     int allow) {
    __shared__ int bad;
    __shared__ double scale_s;
    if (threadIdx.x == 0) {
        bad = (allow && nD >= 3) ? 0 : 1;
        double sc = 0.0;
        for (long long k = 0; k < nD; ++k) sc = fmax(sc, fabs(thick[k]));
        scale_s = sc;
    }
    __syncthreads();
    const double d0 = thick[0];
    const double step = nD > 1 ? (thick[nD - 1] - d0) / (double)(nD - 1) : 0.0;
    if (!bad) {
        int mybad = (step == 0.0 || !isfinite(step)) ? 1 : 0;
        for (long long k = threadIdx.x; k < nD; k += blockDim.x)
            if (!(fabs(thick[k] - (d0 + step * (double)k)) <= 1e-13 * scale_s)) mybad = 1;
        if (mybad) bad = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        info[0] = bad ? 0.0 : 1.0;
        info[1] = step;
    }
}

__global__ void phi_contract_kernel(const double2* __restrict__ chi, const double* __restrict__ phi_deg,
                                    double2* __restrict__ X, long long nE, long long nP, int zzz_phi_quirk) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nE * nP) return;
    const long long e = g % nE, p = g / nE;
    double s, c;
    sincos(phi_deg[p] * (3.141592653589793 / 180.0), &s, &c);
    cplx x[NX];
    contract_phi([&](int r) { return ldc(chi + r * nE + e); }, s, c, zzz_phi_quirk, x);
#pragma unroll
    for (int k = 0; k < NX; ++k) X[(p * NX + k) * nE + e] = make_double2(x[k].re, x[k].im);
}

#ifndef SSHG_SPARSE_MINB
#define SSHG_SPARSE_MINB 3
#endif
__global__ void __launch_bounds__(SNT, SSHG_SPARSE_MINB) yield_sparse_kernel(const SparseArgs a) {
    const long long blk = blockIdx.x;
    const int tile = (int)(blk % a.n_tiles);
    const long long chunk_id = blk / a.n_tiles;
    const long long e = (long long)tile * SNT + threadIdx.x;
    if (e >= a.nE) return;
    const long long nE = a.nE;
    const long long c_lo = a.col0 + chunk_id * a.chunk;
    const long long c_hi = min(c_lo + a.chunk, a.col0 + a.ncols);
    cplx e1[3], e2[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        e1[m] = ldc(a.eps1w + m * nE + e);
        e2[m] = ldc(a.eps2w + m * nE + e);
    }
    const double energy = a.energy[e];
    ColumnStatic cs;
    PhaseWalk walk;
    const bool recur = a.dinfo[0] != 0.0 && a.ph.mref;
    const double d_step = a.dinfo[1];
    int tl_cur = -1;
    int tl = (int)(c_lo / a.nDs), dl = (int)(c_lo % a.nDs) - 1;      // one division per run, not per column
    for (long long col = c_lo; col < c_hi; ++col) {
        if (++dl == a.nDs) {
            dl = 0;
            ++tl;
        }
        const double thick = a.thick_nm[a.d0 + dl];
        const bool fresh = tl != tl_cur;
        if (fresh) {                                       // new angle of incidence
            double st, ct;
            sincos(a.theta_deg[a.t0 + tl] * (3.141592653589793 / 180.0), &st, &ct);
            cs = column_static(e1, e2, energy, st, ct, a.ph);
            tl_cur = tl;
        }
        ColumnDynamic cd;
        if (recur) {
            if (fresh) walk_seed(walk, cs, energy, thick, d_step, a.ph);     // exact at the first column of a run
            else walk_advance(walk);
            cd = column_from_phases(cs, walk_phases(walk, cs, energy, thick, a.ph), a.ph);
        } else {
            cd = column_dynamic(cs, energy, thick, a.ph);
        }
        double* o = a.out + ((long long)tl * a.t_stride + (long long)dl * a.d_stride - a.out_off) + e;
        for (long long p = 0; p < a.nP; ++p) {
            cplx x[NX];
#pragma unroll
            for (int k = 0; k < NX; ++k) x[k] = ldc(a.X + (p * NX + k) * nE + e);
            double y[4];
            sparse_yields(cd, x, y);
#pragma unroll
            for (int k = 0; k < 4; ++k) __stcs(o + k * a.pol_stride + p * a.row_stride, y[k]);
        }
    }
}

// General NumPy-broadcast shape: one thread per point, all four polarisations.
struct PointArgs {
    const double2* eps1w;
    const double2* eps2w;
    const double2* chi;
    const double* energy;
    const long long* eidx;
    const double* theta_deg;
    const double* phi_deg;
    const double* thick_nm;
    double* out;               // [4][n]
    long long n, nE;
    Physics ph;
};

// AMP = false: yields [4][n];  AMP = true: complex amplitudes Gamma * r, [4][n] double2
// (rad_pp / rad_ps / rad_sp / rad_ss of the reference, shg.py:201-378).
template <bool AMP>
__global__ void __launch_bounds__(128) yield_points_kernel(const PointArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const long long e = a.eidx[i], nE = a.nE;
    if (e < 0 || e >= nE) {                     // device-resident index lists are not validated on the host
        const double bad = nan("");
#pragma unroll
        for (int pol = 0; pol < 4; ++pol) {
            if (AMP) reinterpret_cast<double2*>(a.out)[pol * a.n + i] = make_double2(bad, bad);
            else a.out[pol * a.n + i] = bad;
        }
        return;
    }
    const double to_rad = a.ph.radians ? 1.0 : (3.141592653589793 / 180.0);
    double st, ct;
    sincos(a.theta_deg[i] * to_rad, &st, &ct);
    cplx e1[3], e2[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        e1[m] = ldc(a.eps1w + m * nE + e);
        e2[m] = ldc(a.eps2w + m * nE + e);
    }
    const Column c = column_setup(e1, e2, a.energy[e], st, ct, a.thick_nm[i], a.ph);
    double b[6];
    phi_basis_rad(a.phi_deg[i] * to_rad, b);
    auto chi = [&](int r) { return ldc(a.chi + r * nE + e); };
    auto emit = [&](int pol, double re, double im) {
        if (AMP) reinterpret_cast<double2*>(a.out)[pol * a.n + i] = make_double2(re, im);
        else a.out[pol * a.n + i] = fma(re, re, im * im);
    };
#pragma unroll 1
    for (int pol = 0; pol < 3; ++pol) {
        cplx k[7];
        if (pol == 0) coeffs_pp(c, chi, a.ph, k);
        else if (pol == 1) coeffs_ps(c, chi, k);
        else coeffs_sp(c, chi, k);
        double re = k[0].re, im = k[0].im;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            re = fma(k[j + 1].re, b[j], re);
            im = fma(k[j + 1].im, b[j], im);
        }
        emit(pol, re, im);
    }
    {
        cplx k[4];
        coeffs_ss(c, chi, k);
        double re = k[0].re * b[2], im = k[0].im * b[2];
#pragma unroll
        for (int j = 1; j < 4; ++j) {
            re = fma(k[j].re, b[j + 2], re);
            im = fma(k[j].im, b[j + 2], im);
        }
        emit(3, re, im);
    }
}

Physics make_physics(const sshg_physics& p) {
    Physics ph;
    ph.kd = 1e-9 / (p.h_eVs * p.c);
    ph.pref = 100.0 / (p.hbar_eVs * sqrt(2.0 * p.eps0 * p.c * p.c * p.c));
    ph.mref = p.mref;
    ph.sinc_normalized = p.sinc_normalized;
    ph.zzz_phi_quirk = p.zzz_phi_quirk;
    ph.raw = 0;
    ph.radians = (p.flags & SSHG_FLAG_RADIANS) ? 1 : 0;
    return ph;
}

int check_physics(const sshg_physics& p) {
    SSHG_REQUIRE(p.h_eVs > 0 && p.hbar_eVs > 0 && p.eps0 > 0 && p.c > 0,
                 "sshg_physics: h_eVs, hbar_eVs, eps0 and c must be positive");
    return SSHG_OK;
}

constexpr size_t grid_smem(int nt, int chunk) { return (size_t)(chunk * 6 + EPT * NCOL * nt) * sizeof(double); }
// resident 64-thread CTAs per SM: shared memory (228 KB, 1 KB per CTA reserved), at most 8 by registers (128 per thread)
constexpr int K2_SMALL_PER_SM_SMEM = (int)((228 * 1024) / (grid_smem(NT_SMALL, CHUNK_SMALL) + 1024));
constexpr int K2_SMALL_PER_SM = K2_SMALL_PER_SM_SMEM < 8 ? K2_SMALL_PER_SM_SMEM : 8;

}  // namespace

// layout: null = the cube layout [T][D][pol][P][E]; else {t_stride, d_stride, pol_stride, phi_stride} in doubles (energy
// contiguous): the un-broadened yields of the whole grid written straight into a caller-defined device layout
// (sshg_yield_broadcast).
static int yield_impl(sshg_ctx* ctx, const sshg_problem* p, double sigma_out, double* out, int where,
                      const int64_t* layout = nullptr) {
    SSHG_TRACE("sshg_yield");
    SSHG_REQUIRE(ctx && p && out, "sshg_yield: null argument");
    SSHG_REQUIRE(sigma_out >= 0.0 && isfinite(sigma_out), "sshg_yield: sigma_out must be finite and >= 0 (got %g)",
                 sigma_out);
    SSHG_REQUIRE(p->nE > 0 && p->nT > 0 && p->nP > 0 && p->nD > 0,
                 "sshg_yield: nE, nT, nP, nD must be positive (got %lld %lld %lld %lld)",
                 (long long)p->nE, (long long)p->nT, (long long)p->nP, (long long)p->nD);
    SSHG_REQUIRE(p->energy_eV && p->theta_deg && p->phi_deg && p->thick_nm && p->eps1w && p->eps2w && p->chi2,
                 "sshg_yield: null input array");
    if (int rc = check_physics(p->phys)) return rc;
    // shard: t1 = -1 or t0 = t1 = 0 is the full range; any other empty range is a legal no-op (include/sshg.h)
    const bool t_full = p->t1 == -1 || (p->t0 == 0 && p->t1 == 0), d_full = p->d1 == -1 || (p->d0 == 0 && p->d1 == 0);
    const int64_t t0 = t_full ? 0 : p->t0, t1 = t_full ? p->nT : p->t1, d0 = d_full ? 0 : p->d0, d1 = d_full ? p->nD : p->d1;
    SSHG_REQUIRE(0 <= t0 && t0 <= t1 && t1 <= p->nT, "sshg_yield: bad theta shard [%lld,%lld) of %lld",
                 (long long)t0, (long long)t1, (long long)p->nT);
    SSHG_REQUIRE(0 <= d0 && d0 <= d1 && d1 <= p->nD, "sshg_yield: bad thickness shard [%lld,%lld) of %lld",
                 (long long)d0, (long long)d1, (long long)p->nD);
    if (t0 == t1 || d0 == d1) return SSHG_OK;            // empty shard (more ranks than thetas): nothing to do
    SSHG_REQUIRE(p->nP < (1LL << 31) && p->nE < (1LL << 40), "sshg_yield: axis too long");
    SSHG_CUDA(cudaSetDevice(ctx->device));

    const int in_dev = (where & SSHG_IN_DEVICE) != 0, out_dev = (where & SSHG_OUT_DEVICE) != 0;
    const size_t nE = (size_t)p->nE;
    YieldArgs a{};
    const void* d;
    if (int rc = sshg_stage_in(ctx, 0, p->eps1w, 3 * nE * 16, in_dev, &d)) return rc;
    a.eps1w = (const double2*)d;
    if (int rc = sshg_stage_in(ctx, 1, p->eps2w, 3 * nE * 16, in_dev, &d)) return rc;
    a.eps2w = (const double2*)d;
    if (int rc = sshg_stage_in(ctx, 2, p->chi2, SSHG_NCHI * nE * 16, in_dev, &d)) return rc;
    a.chi = (const double2*)d;
    if (int rc = sshg_stage_in(ctx, 3, p->energy_eV, nE * 8, in_dev, &d)) return rc;
    a.energy = (const double*)d;
    if (int rc = sshg_stage_in(ctx, 4, p->theta_deg, (size_t)p->nT * 8, in_dev, &d)) return rc;
    a.theta_deg = (const double*)d;
    if (int rc = sshg_stage_in(ctx, 5, p->thick_nm, (size_t)p->nD * 8, in_dev, &d)) return rc;
    a.thick_nm = (const double*)d;
    const void* dphi;
    if (int rc = sshg_stage_in(ctx, 6, p->phi_deg, (size_t)p->nP * 8, in_dev, &dphi)) return rc;
    SSHG_REQUIRE(((uintptr_t)a.eps1w | (uintptr_t)a.eps2w | (uintptr_t)a.chi) % 16 == 0,
                 "sshg_yield: complex device arrays must be 16-byte aligned");

    // Output broadening (shg.py:433-436): radius as scipy.ndimage (int(4 sigma + 0.5)); a radius of 0 is the
    // identity.  Dense sweeps with radius <= 200 take the fused kernel K3, everything else K2 -> K1 slab by slab.
    // Option "k3" = 0 forces the two-pass path (tests compare both).
    const sshg_options& opt = ctx->opt;                  // test / tuning switches (sshg_ctx_set_option); -1 = automatic
    bool sparse = 4 * p->nP < 64;
    if (opt.k2_sparse >= 0) sparse = opt.k2_sparse != 0;
    const long long radius = sigma_out > 1e-15 ? (long long)(4.0 * sigma_out + 0.5) : 0;
    SSHG_REQUIRE(radius < (1LL << 28), "sshg_yield: sigma_out %g is too large", sigma_out);
    bool fused = radius >= 1 && radius <= BLUR_RMAX && !sparse;
    if (p->phys.flags & SSHG_FLAG_STRICT_BROADENING) fused = false;
    if (opt.k3 == 0) fused = false;
    const bool two_pass = radius >= 1 && !fused;
    if (layout) {
        SSHG_REQUIRE(out_dev && radius == 0 && t0 == 0 && t1 == p->nT && d0 == 0 && d1 == p->nD,
                     "sshg_yield: a caller-defined layout needs device output, no output broadening and the full grid");
    }

    void* tab;
    if (int rc = sshg_scratch(ctx, 7, (size_t)(4 + 12 * p->nP) * 8, &tab)) return rc;
    a.tab = (const double*)tab;
    // Quad items (four rows per evaluation) cut the instructions per point but make every thread write four row streams,
    // two of them descending.  Measured (profiles/r02_k3_quick_quads.jsonl, same process, same box): K2 is store-bound and
    // gets slower with them (config 3: 0.416 against 0.387 ms; config-5 shards +-1 %), so K2 uses them only on request;
    // K3 gains where it is bound by instruction latency -- short launches (config 3: 0.491 against 0.512 ms) and large
    // radii (sigma_out = 12: -3 ... -4 %) -- and loses 1-3 % on the long store-bound launches of config 5 at sigma_out <= 5.
    bool quads = opt.quads == 1;
    // (decided on the WHOLE sweep, not on the shard: a shard must stay bitwise equal to its slice of the full cube)
    const double k3_waves = (double)(p->nT * p->nD) * (double)((p->nE + 255) / 256) / (2.0 * ctx->num_sms);
    if (fused && opt.quads < 0) quads = k3_waves < 4.0 || radius > 32;
    if (fused) phi_harm_table_kernel<<<1, 256, 0, ctx->stream>>>((const double*)dphi, p->nP, (double*)tab, 1, quads ? 1 : 0);
    else phi_table_kernel<<<1, 256, 0, ctx->stream>>>((const double*)dphi, p->nP, (double*)tab, 1, quads ? 1 : 0);
    SSHG_CUDA(cudaGetLastError());
    ctx->launches++;

    a.nE = p->nE;
    a.nP = p->nP;
    a.nTs = (int)(t1 - t0);
    a.nDs = (int)(d1 - d0);
    a.t0 = (int)t0;
    a.d0 = (int)d0;
    // CTA size.  Store-dominated sweeps (many phi rows per column) run 64-thread CTAs: measured on
    // config-5 shards they reach 0.95-0.97 of the HBM peak where 256-thread CTAs reach 0.89-0.92 (shorter
    // tail, more CTAs to interleave), and they tie on the full cube.  Sweeps with few rows per column
    // (config 4: P = 1) are bound by the column setup and run 256-thread CTAs (1.55 ms against 1.96 ms).
    // Option "k2_nt" = 64 | 256 overrides (tuning).
    int nt = (4 * p->nP >= 64) ? NT_SMALL : NT_BIG;
    if (opt.k2_nt == NT_SMALL || opt.k2_nt == NT_BIG) nt = opt.k2_nt;
    a.n_tiles = (int)((p->nE + nt * EPT - 1) / (nt * EPT));
    a.n_seg = 1;
    a.ph = make_physics(p->phys);

    const int64_t cols = (int64_t)a.nTs * a.nDs;
    const int64_t per_col = 4 * p->nP * p->nE;            // doubles per column
    a.d_stride = layout ? layout[1] : per_col;
    a.t_stride = layout ? layout[0] : (int64_t)a.nDs * per_col;
    a.pol_stride = layout ? layout[2] : p->nP * p->nE;
    a.row_stride = layout ? layout[3] : p->nE;
    const bool even_layout = !layout || ((layout[0] | layout[1] | layout[2] | layout[3]) % 2 == 0);
    SSHG_REQUIRE(cols * a.n_tiles < (1LL << 31), "sshg_yield: too many tiles for one launch");
    SSHG_CUDA(cudaFuncSetAttribute(yield_grid_kernel<NT_BIG, CHUNK_BIG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)grid_smem(NT_BIG, CHUNK_BIG)));
    SSHG_CUDA(cudaFuncSetAttribute(yield_grid_kernel<NT_SMALL, CHUNK_SMALL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)grid_smem(NT_SMALL, CHUNK_SMALL)));
    SSHG_CUDA(cudaFuncSetAttribute(yield_grid_kernel<NT_BIG, CHUNK_BIG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)grid_smem(NT_BIG, CHUNK_BIG)));
    SSHG_CUDA(cudaFuncSetAttribute(yield_grid_kernel<NT_SMALL, CHUNK_SMALL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)grid_smem(NT_SMALL, CHUNK_SMALL)));


    // Sweeps with few azimuths per column go through the sparse-phi kernel (phi-contracted chi2,
    // thickness loop inside the thread).  option "k2_sparse" = 0 | 1 overrides (tests run both kernels).
    SparseArgs sa{};
    if (sparse) {
        void* dX;
        if (int rc = sshg_scratch(ctx, SCR_XPHI, (size_t)p->nP * NX * nE * 16, &dX)) return rc;
        const long long tot = (long long)p->nP * p->nE;
        phi_contract_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(a.chi, (const double*)dphi, (double2*)dX,
                                                                                  p->nE, p->nP, a.ph.zzz_phi_quirk);
        SSHG_CUDA(cudaGetLastError());
        ctx->launches++;
        sa.eps1w = a.eps1w; sa.eps2w = a.eps2w; sa.chi = a.chi; sa.X = (const double2*)dX;
        sa.energy = a.energy; sa.theta_deg = a.theta_deg; sa.thick_nm = a.thick_nm; sa.phi_deg = (const double*)dphi;
        sa.nE = p->nE; sa.nP = p->nP; sa.nDs = a.nDs; sa.t0 = a.t0; sa.d0 = a.d0;
        sa.n_tiles = (int)((p->nE + SNT - 1) / SNT);
        sa.t_stride = a.t_stride; sa.d_stride = a.d_stride; sa.pol_stride = a.pol_stride; sa.row_stride = a.row_stride;
        sa.ph = a.ph;
        // uniformly spaced thicknesses -> phase factors by recurrence (option "k2_recurrence" = 0 disables)
        void* dinfo;
        if (int rc = sshg_scratch(ctx, SCR_XPHI + 1, 16, &dinfo)) return rc;
        thick_uniform_kernel<<<1, 256, 0, ctx->stream>>>(a.thick_nm, p->nD, (double*)dinfo,
                                                       opt.k2_recurrence == 0 ? 0 : 1);
        SSHG_CUDA(cudaGetLastError());
        ctx->launches++;
        sa.dinfo = (const double*)dinfo;
    }

    BlurArgs ba{};
    size_t blur_smem = 0;
    // Tile width and shared-memory plan of K3.  A CTA holds wf[2 R + 2] + cols[20][S] + raw[13][S] (S = tile + 2 R) and
    // the phi table (12 doubles per item) behind them; what does not fit is staged in chunks, polarisation by
    // polarisation.  256-energy tiles (128 threads, register-limited to 2 CTAs/SM) halve the halo overhead of
    // 128-energy tiles and fit twice on an SM for every fused radius (R = 64: 101 KB + 11 KB of table), so they are
    // used whenever the axis is longer than one narrow tile.  Option "k3_nt" = 64 | 128 overrides (tests run both).
    int blur_nt = NT_BLUR, blur_per_sm = 1, blur_te = NT_BLUR * EPT;          // threads and energies per CTA
    if (fused) {
        const size_t sm_bytes = (size_t)228 * 1024, cta_max = (size_t)227 * 1024;
        // 256-energy tiles (128 threads, two energies each) whenever the axis is longer than one narrow tile.  (One energy
        // per thread -- 256 threads, 16 warps per SM instead of 8, 64-bit stores -- measured 12-18 % slower:
        // profiles/r02_k3_one_energy_per_thread.jsonl.)
        // Short launches (fewer than 4 waves of wide CTAs: config 3) take 128-energy tiles, FOUR 64-thread CTAs per SM:
        // half of a CTA's life is store-free (setup, harmonics, FIR, strict rows), and four independent CTAs interleave
        // those phases better than two -- config 3 with sigma_out = 5 0.488 -> 0.459 ms; long launches lose 3-4 % to the
        // doubled halo and stay wide (profiles/r02_k3_narrow4.jsonl).
        if (p->nE > NT_BLUR * EPT && !(k3_waves < 4.0 && radius <= 32)) blur_nt = NT_BLUR_WIDE;
        if (opt.k3_nt == NT_BLUR || opt.k3_nt == NT_BLUR_WIDE) blur_nt = opt.k3_nt;
        blur_te = blur_nt * EPT;
        auto fixed = [&]() { return (size_t)blur_fixed_doubles((int)radius, blur_te + 2 * (int)radius) * sizeof(double); };
        auto need = [&]() { return fixed() + (size_t)blur_chunk_doubles(32) * sizeof(double); };
        auto share = [&](int per_sm) { return sm_bytes / per_sm - 1024; };     // 1 KB per CTA is reserved by the system
        SSHG_REQUIRE(need() <= cta_max, "sshg_yield: internal: K3 shared-memory plan");
        const int reg_limit = blur_nt == NT_BLUR ? BLUR_NARROW_PER_SM : 2;                      // registers per thread (launch bounds)
        blur_per_sm = 1;
        for (int k = reg_limit; k >= 1; --k)
            if (need() <= share(k)) { blur_per_sm = k; break; }
        // grow the allocation to this occupancy's share, but not beyond what the whole table needs
        const int items_all = (int)p->nP;                                      // un-paired grid: one item per phi
        size_t total = fixed() + (size_t)blur_chunk_doubles(items_all) * sizeof(double);
        if (total > share(blur_per_sm)) total = share(blur_per_sm);
        if (total > cta_max) total = cta_max;
        int chunk = 32;                                                        // as many items as fit
        while (chunk < items_all && fixed() + (size_t)blur_chunk_doubles(chunk + 1) * sizeof(double) <= total) ++chunk;
        blur_smem = fixed() + (size_t)blur_chunk_doubles(chunk) * sizeof(double);
        ba.radius = (int)radius;
        ba.S = blur_te + 2 * (int)radius;
        ba.chunk = chunk;
        ba.fixup = opt.k3_fixup == 0 ? 0 : 1;
        ba.phi_deg = (const double*)dphi;
        sshg_gaussian_taps(sigma_out, radius, ba.w);
        const void* kfn = blur_nt == NT_BLUR ? (quads ? (const void*)yield_blur_kernel<NT_BLUR, true> : (const void*)yield_blur_kernel<NT_BLUR, false>)
                                             : (quads ? (const void*)yield_blur_kernel<NT_BLUR_WIDE, true> : (const void*)yield_blur_kernel<NT_BLUR_WIDE, false>);
        SSHG_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)blur_smem));
    }
    const int blur_tiles = (int)((p->nE + blur_te - 1) / blur_te);
    SSHG_REQUIRE(cols * blur_tiles < (1LL << 31), "sshg_yield: too many tiles for one launch");

    // Short launches: a CTA that is left alone at the end of a launch still needs its full azimuthal sweep
    // (~40-75 us), and every CTA starts with a store-free setup phase.  Cutting the phi items of a column
    // into segments (one CTA each, the setup is repeated) shortens both.  option "nseg" overrides (tuning).
    // K2 (tools/tune_nseg.py, profiles/r02_tune_nseg.jsonl): segments until the launch has about 4.5 waves, at most 4 --
    // config 3 (1.6 waves un-cut) 3 segments: 0.401 -> 0.365 ms (2: 0.374, 4: 0.382); an 8-theta slab of config 5 (1.4
    // waves) 4 segments: 0.672 -> 0.587 ms; a 23-theta shard (4.1 waves) 2: 1.682 -> 1.652 ms; 46 thetas (8.1 waves) none.
    auto pick_segments = [&](int64_t ctas, int per_sm, double target_waves) -> int {
        if (opt.nseg >= 1) return opt.nseg;
        const double waves = (double)ctas / ((double)ctx->num_sms * per_sm);
        const long long items = p->nP - p->nP / 2;
        int ns = 1;
        while (ns < SSHG_NSEG_MAX && waves * ns < target_waves && items / (ns + 1) >= 32) ++ns;
        return ns;
    };

    // un-broadened yields of columns [c0, c0 + nc) of the shard -> dst
    auto launch_raw = [&](int64_t c0, int64_t nc, double* dst, cudaStream_t st) -> int {
        if (sparse) {
            SparseArgs b = sa;
            b.col0 = c0;
            b.ncols = nc;
            b.out = dst;
            b.out_off = layout ? 0 : c0 * per_col;           // dst points at column c0
            // columns per CTA: enough CTAs for ~8 waves of 3 CTAs per SM, at most 32 columns each
            const long long want = 8LL * 3 * ctx->num_sms;
            long long chunk = nc * b.n_tiles / want;
            b.chunk = (int)(chunk < 1 ? 1 : (chunk > 32 ? 32 : chunk));
            const long long n_chunks = (nc + b.chunk - 1) / b.chunk;
            SSHG_REQUIRE(n_chunks * b.n_tiles < (1LL << 31), "sshg_yield: too many tiles for one launch");
            yield_sparse_kernel<<<(unsigned)(n_chunks * b.n_tiles), SNT, 0, st>>>(b);
            SSHG_CUDA(cudaGetLastError());
            ctx->launches++;
            return SSHG_OK;
        }
        YieldArgs b = a;
        b.col0 = c0;
        b.out = dst;
        b.out_off = layout ? 0 : c0 * per_col;               // dst points at column c0
        b.vec_ok = (p->nE % 2 == 0) && ((uintptr_t)dst % 16 == 0) && even_layout;
        b.n_seg = pick_segments(nc * a.n_tiles, nt == NT_BIG ? 2 : K2_SMALL_PER_SM, SSHG_NSEG_WAVES);
        SSHG_REQUIRE(nc * a.n_tiles * b.n_seg < (1LL << 31), "sshg_yield: too many tiles for one launch");
        const unsigned ctas = (unsigned)(nc * a.n_tiles * b.n_seg);
        // (Staging the warps' runs in shared memory and writing them with cp.async.bulk reaches the sequential-store
        // peak on this address pattern in isolation, but costs resident warps in the real kernel and measured 1.6-1.8x
        // slower: profiles/r02_store_experiments.txt, profiles/r02_k2_bulk_store_*.jsonl, DESIGN.md.)
        if (nt == NT_BIG && quads)
            yield_grid_kernel<NT_BIG, CHUNK_BIG, true><<<ctas, NT_BIG, grid_smem(NT_BIG, CHUNK_BIG), st>>>(b);
        else if (nt == NT_BIG)
            yield_grid_kernel<NT_BIG, CHUNK_BIG, false><<<ctas, NT_BIG, grid_smem(NT_BIG, CHUNK_BIG), st>>>(b);
        else if (quads)
            yield_grid_kernel<NT_SMALL, CHUNK_SMALL, true><<<ctas, NT_SMALL, grid_smem(NT_SMALL, CHUNK_SMALL), st>>>(b);
        else
            yield_grid_kernel<NT_SMALL, CHUNK_SMALL, false><<<ctas, NT_SMALL, grid_smem(NT_SMALL, CHUNK_SMALL), st>>>(b);
        SSHG_CUDA(cudaGetLastError());
        ctx->launches++;
        return SSHG_OK;
    };
    // K3: the same columns, broadened along the energy axis, in one pass (rows next to an azimuthal node are
    // re-evaluated the strict way inside the same CTA; option "k3_fixup" = 0 disables that: tests of the error bound).
    auto launch_blur = [&](int64_t c0, int64_t nc, double* dst, cudaStream_t st) -> int {
        BlurArgs b = ba;
        b.y = a;
        b.y.col0 = c0;
        b.y.out = dst;
        b.y.n_tiles = blur_tiles;
        b.y.vec_ok = (p->nE % 2 == 0) && ((uintptr_t)dst % 16 == 0);
        // K3's setup (halo, harmonics, FIR) is too heavy to repeat: cutting the phi items of a column into segments
        // measured slower, and so did cutting only the items of the last, partial wave of a short launch (config 3:
        // 0.495 against 0.501 ms, config-5 shards 1 % slower: profiles/r02_k3_quick_tail_split.jsonl).  Option "nseg"
        // cuts every item (tests).
        b.y.n_seg = opt.nseg >= 1 ? opt.nseg : 1;
        SSHG_REQUIRE(nc * blur_tiles * b.y.n_seg < (1LL << 31), "sshg_yield: too many tiles for one launch");
        const unsigned ctas = (unsigned)(nc * blur_tiles * b.y.n_seg);
        if (blur_nt == NT_BLUR && quads)
            yield_blur_kernel<NT_BLUR, true><<<ctas, NT_BLUR, blur_smem, st>>>(b);
        else if (blur_nt == NT_BLUR)
            yield_blur_kernel<NT_BLUR, false><<<ctas, NT_BLUR, blur_smem, st>>>(b);
        else if (quads)
            yield_blur_kernel<NT_BLUR_WIDE, true><<<ctas, NT_BLUR_WIDE, blur_smem, st>>>(b);
        else
            yield_blur_kernel<NT_BLUR_WIDE, false><<<ctas, NT_BLUR_WIDE, blur_smem, st>>>(b);
        SSHG_CUDA(cudaGetLastError());
        ctx->launches++;
        return SSHG_OK;
    };

    // slabs of columns: the host-output pipeline and the two-pass broadening work slab by slab
    // (256 MB for host output: slab k is copied back while slab k + 1 is computed; 1 GB when the output stays on the
    // device and the slab is only the intermediate of the two-pass path -- config 4, 288 MB, is then one launch of each
    // kernel).  The slabs are made equal: a short last slab would be a badly filled launch.
    const size_t slab_target = out_dev ? (size_t)1 << 30 : (size_t)256 << 20;
    int64_t cols_per_slab = (int64_t)(slab_target / ((size_t)per_col * 8));
    if (cols_per_slab < 1) cols_per_slab = 1;
    if (cols_per_slab > cols) cols_per_slab = cols;
    {
        const int64_t n_slabs = (cols + cols_per_slab - 1) / cols_per_slab;
        cols_per_slab = (cols + n_slabs - 1) / n_slabs;
    }
    const size_t slab_bytes = (size_t)cols_per_slab * per_col * 8;
    void* tmp = nullptr;                        // K2 output of a slab, input of K1 (two-pass only)
    if (two_pass) {
        if (int rc = sshg_scratch(ctx, 12, slab_bytes, &tmp)) return rc;
    }
    // yields of columns [c0, c0 + nc), broadened if asked, -> dst (device), queued on ctx->stream
    auto produce = [&](int64_t c0, int64_t nc, double* dst) -> int {
        if (fused) return launch_blur(c0, nc, dst, ctx->stream);
        if (!two_pass) return launch_raw(c0, nc, dst, ctx->stream);
        if (int rc = launch_raw(c0, nc, (double*)tmp, ctx->stream)) return rc;
        return sshg_gauss_device(ctx, (const double*)tmp, dst, nc * 4 * p->nP, p->nE, p->nE, 1, sigma_out);
    };

    if (out_dev) {
        if (!two_pass) return produce(0, cols, out);
        for (int64_t c0 = 0; c0 < cols; c0 += cols_per_slab) {
            const int64_t nc = (cols - c0 < cols_per_slab) ? cols - c0 : cols_per_slab;
            if (int rc = produce(c0, nc, out + c0 * per_col)) return rc;
        }
        return SSHG_OK;
    }

    // Host output: compute slabs of columns into two device buffers and copy slab k back on the
    // copy stream while slab k+1 is computed.
    void* buf[2];
    if (int rc = sshg_scratch(ctx, 8, slab_bytes, &buf[0])) return rc;
    {
        // A pageable destination (a fresh NumPy array) goes through the threaded staged copy: the driver's own pageable
        // path runs at 4 GB/s into memory that has never been touched (one thread takes the page faults).
        cudaPointerAttributes at;
        const bool pinned = cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeHost;
        (void)cudaGetLastError();
        if (!pinned) {
            for (int64_t c0 = 0; c0 < cols; c0 += cols_per_slab) {
                const int64_t nc = (cols - c0 < cols_per_slab) ? cols - c0 : cols_per_slab;
                if (int rc = produce(c0, nc, (double*)buf[0])) return rc;
                if (int rc = sshg_copy_out(ctx, out + c0 * per_col, buf[0], (size_t)nc * per_col * 8)) return rc;
            }
            return SSHG_OK;
        }
    }
    if (cols_per_slab < cols) {
        if (int rc = sshg_scratch(ctx, 9, slab_bytes, &buf[1])) return rc;
    } else {
        buf[1] = buf[0];
    }
    // The copy is the bound of a host-output sweep (PCIe), so rows that are known to be equal cross it once: on a grid
    // with (phi, phi + 180 deg) pairs the sS rows of a pair are the SAME BITS -- sS has odd harmonics only, the kernels
    // evaluate the pair once and store the value twice (sweep4 / sweep4_quads; K3: harm_item / harm_quad with SS; the
    // strict re-evaluation and K1 are functions of the row's un-broadened values, which are equal) -- so rows H .. 2 H - 1
    // of sS stay on the device and host threads duplicate rows 0 .. H - 1 while the next slab is on the wire: one eighth
    // less PCIe traffic.  The pairing test is the device's (phi_items_build), repeated here on the host array.
    int64_t Hdup = 0;
    if (!in_dev && !sparse && opt.pair_copy != 0 && p->nP >= 2) {
        const int64_t H = p->nP / 2;
        bool paired = true;
        for (int64_t j = 0; j < H && paired; ++j) paired = p->phi_deg[j + H] == p->phi_deg[j] + 180.0;
        if (paired) Hdup = H;
    }
    const size_t rowb = (size_t)p->nE * 8;                               // bytes per row
    const size_t head_b = (size_t)(3 * p->nP + Hdup) * rowb;             // pp, ps, sp and sS rows 0 .. H - 1
    const size_t tail_off = (size_t)(3 * p->nP + 2 * Hdup) * rowb, tail_b = (size_t)per_col * 8 - tail_off;
    int k = 0;
    bool used[2] = {false, false};
    int64_t prev_c0 = -1, prev_nc = 0;
    int prev_k = 0;
    auto finish_prev = [&]() -> int {          // slab prev has arrived: duplicate its equal rows (the next copy is in flight)
        if (prev_c0 < 0 || !Hdup) return SSHG_OK;
        SSHG_CUDA(cudaEventSynchronize(ctx->ev_copied[prev_k]));
        sshg_parallel_dup((char*)(out + prev_c0 * per_col), (size_t)per_col * 8, (size_t)(3 * p->nP) * rowb,
                          (size_t)(3 * p->nP + Hdup) * rowb, (size_t)Hdup * rowb, prev_nc);
        return SSHG_OK;
    };
    for (int64_t c0 = 0; c0 < cols; c0 += cols_per_slab, k ^= 1) {
        const int64_t nc = (cols - c0 < cols_per_slab) ? cols - c0 : cols_per_slab;
        if (used[k]) SSHG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[k], 0));   // buffer drained
        if (int rc = produce(c0, nc, (double*)buf[k])) return rc;
        SSHG_CUDA(cudaEventRecord(ctx->ev_done[k], ctx->stream));
        SSHG_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[k], 0));
        if (!Hdup) {
            SSHG_CUDA(cudaMemcpyAsync(out + c0 * per_col, buf[k], (size_t)nc * per_col * 8, cudaMemcpyDeviceToHost,
                                      ctx->copy_stream));
        } else {
            SSHG_CUDA(cudaMemcpy2DAsync(out + c0 * per_col, (size_t)per_col * 8, buf[k], (size_t)per_col * 8, head_b, (size_t)nc,
                                        cudaMemcpyDeviceToHost, ctx->copy_stream));
            if (tail_b)
                SSHG_CUDA(cudaMemcpy2DAsync((char*)(out + c0 * per_col) + tail_off, (size_t)per_col * 8, (const char*)buf[k] + tail_off,
                                            (size_t)per_col * 8, tail_b, (size_t)nc, cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
        SSHG_CUDA(cudaEventRecord(ctx->ev_copied[k], ctx->copy_stream));
        used[k] = true;
        if (int rc = finish_prev()) return rc;
        prev_c0 = c0;
        prev_nc = nc;
        prev_k = k;
    }
    if (int rc = finish_prev()) return rc;
    SSHG_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
    return SSHG_OK;
}

extern "C" int sshg_yield(sshg_ctx* ctx, const sshg_problem* p, double* out, int where) {
    return yield_impl(ctx, p, 0.0, out, where);
}

extern "C" int sshg_yield_broadened(sshg_ctx* ctx, const sshg_problem* p, double sigma_out, double* out, int where) {
    return yield_impl(ctx, p, sigma_out, out, where);
}

// The tail of shgyield() for LARGE NumPy broadcasts with outer-product structure (the reference's only way to express a
// sweep: energy[E], theta[T, 1], phi[P, 1, 1], ...; shg.py:381-437): the yields of the grid are written by K2 / K2s
// straight into the layout of the broadcast result ([pol][...result axes...], energy contiguous), broad(., sigma_out)
// runs over EVERY axis of each polarisation's array on the device (scipy.ndimage.gaussian_filter semantics, shg.py:26-28,
// 433-436), and the four arrays go to the host by the threaded staged copy.  No host-side transposes or re-uploads.
extern "C" int sshg_yield_broadcast(sshg_ctx* ctx, const sshg_problem* p, const int64_t strides[4], double sigma_out,
                                    int32_t ndim, const int64_t* shape, double* const out_host[4]) {
    SSHG_TRACE("sshg_yield_broadcast");
    SSHG_REQUIRE(ctx && p && strides && out_host && (shape || ndim == 0), "sshg_yield_broadcast: null argument");
    SSHG_REQUIRE(ndim >= 0 && ndim <= 15, "sshg_yield_broadcast: ndim must be in 0 .. 15 (got %d)", (int)ndim);
    SSHG_REQUIRE(sigma_out >= 0.0 && isfinite(sigma_out), "sshg_yield_broadcast: sigma_out must be finite and >= 0");
    SSHG_REQUIRE(p->nE > 0 && p->nT > 0 && p->nP > 0 && p->nD > 0, "sshg_yield_broadcast: empty grid");
    int64_t total = 1;
    for (int a = 0; a < ndim; ++a) {
        SSHG_REQUIRE(shape[a] > 0, "sshg_yield_broadcast: extents must be positive");
        total *= shape[a];
    }
    SSHG_REQUIRE(total == p->nE * p->nT * p->nP * p->nD, "sshg_yield_broadcast: shape does not match the grid");
    SSHG_REQUIRE(strides[2] == total, "sshg_yield_broadcast: the polarisation stride must be the size of one result array");
    for (int k = 0; k < 4; ++k) SSHG_REQUIRE(out_host[k], "sshg_yield_broadcast: null output array");
    // the largest element index the layout can reach must stay inside one polarisation's array
    SSHG_REQUIRE(strides[0] >= 0 && strides[1] >= 0 && strides[3] >= 0 &&
                     (p->nT - 1) * strides[0] + (p->nD - 1) * strides[1] + (p->nP - 1) * strides[3] + p->nE <= total,
                 "sshg_yield_broadcast: layout exceeds the result array");
    SSHG_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)total * 4 * 8;
    const bool blur = sigma_out > 1e-15 && ndim > 0;
    // intermediates up to 256 MB live in the context's scratch slots (no allocation per call); larger ones are
    // allocated for the call and released again
    void *a = nullptr, *b = nullptr;
    const bool own = bytes > ((size_t)256 << 20);
    auto release = [&]() {
        if (own && a) cudaFree(a);
        if (own && b) cudaFree(b);
    };
    if (!own) {
        if (int rc = sshg_scratch(ctx, SCR_BCAST, bytes, &a)) return rc;
        if (blur) {
            if (int rc = sshg_scratch(ctx, SCR_BCAST + 1, bytes, &b)) return rc;
        }
    } else if (cudaMalloc(&a, bytes) != cudaSuccess || (blur && cudaMalloc(&b, bytes) != cudaSuccess)) {
        (void)cudaGetLastError();
        release();
        sshg_set_error("sshg_yield_broadcast: device allocation of %zu bytes failed", bytes * (blur ? 2 : 1));
        return SSHG_ENOMEM;
    }
    int rc = yield_impl(ctx, p, 0.0, (double*)a, SSHG_HOST | SSHG_OUT_DEVICE, strides);
    const double* res = (const double*)a;
    if (!rc && blur) {
        // one pass per axis of the result, ping-pong between the two buffers (as sshg_gauss_nd_device, which would need a
        // third one); the four polarisations are the leading, un-filtered axis
        double *src = (double*)a, *dst = (double*)b;
        for (int ax = 0; ax < ndim && !rc; ++ax) {
            int64_t outer = 4, inner = 1;
            for (int j = 0; j < ax; ++j) outer *= shape[j];
            for (int j = ax + 1; j < ndim; ++j) inner *= shape[j];
            const int64_t n = shape[ax];
            if (inner == 1) rc = sshg_gauss_device(ctx, src, dst, outer, n, n, 1, sigma_out);
            else rc = sshg_gauss_device_ex(ctx, src, dst, outer * inner, n, 1, inner, inner, n * inner, sigma_out);
            double* t = src;
            src = dst;
            dst = t;
        }
        res = src;
    }
    for (int k = 0; k < 4 && !rc; ++k) rc = sshg_copy_out(ctx, out_host[k], res + (size_t)k * total, (size_t)total * 8);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !rc) rc = sshg_cuda_fail(cudaGetLastError(), "sync", __FILE__, __LINE__);
    release();
    return rc;
}

static int points_impl(sshg_ctx* ctx, const sshg_points* p, double* out, int where, bool amp) {
    SSHG_TRACE("sshg_yield_points");
    SSHG_REQUIRE(ctx && p && out, "sshg_yield_points: null argument");
    SSHG_REQUIRE(p->n > 0 && p->nE > 0, "sshg_yield_points: n and nE must be positive");
    SSHG_REQUIRE(p->energy_eV && p->eidx && p->theta_deg && p->phi_deg && p->thick_nm && p->eps1w && p->eps2w && p->chi2,
                 "sshg_yield_points: null input array");
    if (int rc = check_physics(p->phys)) return rc;
    SSHG_CUDA(cudaSetDevice(ctx->device));
    const int in_dev = (where & SSHG_IN_DEVICE) != 0, out_dev = (where & SSHG_OUT_DEVICE) != 0;
    const size_t nE = (size_t)p->nE, n = (size_t)p->n;
    if (!in_dev) {
        for (size_t i = 0; i < n; ++i)
            SSHG_REQUIRE(p->eidx[i] >= 0 && p->eidx[i] < p->nE, "sshg_yield_points: eidx[%zu] = %lld out of range", i,
                         (long long)p->eidx[i]);
    }
    PointArgs a{};
    const void* d;
    if (int rc = sshg_stage_in(ctx, 0, p->eps1w, 3 * nE * 16, in_dev, &d)) return rc;
    a.eps1w = (const double2*)d;
    if (int rc = sshg_stage_in(ctx, 1, p->eps2w, 3 * nE * 16, in_dev, &d)) return rc;
    a.eps2w = (const double2*)d;
    if (int rc = sshg_stage_in(ctx, 2, p->chi2, SSHG_NCHI * nE * 16, in_dev, &d)) return rc;
    a.chi = (const double2*)d;
    if (int rc = sshg_stage_in(ctx, 3, p->energy_eV, nE * 8, in_dev, &d)) return rc;
    a.energy = (const double*)d;
    if (int rc = sshg_stage_in(ctx, 4, p->theta_deg, n * 8, in_dev, &d)) return rc;
    a.theta_deg = (const double*)d;
    if (int rc = sshg_stage_in(ctx, 5, p->thick_nm, n * 8, in_dev, &d)) return rc;
    a.thick_nm = (const double*)d;
    if (int rc = sshg_stage_in(ctx, 6, p->phi_deg, n * 8, in_dev, &d)) return rc;
    a.phi_deg = (const double*)d;
    if (int rc = sshg_stage_in(ctx, 7, p->eidx, n * 8, in_dev, &d)) return rc;
    a.eidx = (const long long*)d;
    a.n = p->n;
    a.nE = p->nE;
    a.ph = make_physics(p->phys);
    a.ph.raw = amp ? 1 : 0;
    const size_t out_bytes = 4 * n * (amp ? 16 : 8);
    void* dout = out;
    if (!out_dev) {
        if (int rc = sshg_scratch(ctx, 8, out_bytes, &dout)) return rc;
    }
    a.out = (double*)dout;
    if (amp) yield_points_kernel<true><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(a);
    else yield_points_kernel<false><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(a);
    SSHG_CUDA(cudaGetLastError());
    ctx->launches++;
    if (!out_dev) {
        SSHG_CUDA(cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return SSHG_OK;
}

extern "C" int sshg_yield_points(sshg_ctx* ctx, const sshg_points* p, double* out, int where) {
    return points_impl(ctx, p, out, where, false);
}

extern "C" int sshg_amplitudes_points(sshg_ctx* ctx, const sshg_points* p, double* out, int where) {
    return points_impl(ctx, p, out, where, true);
}

// ---- the tail of shgyield() for small broadcasts: yields of a point list + broad(., sigma_out) along every axis --------
// (reference shg.py:428-436; the interactive call shapes of gui.py:286-330).  With resident spectra the call is one
// upload of the packed point arrays from page-locked memory, the yield kernel, one FIR pass per axis and one download,
// recorded as a CUDA graph when a (shape, sigma_out, physics, spectra) combination comes a second time and replayed after.
namespace {

int grow_pinned(sshg_ctx* ctx, void** buf, size_t* have, size_t bytes) {
    if (*have >= bytes) return SSHG_OK;
    if (*buf) {
        SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
        SSHG_CUDA(cudaFreeHost(*buf));
        *buf = nullptr;
        *have = 0;
    }
    size_t cap = bytes < 65536 ? 65536 : bytes + bytes / 2;
    cudaError_t e = cudaMallocHost(buf, cap);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        sshg_set_error("pinned host allocation of %zu bytes failed: %s", cap, cudaGetErrorString(e));
        return SSHG_ENOMEM;
    }
    *have = cap;
    ctx->epoch++;
    return SSHG_OK;
}

// the device part: everything between the upload of the point arrays and the download of the result
int small_call_enqueue(sshg_ctx* ctx, const PointArgs& a, double sigma_out, int ndim, const int64_t* shape, double* d_y,
                       double* d_out, int* n_launches) {
    const size_t n = (size_t)a.n;
    PointArgs k = a;
    const bool blur = sigma_out > 1e-15 && ndim > 0;
    k.out = blur ? d_y : d_out;
    yield_points_kernel<false><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(k);
    SSHG_CUDA(cudaGetLastError());
    ctx->launches++;
    *n_launches = 1;
    if (blur) {
        int64_t full[9];
        full[0] = 4;
        for (int j = 0; j < ndim; ++j) full[1 + j] = shape[j];
        const int64_t before = ctx->launches;
        if (int rc = sshg_gauss_nd_device(ctx, d_y, d_out, 1, ndim + 1, full, sigma_out)) return rc;
        *n_launches += (int)(ctx->launches - before);
    }
    return SSHG_OK;
}

}  // namespace

extern "C" int sshg_yield_points_broadened(sshg_ctx* ctx, const sshg_points* p, double sigma_out, int32_t ndim,
                                           const int64_t* shape, double* out, int where) {
    SSHG_TRACE("sshg_yield_points_broadened");
    SSHG_REQUIRE(ctx && p && out && (shape || ndim == 0), "sshg_yield_points_broadened: null argument");
    SSHG_REQUIRE(p->n > 0 && p->nE > 0, "sshg_yield_points_broadened: n and nE must be positive");
    SSHG_REQUIRE(ndim >= 0 && ndim <= 8, "sshg_yield_points_broadened: ndim must be in 0 .. 8 (got %d)", (int)ndim);
    SSHG_REQUIRE(sigma_out >= 0.0 && isfinite(sigma_out), "sshg_yield_points_broadened: sigma_out must be finite and >= 0 (got %g)",
                 sigma_out);
    SSHG_REQUIRE(where == SSHG_HOST || where == SSHG_SPECTRA_DEVICE || where == SSHG_DEVICE,
                 "sshg_yield_points_broadened: where must be SSHG_HOST, SSHG_SPECTRA_DEVICE or SSHG_DEVICE");
    int64_t prod = 1;
    for (int j = 0; j < ndim; ++j) {
        SSHG_REQUIRE(shape[j] > 0, "sshg_yield_points_broadened: extents must be positive");
        prod *= shape[j];
    }
    SSHG_REQUIRE(prod == p->n, "sshg_yield_points_broadened: prod(shape) = %lld but n = %lld", (long long)prod, (long long)p->n);
    SSHG_REQUIRE(p->energy_eV && p->eidx && p->theta_deg && p->phi_deg && p->thick_nm && p->eps1w && p->eps2w && p->chi2,
                 "sshg_yield_points_broadened: null input array");
    if (int rc = check_physics(p->phys)) return rc;
    SSHG_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)p->n, nE = (size_t)p->nE;
    const bool all_dev = where == SSHG_DEVICE, spectra_dev = all_dev || where == SSHG_SPECTRA_DEVICE;

    PointArgs a{};
    a.n = p->n;
    a.nE = p->nE;
    a.ph = make_physics(p->phys);
    const void* d;
    if (int rc = sshg_stage_in(ctx, 0, p->eps1w, 3 * nE * 16, spectra_dev, &d)) return rc;
    a.eps1w = (const double2*)d;
    if (int rc = sshg_stage_in(ctx, 1, p->eps2w, 3 * nE * 16, spectra_dev, &d)) return rc;
    a.eps2w = (const double2*)d;
    if (int rc = sshg_stage_in(ctx, 2, p->chi2, SSHG_NCHI * nE * 16, spectra_dev, &d)) return rc;
    a.chi = (const double2*)d;
    if (int rc = sshg_stage_in(ctx, 3, p->energy_eV, nE * 8, spectra_dev, &d)) return rc;
    a.energy = (const double*)d;
    SSHG_REQUIRE(((uintptr_t)a.eps1w | (uintptr_t)a.eps2w | (uintptr_t)a.chi) % 16 == 0,
                 "sshg_yield_points_broadened: complex device arrays must be 16-byte aligned");

    void *d_y = nullptr, *d_out = out;
    if (int rc = sshg_scratch(ctx, SCR_PTS_Y, 4 * n * 8, &d_y)) return rc;
    if (all_dev) {                              // everything resident: queue and return
        a.theta_deg = p->theta_deg; a.phi_deg = p->phi_deg; a.thick_nm = p->thick_nm; a.eidx = (const long long*)p->eidx;
        int nl;
        return small_call_enqueue(ctx, a, sigma_out, ndim, shape, (double*)d_y, out, &nl);
    }
    for (size_t i = 0; i < n; ++i)
        SSHG_REQUIRE(p->eidx[i] >= 0 && p->eidx[i] < p->nE, "sshg_yield_points_broadened: eidx[%zu] = %lld out of range", i,
                     (long long)p->eidx[i]);

    if (n > ((size_t)1 << 20)) {
        // a large broadcast: no page-locked staging of hundreds of megabytes, the arrays go up one by one
        if (int rc = sshg_stage_in(ctx, 4, p->theta_deg, n * 8, 0, &d)) return rc;
        a.theta_deg = (const double*)d;
        if (int rc = sshg_stage_in(ctx, 5, p->thick_nm, n * 8, 0, &d)) return rc;
        a.thick_nm = (const double*)d;
        if (int rc = sshg_stage_in(ctx, 6, p->phi_deg, n * 8, 0, &d)) return rc;
        a.phi_deg = (const double*)d;
        if (int rc = sshg_stage_in(ctx, 7, p->eidx, n * 8, 0, &d)) return rc;
        a.eidx = (const long long*)d;
        if (int rc = sshg_scratch(ctx, SCR_PTS_OUT, 4 * n * 8, &d_out)) return rc;
        int nl;
        if (int rc = small_call_enqueue(ctx, a, sigma_out, ndim, shape, (double*)d_y, (double*)d_out, &nl)) return rc;
        SSHG_CUDA(cudaMemcpyAsync(out, d_out, 4 * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
        return SSHG_OK;
    }

    // point arrays: theta | phi | thick | eidx packed in page-locked memory, ONE upload; result through page-locked memory
    void* d_in;
    if (int rc = sshg_scratch(ctx, SCR_PTS_IN, 4 * n * 8, &d_in)) return rc;
    if (int rc = sshg_scratch(ctx, SCR_PTS_OUT, 4 * n * 8, &d_out)) return rc;
    if (int rc = grow_pinned(ctx, &ctx->pin_in, &ctx->pin_in_bytes, 4 * n * 8)) return rc;
    if (int rc = grow_pinned(ctx, &ctx->pin_out, &ctx->pin_out_bytes, 4 * n * 8)) return rc;
    double* hin = (double*)ctx->pin_in;
    memcpy(hin, p->theta_deg, n * 8);
    memcpy(hin + n, p->phi_deg, n * 8);
    memcpy(hin + 2 * n, p->thick_nm, n * 8);
    memcpy(hin + 3 * n, p->eidx, n * 8);
    a.theta_deg = (const double*)d_in;
    a.phi_deg = (const double*)d_in + n;
    a.thick_nm = (const double*)d_in + 2 * n;
    a.eidx = (const long long*)((const double*)d_in + 3 * n);

    auto enqueue_all = [&](int* nl) -> int {
        SSHG_CUDA(cudaMemcpyAsync(d_in, ctx->pin_in, 4 * n * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (int rc = small_call_enqueue(ctx, a, sigma_out, ndim, shape, (double*)d_y, (double*)d_out, nl)) return rc;
        SSHG_CUDA(cudaMemcpyAsync(ctx->pin_out, d_out, 4 * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        return SSHG_OK;
    };

    // ---- replay, or run directly and record for the next call -------------------------------------------------
    const bool graphable = spectra_dev && ctx->opt.graphs != 0 && n <= (1u << 16) && ctx->stream == ctx->own_stream;
    sshg_small_key key;
    memset(&key, 0, sizeof(key));
    key.n = p->n; key.nE = p->nE; key.ndim = ndim;
    for (int j = 0; j < ndim; ++j) key.shape[j] = shape[j];
    key.sigma_out = sigma_out;
    key.phys = p->phys;
    key.energy = a.energy; key.eps1w = a.eps1w; key.eps2w = a.eps2w; key.chi2 = a.chi;
    key.epoch = ctx->epoch;
    sshg_small_graph* hit = nullptr;
    if (graphable)
        for (int k = 0; k < SSHG_NGRAPH; ++k)
            if (ctx->graphs[k].exec && memcmp(&ctx->graphs[k].key, &key, sizeof(key)) == 0) hit = &ctx->graphs[k];
    if (hit) {
        hit->used = ++ctx->graph_clock;
        SSHG_CUDA(cudaGraphLaunch(hit->exec, ctx->stream));
        ctx->launches += hit->launches;
        ctx->graph_replays++;
    } else {
        int nl = 0;
        if (int rc = enqueue_all(&nl)) return rc;
    }
    SSHG_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->pin_out, 4 * n * 8);
    if (hit || !graphable) return SSHG_OK;

    // Every buffer the sequence touches now exists at its final size and the taps are resident, so the same sequence can
    // be captured (capture executes nothing).  Recording costs more than a call, so it waits for the SECOND sighting of a
    // combination: a stream of ever-new combinations (a slider that changes the prepared spectra on every tick) never
    // pays for it.  A failure here only means "no graph": the result is out.
    if (key.epoch != ctx->epoch) return SSHG_OK;      // a buffer moved during the direct run; record next time
    {
        bool seen_before = false;
        for (int k = 0; k < SSHG_NGRAPH; ++k)
            if (ctx->seen_valid[k] && memcmp(&ctx->seen[k], &key, sizeof(key)) == 0) {
                seen_before = true;
                ctx->seen_valid[k] = 0;
            }
        if (!seen_before) {
            ctx->seen[ctx->seen_next] = key;
            ctx->seen_valid[ctx->seen_next] = 1;
            ctx->seen_next = (ctx->seen_next + 1) % SSHG_NGRAPH;
            return SSHG_OK;
        }
    }
    sshg_small_graph* slot = &ctx->graphs[0];
    for (int k = 0; k < SSHG_NGRAPH; ++k) {
        if (!ctx->graphs[k].exec) { slot = &ctx->graphs[k]; break; }
        if (ctx->graphs[k].used < slot->used) slot = &ctx->graphs[k];
    }
    if (slot->exec) {
        cudaGraphExecDestroy(slot->exec);
        slot->exec = nullptr;
    }
    const int64_t launches_before = ctx->launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
        int nl = 0;
        const int rc = enqueue_all(&nl);
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        if (rc == SSHG_OK && e == cudaSuccess && graph && key.epoch == ctx->epoch &&
            cudaGraphInstantiate(&slot->exec, graph, 0) == cudaSuccess) {
            slot->key = key;
            slot->used = ++ctx->graph_clock;
            slot->launches = nl;
        } else {
            slot->exec = nullptr;
        }
        if (graph) cudaGraphDestroy(graph);
    }
    (void)cudaGetLastError();
    ctx->launches = launches_before;                 // nothing ran during the capture
    return SSHG_OK;
}
