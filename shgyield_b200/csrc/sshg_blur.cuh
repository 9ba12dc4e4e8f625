// K3 device code: the yield cube broadened along the energy axis inside the yield kernel.
// Included by sshg_yield.cu inside its anonymous namespace (uses YieldArgs, EPT, NCOL, store2t, ldc and the
// physics of sshg_math.cuh); the host-side plan (tile width, shared-memory budget, fallbacks) is in
// sshg_yield.cu::yield_impl.
#pragma once

// ---- K3: the yield cube broadened along the energy axis (sigma_out > 0), in ONE pass --------------------
// The reference broadens every yield spectrum after the fact (broad(., sigma_out), shg.py:26-28, 433-436;
// sigma_out = 5 is the API default).  Done literally on a cube that is a second pass over HBM (read 8 B +
// write 8 B per point) and 2 radius + 1 FMAs per point -- 3.7x the cost of K2.  Here the broadening is
// moved in front of the azimuthal sweep instead: |r(phi)|^2 is a trigonometric polynomial of degree 6 in
// phi (sshg_math.cuh::harmonics7), the Gaussian FIR along E is linear, so it is applied to the 13 (sS: 7)
// harmonic spectra of a column -- 46 FIRs per (theta, d) column instead of 4 nP -- and each point is then
// a 13-term real sum.  Per point that is 7 FP64 instructions (sS: 3) against 10 (5) in K2, so the broadened
// cube is again bound by its 8 B/point store stream.
//
// A CTA of NT threads owns 2 NT consecutive energies of a column.
//   (A)  column setup for those energies plus a halo of `radius` on each side ('reflect' boundary of
//        scipy.ndimage at the ends of the axis), parked in shared memory: cols[20][S];
//   then, polarisation by polarisation:
//   (A') coefficients + harmonics of the staged energies from the parked columns: raw[13][S];
//   (B)  every thread filters the harmonic spectra for its two energies (one LDS.128 feeds 4 FMAs); the 2 x 13
//        results stay in registers;
//   (C)  the azimuthal sweep over the phi table (in shared memory; chunked only if it does not fit), phi / phi +
//        180 deg pairs sharing the even and odd harmonic sums, 128-bit streaming stores;
//   (D)  strict re-evaluation of the rows the sweep flagged (below).
// Accuracy: the sum of harmonics cancels where the yield has a node in phi, so its error is absolute --
// <= 5e-14 of the column's amplitude bound U -- instead of relative.  The sweep therefore records, branch-free,
// which (polarisation, row) of the tile produced a value below NODE_FRACTION * U (a bit mask per 16 items, OR-ed
// into a shared-memory bitmap), and the CTA itself re-evaluates those rows the reference's way: the un-broadened
// |r(phi)|^2 of the staged energies from the parked columns exactly as K2 computes them, then the Gaussian FIR
// along E.  Everything the harmonic sum cannot resolve to 1e-10 -- including exact symmetry zeros -- is thus
// computed row-wise; no second kernel, no flag memory in HBM, no repeated column setup.
constexpr int NT_BLUR = 64;                     // 128-energy tiles (short energy axes)
constexpr int NT_BLUR_WIDE = 128;               // 256-energy tiles (half the halo overhead), 2 CTAs/SM up to radius 64, 1 beyond
#ifndef SSHG_K3_NARROW_PER_SM
#define SSHG_K3_NARROW_PER_SM 4
#endif
constexpr int BLUR_NARROW_PER_SM = SSHG_K3_NARROW_PER_SM;   // resident 64-thread CTAs per SM the registers allow
constexpr int BLUR_RMAX = 200;                  // largest fused radius (sigma_out <= 50, the range of the reference GUI's slider)
constexpr int FX_RB = 12;                       // flagged rows re-evaluated together (rowbuf aliases raw: <= NH)
#ifndef SSHG_SWEEP_SB
#define SSHG_SWEEP_SB 16
#endif
constexpr int SWEEP_SB = SSHG_SWEEP_SB;         // items per flag mask (8 and 32 measured the same within 2 %)
static_assert(SWEEP_SB >= 1 && SWEEP_SB <= 32, "one 32-bit mask per block of items");

struct BlurArgs {
    YieldArgs y;                                // y.tab: harmonic table, 12 doubles per item
    const double* phi_deg;                      // [nP] (strict re-evaluation of flagged rows)
    int radius;
    int S;                                      // staged energies per CTA = 2 NT + 2 radius
    int chunk;                                  // phi items per table chunk
    int fixup;                                  // 1: re-evaluate the rows next to azimuthal nodes (0: tests of the bound)
    double w[BLUR_RMAX + 1];                    // Gaussian taps, w[0] centre
};

// shared memory of a CTA, in doubles (the host plans with the same function)
__host__ __device__ constexpr int blur_flag_words(int chunk) { return chunk / 32 + 2; }   // per part; a mask may straddle words
__host__ __device__ constexpr int blur_fixed_doubles(int radius, int S) {
    return 2 * radius + 2 + (NCOL + NH) * S + FX_RB * 6;
}
// a chunk of the phi table: 12 doubles per item, in front of it the bitmaps (flagged items; strict rows of the four row
// kinds of an item): 6 x words x 4 B reserved
__host__ __device__ constexpr int blur_chunk_doubles(int chunk) { return chunk * 12 + 3 * blur_flag_words(chunk) + (blur_flag_words(chunk) & 1); }

__global__ void phi_harm_table_kernel(const double* __restrict__ phi, long long nP, double* __restrict__ tab,
                                      int allow_pairs, int allow_quads) {
    const PhiItems it = phi_items_build(phi, nP, tab, allow_pairs, allow_quads);
    const long long items = it.total();
    for (long long i = threadIdx.x; i < items; i += blockDim.x) {
        double t[12];
        phi_harmonics(phi[phi_item_row(it, i)], t);
#pragma unroll
        for (int k = 0; k < 12; ++k) tab[TAB_HDR + i * 12 + k] = t[k];
    }
}

__device__ __forceinline__ long long reflect_index(long long i, long long n) {
    if (i >= 0 && i < n) return i;
    long long m = i % (2 * n);
    if (m < 0) m += 2 * n;
    return m >= n ? 2 * n - 1 - m : m;
}

// (B) Gaussian FIR of NHH harmonic spectra for two adjacent energies: x points at this thread's window
// x[0 .. 2 R + 1] of the first spectrum (spectra are S doubles apart), wf[0 .. 2 R] are the full taps.
// out0 = sum_t wf[t] x[t], out1 = sum_t wf[t] x[t + 1].  Absent taps at the two ends are skipped, not
// multiplied by zero (a NaN sample must not spread beyond the radius).
template <int NHH>
__device__ __forceinline__ void fir_pair(const double* __restrict__ x, int S, int R, const double* __restrict__ wf,
                                         double* A0, double* A1) {
    {
        const double w0 = wf[0], w1 = wf[1];
#pragma unroll
        for (int m = 0; m < NHH; ++m) {
            const double2 v = *reinterpret_cast<const double2*>(x + m * S);
            A0[m] = fma(v.y, w1, v.x * w0);
            A1[m] = v.y * w0;
        }
    }
#pragma unroll 1
    for (int j = 1; j < R; ++j) {
        const double wa = wf[2 * j - 1], wb = wf[2 * j], wc = wf[2 * j + 1];
#pragma unroll
        for (int m = 0; m < NHH; ++m) {
            const double2 v = *reinterpret_cast<const double2*>(x + m * S + 2 * j);
            A0[m] = fma(v.y, wc, fma(v.x, wb, A0[m]));
            A1[m] = fma(v.y, wb, fma(v.x, wa, A1[m]));
        }
    }
    {
        const double wa = wf[2 * R - 1], wb = wf[2 * R];
#pragma unroll
        for (int m = 0; m < NHH; ++m) {
            const double2 v = *reinterpret_cast<const double2*>(x + m * S + 2 * R);
            A0[m] = fma(v.x, wb, A0[m]);
            A1[m] = fma(v.y, wb, fma(v.x, wa, A1[m]));
        }
    }
}

// One azimuth item of the harmonic form for two energies: even and odd harmonic sums.
// tb: {cos, sin of 2, 4, 6 phi | cos, sin of 1, 3, 5 phi}.
template <bool SS>
__device__ __forceinline__ void harm_item(const double* A0, const double* A1, const double* __restrict__ tb, double& ye0,
                                          double& ye1, double& yo0, double& yo1) {
    double t[12];
#pragma unroll
    for (int j = 0; j < (SS ? 3 : 6); ++j) {
        const double2 v = *reinterpret_cast<const double2*>(tb + 2 * j);
        t[2 * j] = v.x;
        t[2 * j + 1] = v.y;
    }
    ye0 = A0[0];
    ye1 = A1[0];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        ye0 = fma(A0[1 + j], t[j], ye0);
        ye1 = fma(A1[1 + j], t[j], ye1);
    }
    yo0 = 0.0;
    yo1 = 0.0;
    if (!SS) {
        yo0 = A0[7] * t[6];
        yo1 = A1[7] * t[6];
#pragma unroll
        for (int j = 1; j < 6; ++j) {
            yo0 = fma(A0[7 + j], t[6 + j], yo0);
            yo1 = fma(A1[7 + j], t[6 + j], yo1);
        }
    }
}

// One QUAD item of the harmonic form for two energies: the four rows phi (a), 180 - phi (b), phi + 180 (c), 360 - phi (d)
// from four partial sums -- even / odd harmonics x cosine / sine terms: 20 FP64 instructions per energy for four points
// (sS, even harmonics only: 8).  tb as in harm_item.
template <bool SS>
__device__ __forceinline__ void harm_quad(const double* A, const double* __restrict__ tb, double& ya, double& yb, double& yc,
                                          double& yd) {
    double t[12];
#pragma unroll
    for (int j = 0; j < (SS ? 3 : 6); ++j) {
        const double2 v = *reinterpret_cast<const double2*>(tb + 2 * j);
        t[2 * j] = v.x;
        t[2 * j + 1] = v.y;
    }
    const double ec = fma(A[5], t[4], fma(A[3], t[2], fma(A[1], t[0], A[0])));       // C0 + C2 c2 + C4 c4 + C6 c6
    const double es = fma(A[6], t[5], fma(A[4], t[3], A[2] * t[1]));                 // S2 s2 + S4 s4 + S6 s6
    const double p = ec + es, m = ec - es;
    if (SS) {
        ya = p; yc = p; yb = m; yd = m;
        return;
    }
    const double oc = fma(A[11], t[10], fma(A[9], t[8], A[7] * t[6]));              // C1 c1 + C3 c3 + C5 c5
    const double os = fma(A[12], t[11], fma(A[10], t[9], A[8] * t[7]));             // S1 s1 + S3 s3 + S5 s5
    const double u = oc + os, v = oc - os;
    ya = p + u;                                 // phi
    yc = p - u;                                 // phi + 180: odd harmonics flip
    yd = m + v;                                 // 360 - phi: sine terms flip
    yb = m - v;                                 // 180 - phi: both
}

// mask bit k = item (it + k) of the chunk; rare path of the sweep
__device__ __forceinline__ void or_flag_bits(unsigned* fl, int it, unsigned m) {
    if (!m) return;
    const int sh = it & 31;
    atomicOr(fl + (it >> 5), m << sh);
    if (sh + SWEEP_SB > 32) atomicOr(fl + (it >> 5) + 1, m >> (32 - sh));
}

// (C) azimuthal sweep of the harmonic form, two energies per thread.  tb: the chunk's items of 12 doubles; items
// below n_pair give the rows i and i + H, the others the un-paired rows.
//
// Node test.  The hot loop is branch-free: besides the stores it compares the HIGH WORDS of the values it produced
// (the high word of a double orders like the double; a negative value is below everything, a positive NaN above)
// with the high word of NODE_FRACTION * U of its two energies and collects the outcome of an item in one bit of a
// mask -- integer minima, one compare and a select, nothing on the FP64 pipe.  After every SWEEP_SB items a non-zero
// mask (rare) is OR-ed into the CTA's bitmap of flagged ITEMS (bit i = item i of the chunk: its row, or its two rows).
// One part of a chunk: PAIR items give the rows i (q0) and i + H (q1), the others one un-paired row (q0).
// One flag bit per ITEM: the minimum of the high words of its (up to four) values against one threshold h -- the
// larger of the thread's two energies; phase (D) sorts out which row and which energy it was.
template <bool SS, bool VEC, bool PAIR>
__device__ __forceinline__ void sweep_part(const double* A0, const double* A1, const double* __restrict__ t, int n, int ibase,
                                           double* q0, double* q1, long long stride, bool valid1, int h, unsigned* fl) {
#pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += SWEEP_SB) {
        const int nb = min(SWEEP_SB, n - i0);
        unsigned mask = 0u, bit = 1u;
#pragma unroll 2
        for (int i = 0; i < nb; ++i) {
            double ye0, ye1, yo0, yo1;
            harm_item<SS>(A0, A1, t, ye0, ye1, yo0, yo1);
            t += 12;
            const double a0 = ye0 + yo0, a1 = ye1 + yo1;
            store2t<VEC>(q0, a0, a1, valid1);
            q0 += stride;
            int m = min(__double2hiint(a0), __double2hiint(a1));
            if (PAIR) {
                const double s0 = ye0 - yo0, s1 = ye1 - yo1;
                store2t<VEC>(q1, s0, s1, valid1);
                q1 += stride;
                if (!SS) m = min(m, min(__double2hiint(s0), __double2hiint(s1)));    // sS: the two rows are equal
            }
            if (m < h) mask |= bit;
            bit += bit;
        }
        if (mask) or_flag_bits(fl, ibase + i0, mask);       // rare: a value of this block is next to a node
    }
}

// The quad items of a chunk: rows ascend from pa (phi) and pc (phi + 180) and descend from pb (180 - phi) and pd (360 - phi).
template <bool SS, bool VEC>
__device__ __forceinline__ void sweep_quads(const double* A0, const double* A1, const double* __restrict__ t, int n, double* pa,
                                            double* pb, double* pc, double* pd, long long stride, bool valid1, int h,
                                            unsigned* fl) {
#pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += SWEEP_SB) {
        const int nb = min(SWEEP_SB, n - i0);
        unsigned mask = 0u, bit = 1u;
#pragma unroll 2
        for (int i = 0; i < nb; ++i) {
            double a0, b0, c0, d0, a1, b1, c1, d1;
            harm_quad<SS>(A0, t, a0, b0, c0, d0);
            harm_quad<SS>(A1, t, a1, b1, c1, d1);
            t += 12;
            store2t<VEC>(pa, a0, a1, valid1);
            store2t<VEC>(pb, b0, b1, valid1);
            store2t<VEC>(pc, c0, c1, valid1);
            store2t<VEC>(pd, d0, d1, valid1);
            pa += stride;
            pc += stride;
            pb -= stride;
            pd -= stride;
            int m = min(min(__double2hiint(a0), __double2hiint(a1)), min(__double2hiint(b0), __double2hiint(b1)));
            if (!SS) m = min(m, min(min(__double2hiint(c0), __double2hiint(c1)), min(__double2hiint(d0), __double2hiint(d1))));
            if (m < h) mask |= bit;
            bit += bit;
        }
        if (mask) or_flag_bits(fl, i0, mask);
    }
}

// A chunk of the item list: nq quads (first quad: rows q0 + 1, ...), np pair items (first: pair item p0), ns un-paired rows
// (first: row 2 H + s0); prow = this thread's energies in row 0 of the polarisation.
template <bool SS, bool VEC, bool QUADS>
__device__ __forceinline__ void sweep_harm(const double* A0, const double* A1, const double* __restrict__ tb, const PhiItems& it,
                                           int nq, long long q0, int np, long long p0, int ns, long long s0, double* prow,
                                           long long nE, bool valid1, int h, unsigned* fl) {
    if (QUADS)
        sweep_quads<SS, VEC>(A0, A1, tb, nq, prow + (q0 + 1) * nE, prow + (it.H - q0 - 1) * nE, prow + (q0 + 1 + it.H) * nE,
                             prow + (2 * it.H - q0 - 1) * nE, nE, valid1, h, fl);
    const long long pj = QUADS ? p0 * it.PSTEP : p0, pstride = QUADS ? it.PSTEP * nE : nE;
    sweep_part<SS, VEC, true>(A0, A1, tb + 12 * nq, np, nq, prow + pj * nE, prow + (pj + it.H) * nE, pstride, valid1, h, fl);
    sweep_part<SS, VEC, false>(A0, A1, tb + 12 * (nq + np), ns, nq + np, prow + (2 * it.H + s0) * nE, nullptr, nE, valid1, h, fl);
}

// the 7 (sS: 4) complex F-basis coefficients of one polarisation (shg.py:218-269, 290-313, 334-351, 372-377)
template <typename ChiFn>
__device__ __forceinline__ void coeffs_of(int pol, const Column& c, ChiFn chi, const Physics& ph, cplx kk[7]) {
    if (pol == 0) coeffs_pp(c, chi, ph, kk);
    else if (pol == 1) coeffs_ps(c, chi, kk);
    else if (pol == 2) coeffs_sp(c, chi, kk);
    else coeffs_ss(c, chi, kk);
}

__device__ __forceinline__ Column parked_column(const double* cols, int S, int i) {
    auto f = [&](int j) { return mk(cols[(2 * j) * S + i], cols[(2 * j + 1) * S + i]); };
    Column c;
    c.Gpp = f(0); c.Gps = f(1); c.Gsp = f(2); c.Gss = f(3);
    c.A = f(4); c.Bz = f(5); c.Q = f(6);
    c.a2 = f(7); c.ab = f(8); c.b2 = f(9);
    return c;
}

// shared memory: wf[2 R + 2] | cols[NCOL][S] | raw[NH][S] (rowbuf[FX_RB][S] during (D)) | pb[FX_RB][6] |
//                bitmaps fl[2][chunk / 32 + 1], st[2][chunk / 32 + 1] | tabs[chunk][12]
// QUADS: the item list may contain four-row items (instantiated apart: the pair form keeps its registers and schedule)
template <int NT, bool QUADS>
__global__ void __launch_bounds__(NT, (NT <= 64 ? BLUR_NARROW_PER_SM : 2)) yield_blur_kernel(const __grid_constant__ BlurArgs b) {
    extern __shared__ double smem[];
    const YieldArgs& a = b.y;
    const int R = b.radius, S = b.S, CH = b.chunk;
    const int FLW = blur_flag_words(CH);
    double* wf = smem;                          // [2 R + 2] full taps (padded to an even count)
    double* cols = wf + 2 * R + 2;              // [NCOL][S] parked columns, struct of arrays
    double* raw = cols + NCOL * S;              // [NH][S] harmonic spectra of the current polarisation
    double* pb = raw + NH * S;                  // [FX_RB][6] phi basis of the rows being re-evaluated
    unsigned* fl = reinterpret_cast<unsigned*>(pb + FX_RB * 6);      // [2][FLW] flagged rows, then [2][FLW] strict rows
    double* tabs = smem + blur_fixed_doubles(R, S) + (blur_chunk_doubles(CH) - CH * 12);     // [CH][12]

    const long long blk = blockIdx.x;
    const int tile = (int)(blk % a.n_tiles);
    const int n_seg = a.n_seg;
    const int seg = (int)((blk / a.n_tiles) % n_seg);
    const long long lcol = blk / a.n_tiles / n_seg;
    const long long col = a.col0 + lcol;
    const int tl = (int)(col / a.nDs), dl = (int)(col % a.nDs);
    const int tid = threadIdx.x;
    const long long nE = a.nE;
    const long long tile_e0 = (long long)tile * NT * EPT;

    for (int t = tid; t <= 2 * R; t += NT) wf[t] = b.w[t < R ? R - t : t - R];

    // ---- (A) column setup of the staged energies (tile + halo), parked -----------------------------------
    {
        double st, ct;
        sincos(a.theta_deg[a.t0 + tl] * (3.141592653589793 / 180.0), &st, &ct);
        const double thick = a.thick_nm[a.d0 + dl];
#pragma unroll 1
        for (int i = tid; i < S; i += NT) {
            const long long g = tile_e0 - R + i;
            if (g >= nE + R) break;             // no output of this axis reaches that far
            const long long e = reflect_index(g, nE);
            cplx e1[3], e2[3];
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                e1[m] = ldc(a.eps1w + m * nE + e);
                e2[m] = ldc(a.eps2w + m * nE + e);
            }
            const Column c = column_setup(e1, e2, a.energy[e], st, ct, thick, a.ph);
            const cplx v[10] = {c.Gpp, c.Gps, c.Gsp, c.Gss, c.A, c.Bz, c.Q, c.a2, c.ab, c.b2};
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                cols[(2 * j) * S + i] = v[j].re;
                cols[(2 * j + 1) * S + i] = v[j].im;
            }
        }
    }

    const long long e0 = tile_e0 + (long long)tid * EPT;
    const bool active = e0 < nE;
    const bool valid1 = e0 + 1 < nE;
    const PhiItems it = read_phi_items(a.tab, a.nP);
    const long long items = it.total();
    const bool vec = a.vec_ok != 0;
    const long long rows_per_pol = a.nP * nE;
    double* out_col = a.out + lcol * 4 * rows_per_pol + e0;
    const long long seg_len = (items + n_seg - 1) / n_seg;
    const long long seg_lo = seg * seg_len, seg_hi = min(items, (seg + 1) * seg_len);
    const bool one_chunk = seg_hi - seg_lo <= CH;               // the whole table of this CTA fits: staged once
    const int INT_LOWEST = (int)0x80000000;
    unsigned* st = fl + FLW;                    // [4][FLW] strict rows of the row kinds a, b, c, d

#pragma unroll 1
    for (int pol = 0; pol < 4; ++pol) {
        // ---- (A') harmonic spectra of this polarisation for the staged energies ----------------------------
        __syncthreads();                        // columns parked (pol 0); raw / rowbuf of the previous polarisation dead
#pragma unroll 1
        for (int i = tid; i < S; i += NT) {
            const long long g = tile_e0 - R + i;
            if (g >= nE + R) break;
            const long long e = reflect_index(g, nE);
            const Column c = parked_column(cols, S, i);
            auto chi = [&](int r) { return ldc(a.chi + r * nE + e); };
            cplx kk[7];
            coeffs_of(pol, c, chi, a.ph, kk);
            double Hh[NH];
            if (pol == 3) {
                harmonics4(kk, Hh);
#pragma unroll
                for (int m = 0; m < NH_SS; ++m) raw[m * S + i] = Hh[m];
            } else {
                harmonics7(kk, Hh);
#pragma unroll
                for (int m = 0; m < NH; ++m) raw[m * S + i] = Hh[m];
            }
        }
        __syncthreads();

        // ---- (B) broadened harmonics of this thread's two energies, in registers ----------------------------
        double A0[NH], A1[NH];
        {
            const double* x = raw + tid * EPT;
            if (pol == 3) fir_pair<NH_SS>(x, S, R, wf, A0, A1);
            else fir_pair<NH>(x, S, R, wf, A0, A1);
        }
        int h = INT_LOWEST;                     // nothing is ever below: no flags
        if (b.fixup && active) {
            const double u0 = amplitude_bound(A0, pol == 3), u1 = valid1 ? amplitude_bound(A1, pol == 3) : 0.0;
            const double u = fmax(u0, u1);      // one threshold for both energies; (D) tests each against its own
            if (u < 1.0e300) h = __double2hiint(NODE_FRACTION * u);         // a NaN / Inf column flags nothing
        }

        // ---- (C) azimuthal sweep, (D) strict re-evaluation of the flagged rows, table chunk by chunk --------
#pragma unroll 1
        for (long long base = seg_lo; base < seg_hi; base += CH) {
            const int n_items = (int)min((long long)CH, seg_hi - base);
            if (!one_chunk || pol == 0)         // (every thread is past the previous chunk: barrier at the end of the loop)
                for (int i = tid; i < n_items * 12; i += NT) tabs[i] = a.tab[TAB_HDR + base * 12 + i];
            for (int i = tid; i < 5 * FLW; i += NT) fl[i] = 0u;     // flagged items and (D)'s strict rows
            __syncthreads();
            // the kinds of item in [base, base + n_items): quads, then pair items, then un-paired rows
            const long long q0 = QUADS ? min(base, it.Q) : 0, q1 = QUADS ? min(base + n_items, it.Q) : 0;
            const long long p0 = min(max(base - it.Q, 0LL), it.NPI), p1 = min(max(base + n_items - it.Q, 0LL), it.NPI);
            const long long s0 = max(base - it.Q - it.NPI, 0LL);
            const int nq = (int)(q1 - q0), np = (int)(p1 - p0), ns = n_items - nq - np;
            double* prow = out_col + pol * rows_per_pol;
            if (active) {
#define SSHG_SWEEP(SS_, VEC_) \
    sweep_harm<SS_, VEC_, QUADS>(A0, A1, tabs, it, nq, q0, np, p0, ns, s0, prow, nE, valid1, h, fl)
                if (vec) {
                    if (pol == 3) SSHG_SWEEP(true, true);
                    else SSHG_SWEEP(false, true);
                } else {
                    if (pol == 3) SSHG_SWEEP(true, false);
                    else SSHG_SWEEP(false, false);
                }
#undef SSHG_SWEEP
            }
            __syncthreads();                    // bitmap complete; nobody reads raw or this chunk's table any more

            // ---- (D) flagged items.  The bitmaps are the same for every thread: uniform control flow.
            // First every thread looks at its own values of each flagged item again (the arithmetic of the sweep: the
            // same bits): a value in the band (DEEP, NODE) * U marks its row in the bitmaps `st`; negative rounding
            // noise (exact symmetry zeros) is stored again as 0.  Rows marked in `st` are then re-evaluated the
            // reference's way, FX_RB rows at a time; rows that were flagged only by exact zeros are done.
            if (!b.fixup) continue;
            // row of kind k (0: phi, 1: 180 - phi, 2: phi + 180, 3: 360 - phi) of chunk item `item`
            auto item_row = [&](int item, int k) -> long long {
                if (QUADS && item < nq) {
                    const long long j = q0 + item + 1;
                    return k == 0 ? j : k == 1 ? it.H - j : k == 2 ? j + it.H : 2 * it.H - j;
                }
                if (item < nq + np) return (p0 + item - nq) * (QUADS ? it.PSTEP : 1) + (k == 2 ? it.H : 0);
                return 2 * it.H + s0 + (item - nq - np);
            };
            {
                double thr0 = 0.0, thr1 = 0.0;  // NODE_FRACTION * U of this thread's energies (computed on first use)
                bool have_thr = false;
#pragma unroll 1
                for (int wi = 0; wi < FLW; ++wi) {
                    unsigned bits = fl[wi];
                    unsigned mine[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
                    while (bits) {
                        const int bp = __ffs(bits) - 1;
                        bits &= bits - 1;
                        if (!active) continue;
                        const int item = wi * 32 + bp;
                        if (!have_thr) {
                            thr0 = NODE_FRACTION * amplitude_bound(A0, pol == 3);
                            thr1 = valid1 ? NODE_FRACTION * amplitude_bound(A1, pol == 3) : 0.0;
                            have_thr = true;
                        }
                        double y0[4], y1[4];    // this thread's values of the item's rows a, b, c, d
                        const double* tb = tabs + 12 * item;
                        unsigned kinds;         // which of the four exist for this item
                        if (QUADS && item < nq) {
                            if (pol == 3) { harm_quad<true>(A0, tb, y0[0], y0[1], y0[2], y0[3]); harm_quad<true>(A1, tb, y1[0], y1[1], y1[2], y1[3]); }
                            else { harm_quad<false>(A0, tb, y0[0], y0[1], y0[2], y0[3]); harm_quad<false>(A1, tb, y1[0], y1[1], y1[2], y1[3]); }
                            kinds = 0xFu;
                        } else {
                            double ye0, ye1, yo0, yo1;
                            if (pol == 3) harm_item<true>(A0, A1, tb, ye0, ye1, yo0, yo1);
                            else harm_item<false>(A0, A1, tb, ye0, ye1, yo0, yo1);
                            y0[0] = ye0 + yo0; y1[0] = ye1 + yo1;
                            y0[2] = ye0 - yo0; y1[2] = ye1 - yo1;
                            y0[1] = y0[3] = y1[1] = y1[3] = 0.0;
                            kinds = item < nq + np ? 0x5u : 0x1u;
                        }
                        const double band = DEEP_FRACTION / NODE_FRACTION;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (!((kinds >> k) & 1u)) continue;
                            const double v0 = y0[k], v1 = y1[k];
                            if (((v0 < thr0) & (v0 > band * thr0)) | (valid1 & (v1 < thr1) & (v1 > band * thr1))) {
                                mine[k] |= 1u << bp;
                            } else if ((v0 < 0.0) | (valid1 & (v1 < 0.0))) {
                                double* dst = prow + item_row(item, k) * nE;
                                if (vec) store2t<true>(dst, fmax(v0, 0.0), fmax(v1, 0.0), valid1);
                                else store2t<false>(dst, fmax(v0, 0.0), fmax(v1, 0.0), valid1);
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (mine[k]) atomicOr(st + k * FLW + wi, mine[k]);
                }
            }
            __syncthreads();                    // `st` complete
            int rows[FX_RB], nb = 0;
#pragma unroll 1
            for (int wi = 0; wi <= 4 * FLW; ++wi) {
                unsigned bits = wi < 4 * FLW ? st[wi] : 0u;
                const bool last = wi == 4 * FLW;
#pragma unroll 1
                while (bits || (last && nb > 0)) {
                    if (bits) {
                        const int bp = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const int kind = wi / FLW;
                        const int item = (wi - kind * FLW) * 32 + bp;
                        rows[nb] = (int)item_row(item, kind);
                        if (tid == nb) {        // basis row of this azimuth from the staged table: cos / sin of phi are its
                            const double* t = tabs + 12 * item;          // first odd harmonic; the row kind sets their signs
                            const double c = (kind == 1 || kind == 2) ? -t[6] : t[6], sn = (kind == 2 || kind == 3) ? -t[7] : t[7];
                            double* bb = pb + nb * 6;
                            bb[0] = c * c;
                            bb[1] = sn * c;
                            bb[2] = bb[0] * c;
                            bb[3] = bb[1] * c;
                            bb[4] = bb[1] * sn;
                            bb[5] = sn * sn * sn;
                        }
                        ++nb;
                        if (nb < FX_RB) continue;
                    }
                    // -- one batch: the un-broadened yields of the staged energies as K2 computes them ...
                    double* rowbuf = raw;       // [FX_RB][S]
                    __syncthreads();
#pragma unroll 1
                    for (int i = tid; i < S; i += NT) {
                        const long long g = tile_e0 - R + i;
                        if (g >= nE + R) break;
                        const long long e = reflect_index(g, nE);
                        const Column c = parked_column(cols, S, i);
                        auto chi = [&](int r) { return ldc(a.chi + r * nE + e); };
                        cplx kk[7];
                        coeffs_of(pol, c, chi, a.ph, kk);
#pragma unroll 1
                        for (int q = 0; q < nb; ++q) {
                            const double* bb = pb + q * 6;
                            double re, im;
                            if (pol == 3) {     // k[0..3] multiply {c^3, s c^2, s^2 c, s^3}
                                re = kk[0].re * bb[2];
                                im = kk[0].im * bb[2];
#pragma unroll
                                for (int j = 1; j < 4; ++j) {
                                    re = fma(kk[j].re, bb[j + 2], re);
                                    im = fma(kk[j].im, bb[j + 2], im);
                                }
                            } else {
                                re = kk[0].re;
                                im = kk[0].im;
#pragma unroll
                                for (int j = 0; j < 6; ++j) {
                                    re = fma(kk[j + 1].re, bb[j], re);
                                    im = fma(kk[j + 1].im, bb[j], im);
                                }
                            }
                            rowbuf[q * S + i] = fma(re, re, im * im);
                        }
                    }
                    __syncthreads();
                    // -- ... filtered along E (four rows at a time: one LDS.128 feeds 4 FMAs) and stored
                    if (active) {
#pragma unroll 1
                        for (int r0 = 0; r0 < nb; r0 += 4) {
                            double o0[4], o1[4];
                            fir_pair<4>(rowbuf + r0 * S + tid * EPT, S, R, wf, o0, o1);
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (r0 + q < nb) {
                                    double* dst = prow + (long long)rows[r0 + q] * nE;
                                    if (vec) store2t<true>(dst, o0[q], o1[q], valid1);
                                    else store2t<false>(dst, o0[q], o1[q], valid1);
                                }
                        }
                    }
                    __syncthreads();            // rowbuf and pb are free again
                    nb = 0;
                }
            }
            __syncthreads();                    // every thread has read the bitmaps before the next chunk clears them
        }
    }
}
