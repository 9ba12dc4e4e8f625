// K3 device code: the yield cube broadened along the energy axis inside the yield kernel.
// Included by sshg_yield.cu inside its anonymous namespace (uses YieldArgs, EPT, store2, ldc and the
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
// A CTA of NT threads owns 2 NT consecutive energies of a column.  (A) it evaluates column setup +
// coefficients + harmonics for those energies plus a halo of `radius` on each side ('reflect' boundary of
// scipy.ndimage at the ends of the axis) into shared memory, raw[46][S]; (B) per polarisation every
// thread filters the harmonic spectra for its two energies (one LDS.128 feeds 4 FMAs) and parks the
// results in shared memory in place of the raw spectra, which are dead by then; (C) the azimuth is swept
// table chunk by table chunk (the chunks live behind the parked harmonics, so they are as long as the
// CTA's shared-memory share allows), four polarisations per chunk, phi / phi + 180 deg pairs sharing the
// even and odd harmonic sums.
// Accuracy: the sum of harmonics cancels where the yield has a node in phi, so its error is absolute --
// <= 5e-14 of the column's amplitude bound -- instead of relative.  Values next to a node are flagged in the
// sweep and re-evaluated row-wise by yield_blur_fixup_kernel (below), which restores the element-wise 1e-10.
constexpr int NT_BLUR = 64;                     // 128-energy tiles
constexpr int NT_BLUR_WIDE = 128;               // 256-energy tiles (half the halo overhead), 2 CTAs/SM up to radius 28
constexpr int NA = 3 * NH + NH_SS;              // harmonic spectra per column (46)
constexpr int BLUR_RMAX = 64;                   // largest fused radius (sigma_out <= 16)

struct BlurArgs {
    YieldArgs y;                                // y.tab: harmonic table, 12 doubles per item
    int radius;
    int S;                                      // staged energies per CTA = 2 NT + 2 radius
    int chunk;                                  // phi items per table chunk (what fits behind the parked harmonics)
    unsigned char* flags;                       // [column][fix-up tile][pol][row]: 1 = re-evaluate strictly (or null)
    int* work;                                  // work[0]: number of (column, tile) blocks with a flag; work[1 + k]: the
                                                // k-th such block; work[1 + blocks + t]: block t already listed
    int fx_tiles;                               // fix-up tiles (FX_TILE energies) per column
    int n_blocks;                               // columns of the launch x fx_tiles
    double w[BLUR_RMAX + 1];                    // Gaussian taps, w[0] centre
};

__global__ void phi_harm_table_kernel(const double* __restrict__ phi, long long nP, double* __restrict__ tab,
                                      int allow_pairs) {
    __shared__ int bad;
    const long long H = nP / 2;
    if (threadIdx.x == 0) bad = (allow_pairs && H > 0) ? 0 : 1;
    __syncthreads();
    if (allow_pairs && H > 0) {
        int mybad = 0;
        for (long long j = threadIdx.x; j < H; j += blockDim.x)
            if (phi[j + H] != phi[j] + 180.0) mybad = 1;
        if (mybad) bad = 1;
    }
    __syncthreads();
    const long long h = bad ? 0 : H;
    if (threadIdx.x == 0) tab[0] = (double)h;
    const long long items = nP - h;
    for (long long i = threadIdx.x; i < items; i += blockDim.x) {
        double t[12];
        phi_harmonics(phi[i < h ? i : i + h], t);
#pragma unroll
        for (int k = 0; k < 12; ++k) tab[2 + i * 12 + k] = t[k];     // items start 16-byte aligned
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
    if (R == 0) {                               // no broadening (experimental sigma_out = 0 path): the harmonics themselves
#pragma unroll
        for (int m = 0; m < NHH; ++m) {
            const double2 v = *reinterpret_cast<const double2*>(x + m * S);
            A0[m] = v.x;
            A1[m] = v.y;
        }
        return;
    }
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

// (C) azimuthal sweep of the harmonic form, two energies per thread.  tb: items of 12 doubles
// {cos, sin of 2, 4, 6 phi | cos, sin of 1, 3, 5 phi}.
// Points whose yield is far below the amplitude bound U >= max_phi y of their column are where the harmonic
// sum cancels: |error| <= 5e-14 U is then no longer small against the value.  They are flagged, per
// (column, fix-up tile, polarisation, row), and `yield_blur_fixup_kernel` re-evaluates those row segments the
// strict way (NODE_FRACTION, amplitude_bound: sshg_math.cuh).
constexpr int FX_NT = 64;                       // fix-up kernel: threads per CTA
constexpr int FX_TILE = FX_NT * EPT;            // ... and energies per tile (flag granularity along E)
constexpr int FX_RB = 4;                        // flagged rows evaluated together

// One azimuth item of the harmonic form for two energies: even and odd harmonic sums.
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

// (C) azimuthal sweep of the harmonic form, two energies per thread.  tb: items of 12 doubles
// {cos, sin of 2, 4, 6 phi | cos, sin of 1, 3, 5 phi}; items below n_pair give the rows i and i + H.
//
// Node test.  The hot loop is branch-free: besides the stores it only tracks the minimum of the HIGH WORDS of the
// values it produced (the high word of a double orders like the double, a negative value is below everything, a
// NaN above), two integer ops per value.  After every SWEEP_SB (16; 8 and 32 measured the same within 2 %) items the minimum is compared with the high word of
// NODE_FRACTION * U; only if a value of the block is that small -- rare -- the block is walked again: values in
// the band (DEEP, NODE) * U flag their row segment for the fix-up kernel (and list the block once), negative
// rounding noise at exact nodes is stored again clamped at 0.
#ifndef SSHG_SWEEP_SB
#define SSHG_SWEEP_SB 16
#endif
constexpr int SWEEP_SB = SSHG_SWEEP_SB;

template <bool SS, bool VEC>
__device__ __forceinline__ void sweep_harm(const double* A0, const double* A1, const double* __restrict__ tb,
                                           int n_pair, int n_single, double* plo, double* phi_, double* psg,
                                           long long stride, bool valid1, unsigned char* flo,
                                           unsigned char* fhi, unsigned char* fsg, int* work, int* listed, int block_id) {
#ifdef SSHG_K3_NOFLAG                            // A/B measurement only: the sweep without the node test
    const bool flagging = false;
#else
    const bool flagging = flo != nullptr;
#endif
    const double u0 = flagging ? amplitude_bound(A0, SS) : 0.0, u1 = (flagging && valid1) ? amplitude_bound(A1, SS) : 0.0;
    const double thr0 = NODE_FRACTION * u0, thr1 = NODE_FRACTION * u1;
    const double deep0 = DEEP_FRACTION * u0, deep1 = DEEP_FRACTION * u1;
    const int INT_LOWEST = (int)0x80000000;     // when not flagging: nothing is ever below
    const int h0 = flagging ? __double2hiint(thr0) : INT_LOWEST, h1 = (flagging && valid1) ? __double2hiint(thr1) : INT_LOWEST;
    auto node = [&](double y0, double y1, unsigned char* flag, double* p) {
        if (!((y0 < thr0) | (y1 < thr1))) return;
        if (((y0 < thr0) & (y0 > deep0)) | ((y1 < thr1) & (y1 > deep1))) {
            *flag = 1;
            if (atomicExch(listed, 1) == 0) work[1 + atomicAdd(work, 1)] = block_id;   // first flag of this block
        }
        if ((y0 < 0.0) | (y1 < 0.0)) store2t<VEC>(p, fmax(y0, 0.0), fmax(y1, 0.0), valid1);
    };
    // pairs first (rows i and i + H), then the un-paired tail; each part in blocks of SWEEP_SB items
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
        const int n = part == 0 ? n_pair : n_single;
        const double* tp = tb + (part == 0 ? 0 : 12 * n_pair);
        double* p0 = part == 0 ? plo : psg;
        double* p1 = phi_;
        unsigned char* f0 = part == 0 ? flo : fsg;
#pragma unroll 1
        for (int i0 = 0; i0 < n; i0 += SWEEP_SB) {
            const int nb = min(SWEEP_SB, n - i0);
            int mn0 = 0x7fffffff, mn1 = 0x7fffffff;
            const double* t = tp;
            double *q0 = p0, *q1 = p1;
#pragma unroll 2
            for (int i = 0; i < nb; ++i) {
                double ye0, ye1, yo0, yo1;
                harm_item<SS>(A0, A1, t, ye0, ye1, yo0, yo1);
                t += 12;
                const double a0 = ye0 + yo0, a1 = ye1 + yo1;
                store2t<VEC>(q0, a0, a1, valid1);
                q0 += stride;
                mn0 = min(mn0, __double2hiint(a0));
                mn1 = min(mn1, __double2hiint(a1));
                if (part == 0) {
                    const double s0 = ye0 - yo0, s1 = ye1 - yo1;
                    store2t<VEC>(q1, s0, s1, valid1);
                    q1 += stride;
                    mn0 = min(mn0, __double2hiint(s0));
                    mn1 = min(mn1, __double2hiint(s1));
                }
            }
            if ((mn0 < h0) | (mn1 < h1)) {      // rare: a value of this block is next to a node
                t = tp;
                q0 = p0;
                q1 = p1;
#pragma unroll 2
                for (int i = 0; i < nb; ++i) {
                    double ye0, ye1, yo0, yo1;
                    harm_item<SS>(A0, A1, t, ye0, ye1, yo0, yo1);          // the same arithmetic: the same bits
                    t += 12;
                    node(ye0 + yo0, ye1 + yo1, f0 + i0 + i, q0);
                    q0 += stride;
                    if (part == 0) {
                        node(ye0 - yo0, ye1 - yo1, fhi + i0 + i, q1);
                        q1 += stride;
                    }
                }
            }
            tp += 12 * nb;
            p0 += (long long)nb * stride;
            if (part == 0) p1 += (long long)nb * stride;
        }
    }
}

// shared memory: wf[2 R + 2] | raw[NA][S], later blur[NA][2 NT] in its place | ... up to the end of the allocation;
// the phi table chunks live behind blur (the tail of raw is dead by then).
template <int NT>
__global__ void __launch_bounds__(NT, (NT <= 64 ? 3 : 2)) yield_blur_kernel(const __grid_constant__ BlurArgs b) {
    extern __shared__ double smem[];
    const YieldArgs& a = b.y;
    const int R = b.radius, S = b.S, BLUR_CHUNK = b.chunk;
    double* wf = smem;                          // [2 R + 2] full taps (padded to an even count)
    double* raw = wf + 2 * R + 2;               // [NA][S]
    double* tabs = raw + NA * NT * EPT;         // [chunk][12], valid once every FIR has been done

    const long long blk = blockIdx.x;
    const int tile = (int)(blk % a.n_tiles);
    const int seg = (int)((blk / a.n_tiles) % a.n_seg);
    const long long lcol = blk / a.n_tiles / a.n_seg;
    const long long col = a.col0 + lcol;
    const int tl = (int)(col / a.nDs), dl = (int)(col % a.nDs);
    const int tid = threadIdx.x;
    const long long nE = a.nE;
    const long long tile_e0 = (long long)tile * NT * EPT;

    for (int t = tid; t <= 2 * R; t += NT) wf[t] = b.w[t < R ? R - t : t - R];

    // ---- (A) harmonic spectra of the staged energies (tile + halo) --------------------------------------
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
            auto chi = [&](int r) { return ldc(a.chi + r * nE + e); };
            double* dst = raw + i;
#pragma unroll 1
            for (int pol = 0; pol < 4; ++pol) {
                if (pol == 3) {
                    cplx kk[4];
                    coeffs_ss(c, chi, kk);
                    double H[NH_SS];
                    harmonics4(kk, H);
#pragma unroll
                    for (int m = 0; m < NH_SS; ++m) dst[(3 * NH + m) * S] = H[m];
                } else {
                    cplx kk[7];
                    if (pol == 0) coeffs_pp(c, chi, a.ph, kk);
                    else if (pol == 1) coeffs_ps(c, chi, kk);
                    else coeffs_sp(c, chi, kk);
                    double H[NH];
                    harmonics7(kk, H);
#pragma unroll
                    for (int m = 0; m < NH; ++m) dst[(pol * NH + m) * S] = H[m];
                }
            }
        }
    }
    __syncthreads();

    const long long e0 = tile_e0 + (long long)tid * EPT;
    const bool active = e0 < nE;
    const bool valid1 = e0 + 1 < nE;
    const long long H = (long long)a.tab[0];
    const long long items = a.nP - H;
    const bool vec = a.vec_ok != 0;
    const long long rows_per_pol = a.nP * nE;
    double* out_col = a.out + lcol * 4 * rows_per_pol + e0;
    const long long seg_len = (items + a.n_seg - 1) / a.n_seg;
    const long long seg_hi = min(items, (seg + 1) * seg_len);

    // ---- (B) broadened harmonics of this thread's two energies, parked in place of the raw spectra -----------
    // blur[m][2 NT] aliases the start of raw: spectrum m's slot lies inside raw rows <= m, which every thread
    // has finished reading once the barrier after that polarisation's FIR is passed; later polarisations'
    // raw rows are untouched.  Each thread only ever reads back the two values it wrote itself.
    constexpr int TE = NT * EPT;
    double* blur = raw + tid * EPT;
#pragma unroll 1
    for (int pol = 0; pol < 4; ++pol) {
        double A0[NH], A1[NH];
        const double* x = raw + (pol * NH) * S + tid * EPT;
        if (pol == 3) fir_pair<NH_SS>(x, S, R, wf, A0, A1);
        else fir_pair<NH>(x, S, R, wf, A0, A1);
        __syncthreads();
        const int nh = pol == 3 ? NH_SS : NH;
#pragma unroll
        for (int m = 0; m < NH; ++m)
            if (m < nh) *reinterpret_cast<double2*>(blur + (pol * NH + m) * TE) = make_double2(A0[m], A1[m]);
    }

    // ---- (C) azimuthal sweep: table chunk by chunk, the four polarisations per chunk -----------------------
    for (long long base = seg * seg_len; base < seg_hi; base += BLUR_CHUNK) {
        const int n_items = (int)min((long long)BLUR_CHUNK, seg_hi - base);
        __syncthreads();                        // previous chunk fully consumed
        for (int i = tid; i < n_items * 12; i += NT) tabs[i] = a.tab[2 + base * 12 + i];
        __syncthreads();
        if (!active) continue;
        const int n_pair = (int)max(0LL, min((long long)n_items, H - base));
        const int n_single = n_items - n_pair;
#pragma unroll 1
        for (int pol = 0; pol < 4; ++pol) {
            double A0[NH], A1[NH];
            const int nh = pol == 3 ? NH_SS : NH;
#pragma unroll
            for (int m = 0; m < NH; ++m)
                if (m < nh) {
                    const double2 v = *reinterpret_cast<const double2*>(blur + (pol * NH + m) * TE);
                    A0[m] = v.x;
                    A1[m] = v.y;
                }
            double* prow = out_col + pol * rows_per_pol;
            double* plo = prow + base * nE;
            double* phi_ = prow + (base + H) * nE;
            double* psg = prow + (base + n_pair + H) * nE;
            unsigned char *flo = nullptr, *fhi = nullptr, *fsg = nullptr;
            int* listed = nullptr;
            int block_id = 0;
            if (b.flags) {                      // flag block of this (column, fix-up tile, polarisation): one byte per row
                block_id = (int)(lcol * b.fx_tiles + e0 / FX_TILE);
                unsigned char* fb = b.flags + ((long long)block_id * 4 + pol) * a.nP;
                flo = fb + base;
                fhi = fb + base + H;
                fsg = fb + base + n_pair + H;
                listed = b.work + 1 + b.n_blocks + block_id;
            }
#define SSHG_SWEEP(SS_, VEC_) \
    sweep_harm<SS_, VEC_>(A0, A1, tabs, n_pair, n_single, plo, phi_, psg, nE, valid1, flo, fhi, fsg, b.work, listed, block_id)
            if (vec) {
                if (pol == 3) SSHG_SWEEP(true, true);
                else SSHG_SWEEP(false, true);
            } else {
                if (pol == 3) SSHG_SWEEP(true, false);
                else SSHG_SWEEP(false, false);
            }
#undef SSHG_SWEEP
        }
    }
}


// ---- K3 fix-up: the flagged row segments, evaluated the strict way ----------------------------------------
// A CTA owns the FX_TILE energies of one (column, fix-up tile).  If none of its 4 nP flags is set it exits.
// Otherwise it evaluates the column setup for its energies + halo (parked in shared memory), per polarisation with
// flagged rows the 7 (sS: 4) complex coefficients of K2, and per flagged row the un-broadened yields |r(phi)|^2
// of the staged energies exactly as K2 does, filters them along E and overwrites K3's values.  The yields it
// writes are relative-accurate next to an azimuthal node, like the reference's.
struct FixArgs {
    YieldArgs y;                                // y.tab unused
    const double* phi_deg;
    const unsigned char* flags;
    const int* work;                            // the list K3 filled (see BlurArgs)
    int radius;
    int S;                                      // FX_TILE + 2 radius
    double w[BLUR_RMAX + 1];
};

__global__ void __launch_bounds__(FX_NT) yield_blur_fixup_kernel(const __grid_constant__ FixArgs b) {
    extern __shared__ double smem[];
    const YieldArgs& a = b.y;
    const int R = b.radius, S = b.S;
    double* wf = smem;                          // [2 R + 2]
    double* cols = wf + 2 * R + 2;              // [NCOL][S]   parked Column, struct of arrays
    double* kb = cols + NCOL * S;               // [14][S]     coefficients of the current polarisation
    double* rowbuf = kb + 14 * S;               // [FX_RB][S]  un-broadened yields of the current batch of rows

    const int tid = threadIdx.x;
    const long long nE = a.nE, nP = a.nP;
    const int n_work = b.work[0];               // blocks with at least one flagged row segment (usually few or none)
    for (int t = tid; t <= 2 * b.radius; t += FX_NT) wf[t] = b.w[t < b.radius ? b.radius - t : t - b.radius];
#pragma unroll 1
    for (int item = blockIdx.x; item < n_work; item += gridDim.x) {
    __syncthreads();                            // the previous block's shared memory is no longer read
    const long long blk = b.work[1 + item];
    const int tile = (int)(blk % a.n_tiles);
    const long long lcol = blk / a.n_tiles;
    const unsigned char* fb = b.flags + blk * 4 * nP;

    const long long col = a.col0 + lcol;
    const int tl = (int)(col / a.nDs), dl = (int)(col % a.nDs);
    const long long tile_e0 = (long long)tile * FX_TILE;
    {
        double st, ct;
        sincos(a.theta_deg[a.t0 + tl] * (3.141592653589793 / 180.0), &st, &ct);
        const double thick = a.thick_nm[a.d0 + dl];
#pragma unroll 1
        for (int i = tid; i < S; i += FX_NT) {
            const long long g = tile_e0 - R + i;
            if (g >= nE + R) break;
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
    __syncthreads();

    const long long e0 = tile_e0 + (long long)tid * EPT;
    const bool active = e0 < nE, valid1 = e0 + 1 < nE;
    const bool vec = a.vec_ok != 0;
    const long long rows_per_pol = nP * nE;
    double* out_col = a.out + lcol * 4 * rows_per_pol + e0;

#pragma unroll 1
    for (int pol = 0; pol < 4; ++pol) {
        int anyp = 0;
        for (long long i = tid; i < nP; i += FX_NT) anyp |= fb[pol * nP + i];
        if (!__syncthreads_or(anyp)) continue;
        const int nk = pol == 3 ? 4 : 7;
#pragma unroll 1
        for (int i = tid; i < S; i += FX_NT) {
            const long long g = tile_e0 - R + i;
            if (g >= nE + R) break;
            const long long e = reflect_index(g, nE);
            auto parked = [&](int j) { return mk(cols[(2 * j) * S + i], cols[(2 * j + 1) * S + i]); };
            Column c;
            c.Gpp = parked(0); c.Gps = parked(1); c.Gsp = parked(2); c.Gss = parked(3);
            c.A = parked(4); c.Bz = parked(5); c.Q = parked(6);
            c.a2 = parked(7); c.ab = parked(8); c.b2 = parked(9);
            auto chi = [&](int r) { return ldc(a.chi + r * nE + e); };
            cplx kk[7];
            if (pol == 0) coeffs_pp(c, chi, a.ph, kk);
            else if (pol == 1) coeffs_ps(c, chi, kk);
            else if (pol == 2) coeffs_sp(c, chi, kk);
            else coeffs_ss(c, chi, kk);
            for (int j = 0; j < nk; ++j) {
                kb[(2 * j) * S + i] = kk[j].re;
                kb[(2 * j + 1) * S + i] = kk[j].im;
            }
        }
        __syncthreads();
        // flagged rows of this polarisation, FX_RB at a time: one pass over the staged energies evaluates the
        // un-broadened yields of the whole batch, one pass filters them (two barriers per batch, FX_RB independent
        // FMA chains per thread)
        long long row = 0;
#pragma unroll 1
        while (row < nP) {
            int rows[FX_RB], nb = 0;
            for (; row < nP && nb < FX_RB; ++row)
                if (fb[pol * nP + row]) rows[nb++] = (int)row;      // the same bytes for every thread: uniform
            if (nb == 0) break;
            double bb[FX_RB][6];
#pragma unroll
            for (int q = 0; q < FX_RB; ++q) phi_basis(b.phi_deg[rows[q < nb ? q : 0]], bb[q]);
            for (int i = tid; i < S; i += FX_NT) {
                double kr[7], ki[7];
#pragma unroll
                for (int j = 0; j < 7; ++j)
                    if (j < nk) {
                        kr[j] = kb[(2 * j) * S + i];
                        ki[j] = kb[(2 * j + 1) * S + i];
                    }
#pragma unroll
                for (int q = 0; q < FX_RB; ++q) {
                    double re, im;
                    if (pol == 3) {                             // k[0..3] multiply {c^3, s c^2, s^2 c, s^3}
                        re = kr[0] * bb[q][2];
                        im = ki[0] * bb[q][2];
#pragma unroll
                        for (int j = 1; j < 4; ++j) {
                            re = fma(kr[j], bb[q][j + 2], re);
                            im = fma(ki[j], bb[q][j + 2], im);
                        }
                    } else {
                        re = kr[0];
                        im = ki[0];
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            re = fma(kr[j + 1], bb[q][j], re);
                            im = fma(ki[j + 1], bb[q][j], im);
                        }
                    }
                    rowbuf[q * S + i] = fma(re, re, im * im);
                }
            }
            __syncthreads();
            if (active) {
                double o0[FX_RB], o1[FX_RB];
                const double* x = rowbuf + tid * EPT;               // x[q S + 0 .. 2 R + 1]
#pragma unroll
                for (int q = 0; q < FX_RB; ++q) {                   // centre tap, then the pairs from the farthest inwards
                    o0[q] = x[q * S + R] * wf[R];
                    o1[q] = x[q * S + R + 1] * wf[R];
                }
                for (int j = R; j >= 1; --j) {
                    const double w = wf[R + j];
#pragma unroll
                    for (int q = 0; q < FX_RB; ++q) {
                        o0[q] = fma(x[q * S + R - j] + x[q * S + R + j], w, o0[q]);
                        o1[q] = fma(x[q * S + R + 1 - j] + x[q * S + R + 1 + j], w, o1[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < FX_RB; ++q)
                    if (q < nb) store2(out_col + (pol * nP + rows[q]) * nE, o0[q], o1[q], vec, valid1);
            }
            __syncthreads();
        }
    }
    }   // work items
}
