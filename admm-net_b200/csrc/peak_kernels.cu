// Grid peak search (reference utils/peakSearchUtils.py:9-173, utils/mathUtils.py:4-21) in fp64.
//
//   surface  Z[iy][ix] = | sum_{p<yb} sum_{q<xb} conj(phi[p*xb+q]) e^{+j2pi fre_p(y_iy)} e^{-j2pi fre_q(x_ix)} |^2
//            evaluated separably: T = S conj(Phi), Z = |T conj(D)^T|^2
//   maxima   8-connected, plateau-aware, borders allowed (skimage.morphology.local_maxima semantics,
//            oracle/peak_oracle.py::local_maxima), listed in row-major order (np.where)
//   refine   `iter` rounds of np.arange windows around every peak, first-occurrence argmax
//            (peakSearchUtils.py:136-171)
//   top-L    stable sort by height, descending (main_for_net.py:119-126)
// All grid positions are produced with numpy's arange/linspace arithmetic so that they are bit-identical
// to the reference's; the double-precision sums differ from BLAS' order at the 1e-16 level only.
#include "common.cuh"

namespace admmnet {

struct PeakArgs {
    const void* phi;       // [B][n] complex64 or complex128
    int phi_is_c128;
    int B, xb, yb;         // n = xb*yb; x <-> tau uses xb, y <-> f uses yb
    const double* axis_x;  // [Gx] coarse grid (np.arange on the host / caller)
    const double* axis_y;  // [Gy]
    int Gx, Gy;
    double xmin, xmax, xstep, ymin, ymax, ystep, reducefactor;
    int iters;
    int pmax;              // capacity of peaks per signal
    int ptile;             // peaks refined per block-wide round (sizes the refinement scratch)
    double* peaks;         // [B][pmax][3]  (x, y, height), discovery order
    int* count;            // [B]
    int topl;              // 0 = skip
    double* top;           // [B][topl][3]
    double* surface;       // optional [B][Gy][Gx]
    int* status;
    const double* surface_in;  // optional [B][Gy][Gx]: debug tap, the local-maximum stage runs on THIS surface
};

#define TWO_PI_D 6.283185307179586   // == 2*np.pi in binary64

__device__ __forceinline__ double np_arange_val(double start, double step, int i) {
    // numpy DOUBLE_fill: buf[0]=start, buf[1]=start+step, buf[i]=start+i*(buf[1]-buf[0])
    if (i == 0) return start;
    const double b1 = __dadd_rn(start, step);
    if (i == 1) return b1;
    const double delta = __dsub_rn(b1, start);
    return __dadd_rn(start, __dmul_rn((double)i, delta));
}
__device__ __forceinline__ int np_arange_len(double start, double stop, double step) {
    const double q = __ddiv_rn(__dsub_rn(stop, start), step);
    const double c = ceil(q);
    return c > 0.0 ? (c > 1.0e6 ? 1000000 : (int)c) : 0;
}
// vander_vec(0,(len-1)*v,len)[k] = exp(1j*2*pi*linspace(0,(len-1)*v,len)[k])
__device__ __forceinline__ double2 steer(double v, int k, int len) {
    double fre = 0.0;
    if (len > 1) {
        const double stop = __dmul_rn((double)(len - 1), v);
        if (k == len - 1) fre = stop;
        else {
            const double step = __ddiv_rn(stop, (double)(len - 1));
            fre = __dmul_rn((double)k, step);
        }
    }
    const double th = __dmul_rn(TWO_PI_D, fre);
    double s, c;
    sincos(th, &s, &c);
    return make_double2(c, s);
}
__device__ __forceinline__ double2 dmul(double2 a, double2 b) {
    return make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)),
                        __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
__device__ __forceinline__ double abs2_np(double2 w) {
    const double a = hypot(w.x, w.y);   // np.abs(.) ** 2
    return a * a;
}

// value of the spectrum at one (x,y): | sum_p s_p(y) * conj( sum_q phi[p][q] d_q(x) ) |^2 ; phis in shared memory
#define PEAK_MAX_BASE 32
__device__ double spectrum_point(const double2* __restrict__ phis, int xb, int yb, double x, double y) {
    double2 dq[PEAK_MAX_BASE];
    for (int q = 0; q < xb; ++q) dq[q] = steer(x, q, xb);
    double2 acc = make_double2(0.0, 0.0);
    for (int p = 0; p < yb; ++p) {
        const double2 sp = steer(y, p, yb);
        double2 row = make_double2(0.0, 0.0);
        for (int q = 0; q < xb; ++q) {
            const double2 ph = phis[p * xb + q];
            row.x += ph.x * dq[q].x - ph.y * dq[q].y;        // phi * d
            row.y += ph.x * dq[q].y + ph.y * dq[q].x;
        }
        acc.x += row.x * sp.x + row.y * sp.y;                // conj(row) * s
        acc.y += row.x * sp.y - row.y * sp.x;
    }
    return abs2_np(acc);
}

__host__ __device__ inline size_t peak_smem_bytes(int Gx, int Gy, int xb, int yb, int pmax) {
    size_t dbl = (((size_t)Gx * Gy + 1) & ~(size_t)1)   // Z (padded: the double2 arrays behind it need 16 B alignment)
                 + 2 * ((size_t)Gy * yb + (size_t)Gx * xb + (size_t)Gy * xb + (size_t)xb * yb)   // Sy, Dx, T, phi
                 + 64;                      // reductions
    size_t bytes = dbl * 8 + (((size_t)Gx * Gy + 15) & ~(size_t)15) /*flags*/ + 16 + (256 + 8) * 4 /*scan*/ +
                   (((size_t)pmax * 4 + 15) & ~(size_t)15);
    return (bytes + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t peak_refine_bytes_per_peak(int xb, int yb) {
    return 16 /*x0,y0*/ + 128 /*vals*/ + (size_t)4 * yb * 16 + (size_t)8 * xb * 16 + 8 /*nx,ny*/;
}

#define PEAK_NT 512
__global__ void __launch_bounds__(PEAK_NT) k_peak_search(PeakArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Gx = a.Gx, Gy = a.Gy, xb = a.xb, yb = a.yb, n = xb * yb, N = Gx * Gy;
    double* Z = reinterpret_cast<double*>(smem_raw);
    double2* Sy = reinterpret_cast<double2*>(Z + ((N + 1) & ~1));   // [Gy][yb]
    double2* Dx = Sy + (size_t)Gy * yb;                        // [Gx][xb]
    double2* T = Dx + (size_t)Gx * xb;                         // [Gy][xb]
    double2* phis = T + (size_t)Gy * xb;                       // [n]
    double* redd = reinterpret_cast<double*>(phis + n);        // [64]
    int* scan = reinterpret_cast<int*>(redd + 64);             // [256+8]
    int* plist = scan + 256 + 8;                               // [pmax] flat pixel index of each peak
    unsigned char* flag = reinterpret_cast<unsigned char*>(plist + ((a.pmax + 3) & ~3));   // [N], 16 B aligned
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int sig = blockIdx.x;

    for (int i = tid; i < n && !a.surface_in; i += PEAK_NT) {
        if (a.phi_is_c128) phis[i] = reinterpret_cast<const double2*>(a.phi)[(size_t)sig * n + i];
        else {
            const float2 v = reinterpret_cast<const float2*>(a.phi)[(size_t)sig * n + i];
            phis[i] = make_double2((double)v.x, (double)v.y);
        }
    }
    for (int i = tid; i < Gy * yb && !a.surface_in; i += PEAK_NT) Sy[i] = steer(a.axis_y[i / yb], i % yb, yb);
    for (int i = tid; i < Gx * xb && !a.surface_in; i += PEAK_NT) Dx[i] = steer(a.axis_x[i % Gx], i / Gx, xb);   // [xb][Gx]
    __syncthreads();
    // T[iy][q] = sum_p conj(phi[p][q]) * Sy[iy][p]
    for (int i = tid; i < Gy * xb && !a.surface_in; i += PEAK_NT) {
        const int iy = i / xb, q = i % xb;
        double2 acc = make_double2(0.0, 0.0);
        for (int p = 0; p < yb; ++p) {
            const double2 ph = phis[p * xb + q], s = Sy[iy * yb + p];
            acc.x += ph.x * s.x + ph.y * s.y;
            acc.y += ph.x * s.y - ph.y * s.x;
        }
        T[i] = acc;
    }
    __syncthreads();
    double lmin = INFINITY, lmax = -INFINITY;
    for (int i = tid; i < N; i += PEAK_NT) {
        const int iy = i / Gx, ix = i % Gx;
        double z;
        if (a.surface_in) {
            z = a.surface_in[(size_t)sig * N + i];
        } else {
            double2 acc = make_double2(0.0, 0.0);
            for (int q = 0; q < xb; ++q) {
                const double2 t = T[iy * xb + q], dq = Dx[q * Gx + ix];
                acc.x += t.x * dq.x + t.y * dq.y;
                acc.y += t.y * dq.x - t.x * dq.y;
            }
            z = abs2_np(acc);
        }
        Z[i] = z;
        lmin = fmin(lmin, z);
        lmax = fmax(lmax, z);
        if (a.surface) a.surface[(size_t)sig * N + i] = z;
    }
    // block min / max (constant image has no maxima)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    if (lane == 0) { redd[wid] = lmin; redd[16 + wid] = lmax; }
    __syncthreads();
    double gmin = redd[0], gmax = redd[16];
    for (int w = 1; w < PEAK_NT / 32; ++w) { gmin = fmin(gmin, redd[w]); gmax = fmax(gmax, redd[16 + w]); }
    const bool constant = !(gmax > gmin);

    // ---- candidates: no strictly greater 8-neighbour (outside the image counts as lower)
    for (int i = tid; i < N; i += PEAK_NT) {
        const int iy = i / Gx, ix = i % Gx;
        const double z = Z[i];
        bool cand = !constant;
        for (int dy = -1; dy <= 1 && cand; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int yy = iy + dy, xx = ix + dx;
                if ((dy | dx) == 0 || yy < 0 || yy >= Gy || xx < 0 || xx >= Gx) continue;
                if (Z[yy * Gx + xx] > z) { cand = false; break; }
            }
        flag[i] = cand ? 1 : 0;
    }
    __syncthreads();
    // ---- plateaus: a candidate with an equal-valued non-candidate neighbour is not a maximum
    for (int guard = 0; guard < N; ++guard) {
        int changed = 0;
        for (int i = tid; i < N; i += PEAK_NT) {
            if (!flag[i]) continue;
            const int iy = i / Gx, ix = i % Gx;
            const double z = Z[i];
            bool kill = false;
            for (int dy = -1; dy <= 1 && !kill; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int yy = iy + dy, xx = ix + dx;
                    if ((dy | dx) == 0 || yy < 0 || yy >= Gy || xx < 0 || xx >= Gx) continue;
                    if (Z[yy * Gx + xx] == z && !flag[yy * Gx + xx]) { kill = true; break; }
                }
            if (kill) { flag[i] = 0; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
    }
    // ---- ordered compaction (row-major == np.where order)
    const int seg = (N + 255) / 256;
    const int s0 = tid < 256 ? min(N, tid * seg) : N, s1 = min(N, s0 + seg);
    int cnt = 0;
    for (int i = s0; i < s1; ++i) cnt += flag[i];
    if (tid < 256) scan[tid] = cnt;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int t = 0; t < 256; ++t) { const int c = scan[t]; scan[t] = run; run += c; }
        scan[256] = run;
    }
    __syncthreads();
    const int P = scan[256];
    const int Pst = min(P, a.pmax);
    if (tid < 256) {
        int o = scan[tid];
        for (int i = s0; i < s1; ++i)
            if (flag[i]) { if (o < a.pmax) plist[o] = i; ++o; }
    }
    if (tid == 0) {
        a.count[sig] = P;
        if (P > a.pmax) atomicOr(a.status, 2);
    }
    __syncthreads();

    // ---- refinement (peakSearchUtils.py:136-171), tiles of PT peaks, block-wide phases per round:
    //   R1 window per peak -> R2 steering vectors of the <=4+4 distinct local y/x values -> R3 row sums
    //   t_q(y) -> R4 the <=16 local values -> R5 first-occurrence argmax.  Windows wider than 4 points per
    //   axis (cannot happen for reducefactor <= 1/2, kept for generality) take the scalar fallback.
    double* out = a.peaks + (size_t)sig * a.pmax * 3;
    const int PT = a.ptile;
    double* rx0 = reinterpret_cast<double*>(flag + ((N + 15) & ~15));      // [PT]
    double* ry0 = rx0 + PT;                                                // [PT]
    double* vals = ry0 + PT;                                               // [PT][16]
    double2* svec = reinterpret_cast<double2*>(vals + 16 * PT);            // [PT][4][yb]
    double2* dvec = svec + (size_t)PT * 4 * yb;                            // [PT][4][xb]
    double2* tvec = dvec + (size_t)PT * 4 * xb;                            // [PT][4][xb]
    int* rnx = reinterpret_cast<int*>(tvec + (size_t)PT * 4 * xb);         // [PT]
    int* rny = rnx + PT;                                                   // [PT]
    for (int k = tid; k < Pst; k += PEAK_NT) {
        const int pix = plist[k];
        out[3 * k] = a.axis_x[pix % Gx];
        out[3 * k + 1] = a.axis_y[pix / Gx];
        out[3 * k + 2] = a.surface_in ? Z[pix] : 0.0;
    }
    __syncthreads();
    for (int base = 0; base < Pst; base += PT) {
        const int np = min(PT, Pst - base);
        double lx = a.xstep, ly = a.ystep;
        for (int it = 0; it < a.iters; ++it) {
            lx = __dmul_rn(a.reducefactor, lx);
            ly = __dmul_rn(a.reducefactor, ly);
            // R1
            for (int j = tid; j < np; j += PEAK_NT) {
                const int k = base + j;
                const double px = out[3 * k], py = out[3 * k + 1];
                const double x0 = fmax(a.xmin, __dsub_rn(px, lx)), x1 = fmin(__dsub_rn(a.xmax, lx), __dadd_rn(px, lx));
                const double y0 = fmax(a.ymin, __dsub_rn(py, ly)), y1 = fmin(__dsub_rn(a.ymax, ly), __dadd_rn(py, ly));
                int nx = 0, ny = 0;
                if (!(x0 >= x1 || y0 >= y1)) {
                    nx = np_arange_len(x0, x1, lx);
                    ny = np_arange_len(y0, y1, ly);
                    if (nx == 0 || ny == 0) nx = ny = 0;
                }
                if (nx > 4 || ny > 4) {                      // scalar fallback, any window size
                    double best = -INFINITY;
                    int bidx = -1;
                    for (int q = 0; q < nx * ny; ++q) {
                        const double z = spectrum_point(phis, xb, yb, np_arange_val(x0, lx, q % nx), np_arange_val(y0, ly, q / nx));
                        if (z > best) { best = z; bidx = q; }
                    }
                    if (bidx >= 0) {
                        out[3 * k] = np_arange_val(x0, lx, bidx % nx);
                        out[3 * k + 1] = np_arange_val(y0, ly, bidx / nx);
                        out[3 * k + 2] = best;
                    }
                    nx = ny = 0;
                }
                rx0[j] = x0; ry0[j] = y0; rnx[j] = nx; rny[j] = ny;
            }
            __syncthreads();
            // R2: steering vectors
            for (int i = tid; i < np * 4 * (yb + xb); i += PEAK_NT) {
                const int j = i / (4 * (yb + xb)), rem = i % (4 * (yb + xb));
                if (rem < 4 * yb) {
                    const int iy = rem / yb, pp = rem % yb;
                    if (iy < rny[j]) svec[((size_t)j * 4 + iy) * yb + pp] = steer(np_arange_val(ry0[j], ly, iy), pp, yb);
                } else {
                    const int r2 = rem - 4 * yb, ix = r2 / xb, qq = r2 % xb;
                    if (ix < rnx[j]) dvec[((size_t)j * 4 + ix) * xb + qq] = steer(np_arange_val(rx0[j], lx, ix), qq, xb);
                }
            }
            __syncthreads();
            // R3: t[iy][q] = sum_p conj(phi[p][q]) s_p(y_iy)
            for (int i = tid; i < np * 4 * xb; i += PEAK_NT) {
                const int j = i / (4 * xb), iy = (i / xb) & 3, qq = i % xb;
                if (iy < rny[j]) {
                    const double2* sv = svec + ((size_t)j * 4 + iy) * yb;
                    double2 acc = make_double2(0.0, 0.0);
                    for (int pp = 0; pp < yb; ++pp) {
                        const double2 ph = phis[pp * xb + qq], sp = sv[pp];
                        acc.x += ph.x * sp.x + ph.y * sp.y;
                        acc.y += ph.x * sp.y - ph.y * sp.x;
                    }
                    tvec[((size_t)j * 4 + iy) * xb + qq] = acc;
                }
            }
            __syncthreads();
            // R4: local surface values
            for (int i = tid; i < np * 16; i += PEAK_NT) {
                const int j = i >> 4, iy = (i >> 2) & 3, ix = i & 3;
                if (iy < rny[j] && ix < rnx[j]) {
                    const double2* tv = tvec + ((size_t)j * 4 + iy) * xb;
                    const double2* dv = dvec + ((size_t)j * 4 + ix) * xb;
                    double2 acc = make_double2(0.0, 0.0);
                    for (int qq = 0; qq < xb; ++qq) {
                        acc.x += tv[qq].x * dv[qq].x + tv[qq].y * dv[qq].y;
                        acc.y += tv[qq].y * dv[qq].x - tv[qq].x * dv[qq].y;
                    }
                    vals[i] = abs2_np(acc);
                }
            }
            __syncthreads();
            // R5: first-occurrence maximum in row-major order
            for (int j = tid; j < np; j += PEAK_NT) {
                const int nx = rnx[j], ny = rny[j];
                if (nx > 0) {
                    double best = -INFINITY;
                    int bx = -1, by = -1;
                    for (int iy = 0; iy < ny; ++iy)
                        for (int ix = 0; ix < nx; ++ix) {
                            const double z = vals[j * 16 + iy * 4 + ix];
                            if (z > best) { best = z; bx = ix; by = iy; }
                        }
                    if (bx >= 0) {
                        const int k = base + j;
                        out[3 * k] = np_arange_val(rx0[j], lx, bx);
                        out[3 * k + 1] = np_arange_val(ry0[j], ly, by);
                        out[3 * k + 2] = best;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (a.topl <= 0) return;
    __syncthreads();
    __threadfence_block();
    // ---- top-L by height (stable: ties keep discovery order)
    double* top = a.top + (size_t)sig * a.topl * 3;
    unsigned char* taken = flag;   // reuse
    for (int i = tid; i < Pst; i += PEAK_NT) taken[i] = 0;
    __syncthreads();
    for (int l = 0; l < a.topl; ++l) {
        double best = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = tid; i < Pst; i += PEAK_NT)
            if (!taken[i]) {
                const double hgt = out[3 * i + 2];
                if (hgt > best) { best = hgt; bi = i; }
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
        }
        __syncthreads();
        if (lane == 0) { redd[wid] = best; scan[wid] = bi; }
        __syncthreads();
        double gb = redd[0];
        int gi = scan[0];
        for (int w = 1; w < PEAK_NT / 32; ++w) {
            const double ob = redd[w];
            const int oi = scan[w];
            if (oi == 0x7fffffff) continue;
            if (gi == 0x7fffffff || ob > gb || (ob == gb && oi < gi)) { gb = ob; gi = oi; }
        }
        if (tid == 0) {
            if (gi != 0x7fffffff) {
                top[3 * l] = out[3 * gi]; top[3 * l + 1] = out[3 * gi + 1]; top[3 * l + 2] = out[3 * gi + 2];
                taken[gi] = 1;
            } else {
                top[3 * l] = 0.0; top[3 * l + 1] = 0.0; top[3 * l + 2] = 0.0;
            }
        }
        __syncthreads();
    }
}

// Spectrum at arbitrary points (peak_search(phi, X, x_base, Y, y_base) with non-tensor grids).
__global__ void __launch_bounds__(256)
k_peak_points(const void* phi, int phi_is_c128, int xb, int yb, const double* __restrict__ X,
              const double* __restrict__ Y, int npts, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* phis = reinterpret_cast<double2*>(smem_raw);
    const int n = xb * yb;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (phi_is_c128) phis[i] = reinterpret_cast<const double2*>(phi)[i];
        else {
            const float2 v = reinterpret_cast<const float2*>(phi)[i];
            phis[i] = make_double2((double)v.x, (double)v.y);
        }
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npts) out[i] = spectrum_point(phis, xb, yb, X[i], Y[i]);
}

}  // namespace admmnet
