/*
 * admmnet_b200 — C ABI of the B200 (sm_100a) hot path of E-J408/admm-net.
 *
 * The reference has no FFI layer: the path sits behind three Python callables
 * (SURVEY.md §8b).  This header is the boundary a binding would target; the Python mirror of the
 * reference API in admm-net_b200/ (ctypes) is its only in-tree caller.  INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; admmnet_last_error() (thread local) explains.
 *   - all data pointers are DEVICE pointers unless the name ends in _host; the caller owns every
 *     buffer; nothing is allocated behind the caller's back (query *_workspace_bytes first).
 *   - work is ordered after everything queued on the caller's stream (cudaStream_t passed as void*)
 *     and the stream waits for it; admmnet_forward additionally fans chunks out over library-owned
 *     non-blocking streams (two chunk lanes + a high-priority lane for the latency-bound QL kernel)
 *     between a fork and a join event.  Those streams and events are kept per (device, caller stream), so calls
 *     are re-entrant across host threads and streams given distinct workspaces and distinct streams
 *     (tests/test_gpu_parity.py::test_forward_is_reentrant_across_host_threads_and_streams).  The device that
 *     owns the buffers must be current (cudaSetDevice) when calling.  No C++ exception crosses the ABI.
 *   - complex64 = interleaved float pairs, complex128 = interleaved double pairs, row-major.
 */
#ifndef ADMMNET_B200_H
#define ADMMNET_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADMMNET_OK 0
#define ADMMNET_ERR_ARG (-1)
#define ADMMNET_ERR_CUDA (-2)
#define ADMMNET_ERR_WORKSPACE (-3)

/* status word bits (admmnet_status / peak status) */
#define ADMMNET_STATUS_EIGH_FAILED 1     /* QL did not converge or rotation stream overflowed rcap */
#define ADMMNET_STATUS_PEAK_OVERFLOW 2   /* more local maxima than pmax (count[] still holds the true number) */

const char* admmnet_last_error(void);
int admmnet_version(void);

/* ---------------------------------------------------------------------------------------------
 * Parameters.  One float record per layer, admmnet_param_stride(n) floats apart; layout in
 * admm-net_b200/csrc/common.cuh (enum ParamOff), written by admm-net_b200/params.py from a
 * reference-format state_dict (keys of admm_net.py:727-739).
 * ------------------------------------------------------------------------------------------- */
int admmnet_param_stride(int n);

/* ---------------------------------------------------------------------------------------------
 * PhiEstADMMNet.forward  (replaces admm_net.py:742-764; layers: 79-105, 134-194, 237-354, 388-474)
 *   y, b     complex64 [B][n], n = Mdim*Ndim <= 127
 *   sigma    float32  [B]
 *   params   float32  [K][admmnet_param_stride(n)]
 *   phi_out  complex64 [B][n]
 *   rcap     per-signal capacity of the plane-rotation stream (float2 entries), multiple of 1024;
 *            0 selects the default 2*d*d rounded up.
 * The ZLayer batch mean (admm_net.py:459) is taken over the B signals of this call, whatever `chunk` is:
 * per-signal state (packed Z, G, phi, h, r: ~83 KB/signal at n=100) is kept for all B signals, the
 * eigen-solver scratch (~200 KB/signal) only for `chunk` signals at a time (chunk <= 0: chunk = B).
 * ------------------------------------------------------------------------------------------- */
int admmnet_forward_workspace_bytes(int B, int chunk, int n, int K, int rcap, size_t* bytes);
int admmnet_forward(const void* y, const void* b, const float* sigma, int B, int chunk, int Mdim, int Ndim, int K,
                    const float* params, void* phi_out, void* ws, size_t ws_bytes, int rcap, void* stream);

/* Split-phase variant (what admmnet_forward loops over).  For exact whole-batch semantics across GPUs
 * (SURVEY.md §8e): for each layer k < K-1 run admmnet_layer_chunk over the local chunks, then
 * admmnet_layer_rsum(k), all-reduce *rsum[k] over the ranks, admmnet_set_mean(k, global count);
 * finally admmnet_final_phi.  y/b/sigma are the full local arrays; sig_off/Bc select the chunk.     */
int admmnet_layer_chunk(const void* y, const void* b, const float* sigma, int B, int chunk, int sig_off, int Bc,
                        int Mdim, int Ndim, int K, int k, const float* params, void* ws, size_t ws_bytes, int rcap,
                        void* stream);
int admmnet_layer_rsum(void* ws, size_t ws_bytes, int B, int chunk, int n, int K, int rcap, int k, void* stream);
/* One whole layer of the local batch the way admmnet_forward runs it: all chunks fanned out over the chunk lanes
 * (forked from and joined back into `stream`), followed by admmnet_layer_rsum(k).  The multi-GPU 'global' norm scope
 * (sharding.py) calls this, all-reduces rsum[k] on `stream`, then admmnet_set_mean — no host synchronisation.  */
int admmnet_layer(const void* y, const void* b, const float* sigma, int B, int chunk, int Mdim, int Ndim, int K, int k,
                  const float* params, void* ws, size_t ws_bytes, int rcap, void* stream);
int admmnet_ws_scalars(void* ws, size_t ws_bytes, int B, int chunk, int n, int K, int rcap, double** rsum,
                       float** mean, int** status);
int admmnet_set_mean(void* ws, size_t ws_bytes, int B, int chunk, int n, int K, int rcap, int k, double count,
                     void* stream);
int admmnet_final_phi(const void* y, const void* b, int B, int chunk, int Mdim, int Ndim, int K, const float* params,
                      void* phi_out, void* ws, size_t ws_bytes, int rcap, void* stream);
/* clears the status word; call once before layer 0 when driving the split-phase API by hand */
int admmnet_reset_status(void* ws, size_t ws_bytes, int B, int chunk, int n, int K, int rcap, void* stream);
/* copies the status word to the host (synchronises the stream) */
int admmnet_status(void* ws, size_t ws_bytes, int B, int chunk, int n, int K, int rcap, void* stream,
                   int* status_host);

/* ---------------------------------------------------------------------------------------------
 * Unit taps of the eigen-solver (replaces torch.linalg.eigh + the two bmm of admm_net.py:303,349)
 *   A        complex64 [B][d][d] row-major, lower triangle read, d <= 128
 *   evals    float32 [B][d]   (QL order, not sorted)
 *   evecs    complex64 [B][d][d] row-major, column i pairs with evals[i]       (may be NULL)
 *   fn_out   complex64 [B][d(d+1)/2] packed lower (row-major) f(A) = U diag(f(l)) U^H (may be NULL)
 *   params   one layer record: f = the learned eigenvalue map (admm_net.py:310-334); NULL: f = id
 * ------------------------------------------------------------------------------------------- */
int admmnet_eigh_workspace_bytes(int B, int d, int rcap, size_t* bytes);
int admmnet_eigh_batched(const void* A, int B, int d, float* evals, void* evecs, void* fn_out, const float* params,
                         void* ws, size_t ws_bytes, int rcap, void* stream, int* status_dev);

/* Unit tap of the layer-0 shortcut: eigen-decomposition of the arrowhead matrices
 * [[diag(h), phi],[phi^H, c0]] that admm_net.py:286-303 hands to eigh when G = Z = 0.
 *   h float32 [B][n], phi complex64 [B][n], c0 float32 [B]
 *   evals float32 [B][n+1] ascending; evecs complex64 [B][n+1][n+1] row-major (may be NULL)
 *   handled int [B]: 0 = declined (poles closer than rounding or a vanishing |phi_i|; admmnet_forward sends
 *   such signals through the general eigen-solver), outputs then undefined                                */
int admmnet_arrow_eigh(const float* h, const void* phi, const float* c0, int B, int n, float* evals, void* evecs,
                       int* handled, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Classical ADMM  (replaces admm.py:6-114 admm_for_us; see SURVEY.md App. A.3 for why the solver
 * is the linear recursion phi_k = M^-1 (y/b + rho phi_{k-1}) for every reachable input)
 *   y, b     complex128 (in_is_c128=1) or complex64 [B][n], n <= 256
 *   phi_out  complex128 [B][n]
 * ------------------------------------------------------------------------------------------- */
int admm_classic_forward(const void* y, const void* b, int in_is_c128, int B, int n, double rho, int n_iter,
                         void* phi_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Peak search  (replaces utils/peakSearchUtils.py:9-60 peak_search[_func] and 63-173 alt_peak_search)
 *   phi       complex64/complex128 [B][xbase*ybase]
 *   axis_x/y  float64 coarse grid axes (np.arange(xmin, xmax-xstep, xstep), np.arange(ymin, ymax-xstep, ystep))
 *   peaks     float64 [B][pmax][3] rows (x, y, height) in row-major discovery order; count[B] = #maxima
 *   top       float64 [B][topl][3] best topl rows by height (stable), zero padded; topl = 0 skips it
 *   surface   optional float64 [B][Gy][Gx]
 * ------------------------------------------------------------------------------------------- */
int peak_search_full(const void* phi, int phi_is_c128, int B, int xbase, int ybase, const double* axis_x, int Gx,
                     const double* axis_y, int Gy, double xmin, double xmax, double xstep, double ymin, double ymax,
                     double ystep, double reducefactor, int iters, int pmax, double* peaks, int* count, int topl,
                     double* top, double* surface, int* status_dev, void* stream);
/* spectrum of ONE phi at npts arbitrary (x,y) points: peak_search(phi, X, x_base, Y, y_base) */
int peak_search_points(const void* phi, int phi_is_c128, int xbase, int ybase, const double* X, const double* Y,
                       int npts, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ADMMNet's learned regression head  (replaces PeakSearchLayer.forward, admm_net.py:570-630, eval mode)
 *   phi           complex64 [B][n], n = M*N <= 128
 *   head_params   float32 [admmnet_head_param_count(n, L)], packed by admm-net_b200/params.py::pack_head
 *   tau, f, conf  float32 [B][L]
 * ------------------------------------------------------------------------------------------- */
int admmnet_head_param_count(int n, int L);
int admmnet_peak_head(const void* phi, int B, int n, int L, const float* head_params, float* tau, float* f,
                      float* conf, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Synthetic inputs on the device (replaces the per-sample Python loop of generate_data.py:133-221 as the input
 * generator of benchmarks; counter-based RNG keyed by (seed, signal index))
 *   y, b   complex64 [B][Nb*Nd], sigma float32 [B]; truth optional float64 [B][L][4] = (tau, f, Re C, Im C)
 * ------------------------------------------------------------------------------------------- */
int admmnet_generate(void* y, void* b, float* sigma, double* truth, int B, int Nb, int Nd, int L, double snr_w_db,
                     double snr_demod_db, unsigned long long seed, void* stream);
/* dataset variant (DatasetGeneratorCreatePhi, generate_data.py:410-463): the noise SNR of every signal is drawn
 * uniformly from [snr_w_lo_db, snr_w_hi_db) and the symbol error rate (percent) is written to ser[B] (optional). */
int admmnet_generate_dataset(void* y, void* b, float* sigma, double* truth, float* ser, int B, int Nb, int Nd, int L,
                             double snr_w_lo_db, double snr_w_hi_db, double snr_demod_db, unsigned long long seed,
                             void* stream);

/* ---------------------------------------------------------------------------------------------
 * Measurement hooks (bench.py): per-kernel CUDA-event timing of the launches issued between
 * admmnet_profile_begin and admmnet_profile_end (summed ms and launch count per kernel kind), and an
 * FP32-FMA peak micro-kernel (roofline denominator of the FP32-pipe-bound eigen-solver kernels).
 * Process-global, not thread safe.
 * ------------------------------------------------------------------------------------------- */
int admmnet_profile_begin(void);
int admmnet_profile_kinds(void);
const char* admmnet_profile_kind_name(int kind);
int admmnet_profile_end(double* ms, long long* launches);
int admmnet_fp32_peak_launch(float* out, int grid, int iters, double* flops, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Debug tap of the tcgen05 / TMEM / TMA layer (csrc/tc.cuh) the tail kernel is built on: one 128 x N x K tf32 tile,
 * out = D0 + (+-A) * B^T, A [128][K], B [N][K], D0/out [128][N] row-major fp32 on the device.
 * flags: 1 A from tensor memory, 2 negate A, 16 operands loaded by TMA, 32 3xTF32 split (arbitrary fp32 inputs,
 * fp32-class result), 64 accumulator at TMEM column 8.  No reference counterpart (unit test of the machinery).
 * ------------------------------------------------------------------------------------------- */
/* Debug tap of the local-maximum stage (skimage.morphology.local_maxima(Z, connectivity=2), peakSearchUtils.py:118-119:
 * 8-connected, plateau-aware, borders allowed, constant image has none) on a caller-supplied surface [B][Gy][Gx]:
 * peaks[B][pmax][3] = (axis_x[ix], axis_y[iy], value) in row-major discovery order, count[B] the number found. */
int peak_surface_maxima(const double* surface, int B, const double* axis_x, int Gx, const double* axis_y, int Gy,
                        int pmax, double* peaks, int* count, int* status_dev, void* stream);

/* shared-memory bytes of the tensor-core tail kernel (k_tail_tc) for matrix order d; -1: d outside its range
 * (33..104, the SIMT tail kernels serve), 0: switched off with ADMMNET_TAILTC=0 */
int admmnet_tail_tc_smem_bytes(int d);
/* tcgen05.mma flops k_tail_tc issues per signal of matrix order d (all 3xTF32 split terms; bench.py's tensor-pipe
 * figure), 0 when the kernel does not serve d */
double admmnet_tail_tc_mma_flops(int d);
/* tuning aid: with ADMMNET_TC_PROF=1 in the environment k_tail_tc's CTA 0 accumulates clock cycles per phase
 * (13 counters, csrc/tail_tc.cu enum TcPhase; the last one counts signals); reads and clears them. */
int admmnet_tail_tc_profile_read(long long* host16);
/* tuning aid: with ADMMNET_DC_PROF=1 k_dc's CTA 0 accumulates clock cycles per (level, phase) of the divide & conquer
 * tridiagonal solver (csrc/dc_kernels.cu: counters 8*level+phase, secular iterations/warps at 72+level / 80+level,
 * signals at 96); reads and clears the 128 counters. */
int admmnet_dc_profile_read(long long* host128);
int admmnet_tc_gemm_probe(const float* A, const float* B, const float* D0, int N, int K, int flags, float* out,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif
