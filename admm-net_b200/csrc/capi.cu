// extern "C" boundary (include/admmnet_b200.h): argument checks, workspace carving, launches.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/admmnet_b200.h"
#include "net_kernels.cu"
#include "arrow_kernels.cu"
#include "classic_kernels.cu"
#include "peak_kernels.cu"
#include "gen_kernels.cu"
#include "head_kernels.cu"
#include "tc_probe.cu"
#include "tail_tc.cu"
#include "big_kernels.cu"
#include "dc_kernels.cu"

using namespace admmnet;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(ADMMNET_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
    } while (0)

// Dynamic shared-memory opt-in, raised monotonically per (device, kernel) under a mutex: two host threads launching
// the same kernel with different footprints (k_head2's stages, two matrix orders) must never LOWER the limit between
// the other thread's cudaFuncSetAttribute and its launch ("invalid argument" at launch).
template <typename F>
static cudaError_t ensure_smem(F* func, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, int> cur;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    int& have = cur[std::make_pair(dev, (const void*)func)];
    if (bytes <= have) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

// ------------------------------------------------------------------------------------ profiling hooks
// Optional per-kernel CUDA-event timing on the launching stream (bench.py's roofline / launch count).
// Process-global (guarded by a mutex); off by default (then the only cost is one branch per launch).
namespace prof {
enum Kind { HEAD = 0, QL, ROT, TAIL, MISC, CLASSIC, PEAK, HEAD2, MERGE, ARROW, BIG, DC, NKINDS };
static const char* kNames[NKINDS] = {"k_head", "k_ql", "k_rot", "k_tail", "misc", "k_classic", "k_peak_search", "k_head2", "k_merge", "k_arrow", "k_big_layer", "k_dc"};
struct Rec { int kind; cudaEvent_t a, b; };
static bool on = false;
static std::mutex mu;               // launches may come from several host threads while profiling is on
static std::vector<Rec> recs;
static std::vector<cudaEvent_t> pool;
static size_t pool_used = 0;
static cudaEvent_t get_event() {
    if (pool_used == pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        pool.push_back(e);
    }
    return pool[pool_used++];
}
struct Scope {
    cudaStream_t st;
    bool active;
    cudaEvent_t eb = nullptr;
    Scope(int kind, cudaStream_t s) : st(s), active(on) {
        if (active) {
            std::lock_guard<std::mutex> lock(mu);
            Rec r{kind, get_event(), get_event()};
            eb = r.b;
            cudaEventRecord(r.a, st);
            recs.push_back(r);
        }
    }
    ~Scope() {
        if (active) cudaEventRecord(eb, st);
    }
};
}  // namespace prof

extern "C" int admmnet_profile_begin(void) {
    prof::recs.clear();
    prof::pool_used = 0;
    prof::on = true;
    return 0;
}
extern "C" int admmnet_profile_kinds(void) { return prof::NKINDS; }
extern "C" const char* admmnet_profile_kind_name(int kind) {
    return (kind >= 0 && kind < prof::NKINDS) ? prof::kNames[kind] : "";
}
// ms[kind] = summed device time of that kind's launches since admmnet_profile_begin, launches[kind] = count
extern "C" int admmnet_profile_end(double* ms, long long* launches) {
    prof::on = false;
    if (!ms || !launches) return fail(ADMMNET_ERR_ARG, "null pointer");
    for (int k = 0; k < prof::NKINDS; ++k) { ms[k] = 0.0; launches[k] = 0; }
    for (auto& r : prof::recs) {
        CK(cudaEventSynchronize(r.b));
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.kind] += t;
        launches[r.kind] += 1;
    }
    prof::recs.clear();
    prof::pool_used = 0;
    return 0;
}

extern "C" const char* admmnet_last_error(void) { return g_err.c_str(); }
extern "C" int admmnet_version(void) { return 100; }
extern "C" int admmnet_param_stride(int n) { return param_stride(n); }

static long long* tc_prof_buffer();

// ------------------------------------------------------------------------------------ workspace
namespace {
struct Ws {
    float2 *Zp, *GV, *rot, *tau, *phi_cur, *Ttr;
    float *Zr, *Zr2, *lam, *dT, *eT, *h_cur, *r, *mean;
    double* rho;
    int *nrot, *status;
    int* handled;      // [chunk] per slot: 1 = layer-0 signal solved by k_arrow
    float2* big;       // d > 128 only: Jacobi scratch [slot][big_grid][2 d^2] (csrc/big_kernels.cu)
    double* rsum;
    size_t bytes;
};
inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
constexpr int NSLOT_MAX = 8;
// number of chunk lanes = scratch slots (chunks of one layer in flight); ADMMNET_NSLOT overrides (1..8)
static int nslot_value() {
    static const int v = [] {
        const char* e = getenv("ADMMNET_NSLOT");
        const int x = e ? atoi(e) : 2;
        return x < 1 ? 1 : (x > NSLOT_MAX ? NSLOT_MAX : x);
    }();
    return v;
}
#define NSLOT (nslot_value())
constexpr int TR_MAX = 88;   // largest trailing block handed to a later tridiagonalisation stage
// orders at which the trailing block is compacted and handed to the next (smaller, higher-occupancy) stage
struct StagePlan { int marks[8]; };
inline const StagePlan& stage_plan() {
    // parsed once; C++11 guarantees the initialisation of a function-local static is thread safe
    static const StagePlan plan = [] {
        StagePlan p = {{80, 64, 48, 32, 0, 0, 0, 0}};
        if (const char* e = getenv("ADMMNET_STAGES")) {   // ADMMNET_STAGES="88,72,56,40" overrides the plan (tuning)
            int n = 0;
            for (const char* q = e; *q && n < 7;) {
                p.marks[n++] = atoi(q);
                while (*q && *q != ',') ++q;
                if (*q == ',') ++q;
            }
            for (; n < 8; ++n) p.marks[n] = 0;
        }
        return p;
    }();
    return plan;
}
inline int next_stage_order(int d) {
    const StagePlan& p = stage_plan();
    for (int i = 0; i < 8 && p.marks[i] > 0; ++i)
        if (p.marks[i] <= TR_MAX && d > p.marks[i] + 8) return p.marks[i];
    return 0;     // finish in this stage
}
// register-resident tridiagonalisation (csrc/trd_reg.cuh): slots per thread, 0 = the shared-memory stages.
// Opt-in (ADMMNET_TRD=1): parity-green, but measured slower on B200 (752 ms against 304 + 333 ms per 131072 x 8
// signal-layers for the staged form: 4 block barriers per Householder step at 2 CTAs per SM).
inline int trd_ns(int d) {
    static const bool on = getenv("ADMMNET_TRD") && atoi(getenv("ADMMNET_TRD")) != 0;
    return !on ? 0 : (d <= 112 ? 7 : 8);
}
inline int stage1_steps(int d) {
    if (trd_ns(d)) return d - 1;
    const int nx = next_stage_order(d);
    return nx ? d - nx : d - 1;
}
inline int default_rcap(int d) { return d > 128 ? 2048 : ((2 * d * d + 2048 + 1023) / 1024) * 1024; }
// d > 128 runs the size-agnostic Jacobi layer kernel (csrc/big_kernels.cu): persistent CTAs, two per SM
inline int big_grid(int C) { return C < 296 ? C : 296; }

// n = signal length (for phi/h), d = matrix order.  Per-signal STATE arrays (Zp, GV, phi_cur, h_cur, r) cover
// the whole batch B; the eigen-solver SCRATCH (Zr, rot, tau, lam, dT, eT, nrot) covers one chunk C <= B.
Ws carve(void* base, int B, int C, int n, int d, int K, int rcap) {
    Ws w;
    unsigned char* p = reinterpret_cast<unsigned char*>(base);
    size_t off = 0;
    const size_t npk = (size_t)d * (d + 1) / 2;
    auto take = [&](size_t nbytes) { unsigned char* q = p + off; off += al(nbytes); return q; };
    w.Zp = (float2*)take(B * npk * sizeof(float2));
    w.GV = (float2*)take(B * npk * sizeof(float2));
    // two scratch slots so that consecutive chunks can be in flight on different streams; the Householder/QL scratch
    // exists for d <= 128 only, the Jacobi scratch for d > 128 only
    const size_t S = d > 128 ? 0 : (size_t)NSLOT * C;
    const size_t zsz = (size_t)d * (4 * ((d + 3) / 4));          // Z^T per signal: d rows of pitch 4*ceil(d/4)
    w.Zr = (float*)take(S * zsz * sizeof(float));
    w.Zr2 = (float*)take(S * zsz * sizeof(float));
    w.rho = (double*)take(S * DC_MAXTEAR * sizeof(double));
    w.rot = (float2*)take(S * rcap * sizeof(float2));
    w.tau = (float2*)take(S * d * sizeof(float2));
    w.Ttr = (float2*)take(S * 2 * TR_MAX * TR_MAX * sizeof(float2));   // ping-pong pair per slot
    w.lam = (float*)take(S * d * sizeof(float));
    w.dT = (float*)take(S * d * sizeof(float));
    w.eT = (float*)take(S * d * sizeof(float));
    w.big = (float2*)take(d > 128 ? (size_t)NSLOT * big_grid(C) * big_scratch_f2(d) * sizeof(float2) : 0);
    w.phi_cur = (float2*)take((size_t)B * n * sizeof(float2));
    w.h_cur = (float*)take((size_t)B * n * sizeof(float));
    w.r = (float*)take((size_t)B * sizeof(float));
    w.nrot = (int*)take((size_t)NSLOT * C * sizeof(int));
    w.handled = (int*)take((size_t)NSLOT * C * sizeof(int));
    w.rsum = (double*)take((size_t)(K + 1) * sizeof(double));
    w.mean = (float*)take((size_t)(K + 1) * sizeof(float));
    w.status = (int*)take(sizeof(int));
    w.bytes = off;
    return w;
}

int check_net_args(int B, int& chunk, int n, int K, int& rcap) {
    if (B <= 0 || K <= 0) return fail(ADMMNET_ERR_ARG, "B and K must be positive");
    if (chunk <= 0 || chunk > B) chunk = B;
    if (n < 2 || n > BIG_NMAX) return fail(ADMMNET_ERR_ARG, "n = M*N must be in [2,256]");
    if (rcap == 0) rcap = default_rcap(n + 1);
    if (rcap < 2048 || rcap % ROT_CHUNK) return fail(ADMMNET_ERR_ARG, "rcap must be a multiple of 1024, >= 2048");
    return 0;
}

// Legacy divide & conquer plan for the QL form of the tridiagonal solver: L levels -> 2^L blocks solved by k_ql/k_rot
// and glued by k_merge (ADMMNET_DC=L with ADMMNET_DCK=0; default 0 = plain QL).  Parity-green (tests run it), but its
// fp64-flavoured secular solves cost more than the rotations they save (1690 ms against 760 ms per step of 131072
// signals x 9 layers); the fused fp32 kernel k_dc (csrc/dc_kernels.cu) superseded it, also for small batches.
inline int dc_levels(int d, int B) {
    static const int env = getenv("ADMMNET_DC") ? atoi(getenv("ADMMNET_DC")) : -1;
    int L = env < 0 ? 0 : (env > 3 ? 3 : env);
    while (L > 0 && (d >> L) < 8) --L;        // blocks of at least 8
    return L;
}
// Hybrid plan for many-chunk batches (ADMMNET_HYB = levels, default below): k_ql + k_rotf solve the 2^L blocks of the torn
// tridiagonal (half / quarter length sweeps: the latency-bound QL chain and the rotation count shrink accordingly) and
// k_dc runs only the top L merge levels.
// ADMMNET_MIX=1: in a many-chunk forward the odd chunk lanes take k_dc and the even ones the QL pair, so that every
// latency-bound k_ql runs beside a throughput-bound lane (experiment; see DESIGN.md §5 for the measurement).
inline bool mix_lanes() {
    static const bool on = getenv("ADMMNET_MIX") && atoi(getenv("ADMMNET_MIX")) != 0;
    return on;
}
inline int hyb_levels(int d) {
    static const int env = getenv("ADMMNET_HYB") ? atoi(getenv("ADMMNET_HYB")) : 0;
    int L = env < 0 ? 0 : (env > 2 ? 2 : env);
    while (L > 0 && (d >> L) < 16) --L;
    return L;
}
inline TearSpec dc_tears(int d, int B, int hyb = 0) {
    TearSpec ts;
    ts.absconv = hyb > 0 ? 1 : 0;
    const int L = hyb > 0 ? hyb : dc_levels(d, B), nb = 1 << L;
    ts.n = nb - 1;
    for (int q = 0; q < ts.n; ++q) ts.pos[q] = (int)(((long long)(q + 1) * d) / nb);
    return ts;
}

// later tridiagonalisation stages on the compacted trailing block (no-op when stage 1 did everything)
int launch_head2(const Ws& w, int B, int d, cudaStream_t st, const int* skip = nullptr) {
    if (trd_ns(d)) return 0;                 // the register-resident form finishes in k_head
    int dcur = next_stage_order(d);          // order handed over by k_head
    if (!dcur) return 0;
    const size_t tr_sz = (size_t)B * TR_MAX * TR_MAX;
    float2* bufs[2] = {w.Ttr, w.Ttr + tr_sz};
    int which = 0;
    while (dcur) {
        const int nx = next_stage_order(dcur);
        Head2Args h;
        h.Tin = bufs[which]; h.Tout = bufs[which ^ 1]; h.GV = w.GV; h.tau = w.tau; h.dT = w.dT; h.eT = w.eT;
        h.B = B; h.d = d; h.d2 = dcur; h.ld2 = dcur | 1; h.k0 = d - dcur; h.k_stop = nx ? dcur - nx : dcur - 1;
        h.skip = skip;
        prof::Scope pscope(prof::HEAD2, st);
        if (dcur > 64) {
            const size_t sm = head2_smem_bytes<256, 128>(h.d2, h.ld2);
            CK(ensure_smem(k_head2<256, 128>, (int)sm));
            k_head2<256, 128><<<B, 256, sm, st>>>(h);
        } else {
            const size_t sm = head2_smem_bytes<128, 64>(h.d2, h.ld2);
            CK(ensure_smem(k_head2<128, 64>, (int)sm));
            k_head2<128, 64><<<B, 128, sm, st>>>(h);
        }
        CK(cudaGetLastError());
        which ^= 1;
        dcur = nx;
    }
    return 0;
}

// Per-phase clock counters of k_dc (CTA 0, summed over launches) when ADMMNET_DC_PROF=1: a tuning aid.
static long long* dc_prof_buffer() {
    static long long* buf = [] {
        long long* p = nullptr;
        if (getenv("ADMMNET_DC_PROF") && atoi(getenv("ADMMNET_DC_PROF")) != 0) {
            if (cudaMalloc(&p, DCP_N * sizeof(long long)) != cudaSuccess) p = nullptr;
            else cudaMemset(p, 0, DCP_N * sizeof(long long));
        }
        return p;
    }();
    return buf;
}
// the eigen-solver launches after the tridiagonal form is in the workspace
int launch_eig_tail(const Ws& w, int B, int n, int d, int rcap, const float* Pk, int with_c, float2* U_out,
                    float* lamp_out, int* status, cudaStream_t st, cudaStream_t qst = nullptr,
                    cudaEvent_t ev_in = nullptr, cudaEvent_t ev_out = nullptr, const int* skip = nullptr,
                    bool lone = true) {
    // Tridiagonal eigenproblem.  Two forms, both parity-tested (tests/test_gpu_parity.py):
    //   * k_dc (csrc/dc_kernels.cu): fused fp32 divide & conquer, one CTA per signal.  Throughput bound; it has no
    //     serial chain, so it wins whenever the launch stands alone - a call of a single chunk (B <= chunk: latency and
    //     small-batch mode, B = 1: 9.8 -> 4.4 ms for K = 10) and the eigh tap of the training path
    //     (measured on B200, d = 101: 0.12 ms against 0.93 ms at B = 1, 1.0 against 4.9 at B = 2368, 6.9 against 9.4
    //     at B = 16384).
    //   * k_ql (fp64 scalar QL chain, one thread per signal, pure latency: 4.1 ms whatever the batch) + k_rotf.  In a
    //     many-chunk forward k_ql runs on the priority lane underneath the other chunk's kernels, which makes this
    //     pair 1 % faster end to end (84.1k against 83.1k signals/s), so chunks of a larger batch keep it.
    // ADMMNET_DCK = 1 / 0 forces one form; ADMMNET_DC = 1..3 adds k_merge levels to the QL form (legacy).
    static const int dck_env = getenv("ADMMNET_DCK") ? atoi(getenv("ADMMNET_DCK")) : -1;
    const bool use_dck = dck_env >= 0 ? dck_env != 0 : lone;
    const int hyb = use_dck ? 0 : hyb_levels(d);
    const float* zfinal = w.Zr;
    auto launch_dc = [&](int nhyb, const TearSpec& ts) -> int {
        DcArgs da;
        da.dT = w.dT; da.eT = w.eT; da.lam = w.lam; da.Zt = w.Zr; da.status = status; da.skip = skip;
        da.prof = dc_prof_buffer();
        da.B = B; da.d = d; da.ldz = 4 * ((d + 3) / 4);
        da.beta_in = w.rho; da.nhyb = nhyb; da.ntear = nhyb ? ts.n : 0; da.tear_stride = DC_MAXTEAR;
        for (int q = 0; q < 7; ++q) da.tear_pos[q] = q < ts.n ? ts.pos[q] : 1;
        const size_t sm = dc_smem_bytes(d, da.ldz);
        CK(ensure_smem(k_dc, (int)sm));
        prof::Scope pscope(prof::DC, st);
        k_dc<<<B, DCK_NT, sm, st>>>(da);
        CK(cudaGetLastError());
        return 0;
    };
    if (use_dck) {
        TearSpec none; none.n = 0; none.absconv = 0;
        if (int e = launch_dc(0, none)) return e;
    } else {
    const bool side = qst != nullptr;
    if (side) {   // k_ql on a high-priority side stream: it is latency bound and co-resides with other kernels
        CK(cudaEventRecord(ev_in, st));
        CK(cudaStreamWaitEvent(qst, ev_in, 0));
    }
    {
        cudaStream_t st_main = st;
        cudaStream_t st = side ? qst : st_main;
        const size_t sm = (size_t)2 * d * QL_THREADS * sizeof(double);
        CK(ensure_smem(k_ql, (int)sm));
        prof::Scope pscope(prof::QL, st);
        k_ql<<<(B + QL_THREADS - 1) / QL_THREADS, QL_THREADS, sm, st>>>(w.dT, w.eT, B, d, w.lam, w.rot, rcap, w.nrot,
                                                                         status, dc_tears(d, B, hyb), w.rho, skip);
        CK(cudaGetLastError());
    }
    if (side) {
        CK(cudaEventRecord(ev_out, qst));
        CK(cudaStreamWaitEvent(st, ev_out, 0));
    }
    {
        const size_t sm = (size_t)2 * ROT_STAGE * sizeof(float2) + (size_t)d * (4 * ((d + 3) / 4)) * sizeof(float);
        CK(ensure_smem(k_rot, (int)sm));
        prof::Scope pscope(prof::ROT, st);
        static const bool fused = !(getenv("ADMMNET_ROTF") && atoi(getenv("ADMMNET_ROTF")) == 0);
        // ADMMNET_ROTP=1: the coordinates of a signal split over two one-warp CTAs (k_rotf_p<2, 2>: 7 instead of 4
        // resident warps per SM, half the packed operations per rotation and warp).  Parity-identical, measured 8 %
        // slower (k_rot 276 -> 297 ms per 131072 signals x 8 layers: the per-rotation overhead - parameter loads, ring
        // indexing - is paid by both warps), so it stays opt-in.
        static const bool panels = getenv("ADMMNET_ROTP") && atoi(getenv("ADMMNET_ROTP")) != 0;
        if (fused && panels) {
            const size_t smp = (size_t)2 * ROT_STAGE * sizeof(float2) + (size_t)d * (2 * ((d + 3) / 4)) * sizeof(float);
            CK(ensure_smem(k_rotf_p<2, 2>, (int)smp));
            k_rotf_p<2, 2><<<2 * B, ROT_THREADS, smp, st>>>(w.rot, rcap, w.nrot, d, w.Zr, skip);
        } else if (fused) {      // same shared-memory size: ring of 4 x 256 entries instead of 2 x 512
            CK(ensure_smem(k_rotf, (int)sm));
            k_rotf<<<B, ROT_THREADS, sm, st>>>(w.rot, rcap, w.nrot, d, w.Zr, skip);
        } else
        k_rot<<<B, ROT_THREADS, sm, st>>>(w.rot, rcap, w.nrot, d, w.Zr, skip);
        CK(cudaGetLastError());
    }
    if (hyb > 0) {
        // the top merge levels in the fused kernel (in place: it stages Z^T in shared memory first)
        if (int e = launch_dc(hyb, dc_tears(d, B, hyb))) return e;
    } else {
        // legacy divide & conquer merges, leaves -> root; Z ping-pongs between the two scratch buffers
        const TearSpec ts = dc_tears(d, B);
        const int L = dc_levels(d, B);
        float* zb[2] = {w.Zr, w.Zr2};
        int cur = 0;
        for (int lev = 1; lev <= L; ++lev) {
            const int half = 1 << (lev - 1), step = 1 << lev;
            // tears handled at this level: 1-based number t with t % step == half; its range is bounded by the
            // tears t-half and t+half.  The ranges of a level partition [0,d): one launch handles all (<= 4).
            MergeArgs m;
            m.Zin = zb[cur]; m.Zout = zb[cur ^ 1]; m.lam = w.lam; m.rho = w.rho; m.status = status;
            m.B = B; m.d = d; m.nr = 0; m.skip = skip;
            for (int i = 0; i < ts.n; ++i) {
                if ((i + 1) % step != half) continue;
                const int lo = (i + 1) - half, hi = (i + 1) + half;
                m.ra[m.nr] = lo == 0 ? 0 : ts.pos[lo - 1];
                m.rb[m.nr] = hi > ts.n ? d : ts.pos[hi - 1];
                m.rp[m.nr] = ts.pos[i];
                m.rt[m.nr] = i;
                ++m.nr;
            }
            const size_t smm = merge_smem_bytes(d);
            CK(ensure_smem(k_merge, (int)smm));
            prof::Scope pscope(prof::MERGE, st);
            k_merge<<<B, MG_THREADS, smm, st>>>(m);
            CK(cudaGetLastError());
            cur ^= 1;
        }
        zfinal = zb[cur];
    }
    }
    {
        TailArgs t;
        t.Zr = zfinal; t.GV = w.GV; t.tau = w.tau; t.lam = w.lam; t.phi_cur = w.phi_cur; t.h_cur = w.h_cur;
        t.Pk = Pk; t.r_out = w.r; t.U_out = U_out; t.lamp_out = lamp_out;
        t.B = B; t.n = n; t.d = d; t.ldu = 4 * ((d + 3) / 4); t.with_c = with_c; t.skip = skip;
        prof::Scope pscope(prof::TAIL, st);
        // tensor-core form (tcgen05/TMEM, Z^T by TMA): production sizes, no eigenvector tap, positive eigenvalue map
        static const bool use_tc = !(getenv("ADMMNET_TAILTC") && atoi(getenv("ADMMNET_TAILTC")) == 0);
        TailTcPlan plan;
        if (use_tc && with_c >= 0 && Pk && !U_out && !lamp_out && !skip && tail_tc_plan(d, plan)) {
            TailTcArgs ta;
            ta.Zr = zfinal; ta.GV = w.GV; ta.tau = w.tau; ta.lam = w.lam; ta.phi_cur = w.phi_cur; ta.h_cur = w.h_cur;
            ta.Pk = Pk; ta.r_out = w.r; ta.B = B; ta.n = n; ta.with_c = with_c; ta.plan = plan;
            ta.prof = tc_prof_buffer();
            CUtensorMap tmZ;
            if (!make_tmap_2d_f32(&tmZ, zfinal, (uint64_t)B * d, plan.ldz, (uint64_t)plan.ldz * 4, d, plan.ldz))
                return fail(ADMMNET_ERR_CUDA, "cuTensorMapEncodeTiled failed for the Z^T scratch");
            int dev = 0, nsm = 0;
            CK(cudaGetDevice(&dev));
            CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
            CK(ensure_smem(k_tail_tc, plan.total));
            k_tail_tc<<<B < nsm ? B : nsm, TC_THREADS, plan.total, st>>>(ta, tmZ);
            CK(cudaGetLastError());
            return 0;
        }
        static const bool persistent = !(getenv("ADMMNET_TAILP") && atoi(getenv("ADMMNET_TAILP")) == 0);
        const size_t smp = tailp_smem_bytes(d, t.ldu);
        if (persistent && d <= 104 && with_c > 0 && !U_out && !lamp_out && !skip && smp <= 227 * 1024) {
            // production form: persistent CTAs (one per SM) with the next signal's operands prefetched
            int dev = 0, nsm = 0;
            CK(cudaGetDevice(&dev));
            CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
            CK(ensure_smem(k_tail_p<13, 416, 8>, (int)smp));
            k_tail_p<13, 416, 8><<<B < nsm ? B : nsm, 416, smp, st>>>(t);
            CK(cudaGetLastError());
            return 0;
        }
        const size_t sm = tail_smem_bytes(d, t.ldu);
        // (a 16-lane row split, k_tail<7, 832, 16>, doubles the warps per SM but measured 25 % slower)
        if (d <= 104) {
            CK(ensure_smem(k_tail<13, 416, 8>, (int)sm));
            k_tail<13, 416, 8><<<B, 416, sm, st>>>(t);
        } else {
            CK(ensure_smem(k_tail<16, 512, 8>, (int)sm));
            k_tail<16, 512, 8><<<B, 512, sm, st>>>(t);
        }
        CK(cudaGetLastError());
    }
    return 0;
}
}  // namespace

// view of the workspace for the chunk [off, off+Bc): state pointers advanced, scratch untouched
static Ws chunk_view(const Ws& w, int off, int n, int d, int slot = 0, int C = 0, int rcap = 0) {
    Ws c = w;
    const size_t npk = (size_t)d * (d + 1) / 2;
    c.Zr += (size_t)slot * C * d * (4 * ((d + 3) / 4));
    c.Zr2 += (size_t)slot * C * d * (4 * ((d + 3) / 4));
    c.rho += (size_t)slot * C * DC_MAXTEAR;
    c.rot += (size_t)slot * C * rcap;
    c.tau += (size_t)slot * C * d;
    c.Ttr += (size_t)slot * C * 2 * TR_MAX * TR_MAX;
    c.lam += (size_t)slot * C * d;
    c.dT += (size_t)slot * C * d;
    c.eT += (size_t)slot * C * d;
    c.nrot += (size_t)slot * C;
    c.handled += (size_t)slot * C;
    if (d > 128) c.big += (size_t)slot * big_grid(C) * big_scratch_f2(d);
    c.Zp += (size_t)off * npk;
    c.GV += (size_t)off * npk;
    c.phi_cur += (size_t)off * n;
    c.h_cur += (size_t)off * n;
    c.r += off;
    return c;
}

extern "C" int admmnet_forward_workspace_bytes(int B, int chunk, int n, int K, int rcap, size_t* bytes) {
    if (!bytes) return fail(ADMMNET_ERR_ARG, "bytes is NULL");
    if (int e = check_net_args(B, chunk, n, K, rcap)) return e;
    *bytes = carve(nullptr, B, chunk, n, n + 1, K, rcap).bytes;
    return 0;
}

extern "C" int admmnet_ws_scalars(void* ws, size_t ws_bytes, int B, int chunk, int n, int K, int rcap, double** rsum,
                                  float** mean, int** status) {
    if (int e = check_net_args(B, chunk, n, K, rcap)) return e;
    Ws w = carve(ws, B, chunk, n, n + 1, K, rcap);
    if (!ws || ws_bytes < w.bytes) return fail(ADMMNET_ERR_WORKSPACE, "workspace too small");
    if (rsum) *rsum = w.rsum;
    if (mean) *mean = w.mean;
    if (status) *status = w.status;
    return 0;
}

// Internal streams for admmnet_forward: chunk lanes (consecutive chunks of a layer overlap) and one
// high-priority side stream per lane for the latency-bound k_ql.  One set per (device, caller stream), created
// lazily under a mutex: calls from different host threads on different streams (with their own workspaces) never
// share a stream or an event, so admmnet_forward is re-entrant across streams as the header states.  Two threads
// enqueueing on the SAME stream would interleave their launches and is not supported (as for any CUDA library).
namespace {
struct Lanes {
    cudaStream_t L[NSLOT_MAX], Q[NSLOT_MAX];
    cudaEvent_t evH[NSLOT_MAX], evQ[NSLOT_MAX], evDone[NSLOT_MAX], evStart, evMean, evEnd;
};
std::mutex g_lanes_mu;
std::map<std::pair<int, cudaStream_t>, Lanes*> g_lanes;
int get_lanes(cudaStream_t caller, Lanes** out) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_lanes_mu);
    auto key = std::make_pair(dev, caller);
    auto it = g_lanes.find(key);
    if (it != g_lanes.end()) { *out = it->second; return 0; }
    Lanes* l = new Lanes();
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (int i = 0; i < NSLOT; ++i) {
        CK(cudaStreamCreateWithPriority(&l->L[i], cudaStreamNonBlocking, lo));
        CK(cudaStreamCreateWithPriority(&l->Q[i], cudaStreamNonBlocking, hi));
        CK(cudaEventCreateWithFlags(&l->evH[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&l->evQ[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&l->evDone[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&l->evStart, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&l->evMean, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&l->evEnd, cudaEventDisableTiming));
    g_lanes[key] = l;
    *out = l;
    return 0;
}
}  // namespace

static int layer_chunk_impl(const void* y, const void* b, const float* sigma, int B, int chunk, int sig_off, int Bc,
                            int Mdim, int Ndim, int K, int k, const float* params, void* ws, size_t ws_bytes, int rcap,
                            cudaStream_t st, int slot, cudaStream_t qst, cudaEvent_t ev_in, cudaEvent_t ev_out);

extern "C" int admmnet_layer_chunk(const void* y, const void* b, const float* sigma, int B, int chunk, int sig_off,
                                   int Bc, int Mdim, int Ndim, int K, int k, const float* params, void* ws,
                                   size_t ws_bytes, int rcap, void* stream) {
    return layer_chunk_impl(y, b, sigma, B, chunk, sig_off, Bc, Mdim, Ndim, K, k, params, ws, ws_bytes, rcap,
                            (cudaStream_t)stream, 0, nullptr, nullptr, nullptr);
}

static int layer_chunk_impl(const void* y, const void* b, const float* sigma, int B, int chunk, int sig_off, int Bc,
                            int Mdim, int Ndim, int K, int k, const float* params, void* ws, size_t ws_bytes, int rcap,
                            cudaStream_t st, int slot, cudaStream_t qst, cudaEvent_t ev_in, cudaEvent_t ev_out) {
    const int n = Mdim * Ndim, d = n + 1;
    if (int e = check_net_args(B, chunk, n, K, rcap)) return e;
    if (!y || !b || !sigma || !params) return fail(ADMMNET_ERR_ARG, "null input pointer");
    if (k < 0 || k >= K - 1) return fail(ADMMNET_ERR_ARG, "admmnet_layer_chunk: k must be in [0, K-2]");
    if (sig_off < 0 || Bc <= 0 || Bc > chunk || sig_off + Bc > B) return fail(ADMMNET_ERR_ARG, "bad chunk range");
    Ws wf = carve(ws, B, chunk, n, d, K, rcap);
    if (!ws || ws_bytes < wf.bytes) return fail(ADMMNET_ERR_WORKSPACE, "workspace too small");
    Ws w = chunk_view(wf, sig_off, n, d, slot, chunk, rcap);
    const int ps = param_stride(n);
    if (d > 128) {
        // size-agnostic layer (cfg 4: n = 144, 196, 256): one persistent kernel per chunk, csrc/big_kernels.cu
        BigArgs g;
        g.y = (const float2*)y + (size_t)sig_off * n; g.b = (const float2*)b + (size_t)sig_off * n; g.sigma = sigma + sig_off;
        g.Zp = w.Zp; g.GV = w.GV; g.phi_cur = w.phi_cur; g.h_cur = w.h_cur; g.r = w.r;
        g.mean_prev = k > 0 ? w.mean + (k - 1) : w.mean;
        g.Pk = params + (size_t)k * ps; g.Pkm1 = k > 0 ? params + (size_t)(k - 1) * ps : params;
        g.scratch = w.big; g.status = w.status; g.B = Bc; g.n = n; g.d = d; g.first = (k == 0);
        prof::Scope pscope(prof::BIG, st);
        k_big_layer<<<big_grid(Bc), BIG_NT, 0, st>>>(g);
        CK(cudaGetLastError());
        return 0;
    }
    HeadArgs h;
    h.y = (const float2*)y + (size_t)sig_off * n; h.b = (const float2*)b + (size_t)sig_off * n; h.sigma = sigma + sig_off;
    h.Zp = w.Zp; h.GV = w.GV; h.phi_cur = w.phi_cur; h.h_cur = w.h_cur; h.r_prev = w.r;
    h.mean_prev = k > 0 ? w.mean + (k - 1) : w.mean;
    h.Pk = params + (size_t)k * ps; h.Pkm1 = k > 0 ? params + (size_t)(k - 1) * ps : params;
    h.tau = w.tau; h.dT = w.dT; h.eT = w.eT;
    h.B = Bc; h.n = n; h.d = d; h.ld = d | 1; h.first = (k == 0);
    h.Ttr = w.Ttr; h.k1 = stage1_steps(d);
    // layer 0: the matrix is an arrowhead; k_arrow solves it directly and the general pipeline only sees the
    // signals it declined (handled[sig] == 0).  ADMMNET_ARROW=0 switches the shortcut off.
    static const bool use_arrow = !(getenv("ADMMNET_ARROW") && atoi(getenv("ADMMNET_ARROW")) == 0);
    const int* skip = nullptr;
    if (h.first && use_arrow) {
        ArrowArgs aa;
        aa.y = h.y; aa.b = h.b; aa.sigma = h.sigma; aa.h_in = nullptr; aa.phi_in = nullptr; aa.c0_in = nullptr;
        aa.Zp = w.Zp; aa.GV = w.GV; aa.phi_cur = w.phi_cur; aa.h_cur = w.h_cur; aa.Pk = h.Pk; aa.r_out = w.r;
        aa.handled = w.handled; aa.lam_out = nullptr; aa.U_out = nullptr;
        aa.B = Bc; aa.n = n; aa.d = d; aa.ldu = 4 * ((d + 3) / 4);
        const size_t sma = arrow_smem_bytes(d, aa.ldu);
        CK(ensure_smem(k_arrow, (int)sma));
        prof::Scope pscope(prof::ARROW, st);
        k_arrow<<<Bc, AR_NT, sma, st>>>(aa);
        CK(cudaGetLastError());
        skip = w.handled;
    }
    h.skip = skip;
    const size_t sm = head_smem_bytes(d, h.ld);
    {
        const int ns = trd_ns(d);
        CK(ensure_smem(ns == 7 ? k_head<7> : ns == 8 ? k_head<8> : k_head<0>, (int)sm));
        prof::Scope pscope(prof::HEAD, st);
        if (ns == 7) k_head<7><<<Bc, 256, sm, st>>>(h);
        else if (ns == 8) k_head<8><<<Bc, 256, sm, st>>>(h);
        else k_head<0><<<Bc, 256, sm, st>>>(h);
    }
    CK(cudaGetLastError());
    if (int e = launch_head2(w, Bc, d, st, skip)) return e;
    return launch_eig_tail(w, Bc, n, d, rcap, h.Pk, 1, nullptr, nullptr, w.status, st, qst, ev_in, ev_out, skip,
                           /*lone=*/B <= chunk || (mix_lanes() && (slot & 1)));
}

extern "C" int admmnet_reset_status(void* ws, size_t ws_bytes, int B, int chunk, int n, int K, int rcap, void* stream) {
    if (int e = check_net_args(B, chunk, n, K, rcap)) return e;
    Ws w = carve(ws, B, chunk, n, n + 1, K, rcap);
    if (!ws || ws_bytes < w.bytes) return fail(ADMMNET_ERR_WORKSPACE, "workspace too small");
    CK(cudaMemsetAsync(w.status, 0, sizeof(int), (cudaStream_t)stream));
    return 0;
}

extern "C" int admmnet_layer_rsum(void* ws, size_t ws_bytes, int B, int chunk, int n, int K, int rcap, int k,
                                  void* stream) {
    if (int e = check_net_args(B, chunk, n, K, rcap)) return e;
    if (k < 0 || k >= K) return fail(ADMMNET_ERR_ARG, "k out of range");
    Ws w = carve(ws, B, chunk, n, n + 1, K, rcap);
    if (!ws || ws_bytes < w.bytes) return fail(ADMMNET_ERR_WORKSPACE, "workspace too small");
    prof::Scope pscope(prof::MISC, (cudaStream_t)stream);
    k_rsum<<<1, 1024, 0, (cudaStream_t)stream>>>(w.r, B, w.rsum + k);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int admmnet_set_mean(void* ws, size_t ws_bytes, int B, int chunk, int n, int K, int rcap, int k,
                                double count, void* stream) {
    if (int e = check_net_args(B, chunk, n, K, rcap)) return e;
    if (k < 0 || k >= K) return fail(ADMMNET_ERR_ARG, "k out of range");
    Ws w = carve(ws, B, chunk, n, n + 1, K, rcap);
    if (!ws || ws_bytes < w.bytes) return fail(ADMMNET_ERR_WORKSPACE, "workspace too small");
    prof::Scope pscope(prof::MISC, (cudaStream_t)stream);
    k_mean_from_sum<<<1, 32, 0, (cudaStream_t)stream>>>(w.rsum + k, count, w.mean + k);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int admmnet_final_phi(const void* y, const void* b, int B, int chunk, int Mdim, int Ndim, int K,
                                 const float* params, void* phi_out, void* ws, size_t ws_bytes, int rcap, void* stream) {
    const int n = Mdim * Ndim, d = n + 1;
    if (int e = check_net_args(B, chunk, n, K, rcap)) return e;
    if (!y || !b || !params || !phi_out) return fail(ADMMNET_ERR_ARG, "null pointer");
    Ws w = carve(ws, B, chunk, n, d, K, rcap);
    if (!ws || ws_bytes < w.bytes) return fail(ADMMNET_ERR_WORKSPACE, "workspace too small");
    const int ps = param_stride(n);
    FinalArgs f;
    f.y = (const float2*)y; f.b = (const float2*)b; f.Zp = w.Zp; f.GV = w.GV; f.phi_prev = w.phi_cur;
    f.r_prev = w.r; f.mean_prev = K > 1 ? w.mean + (K - 2) : w.mean;
    f.Pk = params + (size_t)(K - 1) * ps; f.Pkm1 = K > 1 ? params + (size_t)(K - 2) * ps : params;
    f.phi_out = (float2*)phi_out; f.B = B; f.n = n; f.d = d; f.first = (K == 1);
    cudaStream_t st = (cudaStream_t)stream;
    const long long nthreads = (long long)B * 32;
    prof::Scope pscope(prof::MISC, st);
    k_final_phi<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(f);
    CK(cudaGetLastError());
    return 0;
}

// all chunks of layer k over the lanes (which must already be ordered after the producer of the layer's inputs),
// joined on lane 0 where the residual norms are summed
static int layer_over_lanes(Lanes* ln, const void* y, const void* b, const float* sigma, int B, int chunk, int Mdim,
                            int Ndim, int K, int k, const float* params, void* ws, size_t ws_bytes, int rcap) {
    const int n = Mdim * Ndim;
    int c = 0;
    for (int off = 0; off < B; off += chunk, ++c) {
        const int Bc = B - off < chunk ? B - off : chunk;
        const int s = c % NSLOT;
        if (int e = layer_chunk_impl(y, b, sigma, B, chunk, off, Bc, Mdim, Ndim, K, k, params, ws, ws_bytes, rcap,
                                     ln->L[s], s, ln->Q[s], ln->evH[s], ln->evQ[s]))
            return e;
    }
    for (int i = 1; i < NSLOT; ++i) {
        CK(cudaEventRecord(ln->evDone[i], ln->L[i]));
        CK(cudaStreamWaitEvent(ln->L[0], ln->evDone[i], 0));
    }
    return admmnet_layer_rsum(ws, ws_bytes, B, chunk, n, K, rcap, k, ln->L[0]);
}

static bool lanes_enabled() {
    static const bool use_lanes = !(getenv("ADMMNET_LANES") && atoi(getenv("ADMMNET_LANES")) == 0);
    return use_lanes;
}

extern "C" int admmnet_layer(const void* y, const void* b, const float* sigma, int B, int chunk, int Mdim, int Ndim,
                             int K, int k, const float* params, void* ws, size_t ws_bytes, int rcap, void* stream) {
    const int n = Mdim * Ndim;
    if (int e = check_net_args(B, chunk, n, K, rcap)) return e;
    cudaStream_t cs = (cudaStream_t)stream;
    if (!lanes_enabled()) {
        for (int off = 0; off < B; off += chunk) {
            const int Bc = B - off < chunk ? B - off : chunk;
            if (int e = admmnet_layer_chunk(y, b, sigma, B, chunk, off, Bc, Mdim, Ndim, K, k, params, ws, ws_bytes, rcap,
                                            stream))
                return e;
        }
        return admmnet_layer_rsum(ws, ws_bytes, B, chunk, n, K, rcap, k, stream);
    }
    Lanes* ln = nullptr;
    if (int e = get_lanes(cs, &ln)) return e;
    CK(cudaEventRecord(ln->evStart, cs));
    for (int i = 0; i < NSLOT; ++i) CK(cudaStreamWaitEvent(ln->L[i], ln->evStart, 0));
    if (int e = layer_over_lanes(ln, y, b, sigma, B, chunk, Mdim, Ndim, K, k, params, ws, ws_bytes, rcap)) return e;
    CK(cudaEventRecord(ln->evEnd, ln->L[0]));
    CK(cudaStreamWaitEvent(cs, ln->evEnd, 0));
    return 0;
}

extern "C" int admmnet_forward(const void* y, const void* b, const float* sigma, int B, int chunk, int Mdim, int Ndim,
                               int K, const float* params, void* phi_out, void* ws, size_t ws_bytes, int rcap,
                               void* stream) {
    const int n = Mdim * Ndim;
    if (int e = check_net_args(B, chunk, n, K, rcap)) return e;
    cudaStream_t cs = (cudaStream_t)stream;
    if (int e = admmnet_reset_status(ws, ws_bytes, B, chunk, n, K, rcap, stream)) return e;
    if (!lanes_enabled()) {      // single-stream variant (profiling): every launch on the caller's stream
        for (int k = 0; k < K - 1; ++k) {
            if (int e = admmnet_layer(y, b, sigma, B, chunk, Mdim, Ndim, K, k, params, ws, ws_bytes, rcap, stream))
                return e;
            if (int e = admmnet_set_mean(ws, ws_bytes, B, chunk, n, K, rcap, k, (double)B, stream)) return e;
        }
        return admmnet_final_phi(y, b, B, chunk, Mdim, Ndim, K, params, phi_out, ws, ws_bytes, rcap, stream);
    }
    Lanes* ln = nullptr;
    if (int e = get_lanes(cs, &ln)) return e;
    // fork: the lanes start after everything already queued on the caller's stream
    CK(cudaEventRecord(ln->evStart, cs));
    for (int i = 0; i < NSLOT; ++i) CK(cudaStreamWaitEvent(ln->L[i], ln->evStart, 0));
    for (int k = 0; k < K - 1; ++k) {
        if (int e = layer_over_lanes(ln, y, b, sigma, B, chunk, Mdim, Ndim, K, k, params, ws, ws_bytes, rcap)) return e;
        // the mean is finalised on lane 0; release the other lanes
        if (int e = admmnet_set_mean(ws, ws_bytes, B, chunk, n, K, rcap, k, (double)B, ln->L[0])) return e;
        CK(cudaEventRecord(ln->evMean, ln->L[0]));
        for (int i = 1; i < NSLOT; ++i) CK(cudaStreamWaitEvent(ln->L[i], ln->evMean, 0));
    }
    if (int e = admmnet_final_phi(y, b, B, chunk, Mdim, Ndim, K, params, phi_out, ws, ws_bytes, rcap, ln->L[0])) return e;
    CK(cudaEventRecord(ln->evEnd, ln->L[0]));
    CK(cudaStreamWaitEvent(cs, ln->evEnd, 0));
    return 0;
}

extern "C" int admmnet_status(void* ws, size_t ws_bytes, int B, int chunk, int n, int K, int rcap, void* stream,
                              int* status_host) {
    if (!status_host) return fail(ADMMNET_ERR_ARG, "status_host is NULL");
    if (int e = check_net_args(B, chunk, n, K, rcap)) return e;
    Ws w = carve(ws, B, chunk, n, n + 1, K, rcap);
    if (!ws || ws_bytes < w.bytes) return fail(ADMMNET_ERR_WORKSPACE, "workspace too small");
    CK(cudaMemcpyAsync(status_host, w.status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

// ------------------------------------------------------------------------------------ arrowhead tap
extern "C" int admmnet_arrow_eigh(const float* h, const void* phi, const float* c0, int B, int n, float* evals,
                                  void* evecs, int* handled, void* stream) {
    if (!h || !phi || !c0 || !evals || !handled) return fail(ADMMNET_ERR_ARG, "null pointer");
    if (B <= 0 || n < 2 || n > 127) return fail(ADMMNET_ERR_ARG, "need B > 0 and 2 <= n <= 127");
    ArrowArgs aa;
    aa.y = nullptr; aa.b = nullptr; aa.sigma = nullptr; aa.h_in = h; aa.phi_in = (const float2*)phi; aa.c0_in = c0;
    aa.Zp = nullptr; aa.GV = nullptr; aa.phi_cur = nullptr; aa.h_cur = nullptr; aa.Pk = nullptr; aa.r_out = nullptr;
    aa.handled = handled; aa.lam_out = evals; aa.U_out = (float2*)evecs;
    aa.B = B; aa.n = n; aa.d = n + 1; aa.ldu = 4 * ((n + 1 + 3) / 4);
    const size_t sma = arrow_smem_bytes(aa.d, aa.ldu);
    CK(ensure_smem(k_arrow, (int)sma));
    prof::Scope pscope(prof::ARROW, (cudaStream_t)stream);
    k_arrow<<<B, AR_NT, sma, (cudaStream_t)stream>>>(aa);
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------ eigh taps
extern "C" int admmnet_eigh_workspace_bytes(int B, int d, int rcap, size_t* bytes) {
    if (!bytes) return fail(ADMMNET_ERR_ARG, "bytes is NULL");
    if (B <= 0 || d < 3 || d > BIG_NMAX + 1) return fail(ADMMNET_ERR_ARG, "need B > 0 and 3 <= d <= 257");
    if (rcap == 0) rcap = default_rcap(d);
    *bytes = carve(nullptr, B, B, d - 1, d, 1, rcap).bytes;
    return 0;
}

extern "C" int admmnet_eigh_batched(const void* A, int B, int d, float* evals, void* evecs, void* fn_out,
                                    const float* params, void* ws, size_t ws_bytes, int rcap, void* stream,
                                    int* status_dev) {
    if (B <= 0 || d < 3 || d > BIG_NMAX + 1) return fail(ADMMNET_ERR_ARG, "need B > 0 and 3 <= d <= 257");
    if (!A || !status_dev) return fail(ADMMNET_ERR_ARG, "null pointer");
    if (rcap == 0) rcap = default_rcap(d);
    if (rcap < 2048 || rcap % ROT_CHUNK) return fail(ADMMNET_ERR_ARG, "rcap must be a multiple of 1024, >= 2048");
    Ws w = carve(ws, B, B, d - 1, d, 1, rcap);
    if (!ws || ws_bytes < w.bytes) return fail(ADMMNET_ERR_WORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (d > 128) {       // Jacobi form (csrc/big_kernels.cu); eigenvalues come out unsorted
        prof::Scope pscope(prof::BIG, st);
        k_big_eigh<<<big_grid(B), BIG_NT, 0, st>>>((const float2*)A, B, d, evals, (float2*)evecs, (float2*)fn_out, params,
                                                    w.big, status_dev);
        CK(cudaGetLastError());
        return 0;
    }
    const int ld = d | 1;
    const size_t sm = head_smem_bytes(d, ld);
    {
        prof::Scope pscope(prof::HEAD, st);
        const int ns = trd_ns(d);
        CK(ensure_smem(ns == 7 ? k_tridiag<7> : ns == 8 ? k_tridiag<8> : k_tridiag<0>, (int)sm));
        if (ns == 7) k_tridiag<7><<<B, 256, sm, st>>>((const float2*)A, B, d, ld, w.GV, w.tau, w.dT, w.eT, w.Ttr, d - 1);
        else if (ns == 8) k_tridiag<8><<<B, 256, sm, st>>>((const float2*)A, B, d, ld, w.GV, w.tau, w.dT, w.eT, w.Ttr, d - 1);
        else k_tridiag<0><<<B, 256, sm, st>>>((const float2*)A, B, d, ld, w.GV, w.tau, w.dT, w.eT, w.Ttr, stage1_steps(d));
    }
    CK(cudaGetLastError());
    if (int e = launch_head2(w, B, d, st)) return e;
    if (int e = launch_eig_tail(w, B, d - 1, d, rcap, params, params ? 0 : -1, (float2*)evecs, nullptr, status_dev, st))
        return e;
    const size_t npk = (size_t)d * (d + 1) / 2;
    if (evals) CK(cudaMemcpyAsync(evals, w.lam, (size_t)B * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (fn_out) CK(cudaMemcpyAsync(fn_out, w.GV, (size_t)B * npk * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// ------------------------------------------------------------------------------------ classical ADMM
extern "C" int admm_classic_forward(const void* y, const void* b, int in_is_c128, int B, int n, double rho, int n_iter,
                                    void* phi_out, void* stream) {
    if (!y || !b || !phi_out) return fail(ADMMNET_ERR_ARG, "null pointer");
    if (B <= 0 || n <= 0 || n > 256 || n_iter < 0) return fail(ADMMNET_ERR_ARG, "need B > 0, 0 < n <= 256, n_iter >= 0");
    cudaStream_t st = (cudaStream_t)stream;
    prof::Scope pscope(prof::CLASSIC, st);
    static const int gl_env = getenv("ADMMNET_CLASSIC_GL") ? atoi(getenv("ADMMNET_CLASSIC_GL")) : 0;
    // complex64 input, n even (16-byte bulk-copy granularity), n <= 104: the persistent streaming form
    // (ADMMNET_CLASSIC_P=0 or an explicit ADMMNET_CLASSIC_GL select the plain kernels)
    static const bool use_p = !(getenv("ADMMNET_CLASSIC_P") && atoi(getenv("ADMMNET_CLASSIC_P")) == 0);
    static const int pv = getenv("ADMMNET_CLASSIC_PV") ? atoi(getenv("ADMMNET_CLASSIC_PV")) : 0;
    if (use_p && !gl_env && !in_is_c128 && n <= 104 && n % 2 == 0 && (((uintptr_t)y | (uintptr_t)b | (uintptr_t)phi_out) & 15) == 0) {
        int dev = 0, nsm = 0;
        CK(cudaGetDevice(&dev));
        CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
#define CLASSIC_P_LAUNCH(GL, MAXE, NT, MINB)                                                                        \
        {                                                                                                           \
            const int S = classic_p_tile<GL, NT>(), ntiles = (B + S - 1) / S, cap = nsm * MINB;                     \
            const size_t sm = classic_p_smem_bytes<GL, NT>(n);                                                      \
            CK(ensure_smem(k_classic_p<GL, MAXE, NT, MINB>, (int)sm));                                              \
            k_classic_p<GL, MAXE, NT, MINB><<<ntiles < cap ? ntiles : cap, NT, sm, st>>>(                           \
                (const float2*)y, (const float2*)b, B, n, rho, n_iter, (double2*)phi_out);                          \
        }
        // tile / occupancy variants (tools/classic_ab.py; batch 65536, n = 100, 5 / 100 iterations, us per launch):
        //   0: 8 lanes per signal, 256 threads, one CTA per SM (default)   64.7 / 355
        //   1: 16 lanes, 256 threads, two CTAs per SM                      65.2 / 483
        //   2: 8 lanes, 128 threads, two CTAs per SM                       62.8 / 395
        //   3: 16 lanes, 512 threads, one CTA per SM                       66.5 / 460
        // against 94 / 504 for the best plain kernel: the short run is memory bound whatever the tiling, the long one
        // fp64 bound and prefers the fewest shuffle steps and idle lanes (8 lanes x 13 elements for n = 100)
        if (pv == 1) CLASSIC_P_LAUNCH(16, 7, 256, 2)
        else if (pv == 2) CLASSIC_P_LAUNCH(8, 13, 128, 2)
        else if (pv == 3) CLASSIC_P_LAUNCH(16, 7, 512, 1)
        else CLASSIC_P_LAUNCH(8, 13, 256, 1)
#undef CLASSIC_P_LAUNCH
        CK(cudaGetLastError());
        return 0;
    }
    // plain kernels (complex128 input, odd n, n > 104), measured on B200 at batch 65536, n = 100: 5 iterations
    // 32 lanes 94 us / 16 lanes 114 us / 8 lanes 167 us; 100 iterations 504 / 506 / 610 us
    const int gl = (gl_env == 8 && n <= 104) ? 8 : (gl_env == 16 && n <= 112) ? 16 : (gl_env == 32) ? 32
                   : ((n_iter > 20 && n <= 112) ? 16 : 32);
    const int grid = (int)(((long long)B * gl + 255) / 256);      // one group of gl lanes per signal
#define CLASSIC_LAUNCH(T, E, G) k_classic<T, E, G><<<grid, 256, 0, st>>>((const T*)y, (const T*)b, B, n, rho, n_iter, (double2*)phi_out)
    if (gl == 8) {
        if (in_is_c128) CLASSIC_LAUNCH(double2, 13, 8); else CLASSIC_LAUNCH(float2, 13, 8);
    } else if (gl == 16) {
        if (in_is_c128) CLASSIC_LAUNCH(double2, 7, 16); else CLASSIC_LAUNCH(float2, 7, 16);
    } else if (n <= 128) {
        if (in_is_c128) CLASSIC_LAUNCH(double2, 4, 32); else CLASSIC_LAUNCH(float2, 4, 32);
    } else {
        if (in_is_c128) CLASSIC_LAUNCH(double2, 8, 32); else CLASSIC_LAUNCH(float2, 8, 32);
    }
#undef CLASSIC_LAUNCH
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------ peak search
extern "C" int peak_search_full(const void* phi, int phi_is_c128, int B, int xbase, int ybase, const double* axis_x,
                                int Gx, const double* axis_y, int Gy, double xmin, double xmax, double xstep,
                                double ymin, double ymax, double ystep, double reducefactor, int iters, int pmax,
                                double* peaks, int* count, int topl, double* top, double* surface, int* status_dev,
                                void* stream) {
    if (!phi || !axis_x || !axis_y || !peaks || !count || !status_dev) return fail(ADMMNET_ERR_ARG, "null pointer");
    if (B <= 0 || Gx <= 0 || Gy <= 0 || pmax <= 0 || iters < 0 || topl < 0) return fail(ADMMNET_ERR_ARG, "bad sizes");
    if (xbase < 1 || ybase < 1 || xbase > PEAK_MAX_BASE || ybase > PEAK_MAX_BASE)
        return fail(ADMMNET_ERR_ARG, "xbase/ybase must be in [1,32]");
    if (topl > 0 && !top) return fail(ADMMNET_ERR_ARG, "top is NULL");
    const size_t sm0 = peak_smem_bytes(Gx, Gy, xbase, ybase, pmax);
    const size_t per_peak = peak_refine_bytes_per_peak(xbase, ybase);
    if (sm0 + 4 * per_peak > 227 * 1024) return fail(ADMMNET_ERR_ARG, "coarse grid too large for one CTA's shared memory");
    int ptile = (int)((227 * 1024 - sm0) / per_peak);
    if (ptile > 64) ptile = 64;
    if ((size_t)ptile * per_peak > 72 * 1024) ptile = (int)(72 * 1024 / per_peak);
    if (ptile < 4) ptile = 4;
    const size_t sm = sm0 + (size_t)ptile * per_peak + 16;
    PeakArgs a;
    a.phi = phi; a.phi_is_c128 = phi_is_c128; a.B = B; a.xb = xbase; a.yb = ybase;
    a.axis_x = axis_x; a.axis_y = axis_y; a.Gx = Gx; a.Gy = Gy;
    a.xmin = xmin; a.xmax = xmax; a.xstep = xstep; a.ymin = ymin; a.ymax = ymax; a.ystep = ystep;
    a.reducefactor = reducefactor; a.iters = iters; a.pmax = pmax; a.ptile = ptile; a.peaks = peaks; a.count = count;
    a.topl = topl; a.top = top; a.surface = surface; a.status = status_dev; a.surface_in = nullptr;
    CK(ensure_smem(k_peak_search, (int)sm));
    prof::Scope pscope(prof::PEAK, (cudaStream_t)stream);
    k_peak_search<<<B, PEAK_NT, sm, (cudaStream_t)stream>>>(a);
    CK(cudaGetLastError());
    return 0;
}

// Debug tap: the 8-connected, plateau-aware local-maximum stage of k_peak_search (skimage.morphology.local_maxima as
// peakSearchUtils.py:118 calls it) on a caller-supplied surface; peaks[B][pmax][3] = (axis_x[ix], axis_y[iy], value)
// in row-major discovery order (np.where order, peakSearchUtils.py:119).
extern "C" int peak_surface_maxima(const double* surface, int B, const double* axis_x, int Gx, const double* axis_y,
                                   int Gy, int pmax, double* peaks, int* count, int* status_dev, void* stream) {
    if (!surface || !axis_x || !axis_y || !peaks || !count || !status_dev) return fail(ADMMNET_ERR_ARG, "null pointer");
    if (B <= 0 || Gx <= 0 || Gy <= 0 || pmax <= 0) return fail(ADMMNET_ERR_ARG, "bad sizes");
    const size_t sm0 = peak_smem_bytes(Gx, Gy, 1, 1, pmax);
    const size_t per_peak = peak_refine_bytes_per_peak(1, 1);
    if (sm0 + 4 * per_peak > 227 * 1024) return fail(ADMMNET_ERR_ARG, "surface too large for one CTA's shared memory");
    PeakArgs a;
    a.phi = surface; a.phi_is_c128 = 1; a.B = B; a.xb = 1; a.yb = 1;
    a.axis_x = axis_x; a.axis_y = axis_y; a.Gx = Gx; a.Gy = Gy;
    a.xmin = 0; a.xmax = 1; a.xstep = 1; a.ymin = 0; a.ymax = 1; a.ystep = 1;
    a.reducefactor = 0.1; a.iters = 0; a.pmax = pmax; a.ptile = 4; a.peaks = peaks; a.count = count;
    a.topl = 0; a.top = nullptr; a.surface = nullptr; a.status = status_dev; a.surface_in = surface;
    const size_t sm = sm0 + 4 * per_peak + 16;
    CK(ensure_smem(k_peak_search, (int)sm));
    k_peak_search<<<B, PEAK_NT, sm, (cudaStream_t)stream>>>(a);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int peak_search_points(const void* phi, int phi_is_c128, int xbase, int ybase, const double* X,
                                  const double* Y, int npts, double* out, void* stream) {
    if (!phi || !X || !Y || !out) return fail(ADMMNET_ERR_ARG, "null pointer");
    if (npts <= 0) return 0;
    if (xbase < 1 || ybase < 1 || xbase > PEAK_MAX_BASE || ybase > PEAK_MAX_BASE)
        return fail(ADMMNET_ERR_ARG, "xbase/ybase must be in [1,32]");
    const size_t sm = (size_t)xbase * ybase * sizeof(double2);
    k_peak_points<<<(npts + 255) / 256, 256, sm, (cudaStream_t)stream>>>(phi, phi_is_c128, xbase, ybase, X, Y, npts, out);
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------ tcgen05 / TMEM / TMA probe
extern "C" int admmnet_tc_gemm_probe(const float* A, const float* B, const float* D0, int N, int K, int flags,
                                     float* out, void* stream) {
    if (!A || !B || !out) return fail(ADMMNET_ERR_ARG, "null pointer");
    if (N < 8 || N > 256 || (N % 8) || K < 8 || K > 64 || (K % 8)) return fail(ADMMNET_ERR_ARG, "need N in [8,256] step 8, K in [8,64] step 8");
    if (flags & (TCP_A_MN | TCP_B_MN))
        return fail(ADMMNET_ERR_ARG, "MN-major staging is not supported (the tail kernel only uses K-major tiles)");
    const size_t sm = 1024 + (size_t)(3 * 128 + 3 * N) * K * sizeof(float);
    if (sm > 227 * 1024) return fail(ADMMNET_ERR_ARG, "probe tile does not fit shared memory");
    TcProbeArgs a;
    a.A = A; a.B = B; a.D0 = D0; a.out = out; a.N = N; a.K = K; a.flags = flags;
    CUtensorMap tmA, tmB;
    memset(&tmA, 0, sizeof(tmA));
    memset(&tmB, 0, sizeof(tmB));
    if (flags & TCP_TMA) {
        if (!make_tmap_2d_f32(&tmA, A, 128, K, (uint64_t)K * 4, 128, K) || !make_tmap_2d_f32(&tmB, B, N, K, (uint64_t)K * 4, N, K))
            return fail(ADMMNET_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    }
    CK(ensure_smem(k_tc_probe, (int)sm));
    k_tc_probe<<<1, 128, sm, (cudaStream_t)stream>>>(a, tmA, tmB);
    CK(cudaGetLastError());
    return 0;
}

// Per-phase clock counters of k_tail_tc (CTA 0, summed over launches) when ADMMNET_TC_PROF=1: a tuning aid.
static long long* tc_prof_buffer() {
    static long long* buf = [] {
        long long* p = nullptr;
        if (getenv("ADMMNET_TC_PROF") && atoi(getenv("ADMMNET_TC_PROF")) != 0) {
            if (cudaMalloc(&p, 16 * sizeof(long long)) != cudaSuccess) p = nullptr;
            else cudaMemset(p, 0, 16 * sizeof(long long));
        }
        return p;
    }();
    return buf;
}
extern "C" int admmnet_tail_tc_profile_read(long long* host16) {
    if (!host16) return fail(ADMMNET_ERR_ARG, "null pointer");
    long long* p = tc_prof_buffer();
    if (!p) { memset(host16, 0, 16 * sizeof(long long)); return 0; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(host16, p, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    CK(cudaMemset(p, 0, 16 * sizeof(long long)));
    return 0;
}

extern "C" int admmnet_dc_profile_read(long long* host128) {
    if (!host128) return fail(ADMMNET_ERR_ARG, "null pointer");
    long long* p = dc_prof_buffer();
    if (!p) { memset(host128, 0, DCP_N * sizeof(long long)); return 0; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(host128, p, DCP_N * sizeof(long long), cudaMemcpyDeviceToHost));
    CK(cudaMemset(p, 0, DCP_N * sizeof(long long)));
    return 0;
}

// tcgen05.mma work k_tail_tc issues per signal for matrix order d (flops, 2 per multiply-add, every 3xTF32 split term
// counted): per compact-WY block with window K2 = 2 (dp - a0): GEMM 1 = M 128 x N (96 + 48) x K K2, GEMM 2 = six
// chains M 128 x N K2 x K 24; rebuild = twelve chains M 128 x N dp x K dp.  0 when the kernel does not serve d.
extern "C" double admmnet_tail_tc_mma_flops(int d) {
    TailTcPlan plan;
    if (!tail_tc_plan(d, plan)) return 0.0;
    double f = 0.0;
    for (int j = 0; j < plan.nblk; ++j) {
        const double K2 = 2.0 * (plan.dp - plan.a0[j]);
        f += 2.0 * 128.0 * K2 * (96.0 + 48.0) + 6.0 * 2.0 * 128.0 * K2 * TC_NB;
    }
    f += 12.0 * 2.0 * 128.0 * (double)plan.dp * (double)plan.dp;
    return f;
}

// shared-memory bytes of the tensor-core tail kernel for matrix order d, or -1 when d is outside its range (then the
// SIMT tail kernels serve); 0 when it is switched off (ADMMNET_TAILTC=0)
extern "C" int admmnet_tail_tc_smem_bytes(int d) {
    if (getenv("ADMMNET_TAILTC") && atoi(getenv("ADMMNET_TAILTC")) == 0) return 0;
    TailTcPlan plan;
    return tail_tc_plan(d, plan) ? plan.total : -1;
}

// ------------------------------------------------------------------------------------ FP32 FMA peak
// Roofline denominator for the FP32-pipe-bound eigen-solver kernels (MEASURED_PEAKS.json has no FP32
// figure): 8 independent FFMA chains per thread, grid = 148*8 CTAs of 256 threads.
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
// out must hold grid*256 floats; returns the flop count of one launch in *flops
extern "C" int admmnet_fp32_peak_launch(float* out, int grid, int iters, double* flops, void* stream) {
    if (!out || grid <= 0 || iters <= 0) return fail(ADMMNET_ERR_ARG, "bad arguments");
    k_fma_peak<<<grid, 256, 0, (cudaStream_t)stream>>>(out, iters, 0.999f, 0.001f);
    CK(cudaGetLastError());
    if (flops) *flops = 2.0 * 8 * 16 * (double)iters * 256.0 * grid;
    return 0;
}

// ------------------------------------------------------------------------------------ synthetic inputs
extern "C" int admmnet_generate_dataset(void* y, void* b, float* sigma, double* truth, float* ser, int B, int Nb, int Nd,
                                        int L, double snr_w_lo_db, double snr_w_hi_db, double snr_demod_db,
                                        unsigned long long seed, void* stream);
extern "C" int admmnet_generate(void* y, void* b, float* sigma, double* truth, int B, int Nb, int Nd, int L,
                                double snr_w_db, double snr_demod_db, unsigned long long seed, void* stream) {
    return admmnet_generate_dataset(y, b, sigma, truth, nullptr, B, Nb, Nd, L, snr_w_db, snr_w_db, snr_demod_db, seed,
                                    stream);
}
extern "C" int admmnet_generate_dataset(void* y, void* b, float* sigma, double* truth, float* ser, int B, int Nb, int Nd,
                                        int L, double snr_w_db, double snr_w_hi_db, double snr_demod_db,
                                        unsigned long long seed, void* stream) {
    if (!y || !b || !sigma) return fail(ADMMNET_ERR_ARG, "null pointer");
    if (B <= 0 || Nb < 1 || Nd < 1 || Nb * Nd > 256 || L < 1 || L > GEN_MAXL)
        return fail(ADMMNET_ERR_ARG, "need B > 0, Nb*Nd <= 256, 1 <= L <= 8");
    GenArgs a;
    a.y = (float2*)y; a.b = (float2*)b; a.sigma = sigma; a.truth = truth; a.ser = ser;
    a.B = B; a.Nb = Nb; a.Nd = Nd; a.L = L; a.snr_w_db = snr_w_db; a.snr_w_hi_db = snr_w_hi_db;
    a.snr_demod_db = snr_demod_db; a.seed = seed;
    const long long nthreads = (long long)B * 32;
    prof::Scope pscope(prof::MISC, (cudaStream_t)stream);
    k_generate<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------ ADMMNet regression head
extern "C" int admmnet_head_param_count(int n, int L) { return head_param_count(n, L); }
extern "C" int admmnet_peak_head(const void* phi, int B, int n, int L, const float* head_params, float* tau, float* f,
                                 float* conf, void* stream) {
    if (!phi || !head_params || !tau || !f || !conf) return fail(ADMMNET_ERR_ARG, "null pointer");
    if (B <= 0 || n < 1 || n > 128 || L < 1 || L > 8) return fail(ADMMNET_ERR_ARG, "need B > 0, n <= 128, 1 <= L <= 8");
    HeadArgs2 a;
    a.phi = (const float2*)phi; a.P = head_params; a.tau = tau; a.f = f; a.conf = conf; a.B = B; a.n = n; a.L = L;
    const size_t sm = (size_t)(HS * 2 * n + 3 * HS * HD + HS * HEADS * n) * sizeof(float);
    CK(ensure_smem(k_peak_head, (int)sm));
    prof::Scope pscope(prof::MISC, (cudaStream_t)stream);
    k_peak_head<<<(B + HS - 1) / HS, HD, sm, (cudaStream_t)stream>>>(a);
    CK(cudaGetLastError());
    return 0;
}
