// C-ABI of libb200gp (include/b200gp.h): context, workspace carving, phase orchestration, jitter ladder.
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200gp.h"
#include "kernels.cuh"

using namespace b200gp;

namespace b200gp {
std::atomic<long long> g_launch_count{0};
int configure_hellinger_kernels();
void launch_hellinger_precompute(const double* Xsub, int n, int D, const int4* prog, const int* prog_off,
                                 const double* theta, const int* theta_off, const double* lambda, int M, int S,
                                 double* logdet, double* G, int* status, cudaStream_t st);
void launch_hellinger_pairs(const double* logdet, const double* G, int S, int n, const int* pair_i, const int* pair_j,
                            int npairs, double tol, double* h, int* hstatus, double* dist, int* status, cudaStream_t st);
void launch_slot_logdet(const Batch& b, double* logdet, int* status, cudaStream_t st);
void launch_havg_build(const Batch& b, const double* G, int S, int n, const int* pair_i, const int* pair_j,
                       const int* work_pair, const int* work_s, cudaStream_t st);
void launch_hpair_init(const double* logdet, int S, const int* pair_i, int npairs, double* ldpq, int* hstatus, cudaStream_t st);
void launch_hpair_scatter(const Batch& b, int S, const int* work_pair, const int* work_s, double* ldpq, int* hstatus,
                          cudaStream_t st);
void launch_hpair_finish(const double* logdet, int S, const int* pair_i, const int* pair_j, int npairs, double* ldpq_h,
                         int* hstatus, double* dist, int* status, cudaStream_t st);
void launch_vector_pairs(const double* G, int S, int n, const int* pair_i, const int* pair_j, int npairs, int metric,
                         double* dist, cudaStream_t st);
void launch_pack_data(const double* X, const double* y, int N, int D, int Np, double* Xt, double* yp, cudaStream_t s);
void launch_predict(const Batch& b, int item, const double* Xs, int ns, int include_likelihood, double* scratch, double* mean,
                    double* var, cudaStream_t s);
int configure_predict_kernels();
int launch_alignments(const Batch& b, int nmodels, double* out, cudaStream_t s);
int configure_align_kernels();
}

struct b200gp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int quirks = 0;
  bool bound = false;
  bool profiling = false;
  Batch base{};             // geometry + workspace pointers
  int slots = 0, max_params = 0;
  int panel_tiles = 4;          // look-ahead Cholesky (single large matrix)
  bool panel_user = false;      // b200gp_set_panel_tiles was called: no adaptive choice
  int batch_panel_tiles = 8;    // batched Cholesky / L^-T (measured: 4 -> 26.6, 8 -> 26.2 ms per 256 items at N = 2048)
  int gemm64 = 1;               // Cholesky / L^-T updates on 64 x 64 CTA tiles
  int potrf_la = 1;             // diagonal tiles by the look-ahead register factorisation
  bool kinv64 = true;           // K^-1 on 64 x 64 CTA tiles with CTA-granular skipping (kinv64_kernel) instead of kinv_kernel
  bool fused_grad = false;      // K^-1 and the gradient of the <= 4-leaf programs in one kernel (kinv_grad_kernel)
  int tma = 3;                  // 64 x 64 tile GEMMs fed by the TMA (1 + swizzle mode, tile_gemm64_tma.cuh); 0: cp.async
  bool tma_ok = false;          // tensor maps of the bound workspace are valid
  alignas(64) unsigned char tmaps[256];        // TmaMaps of the bound M1 / M2
  alignas(64) unsigned char tmaps_potrf[256];  // TmaMaps of the matrix of the running b200gp_potrf call
  int n_groups = 2;             // slot groups (streams) of the batched schedule (measured at 256 x N=2048: 1 group 89.2 ms,
                                // 2 / 3 / 4 groups 87.8 ms -- the schedule is not the limiter)
  cudaStream_t gstream[4]{};    // streams of the groups 1..3 (group 0 runs on the context stream)
  cudaEvent_t gev[4]{};
  int la_slots = 65;            // batches below this many slots take the look-ahead / pipelined (latency) schedules
  double* ws_jitter = nullptr;
  int* ws_ids = nullptr;
  int* ws_tries_dummy = nullptr;
  // pinned host staging
  int* h_status = nullptr;
  int* h_ids = nullptr;
  double* h_diag = nullptr;
  double* h_jit = nullptr;
  size_t h_cap = 0;
  cudaEvent_t ev[8]{};
  cudaStream_t side = nullptr;     // high-priority stream for the look-ahead panel of large single factorisations
  cudaStream_t peer = nullptr;     // second stream of the two-group batched schedule
  const double* lut = nullptr;     // OP_LOOKUP table (b200gp_set_lookup)
  int lut_n = 0;
  cudaEvent_t evP = nullptr, evR = nullptr, evS = nullptr, evI = nullptr;
  std::vector<cudaEvent_t> evC;    // Cholesky panel p finished (small batches: the L^-T phase follows panel by panel)
  float phase_ms[6]{};
  int n_phase = 0;
};

static int fail(b200gp_ctx* c, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return -1;
}
#define CUCHECK(c, expr)                                                                     \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) return fail(c, "%s failed: %s", #expr, cudaGetErrorString(_e));  \
  } while (0)

namespace {
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) / 256 * 256;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct Layout {
  double *M1, *M2, *Dinv, *z, *alpha, *logdet_part, *diag_part, *grad_part, *solve_part, *Xt, *y, *jitter;
  int *slot_status, *ids, *klist, *kcount;
  CProg* cprog;
  size_t bytes;
};

Layout carve(void* ws, int N, int D, int slots, int max_params) {
  const int Np = round_up(N, NB), T = Np / NB;
  const size_t S = (size_t)slots;
  Carver c(ws);
  Layout l;
  l.M1 = c.take<double>(S * Np * Np);
  l.M2 = c.take<double>(S * Np * Np);
  l.Dinv = c.take<double>(S * T * NB * NB);
  l.z = c.take<double>(S * Np);
  l.alpha = c.take<double>(S * Np);
  l.logdet_part = c.take<double>(S * T);
  l.diag_part = c.take<double>(S * T);
  l.grad_part = c.take<double>(S * tri_count(T) * 4 * (size_t)(max_params + 1));
  l.solve_part = c.take<double>(S * tri_count(T) * (size_t)NB);
  l.Xt = c.take<double>((size_t)D * Np);
  l.y = c.take<double>(Np);
  l.jitter = c.take<double>(S);
  l.slot_status = c.take<int>(S);
  l.ids = c.take<int>(S);
  l.cprog = c.take<CProg>(S);
  l.klist = c.take<int>(S * CK_KINDS);
  l.kcount = c.take<int>(CK_KINDS);
  l.bytes = c.off + 256;
  return l;
}
}  // namespace

extern "C" {

const char* b200gp_version(void) { return "b200gp 0.2 (sm_100a, FP64 DMMA, TMA-fed tile GEMMs)"; }

int b200gp_create(int device, void* cuda_stream, b200gp_ctx** out) {
  if (!out) return -1;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return -2;
  if (cudaSetDevice(device) != cudaSuccess) return -3;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -3;
  if (prop.major < 10) return -4;   // sm_100a only: there is no fallback path
  if (configure_kernels() != 0 || configure_hellinger_kernels() != 0 || configure_predict_kernels() != 0 ||
      configure_align_kernels() != 0) return -5;
  b200gp_ctx* c = new b200gp_ctx();
  c->device = device;
  c->stream = static_cast<cudaStream_t>(cuda_stream);
  for (auto& e : c->ev) cudaEventCreate(&e);
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi);
  cudaStreamCreateWithFlags(&c->peer, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->evP, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->evR, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->evS, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->evI, cudaEventDisableTiming);
  for (int q = 1; q < 4; q++) {
    cudaStreamCreateWithFlags(&c->gstream[q], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&c->gev[q], cudaEventDisableTiming);
  }
  *out = c;
  return 0;
}

int b200gp_destroy(b200gp_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  for (auto& e : c->ev) cudaEventDestroy(e);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->peer) cudaStreamDestroy(c->peer);
  if (c->evP) { cudaEventDestroy(c->evP); cudaEventDestroy(c->evR); cudaEventDestroy(c->evS); cudaEventDestroy(c->evI); }
  for (auto& e : c->evC) cudaEventDestroy(e);
  for (int q = 1; q < 4; q++) { if (c->gstream[q]) cudaStreamDestroy(c->gstream[q]); if (c->gev[q]) cudaEventDestroy(c->gev[q]); }
  if (c->h_status) cudaFreeHost(c->h_status);
  if (c->h_ids) cudaFreeHost(c->h_ids);
  if (c->h_diag) cudaFreeHost(c->h_diag);
  if (c->h_jit) cudaFreeHost(c->h_jit);
  delete c;
  return 0;
}

const char* b200gp_last_error(b200gp_ctx* c) { return c ? c->err.c_str() : "null context"; }

int b200gp_set_quirks(b200gp_ctx* c, int mask) {
  if (!c) return -1;
  c->quirks = mask & (QUIRK_PER_LENGTHSCALE | QUIRK_LIN_SHIFT);
  return 0;
}

int b200gp_set_lookup(b200gp_ctx* c, const double* d_table, int n) {
  if (!c) return -1;
  if ((d_table == nullptr) != (n == 0) || n < 0) return fail(c, "b200gp_set_lookup: need (table, n > 0) or (NULL, 0)");
  c->lut = d_table;
  c->lut_n = n;
  return 0;
}

int b200gp_set_panel_tiles(b200gp_ctx* c, int tiles) {
  if (!c || tiles < 1 || tiles > 16) return -1;
  c->panel_tiles = tiles;
  c->batch_panel_tiles = tiles;
  c->panel_user = true;
  return 0;
}

int b200gp_set_tile64(b200gp_ctx* c, int enabled) {
  if (!c) return -1;
  c->gemm64 = enabled ? 1 : 0;
  c->kinv64 = enabled != 0;
  c->base.gemm64 = c->gemm64;
  return 0;
}

int b200gp_set_potrf_la(b200gp_ctx* c, int enabled) {
  if (!c) return -1;
  c->potrf_la = enabled < 0 ? 0 : (enabled > 2 ? 2 : enabled);
  c->base.potrf_la = c->potrf_la;
  return 0;
}

static void encode_bound_maps(b200gp_ctx* c) {
  c->tma_ok = false;
  c->base.tma = 0;
  c->base.tmaps = nullptr;
  if (!c->bound || c->tma == 0) return;
  if (tma_maps_bytes() > sizeof c->tmaps) return;
  const Batch& b = c->base;
  if (tma_encode_maps(c->tmaps, b.M1, b.M2, b.Np, (size_t)c->slots * b.Np, (c->tma == 4 ? 3 : c->tma) - 1) != 0) return;
  c->tma_ok = true;
  c->base.tma = c->tma;
  c->base.tmaps = c->tmaps;
}

int b200gp_set_tma(b200gp_ctx* c, int mode) {
  if (!c || mode < 0 || mode > 4) return -1;
  c->tma = mode;
  encode_bound_maps(c);
  return 0;
}

int b200gp_set_fused_grad(b200gp_ctx* c, int enabled) {
  if (!c) return -1;
  c->fused_grad = enabled != 0;
  return 0;
}

int b200gp_set_profiling(b200gp_ctx* c, int enabled) {
  if (!c) return -1;
  c->profiling = enabled != 0;
  return 0;
}

long long b200gp_launch_count(b200gp_ctx*) { return g_launch_count.load(); }

int b200gp_last_phase_ms(b200gp_ctx* c, float* out, int cap) {
  if (!c || !out) return -1;
  int n = c->n_phase < cap ? c->n_phase : cap;
  for (int i = 0; i < n; i++) out[i] = c->phase_ms[i];
  return n;
}

size_t b200gp_workspace_bytes(int N, int D, int slots, int max_params) {
  if (N <= 0 || D <= 0 || slots <= 0 || max_params < 0) return 0;
  return carve(nullptr, N, D, slots, max_params).bytes;
}

int b200gp_bind(b200gp_ctx* c, void* ws, size_t ws_bytes, int slots, int max_params, const double* dX, int N, int D,
                const double* dy) {
  if (!c) return -1;
  if (!ws || !dX || !dy) return fail(c, "b200gp_bind: null pointer");
  if (N <= 0 || D <= 0 || D > MAX_DIMS) return fail(c, "b200gp_bind: need 0 < N, 0 < D <= %d", MAX_DIMS);
  if (slots <= 0) return fail(c, "b200gp_bind: slots must be positive");
  if (max_params < 1 || max_params > MAX_PARAMS) return fail(c, "b200gp_bind: max_params must be in [1, %d]", MAX_PARAMS);
  const size_t need = b200gp_workspace_bytes(N, D, slots, max_params);
  if (ws_bytes < need) return fail(c, "b200gp_bind: workspace too small (%zu < %zu bytes)", ws_bytes, need);
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(c, "b200gp_bind: workspace must be 256-byte aligned");
  CUCHECK(c, cudaSetDevice(c->device));
  Layout l = carve(ws, N, D, slots, max_params);
  Batch& b = c->base;
  b = Batch{};
  b.N = N; b.Np = round_up(N, NB); b.T = b.Np / NB; b.D = D;
  b.mstride = (size_t)b.Np * b.Np;
  b.M1 = l.M1; b.M2 = l.M2; b.Dinv = l.Dinv; b.z = l.z; b.alpha = l.alpha;
  b.logdet_part = l.logdet_part; b.diag_part = l.diag_part; b.grad_part = l.grad_part; b.solve_part = l.solve_part;
  b.pstride = max_params + 1;
  b.gemm64 = c->gemm64;
  b.potrf_la = c->potrf_la;
  b.slot_status = l.slot_status;
  b.Xt = l.Xt; b.y = l.y;
  b.cprog = l.cprog; b.klist = l.klist; b.kcount = l.kcount; b.list_stride = slots;
  c->ws_jitter = l.jitter; c->ws_ids = l.ids;
  c->slots = slots; c->max_params = max_params;
  launch_pack_data(dX, dy, N, D, b.Np, l.Xt, l.y, c->stream);
  CUCHECK(c, cudaGetLastError());
  const size_t cap = (size_t)slots;
  if (cap > c->h_cap) {
    if (c->h_status) { cudaFreeHost(c->h_status); cudaFreeHost(c->h_ids); cudaFreeHost(c->h_diag); cudaFreeHost(c->h_jit); }
    CUCHECK(c, cudaMallocHost(&c->h_status, cap * sizeof(int)));
    CUCHECK(c, cudaMallocHost(&c->h_ids, cap * sizeof(int)));
    CUCHECK(c, cudaMallocHost(&c->h_jit, cap * sizeof(double)));
    c->h_cap = cap;
    c->h_diag = nullptr;
  }
  if (c->h_diag) cudaFreeHost(c->h_diag);
  CUCHECK(c, cudaMallocHost(&c->h_diag, cap * b.T * sizeof(double)));
  c->bound = true;
  encode_bound_maps(c);       // needs c->bound; b.tma / b.tmaps stay 0 / null when the driver cannot encode (cp.async path)
  return 0;
}

}  // extern "C"

// Panelled right-looking schedule shared by the Cholesky and the L^-T phases.
//   panel(g, k0, k1, stream)      finishes tile columns [k0, k1) of slot group g given every contribution from the left
//   update(g, k0, W, j0, j1, st)  applies panel [k0, k0+W) to the columns [j0, j1) to its right, K = 128 W
//  * many slots: the chunk is split into two slot groups that run the same schedule on two streams, so the
//    latency-bound steps of one group (in-register diagonal factorisations, K = 128 triangular solves) overlap the
//    DMMA trailing updates of the other;
//  * few slots (a single large matrix): look-ahead.  The next panel (its priority update, the latency-bound diagonal
//    factorisations and the TRSMs) runs on a high-priority side stream while the main stream is still applying the
//    current panel to the rest of the trailing matrix, so the panel disappears behind the DMMA update.
// Panel width of the look-ahead schedule.  Large single matrices (T > 32) want wide panels (trailing update with
// K = 512: C tiles are re-read 4x less often).  Small batches of moderate order are latency-bound -- the chain
// potrf -> trsm -> update of the next column is the evaluation time -- and want the shortest chain: one tile column per
// panel, so the update in front of every diagonal factorisation has K = 128 (measured at N = 2048: 1 item 2.2 -> 1.9 ms,
// 8 items 4.1 -> 3.5 ms with W = 1; from ~16 items on W = 2 is as good or better).
static int la_panel_width(const b200gp_ctx* c, const Batch& b) {
  if (c->panel_user || b.T > 32) return c->panel_tiles;
  return b.nslots <= 12 ? 1 : 2;
}

template <typename PanelFn, typename UpdateFn>
static void panel_schedule(b200gp_ctx* c, const Batch& b, PanelFn panel, UpdateFn update) {
  const int T = b.T;
  const bool la = (b.nslots < c->la_slots && T >= 2);
  if (!la) {
    const int W = c->batch_panel_tiles;
    const int ng = b.nslots >= 16 ? c->n_groups : 1;
    Batch g[4] = {b, b, b, b};
    cudaStream_t st[4] = {c->stream, c->gstream[1], c->gstream[2], c->gstream[3]};
    if (ng > 1) cudaEventRecord(c->evS, c->stream);
    int first = b.slot0;
    for (int q = 0; q < ng; q++) {
      g[q].slot0 = first;
      g[q].nslots = (b.nslots * (q + 1)) / ng - (b.nslots * q) / ng;
      first += g[q].nslots;
      if (q > 0) cudaStreamWaitEvent(st[q], c->evS, 0);
    }
    for (int q = 0; q < ng; q++) {
      panel(g[q], 0, W < T ? W : T, st[q]);
      for (int k0 = 0; k0 + W < T; k0 += W) {
        const int k1 = k0 + W;
        update(g[q], k0, W, k1, T, st[q]);
        panel(g[q], k1, (k1 + W < T) ? k1 + W : T, st[q]);
      }
    }
    for (int q = 1; q < ng; q++) {
      cudaEventRecord(c->gev[q], st[q]);
      cudaStreamWaitEvent(c->stream, c->gev[q], 0);
    }
    return;
  }
  cudaStream_t s0 = c->stream, s1 = c->side;
  const int W = la_panel_width(c, b);
  cudaEventRecord(c->evS, s0);
  cudaStreamWaitEvent(s1, c->evS, 0);
  panel(b, 0, W < T ? W : T, s1);
  cudaEventRecord(c->evP, s1);
  bool have_r = false;
  for (int k0 = 0; k0 + W < T; k0 += W) {
    const int k1 = k0 + W;
    const int n1 = (k1 + W < T) ? k1 + W : T;                 // end of the next panel
    cudaStreamWaitEvent(s0, c->evP, 0);                       // panel [k0, k1) is finished
    if (have_r) cudaStreamWaitEvent(s1, c->evR, 0);           // columns [k1, n1) have every earlier trailing update
    update(b, k0, W, k1, n1, s1);                             // look-ahead: the next panel's columns first ...
    if (n1 < T) update(b, k0, W, n1, T, s0);                  // ... the rest of the trailing matrix meanwhile
    cudaEventRecord(c->evR, s0);
    have_r = true;
    panel(b, k1, n1, s1);
    cudaEventRecord(c->evP, s1);
  }
  cudaStreamWaitEvent(s0, c->evP, 0);
}

// Right-looking blocked Cholesky of every slot of `b` (M1 in place).  `panel_done(k0, stream)`, if given, is called when
// the tile columns of the panel starting at k0 are final.
template <typename DoneFn>
static void cholesky_phase_impl(b200gp_ctx* c, const Batch& b, DoneFn panel_done) {
  const int T = b.T;
  panel_schedule(
      c, b,
      [&](const Batch& g, int k0, int k1, cudaStream_t s) {
        for (int k = k0; k < k1; k++) {      // left-looking inside the panel: one K = 128 (k - k0) update per column
          if (k > k0) launch_update(g, UpdRange{k0, k - k0, k, k + 1}, s);
          launch_potrf_diag(g, k, s);
          if (k + 1 < T) launch_gemm(g, GM_TRSM, k, s);
        }
        panel_done(k0, s);
      },
      [&](const Batch& g, int k0, int W, int j0, int j1, cudaStream_t s) { launch_update(g, UpdRange{k0, W, j0, j1}, s); });
}

static void cholesky_phase(b200gp_ctx* c, const Batch& b) {
  cholesky_phase_impl(c, b, [](int, cudaStream_t) {});
}

// One panel of the right-looking V = L^-T (upper tiles of M2): column i is finished by V_ji = -S_ji D_i^T (j < i), then
// pushes its contribution S_jm += V_ji L_mi^T to every column m > i.
static void inverse_panel(const Batch& g, int k0, int k1, cudaStream_t s) {
  for (int i = k0; i < k1; i++) {      // left-looking inside the panel, like the Cholesky
    if (i > k0) launch_vupdate(g, UpdRange{k0, i - k0, i, i + 1}, s);
    if (i > 0) launch_gemm(g, GM_VTRSM, i, s);
  }
}

static void inverse_phase(b200gp_ctx* c, const Batch& b) {
  panel_schedule(
      c, b, [&](const Batch& g, int k0, int k1, cudaStream_t s) { inverse_panel(g, k0, k1, s); },
      [&](const Batch& g, int k0, int W, int j0, int j1, cudaStream_t s) { launch_vupdate(g, UpdRange{k0, W, j0, j1}, s); });
}

// Few slots of moderate order (the tail of a lock-step fit: a handful of items at N ~ 2048): every launch is a few CTAs
// on 148 SMs and the evaluation time is the length of the dependency chain, not the flops.  Panel p of L^-T only needs
// panel p of L, so it runs on a third stream as soon as the Cholesky has finished that panel, hidden behind the (longer)
// chain of the next Cholesky panel.
static bool pipelined_inverse(const b200gp_ctx* c, const Batch& b) { return b.nslots < c->la_slots && b.T >= 2 && b.T <= 32; }

static void cholesky_inverse_pipelined(b200gp_ctx* c, const Batch& b) {
  const int T = b.T, W = la_panel_width(c, b);
  const int npanels = (T + W - 1) / W;
  while ((int)c->evC.size() < npanels) {
    cudaEvent_t e;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    c->evC.push_back(e);
  }
  cholesky_phase_impl(c, b, [&](int k0, cudaStream_t s) { cudaEventRecord(c->evC[k0 / W], s); });
  cudaStream_t s2 = c->peer;
  for (int p = 0; p < npanels; p++) {
    const int k0 = p * W, k1 = (k0 + W < T) ? k0 + W : T;
    cudaStreamWaitEvent(s2, c->evC[p], 0);
    inverse_panel(b, k0, k1, s2);
    if (k1 < T) launch_vupdate(b, UpdRange{k0, k1 - k0, k1, T}, s2);
  }
  cudaEventRecord(c->evI, s2);
  cudaStreamWaitEvent(c->stream, c->evI, 0);
}

static int lml_grad_impl(b200gp_ctx* c, const int32_t* prog, const int32_t* prog_off, const double* theta,
                         const int32_t* theta_off, const int32_t* ids, int n, int with_grad, double* lml, double* dlml,
                         int32_t* status, int32_t* tries);

// Enqueue all phases for one chunk.  Phase events are recorded only when profiling is on.
static void run_pipeline(b200gp_ctx* c, const Batch& b, int with_grad) {
  cudaStream_t s = c->stream;
  const bool prof = c->profiling;
  cudaMemsetAsync(b.slot_status, 0, sizeof(int) * b.nslots, s);
  if (prof) cudaEventRecord(c->ev[0], s);
  launch_compile(b, s);
  launch_build_k(b, s);
  if (prof) cudaEventRecord(c->ev[1], s);
  if (pipelined_inverse(c, b)) {
    cholesky_inverse_pipelined(c, b);       // the "cholesky" phase time then covers both, "inverse" reads 0
    if (prof) cudaEventRecord(c->ev[2], s);
  } else {
    cholesky_phase(c, b);
    if (prof) cudaEventRecord(c->ev[2], s);
    inverse_phase(c, b);
  }
  if (prof) cudaEventRecord(c->ev[3], s);
  launch_solve(b, s);
  if (prof) cudaEventRecord(c->ev[4], s);
  if (with_grad) {
    if (c->fused_grad) launch_kinv_grad(b, s);
    else if (c->kinv64) launch_kinv64(b, s);
    else launch_gemm(b, GM_KINV, 0, s);
  }
  if (prof) cudaEventRecord(c->ev[5], s);
  if (with_grad) launch_grad(b, s, c->fused_grad);
  launch_finalize(b, with_grad, s);
  if (prof) cudaEventRecord(c->ev[6], s);
}

static void collect_phase_ms(b200gp_ctx* c) {
  if (!c->profiling) return;
  cudaEventSynchronize(c->ev[6]);
  for (int i = 0; i < 6; i++) cudaEventElapsedTime(&c->phase_ms[i], c->ev[i], c->ev[i + 1]);
  c->n_phase = 6;
}

extern "C" {

int b200gp_build_K(b200gp_ctx* c, const int32_t* prog, const int32_t* prog_off, const double* theta,
                   const int32_t* theta_off, const int32_t* ids, int n, int add_diag, double* K_out) {
  if (!c) return -1;
  if (!c->bound) return fail(c, "b200gp_build_K: call b200gp_bind first");
  if (!prog || !prog_off || !theta || !theta_off || !ids || !K_out) return fail(c, "b200gp_build_K: null pointer");
  CUCHECK(c, cudaSetDevice(c->device));
  const size_t n2 = (size_t)c->base.N * c->base.N;
  for (int off = 0; off < n; off += c->slots) {
    Batch b = c->base;
    b.nslots = (n - off < c->slots) ? n - off : c->slots;
    b.ids = ids + off;
    b.prog = reinterpret_cast<const int4*>(prog); b.prog_off = prog_off; b.theta = theta; b.theta_off = theta_off;
    b.quirks = c->quirks; b.lut = c->lut; b.lut_n = c->lut_n;
    b.add_diag = add_diag;
    launch_compile(b, c->stream);
    launch_build_k(b, c->stream);
    launch_export_sym(b, K_out + (size_t)off * n2, c->stream);
  }
  CUCHECK(c, cudaGetLastError());
  return 0;
}

int b200gp_lml_grad(b200gp_ctx* c, const int32_t* prog, const int32_t* prog_off, const double* theta,
                    const int32_t* theta_off, const int32_t* ids, int n, int with_grad, double* lml, double* dlml,
                    int32_t* status, int32_t* tries) {
  if (!c) return -1;
  if (!c->bound) return fail(c, "b200gp_lml_grad: call b200gp_bind first");
  if (!prog || !prog_off || !theta || !theta_off || !ids || !lml || !status || (with_grad && !dlml))
    return fail(c, "b200gp_lml_grad: null pointer");
  return lml_grad_impl(c, prog, prog_off, theta, theta_off, ids, n, with_grad, lml, dlml, status, tries);
}

}  // extern "C"

// Scoring pipeline + GPy's jitchol ladder for the items ids[0..n), chunked by the bound slot count.  On return the
// stream is synchronised and c->h_status[slot] holds the slot status of the last pipeline run of the last chunk.
static int lml_grad_impl(b200gp_ctx* c, const int32_t* prog, const int32_t* prog_off, const double* theta,
                         const int32_t* theta_off, const int32_t* ids, int n, int with_grad, double* lml, double* dlml,
                         int32_t* status, int32_t* tries) {
  CUCHECK(c, cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  for (int off = 0; off < n; off += c->slots) {
    Batch b = c->base;
    b.nslots = (n - off < c->slots) ? n - off : c->slots;
    b.ids = ids + off;
    b.prog = reinterpret_cast<const int4*>(prog); b.prog_off = prog_off; b.theta = theta; b.theta_off = theta_off;
    b.lml = lml; b.dlml = dlml; b.status = status; b.tries = tries; b.try_no = 0;
    b.jitter = nullptr;
    b.quirks = c->quirks; b.lut = c->lut; b.lut_n = c->lut_n;
    b.add_diag = 1;
    run_pipeline(c, b, with_grad);
    CUCHECK(c, cudaGetLastError());
    CUCHECK(c, cudaMemcpyAsync(c->h_status, b.slot_status, sizeof(int) * b.nslots, cudaMemcpyDeviceToHost, s));
    CUCHECK(c, cudaStreamSynchronize(s));
    collect_phase_ms(c);
    std::vector<int> failed;
    for (int i = 0; i < b.nslots; i++)
      if (c->h_status[i] != 0) failed.push_back(i);
    if (failed.empty()) continue;

    // GPy jitchol ladder (util/linalg.py): jitter = mean(diag(K_y)) * 1e-6 * 10^t, t = 0..4
    CUCHECK(c, cudaMemcpyAsync(c->h_ids, b.ids, sizeof(int) * b.nslots, cudaMemcpyDeviceToHost, s));
    CUCHECK(c, cudaMemcpyAsync(c->h_diag, b.diag_part, sizeof(double) * b.nslots * b.T, cudaMemcpyDeviceToHost, s));
    CUCHECK(c, cudaStreamSynchronize(s));
    std::vector<int> item(failed.size());
    std::vector<double> dmean(failed.size());
    for (size_t f = 0; f < failed.size(); f++) {
      item[f] = c->h_ids[failed[f]];
      double sum = 0.0;
      for (int k = 0; k < b.T; k++) sum += c->h_diag[failed[f] * b.T + k];
      dmean[f] = sum / b.N;
    }
    for (int t = 0; t < 5 && !item.empty(); t++) {
      std::vector<int> it2;
      std::vector<double> dm2;
      int m = 0;
      for (size_t f = 0; f < item.size(); f++) {
        const double jit = dmean[f] * 1e-6 * std::pow(10.0, t);
        if (!std::isfinite(jit)) continue;      // ladder stops for this item; status stays NOT_PD
        c->h_ids[m] = item[f];
        c->h_jit[m] = jit;
        it2.push_back(item[f]);
        dm2.push_back(dmean[f]);
        m++;
      }
      if (m == 0) break;
      CUCHECK(c, cudaMemcpyAsync(c->ws_ids, c->h_ids, sizeof(int) * m, cudaMemcpyHostToDevice, s));
      CUCHECK(c, cudaMemcpyAsync(c->ws_jitter, c->h_jit, sizeof(double) * m, cudaMemcpyHostToDevice, s));
      Batch bf = b;
      bf.nslots = m;
      bf.ids = c->ws_ids;
      bf.jitter = c->ws_jitter;
      bf.try_no = t + 1;
      run_pipeline(c, bf, with_grad);
      CUCHECK(c, cudaGetLastError());
      CUCHECK(c, cudaMemcpyAsync(c->h_status, bf.slot_status, sizeof(int) * m, cudaMemcpyDeviceToHost, s));
      CUCHECK(c, cudaStreamSynchronize(s));
      item.clear();
      dmean.clear();
      for (int i = 0; i < m; i++)
        if (c->h_status[i] != 0) { item.push_back(it2[i]); dmean.push_back(dm2[i]); }
    }
  }
  return 0;
}

extern "C" {

int b200gp_posterior(b200gp_ctx* c, const int32_t* prog, const int32_t* prog_off, const double* theta,
                     const int32_t* theta_off, const int32_t* d_id, int item, const double* dXs, int Ns,
                     int include_likelihood, double* mean, double* var, double* lml, int32_t* status) {
  if (!c) return -1;
  if (!c->bound) return fail(c, "b200gp_posterior: call b200gp_bind first");
  if (!prog || !prog_off || !theta || !theta_off || !d_id || !dXs || !mean || !var || !lml || !status)
    return fail(c, "b200gp_posterior: null pointer");
  if (Ns <= 0 || item < 0) return fail(c, "b200gp_posterior: need Ns > 0 and item >= 0");
  // factorise (with the jitter ladder) exactly as GP.parameters_changed does; slot 0 then holds V = L^-T and alpha
  int rc = lml_grad_impl(c, prog, prog_off, theta, theta_off, d_id, 1, 0, lml, nullptr, status, nullptr);
  if (rc != 0) return rc;
  const double nan_v = NAN;
  if (c->h_status[0] != 0) {        // not PD even with jitter (GPy raises LinAlgError): NaN predictions, status[item] says why
    std::vector<double> nv((size_t)Ns, nan_v);
    CUCHECK(c, cudaMemcpyAsync(mean, nv.data(), sizeof(double) * Ns, cudaMemcpyHostToDevice, c->stream));
    CUCHECK(c, cudaMemcpyAsync(var, nv.data(), sizeof(double) * Ns, cudaMemcpyHostToDevice, c->stream));
    CUCHECK(c, cudaStreamSynchronize(c->stream));
    return 0;
  }
  Batch b = c->base;
  b.nslots = 1;
  b.prog = reinterpret_cast<const int4*>(prog); b.prog_off = prog_off; b.theta = theta; b.theta_off = theta_off;
  b.lut = c->lut; b.lut_n = c->lut_n;
  const int chunk = b.Np < 4096 ? b.Np : 4096;       // K(x*, X) of a chunk lives in M1 of slot 0, scratch in Dinv
  for (int off = 0; off < Ns; off += chunk) {
    const int ns = Ns - off < chunk ? Ns - off : chunk;
    launch_predict(b, item, dXs + (size_t)off * b.D, ns, include_likelihood, b.Dinv, mean + off, var + off, c->stream);
  }
  CUCHECK(c, cudaGetLastError());
  CUCHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

size_t b200gp_potrf_work_bytes(int N) {
  if (N <= 0 || N % NB != 0) return 0;
  const size_t T = N / NB;
  return T * NB * NB * sizeof(double) + T * sizeof(double) + 1024;
}

int b200gp_potrf(b200gp_ctx* c, double* dA, int N, void* work, double* logdet, int32_t* status) {
  if (!c) return -1;
  if (!dA || !work || !logdet || !status) return fail(c, "b200gp_potrf: null pointer");
  if (N <= 0 || N % NB != 0) return fail(c, "b200gp_potrf: N must be a positive multiple of %d", NB);
  CUCHECK(c, cudaSetDevice(c->device));
  const int T = N / NB;
  Carver cv(work);
  Batch b{};
  b.N = N; b.Np = N; b.T = T; b.D = 0; b.nslots = 1; b.mstride = (size_t)N * N;
  b.M1 = dA; b.M2 = nullptr;
  b.gemm64 = c->gemm64;
  b.potrf_la = c->potrf_la;
  if (c->tma != 0 && tma_maps_bytes() <= sizeof c->tmaps_potrf &&
      tma_encode_maps(c->tmaps_potrf, dA, nullptr, N, (size_t)N, (c->tma == 4 ? 3 : c->tma) - 1) == 0) {
    b.tma = c->tma;
    b.tmaps = c->tmaps_potrf;
  }
  b.Dinv = cv.take<double>((size_t)T * NB * NB);
  b.logdet_part = cv.take<double>(T);
  b.slot_status = cv.take<int>(1);
  cudaStream_t s = c->stream;
  CUCHECK(c, cudaMemsetAsync(b.slot_status, 0, sizeof(int), s));
  if (c->profiling) cudaEventRecord(c->ev[0], s);
  cholesky_phase(c, b);
  if (c->profiling) cudaEventRecord(c->ev[1], s);
  CUCHECK(c, cudaGetLastError());
  std::vector<double> parts(T);
  int st = 0;
  CUCHECK(c, cudaMemcpyAsync(parts.data(), b.logdet_part, sizeof(double) * T, cudaMemcpyDeviceToHost, s));
  CUCHECK(c, cudaMemcpyAsync(&st, b.slot_status, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUCHECK(c, cudaStreamSynchronize(s));
  if (c->profiling) { cudaEventElapsedTime(&c->phase_ms[0], c->ev[0], c->ev[1]); c->n_phase = 1; }
  double ld = 0.0;
  if (st == 0) for (int k = 0; k < T; k++) ld += parts[k];
  *logdet = st == 0 ? ld : NAN;
  *status = st;
  return 0;
}

}  // extern "C"

extern "C" {

int b200gp_hellinger_precompute(b200gp_ctx* c, const double* dXsub, int n, int D, const int32_t* prog,
                                const int32_t* prog_off, const double* theta_samples, const int32_t* theta_off,
                                const double* lambda, int M, int S, double* logdet, double* G, int32_t* status) {
  if (!c) return -1;
  if (!dXsub || !prog || !prog_off || !theta_samples || !theta_off || !lambda || !logdet || !status)
    return fail(c, "b200gp_hellinger_precompute: null pointer");
  if (n <= 0 || n > NB) return fail(c, "b200gp_hellinger_precompute: subset size must be in [1, %d]", NB);
  if (D <= 0 || D > MAX_DIMS) return fail(c, "b200gp_hellinger_precompute: need 0 < D <= %d", MAX_DIMS);
  if (M <= 0 || S <= 0 || S > 65535) return fail(c, "b200gp_hellinger_precompute: bad M / S");
  CUCHECK(c, cudaSetDevice(c->device));
  for (int off = 0; off < M; off += 32768) {     // grid.y limit
    const int m = M - off < 32768 ? M - off : 32768;
    // prog_off / theta_off are absolute, so a chunk only shifts the model index base
    launch_hellinger_precompute(dXsub, n, D, reinterpret_cast<const int4*>(prog), prog_off + off, theta_samples,
                                theta_off + off, lambda, m, S, logdet + (size_t)off * S,
                                G ? G + (size_t)off * S * n * n : nullptr, status + (size_t)off * S, c->stream);
  }
  CUCHECK(c, cudaGetLastError());
  return 0;
}

int b200gp_hellinger_pairs(b200gp_ctx* c, const double* logdet, const double* G, int M, int S, int n,
                           const int32_t* pair_i, const int32_t* pair_j, int npairs, double tol_logdet, void* work,
                           double* dist, int32_t* status) {
  if (!c) return -1;
  if (!logdet || !G || !pair_i || !pair_j || !dist || !status || !work)
    return fail(c, "b200gp_hellinger_pairs: null pointer");
  if (n <= 0 || n > NB || M <= 0 || S <= 0 || S > 65535) return fail(c, "b200gp_hellinger_pairs: bad geometry");
  if (npairs <= 0) return 0;
  CUCHECK(c, cudaSetDevice(c->device));
  Carver cv(work);
  double* h = cv.take<double>((size_t)npairs * S);
  int* hs = cv.take<int>((size_t)npairs * S);
  launch_hellinger_pairs(logdet, G, S, n, pair_i, pair_j, npairs, tol_logdet, h, hs, dist, status, c->stream);
  CUCHECK(c, cudaGetLastError());
  return 0;
}

int b200gp_gram_logdet(b200gp_ctx* c, const int32_t* prog, const int32_t* prog_off, const double* theta,
                       const int32_t* theta_off, const int32_t* ids, int n, int diag_mode, double* K_out, double* logdet,
                       int32_t* status) {
  if (!c) return -1;
  if (!c->bound) return fail(c, "b200gp_gram_logdet: call b200gp_bind first");
  if (!prog || !prog_off || !theta || !theta_off || !ids || (logdet && !status))
    return fail(c, "b200gp_gram_logdet: null pointer");
  if (diag_mode < 0 || diag_mode > 2) return fail(c, "b200gp_gram_logdet: diag_mode must be 0, 1 or 2");
  if (!K_out && !logdet) return 0;
  CUCHECK(c, cudaSetDevice(c->device));
  const size_t n2 = (size_t)c->base.N * c->base.N;
  for (int off = 0; off < n; off += c->slots) {
    Batch b = c->base;
    b.nslots = (n - off < c->slots) ? n - off : c->slots;
    b.ids = ids + off;
    b.prog = reinterpret_cast<const int4*>(prog); b.prog_off = prog_off; b.theta = theta; b.theta_off = theta_off;
    b.quirks = c->quirks; b.lut = c->lut; b.lut_n = c->lut_n;
    b.add_diag = diag_mode;
    launch_compile(b, c->stream);
    launch_build_k(b, c->stream);
    if (K_out) launch_export_sym(b, K_out + (size_t)off * n2, c->stream);
    if (logdet) {
      CUCHECK(c, cudaMemsetAsync(b.slot_status, 0, sizeof(int) * b.nslots, c->stream));
      cholesky_phase(c, b);
      launch_slot_logdet(b, logdet, status, c->stream);
    }
  }
  CUCHECK(c, cudaGetLastError());
  return 0;
}

size_t b200gp_hellinger_pairs_large_work_bytes(int npairs, int S) {
  if (npairs <= 0 || S <= 0) return 0;
  return (size_t)npairs * S * (sizeof(double) + sizeof(int)) + 1024;
}

int b200gp_hellinger_pairs_large(b200gp_ctx* c, const double* logdet, const double* G, int M, int S, int n,
                                 const int32_t* pair_i, const int32_t* pair_j, int npairs, const int32_t* work_pair,
                                 const int32_t* work_s, int nwork, void* work, double* dist, int32_t* status) {
  if (!c) return -1;
  if (!c->bound) return fail(c, "b200gp_hellinger_pairs_large: call b200gp_bind (on the subset) first");
  if (!logdet || !G || !pair_i || !pair_j || !dist || !status || !work || (nwork > 0 && (!work_pair || !work_s)))
    return fail(c, "b200gp_hellinger_pairs_large: null pointer");
  if (n != c->base.N) return fail(c, "b200gp_hellinger_pairs_large: the context is bound to %d points, the Gram matrices have %d", c->base.N, n);
  if (M <= 0 || S <= 0 || nwork < 0) return fail(c, "b200gp_hellinger_pairs_large: bad geometry");
  if (npairs <= 0) return 0;
  CUCHECK(c, cudaSetDevice(c->device));
  Carver cv(work);
  double* ldpq = cv.take<double>((size_t)npairs * S);
  int* hs = cv.take<int>((size_t)npairs * S);
  cudaStream_t s = c->stream;
  launch_hpair_init(logdet, S, pair_i, npairs, ldpq, hs, s);
  for (int off = 0; off < nwork; off += c->slots) {
    Batch b = c->base;
    b.nslots = (nwork - off < c->slots) ? nwork - off : c->slots;
    launch_havg_build(b, G, S, n, pair_i, pair_j, work_pair + off, work_s + off, s);
    CUCHECK(c, cudaMemsetAsync(b.slot_status, 0, sizeof(int) * b.nslots, s));
    cholesky_phase(c, b);
    launch_hpair_scatter(b, S, work_pair + off, work_s + off, ldpq, hs, s);
  }
  launch_hpair_finish(logdet, S, pair_i, pair_j, npairs, ldpq, hs, dist, status, s);
  CUCHECK(c, cudaGetLastError());
  return 0;
}

int b200gp_alignment_min_slots(int N, int n) {
  if (N <= 0 || n <= 0) return 0;
  const size_t Np = round_up(N, NB), mp = round_up(n, NB), nsplit = (N + 15) / 16;
  const size_t for_partials = (nsplit * mp * mp + Np * Np - 1) / (Np * Np);       // split-K partials live in M2
  const size_t for_products = ((size_t)n * n + (Np / NB) * NB * NB - 1) / ((Np / NB) * NB * NB);   // centred products in Dinv
  size_t need = (size_t)n;
  if (for_partials > need) need = for_partials;
  if (for_products > need) need = for_products;
  return (int)need;
}

int b200gp_centered_alignments(b200gp_ctx* c, const int32_t* prog, const int32_t* prog_off, const double* theta,
                               const int32_t* theta_off, const int32_t* ids, int n, double* out) {
  if (!c) return -1;
  if (!c->bound) return fail(c, "b200gp_centered_alignments: call b200gp_bind first");
  if (!prog || !prog_off || !theta || !theta_off || !ids || !out) return fail(c, "b200gp_centered_alignments: null pointer");
  if (n <= 0) return 0;
  if (b200gp_alignment_min_slots(c->base.N, n) > c->slots)
    return fail(c, "b200gp_centered_alignments: %d kernels of order %d need %d slots, %d are bound", n, c->base.N,
                b200gp_alignment_min_slots(c->base.N, n), c->slots);
  CUCHECK(c, cudaSetDevice(c->device));
  Batch b = c->base;
  b.nslots = n;
  b.ids = ids;
  b.prog = reinterpret_cast<const int4*>(prog); b.prog_off = prog_off; b.theta = theta; b.theta_off = theta_off;
  b.quirks = c->quirks; b.lut = c->lut; b.lut_n = c->lut_n;
  b.add_diag = 0;
  launch_compile(b, c->stream);
  launch_build_k(b, c->stream);
  b.nslots = c->slots;            // scratch capacity: every bound slot
  const int rc = launch_alignments(b, n, out, c->stream);
  if (rc != 0) return fail(c, "b200gp_centered_alignments: scratch does not fit the bound workspace (code %d)", rc);
  CUCHECK(c, cudaGetLastError());
  return 0;
}

int b200gp_vector_pairs(b200gp_ctx* c, const double* G, int M, int S, int n, const int32_t* pair_i, const int32_t* pair_j,
                        int npairs, int metric, double* dist) {
  if (!c) return -1;
  if (!G || !pair_i || !pair_j || !dist) return fail(c, "b200gp_vector_pairs: null pointer");
  if (n <= 0 || M <= 0 || S <= 0 || (metric != 0 && metric != 1)) return fail(c, "b200gp_vector_pairs: bad geometry / metric");
  if (npairs <= 0) return 0;
  CUCHECK(c, cudaSetDevice(c->device));
  launch_vector_pairs(G, S, n, pair_i, pair_j, npairs, metric, dist, c->stream);
  CUCHECK(c, cudaGetLastError());
  return 0;
}

size_t b200gp_hellinger_pairs_work_bytes(int npairs, int S) {
  if (npairs <= 0 || S <= 0) return 0;
  return (size_t)npairs * S * (sizeof(double) + sizeof(int)) + 1024;
}

}  // extern "C"
