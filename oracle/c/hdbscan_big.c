/* ORACLE (test infrastructure, not product code).
 *
 * Multithreaded versions of oracle/c/hdbscan_ref.c for the sizes the hot path is benchmarked at (BASELINE
 * configs[2]: ~350 k thing points x 5-D per FOR-instance cylinder; reference call site
 * torch_points3d/utils/hdbscan_cluster.py:8-13,117-167).  Same definitions, same float64 operation order
 * (build with -ffp-contract=off), so the results are bit-identical to the single-thread functions:
 *
 *   big_core_distances : brute-force k-th nearest sample counting the sample itself, rows split over threads
 *   big_mst_total_order: Prim on the complete mutual-reachability graph under the strict total order
 *                        (w, min(a,b), max(a,b)); every Prim step updates the frontier in parallel
 *                        (each thread owns a slice of the points that are not in the tree yet) and the
 *                        per-thread minima meet behind a spin barrier.
 *
 * 350 k x 5-D: ~2 x 6e10 pair evaluations -- about a minute on 8 cores; used by scripts/make_golden_hdbscan_big.py
 * to freeze (labels, sha256 of the MST) under tests/golden/, and by tests at 50 k.
 */
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>

static inline double sqd(const double* a, const double* b, int64_t D) {
  double s = 0.0;
  for (int64_t j = 0; j < D; ++j) { const double t = a[j] - b[j]; s += t * t; }
  return s;
}
static inline int kless(double w1, int64_t a1, int64_t b1, double w2, int64_t a2, int64_t b2) {
  if (w1 != w2) return w1 < w2;
  if (a1 != a2) return a1 < a2;
  return b1 < b2;
}

/* ---------------------------------------------------------------- core distances */
typedef struct { const double* X; int64_t n, D, k, lo, hi; double* core; } core_job_t;

static void* core_worker(void* p) {
  core_job_t* J = (core_job_t*)p;
  const int64_t k = J->k, n = J->n, D = J->D;
  double* best = (double*)malloc(sizeof(double) * (size_t)k);
  for (int64_t i = J->lo; i < J->hi; ++i) {
    int64_t m = 0;
    for (int64_t j = 0; j < n; ++j) {
      const double d2 = sqd(J->X + i * D, J->X + j * D, D);
      if (m < k) {
        int64_t q = m++;
        while (q > 0 && best[q - 1] > d2) { best[q] = best[q - 1]; --q; }
        best[q] = d2;
      } else if (d2 < best[k - 1]) {
        int64_t q = k - 1;
        while (q > 0 && best[q - 1] > d2) { best[q] = best[q - 1]; --q; }
        best[q] = d2;
      }
    }
    J->core[i] = sqrt(best[k - 1]);
  }
  free(best);
  return 0;
}

int big_core_distances(const double* X, int64_t n, int64_t D, int64_t k, double* core, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  core_job_t jobs[256];
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = (core_job_t){X, n, D, k, n * t / nthreads, n * (t + 1) / nthreads, core};
    if (pthread_create(&th[t], 0, core_worker, &jobs[t])) return 1;
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
  return 0;
}

/* ---------------------------------------------------------------- Prim */
typedef struct { double w; int64_t a, b; } edge_t;
static int edge_cmp(const void* x, const void* y) {
  const edge_t* e = (const edge_t*)x; const edge_t* f = (const edge_t*)y;
  if (kless(e->w, e->a, e->b, f->w, f->a, f->b)) return -1;
  if (kless(f->w, f->a, f->b, e->w, e->a, e->b)) return 1;
  return 0;
}

typedef struct {
  atomic_int count;
  atomic_int sense;
  int n;
} barrier_t;
static inline void bar_wait(barrier_t* B, int* local) {
  *local = !*local;
  if (atomic_fetch_add(&B->count, 1) == B->n - 1) {
    atomic_store(&B->count, 0);
    atomic_store(&B->sense, *local);
  } else {
    while (atomic_load(&B->sense) != *local) { __builtin_ia32_pause(); }
  }
}

typedef struct {
  const double* X; const double* core; int64_t n, D; double alpha;
  double* bw; int64_t* ba; int64_t* bb;      /* best edge into the tree, per point */
  edge_t* edges;
  barrier_t* bar;
  int nthreads;
  /* per-thread proposals, padded against false sharing */
  struct { double w; int64_t a, b, j; char pad[32]; }* prop;
  volatile int64_t* cur;
} prim_shared_t;
typedef struct { prim_shared_t* S; int tid; } prim_job_t;

static void* prim_worker(void* p) {
  prim_job_t* J = (prim_job_t*)p;
  prim_shared_t* S = J->S;
  const int tid = J->tid, T = S->nthreads;
  const int64_t n = S->n, D = S->D;
  const int64_t lo = n * tid / T, hi = n * (tid + 1) / T;
  /* points of my slice that are not in the tree yet (swap-remove) */
  int64_t* rem = (int64_t*)malloc(sizeof(int64_t) * (size_t)(hi - lo + 1));
  int64_t m = 0;
  for (int64_t j = lo; j < hi; ++j) if (j != 0) rem[m++] = j;
  int local = 0;
  for (int64_t i = 0; i < n - 1; ++i) {
    const int64_t cur = *S->cur;
    const double* xc = S->X + cur * D;
    const double cc = S->core[cur];
    double mw = INFINITY; int64_t ma = INT64_MAX, mb = INT64_MAX, mj = -1, mpos = -1;
    for (int64_t q = 0; q < m; ++q) {
      const int64_t j = rem[q];
      double d = sqrt(sqd(xc, S->X + j * D, D)) / S->alpha;
      double w = cc > S->core[j] ? cc : S->core[j];
      if (d > w) w = d;
      const int64_t a = cur < j ? cur : j, b = cur < j ? j : cur;
      if (kless(w, a, b, S->bw[j], S->ba[j], S->bb[j])) { S->bw[j] = w; S->ba[j] = a; S->bb[j] = b; }
      if (kless(S->bw[j], S->ba[j], S->bb[j], mw, ma, mb)) { mw = S->bw[j]; ma = S->ba[j]; mb = S->bb[j]; mj = j; mpos = q; }
    }
    S->prop[tid].w = mw; S->prop[tid].a = ma; S->prop[tid].b = mb; S->prop[tid].j = mj;
    bar_wait(S->bar, &local);
    /* every thread reduces the proposals the same way (no second broadcast needed for the winner's owner) */
    int win = -1;
    for (int t = 0; t < T; ++t) {
      if (S->prop[t].j < 0) continue;
      if (win < 0 || kless(S->prop[t].w, S->prop[t].a, S->prop[t].b, S->prop[win].w, S->prop[win].a, S->prop[win].b)) win = t;
    }
    if (win == tid) { rem[mpos] = rem[--m]; }
    if (tid == 0) {
      S->edges[i].w = S->prop[win].w; S->edges[i].a = S->prop[win].a; S->edges[i].b = S->prop[win].b;
      *S->cur = S->prop[win].j;
    }
    bar_wait(S->bar, &local);
  }
  free(rem);
  return 0;
}

int big_mst_total_order(const double* X, const double* core, int64_t n, int64_t D, double alpha,
                        int64_t* u, int64_t* v, double* w, int nthreads) {
  if (n < 2) return 0;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (nthreads > n) nthreads = (int)n;
  prim_shared_t S;
  S.X = X; S.core = core; S.n = n; S.D = D; S.alpha = alpha; S.nthreads = nthreads;
  S.bw = (double*)malloc(sizeof(double) * (size_t)n);
  S.ba = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  S.bb = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  S.edges = (edge_t*)malloc(sizeof(edge_t) * (size_t)(n - 1));
  S.prop = malloc(sizeof(*S.prop) * (size_t)nthreads);
  barrier_t bar; atomic_init(&bar.count, 0); atomic_init(&bar.sense, 0); bar.n = nthreads;
  S.bar = &bar;
  volatile int64_t cur = 0;
  S.cur = &cur;
  if (!S.bw || !S.ba || !S.bb || !S.edges || !S.prop) return 1;
  for (int64_t j = 0; j < n; ++j) { S.bw[j] = INFINITY; S.ba[j] = S.bb[j] = INT64_MAX; }
  pthread_t th[256];
  prim_job_t jobs[256];
  for (int t = 0; t < nthreads; ++t) {
    jobs[t].S = &S; jobs[t].tid = t;
    if (pthread_create(&th[t], 0, prim_worker, &jobs[t])) return 1;
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
  qsort(S.edges, (size_t)(n - 1), sizeof(edge_t), edge_cmp);
  for (int64_t i = 0; i < n - 1; ++i) { u[i] = S.edges[i].a; v[i] = S.edges[i].b; w[i] = S.edges[i].w; }
  free(S.bw); free(S.ba); free(S.bb); free(S.edges); free(S.prop);
  return 0;
}
