/*
 * asm_oracle.c -- CPU restatement of the ASM convergence-diagnostic hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under asm_b200/ may include, link, import or call
 * this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference (rbouckaert/asm, Java) ships no tests, golden vectors
 * or recorded logs, Java is not installed here, and its tree metric (BEASTLabs
 * RNNIMetric, `atleast 2.0.0`, un-vendored) is absent from /root/reference.  This file
 * therefore follows the cited Java line by line (same loop order, same operand order,
 * no fused multiply-add: build with -ffp-contract=off) and is cross-checked only
 * against independent restatements and hand-derived known answers (tests/).
 *
 * Each function cites the reference lines it follows, relative to /root/reference.
 * The tree distance is Robinson-Foulds on clade sets (un-halved symmetric difference),
 * which BASELINE.json's north_star substitutes for RNNIMetric at the single call site
 * src/asm/inference/TreeESS.java:179-180.  It is computed here by sorted-set
 * intersection on clade identities, independently of any bitset encoding.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_LAG 2000 /* TraceESS.java:21 */

/* ------------------------------------------------------------------------------------------
 * Clade identity: the reference identifies a clade by the string of its sorted leaf numbers,
 * "3,7,12," (CladeDifference.java:137-141).  A string-keyed dictionary hands out ids in
 * first-seen order.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  char **keys;    /* id -> key string */
  int32_t *sizes; /* id -> number of leaves in the clade */
  int64_t n, cap;
  int64_t *slots; /* open addressing, -1 = empty */
  int64_t n_slots;
} orc_dict;

static uint64_t fnv1a(const char *s) {
  uint64_t h = 1469598103934665603ULL;
  for (; *s; ++s) {
    h ^= (unsigned char)*s;
    h *= 1099511628211ULL;
  }
  return h;
}

static void dict_init(orc_dict *d) {
  d->n = 0;
  d->cap = 1024;
  d->keys = (char **)malloc(sizeof(char *) * d->cap);
  d->sizes = (int32_t *)malloc(sizeof(int32_t) * d->cap);
  d->n_slots = 4096;
  d->slots = (int64_t *)malloc(sizeof(int64_t) * d->n_slots);
  for (int64_t i = 0; i < d->n_slots; i++) d->slots[i] = -1;
}

static void dict_free(orc_dict *d) {
  for (int64_t i = 0; i < d->n; i++) free(d->keys[i]);
  free(d->keys);
  free(d->sizes);
  free(d->slots);
}

static void dict_rehash(orc_dict *d) {
  int64_t ns = d->n_slots * 2;
  int64_t *s = (int64_t *)malloc(sizeof(int64_t) * ns);
  for (int64_t i = 0; i < ns; i++) s[i] = -1;
  for (int64_t id = 0; id < d->n; id++) {
    uint64_t h = fnv1a(d->keys[id]) & (uint64_t)(ns - 1);
    while (s[h] >= 0) h = (h + 1) & (uint64_t)(ns - 1);
    s[h] = id;
  }
  free(d->slots);
  d->slots = s;
  d->n_slots = ns;
}

static int64_t dict_get_or_add(orc_dict *d, const char *key, int32_t size) {
  uint64_t h = fnv1a(key) & (uint64_t)(d->n_slots - 1);
  while (d->slots[h] >= 0) {
    if (strcmp(d->keys[d->slots[h]], key) == 0) return d->slots[h];
    h = (h + 1) & (uint64_t)(d->n_slots - 1);
  }
  if (d->n == d->cap) {
    d->cap *= 2;
    d->keys = (char **)realloc(d->keys, sizeof(char *) * d->cap);
    d->sizes = (int32_t *)realloc(d->sizes, sizeof(int32_t) * d->cap);
  }
  int64_t id = d->n++;
  d->keys[id] = strdup(key);
  d->sizes[id] = size;
  d->slots[h] = id;
  if (d->n * 2 > d->n_slots) dict_rehash(d);
  return id;
}

/* ------------------------------------------------------------------------------------------
 * Tree store: the oracle's stand-in for TraceInfo.trees (TraceInfo.java:27).  Every tree is
 * kept as the ascending list of its clade ids (one per internal node, root included), which is
 * all that the RF distance needs.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_chains, n_taxa;
  orc_dict dict;
  int64_t *n_trees;  /* [n_chains] */
  int64_t *cap;      /* [n_chains] */
  int32_t **ids;     /* [n_chains] -> [n_trees][n_taxa-1] ascending clade ids */
  int32_t **raw_ids; /* [n_chains] -> [n_trees][n_taxa-1] clade ids in traverse order */
} orc_treeset;

orc_treeset *orc_treeset_new(int32_t n_chains, int32_t n_taxa) {
  orc_treeset *ts = (orc_treeset *)calloc(1, sizeof(orc_treeset));
  ts->n_chains = n_chains;
  ts->n_taxa = n_taxa;
  dict_init(&ts->dict);
  ts->n_trees = (int64_t *)calloc(n_chains, sizeof(int64_t));
  ts->cap = (int64_t *)calloc(n_chains, sizeof(int64_t));
  ts->ids = (int32_t **)calloc(n_chains, sizeof(int32_t *));
  ts->raw_ids = (int32_t **)calloc(n_chains, sizeof(int32_t *));
  return ts;
}

void orc_treeset_free(orc_treeset *ts) {
  if (!ts) return;
  for (int c = 0; c < ts->n_chains; c++) {
    free(ts->ids[c]);
    free(ts->raw_ids[c]);
  }
  dict_free(&ts->dict);
  free(ts->n_trees);
  free(ts->cap);
  free(ts->ids);
  free(ts->raw_ids);
  free(ts);
}

typedef struct {
  char *buf;
  size_t cap;
} strbuf;

/* CladeDifference.java:107-144 -- post-order traverse.  A leaf yields {getNr()}; an internal
 * node yields the merge of its children's sorted leaf lists and emits the key string
 * "a,b,c," for that clade (root included).  Returns the clade length; `clade` receives it. */
static int traverse(const int32_t *left, const int32_t *right, int node, int32_t *clade, int32_t *scratch,
                    orc_dict *dict, int32_t *out_ids, int *n_out, strbuf *sb) {
  if (left[node] < 0) { /* node.isLeaf()  :109-111 */
    clade[0] = node;    /* getNr(): leaves are numbered 0..n-1 */
    return 1;
  }
  /* :113-114 */
  int nl = traverse(left, right, left[node], clade, scratch, dict, out_ids, n_out, sb);
  int32_t *rc = clade + nl;
  int nr = traverse(left, right, right[node], rc, scratch, dict, out_ids, n_out, sb);
  /* :117-136 merge (written as a plain two-way merge of two ascending lists) */
  int i = 0, il = 0, ir = 0;
  while (il < nl && ir < nr) scratch[i++] = clade[il] < rc[ir] ? clade[il++] : rc[ir++];
  while (il < nl) scratch[i++] = clade[il++];
  while (ir < nr) scratch[i++] = rc[ir++];
  memcpy(clade, scratch, sizeof(int32_t) * (size_t)(nl + nr));
  /* :137-141 key string */
  size_t need = (size_t)(nl + nr) * 12 + 1;
  if (sb->cap < need) {
    sb->cap = need * 2;
    sb->buf = (char *)realloc(sb->buf, sb->cap);
  }
  char *p = sb->buf;
  for (i = 0; i < nl + nr; i++) p += sprintf(p, "%d,", clade[i]);
  *p = 0;
  out_ids[(*n_out)++] = (int32_t)dict_get_or_add(dict, sb->buf, nl + nr);
  return nl + nr;
}

static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

/* Append trees to a chain.  left/right are [n_trees][2*n_taxa-1] child indices (-1 for the
 * leaves 0..n_taxa-1), root is [n_trees].  Mirrors traceInfo.trees[chain].add(tree)
 * (AutoStopMCMC.java:536) followed by the clade extraction of CladeDifference.process
 * (CladeDifference.java:147-159).  Returns 0, or -1 if a tree is not a binary tree on
 * n_taxa leaves. */
int32_t orc_treeset_add(orc_treeset *ts, int32_t chain, int64_t n_trees, const int32_t *left, const int32_t *right,
                        const int32_t *root) {
  int n = ts->n_taxa, nn = 2 * n - 1, k = n - 1;
  int64_t have = ts->n_trees[chain];
  if (have + n_trees > ts->cap[chain]) {
    int64_t nc = ts->cap[chain] ? ts->cap[chain] : 64;
    while (nc < have + n_trees) nc *= 2;
    ts->ids[chain] = (int32_t *)realloc(ts->ids[chain], sizeof(int32_t) * (size_t)nc * k);
    ts->raw_ids[chain] = (int32_t *)realloc(ts->raw_ids[chain], sizeof(int32_t) * (size_t)nc * k);
    ts->cap[chain] = nc;
  }
  int32_t *clade = (int32_t *)malloc(sizeof(int32_t) * n * 2);
  int32_t *scratch = (int32_t *)malloc(sizeof(int32_t) * n);
  strbuf sb = {NULL, 0};
  int rc = 0;
  for (int64_t t = 0; t < n_trees; t++) {
    int n_out = 0;
    int32_t *dst = ts->raw_ids[chain] + (size_t)(have + t) * k;
    int len = traverse(left + t * nn, right + t * nn, root[t], clade, scratch, &ts->dict, dst, &n_out, &sb);
    if (len != n || n_out != k) {
      rc = -1;
      break;
    }
    int32_t *srt = ts->ids[chain] + (size_t)(have + t) * k;
    memcpy(srt, dst, sizeof(int32_t) * k);
    qsort(srt, k, sizeof(int32_t), cmp_i32);
    ts->n_trees[chain] = have + t + 1;
  }
  free(clade);
  free(scratch);
  free(sb.buf);
  return rc;
}

int64_t orc_num_clades(const orc_treeset *ts) { return ts->dict.n; }
int64_t orc_num_trees(const orc_treeset *ts, int32_t chain) { return ts->n_trees[chain]; }

/* Key string of clade `id` (the reference's own representation, CladeDifference.java:137-141). */
int32_t orc_clade_key(const orc_treeset *ts, int64_t id, char *buf, int64_t cap) {
  if (id < 0 || id >= ts->dict.n) return -1;
  size_t len = strlen(ts->dict.keys[id]);
  if ((int64_t)len + 1 > cap) return -2;
  memcpy(buf, ts->dict.keys[id], len + 1);
  return (int32_t)len;
}

int32_t orc_clade_size(const orc_treeset *ts, int64_t id) { return ts->dict.sizes[id]; }

/* Clade ids of one tree in traverse order (post-order, children before parent). */
void orc_tree_clade_ids(const orc_treeset *ts, int32_t chain, int64_t idx, int32_t *out) {
  memcpy(out, ts->raw_ids[chain] + (size_t)idx * (ts->n_taxa - 1), sizeof(int32_t) * (size_t)(ts->n_taxa - 1));
}

/* Per-chain clade counts over trees [from,to): CladeDifference.process (:147-159) counts, for
 * each chain, how many trees contain each clade.  out is [orc_num_clades]. */
void orc_clade_counts(const orc_treeset *ts, int32_t chain, int64_t from, int64_t to, int32_t *out) {
  int k = ts->n_taxa - 1;
  memset(out, 0, sizeof(int32_t) * (size_t)ts->dict.n);
  for (int64_t t = from; t < to; t++) {
    const int32_t *ids = ts->raw_ids[chain] + (size_t)t * k;
    for (int j = 0; j < k; j++) out[ids[j]]++;
  }
}

/* Robinson-Foulds distance |A|+|B|-2|A n B| on the clade sets of two trees, by merge
 * intersection of the ascending id lists. */
static inline int rf_sorted(const int32_t *a, const int32_t *b, int k) {
  int i = 0, j = 0, common = 0;
  while (i < k && j < k) {
    int32_t x = a[i], y = b[j];
    common += (x == y);
    i += (x <= y);
    j += (y <= x);
  }
  return 2 * k - 2 * common;
}

int32_t orc_rf(const orc_treeset *ts, int32_t c1, int64_t i1, int32_t c2, int64_t i2) {
  int k = ts->n_taxa - 1;
  return rf_sorted(ts->ids[c1] + (size_t)i1 * k, ts->ids[c2] + (size_t)i2 * k, k);
}

/* TreeESS.java:168-186 distancePlusOne, with RF in place of RNNIMetric (:179-180).  The
 * cache (:172-175,181) only memoises, so it is not restated.  Returns float as the Java does. */
static inline float distance_plus_one(const orc_treeset *ts, int ts1, int64_t i1, int ts2, int64_t i2) {
  if (ts1 == ts2 && i1 == i2) return 0 + 1; /* :169-171 */
  float d = (float)orc_rf(ts, ts1, i1, ts2, i2) + 1; /* :180 */
  return d;
}

float orc_distance_plus_one(const orc_treeset *ts, int32_t ts1, int64_t i1, int32_t ts2, int64_t i2) {
  return distance_plus_one(ts, ts1, i1, ts2, i2);
}

/* RF block probe: out[(r-r0)*(c1-c0)+(c-c0)] = RF(tree[chainA][r], tree[chainB][c]). */
void orc_rf_block(const orc_treeset *ts, int32_t chainA, int64_t r0, int64_t r1, int32_t chainB, int64_t c0,
                  int64_t c1, uint16_t *out) {
  int64_t nc = c1 - c0;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t r = r0; r < r1; r++)
    for (int64_t c = c0; c < c1; c++) out[(r - r0) * nc + (c - c0)] = (uint16_t)orc_rf(ts, chainA, r, chainB, c);
}

/* TreePSRF.java:296-301 start(): burn-in rounded up to a multiple of delta. */
static inline int psrf_start(int start, int delta) {
  if (start % delta != 0) start = start + delta - start % delta;
  return start;
}

/* TreePSRF.java:274-294 calcPSRF(treeSet1, treeSet2, k, burnin, end).  var_out (optional)
 * receives {varIn, varBetween}. */
static double calc_psrf(const orc_treeset *ts, int treeSet1, int treeSet2, int k, const int32_t *burnin, int end,
                        int delta, double *var_out) {
  double varIn = 0;
  int start = psrf_start(burnin[treeSet1], delta);
  for (int i = start; i < end; i += delta) {
    double d = distance_plus_one(ts, treeSet1, k, treeSet1, i);
    varIn += d * d;
  }
  double varBetween = 0;
  start = psrf_start(burnin[treeSet2], delta);
  for (int i = start; i < end; i += delta) {
    double d = distance_plus_one(ts, treeSet1, k, treeSet2, i);
    varBetween += d * d;
  }
  double psrf;
  if (varIn != 0) {
    psrf = sqrt(varBetween / varIn);
  } else {
    psrf = sqrt(varBetween);
  }
  if (var_out) {
    var_out[0] = varIn;
    var_out[1] = varBetween;
  }
  return psrf;
}

double orc_calc_psrf(const orc_treeset *ts, int32_t treeSet1, int32_t treeSet2, int32_t k, const int32_t *burnin,
                     int32_t end, int32_t delta, double *var_out) {
  return calc_psrf(ts, treeSet1, treeSet2, k, burnin, end, delta, var_out);
}

/* TreePSRF.java:264-271 mean(): sequential sum, then one divide (0/0 = NaN for no rows). */
static double psrf_mean(const double *psrf, int n) {
  double sum = 0;
  for (int i = 0; i < n; i++) sum += psrf[i];
  return sum / n;
}

/* The row loop of converged2sided (TreePSRF.java:141-145 / :154-157): psrf[(x-start)/delta] =
 * calcPSRF(ts1, ts2, x) for x in [start,end) step delta, then mean().  Rows are independent, so
 * they run under OpenMP; the mean is taken sequentially afterwards as the Java does.
 * psrf_rows (optional) receives the per-row values; sums (optional) receives
 * [(end-start)/delta][2] = {varIn, varBetween}. */
double orc_psrf_rows_mean(const orc_treeset *ts, int32_t ts1, int32_t ts2, int32_t start, const int32_t *burnin,
                          int32_t end, int32_t delta, double *psrf_rows, double *sums) {
  int n = (end - start) / delta; /* :141 */
  if (n < 0) n = 0;
  double *psrf = psrf_rows ? psrf_rows : (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
#pragma omp parallel for schedule(dynamic, 8)
  for (int x = start; x < end; x += delta) {
    double v[2];
    int r = (x - start) / delta;
    if (r < n) {
      psrf[r] = calc_psrf(ts, ts1, ts2, x, burnin, end, delta, v);
      if (sums) {
        sums[2 * r] = v[0];
        sums[2 * r + 1] = v[1];
      }
    }
  }
  double m = psrf_mean(psrf, n);
  if (!psrf_rows) free(psrf);
  return m;
}

/* Criterion state + inputs of TreePSRF (TreePSRF.java:15-38, TreeESS.java:18-41). */
typedef struct {
  /* inputs (Appendix A.4 of SURVEY.md; defaults TreeESS.java:18-26, TreePSRF.java:15-20) */
  double smoothing; /* 1.0 */
  double b;         /* 0.05 */
  int32_t targetESS;  /* 100 */
  int32_t cacheLimit; /* 1024 */
  int32_t sampleSize; /* 10 */
  int32_t twoSided;   /* true */
  int32_t checkESS;   /* false */
  /* state */
  int32_t delta;  /* TreeESS.java:40 */
  int32_t start0; /* TreePSRF.java:28 */
  int32_t prev;   /* TreePSRF.java:31 */
  int32_t n_chains;
  double grValues[64 * 64]; /* TreePSRF.java:51,315-317; [a*n_chains+b] for the C>2 generalisation */
  int32_t n_rows_last;      /* bookkeeping for tests: rows used by the last evaluation */
  int32_t threw;            /* set when the Java would have caught a Throwable (:188-194) */
  int32_t have_indices;     /* TreeESS.java:115 `indices == null` */
  int32_t indices[1024];    /* TreeESS.java:115 */
} orc_psrf_state;

/* TreePSRF.initAndValidate (TreePSRF.java:54-93).  Returns 0 or a negative code for each
 * IllegalArgumentException site. */
int32_t orc_psrf_init(orc_psrf_state *st) {
  if (!(st->smoothing >= 0.0 && 1.0 >= st->smoothing)) return -1; /* :56-58 */
  if (st->targetESS <= 0) return -2;                              /* :61-63 */
  if (st->cacheLimit <= 0) return -3;                             /* :66-68 */
  if (st->cacheLimit <= st->targetESS) return -4;                 /* :69-72 */
  if (st->sampleSize <= 0) return -5;                             /* :79-81 */
  if (st->sampleSize >= st->targetESS) return -6;                 /* :82-84 */
  st->delta = 1;
  st->start0 = -1;
  st->prev = 0;
  st->threw = 0;
  st->have_indices = 0;
  if (st->sampleSize > 1024) return -7; /* oracle capacity, not a reference rule */
  return 0;
}

/* TreePSRF.setup (TreePSRF.java:304-318).  The reference throws for nChains != 2 (:311-313);
 * the generalisation to more chains is accepted here and flagged
 * by the return value 1 so tests can tell the two regimes apart. */
int32_t orc_psrf_setup(orc_psrf_state *st, int32_t nChains) {
  if (nChains < 2 || nChains > 64) return -1;
  st->n_chains = nChains;
  for (int i = 0; i < nChains * nChains; i++) st->grValues[i] = -2.0; /* :315-317 */
  return nChains == 2 ? 0 : 1;
}

static int psrf_converged(orc_psrf_state *st, const orc_treeset *ts, const int32_t *burnin, int end);

/* TreePSRF.java:107-197 converged2sided, literally, for two chains. */
static int converged2sided(orc_psrf_state *st, const orc_treeset *ts, const int32_t *burnin, int end) {
  double lower = 1.0 - st->b, upper = 1.0 + st->b; /* :89-91 */
  if (end % st->delta != 0) return st->prev;       /* :113-116 */
  int start = (int)(end * (1.0 - st->smoothing));  /* :117 */
  start = start - start % st->delta;               /* :118 */
  if (end / st->delta > st->cacheLimit) {          /* :120-124 */
    st->delta *= 2;
    return psrf_converged(st, ts, burnin, end);
  }
  int delta = st->delta;
  st->n_rows_last = (end - start) / delta;
  double psrf1mean = orc_psrf_rows_mean(ts, 0, 1, start, burnin, end, delta, NULL, NULL); /* :141-145 */
  st->grValues[0] = psrf1mean;                                                            /* :146 */
  double psrf2mean = 0;                                                                   /* :147 */
  if (lower < psrf1mean && psrf1mean < upper) {                                           /* :150 */
    if (st->start0 < 0) st->start0 = start;                                               /* :151-153 */
    psrf2mean = orc_psrf_rows_mean(ts, 1, 0, start, burnin, end, delta, NULL, NULL);      /* :154-157 */
    st->grValues[1] = psrf2mean;                                                          /* :158 */
    if (lower < psrf2mean && psrf2mean < upper) {                                         /* :161 */
      if (!st->checkESS) return 1;                                                        /* :162-164 */
      if ((end - st->start0) <= st->targetESS) return 0;                                  /* :168-171 */
      /* :176 `checkESS || pseudoESS(..)`: checkESS is true here, so pseudoESS is never evaluated */
      return 1;
    }
  }
  if (lower <= psrf1mean || psrf1mean >= upper || lower <= psrf2mean || psrf2mean >= upper) { /* :184-187 */
    st->start0 = -1;
  }
  return 0; /* :196 */
}

/* TreePSRF.java:200-261 converged1sided(side = 0).  pseudoESS (:244) needs BEAST's ESS and is
 * restated in orc_pseudo_ess below. */
double orc_pseudo_ess(const orc_treeset *ts, int32_t *indices, int32_t *have_indices, int32_t N, int32_t delta,
                      int32_t treeSet, int32_t cutStart, int32_t cutEnd);

static int converged1sided(orc_psrf_state *st, const orc_treeset *ts, int side, const int32_t *burnin, int end) {
  double lower = 1.0 - st->b, upper = 1.0 + st->b;
  if (end % st->delta != 0) return st->prev;      /* :206-209 */
  int start = (int)(end * (1.0 - st->smoothing)); /* :210 */
  start = start - start % st->delta;              /* :211 */
  if (end / st->delta > st->cacheLimit) {         /* :213-217 */
    st->delta *= 2;
    return psrf_converged(st, ts, burnin, end);
  }
  int delta = st->delta;
  st->n_rows_last = (end - start) / delta;
  double psrf1mean = orc_psrf_rows_mean(ts, side - 0, 1 - side, start, burnin, end, delta, NULL, NULL); /* :219-223 */
  st->grValues[0] = psrf1mean;                                                                          /* :224 */
  if (lower < psrf1mean && psrf1mean < upper) {                                                         /* :227 */
    if (!st->checkESS) return 1;                                                                        /* :228-230 */
    if (st->start0 < 0) st->start0 = start;                                                             /* :232-234 */
    if ((end - st->start0) <= st->targetESS) return 0;                                                  /* :238-240 */
    int cutStart = st->start0;
    cutStart = cutStart - cutStart % delta; /* :241-242 */
    int cutEnd = end;
    if (orc_pseudo_ess(ts, st->indices, &st->have_indices, st->sampleSize, delta, side - 0, cutStart, cutEnd) >=
        st->targetESS) { /* :244-246 */
      return 1;
    }
    return 0;
  }
  if (lower >= psrf1mean || psrf1mean >= upper) st->start0 = -1; /* :249-251 */
  return 0;
}

/* Generalisation to C > 2 chains (following the reference's own
 * every-pair use in tools/TreeConvergenceCheck.java:52-77): the 2-chain statistic is evaluated
 * for every ordered chain pair; converged iff every mean lies strictly inside (lower, upper). */
static int converged_multi(orc_psrf_state *st, const orc_treeset *ts, const int32_t *burnin, int end) {
  double lower = 1.0 - st->b, upper = 1.0 + st->b;
  int C = st->n_chains;
  if (end % st->delta != 0) return st->prev;
  int start = (int)(end * (1.0 - st->smoothing));
  start = start - start % st->delta;
  if (end / st->delta > st->cacheLimit) {
    st->delta *= 2;
    return psrf_converged(st, ts, burnin, end);
  }
  int delta = st->delta;
  st->n_rows_last = (end - start) / delta;
  int all_in = 1;
  for (int a = 0; a < C; a++)
    for (int b = 0; b < C; b++) {
      if (a == b) continue;
      double m = orc_psrf_rows_mean(ts, a, b, start, burnin, end, delta, NULL, NULL);
      st->grValues[a * C + b] = m;
      if (!(lower < m && m < upper)) all_in = 0;
    }
  if (all_in) {
    if (st->start0 < 0) st->start0 = start;
    if (!st->checkESS) return 1;
    if ((end - st->start0) <= st->targetESS) return 0;
    return 1;
  }
  st->start0 = -1;
  return 0;
}

/* TreePSRF.java:96-105 converged(). */
static int psrf_converged(orc_psrf_state *st, const orc_treeset *ts, const int32_t *burnin, int end) {
  if (st->n_chains != 2) {
    st->prev = converged_multi(st, ts, burnin, end);
    return st->prev;
  }
  if (st->twoSided) {
    st->prev = converged2sided(st, ts, burnin, end);
    return st->prev;
  } else {
    st->prev = converged1sided(st, ts, 0, burnin, end);
    return st->prev;
  }
}

int32_t orc_psrf_converged(orc_psrf_state *st, const orc_treeset *ts, const int32_t *burnin, int32_t end) {
  /* The Java indexes trees[chain].get(i) for i < end; an index past the list throws, is caught
   * by catch (Throwable) (:188-194), sets start0 = -1 and returns false. */
  for (int c = 0; c < st->n_chains; c++)
    if (end > ts->n_trees[c] || burnin[c] < 0) {
      st->threw = 1;
      st->start0 = -1;
      st->prev = 0;
      return 0;
    }
  return psrf_converged(st, ts, burnin, end);
}

/* Full S table for the generalised statistic: S[((a*n_rows)+r)*C + c] = sum over columns
 * i in chain c (i from start(burnin[c]) to end step delta) of distancePlusOne(a, x_r, c, i)^2
 * with x_r = start + r*delta.  This is varIn (c == a) / varBetween (c != a) of calcPSRF
 * (TreePSRF.java:275-286), accumulated in double exactly as the Java does. */
void orc_psrf_sums(const orc_treeset *ts, const int32_t *burnin, int32_t start, int32_t end, int32_t delta,
                   double *S) {
  int C = ts->n_chains;
  int n_rows = (end - start) / delta;
  if ((end - start) % delta != 0) n_rows++; /* rows x = start, start+delta, ... < end */
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
  for (int a = 0; a < C; a++)
    for (int r = 0; r < n_rows; r++) {
      int x = start + r * delta;
      for (int c = 0; c < C; c++) {
        double v = 0;
        int s = psrf_start(burnin[c], delta);
        for (int i = s; i < end; i += delta) {
          double d = distance_plus_one(ts, a, x, c, i);
          v += d * d;
        }
        S[((size_t)a * n_rows + r) * C + c] = v;
      }
    }
}

/* ------------------------------------------------------------------------------------------
 * TraceESS
 * ---------------------------------------------------------------------------------------- */

/* TraceESS.java:130-179 ACT(double[] trace, int start, int end), literally: every outer
 * iteration recomputes autoCorrelation[] for all reachable lags (:147-158). */
double orc_act_literal(const double *trace, int64_t start, int64_t end) {
  double sum = 0.0;
  double *squareLaggedSums = (double *)calloc(ORC_MAX_LAG, sizeof(double));
  double *autoCorrelation = (double *)calloc(ORC_MAX_LAG, sizeof(double));
  for (int64_t i = 0; i < (end - start); i++) {
    double traceI = trace[i + start];
    sum += traceI;
    const double mean = sum / (double)(i + 1);
    double sum1 = sum;
    double sum2 = sum;
    int64_t lim = (i + 1) < ORC_MAX_LAG ? (i + 1) : ORC_MAX_LAG;
    for (int64_t lagIndex = 0; lagIndex < lim; lagIndex++) {
      double traceIMinusLagIndex = trace[i + start - lagIndex];
      squareLaggedSums[lagIndex] = squareLaggedSums[lagIndex] + traceIMinusLagIndex * traceI;
      autoCorrelation[lagIndex] =
          squareLaggedSums[lagIndex] - (sum1 + sum2) * mean + mean * mean * (double)(i + 1 - lagIndex);
      autoCorrelation[lagIndex] /= (double)(i + start + 1 - lagIndex);
      sum1 -= traceIMinusLagIndex;
      sum2 -= trace[start + lagIndex];
    }
  }
  const int64_t maxLag = (end - start) < ORC_MAX_LAG ? (end - start) : ORC_MAX_LAG;
  double integralOfACFunctionTimes2 = 0.0;
  for (int64_t lagIndex = 0; lagIndex < maxLag; lagIndex++) {
    if (lagIndex == 0) {
      integralOfACFunctionTimes2 = autoCorrelation[0];
    } else if (lagIndex % 2 == 0) {
      if (autoCorrelation[lagIndex - 1] + autoCorrelation[lagIndex] > 0) {
        integralOfACFunctionTimes2 += 2.0 * (autoCorrelation[lagIndex - 1] + autoCorrelation[lagIndex]);
      } else {
        break;
      }
    }
  }
  double r = integralOfACFunctionTimes2 / autoCorrelation[0];
  free(squareLaggedSums);
  free(autoCorrelation);
  return r;
}

/* The same function with the dead work removed: autoCorrelation[] is overwritten on every
 * outer iteration, so only the pass i = n-1 reaches the integral.  Every value
 * that pass reads is accumulated in the order the literal loop accumulates it:
 *   sum                   += trace[i]                         (:137)   i ascending
 *   squareLaggedSums[lag] += trace[i-lag]*trace[i]            (:149)   i ascending from lag
 *   sum1, sum2            -= trace[n-1-lag], trace[lag]       (:156-157) lag ascending
 * ac_out (optional, [min(n, max_lag)]) receives autoCorrelation[]; lags_used (optional)
 * receives the number of lags the integral loop (:161-172) read before it stopped. */
double orc_act_final(const double *trace, int64_t start, int64_t end, int32_t max_lag, double *ac_out,
                     int32_t *lags_used) {
  int64_t n = end - start;
  const double *x = trace + start;
  int64_t L = n < max_lag ? n : max_lag;
  double *ac = (double *)calloc((size_t)(L > 0 ? L : 1), sizeof(double));
  double sum = 0.0;
  for (int64_t i = 0; i < n; i++) sum += x[i];
  const double mean = sum / (double)n;
  double sum1 = sum, sum2 = sum;
  for (int64_t lag = 0; lag < L; lag++) {
    double sls = 0.0;
    for (int64_t i = lag; i < n; i++) sls = sls + x[i - lag] * x[i];
    double a = sls - (sum1 + sum2) * mean + mean * mean * (double)(n - lag);
    a /= (double)(n + start - lag);
    ac[lag] = a;
    sum1 -= x[n - 1 - lag];
    sum2 -= x[lag];
  }
  double integral = 0.0;
  int32_t used = 0;
  for (int64_t lag = 0; lag < L; lag++) {
    if (lag == 0) {
      integral = ac[0];
      used = 1;
    } else if (lag % 2 == 0) {
      used = (int32_t)lag + 1;
      if (ac[lag - 1] + ac[lag] > 0)
        integral += 2.0 * (ac[lag - 1] + ac[lag]);
      else
        break;
    }
  }
  if (ac_out) memcpy(ac_out, ac, sizeof(double) * (size_t)L);
  if (lags_used) *lags_used = used;
  /* :178; for n = 0 the Java divides 0.0 by autoCorrelation[0] = 0.0, i.e. NaN */
  double r = L > 0 ? integral / ac[0] : NAN;
  free(ac);
  return r;
}

/* TraceESS.java:126-128 calcESS(double[] trace, int start, int end). */
double orc_calc_ess(const double *trace, int64_t start, int64_t end, int32_t literal) {
  double act = literal ? orc_act_literal(trace, start, end) : orc_act_final(trace, start, end, ORC_MAX_LAG, NULL, NULL);
  return (double)(end - start) / act;
}

/* Batched ACT/ESS over independent traces (the shape of BASELINE config 3), OpenMP over
 * traces.  traces is [n_traces][stride]. */
void orc_ess_batch(const double *traces, int32_t n_traces, int64_t n_samples, int64_t stride, int32_t literal,
                   double *act_out, double *ess_out) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int32_t t = 0; t < n_traces; t++) {
    const double *x = traces + (size_t)t * stride;
    double act = literal ? orc_act_literal(x, 0, n_samples) : orc_act_final(x, 0, n_samples, ORC_MAX_LAG, NULL, NULL);
    if (act_out) act_out[t] = act;
    if (ess_out) ess_out[t] = (double)n_samples / act;
  }
}

/* Java Math.min: NaN if either argument is NaN (TraceESS.java:63). */
static double java_min(double a, double b) {
  if (a != a) return a;
  if (b != b) return b;
  return a < b ? a : b;
}

/* TraceESS.java:38-69 converged(burnin, end, useMapping = true).
 * logs[chain] points at a column-major table [n_cols][log_stride] (logLines[chain][col].get(x),
 * TraceInfo.java:21).  map is traceInfo.getMap() (TraceInfo.java:52-68).  ess_out is
 * currentESSs (:52).  Returns minESS >= targetESS * nChains (:68). */
int32_t orc_trace_ess_converged(const double *const *logs, int64_t log_stride, int32_t nChains, const int32_t *burnin,
                                int32_t end, const int32_t *map, int32_t n_map, int32_t targetESS, int32_t literal,
                                double *ess_out, double *min_out) {
  int64_t total = 0;
  for (int i = 0; i < nChains; i++) total += end - burnin[i]; /* :40-43 */
  double *trace = (double *)malloc(sizeof(double) * (size_t)(total > 0 ? total : 1)); /* :47 */
  double minESS = INFINITY;                                                           /* :50 */
  for (int j = 0; j < n_map; j++) {                                                   /* :53 */
    int traceIndex = map[j];
    int64_t k = 0;
    for (int i = 0; i < nChains; i++)                  /* :56-60 */
      for (int x = burnin[i]; x < end; x++) trace[k++] = logs[i][(size_t)traceIndex * log_stride + x];
    double ess = orc_calc_ess(trace, 0, total, literal); /* :61 */
    if (ess_out) ess_out[j] = ess;                       /* :62 */
    minESS = java_min(ess, minESS);                      /* :63 */
  }
  free(trace);
  if (min_out) *min_out = minESS;
  return minESS >= (double)(targetESS * nChains); /* :68 */
}

/* ------------------------------------------------------------------------------------------
 * BurnInDetector
 * ---------------------------------------------------------------------------------------- */

/* BurnInDetector.java:51-78 burnIn(List<Double> trace, int end); WINDOW_SIZE = 10 (:49). */
int32_t orc_burnin_trace(const double *trace, int32_t end) {
  const int WINDOW_SIZE = 10;
  double m2 = 0, sq2 = 0;
  int lb = 3 * end / 4;
  for (int i = lb; i < end; i++) m2 += trace[i];
  m2 = m2 / (double)(end - lb);
  for (int i = lb; i < end; i++) {
    double d = trace[i];
    sq2 += (d - m2) * (d - m2);
  }
  double stdev2 = sqrt(sq2 / (double)(end - lb - 1));
  for (int i = 0; i + WINDOW_SIZE < lb; i++) {
    double sum = 0;
    for (int j = 0; j < WINDOW_SIZE; j++) sum += trace[i + j];
    sum /= WINDOW_SIZE;
    if (m2 - stdev2 < sum && sum < m2 + stdev2) return i + WINDOW_SIZE;
  }
  return lb;
}

/* BurnInDetector.java:21-47 burnIn(end).  strategy 0 = Automatic (max over mapped columns of
 * the per-trace burn-in), 1 = Running (ceil(percent/100 * end) for every chain, :30-34). */
void orc_burnin(const double *const *logs, int64_t log_stride, int32_t nChains, const int32_t *map, int32_t n_map,
                int32_t strategy, int32_t percent, int32_t end, int32_t *burnin_out) {
  if (strategy == 0) {
    for (int i = 0; i < nChains; i++) {
      int maxBurnin = 0;
      for (int j = 0; j < n_map; j++) {
        int b = orc_burnin_trace(logs[i] + (size_t)map[j] * log_stride, end);
        maxBurnin = b > maxBurnin ? b : maxBurnin;
      }
      burnin_out[i] = maxBurnin;
    }
  } else {
    double burnPercent = (double)percent / 100;
    for (int i = 0; i < nChains; i++) burnin_out[i] = (int)ceil(burnPercent * end);
  }
}

/* ------------------------------------------------------------------------------------------
 * TreeESS.pseudoESS (TreeESS.java:117-166).  ESS.calcESS(Double[], 1) belongs to BEAST.base
 * (>= 2.7.5, un-vendored); TraceESS.ACT is the reference's own copy of that
 * routine (TraceESS.java:130-179), so it stands in for it here with sampleInterval = 1.
 * ---------------------------------------------------------------------------------------- */
double orc_pseudo_ess(const orc_treeset *ts, int32_t *indices, int32_t *have_indices, int32_t N, int32_t delta,
                      int32_t treeSet, int32_t cutStart, int32_t cutEnd) {
  if (!*have_indices) { /* :120-126 */
    for (int i = 0; i < N; i++) {
      int offset = (cutEnd - cutStart) * i / N;
      offset = offset - offset % delta;
      indices[i] = cutStart + offset;
    }
    *have_indices = 1;
  } else { /* :127-132 */
    while (indices[0] < cutStart) {
      memmove(indices, indices + 1, sizeof(int32_t) * (size_t)(N - 1));
      indices[N - 1] = cutEnd - cutEnd % delta;
    }
  }
  int len = (cutEnd - cutStart) / delta; /* :137 */
  double *trace = (double *)malloc(sizeof(double) * (size_t)(len > 0 ? len : 1));
  double sum = 0;
  for (int j = 0; j < N; j++) { /* :139-156, one reference tree at a time */
    int k = 0;
    for (int i = cutStart; i < cutEnd; i += delta) {
      if (k >= len) break; /* the Java would index past trace[j] and throw */
      if (i != indices[j])
        trace[k++] = (double)distance_plus_one(ts, treeSet, i, treeSet, indices[j]);
      else
        trace[k++] = 0.0;
    }
    double ess = (double)len / orc_act_final(trace, 0, len, ORC_MAX_LAG, NULL, NULL);
    sum += ess; /* :158-161 */
  }
  free(trace);
  return sum / N; /* :162 */
}

int32_t orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Launchers such as torchrun export OMP_NUM_THREADS=1 to their children; the timing harness asks for
 * the host's cores explicitly. */
void orc_set_num_threads(int32_t n) {
#ifdef _OPENMP
  if (n >= 1) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------
 * GelmanRubin (GelmanRubin.java) and CladeDifference (CladeDifference.java) criteria
 * ---------------------------------------------------------------------------------------- */

/* GelmanRubin.java:55-96 calcGRStat(sampleCount, trace1, trace2), literally. */
double orc_gr_stat(int32_t sampleCount, const double *trace1, const double *trace2) {
  double mean1 = 0, mean2 = 0, sumsq1 = 0, sumsq2 = 0;
  for (int i = 0; i < sampleCount; i++) {
    double d = trace1[i];
    mean1 += d;
    sumsq1 += d * d;
  }
  mean1 /= sampleCount;
  for (int i = 0; i < sampleCount; i++) {
    double d = trace2[i];
    mean2 += d;
    sumsq2 += d * d;
  }
  mean2 /= sampleCount;
  double var1 = (sumsq1 - mean1 * mean1 * sampleCount) / (sampleCount - 1);
  double var2 = (sumsq2 - mean2 * mean2 * sampleCount) / (sampleCount - 1);
  double fW = (var1 + var2) / 2;
  if (fW == 0) return 1;
  double totalMean = (mean1 + mean2) / 2;
  double totalSq = mean1 * mean1 + mean2 * mean2;
  double fB = (totalSq - totalMean * totalMean * 2);
  double varR = ((sampleCount - 1.0) / sampleCount) + (fB / fW) * (1.0 / sampleCount);
  return sqrt(varR);
}

/* GelmanRubin.java:32-51 converged(burnin, available): every chain pair i<j, every item k (all columns, no
 * burn-in), false at the first statistic above the threshold.  gr_out (optional, [pairs][nItems] in loop
 * order) receives every statistic regardless of the early return. */
int32_t orc_gelman_rubin_converged(const double *const *logs, int64_t log_stride, int32_t nChains, int32_t nItems,
                                   int32_t available, double threshold, double *gr_out) {
  int result = 1, decided = 0, p = 0;
  for (int i = 0; i < nChains; i++)
    for (int j = i + 1; j < nChains; j++)
      for (int k = 0; k < nItems; k++) {
        double g = orc_gr_stat(available, logs[i] + (size_t)k * log_stride, logs[j] + (size_t)k * log_stride);
        if (gr_out) gr_out[p] = g;
        p++;
        if (!decided && g > threshold) {
          result = 0;
          decided = 1;
        }
      }
  return result;
}

/* CladeDifference.java:85-103 calcMaxCladeDifference(iThread1, iThread2) on the clade counts of the first
 * `available` trees of each chain (process() has been called for trees [0, available), :48-53).
 * m_nClades is the LENGTH OF THE ROOT CLADE = number of taxa (:149), and nTotal/m_nClades is an int division. */
double orc_max_clade_difference(const orc_treeset *ts, int32_t iThread1, int32_t iThread2, int32_t available) {
  if (iThread1 == iThread2) return 0;
  int64_t D = ts->dict.n;
  int32_t *c1 = (int32_t *)calloc((size_t)(D > 0 ? D : 1), sizeof(int32_t));
  int32_t *c2 = (int32_t *)calloc((size_t)(D > 0 ? D : 1), sizeof(int32_t));
  orc_clade_counts(ts, iThread1, 0, available, c1);
  orc_clade_counts(ts, iThread2, 0, available, c2);
  int nTotal = 0, nMax = 0;
  for (int64_t id = 0; id < D; id++) {
    if (c1[id] == 0) continue; /* map1.keySet() */
    int i1 = c1[id], i2 = c2[id];
    nTotal += i1;
    int a = i1 - i2;
    if (a < 0) a = -a;
    if (a > nMax) nMax = a;
  }
  free(c1);
  free(c2);
  int m_nClades = available > 0 ? ts->n_taxa : 1; /* :25 initial value 1 until a tree is processed */
  return nMax / (double)(nTotal / m_nClades);
}

/* Java Math.max(double, double): NaN if either is NaN. */
static double java_max(double a, double b) {
  if (a != a) return a;
  if (b != b) return b;
  return a > b ? a : b;
}

/* CladeDifference.java:41-65 converged(burnin, available). */
int32_t orc_clade_difference_converged(const orc_treeset *ts, int32_t nChains, int32_t available, double threshold,
                                       double *max_out) {
  double fMaxCladeProbDiff = 0;
  for (int i = 0; i < nChains; i++)
    for (int k = 0; k < nChains; k++)
      fMaxCladeProbDiff = java_max(fMaxCladeProbDiff, orc_max_clade_difference(ts, i, k, available));
  if (max_out) *max_out = fMaxCladeProbDiff;
  return fMaxCladeProbDiff < threshold;
}
