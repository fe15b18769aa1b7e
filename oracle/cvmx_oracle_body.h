/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Included twice by cvmx_oracle.c with
 * REAL = double / float and SUF = f64 / f32.  Plain-C, order-explicit restatement of
 * the reference path; every floating-point operation is one rounded operation in REAL
 * (the file is compiled with -ffp-contract=off, no -ffast-math).
 *
 * Reference (file:line under /root/reference):
 *   pairwise weight sums ..... numpy pairwise summation behind cvmatrix/cvmatrix.py:617, 1225
 *   sequential column sums ... numpy axis-0 reduction behind cvmatrix/cvmatrix.py:709-737, 1231-1241
 *   fit ...................... cvmatrix/cvmatrix.py:1193-1243
 *   fold statistics .......... cvmatrix/cvmatrix.py:589-752, 1012-1129
 *   fold kernel matrices ..... cvmatrix/cvmatrix.py:898-1010
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* numpy pairwise summation of f(i), i in [0, n) ------------------------------------ */
typedef REAL (*FN(elem_fn))(const void* ctx, int64_t i);

static REAL FN(pairwise)(FN(elem_fn) f, const void* ctx, int64_t off, int64_t n) {
  if (n < 8) {
    REAL r = (REAL)0;
    for (int64_t i = 0; i < n; ++i) r = r + f(ctx, off + i);
    return r;
  }
  if (n <= 128) {
    REAL r[8];
    for (int j = 0; j < 8; ++j) r[j] = f(ctx, off + j);
    int64_t stop = n - (n % 8);
    for (int64_t i = 8; i < stop; i += 8)
      for (int j = 0; j < 8; ++j) r[j] = r[j] + f(ctx, off + i + j);
    REAL res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (int64_t i = stop; i < n; ++i) res = res + f(ctx, off + i);
    return res;
  }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  REAL a = FN(pairwise)(f, ctx, off, n2);
  REAL b = FN(pairwise)(f, ctx, off + n2, n - n2);
  return a + b;
}

typedef struct {
  const REAL* Z;      /* X or Y */
  const REAL* w;      /* NULL = unweighted */
  const int64_t* idx; /* NULL = identity */
  int64_t ld;
  int64_t col;
  int kind; /* 0: w   1: w*z   2: (w*z)*z */
} FN(ctx_t);

static REAL FN(elem)(const void* vctx, int64_t i) {
  const FN(ctx_t)* c = (const FN(ctx_t)*)vctx;
  int64_t r = c->idx ? c->idx[i] : i;
  if (c->kind == 0) return c->w[r];
  REAL z = c->Z[r * c->ld + c->col];
  REAL wz = c->w ? (REAL)(z * c->w[r]) : z;
  if (c->kind == 1) return wz;
  return (REAL)(wz * z);
}

/* column sums of (w*Z)[idx] (kind 1) or ((w*Z)*Z)[idx] (kind 2), numpy order ----------- */
static void FN(colsums)(const REAL* Z, const REAL* w, const int64_t* idx, int64_t n, int64_t C,
                        int kind, REAL* out) {
  if (C == 1) {
    FN(ctx_t) c = {Z, w, idx, 1, 0, kind};
    out[0] = (REAL)0 + FN(pairwise)(FN(elem), &c, 0, n);
    return;
  }
  for (int64_t j = 0; j < C; ++j) out[j] = (REAL)0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t r = idx ? idx[i] : i;
    const REAL* z = Z + r * C;
    if (w) {
      REAL wr = w[r];
      if (kind == 1)
        for (int64_t j = 0; j < C; ++j) out[j] = out[j] + (REAL)(z[j] * wr);
      else
        for (int64_t j = 0; j < C; ++j) out[j] = out[j] + (REAL)((REAL)(z[j] * wr) * z[j]);
    } else {
      if (kind == 1)
        for (int64_t j = 0; j < C; ++j) out[j] = out[j] + z[j];
      else
        for (int64_t j = 0; j < C; ++j) out[j] = out[j] + (REAL)(z[j] * z[j]);
    }
  }
}

/* G[i][j] = sum_r (w_r x_ri) b_rj over the listed rows, accumulated in row order -------- */
static void FN(gram)(const REAL* X, const REAL* B, const REAL* w, const int64_t* idx, int64_t n,
                     int64_t K, int64_t C, REAL* G) {
  for (int64_t e = 0; e < K * C; ++e) G[e] = (REAL)0;
  for (int64_t t = 0; t < n; ++t) {
    int64_t r = idx ? idx[t] : t;
    const REAL* x = X + r * K;
    const REAL* b = B + r * C;
    for (int64_t i = 0; i < K; ++i) {
      REAL a = w ? (REAL)(x[i] * w[r]) : x[i];
      REAL* g = G + i * C;
      for (int64_t j = 0; j < C; ++j) g[j] = g[j] + (REAL)(a * b[j]);
    }
  }
}

/*
 * fit: totals and moment sums.  flags = cX | cY<<1 | sX<<2 | sY<<3.  Outputs that the
 * reference would leave as None are not written.  Returns 1 for a negative weight.
 */
int FN(orc_fit)(const REAL* X, const REAL* Y, const REAL* w, int64_t N, int64_t K, int64_t M,
                uint32_t flags, REAL* XTX, REAL* XTY, REAL* sum_X, REAL* sum_Y, REAL* sum_sq_X,
                REAL* sum_sq_Y, REAL* sum_w, int64_t* nnz_w) {
  int cX = flags & 1, cY = (flags >> 1) & 1, sX = (flags >> 2) & 1, sY = (flags >> 3) & 1;
  if (w)
    for (int64_t i = 0; i < N; ++i)
      if (w[i] < 0) return 1;
  FN(gram)(X, X, w, NULL, N, K, K, XTX);
  if (Y) FN(gram)(X, Y, w, NULL, N, K, M, XTY);
  if (cX || cY || sX || sY) {
    if (w) {
      FN(ctx_t) c = {NULL, w, NULL, 1, 0, 0};
      *sum_w = (REAL)0 + FN(pairwise)(FN(elem), &c, 0, N);
      int64_t nz = 0;
      for (int64_t i = 0; i < N; ++i) nz += (w[i] != 0);
      *nnz_w = nz;
    } else {
      *sum_w = (REAL)N;
      *nnz_w = N;
    }
  }
  if (cX || cY || sX) FN(colsums)(X, w, NULL, N, K, 1, sum_X);
  if (Y && (cX || cY || sY)) FN(colsums)(Y, w, NULL, N, M, 1, sum_Y);
  if (sX) FN(colsums)(X, w, NULL, N, K, 2, sum_sq_X);
  if (Y && sY) FN(colsums)(Y, w, NULL, N, M, 2, sum_sq_Y);
  return 0;
}

static void FN(std_row)(const REAL* q_total, const REAL* q_val, const REAL* mean, const REAL* s_train,
                        REAL sw, REAL div, REAL resolution, int64_t C, REAL* out) {
  for (int64_t j = 0; j < C; ++j) {
    REAL q = q_total[j] - q_val[j];
    REAL t1 = (REAL)((REAL)((REAL)-2 * mean[j]) * s_train[j]);
    REAL t2 = (REAL)(sw * (REAL)(mean[j] * mean[j]));
    REAL var = (REAL)((REAL)((REAL)(t1 + t2) + q) / div);
    if (!(var >= (REAL)0) && var == var) var = (REAL)0; /* np.maximum keeps NaN */
    REAL sd = (REAL)SQRT(var);
    out[j] = (sd <= resolution) ? (REAL)1 : sd;
  }
}

/*
 * One fold.  want = XTX | XTY<<1.  stats = 4 rows [X_mean(K) | X_std(K) | Y_mean(M) | Y_std(M)]
 * packed back to back; the full set the flags allow is written (callers select which to
 * expose).  scal = {sum_w_train, nnz_train}.  status: 0 ok, 1 nnz_train == 0 (weighted only),
 * 2 nnz_train <= ddof (only reported when a std is computed).  Needs the totals from orc_fit.
 */
int FN(orc_fold)(const REAL* X, const REAL* Y, const REAL* w, int64_t N, int64_t K, int64_t M,
                 uint32_t flags, int64_t ddof, REAL resolution, const REAL* T_XX, const REAL* T_XY,
                 const REAL* sum_X, const REAL* sum_Y, const REAL* sum_sq_X, const REAL* sum_sq_Y,
                 REAL sum_w, int64_t nnz_w, const int64_t* val_in, int64_t n_val, uint32_t want,
                 REAL* out_XTX, REAL* out_XTY, REAL* stats, REAL* scal, int32_t* status) {
  int cX = flags & 1, cY = (flags >> 1) & 1, sX = (flags >> 2) & 1, sY = (flags >> 3) & 1;
  int wXX = want & 1, wXY = (want >> 1) & 1;
  int hasY = Y != NULL;
  *status = 0;
  int64_t* val = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_val > 0 ? n_val : 1));
  for (int64_t i = 0; i < n_val; ++i) {
    int64_t v = val_in[i];
    if (v < 0) v += N; /* numpy wrap-around */
    if (v < 0 || v >= N) { free(val); return 3; }
    val[i] = v;
  }
  REAL* buf = (REAL*)calloc((size_t)(4 * (K + (hasY ? M : 0)) + 8), sizeof(REAL));
  REAL *sXv = buf, *qXv = buf + K, *sXt = buf + 2 * K, *sYv = buf + 3 * K;
  REAL *qYv = sYv + M, *sYt = sYv + 2 * M;
  REAL *Xm = stats, *Xs = stats + K, *Ym = stats + 2 * K, *Ys = stats + 2 * K + M;

  int any = cX || cY || sX || sY;
  REAL sw = (REAL)0, nz = (REAL)0, div = (REAL)0;
  int needXm = cX || cY || sX, needYm = hasY && (cX || cY || sY);
  if (any) {
    if (w) {
      FN(ctx_t) c = {NULL, w, val, 1, 0, 0};
      REAL swv = (REAL)0 + FN(pairwise)(FN(elem), &c, 0, n_val);
      int64_t nzv = 0;
      for (int64_t i = 0; i < n_val; ++i) nzv += (w[val[i]] != 0);
      sw = (REAL)(sum_w - swv);
      nz = (REAL)(nnz_w - nzv);
      if (nz == (REAL)0) *status = 1;
    } else {
      sw = nz = (REAL)(N - n_val);
    }
    if (needXm) {
      FN(colsums)(X, w, val, n_val, K, 1, sXv);
      for (int64_t j = 0; j < K; ++j) { sXt[j] = sum_X[j] - sXv[j]; Xm[j] = sXt[j] / sw; }
    }
    if (needYm) {
      FN(colsums)(Y, w, val, n_val, M, 1, sYv);
      for (int64_t j = 0; j < M; ++j) { sYt[j] = sum_Y[j] - sYv[j]; Ym[j] = sYt[j] / sw; }
    }
    if (sX || (hasY && sY)) {
      if (*status == 0 && nz <= (REAL)ddof) *status = 2;
      div = (REAL)((REAL)((REAL)(nz - (REAL)ddof) * sw) / nz);
    }
    if (sX) {
      FN(colsums)(X, w, val, n_val, K, 2, qXv);
      FN(std_row)(sum_sq_X, qXv, Xm, sXt, sw, div, resolution, K, Xs);
    }
    if (hasY && sY) {
      FN(colsums)(Y, w, val, n_val, M, 2, qYv);
      FN(std_row)(sum_sq_Y, qYv, Ym, sYt, sw, div, resolution, M, Ys);
    }
  }
  scal[0] = sw;
  scal[1] = nz;

  if (wXX) {
    FN(gram)(X, X, w, val, n_val, K, K, out_XTX);
    for (int64_t i = 0; i < K; ++i)
      for (int64_t j = 0; j < K; ++j) {
        REAL a = T_XX[i * K + j] - out_XTX[i * K + j];
        if (cX) a = a - (REAL)(sw * (REAL)(Xm[i] * Xm[j]));
        if (sX) a = a / (REAL)(Xs[i] * Xs[j]);
        out_XTX[i * K + j] = a;
      }
  }
  if (wXY && hasY) {
    FN(gram)(X, Y, w, val, n_val, K, M, out_XTY);
    for (int64_t i = 0; i < K; ++i)
      for (int64_t j = 0; j < M; ++j) {
        REAL a = T_XY[i * M + j] - out_XTY[i * M + j];
        if (cX || cY) a = a - (REAL)(sw * (REAL)(Xm[i] * Ym[j]));
        if (sX && sY) a = a / (REAL)(Xs[i] * Ys[j]);
        else if (sX) a = a / Xs[i];
        else if (sY) a = a / Ys[j];
        out_XTY[i * M + j] = a;
      }
  }
  free(buf);
  free(val);
  return 0;
}

/* exported helpers so tests can pin the two summation orders on their own */
REAL FN(orc_pairwise_sum)(const REAL* a, int64_t n) {
  FN(ctx_t) c = {NULL, a, NULL, 1, 0, 0};
  return (REAL)0 + FN(pairwise)(FN(elem), &c, 0, n);
}
void FN(orc_colsum)(const REAL* A, int64_t n, int64_t C, REAL* out) {
  FN(colsums)(A, NULL, NULL, n, C, 1, out);
}

#undef FN
#undef CAT
#undef CAT_
