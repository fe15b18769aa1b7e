// Statistics kernels: numpy-order weight sums and column moments, per-fold means / stds.
//
// Parity with the reference hinges on the ORDER of these reductions:
//   * sum of weights (np.sum of an (n,1) array)        -> numpy pairwise summation
//   * column sums over rows (np.sum(A, axis=0), C >= 2) -> strictly sequential in row order
//   * C == 1                                            -> pairwise again
// and on every elementwise op being individually rounded (Rn<T>, no FMA contraction).
// Reference: cvmatrix/cvmatrix.py:589-630 (weight mass), :632-752 (fold statistics),
// :1045-1129 (divisor, std), :1219-1243 (fit moments).
#pragma once
#include "common.cuh"

namespace cvmx {

constexpr int PW_LEVELS_BIG = 10;   // 1024 sub-trees of the pairwise recursion per block (long folds, fit)
constexpr int PW_LEVELS_SMALL = 5;  // one warp per fold (folds of <= 1024 rows)

// Functor over "element i of the reduced vector" for the three quantities numpy sums pairwise.
template <typename T>
struct PwSrc {
  const T* Z;          // N x ld, [X | Y | pad]
  const T* w;          // N (all ones when the model is unweighted)
  const int64_t* idx;  // row list (nullptr: identity)
  int64_t ld;
  int64_t col;
  int kind;  // 0: w   1: rn(w*z)   2: rn(rn(w*z)*z)
  __device__ __forceinline__ T operator()(int64_t i) const {
    const int64_t r = idx ? idx[i] : i;
    const T wr = w[r];
    if (kind == 0) return wr;
    const T z = Z[r * ld + col];
    const T wz = Rn<T>::mul(z, wr);
    return kind == 1 ? wz : Rn<T>::mul(wz, z);
  }
};

template <typename T, typename F>
__device__ T pw_leaf(const F& f, int64_t off, int64_t n) {
  if (n < 8) {
    T r = T(0);
    for (int64_t i = 0; i < n; ++i) r = Rn<T>::add(r, f(off + i));
    return r;
  }
  T r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = f(off + j);
  const int64_t stop = n - (n % 8);
  for (int64_t i = 8; i < stop; i += 8) {
    T v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = f(off + i + j);
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = Rn<T>::add(r[j], v[j]);
  }
  T res = Rn<T>::add(Rn<T>::add(Rn<T>::add(r[0], r[1]), Rn<T>::add(r[2], r[3])),
                     Rn<T>::add(Rn<T>::add(r[4], r[5]), Rn<T>::add(r[6], r[7])));
  for (int64_t i = stop; i < n; ++i) res = Rn<T>::add(res, f(off + i));
  return res;
}

// numpy's recursion (split at n/2 rounded down to a multiple of 8 while n > 128), walked with an
// explicit stack so no device call stack is needed.
template <typename T, typename F>
__device__ T pw_serial(const F& f, int64_t off0, int64_t n0) {
  constexpr int MAXD = 48;
  int64_t offs[MAXD], lens[MAXD];
  T accs[MAXD];
  signed char st[MAXD];
  int sp = 0;
  offs[0] = off0; lens[0] = n0; st[0] = 0; sp = 1;
  T ret = T(0);
  while (sp > 0) {
    const int top = sp - 1;
    const int64_t n = lens[top];
    if (st[top] == 0) {
      if (n <= 128) { ret = pw_leaf<T>(f, offs[top], n); --sp; continue; }
      int64_t n2 = n / 2; n2 -= n2 % 8;
      st[top] = 1;
      offs[sp] = offs[top]; lens[sp] = n2; st[sp] = 0; ++sp;
    } else if (st[top] == 1) {
      accs[top] = ret;
      int64_t n2 = n / 2; n2 -= n2 % 8;
      st[top] = 2;
      offs[sp] = offs[top] + n2; lens[sp] = n - n2; st[sp] = 0; ++sp;
    } else {
      ret = Rn<T>::add(accs[top], ret);
      --sp;
    }
  }
  return ret;
}

// Whole-block evaluation of the same tree: thread j walks PW_LEVELS levels down along the bits of j,
// sums its sub-tree serially, then the sub-tree values are combined bottom-up in exactly the order the
// recursion would (left + right).  Result is valid in thread 0 (and in s_val[0]).
template <typename T, int LEVELS, typename F>
__device__ T block_pairwise(const F& f, int64_t n, T* s_val, unsigned char* s_valid) {
  constexpr int PW_LEVELS = LEVELS, PW_THREADS = 1 << LEVELS;
  const int tid = threadIdx.x;
  int64_t off = 0, len = n;
  bool owner = true;
  for (int level = 0; level < PW_LEVELS; ++level) {
    if (len <= 128) {
      owner = (tid & ((1 << (PW_LEVELS - level)) - 1)) == 0;
      break;
    }
    int64_t n2 = len / 2; n2 -= n2 % 8;
    if ((tid >> (PW_LEVELS - 1 - level)) & 1) { off += n2; len -= n2; } else { len = n2; }
  }
  __syncthreads();  // s_val may be reused between calls
  s_val[tid] = owner ? pw_serial<T>(f, off, len) : T(0);
  s_valid[tid] = owner ? 1 : 0;
  __syncthreads();
  for (int s = 1; s < PW_THREADS; s <<= 1) {
    if ((tid & (2 * s - 1)) == 0 && s_valid[tid + s]) s_val[tid] = Rn<T>::add(s_val[tid], s_val[tid + s]);
    __syncthreads();
  }
  return Rn<T>::add(T(0), s_val[0]);  // numpy seeds add-reductions with +0
}

struct FitScalars {  // device-resident fit totals
  double sum_w;      // model-dtype value widened to double
  int64_t nnz_w;
  int32_t neg_weight;  // any(w < 0)
  int32_t pad;
};

// One block per fold (or one block for the whole data set when fit_mode): pairwise weight sum,
// non-zero count, the derived training scalars, and - for K == 1 / M == 1 - the pairwise column sums.
//   pw_cols[f][4] = { sum wx, sum wx*x, sum wy, sum wy*y } over the fold rows (only the C == 1 ones are used)
template <typename T, int LEVELS>
__global__ void __launch_bounds__(1 << LEVELS)
k_weight_mass(const T* __restrict__ Z, const T* __restrict__ w, int64_t ld, int64_t N, int64_t K, int64_t M,
              int weighted, const int64_t* __restrict__ offsets, const int64_t* __restrict__ indices,
              int64_t fold0, int fit_mode, int64_t ddof, FitScalars* __restrict__ fit, FoldScalars* __restrict__ fs,
              T* __restrict__ pw_cols) {
  constexpr int PW_THREADS = 1 << LEVELS;
  __shared__ T s_val[PW_THREADS];
  __shared__ unsigned char s_valid[PW_THREADS];
  __shared__ long long s_cnt[2];
  const int tid = threadIdx.x;
  const int64_t f = blockIdx.x;
  const int64_t beg = fit_mode ? 0 : offsets[fold0 + f];
  const int64_t n = fit_mode ? N : offsets[fold0 + f + 1] - beg;
  const int64_t* idx = fit_mode ? nullptr : indices + beg;

  if (tid < 2) s_cnt[tid] = 0;
  __syncthreads();
  T swv = T(0);
  if (weighted) {
    long long nz = 0, neg = 0;
    // 8 independent (index -> weight) load chains per trip: the loop is latency-, not bandwidth-bound
    for (int64_t i0 = tid; i0 < n; i0 += 8 * PW_THREADS) {
      int64_t r[8];
      T v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t i = i0 + (int64_t)u * PW_THREADS;
        r[u] = i < n ? (idx ? idx[i] : i) : -1;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = r[u] >= 0 ? w[r[u]] : T(0);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        nz += (v[u] != T(0));
        neg += (v[u] < T(0));
      }
    }
    if (nz) atomicAdd((unsigned long long*)&s_cnt[0], (unsigned long long)nz);
    if (neg) atomicAdd((unsigned long long*)&s_cnt[1], (unsigned long long)neg);
    PwSrc<T> src{Z, w, idx, ld, 0, 0};
    swv = block_pairwise<T, LEVELS>(src, n, s_val, s_valid);
  }
  T cols[4] = {T(0), T(0), T(0), T(0)};
  if (K == 1) {
    PwSrc<T> a{Z, w, idx, ld, 0, 1}, b{Z, w, idx, ld, 0, 2};
    cols[0] = block_pairwise<T, LEVELS>(a, n, s_val, s_valid);
    cols[1] = block_pairwise<T, LEVELS>(b, n, s_val, s_valid);
  }
  if (M == 1) {
    PwSrc<T> a{Z, w, idx, ld, K, 1}, b{Z, w, idx, ld, K, 2};
    cols[2] = block_pairwise<T, LEVELS>(a, n, s_val, s_valid);
    cols[3] = block_pairwise<T, LEVELS>(b, n, s_val, s_valid);
  }
  __syncthreads();
  if (tid != 0) return;
  if (pw_cols)
    for (int j = 0; j < 4; ++j) pw_cols[f * 4 + j] = cols[j];
  if (fit_mode) {
    fit->sum_w = weighted ? (double)swv : (double)N;
    fit->nnz_w = weighted ? (int64_t)s_cnt[0] : N;
    fit->neg_weight = s_cnt[1] != 0;
    fit->pad = 0;
    return;
  }
  T sw, nz;
  int32_t status = 0;
  if (weighted) {
    sw = Rn<T>::sub((T)fit->sum_w, swv);
    nz = (T)(fit->nnz_w - (int64_t)s_cnt[0]);
    if (nz == T(0)) status |= 1;
  } else {
    sw = nz = (T)(N - n);
  }
  if (nz <= (T)ddof) status |= 2;
  const T dv = Rn<T>::div(Rn<T>::mul(Rn<T>::sub(nz, (T)ddof), sw), nz);
  FoldScalars o;
  o.sw = (double)sw; o.nz = (double)nz; o.div = (double)dv; o.status = status; o.pad = 0;
  fs[f] = o;
}

// ---- per-column finalisation shared by both moment kernels ----------------------------------------
template <typename T>
struct MomentParams {
  const T* Z; const T* w; int64_t ld; int64_t K; int64_t M;
  const int64_t* offsets; const int64_t* indices; int64_t fold0;  // fit mode: offsets == nullptr, rows [0, N)
  int64_t N;
  uint32_t flags;
  T resolution;
  T* sum_z; T* sumsq_z;        // [ld] totals (written in fit mode, read otherwise)
  const FoldScalars* fs;       // per fold (fold mode)
  const T* pw_cols;            // per fold pairwise column sums for K == 1 / M == 1
  T* stats;                    // [P][2][ld]: mean, std
  int grp0 = 0, grp_stride = 1; // column-group sharding (multi-GPU): block b handles group grp0 + b * grp_stride
  T* raw = nullptr;             // [P][2][ld]: if set, store the raw fold sums (s, q) here and leave mean / std to
                                // k_finalize_stats - the chains then do not depend on the weight-mass kernel
  int stages = 4;               // ring depth of k_moments_pipe (runtime: few long chains want more bytes in flight)
  int64_t row0 = 0;             // fit mode: rows [row0, row0 + N) ...
  int accumulate = 0;           // ... continuing the chains stored in sum_z / sumsq_z (chunk-pipelined fit)
  const int64_t* ranges = nullptr;   // fold mode, optional: [fold][2] = begin, end positions in `indices` (a slice of the
                                     // fold's rows: chunk-pipelined fit + folds), instead of offsets[fold0 + fold]
  const int* scan_ok = nullptr; // [folds][scan_groups]: groups already finished by the binade scan (kernels_scan.cuh);
  int scan_groups = 0;          // k_moments_pipe skips them
};

template <typename T>
__device__ __forceinline__ void finalize_column(const MomentParams<T>& p, int64_t f, int64_t c, T s, T q) {
  const int64_t C = p.K + p.M;
  if (c >= C) return;
  const bool isX = c < p.K;
  if ((p.offsets || p.ranges) && p.raw) {  // deferred: k_finalize_stats turns the sums into mean / std once the fold scalars exist
    T* o = p.raw + (size_t)f * 2 * p.ld;
    o[c] = s;
    o[p.ld + c] = q;
    return;
  }
  if (p.pw_cols) {
    if (isX && p.K == 1) { s = p.pw_cols[f * 4 + 0]; q = p.pw_cols[f * 4 + 1]; }
    if (!isX && p.M == 1) { s = p.pw_cols[f * 4 + 2]; q = p.pw_cols[f * 4 + 3]; }
  }
  if (!p.offsets && !p.ranges) {  // fit mode
    p.sum_z[c] = s;
    p.sumsq_z[c] = q;
    return;
  }
  const bool cX = p.flags & 1, cY = p.flags & 2, sX = p.flags & 4, sY = p.flags & 8;
  const bool need_mean = isX ? (cX || cY || sX) : (cX || cY || sY);
  const bool need_std = isX ? sX : sY;
  T mean = T(0), sd = T(0);
  if (need_mean) {
    const T sw = (T)p.fs[f].sw;
    const T s_train = Rn<T>::sub(p.sum_z[c], s);
    mean = Rn<T>::div(s_train, sw);
    if (need_std) {
      const T dv = (T)p.fs[f].div;
      const T q_train = Rn<T>::sub(p.sumsq_z[c], q);
      const T t1 = Rn<T>::mul(Rn<T>::mul(T(-2), mean), s_train);
      const T t2 = Rn<T>::mul(sw, Rn<T>::mul(mean, mean));
      T var = Rn<T>::div(Rn<T>::add(Rn<T>::add(t1, t2), q_train), dv);
      if (!(var >= T(0)) && var == var) var = T(0);  // np.maximum(var, 0) keeps NaN
      sd = Rn<T>::sqrt(var);
      if (sd <= p.resolution) sd = T(1);
    }
  }
  T* o = p.stats + (size_t)f * 2 * p.ld;
  o[c] = mean;
  o[p.ld + c] = sd;
}

// ---- small folds: one thread per (fold, column), rows read straight from global ---------------------
template <typename T>
__global__ void __launch_bounds__(128) k_moments_direct(MomentParams<T> p) {
  const int64_t f = blockIdx.y;
  const int64_t c = (int64_t)(p.grp0 + blockIdx.x * p.grp_stride) * 128 + threadIdx.x;
  if (c >= p.ld) return;
  const int64_t beg = p.ranges ? p.ranges[2 * f] : (p.offsets ? p.offsets[p.fold0 + f] : 0);
  const int64_t n = p.ranges ? p.ranges[2 * f + 1] - beg : (p.offsets ? p.offsets[p.fold0 + f + 1] - beg : p.N);
  const int64_t* idx = (p.offsets || p.ranges) ? p.indices + beg : nullptr;
  T s = T(0), q = T(0);
  if (p.accumulate) {
    if (p.ranges) { s = p.raw[(size_t)f * 2 * p.ld + c]; q = p.raw[(size_t)f * 2 * p.ld + p.ld + c]; }
    else { s = p.sum_z[c]; q = p.sumsq_z[c]; }
  }
  int64_t i = 0;
  for (; i + 4 <= n; i += 4) {
    int64_t r[4]; T z[4], wv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) r[j] = idx ? idx[i + j] : p.row0 + i + j;
#pragma unroll
    for (int j = 0; j < 4; ++j) { z[j] = p.Z[r[j] * p.ld + c]; wv[j] = p.w[r[j]]; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const T wz = Rn<T>::mul(z[j], wv[j]);
      s = Rn<T>::add(s, wz);
      q = Rn<T>::add(q, Rn<T>::mul(wz, z[j]));
    }
  }
  for (; i < n; ++i) {
    const int64_t r = idx ? idx[i] : p.row0 + i;
    const T z = p.Z[r * p.ld + c];
    const T wz = Rn<T>::mul(z, p.w[r]);
    s = Rn<T>::add(s, wz);
    q = Rn<T>::add(q, Rn<T>::mul(wz, z));
  }
  finalize_column<T>(p, f, c, s, q);
}

// ---- long folds: one producer warp (cp.async) feeds two consumer warps that own 32 sequential column chains ---
// The chain - one dependent DADD per row and column, 8 cycles on B200 - is the critical path, and the FP64 pipe
// of one SM sub-partition issues a warp-wide DADD/DMUL only every ~2 cycles, so the sum chain (DMUL, DADD) and the
// sum-of-squares chain (DMUL, DMUL, DADD) run in two different warps (two sub-partitions); each stays below the
// 8-cycle chain latency (measured 14.6 cycles per row; a variant with extra "product" warps forming rn(rn(w z) z)
// ahead of the chain was tried and was slower, 1.8 ms vs 1.5 ms at cfg 2).  Two producer warps gather 64 rows per stage with coalesced 16-byte cp.async (a warp-wide
// instruction covers whole row segments; measured alternatives: per-row 256-byte TMA bulk copies are issue-bound,
// 6.5 ms vs 2.3 ms at cfg 2, and a lane-per-row mapping is sector-request-bound, 7.4 ms) into an MOM_STAGES-deep
// ring, each lane signalling the stage's mbarrier when its copies land (cp.async.mbarrier.arrive.noinc).  Row byte
// offsets are computed once per stage and broadcast by shuffle; row indices are fetched a group of stages early.
constexpr int MOM_COLS = 32;
constexpr int MOM_ROWS = 64;                           // rows per stage: amortises the ~160 cycles of per-stage bookkeeping (ncu)
constexpr int MOM_STAGES = 4;                          // default ring depth (MomentParams::stages overrides it)
constexpr int MOM_PRODUCERS = 2;                       // loader warps (4 were measured slower: 1.83 ms vs 1.50 ms at cfg 2)
constexpr int MOM_THREADS = 32 * (2 + MOM_PRODUCERS);  // warp 0: sum chain, warp 1: sum-of-squares chain, warps 2..: loaders
constexpr int MOM_GROUP = 4;                           // stages whose row indices are fetched together
template <typename T> struct MomCfg { static constexpr int PITCH = MOM_COLS + 16 / sizeof(T); };  // 16-byte pad per row

template <typename T>
__global__ void __launch_bounds__(MOM_THREADS) k_moments_pipe(MomentParams<T> p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int PITCH = MomCfg<T>::PITCH;
  T* sz = reinterpret_cast<T*>(smem_raw);                       // [STAGES][ROWS][PITCH]
  const int NST = p.stages;
  T* swt = sz + (size_t)NST * MOM_ROWS * PITCH;                  // [STAGES][ROWS]
  T* sres = swt + NST * MOM_ROWS;                                // [COLS] sum-of-squares hand-over
  uint64_t* full = reinterpret_cast<uint64_t*>(sres + MOM_COLS);
  uint64_t* empty = full + NST;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t f = blockIdx.y;
  if (p.scan_ok && p.scan_ok[f * p.scan_groups + (p.grp0 + blockIdx.x * p.grp_stride)]) return;   // whole CTA
  const int64_t c0 = (int64_t)(p.grp0 + blockIdx.x * p.grp_stride) * MOM_COLS;
  const int64_t beg = p.ranges ? p.ranges[2 * f] : (p.offsets ? p.offsets[p.fold0 + f] : 0);
  const int64_t n = p.ranges ? p.ranges[2 * f + 1] - beg : (p.offsets ? p.offsets[p.fold0 + f + 1] - beg : p.N);
  const int64_t* idx = (p.offsets || p.ranges) ? p.indices + beg : nullptr;
  const int64_t nst = (n + MOM_ROWS - 1) / MOM_ROWS;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(full + s, 32 * MOM_PRODUCERS); mbar_init(empty + s, 2); }
    mbar_fence_init();
  }
  __syncthreads();

  T acc = T(0);
  if (p.accumulate && warp < 2) {
    if (p.ranges) acc = p.raw[(size_t)f * 2 * p.ld + (warp == 0 ? 0 : p.ld) + c0 + lane];   // fold slices continue the fold's raw sums
    else acc = warp == 0 ? p.sum_z[c0 + lane] : p.sumsq_z[c0 + lane];
  }
  if (warp >= 2) {
    // ---------------- producers: warp 2 + pw owns rows [PROWS pw, PROWS (pw + 1)) of every stage ----------------
    // A warp-wide cp.async covers RPI whole row segments (coalesced 16-byte chunks); the byte offset of each row is
    // computed once per stage by the lane that fetched its index and broadcast with a shuffle.
    constexpr int PROWS = MOM_ROWS / MOM_PRODUCERS;          // rows of a stage owned by this warp
    constexpr int EPC = 16 / sizeof(T);                      // elements per 16-byte chunk
    constexpr int CPR = MOM_COLS / EPC;                      // chunks (lanes) per row segment
    constexpr int RPI = 32 / CPR;                            // rows covered by one warp-wide cp.async
    constexpr int ITER = PROWS / RPI;
    const int pw = warp - 2;
    const int lr = lane / CPR, lc = lane % CPR;
    const char* zbase = reinterpret_cast<const char*>(p.Z + c0 + lc * EPC);
    const long long row_bytes = (long long)p.ld * (long long)sizeof(T);
    int64_t cur[MOM_GROUP], nxt[MOM_GROUP];
    auto fetch = [&](int64_t st0, int64_t (&g)[MOM_GROUP]) {     // lane l holds the row of local row (l % PROWS)
#pragma unroll
      for (int j = 0; j < MOM_GROUP; ++j) {
        const int64_t row = (st0 + j) * MOM_ROWS + pw * PROWS + (lane % PROWS);
        g[j] = row < n ? (idx ? idx[row] : p.row0 + row) : -1;
      }
    };
    fetch(0, cur);
    for (int64_t st0 = 0; st0 < nst; st0 += MOM_GROUP) {
      fetch(st0 + MOM_GROUP, nxt);
#pragma unroll
      for (int j = 0; j < MOM_GROUP; ++j) {
        const int64_t st = st0 + j;
        if (st < nst) {
          const int slot = (int)(st % NST);
          const unsigned round = (unsigned)(st / NST);
          if (round > 0) mbar_wait(empty + slot, (round & 1) ^ 1);
          const int64_t g = cur[j];
          const long long my_off = g >= 0 ? g * row_bytes : -1;   // byte offset of "my" row, -1: past the end
          T* dst = sz + ((size_t)slot * MOM_ROWS + pw * PROWS + lr) * PITCH + lc * EPC;
#pragma unroll
          for (int i = 0; i < ITER; ++i) {
            const long long off = __shfl_sync(0xffffffffu, my_off, lr + i * RPI);
            cp_async16(dst + (size_t)i * RPI * PITCH, zbase + (off >= 0 ? off : 0), off >= 0 ? 16 : 0);
          }
          if (lane < PROWS) {
            T* wd = swt + slot * MOM_ROWS + pw * PROWS + lane;
            if (sizeof(T) == 8) cp_async8(wd, p.w + (g >= 0 ? g : 0), g >= 0 ? 8 : 0);
            else cp_async4(wd, p.w + (g >= 0 ? g : 0), g >= 0 ? 4 : 0);
          }
          cp_async_mbar_arrive_noinc(full + slot);
        }
      }
#pragma unroll
      for (int j = 0; j < MOM_GROUP; ++j) cur[j] = nxt[j];
    }
    cp_async_wait<0>();
  } else {
    // ---------------- consumers: warp 0 sums rn(w z), warp 1 sums rn(rn(w z) z) ----------------
    int slot = 0;
    unsigned parity = 0;
    int64_t left = n;
    {
      const T* zr = sz + lane;
      const T* wr = swt;
      for (int64_t st = 0; st < nst; ++st) {
        mbar_wait(full + slot, parity);
        if (left >= MOM_ROWS) {
          if (warp == 0) {
#pragma unroll
            for (int r = 0; r < MOM_ROWS; ++r) acc = Rn<T>::add(acc, Rn<T>::mul(zr[r * PITCH], wr[r]));
          } else {
#pragma unroll
            for (int r = 0; r < MOM_ROWS; ++r) {
              const T z = zr[r * PITCH];
              acc = Rn<T>::add(acc, Rn<T>::mul(Rn<T>::mul(z, wr[r]), z));
            }
          }
        } else {
          for (int r = 0; r < (int)left; ++r) {
            const T z = zr[r * PITCH];
            const T wz = Rn<T>::mul(z, wr[r]);
            acc = Rn<T>::add(acc, warp == 0 ? wz : Rn<T>::mul(wz, z));
          }
        }
        left -= MOM_ROWS;
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);
        zr += MOM_ROWS * PITCH;
        wr += MOM_ROWS;
        if (++slot == NST) { slot = 0; parity ^= 1; zr = sz + lane; wr = swt; }
      }
    }
    if (warp == 1) sres[lane] = acc;
  }
  __syncthreads();
  if (warp == 0) finalize_column<T>(p, f, c0 + lane, acc, sres[lane]);
}

template <typename T>
constexpr size_t moments_pipe_smem(int stages) {
  return sizeof(T) * ((size_t)stages * MOM_ROWS * MomCfg<T>::PITCH + (size_t)stages * MOM_ROWS + MOM_COLS) +
         2 * stages * sizeof(uint64_t);
}
constexpr int MOM_STAGES_DEEP = 8;      // few long chains: ~140 KB of gathered rows in flight per CTA

// Validation rows of a fold for the caller's next step (prediction on the held-out rows): out[i][c] = Z[idx[i]][c0 + c],
// optionally (z - mean) / std with the training-set statistics, each op individually rounded like numpy's
// (X[val] - X_mean) / X_std.
template <typename T>
__global__ void k_validation_rows(const T* __restrict__ Z, int64_t ld, const int64_t* __restrict__ idx, int64_t n, int64_t c0,
                                  int64_t ncols, const T* __restrict__ mean, const T* __restrict__ sdev, T* __restrict__ out) {
  const int64_t total = n * ncols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / ncols, c = e - r * ncols;
    T v = Z[idx[r] * ld + c0 + c];
    if (mean) v = Rn<T>::sub(v, mean[c0 + c]);
    if (sdev) v = Rn<T>::div(v, sdev[c0 + c]);
    out[e] = v;
  }
}

// mean / std of every (fold, column) from the raw fold sums (deferred finalisation, see MomentParams::raw)
template <typename T>
__global__ void k_finalize_stats(MomentParams<T> p, int64_t nfolds) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t f = blockIdx.y;
  if (c >= p.ld || f >= nfolds) return;
  // column-sharded evaluation: only this shard's groups carry sums (p.stages = columns per group here); foreign
  // entries stay zero so that the all-reduce across ranks assembles the rows
  if (p.grp_stride > 1 && (int)((c / p.stages) % p.grp_stride) != p.grp0) return;
  const T* r = p.raw + (size_t)f * 2 * p.ld;
  MomentParams<T> q = p;
  q.raw = nullptr;
  finalize_column<T>(q, f, c, r[c], r[p.ld + c]);
}

}  // namespace cvmx
