// Speculative binade scan: numpy-order (strictly sequential) column sums WITHOUT the dependent-add chain.
//
// np.sum(A, axis=0) adds the rows one after another (cvmatrix/cvmatrix.py:709, 716, 727, 737, 1231-1241), so a
// bit-identical GPU result looks like one dependent DADD per row and column (k_moments_pipe: 8-cycle latency per
// row, 15 cycles measured) - a floor that no amount of SMs removes.  This file removes it for float64:
//
//   While the running sum s stays inside one binade [2^E, 2^(E+1)) every representable value there is a multiple
//   of u = 2^(E-52), and fl(s + x) is "round s + x to a multiple of u, ties to the even multiple".  The rounded
//   increment therefore depends on s only through the PARITY of s / u.  Consequence: a chain started from ANY
//   value of the same binade with the same parity makes exactly the same rounding decisions, step for step.
//
//   pass 1  k_scan_segsums : rows are cut into segments of L rows; every (segment, column) gets an approximate sum
//                            S and the sum of magnitudes A, in parallel.
//   pass 2  k_scan_prefix  : per column, prefix P_j of the S (cheap: n / L terms).  With the classical error bound
//                            |true running sum - P_j| <= 2^-24 * (sum of magnitudes so far), the interval
//                            [P_j - A_j - margin, P_j + A_j + margin] contains every value the true chain takes
//                            inside segment j.  If the interval lies inside one binade the segment is "fast" and
//                            gets the proxy start B = +-1.5 * 2^E; otherwise it is "slow".
//   pass 3  k_scan_delta   : every fast (segment, column) runs the real chain from the two proxies B (even) and
//                            B + u (odd) - independent chains, one per thread, thousands in flight - and stores the
//                            two exact increments d0 = end0 - B, d1 = end1 - (B + u)  (multiples of u: exact).
//   pass 4  k_scan_chain   : per column, one step per SEGMENT: s += parity(s) ? d1 : d0 (an exact addition);
//                            slow segments (the first few, and the ~1 per binade crossing) are added row by row.
//
// The result is bit-identical to the sequential chain for every input (slow segments fall back to it; NaN / inf /
// zero / cancelling sums are simply never fast).  tests/test_scan_model.py holds a numpy model of the same four
// passes and its adversarial cases; tests/test_gpu_scan.py compares the scan with the chains bit for bit.
// Groups of 32 columns whose segments are mostly slow (e.g. mean-centred data: the sum wanders around zero) are
// handed to k_moments_pipe instead (flag written by pass 2, read by passes 3 / 4 and by k_moments_pipe).
#pragma once
#include "kernels_stats.cuh"

namespace cvmx {

constexpr int SCAN_COLS = MOM_COLS;   // same column groups as k_moments_pipe (column sharding is per group)
constexpr int SCAN_WARPS = 4;         // segments per CTA in the two streaming passes
constexpr int SCAN_L = 128;           // rows per segment (multiple of 32; compile-time in pass 4 and in k_scan_fused)

struct ScanParams {
  MomentParams<double> p;
  double* seg;          // [folds][chain: sum, sum of squares][slot 0, 1][ld][max_segs]: segment index fastest, so the
                        // per-column passes 2 / 4 read contiguous memory; the 4 warps of a pass-1 / 3 CTA (4
                        // consecutive segments) fill one 32-byte sector per column
  int* ok;              // [folds][groups_total]: 1 = the scan result stands, 0 = k_moments_pipe recomputes the group
  int* slow_list;       // [folds][chain][ld][slow_cap]: the slow segments of every column, ascending (written by pass 2)
  int* slow_cnt;        // [folds][chain][ld]
  int slow_cap;
  int L;
  int64_t max_segs;
  int groups_total;
};

// element (segment s, column c) of plane (fold f, chain, slot)
__device__ __forceinline__ double* scan_plane(const ScanParams& sp, int64_t f, int chain, int slot, int64_t c) {
  return sp.seg + ((((size_t)f * 2 + chain) * 2 + slot) * (size_t)sp.p.ld + (size_t)c) * (size_t)sp.max_segs;
}

// rows [r0, r0 + cnt) of a fold, 32 at a time: lane l fetches the index and weight of row l of the batch, the
// column values are loaded 8 rows deep, and fn(t = rn(w z), q = rn(t z)) is called in row order
template <typename F>
__device__ __forceinline__ void scan_rows(const MomentParams<double>& p, const int64_t* idx, int64_t r0, int cnt, int lane,
                                          const double* zc, F&& fn) {
  for (int base = 0; base < cnt; base += 32) {
    const int m = min(32, cnt - base);
    long long row = 0;
    double wv = 0.0;
    if (lane < m) {
      row = idx ? idx[r0 + base + lane] : p.row0 + r0 + base + lane;
      wv = p.w[row];
    }
    const long long off = row * p.ld;
#pragma unroll 1
    for (int k = 0; k < m; k += 8) {
      double z[8], wk[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long o = __shfl_sync(0xffffffffu, off, (k + u) & 31);
        wk[u] = __shfl_sync(0xffffffffu, wv, (k + u) & 31);
        z[u] = (k + u < m) ? zc[o] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (k + u < m) {
          const double t = __dmul_rn(z[u], wk[u]);
          fn(t, __dmul_rn(t, z[u]));
        }
      }
    }
  }
}

// ---- pass 1 -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * SCAN_WARPS) k_scan_segsums(ScanParams sp) {
  const MomentParams<double>& p = sp.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t f = blockIdx.z;
  const int64_t c = (int64_t)(p.grp0 + blockIdx.x * p.grp_stride) * SCAN_COLS + lane;
  const int64_t beg = p.offsets ? p.offsets[p.fold0 + f] : 0;
  const int64_t n = p.offsets ? p.offsets[p.fold0 + f + 1] - beg : p.N;
  const int64_t* idx = p.offsets ? p.indices + beg : nullptr;
  if (blockIdx.y == 0 && threadIdx.x == 0) sp.ok[f * sp.groups_total + p.grp0 + blockIdx.x * p.grp_stride] = 1;   // pass 2 clears it
  // gridDim.y may be smaller than the number of segment quads (a grid capped to a few SMs' worth of CTAs while a
  // Gram kernel owns the rest of the GPU): CTAs stride over the quads
  for (int64_t s = (int64_t)blockIdx.y * SCAN_WARPS + warp; s * sp.L < n; s += (int64_t)gridDim.y * SCAN_WARPS) {
    const int64_t r0 = s * sp.L;
    const int cnt = (int)min((int64_t)sp.L, n - r0);
    double S = 0.0, A = 0.0, Q = 0.0;
    unsigned negq = 0;
    scan_rows(p, idx, r0, cnt, lane, p.Z + c, [&](double t, double q) {
      S = __dadd_rn(S, t);
      A = __dadd_rn(A, fabs(t));
      Q = __dadd_rn(Q, fabs(q));
      negq |= (unsigned)(__double2hiint(q) & 0x80000000);
    });
    // q = rn(rn(w z) z) is non-negative unless a weight is negative (fit rejects those) - then the squares chain is
    // never fast (NaN magnitude)
    scan_plane(sp, f, 0, 0, c)[s] = S;
    scan_plane(sp, f, 0, 1, c)[s] = A;
    scan_plane(sp, f, 1, 0, c)[s] = Q;
    scan_plane(sp, f, 1, 1, c)[s] = negq ? __longlong_as_double(0x7ff8000000000000LL) : Q;
  }
}

// ---- pass 2 -----------------------------------------------------------------------------------------------------
// SCAN_PREFIX_CTAS CTAs per (column group, fold), one warp per column.  256 segments per trip: every lane owns 8 consecutive
// segments (16 independent loads in flight), prefixes them locally, and a warp scan supplies the lane offsets (the
// approximation may use any summation order).  Overwrites slot 0 with the proxy start (+0: slow, -0: the segment
// holds only zeros, else +-1.5 * 2^E), lists the slow segments of the column in ascending order, and flags the group.
constexpr int SCAN_PREFIX_CTAS = 4;                                   // per column group
constexpr int SCAN_PREFIX_THREADS = 32 * SCAN_COLS / SCAN_PREFIX_CTAS;
constexpr int SCAN_PER_LANE = 8;                                      // max_segs is a multiple of this (16-byte loads)
__global__ void __launch_bounds__(SCAN_PREFIX_THREADS) k_scan_prefix(ScanParams sp) {
  const MomentParams<double>& p = sp.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t f = blockIdx.y;
  const int g = p.grp0 + (blockIdx.x / SCAN_PREFIX_CTAS) * p.grp_stride;
  const int64_t c = (int64_t)g * SCAN_COLS + (blockIdx.x % SCAN_PREFIX_CTAS) * (SCAN_COLS / SCAN_PREFIX_CTAS) + warp;
  const int64_t beg = p.offsets ? p.offsets[p.fold0 + f] : 0;
  const int64_t n = p.offsets ? p.offsets[p.fold0 + f + 1] - beg : p.N;
  const int64_t nseg = (n + sp.L - 1) / sp.L;
  int* okp = sp.ok + f * sp.groups_total + g;
  if (nseg < 4 && threadIdx.x == 0) atomicAnd(okp, 0);
  if (c < p.K + p.M) {
    for (int chain = 0; chain < 2; ++chain) {
      double P = 0.0, tot = 0.0;
      if (p.accumulate) P = chain == 0 ? p.sum_z[c] : p.sumsq_z[c];
      double* dS = scan_plane(sp, f, chain, 0, c);
      const double* dA = scan_plane(sp, f, chain, 1, c);
      int* list = sp.slow_list + ((size_t)(f * 2 + chain) * p.ld + c) * sp.slow_cap;
      int nslow = 0;
      for (int64_t J = 0; J < nseg; J += 32 * SCAN_PER_LANE) {
        const int64_t j0 = J + lane * SCAN_PER_LANE;
        double Sj[SCAN_PER_LANE], Aj[SCAN_PER_LANE];
#pragma unroll
        for (int e = 0; e < SCAN_PER_LANE; e += 2) {   // the planes are padded to a multiple of SCAN_PER_LANE segments
          const bool in = j0 + e < sp.max_segs;
          const double2 s2 = in ? *reinterpret_cast<const double2*>(dS + j0 + e) : make_double2(0.0, 0.0);
          const double2 a2 = in ? *reinterpret_cast<const double2*>(dA + j0 + e) : make_double2(0.0, 0.0);
          Sj[e] = j0 + e < nseg ? s2.x : 0.0;
          Sj[e + 1] = j0 + e + 1 < nseg ? s2.y : 0.0;
          Aj[e] = j0 + e < nseg ? a2.x : 0.0;
          Aj[e + 1] = j0 + e + 1 < nseg ? a2.y : 0.0;
        }
        double lS = 0.0, lA = 0.0;
#pragma unroll
        for (int e = 0; e < SCAN_PER_LANE; ++e) { lS += Sj[e]; lA += Aj[e]; }
        double sS = lS, sA = lA;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double tS = __shfl_up_sync(0xffffffffu, sS, o), tA = __shfl_up_sync(0xffffffffu, sA, o);
          if (lane >= o) { sS += tS; sA += tA; }
        }
        double Pj = __shfl_up_sync(0xffffffffu, sS, 1), totj = __shfl_up_sync(0xffffffffu, sA, 1);
        Pj = P + (lane == 0 ? 0.0 : Pj);
        totj = tot + (lane == 0 ? 0.0 : totj);
        unsigned slowbits = 0;
#pragma unroll
        for (int e = 0; e < SCAN_PER_LANE; ++e) {
          const double A1 = Aj[e] * (1.0 + 0x1p-20);
          const double margin = 0x1p-24 * totj + 0x1p-40 * fabs(Pj);
          const double lo = (Pj - A1) - margin, hi = (Pj + A1) + margin;
          const long long blo = __double_as_longlong(lo), bhi = __double_as_longlong(hi);
          const int elo = (int)((blo >> 52) & 0x7ff), ehi = (int)((bhi >> 52) & 0x7ff);
          const bool same = ((blo ^ bhi) >= 0) && elo == ehi && elo >= 1 && elo <= 2046;
          double B = 0.0;
          if (Aj[e] == 0.0) B = -0.0;                                                   // only zeros: s + (+-0)
          else if (same) B = __longlong_as_double((blo & (long long)0xfff0000000000000ULL) | 0x0008000000000000LL);
          if (j0 + e < nseg) {
            dS[j0 + e] = B;
            if (__double_as_longlong(B) == 0) slowbits |= 1u << e;
          }
          Pj += Sj[e];
          totj += Aj[e];
        }
        // ascending list of slow segments: lane-major order is segment order
        const int mine = __popc(slowbits);
        int before = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, before, o);
          if (lane >= o) before += t;
        }
        const int total = __shfl_sync(0xffffffffu, before, 31);
        int pos = nslow + before - mine;
        for (unsigned bits = slowbits; bits; bits &= bits - 1, ++pos)
          if (pos < sp.slow_cap) list[pos] = (int)(j0 + __ffs(bits) - 1);
        nslow += total;
        P = P + __shfl_sync(0xffffffffu, sS, 31);
        tot = tot + __shfl_sync(0xffffffffu, sA, 31);
      }
      if (lane == 0) {
        sp.slow_cnt[(size_t)(f * 2 + chain) * p.ld + c] = nslow;
        if (nslow > nseg / 4 || nslow > sp.slow_cap) atomicAnd(okp, 0);   // a mostly-slow column: leave the group to the chains
      }
    }
  }
}

// ---- pass 3 -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * SCAN_WARPS) k_scan_delta(ScanParams sp) {
  const MomentParams<double>& p = sp.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t f = blockIdx.z;
  const int g = p.grp0 + blockIdx.x * p.grp_stride;
  if (!sp.ok[f * sp.groups_total + g]) return;
  const int64_t c = (int64_t)g * SCAN_COLS + lane;
  const int64_t beg = p.offsets ? p.offsets[p.fold0 + f] : 0;
  const int64_t n = p.offsets ? p.offsets[p.fold0 + f + 1] - beg : p.N;
  const int64_t* idx = p.offsets ? p.indices + beg : nullptr;
  for (int64_t s = (int64_t)blockIdx.y * SCAN_WARPS + warp; s * sp.L < n; s += (int64_t)gridDim.y * SCAN_WARPS) {
    const int64_t r0 = s * sp.L;
    const int cnt = (int)min((int64_t)sp.L, n - r0);
    double *o00 = scan_plane(sp, f, 0, 0, c) + s, *o01 = scan_plane(sp, f, 0, 1, c) + s;
    double *o10 = scan_plane(sp, f, 1, 0, c) + s, *o11 = scan_plane(sp, f, 1, 1, c) + s;
    const double Bs = *o00, Bq = *o10;
    const long long bs = __double_as_longlong(Bs), bq = __double_as_longlong(Bq);
    if (__all_sync(0xffffffffu, bs == 0 && bq == 0)) continue;
    // identity segments (B = -0) use the same start for both parities; otherwise the odd proxy is B + u
    const bool ids = bs == (long long)0x8000000000000000ULL, idq = bq == (long long)0x8000000000000000ULL;
    const double Bs1 = ids ? Bs : __longlong_as_double(bs | 1), Bq1 = idq ? Bq : __longlong_as_double(bq | 1);
    double c0 = Bs, c1 = Bs1, e0 = Bq, e1 = Bq1;
    scan_rows(p, idx, r0, cnt, lane, p.Z + c, [&](double t, double q) {
      c0 = __dadd_rn(c0, t);
      c1 = __dadd_rn(c1, t);
      e0 = __dadd_rn(e0, q);
      e1 = __dadd_rn(e1, q);
    });
    if (bs != 0) {
      *o00 = ids ? c0 : __dsub_rn(c0, Bs);
      *o01 = ids ? c1 : __dsub_rn(c1, Bs1);
    }
    if (bq != 0) {
      *o10 = idq ? e0 : __dsub_rn(e0, Bq);
      *o11 = idq ? e1 : __dsub_rn(e1, Bq1);
    }
  }
}

// ---- pass 4 -----------------------------------------------------------------------------------------------------
// One warp per (column, chain); the running sum is warp-uniform.  A fast segment costs one exact addition (operands
// staged 32 segments at a time through shared memory, the next 32 in flight).  Slow segments come from the list of
// pass 2 and run a three-stage software pipeline: row indices of slow segment k + 2 and the gathered values of k + 1
// are in flight while the 32 lanes' products of segment k are added row by row, in the reference order.
// Columns per CTA (x 2 chains = warps): 2 puts one warp on each SM sub-partition, so the dependent adds do not
// queue behind each other on the FP64 pipe (8 columns per CTA: 134 us instead of 90 us at cfg 2 / 8 shards).
constexpr int SCAN_NB = SCAN_L / 32;
template <int COLS> constexpr size_t scan_chain_smem() { return (size_t)2 * COLS * (SCAN_L + 64) * sizeof(double); }
static_assert(SCAN_L % 32 == 0, "segments are gathered 32 rows at a time");

template <int SCAN_CHAIN_COLS>
__global__ void __launch_bounds__(64 * SCAN_CHAIN_COLS) k_scan_chain(ScanParams sp) {
  extern __shared__ __align__(16) unsigned char scan_smem[];
  __shared__ double sres[SCAN_CHAIN_COLS];
  const MomentParams<double>& p = sp.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, chain = warp & 1, lc = warp >> 1;
  constexpr int CPG = SCAN_COLS / SCAN_CHAIN_COLS;   // CTAs per column group
  const int64_t f = blockIdx.y;
  const int g = p.grp0 + (blockIdx.x / CPG) * p.grp_stride;
  if (!sp.ok[f * sp.groups_total + g]) return;
  const int64_t c = (int64_t)g * SCAN_COLS + (blockIdx.x % CPG) * SCAN_CHAIN_COLS + lc;
  const int64_t beg = p.offsets ? p.offsets[p.fold0 + f] : 0;
  const int64_t n = p.offsets ? p.offsets[p.fold0 + f + 1] - beg : p.N;
  const int64_t* idx = p.offsets ? p.indices + beg : nullptr;
  const int64_t nseg = (n + SCAN_L - 1) / SCAN_L;
  double* sv = reinterpret_cast<double*>(scan_smem) + (size_t)warp * (SCAN_L + 64);   // [SCAN_L] slow-segment values
  double2* sd = reinterpret_cast<double2*>(sv + SCAN_L);                               // [32] (d0, d1) of the current trip
  double acc = 0.0;
  if (c < p.K + p.M) {
    if (p.accumulate) acc = chain == 0 ? p.sum_z[c] : p.sumsq_z[c];
    const double* d0p = scan_plane(sp, f, chain, 0, c);
    const double* d1p = scan_plane(sp, f, chain, 1, c);
    const double* zc = p.Z + c;
    const int* list = sp.slow_list + ((size_t)(f * 2 + chain) * p.ld + c) * sp.slow_cap;
    const int nslow = sp.slow_cnt[(size_t)(f * 2 + chain) * p.ld + c];

    // slow-segment pipeline state: rows of segment k + 1 (after advance: k + 2), gathered z / w of segment k (k + 1)
    long long rowA[SCAN_NB];
    double zB[SCAN_NB], wB[SCAN_NB];
    auto load_rows = [&](int k) {
      const int64_t r0 = k < nslow ? (int64_t)list[k] * SCAN_L : n;
#pragma unroll
      for (int b = 0; b < SCAN_NB; ++b) {
        const int64_t r = r0 + 32 * b + lane;
        rowA[b] = (r < n && 32 * b + lane < SCAN_L) ? (idx ? idx[r] : p.row0 + r) : -1;
      }
    };
    auto load_vals = [&]() {
#pragma unroll
      for (int b = 0; b < SCAN_NB; ++b) {
        zB[b] = rowA[b] >= 0 ? zc[rowA[b] * p.ld] : 0.0;
        wB[b] = rowA[b] >= 0 ? p.w[rowA[b]] : 0.0;
      }
    };
    int k = 0;
    load_rows(0);
    load_vals();
    load_rows(1);
    int64_t next_slow = nslow > 0 ? list[0] : nseg;

    double x0 = lane < nseg ? d0p[lane] : 0.0, x1 = lane < nseg ? d1p[lane] : 0.0;
    for (int64_t J = 0; J < nseg; J += 32) {
      __syncwarp();
      sd[lane] = make_double2(x0, x1);
      {
        const int64_t jn = J + 32 + lane;                             // next 32 segments: loads fly during this trip
        x0 = jn < nseg ? d0p[jn] : 0.0;
        x1 = jn < nseg ? d1p[jn] : 0.0;
      }
      __syncwarp();
      const int m = (int)min((int64_t)32, nseg - J);
      int u = 0;
      while (u < m) {
        const int stop = (int)min((int64_t)m, next_slow - J);
#pragma unroll 4
        for (; u < stop; ++u) {
          const double2 a = sd[u];
          const double r0v = __dadd_rn(acc, a.x), r1v = __dadd_rn(acc, a.y);
          acc = (__double2loint(acc) & 1) ? r1v : r0v;
        }
        if (u < m) {   // segment J + u is slow: its gathered values are in zB / wB
          const int64_t r0 = (J + u) * SCAN_L;
          const int cnt = (int)min((int64_t)SCAN_L, n - r0);
#pragma unroll
          for (int b = 0; b < SCAN_NB; ++b) {
            const double t = __dmul_rn(zB[b], wB[b]);
            sv[32 * b + lane] = chain == 0 ? t : __dmul_rn(t, zB[b]);
          }
          __syncwarp();
          ++k;
          load_vals();          // segment k (rows already here)
          load_rows(k + 1);
          next_slow = k < nslow ? list[k] : nseg;
          int i = 0;
          for (; i + 4 <= cnt; i += 4) {
            const double2 v01 = *reinterpret_cast<const double2*>(sv + i), v23 = *reinterpret_cast<const double2*>(sv + i + 2);
            acc = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(acc, v01.x), v01.y), v23.x), v23.y);
          }
          for (; i < cnt; ++i) acc = __dadd_rn(acc, sv[i]);
          __syncwarp();
          ++u;
        }
      }
    }
  }
  if (chain == 1 && lane == 0) sres[lc] = acc;
  __syncthreads();
  if (chain == 0 && lane == 0) finalize_column<double>(p, f, c, acc, sres[lc]);
}

// ---- passes 1 - 3 in ONE read of the rows (decoupled look-back) ----------------------------------------------------
// The four-pass form reads every row twice (pass 1 and pass 3), and that second read is what the scan costs beside a
// Gram kernel that owns the tensor pipe but not the memory system.  k_scan_fused stages one segment (SCAN_L rows x 32
// columns) in shared memory ONCE and does everything that needs the rows from there:
//   1. segment sums S, A, Q (any order) from shared memory;
//   2. the prefix over all earlier segments of the same (fold, column group) by decoupled look-back (Merrill & Garland):
//      the CTA publishes its aggregate, then a warp polls the status words of up to 32 predecessors at a time
//      (lane = predecessor), adds their aggregates down to the nearest published inclusive prefix (lane = column) and
//      publishes its own inclusive prefix;
//   3. the proxy starts (same interval test as k_scan_prefix) and the four proxy chains per column (2 sums x 2 parities,
//      one warp each) - again from shared memory.
// CTAs are persistent and take (sequence, segment) tasks in ticket order from an atomic counter, dealt round-robin over
// the (fold, group) sequences so that every sequence advances together; a CTA only ever waits for tickets smaller
// than its own, and the smallest unfinished ticket is always being processed, so look-back cannot deadlock.  Every
// CTA prefetches the rows of its NEXT task (cp.async into the other half of a double buffer) before it starts on the
// current one: without that, look-back couples each task to its 32 predecessors and the whole GPU falls into
// lock-step load / compute phases (measured: 2 TB/s; the first version of this kernel).
// Slow segments are marked in the d0 plane with a reserved NaN payload (a fast segment's increments are finite);
// k_scan_lists turns the marks into the ascending per-column lists k_scan_chain walks.  The approximate prefix now
// depends on the order in which look-back happened to add the aggregates; that only moves segments between "fast" and
// "slow" - the error margin covers any summation order - never the result, which stays bit-identical to the chain.
constexpr int SF_THREADS = 128;
constexpr unsigned long long SCAN_SLOW_MARK = 0x7ff8dead00000000ULL;

struct ScanLook {
  double* agg;        // [nseq][max_segs][3][32]: S, A, Q of a segment
  double* inc;        // same shape: inclusive prefixes
  int* status;        // [nseq][max_segs]: 0 nothing yet, 1 aggregate published, 2 inclusive prefix published
  unsigned* ticket;   // work counter (zeroed with the status words before every launch)
  int nseq;           // sequences = folds x own column groups
  int mine;           // own column groups per fold
  unsigned total;     // tasks = nseq x segments of the longest fold
};

constexpr size_t scan_fused_smem() {
  return 2 * ((size_t)SCAN_L * SCAN_COLS * sizeof(double) + (size_t)SCAN_L * sizeof(double))   // two segments + their weights
         + 3072                                           // row numbers, later the cross-warp sums [4][3][32]
         + 3 * 32 * sizeof(double)                        // exclusive prefix
         + 3 * 32 * sizeof(double)                        // segment sums
         + 2 * 32 * sizeof(double)                        // proxy starts
         + 64;
}
static_assert(SCAN_L * 8 <= 3072 && SCAN_L % 32 == 0, "row-number scratch");

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

struct ScanTask {       // one (sequence, segment)
  int seq; int g; int cnt; bool valid;
  int64_t seg, f, c0, n, r0;
  const int64_t* idx;
};

__global__ void __launch_bounds__(SF_THREADS, 3) k_scan_fused(ScanParams sp, ScanLook lk) {
  extern __shared__ __align__(128) unsigned char sf_smem[];
  const MomentParams<double>& p = sp.p;
  double* tile0 = reinterpret_cast<double*>(sf_smem);                      // [2][SCAN_L][32]
  double* swt0 = tile0 + (size_t)2 * SCAN_L * SCAN_COLS;                   // [2][SCAN_L]
  long long* srow = reinterpret_cast<long long*>(swt0 + 2 * SCAN_L);       // [SCAN_L]   (only while copies are issued)
  double* red = reinterpret_cast<double*>(srow);                           // [4][3][32] (aliases srow)
  double* ex = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(srow) + 3072);   // [3][32]
  double* segs = ex + 96;                                                  // [3][32]
  double* prox = segs + 96;                                                // [2][32]
  unsigned* misc = reinterpret_cast<unsigned*>(prox + 64);                 // [0] ticket, [1] negative-square flags
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // next valid task in ticket order (tasks past the end of a short fold are skipped); block-uniform
  auto take = [&]() -> ScanTask {
    ScanTask t;
    t.valid = false;
    while (true) {
      __syncthreads();
      if (tid == 0) misc[0] = atomicAdd(lk.ticket, 1u);
      __syncthreads();
      const unsigned ticket = misc[0];
      if (ticket >= lk.total) return t;
      t.seq = (int)(ticket % (unsigned)lk.nseq);
      t.seg = (int64_t)(ticket / (unsigned)lk.nseq);
      t.f = t.seq / lk.mine;
      t.g = p.grp0 + (t.seq % lk.mine) * p.grp_stride;
      t.c0 = (int64_t)t.g * SCAN_COLS;
      const int64_t beg = p.offsets ? p.offsets[p.fold0 + t.f] : 0;
      t.n = p.offsets ? p.offsets[p.fold0 + t.f + 1] - beg : p.N;
      t.idx = p.offsets ? p.indices + beg : nullptr;
      if (t.seg == 0 && tid == 0) sp.ok[t.f * sp.groups_total + t.g] = 1;  // k_scan_lists may clear it
      t.r0 = t.seg * SCAN_L;
      if (t.r0 >= t.n) continue;
      t.cnt = (int)min((int64_t)SCAN_L, t.n - t.r0);
      t.valid = true;
      return t;
    }
  };
  // stage the rows of a task: row numbers and weights first, then 16-byte chunks (16 lanes cover one 256-byte row piece)
  auto issue = [&](const ScanTask& t, int buf) {
    double* tile = tile0 + (size_t)buf * SCAN_L * SCAN_COLS;
    double* swt = swt0 + buf * SCAN_L;
    __syncthreads();                                       // srow / red free again
    for (int r = tid; r < SCAN_L; r += SF_THREADS) {
      const long long row = r < t.cnt ? (t.idx ? t.idx[t.r0 + r] : p.row0 + t.r0 + r) : -1;
      srow[r] = row;
      cp_async8(swt + r, p.w + (row >= 0 ? row : 0), row >= 0 ? 8 : 0);
    }
    __syncthreads();
    const int chunk = tid & 15, rr = tid >> 4;
    const char* zbase = reinterpret_cast<const char*>(p.Z + t.c0 + chunk * 2);
    const long long row_bytes = (long long)p.ld * 8;
#pragma unroll 8
    for (int k = 0; k < SCAN_L / 8; ++k) {
      const int r = rr + 8 * k;
      const long long row = srow[r];
      cp_async16(tile + (size_t)r * SCAN_COLS + chunk * 2, zbase + (row >= 0 ? row : 0) * row_bytes, row >= 0 ? 16 : 0);
    }
    cp_async_commit();
  };

  ScanTask cur = take();
  int buf = 0;
  if (cur.valid) issue(cur, buf);
  while (cur.valid) {
    ScanTask nxt = take();
    if (nxt.valid) { issue(nxt, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const double* tile = tile0 + (size_t)buf * SCAN_L * SCAN_COLS;
    const double* swt = swt0 + buf * SCAN_L;
    const int seq = cur.seq, cnt = cur.cnt;
    const int64_t seg = cur.seg, f = cur.f, c0 = cur.c0;
    if (tid == 0) misc[1] = 0u;

    // ---- 1. segment sums (rows past the end are zero-filled: t = q = +0) -------------------------------------------
    {
      double S[4] = {0, 0, 0, 0}, A[4] = {0, 0, 0, 0}, Q[4] = {0, 0, 0, 0};
      unsigned neg = 0;
      const double* zt = tile + (size_t)(warp * (SCAN_L / 4)) * SCAN_COLS + lane;
      const double* wt = swt + warp * (SCAN_L / 4);
#pragma unroll 4
      for (int r = 0; r < SCAN_L / 4; r += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double z = zt[(size_t)(r + u) * SCAN_COLS];
          const double t = __dmul_rn(z, wt[r + u]);
          const double q = __dmul_rn(t, z);
          S[u] = __dadd_rn(S[u], t);
          A[u] = __dadd_rn(A[u], fabs(t));
          Q[u] = __dadd_rn(Q[u], fabs(q));
          neg |= (unsigned)(__double2hiint(q) & 0x80000000);
        }
      }
      __syncthreads();                                     // misc[1] cleared
      red[(warp * 3 + 0) * 32 + lane] = (S[0] + S[1]) + (S[2] + S[3]);
      red[(warp * 3 + 1) * 32 + lane] = (A[0] + A[1]) + (A[2] + A[3]);
      red[(warp * 3 + 2) * 32 + lane] = (Q[0] + Q[1]) + (Q[2] + Q[3]);
      if (neg) atomicOr(&misc[1], 1u << lane);
    }
    __syncthreads();

    // ---- 2. look-back (warp 0; lane = column for values, lane = predecessor for status words) ----------------------
    if (warp == 0) {
      double a[3];
#pragma unroll
      for (int v = 0; v < 3; ++v) a[v] = (red[(0 * 3 + v) * 32 + lane] + red[(1 * 3 + v) * 32 + lane]) + (red[(2 * 3 + v) * 32 + lane] + red[(3 * 3 + v) * 32 + lane]);
      // q = rn(rn(w z) z) is non-negative unless a weight is negative (fit rejects those): then the squares chain is never fast
      if ((misc[1] >> lane) & 1u) a[2] = __longlong_as_double(0x7ff8000000000000LL);
      const int64_t c = c0 + lane;
      const size_t base = ((size_t)seq * sp.max_segs) * 96;
      double* myagg = lk.agg + base + (size_t)seg * 96;
      double* myinc = lk.inc + base + (size_t)seg * 96;
      int* st = lk.status + (size_t)seq * sp.max_segs;
      double e[3] = {0.0, 0.0, 0.0};
      if (seg == 0) {
        if (p.accumulate && c < p.K + p.M) { e[0] = p.sum_z[c]; e[2] = p.sumsq_z[c]; }   // chunked fit: the chains continue
      } else {
#pragma unroll
        for (int v = 0; v < 3; ++v) __stcg(myagg + v * 32 + lane, a[v]);
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release_gpu(st + seg, 1);
        int64_t j = seg - 1;                     // nearest predecessor not yet added
        unsigned long long spins = 0;
        while (true) {
          const int64_t pj = j - lane;
          const int s = pj >= 0 ? ld_acquire_gpu(st + pj) : 3;
          const unsigned has_inc = __ballot_sync(0xffffffffu, s == 2);
          const int stop = has_inc ? __ffs(has_inc) - 1 : 32;              // predecessors j .. j - stop (the last one inclusive)
          const unsigned need = stop >= 31 ? 0xffffffffu : ((2u << stop) - 1u);
          const unsigned ready = __ballot_sync(0xffffffffu, s != 0);
          if ((ready & need) != need) {
            if (++spins > (1ull << 24)) __trap();                          // a lost predecessor must not hang the GPU
            __nanosleep(40);
            continue;
          }
          __threadfence();
          const int last = has_inc ? stop : 31;
          for (int k0 = 0; k0 <= last; k0 += 8) {
            double v[8][3];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int k = k0 + u;
              const bool on = k <= last && j - k >= 0;
              const double* src = ((has_inc && k == stop) ? lk.inc : lk.agg) + base + (size_t)(on ? j - k : 0) * 96;
#pragma unroll
              for (int w3 = 0; w3 < 3; ++w3) v[u][w3] = on ? __ldcg(src + w3 * 32 + lane) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
              for (int w3 = 0; w3 < 3; ++w3) e[w3] += v[u][w3];
          }
          if (has_inc) break;
          j -= 32;
        }
      }
#pragma unroll
      for (int v = 0; v < 3; ++v) { ex[v * 32 + lane] = e[v]; segs[v * 32 + lane] = a[v]; __stcg(myinc + v * 32 + lane, e[v] + a[v]); }
      __threadfence();
      __syncwarp();
      if (lane == 0) st_release_gpu(st + seg, 2);
    }
    __syncthreads();

    // ---- 3a. proxy starts (threads 0..63: chain = warp, column = lane; same test as k_scan_prefix) ------------------
    if (warp < 2) {
      const int chain = warp;
      const double Pj = chain == 0 ? ex[lane] : ex[64 + lane];
      const double totj = chain == 0 ? ex[32 + lane] : ex[64 + lane];
      const double Aj = chain == 0 ? segs[32 + lane] : segs[64 + lane];
      const double A1 = Aj * (1.0 + 0x1p-20);
      const double margin = 0x1p-24 * fabs(totj) + 0x1p-40 * fabs(Pj);
      const double lo = (Pj - A1) - margin, hi = (Pj + A1) + margin;
      const long long blo = __double_as_longlong(lo), bhi = __double_as_longlong(hi);
      const int elo = (int)((blo >> 52) & 0x7ff), ehi = (int)((bhi >> 52) & 0x7ff);
      const bool same = ((blo ^ bhi) >= 0) && elo == ehi && elo >= 1 && elo <= 2046;
      double B = 0.0;
      if (Aj == 0.0) B = -0.0;                                                          // only zeros: s + (+-0)
      else if (same) B = __longlong_as_double((blo & (long long)0xfff0000000000000ULL) | 0x0008000000000000LL);
      if (c0 + lane >= p.K + p.M) B = 0.0;
      prox[chain * 32 + lane] = B;
    }
    __syncthreads();

    // ---- 3b. proxy chains from shared memory: warp = (chain, parity), lane = column --------------------------------
    {
      const int chain = warp >> 1, parity = warp & 1;
      const double B = prox[chain * 32 + lane];
      const long long bb = __double_as_longlong(B);
      const bool slow = bb == 0;
      const bool ident = bb == (long long)0x8000000000000000ULL;
      const double start = (parity && !ident) ? __longlong_as_double(bb | 1) : B;
      double acc = start;
      const double* zt = tile + lane;
      if (!__all_sync(0xffffffffu, slow)) {
        if (cnt == SCAN_L) {
#pragma unroll 2
          for (int r = 0; r < SCAN_L; r += 8) {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const double z = zt[(size_t)(r + u) * SCAN_COLS];
              t[u] = __dmul_rn(z, swt[r + u]);
              if (chain) t[u] = __dmul_rn(t[u], z);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, t[u]);
          }
        } else {
          for (int r = 0; r < cnt; ++r) {
            const double z = zt[(size_t)r * SCAN_COLS];
            double t = __dmul_rn(z, swt[r]);
            if (chain) t = __dmul_rn(t, z);
            acc = __dadd_rn(acc, t);
          }
        }
      }
      const int64_t c = c0 + lane;
      if (c < p.K + p.M) {
        double d = ident ? acc : __dsub_rn(acc, start);
        if (slow) d = __longlong_as_double((long long)SCAN_SLOW_MARK);
        scan_plane(sp, f, chain, parity, c)[seg] = d;
      }
    }
    cur = nxt;
    buf ^= 1;
  }
}

// Ascending list of the slow segments of every (fold, chain, column) from the marks k_scan_fused left in the d0 plane;
// groups that are mostly slow (or too short to be worth it) are handed back to k_moments_pipe.  One warp per column chain.
__global__ void __launch_bounds__(128) k_scan_lists(ScanParams sp, int mine) {
  const MomentParams<double>& p = sp.p;
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);        // (own group, column, chain)
  const int64_t f = blockIdx.y;
  if (wid >= (int64_t)mine * SCAN_COLS * 2) return;
  const int chain = (int)(wid & 1), col = (int)((wid >> 1) % SCAN_COLS), gi = (int)(wid / (2 * SCAN_COLS));
  const int g = p.grp0 + gi * p.grp_stride;
  const int64_t c = (int64_t)g * SCAN_COLS + col;
  const int64_t beg = p.offsets ? p.offsets[p.fold0 + f] : 0;
  const int64_t n = p.offsets ? p.offsets[p.fold0 + f + 1] - beg : p.N;
  const int64_t nseg = (n + sp.L - 1) / sp.L;
  int* okp = sp.ok + f * sp.groups_total + g;
  if (nseg < 4 && lane == 0) atomicAnd(okp, 0);
  if (c >= p.K + p.M) return;
  const double* d0 = scan_plane(sp, f, chain, 0, c);
  int* list = sp.slow_list + ((size_t)(f * 2 + chain) * p.ld + c) * sp.slow_cap;
  int nslow = 0;
  for (int64_t J = 0; J < nseg; J += 32) {
    const int64_t j = J + lane;
    const bool slow = j < nseg && (unsigned long long)__double_as_longlong(d0[j]) == SCAN_SLOW_MARK;
    const unsigned m = __ballot_sync(0xffffffffu, slow);
    if (slow) {
      const int pos = nslow + __popc(m & ((1u << lane) - 1u));
      if (pos < sp.slow_cap) list[pos] = (int)j;
    }
    nslow += __popc(m);
  }
  if (lane == 0) {
    sp.slow_cnt[(size_t)(f * 2 + chain) * p.ld + c] = nslow;
    if (nslow > nseg / 4 || nslow > sp.slow_cap) atomicAnd(okp, 0);
  }
}

}  // namespace cvmx
