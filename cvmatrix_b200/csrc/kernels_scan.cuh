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
constexpr int SCAN_L = 256;           // rows per segment (multiple of 32; compile-time in pass 4)

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
  // row-slab mode, decoupled chain (cvmx_slab_scan_*): the rows before this slab live on other ranks.
  //   start [folds][chain][2: prefix, magnitudes][ld]: APPROXIMATE running sum and sum of magnitudes at the slab's first row
  //         (any summation order - pass 2 only classifies segments with them, inside its usual error margin)
  //   carry [folds][chain][ld]: the EXACT chains after the previous slab - pass 4 continues them
  const double* start = nullptr;
  const double* carry = nullptr;
};

// element (segment s, column c) of plane (fold f, chain, slot)
__device__ __forceinline__ double* scan_plane(const ScanParams& sp, int64_t f, int chain, int slot, int64_t c) {
  return sp.seg + ((((size_t)f * 2 + chain) * 2 + slot) * (size_t)sp.p.ld + (size_t)c) * (size_t)sp.max_segs;
}

// spec planes: [folds][chain][3: guessed proxy, d0, d1][ld][max_segs]
__device__ __forceinline__ double* scan_spec_plane(const ScanParams& sp, double* spec, int64_t f, int chain, int which, int64_t c) {
  return spec + ((((size_t)f * 2 + chain) * 3 + which) * (size_t)sp.p.ld + (size_t)c) * (size_t)sp.max_segs;
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

// ---- pass 1b (row slabs) ----------------------------------------------------------------------------------------------
// Slab totals of the pass-1 planes: out[fold][chain][0] = sum of S, [1] = sum of A over the slab's segments, per column.
// The ranks exchange these (tiny) rows; the totals of the slabs before a rank are its `start`.  One warp per column.
__global__ void __launch_bounds__(128) k_scan_slab_totals(ScanParams sp, double* __restrict__ out) {
  const MomentParams<double>& p = sp.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t f = blockIdx.y;
  const int64_t c = (int64_t)blockIdx.x * 4 + warp;
  if (c >= p.ld) return;
  const int64_t beg = p.offsets ? p.offsets[p.fold0 + f] : 0;
  const int64_t n = p.offsets ? p.offsets[p.fold0 + f + 1] - beg : p.N;
  const int64_t nseg = (n + sp.L - 1) / sp.L;
  for (int chain = 0; chain < 2; ++chain) {
    const double* dS = scan_plane(sp, f, chain, 0, c);
    const double* dA = scan_plane(sp, f, chain, 1, c);
    double S = 0.0, A = 0.0;
    if (c < p.K + p.M)
      for (int64_t s = lane; s < nseg; s += 32) { S += dS[s]; A += dA[s]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) { S += __shfl_xor_sync(0xffffffffu, S, o); A += __shfl_xor_sync(0xffffffffu, A, o); }
    if (lane == 0) {
      out[(((size_t)f * 2 + chain) * 2 + 0) * p.ld + c] = S;
      out[(((size_t)f * 2 + chain) * 2 + 1) * p.ld + c] = A;
    }
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
// SPEC: the proxy chains already ran from guessed proxies (k_scan_spec); a segment is fast iff the true proxy start
// equals the guess, and its increments are copied from the spec planes (pass 3 is skipped).
template <bool SPEC>
__global__ void __launch_bounds__(SCAN_PREFIX_THREADS) k_scan_prefix(ScanParams sp, const double* __restrict__ spec) {
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
      if (sp.start) {
        P = sp.start[(((size_t)f * 2 + chain) * 2 + 0) * p.ld + c];
        tot = sp.start[(((size_t)f * 2 + chain) * 2 + 1) * p.ld + c];
      } else if (p.accumulate) P = chain == 0 ? p.sum_z[c] : p.sumsq_z[c];
      double* dS = scan_plane(sp, f, chain, 0, c);
      double* dA = scan_plane(sp, f, chain, 1, c);
      const double* gB = SPEC ? scan_spec_plane(sp, const_cast<double*>(spec), f, chain, 0, c) : nullptr;
      const double* gd0 = SPEC ? scan_spec_plane(sp, const_cast<double*>(spec), f, chain, 1, c) : nullptr;
      const double* gd1 = SPEC ? scan_spec_plane(sp, const_cast<double*>(spec), f, chain, 2, c) : nullptr;
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
            if (SPEC) {
              // accept the speculative increments only where the guess IS the true proxy start (same sign and binade;
              // -0 = identity segment on both sides); everything else is added row by row in pass 4
              const bool hit = __double_as_longlong(B) != 0 && __double_as_longlong(B) == __double_as_longlong(gB[j0 + e]);
              dS[j0 + e] = hit ? gd0[j0 + e] : 0.0;
              dA[j0 + e] = hit ? gd1[j0 + e] : 0.0;
              if (!hit) slowbits |= 1u << e;
            } else {
              dS[j0 + e] = B;
              if (__double_as_longlong(B) == 0) slowbits |= 1u << e;
            }
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
    if (sp.carry) acc = sp.carry[((size_t)f * 2 + chain) * p.ld + c];
    else if (p.accumulate) acc = chain == 0 ? p.sum_z[c] : p.sumsq_z[c];
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

// ---- passes 1 and 3 in ONE read of the rows: speculative proxies ---------------------------------------------------
// The four-pass form reads every row twice (pass 1 and pass 3), because the proxy start of a segment needs the prefix
// over all earlier segments.  For fold statistics the data set's own totals are already known (fit), and they predict
// that prefix well enough to GUESS the binade: after k of the N rows the running sum of column c is about
// sum_z[c] * k / N.  k_scan_spec computes the segment sums S, A, Q of pass 1 AND runs the four proxy chains of pass 3
// from the guessed binade in the same sweep; k_scan_prefix<true> then derives the true proxy start exactly as before
// and accepts the speculative increments only where its binade (and sign) EQUALS the guess - bit for bit - and marks
// the segment slow otherwise.  A wrong guess therefore costs one slow segment, never a wrong bit; rows that are not
// exchangeable (sorted data) simply fall back to slow segments / the chain kernel through the usual group flag.
// (Measured alternatives, both slower than the four passes: a single-pass decoupled look-back scan with the segment
// staged in shared memory, 249 us vs 271 us for the 8-way shard of cfg 2 - look-back couples every task to its 32
// predecessors and the GPU falls into lock-step load / compute phases - and its persistent double-buffered form, 427 us.)
__device__ __forceinline__ double scan_guess_proxy(double est) {
  const long long b = __double_as_longlong(est);
  const int e = (int)((b >> 52) & 0x7ff);
  if (e < 1 || e > 2046) return 0.0;                       // zero, denormal, inf, nan: no guess
  return __longlong_as_double((b & (long long)0xfff0000000000000ULL) | 0x0008000000000000LL);
}

__global__ void __launch_bounds__(32 * SCAN_WARPS) k_scan_spec(ScanParams sp, double* __restrict__ spec) {
  const MomentParams<double>& p = sp.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t f = blockIdx.z;
  const int64_t c = (int64_t)(p.grp0 + blockIdx.x * p.grp_stride) * SCAN_COLS + lane;
  const int64_t beg = p.offsets[p.fold0 + f];
  const int64_t n = p.offsets[p.fold0 + f + 1] - beg;
  const int64_t* idx = p.indices + beg;
  if (blockIdx.y == 0 && threadIdx.x == 0) sp.ok[f * sp.groups_total + p.grp0 + blockIdx.x * p.grp_stride] = 1;   // pass 2 clears it
  const bool live = c < p.K + p.M;
  const double tot_s = live ? p.sum_z[c] : 0.0, tot_q = live ? p.sumsq_z[c] : 0.0;
  const double invN = 1.0 / (double)p.N;
  for (int64_t s = (int64_t)blockIdx.y * SCAN_WARPS + warp; s * sp.L < n; s += (int64_t)gridDim.y * SCAN_WARPS) {
    const int64_t r0 = s * sp.L;
    const int cnt = (int)min((int64_t)sp.L, n - r0);
    const double frac = ((double)r0 + 0.5 * (double)cnt) * invN;
    const double Bs = scan_guess_proxy(tot_s * frac), Bq = scan_guess_proxy(tot_q * frac);
    const long long bs = __double_as_longlong(Bs), bq = __double_as_longlong(Bq);
    const double Bs1 = bs ? __longlong_as_double(bs | 1) : 0.0, Bq1 = bq ? __longlong_as_double(bq | 1) : 0.0;
    double S = 0.0, A = 0.0, Q = 0.0, c0 = Bs, c1 = Bs1, e0 = Bq, e1 = Bq1;
    unsigned negq = 0, all_neg_t = 0x80000000u, all_neg_q = 0x80000000u;
    scan_rows(p, idx, r0, cnt, lane, p.Z + c, [&](double t, double q) {
      S = __dadd_rn(S, t);
      A = __dadd_rn(A, fabs(t));
      Q = __dadd_rn(Q, fabs(q));
      c0 = __dadd_rn(c0, t);
      c1 = __dadd_rn(c1, t);
      e0 = __dadd_rn(e0, q);
      e1 = __dadd_rn(e1, q);
      const unsigned st = (unsigned)__double2hiint(t), sq = (unsigned)__double2hiint(q);
      negq |= sq & 0x80000000u;
      all_neg_t &= st;
      all_neg_q &= sq;
    });
    scan_plane(sp, f, 0, 0, c)[s] = S;
    scan_plane(sp, f, 0, 1, c)[s] = A;
    scan_plane(sp, f, 1, 0, c)[s] = Q;
    scan_plane(sp, f, 1, 1, c)[s] = negq ? __longlong_as_double(0x7ff8000000000000LL) : Q;
    // a segment that holds only zeros adds (+-0) to the running sum: its increment is -0 iff every term is -0
    const bool zs = A == 0.0, zq = Q == 0.0 && !negq;
    const double ids = (all_neg_t & 0x80000000u) ? -0.0 : 0.0, idq = (all_neg_q & 0x80000000u) ? -0.0 : 0.0;
    scan_spec_plane(sp, spec, f, 0, 0, c)[s] = zs ? -0.0 : Bs;
    scan_spec_plane(sp, spec, f, 0, 1, c)[s] = zs ? ids : __dsub_rn(c0, Bs);
    scan_spec_plane(sp, spec, f, 0, 2, c)[s] = zs ? ids : __dsub_rn(c1, Bs1);
    scan_spec_plane(sp, spec, f, 1, 0, c)[s] = zq ? -0.0 : Bq;
    scan_spec_plane(sp, spec, f, 1, 1, c)[s] = zq ? idq : __dsub_rn(e0, Bq);
    scan_spec_plane(sp, spec, f, 1, 2, c)[s] = zq ? idq : __dsub_rn(e1, Bq1);
  }
}

}  // namespace cvmx
