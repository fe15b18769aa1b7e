// float32 Gram kernel on the 5th-generation tensor cores: tcgen05.mma kind::tf32, accumulators in tensor memory.
//
//   G[i][j] = sum_r rn(w_r x_ri) z_rj          (cvmatrix/cvmatrix.py:1215-1217 in fit, :1001 per fold; float32 model)
//
// tcgen05 has no float64 kind, so the float64 headline stays on DMMA (kernels_gram.cuh); a float32 model is where the
// UMMA pipe applies.  float32 accuracy is kept with the 3xTF32 split: a = a_hi + a_lo, z = z_hi + z_lo with the hi parts
// rounded to TF32 and the lo parts the exact float32 remainders, a z ~= a_lo z_hi + a_hi z_lo + a_hi z_hi (the dropped
// lo lo term is 2^-22 of the product).  One CTA = one 128 x 128 output tile of one (fold, row-split) unit, as in k_gram:
//
//   warps 8-10  producers   TMA bulk copies gather TBK = 32 rows x {A block, B block} (512 bytes per row piece) into a
//                           staging ring, row weights by cp.async; full / empty mbarriers                       (UBLKCP)
//   warps 0-7   converters  staging -> UMMA operands: rn(w x) for A, the hi / lo split, and the TRANSPOSE the tensor
//                           core wants - operands are K-major (K = the data-row index is contiguous: 32 TF32 = one
//                           128-byte row) in the canonical 128-byte-swizzled layout, written with 16-byte stores;
//                           fence.proxy.async hands them to the tensor core's async proxy
//   warp 11     issuer      one elected thread issues 3 x 4 tcgen05.mma (M = N = 128, K = 8) per stage into a TMEM
//                           accumulator (128 lanes x 128 float32 columns); tcgen05.commit releases the operand set
//                           and, every TFLUSH stages, hands the accumulator to the converters                   (UTCHMMA)
//   warps 0-7   flush       tcgen05.ld their 64 accumulator values (lane = output row) and add them into float64
//                           REGISTER accumulators; two TMEM accumulators alternate, so the flush of window w overlaps
//                           the MMAs of window w + 1                                                         (LDTM)
//
// The tensor core accumulates in float32 and TRUNCATES after every MMA (a bias of about -2e-8 of the accumulator per MMA;
// measured: 48 MMAs per window into one accumulator gave -1.1e-6), hence (a) short windows for the main term a_hi z_hi -
// 128 rows = 16 MMAs per window, ~3.5e-7 of a window's partial sum, independent of the fold length - and (b) a SEPARATE
// accumulator for the two small terms (2^-11 of the main one: their truncation is negligible even over a whole unit), read once
// at the end.
// The float64 accumulators leave through the same partial-buffer / epilogue code as every other variant
// (fragment map 1: thread tid owns output row 32 (warp % 4) + lane and columns 64 (warp / 4) + 0..63).
#pragma once
#include "kernels_gram.cuh"

namespace cvmx {

constexpr int TBK = 32;          // data rows per stage = K extent of one swizzle atom (32 TF32 = 128 bytes)
constexpr int TSTAGES = 3;       // staging ring depth (one producer warp per slot)
constexpr int TOPS = 2;          // operand buffer sets
constexpr int TFLUSH = 4;        // stages per accumulator window (128 rows)
constexpr int TOP_BYTES = GB * 128;                       // one operand buffer: 128 rows x 128 bytes = 16 KB
constexpr int TSTAGE_BYTES = 2 * TBK * GB * 4;            // A rows + B rows of a stage: 32 KB
constexpr int TMEM_COLS = 512;                            // float32 accumulators of 128 columns: main term x 2 (windows alternate), small terms x 1
// (no setmaxnreg split here: the converters' 64 float64 accumulators + conversion temporaries fit the 168 registers the
// CTA is launched with, and the issuer's descriptor arithmetic does not fit the 64 a trimmed producer warpgroup would keep)

constexpr size_t gram_tc_smem_bytes() {
  return 1024 /* alignment slack */ + (size_t)TOPS * 4 * TOP_BYTES + (size_t)TSTAGES * TSTAGE_BYTES + TSTAGES * TBK * 4 + 16 * sizeof(uint64_t) + 16;
}

// ---- tcgen05 wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, M = N = 128, K = 8 (TF32)
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// Shared-memory operand descriptor: K-major, 128-byte swizzle.  A row of the operand (one output row / column, 32 TF32
// along K) is 128 bytes; 8 rows form one 1024-byte swizzle atom; atoms follow each other along M / N (stride byte offset
// 1024).  Bits: [0,14) address >> 4, [16,30) leading byte offset >> 4 (unused for K-major swizzled layouts), [32,46)
// stride byte offset >> 4, [46,48) descriptor version 1 (sm_100), [61,64) layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t tc_smem_desc(const void* p) {
  const uint64_t addr = (uint64_t)(smem_u32(p) & 0x3ffff) >> 4;
  return addr | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor, kind::tf32: [4,6) D format 1 = F32, [7,10) A format 2 = TF32, [10,13) B format 2 = TF32,
// [15] A major 0 = K, [16] B major 0 = K, [17,23) N >> 3, [24,29) M >> 4.
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GB >> 3) << 17) | ((uint32_t)(GB >> 4) << 24);

// bounded wait: a protocol error must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, unsigned parity) {
  for (unsigned long long spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1ull << 26)) __trap();
}

__global__ void __launch_bounds__(GLAUNCH, 1) k_gram_tc(const GramParams<float> p) {
  extern __shared__ unsigned char tc_smem_unaligned[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_unaligned) + 1023) & ~(uintptr_t)1023);
  unsigned char* s_ops = smem;                                             // [TOPS][A_hi, A_lo, B_hi, B_lo][16 KB], 1024-aligned
  float* s_stage = reinterpret_cast<float*>(smem + (size_t)TOPS * 4 * TOP_BYTES);   // [TSTAGES][A: TBK x 128 | B: TBK x 128]
  float* s_w = s_stage + (size_t)TSTAGES * TSTAGE_BYTES / 4;               // [TSTAGES][TBK]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + TSTAGES * TBK);
  uint64_t *full_s = bars, *empty_s = bars + 3, *op_full = bars + 6, *op_free = bars + 8, *acc_full = bars + 10, *acc_free = bars + 12;
  uint64_t* done = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x % p.ntiles;
  const GramUnit unit = p.units[blockIdx.x / p.ntiles];
  const int2 tl = p.tiles[tile];
  const int bi = tl.x, bj = tl.y;
  const bool diag = bi == bj;
  const int64_t ld = p.ld;
  const int64_t nrows = unit.row_end - unit.row_begin;
  const int64_t nk = (nrows + TBK - 1) / TBK;
  const int64_t nwin = (nk + TFLUSH - 1) / TFLUSH;

  if (tid == 0) {
    for (int s = 0; s < TSTAGES; ++s) { mbar_init(full_s + s, 33); mbar_init(empty_s + s, GTHREADS / 32); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(op_full + s, GTHREADS / 32); mbar_init(op_free + s, 1);
      mbar_init(acc_full + s, 1); mbar_init(acc_free + s, GTHREADS / 32);
    }
    mbar_init(done, GTHREADS / 32);
    mbar_fence_init();
  }
  if (warp == 11) tc_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= GTHREADS / 32) {
    if (warp < 11) {
      // ---------------- producers: lane = row of the stage.  Warp 8 + s owns staging slot s (a slot must have ONE producer:
      // a parity wait can only tell "the previous phase" from "this one", so a third warp sharing two slots would run
      // two phases ahead and overwrite a stage that has not been consumed). ----------------
      const int64_t acol = (int64_t)bi * GB, bcol = (int64_t)bj * GB;
      const unsigned a_bytes = (unsigned)(min((int64_t)GB, ld - acol) * 4), b_bytes = diag ? 0u : (unsigned)(min((int64_t)GB, ld - bcol) * 4);
      for (int64_t kt = warp - GTHREADS / 32; kt < nk && warp - GTHREADS / 32 < TSTAGES; kt += TSTAGES) {
        const int slot = (int)(kt % TSTAGES);
        const unsigned round = (unsigned)(kt / TSTAGES);
        const int64_t pos = unit.row_begin + kt * TBK + lane;
        const int64_t grow = pos < unit.row_end ? (p.indices ? p.indices[pos] : pos) : -1;
        if (round > 0) mbar_wait_bounded(empty_s + slot, (round & 1) ^ 1);
        const int rows = (int)min((int64_t)TBK, nrows - kt * TBK);
        if (lane == 0) mbar_arrive_expect_tx(full_s + slot, (unsigned)rows * (a_bytes + b_bytes));
        float* stA = s_stage + (size_t)slot * (TSTAGE_BYTES / 4);
        if (grow >= 0) {
          bulk_g2s(stA + lane * GB, p.Z + grow * ld + acol, a_bytes, full_s + slot);
          if (!diag) bulk_g2s(stA + (TBK + lane) * GB, p.Z + grow * ld + bcol, b_bytes, full_s + slot);
        }
        cp_async4(s_w + slot * TBK + lane, p.w + (grow >= 0 ? grow : 0), grow >= 0 ? 4 : 0);
        cp_async_mbar_arrive_noinc(full_s + slot);
      }
      cp_async_wait<0>();
    } else {
      // ---------------- MMA issuer (warp 11; lane 0 issues) --------------------------------------------------------------
      for (int64_t kt = 0; kt < nk; ++kt) {
        const int ob = (int)(kt % TOPS);
        const int64_t win = kt / TFLUSH;
        const int ab = (int)(win & 1);
        const bool first = (kt % TFLUSH) == 0;
        if (first && win >= 2) mbar_wait_bounded(acc_free + ab, (unsigned)((win / 2 - 1) & 1));
        mbar_wait_bounded(op_full + ob, (unsigned)((kt / TOPS) & 1));
        tc_fence_after();
        if (lane == 0) {
          const unsigned char* ops = s_ops + (size_t)ob * 4 * TOP_BYTES;
          const uint64_t a_hi = tc_smem_desc(ops), a_lo = tc_smem_desc(ops + TOP_BYTES);
          const uint64_t b_hi = tc_smem_desc(ops + 2 * TOP_BYTES), b_lo = tc_smem_desc(ops + 3 * TOP_BYTES);
          const uint32_t d = tmem_base + (uint32_t)ab * GB, d_small = tmem_base + 2u * GB;
#pragma unroll
          for (int ks = 0; ks < TBK / 8; ++ks) {                       // 32 bytes along K per step: address field + 2
            const uint64_t o = (uint64_t)(ks * 2);
            tc_mma_tf32(d_small, a_lo + o, b_hi + o, TC_IDESC, (kt == 0 && ks == 0) ? 0u : 1u);
            tc_mma_tf32(d_small, a_hi + o, b_lo + o, TC_IDESC, 1u);
            tc_mma_tf32(d, a_hi + o, b_hi + o, TC_IDESC, (first && ks == 0) ? 0u : 1u);
          }
          tc_commit(op_free + ob);                                    // operand set reusable when these MMAs have read it
          if ((kt % TFLUSH) == TFLUSH - 1 || kt == nk - 1) tc_commit(acc_full + ab);
        }
        __syncwarp();
      }
      // the TMEM allocation is released once every converter has read its last window
      mbar_wait_bounded(done, 0);
      tc_fence_after();
      tc_dealloc(tmem_base, TMEM_COLS);
    }
    return;
  }

  // ---------------- converters + flush (warps 0-7) ------------------------------------------------------------------------
  double acc[8][4][2];
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;

  // this thread's 64 values of the accumulator at TMEM column `col0` are added into the float64 registers
  auto drain = [&](uint32_t col0) {
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + col0 + (uint32_t)(64 * (warp >> 2));
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t r[16];
      tc_ld16(taddr + ch * 16, r);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int lin = ch * 16 + j;
        acc[lin >> 3][(lin >> 1) & 3][lin & 1] += (double)__uint_as_float(r[j]);
      }
    }
  };
  auto flush = [&](int64_t win) {
    const int ab = (int)(win & 1);
    mbar_wait_bounded(acc_full + ab, (unsigned)((win / 2) & 1));
    tc_fence_after();
    drain((uint32_t)(ab * GB));
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(acc_free + ab);
  };

  // thread -> operand row i = tid % 128 (output row of A / column of B) and K chunks rg = tid / 128 + 2 j (4 data rows each)
  int64_t flushed = 0;
  const int ci = tid & 127, rg0 = tid >> 7;
  const uint32_t op_row = (uint32_t)((ci >> 3) * 1024 + (ci & 7) * 128);
#pragma unroll 1
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int slot = (int)(kt % TSTAGES), ob = (int)(kt % TOPS);
    const int rows = (int)min((int64_t)TBK, nrows - kt * TBK);
    mbar_wait_bounded(full_s + slot, (unsigned)((kt / TSTAGES) & 1));
    if (kt >= TOPS) mbar_wait_bounded(op_free + ob, (unsigned)((kt / TOPS - 1) & 1));
    const float* stA = s_stage + (size_t)slot * (TSTAGE_BYTES / 4);
    const float* stB = diag ? stA : stA + TBK * GB;
    const float* wv = s_w + slot * TBK;
    unsigned char* ops = s_ops + (size_t)ob * 4 * TOP_BYTES;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rg = rg0 + 2 * j;
      uint32_t ah[4], al[4], bh[4], bl[4];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int r = 4 * rg + rr;
        const bool ok = r < rows;                                      // rows past the end of the unit were never copied
        const float x = ok ? stA[r * GB + ci] : 0.f, z = ok ? stB[r * GB + ci] : 0.f, w = ok ? wv[r] : 0.f;
        split_tf32(__fmul_rn(x, w), ah[rr], al[rr]);                   // rn(w x) in float32 == WX of the reference
        split_tf32(z, bh[rr], bl[rr]);
      }
      const uint32_t off = op_row + (uint32_t)(((rg ^ (ci & 7)) & 7) * 16);   // 128-byte swizzle: chunk ^= row % 8
      *reinterpret_cast<uint4*>(ops + off) = make_uint4(ah[0], ah[1], ah[2], ah[3]);
      *reinterpret_cast<uint4*>(ops + TOP_BYTES + off) = make_uint4(al[0], al[1], al[2], al[3]);
      *reinterpret_cast<uint4*>(ops + 2 * TOP_BYTES + off) = make_uint4(bh[0], bh[1], bh[2], bh[3]);
      *reinterpret_cast<uint4*>(ops + 3 * TOP_BYTES + off) = make_uint4(bl[0], bl[1], bl[2], bl[3]);
    }
    fence_proxy_async();                                               // generic-proxy stores -> the tensor core's async proxy
    __syncwarp();
    if (lane == 0) { mbar_arrive(op_full + ob); mbar_arrive(empty_s + slot); }
    // flush finished windows one stage late: the issuer has then had a whole conversion to complete their last MMAs
    while (kt >= 1 && flushed <= (kt - 1) / TFLUSH - 1) flush(flushed++);
  }
  while (flushed < nwin) flush(flushed++);
  if (nk > 0) { drain(2u * GB); tc_fence_before(); }   // the small terms (complete: the last acc_full commit covers every MMA)
  __syncwarp();
  if (lane == 0) mbar_arrive(done);
  compute_barrier();   // every stage converted and every window read: staging / operand memory can be reused as the epilogue tile

  if (unit.nsplit == 1 && !p.force_partials) {
    gram_epilogue<float, 4>(acc, reinterpret_cast<float*>(smem), p.epi, unit.fold, bi, bj, 1);
  } else {
    double* dst = p.partials + ((size_t)(unit.part_base + unit.split) * p.ntiles + tile) * (size_t)(GACC * GTHREADS);
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dst[((t * 4 + u) * 2 + 0) * GTHREADS + tid] = acc[t][u][0];
        dst[((t * 4 + u) * 2 + 1) * GTHREADS + tid] = acc[t][u][1];
      }
  }
}

}  // namespace cvmx
