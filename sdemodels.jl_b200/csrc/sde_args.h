/* sde_args.h — plain-old-data kernel argument blocks shared by the host C-ABI layer and the device
 * kernels (AOT-compiled by nvcc and JIT-compiled by NVRTC from the same header).  C types only. */
#ifndef SDE_ARGS_H
#define SDE_ARGS_H

#if defined(__CUDACC_RTC__)
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
#else
#include <stdint.h>
#endif

#define SDE_MAXD 4        /* state dimension  (reference type parameter D of AbstractSDE{D,M}) */
#define SDE_MAXM 4        /* noise dimension  (M) */
#define SDE_MAXP 16       /* model parameters (struct fields of the generated model type) */
#define SDE_MAXC 32       /* per-thread prepared constants (raw parameters + hoisted sub-expressions) */
#define SDE_MAXF 4        /* payoff features reduced per path */
#ifndef SDE_BLOCK
#define SDE_BLOCK 128     /* threads per CTA for every path kernel */
#endif
#ifndef SDE_PPT
#define SDE_PPT 4         /* paths per thread in the terminal-value kernel (instruction-level parallelism); a CTA works on chunks
                             of SDE_BLOCK * SDE_PPT = 512 consecutive paths, whose summation tree defines the payoff partials */
#endif
#ifndef SDE_PPT_V2
#define SDE_PPT_V2 4      /* the same for the FP64 kernel of normal stream v2 as its own switch: with 2, the same 512-path chunks
                             and summation tree run in SDE_BLOCK * SDE_PPT / SDE_PPT_V2 = 256-thread CTAs (sample_body's staged
                             form): 118 registers / 16 warps per SM instead of 204 / 8.  Measured (profiles/r02y_ab_loop.txt):
                             295.2 G path-steps/s with 4, 292.4 with 2 */
#endif
#define SDE_SAMPLE_THREADS(v2_f64) ((v2_f64) ? SDE_BLOCK * SDE_PPT / SDE_PPT_V2 : SDE_BLOCK)
#define SDE_TAB_DOUBLES 2304 /* FP64 transform tables in shared memory: 256 (log) + 2048 (rotation) */
/* SDE_V2_FINE = 1 (default): normal stream v2 evaluates its transform on finer tables — 512 log cells (|r| <= 2^-10: three
 * FMAs instead of four) and 2048 rotation cells (|phi| <= pi/2048: one-term sine residual) — 40 KB of shared memory per
 * CTA for 2 FP64 instructions less per pair.  The SPECIFICATION (u1, angle -> normals) does not depend on it. */
#ifndef SDE_V2_FINE
#define SDE_V2_FINE 1
#endif
#if SDE_V2_FINE
#define SDE_V2_LOGN 512
#define SDE_V2_ROTN 2048
#else
#define SDE_V2_LOGN 128
#define SDE_V2_ROTN 1024
#endif
#define SDE_TAB2_DOUBLES (2 * SDE_V2_LOGN + 2 * SDE_V2_ROTN)
#define SDE_SIMREF_PW 1   /* paths per lane in the reference-layout trajectory kernels (2 measured slower: 128 registers) */
#define SDE_SOA_VEC 4  /* consecutive paths per thread in the [time][D][path] trajectory kernel (8 measured no better in FP32) */
#ifndef SDE_SIMW_ROWBYTES
#define SDE_SIMW_ROWBYTES 128 /* contiguous bytes per path and tile in the Wiener-driven kernel */
#define SDE_SIMW_MINBLOCKS 5
#endif
#define SDE_NBINS 68      /* 32-bit limbs of the exact (Kulisch-style) accumulator, held in int64 */
#define SDE_SHARD_ALIGN 1024 /* shards start on multiples of this many paths => chunking is G-invariant */

/* payoff kinds (new functionality on top of the reference) */
#define SDE_PAYOFF_MOMENTS 0   /* features f_d = x_d, d < D */
#define SDE_PAYOFF_CALL 1      /* disc * max(x_k - K, 0) */
#define SDE_PAYOFF_PUT 2       /* disc * max(K - x_k, 0) */
#define SDE_PAYOFF_COMPONENT 3 /* x_k */

/* full-path layouts */
#define SDE_LAYOUT_REFERENCE 0 /* [path][time][D]  == Julia Array{SVector{D}}(nsteps+1, npaths)   */
#define SDE_LAYOUT_SOA 1       /* [time][D][path]  coalesced device layout */

/* a CUtensorMap (TMA descriptor), opaque to device code; built by the host with cuTensorMapEncodeTiled */
typedef struct {
#if defined(__cplusplus)
  alignas(64)
#endif
  unsigned long long v[16];
} sde_tmap;

/* Geometry of the TMA trajectory tiles (device: sde_device.cuh, host: the tensor maps in sde_api.cu).  A tensor
 * map's strides must be multiples of 16 bytes but the path pitch nrows*D*sizeof(T) usually is not, so g = 1, 2 or 4
 * consecutive paths form one row of the map; the box of sub-path j starts at the first time row h_j whose
 * address is 16-byte aligned (TMA box starts must be). */
#if defined(__CUDACC__)
#define SDE_ARGS_HD __host__ __device__ static inline
#else
#define SDE_ARGS_HD static inline
#endif
SDE_ARGS_HD int sde_tma_group(int64_t pitch_bytes, int align = 16) {
  int g = 1;
  while ((g * pitch_bytes) % align) g *= 2;
  return g; /* the pitch is a multiple of sizeof(T) >= 4, hence g <= align / 4 */
}
SDE_ARGS_HD int sde_tma_head(int j, int64_t pitch_bytes, int rowb, int align = 16) {
  int h = 0;
  while (((int64_t)j * pitch_bytes + (int64_t)h * rowb) % align) h++;
  return h;
}
/* alignment K2u gives the row segments of its tiles: 16 bytes is what a tensor store needs; the 64-byte tiles align to the
 * 32-byte DRAM sector (g <= 8 paths per map row, h_j <= 7), because a 64-byte segment that starts in the middle of a sector
 * leaves two of its three sectors partial (measured on 513-row FP32 paths: 2.1 TB/s against 3.66 TB/s on 512-row paths;
 * 3.40 TB/s once aligned).  SDE_SIMRT2_ALIGN_F64 / _F32 below. */
/* bytes per path and tile of K2u (tile = 32 paths x this; 128: 128-byte swizzle, 64: 64-byte swizzle).  The FP32 kernel is bound
 * by instruction issue, not by its stores (profiles/r02m: 6.91 ms with tensor stores, 6.62 ms with none at all, 6.79 ms with
 * stores that stay in L2; 12 warps per SM at 16 KB of ring per warp): half-size tiles double the resident warps. */
#ifndef SDE_SIMRT2_ROWB_F64
#define SDE_SIMRT2_ROWB_F64 128
#endif
#ifndef SDE_SIMRT2_ROWB_F32
#define SDE_SIMRT2_ROWB_F32 64
#endif
#ifndef SDE_SIMRT2_ALIGN_F64
#define SDE_SIMRT2_ALIGN_F64 32 /* measured on C3: 3.63 -> 3.84 TB/s, wiener! 3.70 -> 4.13 (profiles/r02q_alignment_variants.txt) */
#endif
#ifndef SDE_SIMRT2_ALIGN_F32
#define SDE_SIMRT2_ALIGN_F32 (SDE_SIMRT2_ROWB_F32 == 64 ? 32 : 16)
#endif
#ifndef SDE_SIMRT2_PREFETCH
#define SDE_SIMRT2_PREFETCH 0 /* K2u: generate the next batch's normals alongside the current batch's steps.  Measured on C3
                                 (profiles/r02u_k2u_prefetch_ab.txt): FP32 6.00 vs 5.99 ms (ptxas keeps the two chains apart anyway),
                                 FP64 14.9 vs 10.6 ms (eight more live doubles at the 128-register cap): off */
#endif
/* K2u ring depth (tiles in flight per warp).  Measured on C3 after the sector-aligned tiles (profiles/r02zz_c3_ring_depth_*.txt):
 * FP32 (2 KB tiles) 2 / 3 / 4 / 5 / 6 / 8 stages: 3.31 / 3.32 / 3.43 / 3.51 / 3.67 / 3.39 TB/s — 6 stages = 48 KB per CTA, 16
 * warps per SM; FP64 (4 KB tiles) 2 / 3 / 4 / 6 stages: 3.89 / 3.24 / 3.18 / 2.33 — the FP64 form needs its warps more than
 * stores in flight */
#ifndef SDE_SIMRT2_STAGES_F64
#define SDE_SIMRT2_STAGES_F64 2
#endif
#ifndef SDE_SIMRT2_STAGES_F32
#define SDE_SIMRT2_STAGES_F32 6
#endif
/* K2u: how a finished tile leaves shared memory — 0: one tensor store by lane 0 (asynchronous, STAGES deep ring),
 * 1: 8 x (LDS.128 + STG.128) by the whole warp (synchronous, one stage) */
#ifndef SDE_SIMRT2_MANUAL_F64
#define SDE_SIMRT2_MANUAL_F64 0
#endif
#ifndef SDE_SIMRT2_MANUAL_F32
#define SDE_SIMRT2_MANUAL_F32 0
#endif
#ifndef SDE_LOGTT_STAGES
#define SDE_LOGTT_STAGES 4 /* ring depth of the TMA-load logtransition kernel K7t: 4 KB per stage and warp, 3 CTAs per SM */
#endif
/* alignment of the row segments of the K3t / K7t tiles: 16 bytes, TMA's own requirement.  Sector alignment (32 bytes, g <= 8
 * sub-paths of 32/g map rows per tile) was built and measured on 513-row paths (profiles/r02q_alignment_variants.txt): FP64
 * BlackScholes 5.57 vs 5.58, Heston 5.54 vs 5.37, FP32 BlackScholes 4.99 vs 5.62 TB/s, logtransition 4.40 vs 4.95 — twice the
 * tensor operations of half the size cost more than the partial sectors of these load-dominated kernels, and one ragged test
 * shape hung: not shipped, the geometry code below stays general in g. */
#define SDE_SIMWT_ALIGN 16
#ifndef SDE_SIMWT_STAGES
#define SDE_SIMWT_STAGES 6 /* ring depth of the TMA Wiener-driven kernel: 4 KB per stage and warp */
#endif

typedef struct {
  double params[SDE_MAXP];
  double x0[SDE_MAXD];
  double t0, dt;
  int64_t nsteps;    /* time steps (for mlmc: number of FINE steps = ncoarse * substeps) */
  int64_t npaths;    /* paths handled by this launch */
  int64_t path0;     /* global index of the first path of this launch (Philox counter, sharding) */
  uint32_t seed_lo, seed_hi, stream;
  int32_t payoff_kind, payoff_comp;
  double payoff_k, payoff_disc;
  double lk[6];      /* noise_var-scaled constants of the FP64 normal transform (make_log_consts) */
  double noise_var;  /* variance of the generated increments: dt for the step schemes, 1 for Exact (plain N(0,1)) */
  /* mlmc only */
  double dt_coarse;
  int64_t substeps;
  /* outputs (any may be null) */
  void *out;         /* sample: [npaths][D]; simulate: see layout; mlmc: fine terminals [npaths][D] */
  void *out2;        /* mlmc: coarse terminals [npaths][D]; simulate_wiener: the Wiener input w */
  double *partials;  /* [nchunks][2*NF] per-CTA sums, then exactly accumulated by sde_accumulate */
  unsigned long long *nonfinite; /* number of paths that ended non-finite (reference: DomainError) */
  int64_t ld_paths;  /* SoA layout: leading dimension (total paths of the buffer) */
  int64_t path_off;  /* SoA layout: column offset of this launch inside the buffer */
  /* TMA trajectory kernels only: tensor maps of the w (input) and x (output) buffers in reference order */
  sde_tmap tm_w, tm_x;
} sde_args;

#endif
