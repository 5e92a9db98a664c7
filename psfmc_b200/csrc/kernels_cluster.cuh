// Fused path for 256 x 256 frames: one thread-block CLUSTER of four CTAs per walker.
// The packed frame z = raw + i*wsc*raw^2 (512 KB of complex64) does not fit one SM,
// so it is split over the shared memory of four SMs and transposed between the row
// and the column passes through distributed shared memory (st.shared::cluster) --
// like the 128 x 128 kernel it touches L2/HBM only for the walker's parameters, the
// PSF spectra (shared by all walkers), the observation and one double of output.
//
// CTA r of the cluster owns rows 64r .. 64r+63 during the row passes and 64 columns
// during the column passes; the columns are owned in MIRROR PAIRS (kx, 256 - kx): CTA q
// holds kx = 32q + s in slot s < 32 and kx = 256 - (32q + s) in slot s + 32 (CTA 0:
// slot 0 = column 0, slot 32 = column 128, the two self-mirrored columns). With both
// Z[k] and Z[-k] in the registers of one thread the two real images packed in z are
// multiplied by their own spectra without ever being separated:
//     Y[k] = A[k] P[k] + i B[k] V[k],  A = (Z[k] + conj Z[-k]) / 2,  B = (Z[k] - conj Z[-k]) / 2i
//          = Z[k] S+[k] + conj(Z[-k]) S-[k],   S+- = (P +- V) / 2
// (P, V: spectra of the PSF and of its variance map, Hermitian), so there are no
// split / Hermitian-rebuild passes and no packed DC/Nyquist columns.
//
// Length-256 transforms are 16 x 16 (radix-16 butterflies in registers, twiddles
// W256^(n2 k1), one exchange). One main tile of 128 KB per CTA serves both layouts:
//   C (columns): element (y, slot) at 64 y + slot            y = 0..255
//   R (rows):    element (y_local, kx -> (q, slot)) at 64 (64 q + y_local) + slot
// i.e. the transposition CTA r <-> CTA q moves 64 x 64 blocks without changing the
// offset inside a block, and a value is always written by its PRODUCER into the
// consumer's tile:
//   rows fwd   render 16 px/thread -> radix-16 -> exchange (2 KB warp-private scratch
//              per row) -> radix-16 -> push Z[y][kx] to the column owners (C layout)
//   -- cluster barrier --
//   columns    radix-16 | radix-16, spectrum multiply, inverse radix-16 | twiddle +
//              inverse radix-16 in registers -- cluster barrier (everybody has read its
//              tile) -- push to the row owners (R layout)
//   -- cluster barrier --
//   rows inv   gather the row from the four 64-slot blocks, inverse radix-16 ->
//              exchange -> inverse radix-16 -> residual, IVM, masked chi-square terms
//              (float64 accumulation) -> CTA partial -> CTA 0 of the cluster
//   -- cluster barrier (arrive ... wait around the next walker's render + forward
//      transform: only its pushes have to wait for the other CTAs' inverse rows) --
//
// Reference arithmetic: see render.cuh and kernels_staged.cuh; this file only
// re-schedules it.
#pragma once
#include "kernels_fused.cuh"

namespace psfmc {

#define PSFMC_CL_N 256
#define PSFMC_CL_CTAS 4
#define PSFMC_CL_THREADS 512
#define PSFMC_CL_TILE_BYTES (256 * 64 * 8)
#define PSFMC_CL_X_OFF PSFMC_CL_TILE_BYTES           // row exchange scratch: 32 rows x 2 KB
#define PSFMC_CL_TW_OFF (PSFMC_CL_X_OFF + 32 * 2048)  // W256^(j k1) as [k1][j], 2 KB
#define PSFMC_CL_RED_OFF (PSFMC_CL_TW_OFF + 2048)     // 16 warp partials + 4 CTA partials
// per-walker parameters staged one walker ahead: header (PSF index, packing scale),
// float32 render constants, float64 constants (point-source taps)
#define PSFMC_CL_PAR_OFF (PSFMC_CL_RED_OFF + 256)
#define PSFMC_CL_PAR_RC (PSFMC_CL_PAR_OFF + 64)
#define PSFMC_CL_PAR_DER (PSFMC_CL_PAR_RC + PSFMC_MAX_COMPONENTS * PSFMC_RC_STRIDE * 4)
#define PSFMC_CL_SMEM (PSFMC_CL_PAR_DER + PSFMC_MAX_COMPONENTS * PSFMC_DERIVED_STRIDE * 8)

struct ClusterParams {
  FusedParams f;          // render constants, observation, outputs (spec/specx unused)
  const float4 *spec4;    // [K][4 q][256 ky][32 lanes]: (S+, S-) of column kx = 32 q + lane
  const float4 *specx4;   // [K][2][256 ky]: (S+, S-) of the self-mirrored columns 0 and 128
  const float2 *tw;       // [16 k1][16 j]: W256^(j k1)
};

// ---- cluster primitives (PTX; emulated by tests/emu/cuda_emu.h) -----------------
__device__ __forceinline__ unsigned cl_rank() {
#ifdef PSFMC_EMU
  return emu::cluster_ctarank();
#else
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
#endif
}
// Generic pointer to the same shared-memory location in CTA `rank` of the cluster.
// Remote stores go through these 64-bit pointers (plain generic stores): a
// st.shared::cluster on a 32-bit address costs one S2R (shared-window base) PER STORE.
__device__ __forceinline__ unsigned char *cl_map(unsigned char *ptr, unsigned rank) {
#ifdef PSFMC_EMU
  return reinterpret_cast<unsigned char *>(emu::map_shared_rank((uintptr_t)ptr, rank));
#else
  unsigned long long in = reinterpret_cast<unsigned long long>(ptr), out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(in), "r"(rank));
  return reinterpret_cast<unsigned char *>(out);
#endif
}
__device__ __forceinline__ void stg64(unsigned char *ptr, cplx<float> v) {
  *reinterpret_cast<float2 *>(ptr) = make_float2(v.x, v.y);
}
__device__ __forceinline__ void cl_arrive() {
#ifdef PSFMC_EMU
  emu::cluster_arrive();
#else
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ void cl_wait() {
#ifdef PSFMC_EMU
  emu::cluster_wait();
#else
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ float4 ldg128(const float4 *p) {
#ifdef PSFMC_EMU
  return *p;
#else
  return __ldg(p);
#endif
}

// x at frequency k, y at -k, f = (S+, S-) at k:
//   x' = x S+ + conj(y) S-,   y' = y conj(S+) + conj(x) conj(S-)
__device__ __forceinline__ void cl_pair_mul(cplx<float> &x, cplx<float> &y, float4 f) {
  const cplx<float> sp = mk<float>(f.x, f.y), sm = mk<float>(f.z, f.w);
  const cplx<float> t1 = cmul_conj(y, sm);   // y conj(S-)  = conj(conj(y) S-)
  const cplx<float> t2 = x * sm;             // x S-        = conj(conj(x) conj(S-))
  const cplx<float> nx = x * sp + mk<float>(t1.x, -t1.y);
  const cplx<float> ny = cmul_conj(y, sp) + mk<float>(t2.x, -t2.y);
  x = nx;
  y = ny;
}
// self-mirrored frequency (k = -k)
__device__ __forceinline__ void cl_self_mul(cplx<float> &x, float4 f) {
  const cplx<float> sp = mk<float>(f.x, f.y), sm = mk<float>(f.z, f.w);
  const cplx<float> t1 = cmul_conj(x, sm);
  x = x * sp + mk<float>(t1.x, -t1.y);
}

// Per-thread constants of the row passes: 2 rows per warp, 16 threads per row.
struct ClRowRole {
  int w, rr, l;
  bool l0;
  smem_addr_t xrow;   // this row's 2 KB exchange scratch
  smem_addr_t twl;    // &TW[0][l]
  unsigned lx;        // 8 l
};

// Tile address of element kx = l + 16 k2 of a row whose four 64-slot blocks start at
// base[q] (C layout of the column owners for the pushes, R layout of the local tile
// for the gathers): ra[q] = base[q] + 8 l, rm[q] = base[q] - 8 l.
template <int K2, typename A = smem_addr_t>
__device__ __forceinline__ A cl_row_addr(const A *ra, const A *rm, bool l0) {
  if (K2 < 8) return ra[K2 >> 1] + 128 * (K2 & 1);          // kx < 128: slot 16 (k2 & 1) + l
  constexpr int m = 15 - K2;                                 // kx' = 256 - kx = 16 m + 16 - l
  if ((m & 1) == 0) return rm[m >> 1] + 8 * 48;              // slot 48 - l
  // m odd: slot 64 - l for l >= 1; l = 0 is kx' = 16 (m + 1), the first mirrored slot
  // of the next owner (kx = 128 -> CTA 0, slot 32)
  return l0 ? rm[((m + 1) >> 1) & 3] + 8 * 32 : rm[m >> 1] + 8 * 64;
}

// Destination of element kx = l + 16 k2 of a row: owner CTA q, base (row start + 8 l
// or row start - 8 l) and a compile-time byte offset; lane l = 0 of the odd mirrored
// blocks goes to the first mirrored slot of the NEXT owner (see cl_row_addr).
template <int K2>
struct ClRowMap {
  static constexpr int m = 15 - K2;
  static constexpr bool minus = K2 >= 8;
  static constexpr int q = K2 < 8 ? (K2 >> 1) : (m >> 1);
  static constexpr int imm = K2 < 8 ? 128 * (K2 & 1) : ((m & 1) ? 8 * 64 : 8 * 48);
  static constexpr bool l0_special = (K2 >= 8) && ((m & 1) != 0);
  static constexpr int q0 = ((m + 1) >> 1) & 3;
  static constexpr int imm0 = 8 * 32;
};

// push: ga[q] / gm[q] = generic pointers into the OWNERS' tiles (row start +- 8 l), la /
// lm the same row in this CTA's tile: the quarter that stays here goes through the
// ordinary shared-memory path, the rest through distributed shared memory
template <int K2>
struct ClRowLoop {
  static __device__ __forceinline__ void push(unsigned char *const *ga, unsigned char *const *gm,
                                              smem_addr_t la, smem_addr_t lm, bool l0,
                                              int rank, const cplx<float> *u) {
    typedef ClRowMap<K2> M;
    if (M::l0_special && l0) {
      if (M::q0 == rank)
        sts64(lm + M::imm0, u[K2]);
      else
        stg64(gm[M::q0] + M::imm0, u[K2]);
    } else if (M::q == rank) {
      sts64((M::minus ? lm : la) + M::imm, u[K2]);
    } else {
      stg64((M::minus ? gm[M::q] : ga[M::q]) + M::imm, u[K2]);
    }
    ClRowLoop<K2 + 1>::push(ga, gm, la, lm, l0, rank, u);
  }
  static __device__ __forceinline__ void gather(const smem_addr_t *ra, const smem_addr_t *rm,
                                                bool l0, cplx<float> *u) {
    u[K2] = lds64(cl_row_addr<K2>(ra, rm, l0));
    ClRowLoop<K2 + 1>::gather(ra, rm, l0, u);
  }
  // u += the same row elements of a row in ANOTHER CTA's tile (generic mapped pointers)
  static __device__ __forceinline__ void gather_add(unsigned char *const *ga,
                                                    unsigned char *const *gm, bool l0,
                                                    cplx<float> *u) {
    const float2 v = *reinterpret_cast<const float2 *>(
        cl_row_addr<K2, unsigned char *>(ga, gm, l0));
    u[K2] = u[K2] + mk<float>(v.x, v.y);
    ClRowLoop<K2 + 1>::gather_add(ga, gm, l0, u);
  }
};
template <>
struct ClRowLoop<16> {
  static __device__ __forceinline__ void push(unsigned char *const *, unsigned char *const *,
                                              smem_addr_t, smem_addr_t, bool, int,
                                              const cplx<float> *) {}
  static __device__ __forceinline__ void gather(const smem_addr_t *, const smem_addr_t *, bool,
                                                cplx<float> *) {}
  static __device__ __forceinline__ void gather_add(unsigned char *const *,
                                                    unsigned char *const *, bool,
                                                    cplx<float> *) {}
};

// render + forward row transform of row batch `it` of walker b; u[k2] = Z[y][l + 16 k2]
template <bool PADDED>
__device__ __forceinline__ void cl_rows_forward(const FusedParams &P, const ClRowRole &R,
                                                const float *rc_s, const double *der_s, int y,
                                                float wsc, cplx<float> *u,
                                                const FoldParams &F) {
  {
    cplx<float> v[16];
    if (!PADDED || y < F.Hr) {
      fused_render16<16, true>(P, rc_s, der_s, y, R.l, wsc, v);
      if (PADDED) {   // nothing is rendered outside the observation frame
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (R.l + 16 * j >= F.Wr) v[j] = mk<float>(0.0f, 0.0f);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = mk<float>(0.0f, 0.0f);
    }
    dft16<false>(v);
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) v[k1] = v[k1] * lds64(R.twl + 128 * k1);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1)
      sts64(R.xrow + 128 * k1 + (R.lx ^ (unsigned)(8 * k1)), v[k1]);
  }
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2)
    u[n2] = lds64(R.xrow + 16 * R.lx + (R.lx ^ (unsigned)(8 * n2)));
  __syncwarp();
  dft16<false>(u);
}

// inverse row transform + chi-square terms of one row; u[k2] = Y[y][l + 16 k2] on entry.
// Returns this thread's float64 partial sum over its 16 pixels.
template <bool PADDED>
__device__ __forceinline__ double cl_rows_inverse(const FusedParams &P, const ClRowRole &R,
                                                  int y, float unscale, cplx<float> *u,
                                                  const float2 *o, const FoldParams &F) {
  dft16<true>(u);   // over k2 -> n2
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2)
    sts64(R.xrow + 16 * R.lx + (R.lx ^ (unsigned)(8 * n2)), u[n2]);
  __syncwarp();
  cplx<float> v[16];
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1)
    v[k1] = lds64(R.xrow + 128 * k1 + (R.lx ^ (unsigned)(8 * k1)));
  __syncwarp();
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul_conj(v[k1], lds64(R.twl + 128 * k1));
  dft16<true>(v);   // v[j]: pixel x = l + 16 j = (convolved model, scaled model variance)
  if (PADDED) {
    // fold the linear convolution back modulo Wr (see Frame) through the row's exchange
    // scratch (free: the barrier above), pixel x at position x
#pragma unroll
    for (int j = 0; j < 16; ++j) sts64(R.xrow + 8u * (R.l + 16 * j), v[j]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int x = R.l + 16 * j;
      if (x <= F.fx_hi) v[j] = v[j] + lds64(R.xrow + 8u * (x + F.Wr));
      if (x >= F.fx_lo && x < F.Wr) v[j] = v[j] + lds64(R.xrow + 8u * (PSFMC_CL_N + x - F.Wr));
    }
    __syncwarp();
  }
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float resid, ivm;
    const double t = Epilogue<float>::term(v[j].x, v[j].y * unscale, o[j].x, fabsf(o[j].y),
                                           &resid, &ivm);
    if (__float_as_int(o[j].y) >= 0) acc += t;
  }
  (void)P;
  (void)y;
  return acc;
}

template <bool PADDED>
__global__ void __launch_bounds__(PSFMC_CL_THREADS, 1)
cluster256_lnlike_kernel(const ClusterParams CP, const FoldParams F) {
  PSFMC_DYN_SMEM(smem_raw);
  const FusedParams &P = CP.f;
  const smem_addr_t tile = smem_base(smem_raw);
  const smem_addr_t twb = tile + PSFMC_CL_TW_OFF;
  double *red_s = reinterpret_cast<double *>(smem_raw + PSFMC_CL_RED_OFF);   // [16] + [4]
  int *cnt_s = reinterpret_cast<int *>(smem_raw + PSFMC_CL_RED_OFF + 8 * 20);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned rank = cl_rank();
  const long long cluster_id = blockIdx.x / PSFMC_CL_CTAS;
  const long long n_clusters = gridDim.x / PSFMC_CL_CTAS;
  if (tid < 256) {
    const float2 t = CP.tw[tid];
    sts64(twb + 8u * tid, mk<float>(t.x, t.y));
  }
  if (tid == 0) *cnt_s = 0;
  // staged parameters of the walker the next forward pass renders
  int *par_sel = reinterpret_cast<int *>(smem_raw + PSFMC_CL_PAR_OFF);
  double *par_wsc = reinterpret_cast<double *>(smem_raw + PSFMC_CL_PAR_OFF + 8);
  float *rc_s = reinterpret_cast<float *>(smem_raw + PSFMC_CL_PAR_RC);
  double *der_s = reinterpret_cast<double *>(smem_raw + PSFMC_CL_PAR_DER);
  auto stage_params = [&](long long bs) {
    if (bs >= P.n_batch) return;
    if (tid < P.ncomp * PSFMC_RC_STRIDE)
      rc_s[tid] = __ldg(P.rconst + bs * P.ncomp * PSFMC_RC_STRIDE + tid);
    for (int k = tid; k < P.ncomp * PSFMC_DERIVED_STRIDE; k += PSFMC_CL_THREADS)
      der_s[k] = __ldg(P.derived + bs * P.ncomp * PSFMC_DERIVED_STRIDE + k);
    if (tid == PSFMC_CL_THREADS - 1) {
      *par_sel = P.psf_sel[bs];
      *par_wsc = P.wscale[bs];
    }
  };
  stage_params(cluster_id);
  // this CTA's tile as seen by each CTA of the cluster (index = owner rank)
  unsigned char *mb[PSFMC_CL_CTAS];
#pragma unroll
  for (int q = 0; q < PSFMC_CL_CTAS; ++q) mb[q] = cl_map(smem_raw, q);
  double *red0 = reinterpret_cast<double *>(cl_map(smem_raw + PSFMC_CL_RED_OFF + 8 * 16, 0));

  ClRowRole R;
  R.w = w;
  R.rr = lane >> 4;
  R.l = lane & 15;
  R.l0 = (R.l == 0);
  R.lx = 8u * R.l;
  R.xrow = tile + PSFMC_CL_X_OFF + 2048u * (2 * w + R.rr);
  R.twl = twb + R.lx;
  const int yl0 = 2 * w + R.rr;   // local row of batch 0; batch 1 is 32 rows further

  __syncthreads();
  // every CTA of the cluster is resident before anybody stores into its tile
  cl_arrive();
  cl_wait();

  // column-pass roles. Passes 1 and 3: units (slot = lane, n2 = w) and (lane + 32, w).
  // Pass 2: the mirror pair (lane, k1 = w), (lane + 32, k1 = -w mod 16); lane 0 of CTA 0
  // holds the self-mirrored columns 0 (warps 0..7) and 128 (warps 8..15) and takes the
  // pair k1 = j, 16 - j of ONE column (j = w & 7; j = 0: k1 = 0 and k1 = 8)
  const bool special = (rank == 0) && (lane == 0);
  const int j8 = w & 7;
  const int colA = special ? (w < 8 ? 0 : 32) : lane;
  const int colB = special ? colA : lane + 32;
  const int k1A = special ? j8 : w;
  const int k1B = special ? (j8 ? 16 - j8 : 8) : ((16 - w) & 15);
  const bool rule0 = (k1A == 0) && !special;   // partner of a[k2] is b[(16 - k2) & 15]
  const bool special0 = special && (j8 == 0);  // pairs inside a and inside b

  // The last cluster barrier of a walker (all inverse rows done, CTA partials delivered
  // to CTA 0) is only ARRIVED at; the wait is taken after the next walker's render and
  // forward transform, right before its first push -- `pending` is the walker whose
  // lnL CTA 0 still has to assemble then.
  long long pending = -1;
  bool pending_invalid = false;
  auto finish_pending = [&]() {
    cl_wait();
    if (rank == 0 && tid == 0) {
      double tot = 0.0;
      for (int k = 0; k < PSFMC_CL_CTAS; ++k) tot += red_s[16 + k];
      double val = -0.5 * tot;
      if (!isfinite(val) || pending_invalid) val = -INFINITY;
      P.lnl[pending] = val;
    }
    pending = -1;
  };
#pragma unroll 1
  for (long long b = cluster_id; b < P.n_batch; b += n_clusters) {
    int sel = *par_sel;
    const bool invalid = sel < 0;
    if (invalid) sel = 0;
    const double wscale_b = *par_wsc;
    const float unscale = (float)(P.vscale_inv[sel] / wscale_b);

    // ------------------------------------------ rows: render + forward + push --
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
      const int yl = yl0 + 32 * it;
      const int y = 64 * (int)rank + yl;
      cplx<float> u[16];
      // (padded frames: a warp whose two rows lie outside the observation frame has
      // nothing to render or transform -- it only delivers the zeros)
      if (PADDED && y - R.rr >= F.Hr) {
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) u[k2] = mk<float>(0.0f, 0.0f);
      } else {
        cl_rows_forward<PADDED>(P, R, rc_s, der_s, y, (float)wscale_b, u, F);
      }
      if (it == 0 && pending >= 0) finish_pending();   // the other CTAs are done with
                                                       // the last walker's rows
      unsigned char *ga[PSFMC_CL_CTAS], *gm[PSFMC_CL_CTAS];
#pragma unroll
      for (int q = 0; q < PSFMC_CL_CTAS; ++q) {
        unsigned char *rowb = mb[q] + 512 * y;
        ga[q] = rowb + R.lx;
        gm[q] = rowb - R.lx;
      }
      const smem_addr_t lrow = tile + 512u * (unsigned)y;
      ClRowLoop<0>::push(ga, gm, lrow + R.lx, lrow - R.lx, R.l0, (int)rank, u);
    }
    cl_arrive();
    cl_wait();
    // every thread of the CTA is past its forward rows: stage the next walker's
    // parameters (read again after the CTA barrier of this walker's reduction)
    stage_params(b + n_clusters);

    // spectra of unit A of the multiply pass at ky = k1A + 16 k2 (of unit B for the
    // second half of special0)
    const float4 *sp;
    int sstride;
    if (special) {
      sp = CP.specx4 + ((size_t)sel * 2 + (w < 8 ? 0 : 1)) * 256 + k1A;
      sstride = 16;
    } else {
      sp = CP.spec4 + (((size_t)sel * PSFMC_CL_CTAS + rank) * 256 + k1A) * 32 + lane;
      sstride = 16 * 32;
    }

    // ------------------------------------------------- columns: radix-16 (n1) --
#pragma unroll
    for (int unit = 0; unit < 2; ++unit) {
      const smem_addr_t base = tile + 512u * w + 8u * (lane + 32 * unit);
      cplx<float> v[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) v[n1] = lds64(base + 8192 * n1);
      dft16<false>(v);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v[k1] = v[k1] * lds64(twb + 8u * w + 128 * k1);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) sts64(base + 8192 * k1, v[k1]);
    }
    // first half of the spectrum values: their L2 latency overlaps the CTA barrier
    float4 f[8];
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) f[k2] = ldg128(sp + k2 * sstride);
    __syncthreads();

    // -------------- columns: radix-16 (n2), spectrum multiply, inverse radix-16 --
    {
      const smem_addr_t baseA = tile + 8192u * k1A + 8u * colA;
      const smem_addr_t baseB = tile + 8192u * k1B + 8u * colB;
      cplx<float> a[16], bb[16];
#pragma unroll
      for (int n2 = 0; n2 < 16; ++n2) {
        a[n2] = lds64(baseA + 512 * n2);
        bb[n2] = lds64(baseB + 512 * n2);
      }
      dft16<false>(a);    // a[k2]  = Z[k1A + 16 k2][colA]
      dft16<false>(bb);   // bb[k2] = Z[k1B + 16 k2][colB]
      if (special0) {
        cl_self_mul(a[0], ldg128(sp));
        cl_self_mul(a[8], ldg128(sp + 8 * sstride));
#pragma unroll
        for (int k2 = 1; k2 < 8; ++k2) cl_pair_mul(a[k2], a[16 - k2], ldg128(sp + k2 * sstride));
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2)
          cl_pair_mul(bb[k2], bb[15 - k2], ldg128(sp + 8 + k2 * sstride));
      } else if (rule0) {
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
          cl_pair_mul(a[k2], bb[(16 - k2) & 15], f[k2 & 7]);
          if (k2 < 8) f[k2] = ldg128(sp + (k2 + 8) * sstride);
        }
      } else {
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
          cl_pair_mul(a[k2], bb[15 - k2], f[k2 & 7]);
          if (k2 < 8) f[k2] = ldg128(sp + (k2 + 8) * sstride);
        }
      }
      dft16<true>(a);     // over k2 -> n2
      dft16<true>(bb);
#pragma unroll
      for (int n2 = 0; n2 < 16; ++n2) {
        sts64(baseA + 512 * n2, a[n2]);
        sts64(baseB + 512 * n2, bb[n2]);
      }
    }
    __syncthreads();

    // --------------- columns: inverse radix-16 (k1) in registers, then push rows --
    {
      cplx<float> v0[16], v1[16];
      const smem_addr_t base = tile + 512u * w + 8u * lane;
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) v0[k1] = lds64(base + 8192 * k1);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) v1[k1] = lds64(base + 256 + 8192 * k1);
      cl_arrive();        // this thread has read its part of the tile
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) {
        const cplx<float> tw = lds64(twb + 8u * w + 128 * k1);
        v0[k1] = cmul_conj(v0[k1], tw);
        v1[k1] = cmul_conj(v1[k1], tw);
      }
      dft16<true>(v0);    // v[n1]: row y = w + 16 n1
      cl_wait();          // every CTA has read its tile: the tiles may be overwritten
      const unsigned off = 512u * (64u * rank + w) + 8u * lane;
      unsigned char *pb[PSFMC_CL_CTAS];
#pragma unroll
      for (int q = 0; q < PSFMC_CL_CTAS; ++q) pb[q] = mb[q] + off;
      const smem_addr_t lb = tile + off;
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        if ((n1 >> 2) == (int)rank)
          sts64(lb + 8192 * (n1 & 3), v0[n1]);
        else
          stg64(pb[n1 >> 2] + 8192 * (n1 & 3), v0[n1]);
      }
      dft16<true>(v1);    // overlaps the first unit's stores in flight
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        if ((n1 >> 2) == (int)rank)
          sts64(lb + 256 + 8192 * (n1 & 3), v1[n1]);
        else
          stg64(pb[n1 >> 2] + 256 + 8192 * (n1 & 3), v1[n1]);
      }
    }
    cl_arrive();
    // observation + signed variance of the first row batch: their L2 latency overlaps
    // the cluster barrier
    float2 o[16];
    {
      // (pixel x = l + 16 j sits at position 32 (j >> 1) + 2 l + (j & 1) of its row, see
      // cluster_ow_index: a thread's pixels j = 2 i, 2 i + 1 are one 16-byte load)
      const float4 *owr = reinterpret_cast<const float4 *>(
          P.ow + (64 * (int)rank + yl0) * PSFMC_CL_N + 2 * R.l);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 q = __ldg(owr + 16 * i);
        o[2 * i] = make_float2(q.x, q.y);
        o[2 * i + 1] = make_float2(q.z, q.w);
      }
    }
    cl_wait();

    // ------------------------------------------- rows: inverse + chi-square --
    double acc = 0.0;
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
      const int yl = yl0 + 32 * it;
      const int y = 64 * (int)rank + yl;
      if (PADDED && y - R.rr >= F.Hr) continue;   // both rows of the warp are padding
      if (it) {
        const float4 *owr = reinterpret_cast<const float4 *>(P.ow + y * PSFMC_CL_N + 2 * R.l);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 q = __ldg(owr + 16 * i);
          o[2 * i] = make_float2(q.x, q.y);
          o[2 * i + 1] = make_float2(q.z, q.w);
        }
      }
      smem_addr_t ra[PSFMC_CL_CTAS], rm[PSFMC_CL_CTAS];
#pragma unroll
      for (int q = 0; q < PSFMC_CL_CTAS; ++q) {
        const smem_addr_t rowb = tile + 512u * (unsigned)(64 * q + yl);
        ra[q] = rowb + R.lx;
        rm[q] = rowb - R.lx;
      }
      cplx<float> u[16];
      ClRowLoop<0>::gather(ra, rm, R.l0, u);
      if (PADDED && y < F.Hr) {
        // fold the linear convolution back modulo Hr (see Frame): row y also receives
        // rows y + Hr and 256 + y - Hr, which other CTAs of the cluster own -- read
        // straight from their tiles (final since the last cluster barrier, overwritten
        // only after the next one)
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
          const bool take = which == 0 ? (y <= F.fy_hi) : (y >= F.fy_lo);
          if (!take) continue;
          const int yq = which == 0 ? y + F.Hr : PSFMC_CL_N + y - F.Hr;
          const int oq = yq >> 6;   // owner of row yq (selects keep mb[] in registers)
          unsigned char *ob = oq == 0 ? mb[0] : oq == 1 ? mb[1] : oq == 2 ? mb[2] : mb[3];
          unsigned char *ga[PSFMC_CL_CTAS], *gm[PSFMC_CL_CTAS];
#pragma unroll
          for (int q = 0; q < PSFMC_CL_CTAS; ++q) {
            unsigned char *rowb = ob + 512 * (64 * q + (yq & 63));
            ga[q] = rowb + R.lx;
            gm[q] = rowb - R.lx;
          }
          ClRowLoop<0>::gather_add(ga, gm, R.l0, u);
        }
      }
      acc += cl_rows_inverse<PADDED>(P, R, y, unscale, u, o, F);
    }
    // float64 reduction: warp shuffles; the last warp to arrive sums the 16 warp
    // partials in fixed order (deterministic) and delivers the CTA partial to CTA 0
    // -- before its own barrier arrival below, which publishes it
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      volatile double *red = red_s;
      red[w] = acc;
      __threadfence_block();
      const int prev = atomicAdd(cnt_s, 1);
      if (prev == PSFMC_CL_THREADS / 32 - 1) {
        __threadfence_block();
        double tot = 0.0;
        for (int k = 0; k < PSFMC_CL_THREADS / 32; ++k) tot += red[k];
        red0[rank] = tot;
        *cnt_s = 0;
      }
    }
    // this CTA is done with its tile (rows) and has delivered its partial
    cl_arrive();
    pending = b;
    pending_invalid = invalid;
  }
  if (pending >= 0) finish_pending();
}

// -------------------------------------------------------------- host side --

template <typename T>
inline bool cluster_path_available(const StagedPlan &plan) {
  return sizeof(T) == 4 && plan.fr.H == PSFMC_CL_N && plan.fr.W == PSFMC_CL_N;
}

// W256^(j k1) as [k1][j], computed in long double and rounded once
inline void cluster_twiddles(float2 *tw) {
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int k1 = 0; k1 < 16; ++k1)
    for (int j = 0; j < 16; ++j) {
      const long double ang = -2.0L * pi * (long double)(j * k1) / 256.0L;
      tw[k1 * 16 + j].x = (float)cosl(ang);
      tw[k1 * 16 + j].y = (float)sinl(ang);
    }
}

// Re-layout of the float64 spectra [K][2*Wc][H] (column-major, see kernels_staged.cuh)
// as S+- = (P +- vscale V) / 2 per column:
//   spec4 [K][q][ky][lane] = (S+, S-) at kx = 32 q + lane, specx4[K][t][ky] at kx = 128 t
inline void cluster_spectrum_layout(const cplx<double> *spec64, int n_psf, const double *vscale,
                                    float4 *spec4, float4 *specx4) {
  constexpr int N = PSFMC_CL_N, Wc = N / 2 + 1;
  for (int k = 0; k < n_psf; ++k) {
    const cplx<double> *src = spec64 + (size_t)k * 2 * Wc * N;
    auto at = [&](int kx, int ky) {
      const cplx<double> &p = src[((size_t)0 * Wc + kx) * N + ky];
      const cplx<double> &v = src[((size_t)1 * Wc + kx) * N + ky];
      float4 o;
      o.x = (float)(0.5 * (p.x + vscale[k] * v.x));
      o.y = (float)(0.5 * (p.y + vscale[k] * v.y));
      o.z = (float)(0.5 * (p.x - vscale[k] * v.x));
      o.w = (float)(0.5 * (p.y - vscale[k] * v.y));
      return o;
    };
    for (int q = 0; q < PSFMC_CL_CTAS; ++q)
      for (int ky = 0; ky < N; ++ky)
        for (int lane = 0; lane < 32; ++lane)
          spec4[(((size_t)k * PSFMC_CL_CTAS + q) * N + ky) * 32 + lane] = at(32 * q + lane, ky);
    for (int t = 0; t < 2; ++t)
      for (int ky = 0; ky < N; ++ky) specx4[((size_t)k * 2 + t) * N + ky] = at(128 * t, ky);
  }
}

struct ClusterBuffers {
  float *rconst = nullptr;
  const float4 *spec4 = nullptr, *specx4 = nullptr;
  const float2 *ow = nullptr, *tw = nullptr;
  int n_clusters = 32;   // clusters that can be resident at once
};

inline int cluster_prepare_device(int *n_clusters_out) {
#ifdef PSFMC_EMU
  *n_clusters_out = 2;
  return 0;
#else
  if (cudaFuncSetAttribute(cluster256_lnlike_kernel<false>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           PSFMC_CL_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(cluster256_lnlike_kernel<true>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           PSFMC_CL_SMEM) != cudaSuccess)
    return 1;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(PSFMC_CL_CTAS * 64);
  cfg.blockDim = dim3(PSFMC_CL_THREADS);
  cfg.dynamicSmemBytes = PSFMC_CL_SMEM;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PSFMC_CL_CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, cluster256_lnlike_kernel<false>, &cfg) != cudaSuccess ||
      n < 1)
    return 1;
  *n_clusters_out = n;
  return 0;
#endif
}

// theta -> lnL for n_batch walkers: prepare kernel + one persistent cluster kernel.
template <typename T>
inline int launch_cluster_lnlike(const StagedPlan &plan, const StagedBuffers<T> &buf,
                                 const ClusterBuffers &cb, const Program &prog_h,
                                 const double *theta, long long n_batch, long long ld,
                                 double *lnl, cudaStream_t stream,
                                 cudaEvent_t ev_begin = nullptr, cudaEvent_t ev_end = nullptr) {
  if (n_batch <= 0) return 0;
  const int ncomp = prog_h.n_components;
  launch_prepare(*buf.prog_host, theta, n_batch, ld, plan.fr.Hr, plan.fr.Wr, ncomp, buf.derived,
                 buf.psf_sel, buf.wscale, cb.rconst, stream);
  ClusterParams CP;
  FoldParams F;
  F.Hr = plan.fr.Hr;
  F.Wr = plan.fr.Wr;
  F.fy_hi = plan.fr.fy_hi;
  F.fy_lo = plan.fr.fy_lo;
  F.fx_hi = plan.fr.fx_hi;
  F.fx_lo = plan.fr.fx_lo;
  FusedParams &P = CP.f;
  P.rconst = cb.rconst;
  P.derived = buf.derived;
  P.wscale = buf.wscale;
  P.psf_sel = buf.psf_sel;
  P.vscale_inv = buf.vscale_inv;
  P.spec4 = nullptr;
  P.specx4 = nullptr;
  P.ow = cb.ow;
  P.maskw = nullptr;
  P.lnl_const = 0.0;
  P.n_peer = 0;
  P.hot = nullptr;
  P.nan_marks = 0;
  P.kpv = nullptr;
  P.sub_spec = nullptr;
  P.partials = nullptr;
  P.skip_tab = nullptr;
  P.lnl = lnl;
  P.n_batch = n_batch;
  P.ncomp = ncomp;
  P.kind_bits = 0;
  for (int c = 0; c < ncomp; ++c) P.kind_bits |= (unsigned long long)(prog_h.kind[c] & 3) << (2 * c);
  CP.spec4 = cb.spec4;
  CP.specx4 = cb.specx4;
  CP.tw = cb.tw;
  const unsigned nclus = (unsigned)(n_batch < cb.n_clusters ? n_batch : cb.n_clusters);
  if (ev_begin) cudaEventRecord(ev_begin, stream);
  if (plan.fr.padded)
    launch_kernel_cluster(cluster256_lnlike_kernel<true>, dim3(nclus * PSFMC_CL_CTAS),
                          dim3(PSFMC_CL_THREADS), (size_t)PSFMC_CL_SMEM, stream, PSFMC_CL_CTAS,
                          CP, F);
  else
    launch_kernel_cluster(cluster256_lnlike_kernel<false>, dim3(nclus * PSFMC_CL_CTAS),
                          dim3(PSFMC_CL_THREADS), (size_t)PSFMC_CL_SMEM, stream, PSFMC_CL_CTAS,
                          CP, F);
  if (ev_end) cudaEventRecord(ev_end, stream);
  return 2;
}

}  // namespace psfmc
