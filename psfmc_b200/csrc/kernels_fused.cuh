// Fused single-kernel shared-memory path: 128 x 128 frames, float32 render + FFT,
// float64 chi-square accumulation. One persistent CTA per SM walks over the
// walkers of the batch; the whole packed frame z = raw + i*raw^2 (128 KB of
// complex64) lives in shared memory from render to reduction, so per walker the
// kernel touches HBM/L2 only for its parameters, the PSF spectra (128 KB, L2
// resident, shared by all walkers), the observation (128 KB, L2 resident) and
// one double of output.
//
// Length-128 transforms are split 16 x 8: a radix-16 butterfly in registers, the
// inter-stage twiddles W128^(n2*k1), one exchange through shared memory, a radix-8
// butterfly in registers (index maps n = n2 + 8*n1, k = k1 + 16*k2).
//
//   rows fwd  (per warp, 4 rows at a time, no CTA barrier): render 16 px/thread
//             -> radix-16 -> twiddle -> exchange (warp-private) -> radix-8 -> the row
//             spectrum Z[kx], kx = 0..127, of the PACKED row goes to column kx as it is
//   cols      radix-16 -> exchange inside a 4-warp column group -> radix-8 ->
//             mirror-pair product with the PSF / PSF-variance spectra (registers):
//                 Y[k] = Z[k] S+[k] + conj(Z[-k]) S-[k],   S+- = (P +- V) / 2
//             (the two real images packed in z are multiplied by their own spectra
//             without ever being separated: no real-pair split, no Hermitian rebuild,
//             no packed DC/Nyquist columns; a thread owns the columns kx and -kx and
//             the frequencies ky and -ky, so both partners sit in its registers)
//             -> inverse radix-8 -> exchange -> inverse radix-16
//   rows inv  inverse radix-8 -> exchange -> inverse radix-16 -> Re = convolved
//             model, Im = model variance -> residual, composite IVM, masked
//             chi-square terms in registers -> float64 warp + CTA reduction -> lnL
//
// Two CTA-wide barriers and two 128-thread named barriers per walker; warps drift
// apart between them, which overlaps the SFU-bound render of one warp with the
// shared-memory-bound butterflies of another.
//
// Reference arithmetic: see render.cuh (components) and kernels_staged.cuh
// (convolution / likelihood); this file only re-schedules it.
#pragma once
#include "pipeline.cuh"
#include "tma.cuh"
#include "twiddle128.cuh"

namespace psfmc {

#define PSFMC_FUSED_N 128
#define PSFMC_FUSED_THREADS 512
// Bytes per tile row: 128 complex64 values (column layout: column kx at 8 kx) + room for
// the row pass's exchange layout (16 x 10 complex positions, see lds64 below). The
// pitch is an ODD multiple of 64 bytes, so the two rows of a half-warp fall on different
// banks in the column layout without any swizzle.
#define PSFMC_FUSED_ROWB 1344
#define PSFMC_FUSED_SMEM (PSFMC_FUSED_N * PSFMC_FUSED_ROWB)

struct FusedParams {
  const float *rconst;       // [B][ncomp][PSFMC_RC_STRIDE]
  const double *derived;     // [B][ncomp][PSFMC_DERIVED_STRIDE] (point-source taps)
  const double *wscale;      // [B]
  const int *psf_sel;        // [B]
  const double *vscale_inv;  // [K]
  const float4 *spec4;       // [K][ky=128][slot=64]: (S+, S-) of column kx = slot, see
                             // fused_spectrum_layout()
  const float4 *specx4;      // [K][ky=128]: the same for the self-mirrored column kx = 64
  const float2 *ow;          // [128*128]: (obs, ovar); (0, 1e30) at excluded pixels
  const unsigned short *maskw;  // [128 rows][8]: bit j of entry (y, l): pixel (y, l + 8 j) good
  double lnl_const;          // ln(2 pi) * number of good pixels
  double *lnl;               // [B]
  long long n_batch;
  int ncomp;
  unsigned skip_quads;       // bit q: rows 4q .. 4q+3 hold no good pixel (mask, bad pixels,
                             // padding): their inverse row transform and epilogue are skipped
  // 512 x 512 frames as 4 x 4 interleaved 128 x 128 sub-images (kernels_tiled.cuh): the
  // forward / inverse halves of this kernel exchange the sub-images' spectra through
  // `sub_spec` [jobs][ky=128][kx=128] (job = 16 * walker + sub-image); the inverse half
  // writes one partial sum per job; row-skip bits per sub-image
  cplx<float> *sub_spec;
  double *partials;
  const unsigned *skip_tab;
  // Hot pixels (prepare_kernel): hot[b][c] = (py << 16 | px) of the pixel that component c
  // of walker b has taken out of the transform, or -1 (null: feature off); kpv[K][dy][dx] =
  // (PSF, v * PSF variance) real-space kernels of the circular convolution, lag (dy, dx)
  const int *hot;
  const float2 *kpv;
  // nan_marks (host calls of an engine with the float64 repeat): a result that came out
  // non-finite is reported as NaN = "worth repeating in float64"; -inf is then reserved
  // for walkers that are dead by construction (PSF index out of range, Sersic constants
  // poisoned by the prepare kernel: the reference is -inf there too), which are final.
  // The host turns whatever NaN is left into -inf.
  int nan_marks;
  // lnL gather over peer memory (engine.cu, psfmc_lnlike_batch_exchange): with n_peer > 0
  // the result of walker b goes to lnl_peer[p][b] for every rank p (the ranks' mailboxes,
  // mapped through CUDA IPC; plain stores over NVLink) instead of lnl[b]
  double *lnl_peer[16];
  int n_peer;
  unsigned long long kind_bits;   // 2 bits per component (a kernel-parameter ARRAY indexed
                                  // by a loop variable is copied to local memory)
  // Blob images (IMAGES instance only; psfmc_render_batch / psfmc_accumulate_batch,
  // psfMC/models.py:213-226): [B][128 * 128] float each, or null. The observation and its
  // variance as the staged kernels see them (img_obs / img_ovar: no (0, 1e30) at excluded
  // pixels -- the images cover every pixel). ps_only: only the point sources are rendered;
  // img_resid is then the point-source-subtracted image (models.py:296-306).
  float *img_raw, *img_conv, *img_resid, *img_ivm;
  const float *img_obs, *img_ovar;
  int ps_only;
};

// padded frames (see Frame in common.cuh): observation frame and fold bounds
struct FoldParams {
  int Hr, Wr, fy_hi, fy_lo, fx_hi, fx_lo;
};

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
#ifdef PSFMC_EMU
  emu::named_barrier(id, nthreads);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}

__device__ __forceinline__ cplx<float> tw128(int n2, int k1) {
  const int idx = n2 * 16 + k1;
  return mk<float>(c_tw128[idx][0], c_tw128[idx][1]);
}

// 16-point DFT in registers, natural order in and out: two radix-4 layers
// (n = 4*n1 + n2, k = k1 + 4*k2) with the W16 twiddles in between.
template <bool INV>
__device__ __forceinline__ void dft16(cplx<float> *v) {
  const float C = 0.92387953251128675613f, S = 0.38268343236508977173f;
  const float h = 0.70710678118654752440f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4<float, INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  // v[4*k1 + n2] *= W16^(n2*k1) = (cos, -sin)(pi e / 8), conjugated for the inverse
  v[4 * 1 + 1] = crot<float, INV>(v[4 * 1 + 1], C, S);     // e = 1
  v[4 * 1 + 2] = crot<float, INV>(v[4 * 1 + 2], h, h);     // e = 2
  v[4 * 1 + 3] = crot<float, INV>(v[4 * 1 + 3], S, C);     // e = 3
  v[4 * 2 + 1] = crot<float, INV>(v[4 * 2 + 1], h, h);     // e = 2
  v[4 * 2 + 2] = INV ? mul_pos_i(v[4 * 2 + 2]) : mul_neg_i(v[4 * 2 + 2]);  // e = 4
  v[4 * 2 + 3] = crot<float, INV>(v[4 * 2 + 3], -h, h);    // e = 6
  v[4 * 3 + 1] = crot<float, INV>(v[4 * 3 + 1], S, C);     // e = 3
  v[4 * 3 + 2] = crot<float, INV>(v[4 * 3 + 2], -h, h);    // e = 6
  v[4 * 3 + 3] = crot<float, INV>(v[4 * 3 + 3], -C, -S);   // e = 9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
    dft4<float, INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // X[k1 + 4*k2] sits in v[4*k1 + k2]: transpose to natural order
  cplx<float> t[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) t[i] = v[i];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = t[4 * (k & 3) + (k >> 2)];
}

// Shared-memory tile access by BYTE offset (32-bit address arithmetic). Exchange layout
// of a row between its radix-16 and radix-8 sides: element (k1, n2) of row y sits at
// complex position
//     18 (k1 >> 1) + 8 (k1 & 1) + n2
// inside the row's own 1344 bytes (an odd multiple of 64: consecutive rows are half a
// bank cycle apart). Radix-16 side (one k1 per instruction, lanes over n2, two rows per
// half-warp): 2 x 8 consecutive positions on opposite bank halves = one conflict-free
// 128-byte wavefront. Radix-8 side: thread l owns k1 = 2 l and 2 l + 1 -- with the
// mirror-pair product no thread needs a frequency's Hermitian partner any more, so the
// pairing is free -- i.e. 128 contiguous, 16-byte aligned bytes = eight 128-bit accesses,
// and the eight threads of a row (pitch 144 bytes) cover all eight 16-byte bank groups
// (9 l mod 8 is a permutation). Its results, the row frequencies kx = 2 l + 16 k2 and
// kx + 1, are neighbours in the column layout too: 128-bit accesses there as well. Every
// address is a per-thread base plus a compile-time immediate (the XOR-swizzled layout of
// round 1 cost one LOP3 per access and twice the number of accesses).
#ifdef PSFMC_EMU
typedef uintptr_t smem_addr_t;
#else
typedef unsigned smem_addr_t;
#endif

__device__ __forceinline__ cplx<float> lds64(smem_addr_t addr) {
#ifdef PSFMC_EMU
  return *reinterpret_cast<const cplx<float> *>(addr);
#else
  cplx<float> v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void sts64(smem_addr_t addr, cplx<float> v) {
#ifdef PSFMC_EMU
  *reinterpret_cast<cplx<float> *>(addr) = v;
#else
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
#endif
}
// Global accesses that are used exactly once (the sub-spectra of the tiled path):
// streaming cache hints keep them from evicting the L1 / L2 lines the kernel reuses.
__device__ __forceinline__ void stream_store(cplx<float> *p, cplx<float> v) {
#if defined(PSFMC_EMU) || defined(PSFMC_NO_STREAM_HINTS)
  *p = v;
#else
  __stcs(reinterpret_cast<float2 *>(p), make_float2(v.x, v.y));
#endif
}
__device__ __forceinline__ cplx<float> stream_load(const cplx<float> *p) {
#if defined(PSFMC_EMU) || defined(PSFMC_NO_STREAM_HINTS)
  return *p;
#else
  const float2 v = __ldcs(reinterpret_cast<const float2 *>(p));
  return mk<float>(v.x, v.y);
#endif
}

// Tiled path: the sub-spectrum of a job travels between global memory and the tile ROW BY
// ROW (row ky: 128 complex64 values = 1 KB, contiguous on both sides) with bulk
// asynchronous copies (TMA, cp.async.bulk): no registers, no LSU instructions, and --
// the point -- asynchronous, so that the inverse half fetches the NEXT job's rows while it
// still works on this job's epilogue, and the forward half's results drain while the next
// job is being rendered. The emulator build copies synchronously at the same places.
#define PSFMC_SUBROW_BYTES (PSFMC_FUSED_N * 8)
// one lane: global row -> tile row; completion is counted on `bar` (device builds)
__device__ __forceinline__ void tile_row_fetch(unsigned char *tile_row,
                                               const cplx<float> *grow,
                                               unsigned long long *bar) {
#ifdef PSFMC_EMU
  (void)bar;
  cplx<float> *dst = reinterpret_cast<cplx<float> *>(tile_row);
  for (int k = 0; k < PSFMC_FUSED_N; ++k) dst[k] = grow[k];
#else
  bulk_load(tile_row, grow, PSFMC_SUBROW_BYTES, bar);
#endif
}
// one lane: tile row -> global row (joins the lane's current bulk group, device builds)
__device__ __forceinline__ void tile_row_store(cplx<float> *grow, const unsigned char *tile_row) {
#ifdef PSFMC_EMU
  const cplx<float> *src = reinterpret_cast<const cplx<float> *>(tile_row);
  for (int k = 0; k < PSFMC_FUSED_N; ++k) grow[k] = src[k];
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(grow),
               "r"(smem_u32(tile_row)), "r"((unsigned)PSFMC_SUBROW_BYTES)
               : "memory");
#endif
}
__device__ __forceinline__ void tile_store_commit() {
#ifndef PSFMC_EMU
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#endif
}
// the lane's bulk stores have READ their shared-memory source: the tile may be rewritten
__device__ __forceinline__ void tile_store_wait_read() {
#ifndef PSFMC_EMU
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
// generic-proxy accesses of shared memory before this point are ordered before later
// asynchronous-proxy (TMA) accesses of the same addresses
__device__ __forceinline__ void async_proxy_fence() {
#ifndef PSFMC_EMU
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

struct cplx2f {   // two adjacent complex64 values (one 128-bit shared-memory access)
  cplx<float> lo, hi;
};
__device__ __forceinline__ cplx2f lds128(smem_addr_t addr) {
  cplx2f v;
#ifdef PSFMC_EMU
  v.lo = reinterpret_cast<const cplx<float> *>(addr)[0];
  v.hi = reinterpret_cast<const cplx<float> *>(addr)[1];
#else
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.lo.x), "=f"(v.lo.y), "=f"(v.hi.x), "=f"(v.hi.y)
               : "r"(addr)
               : "memory");
#endif
  return v;
}
__device__ __forceinline__ void sts128(smem_addr_t addr, cplx<float> lo, cplx<float> hi) {
#ifdef PSFMC_EMU
  reinterpret_cast<cplx<float> *>(addr)[0] = lo;
  reinterpret_cast<cplx<float> *>(addr)[1] = hi;
#else
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(lo.x), "f"(lo.y),
               "f"(hi.x), "f"(hi.y)
               : "memory");
#endif
}
__device__ __forceinline__ smem_addr_t smem_base(unsigned char *ptr) {
#ifdef PSFMC_EMU
  return (smem_addr_t)ptr;
#else
  return (smem_addr_t)__cvta_generic_to_shared(ptr);
#endif
}

// Raw model at the 16 pixels x = l + XS*j of row y -> packed z = raw + i*wsc*raw^2
// (XS = 8: 128-wide rows, XS = 16: 256-wide rows of the cluster kernel).
// Pixels are rendered in pairs (j, j+1) with element-wise pair arithmetic.
template <int XS, bool STAGED, bool IMAGES = false>
__device__ __forceinline__ void fused_render16(const FusedParams &P, const float *rc0,
                                               const double *der0, int y, int l, float wsc,
                                               cplx<float> *v) {
  // rc0 / der0: this walker's render constants / float64 constants; STAGED = they were
  // copied to shared memory beforehand (plain loads), else read-only global loads
  cplx<float> acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = mk<float>(0.0f, 0.0f);
  for (int c = 0; c < P.ncomp; ++c) {
    const int kind = (int)((P.kind_bits >> (2 * c)) & 3ull);
    const float *rc = rc0 + c * PSFMC_RC_STRIDE;
    if (IMAGES && P.ps_only && kind != PSFMC_POINT) continue;
    if (kind == PSFMC_SKY) {
      const cplx<float> adu = bcast(STAGED ? *rc : __ldg(rc));
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = acc[i] + adu;
    } else if (kind == PSFMC_SERSIC) {
      SersicF32 s;
      const float4 *rc4 = reinterpret_cast<const float4 *>(rc);
      const float4 q0 = STAGED ? rc4[0] : __ldg(rc4);
      const float4 q1 = STAGED ? rc4[1] : __ldg(rc4 + 1);
      const float4 q2 = STAGED ? rc4[2] : __ldg(rc4 + 2);
      s.xi = q0.x; s.xf = q0.y; s.yi = q0.z; s.yf = q0.w;
      s.a00 = q1.x; s.a01 = q1.y; s.a10 = q1.z; s.a11 = q1.w;
      s.p = q2.x; s.c0 = q2.y; s.c1 = q2.z; s.kq = q2.w;
      const float dy = ((float)y - s.yi) - s.yf;
      const float cu = s.a01 * dy, cv = s.a11 * dy, dy2 = dy * dy;
      // (x - xi) is an exact small integer; subtracting the fraction LAST keeps dx
      // accurate to an ulp of dx itself next to the centre (where the profile and its
      // centroid correction vary as a power of the distance)
      const cplx<float> dxi = bcast((float)l - s.xi), nxf = bcast(-s.xf);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const cplx<float> dx =
            (dxi + mk<float>((float)(2 * XS * i), (float)(2 * XS * i + XS))) + nxf;
        if ((PSFMC_RCP_PATTERN >> i) & 1)
          acc[i] = sersic_pair_f32<true>(s, dx, cu, cv, dy2, acc[i]);
        else
          acc[i] = sersic_pair_f32<false>(s, dx, cu, cv, dy2, acc[i]);
      }
    } else {  // point source: at most 7 x 7 pixels of the frame, float64 taps
      const double *d = der0 + c * PSFMC_DERIVED_STRIDE;
      const int ymin = (int)(STAGED ? d[D_PS_YMIN] : __ldg(d + D_PS_YMIN));
      const int ymax = (int)(STAGED ? d[D_PS_YMAX] : __ldg(d + D_PS_YMAX));
      if (y >= ymin && y <= ymax) {
        const int xmin = (int)(STAGED ? d[D_PS_XMIN] : __ldg(d + D_PS_XMIN));
        const int xmax = (int)(STAGED ? d[D_PS_XMAX] : __ldg(d + D_PS_XMAX));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int x = l + 2 * XS * i;
          if (x >= xmin && x <= xmax) acc[i].x += (float)point_pixel(d, x, y);
          if (x + XS >= xmin && x + XS <= xmax) acc[i].y += (float)point_pixel(d, x + XS, y);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const cplx<float> sq = pmul(pmul(acc[i], acc[i]), bcast(wsc));
    v[2 * i] = mk<float>(acc[i].x, sq.x);
    v[2 * i + 1] = mk<float>(acc[i].y, sq.y);
  }
}

__device__ __forceinline__ cplx<float> ldc2(const cplx<float> *p) {
#ifdef PSFMC_EMU
  return *p;
#else
  const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
  return mk<float>(v.x, v.y);
#endif
}

// Mirror-pair product. z = raw + i w raw^2 packs two real images; with Z = FFT2(z),
// A = FFT2(raw) = (Z[k] + conj Z[-k]) / 2 and B = FFT2(w raw^2) = (Z[k] - conj Z[-k]) / 2i,
// so the spectrum of (raw * psf) + i (w raw^2 * psf_var) is
//     Y[k]  = A P + i B V = Z[k] S+ + conj(Z[-k]) S-,          S+- = (P +- V) / 2,
//     Y[-k] = conj( conj(Z[-k]) S+ + Z[k] S- )                 (P, V Hermitian).
// zu = Z[k], zm = Z[-k], s = (S+.re, S+.im, S-.re, S-.im) at k. Eight packed
// instructions per pair: the per-half sign flips all sit on the accumulator operand
// (FFMA2's .NP / .PN operand modifiers), the swaps and full negations on the others.
__device__ __forceinline__ void mirror_pair(cplx<float> &zu, cplx<float> &zm, float4 s) {
#ifdef PSFMC_EMU
  const cplx<float> sp = mk<float>(s.x, s.y), sm = mk<float>(s.z, s.w);
  const cplx<float> zc = mk<float>(zm.x, -zm.y);
  const cplx<float> yu = zu * sp + zc * sm;
  const cplx<float> w2 = zc * sp + zu * sm;
  zu = yu;
  zm = mk<float>(w2.x, -w2.y);
#else
  const u64_t SP = pk2(s.x, s.y), SPs = pk2(s.y, s.x), SM = pk2(s.z, s.w), SMs = pk2(s.w, s.z);
  const u64_t zux = pk2(zu.x, zu.x), zuy = pk2(zu.y, zu.y);
  const u64_t zmx = pk2(zm.x, zm.x), zmy = pk2(zm.y, zm.y);
  // Y[k]: h = zu.y (py, px) - zm.y (my, mx);  yu = zu.x (px, py) + zm.x (mx, my) + (-h.x, h.y)
  const cplx<float> h1 = upk2(mul2(zmy, SMs));
  const cplx<float> h = upk2(fma2(zuy, SPs, pk2(-h1.x, -h1.y)));
  const u64_t g = fma2(zmx, SM, pk2(-h.x, h.y));
  const cplx<float> yu = upk2(fma2(zux, SP, g));
  // Y[-k]: b = zm.x (px, py) + zu.x (mx, my);  ym = zm.y (py, px) - zu.y (my, mx) + (b.x, -b.y)
  const cplx<float> b = upk2(fma2(zmx, SP, mul2(zux, SM)));
  const cplx<float> nms = upk2(SMs);
  const u64_t a1 = fma2(zuy, pk2(-nms.x, -nms.y), pk2(b.x, -b.y));
  const cplx<float> ym = upk2(fma2(zmy, SPs, a1));
  zu = yu;
  zm = ym;
#endif
}
// a self-mirrored element (k = -k: kx in {0, 64} and ky in {0, 64})
__device__ __forceinline__ void mirror_self(cplx<float> &z, float4 s) {
  cplx<float> zm = z;
  mirror_pair(z, zm, s);
}

// Hot pixels of the two walkers a CTA has in flight (the one whose inverse rows run and the
// next one, whose forward rows run in the same phase): positions from the prepare kernel,
// the values the render found there ((raw, w raw^2) of the whole pixel, which then enters
// the transform as zero).
struct HotState {
  int pos[2][PSFMC_MAX_COMPONENTS];
  float val[2][PSFMC_MAX_COMPONENTS][2];
  int any[2];
  int dead[2];   // a component's float32 constants are poisoned (prepare_kernel): -inf, final
};

// Per-thread constants of the row passes (4 rows per warp, 8 threads per row).
struct RowRole {
  int w, rr, l;
  unsigned t16;     // radix-16 side of the exchange layout: 8 l, k1 at + 144 (k1 >> 1) + 64 (k1 & 1)
  unsigned qa;      // radix-8 side: 144 l; k1 = 2 l at + 8 n2, k1 = 2 l + 1 at + 64 + 8 n2
  unsigned fa;      // column layout: 16 l; columns 2 l + 16 k2 and + 1 at + 128 k2
};

// render + forward row transform of row batch `it` of walker b
// (rc0 / der0: the walker's constants, STAGED = in shared memory, see fused_render16)
// PADDED: the observation frame is P.Hr x P.Wr in the corner of the 128 x 128 transform
// frame; nothing is rendered outside it.
// TILED: the tile is sub-image `sub` = 4 ry + rx of a 512 x 512 frame: tile pixel (y, x)
// is frame pixel (4 y + ry, 4 x + rx).
template <bool STAGED, bool PADDED = false, bool TILED = false, bool IMAGES = false>
__device__ __forceinline__ void fused_rows_forward(const FusedParams &P, smem_addr_t tile,
                                                   const RowRole &R, smem_addr_t twl,
                                                   const float *rc0, const double *der0,
                                                   int it, float wsc,
                                                   const FoldParams *F = nullptr, int sub = 0,
                                                   bool drain = false, HotState *hs = nullptr,
                                                   int hpar = 0, long long bw = 0) {
  // bw (IMAGES): the walker whose rows are rendered
  // drain (tiled forward half, first row batch of a job): the previous job's results are
  // still leaving the tile through bulk stores issued by threads 0..127; the tile may
  // only be written once those have read it (see the kernel)
  const int y = it * 64 + R.w * 4 + R.rr;
  const smem_addr_t rb = tile + (unsigned)y * PSFMC_FUSED_ROWB;
  if (PADDED && y - R.rr >= F->Hr) {
    // all four rows of the warp lie outside the observation frame: their row spectra
    // are zero (the tile still holds the previous walker's values there)
    __syncwarp();
    const cplx<float> zero = mk<float>(0.0f, 0.0f);
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) sts128(rb + R.fa + 8 * 16 * k2, zero, zero);
    return;
  }
  {
    cplx<float> v[16];
    if (!PADDED || y < F->Hr) {
      if (TILED)
        fused_render16<32, STAGED>(P, rc0, der0, 4 * y + (sub >> 2), 4 * R.l + (sub & 3), wsc,
                                   v);
      else
        fused_render16<8, STAGED, IMAGES>(P, rc0, der0, y, R.l, wsc, v);
      if (IMAGES && P.img_raw) {
        float *dst = P.img_raw + ((size_t)bw * PSFMC_FUSED_N + y) * PSFMC_FUSED_N + R.l;
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[8 * j] = v[j].x;
      }
      if (!PADDED && !TILED && __builtin_expect(hs != nullptr, 0)) {
        // take the walker's hot pixels out of the transform (see prepare_kernel): the
        // thread that rendered one records its value and zeroes it
#pragma unroll 1
        for (int c = 0; c < P.ncomp; ++c) {
          const int hp = hs->pos[hpar][c];
          if (hp >= 0 && (hp >> 16) == y && ((hp & 7) == R.l)) {
            const int jj = (hp & 0xffff) >> 3;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j == jj) {
                hs->val[hpar][c][0] = v[j].x;
                hs->val[hpar][c][1] = v[j].y;
                v[j] = mk<float>(0.0f, 0.0f);
              }
          }
        }
      }
      if (PADDED) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (R.l + 8 * j >= F->Wr) v[j] = mk<float>(0.0f, 0.0f);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = mk<float>(0.0f, 0.0f);
    }
    dft16<false>(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; k1 += 2) {
      const cplx2f w = lds128(twl + 16 * (k1 >> 1));
      if (k1 > 0) v[k1] = v[k1] * w.lo;
      v[k1 + 1] = v[k1 + 1] * w.hi;
    }
    if (TILED && drain) {
      if (threadIdx.x < PSFMC_FUSED_N) tile_store_wait_read();
      group_barrier(6, PSFMC_FUSED_THREADS);
    }
    __syncwarp();   // every lane is done reading this row (previous walker)
    const smem_addr_t rt = rb + R.t16;
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) sts64(rt + 144 * (k1 >> 1) + 64 * (k1 & 1), v[k1]);
  }
  __syncwarp();
  cplx<float> a[8], bb[8];
  {
    const smem_addr_t ra = rb + R.qa;
#pragma unroll
    for (int n2 = 0; n2 < 8; n2 += 2) {
      const cplx2f pa = lds128(ra + 8 * n2), pb = lds128(ra + 64 + 8 * n2);
      a[n2] = pa.lo;
      a[n2 + 1] = pa.hi;
      bb[n2] = pb.lo;
      bb[n2 + 1] = pb.hi;
    }
  }
  __syncwarp();
  dft8<float, false>(a);    // a[k2]  = Z[2 l + 16 k2]
  dft8<float, false>(bb);   // bb[k2] = Z[2 l + 1 + 16 k2]
  // the row spectrum of the packed row goes to the columns as it is (column kx at 8 kx);
  // the two images are taken apart only implicitly, by mirror_pair
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) sts128(rb + R.fa + 8 * 16 * k2, a[k2], bb[k2]);
}

// inverse row transform + chi-square terms of row batch `it`;
// returns this thread's float64 partial sum over its 16 pixels. PREFETCH: issue the
// observation loads before the transform (needs 32 registers for its duration).
template <bool PREFETCH, bool PADDED = false, bool IMAGES = false>
__device__ __forceinline__ double fused_rows_inverse(const FusedParams &P, smem_addr_t tile,
                                                     const RowRole &R, smem_addr_t twl,
                                                     int it, float unscale,
                                                     const FoldParams *F = nullptr, int sub = 0,
                                                     unsigned skip_quads = 0,
                                                     const cplx<float> *gnext = nullptr,
                                                     unsigned long long *bar = nullptr,
                                                     unsigned char *tile_ptr = nullptr,
                                                     const HotState *hs = nullptr, int hpar = 0,
                                                     const float2 *kpv = nullptr,
                                                     long long bw = 0) {
  // bw (IMAGES): the walker whose rows come back
  const int y = it * 64 + R.w * 4 + R.rr;
  // gnext (tiled inverse half): once the warp has taken its four rows out of the tile,
  // lanes 0..3 fetch the same rows of the NEXT job's sub-spectrum into their place
  auto fetch_next = [&]() {
    if (!gnext) return;
    __syncwarp();
    const int lane = threadIdx.x & 31;
    if (lane < 4) {
      async_proxy_fence();
      const int row = y - R.rr + lane;
      tile_row_fetch(tile_ptr + (size_t)row * PSFMC_FUSED_ROWB, gnext + row * PSFMC_FUSED_N, bar);
    }
  };
  if (PADDED && y - R.rr >= F->Hr) return 0.0;   // the warp's four rows are padding
  // ... or hold no unmasked pixel: nothing of them enters the sum (models.py:233-236)
  if ((skip_quads >> (it * 16 + R.w)) & 1u) {
    fetch_next();
    return 0.0;
  }
  const smem_addr_t rb = tile + (unsigned)y * PSFMC_FUSED_ROWB;
  // observation + signed variance of this thread's 16 pixels: issued first, used
  // last (L2 latency hidden behind the whole inverse transform)
  float2 o[16];
  // (sub-image tables of a tiled frame follow each other)
  // (pixel x = l + 8 j of a row sits at position 16 (j >> 1) + 2 l + (j & 1) of the row's
  // table, see fused_ow_index: a thread's pixels j = 2 i, 2 i + 1 are one 16-byte load, the
  // eight threads of a row read 128 contiguous bytes)
  const float4 *owr = reinterpret_cast<const float4 *>(
      P.ow + (sub * PSFMC_FUSED_N + y) * PSFMC_FUSED_N + 2 * R.l);
  const unsigned mw = __ldg(P.maskw + (sub * PSFMC_FUSED_N + y) * 8 + R.l);   // bit j: pixel
                                                                         // x = l + 8 j is good
  if (PREFETCH) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 q = __ldg(owr + 8 * i);
      o[2 * i] = make_float2(q.x, q.y);
      o[2 * i + 1] = make_float2(q.z, q.w);
    }
  }
  cplx<float> a[8], bb[8];   // a[k2] = Y[2 l + 16 k2], bb[k2] = Y[2 l + 1 + 16 k2]
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) {
    const cplx2f pr = lds128(rb + R.fa + 8 * 16 * k2);
    a[k2] = pr.lo;
    bb[k2] = pr.hi;
  }
  __syncwarp();
  dft8<float, true>(a);     // a[n2]
  dft8<float, true>(bb);
  {
    const smem_addr_t ra = rb + R.qa;
#pragma unroll
    for (int n2 = 0; n2 < 8; n2 += 2) {
      sts128(ra + 8 * n2, a[n2], a[n2 + 1]);
      sts128(ra + 64 + 8 * n2, bb[n2], bb[n2 + 1]);
    }
  }
  __syncwarp();
  cplx<float> v[16];
  {
    const smem_addr_t rt = rb + R.t16;
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) v[k1] = lds64(rt + 144 * (k1 >> 1) + 64 * (k1 & 1));
  }
  fetch_next();   // (the rows' last access to the tile was just made)
#pragma unroll
  for (int k1 = 0; k1 < 16; k1 += 2) {
    const cplx2f w = lds128(twl + 16 * (k1 >> 1));
    if (k1 > 0) v[k1] = cmul_conj(v[k1], w.lo);
    v[k1 + 1] = cmul_conj(v[k1 + 1], w.hi);
  }
  dft16<true>(v);           // v[j] = (convolved model, scaled model variance)
  if (PADDED) {
    // fold the linear convolution back modulo Wr (see Frame): through the row's own
    // 1 KB of the tile (free once every lane has taken its exchange values), pixel x at
    // position x
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) sts64(rb + 8u * (R.l + 8 * j), v[j]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int x = R.l + 8 * j;
      if (x <= F->fx_hi) v[j] = v[j] + lds64(rb + 8u * (x + F->Wr));
      if (x >= F->fx_lo && x < F->Wr)
        v[j] = v[j] + lds64(rb + 8u * (PSFMC_FUSED_N + x - F->Wr));
    }
    __syncwarp();
  }
  if (!PREFETCH) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 q = __ldg(owr + 8 * i);
      o[2 * i] = make_float2(q.x, q.y);
      o[2 * i + 1] = make_float2(q.z, q.w);
    }
  }
  if (!PADDED && __builtin_expect(hs != nullptr, 0)) {
    // the hot pixels' own contribution, convolved exactly: value x real-space kernel
#pragma unroll 1
    for (int c = 0; c < P.ncomp; ++c) {
      const int hp = hs->pos[hpar][c];
      if (hp < 0) continue;
      const float c1 = hs->val[hpar][c][0], c2 = hs->val[hpar][c][1];
      const int dy = (y - (hp >> 16)) & (PSFMC_FUSED_N - 1);
      const float2 *krow = kpv + dy * PSFMC_FUSED_N;
      const int x0 = R.l - (hp & 0xffff);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 k = __ldg(krow + ((x0 + 8 * j) & (PSFMC_FUSED_N - 1)));
        v[j].x = fmaf(c1, k.x, v[j].x);
        v[j].y = fmaf(c2, k.y, v[j].y);
      }
    }
  }
  if (IMAGES) {
    // convolved model, residual and composite IVM of EVERY pixel, with the staged
    // kernels' float32 expressions (Epilogue<float>::term)
    const size_t g0 = (size_t)y * PSFMC_FUSED_N + R.l, w0 = (size_t)bw * PSFMC_FUSED_N * PSFMC_FUSED_N;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const size_t g = g0 + 8 * j;
      const float conv = v[j].x, mvar = v[j].y * unscale;
      const float resid = __ldg(P.img_obs + g) - conv;
      const float ivm = fast_rcp(mvar + __ldg(P.img_ovar + g));
      if (P.img_conv) P.img_conv[w0 + g] = conv;
      if (P.img_resid) P.img_resid[w0 + g] = resid;
      if (P.img_ivm) P.img_ivm[w0 + g] = ivm;
    }
  }
  // Chi-square terms (psfMC/models.py:233-236) of the 16 pixels, float32, summed in
  // float32 over 8 pixels before they enter the float64 sum:
  //     resid^2 ivm - log(ivm / 2 pi) = resid^2 / tot + ln2 log2(tot) + ln(2 pi),
  // tot = model variance + observation variance; ln(2 pi) times the number of good pixels
  // is added once per walker (FusedParams::lnl_const). One packed FFMA2 gives (resid, tot);
  // masked pixels are left out by predication (bit j of the thread's mask word; their
  // table entry is (0, 1e30), so nothing non-finite arises from the observation there).
  // The float32 transform must not leave the convolved model variance below -2^-8 of the
  // pixel's own variance (see Epilogue<float>::term): the minimum of
  // tot - (1 - 2^-8) ovar over the pixels is checked once per thread.
  const cplx<float> cu = mk<float>(-1.0f, unscale);
  double acc = 0.0;
  float s1 = 0.0f, s2 = 0.0f, mn = 3.0e38f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const cplx<float> rt = pfma(v[j], cu, mk<float>(o[j].x, o[j].y));   // (resid, tot)
    const float ivm = fast_rcp(rt.y), l2 = fast_lg2(rt.y);
    const float r2 = rt.x * rt.x;
    mn = fminf(mn, fmaf(-(1.0f - PSFMC_VAR_NOISE_TOL), o[j].y, rt.y));
    if ((mw >> j) & 1u) {
      s1 = fmaf(r2, ivm, s1);
      s2 += l2;
    }
    if ((j & 7) == 7) {
      acc += (double)fmaf(0.69314718055994530942f, s2, s1);
      s1 = s2 = 0.0f;
    }
  }
  if (mn < 0.0f) acc = NAN;
  return acc;
}

// MODE 0: the whole lnL of a 128 x 128 frame. MODE 1 / 2: the forward / inverse half for
// the 4 x 4 sub-images of a 512 x 512 frame (kernels_tiled.cuh): a "walker" of the loop
// below is then a job = 16 * walker + sub-image; the forward half stops after the
// columns' forward radix-8 (its results go to P.sub_spec instead of the spectrum
// product), the inverse half starts there and ends with one partial sum per job.
#define PSFMC_MODE_FULL 0
#define PSFMC_MODE_FWD 1
#define PSFMC_MODE_INV 2
// IMAGES: the blob images of every walker are written out on the way (FusedParams::img_*).
template <bool PADDED, int MODE = PSFMC_MODE_FULL, bool IMAGES = false>
__global__ void __launch_bounds__(PSFMC_FUSED_THREADS, 1)
fused_lnlike_kernel(const FusedParams P, const FoldParams F) {
  constexpr bool TILED = MODE != PSFMC_MODE_FULL;
  PSFMC_DYN_SMEM(smem_raw);
  const smem_addr_t tile = smem_base(smem_raw);
  __shared__ double red_s[PSFMC_FUSED_THREADS / 32];
  __shared__ int cnt_s;
  constexpr int N = PSFMC_FUSED_N;
  constexpr unsigned ROWB = PSFMC_FUSED_ROWB;   // bytes per tile row
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) cnt_s = 0;

  // row-pass role
  RowRole R;
  R.w = w;
  R.rr = lane >> 3;
  R.l = lane & 7;
  {
    R.t16 = 8u * R.l;
    R.qa = 144u * R.l;
    R.fa = 16u * R.l;
  }
  // twiddles of the row passes W128^(l*k1) in shared memory as [l][k1] with a pitch of
  // 144 bytes: a thread takes its sixteen values with eight 128-bit loads (twl + 16 i =
  // k1 2i, 2i+1); the eight threads of a row hit the eight 16-byte bank groups
  __shared__ __align__(16) float tw_s[8][36];   // row l: 16 complex values + 16 bytes
  if (tid < 128) {
    const int k1 = tid & 15, ll = tid >> 4;
    tw_s[ll][2 * k1] = c_tw128[ll * 16 + k1][0];
    tw_s[ll][2 * k1 + 1] = c_tw128[ll * 16 + k1][1];
  }
  const smem_addr_t twl = smem_base(reinterpret_cast<unsigned char *>(&tw_s[0][0])) + 144u * R.l;
  // constants of the walker the next forward pass renders, staged one walker ahead
  // (during the column passes) so that the render does not wait for L2
  __shared__ __align__(16) float rc_s[PSFMC_MAX_COMPONENTS * PSFMC_RC_STRIDE];
  __shared__ double der_s[PSFMC_MAX_COMPONENTS * PSFMC_DERIVED_STRIDE];
  __shared__ float wsc_s;
  __shared__ HotState hot_s;
#ifdef PSFMC_HOT_COMPILED_OUT
  constexpr bool use_hot = false;
#else
  const bool use_hot = MODE == PSFMC_MODE_FULL && !PADDED && P.hot != nullptr;
#endif
  int hpar = 1;   // hot-state half of the walker whose inverse rows run (the next: hpar ^ 1)
  if (tid == 1) hot_s.any[1] = hot_s.dead[1] = 0;   // (half 0 is staged below)
  auto stage_params = [&](long long bs) {
    if (bs >= P.n_batch || MODE == PSFMC_MODE_INV) return;
    if (use_hot && tid >= 64 && tid < 64 + 32) {
      // one warp: the next walker's hot-pixel positions and whether its float32 constants
      // are poisoned (flags by warp votes)
      const int c = tid - 64, half = hpar ^ 1;
      const int hp = c < P.ncomp ? __ldg(P.hot + bs * P.ncomp + c) : -1;
      hot_s.pos[half][c] = hp;
      hot_s.val[half][c][0] = 0.0f;
      hot_s.val[half][c][1] = 0.0f;
      const int some = __any_sync(0xffffffffu, hp >= 0);
      const float c0 = c < P.ncomp ? __ldg(P.rconst + (bs * P.ncomp + c) * PSFMC_RC_STRIDE + 9)
                                   : 0.0f;
      const int dead = __any_sync(0xffffffffu, c0 != c0);
      if (c == 0) {
        hot_s.any[half] = some;
        hot_s.dead[half] = dead;
      }
    }
    if (TILED) bs >>= 4;   // job -> walker
    if (tid < P.ncomp * PSFMC_RC_STRIDE)
      rc_s[tid] = __ldg(P.rconst + bs * P.ncomp * PSFMC_RC_STRIDE + tid);
    for (int k = tid; k < P.ncomp * PSFMC_DERIVED_STRIDE; k += PSFMC_FUSED_THREADS)
      der_s[k] = __ldg(P.derived + bs * P.ncomp * PSFMC_DERIVED_STRIDE + k);
    if (tid == PSFMC_FUSED_THREADS - 1) wsc_s = (float)P.wscale[bs];
  };
  stage_params(blockIdx.x);
  __syncthreads();
  // tiled inverse half: the sub-spectrum of a job arrives in the tile by bulk copies whose
  // bytes are counted on this barrier (one phase per job)
  __shared__ __align__(8) unsigned long long tile_bar;
  unsigned tile_phase = 0;
  if (MODE == PSFMC_MODE_INV) {
#ifndef PSFMC_EMU
    if (tid == 0) mbar_init(&tile_bar, 1);
    __syncthreads();
    if (tid == 0 && (long long)blockIdx.x < P.n_batch)
      mbar_expect_tx(&tile_bar, N * PSFMC_SUBROW_BYTES);
    __syncthreads();
#endif
    // the first job's rows: warp w fetches the rows it will transform (4 w .. and 64 + 4 w ..)
    if ((long long)blockIdx.x < P.n_batch && lane < 8) {
      const int row = (lane >> 2) * 64 + 4 * w + (lane & 3);
      tile_row_fetch(smem_raw + (size_t)row * ROWB,
                     P.sub_spec + ((size_t)blockIdx.x * N + row) * N, &tile_bar);
    }
  }

  // Column-pass roles. The four warps of a column group (which meet at the named
  // barriers) sit on four DIFFERENT schedulers (warp w runs on scheduler w & 3), so
  // every scheduler hosts one warp of each group and its warps are not phase-locked
  // to each other. Group cg owns the 16 "slots" p = 16 cg .. 16 cg + 15: slot p is the
  // mirror pair of columns (p, 128 - p); slot 0 is the two self-mirrored columns (0, 64).
  //  * radix-16 sub-passes: thread = (column c, residues n2 = m and m + 4); lanes 0..15
  //    run over the slots' first columns, lanes 16..31 over their mirrors (two
  //    conflict-free half-warp wavefronts), m is warp-uniform (twiddles from the
  //    constant bank).
  //  * radix-8 sub-pass: thread = (slot p, m8 = 0..7), two rounds of two 8-point units
  //    chosen so that every element and its mirror (-ky, -kx) are in the same thread:
  //        m8 > 0:  round 0: a = (p, k1 = m8),      bb = (-p, k1 = 16 - m8)
  //                 round 1: a = (p, k1 = 16 - m8), bb = (-p, k1 = m8)
  //                 mirror of a[k2] (ky = k1 + 16 k2) is bb[7 - k2]
  //        m8 = 0:  round 0: a = (p, 0), bb = (-p, 0): mirror of a[k2] is bb[(8 - k2) & 7]
  //                 round 1: a = (p, 8), bb = (-p, 8): mirror of a[k2] is bb[7 - k2]
  //        slot 0:  round 0 works on column 0, round 1 on column 64 (each its own
  //                 mirror): a = (c, k1), bb = (c, -k1); for m8 = 0: a = (c, 0), bb = (c, 8)
  //                 with the mirrors inside a (k2 <-> 8 - k2) and inside bb (k2 <-> 7 - k2).
  const int cg = w >> 2, m = w & 3;
  const int c = lane < 16 ? 16 * cg + lane
                          : ((16 * cg + lane - 16) == 0 ? 64 : 128 - (16 * cg + lane - 16));
  const unsigned cev = 8u * c;   // byte offset of column c in a row
  const int slot = 16 * cg + (lane & 15), m8 = 2 * m + (lane >> 4);
  const bool slot0 = slot == 0;
  const bool zpat = slot0 && m8 == 0;          // one thread of the CTA
  const bool xpat = !slot0 && m8 == 0;         // lanes 0..15 of the warps with m = 0

  // Order of the row work between two column passes. Every warp owns two row
  // batches; per batch the inverse pass of walker b must precede the forward pass
  // of the next walker (same rows). Half of the warps of each scheduler interleave
  // (inv 0, fwd 0, inv 1, fwd 1), the other half run (inv 0, inv 1, fwd 0, fwd 1),
  // so the SFU-bound render of some warps overlaps the FMA/LSU-bound inverse
  // transforms of the others.
  const bool interleave = ((w >> 2) & 1) != 0;

  // The first pass through the loop (b < 0) only renders and forward-transforms the
  // CTA's first walker; every later pass runs the column passes of walker b, then
  // its inverse rows interleaved with the forward rows of the CTA's next walker.
#pragma unroll 1
  for (long long b = (long long)blockIdx.x -
                     (MODE == PSFMC_MODE_INV ? 0ll : (long long)gridDim.x);
       b < P.n_batch; b += gridDim.x) {
    const bool cur = b >= 0;
    const long long wb = TILED ? (b >> 4) : b;        // walker of this job
    const int sub = TILED ? (int)(b & 15) : 0;        // sub-image of this job
    int sel = (cur && MODE != PSFMC_MODE_FWD) ? P.psf_sel[wb] : 0;
    const bool invalid = sel < 0;
    if (invalid) sel = 0;
    const double wscale_b = (cur && MODE != PSFMC_MODE_FWD) ? P.wscale[wb] : 1.0;
    const float unscale = (float)(P.vscale_inv[sel] / wscale_b);
    if (cur) {

    __syncthreads();
    // every warp is past the forward rows of walker b: stage walker b + grid
    stage_params(b + gridDim.x);
    if (MODE == PSFMC_MODE_INV) {
#ifndef PSFMC_EMU
      // this job's sub-spectrum has landed in the tile; the next job's bytes are expected
      // on the barrier's next phase (its fetches are issued by the row passes below)
      mbar_wait(&tile_bar, tile_phase);
      tile_phase ^= 1u;
      if (tid == 0 && b + (long long)gridDim.x < P.n_batch)
        mbar_expect_tx(&tile_bar, N * PSFMC_SUBROW_BYTES);
#endif
    }

    // ------------------------------------------------- columns: radix-16 --
    // both residues n2 = m and m + 4 of this thread in flight at once (ILP)
    if (MODE != PSFMC_MODE_INV) {
      const smem_addr_t cb0 = tile + (unsigned)m * ROWB + cev;
      const smem_addr_t cb1 = cb0 + 4 * ROWB;
      cplx<float> v0[16], v1[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v0[j] = lds64(cb0 + 8 * j * ROWB);
#pragma unroll
      for (int j = 0; j < 16; ++j) v1[j] = lds64(cb1 + 8 * j * ROWB);
      dft16<false>(v0);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v0[k1] = v0[k1] * tw128(m, k1);
      dft16<false>(v1);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) sts64(cb0 + 8 * k1 * ROWB, v0[k1]);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v1[k1] = v1[k1] * tw128(m + 4, k1);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) sts64(cb1 + 8 * k1 * ROWB, v1[k1]);
    }
    // spectrum values of round 0 of the mirror-pair product, issued before the group
    // barrier so that their L2 latency overlaps the wait: one 16-byte load (S+, S-) per
    // pair of elements
    const float4 *sp4 = P.spec4 + (size_t)sel * N * 64 + slot;
    const float4 *spx = P.specx4 + (size_t)sel * N;
    int kA = m8, kB = (16 - m8) & 15;          // round 0 (m8 = 0: both 0)
    if (zpat) kB = 8;
    float4 S[8];
    if (!TILED) {
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) S[k2] = __ldg(sp4 + (kA + 16 * k2) * 64);
    }
    if (MODE != PSFMC_MODE_INV) group_barrier(1 + cg, 128);
    // this job's spectrum tile in global memory (tiled frames)
    cplx<float> *gsub = TILED ? P.sub_spec + (size_t)b * N * N : nullptr;
    // Tiled halves keep the 2 x 16 values of both rounds in registers: the inverse half
    // finds its input in the tile in FREQUENCY order (row ky, fetched by TMA) and writes
    // rows 8 k1 + n2 -- the forward half the other way round -- so within a column group
    // every read has to precede every write (one more group barrier).
    cplx<float> ga[2][8], gb[2][8];
    if (MODE == PSFMC_MODE_INV) {
      int ka = kA, kb = kB;
#pragma unroll
      for (int round = 0; round < 2; ++round) {
        const int ca = slot0 ? 64 * round : slot;
        const int cb = slot0 ? 64 * round : 128 - slot;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
          ga[round][k2] = lds64(tile + (unsigned)(ka + 16 * k2) * ROWB + 8u * ca);
          gb[round][k2] = lds64(tile + (unsigned)(kb + 16 * k2) * ROWB + 8u * cb);
        }
        int ka1 = kb, kb1 = ka;
        if (m8 == 0) ka1 = kb1 = 8;
        if (zpat) ka1 = 0;
        ka = ka1;
        kb = kb1;
      }
      group_barrier(1 + cg, 128);
    }

    // ---------- columns: radix-8, mirror-pair spectrum product, inverse radix-8 --
#pragma unroll
    for (int round = 0; round < 2; ++round) {
      // columns of the two units: (slot, mirror), slot 0: column 0, then column 64
      const int ca = slot0 ? 64 * round : slot;
      const int cb = slot0 ? 64 * round : 128 - slot;
      const smem_addr_t ba = tile + 8u * kA * ROWB, bq = tile + 8u * kB * ROWB;
      const unsigned aev = 8u * ca, bev = 8u * cb;
      cplx<float> a[8], bb[8];
      if (MODE != PSFMC_MODE_INV) {
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
          a[n2] = lds64(ba + (unsigned)n2 * ROWB + aev);
          bb[n2] = lds64(bq + (unsigned)n2 * ROWB + bev);
        }
        dft8<float, false>(a);    // a[k2]  = U[kA + 16 k2][ca]
        dft8<float, false>(bb);   // bb[k2] = U[kB + 16 k2][cb]
      }
      if (MODE == PSFMC_MODE_FWD) {
        // forward half: the column spectrum is kept until the group has read all its input
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
          ga[round][k2] = a[k2];
          gb[round][k2] = bb[k2];
        }
      } else if (MODE == PSFMC_MODE_INV) {
        // inverse half: ... and comes back multiplied by the PSF spectra
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
          a[k2] = ga[round][k2];
          bb[k2] = gb[round][k2];
        }
      } else if (zpat) {
        // columns 0 / 64, k1 = 0 (a) and 8 (bb): every mirror is in the same array
        mirror_self(a[0], S[0]);
        mirror_self(a[4], S[4]);
#pragma unroll
        for (int k2 = 1; k2 < 4; ++k2) mirror_pair(a[k2], a[8 - k2], S[k2]);
        const float4 *s8 = round == 0 ? sp4 : spx;
        const int st = round == 0 ? 64 : 1;
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2)
          mirror_pair(bb[k2], bb[7 - k2], __ldg(s8 + (8 + 16 * k2) * st));
      } else if (xpat && round == 0) {
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) mirror_pair(a[k2], bb[(8 - k2) & 7], S[k2]);
      } else {
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) mirror_pair(a[k2], bb[7 - k2], S[k2]);
      }
      int kA1 = kB, kB1 = kA;                  // round 1
      if (m8 == 0) kA1 = kB1 = 8;
      if (zpat) kA1 = 0;
      if (round == 0 && !TILED) {   // prefetch round 1's spectrum values
        const float4 *s1 = slot0 ? spx : sp4;
        const int st = slot0 ? 1 : 64;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) S[k2] = __ldg(s1 + (kA1 + 16 * k2) * st);
      }
      if (MODE != PSFMC_MODE_FWD) {
        dft8<float, true>(a);     // a[n2]: inverse over k2
        dft8<float, true>(bb);
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
          sts64(ba + (unsigned)n2 * ROWB + aev, a[n2]);
          sts64(bq + (unsigned)n2 * ROWB + bev, bb[n2]);
        }
      }
      kA = kA1;
      kB = kB1;
    }
    if (MODE != PSFMC_MODE_FWD) group_barrier(1 + cg, 128);
    if (MODE == PSFMC_MODE_FWD) {
      // forward half: the column spectrum goes back into the tile in frequency order (row
      // ky, column kx) and leaves for the 4 x 4 combine kernel row by row (bulk stores)
      group_barrier(1 + cg, 128);
      int ka = m8, kb = zpat ? 8 : (16 - m8) & 15;
#pragma unroll
      for (int round = 0; round < 2; ++round) {
        const int ca = slot0 ? 64 * round : slot;
        const int cb = slot0 ? 64 * round : 128 - slot;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
          sts64(tile + (unsigned)(ka + 16 * k2) * ROWB + 8u * ca, ga[round][k2]);
          sts64(tile + (unsigned)(kb + 16 * k2) * ROWB + 8u * cb, gb[round][k2]);
        }
        int ka1 = kb, kb1 = ka;
        if (m8 == 0) ka1 = kb1 = 8;
        if (zpat) ka1 = 0;
        ka = ka1;
        kb = kb1;
      }
      async_proxy_fence();
    }

    // ----------------------------------------- columns: inverse radix-16 --
    if (MODE != PSFMC_MODE_FWD) {
      const smem_addr_t cb0 = tile + (unsigned)m * ROWB + cev;
      const smem_addr_t cb1 = cb0 + 4 * ROWB;
      cplx<float> v0[16], v1[16];
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) v0[k1] = lds64(cb0 + 8 * k1 * ROWB);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) v1[k1] = lds64(cb1 + 8 * k1 * ROWB);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v0[k1] = cmul_conj(v0[k1], tw128(m, k1));
      dft16<true>(v0);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v1[k1] = cmul_conj(v1[k1], tw128(m + 4, k1));
#pragma unroll
      for (int j = 0; j < 16; ++j) sts64(cb0 + 8 * j * ROWB, v0[j]);
      dft16<true>(v1);
#pragma unroll
      for (int j = 0; j < 16; ++j) sts64(cb1 + 8 * j * ROWB, v1[j]);
    }
    __syncthreads();
    if (MODE == PSFMC_MODE_FWD && tid < N) {
      tile_row_store(gsub + (size_t)tid * N, smem_raw + (size_t)tid * ROWB);
      tile_store_commit();
    }
    if (PADDED) {
      // fold the linear convolution back modulo Hr (see Frame): row p receives rows
      // p + Hr (p <= fy_hi) and 128 + p - Hr (p >= fy_lo); the sources lie at or beyond
      // row Hr, the targets below it.
      for (int e = tid; e < F.Hr * N; e += PSFMC_FUSED_THREADS) {
        const int p = e >> 7, cc = e & (N - 1);
        const bool hi = p <= F.fy_hi, lo = p >= F.fy_lo;
        if (hi || lo) {
          const smem_addr_t dst = tile + (unsigned)p * ROWB + 8u * cc;
          cplx<float> acc = lds64(dst);
          if (hi) {
            const int q = p + F.Hr;
            acc = acc + lds64(tile + (unsigned)q * ROWB + 8u * cc);
          }
          if (lo) {
            const int q = N + p - F.Hr;
            acc = acc + lds64(tile + (unsigned)q * ROWB + 8u * cc);
          }
          sts64(dst, acc);
        }
      }
      __syncthreads();
    }
    }  // if (cur)

    // ------ rows: inverse + chi-square of walker b, render + forward of the next --
    const long long bn = b + gridDim.x;
    const bool has_next = bn < P.n_batch;
    // (hot pixels are rare: the row passes get a null pointer unless this walker / the next
    // one has any)
    HotState *hot_cur = (use_hot && cur && hot_s.any[hpar]) ? &hot_s : nullptr;
    HotState *hot_nxt = (use_hot && has_next && hot_s.any[hpar ^ 1]) ? &hot_s : nullptr;
    const float wsc_next = has_next ? wsc_s : 0.0f;
    double acc = 0.0;
#pragma unroll 1
    for (int step = 0; step < 4; ++step) {
      // interleave: I0 F0 I1 F1 ; otherwise: I0 I1 F0 F1
      const bool fwd = interleave ? (step & 1) : (step >= 2);
      const int it = interleave ? (step >> 1) : (step & 1);
      if (!fwd) {
        if (cur && MODE != PSFMC_MODE_FWD)
          acc += fused_rows_inverse<true, PADDED, IMAGES>(
              P, tile, R, twl, it, unscale, &F, sub,
              MODE == PSFMC_MODE_INV ? __ldg(P.skip_tab + sub) : P.skip_quads,
              (MODE == PSFMC_MODE_INV && has_next) ? P.sub_spec + (size_t)bn * N * N : nullptr,
              &tile_bar, smem_raw, hot_cur, hpar, P.kpv + (size_t)sel * N * N, b);
      } else if (has_next && MODE != PSFMC_MODE_INV) {
        fused_rows_forward<true, PADDED, TILED, IMAGES>(
            P, tile, R, twl, rc_s, der_s, it, wsc_next, &F, TILED ? (int)(bn & 15) : 0,
            MODE == PSFMC_MODE_FWD && cur && it == 0, hot_nxt, hpar ^ 1, bn);
      }
      if (cur && MODE != PSFMC_MODE_FWD &&
          ((interleave && step == 2) || (!interleave && step == 1))) {
        // both inverse batches of this warp are done: float64 reduction. Warp
        // shuffles, then the last warp to arrive sums the per-warp partials in
        // fixed order (deterministic) and writes lnL.
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
          acc += __shfl_down_sync(0xffffffffu, acc, off);
        if (lane == 0) {
          volatile double *red = red_s;
          red[w] = acc;
          __threadfence_block();
          const int prev = atomicAdd(&cnt_s, 1);
          if (prev == PSFMC_FUSED_THREADS / 32 - 1) {
            __threadfence_block();
            double tot = 0.0;
            for (int k = 0; k < PSFMC_FUSED_THREADS / 32; ++k) tot += red[k];
            if (MODE == PSFMC_MODE_INV) {
              // one partial per job; finalize_kernel sums the 16 of a walker in fixed
              // order (the constant term rides on sub-image 0)
              P.partials[b] = sub == 0 ? tot + P.lnl_const : tot;
            } else {
              double val = -0.5 * (tot + P.lnl_const);
              if (!isfinite(val)) val = P.nan_marks ? NAN : -INFINITY;
              if (invalid || (use_hot && hot_s.dead[hpar])) val = -INFINITY;
              if (P.n_peer > 0) {
#pragma unroll
                for (int p = 0; p < 16; ++p)
                  if (p < P.n_peer) P.lnl_peer[p][b] = val;
              } else {
                P.lnl[b] = val;
              }
            }
            cnt_s = 0;
          }
        }
      }
    }
    hpar ^= 1;   // the next walker becomes the current one
  }
#ifndef PSFMC_EMU
  if (MODE == PSFMC_MODE_FWD && tid < N)
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the last job's rows
#endif
}

// -------------------------------------------------------------- host side --

// (the TRANSFORM frame must be 128 x 128: the observation frame itself, or any smaller
// frame whose padded transform frame is -- every frame with height + psf_height - 1 <=
// 128 and width + psf_width - 1 <= 128 that is not a power of two)
template <typename T>
inline bool fused_path_available(const StagedPlan &plan, const Program &) {
  return sizeof(T) == 4 && plan.fr.H == PSFMC_FUSED_N && plan.fr.W == PSFMC_FUSED_N;
}

template <typename T>
inline int fused_prepare_device(const StagedPlan &) {
#ifndef PSFMC_EMU
  if (cudaFuncSetAttribute(fused_lnlike_kernel<false>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           PSFMC_FUSED_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(fused_lnlike_kernel<true>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           PSFMC_FUSED_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(fused_lnlike_kernel<false, PSFMC_MODE_FULL, true>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           PSFMC_FUSED_SMEM) != cudaSuccess)
    return 1;
#endif
  return 0;
}

// Re-layout of the float64 spectra [K][2*Wc][H] (column-major, see
// kernels_staged.cuh; P = PSF, V = PSF variance, normalisation and ifftshift sign
// folded in) for the mirror-pair product of the fused kernel:
//   spec4 [K][ky][slot]: (S+, S-) = (P +- v V) / 2 at (ky, kx = slot), slot = 0..63
//   specx4[K][ky]      : the same at kx = 64
// `vscale[k]` = v multiplies the V channel (power of two, undone in the epilogue).
// The mirrored element (-ky, -kx) uses the complex conjugates (P and V are spectra of
// real images); at the four self-mirrored elements (kx, ky in {0, 64}) the spectra are
// real, their imaginary rounding residue is dropped like a c2r transform drops it.
inline void fused_spectrum_layout(const cplx<double> *spec64, int n_psf,
                                  const double *vscale, float4 *spec4, float4 *specx4) {
  constexpr int N = PSFMC_FUSED_N, Wc = N / 2 + 1;
  for (int k = 0; k < n_psf; ++k) {
    const cplx<double> *src = spec64 + (size_t)k * 2 * Wc * N;
    for (int ky = 0; ky < N; ++ky)
      for (int kx = 0; kx <= 64; ++kx) {
        const cplx<double> &pp = src[((size_t)0 * Wc + kx) * N + ky];
        const cplx<double> &vv = src[((size_t)1 * Wc + kx) * N + ky];
        const bool self = (kx == 0 || kx == 64) && (ky == 0 || ky == 64);
        const double vy = self ? 0.0 : vv.y, py = self ? 0.0 : pp.y;
        float4 o;
        o.x = (float)(0.5 * (pp.x + vscale[k] * vv.x));
        o.y = (float)(0.5 * (py + vscale[k] * vy));
        o.z = (float)(0.5 * (pp.x - vscale[k] * vv.x));
        o.w = (float)(0.5 * (py - vscale[k] * vy));
        if (kx < 64)
          spec4[((size_t)k * N + ky) * 64 + kx] = o;
        else
          specx4[(size_t)k * N + ky] = o;
      }
  }
}

// position of pixel x in a row of the (obs, ovar) table of the row passes' epilogue
inline size_t fused_ow_index(size_t x) {
  const size_t l = x & 7, j = x >> 3;
  return 16 * (j >> 1) + 2 * l + (j & 1);
}

struct FusedBuffers {
  float *rconst = nullptr;
  const float4 *spec4 = nullptr, *specx4 = nullptr;
  const float2 *ow = nullptr;
  const unsigned short *maskw = nullptr;
  double lnl_const = 0.0;
  int n_sms = 148;
  double *const *lnl_peer = nullptr;   // [n_peer] destinations of the results, or null
  int n_peer = 0;
  int *hot = nullptr;                  // [B][ncomp] hot-pixel flags (prepare kernel), or null
  const float2 *kpv = nullptr;         // [K][128][128] real-space kernels
  bool nan_marks = false;              // see FusedParams
  unsigned skip_quads = 0;   // see FusedParams
};

// Blob images wanted from a launch (unpadded 128 x 128 frames): [n_batch][128 * 128] float
// each, null = not wanted; obs / ovar: the staged path's observation arrays.
struct FusedImages {
  float *raw = nullptr, *conv = nullptr, *resid = nullptr, *ivm = nullptr;
  const float *obs = nullptr, *ovar = nullptr;
  bool ps_only = false;
};

// theta -> lnL for n_batch walkers: prepare kernel + one persistent fused kernel.
// Returns the number of kernels launched.
template <typename T>
inline int launch_fused_lnlike(const StagedPlan &plan, const StagedBuffers<T> &buf,
                               const FusedBuffers &fb, const Program &prog_h,
                               const double *theta, long long n_batch, long long ld,
                               double *lnl, cudaStream_t stream,
                               cudaEvent_t ev_begin = nullptr, cudaEvent_t ev_end = nullptr,
                               const FusedImages *img = nullptr) {
  if (n_batch <= 0) return 0;
  const int ncomp = prog_h.n_components;
  // (images: every pixel of every walker is wanted as the float32 kernels compute it --
  // no hot pixels taken out, no row skipped)
  int *hot = (fb.hot && fb.kpv && !plan.fr.padded && !img) ? fb.hot : nullptr;
  launch_prepare(*buf.prog_host, theta, n_batch, ld, plan.fr.Hr, plan.fr.Wr, ncomp, buf.derived,
                 buf.psf_sel, buf.wscale, fb.rconst, stream, hot);
  FusedParams P;
  FoldParams F;
  F.Hr = plan.fr.Hr;
  F.Wr = plan.fr.Wr;
  F.fy_hi = plan.fr.fy_hi;
  F.fy_lo = plan.fr.fy_lo;
  F.fx_hi = plan.fr.fx_hi;
  F.fx_lo = plan.fr.fx_lo;
  P.rconst = fb.rconst;
  P.derived = buf.derived;
  P.wscale = buf.wscale;
  P.psf_sel = buf.psf_sel;
  P.vscale_inv = buf.vscale_inv;
  P.spec4 = fb.spec4;
  P.specx4 = fb.specx4;
  P.ow = fb.ow;
  P.maskw = fb.maskw;
  P.lnl_const = fb.lnl_const;
  P.hot = hot;
  P.kpv = fb.kpv;
  P.nan_marks = (fb.nan_marks && hot) ? 1 : 0;
  P.sub_spec = nullptr;
  P.partials = nullptr;
  P.skip_tab = nullptr;
  P.n_peer = fb.n_peer;
  for (int p = 0; p < 16; ++p) P.lnl_peer[p] = p < fb.n_peer ? fb.lnl_peer[p] : nullptr;
  P.lnl = lnl;
  P.n_batch = n_batch;
  P.ncomp = ncomp;
  P.skip_quads = img ? 0u : fb.skip_quads;
  P.img_raw = img ? img->raw : nullptr;
  P.img_conv = img ? img->conv : nullptr;
  P.img_resid = img ? img->resid : nullptr;
  P.img_ivm = img ? img->ivm : nullptr;
  P.img_obs = img ? img->obs : nullptr;
  P.img_ovar = img ? img->ovar : nullptr;
  P.ps_only = (img && img->ps_only) ? 1 : 0;
  P.kind_bits = 0;
  for (int c = 0; c < ncomp; ++c) P.kind_bits |= (unsigned long long)(prog_h.kind[c] & 3) << (2 * c);
  unsigned grid = (unsigned)(n_batch < fb.n_sms ? n_batch : fb.n_sms);
  if (ev_begin) cudaEventRecord(ev_begin, stream);
  if (img)
    launch_kernel(fused_lnlike_kernel<false, PSFMC_MODE_FULL, true>, dim3(grid),
                  dim3(PSFMC_FUSED_THREADS), (size_t)PSFMC_FUSED_SMEM, stream, P, F);
  else if (plan.fr.padded)
    launch_kernel(fused_lnlike_kernel<true>, dim3(grid), dim3(PSFMC_FUSED_THREADS),
                  (size_t)PSFMC_FUSED_SMEM, stream, P, F);
  else
    launch_kernel(fused_lnlike_kernel<false>, dim3(grid), dim3(PSFMC_FUSED_THREADS),
                  (size_t)PSFMC_FUSED_SMEM, stream, P, F);
  if (ev_end) cudaEventRecord(ev_end, stream);
  return 2;
}

}  // namespace psfmc
