// Kernel family (1): the 1-loop mode-coupling integrals as dense quadratures over the
// extrapolated log-k power spectrum (replaces redTime.cc:740-1282).
//
//   k_extrap           Pab extrapolation + window (rt:181-232, 772-778)        HBM/latency bound
//   k_bilinear         J_n(k_i;A,B) = sum_{q1,q2} a(q1) b(q2) T_n[i-q1][i-q2]   FP64 pipe bound (DMMA)
//   (compact_mask)     list of the cosmologies a launch is masked in for, built inside k_extrap
//   k_jlo              J_0 at the low-k row nloMR (rt:1252,1267-1272)
//   k_pz               P13-type log-convolutions PZ_n (rt:689-727)
//   k_assemble         A_{acd,bef}, R^l_{abc}, P_T,jm, P_MR,n from the table of assembly_table.cc
//
// k_bilinear design (sm_100a): see the comment above the kernel -- per (kernel, row block of 8 rows,
// slot) a dense matrix product T' x Hankel(spectrum) on the FP64 pipe as DMMA.8x8x4, T stored as 8 x 8
// tiles in operand order and streamed from L2, the spectra staged in shared memory by TMA bulk
// copies, followed by the alpha-side dot products.  Only the (kernel, spectrum) sets the requested
// output groups consume are computed; one set is nk (2 nsup^2 + 6 nsup) FLOP useful, 2 nk NV^2
// executed (NV = nsup + 7 rounded up to whole 8-lag tiles).
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "rtrg_device.h"
#include "stage_device.h"

namespace rtrg {

// ---------------------------------------------------------------------------- helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                             unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// ---------------------------------------------------------------------------- k_extrap
// act[0..*nact) = the unmasked cosmologies of a launch in ascending order: the slot list of
// k_bilinear.  One warp (ballot + popc prefix); done by the first warp of k_extrap's first block,
// which saves a launch per evaluation.
__device__ __forceinline__ void compact_mask(const int *__restrict__ mask, int B, int *__restrict__ act,
                                             int *__restrict__ nact, int lane) {
  int off = 0;
  for (int base = 0; base < B; base += 32) {
    const int b = base + lane;
    const int m = (b < B) && (!mask || mask[b]);
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    if (m) act[off + __popc(bal & ((1u << lane) - 1u))] = b;
    off += __popc(bal);
  }
  if (lane == 0) *nact = off;
}
// y: state vectors [B][ystride]; only the three ln P components are read.
__global__ void k_extrap(IntegralTabs tb, const Cosmo *__restrict__ cosmo,
                         const double *__restrict__ y, long long ystride,
                         double *__restrict__ P3, double *__restrict__ Prev,
                         const int *__restrict__ mask, long long *__restrict__ matvecs, int units,
                         int *__restrict__ act, int *__restrict__ nact) {
  if (act && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 32) compact_mask(mask, gridDim.y, act, nact, threadIdx.x);
  const int b = blockIdx.y;
  if (mask && !mask[b]) return;
  const int ip = blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= tb.np) return;
  if (ip == 0 && matvecs) matvecs[b] += units;  // one launch at a time per stream: no race
  const double *yb = y + (long long)b * ystride;
  const int n0 = tb.ex_n0[ip];
  const double w0 = tb.ex_w[4 * ip], w1 = tb.ex_w[4 * ip + 1], w2 = tb.ex_w[4 * ip + 2],
               w3 = tb.ex_w[4 * ip + 3];
  const double tail = (cosmo[b].ns - 3.0) * tb.ex_dx[ip];
  const double win = tb.WP[ip], kk = tb.kpad[ip], k2 = kk * kk;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const double *f = yb + c * tb.nk + n0;
    double lnP = tail;
    RT_ASSERT((w0 == 0.0 || (n0 >= 0 && n0 < tb.nk)) && (w3 == 0.0 || (n0 + 3 >= 0 && n0 + 3 < tb.nk)));
    RT_ASSERT((w1 == 0.0 || (n0 + 1 >= 0 && n0 + 1 < tb.nk)) && (w2 == 0.0 || (n0 + 2 >= 0 && n0 + 2 < tb.nk)));
    if (w0 != 0.0) lnP += w0 * f[0];
    if (w1 != 0.0) lnP += w1 * f[1];
    if (w2 != 0.0) lnP += w2 * f[2];
    if (w3 != 0.0) lnP += w3 * f[3];
    const double P = (win > 0.0) ? exp(lnP) * win : 0.0;
    P3[((long long)b * 3 + c) * tb.np + ip] = P;
    if (ip >= tb.jlo) {
      RT_ASSERT((BIL_R - 1) + (tb.np - 1 - ip) < tb.LP);
      Prev[((long long)b * 3 + c) * tb.LP + (BIL_R - 1) + (tb.np - 1 - ip)] = P * k2;
    }
  }
}

// One work item of a k_bilinear launch: kernel n applied to the beta-side spectra cd[0..ncd).
// Only the (kernel, spectrum) combinations the requested outputs consume are computed
// (assembly_needs): e.g. the RHS needs A and R = all of J but only 3 of the 7 Jn0 kernels,
// the default output columns need P_T,jm = 12 kernels x 2 spectra.  replicate: the three
// spectra are identical (1-loop cache at z1l, rt:1303-1305), one product serves all 9 pairs.
struct BilItem {
  short n, ncd, cd[3];
  short tn;    // table the matrix-vector product runs on: n, or the transposed copy of T_n
  short swap;  // 1: transposed -- cd[] are alpha-side spectra, the dot products run over the beta side
  short abmask;  // spectra of the dot-product side the requested outputs consume (bit a)
};
struct BilLaunch {
  BilItem it[N_JKERN];
  int replicate;
};

// ---------------------------------------------------------------------------- k_bilinear
// For one kernel n, one row block (R = 8 consecutive output rows, first row i0) and one slot (a
// cosmology with one beta-side spectrum b) the work is a small dense matrix product,
//   S[u][r] = sum_v T'[u][v] W[v][r],   T'[u][v] = T_n[i0 + u][i0 + v],  W[v][r] = b_rev[v - r + 7],
// over the NV alpha-side lags u and beta-side lags v the windowed spectra reach, followed by the
// alpha-side dot products out[ab][r] = sum_u a_rev,ab[u - r + 7] S[u][r].  W is a Hankel matrix of the
// spectrum, so nothing but the spectrum itself is staged.  The product runs on the FP64 pipe as
// DMMA.8x8x4 (mma.sync.m8n8k4.f64): the same 64 FMA/clk/SM units as DFMA, but 256 FMAs per issued
// instruction with 4 operand registers -- measured 37.2 TFLOP/s = 99.9 % of 148 SM x 64 x 2 x 1.965 GHz
// against 34.3 for a register-resident DFMA loop and 25.2 for the best DFMA form of this kernel
// (tools/dmma_probe.cu, profiles/r02_bilinear_experiments.txt).
//
// A warp owns MT = 4 tiles of 8 alpha-side lags and NS = 3 slots: 12 accumulator tiles (8 x 8, two
// doubles per lane).  Per chunk of 8 beta-side lags it issues 4 LDG.128 -- the T tiles are stored in
// operand order, a warp reads 512 contiguous bytes per tile, prefetched one chunk ahead --, 2 NS
// LDS.64 for the B fragments (lane (g, t): b_rev[v0 + 2 t - g + 7 + h]) and 8 NS DMMAs.  Every T
// element fetched from L2 feeds 8 rows x NS slots FMAs.  After the last chunk the lane's 2 x 3 NS
// alpha-side products are formed as independent DFMA chains and reduced over the 8 lane groups by
// recursive halving, then over the warps of the CTA through shared memory in a fixed pairwise order
// (the scalar FP64 instructions of this part queue behind the DMMAs of the warps still in their main
// loops, so the depth of its dependency chains is what counts: 7 instead of ~60 for the naive form).
//
// Work distribution.  The (row block, lag tile) pairs of ALL row blocks lie on one axis, item = rb NV +
// lag, cut into CTAs of IPC = 32 NWARP consecutive items (NWARP warps of 4 tiles).  A CTA touches up
// to three row blocks (a warp at most two); that costs nothing inside the main loop because the B
// fragments do not depend on the row block, and the epilogue reduces the lags of each row block
// separately.  The item axis is global (independent of k-sharding), so a sharded run adds the same
// numbers in the same order as the unsharded one; per-slot arithmetic does not depend on which
// slots share a CTA.  The slots of a CTA are consecutive entries of {active cosmologies} x {spectra
// of the item}; act[0..*nact) lists the unmasked cosmologies in ascending order (compact_mask).
// Grid: (CTAs along the item axis x v_split, items of the launch, slot groups).
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

template <int NWARP, int MT, int NS>
__device__ __forceinline__ void bil_body(const IntegralTabs &tb, const Cosmo *__restrict__ cosmo, const double *__restrict__ Prev,
               double *__restrict__ Jpart, const BilLaunch &L, int rb_lo, int rb_hi, int c_lo,
               const int *__restrict__ act, const int *__restrict__ nact) {
  constexpr int R = BIL_R, VC = 8;                  // rows per row block; lags per chunk
  constexpr int IPC = NWARP * MT * 8;               // items (alpha-side lags) per CTA
  constexpr int NSL = 3 * NS;                       // (slot, ab) sums per row
  static_assert(R == 8, "DMMA tiles are 8 x 8");
  const BilItem item = L.it[blockIdx.y];
  const int n = item.n, ncd = item.ncd;
  const int nslots = (*nact) * ncd;
  const int slot0 = NS * blockIdx.z;
  if (slot0 >= nslots) return;
  // blockIdx.x = (CTA along the item axis, split of the beta-side lags)
  const int cx = blockIdx.x / tb.vsplit, vs = blockIdx.x - cx * tb.vsplit;
  const int c = c_lo + cx;
  extern __shared__ __align__(128) double sm[];
  double *s_a = sm;                     // [NS][3][LP]: the three spectra of each slot's cosmology
  double *s_red = sm + NS * 3 * tb.LP;  // [NWARP][2][NSL R]
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ int s_wrb[NWARP][2];       // row block(s) the tiles of each warp belong to
  __shared__ int s_e[NS], s_cd[NS], s_ok[NS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;  // DMMA fragment coordinates of this lane
  const int LP = tb.LP;
  const uint32_t bytes = 3u * (uint32_t)LP * 8u;

  int e_q[NS], cd_q[NS];
#pragma unroll
  for (int q = 0; q < NS; q++) {
    const int sl = min(slot0 + q, nslots - 1);
    e_q[q] = act[sl / ncd];
    cd_q[q] = item.cd[sl - (sl / ncd) * ncd];
    if (tid == 0) s_e[q] = e_q[q], s_cd[q] = cd_q[q], s_ok[q] = (slot0 + q) < nslots;
  }

  if (tid == 0) mbar_init(&mbar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&mbar, NS * bytes);
#pragma unroll
    for (int q = 0; q < NS; q++) tma_bulk_g2s(s_a + q * 3 * LP, Prev + (long long)e_q[q] * 3 * LP, bytes, &mbar);
  }

  // the warp's tiles: global tile index -> (row block, first lag of the tile)
  const int NVT = tb.NV >> 3, NUT = tb.ldT >> 3;
  const int tile0 = (c * IPC + warp * (MT * 8)) >> 3;
  const size_t cstride = (size_t)NUT * 32;  // one chunk of beta-side lags further, in 16-byte elements
  int rb_m[MT], tu_m[MT];
  const double2 *Tp[MT];
#pragma unroll
  for (int m = 0; m < MT; m++) {
    const int gt = tile0 + m;
    rb_m[m] = gt / NVT;
    tu_m[m] = (gt - rb_m[m] * NVT) << 3;
    const int rbc = min(max(rb_m[m], rb_lo), rb_hi - 1);  // idle tiles stream a valid column, are never added
    // tile (v''/8, u''/8) with u'' = 8 rbc + tu, v'' = 8 rbc + v
    Tp[m] = reinterpret_cast<const double2 *>(tb.Tc) +
            (((size_t)item.tn * (tb.NUp >> 3) + rbc) * NUT + rbc + (tu_m[m] >> 3)) * 32 + lane;
    RT_ASSERT(n >= 0 && n < N_JKERN && 8 * rbc + tu_m[m] + 8 <= tb.ldT && 8 * rbc + tb.NVp <= tb.NUp);
    RT_ASSERT((((size_t)item.tn * (tb.NUp >> 3) + rbc + (tb.NVp >> 3) - 1) * NUT + rbc + (tu_m[m] >> 3)) * 64 + 63 < (size_t)tb.n_Tc);
  }
  const double *s_c[NS];
#pragma unroll
  for (int q = 0; q < NS; q++) {
    s_c[q] = s_a + (q * 3 + cd_q[q]) * LP + (R - 1) + 2 * t - g;  // B fragment of this lane at lag 0
    RT_ASSERT(e_q[q] >= 0 && cd_q[q] >= 0 && cd_q[q] < 3);
  }
  RT_ASSERT(tb.NVp + R <= LP && tb.NV <= tb.NVp);

  double acc[MT][NS][2];
#pragma unroll
  for (int m = 0; m < MT; m++)
#pragma unroll
    for (int q = 0; q < NS; q++) acc[m][q][0] = acc[m][q][1] = 0.0;
  // this CTA's share of the beta-side lags (v_split > 1 shortens the serial chain of a CTA when
  // the grid is too small to fill the GPU: single cosmology, k-sharded ranks)
  const int vlen = ((tb.NVp / VC + tb.vsplit - 1) / tb.vsplit) * VC;
  const int tv_begin = vs * vlen, NVp = min(tb.NVp, tv_begin + vlen);
  double2 tcur[MT], tnxt[MT];
#pragma unroll
  for (int m = 0; m < MT; m++) {
    tcur[m] = make_double2(0.0, 0.0);
    if (tv_begin < tb.NVp) tcur[m] = __ldg(Tp[m] + (size_t)(tv_begin >> 3) * cstride);  // (an empty split adds zeros)
    tnxt[m] = tcur[m];
  }

  mbar_wait(&mbar, 0);

  for (int tv0 = tv_begin; tv0 < NVp; tv0 += VC) {
    if (tv0 + VC < NVp) {
#pragma unroll
      for (int m = 0; m < MT; m++) tnxt[m] = __ldg(Tp[m] + (size_t)((tv0 + VC) >> 3) * cstride);
    }
    double bf[NS][2];
#pragma unroll
    for (int q = 0; q < NS; q++) bf[q][0] = s_c[q][tv0], bf[q][1] = s_c[q][tv0 + 1];
    // (half, slot, tile) order: an accumulator is touched again 4 NS DMMAs later
#pragma unroll
    for (int q = 0; q < NS; q++)
#pragma unroll
      for (int m = 0; m < MT; m++) dmma884(acc[m][q], tcur[m].x, bf[q][0]);  // lags tv0 + 0, 2, 4, 6
#pragma unroll
    for (int q = 0; q < NS; q++)
#pragma unroll
      for (int m = 0; m < MT; m++) dmma884(acc[m][q], tcur[m].y, bf[q][1]);  // lags tv0 + 1, 3, 5, 7
#pragma unroll
    for (int m = 0; m < MT; m++) tcur[m] = tnxt[m];
  }

  // alpha side, per slot: out[ab][r] = sum_u arev_ab[u - r + 7] * S[u][r], the sum running over the
  // lags of ONE row block.  Lane (g, t) holds S[tu_m + g][2 t] and S[tu_m + g][2 t + 1] of its tiles.
  // A warp holds tiles of one row block, or of two when it straddles a block boundary: side 0 = the
  // block of tile 0, side 1 = that of tile 3.
  // The FP64 instructions of this part queue behind the DMMAs of the warps that are still in their
  // main loops, so what counts is the length of its dependency chains: the 2 NSL partial sums of a
  // lane are built as independent chains (one DFMA per tile each), then reduced over the 8 lane
  // groups by recursive halving (3 exchange levels, 7/8 NSL' shuffle + add pairs in all).
  const int rbA = rb_m[0], rbB = rb_m[MT - 1];
  const int nside = (rbA == rbB) ? 1 : 2;
  if (lane == 0) s_wrb[warp][0] = rbA, s_wrb[warp][1] = (nside == 2) ? rbB : -1;
  const int abmask = item.abmask;
  constexpr int NPV = (2 * NSL + 7) / 8 * 8;  // values per lane, padded to a multiple of 8
  for (int side = 0; side < nside; side++) {
    const int rbs = side ? rbB : rbA;
    if (rbs < rb_lo || rbs >= rb_hi) continue;  // (never summed: the row block is another rank's)
    double pv[NPV];
#pragma unroll
    for (int i = 0; i < NPV; i++) pv[i] = 0.0;
#pragma unroll
    for (int m = 0; m < MT; m++) {
      if (rb_m[m] != rbs) continue;
      const double *ap = s_a + (R - 1) + g - 2 * t + tu_m[m];
      RT_ASSERT((R - 1) + g - 2 * t + tu_m[m] - 1 >= 0 && (R - 1) + g - 2 * t + tu_m[m] < LP);
#pragma unroll
      for (int q = 0; q < NS; q++)
#pragma unroll
        for (int ab = 0; ab < 3; ab++) {
          if (!((abmask >> ab) & 1)) continue;  // (warp uniform)
          const double2 w = make_double2(ap[(q * 3 + ab) * LP], ap[(q * 3 + ab) * LP - 1]);
          pv[(q * 3 + ab) * 2] = fma(w.x, acc[m][q][0], pv[(q * 3 + ab) * 2]);
          pv[(q * 3 + ab) * 2 + 1] = fma(w.y, acc[m][q][1], pv[(q * 3 + ab) * 2 + 1]);
        }
    }
    // after level l (lane bit 16 >> l) a lane keeps the half of its values selected by that bit:
    // lane group (b16, b8, b4) ends with the totals of values b16 NPV/2 + b8 NPV/4 + b4 NPV/8 + i
#pragma unroll
    for (int lvl = 0; lvl < 3; lvl++) {
      const int half = NPV >> (lvl + 1), bit = 16 >> lvl;
      const bool up = (lane & bit) != 0;
#pragma unroll
      for (int i = 0; i < half; i++) {
        const double send = up ? pv[i] : pv[i + half];
        const double keep = up ? pv[i + half] : pv[i];
        pv[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
      }
    }
    const int v0 = ((lane >> 4) & 1) * (NPV / 2) + ((lane >> 3) & 1) * (NPV / 4) + ((lane >> 2) & 1) * (NPV / 8);
#pragma unroll
    for (int i = 0; i < NPV / 8; i++) {
      const int v = v0 + i;  // = (q 3 + ab) 2 + (r & 1), r = 2 t + (r & 1)
      if (v < 2 * NSL) s_red[(warp * 2 + side) * (NSL * R) + (v >> 1) * R + 2 * t + (v & 1)] = pv[i];
    }
  }
  __syncthreads();
  // one thread per (row block of this CTA, slot, ab, row): add the warps' partial sums in warp
  // order and store this CTA's part of the row block
  const int rb_first = (c * IPC) / tb.NV, nrbl = (c * IPC + IPC - 1) / tb.NV - rb_first + 1;
  const int nch = tb.nchunk * tb.vsplit;
  for (int idx = tid; idx < nrbl * NSL * R; idx += NWARP * 32) {
    const int rbl = idx / (NSL * R), x = idx - rbl * (NSL * R), rbx = rb_first + rbl;
    if (rbx < rb_lo || rbx >= rb_hi) continue;
    const int slot = x / R, r = x - slot * R, q = slot / 3, ab = slot - 3 * q;
    const int e = s_e[q];
    // Jn0 only feeds the RSD terms (rt:804): cosmologies without them keep their old entries
    if (!s_ok[q] || !((abmask >> ab) & 1) || (n >= 7 && !cosmo[e].sw_pr)) continue;
    // the warps' parts in a fixed pairwise order (x + 0.0 is exact for the warps of other blocks)
    double sw[NWARP];
#pragma unroll
    for (int wv = 0; wv < NWARP; wv++) {
      const double s0 = (s_wrb[wv][0] == rbx) ? s_red[(wv * 2) * (NSL * R) + x] : 0.0;
      const double s1 = (s_wrb[wv][1] == rbx) ? s_red[(wv * 2 + 1) * (NSL * R) + x] : 0.0;
      sw[wv] = s0 + s1;
    }
#pragma unroll
    for (int step = 1; step < NWARP; step <<= 1)
#pragma unroll
      for (int wv = 0; wv + step < NWARP; wv += 2 * step) sw[wv] += sw[wv + step];
    const double s = sw[0];
    const int part = vs * tb.nchunk + (c - (rbx * tb.NV) / IPC);
    RT_ASSERT(part >= 0 && part < nch && c - (rbx * tb.NV) / IPC < tb.nchunk && rbx * R + r < tb.nk);
    double *dst = Jpart + (((long long)e * N_JKERN + n) * nch + part) * 9 * tb.nk + rbx * R + r;
    if (L.replicate) {
#pragma unroll
      for (int pair = 0; pair < 9; pair++) dst[(long long)pair * tb.nk] = s;
    } else {
      // (transposed items: the slot's spectrum is the alpha side of the pair)
      dst[(long long)(item.swap ? s_cd[q] * 3 + ab : ab * 3 + s_cd[q]) * tb.nk] = s;
    }
  }
}

template <int NWARP, int MT, int NS, int MINB>
__global__ void __launch_bounds__(NWARP * 32, MINB)
    k_bilinear(IntegralTabs tb, const Cosmo *__restrict__ cosmo, const double *__restrict__ Prev,
               double *__restrict__ Jpart, const __grid_constant__ BilLaunch L, int rb_lo, int rb_hi, int c_lo,
               const int *__restrict__ act, const int *__restrict__ nact) {
  bil_body<NWARP, MT, NS>(tb, cosmo, Prev, Jpart, L, rb_lo, rb_hi, c_lo, act, nact);
}
// the same with a register cap instead of an occupancy target (experiments: room for other kernels)
template <int NWARP, int MT, int NS, int MAXREG>
__global__ void __maxnreg__(MAXREG)
    k_bilinear_r(IntegralTabs tb, const Cosmo *__restrict__ cosmo, const double *__restrict__ Prev,
                 double *__restrict__ Jpart, const __grid_constant__ BilLaunch L, int rb_lo, int rb_hi, int c_lo,
                 const int *__restrict__ act, const int *__restrict__ nact) {
  bil_body<NWARP, MT, NS>(tb, cosmo, Prev, Jpart, L, rb_lo, rb_hi, c_lo, act, nact);
}

// ---------------------------------------------------------------------------- k_jlo
// J_{000}(P00,P00) at the padded row nloMR, needed by P_MR,4..6 (rt:1252,1267-1272)
__global__ void __launch_bounds__(256)
    k_jlo(IntegralTabs tb, double kfac_lo, const double *__restrict__ Prev,
          double *__restrict__ Jlo, const int *__restrict__ mask) {
  const int e = blockIdx.x;
  if (mask && !mask[e]) return;
  extern __shared__ double s_p[];  // [nsup]
  __shared__ double s_w[8];
  const double *a = Prev + (long long)e * 3 * tb.LP + (BIL_R - 1);
  for (int j = threadIdx.x; j < tb.nsup; j += blockDim.x) s_p[j] = a[j];
  __syncthreads();
  double tot = 0.0;
  for (int jj = threadIdx.x; jj < tb.nsup; jj += blockDim.x) {
    double acc = 0.0;
    const double *Tc = tb.Tlo + jj;
    for (int ll = 0; ll < tb.nsup; ll++) acc = fma(__ldg(Tc + (size_t)ll * tb.nsup), s_p[ll], acc);
    tot += acc * s_p[jj];
  }
  tot = warp_sum(tot);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); wv++) s += s_w[wv];
    Jlo[e] = kfac_lo * s;
  }
}

// ---------------------------------------------------------------------------- k_pz
// PZb[e][n][ab][i] = dlnk/(2 pi^2) k_i^3 P00(k_i) sum_m P_ab(q_m) G_n[i_pad - m]
// One warp per output row: the lanes stride over the samples m (coalesced reads of the reversed
// kernel row, conflict-free shared-memory reads of the spectrum) and the 32 partial sums are
// added by a butterfly -- a chain of (np - jlo)/32 FMAs instead of np - jlo per row, which is what
// a single cosmology (k-sharded ranks: a few dozen rows) waits for.  The order of the additions
// does not depend on the batch or on the k-sharding.
enum { PZ_ROWS = 16 };  // rows per CTA (4 warps x 4 rows)
__global__ void __launch_bounds__(128)
    k_pz(IntegralTabs tb, double pre, const double *__restrict__ P3, double *__restrict__ PZb,
         int row0, int nrows, const int *__restrict__ mask, unsigned int need) {
  const int e = blockIdx.y;
  if (mask && !mask[e]) return;
  if (!((need >> blockIdx.x) & 1u)) return;  // no requested output consumes PZ_n(P_ab)
  const int n = blockIdx.x / 3, ab = blockIdx.x - 3 * n;
  extern __shared__ double s_p[];  // [np]
  const double *Pab = P3 + ((long long)e * 3 + ab) * tb.np;
  for (int m = threadIdx.x; m < tb.np; m += blockDim.x) s_p[m] = Pab[m];
  __syncthreads();
  const double *G = tb.G + (long long)n * (2 * tb.np - 1) + (tb.np - 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r_end = min(nrows, (int)(blockIdx.z + 1) * PZ_ROWS);
  for (int ii = blockIdx.z * PZ_ROWS + warp; ii < r_end; ii += 4) {
    const int i = row0 + ii, ipad = tb.nshift + i;
    double acc = 0.0;
    RT_ASSERT(ipad - tb.jlo <= tb.np - 1 && ipad - (tb.np - 1) >= -(tb.np - 1) && i < tb.nk);
    for (int m = tb.jlo + lane; m < tb.np; m += 32) acc = fma(s_p[m], __ldg(G + ipad - m), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const double k = tb.kpad[ipad];
      PZb[(((long long)e * N_ZKERN + n) * 3 + ab) * tb.nk + i] =
          pre * (k * k * k) * P3[(long long)e * 3 * tb.np + ipad] * acc;
    }
  }
}

// ---------------------------------------------------------------------------- k_assemble
// (the pieces live in stage_device.h: k_stage_post runs the same code in front of the right-hand side)
template <int ASM_ROWS>
__global__ void __launch_bounds__(ASM_ROWS == 4 ? 768 : 256)
    k_assemble(IntegralTabs tb, const Cosmo *__restrict__ cosmo, const double *__restrict__ Jpart,
               const double *__restrict__ PZb, const double *__restrict__ P3,
               const double *__restrict__ Jlo, double *__restrict__ src, double *__restrict__ raw,
               int row0, int nrows, const int *__restrict__ mask, int groups) {
  const int e = blockIdx.y;
  if (mask && !mask[e]) return;
  __shared__ AsmShared<ASM_ROWS> sa;
  const int r0 = row0 + blockIdx.x * ASM_ROWS;
  const int rows = min(ASM_ROWS, row0 + nrows - r0);
  asm_load_table(tb, sa);
  asm_gather_vals(tb, cosmo[e].sw_pr, Jpart, PZb, P3, Jlo, raw, e, r0, rows, groups, sa);
  __syncthreads();
  for (int idx = threadIdx.x; idx < N_SRC * ASM_ROWS; idx += blockDim.x) {
    const int o = idx / ASM_ROWS, rr = idx - o * ASM_ROWS;
    if (rr >= rows || !asm_wanted(o, groups)) continue;
    src[((long long)e * N_SRC + o) * tb.nk + r0 + rr] = asm_source(tb, sa, o, rr, r0 + rr);
  }
}

// ---------------------------------------------------------------------------- launchers
// Tuning variants of the bilinear kernel (threads per CTA, slots per CTA, CTAs per SM), selected
// once per process by RTRG_BIL_VARIANT; 0 is the production configuration.  Results do not depend on
// the slot count; they depend on the threads per CTA at round-off level (which lags share a
// partial sum).
struct BilVariant {
  int nwarp, mt, ns;
  int ipc() const { return nwarp * mt * 8; }  // items (alpha-side lags) per CTA
};
static const BilVariant kBilVariants[] = {{8, 4, 3}, {8, 2, 3}, {16, 2, 3}, {4, 4, 3}, {8, 2, 4}, {4, 2, 3}, {16, 1, 3}, {8, 3, 3}, {8, 3, 3}, {8, 4, 3}};
static int bil_variant_index() {
  static const int v = [] {
    const char *e = std::getenv("RTRG_BIL_VARIANT");
    const int i = e ? std::atoi(e) : 0;
    return (i >= 0 && i < (int)(sizeof kBilVariants / sizeof kBilVariants[0])) ? i : 0;
  }();
  return v;
}
int bilinear_tpb() { return kBilVariants[bil_variant_index()].ipc(); }
size_t bilinear_smem_bytes(const IntegralTabs &tb) {
  const BilVariant v = kBilVariants[bil_variant_index()];
  return (size_t)(3 * v.ns * tb.LP + v.nwarp * 2 * 3 * v.ns * BIL_R) * sizeof(double);
}
template <class F>
static auto bil_dispatch(F &&f) {
  switch (bil_variant_index()) {
    case 1: return f(k_bilinear<8, 2, 3, 4>, kBilVariants[1]);
    case 2: return f(k_bilinear<16, 2, 3, 2>, kBilVariants[2]);
    case 3: return f(k_bilinear<4, 4, 3, 4>, kBilVariants[3]);
    case 4: return f(k_bilinear<8, 2, 4, 3>, kBilVariants[4]);
    case 5: return f(k_bilinear<4, 2, 3, 8>, kBilVariants[5]);
    case 6: return f(k_bilinear<16, 1, 3, 4>, kBilVariants[6]);
    case 7: return f(k_bilinear<8, 3, 3, 2>, kBilVariants[7]);
    case 8: return f(k_bilinear_r<8, 3, 3, 96>, kBilVariants[8]);
    case 9: return f(k_bilinear_r<8, 4, 3, 104>, kBilVariants[9]);
    default: return f(k_bilinear<8, 4, 3, 2>, kBilVariants[0]);
  }
}

// Evaluation for every (unmasked) cosmology: y -> the source rows of the requested output
// groups (GRP_* bits; GRP_ALL also fills raw J/PZ/Jn0 when raw != nullptr).  identical != 0:
// the three spectra in y are the same array (1-loop cache).  Returns the number of launches.
// side: optional second stream + fork/join events.  The P13-type convolutions (k_pz) and k_jlo only
// need the extrapolated spectra, not the bilinear sums, so they run beside k_bilinear (a parallel
// branch of the captured graph) and k_assemble joins the two: off the critical path of a single
// cosmology, where every kernel of an evaluation is latency bound.  Not used while per-kernel
// event timing is on.
int launch_integrals(const IntegralTabs &tb, const Batch &S, const double *y, long long ystride,
                     double *src, double *raw, const int *mask, int groups, int identical,
                     cudaStream_t st, Profiler *prof, const SideStream *side, bool assemble) {
  const int B = S.B, row0 = S.k_lo, nrows = S.k_hi - S.k_lo;
  int launches = 0;
  // work items: kernel n with the beta-side spectra the requested groups consume
  BilLaunch L;
  int nitems = 0, units = 0;
  for (int n = 0; n < N_JKERN; n++) {
    int need = 0;
    for (int gi = 0; gi < 4; gi++)
      if (groups & (1 << gi)) need |= tb.need_cd[gi][n];
    if (groups & GRP_RAW) need = 7;  // parity hooks: every product, used by an output or not
    if (!need) continue;
    if (identical) need = 1;
    BilItem it;
    it.n = it.tn = (short)n;
    it.swap = 0;
    // J_n(ab, cd) = a_ab^T T_n b_cd: the matrix-vector product may run on either side.  Kernels with a
    // transposed copy take the side with fewer spectra (the default output columns need (ab, cd) =
    // (2, 1), (2, 2) of n = 7, 8, 9: one product T_n^T a_2 instead of two)
    if (!identical && !(groups & GRP_RAW) && n >= TKERN_FIRST && n < TKERN_FIRST + N_TKERN) {
      int need_a = 0;
      for (int gi = 0; gi < 4; gi++)
        if (groups & (1 << gi)) need_a |= tb.need_ab[gi][n];
      if (__builtin_popcount(need_a) < __builtin_popcount(need)) {
        need = need_a;
        it.tn = (short)(N_JKERN + n - TKERN_FIRST);
        it.swap = 1;
      }
    }
    // the other side: only the dot products some requested output reads
    int other = 0;
    for (int gi = 0; gi < 4; gi++)
      if (groups & (1 << gi)) other |= it.swap ? tb.need_cd[gi][n] : tb.need_ab[gi][n];
    it.abmask = (short)((groups & GRP_RAW) ? 7 : identical ? 1 : other);
    it.ncd = 0;
    for (int c = 0; c < 3; c++)
      if (need & (1 << c)) it.cd[it.ncd++] = (short)c;
    for (int c = it.ncd; c < 3; c++) it.cd[c] = 0;
    L.it[nitems++] = it;
    units += it.ncd;
  }
  // heaviest items first (CTAs are dispatched in block order; a CTA's duration grows with the number of
  // spectra of its item): the light ones fill the tail of a small launch
  std::stable_sort(L.it, L.it + nitems, [](const BilItem &a, const BilItem &b) { return a.ncd > b.ncd; });
  L.replicate = identical ? 1 : 0;
  {
    dim3 g((tb.np + 127) / 128, B);
    RT_TIC(prof, PC_EXTRAP, st);
    k_extrap<<<g, 128, 0, st>>>(tb, S.cosmo, y, ystride, S.P3, S.Prev, mask, S.matvecs, units, S.act, S.nact);
    RT_TOC(prof, st);
    launches++;
  }
  const bool fork = side && side->stream && !prof;
  cudaStream_t st_side = fork ? side->stream : st;
  if (fork) {
    cudaEventRecord(side->fork, st);
    cudaStreamWaitEvent(st_side, side->fork, 0);
  }
  RT_TIC(prof, PC_BILINEAR, st);
  if (nitems) {
    // z covers the slot triples of the widest item of this launch; CTAs beyond an item's own
    // slot count exit at once
    int maxncd = 1;
    for (int i = 0; i < nitems; i++) maxncd = std::max(maxncd, (int)L.it[i].ncd);
    // CTAs along the global (row block, lag) axis that hold lags of this rank's row blocks
    const int rb_lo = row0 / BIL_R, rb_hi = (row0 + nrows) / BIL_R;
    bil_dispatch([&](auto kern, const BilVariant &v) {
      const int ipc = v.ipc(), ns = v.ns, tpb = v.nwarp * 32;
      const int c_lo = (rb_lo * tb.NV) / ipc, c_hi = (rb_hi * tb.NV - 1) / ipc;
      dim3 g((c_hi - c_lo + 1) * tb.vsplit, nitems, (B * maxncd + ns - 1) / ns);
      kern<<<g, tpb, bilinear_smem_bytes(tb), st>>>(tb, S.cosmo, S.Prev, S.Jpart, L, rb_lo, rb_hi, c_lo, S.act, S.nact);
      return 0;
    });
    launches++;
  }
  RT_TOC(prof, st);
  if (groups & GRP_PMR) {
    RT_TIC(prof, PC_JLO, st);
    k_jlo<<<B, 256, tb.nsup * sizeof(double), st_side>>>(tb, tb.kfac_lo, S.Prev, S.Jlo, mask);
    RT_TOC(prof, st);
    launches++;
  }
  // P13-type log-convolutions: only those the requested groups consume (P_T,jm needs none)
  unsigned int need_pz = 0;
  for (int gi = 0; gi < 4; gi++)
    if (groups & (1 << gi)) need_pz |= tb.need_pz[gi];
  if (groups & GRP_RAW) need_pz = (1u << (N_ZKERN * 3)) - 1u;
  if (need_pz) {
    dim3 g(N_ZKERN * 3, B, (nrows + PZ_ROWS - 1) / PZ_ROWS);
    const double pre = tb.dlnk / (2.0 * M_PI * M_PI);  // rt:719
    RT_TIC(prof, PC_PZ, st);
    k_pz<<<g, 128, tb.np * sizeof(double), st_side>>>(tb, pre, S.P3, S.PZb, row0, nrows, mask, need_pz);
    RT_TOC(prof, st);
    launches++;
  }
  if (fork) {
    cudaEventRecord(side->join, st_side);
    cudaStreamWaitEvent(st, side->join, 0);
  }
  if (assemble) {
    RT_TIC(prof, PC_ASSEMBLE, st);
    if ((long long)B * nrows <= 2048)
      k_assemble<4><<<dim3((nrows + 3) / 4, B), 768, 0, st>>>(tb, S.cosmo, S.Jpart, S.PZb, S.P3, S.Jlo, src, raw, row0,
                                                             nrows, mask, groups);
    else
      k_assemble<16><<<dim3((nrows + 15) / 16, B), 256, 0, st>>>(tb, S.cosmo, S.Jpart, S.PZb, S.P3, S.Jlo, src, raw,
                                                               row0, nrows, mask, groups);
    RT_TOC(prof, st);
    launches++;
  }
  return launches;
}

void launch_extrap_only(const IntegralTabs &tb, const Batch &S, const double *y, long long ystride,
                        const int *mask, cudaStream_t st) {
  dim3 g((tb.np + 127) / 128, S.B);
  k_extrap<<<g, 128, 0, st>>>(tb, S.cosmo, y, ystride, S.P3, S.Prev, mask, nullptr, 0, nullptr, nullptr);
}

int integrals_configure(const IntegralTabs &tb) {
  // opt in to the dynamic shared memory the bilinear kernel needs on large grids
  return bil_dispatch([&](auto kern, const BilVariant &) {
    return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bilinear_smem_bytes(tb));
  });
}

}  // namespace rtrg
