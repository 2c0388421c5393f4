// binary-spgemm_b200/csrc/tu_ell.cu — translation unit of the ELL fast path: the CSR -> ELL re-layout of B (k_build_ell) and the
// ordered-table kernel k_fused_ell (fused_ell.cuh); the sorting-network kernels live in tu_sort_w*.cu, one per ELL width.
#include "ctx.h"
#include "fused_ell.cuh"
#include "fused_sort.cuh"      // host-side plan helpers only (sort_plan_flt): the sort kernels are instantiated in tu_sort_w*.cu

int set_attrs_ell(int smem_optin) {
#define ATTR_E(Wv) BSP_ATTR((k_fused_ell<Wv, 1>)); BSP_ATTR((k_fused_ell<Wv, 2>)); BSP_ATTR((k_fused_ell<Wv, 4>)); BSP_ATTR((k_fused_ell<Wv, 8>))
  ATTR_E(4); ATTR_E(8); ATTR_E(16); ATTR_E(32);
#undef ATTR_E
  return BSPGEMM_OK;
}

int launch_sort(bspgemm_dev* d, int* ccol) {
  switch (d->ell_W) { case 4: return launch_sort_w4(d, ccol); case 8: return launch_sort_w8(d, ccol); case 16: return launch_sort_w16(d, ccol); default: return launch_sort_w32(d, ccol); }
}

u32 sort_pad_for(int W, int LAL, int Bm) { return sort_plan_flt(W, LAL, Bm) ? EMPTY_F : EMPTY; }

int build_ell(bspgemm_dev* d, int W, bool sorted, u32 pad) {
  const MulArgs& a = d->a;
  d->pb.ell_W = 0;                                     // whatever copy `bell` held is gone (set again below / by bspgemm_dev_prepare_b)
  d->ell_pad = pad;
  CKS(d->bell.ensure(((size_t)a.m.Bn + 1) * W + 4));
  const long long threads = (((long long)a.m.Bn + ELL_RPT) / ELL_RPT) * (W / 4);     // ELL_RPT rows per thread
  const int grid = (int)((threads + 255) / 256);
#define BE(Wv) do { if (sorted) k_build_ell<Wv, true><<<grid, 256, 0, d->stream>>>(a.m.Brow, a.m.Bcol, a.m.Bn, (u32)a.m.Bm, d->bell.p, d->d_sc, pad); \
                    else k_build_ell<Wv, false><<<grid, 256, 0, d->stream>>>(a.m.Brow, a.m.Bcol, a.m.Bn, (u32)a.m.Bm, d->bell.p, d->d_sc, pad); } while (0)
  switch (W) { case 4: BE(4); break; case 8: BE(8); break; case 16: BE(16); break; default: BE(32); break; }
#undef BE
  d->launches++;
  CK(cudaGetLastError());
  return BSPGEMM_OK;
}

int launch_ell(bspgemm_dev* d) {
  const MulArgs& a = d->a;
  int* ccol = d->user_ccol ? d->user_ccol : d->ccol.p;
  const int W = d->ell_W, R = d->ell_R;
  // the ELL copy of a prepared B (bspgemm_dev_prepare_b) is reused when its width and padding are what this plan needs
  // (sorted rows serve both kernels); otherwise it is (re)built — and stays the prepared copy if B is the prepared matrix
  const u32 pad = d->use_sort ? sort_pad_for(W, d->sort_LAL, a.m.Bm) : EMPTY;
  const bool prep = b_prepared(d);
  if (!(prep && d->pb.ell_W == W && d->ell_pad == pad)) {
    CKS(build_ell(d, W, d->use_sort || prep, pad));
    if (prep) d->pb.ell_W = W;
  }
  if (d->use_sort) return launch_sort(d, ccol);
  const u32 ntiles = (u32)(((size_t)a.m.An + R - 1) / R);
  const int warps = d->ell_warps;
  const u32 SW = (u32)R * d->ell_maxA * (u32)W;
  const size_t smem = ((size_t)ell_warp_words(R, d->ell_TW, SW) * warps + ELL_CTA_WORDS) * 4;
  const long long want = ((long long)ntiles + warps - 1) / warps;
  const int grid = (int)std::max<long long>(1, std::min<long long>(want, d->sm_count));
  const size_t niter = ((size_t)ntiles + (size_t)grid * warps - 1) / ((size_t)grid * warps);
  const size_t nblocks = niter * grid + 1;
  u64* chain = nullptr;
  CKS(chain_reserve(d, nblocks, &chain));
  CK(cudaEventRecord(d->ev[3], d->stream));
  EllArgs p{};
  p.blk_status = chain;
  p.Arow = a.m.Arow; p.Acol = a.m.Acol; p.Bell = d->bell.p; p.An = a.m.An; p.Bn = a.m.Bn;
  { const double inv = 4294967296.0 / (double)a.m.Bm * (1.0 - 1.0 / 1048576.0); float f = (float)inv; if ((double)f > inv) f = nextafterf(f, 0.0f); p.inv_bm = f; }
  p.SW = SW; p.lf16 = d->ell_lf16;
  p.TW = d->ell_TW; p.Bm = (u32)a.m.Bm; p.Crow = a.dCrow; p.is64 = a.is64; p.Ccol = ccol; p.sc = d->d_sc;
  p.ntiles = ntiles;
#ifdef BSPGEMM_DEBUG_KNOBS
  p.debug_nochain = getenv("BSPGEMM_DEBUG_NOCHAIN") ? (u32)(R * d->ell_maxA * W) : 0u;   // WRONG RESULTS: timing experiments only
#endif
#define LE(Wv, Rv) k_fused_ell<Wv, Rv><<<grid, (warps + 1) * 32, smem, d->stream>>>(p)   /* + the chain helper warp */
#define LER(Wv) do { switch (R) { case 1: LE(Wv, 1); break; case 2: LE(Wv, 2); break; case 4: LE(Wv, 4); break; default: LE(Wv, 8); break; } } while (0)
  switch (W) { case 4: LER(4); break; case 8: LER(8); break; case 16: LER(16); break; default: LER(32); break; }
#undef LER
#undef LE
  d->launches++;
  CK(cudaGetLastError());
  return BSPGEMM_OK;
}

