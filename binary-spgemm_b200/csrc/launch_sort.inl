// binary-spgemm_b200/csrc/launch_sort.inl — launcher of the sorting-network kernels (fused_sort.cuh) for ONE ELL width:
// included by tu_sort_w{4,8,16,32}.cu with SORT_W defined, so that the four widths compile in parallel.
#include "ctx.h"
#include "fused_sort.cuh"

// Compute warps per CTA for the persistent fused kernels (one more warp, the chain helper, is added at launch).
// Shared memory comes out of the SM's 256 KB unified array in steps (.., 164, 196, 228 KB); what is left is L1, which
// the gathers of B want: stay one step below the maximum unless that costs more than a fifth of the warps.
static int pick_compute_warps(size_t per_warp, size_t fixed, int max_warps, size_t optin) {
  auto fit = [&](size_t cap) { return cap > fixed ? (int)std::min<size_t>((size_t)max_warps, (cap - fixed) / per_warp) : 0; };
  const int w_max = fit(optin), w_step = fit(196 * 1024 - 1024);
  int w = (w_step * 5 >= w_max * 4) ? w_step : w_max;
  if (const char* e = getenv("BSPGEMM_WARPS")) w = std::max(1, std::min(w_max, atoi(e)));   // tuning knob
  return w;
}

template <int W, int LAL, bool ASYNC, bool FLT> static int launch_sort_t(bspgemm_dev* d, int* ccol) {
  const MulArgs& a = d->a;
  constexpr SortGeom G = sort_geom<W, LAL>();
  const u32 ntiles = (u32)(((size_t)a.m.An + G.R - 1) / G.R);
  // warps per CTA: what the kernel's register count allows (registers are allocated per SM sub-partition: 80 -> 6 warps
  // each, 81..102 -> 5), one of them the chain helper
  auto kern = ASYNC ? k_fused_sort_async<W, LAL, FLT> : k_fused_sort<W, LAL>;
  static cudaFuncAttributes fa; static bool have_fa = false;     // per instantiation: the query costs microseconds on every launch otherwise
  if (!have_fa) { CK(cudaFuncGetAttributes(&fa, kern)); have_fa = true; }
  const int max_compute = std::max(1, std::min(SORT_MAX_WARPS, fa.maxThreadsPerBlock / 32) - 1);
  // staging buffers per warp: 2 = commit one tile later; small tiles get 3 (commit lag 2) as long as that costs no warps.
  // ASYNC: the input buffer the cp.async copies land in + 2 staging buffers (the commit of tile t-2 comes before tile t is
  // staged, see fused_sort.cuh); nothing is kept back for L1, which the copies bypass.
  const size_t one_buf = (size_t)(ASYNC ? sort_stage_words_a(G.R, G.LA, W) : sort_stage_words(G.R, G.LA, W)) * 4, in_buf = (size_t)sort_input_words(G.R, G.LA, W) * 4, fixed = ELL_CTA_WORDS * 4 + 64;
  auto warp_bytes = [&](int nb) { return ASYNC ? in_buf + (size_t)(nb - 1) * one_buf : (size_t)nb * one_buf + 256; };   // sync: + the next tile's <= 64 A nonzeros
  int nbuf = ASYNC ? 3 : 2;
  if (!ASYNC) {
    const int w2 = pick_compute_warps(warp_bytes(2), fixed, max_compute, d->smem_optin);
    while (nbuf < 3 && pick_compute_warps(warp_bytes(nbuf + 1), fixed, max_compute, d->smem_optin) >= w2) ++nbuf;
  }
  if (const char* e = getenv("BSPGEMM_NBUF")) nbuf = std::max(2, std::min(4, atoi(e)));   // tuning knob
  const size_t per_warp = warp_bytes(nbuf);
  int warps = pick_compute_warps(per_warp, fixed, max_compute, d->smem_optin);
  if (ASYNC && !getenv("BSPGEMM_WARPS")) warps = (int)std::min<size_t>((size_t)max_compute, (d->smem_optin - fixed) / per_warp);
  if (warps < 1) return fail(BSPGEMM_ERR_CUDA, "sort kernel does not fit on an SM");
  const size_t smem = per_warp * warps + ELL_CTA_WORDS * 4;
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
  p.Bm = (u32)a.m.Bm; p.Crow = a.dCrow; p.is64 = a.is64; p.Ccol = ccol; p.sc = d->d_sc; p.ntiles = ntiles; p.nbuf = (u32)nbuf;
#ifdef BSPGEMM_DEBUG_KNOBS
  p.debug_nochain = getenv("BSPGEMM_DEBUG_NOCHAIN") ? (u32)(G.R * G.LA * W) : 0u;   // WRONG RESULTS: timing experiments only
#endif
  d->st.rows_per_tile = G.R; d->st.variant = 2; d->st.kernel_flags = (ASYNC ? 1 : 0) | (FLT ? 2 : 0);
  p.one = 1u; p.mone = 0xffffffffu;
  if (getenv("BSPGEMM_VERBOSE")) {
    fprintf(stderr, "k_fused_sort%s<%d,%d>%s: regs %d, maxThreadsPerBlock %d, static smem %zu, launch %d x %d threads, dyn smem %zu, nbuf %d\n",
            ASYNC ? "_async" : "", W, LAL, FLT ? " (floating-point network)" : "", fa.numRegs, fa.maxThreadsPerBlock, fa.sharedSizeBytes, grid, (warps + 1) * 32, smem, nbuf);
  }
  kern<<<grid, (warps + 1) * 32, smem, d->stream>>>(p);        // + the chain helper warp
  d->launches++;
  CK(cudaGetLastError());
  return BSPGEMM_OK;
}

#define SORT_CAT2(a, b) a##b
#define SORT_CAT(a, b) SORT_CAT2(a, b)

// the floating-point network exists for the big-tile geometries only (sort_big_tile)
template <int Lv> static int set_attrs_l(int smem_optin) {
  BSP_ATTR((k_fused_sort<SORT_W, Lv>));
  BSP_ATTR((k_fused_sort_async<SORT_W, Lv, false>));
  if constexpr (sort_big_tile(SORT_W, Lv)) BSP_ATTR((k_fused_sort_async<SORT_W, Lv, true>));
  return BSPGEMM_OK;
}
int SORT_CAT(set_attrs_sort_w, SORT_W)(int smem_optin) {
  CKS(set_attrs_l<2>(smem_optin)); CKS(set_attrs_l<3>(smem_optin)); CKS(set_attrs_l<4>(smem_optin)); CKS(set_attrs_l<5>(smem_optin));
  return BSPGEMM_OK;
}

template <int Lv> static int launch_sort_l(bspgemm_dev* d, int* ccol) {
  const bool async = sort_plan_async(SORT_W, Lv);
  const bool flt = d->ell_pad == EMPTY_F;                      // the ELL copy was built for the floating-point network (launch_ell / prepare_b)
  if (flt != sort_plan_flt(SORT_W, Lv, d->a.m.Bm)) return fail(BSPGEMM_ERR_CUDA, "internal: ELL padding does not match the sort kernel of this plan");
  if constexpr (sort_big_tile(SORT_W, Lv)) { if (flt) return launch_sort_t<SORT_W, Lv, true, true>(d, ccol); }
  return async ? launch_sort_t<SORT_W, Lv, true, false>(d, ccol) : launch_sort_t<SORT_W, Lv, false, false>(d, ccol);
}
int SORT_CAT(launch_sort_w, SORT_W)(bspgemm_dev* d, int* ccol) {
  // Big tiles (32 keys per lane, one pass per tile: config 3) take the cp.async kernel, the others the register-prefetch
  // one (config 2: 0.205 ms against 0.24 ms); see sort_plan_async / sort_plan_flt (fused_sort.cuh).
  switch (d->sort_LAL) { case 2: return launch_sort_l<2>(d, ccol); case 3: return launch_sort_l<3>(d, ccol); case 4: return launch_sort_l<4>(d, ccol); default: return launch_sort_l<5>(d, ccol); }
}
