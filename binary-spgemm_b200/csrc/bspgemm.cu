// binary-spgemm_b200/csrc/bspgemm.cu — host side of libbspgemm.so: the C ABI declared in include/bspgemm.h.
//
// Layering (replaces final/SpGEMM_mpi_omp.c L1-L3):
//   DevCtx      one GPU: workspace, output arena, the launch sequence (estimate -> bins -> scan -> fill)
//   Global ctx  the "communicator": G DevCtx + NCCL comms (dlopen'ed) for the B broadcast
//   C ABI       host-pointer operators (upload / shard / gather) and the device-resident operator
// No CPU fallback anywhere: every failure is returned as a status code.
#include "ctx.h"
#include "fused_ell.cuh"       // ELL_CTA_WORDS, ell_table_limit (planning only: the ELL kernels are instantiated in tu_ell.cu / tu_sort_w*.cu)
#include <chrono>
#include "rows_window.cuh"
#include "rows_sort.cuh"
#include "rows_bm.cuh"
#include "band.cuh"
#include "coo2csc.cuh"
#include "mask.cuh"

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
  return code;
}
extern "C" const char* bspgemm_strerror(int s) {
  switch (s) {
    case BSPGEMM_OK: return "ok";
    case BSPGEMM_ERR_CUDA: return "CUDA error";
    case BSPGEMM_ERR_NCCL: return "NCCL error";
    case BSPGEMM_ERR_OOM: return "out of memory";
    case BSPGEMM_ERR_OVERFLOW32: return "nnz(C) >= 2^31 on the 32-bit ABI (use the _i64 entry point)";
    case BSPGEMM_ERR_BADARG: return "bad argument";
    case BSPGEMM_ERR_NOGPU: return "no CUDA device (no CPU fallback exists)";
    case BSPGEMM_ERR_CAPACITY: return "output buffer too small";
    case BSPGEMM_ERR_STATE: return "bspgemm_init not called";
    default: return "unknown status";
  }
}
extern "C" const char* bspgemm_last_error(void) { return g_err; }
extern "C" const char* bspgemm_version(void) { return "bspgemm-b200 0.1 (sm_100a)"; }


static int wait_stream(cudaStream_t s);

static int g_cap_s_max() { const char* e = getenv("BSPGEMM_CAP_S"); int v = e ? atoi(e) : 512; if (v < 32) v = 32; if (v > 1024) v = 1024; int p = 32; while (p < v) p <<= 1; return p; }
static const u32 CAP_M1 = 2048;
// rows above CAP_M2 intermediate products take the windowed-bitmap kernel, rows up to it the CTA-wide sort (BSPGEMM_CAP_M2: tuning knob)
static u32 cap_m2_value() { const char* e = getenv("BSPGEMM_CAP_M2"); const int v = e ? atoi(e) : 16384; return v == 4096 ? 4096u : v == 8192 ? 8192u : 16384u; }
#define CAP_M2 (cap_m2_value())

// Every kernel that uses dynamic shared memory gets the opt-in maximum once, at context creation (BSP_ATTR, ctx.h).
static int set_kernel_attributes(int smem_optin) {
#define ATTR(k) BSP_ATTR(k)
#define ATTR_G(Gv) ATTR((k_rows_warp<Gv, MODE_COUNT>)); ATTR((k_rows_warp<Gv, MODE_FILL>)); ATTR((k_fused<Gv, true>)); ATTR((k_fused<Gv, false>))
  ATTR_G(4); ATTR_G(8); ATTR_G(16); ATTR_G(32);
  ATTR((k_rows_sort<32, 512, MODE_COUNT>)); ATTR((k_rows_sort<32, 512, MODE_FILL>)); ATTR((k_rows_sort<32, 512, MODE_STAGE>));
  ATTR((k_rows_sort<8, 256, MODE_COUNT>)); ATTR((k_rows_sort<8, 256, MODE_FILL>)); ATTR((k_rows_sort<8, 256, MODE_STAGE>));
  ATTR(k_rows_window<MODE_COUNT>); ATTR(k_rows_window<MODE_FILL>); ATTR(k_rows_window<MODE_STAGE>);
  ATTR((k_rows_bm<MODE_COUNT, false>)); ATTR((k_rows_bm<MODE_FILL, false>)); ATTR((k_rows_bm<MODE_STAGE, false>));
  ATTR((k_rows_bm<MODE_COUNT, true>)); ATTR((k_rows_bm<MODE_FILL, true>)); ATTR((k_rows_bm<MODE_STAGE, true>));
#undef ATTR_G
#undef ATTR
  CKS(set_attrs_ell(smem_optin));
  CKS(set_attrs_sort_w4(smem_optin)); CKS(set_attrs_sort_w8(smem_optin)); CKS(set_attrs_sort_w16(smem_optin)); CKS(set_attrs_sort_w32(smem_optin));
  return BSPGEMM_OK;
}

static int gidx(int G) { return G == 4 ? 0 : G == 8 ? 1 : G == 16 ? 2 : 3; }

template <int MODE> static int launch_rows_warp(bspgemm_dev* d) {
  const MulArgs& a = d->a;
  int* ccol = d->user_ccol ? d->user_ccol : d->ccol.p;
  // rows of up to 64 products: their own register-only kernel (BSPGEMM_NO_TINY: everything through k_rows_warp)
  const u32 tiny_max = getenv("BSPGEMM_NO_TINY") ? 0u : std::min<u32>(TINY_MAX, d->cap_s);      // (never beyond the warp bin: larger rows are on the CTA lists)
  if (tiny_max) {
    const long long want = ((long long)a.m.An + 7) / 8;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)d->sm_count * 8 * 2));
    k_rows_tiny<MODE><<<grid, 256, 0, d->stream>>>(a.m, d->ip.p, d->cnt.p, a.dCrow, a.is64, ccol, d->d_sc, tiny_max);
    d->launches++;
    CK(cudaGetLastError());
  }
  const size_t smem = (size_t)WARPS_S * (tab_words(d->cap_s) + d->cap_s) * sizeof(u32);
  int bps = 0;
#define LW(Gv) do { \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_rows_warp<Gv, MODE>, WARPS_S * 32, smem)); \
    const long long want = ((long long)a.m.An + WARPS_S - 1) / WARPS_S; \
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)d->sm_count * std::max(bps, 1) * 2)); \
    k_rows_warp<Gv, MODE><<<grid, WARPS_S * 32, smem, d->stream>>>(a.m, d->ip.p, d->cnt.p, d->cap_s, a.dCrow, a.is64, ccol, d->d_sc, tiny_max); } while (0)
  switch (d->G) { case 4: LW(4); break; case 8: LW(8); break; case 16: LW(16); break; default: LW(32); break; }
#undef LW
  d->launches++;
  CK(cudaGetLastError());
  return BSPGEMM_OK;
}

static int launch_fused(bspgemm_dev* d, u32 ntiles, int acc_ip) {
  const MulArgs& a = d->a;
  int* ccol = d->user_ccol ? d->user_ccol : d->ccol.p;
  // one persistent CTA per SM; every warp is an independent worker with its own shared-memory region
  const size_t per_warp = (size_t)fused_warp_words(d->cap_s) * sizeof(u32);
  const size_t avail = d->smem_optin - 64;
  int warps = (int)std::min<size_t>(32, avail / per_warp);
  if (warps < 1) return fail(BSPGEMM_ERR_CUDA, "fused kernel does not fit on an SM (%zu bytes per warp)", per_warp);
  const size_t smem = per_warp * warps;
  const long long want = ((long long)ntiles + warps - 1) / warps;
  const int grid = (int)std::max<long long>(1, std::min<long long>(want, d->sm_count));
  const bool notail = d->max_len_b <= (u32)d->G;
#define LF(Gv) do { \
    if (notail) k_fused<Gv, true><<<grid, warps * 32, smem, d->stream>>>(a.m, d->cnt.p, d->cap_s, a.dCrow, a.is64, ccol, d->status.p + SC_WORDS, d->d_sc, ntiles, acc_ip); \
    else        k_fused<Gv, false><<<grid, warps * 32, smem, d->stream>>>(a.m, d->cnt.p, d->cap_s, a.dCrow, a.is64, ccol, d->status.p + SC_WORDS, d->d_sc, ntiles, acc_ip); } while (0)
  switch (d->G) { case 4: LF(4); break; case 8: LF(8); break; case 16: LF(16); break; default: LF(32); break; }
#undef LF
  d->launches++;
  CK(cudaGetLastError());
  return BSPGEMM_OK;
}

template <int MODE> static int launch_bins_ml(bspgemm_dev* d) {
  const MulArgs& a = d->a;
  int* ccol = d->user_ccol ? d->user_ccol : d->ccol.p;
  const size_t An = (size_t)a.m.An;
  u32* l1 = d->lists.p, *l2 = d->lists.p + An, *l3 = d->lists.p + 2 * An;
  u32* ctr = d->d_sc->win_ctr + (MODE == MODE_FILL ? 3 : 0);
  int bps = 0;
  const u64* tofs = d->tofs.p;
  if (MODE == MODE_STAGE) {
    if (d->have_l && !d->use_window && !d->use_bm) return fail(BSPGEMM_ERR_CUDA, "internal: the global-bitmap kernel has no staged mode");
    ccol = d->temp.p;               // the kernels write row i at temp[tofs[i] ..)
  }
  // largest rows first; rows are handed out dynamically inside every kernel
  // compressed single pass for rows of up to 16384 products: only where a row can need more than one window (BSPGEMM_BM_NO_COMP: off)
  const u32 bm_cmode = ((u64)a.m.Bm > (u64)BM_WORDS * 32ull && (u32)a.m.Bm <= BM_COMP_MAX_BM && !getenv("BSPGEMM_BM_NO_COMP")) ? 1u : 0u;
  const bool bm_m2 = d->use_bm && getenv("BSPGEMM_BM_L_ONLY") == nullptr;        // (tuning knob: the M2 list stays on the CTA-wide sort)
  if (d->use_bm) {           // long rows (and rows of more than 1024 A entries), then the medium rows: windowed shared-memory bitmap (rows_bm.cuh)
    k_rows_bm<MODE, false><<<d->sm_count, BM_THREADS, BM_SMEM, d->stream>>>(a.m, l3, &d->d_sc->n_l, ctr + 0, d->cnt.p, a.dCrow, a.is64, ccol, tofs, d->d_sc, 0u);
    d->launches++;
    CK(cudaGetLastError());
    if (d->have_m2 && bm_m2) {
      k_rows_bm<MODE, true><<<d->sm_count, BM_THREADS, BM_SMEM, d->stream>>>(a.m, l2, &d->d_sc->n_m2, ctr + 1, d->cnt.p, a.dCrow, a.is64, ccol, tofs, d->d_sc, bm_cmode);
      d->launches++;
      CK(cudaGetLastError());
    }
  }
  if (d->have_l && !d->use_bm) {
    if (d->use_window)   // windowed shared-memory bitmap (rows_window.cuh)
      k_rows_window<MODE><<<d->sm_count, 1024, (size_t)WIN_WORDS * 4, d->stream>>>(a.m, l3, &d->d_sc->n_l, ctr + 0, d->cnt.p, d->G_big, WIN_WORDS, a.dCrow, a.is64, ccol, tofs, d->d_sc);
    else                 // matrices with more columns than WIN_MAX_WINDOWS windows: bitmap over [0,Bm) in global memory
      k_rows_gbitmap<MODE><<<d->l_grid, 1024, 0, d->stream>>>(a.m, l3, &d->d_sc->n_l, d->cnt.p, d->G_big, d->bitmaps.p, d->bm_words, a.dCrow, a.is64, ccol, d->d_sc);
    d->launches++;
    CK(cudaGetLastError());
  }
  if (d->have_m2 && !bm_m2) {      // 2048 < IP <= 16384: 512 threads, up to 32 keys per thread (rows_sort.cuh)
    const size_t smem = (size_t)CAP_M2 * 4;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_rows_sort<32, 512, MODE>, 512, smem));
    k_rows_sort<32, 512, MODE><<<d->sm_count * std::max(bps, 1), 512, smem, d->stream>>>(a.m, l2, &d->d_sc->n_m2, ctr + 1, d->ip.p, d->cnt.p, d->G_big, a.dCrow, a.is64, ccol, tofs, d->d_sc);
    d->launches++;
    CK(cudaGetLastError());
  }
  if (d->have_m) {       // cap_s < IP <= 2048: 256 threads, up to 8 keys per thread
    const size_t smem = (size_t)CAP_M1 * 4;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_rows_sort<8, 256, MODE>, 256, smem));
    k_rows_sort<8, 256, MODE><<<d->sm_count * std::max(bps, 1), 256, smem, d->stream>>>(a.m, l1, &d->d_sc->n_m1, ctr + 2, d->ip.p, d->cnt.p, d->G_big, a.dCrow, a.is64, ccol, tofs, d->d_sc);
    d->launches++;
    CK(cudaGetLastError());
  }
  return BSPGEMM_OK;
}

// MODE_STAGE epilogue: staged big rows -> their final position in Ccol
static int launch_copy_rows(bspgemm_dev* d) {
  const MulArgs& a = d->a;
  int* ccol = d->user_ccol ? d->user_ccol : d->ccol.p;
  const size_t An = (size_t)a.m.An;
  u32* ls[3] = { d->lists.p + 2 * An, d->lists.p + An, d->lists.p };
  const u32* nl[3] = { &d->d_sc->n_l, &d->d_sc->n_m2, &d->d_sc->n_m1 };
  const bool have[3] = { d->have_l, d->have_m2, d->have_m };
  for (int b = 0; b < 3; ++b) {
    if (!have[b]) continue;
    if (b == 2) k_copy_rows<true><<<d->sm_count * 8, 256, 0, d->stream>>>(ls[b], nl[b], d->cnt.p, d->tofs.p, d->temp.p, a.dCrow, a.is64, ccol);   // M1: warp per row
    else        k_copy_rows<false><<<d->sm_count * 8, 256, 0, d->stream>>>(ls[b], nl[b], d->cnt.p, d->tofs.p, d->temp.p, a.dCrow, a.is64, ccol);
    d->launches++;
    CK(cudaGetLastError());
  }
  return BSPGEMM_OK;
}

// ---- ELL fast path (fused_ell.cuh): usable when every B row has <= 32 columns and padding B to W columns
// per row at most doubles it; every output row then has IP <= max_len(A)*W and fits one warp's table.
static bool ell_plan(bspgemm_dev* d) {
  const MulArgs& a = d->a;
  const DevScalars& h = *d->h_sc;
  d->use_ell = false;
  // explicit bin / estimate overrides select the CSR-gather kernels
  if (d->mode == BSPGEMM_MODE_TWOPHASE || getenv("BSPGEMM_NO_ELL") || getenv("BSPGEMM_CAP_S") || getenv("BSPGEMM_FORCE_ESTIMATE")) return false;
  if (h.max_len_a == 0 || h.max_len_b == 0 || h.max_len_b > 32) return false;
  const bool clustered = h.span_rows > 0 && 2u * h.span_narrow > h.span_rows;   // most rows fit a 2^15-column window (k_probe_span)
  int W = 4; while (W < (int)h.max_len_b) W <<= 1;
  if ((u64)a.m.Bn * (u64)W > 4ull * (u64)a.Bnnz + 4096ull && !getenv("BSPGEMM_FORCE_ELL")) return false;   // padding waste
  u32 lf16 = 28; if (const char* e = getenv("BSPGEMM_ELL_LF16")) lf16 = (u32)std::max(16, std::min(64, atoi(e)));   // tuning knob
  const u32 TW = ell_table_limit(h.max_len_a, (u32)W, lf16);
  if (TW > 8192u || (u64)a.m.Bm < 4ull * TW) return false;
  const size_t avail = d->smem_optin - 64 - ELL_CTA_WORDS * 4;
  const int64_t avgA = std::max<int64_t>(1, (a.Annz + a.m.An - 1) / std::max(a.m.An, 1));
  // R rows per tile: as many as keep (a) the tile's A nonzeros within one 64-entry chunk on average and (b) the CTA at
  // ELL_MAX_WARPS warps (shared memory per warp grows with R; occupancy matters more than amortising the tile overhead)
  auto warps_for = [&](int r) { return (int)std::min<size_t>(ELL_MAX_WARPS - 1, avail / ((size_t)ell_warp_words(r, TW, r * h.max_len_a * W) * 4)); };
  int R = 8;
  while (R > 1 && ((int64_t)R * avgA > 64 || warps_for(R) < ELL_MAX_WARPS - 1)) R >>= 1;
  if (const char* e = getenv("BSPGEMM_ELL_R")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) R = v; }   // tuning knob
  const int warps = warps_for(R);
  if (warps < 4) return false;
  // register-sort variant: regular matrices (every row close to the padded size LA*W <= 512)
  d->use_sort = false;
  if (!getenv("BSPGEMM_NO_SORT")) {
    int lal = 2; while ((1u << lal) < h.max_len_a) ++lal;
    const int LA = 1 << lal;
    const bool fits = lal <= 5 && LA * W <= 1024;
    const bool regular = (u64)a.Annz * 2ull >= (u64)a.m.An * (u64)LA || getenv("BSPGEMM_FORCE_SORT");
    if (fits && regular) { d->use_sort = true; d->sort_LAL = lal; }
  }
  if (clustered && !d->use_sort && !getenv("BSPGEMM_FORCE_ELL")) return false;   // narrow rows: the bitmap kernels (kernels.cuh), not the global slot map
  d->use_ell = true; d->ell_W = W; d->ell_R = R; d->ell_TW = TW; d->ell_warps = warps; d->ell_maxA = h.max_len_a; d->ell_lf16 = lf16;
  return true;
}

// Banded / block-diagonal fast path (band.cuh): B rows become (first, len) descriptors, output rows 128-bit bitmaps.
static int build_desc(bspgemm_dev* d) {
  const MulArgs& a = d->a;
  d->pb.desc = false;                                  // whatever `bdesc` held is gone (bspgemm_dev_prepare_b sets it again)
  CKS(d->bdesc.ensure(((size_t)a.m.Bn + 1) * 2 + 4));
  const long long threads = (((long long)a.m.Bn + 31) / 32) * 32;
  k_build_desc<<<(int)std::max<long long>(1, (threads + 255) / 256), 256, 0, d->stream>>>(a.m.Brow, a.m.Bcol, a.m.Bn, (u32)a.m.Bm, reinterpret_cast<uint2*>(d->bdesc.p), d->d_sc);
  d->launches++;
  CK(cudaGetLastError());
  return BSPGEMM_OK;
}
static int launch_band(bspgemm_dev* d) {
  const MulArgs& a = d->a;
  int* ccol = d->user_ccol ? d->user_ccol : d->ccol.p;
  if (!(b_prepared(d) && d->pb.desc)) CKS(build_desc(d));
  uint2* desc = reinterpret_cast<uint2*>(d->bdesc.p);
  const u32 ntiles = (u32)(((size_t)a.m.An + BAND_THREADS - 1) / BAND_THREADS);
  u64* chain = nullptr;
  CKS(chain_reserve(d, (size_t)ntiles + 1, &chain));
  CK(cudaEventRecord(d->ev[3], d->stream));
  int bps = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_band, BAND_THREADS, 0));
  const int grid = (int)std::max<long long>(1, std::min<long long>((long long)ntiles, (long long)d->sm_count * std::max(bps, 1)));
  BandArgs p{};
  p.Arow = a.m.Arow; p.Acol = a.m.Acol; p.desc = desc; p.An = a.m.An; p.Bn = a.m.Bn;
  p.Crow = a.dCrow; p.is64 = a.is64; p.Ccol = ccol; p.status = chain; p.sc = d->d_sc; p.ntiles = ntiles;
  k_band<<<grid, BAND_THREADS, 0, d->stream>>>(p);
  d->launches++;
  CK(cudaGetLastError());
  return BSPGEMM_OK;
}

static int pick_group(int64_t nnz, int64_t rows) {
  const int64_t avg = rows > 0 ? (nnz + rows - 1) / rows : 1;
  return avg <= 4 ? 4 : avg <= 8 ? 8 : avg <= 16 ? 16 : 32;
}

static int launch_estimate_kernel(bspgemm_dev* d) {
  const MulArgs& a = d->a;
  const size_t An = (size_t)a.m.An;
  CKS(d->ip.ensure(An + 1));
  const int ga = pick_group(a.Annz, a.m.An);
  const long long threads = (long long)An * ga;
  const int grid = (int)((threads + 255) / 256);
  switch (ga) {
    case 4:  k_estimate<4><<<grid, 256, 0, d->stream>>>(a.m, d->ip.p, d->d_sc); break;
    case 8:  k_estimate<8><<<grid, 256, 0, d->stream>>>(a.m, d->ip.p, d->d_sc); break;
    case 16: k_estimate<16><<<grid, 256, 0, d->stream>>>(a.m, d->ip.p, d->d_sc); break;
    default: k_estimate<32><<<grid, 256, 0, d->stream>>>(a.m, d->ip.p, d->d_sc); break;
  }
  d->launches++;
  CK(cudaGetLastError());
  return BSPGEMM_OK;
}

// phase 0: longest rows of A and B (decides whether the work-estimation pass can be skipped)
static int mul_launch_probe(bspgemm_dev* d) {
  NvtxRange nvtx_("bspgemm.probe");
  const MulArgs& a = d->a;
  CK(cudaSetDevice(d->device));
  d->launches = 0;
  d->fast = false;
  memset(&d->st, 0, sizeof d->st);
  CKS(d->status.ensure(SC_WORDS + (size_t)a.m.An / 4 + 8192));      // scalars + the longest look-back chain of any pipeline
  d->d_sc = reinterpret_cast<DevScalars*>(d->status.p);
  CK(cudaEventRecord(d->ev[0], d->stream));
  CK(cudaMemsetAsync(d->d_sc, 0, sizeof(DevScalars), d->stream));
  const bool bprep = b_prepared(d);                                    // B's longest row is known (mul_launch_estimate)
  const int nmax = bprep ? a.m.An : std::max(a.m.An, a.m.Bn);
  k_maxlen<<<std::max(1, std::min((nmax + 255) / 256, d->sm_count * 8)), 256, 0, d->stream>>>(a.m.Arow, a.m.An, a.m.Brow, bprep ? 0 : a.m.Bn, d->d_sc);
  d->launches++;
  CK(cudaGetLastError());
  if (a.m.An > 0 && a.Bnnz > 0) {                                  // are the rows' columns clustered (banded / block-diagonal)?
    const int nsamples = std::min(a.m.An, 2048);
    k_probe_span<<<(nsamples * 32 + 255) / 256, 256, 0, d->stream>>>(a.m, nsamples, d->d_sc);
    d->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
  d->phase = 1;
  return BSPGEMM_OK;
}

// phase 1: work estimation (north-star step 1) — skipped when maxlen(A)*maxlen(B) bounds every row's IP
// by the S-bin capacity (then no row can leave the S bin and Σip <= nnzA*maxlen(B)).
static int mul_launch_estimate(bspgemm_dev* d) {
  NvtxRange nvtx_("bspgemm.estimate");
  const MulArgs& a = d->a;
  CK(cudaSetDevice(d->device));
  CK(cudaStreamSynchronize(d->stream));
  if (b_prepared(d)) d->h_sc->max_len_b = d->pb.max_len_b;
  const DevScalars& h = *d->h_sc;
  const u64 bound = (u64)h.max_len_a * (u64)h.max_len_b;
  d->max_len_b = h.max_len_b;
  d->skip_estimate = d->mode != BSPGEMM_MODE_TWOPHASE && bound <= (u64)g_cap_s_max() && !getenv("BSPGEMM_FORCE_ESTIMATE");
  if (ell_plan(d)) d->skip_estimate = true;
  // every sampled row is a union of runs of consecutive columns inside a 128-column window: banded / block-diagonal
  d->use_band = !d->no_band && d->mode != BSPGEMM_MODE_TWOPHASE && h.span_rows > 0 && h.span_runs == h.span_rows &&
                !getenv("BSPGEMM_NO_BAND") && !getenv("BSPGEMM_NO_ELL") && !getenv("BSPGEMM_CAP_S") && !getenv("BSPGEMM_FORCE_ESTIMATE") &&
                !getenv("BSPGEMM_FORCE_ELL") && !getenv("BSPGEMM_FORCE_SORT") && !getenv("BSPGEMM_NO_SORT");
  if (d->use_band) d->skip_estimate = true;
  d->row_ip_bound = (u32)std::min<u64>(bound, 0xfffffffeull);
  CK(cudaEventRecord(d->ev[6], d->stream));
  if (!d->skip_estimate) {
    CKS(launch_estimate_kernel(d));
    CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
  }
  CK(cudaEventRecord(d->ev[1], d->stream));
  d->phase = 2;
  return BSPGEMM_OK;
}

// phase 2: bins, symbolic, scan / fused fill
static int mul_launch_main(bspgemm_dev* d) {
  NvtxRange nvtx_("bspgemm.main");
  const MulArgs& a = d->a;
  CK(cudaSetDevice(d->device));
  const size_t An = (size_t)a.m.An;
  u64 ip_bound;                      // upper bound of nnz(C) used to size the fused output arena
  u32 max_ip;
  if (d->use_band) {
    ip_bound = std::min<u64>((u64)a.Annz * (u64)d->max_len_b, (u64)a.m.An * (u64)BAND_BITS);   // an output row has at most BAND_BITS columns
    bool fits;
    if (d->user_ccol) fits = (u64)d->user_cap >= ip_bound;
    else if (ip_bound <= (u64)d->ccol.cap) fits = true;
    else { size_t fr = 0, tot = 0; CK(cudaMemGetInfo(&fr, &tot)); fits = ip_bound * 4ull + (u64)a.m.Bn * 8ull <= (u64)d->ccol.cap * 4ull + (u64)(fr / 2); }
    if (fits) {
      d->cap_s = BAND_BITS; d->G = 1; d->have_m = d->have_m2 = d->have_l = false;
      d->st.cap_s = (int)BAND_BITS; d->st.group = 1; d->st.variant = 3; d->st.rows_per_tile = BAND_THREADS;
      d->used_mode = BSPGEMM_MODE_FUSED; d->st.mode = BSPGEMM_MODE_FUSED;
      CK(cudaEventRecord(d->ev[2], d->stream));
      if (!d->user_ccol) CKS(d->ccol.ensure((size_t)std::max<u64>(ip_bound, 1)));
      CKS(launch_band(d));                              // records ev[3] between the descriptor build and the band kernel
      CK(cudaEventRecord(d->ev[4], d->stream));
      CK(cudaEventRecord(d->ev[5], d->stream));
      CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
      d->phase = 3;
      return BSPGEMM_OK;
    }
    d->use_band = false;
    if (!d->use_ell) {
      d->skip_estimate = false;
      CKS(launch_estimate_kernel(d));
      CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
      CK(cudaEventRecord(d->ev[1], d->stream));
      return mul_launch_main(d);
    }
  }
  if (d->use_ell) {
    // ELL fast path: every row fits one warp's table by construction; needs the Σip bound to fit the output arena
    ip_bound = (u64)a.Annz * (u64)d->max_len_b;
    bool fits;
    if (d->user_ccol) fits = (u64)d->user_cap >= ip_bound;
    else if (ip_bound <= (u64)d->ccol.cap) fits = true;
    else { size_t fr = 0, tot = 0; CK(cudaMemGetInfo(&fr, &tot)); fits = ip_bound * 4ull + (u64)a.m.Bn * d->ell_W * 4ull <= (u64)d->ccol.cap * 4ull + (u64)(fr / 2); }
    if (fits) {
      d->cap_s = d->ell_TW; d->G = d->ell_W; d->have_m = d->have_m2 = d->have_l = false;
      d->st.cap_s = (int)d->ell_TW; d->st.group = d->ell_W; d->st.variant = 1; d->st.rows_per_tile = d->ell_R;
      d->used_mode = BSPGEMM_MODE_FUSED; d->st.mode = BSPGEMM_MODE_FUSED;
      CK(cudaEventRecord(d->ev[2], d->stream));
      if (!d->user_ccol) CKS(d->ccol.ensure((size_t)std::max<u64>(ip_bound, 1)));
      CKS(launch_ell(d));                               // records ev[3] between the ELL build and the fused kernel
      CK(cudaEventRecord(d->ev[4], d->stream));
      CK(cudaEventRecord(d->ev[5], d->stream));
      CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
      d->phase = 3;
      return BSPGEMM_OK;
    }
    d->use_ell = false;
    d->skip_estimate = false;
    CKS(launch_estimate_kernel(d));
    CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
    CK(cudaEventRecord(d->ev[1], d->stream));
    return mul_launch_main(d);
  }
  if (d->skip_estimate) {
    max_ip = d->row_ip_bound;
    ip_bound = (u64)a.Annz * (u64)d->max_len_b;
  } else {
    CK(cudaStreamSynchronize(d->stream));
    const DevScalars& h = *d->h_sc;
    if (h.err & 1u) return fail(BSPGEMM_ERR_BADARG, "a column index of A is outside [0,Bn=%d)", a.m.Bn);
    max_ip = h.max_ip;
    ip_bound = h.total_ip;
    d->st.ip = (int64_t)h.total_ip;
    for (int b = 0; b < 34; ++b) {
      const u64 top = b ? (1ull << (b - 1)) : 0;                       // largest IP in the bin
      d->hist_rows[b] = h.hist[b];
      (void)top;
    }
  }
  // bin thresholds
  u32 cap = 32; const u32 cap_max = (u32)g_cap_s_max();
  while (cap < max_ip && cap < cap_max) cap <<= 1;
  d->cap_s = cap;
  d->G = pick_group(a.Bnnz, a.m.Bn);
  d->G_big = d->skip_estimate ? d->G : std::max(d->G, pick_group((int64_t)ip_bound, a.Annz));   // power-law: long B rows are selected more often
  d->have_m = max_ip > cap;
  d->have_m2 = max_ip > CAP_M1;
  d->have_l = max_ip > CAP_M2;
  d->st.cap_s = (int)cap; d->st.group = d->G;
  CKS(d->cnt.ensure(An + 1));
  // Rows above 2048 products of matrices of up to BM_MAX_WINDOWS windows of columns: rows_bm.cuh (BSPGEMM_NO_BM: the round-1 kernels).
  // Its medium-row kernel keeps a row's products in registers and wants at most 1024 A entries per row: k_build_lists sends longer
  // rows to the long-row list whatever their product count, so that list may be non-empty even when max_ip <= CAP_M2.
  d->use_bm = d->have_m2 && (u64)a.m.Bm <= (u64)BM_MAX_WINDOWS * BM_WORDS * 32ull && !getenv("BSPGEMM_NO_BM");
  if (d->use_bm) d->have_l = true;
  if (d->have_m) {
    CKS(d->lists.ensure(3 * An + 3));
    CKS(d->tofs.ensure(An + 1));
    k_build_lists<<<(int)((An + 255) / 256), 256, 0, d->stream>>>(d->ip.p, a.m.An, cap, CAP_M1, CAP_M2,
        d->lists.p, d->lists.p + An, d->lists.p + 2 * An, d->tofs.p, d->d_sc, a.m.Arow, d->use_bm ? BM_CHUNK : 0xffffffffu);
    d->launches++;
    CK(cudaGetLastError());
  }
  // Matrices of up to WIN_MAX_WINDOWS windows of columns: every big row goes through the windowed bitmap kernel
  d->use_window = (u64)a.m.Bm <= (u64)WIN_MAX_WINDOWS * WIN_WORDS * 32ull && !getenv("BSPGEMM_NO_WINDOW");
  if (d->have_l && !d->use_window && !d->use_bm) {
    d->bm_words = (u32)(((size_t)a.m.Bm + 31) / 32);
    d->l_grid = d->sm_count;
    const size_t need = (size_t)d->l_grid * d->bm_words;
    if (need > d->bitmaps.cap) { CKS(d->bitmaps.ensure(need)); CK(cudaMemsetAsync(d->bitmaps.p, 0, d->bitmaps.cap * sizeof(u32), d->stream)); }
  }
  // mode
  int mode = d->mode;
  if (mode == BSPGEMM_MODE_AUTO) {
    const u64 have_now = d->user_ccol ? (u64)d->user_cap : (u64)d->ccol.cap;
    if (ip_bound <= have_now) mode = BSPGEMM_MODE_FUSED;
    else if (d->user_ccol) mode = BSPGEMM_MODE_TWOPHASE;
    else {
      size_t fr = 0, tot = 0; CK(cudaMemGetInfo(&fr, &tot));
      mode = (ip_bound * 4ull <= (u64)d->ccol.cap * 4ull + (u64)(fr / 2)) ? BSPGEMM_MODE_FUSED : BSPGEMM_MODE_TWOPHASE;
    }
  }
  if (mode == BSPGEMM_MODE_FUSED && d->user_ccol && (u64)d->user_cap < ip_bound) mode = BSPGEMM_MODE_TWOPHASE;
  if (mode == BSPGEMM_MODE_TWOPHASE && d->skip_estimate) {            // two-phase needs ip[]: run the estimate after all
    d->skip_estimate = false;
    CKS(launch_estimate_kernel(d));
    CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
    CK(cudaEventRecord(d->ev[1], d->stream));
    return mul_launch_main(d);
  }
  d->used_mode = mode; d->st.mode = mode;
  CK(cudaEventRecord(d->ev[2], d->stream));
  if (mode == BSPGEMM_MODE_FUSED) {
    if (!d->user_ccol) CKS(d->ccol.ensure((size_t)std::max<u64>(ip_bound, 1)));
    // Big rows in ONE pass when memory allows: count + sorted row into a staging arena (Σ of their IP <= Σip words), moved to
    // Ccol once the fused kernel has produced the row pointers — instead of a symbolic and a numeric pass that both gather
    // and de-duplicate the row.
    bool staged = false;
    if (d->have_m && (!d->have_l || d->use_window || d->use_bm) && !getenv("BSPGEMM_NO_STAGE")) {
      const size_t need = (size_t)std::max<u64>(ip_bound, 1) + 3u * (size_t)An + 4u;      // every staged row is rounded up to 4 words
      if (d->temp.cap >= need) staged = true;
      else {
        size_t fr = 0, tot = 0; CK(cudaMemGetInfo(&fr, &tot));
        if ((u64)fr + (u64)d->temp.cap * 4ull >= (u64)need * 4ull + (2ull << 30)) { d->temp.release(); if (d->temp.ensure(need) == BSPGEMM_OK) staged = true; }
      }
    }
    d->staged = staged;
    if (d->have_m) CKS(staged ? launch_bins_ml<MODE_STAGE>(d) : launch_bins_ml<MODE_COUNT>(d));
    CK(cudaEventRecord(d->ev[3], d->stream));
    // ... and matrices whose rows are all small and cheap (the report's sprand matrices: Poisson(5) rows, 25 products per row):
    // the ordered kernel is bound by its look-back chain there (1.25 M tiles of 4 rows: 24 ms at n = 5e6), the two unordered
    // passes by the rows themselves.
    const bool cheap = !d->have_m && An >= 65536 && ip_bound / (u64)An <= 96ull;
    if (((staged && !d->skip_estimate) || cheap) && !getenv("BSPGEMM_FUSED_S")) {
      if (d->skip_estimate) {
        CKS(launch_estimate_kernel(d));                               // the warp-per-row kernels read ip[]
        if (cheap) {
          // The bin capacity was sized from the bound max_len(A) x max_len(B) (Poisson(5) rows: ~20 x 20 -> 512), the rows hold a
          // fraction of it (max ~110 -> 128): with the real maximum the per-warp table is 2 KB instead of 6.6 KB and twice as many
          // rows are in flight per SM (64 instead of 32 warps) — one 20 us host round trip against milliseconds.
          CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
          CKS(wait_stream(d->stream));
          u32 capr = 32; while (capr < d->h_sc->max_ip) capr <<= 1;
          if (capr < d->cap_s) { d->cap_s = capr; d->st.cap_s = (int)capr; }
        }
      }
      // Skewed matrices (big rows exist and were just staged with their counts): the S rows are counted and filled in two
      // UNORDERED passes around the device scan instead of the ordered one-pass kernel.  Their intermediate products are a
      // few percent of the total (R-MAT scale 22: 3e8 of 1.2e10), so walking them twice is cheap, while k_fused's in-order
      // look-back made every warp wait for the slowest earlier tile when row costs differ by orders of magnitude (85 % of its
      // warp time in the spin, profiles/r02_rmat20_kfused_ncu_summary.txt: 112 of 527 ms per step at config 4).
      d->st.kernel_flags |= 4;                                       // small rows: count -> scan -> fill (bench.py labels the kernel from this)
      CKS(launch_rows_warp<MODE_COUNT>(d));
      const u32 nt = (u32)((An + SCAN_THREADS * SCAN_ITEMS - 1) / (SCAN_THREADS * SCAN_ITEMS));
      u64* chain = nullptr;
      CKS(chain_reserve(d, (size_t)nt + 1, &chain));
      k_scan<<<nt, SCAN_THREADS, 0, d->stream>>>(d->cnt.p, a.m.An, a.dCrow, a.is64, chain, d->d_sc, nt);
      d->launches++;
      CK(cudaGetLastError());
      CKS(launch_rows_warp<MODE_FILL>(d));
      CK(cudaEventRecord(d->ev[4], d->stream));
      if (staged) CKS(launch_copy_rows(d));
      CK(cudaEventRecord(d->ev[5], d->stream));
      CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
      d->phase = 3;
      return BSPGEMM_OK;
    }
    const u32 rows_per_tile = FUSED_R;
    const u32 ntiles = (u32)((An + rows_per_tile - 1) / rows_per_tile);
    { u64* chain = nullptr; CKS(chain_reserve(d, (size_t)ntiles + 1, &chain)); }
    CKS(launch_fused(d, ntiles, d->skip_estimate ? 1 : 0));
    CK(cudaEventRecord(d->ev[4], d->stream));
    if (d->have_m) CKS(staged ? launch_copy_rows(d) : launch_bins_ml<MODE_FILL>(d));
    CK(cudaEventRecord(d->ev[5], d->stream));
  } else {
    CKS(launch_rows_warp<MODE_COUNT>(d));
    if (d->have_m) CKS(launch_bins_ml<MODE_COUNT>(d));
    CK(cudaEventRecord(d->ev[3], d->stream));
    const u32 ntiles = (u32)((An + SCAN_THREADS * SCAN_ITEMS - 1) / (SCAN_THREADS * SCAN_ITEMS));
    u64* chain = nullptr;
    CKS(chain_reserve(d, (size_t)ntiles + 1, &chain));
    k_scan<<<ntiles, SCAN_THREADS, 0, d->stream>>>(d->cnt.p, a.m.An, a.dCrow, a.is64, chain, d->d_sc, ntiles);
    d->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
  d->phase = 3;
  return BSPGEMM_OK;
}

// Wait for the product's stream.  cudaStreamSynchronize parks the host thread when the context schedules blocking syncs (what a
// framework that owns the context may have chosen), and the wake-up is then tens of microseconds — a tenth of a 0.45 ms step at 8
// GPUs.  The stream is polled instead for the duration of a short product, then the thread blocks like before.
static int wait_stream(cudaStream_t s) {
  const auto t0 = std::chrono::steady_clock::now();
  for (int spins = 0;; ++spins) {
    const cudaError_t e = cudaStreamQuery(s);
    if (e == cudaSuccess) { if (spins) (void)cudaGetLastError(); return BSPGEMM_OK; }     // (a "not ready" answer may linger as the last error)
    if (e != cudaErrorNotReady) { CK(e); }
    if ((spins & 63) == 63 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(20)) break;
  }
  CK(cudaStreamSynchronize(s));
  return BSPGEMM_OK;
}

// phase 3 (two-phase mode only): numeric fill at the scanned row pointers
static int mul_launch_fill(bspgemm_dev* d) {
  NvtxRange nvtx_("bspgemm.fill");
  const MulArgs& a = d->a;
  CK(cudaSetDevice(d->device));
  CKS(wait_stream(d->stream));
  const DevScalars& h = *d->h_sc;
  if (!d->fast && d->use_band && h.band_fail) {
    // the optimistic run/bitmap kernel met a B row that is not a run of consecutive columns, or an output row wider than
    // its register bitmap: nothing it wrote is used — the whole product is redone by the general kernels
    d->no_band = true;
    int rc = mul_launch_probe(d);
    if (rc == BSPGEMM_OK) rc = mul_launch_estimate(d);
    if (rc == BSPGEMM_OK) rc = mul_launch_main(d);
    if (rc == BSPGEMM_OK) rc = mul_launch_fill(d);
    d->no_band = false;
    return rc;
  }
  if (d->fast && ((h.err & 8u) || (d->use_band && h.band_fail))) {
    // the cached plan does not fit this A (a row longer than the plan's LA / an output row wider than the register bitmap):
    // nothing the kernel wrote is used — forget the plan and redo the product with fresh probes
    d->pb.variant = -1;
    int rc = mul_launch_probe(d);
    if (rc == BSPGEMM_OK) rc = mul_launch_estimate(d);
    if (rc == BSPGEMM_OK) rc = mul_launch_main(d);
    if (rc == BSPGEMM_OK) rc = mul_launch_fill(d);
    return rc;
  }
  if (h.err & 1u) return fail(BSPGEMM_ERR_BADARG, "a column index of A is outside [0,Bn=%d)", a.m.Bn);
  if (h.err & 4u) return fail(BSPGEMM_ERR_BADARG, "a column index of B is outside [0,Bm=%d)", a.m.Bm);
  if (h.err & 2u) return fail(BSPGEMM_ERR_OVERFLOW32, "nnz(C) = %llu does not fit 32-bit row pointers", (unsigned long long)h.total_nnz);
  d->st.nnz = (int64_t)h.total_nnz;
  d->st.ip = (int64_t)h.total_ip;
  if (b_prepared(d) && !d->fast) {            // remember what this B and this kind of A needed: the next product starts from it
    d->pb.variant = (d->st.variant == 2 || d->st.variant == 3) ? d->st.variant : -1;
    d->pb.sort_LAL = d->sort_LAL;
  }
  if (d->used_mode == BSPGEMM_MODE_FUSED) { d->phase = 5; return BSPGEMM_OK; }
  if (d->user_ccol) { if ((u64)d->user_cap < h.total_nnz) return fail(BSPGEMM_ERR_CAPACITY, "output capacity %lld < nnz(C) %llu", (long long)d->user_cap, (unsigned long long)h.total_nnz); }
  else CKS(d->ccol.ensure((size_t)std::max<u64>(h.total_nnz, 1)));
  CKS(launch_rows_warp<MODE_FILL>(d));
  CK(cudaEventRecord(d->ev[4], d->stream));
  if (d->have_m) CKS(launch_bins_ml<MODE_FILL>(d));
  CK(cudaEventRecord(d->ev[5], d->stream));
  d->phase = 4;
  return BSPGEMM_OK;
}

static int mul_finish(bspgemm_dev* d) {
  const MulArgs& a = d->a;
  CK(cudaSetDevice(d->device));
  if (d->phase == 4) { CK(cudaStreamSynchronize(d->stream)); d->phase = 5; }
  CK(cudaGetLastError());
  bspgemm_stats& s = d->st;
  s.launches = d->launches;
  if (d->skip_estimate) s.rows_s = a.m.An;
  else for (int b = 0; b < 34; ++b) {
    const u64 top = b ? (1ull << (b - 1)) : 0;                         // largest IP in the bin
    if (top <= d->cap_s) s.rows_s += d->hist_rows[b]; else if (top <= CAP_M2) s.rows_m += d->hist_rows[b]; else s.rows_l += d->hist_rows[b];
  }
  float ms = 0;
  s.plan_cached = d->fast ? 1 : 0;
  s.b_prepared = b_prepared(d) ? 1 : 0;
  if (d->fast) {                               // replayed from the cached plan: one kernel between ev[3] and ev[4]
    cudaEventElapsedTime(&ms, d->ev[0], d->ev[4]); s.ms_total = ms;
    cudaEventElapsedTime(&ms, d->ev[3], d->ev[4]); s.ms_main = ms;
  } else {
    cudaEventElapsedTime(&ms, d->ev[0], d->ev[5]); s.ms_total = ms;
    cudaEventElapsedTime(&ms, d->ev[6], d->ev[1]); s.ms_estimate = ms;
    cudaEventElapsedTime(&ms, d->ev[2], d->ev[3]); s.ms_symbolic = ms;
    cudaEventElapsedTime(&ms, d->ev[3], d->ev[4]); s.ms_main = ms;
    cudaEventElapsedTime(&ms, d->ev[4], d->ev[5]); s.ms_numeric = ms;
  }
  const int64_t rp = a.is64 ? 8 : 4;
  s.algorithmic_bytes = 4 * ((int64_t)a.m.An + 1) + 12 * a.Annz + 4 * s.ip + 4 * s.nnz + rp * ((int64_t)a.m.An + 1);
  return BSPGEMM_OK;
}

// Product with a prepared B whose last product ran the sorting-network or the band kernel: launched straight from that plan —
// no probe kernels, no ELL / descriptor build, no host round trip before the launch: one memset (scalars + look-back words),
// the kernel, the read-back of the scalars.  What the probes would have established is checked where it is used: the kernels
// flag a row of A longer than the plan's LA (err bit 3) or an output row the band kernel cannot hold (band_fail), and
// mul_launch_fill then redoes the product the long way.  The host-side conditions of the plan are re-evaluated here.
static bool env_overrides() {
  static const char* names[] = {"BSPGEMM_NO_ELL", "BSPGEMM_CAP_S", "BSPGEMM_FORCE_ESTIMATE", "BSPGEMM_FORCE_ELL", "BSPGEMM_NO_SORT", "BSPGEMM_FORCE_SORT",
                                "BSPGEMM_NO_BAND", "BSPGEMM_NO_PLAN_CACHE"};
  for (const char* n : names) if (getenv(n)) return true;
  return false;
}
static int mul_launch_fast(bspgemm_dev* d, bool* taken) {
  NvtxRange nvtx_("bspgemm.fast");
  const MulArgs& a = d->a;
  *taken = false;
  if (!b_prepared(d) || d->pb.variant < 0 || d->mode == BSPGEMM_MODE_TWOPHASE || d->user_ccol || env_overrides()) return BSPGEMM_OK;
  const auto& pb = d->pb;
  u64 ip_bound;
  if (pb.variant == 2) {
    const int LA = 1 << pb.sort_LAL;
    if (pb.ell_W == 0 || (u64)a.Annz * 2ull < (u64)a.m.An * (u64)LA) return BSPGEMM_OK;      // "regular" rule of ell_plan
    if (d->ell_pad != sort_pad_for(pb.ell_W, pb.sort_LAL, a.m.Bm)) return BSPGEMM_OK;          // the ELL copy's padding must be this plan's kernel's
    ip_bound = (u64)a.Annz * (u64)pb.max_len_b;
  } else {
    if (!pb.desc) return BSPGEMM_OK;
    ip_bound = std::min<u64>((u64)a.Annz * (u64)pb.max_len_b, (u64)a.m.An * (u64)BAND_BITS);
  }
  if (ip_bound > (u64)d->ccol.cap) return BSPGEMM_OK;      // the arena of the earlier product is reused, never grown here
  CK(cudaSetDevice(d->device));
  d->launches = 0;
  memset(&d->st, 0, sizeof d->st);
  d->fast = true;
  d->no_band = false;
  d->skip_estimate = true; d->max_len_b = pb.max_len_b;
  d->have_m = d->have_m2 = d->have_l = false;
  d->used_mode = BSPGEMM_MODE_FUSED; d->st.mode = BSPGEMM_MODE_FUSED;
  CKS(d->status.ensure(SC_WORDS + (size_t)a.m.An / 4 + 8192));
  d->d_sc = reinterpret_cast<DevScalars*>(d->status.p);
  CK(cudaEventRecord(d->ev[0], d->stream));
  if (pb.variant == 2) {
    d->use_band = false; d->use_ell = true; d->use_sort = true; d->ell_W = pb.ell_W; d->sort_LAL = pb.sort_LAL;
    d->cap_s = 0; d->G = pb.ell_W;
    d->st.cap_s = (1 << pb.sort_LAL) * pb.ell_W; d->st.group = pb.ell_W; d->st.variant = 2;
    CKS(launch_sort(d, d->ccol.p));                       // chain_reserve clears scalars + chain, records ev[3]
  } else {
    d->use_band = true; d->use_ell = false; d->use_sort = false;
    d->cap_s = BAND_BITS; d->G = 1;
    d->st.cap_s = (int)BAND_BITS; d->st.group = 1; d->st.variant = 3; d->st.rows_per_tile = BAND_THREADS;
    CKS(launch_band(d));
  }
  CK(cudaEventRecord(d->ev[4], d->stream));
  CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
  d->phase = 3;
  *taken = true;
  return BSPGEMM_OK;
}

static int mul_run_to_completion(bspgemm_dev* d) {
  if (d->a.m.An == 0) {   // nothing to do: Crow[0] = 0
    CK(cudaSetDevice(d->device));
    CK(cudaMemsetAsync(d->a.dCrow, 0, d->a.is64 ? 8 : 4, d->stream));
    CK(cudaStreamSynchronize(d->stream));
    memset(&d->st, 0, sizeof d->st); d->phase = 5;
    return BSPGEMM_OK;
  }
  bool fast = false;
  CKS(mul_launch_fast(d, &fast));
  if (!fast) {
    CKS(mul_launch_probe(d));
    CKS(mul_launch_estimate(d));
    CKS(mul_launch_main(d));
  }
  CKS(mul_launch_fill(d));
  return mul_finish(d);
}

static int dev_create(bspgemm_dev** out, int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return fail(BSPGEMM_ERR_NOGPU, "no CUDA device visible; this library has no CPU fallback"); }
  if (device < 0 || device >= n) return fail(BSPGEMM_ERR_BADARG, "device %d out of range (0..%d)", device, n - 1);
  CK(cudaSetDevice(device));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, device));
  if (p.major < 10) return fail(BSPGEMM_ERR_NOGPU, "device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
  // L2 fetch granularity: the gathers of B rows are 64-byte (config 3) random accesses; the device default fetches more than the
  // row on a miss (measured: no effect, profiles/r02_sweeps.txt item 2).  BSPGEMM_L2_FETCH=32|64|128 overrides (device-wide limit).
  if (const char* e = getenv("BSPGEMM_L2_FETCH")) { const int v = atoi(e); if (v == 32 || v == 64 || v == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v); cudaGetLastError(); }
  bspgemm_dev* d = new bspgemm_dev();
  d->device = device; d->sm_count = p.multiProcessorCount; d->smem_optin = p.sharedMemPerBlockOptin;
  CK(cudaStreamCreateWithFlags(&d->own_stream, cudaStreamNonBlocking));
  d->stream = d->own_stream;
  CKS(d->status.ensure(SC_WORDS + 8192));
  d->d_sc = reinterpret_cast<DevScalars*>(d->status.p);
  CK(cudaMallocHost((void**)&d->h_sc, sizeof(DevScalars)));
  for (auto& e : d->ev) CK(cudaEventCreate(&e));
  CKS(set_kernel_attributes((int)d->smem_optin));
  const char* m = getenv("BSPGEMM_MODE");
  if (m) d->mode = !strcmp(m, "fused") ? BSPGEMM_MODE_FUSED : !strcmp(m, "twophase") ? BSPGEMM_MODE_TWOPHASE : BSPGEMM_MODE_AUTO;
  *out = d;
  return BSPGEMM_OK;
}

static void dev_destroy(bspgemm_dev* d) {
  if (!d) return;
  cudaSetDevice(d->device);
  cudaStreamSynchronize(d->stream);
  d->ip.release(); d->cnt.release(); d->lists.release(); d->bitmaps.release(); d->status.release(); d->ccol.release(); d->temp.release(); d->tofs.release(); d->bell.release(); d->bdesc.release();
  d->m_crow.release(); d->m_frow.release(); d->m_fcol.release(); d->m_out.release();
  d->in_arow.release(); d->in_acol.release(); d->in_brow.release(); d->in_bcol.release(); d->crow_dev.release(); d->crow_tmp.release();
  if (d->h_sc) cudaFreeHost(d->h_sc);
  for (auto& e : d->ev) if (e) cudaEventDestroy(e);
  if (d->own_stream) cudaStreamDestroy(d->own_stream);
  delete d;
}

// ------------------------------------------------------------------------------------------------ device-resident C ABI
extern "C" int bspgemm_dev_create(bspgemm_dev** h, int device) { if (!h) return fail(BSPGEMM_ERR_BADARG, "null handle"); return dev_create(h, device); }
extern "C" int bspgemm_dev_destroy(bspgemm_dev* h) { dev_destroy(h); return BSPGEMM_OK; }
extern "C" int bspgemm_dev_set_mode(bspgemm_dev* h, int mode) { if (!h || mode < 0 || mode > 2) return fail(BSPGEMM_ERR_BADARG, "bad mode"); h->mode = mode; return BSPGEMM_OK; }
extern "C" int bspgemm_dev_get_stats(bspgemm_dev* h, bspgemm_stats* out) { if (!h || !out) return fail(BSPGEMM_ERR_BADARG, "null"); *out = h->st; return BSPGEMM_OK; }

extern "C" int bspgemm_dev_multiply(bspgemm_dev* h, void* stream,
                                    const int* dAcol, const int* dArow, int An, int64_t Annz,
                                    const int* dBcol, const int* dBrow, int Bn, int Bm, int64_t Bnnz,
                                    void* dCrow, int crow_is_i64, int** dCcol_out, int64_t* nnz_out) {
  if (!h || !dArow || !dBrow || !dCrow || An < 0 || Bn < 0 || Bm < 0 || Annz < 0 || Bnnz < 0) return fail(BSPGEMM_ERR_BADARG, "null pointer or negative size");
  if ((Annz > 0 && !dAcol) || (Bnnz > 0 && !dBcol)) return fail(BSPGEMM_ERR_BADARG, "null column array");
  // NULL is the CUDA legacy default stream (what torch's default stream reports): the launches are then ordered after the
  // caller's earlier work on it (the kernels that produced dArow/dAcol, the fill of dCrow); the handle's private
  // non-blocking stream is only used by the host-pointer operators, which own their buffers.
  h->stream = stream ? (cudaStream_t)stream : cudaStreamLegacy;
  h->a.m = Csr{dArow, dAcol, dBrow, dBcol, An, Bn, Bm};
  h->a.Annz = Annz; h->a.Bnnz = Bnnz; h->a.dCrow = dCrow; h->a.is64 = crow_is_i64 ? 1 : 0;
  h->user_ccol = nullptr; h->user_cap = 0;
  CKS(mul_run_to_completion(h));
  if (dCcol_out) *dCcol_out = h->ccol.p;
  if (nnz_out) *nnz_out = h->st.nnz;
  return BSPGEMM_OK;
}

// ------------------------------------------------------------------------------------------------ masked product
// C = F .* (A·B): replaces SpGEMM_masked (final/SpGEMM_mpi_omp.c:232-288).  The unmasked rows come from the product pipeline
// (ascending, distinct, in the arena, 64-bit row pointers in m_crow); the mask is applied by k_mask_rows (count -> k_scan ->
// fill) into m_out.  A mask with unsorted / repeated columns (legal for the reference, whose mask is a flag array) is first
// canonicalised with the product kernels themselves: F' = I·F.
static int masked_multiply(bspgemm_dev* d, const int* dFcol, const int* dFrow, int64_t Fnnz, void* dCrow, int is64, int** dCcol_out, int64_t* nnz_out) {
  const MulArgs user = d->a;                             // A, B as given; dCrow / is64 below are the caller's
  const int An = user.m.An, Bm = user.m.Bm;
  if (Fnnz < 0 || !dFrow || (Fnnz > 0 && !dFcol)) return fail(BSPGEMM_ERR_BADARG, "masked product: null mask");
  CKS(d->m_crow.ensure((size_t)An + 1));
  if (An == 0) {
    CK(cudaMemsetAsync(dCrow, 0, is64 ? 8 : 4, d->stream)); CK(cudaStreamSynchronize(d->stream));
    memset(&d->st, 0, sizeof d->st);
    if (dCcol_out) *dCcol_out = d->m_out.p;
    if (nnz_out) *nnz_out = 0;
    return BSPGEMM_OK;
  }
  // 1. is the mask canonical (rows strictly ascending, columns in range)?
  CK(cudaMemsetAsync(d->d_sc, 0, sizeof(DevScalars), d->stream));
  k_mask_check<<<(An + 7) / 8, 256, 0, d->stream>>>(dFrow, dFcol, An, (u32)Bm, d->d_sc);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
  CK(cudaStreamSynchronize(d->stream));
  if (d->h_sc->err & 32u) return fail(BSPGEMM_ERR_BADARG, "a column index of the mask is outside [0,Bm=%d)", Bm);
  const int* frow = dFrow; const int* fcol = dFcol;
  if (d->h_sc->err & 16u) {
    // F' = I·F: the product kernels sort and de-duplicate every row of F.  I = identity of order An, built in m_out / m_crow.
    CKS(d->m_out.ensure((size_t)An));
    CKS(d->m_frow.ensure((size_t)An + 1));
    k_iota<<<(An + 1 + 255) / 256, 256, 0, d->stream>>>(d->m_frow.p, An + 1);       // row pointers 0..An
    k_iota<<<(An + 255) / 256, 256, 0, d->stream>>>(d->m_out.p, An);                 // columns 0..An-1
    CK(cudaGetLastError());
    d->a.m = Csr{d->m_frow.p, d->m_out.p, dFrow, dFcol, An, An, Bm};
    d->a.Annz = An; d->a.Bnnz = Fnnz; d->a.dCrow = d->m_crow.p; d->a.is64 = 1;
    CKS(mul_run_to_completion(d));
    const int64_t n1 = d->st.nnz;
    if (n1 > 0x7fffffffLL) return fail(BSPGEMM_ERR_OVERFLOW32, "mask too large");
    CKS(d->m_fcol.ensure((size_t)std::max<int64_t>(n1, 1)));
    CK(cudaMemcpyAsync(d->m_fcol.p, d->ccol.p, (size_t)n1 * 4, cudaMemcpyDeviceToDevice, d->stream));
    k_narrow_rowptr<<<(An + 1 + 255) / 256, 256, 0, d->stream>>>(d->m_crow.p, d->m_frow.p, An + 1);
    CK(cudaGetLastError());
    frow = d->m_frow.p; fcol = d->m_fcol.p;
  }
  // 2. the unmasked product, rows in the arena, 64-bit row pointers
  d->a = user;
  d->a.dCrow = d->m_crow.p; d->a.is64 = 1;
  CKS(mul_run_to_completion(d));
  const bspgemm_stats prod = d->st;
  // 3. count, scan, fill
  CKS(d->cnt.ensure((size_t)An + 1));
  const int grid = std::max(1, std::min((An + 7) / 8, d->sm_count * 16));
  k_mask_rows<MODE_COUNT><<<grid, 256, 0, d->stream>>>(frow, fcol, d->m_crow.p, d->ccol.p, An, d->cnt.p, nullptr, 0, nullptr);
  CK(cudaGetLastError());
  const u32 nt = (u32)(((size_t)An + SCAN_THREADS * SCAN_ITEMS - 1) / (SCAN_THREADS * SCAN_ITEMS));
  d->fast = false;
  CK(cudaMemsetAsync(d->d_sc, 0, sizeof(DevScalars), d->stream));
  u64* chain = nullptr;
  CKS(chain_reserve(d, (size_t)nt + 1, &chain));
  k_scan<<<nt, SCAN_THREADS, 0, d->stream>>>(d->cnt.p, An, dCrow, is64, chain, d->d_sc, nt);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
  CK(cudaStreamSynchronize(d->stream));
  if (d->h_sc->err & 2u) return fail(BSPGEMM_ERR_OVERFLOW32, "nnz(C) = %llu does not fit 32-bit row pointers", (unsigned long long)d->h_sc->total_nnz);
  const int64_t nnz = (int64_t)d->h_sc->total_nnz;
  CKS(d->m_out.ensure((size_t)std::max<int64_t>(nnz, 1)));
  k_mask_rows<MODE_FILL><<<grid, 256, 0, d->stream>>>(frow, fcol, d->m_crow.p, d->ccol.p, An, nullptr, dCrow, is64, d->m_out.p);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(d->stream));
  d->st = prod; d->st.nnz = nnz; d->st.launches = prod.launches + 4;
  if (dCcol_out) *dCcol_out = d->m_out.p;
  if (nnz_out) *nnz_out = nnz;
  return BSPGEMM_OK;
}

extern "C" int bspgemm_dev_multiply_masked(bspgemm_dev* h, void* stream,
                                           const int* dAcol, const int* dArow, int An, int64_t Annz,
                                           const int* dBcol, const int* dBrow, int Bn, int Bm, int64_t Bnnz,
                                           const int* dFcol, const int* dFrow, int64_t Fnnz,
                                           void* dCrow, int crow_is_i64, int** dCcol_out, int64_t* nnz_out) {
  if (!h || !dArow || !dBrow || !dCrow || An < 0 || Bn < 0 || Bm < 0 || Annz < 0 || Bnnz < 0) return fail(BSPGEMM_ERR_BADARG, "null pointer or negative size");
  if ((Annz > 0 && !dAcol) || (Bnnz > 0 && !dBcol)) return fail(BSPGEMM_ERR_BADARG, "null column array");
  CK(cudaSetDevice(h->device));
  h->stream = stream ? (cudaStream_t)stream : cudaStreamLegacy;
  h->a.m = Csr{dArow, dAcol, dBrow, dBcol, An, Bn, Bm};
  h->a.Annz = Annz; h->a.Bnnz = Bnnz; h->a.dCrow = dCrow; h->a.is64 = crow_is_i64 ? 1 : 0;
  h->user_ccol = nullptr; h->user_cap = 0;
  return masked_multiply(h, dFcol, dFrow, Fnnz, dCrow, crow_is_i64 ? 1 : 0, dCcol_out, nnz_out);
}

// B prepared once for many products.  The reference replicates B once, outside its timed region (every rank parses the file,
// final/SpGEMM_mpi_omp.c:309, before the loop :318-328); on the GPU "B resident" includes its gather-friendly copy: the ELL
// re-layout (rows sorted, 4W-byte aligned) or, when every row is a run of consecutive columns, the (first, len) descriptors.
extern "C" int bspgemm_dev_prepare_b(bspgemm_dev* h, void* stream, const int* dBcol, const int* dBrow, int Bn, int Bm, int64_t Bnnz) {
  if (!h || !dBrow || Bn < 0 || Bm < 0 || Bnnz < 0 || (Bnnz > 0 && !dBcol)) return fail(BSPGEMM_ERR_BADARG, "prepare_b: null pointer or negative size");
  bspgemm_dev* d = h;
  CK(cudaSetDevice(d->device));
  d->stream = stream ? (cudaStream_t)stream : cudaStreamLegacy;
  d->pb = bspgemm_dev::PreparedB{};
  d->fast = false;
  d->a = MulArgs{};
  d->a.m = Csr{dBrow, dBcol, dBrow, dBcol, 0, Bn, Bm};      // only the B side is used below
  d->a.Bnnz = Bnnz;
  d->d_sc = reinterpret_cast<DevScalars*>(d->status.p);
  CK(cudaMemsetAsync(d->d_sc, 0, sizeof(DevScalars), d->stream));
  if (Bn > 0) {
    k_maxlen<<<std::max(1, std::min((Bn + 255) / 256, d->sm_count * 8)), 256, 0, d->stream>>>(dBrow, 0, dBrow, Bn, d->d_sc);
    CK(cudaGetLastError());
    if (Bnnz > 0) CKS(build_desc(d));                         // proves (or refutes) "every row is a run" and validates the columns of run rows
  }
  CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
  CK(cudaStreamSynchronize(d->stream));
  const DevScalars hs = *d->h_sc;
  auto& pb = d->pb;
  pb.brow = dBrow; pb.bcol = dBcol; pb.Bn = Bn; pb.Bm = Bm; pb.Bnnz = Bnnz; pb.max_len_b = hs.max_len_b;
  if (Bn > 0 && Bnnz > 0 && hs.band_fail == 0) pb.desc = true;
  else if (hs.max_len_b >= 1 && hs.max_len_b <= 32) {
    int W = 4; while (W < (int)hs.max_len_b) W <<= 1;
    if ((u64)Bn * (u64)W <= 4ull * (u64)Bnnz + 4096ull) {     // the padding rule of ell_plan
      CK(cudaMemsetAsync(d->d_sc, 0, sizeof(DevScalars), d->stream));
      // padding: what the sort kernel of the likely plan wants — only rows of 16+ columns reach the big-tile geometries that have
      // the floating-point network (fused_sort.cuh); a first product that needs the other value rebuilds the copy once
      CKS(build_ell(d, W, true, (W >= 16 && (u32)Bm <= (1u << 23) && !getenv("BSPGEMM_SORT_INT")) ? 0x3F800000u : EMPTY));
      CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
      CK(cudaStreamSynchronize(d->stream));
      if (d->h_sc->err & 4u) return fail(BSPGEMM_ERR_BADARG, "a column index of B is outside [0,Bm=%d)", Bm);
      pb.ell_W = W;
    }
  }
  pb.valid = true;
  return BSPGEMM_OK;
}
extern "C" int bspgemm_dev_forget_b(bspgemm_dev* h) {
  if (!h) return fail(BSPGEMM_ERR_BADARG, "null handle");
  h->pb = bspgemm_dev::PreparedB{};
  return BSPGEMM_OK;
}

// ------------------------------------------------------------------------------------------------ global context (the "communicator")
struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
struct Global {
  std::vector<bspgemm_dev*> devs;
  std::vector<ncclComm_t> comms;
  Nccl nccl;
  bool inited = false;
};
static Global g;
static std::mutex g_mu;

static int nccl_load(Nccl& n) {
  // The drivers' stdout is the reference's CSV line and nothing else, but NCCL logs to stdout by default and ignores
  // NCCL_DEBUG_FILE at level VERSION (which some boxes export): VERSION becomes WARN (same banner, no further output unless
  // something goes wrong), and whatever level is asked for, the log goes to stderr unless the caller chose a file.
  const char* dbg = getenv("NCCL_DEBUG");
  if (dbg && !strcasecmp(dbg, "VERSION")) setenv("NCCL_DEBUG", "WARN", 1);
  if (!getenv("NCCL_DEBUG_FILE")) setenv("NCCL_DEBUG_FILE", "/dev/stderr", 1);
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) { n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (n.lib) break; }
  if (!n.lib) return fail(BSPGEMM_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name) do { *(void**)(&n.field) = dlsym(n.lib, name); if (!n.field) return fail(BSPGEMM_ERR_NCCL, "libnccl lacks %s", name); } while (0)
  SYM(CommInitAll, "ncclCommInitAll"); SYM(CommDestroy, "ncclCommDestroy"); SYM(Broadcast, "ncclBroadcast"); SYM(AllGather, "ncclAllGather");
  SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return BSPGEMM_OK;
}
#define NK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(BSPGEMM_ERR_NCCL, "%s failed: %s", #call, g.nccl.GetErrorString ? g.nccl.GetErrorString(r_) : "?"); } while (0)

// One context ("task") per entry of the device list.  A device may be listed more than once (BSPGEMM_DEVICES=0,0,0 or
// bspgemm_init_devices): several row-block shards then share one GPU — no NCCL communicator (NCCL refuses duplicate devices);
// B is uploaded once per distinct device and shared by its shards.  This is how the sharding / displacement / gather path is
// exercised on a single-GPU box; with distinct devices B is replicated by ncclBroadcast over NVLink.
static int init_devices_locked(const int* devices, int ngpus) {
  if (g.inited) return fail(BSPGEMM_ERR_STATE, "bspgemm_init called twice");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return fail(BSPGEMM_ERR_NOGPU, "no CUDA device visible; this library has no CPU fallback"); }
  std::vector<int> ids;
  if (devices) ids.assign(devices, devices + ngpus);
  else if (const char* e = getenv("BSPGEMM_DEVICES")) {        // explicit task -> device map, overrides the count
    for (const char* q = e; *q;) { char* end; const long v = strtol(q, &end, 10); if (end == q) break; ids.push_back((int)v); q = (*end == ',') ? end + 1 : end; }
    if (ids.empty()) return fail(BSPGEMM_ERR_BADARG, "BSPGEMM_DEVICES=\"%s\" holds no device id", e);
  } else {
    if (ngpus <= 0) ngpus = n;
    if (ngpus > n) return fail(BSPGEMM_ERR_BADARG, "%d GPUs requested, %d visible", ngpus, n);
    for (int i = 0; i < ngpus; ++i) ids.push_back(i);
  }
  ngpus = (int)ids.size();
  bool distinct = true;
  for (int i = 0; i < ngpus; ++i) for (int k = 0; k < i; ++k) if (ids[i] == ids[k]) distinct = false;
  auto undo = [&](int st) { for (auto* x : g.devs) dev_destroy(x); g.devs.clear(); g.comms.clear(); return st; };   // nothing half-built stays behind
  for (int i = 0; i < ngpus; ++i) {
    bspgemm_dev* d = nullptr;
    int s = dev_create(&d, ids[i]);
    if (s != BSPGEMM_OK) return undo(s);
    g.devs.push_back(d);
  }
  if (ngpus > 1 && distinct) {
    int s = nccl_load(g.nccl);
    if (s == BSPGEMM_OK) {
      g.comms.assign(ngpus, nullptr);
      const ncclResult_t r = g.nccl.CommInitAll(g.comms.data(), ngpus, ids.data());
      if (r != ncclSuccess) s = fail(BSPGEMM_ERR_NCCL, "ncclCommInitAll failed: %s", g.nccl.GetErrorString ? g.nccl.GetErrorString(r) : "?");
    }
    if (s != BSPGEMM_OK) return undo(s);
  }
  g.inited = true;
  return BSPGEMM_OK;
}
extern "C" int bspgemm_init(int ngpus) { std::lock_guard<std::mutex> lk(g_mu); return init_devices_locked(nullptr, ngpus); }
extern "C" int bspgemm_init_devices(const int* devices, int ngpus) {
  if (!devices || ngpus <= 0) return fail(BSPGEMM_ERR_BADARG, "null device list");
  std::lock_guard<std::mutex> lk(g_mu); return init_devices_locked(devices, ngpus);
}

extern "C" int bspgemm_finalize(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g.inited) return BSPGEMM_OK;
  for (auto c : g.comms) if (g.nccl.CommDestroy) g.nccl.CommDestroy(c);
  g.comms.clear();
  for (auto* d : g.devs) dev_destroy(d);
  g.devs.clear();
  g.inited = false;
  return BSPGEMM_OK;
}
extern "C" int bspgemm_num_gpus(void) { return g.inited ? (int)g.devs.size() : 0; }

static int ensure_init() {
  if (g.inited) return BSPGEMM_OK;
  const char* e = getenv("BSPGEMM_GPUS");
  return bspgemm_init(e ? atoi(e) : 1);
}

// ------------------------------------------------------------------------------------------------ host-pointer operators
// Shards rows [0,An) over `ng` GPUs as contiguous blocks (final/SpGEMM_mpi_omp.c:165-171), replicates B
// (ncclBroadcast from GPU 0), runs all shards concurrently, gathers Ccol/Crow to the host at the
// displacements (replaces :178-223).
struct bspgemm_result {            // a product left sharded on the GPUs (bspgemm_csr_sharded): metadata + the all-gathered copies
  int ng = 0, An = 0, Bm = 0, is64 = 0;
  std::vector<int> r0;             // shard q owns rows [r0[q], r0[q+1])
  std::vector<int64_t> disp;       // ... and columns [disp[q], disp[q+1]) of the global Ccol
  std::vector<int*> full_col;      // per task, after bspgemm_result_allgather
  std::vector<void*> full_row;
  unsigned long long epoch = 0;    // the shards live in the contexts' arenas: stale after the next product
};
static unsigned long long g_epoch = 0;

static int host_multiply(const int* Acol, const int* Arow, int An, const int* Bcol, const int* Brow, int Bn, int Bm,
                         int** Ccol_malloc, int* Ccol_buf, int64_t capacity, void* Crow, int is64, int64_t* nnz_out, int ng_limit,
                         bspgemm_result* keep = nullptr) {
  NvtxRange nvtx_("bspgemm.host_multiply");
  if (!Arow || !Brow || (!Crow && !keep) || An < 0 || Bn < 0 || Bm < 0) return fail(BSPGEMM_ERR_BADARG, "null pointer or negative size");
  ++g_epoch;
  CKS(ensure_init());
  // Every GPU of the communicator takes part (a collective on a subset of its ranks never completes): with fewer rows than
  // GPUs the surplus shards are empty — they join the broadcast of B and skip the product.  ng_limit == 1 (slice form) runs
  // on GPU 0 alone and issues no collective.
  const int ng = std::min<int>((int)g.devs.size(), ng_limit);
  const int64_t a_base = Arow[0], Annz = (int64_t)Arow[An] - a_base;
  const int64_t b_base = Brow[0], Bnnz = (int64_t)Brow[Bn] - b_base;
  if (Annz < 0 || Bnnz < 0 || b_base != 0) return fail(BSPGEMM_ERR_BADARG, "row pointers not monotone / Brow[0] != 0");
  if ((Annz > 0 && !Acol) || (Bnnz > 0 && !Bcol)) return fail(BSPGEMM_ERR_BADARG, "null column array");
  const size_t rp = is64 ? 8 : 4;
  // Row blocks: equal rows per task like the reference (tasksize = An / numtasks, :165), or — BSPGEMM_SPLIT=ip — contiguous blocks
  // of equal WORK: the split points are taken on the prefix sum of the rows' intermediate products (equal-row
  // blocks leave the hub rows of a power-law matrix to one task).  The result is identical either way.
  std::vector<int> r0(ng + 1);
  for (int q = 0; q <= ng; ++q) r0[q] = (int)((int64_t)An * q / ng);
  if (ng > 1 && An > 0) if (const char* sp = getenv("BSPGEMM_SPLIT")) if (!strcmp(sp, "ip")) {
    std::vector<int64_t> pre((size_t)An + 1, 0);
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)16, (int64_t)std::thread::hardware_concurrency(), (int64_t)An >> 14}));
    auto work = [&](int t) {
      const int b = (int)((int64_t)An * t / nt), e = (int)((int64_t)An * (t + 1) / nt);
      for (int i = b; i < e; ++i) {
        int64_t w = 1;                                    // an empty row still costs a row
        for (int64_t p = Arow[i]; p < Arow[i + 1]; ++p) { const int j = Acol[p]; if ((unsigned)j < (unsigned)Bn) w += Brow[j + 1] - Brow[j]; }
        pre[(size_t)i + 1] = w;
      }
    };
    if (nt == 1) work(0); else { std::vector<std::thread> th; for (int t = 0; t < nt; ++t) th.emplace_back(work, t); for (auto& x : th) x.join(); }
    for (int i = 0; i < An; ++i) pre[(size_t)i + 1] += pre[i];
    for (int q = 1; q < ng; ++q) {
      const int64_t target = pre[An] * q / ng;
      r0[q] = (int)(std::lower_bound(pre.begin(), pre.end(), target) - pre.begin());
      if (r0[q] < r0[q - 1]) r0[q] = r0[q - 1];
      if (r0[q] > An) r0[q] = An;
    }
  }

  // upload A shards; B to task 0 and — without a communicator (tasks sharing GPUs) — to the first task of every other device
  const int64_t bchunk = (Bnnz + ng - 1) / ng;           // Bcol is uploaded in ng chunks (one per GPU) when a communicator exists
  std::vector<int> b_owner(ng);                       // task whose copy of B task q reads
  for (int q = 0; q < ng; ++q) {
    b_owner[q] = q;
    if (g.comms.empty()) for (int k = 0; k < q; ++k) if (g.devs[k]->device == g.devs[q]->device) { b_owner[q] = k; break; }
  }
  for (int q = 0; q < ng; ++q) {
    bspgemm_dev* d = g.devs[q];
    CK(cudaSetDevice(d->device));
    d->stream = d->own_stream;
    const int rows = r0[q + 1] - r0[q];
    const int64_t lo = Arow[r0[q]], hi = Arow[r0[q + 1]];
    CKS(d->in_arow.ensure((size_t)rows + 1));
    CKS(d->in_acol.ensure((size_t)std::max<int64_t>(hi - lo, 1)));
    CKS(d->crow_dev.ensure(((size_t)rows + 1) * rp));
    CK(cudaMemcpyAsync(d->in_arow.p, Arow + r0[q], ((size_t)rows + 1) * 4, cudaMemcpyHostToDevice, d->stream));
    if (hi > lo) CK(cudaMemcpyAsync(d->in_acol.p, Acol + lo, (size_t)(hi - lo) * 4, cudaMemcpyHostToDevice, d->stream));
    if (b_owner[q] != q) continue;
    CKS(d->in_brow.ensure((size_t)Bn + 1));
    CKS(d->in_bcol.ensure((size_t)std::max<int64_t>(bchunk * ng, 1)));
    if (g.comms.empty() || ng == 1) {
      CK(cudaMemcpyAsync(d->in_brow.p, Brow, ((size_t)Bn + 1) * 4, cudaMemcpyHostToDevice, d->stream));
      if (Bnnz > 0) CK(cudaMemcpyAsync(d->in_bcol.p, Bcol, (size_t)Bnnz * 4, cudaMemcpyHostToDevice, d->stream));
    } else {
      // B crosses PCIe ONCE in total: GPU q uploads the q-th of ng equal chunks of Bcol over its own link (GPU 0 also Brow),
      // NVLink replicates (ncclAllGather in place / ncclBroadcast) — the reference: every rank parses the whole file (:309)
      if (q == 0) CK(cudaMemcpyAsync(d->in_brow.p, Brow, ((size_t)Bn + 1) * 4, cudaMemcpyHostToDevice, d->stream));
      const int64_t c0 = bchunk * q, c1 = std::min<int64_t>(Bnnz, c0 + bchunk);
      if (c1 > c0) CK(cudaMemcpyAsync(d->in_bcol.p + c0, Bcol + c0, (size_t)(c1 - c0) * 4, cudaMemcpyHostToDevice, d->stream));
    }
  }
  if (ng > 1 && !g.comms.empty()) {
    NK(g.nccl.GroupStart());
    for (int q = 0; q < ng; ++q) NK(g.nccl.Broadcast(g.devs[0]->in_brow.p, g.devs[q]->in_brow.p, (size_t)Bn + 1, ncclInt32, 0, g.comms[q], g.devs[q]->stream));
    NK(g.nccl.GroupEnd());
    if (Bnnz > 0) {
      NK(g.nccl.GroupStart());
      for (int q = 0; q < ng; ++q) NK(g.nccl.AllGather(g.devs[q]->in_bcol.p + bchunk * q, g.devs[q]->in_bcol.p, (size_t)bchunk, ncclInt32, g.comms[q], g.devs[q]->stream));
      NK(g.nccl.GroupEnd());
    }
  } else if (ng > 1) {                // shards sharing a GPU read the owner's copy once its upload has landed
    for (int q = 0; q < ng; ++q) if (b_owner[q] == q) { CK(cudaSetDevice(g.devs[q]->device)); CK(cudaStreamSynchronize(g.devs[q]->stream)); }
  }
  // run the shards concurrently: launch phase k on every GPU, then wait
  for (int q = 0; q < ng; ++q) {
    bspgemm_dev* d = g.devs[q];
    const int rows = r0[q + 1] - r0[q];
    const int64_t lo = Arow[r0[q]], hi = Arow[r0[q + 1]];
    // Arow holds absolute offsets; shift the base pointer so that Acol_dev[Arow[i]] addresses the shard copy
    const int* acol_dev = (const int*)((uintptr_t)d->in_acol.p - (uintptr_t)lo * 4u);
    d->a.m = Csr{d->in_arow.p, acol_dev, g.devs[b_owner[q]]->in_brow.p, g.devs[b_owner[q]]->in_bcol.p, rows, Bn, Bm};
    d->a.Annz = hi - lo; d->a.Bnnz = Bnnz; d->a.dCrow = d->crow_dev.p; d->a.is64 = is64;
    d->user_ccol = nullptr; d->user_cap = 0;
    d->phase = 0;
  }
  auto for_all = [&](int (*fn)(bspgemm_dev*)) -> int {
    for (int q = 0; q < ng; ++q) { bspgemm_dev* d = g.devs[q]; if (d->a.m.An == 0) continue; CKS(fn(d)); }
    return BSPGEMM_OK;
  };
  CKS(for_all(mul_launch_probe));
  CKS(for_all(mul_launch_estimate));
  CKS(for_all(mul_launch_main));
  CKS(for_all(mul_launch_fill));
  CKS(for_all(mul_finish));

  // displacements (host exclusive scan, replaces :189-196), allocation (:200), gather (:203-204) with the
  // row-pointer offset applied on the device (replaces :211-223)
  std::vector<int64_t> disp(ng + 1, 0);
  for (int q = 0; q < ng; ++q) disp[q + 1] = disp[q] + (g.devs[q]->a.m.An ? g.devs[q]->st.nnz : 0);
  const int64_t nnz = disp[ng];
  if (nnz_out) *nnz_out = nnz;
  if (!is64 && nnz > 0x7fffffffLL) return fail(BSPGEMM_ERR_OVERFLOW32, "nnz(C) = %lld does not fit 32-bit row pointers", (long long)nnz);
  if (keep) {                      // distributed consumer: nothing is gathered (replaces the gather-to-root :203-223)
    keep->ng = ng; keep->An = An; keep->Bm = Bm; keep->is64 = is64; keep->r0 = r0; keep->disp = disp; keep->epoch = g_epoch;
    return BSPGEMM_OK;
  }
  int* out = Ccol_buf;
  if (Ccol_malloc) {
    out = (int*)malloc((size_t)std::max<int64_t>(nnz, 1) * sizeof(int));
    if (!out) return fail(BSPGEMM_ERR_OOM, "malloc of %lld ints failed", (long long)nnz);
  } else if (capacity < nnz) return fail(BSPGEMM_ERR_CAPACITY, "output capacity %lld < nnz(C) %lld", (long long)capacity, (long long)nnz);
  if (is64) ((int64_t*)Crow)[0] = 0; else ((int*)Crow)[0] = 0;
  for (int q = 0; q < ng; ++q) {
    bspgemm_dev* d = g.devs[q];
    const int rows = r0[q + 1] - r0[q];
    if (rows == 0) continue;
    CK(cudaSetDevice(d->device));
    const char* src = (const char*)d->crow_dev.p;
    if (disp[q] != 0) {
      CKS(d->crow_tmp.ensure(((size_t)rows + 1) * rp));
      k_offset_rowptr<<<(rows + 1 + 255) / 256, 256, 0, d->stream>>>(d->crow_dev.p, d->crow_tmp.p, is64, (long long)rows + 1, (long long)disp[q]);
      CK(cudaGetLastError());
      src = (const char*)d->crow_tmp.p;
    }
    CK(cudaMemcpyAsync((char*)Crow + ((size_t)r0[q] + 1) * rp, src + rp, (size_t)rows * rp, cudaMemcpyDeviceToHost, d->stream));
    if (d->st.nnz > 0) CK(cudaMemcpyAsync(out + disp[q], d->ccol.p, (size_t)d->st.nnz * 4, cudaMemcpyDeviceToHost, d->stream));
  }
  for (int q = 0; q < ng; ++q) { CK(cudaSetDevice(g.devs[q]->device)); CK(cudaStreamSynchronize(g.devs[q]->stream)); }
  if (Ccol_malloc) *Ccol_malloc = out;
  return BSPGEMM_OK;
}

// ------------------------------------------------------------------------------------------------ distributed consumer
// The reference gathers every rank's slice to rank 0 (MPI_Gatherv + MPI_Gather + a serial fix-up, final/SpGEMM_mpi_omp.c:203-223)
// — the step its report blames for the multi-node slow-down.  Here the product can stay where it was computed: the shards
// are handed out as device pointers, all-gathered GPU-to-GPU over NVLink (every GPU ends with the whole CSR, row pointers
// already offset), or written one file per shard.
extern "C" int bspgemm_csr_sharded(const int* Acol, const int* Arow, int An, const int* Bcol, const int* Brow, int Bn, int Bm,
                                   int crow_is_i64, bspgemm_result** out) {
  if (!out) return fail(BSPGEMM_ERR_BADARG, "null result handle");
  bspgemm_result* r = new bspgemm_result();
  int64_t nnz = 0;
  const int st = host_multiply(Acol, Arow, An, Bcol, Brow, Bn, Bm, nullptr, nullptr, 0, nullptr, crow_is_i64 ? 1 : 0, &nnz, 1 << 30, r);
  if (st != BSPGEMM_OK) { delete r; return st; }
  *out = r;
  return BSPGEMM_OK;
}
static int result_live(const bspgemm_result* r) {
  if (!r) return fail(BSPGEMM_ERR_BADARG, "null result");
  if (!g.inited || r->epoch != g_epoch) return fail(BSPGEMM_ERR_STATE, "the sharded result was overwritten by a later product (or the context was finalized)");
  return BSPGEMM_OK;
}
extern "C" int bspgemm_result_shards(const bspgemm_result* r) { return r ? r->ng : 0; }
extern "C" int64_t bspgemm_result_nnz(const bspgemm_result* r) { return r ? r->disp[r->ng] : 0; }
extern "C" int bspgemm_result_shard(const bspgemm_result* r, int q, int* device, int* row0, int* rows, int64_t* nnz, int64_t* disp,
                                    const int** dCcol, const void** dCrow) {
  CKS(result_live(r));
  if (q < 0 || q >= r->ng) return fail(BSPGEMM_ERR_BADARG, "shard %d out of range", q);
  const bspgemm_dev* d = g.devs[q];
  if (device) *device = d->device;
  if (row0) *row0 = r->r0[q];
  if (rows) *rows = r->r0[q + 1] - r->r0[q];
  if (nnz) *nnz = r->disp[q + 1] - r->disp[q];
  if (disp) *disp = r->disp[q];
  if (dCcol) *dCcol = d->ccol.p;
  if (dCrow) *dCrow = d->crow_dev.p;       // slice-relative, rows+1 entries (the reference's Crow_slice, :160-174)
  return BSPGEMM_OK;
}
// Every task's GPU gets the whole CSR: Ccol (nnz ints) and Crow (An+1 row pointers, shard offsets applied on the owning GPU by
// k_offset_rowptr before they travel).  Over a communicator: one grouped ncclBroadcast per shard ("all-gather-v"); tasks that
// share GPUs: device-to-device copies.  The buffers belong to the result handle.
extern "C" int bspgemm_result_allgather(bspgemm_result* r, int** dCcol_per_task, void** dCrow_per_task) {
  CKS(result_live(r));
  const int ng = r->ng; const size_t rp = r->is64 ? 8 : 4;
  const int64_t nnz = r->disp[ng];
  if (r->full_col.empty()) {
    r->full_col.assign(ng, nullptr); r->full_row.assign(ng, nullptr);
    for (int q = 0; q < ng; ++q) {
      CK(cudaSetDevice(g.devs[q]->device));
      CK(cudaMalloc((void**)&r->full_col[q], (size_t)std::max<int64_t>(nnz, 1) * 4));
      CK(cudaMalloc(&r->full_row[q], ((size_t)r->An + 1) * rp));
    }
  }
  for (int q = 0; q < ng; ++q) {             // the owner writes its offset row pointers into its own full array
    bspgemm_dev* d = g.devs[q];
    CK(cudaSetDevice(d->device));
    CK(cudaMemsetAsync(r->full_row[q], 0, rp, d->stream));
    const int rows = r->r0[q + 1] - r->r0[q];
    if (rows > 0) {
      k_offset_rowptr<<<(rows + 255) / 256, 256, 0, d->stream>>>((const char*)d->crow_dev.p + rp, (char*)r->full_row[q] + ((size_t)r->r0[q] + 1) * rp,
                                                                  r->is64, (long long)rows, (long long)r->disp[q]);
      CK(cudaGetLastError());
    }
  }
  if (!g.comms.empty()) {
    for (int s = 0; s < ng; ++s) {
      const int rows = r->r0[s + 1] - r->r0[s]; const int64_t n_s = r->disp[s + 1] - r->disp[s];
      if (rows == 0) continue;
      NK(g.nccl.GroupStart());
      for (int q = 0; q < ng; ++q) {
        char* seg = (char*)r->full_row[q] + ((size_t)r->r0[s] + 1) * rp;
        NK(g.nccl.Broadcast((char*)r->full_row[s] + ((size_t)r->r0[s] + 1) * rp, seg, (size_t)rows, r->is64 ? ncclInt64 : ncclInt32, s, g.comms[q], g.devs[q]->stream));
        if (n_s > 0) NK(g.nccl.Broadcast(g.devs[s]->ccol.p, r->full_col[q] + r->disp[s], (size_t)n_s, ncclInt32, s, g.comms[q], g.devs[q]->stream));
      }
      NK(g.nccl.GroupEnd());
    }
  } else {
    for (int s = 0; s < ng; ++s) { CK(cudaSetDevice(g.devs[s]->device)); CK(cudaStreamSynchronize(g.devs[s]->stream)); }
    for (int s = 0; s < ng; ++s) {
      const int rows = r->r0[s + 1] - r->r0[s]; const int64_t n_s = r->disp[s + 1] - r->disp[s];
      if (rows == 0) continue;
      for (int q = 0; q < ng; ++q) {
        CK(cudaSetDevice(g.devs[q]->device));
        if (q != s) CK(cudaMemcpyPeerAsync((char*)r->full_row[q] + ((size_t)r->r0[s] + 1) * rp, g.devs[q]->device,
                                           (char*)r->full_row[s] + ((size_t)r->r0[s] + 1) * rp, g.devs[s]->device, (size_t)rows * rp, g.devs[q]->stream));
        if (n_s > 0) CK(cudaMemcpyPeerAsync(r->full_col[q] + r->disp[s], g.devs[q]->device, g.devs[s]->ccol.p, g.devs[s]->device, (size_t)n_s * 4, g.devs[q]->stream));
      }
    }
  }
  for (int q = 0; q < ng; ++q) { CK(cudaSetDevice(g.devs[q]->device)); CK(cudaStreamSynchronize(g.devs[q]->stream)); }
  for (int q = 0; q < ng; ++q) { if (dCcol_per_task) dCcol_per_task[q] = r->full_col[q]; if (dCrow_per_task) dCrow_per_task[q] = r->full_row[q]; }
  return BSPGEMM_OK;
}
// One file per shard, written from the shard's own GPU: <prefix>.shard<q>.bin (format 0: 8 x int64 header {magic, An, Bm, row0,
// rows, nnz, disp, 0}, then rows+1 slice-relative int64 row pointers, then nnz int32 columns) or <prefix>.shard<q>.mtx
// (format 1: Matrix Market coordinate pattern, global dimensions, this shard's entries in the reference's transposed
// convention — the text line of in-memory (row r, col c) is "c+1 r+1", so that readCOO gives the rows back, final/utils.c:66-77).
extern "C" int bspgemm_result_write(const bspgemm_result* r, const char* prefix, int format) {
  CKS(result_live(r));
  if (!prefix || (format != 0 && format != 1)) return fail(BSPGEMM_ERR_BADARG, "bad prefix / format");
  const size_t rp = r->is64 ? 8 : 4;
  for (int q = 0; q < r->ng; ++q) {
    bspgemm_dev* d = g.devs[q];
    const int rows = r->r0[q + 1] - r->r0[q]; const int64_t n_q = r->disp[q + 1] - r->disp[q];
    std::vector<char> rowbuf(((size_t)rows + 1) * rp); std::vector<int> col((size_t)std::max<int64_t>(n_q, 1));
    CK(cudaSetDevice(d->device));
    if (rows > 0) CK(cudaMemcpy(rowbuf.data(), d->crow_dev.p, ((size_t)rows + 1) * rp, cudaMemcpyDeviceToHost));
    else memset(rowbuf.data(), 0, rowbuf.size());
    if (n_q > 0) CK(cudaMemcpy(col.data(), d->ccol.p, (size_t)n_q * 4, cudaMemcpyDeviceToHost));
    std::vector<int64_t> row64((size_t)rows + 1);
    for (int i = 0; i <= rows; ++i) row64[i] = r->is64 ? ((const int64_t*)rowbuf.data())[i] : (int64_t)((const int*)rowbuf.data())[i];
    char path[4096];
    snprintf(path, sizeof path, "%s.shard%d.%s", prefix, q, format ? "mtx" : "bin");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(BSPGEMM_ERR_BADARG, "cannot open %s", path);
    bool ok = true;
    if (format == 0) {
      const int64_t hdr[8] = {0x3152534347505342LL /* "BSPGCSR1" */, r->An, r->Bm, r->r0[q], rows, n_q, r->disp[q], 0};
      ok = fwrite(hdr, sizeof hdr, 1, f) == 1 && fwrite(row64.data(), 8, (size_t)rows + 1, f) == (size_t)rows + 1 &&
           (n_q == 0 || fwrite(col.data(), 4, (size_t)n_q, f) == (size_t)n_q);
    } else {
      std::vector<char> buf(1 << 20);
      setvbuf(f, buf.data(), _IOFBF, buf.size());
      fprintf(f, "%%%%MatrixMarket matrix coordinate pattern general\n%% rows [%d,%d) of C, shard %d of %d\n%d %d %lld\n", r->r0[q], r->r0[q + 1], q, r->ng,
              r->Bm, r->An, (long long)n_q);
      for (int i = 0; i < rows; ++i)
        for (int64_t p = row64[i]; p < row64[i + 1]; ++p) fprintf(f, "%d %d\n", col[p] + 1, r->r0[q] + i + 1);
      ok = !ferror(f);
      fflush(f); setvbuf(f, nullptr, _IONBF, 0);
    }
    if (fclose(f) != 0 || !ok) return fail(BSPGEMM_ERR_BADARG, "write to %s failed", path);
  }
  return BSPGEMM_OK;
}
extern "C" int bspgemm_result_free(bspgemm_result* r) {
  if (!r) return BSPGEMM_OK;
  if (g.inited && r->epoch != 0)
    for (size_t q = 0; q < r->full_col.size(); ++q) {
      if (q < g.devs.size()) cudaSetDevice(g.devs[q]->device);
      if (r->full_col[q]) cudaFree(r->full_col[q]);
      if (r->full_row[q]) cudaFree(r->full_row[q]);
    }
  delete r;
  return BSPGEMM_OK;
}

extern "C" int bspgemm_csr(const int* Acol, const int* Arow, int An, const int* Bcol, const int* Brow, int Bn, int Bm, int** Ccol, int* Crow) {
  if (!Ccol) return fail(BSPGEMM_ERR_BADARG, "null Ccol");
  int64_t nnz = 0;
  return host_multiply(Acol, Arow, An, Bcol, Brow, Bn, Bm, Ccol, nullptr, 0, Crow, 0, &nnz, 1 << 30);
}
extern "C" int bspgemm_csr_i64(const int* Acol, const int* Arow, int An, const int* Bcol, const int* Brow, int Bn, int Bm, int** Ccol, int64_t* Crow) {
  if (!Ccol) return fail(BSPGEMM_ERR_BADARG, "null Ccol");
  int64_t nnz = 0;
  return host_multiply(Acol, Arow, An, Bcol, Brow, Bn, Bm, Ccol, nullptr, 0, Crow, 1, &nnz, 1 << 30);
}
extern "C" int bspgemm_csr_into(const int* Acol, const int* Arow, int An, const int* Bcol, const int* Brow, int Bn, int Bm,
                                int* Ccol_buf, int64_t capacity, int* Crow, int64_t* nnz_out) {
  if (!Ccol_buf && capacity > 0) return fail(BSPGEMM_ERR_BADARG, "null Ccol_buf");
  return host_multiply(Acol, Arow, An, Bcol, Brow, Bn, Bm, nullptr, Ccol_buf, capacity, Crow, 0, nnz_out, 1 << 30);
}
extern "C" int bspgemm_csr_slice(const int* Acol, const int* Arow, int An, const int* Bcol, const int* Brow, int Bn, int Bm,
                                 int** Ccol, int* Crow, int start_row, int end_row) {
  if (!Ccol || start_row < 0 || end_row < start_row || end_row > An) return fail(BSPGEMM_ERR_BADARG, "bad slice [%d,%d) of %d rows", start_row, end_row, An);
  int64_t nnz = 0;
  return host_multiply(Acol, Arow + start_row, end_row - start_row, Bcol, Brow, Bn, Bm, Ccol, nullptr, 0, Crow, 0, &nnz, 1);
}

// Host-pointer masked product on GPU 0 (the reference's SpGEMM_masked is serial, :227-228): upload A, B, F; download C.
extern "C" int bspgemm_csr_masked(const int* Acol, const int* Arow, int An, const int* Bcol, const int* Brow, int Bn, int Bm,
                                  const int* Fcol, const int* Frow, int** Ccol, int* Crow) {
  if (!Arow || !Brow || !Frow || !Crow || !Ccol || An < 0 || Bn < 0 || Bm < 0) return fail(BSPGEMM_ERR_BADARG, "null pointer or negative size");
  CKS(ensure_init());
  bspgemm_dev* d = g.devs[0];
  CK(cudaSetDevice(d->device));
  d->stream = d->own_stream;
  const int64_t alo = Arow[0], Annz = (int64_t)Arow[An] - alo, Bnnz = (int64_t)Brow[Bn] - Brow[0], flo = Frow[0], Fnnz = (int64_t)Frow[An] - flo;
  if (Annz < 0 || Bnnz < 0 || Fnnz < 0 || Brow[0] != 0) return fail(BSPGEMM_ERR_BADARG, "row pointers not monotone / Brow[0] != 0");
  DevBuf<int> frow, fcol;
  struct Guard { DevBuf<int>& a; DevBuf<int>& b; ~Guard() { a.release(); b.release(); } } guard{frow, fcol};
  CKS(d->in_arow.ensure((size_t)An + 1)); CKS(d->in_acol.ensure((size_t)std::max<int64_t>(Annz, 1)));
  CKS(d->in_brow.ensure((size_t)Bn + 1)); CKS(d->in_bcol.ensure((size_t)std::max<int64_t>(Bnnz, 1)));
  CKS(frow.ensure((size_t)An + 1)); CKS(fcol.ensure((size_t)std::max<int64_t>(Fnnz, 1)));
  CKS(d->crow_dev.ensure(((size_t)An + 1) * 4));
  CK(cudaMemcpyAsync(d->in_arow.p, Arow, ((size_t)An + 1) * 4, cudaMemcpyHostToDevice, d->stream));
  if (Annz) CK(cudaMemcpyAsync(d->in_acol.p, Acol + alo, (size_t)Annz * 4, cudaMemcpyHostToDevice, d->stream));
  CK(cudaMemcpyAsync(d->in_brow.p, Brow, ((size_t)Bn + 1) * 4, cudaMemcpyHostToDevice, d->stream));
  if (Bnnz) CK(cudaMemcpyAsync(d->in_bcol.p, Bcol, (size_t)Bnnz * 4, cudaMemcpyHostToDevice, d->stream));
  CK(cudaMemcpyAsync(frow.p, Frow, ((size_t)An + 1) * 4, cudaMemcpyHostToDevice, d->stream));
  if (Fnnz) CK(cudaMemcpyAsync(fcol.p, Fcol + flo, (size_t)Fnnz * 4, cudaMemcpyHostToDevice, d->stream));
  const int* acol_dev = (const int*)((uintptr_t)d->in_acol.p - (uintptr_t)alo * 4u);
  const int* fcol_dev = (const int*)((uintptr_t)fcol.p - (uintptr_t)flo * 4u);
  d->a.m = Csr{d->in_arow.p, acol_dev, d->in_brow.p, d->in_bcol.p, An, Bn, Bm};
  d->a.Annz = Annz; d->a.Bnnz = Bnnz; d->a.dCrow = d->crow_dev.p; d->a.is64 = 0;
  d->user_ccol = nullptr; d->user_cap = 0;
  int* dC = nullptr; int64_t nnz = 0;
  CKS(masked_multiply(d, fcol_dev, frow.p, Fnnz, d->crow_dev.p, 0, &dC, &nnz));
  int* out = (int*)malloc((size_t)std::max<int64_t>(nnz, 1) * sizeof(int));
  if (!out) return fail(BSPGEMM_ERR_OOM, "malloc of %lld ints failed", (long long)nnz);
  CK(cudaMemcpyAsync(Crow, d->crow_dev.p, ((size_t)An + 1) * 4, cudaMemcpyDeviceToHost, d->stream));
  if (nnz) CK(cudaMemcpyAsync(out, dC, (size_t)nnz * 4, cudaMemcpyDeviceToHost, d->stream));
  CK(cudaStreamSynchronize(d->stream));
  *Ccol = out;
  return BSPGEMM_OK;
}

extern "C" int bspgemm_intermediate_products(const int* Acol, const int* Arow, int An, const int* Brow, int Bn, int64_t* ip_out) {
  if (!Arow || !Brow || !ip_out || An < 0 || Bn < 0) return fail(BSPGEMM_ERR_BADARG, "null pointer or negative size");
  CKS(ensure_init());
  bspgemm_dev* d = g.devs[0];
  CK(cudaSetDevice(d->device));
  d->stream = d->own_stream;
  const int64_t lo = Arow[0], hi = Arow[An];
  *ip_out = 0;
  if (An == 0) return BSPGEMM_OK;
  CKS(d->in_arow.ensure((size_t)An + 1)); CKS(d->in_acol.ensure((size_t)std::max<int64_t>(hi - lo, 1))); CKS(d->in_brow.ensure((size_t)Bn + 1));
  CK(cudaMemcpyAsync(d->in_arow.p, Arow, ((size_t)An + 1) * 4, cudaMemcpyHostToDevice, d->stream));
  if (hi > lo) CK(cudaMemcpyAsync(d->in_acol.p, Acol + lo, (size_t)(hi - lo) * 4, cudaMemcpyHostToDevice, d->stream));
  CK(cudaMemcpyAsync(d->in_brow.p, Brow, ((size_t)Bn + 1) * 4, cudaMemcpyHostToDevice, d->stream));
  const int* acol_dev = (const int*)((uintptr_t)d->in_acol.p - (uintptr_t)lo * 4u);
  d->a.m = Csr{d->in_arow.p, acol_dev, d->in_brow.p, nullptr, An, Bn, 0};
  d->a.Annz = hi - lo; d->a.Bnnz = 0; d->a.dCrow = nullptr; d->a.is64 = 0;
  CK(cudaMemsetAsync(d->d_sc, 0, sizeof(DevScalars), d->stream));
  d->launches = 0;
  CKS(launch_estimate_kernel(d));
  CK(cudaMemcpyAsync(d->h_sc, d->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, d->stream));
  CK(cudaStreamSynchronize(d->stream));
  d->phase = 0;
  if (d->h_sc->err & 1u) return fail(BSPGEMM_ERR_BADARG, "a column index of A is outside [0,Bn=%d)", Bn);
  *ip_out = (int64_t)d->h_sc->total_ip;
  return BSPGEMM_OK;
}

// ---- legacy-signature drop-ins (void, exit(1) on failure like final/utils.c:54-61) ----
// The reference's signature has no "rows of B": Bn = max(Acol) + 1 (any valid CSR pair has at least that many B rows).  One
// pass over Acol on up to 16 host threads (config 3: 6.7e7 entries).  The explicit-Bn entry points avoid it.
static int derive_bn(const int* Acol, const int* Arow, int An) {
  const int64_t lo = Arow[0], hi = Arow[An], n = hi - lo;
  if (n <= 0) return 0;
  const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)16, (int64_t)std::thread::hardware_concurrency(), n >> 20}));
  std::vector<int> part(nt, -1);
  auto scan = [&](int t) {
    const int64_t b = lo + n * t / nt, e = lo + n * (t + 1) / nt;
    int mx = -1;
    for (int64_t p = b; p < e; ++p) mx = std::max(mx, Acol[p]);
    part[t] = mx;
  };
  if (nt == 1) scan(0);
  else { std::vector<std::thread> th; for (int t = 0; t < nt; ++t) th.emplace_back(scan, t); for (auto& x : th) x.join(); }
  return *std::max_element(part.begin(), part.end()) + 1;
}
static void die_on(int s, const char* who) {
  if (s == BSPGEMM_OK) return;
  fprintf(stderr, "%s: %s: %s\n", who, bspgemm_strerror(s), bspgemm_last_error());
  exit(1);
}
extern "C" void bspgemm_SpGEMM_mpi(int* Acol, int* Arow, int An, int* Bcol, int* Brow, int Bm, int** Ccol, int* Crow, int tBlock) {
  (void)tBlock;
  die_on(bspgemm_csr(Acol, Arow, An, Bcol, Brow, derive_bn(Acol, Arow, An), Bm, Ccol, Crow), "SpGEMM_mpi");
}
extern "C" void bspgemm_SpGEMM_masked(int* Acol, int* Arow, int An, int* Bcol, int* Brow, int Bm, int* Fcol, int* Frow, int** Ccol, int* Crow, int* Csize) {
  // same argument list as the reference (final/SpGEMM_mpi_omp.c:232-235); the caller's growable *Ccol is replaced by an exact-size one
  int* fresh = nullptr;
  die_on(bspgemm_csr_masked(Acol, Arow, An, Bcol, Brow, derive_bn(Acol, Arow, An), Bm, Fcol, Frow, &fresh, Crow), "SpGEMM_masked");
  if (Ccol) { free(*Ccol); *Ccol = fresh; } else free(fresh);
  if (Csize) *Csize = Crow[An];
}
extern "C" void bspgemm_SpGEMM_omp(int* Acol, int* Arow, int An, int* Bcol, int* Brow, int Bm, int** Ccol, int* Crow, int tBlock) {
  (void)tBlock;
  die_on(bspgemm_csr_slice(Acol, Arow, An, Bcol, Brow, derive_bn(Acol, Arow, An), Bm, Ccol, Crow, 0, An), "SpGEMM_omp");
}
extern "C" void bspgemm_SpGEMM_bigslice(int* Acol, int* Arow, int An, int* Bcol, int* Brow, int Bm, int** Ccol, int* Crow, int* Csize,
                                        int start_row, int end_row) {
  // The reference appends into a caller-owned growable *Ccol (:28-31); here the old buffer is released
  // and replaced by an exact-size one, and *Csize is updated to its capacity.
  int* fresh = nullptr;
  die_on(bspgemm_csr_slice(Acol, Arow, An, Bcol, Brow, derive_bn(Acol, Arow + start_row, end_row - start_row), Bm, &fresh, Crow, start_row, end_row), "SpGEMM_bigslice");
  if (Ccol) { free(*Ccol); *Ccol = fresh; } else free(fresh);
  if (Csize) *Csize = Crow[end_row - start_row];
}

// ------------------------------------------------------------------------------------------------ COO -> CSC on the device
// Replaces final/coo2csc.c:22-64 (see coo2csc.cuh).  Works on the current device, needs no bspgemm_init.
namespace {
struct C2cTemps {
  u32 *keys[2] = {nullptr, nullptr}, *vals[2] = {nullptr, nullptr}, *hist = nullptr, *cnt = nullptr;
  int* pos = nullptr; u64* status = nullptr; DevScalars* sc = nullptr;
  ~C2cTemps() { for (void* q : {(void*)keys[0], (void*)keys[1], (void*)vals[0], (void*)vals[1], (void*)hist, (void*)cnt, (void*)pos, (void*)status, (void*)sc}) if (q) cudaFree(q); }
};
static int c2c_scan(const u32* in, size_t len, int* out, C2cTemps& t, cudaStream_t st) {      // out[0..len]: exclusive prefix, out[len] = total
  const u32 ntiles = (u32)((len + (size_t)SCAN_THREADS * SCAN_ITEMS - 1) / ((size_t)SCAN_THREADS * SCAN_ITEMS));
  CK(cudaMemsetAsync(t.status, 0, (size_t)ntiles * sizeof(u64), st));
  CK(cudaMemsetAsync(t.sc, 0, sizeof(DevScalars), st));
  if (ntiles == 0) { CK(cudaMemsetAsync(out, 0, sizeof(int), st)); return BSPGEMM_OK; }
  k_scan<<<ntiles, SCAN_THREADS, 0, st>>>(in, (int)len, out, 0, t.status, t.sc, ntiles);
  CK(cudaGetLastError());
  return BSPGEMM_OK;
}
}  // namespace

extern "C" int bspgemm_coo2csc_dev(void* stream, uint32_t* d_row, uint32_t* d_col, const uint32_t* d_row_coo, const uint32_t* d_col_coo,
                                   uint32_t nnz, uint32_t n, uint32_t isOneBased) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(BSPGEMM_ERR_NOGPU, "no CUDA device"); }
  if (!d_col || (nnz && (!d_row || !d_row_coo || !d_col_coo))) return fail(BSPGEMM_ERR_BADARG, "coo2csc: null pointer");
  if (nnz > 0x7fffffffu || n > 0x7ffffff0u || isOneBased > 1u) return fail(BSPGEMM_ERR_BADARG, "coo2csc: nnz and n must be below 2^31, isOneBased 0 or 1");
  cudaStream_t st = (cudaStream_t)stream;
  C2cTemps t;
  const u32 nblocks = (u32)(((size_t)nnz + C2C_CHUNK - 1) / C2C_CHUNK);
  const size_t hlen = (size_t)256 * nblocks, slen = std::max<size_t>(hlen, (size_t)n);
  if (hlen > 0x7fffffffull) return fail(BSPGEMM_ERR_BADARG, "coo2csc: too many entries");
  const size_t stiles = (slen + (size_t)SCAN_THREADS * SCAN_ITEMS - 1) / ((size_t)SCAN_THREADS * SCAN_ITEMS) + 1;
  CK(cudaMalloc((void**)&t.cnt, ((size_t)n + 1) * sizeof(u32)));
  CK(cudaMalloc((void**)&t.status, stiles * sizeof(u64)));
  CK(cudaMalloc((void**)&t.sc, sizeof(DevScalars)));
  // pointer array: per-column counts, then their scan straight into col[0..n]
  CK(cudaMemsetAsync(t.cnt, 0, ((size_t)n + 1) * sizeof(u32), st));
  CK(cudaMemsetAsync(t.sc, 0, sizeof(DevScalars), st));
  u32* d_err = &t.sc->err;
  if (nnz) {
    k_c2c_count<<<(int)std::min<size_t>(((size_t)nnz + 255) / 256, 148 * 16), 256, 0, st>>>(d_col_coo, nnz, n, isOneBased, t.cnt, d_err);
    CK(cudaGetLastError());
    u32 h_err = 0;
    CK(cudaMemcpyAsync(&h_err, d_err, sizeof(u32), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_err) return fail(BSPGEMM_ERR_BADARG, "coo2csc: a column index lies outside [0,n) (the reference's behaviour is undefined there)");
  }
  CKS(c2c_scan(t.cnt, n, (int*)d_col, t, st));
  if (nnz == 0) { CK(cudaStreamSynchronize(st)); return BSPGEMM_OK; }
  // stable LSD radix sort of (col_coo, row_coo) by the key's low `bits` bits
  int bits = 1; while (bits < 32 && (1ull << bits) < (unsigned long long)n) ++bits;
  const int passes = (bits + 7) / 8;
  CK(cudaMalloc((void**)&t.hist, hlen * sizeof(u32)));
  CK(cudaMalloc((void**)&t.pos, (hlen + 1) * sizeof(int)));
  if (passes > 1) for (int i = 0; i < 2; ++i) {
    CK(cudaMalloc((void**)&t.keys[i], (size_t)nnz * sizeof(u32)));
    CK(cudaMalloc((void**)&t.vals[i], (size_t)nnz * sizeof(u32)));
  }
  const u32 *kin = d_col_coo, *vin = d_row_coo;
  u32 base = isOneBased;
  for (int pass = 0; pass < passes; ++pass) {
    const bool last = pass == passes - 1;
    u32* kout = last ? nullptr : t.keys[pass & 1];
    u32* vout = last ? d_row : t.vals[pass & 1];
    k_c2c_hist<<<nblocks, 32 * C2C_WARPS, 0, st>>>(kin, nnz, base, 8 * pass, t.hist, nblocks);
    CK(cudaGetLastError());
    CKS(c2c_scan(t.hist, hlen, t.pos, t, st));
    k_c2c_scatter<<<nblocks, 32 * C2C_WARPS, 0, st>>>(kin, vin, nnz, base, base, 8 * pass, t.pos, nblocks, kout, vout);
    CK(cudaGetLastError());
    kin = kout; vin = vout; base = 0;
  }
  CK(cudaStreamSynchronize(st));
  return BSPGEMM_OK;
}

extern "C" int bspgemm_coo2csc(uint32_t* row, uint32_t* col, const uint32_t* row_coo, const uint32_t* col_coo,
                               uint32_t nnz, uint32_t n, uint32_t isOneBased) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(BSPGEMM_ERR_NOGPU, "no CUDA device"); }
  if (!col || (nnz && (!row || !row_coo || !col_coo))) return fail(BSPGEMM_ERR_BADARG, "coo2csc: null pointer");
  struct Bufs { u32 *I = nullptr, *J = nullptr, *r = nullptr, *c = nullptr; ~Bufs() { for (u32* q : {I, J, r, c}) if (q) cudaFree(q); } } b;
  const size_t eb = std::max<size_t>(1, nnz) * sizeof(u32);
  CK(cudaMalloc((void**)&b.I, eb)); CK(cudaMalloc((void**)&b.J, eb)); CK(cudaMalloc((void**)&b.r, eb));
  CK(cudaMalloc((void**)&b.c, ((size_t)n + 1) * sizeof(u32)));
  if (nnz) {
    CK(cudaMemcpy(b.I, row_coo, (size_t)nnz * sizeof(u32), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b.J, col_coo, (size_t)nnz * sizeof(u32), cudaMemcpyHostToDevice));
  }
  CKS(bspgemm_coo2csc_dev(nullptr, b.r, b.c, b.I, b.J, nnz, n, isOneBased));
  if (nnz) CK(cudaMemcpy(row, b.r, (size_t)nnz * sizeof(u32), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(col, b.c, ((size_t)n + 1) * sizeof(u32), cudaMemcpyDeviceToHost));
  return BSPGEMM_OK;
}
