// binary-spgemm_b200/csrc/ctx.h — host-side context shared by the translation units of libbspgemm.so.
//
// The library is compiled as several translation units (bspgemm.cu: dispatch + C ABI + the CSR-gather kernels;
// tu_ell.cu: ELL re-layout + ordered-table kernel; tu_sort_w*.cu: the sorting-network kernels, one ELL width each) so that
// `make -j` builds them in parallel; every kernel is launched from the unit that instantiates it.
#pragma once
#include "../../include/bspgemm.h"
#include "kernels.cuh"

#define BSP_HIDDEN __attribute__((visibility("hidden")))

#include <nccl.h>      // types only; the library itself is dlopen'ed so libbspgemm.so loads without it
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <stdarg.h>
#include <math.h>
#include <algorithm>
#include <nvtx3/nvToolsExt.h>   // header-only: ranges are no-ops unless a tool (ncu --nvtx, nsys) injects its library

// One NVTX range per phase of a product ("bspgemm.probe", ".estimate", ".main", ".fill", ".fast", ".host_multiply"; no "/" in the names, it is ncu's range separator): lets
// `ncu --nvtx --nvtx-include "bspgemm.main/"` pick the launches of one phase.  RAII, so every early return closes its range.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#include <vector>
#include <mutex>
#include <thread>

using namespace bsk;

// ------------------------------------------------------------------------------------------------ errors
BSP_HIDDEN int fail(int code, const char* fmt, ...);      // records the message for bspgemm_last_error() and returns `code` (bspgemm.cu)
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return fail(e_ == cudaErrorMemoryAllocation ? BSPGEMM_ERR_OOM : BSPGEMM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define CKS(expr) do { int s_ = (expr); if (s_ != BSPGEMM_OK) return s_; } while (0)

// ------------------------------------------------------------------------------------------------ per-GPU context
template <class T> struct DevBuf {
  T* p = nullptr; size_t cap = 0;
  int ensure(size_t n, bool keep = false) {           // grow-only; contents dropped unless keep
    if (n <= cap) return BSPGEMM_OK;
    size_t want = n + n / 16 + 64;
    T* q = nullptr;
    cudaError_t e = cudaMalloc((void**)&q, want * sizeof(T));
    if (e != cudaSuccess) { want = n; e = cudaMalloc((void**)&q, want * sizeof(T)); }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(BSPGEMM_ERR_OOM, "cudaMalloc of %zu bytes failed: %s", want * sizeof(T), cudaGetErrorString(e)); }
    if (keep && p && cap) cudaMemcpy(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice);
    if (p) cudaFree(p);
    p = q; cap = want;
    return BSPGEMM_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct MulArgs {
  Csr m; int64_t Annz, Bnnz; void* dCrow; int is64;
};

struct bspgemm_dev {
  int device = 0, sm_count = 0;
  size_t smem_optin = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  int mode = BSPGEMM_MODE_AUTO;
  // workspace
  DevBuf<u32> ip, cnt, lists, bitmaps, bell;
  DevBuf<u64> status;
  DevBuf<int> ccol;                 // output arena
  DevBuf<int> temp;                 // staging arena of the big rows (MODE_STAGE): Σ IP of the M/L rows
  DevBuf<u64> tofs;                 // per row: offset of its staged columns in temp
  DevScalars* d_sc = nullptr;       // = status.p: the scalars sit in front of the look-back words, so one memset clears both
  DevScalars* h_sc = nullptr;       // pinned
  // B prepared once for many products (bspgemm_dev_prepare_b): its re-layout and the plan of the last product with it
  struct PreparedB {
    bool valid = false;
    const int* brow = nullptr; const int* bcol = nullptr; int Bn = 0, Bm = 0; int64_t Bnnz = 0;   // identity of B (the caller keeps its contents unchanged)
    u32 max_len_b = 0;
    int ell_W = 0;                  // ELL copy of B in `bell` (every row sorted, padded with d->ell_pad), 0: none
    bool desc = false;              // every B row is a run of consecutive columns: descriptors in `bdesc`
    int variant = -1, sort_LAL = 0; // plan of the last successful product with this B that can be replayed without probes (2 sort, 3 band)
  } pb;
  DevBuf<u32> bdesc;
  // masked product (mask.cuh): 64-bit row pointers of the unmasked product, the canonical mask, the masked output
  DevBuf<long long> m_crow; DevBuf<int> m_frow, m_fcol, m_out;
  u32 ell_pad = EMPTY;              // padding value of the ELL copy in `bell`: EMPTY, or EMPTY_F for the floating-point sort network
  bool fast = false;                // the product in flight was launched from the cached plan (no probes, no host round trip before the launch)
  cudaEvent_t ev[8] = {};
  // per-call state
  MulArgs a{};
  int phase = 0;                    // 0 idle, 1 estimate in flight, 2 main in flight, 3 fill in flight, 4 done
  int used_mode = 0, G = 16, G_big = 16, launches = 0;   // G: lanes per B row from mean len(B); G_big: from the mean length of the SELECTED B rows (Σip / nnzA)
  u32 cap_s = 0, cap_m1 = 0, cap_m2 = 0;
  bool have_m = false, have_m2 = false, have_l = false;
  bool use_band = false, no_band = false;   // run/bitmap kernel for banded matrices (band.cuh); no_band: it failed on this input, redo generally
  bool staged = false;              // big rows went through the staging arena (one pass) in the last multiply
  bool use_bm = false;              // M2/L bins: windowed shared-memory bitmap + summary, load-balanced walk (rows_bm.cuh)
  bool use_window = false;          // M/L bins: windowed shared-memory bitmap (rows_window.cuh) instead of table / global bitmap
  u32 bm_words = 0; int l_grid = 0;
  bool skip_estimate = false; u32 row_ip_bound = 0, max_len_b = 0;
  bool use_ell = false; int ell_W = 0, ell_R = 0, ell_warps = 0; u32 ell_TW = 0, ell_maxA = 0, ell_lf16 = 28;
  bool use_sort = false; int sort_LAL = 0;            // register-sort variant of the ELL path (fused_sort.cuh)   // ELL fast path plan (fused_ell.cuh)
  u32 hist_rows[34] = {};
  int fused_bps[4][16] = {};        // cached occupancy of k_fused<G> per log2(cap)
  int* user_ccol = nullptr; int64_t user_cap = 0;   // caller-provided output (device) or null -> arena
  bspgemm_stats st{};
  // input staging for the host-pointer API
  DevBuf<int> in_arow, in_acol, in_brow, in_bcol;
  DevBuf<char> crow_dev;            // device row pointers for the host API
  DevBuf<char> crow_tmp;
};


// DevScalars live in the first SC_WORDS 64-bit words of `status`, the look-back chain words follow.
constexpr size_t SC_WORDS = (sizeof(DevScalars) + 127) / 128 * 16;
// Room for n chain words, zeroed.  A product replayed from the cached plan (d->fast) has not touched the scalars yet: they are
// cleared by the same memset.  Growing the buffer in the middle of a product keeps the scalars (rare: it is sized per product).
inline int chain_reserve(bspgemm_dev* d, size_t n, u64** chain) {
  if (SC_WORDS + n > d->status.cap) { CK(cudaStreamSynchronize(d->stream)); CKS(d->status.ensure(SC_WORDS + n, true)); }
  d->d_sc = reinterpret_cast<DevScalars*>(d->status.p);
  *chain = d->status.p + SC_WORDS;
  if (d->fast) CK(cudaMemsetAsync(d->status.p, 0, (SC_WORDS + n) * sizeof(u64), d->stream));
  else         CK(cudaMemsetAsync(*chain, 0, n * sizeof(u64), d->stream));
  return BSPGEMM_OK;
}
inline bool b_prepared(const bspgemm_dev* d) {
  const auto& b = d->pb; const Csr& m = d->a.m;
  return b.valid && b.brow == m.Brow && b.bcol == m.Bcol && b.Bn == m.Bn && b.Bm == m.Bm && b.Bnnz == d->a.Bnnz;
}

// ---- launchers defined next to the kernels they instantiate
BSP_HIDDEN int build_ell(bspgemm_dev* d, int W, bool sorted, u32 pad);   // tu_ell.cu: CSR -> ELL copy of the current B into d->bell
BSP_HIDDEN u32 sort_pad_for(int W, int LAL, int Bm);                // tu_ell.cu: padding the sort kernel of this plan expects (sort_plan_flt, fused_sort.cuh)
BSP_HIDDEN int set_attrs_ell(int smem_optin);                      // tu_ell.cu
BSP_HIDDEN int launch_ell(bspgemm_dev* d);                         // tu_ell.cu: k_build_ell, then k_fused_ell or the sort kernel
BSP_HIDDEN int launch_sort(bspgemm_dev* d, int* ccol);             // tu_ell.cu: dispatch on the ELL width ...
#define BSP_DECL_SORT(Wv) BSP_HIDDEN int set_attrs_sort_w##Wv(int smem_optin); BSP_HIDDEN int launch_sort_w##Wv(bspgemm_dev* d, int* ccol);
BSP_DECL_SORT(4) BSP_DECL_SORT(8) BSP_DECL_SORT(16) BSP_DECL_SORT(32)   // ... tu_sort_w*.cu
#undef BSP_DECL_SORT

// Every kernel that uses dynamic shared memory gets the opt-in maximum once, at context creation
// (dynamic + static shared memory must stay within the opt-in limit).
#define BSP_ATTR(k) do { cudaFuncAttributes fa_; CK(cudaFuncGetAttributes(&fa_, k)); \
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - (int)fa_.sharedSizeBytes)); } while (0)
