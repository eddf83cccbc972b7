// The launch plan of the fused head kernel: which instantiation serves a channel count and how the image is cut into
// warp units.  Pure host arithmetic (no CUDA runtime calls), shared by head.cu and the CPU emulation harness of
// tests/cpu_simt.  Included at the end of head_kernel.cuh.
#pragma once

namespace simt {

struct PlanTuning { int ur, rs, lpr; };   // simt_head_set_tuning overrides; 0 = automatic

inline int choose_config(int CK, int lpr_req, Plan* P) {
  struct Cfg { int cpl, lpr, nt, minb; };
  static const Cfg cfgs[] = {
#define X(cpl, lpr, nt, minb_fwd, minb_bwd) {cpl, lpr, nt, minb_fwd},
      SIMT_HEAD_CONFIGS(X)
#undef X
  };
  const Cfg* best = nullptr;
  for (const Cfg& c : cfgs) {
    if (c.cpl * c.lpr < CK) continue;
    if (lpr_req > 0 && c.lpr != lpr_req) continue;
    // prefer the fewest lanes per cell, then the least channel padding
    if (!best || c.lpr < best->lpr || (c.lpr == best->lpr && c.cpl * c.lpr < best->cpl * best->lpr)) best = &c;
  }
  if (!best && lpr_req > 0) return choose_config(CK, 0, P);
  if (!best) return SIMT_EUNSUPPORTED;
  P->CPL = best->cpl; P->LPR = best->lpr; P->NT = best->nt; P->MINB = best->minb;
  P->CKP = best->cpl * best->lpr;
  return 0;
}

// sm_count <= 0: unknown (no row splitting)
inline int make_plan_for(int mode, int B, int CK, int C, int h, int w, int H, int W, const PlanTuning& tune, int sm_count,
                         HeadArgs* A, Plan* P) {
  int rc = choose_config(CK, tune.lpr, P);
  if (rc) return rc;
  A->sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
  A->sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
  A->ncy = h > 1 ? h - 1 : 1;
  A->ncx = w > 1 ? w - 1 : 1;
  // cell-rows per unit: ~8 pixel rows per unit keeps the per-unit overhead amortised
  int ur = tune.ur;
  if (ur <= 0) {
    const double rows_per_cell = (double)H / (double)A->ncy;
    ur = (int)(8.0 / rows_per_cell + 0.5);
    if (ur < 1) ur = 1;
    if (ur > 32) ur = 32;
  }
  const int cpw = 32 / P->LPR;
  A->ur = ur;
  // split each cell-row over rs units when the grid would otherwise see only a few units per warp
  int rs = tune.rs > 0 ? tune.rs : 0;
  if (rs <= 0 && sm_count > 0) {
    const double warps = (double)sm_count * 3.0 * (P->NT / 32);   // ~3 CTAs per SM resident
    const double base_units = (double)B * ((A->ncy + ur - 1) / ur) * ((A->ncx + cpw - 1) / cpw);
    rs = (int)(4.0 * warps / base_units + 0.5);                   // aim at >= ~4 units per warp
    if (rs > 4) rs = 4;
  }
  if (ur > 1) rs = 1;
  A->rs = rs < 1 ? 1 : rs;
  A->units_y = ((A->ncy + ur - 1) / ur) * A->rs;
  A->units_x = (A->ncx + cpw - 1) / cpw;
  A->nunits = (long long)B * A->units_y * A->units_x;
  const bool bwd = mode != MODE_FWD;
  const size_t nw = (size_t)(P->NT / 32);
  P->smem = nw * 4 * (P->CPL / 2) * 32 * 8 + (size_t)C * P->CKP * 4 +
            (bwd ? nw * kEdgeRows * (P->CKP + 1) * 4 : 0) + (size_t)(A->ncx + A->ncy + 2) * 4;
#ifdef SIMT_EXP_LXTAB
  P->smem += (size_t)(W + H) * 4;
#endif
  return 0;
}

}  // namespace simt
