// Lane-emulation build of the device core for CPU-only tests (never linked into the product
// library): every phase of csrc/okin_core.cuh runs as a loop over 32 lanes.
#define OKIN_LANE_EMU 1
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/okin.h"
#include "okin_core.cuh"

extern "C" int okin_emu_sweep(const int32_t* hdr, const int32_t* ib, const double* fb, long n_instances,
                              int n_steps, const okin_solver_cfg* c, const okin_batch_io* io) {
  if (hdr[OKIN_H_MAGIC] != OKIN_MAGIC) return -1;
  const int32_t* sec[OKIN_S_COUNT];
  for (int s = 0; s < OKIN_S_COUNT; ++s) okin_resolve_section(hdr, ib, ib, s, sec);
  OkinProgram pr{hdr, ib, fb, ib, sec};
  OkinSolverCfg cfg{c->step_tol, c->coarse_tol, c->fine_tol, c->residual_tol, c->mu_init, c->max_iter,
                    c->use_predictor};
  const int nin = hdr[OKIN_H_NIN], nout = hdr[OKIN_H_NOUT], nt = hdr[OKIN_H_NT], n = 3 * hdr[OKIN_H_NF];
  std::vector<double> sm(hdr[OKIN_H_SMEM_DOUBLES]);
  // the library's launch() rule: lean instantiation when no per-state tangent / metric / diagnostic
  // output is wanted (it only owns the lean part of the slice)
  const bool full = io->tangents || io->velocities || io->tangent_health || io->metrics || io->diagnostics;
  const size_t live = full ? (size_t)hdr[OKIN_H_SMEM_DOUBLES] : (size_t)hdr[OKIN_H_SMEM_DOUBLES_LEAN];
  for (long i = 0; i < n_instances; ++i) {
    std::fill(sm.begin(), sm.begin() + live, 0.0);
    std::fill(sm.begin() + live, sm.end(), NAN);     // a lean run must not touch the rest
    OkinOutputs out;
    out.positions = io->positions ? io->positions + (size_t)i * n_steps * 3 * nout : nullptr;
    out.iters = io->iters ? io->iters + (size_t)i * n_steps : nullptr;
    out.max_residual = io->max_residual ? io->max_residual + (size_t)i * n_steps : nullptr;
    out.tangents = io->tangents ? io->tangents + (size_t)i * n_steps * nt * n : nullptr;
    out.velocities = io->velocities ? io->velocities + (size_t)i * n_steps * nt * 3 * nout : nullptr;
    out.health = io->tangent_health ? io->tangent_health + (size_t)i * n_steps * 2 : nullptr;
    out.metrics = io->metrics ? io->metrics + (size_t)i * n_steps * hdr[OKIN_H_NM] : nullptr;
    out.design = io->design ? io->design + (size_t)i * 3 * nout : nullptr;
    const int nd = hdr[OKIN_H_NDIAG];
    out.diagnostics = (io->diagnostics && nd) ? io->diagnostics + (size_t)i * n_steps * nd : nullptr;
    out.status = io->status + i;
    out.failed_step = io->failed_step + i;
    out.worst_row = io->worst_row ? io->worst_row + i : nullptr;
    const double* tv = io->instance_targets ? io->instance_targets + (size_t)i * nt * n_steps : io->target_values;
    const double* par = io->params ? io->params + (size_t)i * hdr[OKIN_H_NPARAM] : nullptr;
    if (full)
      okin_sweep<true, true>(pr, sm.data(), io->hardpoints + (size_t)i * 3 * nin, par, tv, n_steps, cfg, out);
    else
      okin_sweep<false, true>(pr, sm.data(), io->hardpoints + (size_t)i * 3 * nin, par, tv, n_steps, cfg, out);
    for (size_t k = live; k < sm.size(); ++k)
      if (sm[k] == sm[k]) return -2;                   // lean instantiation wrote outside its slice
    if (out.diagnostics && n_steps > 0) {  // continuity pass, as the product's second kernel does
      const int stride = (n_steps - 1) | 1;
      std::vector<double> scratch((size_t)32 * stride + 64);
      const int failed = io->failed_step[i];
      okin_continuity(pr, scratch.data(), stride, io->positions + (size_t)i * n_steps * 3 * nout, n_steps,
                      failed < 0 ? n_steps : failed, out.diagnostics,
                      io->jumps ? io->jumps + (size_t)i * n_steps * (n / 3) : nullptr);
    }
  }
  return 0;
}

// One constraint row of the generated family functions (csrc/okin_gen_constraints.cuh): residual via
// both entry points and the gradient, for the family-row golden vectors.
extern "C" int okin_emu_family(int fam, const double* p, const double* c, double* res, double* res_only, double* g) {
  for (int k = 0; k < 12; ++k) g[k] = 0.0;
  *res = okin_family_resgrad(fam, p, c, g);
  *res_only = okin_family_res(fam, p, c);
  return 0;
}
