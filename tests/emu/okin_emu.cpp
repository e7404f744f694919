// Lane-emulation build of the device core for CPU-only tests (never linked into the product
// library): every phase of csrc/okin_core.cuh runs as a loop over 32 lanes.
#define OKIN_LANE_EMU 1
#include <cstdlib>
#include <cstring>
#include <vector>

#include "okin_core.cuh"

extern "C" int okin_emu_sweep(const int32_t* hdr, const int32_t* ib, const double* fb, long n_instances,
                              int n_steps, const double* hardpoints, const double* params, const double* tvals,
                              double step_tol,
                              double coarse_tol, double fine_tol, double residual_tol, double mu_init, int max_iter, int use_predictor,
                              double* positions, int32_t* iters, double* max_residual, double* tangents,
                              double* metrics, double* design, int32_t* status, int32_t* failed_step) {
  if (hdr[OKIN_H_MAGIC] != OKIN_MAGIC) return -1;
  OkinProgram pr{hdr, ib, fb};
  OkinSolverCfg cfg{step_tol, coarse_tol, fine_tol, residual_tol, mu_init, max_iter, use_predictor};
  const int nin = hdr[OKIN_H_NIN], nout = hdr[OKIN_H_NOUT], nt = hdr[OKIN_H_NT], n = 3 * hdr[OKIN_H_NF];
  std::vector<double> sm(hdr[OKIN_H_SMEM_DOUBLES]);
  for (long i = 0; i < n_instances; ++i) {
    std::fill(sm.begin(), sm.end(), 0.0);
    OkinOutputs out;
    out.positions = positions ? positions + (size_t)i * n_steps * 3 * nout : nullptr;
    out.iters = iters ? iters + (size_t)i * n_steps : nullptr;
    out.max_residual = max_residual ? max_residual + (size_t)i * n_steps : nullptr;
    out.tangents = tangents ? tangents + (size_t)i * n_steps * nt * n : nullptr;
    out.metrics = metrics ? metrics + (size_t)i * n_steps * hdr[OKIN_H_NM] : nullptr;
    out.design = design ? design + (size_t)i * 3 * nout : nullptr;
    out.status = status + i;
    out.failed_step = failed_step + i;
    okin_sweep(pr, sm.data(), hardpoints + (size_t)i * 3 * nin,
               params ? params + (size_t)i * hdr[OKIN_H_NPARAM] : nullptr, tvals, n_steps, cfg, out);
  }
  return 0;
}
