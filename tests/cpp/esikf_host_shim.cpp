// Test-only extern "C" shim around the PRODUCT's host update engine (better_fastlio2_b200/csrc/esikf_host.hpp — pure C++,
// no CUDA), so that tests/test_esikf_host_cpu.py can drive it on the CPU with measurement passes computed by the oracle
// and compare the posterior with the oracle's own update_iterated_dyn_share_modified restatement.
#include "esikf_host.hpp"

using flb::host::IteratedUpdate;

extern "C" {
void* iu_create(const double* state26, const double* P23, double R, int max_iter, const double* limit) {
  return new IteratedUpdate(state26, P23, R, max_iter, limit);
}
void iu_destroy(void* h) { delete static_cast<IteratedUpdate*>(h); }
int iu_more(void* h) { return static_cast<IteratedUpdate*>(h)->more() ? 1 : 0; }
int iu_need_search(void* h) { return static_cast<IteratedUpdate*>(h)->need_search() ? 1 : 0; }
void iu_current_state(void* h, double* s26) { static_cast<IteratedUpdate*>(h)->current_state(s26); }
void iu_step(void* h, const double* HTH, const double* HTh) { static_cast<IteratedUpdate*>(h)->step(HTH, HTh); }
void iu_step_rows(void* h, const double* hx, const double* hv, int M) { static_cast<IteratedUpdate*>(h)->step_rows(hx, hv, M); }
void iu_skip(void* h) { static_cast<IteratedUpdate*>(h)->skip(); }
void iu_result(void* h, double* state26, double* P23) { static_cast<IteratedUpdate*>(h)->result(state26, P23); }
int iu_converged_count(void* h) { return static_cast<IteratedUpdate*>(h)->converged_count(); }
}
