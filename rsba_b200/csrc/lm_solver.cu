// Levenberg-Marquardt driver (placeholder until K2-K4 land in this round).
#include "problem.cuh"
namespace rsba {
struct LmState {};
void lm_state_free(LmState* s) { delete s; }
}  // namespace rsba
extern "C" {
int rsba_cuda_solve(rsba_problem*, const rsba_solve_options*, rsba_solve_summary*) {
  rsba::set_last_error("rsba_cuda_solve: not built yet");
  return RSBA_ERR_STATE;
}
int rsba_cuda_linearize_and_step(rsba_problem*, const rsba_solve_options*, double, double*, double*,
                                 double*, double*, double*) {
  rsba::set_last_error("rsba_cuda_linearize_and_step: not built yet");
  return RSBA_ERR_STATE;
}
int rsba_cuda_nccl_unique_id(unsigned char*) { return RSBA_ERR_NCCL; }
int rsba_cuda_comm_init(rsba_problem*, int, int, const unsigned char*) { return RSBA_ERR_NCCL; }
}
