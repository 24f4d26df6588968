// K1 dispatch: register-resident path (k1_reg.cuh) when the model fits it, the 8-lanes-per-node path
// (k1_small.cu) for small all-dense models, else the shared-memory fast path (k1_fast.cu), else the generic
// kernel (k1_pdhg.cu).
#include "device.h"

namespace moip {

int launch_k1_reg_kd2(const DevModel&, const LpBatch&, const LpParams&, int, cudaStream_t);
int launch_k1_reg_kd3(const DevModel&, const LpBatch&, const LpParams&, int, cudaStream_t);
int launch_k1_reg_kd4(const DevModel&, const LpBatch&, const LpParams&, int, cudaStream_t);
int launch_k1_reg_kd5(const DevModel&, const LpBatch&, const LpParams&, int, cudaStream_t);

int launch_k1_reg(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st) {
  if (b.B <= 0) return MOIP_OK;
  if (!b.B_dev) MOIP_CUDA(cudaMemsetAsync(b.work_counter, 0, sizeof(int), st));   // (chained rounds: K5 resets it)
  switch (dm.KD) {
    case 2: return launch_k1_reg_kd2(dm, b, p, num_sms, st);
    case 3: return launch_k1_reg_kd3(dm, b, p, num_sms, st);
    case 4: return launch_k1_reg_kd4(dm, b, p, num_sms, st);
    case 5: return launch_k1_reg_kd5(dm, b, p, num_sms, st);
  }
  return MOIP_ERR_UNSUPPORTED;
}

int launch_k1_any(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st) {
  if (dm.reg_ok) return launch_k1_reg(dm, b, p, num_sms, st);
  if (k1_small_applies(dm)) return launch_k1_small(dm, b, p, num_sms, st);
  if (dm.fast_ok) return launch_k1_fast(dm, b, p, num_sms, st);
  return launch_k1(dm, b, p, num_sms, st);
}

}  // namespace moip
