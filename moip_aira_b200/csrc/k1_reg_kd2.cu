// K1 register-resident path, instantiations for KD = 2 dense rows (see k1_reg.cuh)
#include "k1_reg.cuh"
namespace moip {
int launch_k1_reg_kd2(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st) {
  return k1reg::launch_reg_kd<2>(dm, b, p, num_sms, st);
}
}  // namespace moip
