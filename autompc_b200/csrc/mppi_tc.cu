// tcgen05 (UMMA) bf16 MPPI rollout path -- placeholder until the kernel lands.
#include "ampc_common.cuh"

struct AmpcTcPlan { int unused; };

int ampc_mppi_tc_supported(const ampc_mppi_cfg *, const ampc_mlp_desc *, const char **why) {
  *why = "tcgen05 path not built in this revision";
  return 0;
}
int ampc_mppi_tc_create(AmpcTcPlan **, const ampc_mppi_cfg *, const ampc_mlp_desc *) { return AMPC_ERR_UNSUPPORTED; }
void ampc_mppi_tc_destroy(AmpcTcPlan *) {}
int ampc_mppi_tc_grid(const AmpcTcPlan *) { return 0; }
int ampc_mppi_tc_launch(AmpcTcPlan *, const AmpcMppiParams &, cudaStream_t) { return AMPC_ERR_UNSUPPORTED; }
