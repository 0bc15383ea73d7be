// ONE instantiation of the tcgen05 MPPI kernel per compilation of this file (see mppi_tc_kernel.cuh):
//   nvcc ... -DAMPC_TC_INST_CG=<1|2> -DAMPC_TC_INST_NXP=<4|8|16|24|32> -DAMPC_TC_INST_RELU=<0|1> -DAMPC_TC_INST_F16=<0|1> -DAMPC_TC_INST_TRACE=<0|1> -DAMPC_TC_INST_DZ=<0|1>
// autompc_b200/build.py compiles the whole matrix in parallel (ptxas needs ~30 s per instantiation).
#include "mppi_tc_kernel.cuh"

#ifndef AMPC_TC_INST_CG
#error "compile with -DAMPC_TC_INST_CG / _NXP / _RELU / _TRACE (autompc_b200/build.py does)"
#endif
#ifndef AMPC_TC_INST_F16
#define AMPC_TC_INST_F16 0
#endif
#ifndef AMPC_TC_INST_DZ
#define AMPC_TC_INST_DZ 0
#endif
#define AMPC_TC_GETTER_(cg, nxp, relu, f16, tr, dz) ampc_tc_kernel_cg##cg##_nxp##nxp##_relu##relu##_f16##f16##_trace##tr##_dz##dz
#define AMPC_TC_GETTER(cg, nxp, relu, f16, tr, dz) AMPC_TC_GETTER_(cg, nxp, relu, f16, tr, dz)

ampc_tc::TcKernel AMPC_TC_GETTER(AMPC_TC_INST_CG, AMPC_TC_INST_NXP, AMPC_TC_INST_RELU, AMPC_TC_INST_F16, AMPC_TC_INST_TRACE,
                                 AMPC_TC_INST_DZ)() {
  return (ampc_tc::TcKernel)ampc_tc::mppi_rollout_tc_kernel<AMPC_TC_INST_CG, AMPC_TC_INST_NXP, (AMPC_TC_INST_RELU != 0),
                                                            (AMPC_TC_INST_F16 != 0), (AMPC_TC_INST_TRACE != 0),
                                                            (AMPC_TC_INST_DZ != 0)>;
}
