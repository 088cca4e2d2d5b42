// b200mpc: one instantiation set of ocp_ipm_kernel per translation unit (-DOCP_INST_SET=0..4), see ocp_launch.cuh
#include "ocp_inst.cuh"
