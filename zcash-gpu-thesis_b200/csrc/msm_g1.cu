// G1 (Fq coordinates) instantiation of the MSM pipeline, see msm_impl.cuh
#include "msm_impl.cuh"
namespace b200zk {
B200ZK_MSM_INSTANTIATE(g1, fq_t)
}
