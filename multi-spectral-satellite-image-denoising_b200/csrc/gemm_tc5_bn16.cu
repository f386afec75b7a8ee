// tcgen05 GEMM instantiation for BLOCK_N = 16 (one translation unit per tile width: parallel build)
#include "gemm_tc5.cuh"
namespace gemm { namespace tc5 { template int launch_bn<16>(const Problem&, cudaStream_t); } }
