// placeholder — replaced by the tcgen05/TMEM/TMA kernel
#include "../api.h"
namespace cpd {
int gemm_tc_cuda(const void*, const void*, float*, int, int, int, int, int, bool, int, int, const float*, cudaStream_t) { return -1; }
}
